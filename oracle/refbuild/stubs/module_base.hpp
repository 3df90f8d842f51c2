// Stand-in for the reference's modules/module_base.hpp (Boost-based): the members PBSM3D uses, with is_nan / is_water
// following module_base.hpp:471-491.  Test infrastructure (oracle/), never linked into the product.
#pragma once
#include "triangulation.hpp"
#include "interpolation.hpp"
#include <cmath>
#include <memory>
#include <string>
#include <vector>

typedef ptree_stub config_file;

class netcdf {
public:
    std::map<std::string, std::vector<double>> vars;
    void create_variable1D(const std::string& n, std::size_t len) { vars[n].assign(len, 0.0); }
    void put_var1D(const std::string& n, std::size_t i, double v) { vars.at(n).at(i) = v; }
    double get_var1D(const std::string& n, std::size_t i) { return vars.at(n).at(i); }
};

class module_base {
public:
    enum class parallel { data, domain };
    std::string ID_name;
    int ID = 0;
    int IDnum = 0;
    config_file cfg;
    std::shared_ptr<global> global_param;
    std::vector<std::string> _depends, _provides, _vectors;

    module_base(std::string name, parallel type, config_file c) : ID_name(name), cfg(c), _parallel_type(type) {}
    virtual ~module_base() {}
    virtual void run(mesh&) {}
    virtual void init(mesh&) {}
    virtual void checkpoint(mesh&, netcdf&) {}
    virtual void load_checkpoint(mesh&, netcdf&) {}

    enum class SpatialType { local, neighbor, distance };  // module_base.hpp:77
    void depends(const std::string& n) { _depends.push_back(n); }
    void depends(const std::string& n, SpatialType) { _depends.push_back(n); }
    void provides(const std::string& n) { _provides.push_back(n); }
    void provides_vector(const std::string& n) { _vectors.push_back(n); }
    // module_base.hpp:416-435: an optional input counts as found when another module provides it; the harness says which
    std::map<std::string, bool> _optional_found;
    void optional(const std::string& n) { _optional_found.emplace(n, false); }
    bool has_optional(const std::string& n) { auto it = _optional_found.find(n); return it != _optional_found.end() && it->second; }
    virtual void run(mesh_elem&) {}

    bool is_nan(const double& variable)
    {
        if (std::fabs(variable - -9999.0) < 1e-5) return true;
        if (std::isnan(variable)) return true;
        return false;
    }
    bool is_water(mesh_elem& face)
    {
        bool is = false;
        if (face->has_parameter("landcover")) {
            int LC = face->parameter("landcover");
            is = global_param->parameters.get<bool>("landcover." + std::to_string(LC) + ".is_water", false);
        }
        return is;
    }
protected:
    parallel _parallel_type;
};

#define REGISTER_MODULE_HPP(Implementation) static_assert(true, "")
#define REGISTER_MODULE_CPP(Implementation) static_assert(true, "")
