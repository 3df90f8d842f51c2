// chm_shim.hpp — stand-in for the handful of CHM declarations PBSM3D_gpu touches, so the adaptor can be compiled
// and run in this repository, where CHM's dependencies (Boost, CGAL, ...) do not exist.  Names, signatures and
// meaning follow CHM; nothing here is copied from it.  Inside CHM this header is not used (see PBSM3D_gpu.hpp).
//
//   CHM declaration                                        stand-in
//   -----------------------------------------------------  ----------------------------------------------
//   config_file = pt::ptree, cfg.get<T>(key, default)      config_file (string map with typed get)
//   module_base(name, parallel, cfg), depends/provides,    module_base
//     global_param->dt(), is_water(face), cfg               (src/modules/module_base.hpp:58-540)
//   REGISTER_MODULE_HPP / _CPP                             no-ops (factory registration, module_base.hpp:534-540)
//   mesh = shared_ptr<triangulation>; size_faces(),        triangulation / face
//     size_global_faces(), face(i); face->neighbor(j),      (src/mesh/triangulation.hpp:1173, :237-560)
//     is_ghost, owner, cell_global_id, cell_local_id,
//     vertex(v)->point(), has_parameter, get_area,
//     has_vegetation, veg_attribute, (*face)["var"_s]
//   module_error + CHM_THROW_EXCEPTION                     std::runtime_error subclass (src/exception.hpp:133-134)
//   netcdf::create_variable1D / put_var1D / get_var1D      in-memory map (src/netcdf.hpp)
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define SPDLOG_DEBUG(...) ((void)0)
#define REGISTER_MODULE_HPP(name)
#define REGISTER_MODULE_CPP(name)

struct module_error : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
#define CHM_THROW_EXCEPTION(type, msg) throw type(msg)

inline std::string operator""_s(const char* s, size_t n) { return std::string(s, n); }

class config_file
{
  public:
    std::map<std::string, std::string> kv;
    template <typename T> T get(const std::string& key, T def) const
    {
        auto it = kv.find(key);
        if (it == kv.end())
            return def;
        if (it->second == "true")
            return (T)1;
        if (it->second == "false")
            return (T)0;
        std::istringstream ss(it->second);
        double v = 0;
        ss >> v;
        return (T)v;
    }
    int get(const std::string& key, int def) const { return get<int>(key, def); }
    double get(const std::string& key, double def) const { return get<double>(key, def); }
    bool get(const std::string& key, bool def) const { return get<int>(key, def ? 1 : 0) != 0; }
};

struct shim_point
{
    double _x, _y, _z;
    double x() const { return _x; }
    double y() const { return _y; }
    double z() const { return _z; }
};
struct shim_vertex
{
    shim_point p;
    const shim_point& point() const { return p; }
};

class shim_face
{
  public:
    bool is_ghost = false;
    int owner = 0;
    size_t cell_global_id = 0, cell_local_id = 0;
    shim_face* nb[3] = {nullptr, nullptr, nullptr};
    shim_vertex* vtx[3] = {nullptr, nullptr, nullptr};
    std::map<std::string, double> params, veg, vars;
    bool water = false;

    shim_face* neighbor(int j) const { return nb[j]; }
    shim_vertex* vertex(int v) const { return vtx[v]; }
    bool has_parameter(const std::string& k) const { return params.count(k) > 0; }
    double get_area() const { return params.at("area"); }
    bool has_vegetation() const { return !veg.empty(); }
    double veg_attribute(const std::string& k) const
    {
        auto it = veg.find(k);
        if (it == veg.end())
            throw module_error("Parameter " + k + " does not exist.");
        return it->second;
    }
    double& operator[](const std::string& name)
    {
        auto it = vars.find(name);
        if (it == vars.end())
            it = vars.emplace(name, -9999.0).first; // variablestorage default
        return it->second;
    }
};
typedef shim_face* mesh_elem;

class triangulation
{
  public:
    std::vector<std::unique_ptr<shim_face>> faces, ghosts;
    std::vector<std::unique_ptr<shim_vertex>> vertices;
    size_t n_global = 0;
    size_t size_faces() const { return faces.size(); }
    size_t size_global_faces() const { return n_global; }
    bool is_geographic() const { return geographic; } // triangulation.cpp:98-101
    bool geographic = false;
    mesh_elem face(size_t i) const { return faces[i].get(); }
};
typedef std::shared_ptr<triangulation> mesh;

class netcdf
{
  public:
    std::map<std::string, std::vector<double>> data;
    void create_variable1D(const std::string& n, size_t len) { data[n].assign(len, 0.0); }
    void put_var1D(const std::string& n, size_t i, double v) { data[n][i] = v; }
    double get_var1D(const std::string& n, size_t i) { return data.at(n)[i]; }
};

namespace parallel
{
enum type
{
    data,
    domain
};
}

struct shim_global
{
    double _dt = 3600.0;
    double dt() const { return _dt; }
};

class module_base
{
  public:
    std::string ID;
    config_file cfg;
    std::shared_ptr<shim_global> global_param = std::make_shared<shim_global>();
    std::vector<std::string> _depends, _provides;
    module_base(std::string name, parallel::type, config_file c) : ID(std::move(name)), cfg(std::move(c)) {}
    virtual ~module_base() {}
    void depends(const std::string& v) { _depends.push_back(v); }
    void provides(const std::string& v) { _provides.push_back(v); }
    bool is_water(mesh_elem f) const { return f->water; }
    bool is_nan(double v) const { return std::fabs(v - -9999.0) < 1e-5 || std::isnan(v); }
};
