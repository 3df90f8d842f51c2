// Stand-in for the reference's mesh/triangulation.hpp (CGAL-based; CGAL is absent).  Only the members PBSM3D.cpp and
// module_base.hpp touch are provided, each following the reference implementation it replaces:
//   face::edge / edge_unit_normal / edge_length   triangulation.hpp:1443-1474, 1492-1498
//   face::center / get_z                          triangulation.hpp:1577-1589, 1774-1780
//   face::get_area                                triangulation.hpp:1830-1856
//   face::has_vegetation / veg_attribute          triangulation.hpp:1656-1697
//   variable store default -9999                  triangulation.cpp:2536-2560
//   face::normal / slope (snow_slide)             triangulation.hpp:1501-1523, 1549-1574 (CGAL::unit_normal, arma::norm_dot)
// Test infrastructure (oracle/), never linked into the product.
#pragma once
#include <CGAL/Exact_predicates_inexact_constructions_kernel.h>
#include <boost/function.hpp>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

typedef CGAL::Exact_predicates_inexact_constructions_kernel::Vector_2 Vector_2;
typedef CGAL::Exact_predicates_inexact_constructions_kernel::Vector_3 Vector_3;
typedef CGAL::Exact_predicates_inexact_constructions_kernel::Point_2 Point_2;
typedef CGAL::Exact_predicates_inexact_constructions_kernel::Point_3 Point_3;

// "name"_s: the reference hashes the name at compile time (utility/xxh64.hpp:104-116); a string key is equivalent here.
inline std::string operator"" _s(const char* s, std::size_t len) { return std::string(s, len); }

// CHM's exception machinery (exception.hpp:46,133), reduced to what the path throws.
struct module_error : public std::runtime_error { using std::runtime_error::runtime_error; };
#define CHM_THROW_EXCEPTION(exception_type, message) throw exception_type(std::string(message))

// boost::property_tree::ptree as used on the path: typed get(path, default) / get<T>(path).  Like ptree's stream
// translator, a value that does not parse COMPLETELY as T yields the default (e.g. "820.5" read as int).
class ptree_stub {
public:
    std::map<std::string, std::string> kv;
    bool parse(const std::string& s, bool& out) const
    {
        if (s == "true" || s == "1") { out = true; return true; }
        if (s == "false" || s == "0") { out = false; return true; }
        return false;
    }
    bool parse(const std::string& s, int& out) const
    {
        try { std::size_t p; long v = std::stol(s, &p); if (p != s.size()) return false; out = (int)v; return true; } catch (...) { return false; }
    }
    bool parse(const std::string& s, double& out) const
    {
        try { std::size_t p; double v = std::stod(s, &p); if (p != s.size()) return false; out = v; return true; } catch (...) { return false; }
    }
    template <class T> T get(const std::string& key, const T& def) const
    {
        auto it = kv.find(key);
        T out;
        if (it == kv.end() || !parse(it->second, out)) return def;
        return out;
    }
    template <class T> T get(const std::string& key) const
    {
        auto it = kv.find(key);
        T out;
        if (it == kv.end() || !parse(it->second, out)) throw std::out_of_range("ptree_bad_path: " + key);
        return out;
    }
};

class global {
public:
    ptree_stub parameters;
    double _dt = 3600;
    bool _is_point_mode = false;
    double dt() const { return _dt; }
    bool is_point_mode() const { return _is_point_mode; }  // global.hpp:65-71
};

class face_info { public: virtual ~face_info() {} };
class triangulation;

class face_stub {
public:
    triangulation* _domain = nullptr;
    std::size_t cell_global_id = 0, cell_local_id = 0;
    bool is_ghost = false, _is_ghost = false;
    int owner = 0;
    double vx[3], vy[3], vz[3];
    face_stub* _neigh[3] = {nullptr, nullptr, nullptr};
    std::unordered_map<std::string, double> _variables, _parameters;
    std::unordered_map<std::string, Vector_3> _vectors;
    std::vector<std::unique_ptr<face_info>> _module_data;

    double& operator[](const std::string& v)
    {
        auto it = _variables.find(v);
        if (it == _variables.end()) it = _variables.emplace(v, -9999.0).first;
        return it->second;
    }
    face_stub* neighbor(int i) const { return _neigh[i]; }
    bool has_parameter(const std::string& p) const { return _parameters.count(p) != 0; }
    double parameter(const std::string& p) const
    {
        auto it = _parameters.find(p);
        if (it == _parameters.end()) throw module_error("Parameter " + p + " does not exist.");
        return it->second;
    }
    bool has_vegetation() const { return has_parameter("landcover") || has_parameter("canopyType") || has_parameter("CanopyHeight"); }
    double veg_attribute(const std::string& variable);

    Vector_2 edge(int i) const
    {
        const int a = (i + 1) % 3 /*ccw(i)*/, b = (i + 2) % 3 /*cw(i)*/;
        return Vector_2(vx[b] - vx[a], vy[b] - vy[a]);
    }
    Vector_2 edge_unit_normal(int i) const
    {
        auto e = edge(i);
        auto e1 = edge((i + 1) % 3);
        Vector_2 n(e.y(), -e.x());
        double D = e1.x() * n.x() + e1.y() * n.y();
        if (D > 0) n = -n;
        return n / CGAL::sqrt(n.squared_length());
    }
    double edge_length(int i) const { return CGAL::sqrt(edge(i).squared_length()); }
    Point_3 center() const { return Point_3((vx[0] + vx[1] + vx[2]) / 3, (vy[0] + vy[1] + vy[2]) / 3, (vz[0] + vz[1] + vz[2]) / 3); }
    double get_z() const { return center().z(); }
    // CGAL::unit_normal(p, q, r) = cross(q - p, r - p) / |.|  (projected mesh; the geographic branch scales x, y by 1e5)
    Vector_3 normal() const
    {
        const double ax = vx[1] - vx[0], ay = vy[1] - vy[0], az = vz[1] - vz[0];
        const double bx = vx[2] - vx[0], by = vy[2] - vy[0], bz = vz[2] - vz[0];
        const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const double len = CGAL::sqrt(nx * nx + ny * ny + nz * nz);
        return Vector_3(nx / len, ny / len, nz / len);
    }
    // acos(arma::norm_dot(normal, (0,0,1))), norm_dot(a,b) = dot(a,b) / (norm(a) norm(b))
    double slope() const
    {
        const Vector_3 n = normal();
        const double dot = n.x() * 0.0 + n.y() * 0.0 + n.z() * 1.0;
        const double na = std::sqrt(n.x() * n.x() + n.y() * n.y() + n.z() * n.z()), nb = std::sqrt(0.0 * 0.0 + 0.0 * 0.0 + 1.0 * 1.0);
        return std::acos(dot / (na * nb));
    }
    double get_x() const { return center().x(); }  // triangulation.hpp:1755-1771
    double get_y() const { return center().y(); }
    // triangulation.hpp:1543-1546 -> triangulation.cpp:170-186: nearest face CENTRE to the point `distance` along `azimuth`
    face_stub* find_closest_face(double azimuth, double distance);
    double get_area() const
    {
        if (has_parameter("area")) return parameter("area");
        return ((vx[1] - vx[0]) * (vy[2] - vy[0]) - (vx[2] - vx[0]) * (vy[1] - vy[0])) / 2;
    }
    void set_face_vector(const std::string& name, Vector_3 v) { _vectors[name] = v; }

    template <class T> T& make_module_data(std::size_t id)
    {
        if (_module_data.size() <= id) _module_data.resize(id + 1);
        _module_data[id].reset(new T);
        return *static_cast<T*>(_module_data[id].get());
    }
    template <class T> T& get_module_data(std::size_t id) { return *static_cast<T*>(_module_data.at(id).get()); }
};
typedef face_stub* mesh_elem;

class triangulation {
public:
    std::vector<std::unique_ptr<face_stub>> _faces;  // locally owned faces in ascending cell_global_id
    std::shared_ptr<global> _global;
    std::size_t _n_global = 0;
    mesh_elem face(std::size_t i) { return _faces[i].get(); }
    std::size_t size_faces() const { return _faces.size(); }
    std::size_t size_global_faces() const { return _n_global; }
    void ghost_neighbors_communicate_variable(const std::string&) {}  // single rank in the harness: nothing to exchange
    void ghost_to_neighbors_communicate_variable(const std::string&) {}
    std::vector<std::unique_ptr<face_stub>> _ghosts;  // is_ghost faces hanging off owned faces (snow_slide harness): not in _faces
    void print_ghost_neighbor_info() {}
};
typedef std::shared_ptr<triangulation> mesh;

// CGAL's kd-tree k=1 query over the centres of _faces (triangulation.cpp:1037-1058), restated as an exhaustive search.
namespace math { namespace gis { extern boost::function<Point_2(Point_3 src, double bearing, double distance)> point_from_bearing; } }
inline face_stub* face_stub::find_closest_face(double azimuth, double distance)
{
    const Point_2 q = math::gis::point_from_bearing(center(), azimuth, distance);
    face_stub* best = nullptr;
    double bd = 0;
    for (auto& f : _domain->_faces) {
        const Point_3 c = f->center();
        const double d = (c.x() - q.x()) * (c.x() - q.x()) + (c.y() - q.y()) * (c.y() - q.y());
        if (!best || d < bd) { best = f.get(); bd = d; }
    }
    return best;
}

inline double face_stub::veg_attribute(const std::string& variable)
{
    if (has_parameter(variable)) return parameter(variable);
    if (has_parameter("landcover")) {
        int LC = (int)parameter("landcover");
        try { return _domain->_global->parameters.get<double>("landcover." + std::to_string(LC) + "." + variable); }
        catch (const std::out_of_range&) { CHM_THROW_EXCEPTION(module_error, "Parameter " + variable + " does not exist."); }
    }
    CHM_THROW_EXCEPTION(module_error, "Parameter " + variable + " does not exist.");
}
