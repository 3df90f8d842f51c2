// Stand-in for the CGAL Epick kernel types the hot path touches (CGAL is absent).  For every construction used by
// PBSM3D / coordinates.cpp Epick is plain fp64 arithmetic; only those members are provided.
#pragma once
#include <cmath>
namespace CGAL {
struct Vector_2 {
    double _x = 0, _y = 0;
    Vector_2() {}
    Vector_2(double x, double y) : _x(x), _y(y) {}
    double x() const { return _x; }
    double y() const { return _y; }
    Vector_2 operator-() const { return Vector_2(-_x, -_y); }
    Vector_2 operator/(double s) const { return Vector_2(_x / s, _y / s); }
    double squared_length() const { return _x * _x + _y * _y; }
};
struct Vector_3 {
    double _x = 0, _y = 0, _z = 0;
    Vector_3() {}
    Vector_3(double x, double y, double z) : _x(x), _y(y), _z(z) {}
    double x() const { return _x; }
    double y() const { return _y; }
    double z() const { return _z; }
};
struct Point_2 {
    double _x = 0, _y = 0;
    Point_2() {}
    Point_2(double x, double y) : _x(x), _y(y) {}
    double x() const { return _x; }
    double y() const { return _y; }
};
struct Point_3 {
    double _x = 0, _y = 0, _z = 0;
    Point_3() {}
    Point_3(double x, double y, double z) : _x(x), _y(y), _z(z) {}
    double x() const { return _x; }
    double y() const { return _y; }
    double z() const { return _z; }
};
inline double sqrt(double v) { return std::sqrt(v); }
inline double squared_distance(const Point_2& a, const Point_2& b)
{
    const double dx = a.x() - b.x(), dy = a.y() - b.y();
    return dx * dx + dy * dy;
}
struct Exact_predicates_inexact_constructions_kernel {
    typedef CGAL::Point_2 Point_2;
    typedef CGAL::Point_3 Point_3;
    typedef CGAL::Vector_2 Vector_2;
    typedef CGAL::Vector_3 Vector_3;
};
}
