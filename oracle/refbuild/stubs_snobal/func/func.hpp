// Stand-in for FunC (third party, not in the reference tree): sno.cpp builds two lookup tables of the saturation vapour
// pressure with it (sno.cpp:34-36).  They are not on the _adj_snow path; the stand-in just evaluates the function.
// Test infrastructure (oracle/), never linked into the product.
#pragma once
#include <initializer_list>
namespace func {
template <typename T> struct LookupTableParameters { T lo, hi, step; };
template <typename T> class UniformLinearRawInterpTable {
    T (*f)(T);
public:
    UniformLinearRawInterpTable(std::initializer_list<T (*)(T)> fs, LookupTableParameters<T>) : f(*fs.begin()) {}
    T operator()(T x) const { return f(x); }
};
}
