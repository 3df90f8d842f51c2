// Stand-in for oneTBB's parallel_sort (TBB is absent).  Like tbb::parallel_sort, std::sort is not stable: the order of faces with
// EQUAL keys is unspecified in the reference; the fixtures keep the keys of neighbouring active faces distinct.
#pragma once
#include <algorithm>
namespace tbb { template <class It, class Cmp> void parallel_sort(It a, It b, Cmp c) { std::sort(a, b, c); } }
