// Stand-in for oneTBB's concurrent_vector (TBB is absent): snow_slide.cpp sizes it up front and writes element i from iteration i.
#pragma once
#include <vector>
namespace tbb { template <class T> using concurrent_vector = std::vector<T>; }
