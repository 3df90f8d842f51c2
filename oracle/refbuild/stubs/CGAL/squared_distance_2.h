// Stand-in: squared_distance lives in the kernel stand-in.
#pragma once
#include <CGAL/Exact_predicates_inexact_constructions_kernel.h>
