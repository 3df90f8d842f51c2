// Stand-in for interpolation/TPSpline.hpp (Eigen, GSL and FunC are absent): fetchr.hpp:29 includes it and uses nothing of it;
// the interpolant itself is restated in stubs/interpolation.hpp.
#pragma once
