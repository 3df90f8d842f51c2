// Stand-in for interpolation.hpp: PBSM3D.hpp:27 includes it but the module never interpolates.
#pragma once
