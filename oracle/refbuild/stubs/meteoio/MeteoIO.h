// Stand-in for <meteoio/MeteoIO.h>.  MeteoIO (version unpinned, reference spack.yaml:29) is NOT in /root/reference, so
// this one function is a restatement of MeteoIO's published Atmosphere::stdAirPressure / stdDryAirDensity
// (ICAO standard atmosphere with geopotential height), constants as in MeteoIO's Meteoconst.h.  It is the one piece of
// arithmetic on the path that the reference sources cannot pin (call site PBSM3D.cpp:785); it scales c_salt uniformly.
#pragma once
#include <cmath>
namespace mio {
namespace Cst {
const double stefan_boltzmann = 5.670373e-8;
const double gravity = 9.80665;
const double std_press = 101325.;
const double std_temp = 288.15;
const double mean_adiabatique_lapse_rate = 0.0065;
const double earth_R0 = 6356766.0;
const double gaz_constant_dry_air = 287.058;
const double gaz_constant = 8.31451;
const double l_water_sublimation = 2.838e6;
const double specific_heat_air = 1004.67;
const double specific_heat_ice = 2100.0;
}
namespace Atmosphere {
inline double stdAirPressure(const double& altitude)
{
    const double expo = Cst::gravity / (Cst::mean_adiabatique_lapse_rate * Cst::gaz_constant_dry_air);
    return Cst::std_press * std::pow(1. - ((Cst::mean_adiabatique_lapse_rate * Cst::earth_R0 * altitude) / (Cst::std_temp * (Cst::earth_R0 + altitude))), expo);
}
inline double stdDryAirDensity(const double& altitude, const double& temperature)
{
    return stdAirPressure(altitude) / (Cst::gaz_constant_dry_air * temperature);
}
}
}
