// Device restatement of the scalar helpers PBSM3D's hot loop calls.
// Written from the formulas, not from the reference's code layout; each cites what it must agree with.
#pragma once
#include <cmath>
#include <cstdint>

#include "pbsm3d_math.cuh"

namespace pbsm3d {

constexpr double kKappa = 0.4;     // PhysConst::kappa        (physics/PhysConst.h:31)
constexpr double kRhoIce = 917.0;  // PhysConst::rho_ice      (physics/PhysConst.h:39)
constexpr double kZUR = 50.0;      // Atmosphere::Z_U_R       (physics/Atmosphere.h:31)
constexpr double kZ0Snow = 0.01;   // Snow::Z0_SNOW           (physics/Snow.h:31)
constexpr double kPi = 3.14159265358979323846;

// module_base::is_nan (modules/module_base.hpp:471-479): the -9999 sentinel or a real NaN.
__device__ __forceinline__ bool chm_is_nan(double v) { return fabs(v - -9999.0) < 1e-5 || isnan(v); }

// -math::gis::bearing_to_cartesian(phi) (math/coordinates.cpp:112-131): unit vector the wind blows TOWARDS.
__device__ __forceinline__ void wind_unit_vector(double bearing, double& vx, double& vy) {
    double h = 450.0 - bearing;
    if (h > 360.0) h = h - 360.0;
    double th = h * kPi / 180.0;
    double s, c;
    sincos(th, &s, &c);
    vx = -c;
    vy = -s;
}

// Atmosphere::log_scale_wind (physics/Atmosphere.cpp:32-38).
__device__ __forceinline__ double log_scale_wind(double u, double Z_in, double Z_out, double sd, double z0) {
    return u * log((Z_out - (sd + z0)) / z0) / log((Z_in - (sd + z0)) / z0);
}

// Atmosphere::saturatedVapourPressure (physics/Atmosphere.cpp:62-80) as PBSM3D calls it: the argument is
// Kelvin and the branch test `T >= 0` therefore always selects the over-water coefficients.
__device__ __forceinline__ double saturated_vapour_pressure(double t_kelvin) {
    double TA = t_kelvin - 273.15;
    return 611.21 * exp((17.502 * TA) / (240.97 + TA));
}

// mio::Atmosphere::stdDryAirDensity(altitude, T) — MeteoIO is not vendored in the reference tree; restated
// from its published source (standard-atmosphere pressure over R_dry*T).  See DESIGN.md "third-party arithmetic".
__device__ __forceinline__ double std_dry_air_density(double z, double t_kelvin) {
    const double R0 = 6356766.0, g = 9.80665, Rd = 287.058, lapse = 0.0065, T0 = 288.15, p0 = 101325.0;
    const double expo = g / (lapse * Rd);
    double p = p0 * pow(1.0 - ((lapse * R0 * z) / (T0 * (R0 + z))), expo);
    return p / (Rd * t_kelvin);
}

// The same with exp(y ln x) from pbsm3d_math.cuh in place of pow() (assembly prelude; ~1e-15 relative of the above).
__device__ __forceinline__ double std_dry_air_density_fast(double z, double t_kelvin) {
    const double R0 = 6356766.0, g = 9.80665, Rd = 287.058, lapse = 0.0065, T0 = 288.15, p0 = 101325.0;
    const double expo = g / (lapse * Rd);
    double p = p0 * fpow(1.0 - ((lapse * R0 * z) * frcp(T0 * (R0 + z))), expo);
    return p * frcp(Rd * t_kelvin);
}

}  // namespace pbsm3d

