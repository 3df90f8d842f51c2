// Harness around the reference's third_party/snobal/sno.cpp (compiled unmodified, where it lies): the step on the far side of
// PBSM3D — snobal applies `drift_mass` to each face's snowpack.
//   module glue restated here (it lives in src/modules/snobal.cpp:363-385, which needs all of CHM to compile):
//       mass = is_nan(drift_mass) ? 0 : drift_mass;
//       transport_density = mass < 0 ? sbal->rho : drift_density;        // erosion at the pack's density, deposition at drift_density
//       sbal->_adj_snow(mass / transport_density, mass);
//       _adj_snow(delta_avalanche_snowdepth / area, delta_avalanche_mass / area * 1000);   // snobal.cpp:389-408 (snow_slide's output)
//   the reference's own code: sno::_adj_snow (sno.cpp:2527-2575), _adj_layers (:2617-2696), _calc_layers (:2366-2405),
//   _layer_mass (:1564-1580), _cold_content (:2321-2329).
// TEST INFRASTRUCTURE (oracle/): generates tests/golden/golden_snobal_drift.npz and pins oracle/snobal_oracle.py.
#include <cmath>
#include <cstddef>
#include "sno.h"

// SoA state, field f of face i at state[f * n + i]; order fixed by oracle/snobal_oracle.py:FIELDS
enum { F_Z_S, F_M_S, F_RHO, F_LAYERS, F_Z_S_0, F_Z_S_L, F_M_S_0, F_M_S_L, F_CC_S, F_CC_S_0, F_CC_S_L, F_T_S, F_T_S_0, F_T_S_L,
       F_H2O_TOTAL, F_H2O_VOL, F_H2O, F_H2O_MAX, F_H2O_SAT, F_COUNT };

static bool chm_is_nan(double v) { return std::fabs(v - -9999.0) < 1e-5 || std::isnan(v); }  // module_base.hpp:471-479

extern "C" int chmref_sno_fields(void) { return F_COUNT; }

static void load(sno& s, double* state, int n, int i)
{
    auto at = [&](int f) -> double& { return state[(std::size_t)f * n + i]; };
    s.z_s = at(F_Z_S); s.m_s = at(F_M_S); s.rho = at(F_RHO); s.layer_count = (int)at(F_LAYERS);
    s.z_s_0 = at(F_Z_S_0); s.z_s_l = at(F_Z_S_L); s.m_s_0 = at(F_M_S_0); s.m_s_l = at(F_M_S_L);
    s.cc_s = at(F_CC_S); s.cc_s_0 = at(F_CC_S_0); s.cc_s_l = at(F_CC_S_L);
    s.T_s = at(F_T_S); s.T_s_0 = at(F_T_S_0); s.T_s_l = at(F_T_S_L);
    s.h2o_total = at(F_H2O_TOTAL); s.h2o_vol = at(F_H2O_VOL); s.h2o = at(F_H2O); s.h2o_max = at(F_H2O_MAX); s.h2o_sat = at(F_H2O_SAT);
}
static void store(const sno& s, double* state, int n, int i)
{
    auto at = [&](int f) -> double& { return state[(std::size_t)f * n + i]; };
    at(F_Z_S) = s.z_s; at(F_M_S) = s.m_s; at(F_RHO) = s.rho; at(F_LAYERS) = (double)s.layer_count;
    at(F_Z_S_0) = s.z_s_0; at(F_Z_S_L) = s.z_s_l; at(F_M_S_0) = s.m_s_0; at(F_M_S_L) = s.m_s_l;
    at(F_CC_S) = s.cc_s; at(F_CC_S_0) = s.cc_s_0; at(F_CC_S_L) = s.cc_s_l;
    at(F_T_S) = s.T_s; at(F_T_S_0) = s.T_s_0; at(F_T_S_L) = s.T_s_l;
    at(F_H2O_TOTAL) = s.h2o_total; at(F_H2O_VOL) = s.h2o_vol; at(F_H2O) = s.h2o; at(F_H2O_MAX) = s.h2o_max; at(F_H2O_SAT) = s.h2o_sat;
}

extern "C" int chmref_sno_apply_drift(int n, double* state, const double* drift_mass, double drift_density, double threshold,
                                      double max_z_s_0)
{
    sno s;
    s.tstep_info[SMALL_TSTEP].threshold = threshold;  // snobal.cpp:190
    s.max_z_s_0 = max_z_s_0;                          // snobal.cpp:101
    for (int i = 0; i < n; ++i) {
        load(s, state, n, i);
        double mass = drift_mass[i];
        mass = chm_is_nan(mass) ? 0 : mass;
        const double transport_density = mass < 0. ? s.rho : drift_density;
        s._adj_snow(mass / transport_density, mass);
        store(s, state, n, i);
    }
    return 0;
}

// snobal.cpp:389-408: snow_slide's delta_avalanche_snowdepth (a volume) and delta_avalanche_mass (a swe volume) per face
extern "C" int chmref_sno_apply_avalanche(int n, double* state, const double* delta_avalanche_snowdepth, const double* delta_avalanche_swe,
                                          const double* area, double threshold, double max_z_s_0)
{
    sno s;
    s.tstep_info[SMALL_TSTEP].threshold = threshold;
    s.max_z_s_0 = max_z_s_0;
    for (int i = 0; i < n; ++i) {
        load(s, state, n, i);
        double d_depth = delta_avalanche_snowdepth[i] / area[i];
        double d_mass = delta_avalanche_swe[i] / area[i] * 1000;
        s._adj_snow(d_depth, d_mass);
        store(s, state, n, i);
    }
    return 0;
}
