// The step on the far side of PBSM3D: snobal applies `drift_mass` -- and, when snow_slide runs, the
// avalanche volume/mass deltas -- to each face's snowpack with sno::_adj_snow.
//   src/modules/snobal.cpp:363-387        mass = is_nan(drift_mass) ? 0 : drift_mass; erosion removes depth at the pack's density,
//                                         deposition adds depth at `drift_density`; _adj_snow(mass / density, mass)
//   src/modules/snobal.cpp:389-408        _adj_snow(delta_avalanche_snowdepth / area, delta_avalanche_mass / area * 1000)
//   third_party/snobal/sno.cpp:2527-2575  _adj_snow      :2617-2696  _adj_layers     :2366-2405  _calc_layers
//                             :1564-1580  _layer_mass    :2321-2329  _cold_content   snomacros.h:209,359,522,546
// One thread per face, SoA state in CHM face order (the caller's arrays, host or device), every field read once and written
// once: ≈ 312 B/face, HBM-bound.  The arithmetic is the reference's operation for operation with IEEE round-to-nearest
// intrinsics (no FMA contraction), so the result is bit-identical to the compiled sno.cpp (tests/test_gpu_snobal.py).
#pragma once
#include <cuda_runtime.h>

namespace pbsm3d {

struct SnowpackPtrs {  // field order = pbsm3d_snowpack (include/pbsm3d.h)
    double *z_s, *m_s, *rho;
    int* layer_count;
    double *z_s_0, *z_s_l, *m_s_0, *m_s_l, *cc_s, *cc_s_0, *cc_s_l, *T_s, *T_s_0, *T_s_l, *h2o_total, *h2o_vol, *h2o, *h2o_max, *h2o_sat;
};

struct SnobalConst {
    double drift_density;  // snobal.cpp:83
    double threshold;      // tstep_info[SMALL_TSTEP].threshold, snobal.cpp:190
    double max_z_s_0;      // "max_active_layer", snobal.cpp:101
};

constexpr double kSnoFreeze = 2.7316e2;       // FREEZE
constexpr double kSnoMaxDensity = 750.0;      // MAX_SNOW_DENSITY
constexpr double kSnoMinTemp = -75.0;         // MIN_SNOW_TEMP

__device__ __forceinline__ bool sno_is_nan(double v) { return fabs(__dsub_rn(v, -9999.0)) < 1e-5 || v != v; }  // module_base.hpp:471-479

// _cold_content: heat_stor(CP_ICE(temp), mass, temp - FREEZE) = cp * mass * tdiff, CP_ICE(t) = CAL_TO_J(0.024928 + 0.00176 t) / G_TO_KG(1)
__device__ __forceinline__ double sno_cold_content(double temp, double mass) {
    if (!(temp < kSnoFreeze)) return 0.0;
    const double cp = __ddiv_rn(__dmul_rn(__dadd_rn(0.024928, __dmul_rn(0.00176, temp)), 4.186798188), __dmul_rn(1.0, 0.001));
    return __dmul_rn(__dmul_rn(cp, mass), __dsub_rn(temp, kSnoFreeze));
}

// mode 0: drift (a = drift_mass in CHM order, or slot-ordered through iperm when `a_slots`); mode 1: avalanche (a = volume, b = mass
// as swe volume; `area` slot-ordered through iperm)
template <int MODE>
__global__ void __launch_bounds__(256) snobal_adj_snow_kernel(int T, SnowpackPtrs p, SnobalConst k, const double* __restrict__ a,
                                                              const double* __restrict__ b, const int* __restrict__ iperm,
                                                              int a_slots, const double* __restrict__ area,
                                                              double* __restrict__ swe_out, double* __restrict__ depth_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    double z_s = p.z_s[i], m_s = p.m_s[i], rho = p.rho[i];
    int layer_count = p.layer_count[i];
    double z_s_0 = p.z_s_0[i], z_s_l = p.z_s_l[i], m_s_0 = p.m_s_0[i], m_s_l = p.m_s_l[i];
    double cc_s = p.cc_s[i], cc_s_0 = p.cc_s_0[i], cc_s_l = p.cc_s_l[i];
    double T_s = p.T_s[i], T_s_0 = p.T_s_0[i], T_s_l = p.T_s_l[i];
    double h2o_total = p.h2o_total[i], h2o_vol = p.h2o_vol[i], h2o = p.h2o[i], h2o_max = p.h2o_max[i], h2o_sat = p.h2o_sat[i];

    double dz, dm;
    if (MODE == 0) {
        double mass = a[a_slots ? iperm[i] : i];
        mass = sno_is_nan(mass) ? 0.0 : mass;
        dz = __ddiv_rn(mass, mass < 0.0 ? rho : k.drift_density);
        dm = mass;
    } else {
        const double ar = area[iperm[i]];
        dz = __ddiv_rn(a[i], ar);
        dm = __dmul_rn(__ddiv_rn(b[i], ar), 1000.0);
    }
    // ---- _adj_snow
    z_s = __dadd_rn(z_s, dz);
    m_s = __dadd_rn(m_s, dm);
    if (m_s < 0 || z_s < 0) { m_s = cc_s = 0.0; m_s_0 = cc_s_0 = 0.0; }
    rho = z_s != 0.0 ? __ddiv_rn(m_s, z_s) : 0.0;
    bool adj_layers = dz != 0.0;
    if (rho > kSnoMaxDensity) {
        rho = kSnoMaxDensity;
        z_s = __ddiv_rn(m_s, rho);
        adj_layers = true;
    }
    if (adj_layers) {
        const int prev = layer_count;
        // ---- _calc_layers
        if (m_s <= k.threshold) {
            layer_count = 0;
            z_s = z_s_0 = z_s_l = 0.0;
        } else if (z_s < k.max_z_s_0) {
            layer_count = 1;
            z_s_0 = z_s;
            z_s_l = 0.0;
        } else {
            layer_count = 2;
            z_s_0 = k.max_z_s_0;
            z_s_l = __dsub_rn(z_s, z_s_0);
            if (__dmul_rn(z_s_l, rho) < k.threshold) {
                layer_count = 1;
                z_s_0 = z_s;
                z_s_l = 0.0;
            }
        }
        if (layer_count == 0) {
            rho = 0.0;
            if (m_s > 0.0) h2o_total = __dadd_rn(h2o_total, m_s);  // below threshold: this little bit of mass becomes water
            m_s = cc_s = 0.0;
            m_s_0 = cc_s_0 = 0.0;
            T_s = T_s_0 = kSnoMinTemp + kSnoFreeze;
            if (prev == 2) {
                m_s_l = cc_s_l = 0.0;
                T_s_l = kSnoMinTemp + kSnoFreeze;
            }
            h2o_vol = h2o = h2o_max = h2o_sat = 0.0;
        } else {
            m_s_0 = __dmul_rn(rho, z_s_0);
            m_s_l = layer_count == 2 ? __dmul_rn(rho, z_s_l) : 0.0;
            if (prev == 1 && layer_count == 2) {
                T_s_l = T_s;
                cc_s_l = sno_cold_content(T_s_l, m_s_l);
            } else if (prev == 2 && layer_count == 1) {
                T_s_l = kSnoMinTemp + kSnoFreeze;
                cc_s_l = 0.0;
            }
        }
    } else {
        // ---- _layer_mass only (a change of mass without a change of depth)
        if (layer_count == 0) {
            m_s_0 = 0.0;
            m_s_l = 0.0;
        } else {
            m_s_0 = __dmul_rn(rho, z_s_0);
            m_s_l = layer_count == 2 ? __dmul_rn(rho, z_s_l) : 0.0;
        }
    }
    p.z_s[i] = z_s; p.m_s[i] = m_s; p.rho[i] = rho; p.layer_count[i] = layer_count;
    p.z_s_0[i] = z_s_0; p.z_s_l[i] = z_s_l; p.m_s_0[i] = m_s_0; p.m_s_l[i] = m_s_l;
    p.cc_s[i] = cc_s; p.cc_s_0[i] = cc_s_0; p.cc_s_l[i] = cc_s_l;
    p.T_s[i] = T_s; p.T_s_0[i] = T_s_0; p.T_s_l[i] = T_s_l;
    p.h2o_total[i] = h2o_total; p.h2o_vol[i] = h2o_vol; p.h2o[i] = h2o; p.h2o_max[i] = h2o_max; p.h2o_sat[i] = h2o_sat;
    // what snobal hands back to PBSM3D's next step (snobal.cpp:468,491 write these after its energy balance; here: after the
    // mass adjustment, for callers that run PBSM3D without the energy balance)
    if (swe_out) swe_out[i] = m_s;
    if (depth_out) depth_out[i] = z_s;
}

}  // namespace pbsm3d
