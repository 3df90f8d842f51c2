"""CPU oracle (numpy, fp64) for the PBSM3D hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product (chm_b200/) never does and has no CPU fallback.

PARITY PINNED AGAINST THE REFERENCE'S OWN CODE.  The reference ships no test, golden vector or expected output
for PBSM3D and its build (CMake + Boost, CGAL, Armadillo, GSL, Trilinos, MeteoIO ...) cannot run
here — but the path's own sources (src/modules/PBSM3D.cpp, src/physics/Atmosphere.cpp, src/math/coordinates.cpp)
compile unmodified against small stand-in headers (oracle/refbuild → oracle/_ref/libchmref.so).  The committed golden
vectors (tests/golden/golden_*.npz, generator tests/golden/make_golden.py) are outputs of that library; this file
reproduces every assembled coefficient of both linear systems to <= 1e-13 and every output to <= 1e-11 over 16
config/vegetation/water/missing-value variants and two bundled meshes (tests/test_reference_pin.py), and is
re-checked against the live library on fresh random cases wherever the .so is present.  NOT pinned by reference
code (absent from /root/reference, restated): MeteoIO's stdDryAirDensity, the Belos/Ifpack2 Krylov iteration (replaced
by its mathematical contract), CGAL's fp64 geometry constructions.  Each function cites the file:line it follows.

Third-party arithmetic restated from published sources, not from this tree:
* MeteoIO ``mio::Atmosphere::stdDryAirDensity`` (PBSM3D.cpp:785; version unpinned in spack.yaml:29):
  rho = p_std(z)/(R_d T),  p_std(z) = 101325 (1 - 0.0065 R0 z /(288.15 (R0+z)))^(g/(0.0065 R_d)),
  R0 = 6356766 m, g = 9.80665, R_d = 287.058.   An assumption; it is a uniform scale on c_salt.
* Trilinos 15.0.0 Belos GMRES + Ifpack2 ILUT (LinearAlgebra.cpp:164-195,228-252): only the
  mathematical contract is used (solve A x = b from x0 = 0 to ||b-Ax||2/||b||2 <= 1e-8).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

KAPPA = 0.4  # PhysConst.h:31
RHO_ICE = 917.0  # PhysConst.h:39
Z_U_R = 50.0  # Atmosphere.h:31
Z0_SNOW = 0.01  # Snow.h:31
SUSP_THRESHOLD = 1e-12  # PBSM3D.hpp:417-418
DEP_THRESHOLD = 1e-12


@dataclass
class Config:
    """PBSM3D config keys and code defaults (PBSM3D.cpp:123-145, 223-258)."""
    nLayer: int = 10
    do_fixed_settling: bool = False
    settling_velocity: float = 0.5
    do_sublimation: bool = True
    do_lateral_diff: bool = True
    smooth_coeff: float = 820.0
    min_sd_trans: float = 0.1
    cutoff: float = 0.3
    snow_diffusion_const: float = 0.3
    rouault_diffusion_coef: bool = False
    enable_veg: bool = True
    iterative_subl: bool = False
    use_exp_fetch: bool = False
    use_tanh_fetch: bool = True
    use_PomLi_probability: bool = False
    z0_ustar_coupling: bool = False
    use_subgrid_topo: bool = False
    use_subgrid_topo_V2: bool = False
    use_R94_lambda: bool = True
    debug_output: bool = False

    @staticmethod
    def functional_test(nLayer: int = 10) -> "Config":
        """The block in functional_tests/mesh_versioning/json_mesh.json:82-99."""
        return Config(nLayer=nLayer, smooth_coeff=6500.0, do_fixed_settling=True, settling_velocity=0.5,
                      use_R94_lambda=False)


def is_nan(x):
    """module_base::is_nan (module_base.hpp:471-479): -9999 sentinel or NaN."""
    x = np.asarray(x, dtype=np.float64)
    return (np.abs(x - -9999.0) < 1e-5) | np.isnan(x)


def bearing_to_cartesian(bearing):
    """math::gis::bearing_to_cartesian (coordinates.cpp:112-131)."""
    h = 450.0 - bearing
    h = np.where(h > 360.0, h - 360.0, h)
    th = h * np.pi / 180.0
    return np.cos(th), np.sin(th)


def log_scale_wind(u, Z_in, Z_out, sd, z0=Z0_SNOW):
    """Atmosphere::log_scale_wind (Atmosphere.cpp:32-38)."""
    return u * np.log((Z_out - (sd + z0)) / z0) / np.log((Z_in - (sd + z0)) / z0)


def saturated_vapour_pressure(t_kelvin):
    """Atmosphere::saturatedVapourPressure (Atmosphere.cpp:62-80).  The branch test compares the
    KELVIN argument with 0, so the over-water branch is always taken (SURVEY a-note 4)."""
    TA = t_kelvin - 273.15
    return 611.21 * np.exp((17.502 * TA) / (240.97 + TA))


def std_dry_air_density(z, t_kelvin):
    """mio::Atmosphere::stdDryAirDensity – restated from MeteoIO (see module docstring)."""
    R0, g, Rd, lapse, T0, p0 = 6356766.0, 9.80665, 287.058, 0.0065, 288.15, 101325.0
    expo = g / (lapse * Rd)
    p = p0 * np.power(1.0 - ((lapse * R0 * z) / (T0 * (R0 + z))), expo)
    return p / (Rd * t_kelvin)


@dataclass
class State:
    """Per-face state that survives between steps (PBSM3D.hpp:367-408 `data`, face variables)."""
    sum_drift: np.ndarray
    sum_subl: np.ndarray
    drift_mass: np.ndarray  # stale on no-deposition steps (SURVEY a-note 8)
    pbsm_more_than_avail: np.ndarray  # never reset
    blowingsnow_probability: Optional[np.ndarray] = None  # face variable, written only on saltating faces with use_PomLi_probability


@dataclass
class Assembled:
    """Extruded-ELL view of the suspension system of one rank, local layout [L,T]."""
    diag: np.ndarray  # [L,T]
    lat: np.ndarray  # [3,L,T] coefficient of neighbour j in the same layer (0 where none)
    below: np.ndarray  # [L,T] (row z, column z-1); below[0] = 0
    above: np.ndarray  # [L,T] (row z, column z+1); above[L-1] = 0
    rhs: np.ndarray  # [L,T] (only z = 0 non-zero)
    u_z: np.ndarray  # [L,T]
    csubl: np.ndarray  # [L,T]
    Qsalt: np.ndarray  # [T]
    c_salt: np.ndarray
    saltation: np.ndarray  # bool [T] (after the c_salt clamp)
    hs: np.ndarray
    ustar: np.ndarray
    z0: np.ndarray
    prob: Optional[np.ndarray] = None  # [T] Pomeroy-Li probability where it was written this step, NaN elsewhere


class PBSM3DOracle:
    """Restatement of PBSM3D::init / PBSM3D::run for one (global or rank-local) mesh.

    `geo` is chm_b200.mesh.FaceGeometry, `neigh` local neighbour ids (-1 none, >= T ghost).
    """

    def __init__(self, cfg: Config, neigh, geo, global_id, n_global, params: Optional[Dict[str, np.ndarray]] = None,
                 is_water=None):
        # ---- PBSM3D::init (PBSM3D.cpp:221-398)
        for k in ("iterative_subl", "use_subgrid_topo", "use_subgrid_topo_V2", "z0_ustar_coupling", "debug_output"):
            if getattr(cfg, k):
                raise NotImplementedError(f"optional path {k} is not restated (SURVEY.md §8a-notes)")
        if cfg.use_exp_fetch and cfg.use_tanh_fetch:
            raise ValueError("PBSM3d: Cannot specify both exp_fetch and tanh_fetch")  # PBSM3D.cpp:132-135
        if cfg.settling_velocity < 0:
            raise ValueError("PBSM3D settling velocity must be positive")  # :239-242
        self.cfg = cfg
        self.L = int(cfg.nLayer)
        self.dz = 5.0 / float(cfg.nLayer)  # susp_depth / nLayer, :225-226
        self.l_max = 40.0
        self.neigh = np.asarray(neigh, dtype=np.int64)
        self.T = self.neigh.shape[0]
        self.geo = geo
        self.gid = np.asarray(global_id, dtype=np.int64)
        self.G = int(n_global)
        params = params or {}
        T = self.T
        # vegetation (:284-324): any face without vegetation info turns veg off for the whole run
        has_veg = ("CanopyHeight" in params) or ("landcover" in params) or ("canopyType" in params)
        self.enable_veg = bool(cfg.enable_veg and has_veg)
        if self.enable_veg:
            self.CanopyHeight = np.asarray(params["CanopyHeight"], dtype=np.float64)
            if cfg.use_R94_lambda:
                self.LAI = np.asarray(params["LAI"], dtype=np.float64)
                self.N = np.zeros(T)
                self.dv = np.zeros(T)
            else:
                self.LAI = np.zeros(T)
                self.N = np.asarray(params.get("stalk_number", np.ones(T)), dtype=np.float64)  # default 1
                self.dv = np.asarray(params.get("stalk_diameter", np.full(T, 0.8)), dtype=np.float64)  # default 0.8
        else:
            self.CanopyHeight = np.zeros(T)
            self.LAI = np.zeros(T)
            self.N = np.zeros(T)
            self.dv = np.zeros(T)
        self.is_water = np.zeros(T, dtype=bool) if is_water is None else np.asarray(is_water, dtype=bool)
        self.face_neigh = self.neigh >= 0  # [T,3]
        self.state = State(np.zeros(T), np.zeros(T), np.full(T, -9999.0), np.zeros(T), np.full(T, -9999.0))

    # ------------------------------------------------------------------ hot loop 1
    def assemble(self, F: Dict[str, np.ndarray], dt: float) -> Assembled:
        """Saltation + suspension assembly, PBSM3D.cpp:417-1410."""
        cfg, geo, T, L, dz = self.cfg, self.geo, self.T, self.L, self.dz
        fetch = np.asarray(F["fetch"], dtype=np.float64) if (cfg.use_exp_fetch or cfg.use_tanh_fetch) else np.full(T, 1000.0)
        uref = np.asarray(F["U_R"], dtype=np.float64)
        sd = np.asarray(F["snowdepthavg"], dtype=np.float64)
        sd = np.where(is_nan(sd), 0.0, sd)  # :446-447
        u2 = np.asarray(F["U_2m_above_srf"], dtype=np.float64)
        swe = np.asarray(F["swe"], dtype=np.float64)
        swe = np.where(is_nan(swe), 0.0, swe)  # :468-469
        Tc = np.asarray(F["t"], dtype=np.float64)
        phi = np.asarray(F["vw_dir"], dtype=np.float64)

        height_diff = np.maximum(0.0, self.CanopyHeight - sd)  # :473
        if not self.enable_veg:
            height_diff = np.zeros(T)
        ust_th = 0.35 + (1.0 / 150.0) * Tc + (1.0 / 8200.0) * Tc * Tc  # :671-672

        cand = (height_diff <= cfg.cutoff) & (sd >= cfg.min_sd_trans) & (~self.is_water)  # :677
        if cfg.use_R94_lambda:
            lam = np.where(cand, 0.5 * self.LAI * height_diff, 0.0)  # :684
        else:
            lam = np.where(cand, self.N * self.dv * height_diff, 0.0)  # :686
        ustar_c = u2 * KAPPA / np.log(2.0 / 0.0002)  # :720
        salt = cand & (ustar_c >= ust_th)  # :723-725
        z0 = np.full(T, Z0_SNOW)  # :740 / :748, then max(Z0_SNOW, z0) :753
        ustar = np.where(salt, ustar_c, np.maximum(0.01, KAPPA * uref / np.log(Z_U_R / z0)))  # :749
        ustar = np.maximum(0.01, ustar)  # :754
        hs = np.where(salt, 0.08436 * np.power(ustar, 1.27), 0.0)  # :761-765

        t = Tc + 273.15  # :775
        vx, vy = bearing_to_cartesian(phi)
        vx, vy = -vx, -vy  # :881
        # saltation block :781-920 (only faces with salt == True at entry)
        rho_f = std_dry_air_density(geo.cz[:T], t)
        mB = 0.16 * 202.0
        tau_n_ratio = (mB * lam) / (1.0 + mB * lam)  # :807
        with np.errstate(all="ignore"):
            c_salt = rho_f / (3.29 * ustar) * (1.0 - tau_n_ratio - (ust_th * ust_th) / (ustar * ustar))  # :814-816
        bad = (c_salt < 0) | np.isnan(c_salt)  # :820-824
        c_salt = np.where(bad, 0.0, c_salt)
        salt_after = salt & (~bad)
        if cfg.use_exp_fetch:
            c_salt = np.where(fetch < 500.0, c_salt * (1.0 - np.exp(-3.0 * fetch / 500.0)), c_salt)  # :833-838
        elif cfg.use_tanh_fetch:
            Lc = 0.5 * np.tanh(0.1333333333e-1 * 300.0 - 2.0) + 0.5  # fetch_ref, not fetch: :839-845
            c_salt = np.where(fetch <= 300.0, c_salt * Lc, c_salt)
        prob = None
        if cfg.use_PomLi_probability:  # Pomeroy & Li 2000 upscaled probability of blowing snow, :848-866
            A = np.asarray(F["p_snow_hours"], dtype=np.float64)  # hours since the last snowfall
            z10 = 10.0 + sd
            with np.errstate(all="ignore"):
                u10 = np.where(z10 < Z_U_R, uref * np.log((z10 - (sd + Z0_SNOW)) / Z0_SNOW) / np.log((Z_U_R - (sd + Z0_SNOW)) / Z0_SNOW), uref)  # :451-463
                u_mean = 11.2 + 0.365 * Tc + 0.00706 * Tc * Tc + 0.9 * np.log(A)  # eqn 10
                delta = 0.145 * Tc + 0.00196 * Tc * Tc + 4.3  # eqn 11
                z0v = (self.N * self.dv * height_diff) / 2.0  # eqn 14
                us = u10 / np.sqrt(1.0 + 340.0 * z0v)  # eqn 13
                Pu10 = 1.0 / (1.0 + np.exp((np.sqrt(np.pi) * (u_mean - us)) / delta))  # eqn 12
            prob = np.where(salt, Pu10, np.nan)  # written on every face that entered the saltation block
            c_salt = c_salt * Pu10
        uhs = 2.8 * ust_th  # :874
        Qsalt = c_salt * uhs * hs  # :877
        mass = np.zeros(T)
        for j in range(3):  # :895-900
            udotm = vx * geo.nx[j] + vy * geo.ny[j]
            mass = mass + (-geo.elen[j] * Qsalt * udotm)
        mass = mass / geo.area * dt  # :902
        reset = (mass < 0) & (np.abs(mass) > swe)  # :910-919 (saltation flag NOT cleared)
        c_salt = np.where(reset, 0.0, c_salt)
        Qsalt = np.where(reset, 0.0, Qsalt)
        # faces that never entered the block
        c_salt = np.where(salt, c_salt, 0.0)
        Qsalt = np.where(salt, Qsalt, 0.0)
        saltation = salt_after

        rh = np.asarray(F["rh"], dtype=np.float64) / 100.0  # :927
        es = saturated_vapour_pressure(t)  # :928
        rho_p = RHO_ICE

        diag = np.zeros((L, T))
        lat = np.zeros((3, L, T))
        below = np.zeros((L, T))
        above = np.zeros((L, T))
        rhs = np.zeros((L, T))
        u_z_all = np.zeros((L, T))
        csubl_all = np.zeros((L, T))
        A = [geo.elen[j] * dz for j in range(3)]  # d.A[j] :361-362
        area = geo.area
        for z in range(L):
            cz = z * dz + hs + dz / 2.0  # :937
            hz = cz + sd  # :943
            in_canopy = cz < height_diff
            with np.errstate(all="ignore"):
                u_log = np.maximum(0.01, log_scale_wind(uref, Z_U_R, hz, sd, z0))  # :978
            u_above = np.where(hz < Z_U_R, u_log, np.maximum(0.01, uref))  # :975-983
            u_z = np.where(saltation & in_canopy, 2.8 * ust_th, np.where(in_canopy, 0.01, u_above))  # :948-984
            u_z_all[z] = u_z

            rm = 4.6e-5 * np.power(cz, -0.258)  # :1003
            mm_alpha = 4.08 + 12.6 * cz  # :1012
            mm = 4.0 / 3.0 * np.pi * rho_p * rm * rm * rm * (1.0 + 3.0 / mm_alpha + 2.0 / (mm_alpha * mm_alpha))  # :1013-1014
            r_z = np.power((3.0 * mm) / (4 * np.pi * rho_p), 0.3333333)  # :1017
            xrz = 0.005 * np.power(u_z, 1.36)  # :1021
            if cfg.do_fixed_settling:
                omega = np.full(T, cfg.settling_velocity)  # :1023
            else:
                omega = 1.1e7 * np.power(r_z, 1.8)  # :1027
            Vr = omega + 3.0 * xrz * np.cos(np.pi / 4.0)  # :1032
            v = 1.88e-5
            Re = 2.0 * r_z * Vr / v  # :1036
            Nu = 1.79 + 0.606 * np.power(Re, 0.5)  # :1039
            Sh = Nu
            D = 2.06e-5 * np.power(t / 273.15, 1.75)  # :1044
            lambda_t = 0.000063 * t + 0.00673  # :1049-1050
            Ls = 2.838e6  # :1057
            M = 18.01
            R = 8313.0
            sigma = (rh - 1.0) * (1.019 + 0.027 * np.log(cz))  # :1108-1109
            rho = (M * es) / (R * t)  # :1111
            Qr = 0.9 * np.pi * rm * rm * 120.0  # :1114
            dmdtz = Sh * rho * D * (6.283185308 * Nu * R * r_z * sigma * t * t * lambda_t - Ls * M * Qr + Qr * R * t) / \
                (D * Ls * Sh * (Ls * M - R * t) * rho + lambda_t * t * t * Nu * R)  # :1121-1123
            csubl = dmdtz / mm  # :1130
            if not cfg.do_sublimation:
                csubl = np.zeros(T)  # :1228-1231
            csubl_all[z] = csubl

            if cfg.do_lateral_diff:
                alpha = [A[a] * 0.00001 for a in range(3)]  # :1143-1154
            else:
                alpha = [np.zeros(T) for _ in range(3)]
            lmix = KAPPA * (cz + z0) * self.l_max / (KAPPA * (cz + z0) + self.l_max)  # :1156
            w = omega
            diffusion_coeff = cfg.snow_diffusion_const
            if cfg.rouault_diffusion_coef:
                diffusion_coeff = 1.0 / (1.0 + (1.0 * w * w) / (1.56 * ustar * ustar))  # :1167-1172
            K = diffusion_coeff * ustar * lmix  # :1180
            alpha3 = area * K / dz  # :1185
            alpha4 = area * K / dz  # :1187

            nrm = np.sqrt(vx * vx + vy * vy)
            s = u_z / nrm  # :1200
            ux, uy = vx * s, vy * s
            udotm = [ux * geo.nx[j] + uy * geo.ny[j] for j in range(3)]
            udotm3 = -w  # (0,0,1)·(ux,uy,-w)
            udotm4 = w  # (0,0,-1)·(ux,uy,-w)
            V = area * dz  # :1222
            V = V / 5.0  # :1226
            Vc = V * csubl

            d = np.zeros(T)
            for f in range(3):  # :1236-1283
                out = udotm[f] > 0
                has = self.face_neigh[:, f]
                d_has_out = Vc - A[f] * udotm[f] - alpha[f]
                d_no_out = -0.1e-1 * alpha[f] - 1.0 * A[f] * udotm[f] + Vc
                d_has_in = Vc - alpha[f]
                d_no_in = -0.1e-1 * alpha[f] - 0.99 * A[f] * udotm[f] + Vc
                d = d + np.where(out, np.where(has, d_has_out, d_no_out), np.where(has, d_has_in, d_no_in))
                lat[f, z] = np.where(has, np.where(out, alpha[f], -A[f] * udotm[f] + alpha[f]), 0.0)

            def top_face():
                if_out = (Vc - area * udotm3 - alpha3, alpha3)
                if_in = (Vc - alpha3, -area * udotm3 + alpha3)
                o = udotm3 > 0
                return np.where(o, if_out[0], if_in[0]), np.where(o, if_out[1], if_in[1])

            def bottom_face():
                if_out = (Vc - area * udotm4 - alpha4, alpha4)
                if_in = (Vc - alpha4, -area * udotm4 + alpha4)
                o = udotm4 > 0
                return np.where(o, if_out[0], if_in[0]), np.where(o, if_out[1], if_in[1])

            if z == 0:  # :1286-1322
                alpha4p = area * K / (hs / 2.0 + dz / 2.0)  # :1289
                d = d + (Vc - area * udotm4 - alpha4p)  # :1295-1296
                rhs[z] = -alpha4p * c_salt  # :1298-1299
                dt_, up = top_face()
                d = d + dt_
                above[z] = up
            elif z == L - 1:  # :1323-1368, cprecip = 0
                o = udotm3 > 0
                d = d + np.where(o, Vc - area * udotm3 - alpha3, Vc - alpha3)
                db_, lo = bottom_face()
                d = d + db_
                below[z] = lo
            else:  # :1369-1404
                dt_, up = top_face()
                d = d + dt_
                above[z] = up
                db_, lo = bottom_face()
                d = d + db_
                below[z] = lo
            diag[z] = d
        return Assembled(diag, lat, below, above, rhs, u_z_all, csubl_all, Qsalt, c_salt, saltation, hs, ustar, z0, prob)

    # ------------------------------------------------------------------ matrices in the reference ordering
    def suspension_csr(self, asm: Assembled, n_cols_ghost: int = 0):
        """Rows in the local layout z*T + local_id (LinearAlgebra.cpp:81); columns likewise, with
        ghost faces (neigh >= T) mapped to z*T_ext + id when n_cols_ghost > 0.  For a global mesh this
        equals the reference's global numbering z*G + cell_global_id (:51-55,:80)."""
        T, L = self.T, self.L
        Text = T + n_cols_ghost
        rows, cols, vals = [], [], []
        idx = np.arange(T)
        for z in range(L):
            r = z * T + idx
            c0 = z * Text + idx
            rows.append(r); cols.append(c0); vals.append(asm.diag[z])
            for f in range(3):
                has = self.face_neigh[:, f]
                rows.append(r[has]); cols.append(z * Text + self.neigh[has, f]); vals.append(asm.lat[f, z][has])
            if z > 0:
                rows.append(r); cols.append(c0 - Text); vals.append(asm.below[z])
            if z < L - 1:
                rows.append(r); cols.append(c0 + Text); vals.append(asm.above[z])
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(L * T, L * Text))
        return A

    def flux_integrate(self, x: np.ndarray, asm: Assembled, dt: float):
        """PBSM3D.cpp:1467-1503."""
        L, T, dz = self.L, self.T, self.dz
        c = x.reshape(L, T)
        c = np.where((c < 0) | is_nan(c), 0.0, c)  # :1478
        Qsusp = np.zeros(T)
        Qsubl = np.zeros(T)
        for z in range(L):
            Qsusp = Qsusp + c[z] * asm.u_z[z] * dz  # :1483
            Qsubl = Qsubl + asm.csubl[z] * c[z] * dz  # :1494
        return Qsusp, Qsubl

    def deposition_system(self, F, Qsusp_ext, Qsalt_ext):
        """PBSM3D.cpp:1516-1658.  *_ext are [T + n_ghost] (ghost values after the halo exchange)."""
        cfg, geo, T = self.cfg, self.geo, self.T
        eps = cfg.smooth_coeff
        phi = np.asarray(F["vw_dir"], dtype=np.float64)
        vx, vy = bearing_to_cartesian(phi)
        vx, vy = -vx, -vy
        diag = geo.area.copy()  # matrixReplaceGlobalValues(row,row,V) :1546
        off = np.zeros((3, T))
        rhs = np.zeros(T)
        for j in range(3):
            udotm = vx * geo.nx[j] + vy * geo.ny[j]  # :1552
            E = geo.elen[j]
            has = self.face_neigh[:, j]
            nb = np.where(has, self.neigh[:, j], 0)
            own_t, own_s = Qsusp_ext[:T], Qsalt_ext[:T]
            nb_t = Qsusp_ext[nb]
            nb_s = Qsalt_ext[nb]
            nb_s = np.where(is_nan(nb_s), 0.0, nb_s)  # :1586-1592
            use_nb = (~(udotm > 0)) & has
            Qtj = np.where(use_nb, nb_t, own_t)
            Qsj = np.where(use_nb, nb_s, own_s)
            coef = np.where(has, eps * E / geo.dx[j], 0.0)  # :1613-1627
            diag = diag + coef
            off[j] = -coef
            rhs = rhs + (-E * (Qtj + Qsj) * udotm)  # :1631,1656
        return diag, off, rhs

    def deposition_csr(self, diag, off, n_cols_ghost: int = 0):
        T = self.T
        idx = np.arange(T)
        rows, cols, vals = [idx], [idx], [diag]
        for j in range(3):
            has = self.face_neigh[:, j]
            rows.append(idx[has]); cols.append(self.neigh[has, j]); vals.append(off[j][has])
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(T, T + n_cols_ghost))

    def drift_update(self, q, F, saltation, dt):
        """PBSM3D.cpp:1710-1740."""
        st = self.state
        q = np.where(is_nan(q), 0.0, q)
        mass = q * dt
        swe = np.asarray(F["swe"], dtype=np.float64)
        swe = np.where(is_nan(swe), 0.0, swe)
        over = (mass < 0) & (np.abs(mass) > swe)
        st.pbsm_more_than_avail = np.where(over, 1.0, st.pbsm_more_than_avail)
        mass = np.where(over, -swe, mass)
        mass = np.where((mass < 0) & (~saltation), 0.0, mass)
        st.drift_mass = mass
        st.sum_drift = st.sum_drift + mass

    # ------------------------------------------------------------------ one timestep, single rank / global mesh
    def step(self, F: Dict[str, np.ndarray], dt: float, solver: str = "direct", tol: float = 1e-8) -> Dict[str, np.ndarray]:
        """PBSM3D::run for a global (unpartitioned) mesh."""
        if self.neigh.max(initial=-1) >= self.T:
            raise ValueError("step() wants a global mesh; partitioned runs are composed in the tests")
        asm = self.assemble(F, dt)
        T, L = self.T, self.L
        info = {"susp_iters": 0, "dep_iters": 0}
        rhs_max = np.abs(asm.rhs).max(initial=0.0)  # getRhsMax :1424
        suspension_present = rhs_max > SUSP_THRESHOLD
        x = np.zeros(L * T)
        if suspension_present:
            A = self.suspension_csr(asm)
            x, info["susp_iters"] = solve(A, asm.rhs.reshape(-1), solver, tol)
        Qsusp, Qsubl = self.flux_integrate(x, asm, dt)
        st = self.state
        if asm.prob is not None:
            st.blowingsnow_probability = np.where(np.isnan(asm.prob), st.blowingsnow_probability, asm.prob)
        Qsubl_mass = Qsubl * dt
        st.sum_subl = st.sum_subl + Qsubl_mass
        diag, off, drhs = self.deposition_system(F, Qsusp, asm.Qsalt)
        deposition_present = suspension_present and (np.abs(drhs).max(initial=0.0) > DEP_THRESHOLD)  # :1661-1664
        q = np.zeros(T)
        if deposition_present:
            Ad = self.deposition_csr(diag, off)
            q, info["dep_iters"] = solve(Ad, drhs, solver, tol)
            self.drift_update(q, F, asm.saltation, dt)
        return {"c": x.reshape(L, T), "Qsusp": Qsusp, "Qsalt": asm.Qsalt, "Qsubl": Qsubl, "Qsubl_mass": Qsubl_mass,
                "sum_subl": st.sum_subl.copy(), "drift_mass": st.drift_mass.copy(), "sum_drift": st.sum_drift.copy(),
                "pbsm_more_than_avail": st.pbsm_more_than_avail.copy(), "blowingsnow_probability": st.blowingsnow_probability.copy(),
                "q_dep": q, "asm": asm,
                "dep": (diag, off, drhs), "suspension_present": suspension_present,
                "deposition_present": deposition_present, **info}


# ---------------------------------------------------------------------- linear solves
def gmres_right(A, b, M_solve, tol=1e-8, restart=30, maxiter=1000):
    """Restarted, right-preconditioned GMRES from x0 = 0 with the reference's stopping rule
    (LinearAlgebra.cpp:164-168: Num Blocks 30, Maximum Iterations 1000, Convergence Tolerance 1e-8;
    Belos tests the implicit residual scaled by ||r0|| = ||b||).  Returns (x, iterations)."""
    n = b.shape[0]
    x = np.zeros(n)
    bnorm = np.linalg.norm(b)
    if bnorm == 0:
        return x, 0
    its = 0
    while its < maxiter:
        r = b - A @ x
        beta = np.linalg.norm(r)
        if beta / bnorm <= tol:
            break
        V = np.zeros((restart + 1, n))
        H = np.zeros((restart + 1, restart))
        cs = np.zeros(restart)
        sn = np.zeros(restart)
        g = np.zeros(restart + 1)
        g[0] = beta
        V[0] = r / beta
        k_used = 0
        for k in range(restart):
            w = A @ M_solve(V[k])
            for i in range(k + 1):  # modified Gram-Schmidt
                H[i, k] = np.dot(w, V[i])
                w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] > 0:
                V[k + 1] = w / H[k + 1, k]
            for i in range(k):
                tmp = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = tmp
            den = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / den, H[k + 1, k] / den
            H[k, k] = den
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_used = k + 1
            if abs(g[k + 1]) / bnorm <= tol or its >= maxiter:
                break
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        x = x + M_solve(V[:k_used].T @ y)
    return x, its


def solve(A, b, solver="direct", tol=1e-8):
    """'direct' = scipy splu (ground truth); 'gmres_ilu' = GMRES(30) + incomplete LU with the
    reference's drop tolerance / fill (scipy spilu standing in for Ifpack2 ILUT, LinearAlgebra.cpp:178-186)."""
    A = A.tocsc()
    if solver == "direct":
        return spla.splu(A).solve(b), 1
    if solver == "gmres_ilu":
        ilu = spla.spilu(A, drop_tol=1e-4, fill_factor=3.0)
        return gmres_right(A.tocsr(), b, ilu.solve, tol=tol)
    raise ValueError(solver)
