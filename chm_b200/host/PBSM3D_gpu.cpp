// See PBSM3D_gpu.hpp.  Reference being replaced: src/modules/PBSM3D.cpp (ctor :103-219, init :221-398,
// run :400-1748, checkpoint :1753-1773) and src/math/LinearAlgebra.cpp (NearestNeighborProblem).
#include "PBSM3D_gpu.hpp"

#include <algorithm>
#include <cstring>

REGISTER_MODULE_CPP(PBSM3D_gpu);

namespace
{
void check(int rc)
{
    if (rc != PBSM3D_OK)
    {
        CHM_THROW_EXCEPTION(module_error, std::string("PBSM3D_gpu: ") + pbsm3d_last_error());
    }
}
pbsm3d_handle* g_shared_handle = nullptr; // one PBSM3D_gpu instance per process (core.cpp creates each module once)
} // namespace

pbsm3d_handle* PBSM3D_gpu::shared_handle() { return g_shared_handle; }

PBSM3D_gpu::PBSM3D_gpu(config_file cfg) : module_base("PBSM3D_gpu", parallel::domain, cfg)
{
    // identical dependency declarations to PBSM3D::PBSM3D (PBSM3D.cpp:105-202), except that with "fuse_providers"
    // U_2m_above_srf and fetch are derived on the device from U_R / snowdepthavg / vw_dir (scale_wind_vert.cpp, fetchr.cpp)
    // and no longer come from other modules
    _fuse = cfg.get("fuse_providers", false);
    if (!_fuse)
        depends("U_2m_above_srf");
    depends("vw_dir");
    depends("swe");
    depends("t");
    depends("rh");
    depends("U_R");

    provides("pbsm_more_than_avail");
    provides("global_cell_id");

    pbsm3d_config_defaults(&_c);
    _c.use_exp_fetch = cfg.get("use_exp_fetch", false);
    _c.use_tanh_fetch = cfg.get("use_tanh_fetch", true);
    _c.use_PomLi_probability = cfg.get("use_PomLi_probability", false);
    _c.z0_ustar_coupling = cfg.get("z0_ustar_coupling", false);
    _c.use_subgrid_topo = cfg.get("use_subgrid_topo", false);
    _c.use_subgrid_topo_V2 = cfg.get("use_subgrid_topo_V2", false);

    if (_c.use_exp_fetch && _c.use_tanh_fetch)
    {
        CHM_THROW_EXCEPTION(module_error, "PBSM3d: Cannot specify both exp_fetch and tanh_fetch");
    }
    _use_fetch = _c.use_exp_fetch || _c.use_tanh_fetch;
    if (_use_fetch)
    {
        if (!_fuse)
            depends("fetch");
    }
    else
        depends("p_snow_hours");
    _c.use_R94_lambda = cfg.get("use_R94_lambda", true);

    provides("blowingsnow_probability");
    _c.debug_output = cfg.get("debug_output", false);

    provides("Qsubl");
    provides("Qsubl_mass");
    provides("sum_subl");
    provides("drift_mass");
    provides("Qsusp");
    provides("Qsalt");
    provides("sum_drift");
}

void PBSM3D_gpu::init(mesh& domain)
{
    // config keys and defaults of PBSM3D::init (PBSM3D.cpp:223-258)
    _c.nLayer = cfg.get("nLayer", 10);
    _c.do_fixed_settling = cfg.get("do_fixed_settling", false);
    _c.settling_velocity = cfg.get("settling_velocity", 0.5);
    _c.do_sublimation = cfg.get("do_sublimation", true);
    _c.do_lateral_diff = cfg.get("do_lateral_diff", true);
    _c.smooth_coeff = cfg.get("smooth_coeff", 820);
    _c.min_sd_trans = cfg.get("min_sd_trans", 0.1);
    _c.cutoff = cfg.get("cutoff", 0.3);
    _c.snow_diffusion_const = cfg.get("snow_diffusion_const", 0.3);
    _c.rouault_diffusion_coef = cfg.get("rouault_diffusion_coef", false);
    _c.enable_veg = cfg.get("enable_veg", true);
    _c.iterative_subl = cfg.get("iterative_subl", false);
    // not CHM keys: the reference hard-codes them in LinearAlgebra.cpp:164-168
    _c.tolerance = cfg.get("tolerance", 1e-8);
    _c.max_iterations = cfg.get("max_iterations", 1000);
    _c.solver = cfg.get("solver", (int)PBSM3D_SOLVER_AUTO);
    _c.deposition_solver = cfg.get("deposition_solver", (int)PBSM3D_DEP_AUTO);
    _c.fp32_sweep_streams = cfg.get("fp32_sweep_streams", true);

    const size_t ntri = _ntri = domain->size_faces();

    // ---- flatten the triangulation: owned faces in domain->face(i) order, then the NEIGH ghosts they touch,
    //      sorted by cell_global_id (the order triangulation::_ghost_neighbors already has, triangulation.cpp:1755-1762)
    std::vector<mesh_elem> ghosts;
    std::vector<int32_t> neigh(3 * ntri, -1);
    {
        std::vector<mesh_elem> seen;
        for (size_t i = 0; i < ntri; i++)
        {
            auto face = domain->face(i);
            for (int j = 0; j < 3; ++j)
            {
                auto n = face->neighbor(j);
                if (n != nullptr && n->is_ghost)
                    seen.push_back(n);
            }
        }
        std::sort(seen.begin(), seen.end(),
                  [](const mesh_elem& a, const mesh_elem& b) { return a->cell_global_id < b->cell_global_id; });
        seen.erase(std::unique(seen.begin(), seen.end(),
                               [](const mesh_elem& a, const mesh_elem& b) { return a->cell_global_id == b->cell_global_id; }),
                   seen.end());
        ghosts.swap(seen);
    }
    const size_t nghost = ghosts.size();
    std::vector<int64_t> gid(ntri + nghost);
    std::vector<int32_t> owner(nghost);
    std::vector<double> verts(9 * (ntri + nghost));
    auto put_vertices = [&](size_t k, mesh_elem f) {
        for (int v = 0; v < 3; ++v)
        {
            auto p = f->vertex(v)->point();
            verts[9 * k + 3 * v + 0] = p.x();
            verts[9 * k + 3 * v + 1] = p.y();
            verts[9 * k + 3 * v + 2] = p.z();
        }
    };
    for (size_t g = 0; g < nghost; ++g)
    {
        gid[ntri + g] = ghosts[g]->cell_global_id;
        owner[g] = ghosts[g]->owner;
        put_vertices(ntri + g, ghosts[g]);
    }
    auto ghost_index = [&](mesh_elem n) -> int32_t {
        auto it = std::lower_bound(ghosts.begin(), ghosts.end(), n, [](const mesh_elem& a, const mesh_elem& b) {
            return a->cell_global_id < b->cell_global_id;
        });
        return (int32_t)(ntri + (it - ghosts.begin()));
    };

    bool has_area = true, has_veg = true, any_wveg = false, has_wlai = true;
    std::vector<double> area(ntri), canopy(ntri), lai(ntri), sn(ntri, 1.0), sdv(ntri, 0.8);
    std::vector<double> wcanopy(_fuse ? ntri : 0, 0.0), wlai(_fuse ? ntri : 0, 0.0);
    std::vector<uint8_t> water(ntri, 0);
    for (size_t i = 0; i < ntri; i++)
    {
        auto face = domain->face(i);
        gid[i] = face->cell_global_id;
        put_vertices(i, face);
        for (int j = 0; j < 3; ++j)
        {
            auto n = face->neighbor(j);
            if (n == nullptr)
                neigh[3 * i + j] = -1;
            else if (n->is_ghost)
                neigh[3 * i + j] = ghost_index(n);
            else
                neigh[3 * i + j] = (int32_t)n->cell_local_id;
        }
        if (face->has_parameter("area"_s))
            area[i] = face->get_area();
        else
            has_area = false;
        water[i] = is_water(face) ? 1 : 0;
        if (_fuse && face->has_vegetation())
        { // the providers look at vegetation face by face (scale_wind_vert.cpp:60-63, fetchr.cpp:63-72,86-90)
            any_wveg = true;
            wcanopy[i] = face->veg_attribute("CanopyHeight");
            try
            {
                wlai[i] = face->veg_attribute("LAI");
            }
            catch (module_error& e)
            {
                has_wlai = false;
            }
        }
        if (!face->has_vegetation())
            has_veg = false; // one face without vegetation data turns veg off globally (PBSM3D.cpp:317-324)
        else if (_c.enable_veg && has_veg)
        {
            canopy[i] = face->veg_attribute("CanopyHeight");
            if (_c.use_R94_lambda)
                lai[i] = face->veg_attribute("LAI");
            else
            {
                try
                {
                    sn[i] = face->veg_attribute("stalk_number");
                    sdv[i] = face->veg_attribute("stalk_diameter");
                }
                catch (module_error& e)
                {
                    sn[i] = 1;
                    sdv[i] = 0.8; // PBSM3D.cpp:303-312
                }
            }
        }
        (*face)["sum_drift"_s] = 0;
    }

    pbsm3d_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.n_global = (int64_t)domain->size_global_faces();
    m.n_local = (int32_t)ntri;
    m.n_ghost = (int32_t)nghost;
    m.global_id = gid.data();
    m.ghost_owner = nghost ? owner.data() : nullptr;
    m.neigh = neigh.data();
    m.vertices = verts.data();
    m.area = has_area ? area.data() : nullptr;
    const bool veg = _c.enable_veg && has_veg;
    m.canopy_height = veg ? canopy.data() : nullptr;
    m.lai = (veg && _c.use_R94_lambda) ? lai.data() : nullptr;
    m.stalk_number = (veg && !_c.use_R94_lambda) ? sn.data() : nullptr;
    m.stalk_diameter = (veg && !_c.use_R94_lambda) ? sdv.data() : nullptr;
    m.is_water = water.data();
    m.is_geographic = domain->is_geographic() ? 1 : 0; // triangulation.cpp:98-101; selects math::gis::distance (core.cpp:809-821)
    if (_fuse && any_wveg && !veg)
    { // PBSM3D's own vegetation is off (disabled, or a face lacks the data) but the providers still see the canopy where it exists
        _c.enable_veg = 0;
        m.canopy_height = wcanopy.data();
        m.lai = has_wlai ? wlai.data() : nullptr;
    }
    else if (_fuse && veg && !m.lai && has_wlai)
        m.lai = wlai.data(); // stalk-based lambda for PBSM3D, LAI for scale_wind_vert's canopy profile

    pbsm3d_comm comm;
    std::memset(&comm, 0, sizeof(comm));
    pbsm3d_comm* pcomm = nullptr;
    char uid[128];
#ifdef USE_MPI
    comm.rank = domain->_comm_world.rank();
    comm.n_ranks = domain->_comm_world.size();
    if (comm.n_ranks > 1)
    {
        if (comm.rank == 0)
            check(pbsm3d_nccl_unique_id(uid));
        boost::mpi::broadcast(domain->_comm_world, uid, 128, 0);
        comm.nccl_unique_id = uid;
        pcomm = &comm;
    }
#endif
    // one GPU per rank of a node: local rank = rank modulo the GPUs visible to this process
    int device = cfg.get("device", -1);
    if (device < 0)
        device = pcomm ? comm.rank % std::max(1, cfg.get("gpus_per_node", 8)) : 0;
    check(pbsm3d_create(&_c, &m, device, pcomm, &_h));
    g_shared_handle = _h;
    if (_fuse)
    { // the provider modules' own config keys (scale_wind_vert.cpp:161, fetchr.cpp:34-43)
        pbsm3d_wind_config w;
        pbsm3d_wind_config_defaults(&w);
        w.ignore_canopy = cfg.get("ignore_canopy", false);
        w.fetch_steps = cfg.get("steps", 10);
        w.fetch_max_distance = cfg.get("max_distance", 1000.0);
        w.fetch_I = cfg.get("I", 0.06);
        w.fetch_incl_veg = cfg.get("incl_veg", true);
        check(pbsm3d_set_providers(_h, &w));
    }

    _stage = (double*)pbsm3d_host_alloc(18 * ntri * sizeof(double));
    if (!_stage)
        CHM_THROW_EXCEPTION(module_error, std::string("PBSM3D_gpu: ") + pbsm3d_last_error());
    std::memset(_stage, 0, 18 * ntri * sizeof(double));
    double** slots[18] = {&_U_R,   &_U2,    &_sd,    &_swe,        &_t,        &_rh,         &_vw_dir,   &_fetch,     &_psh,
                          &_Qsalt, &_Qsusp, &_Qsubl, &_Qsubl_mass, &_sum_subl, &_drift_mass, &_sum_drift, &_more,     &_prob};
    for (int k = 0; k < 18; ++k)
        *slots[k] = _stage + (size_t)k * ntri;
}

void PBSM3D_gpu::run(mesh& domain)
{
    const size_t ntri = _ntri;
    // gather: the eight variables PBSM3D::run reads (PBSM3D.cpp:436-449,468,670,880,927)
#pragma omp parallel for
    for (size_t i = 0; i < ntri; i++)
    {
        auto face = domain->face(i);
        _U_R[i] = (*face)["U_R"_s];
        if (!_fuse)
            _U2[i] = (*face)["U_2m_above_srf"_s];
        _sd[i] = (*face)["snowdepthavg"_s];
        _swe[i] = (*face)["swe"_s];
        _t[i] = (*face)["t"_s];
        _rh[i] = (*face)["rh"_s];
        _vw_dir[i] = (*face)["vw_dir"_s];
        if (_use_fetch && !_fuse)
            _fetch[i] = (*face)["fetch"_s];
        if (_c.use_PomLi_probability)
            _psh[i] = (*face)["p_snow_hours"_s]; // read even when only "fetch" was declared (PBSM3D.cpp:137-141 vs :853)
    }
    pbsm3d_forcing f{_U_R, _fuse ? nullptr : _U2, _sd, _swe, _t, _rh, _vw_dir, (_use_fetch && !_fuse) ? _fetch : nullptr,
                     _c.use_PomLi_probability ? _psh : nullptr};
    pbsm3d_outputs o{_Qsalt, _Qsusp, _Qsubl, _Qsubl_mass, _sum_subl, _drift_mass, _sum_drift, _more,
                     _c.use_PomLi_probability ? _prob : nullptr};
    check(pbsm3d_step(_h, global_param->dt(), &f, &o, &_stats));
    SPDLOG_DEBUG("  suspension iterations: {} residual: {}", _stats.suspension_iterations, _stats.suspension_residual);
    SPDLOG_DEBUG("  deposition iterations: {} residual: {}", _stats.deposition_iterations, _stats.deposition_residual);

    // scatter: what PBSM3D::run writes (PBSM3D.cpp:925,1496-1501,1727,1738-1739)
    const bool dep = _stats.deposition_present != 0;
#pragma omp parallel for
    for (size_t i = 0; i < ntri; i++)
    {
        auto face = domain->face(i);
        (*face)["Qsalt"_s] = _Qsalt[i];
        (*face)["Qsusp"_s] = _Qsusp[i];
        (*face)["Qsubl"_s] = _Qsubl[i];
        (*face)["Qsubl_mass"_s] = _Qsubl_mass[i];
        (*face)["sum_subl"_s] = _sum_subl[i];
        if (_c.use_PomLi_probability && _prob[i] != -9999.0)
            (*face)["blowingsnow_probability"_s] = _prob[i]; // only faces that have saltated carry a value (PBSM3D.cpp:861)
        if (dep)
        { // untouched on steps without a deposition solve (PBSM3D.cpp:1675,1742-1745)
            if (_more[i] > 0)
                (*face)["pbsm_more_than_avail"_s] = 1;
            (*face)["drift_mass"_s] = _drift_mass[i];
            (*face)["sum_drift"_s] = _sum_drift[i];
        }
    }
}

PBSM3D_gpu::~PBSM3D_gpu()
{
    if (g_shared_handle == _h)
        g_shared_handle = nullptr;
    pbsm3d_destroy(_h);
    pbsm3d_host_free(_stage);
}

void PBSM3D_gpu::checkpoint(mesh& domain, netcdf& chkpt)
{
    chkpt.create_variable1D("PBSM3D:sum_drift", domain->size_faces());
    for (size_t i = 0; i < domain->size_faces(); i++)
        chkpt.put_var1D("PBSM3D:sum_drift", i, (*domain->face(i))["sum_drift"_s]);
}

void PBSM3D_gpu::load_checkpoint(mesh& domain, netcdf& chkpt)
{
    std::vector<double> sd(domain->size_faces());
    for (size_t i = 0; i < domain->size_faces(); i++)
    {
        sd[i] = chkpt.get_var1D("PBSM3D:sum_drift", i);
        (*domain->face(i))["sum_drift"_s] = sd[i];
    }
    check(pbsm3d_set_state(_h, sd.data(), nullptr, nullptr, nullptr));
}
