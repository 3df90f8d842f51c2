#include "snow_slide_gpu.hpp"

#include <cstring>

REGISTER_MODULE_CPP(snow_slide_gpu);

namespace
{
void check(int rc)
{
    if (rc != 0)
        CHM_THROW_EXCEPTION(module_error, std::string("snow_slide_gpu: ") + pbsm3d_last_error());
}
} // namespace

snow_slide_gpu::snow_slide_gpu(config_file cfg) : module_base("snow_slide_gpu", parallel::domain, cfg)
{
    // snow_slide.cpp:30-50 (CHM declares the inputs with SpatialType::neighbor; the library exchanges them itself)
    depends("snowdepthavg");
    depends("swe");

    use_vertical_snow = cfg.get("use_vertical_snow", true);

    provides("delta_avalanche_mass");
    provides("delta_avalanche_snowdepth");

    provides("delta_avalanche_mass_sum");
    provides("delta_avalanche_snowdepth_sum");

    provides("maxDepth");
    // the ghost_ss_* scratch variables of the reference carry its MPI exchange (snow_slide.cpp:41-48); that exchange lives in the
    // library, so they are not declared here
}

snow_slide_gpu::~snow_slide_gpu() { pbsm3d_host_free(_stage); }

void snow_slide_gpu::init(mesh& domain)
{
    _h = PBSM3D_gpu::shared_handle();
    if (!_h)
        CHM_THROW_EXCEPTION(module_error, "snow_slide_gpu: no PBSM3D_gpu handle (PBSM3D_gpu must be initialised first: it owns the device mesh)");
    pbsm3d_slide_config sc;
    pbsm3d_slide_config_defaults(&sc);
    sc.avalache_mult = cfg.get("avalache_mult", 3178.4); // snow_slide.cpp:409-410
    sc.avalache_pow = cfg.get("avalache_pow", -1.998);
    sc.use_vertical_snow = use_vertical_snow ? 1 : 0;
    check(pbsm3d_slide_init(_h, &sc));

    const size_t ntri = _ntri = domain->size_faces();
    _stage = (double*)pbsm3d_host_alloc(8 * ntri * sizeof(double));
    if (!_stage)
        CHM_THROW_EXCEPTION(module_error, std::string("snow_slide_gpu: ") + pbsm3d_last_error());
    std::memset(_stage, 0, 8 * ntri * sizeof(double));
    double** slots[8] = {&_sd, &_sdv, &_swe, &_dsd, &_dmass, &_sum_sd, &_sum_mass, &_maxd};
    for (int k = 0; k < 8; ++k)
        *slots[k] = _stage + (size_t)k * ntri;

    check(pbsm3d_slide_get_constants(_h, _maxd, nullptr));
    for (size_t i = 0; i < ntri; i++)
    { // snow_slide.cpp:441-443
        auto face = domain->face(i);
        (*face)["maxDepth"_s] = _maxd[i];
        (*face)["delta_avalanche_snowdepth_sum"_s] = 0;
        (*face)["delta_avalanche_mass_sum"_s] = 0;
    }
}

void snow_slide_gpu::run(mesh& domain)
{
    const size_t ntri = _ntri;
#pragma omp parallel for
    for (size_t i = 0; i < ntri; i++)
    { // snow_slide.cpp:120-122: snowdepthavg_vert is read without being declared
        auto face = domain->face(i);
        _sd[i] = (*face)["snowdepthavg"_s];
        _sdv[i] = (*face)["snowdepthavg_vert"_s];
        _swe[i] = (*face)["swe"_s];
    }
    check(pbsm3d_slide_run(_h, _sd, _sdv, _swe, _dsd, _dmass, _sum_sd, _sum_mass, nullptr, &_stats, 0));
    SPDLOG_DEBUG("[SnowSlide] needed {} iterations", _stats.iterations);
#pragma omp parallel for
    for (size_t i = 0; i < ntri; i++)
    { // snow_slide.cpp:352-357
        auto face = domain->face(i);
        (*face)["delta_avalanche_snowdepth"_s] = _dsd[i];
        (*face)["delta_avalanche_mass"_s] = _dmass[i];
        (*face)["delta_avalanche_snowdepth_sum"_s] = _sum_sd[i];
        (*face)["delta_avalanche_mass_sum"_s] = _sum_mass[i];
    }
}

void snow_slide_gpu::checkpoint(mesh& domain, netcdf& chkpt)
{ // snow_slide.cpp:59-76
    const size_t n = domain->size_faces();
    std::vector<double> a(n), b(n), c(n), d(n);
    check(pbsm3d_slide_get_state(_h, a.data(), b.data(), c.data(), d.data()));
    chkpt.create_variable1D("snow_slide:delta_avalanche_snowdepth", n);
    chkpt.create_variable1D("snow_slide:delta_avalanche_mass", n);
    chkpt.create_variable1D("snow_slide:delta_avalanche_snowdepth_sum", n);
    chkpt.create_variable1D("snow_slide:delta_avalanche_mass_sum", n);
    for (size_t i = 0; i < n; i++)
    {
        chkpt.put_var1D("snow_slide:delta_avalanche_snowdepth", i, a[i]);
        chkpt.put_var1D("snow_slide:delta_avalanche_mass", i, b[i]);
        chkpt.put_var1D("snow_slide:delta_avalanche_snowdepth_sum", i, c[i]);
        chkpt.put_var1D("snow_slide:delta_avalanche_mass_sum", i, d[i]);
    }
}

void snow_slide_gpu::load_checkpoint(mesh& domain, netcdf& chkpt)
{ // snow_slide.cpp:78-92
    const size_t n = domain->size_faces();
    std::vector<double> a(n), b(n), c(n), d(n);
    for (size_t i = 0; i < n; i++)
    {
        a[i] = chkpt.get_var1D("snow_slide:delta_avalanche_snowdepth", i);
        b[i] = chkpt.get_var1D("snow_slide:delta_avalanche_mass", i);
        c[i] = chkpt.get_var1D("snow_slide:delta_avalanche_snowdepth_sum", i);
        d[i] = chkpt.get_var1D("snow_slide:delta_avalanche_mass_sum", i);
        (*domain->face(i))["delta_avalanche_snowdepth_sum"_s] = c[i];
        (*domain->face(i))["delta_avalanche_mass_sum"_s] = d[i];
    }
    check(pbsm3d_slide_set_state(_h, a.data(), b.data(), c.data(), d.data()));
}
