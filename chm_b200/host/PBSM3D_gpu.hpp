// PBSM3D_gpu — the adaptor module a CHM maintainer drops into src/modules/ to run the PBSM3D path on a B200.
//
// It IS a CHM module: same base class, same registration macros, same constructor / init / run / checkpoint
// signatures, same depends()/provides() lists and config keys as src/modules/PBSM3D.{hpp,cpp}; the body is
// gather → one call into libpbsm3d_b200.so (include/pbsm3d.h) → scatter.
//
// Two build modes, one source:
//   * inside CHM:   add this file + PBSM3D_gpu.cpp to src/modules/, link -lpbsm3d_b200; it includes CHM's own
//                   "module_base.hpp" / "triangulation.hpp" (what REGISTER_MODULE_* and `mesh` come from).
//   * stand-alone:  -DPBSM3D_GPU_STANDALONE uses chm_shim.hpp, a minimal stand-in that declares only the members
//                   of module_base / triangulation / face this adaptor touches, with the same names and meaning,
//                   so the adaptor can be compiled and tested here where Boost/CGAL/... are absent.
#pragma once

#ifdef PBSM3D_GPU_STANDALONE
#include "chm_shim.hpp"
#else
#include "logger.hpp"
#include "module_base.hpp"
#include "triangulation.hpp"
#endif

#include <memory>
#include <string>
#include <vector>

#include "pbsm3d.h"

class PBSM3D_gpu : public module_base
{
    REGISTER_MODULE_HPP(PBSM3D_gpu);

  public:
    PBSM3D_gpu(config_file cfg);
    ~PBSM3D_gpu();
    void init(mesh& domain);
    void run(mesh& domain);
    void checkpoint(mesh& domain, netcdf& chkpt);
    void load_checkpoint(mesh& domain, netcdf& chkpt);

    // last step's solver statistics (iterations, residuals, CUDA-event times)
    const pbsm3d_stats& stats() const { return _stats; }
    // the device mesh of this rank, for the modules that run on it (snow_slide_gpu); nullptr before init()
    static pbsm3d_handle* shared_handle();

  private:
    pbsm3d_config _c;
    pbsm3d_handle* _h = nullptr;
    pbsm3d_stats _stats{};
    bool _use_fetch = true;
    bool _fuse = false; // "fuse_providers": scale_wind_vert and fetchr run inside the library (pbsm3d_set_providers)
    size_t _ntri = 0;
    // SoA staging buffers (host, page-locked so the library can overlap PCIe with compute): forcing in, outputs out.
    // One block of 16 arrays, allocated once in init() with pbsm3d_host_alloc.
    double* _stage = nullptr;
    double *_U_R = nullptr, *_U2 = nullptr, *_sd = nullptr, *_swe = nullptr, *_t = nullptr, *_rh = nullptr, *_vw_dir = nullptr,
           *_fetch = nullptr, *_psh = nullptr;
    double *_Qsalt = nullptr, *_Qsusp = nullptr, *_Qsubl = nullptr, *_Qsubl_mass = nullptr, *_sum_subl = nullptr,
           *_drift_mass = nullptr, *_sum_drift = nullptr, *_more = nullptr, *_prob = nullptr;
};
