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

  private:
    pbsm3d_config _c;
    pbsm3d_handle* _h = nullptr;
    pbsm3d_stats _stats{};
    bool _use_fetch = true;
    size_t _ntri = 0;
    // SoA staging buffers (host): forcing in, outputs out.  Allocated once in init().
    std::vector<double> _U_R, _U2, _sd, _swe, _t, _rh, _vw_dir, _fetch;
    std::vector<double> _Qsalt, _Qsusp, _Qsubl, _Qsubl_mass, _sum_subl, _drift_mass, _sum_drift, _more;
};
