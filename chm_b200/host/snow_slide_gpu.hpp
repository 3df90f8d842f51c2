// snow_slide_gpu — the adaptor module for src/modules/snow_slide.{hpp,cpp} on a B200.
//
// A CHM module like PBSM3D_gpu: same base class, registration macros, constructor / init / run / checkpoint signatures, same
// depends()/provides() lists (snow_slide.cpp:27-50) and config keys (:33, :409-410); the body is gather -> pbsm3d_slide_run ->
// scatter.  The device mesh (slot order, neighbour table, partition, ghost plan) is the one PBSM3D_gpu::init already flattened:
// snow_slide_gpu takes that module's handle (PBSM3D_gpu::shared_handle()), so PBSM3D_gpu must come earlier in the module list.
// The exchanges of the reference (ghost_neighbors_communicate_variable / ghost_to_neighbors_communicate_variable /
// all_reduce, snow_slide.cpp:166-169, 332-338, 381-402) happen inside the library.
#pragma once
#include "PBSM3D_gpu.hpp"

class snow_slide_gpu : public module_base
{
    REGISTER_MODULE_HPP(snow_slide_gpu);

  public:
    snow_slide_gpu(config_file cfg);
    ~snow_slide_gpu();
    void init(mesh& domain);
    void run(mesh& domain);
    void checkpoint(mesh& domain, netcdf& chkpt);
    void load_checkpoint(mesh& domain, netcdf& chkpt);
    const pbsm3d_slide_stats& stats() const { return _stats; }

  private:
    pbsm3d_handle* _h = nullptr;  // borrowed from PBSM3D_gpu
    pbsm3d_slide_stats _stats{};
    bool use_vertical_snow = true;
    size_t _ntri = 0;
    double* _stage = nullptr;  // pinned: 3 inputs + 5 outputs
    double *_sd = nullptr, *_sdv = nullptr, *_swe = nullptr, *_dsd = nullptr, *_dmass = nullptr, *_sum_sd = nullptr, *_sum_mass = nullptr,
           *_maxd = nullptr;
};
