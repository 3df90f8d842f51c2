// Stand-alone driver of the CHM adaptor (PBSM3D_gpu) over the shim types: builds a triangulation from flat binary
// files, then drives the module exactly as CHM's core does — ctor(config) → init(mesh) → run(mesh) per step — and
// writes the face variables the module provides.  Used by tests/test_adaptor_cpp.py; no CHM code involved.
//
//   standalone_driver <dir> <nsteps>      dir holds meta.txt, vertex.bin, elem.bin, neigh.bin, [area.bin],
//                                         config.txt (key value per line), forcing_<k>_<name>.bin
//   with slide_config.txt + slide_{snowdepthavg,snowdepthavg_vert,swe}.bin present, snow_slide_gpu runs twice after the PBSM3D
//   steps (ctor -> init -> run -> run -> checkpoint) and its provided variables are written as slide_out_<run>_<name>.bin
#include <cstdio>
#include <fstream>
#include <iostream>

#include "PBSM3D_gpu.hpp"
#include "snow_slide_gpu.hpp"

template <typename T> static std::vector<T> slurp(const std::string& path, bool required = true)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f)
    {
        if (required)
            throw std::runtime_error("cannot open " + path);
        return {};
    }
    size_t n = (size_t)f.tellg();
    std::vector<T> v(n / sizeof(T));
    f.seekg(0);
    f.read((char*)v.data(), n);
    return v;
}
static void dump(const std::string& path, const std::vector<double>& v)
{
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)v.data(), v.size() * sizeof(double));
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        std::cerr << "usage: standalone_driver <dir> <nsteps>\n";
        return 2;
    }
    const std::string dir = argv[1];
    const int nsteps = std::atoi(argv[2]);
    try
    {
        auto vertex = slurp<double>(dir + "/vertex.bin");
        auto elem = slurp<int32_t>(dir + "/elem.bin");
        auto neigh = slurp<int32_t>(dir + "/neigh.bin");
        auto area = slurp<double>(dir + "/area.bin", false);
        const size_t T = elem.size() / 3, nv = vertex.size() / 3;

        auto tri = std::make_shared<triangulation>();
        tri->n_global = T;
        for (size_t v = 0; v < nv; ++v)
        {
            tri->vertices.emplace_back(new shim_vertex{{vertex[3 * v], vertex[3 * v + 1], vertex[3 * v + 2]}});
        }
        for (size_t i = 0; i < T; ++i)
        {
            tri->faces.emplace_back(new shim_face());
            auto* f = tri->faces.back().get();
            f->cell_global_id = f->cell_local_id = i;
            for (int k = 0; k < 3; ++k)
                f->vtx[k] = tri->vertices[elem[3 * i + k]].get();
            if (!area.empty())
                f->params["area"] = area[i];
        }
        for (size_t i = 0; i < T; ++i)
            for (int k = 0; k < 3; ++k)
                if (neigh[3 * i + k] >= 0)
                    tri->faces[i]->nb[k] = tri->faces[neigh[3 * i + k]].get();

        config_file cfg;
        {
            std::ifstream c(dir + "/config.txt");
            std::string k, v;
            while (c >> k >> v)
                cfg.kv[k] = v;
        }
        PBSM3D_gpu mod(cfg); // CHM: factory->create("PBSM3D_gpu", cfg)   (core.cpp:278)
        mesh domain = tri;
        mod.init(domain);     // core.cpp: module->init(_mesh)
        const char* in_names[] = {"U_R", "U_2m_above_srf", "snowdepthavg", "swe", "t", "rh", "vw_dir", "fetch"};
        const char* out_names[] = {"Qsalt", "Qsusp", "Qsubl", "Qsubl_mass", "sum_subl", "drift_mass", "sum_drift",
                                   "pbsm_more_than_avail"};
        for (int k = 0; k < nsteps; ++k)
        {
            for (const char* n : in_names)
            {
                auto a = slurp<double>(dir + "/forcing_" + std::to_string(k) + "_" + n + ".bin");
                for (size_t i = 0; i < T; ++i)
                    (*tri->face(i))[n] = a[i];
            }
            mod.run(domain); // core.cpp:2181-2184
            for (const char* n : out_names)
            {
                std::vector<double> a(T);
                for (size_t i = 0; i < T; ++i)
                    a[i] = (*tri->face(i))[n];
                dump(dir + "/out_" + std::to_string(k) + "_" + n + ".bin", a);
            }
            std::printf("step %d: suspension %d it (res %.3e), deposition %d it (res %.3e), %.3f ms\n", k,
                        mod.stats().suspension_iterations, mod.stats().suspension_residual, mod.stats().deposition_iterations,
                        mod.stats().deposition_residual, mod.stats().ms_total);
        }
        netcdf chk;
        mod.checkpoint(domain, chk);
        dump(dir + "/checkpoint_sum_drift.bin", chk.data.at("PBSM3D:sum_drift"));
        if (std::ifstream(dir + "/slide_config.txt"))
        {
            config_file scfg;
            std::ifstream c(dir + "/slide_config.txt");
            std::string k, v;
            while (c >> k >> v)
                scfg.kv[k] = v;
            snow_slide_gpu slide(scfg);
            slide.init(domain);
            const char* sin[] = {"snowdepthavg", "snowdepthavg_vert", "swe"};
            const char* sout[] = {"delta_avalanche_snowdepth", "delta_avalanche_mass", "delta_avalanche_snowdepth_sum",
                                  "delta_avalanche_mass_sum", "maxDepth"};
            for (const char* n : sin)
            {
                auto a = slurp<double>(dir + "/slide_" + n + ".bin");
                for (size_t i = 0; i < T; ++i)
                    (*tri->face(i))[n] = a[i];
            }
            for (int run = 0; run < 2; ++run)
            {
                slide.run(domain);
                for (const char* n : sout)
                {
                    std::vector<double> a(T);
                    for (size_t i = 0; i < T; ++i)
                        a[i] = (*tri->face(i))[n];
                    dump(dir + "/slide_out_" + std::to_string(run) + "_" + n + ".bin", a);
                }
                std::printf("snow_slide run %d: %d iteration(s), %d faces shed snow in %d rounds, %.3f ms\n", run, slide.stats().iterations,
                            slide.stats().faces_fired, slide.stats().wavefront_rounds, slide.stats().ms_device);
            }
            netcdf schk;
            slide.checkpoint(domain, schk);
            dump(dir + "/slide_checkpoint_mass_sum.bin", schk.data.at("snow_slide:delta_avalanche_mass_sum"));
        }
    }
    catch (std::exception& e)
    {
        std::cerr << "module_error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
