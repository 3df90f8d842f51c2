// Harness around the UNMODIFIED reference sources (compiled where they lie under /root/reference by oracle/refbuild/Makefile):
//     src/modules/PBSM3D.cpp   src/math/coordinates.cpp   src/physics/Atmosphere.cpp
//     src/modules/scale_wind_vert.cpp   src/modules/fetchr.cpp        (the two per-face providers of PBSM3D inputs)
//     src/modules/snow_slide.cpp                                       (gravitational redistribution)
// against the stand-in headers in stubs/.  This file holds (1) the stand-in NearestNeighborProblem implementation
// (pattern + numbering of LinearAlgebra.cpp:31-152; Solve() = a registered sparse direct solve) and (2) a small C API
// that builds a CHM-like mesh from flat arrays, runs PBSM3D::init/run and reads variables / assembled systems back.
// Test infrastructure only: used by tests/ and tests/golden/make_ref_golden.py to pin oracle/pbsm3d_oracle.py.
#include "PBSM3D.hpp"
#include "fetchr.hpp"
#include "scale_wind_vert.hpp"
#include "snow_slide.hpp"

#include <algorithm>
#include <cstring>
#include <sstream>

namespace {
std::vector<math::LinearAlgebra::NearestNeighborProblem*> g_nnp_registry;  // construction order: suspension, deposition
typedef int (*solve_cb_t)(int n, const int* rowptr, const int* col, const double* val, const double* rhs, double* x);
solve_cb_t g_solve_cb = nullptr;
std::string g_last_error;
}

namespace math { namespace LinearAlgebra {

NearestNeighborProblem::NearestNeighborProblem(mesh& domain, int nLayer) : m_domain(domain), m_nLayer(nLayer)
{
    m_ntri = domain->size_faces();
    m_nglobal = domain->size_global_faces();
    const std::size_t n = m_ntri * nLayer;
    std::vector<std::vector<global_ordinal_type>> cols(n);
    for (std::size_t i = 0; i < m_ntri; ++i) {
        auto face = domain->face(i);
        const global_ordinal_type g = face->cell_global_id;
        const std::size_t l = face->cell_local_id;
        for (int layer = 0; layer < nLayer; ++layer) {
            auto& c = cols[m_ntri * layer + l];
            c.push_back(m_nglobal * layer + g);
            for (int f = 0; f < 3; ++f)
                if (face->neighbor(f) != nullptr) c.push_back(m_nglobal * layer + face->neighbor(f)->cell_global_id);
        }
        for (int layer = 1; layer < nLayer; ++layer) cols[m_ntri * layer + l].push_back(m_nglobal * (layer - 1) + g);
        for (int layer = 0; layer < nLayer - 1; ++layer) cols[m_ntri * layer + l].push_back(m_nglobal * (layer + 1) + g);
    }
    rowptr.assign(n + 1, 0);
    for (std::size_t r = 0; r < n; ++r) rowptr[r + 1] = rowptr[r] + (int)cols[r].size();
    colgid.reserve(rowptr[n]);
    for (auto& c : cols) colgid.insert(colgid.end(), c.begin(), c.end());
    values.assign(colgid.size(), 0.0);
    rhs.assign(n, 0.0);
    solution.assign(n, 0.0);
    g_nnp_registry.push_back(this);
}

NearestNeighborProblem::~NearestNeighborProblem()
{
    for (auto& p : g_nnp_registry)
        if (p == this) p = nullptr;
}

void NearestNeighborProblem::zeroSystem()
{
    std::fill(values.begin(), values.end(), 0.0);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    std::fill(solution.begin(), solution.end(), 0.0);
}

// Single rank in the harness: owned global ids are 0..ntri-1 == local ids, so global row layer*G+g is local row layer*ntri+g.
std::size_t NearestNeighborProblem::local_row(global_ordinal_type g) const
{
    const std::size_t layer = g / m_nglobal, id = g % m_nglobal;
    if (id >= m_ntri || layer >= (std::size_t)m_nLayer) {
        std::fprintf(stderr, "refharness: row %lld is not owned\n", g);
        std::abort();
    }
    return layer * m_ntri + id;
}

double* NearestNeighborProblem::entry(global_ordinal_type row, global_ordinal_type col)
{
    const std::size_t r = local_row(row);
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k)
        if (colgid[k] == col) return &values[k];
    std::fprintf(stderr, "refharness: (%lld,%lld) is outside the static sparsity pattern\n", row, col);
    std::abort();
}

void NearestNeighborProblem::matrixReplaceGlobalValues(global_ordinal_type row, global_ordinal_type col, double val) { *entry(row, col) = val; }
void NearestNeighborProblem::matrixSumIntoGlobalValues(global_ordinal_type row, global_ordinal_type col, double val) { *entry(row, col) += val; }
void NearestNeighborProblem::rhsSumIntoGlobalValue(global_ordinal_type idx, double val) { rhs[local_row(idx)] += val; }

double NearestNeighborProblem::getRhsMax()
{
    double m = 0;
    for (double v : rhs) m = std::max(m, std::fabs(v));
    return m;
}
double NearestNeighborProblem::getSolutionMax()
{
    double m = 0;
    for (double v : solution) m = std::max(m, std::fabs(v));
    return m;
}

SolveConverge NearestNeighborProblem::Solve()
{
    if (!g_solve_cb) throw Belos::StatusTestError("refharness: no solver registered (chmref_set_solver)");
    std::vector<int> col(colgid.size());
    for (std::size_t k = 0; k < colgid.size(); ++k) col[k] = (int)local_row(colgid[k]);
    const int rc = g_solve_cb((int)rhs.size(), rowptr.data(), col.data(), values.data(), rhs.data(), solution.data());
    if (rc != 0) CHM_THROW_EXCEPTION(module_error, "Belos solver failed to converge");
    SolveConverge r;
    r.numIters = 1;
    r.residual = 0;
    return r;
}
}}

namespace {
struct Harness {
    mesh domain;
    std::shared_ptr<global> glob;
    std::unique_ptr<PBSM3D> mod;
    math::LinearAlgebra::NearestNeighborProblem* nnp[2] = {nullptr, nullptr};
    netcdf chk;
    std::unique_ptr<snow_slide> slide;
};

void parse_kv(const char* text, ptree_stub& out)
{
    if (!text) return;
    std::istringstream ss(text);
    std::string line;
    while (std::getline(ss, line)) {
        auto p = line.find('=');
        if (p == std::string::npos) continue;
        out.kv[line.substr(0, p)] = line.substr(p + 1);
    }
}
}

extern "C" {

const char* chmref_last_error() { return g_last_error.c_str(); }
void chmref_set_solver(solve_cb_t cb) { g_solve_cb = cb; }

// vx,vy,vz: [T][3] vertex coordinates of each face in the mesh file's vertex order; neigh: [T][3], -1 = none
// (neighbour i is opposite vertex i, as in CHM's .mesh files); params: n_params arrays of [T]; NaN = parameter absent
// on that face.  cfg_kv: PBSM3D config, one key=value per line.  global_kv: global parameters (landcover table) the same way.
void* chmref_create(int T, const double* vx, const double* vy, const double* vz, const int* neigh, int n_params,
                    const char* const* param_names, const double* param_values, const char* cfg_kv, const char* global_kv)
{
    try {
        math::gis::distance = math::gis::distance_UTM;  // what core.cpp installs for a projected (UTM) mesh
        math::gis::point_from_bearing = math::gis::point_from_bearing_UTM;
        auto h = new Harness;
        h->glob = std::make_shared<global>();
        parse_kv(global_kv, h->glob->parameters);
        h->domain = std::make_shared<triangulation>();
        h->domain->_global = h->glob;
        h->domain->_n_global = T;
        h->domain->_faces.resize(T);
        for (int i = 0; i < T; ++i) {
            auto f = new face_stub;
            f->_domain = h->domain.get();
            f->cell_global_id = f->cell_local_id = i;
            for (int k = 0; k < 3; ++k) {
                f->vx[k] = vx[3 * i + k];
                f->vy[k] = vy[3 * i + k];
                f->vz[k] = vz[3 * i + k];
            }
            for (int p = 0; p < n_params; ++p) {
                const double v = param_values[(std::size_t)p * T + i];
                if (!std::isnan(v)) f->_parameters[param_names[p]] = v;
            }
            h->domain->_faces[i].reset(f);
        }
        for (int i = 0; i < T; ++i)
            for (int k = 0; k < 3; ++k) {
                const int n = neigh[3 * i + k];
                h->domain->_faces[i]->_neigh[k] = n < 0 ? nullptr : h->domain->_faces[n].get();
            }
        config_file cfg;
        parse_kv(cfg_kv, cfg);
        const std::size_t before = g_nnp_registry.size();
        h->mod.reset(new PBSM3D(cfg));
        h->mod->global_param = h->glob;
        h->mod->init(h->domain);
        if (g_nnp_registry.size() != before + 2) throw std::runtime_error("refharness: expected two NearestNeighborProblems from init()");
        h->nnp[0] = g_nnp_registry[before];
        h->nnp[1] = g_nnp_registry[before + 1];
        return h;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return nullptr;
    }
}

void chmref_destroy(void* hv) { delete static_cast<Harness*>(hv); }

int chmref_set_var(void* hv, const char* name, const double* vals)
{
    auto h = static_cast<Harness*>(hv);
    for (std::size_t i = 0; i < h->domain->size_faces(); ++i) (*h->domain->face(i))[name] = vals[i];
    return 0;
}

int chmref_get_var(void* hv, const char* name, double* out)
{
    auto h = static_cast<Harness*>(hv);
    for (std::size_t i = 0; i < h->domain->size_faces(); ++i) out[i] = (*h->domain->face(i))[name];
    return 0;
}

int chmref_run(void* hv, double dt)
{
    auto h = static_cast<Harness*>(hv);
    try {
        h->glob->_dt = dt;
        h->mod->run(h->domain);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    }
}

// which: 0 = suspension system, 1 = deposition system (state after the last run()).
int chmref_system_size(void* hv, int which, int* n_rows, int* nnz)
{
    auto p = static_cast<Harness*>(hv)->nnp[which];
    *n_rows = (int)p->rhs.size();
    *nnz = (int)p->values.size();
    return 0;
}

int chmref_system(void* hv, int which, int* rowptr, long long* col, double* val, double* rhs, double* sol)
{
    auto p = static_cast<Harness*>(hv)->nnp[which];
    std::memcpy(rowptr, p->rowptr.data(), p->rowptr.size() * sizeof(int));
    std::memcpy(col, p->colgid.data(), p->colgid.size() * sizeof(long long));
    std::memcpy(val, p->values.data(), p->values.size() * sizeof(double));
    std::memcpy(rhs, p->rhs.data(), p->rhs.size() * sizeof(double));
    std::memcpy(sol, p->solution.data(), p->solution.size() * sizeof(double));
    return 0;
}

int chmref_n_depends(void* hv) { return (int)static_cast<Harness*>(hv)->mod->_depends.size(); }
const char* chmref_depend(void* hv, int i) { return static_cast<Harness*>(hv)->mod->_depends.at(i).c_str(); }
int chmref_n_provides(void* hv) { return (int)static_cast<Harness*>(hv)->mod->_provides.size(); }
const char* chmref_provide(void* hv, int i) { return static_cast<Harness*>(hv)->mod->_provides.at(i).c_str(); }

// checkpoint round trip through the reference's own checkpoint()/load_checkpoint() (PBSM3D.cpp:1753-1773)
int chmref_checkpoint(void* hv, double* sum_drift)
{
    auto h = static_cast<Harness*>(hv);
    h->mod->checkpoint(h->domain, h->chk);
    auto& v = h->chk.vars.at("PBSM3D:sum_drift");
    std::memcpy(sum_drift, v.data(), v.size() * sizeof(double));
    return 0;
}
int chmref_load_checkpoint(void* hv, const double* sum_drift)
{
    auto h = static_cast<Harness*>(hv);
    h->chk.create_variable1D("PBSM3D:sum_drift", h->domain->size_faces());
    for (std::size_t i = 0; i < h->domain->size_faces(); ++i) h->chk.put_var1D("PBSM3D:sum_drift", i, sum_drift[i]);
    h->mod->load_checkpoint(h->domain, h->chk);
    return 0;
}

// scale_wind_vert in domain mode (ctor -> init(mesh) -> run(mesh), scale_wind_vert.cpp:27-229) on the harness mesh; reads U_R
// (+ snowdepthavg when has_snowdepth, i.e. when another module provides the optional input), writes U_2m_above_srf.
// point_only: run(face) per face instead, i.e. point_scale without the neighbour spline (data-parallel mode).
int chmref_run_scale_wind_vert(void* hv, const char* cfg_kv, int has_snowdepth, int point_only)
{
    auto h = static_cast<Harness*>(hv);
    try {
        config_file cfg;
        parse_kv(cfg_kv, cfg);
        scale_wind_vert m(cfg);
        m.global_param = h->glob;
        m.ID = 7;  // module-data slot distinct from PBSM3D's
        if (has_snowdepth) m._optional_found["snowdepthavg"] = true;
        m.init(h->domain);
        if (point_only) {
            for (std::size_t i = 0; i < h->domain->size_faces(); ++i) { auto f = h->domain->face(i); m.run(f); }
        } else {
            m.run(h->domain);
        }
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    }
}

// fetchr (parallel::data: core calls run(face) for every face, fetchr.cpp:54-119); reads vw_dir, writes fetch.
int chmref_run_fetchr(void* hv, const char* cfg_kv)
{
    auto h = static_cast<Harness*>(hv);
    try {
        config_file cfg;
        parse_kv(cfg_kv, cfg);
        fetchr m(cfg);
        m.global_param = h->glob;
        for (std::size_t i = 0; i < h->domain->size_faces(); ++i) { auto f = h->domain->face(i); m.run(f); }
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    }
}

// ---- snow_slide (src/modules/snow_slide.cpp, compiled unmodified; USE_MPI undefined, so its exchanges are compiled out) ----
// Ghost faces for a rank-local view (is_ghost = true; not part of domain->face(i)): vx,vy,vz [nG][3]; area [nG] (NaN = compute from
// the vertices); ghost attach_ghost[k] becomes neighbour attach_edge[k] of owned face attach_face[k] (a ghost may touch two faces).
int chmref_add_ghosts(void* hv, int nG, const double* vx, const double* vy, const double* vz, const double* area, int n_attach,
                      const int* attach_face, const int* attach_edge, const int* attach_ghost)
{
    auto h = static_cast<Harness*>(hv);
    for (int k = 0; k < nG; ++k) {
        auto f = new face_stub;
        f->_domain = h->domain.get();
        f->is_ghost = f->_is_ghost = true;
        for (int j = 0; j < 3; ++j) { f->vx[j] = vx[3 * k + j]; f->vy[j] = vy[3 * k + j]; f->vz[j] = vz[3 * k + j]; }
        if (!std::isnan(area[k])) f->_parameters["area"] = area[k];
        h->domain->_ghosts.emplace_back(f);
    }
    for (int k = 0; k < n_attach; ++k) h->domain->_faces[attach_face[k]]->_neigh[attach_edge[k]] = h->domain->_ghosts[attach_ghost[k]].get();
    return 0;
}
int chmref_set_ghost_var(void* hv, const char* name, const double* vals)
{
    auto h = static_cast<Harness*>(hv);
    for (std::size_t i = 0; i < h->domain->_ghosts.size(); ++i) (*h->domain->_ghosts[i])[name] = vals[i];
    return 0;
}
int chmref_get_ghost_var(void* hv, const char* name, double* out)
{
    auto h = static_cast<Harness*>(hv);
    for (std::size_t i = 0; i < h->domain->_ghosts.size(); ++i) out[i] = (*h->domain->_ghosts[i])[name];
    return 0;
}
int chmref_slide_init(void* hv, const char* cfg_kv)
{
    auto h = static_cast<Harness*>(hv);
    try {
        config_file cfg;
        parse_kv(cfg_kv, cfg);
        h->slide.reset(new snow_slide(cfg));
        h->slide->global_param = h->glob;
        h->slide->ID = 9;  // module-data slot distinct from PBSM3D's and scale_wind_vert's
        h->slide->init(h->domain);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    }
}
int chmref_slide_run(void* hv)
{
    auto h = static_cast<Harness*>(hv);
    try {
        h->slide->run(h->domain);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    }
}
int chmref_slide_n_depends(void* hv) { return (int)static_cast<Harness*>(hv)->slide->_depends.size(); }
const char* chmref_slide_depend(void* hv, int i) { return static_cast<Harness*>(hv)->slide->_depends.at(i).c_str(); }
int chmref_slide_n_provides(void* hv) { return (int)static_cast<Harness*>(hv)->slide->_provides.size(); }
const char* chmref_slide_provide(void* hv, int i) { return static_cast<Harness*>(hv)->slide->_provides.at(i).c_str(); }
// checkpoint round trip through snow_slide::checkpoint / load_checkpoint (snow_slide.cpp:59-93): 4 arrays [T]
int chmref_slide_checkpoint(void* hv, double* out4)
{
    auto h = static_cast<Harness*>(hv);
    h->slide->checkpoint(h->domain, h->chk);
    const char* names[4] = {"snow_slide:delta_avalanche_snowdepth", "snow_slide:delta_avalanche_mass", "snow_slide:delta_avalanche_snowdepth_sum",
                            "snow_slide:delta_avalanche_mass_sum"};
    const std::size_t T = h->domain->size_faces();
    for (int k = 0; k < 4; ++k) std::memcpy(out4 + k * T, h->chk.vars.at(names[k]).data(), T * sizeof(double));
    return 0;
}
double chmref_face_slope(void* hv, int i) { return static_cast<Harness*>(hv)->domain->face(i)->slope(); }

// The interpolant scale_wind_vert uses (stubs/interpolation.hpp, a restatement of TPSpline.cpp): n samples (x,y,v), one query.
double chmref_tpspline(int n, const double* xyv, const double* query)
{
    std::vector<boost::tuple<double, double, double>> s;
    for (int i = 0; i < n; ++i) s.push_back(boost::make_tuple(xyv[3 * i], xyv[3 * i + 1], xyv[3 * i + 2]));
    auto q = boost::make_tuple(query[0], query[1], 0.0);
    interpolation it(interp_alg::tpspline, n);
    return it(s, q);
}

// The reference's scalar helpers, straight from the compiled reference objects (Atmosphere.cpp, coordinates.cpp).
double chmref_log_scale_wind(double u, double Z_in, double Z_out, double sd, double z0) { return Atmosphere::log_scale_wind(u, Z_in, Z_out, sd, z0); }
double chmref_saturatedVapourPressure(double T) { return Atmosphere::saturatedVapourPressure(T); }
void chmref_bearing_to_cartesian(double bearing, double* xy)
{
    auto v = math::gis::bearing_to_cartesian(bearing);
    xy[0] = v.x();
    xy[1] = v.y();
}
double chmref_distance_UTM(const double* a, const double* b) { return math::gis::distance_UTM(Point_3(a[0], a[1], a[2]), Point_3(b[0], b[1], b[2])); }
}
