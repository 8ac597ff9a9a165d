// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libstructured_ref.so.
//
// Reference-as-library: a unity translation unit that #includes the UNMODIFIED reference
// sources where they lie under /root/reference/src (never copied into this repo; every
// reference .cpp carries an include guard, e.g. src/model/flux.cpp:1-2) and exposes the
// hot path through a small C ABI so tests/ and bench.py can use the reference's own CPU
// code as the checker / CPU baseline:
//   EulerEquation<double,double>::calc_residual   src/model/eulerequation.cpp:202-232
//   EulerEquation<double,double>::calc_dt         src/model/eulerequation.cpp:237-259
//   ADOL-C sparse_jac stand-in (see oracle/jacdriver.hpp) applied to the SAME templates
//   instantiated with Tad = oad::Dual<N> / oad::DepSet  (src/solver/solver.cpp:72-90,156)
// Built only where /root/reference exists (oracle/ref/Makefile); the .so travels to the GPU box.
#include <limits.h>
#include <climits>
#undef CHAR_WIDTH   // vendored fmt uses CHAR_WIDTH as an identifier (spdlog/fmt/bundled/format.h:2198)

#include "../adtypes.hpp"
#include "../jacdriver.hpp"

#ifndef REF_DUAL_LANES
#define REF_DUAL_LANES 8
#endif
typedef oad::Dual<REF_DUAL_LANES> RDual;
typedef oad::DepSet RDep;
using oad::sqrt; using oad::fabs; using oad::pow; using oad::abs;

// io.cpp:224-234 calls value(Tad): the generic template there returns Tad; these overloads
// must be visible before io.cpp is parsed.
inline double value(const RDual x) { return x.v; }
inline double value(const RDep x) { return x.v; }

#include "common.h"
#include "utils/config.cpp"
#include "model/fluid.cpp"
#include "model/reconstruction.cpp"
#include "model/flux.cpp"
#include "utils/mesh.cpp"
#include "solver/solution.cpp"
#include "model/bc.cpp"
#include "utils/io.cpp"
#include "model/eulerequation.cpp"
#include "thirdparty/cnpy.cpp"

// Mesh<Tx,Tad>::~Mesh is declared (src/utils/mesh.h) and defined at the end of mesh.cpp for the
// stock instantiations; the unity TU sees the template definition, so other Tad instantiate implicitly.

#include <unistd.h>
#include <string>
#include <memory>
#include <omp.h>

namespace {

template <class Tad>
struct Side {
    std::shared_ptr<Mesh<double, Tad>> mesh;
    void build(std::shared_ptr<Config<double>> cfg) {
        mesh = std::make_shared<Mesh<double, Tad>>(cfg);
        mesh->label = "";
        mesh->setup();
    }
};

struct RefHandle {
    std::shared_ptr<Config<double>> config;
    Side<double> d;
    Side<RDual> dual;
    Side<RDep> dep;
    std::string dir;
    size_t nic, njc, nv;
};

struct CwdGuard {
    char old[4096];
    bool ok;
    explicit CwdGuard(const std::string& dir) { ok = getcwd(old, sizeof(old)) != nullptr; if (!dir.empty()) { int r = chdir(dir.c_str()); (void)r; } }
    ~CwdGuard() { if (ok) { int r = chdir(old); (void)r; } }
};

void ensure_logger() {
    if (!spdlog::get("console")) {
        auto logger = spdlog::stdout_logger_mt("console", true);
        logger->set_level(spdlog::level::off);
    }
}

template <class Tad>
void eval_residual(Side<Tad>& s, const Tad* q, Tad* rhs, bool lhs) {
    auto sol = s.mesh->solution;
    const size_t n = sol->nt;
    Tad* aq = sol->a_q.data();
    for (size_t c = 0; c < n; c++) aq[c] = q[c];
    s.mesh->equation->calc_residual(sol->a_q.const_ref(), sol->a_rhs, lhs);
    const Tad* ar = sol->a_rhs.data();
    for (size_t c = 0; c < n; c++) rhs[c] = ar[c];
}

} // namespace

extern "C" {

// config_path: a reference .inp (TOML) file; its geometry.filename is resolved relative to the
// directory of the .inp (the reference resolves it relative to the cwd, src/utils/mesh.cpp:357-362).
void* ref_create(const char* config_path) {
    ensure_logger();
    std::string path(config_path);
    std::string dir, base = path;
    size_t pos = path.find_last_of('/');
    if (pos != std::string::npos) { dir = path.substr(0, pos); base = path.substr(pos + 1); }
    CwdGuard g(dir);
    auto h = new RefHandle();
    h->dir = dir;
    char* argv0[] = {(char*)"ref", nullptr};
    // "convergence"/history loggers are created by Solver only; Config just prints through "console".
    h->config = std::make_shared<Config<double>>(base, 1, argv0);
    h->d.build(h->config);
    h->nic = h->d.mesh->nic; h->njc = h->d.mesh->njc;
    h->nv = h->d.mesh->solution->nq + h->d.mesh->solution->ntrans;
    return h;
}

void ref_destroy(void* hv) { delete (RefHandle*)hv; }

void ref_dims(void* hv, int* ni, int* nj, int* nv) {
    auto h = (RefHandle*)hv;
    *ni = (int)h->d.mesh->ni; *nj = (int)h->d.mesh->nj; *nv = (int)h->nv;
}

// xv, yv: [ni][nj] row-major, exactly Mesh::xv.data()
void ref_get_grid(void* hv, double* xv, double* yv) {
    auto h = (RefHandle*)hv;
    auto m = h->d.mesh;
    std::memcpy(xv, m->xv.data(), sizeof(double)*m->ni*m->nj);
    std::memcpy(yv, m->yv.data(), sizeof(double)*m->ni*m->nj);
}

// metrics as Mesh::calc_metrics leaves them (src/utils/mesh.cpp:172-205)
void ref_get_metrics(void* hv, double* normal_chi, double* normal_eta, double* volume) {
    auto h = (RefHandle*)hv;
    auto m = h->d.mesh;
    std::memcpy(normal_chi, m->normal_chi.data(), sizeof(double)*m->ni*m->njc*2);
    std::memcpy(normal_eta, m->normal_eta.data(), sizeof(double)*m->nic*m->nj*2);
    std::memcpy(volume, m->volume.data(), sizeof(double)*m->nic*m->njc);
}

// initial state after EulerEquation::initialize (src/model/eulerequation.cpp:262-291)
void ref_get_q(void* hv, double* q) {
    auto h = (RefHandle*)hv;
    std::memcpy(q, h->d.mesh->solution->q.data(), sizeof(double)*h->d.mesh->solution->nt);
}

// rhs = calc_residual(q, lhs) ; q, rhs: [nic][njc][nv]
void ref_residual(void* hv, const double* q, double* rhs, int lhs) {
    auto h = (RefHandle*)hv;
    auto sol = h->d.mesh->solution;
    std::memcpy(sol->q.data(), q, sizeof(double)*sol->nt);
    h->d.mesh->equation->calc_residual(sol->q.const_ref(), sol->rhs, lhs != 0);
    std::memcpy(rhs, sol->rhs.data(), sizeof(double)*sol->nt);
}

// padded primitive arrays after calc_intermediates of the last ref_residual: [nic+2][njc+2] each
void ref_get_primitives(void* hv, double* rho, double* u, double* v, double* p, double* T) {
    auto h = (RefHandle*)hv;
    auto e = h->d.mesh->equation;
    const size_t n = (h->nic + 2)*(h->njc + 2);
    std::memcpy(rho, e->rho.data(), sizeof(double)*n);
    std::memcpy(u, e->u.data(), sizeof(double)*n);
    std::memcpy(v, e->v.data(), sizeof(double)*n);
    std::memcpy(p, e->p.data(), sizeof(double)*n);
    std::memcpy(T, e->T.data(), sizeof(double)*n);
}

// IOManager::write_surface (src/utils/io.cpp:182-255) run by the reference itself: calc_residual at q_res (fills
// EulerEquation::grad_{u,v}_eta), Solution::q = q_fin, then IOManager::write's own first statement (primvars into
// Solution::p, io.cpp:41) and write_surface(), which writes "<label>.surface" (text, 6 significant digits: xw cp cf per
// line) into the config's directory.  wall[6][nic] receives, in full precision, the reference's own arrays the routine
// reads: grad_u_eta[i][0][0..1], grad_v_eta[i][0][0..1], Solution::p[i][0], p[i][1].  Returns the label length written
// into label_out (the caller reads the file).
int ref_surface(void* hv, const double* q_res, const double* q_fin, double* wall, char* label_out, int label_cap) {
    auto h = (RefHandle*)hv;
    auto m = h->d.mesh;
    auto sol = m->solution;
    std::memcpy(sol->q.data(), q_res, sizeof(double)*sol->nt);
    m->equation->calc_residual(sol->q.const_ref(), sol->rhs, false);
    std::memcpy(sol->q.data(), q_fin, sizeof(double)*sol->nt);
    m->fluid_model->primvars(sol->q.const_ref(), sol->rho, sol->u, sol->v, sol->p, sol->T);
    {
        CwdGuard g(h->dir);
        m->iomanager->write_surface();
    }
    const size_t nic = h->nic;
    auto e = m->equation;
    if (wall) for (size_t i = 0; i < nic; i++) {
        wall[i] = e->grad_u_eta[i][0][0]; wall[nic + i] = e->grad_u_eta[i][0][1];
        wall[2*nic + i] = e->grad_v_eta[i][0][0]; wall[3*nic + i] = e->grad_v_eta[i][0][1];
        wall[4*nic + i] = sol->p[i][0]; wall[5*nic + i] = sol->p[i][1];
    }
    const std::string& lab = m->iomanager->label;
    if (label_out && label_cap > 0) { std::strncpy(label_out, lab.c_str(), label_cap - 1); label_out[label_cap - 1] = 0; }
    return (int)lab.size();
}

// dt[nic][njc][nv] = calc_dt(cfl) at state q (entries k >= nq are never written by the reference)
void ref_calc_dt(void* hv, const double* q, double cfl, double* dt) {
    auto h = (RefHandle*)hv;
    auto sol = h->d.mesh->solution;
    std::memcpy(sol->q.data(), q, sizeof(double)*sol->nt);
    for (size_t c = 0; c < sol->nt; c++) sol->dt.data()[c] = 0.0;
    h->d.mesh->equation->calc_dt(cfl);
    std::memcpy(dt, sol->dt.data(), sizeof(double)*sol->nt);
}

// Times `reps` calls of calc_residual on the reference's own arrays; returns seconds (wall).
double ref_time_residual(void* hv, const double* q, int reps, int lhs) {
    auto h = (RefHandle*)hv;
    auto sol = h->d.mesh->solution;
    std::memcpy(sol->q.data(), q, sizeof(double)*sol->nt);
    Timer t; t.reset();
    for (int r = 0; r < reps; r++)
        h->d.mesh->equation->calc_residual(sol->q.const_ref(), sol->rhs, lhs != 0);
    return (double)t.diff();
}

// Jacobian of calc_residual(q, lhs) in COO, malloc'd like sparse_jac's outputs; caller frees with ref_free.
int ref_jacobian(void* hv, const double* q, int lhs, int* nnz, unsigned int** rind, unsigned int** cind,
                 double** values, int* ncolors) {
    auto h = (RefHandle*)hv;
    {
        CwdGuard g(h->dir);
        if (!h->dual.mesh) h->dual.build(h->config);
        if (!h->dep.mesh) h->dep.build(h->config);
    }
    const size_t n = h->d.mesh->solution->nt;
    const bool l = lhs != 0;
    oad::Coo coo = oad::sparse_jacobian<REF_DUAL_LANES>(n, q,
        [&](const RDep* a, RDep* r) { eval_residual(h->dep, a, r, l); },
        [&](const RDual* a, RDual* r) { eval_residual(h->dual, a, r, l); });
    *nnz = coo.nnz; *rind = coo.rind; *cind = coo.cind; *values = coo.values;
    if (ncolors) *ncolors = coo.ncolors;
    return 0;
}

double ref_time_jacobian(void* hv, const double* q, int lhs, int* nnz_out) {
    int nnz; unsigned int *r, *c; double* v; int nc;
    Timer t; t.reset();
    ref_jacobian(hv, q, lhs, &nnz, &r, &c, &v, &nc);
    double s = (double)t.diff();
    if (nnz_out) *nnz_out = nnz;
    free(r); free(c); free(v);
    return s;
}

void ref_free(void* p) { free(p); }

} // extern "C"
