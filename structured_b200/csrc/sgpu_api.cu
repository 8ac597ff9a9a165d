// C ABI of the B200 residual + Jacobian path (include/structured_gpu.h).  Host-side glue only: every
// compute entry point launches the CUDA kernels in this directory; there is no CPU fallback.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <type_traits>

#include "../../include/structured_gpu.h"
#include "common.cuh"
#include "aux_kernels.cuh"
#include "residual_kernel.cuh"
#include "jacobian_kernel.cuh"
#include "jacobian_march.cuh"
#include "linsolve_kernels.cuh"

using namespace sg;

static thread_local std::string g_create_error;

struct LinWork;
static void lin_free(LinWork*& L);

struct sgpu_ctx {
    sgpu_desc d{};
    std::vector<sgpu_bc> bcs;
    View v{};
    Gas g{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool viscous = false;
    double eps_chi = 0, eps_eta = 0;
    // device planes
    double* q[2] = {nullptr, nullptr};
    double* q_scratch = nullptr;           // second q_tmp of the fused Runge-Kutta stages (sgpu_explicit_step), allocated on first use
    double* rhs = nullptr;
    double* dt = nullptr;
    double* xv = nullptr; double* yv = nullptr;
    double* met = nullptr;                 // 5 planes: ncx ncy nex ney vol
    double* wdist = nullptr; double* beta = nullptr;
    double* partial = nullptr; size_t partial_cap = 0;
    double* l2sq_dev = nullptr;
    double* stage = nullptr; size_t stage_cap = 0;
    double* halo_recv[2] = {nullptr, nullptr};
    double* halo_peer[2] = {nullptr, nullptr};
    bool halo_peer_ipc[2] = {false, false};
    unsigned long long halo_seq = 0;
    int* halo_err = nullptr;               // device word raised by a halo wait that timed out
    JacStore jac{};
    LinWork* lin = nullptr;              // device linear solve workspace (linsolve_api.inl)
    // pipelined host path
    bool pipe_init = false;
    cudaStream_t pipe_stream[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // transpose-in, compute, transpose-out, H2D copies, D2H copies
    cudaEvent_t pipe_up[64] = {}, pipe_cmp[64] = {}, pipe_h2d[64] = {}, pipe_tout[64] = {}, pipe_d2h[64] = {}, pipe_start = nullptr;
    double* pipe_stage[4] = {nullptr, nullptr, nullptr, nullptr}; size_t pipe_stage_cap = 0;
    double* jac_scratch = nullptr; size_t jac_scratch_cap = 0; bool jac_two_stage = false;
    double* jgeo = nullptr; bool jgeo_valid = false;   // static face-geometry weight planes of the Jacobian build (jac_geom_kernel)
    void* ghost_tab = nullptr;
    int* jac_err = nullptr;
    bool have_grid = false, have_dt = false;
    double* wall_last = nullptr; bool wall_track = false, wall_valid = false;   // wall rows of the last residual evaluation (sgpu_track_wall)
    std::string err;
    long long launches = 0;
    // kernel timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t ev_used = 0;
};

#define FAIL(ctx, code, ...) do { char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__); (ctx)->err = _b; return (code); } while (0)
#define CK(ctx, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
    char _b[512]; snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    (ctx)->err = _b; return SGPU_ERR_CUDA; } } while (0)
#define CKL(ctx) CK(ctx, cudaGetLastError())

static Metrics metrics_of(const sgpu_ctx* c) {
    Metrics m; const size_t pl = c->v.plane;
    m.ncx = c->met; m.ncy = c->met + pl; m.nex = c->met + 2*pl; m.ney = c->met + 3*pl; m.vol = c->met + 4*pl;
    return m;
}

static int ensure_stage(sgpu_ctx* c, size_t doubles) {
    if (doubles <= c->stage_cap) return SGPU_OK;
    if (c->stage) { CK(c, cudaFree(c->stage)); c->stage = nullptr; c->stage_cap = 0; }
    CK(c, cudaMalloc(&c->stage, doubles*sizeof(double)));
    c->stage_cap = doubles;
    return SGPU_OK;
}

extern "C" {

const char* sgpu_last_error(const sgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sgpu_create(const sgpu_desc* d, sgpu_ctx** out) {
    if (!d || !out) { g_create_error = "null argument"; return SGPU_ERR_ARG; }
    *out = nullptr;
    auto fail = [&](const char* msg) { g_create_error = msg; return SGPU_ERR_ARG; };
    if (d->ni < 3 || d->nj < 3) return fail("need at least 2x2 cells");
    if (d->ntrans != 0 && d->ntrans != 1) return fail("ntrans must be 0 (laminar) or 1 (SA)");
    if ((d->order != 1 && d->order != 2) || (d->lhs_order != 1 && d->lhs_order != 2)) return fail("Reconstruction not found.");   // eulerequation.cpp:108,119
    if (d->flux != SGPU_FLUX_ROE && d->flux != SGPU_FLUX_AUSM) return fail("Flux not found.");                                      // eulerequation.cpp:128
    if (d->ntrans == 1 && !(d->mu_inf > 1e-15)) return fail("the SA extension needs a viscous case (mu_inf > 0)");
    const int nic = d->ni - 1, njc = d->nj - 1;
    int j0 = d->j_begin, j1 = d->j_end;
    if (j0 == 0 && j1 == 0) j1 = njc;
    if (j0 < 0 || j1 > njc || j1 - j0 < 2) return fail("bad slab [j_begin, j_end): need at least 2 rows inside the grid");
    const bool slabbed = (j0 != 0 || j1 != njc);
    for (int n = 0; n < d->n_bc; n++) {
        const sgpu_bc& b = d->bc[n];
        const bool horiz = b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP;
        if (b.face < 0 || b.face > 3) return fail("boundary face must be bottom/right/top/left");
        if (b.type < 0 || b.type > SGPU_BC_PERIODIC) return fail("Wrong type of BC.");                                             // bc.cpp:523
        if (!horiz && (b.type == SGPU_BC_SLIPWALL || b.type == SGPU_BC_ISOTHERMALWALL || b.type == SGPU_BC_WAKE))
            return fail("Boundary condition not implemented! (slipwall/isothermalwall/wake exist on bottom/top only, bc.cpp:116-126,251-261,415-425)");
        if (b.type == SGPU_BC_OUTFLOW && b.face != SGPU_FACE_RIGHT)
            return fail("Boundary not implemented! (outflow exists on the right face only, bc.cpp:286-309)");
        if (slabbed && horiz && b.type == SGPU_BC_PERIODIC) return fail("periodic bottom/top is not supported on a j-slab partition");
        if (slabbed && b.type == SGPU_BC_WAKE && b.face == SGPU_FACE_TOP) return fail("wake on the top face is not supported on a j-slab partition");
    }
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice failed: ") + cudaGetErrorString(e); return SGPU_ERR_CUDA; }

    sgpu_ctx* c = new sgpu_ctx();
    c->d = *d; c->bcs.assign(d->bc, d->bc + d->n_bc); c->d.bc = c->bcs.data();
    c->device = d->device;
    for (auto& b : c->bcs) {                                       // BoundaryContainer::get_index, bc.cpp:436-457
        if (b.end < 0) b.end = ((b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP) ? nic : njc) + 2 + b.end;
        // a periodic table fills BOTH ghost lines of its direction whichever face it names (bc.cpp:329-365): store the
        // canonical face so that face-based filters (column / row chunks of the host pipeline) treat both spellings alike
        if (b.type == SGPU_BC_PERIODIC) b.face = (b.face == SGPU_FACE_RIGHT) ? SGPU_FACE_LEFT : (b.face == SGPU_FACE_TOP ? SGPU_FACE_BOTTOM : b.face);
    }
    View& v = c->v;
    v.nic = nic; v.njc = njc; v.ni = d->ni; v.nj = d->nj; v.j0 = j0; v.j1 = j1; v.njl = j1 - j0; v.nv = 4 + d->ntrans;
    v.pitch = ((nic + 2*IOFF + 15)/16)*16; v.rows = v.njl + 2*JOFF; v.plane = (size_t)v.rows*v.pitch;
    Gas& g = c->g;
    g.R = d->p_inf/d->rho_inf/d->T_inf;                            // fluid.cpp:12
    g.cp = GAMMA*g.R/(GAMMA - 1.0);                                // fluid.cpp:14
    g.pr = d->pr_inf; g.mu_ref = d->mu_inf; g.T_ref = d->T_inf;
    g.rho_inf = d->rho_inf; g.u_inf = d->u_inf; g.v_inf = d->v_inf; g.p_inf = d->p_inf;
    g.cp_over_pr = g.cp/g.pr; g.iR = 1.0/g.R; g.iT_ref = 1.0/g.T_ref; g.cp_over_prt = g.cp/SA_PRT;
    c->viscous = d->mu_inf > 1e-15;                                // config.cpp:41
    c->eps_chi = std::pow(10.0/nic, 3); c->eps_eta = std::pow(10.0/njc, 3);   // reconstruction.cpp:62-63 (GLOBAL counts)

    auto alloc = [&](double** p, size_t n) { return cudaMalloc(p, n*sizeof(double)); };
    const size_t pl = v.plane, vpl = (size_t)(v.rows + 1)*v.pitch;
    cudaError_t ce = cudaSuccess;
    if (ce == cudaSuccess) ce = alloc(&c->q[0], pl*v.nv);
    if (ce == cudaSuccess) ce = alloc(&c->q[1], pl*v.nv);
    if (ce == cudaSuccess) ce = alloc(&c->rhs, pl*v.nv);
    if (ce == cudaSuccess) ce = alloc(&c->dt, pl);
    if (ce == cudaSuccess) ce = alloc(&c->xv, vpl);
    if (ce == cudaSuccess) ce = alloc(&c->yv, vpl);
    if (ce == cudaSuccess) ce = alloc(&c->met, pl*5);
    if (ce == cudaSuccess) ce = alloc(&c->wdist, pl);
    if (ce == cudaSuccess) ce = alloc(&c->beta, pl);
    if (ce == cudaSuccess) ce = alloc(&c->l2sq_dev, 8);
    if (ce == cudaSuccess) ce = alloc(&c->halo_recv[0], (size_t)4*v.nv*nic + 2);
    if (ce == cudaSuccess) ce = alloc(&c->halo_recv[1], (size_t)4*v.nv*nic + 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->halo_err, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(c->halo_err, 0, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(c->halo_recv[0], 0, ((size_t)4*v.nv*nic + 2)*sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemset(c->halo_recv[1], 0, ((size_t)4*v.nv*nic + 2)*sizeof(double));
    if (ce != cudaSuccess) {
        g_create_error = std::string("cudaMalloc failed: ") + cudaGetErrorString(ce);
        sgpu_destroy(c); return SGPU_ERR_CUDA;
    }
    // benign fill (freestream) so that never-written pad cells cannot produce NaNs that leak into sums
    const double q0 = d->rho_inf, q1 = d->rho_inf*d->u_inf, q2 = d->rho_inf*d->v_inf;
    const double q3 = d->p_inf/(GAMMA - 1.0) + 0.5*d->rho_inf*(d->u_inf*d->u_inf + d->v_inf*d->v_inf);
    const double fillv[5] = {q0, q1, q2, q3, 3.0*d->mu_inf};
    for (int s = 0; s < 2; s++) for (int k = 0; k < v.nv; k++) fill_kernel<<<296, 256>>>(c->q[s] + k*pl, pl, fillv[k]);
    fill_kernel<<<296, 256>>>(c->rhs, pl*v.nv, 0.0);
    fill_kernel<<<296, 256>>>(c->dt, pl, 0.0);
    fill_kernel<<<296, 256>>>(c->wdist, pl, 1.0);
    fill_kernel<<<296, 256>>>(c->beta, pl, 1.0);
    fill_kernel<<<296, 256>>>(c->met, pl*5, 1.0);
    fill_kernel<<<296, 256>>>(c->xv, vpl, 0.0);
    fill_kernel<<<296, 256>>>(c->yv, vpl, 0.0);
    c->launches += 2*v.nv + 7;
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { g_create_error = std::string("init kernels failed: ") + cudaGetErrorString(ce); sgpu_destroy(c); return SGPU_ERR_CUDA; }
    *out = c;
    return SGPU_OK;
}

int sgpu_destroy(sgpu_ctx* c) {
    if (!c) return SGPU_OK;
    cudaSetDevice(c->device);
    for (double* p : {c->q[0], c->q[1], c->rhs, c->dt, c->xv, c->yv, c->met, c->wdist, c->beta, c->partial, c->l2sq_dev,
                      c->stage, c->halo_recv[0], c->halo_recv[1]})
        if (p) cudaFree(p);
    for (int k = 0; k < 2; k++) if (c->halo_peer_ipc[k] && c->halo_peer[k]) cudaIpcCloseMemHandle(c->halo_peer[k]);
    jac_free(c->jac);
    lin_free(c->lin);
    if (c->jac_scratch) cudaFree(c->jac_scratch);
    if (c->q_scratch) cudaFree(c->q_scratch);
    if (c->jgeo) cudaFree(c->jgeo);
    for (int k = 0; k < 4; k++) if (c->pipe_stage[k]) cudaFree(c->pipe_stage[k]);
    if (c->pipe_init) {
        for (int k = 0; k < 5; k++) cudaStreamDestroy(c->pipe_stream[k]);
        for (int k = 0; k < 64; k++) { cudaEventDestroy(c->pipe_up[k]); cudaEventDestroy(c->pipe_cmp[k]); cudaEventDestroy(c->pipe_h2d[k]); cudaEventDestroy(c->pipe_tout[k]); cudaEventDestroy(c->pipe_d2h[k]); }
        cudaEventDestroy(c->pipe_start);
    }
    if (c->ghost_tab) cudaFree(c->ghost_tab);
    if (c->wall_last) cudaFree(c->wall_last);
    if (c->jac_err) cudaFree(c->jac_err);
    if (c->halo_err) cudaFree(c->halo_err);
    for (auto& p : c->ev) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    delete c;
    return SGPU_OK;
}

int sgpu_set_stream(sgpu_ctx* c, void* s) { if (!c) return SGPU_ERR_ARG; c->stream = (cudaStream_t)s; return SGPU_OK; }
int sgpu_synchronize(sgpu_ctx* c) {
    if (!c) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device)); CK(c, cudaStreamSynchronize(c->stream));
    if (c->halo_seq) {                                             // a halo wait that gave up (neighbour never pushed)
        int e = 0; CK(c, cudaMemcpy(&e, c->halo_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) { CK(c, cudaMemset(c->halo_err, 0, sizeof(int))); FAIL(c, SGPU_ERR_STATE, "halo exchange timed out %d time(s): a neighbour slab did not push its boundary rows", e); }
    }
    return SGPU_OK;
}
int sgpu_dims(const sgpu_ctx* c, int* nic, int* njc, int* nv, int* j_begin, int* j_end) {
    if (!c) return SGPU_ERR_ARG;
    if (nic) *nic = c->v.nic; if (njc) *njc = c->v.njc; if (nv) *nv = c->v.nv; if (j_begin) *j_begin = c->v.j0; if (j_end) *j_end = c->v.j1;
    return SGPU_OK;
}
long long sgpu_launch_count(const sgpu_ctx* c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------------------------- grid
int sgpu_set_grid_window(sgpu_ctx* c, const double* xv, const double* yv, int jv_first, int jv_count) {
    if (!c || !xv || !yv) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    const int ja = std::max(v.j0 - 2, 0), jb = std::min(v.j1 + 2, v.nj - 1);       // vertex rows [ja, jb]
    if (jv_first > ja || jv_first + jv_count < jb + 1) FAIL(c, SGPU_ERR_ARG, "vertex window [%d,%d) does not cover rows [%d,%d]", jv_first, jv_first + jv_count, ja, jb);
    const int nrows = jb - ja + 1, r0 = ja - v.j0 + JOFF;
    if (int rc = ensure_stage(c, (size_t)v.ni*nrows)) return rc;
    const dim3 blk(32, 8), grd((nrows + 31)/32, (v.ni + 31)/32);
    const double* src[2] = {xv, yv}; double* dst[2] = {c->xv, c->yv};
    for (int n = 0; n < 2; n++) {
        CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*nrows, src[n] + (ja - jv_first), sizeof(double)*jv_count, sizeof(double)*nrows, v.ni,
                                cudaMemcpyHostToDevice, c->stream));
        vertex_to_plane_kernel<<<grd, blk, 0, c->stream>>>(v, c->stage, dst[n], r0, nrows);
        CKL(c); c->launches++;
    }
    Metrics m = metrics_of(c);
    metrics_kernel<<<dim3((v.pitch + 127)/128, v.rows), 128, 0, c->stream>>>(v, c->xv, c->yv, (double*)m.ncx, (double*)m.ncy,
                                                                            (double*)m.nex, (double*)m.ney, (double*)m.vol);
    CKL(c); c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    c->have_grid = true; c->jgeo_valid = false;
    return SGPU_OK;
}
int sgpu_set_grid(sgpu_ctx* c, const double* xv, const double* yv) {
    if (!c) return SGPU_ERR_ARG;
    return sgpu_set_grid_window(c, xv, yv, 0, c->v.nj);
}

// Binary vertex file (structured_b200/cases.py::write_grid_bin): "SGRIDF64", int32 ni, int32 nj, x[nj][ni], y[nj][ni]
// float64 -- the p3d order (Mesh::plot3d_loader, src/utils/mesh.cpp:146-169) without the ASCII.  A file row IS a row of
// the vertex planes, so only this slab's window of rows is read and it goes to the device with one 2-D copy per array:
// no host transpose, no staging kernel.
int sgpu_set_grid_file(sgpu_ctx* c, const char* path) {
    if (!c || !path) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    FILE* f = fopen(path, "rb");
    if (!f) FAIL(c, SGPU_ERR_ARG, "cannot open grid file %s", path);
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    char magic[8]; int hdr[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SGRIDF64", 8) != 0) FAIL(c, SGPU_ERR_ARG, "file format not found! (%s is not a binary vertex file)", path);   // mesh.cpp:364
    if (fread(hdr, sizeof(int), 2, f) != 2 || hdr[0] != v.ni || hdr[1] != v.nj) FAIL(c, SGPU_ERR_ARG, "grid file %s holds %d x %d vertices, the context %d x %d", path, hdr[0], hdr[1], v.ni, v.nj);
    const int ja = std::max(v.j0 - 2, 0), jb = std::min(v.j1 + 2, v.nj - 1);       // vertex rows [ja, jb]
    const int nrows = jb - ja + 1, r0 = ja - v.j0 + JOFF;
    std::vector<double> h((size_t)nrows*v.ni);
    double* dst[2] = {c->xv, c->yv};
    for (int n = 0; n < 2; n++) {
        const long long off = 16 + ((long long)n*v.nj + ja)*(long long)v.ni*(long long)sizeof(double);
        if (fseeko(f, (off_t)off, SEEK_SET) != 0 || fread(h.data(), sizeof(double), h.size(), f) != h.size()) FAIL(c, SGPU_ERR_ARG, "grid file %s is truncated", path);
        CK(c, cudaMemcpy2D(dst[n] + v.at(r0, IOFF), sizeof(double)*v.pitch, h.data(), sizeof(double)*v.ni, sizeof(double)*v.ni, nrows, cudaMemcpyHostToDevice));
    }
    Metrics m = metrics_of(c);
    metrics_kernel<<<dim3((v.pitch + 127)/128, v.rows), 128, 0, c->stream>>>(v, c->xv, c->yv, (double*)m.ncx, (double*)m.ncy,
                                                                            (double*)m.nex, (double*)m.ney, (double*)m.vol);
    CKL(c); c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    c->have_grid = true; c->jgeo_valid = false;
    return SGPU_OK;
}

int sgpu_set_field_window(sgpu_ctx* c, const char* name, const double* f, int j_first, int j_count) {
    if (!c || !name || !f) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    double* dst = !strcmp(name, "wall_distance") ? c->wdist : (!strcmp(name, "beta") ? c->beta : nullptr);
    if (!dst) FAIL(c, SGPU_ERR_ARG, "unknown field '%s' (wall_distance | beta)", name);
    const View& v = c->v;
    if (j_first > v.j0 || j_first + j_count < v.j1) FAIL(c, SGPU_ERR_ARG, "field window [%d,%d) does not cover the owned rows [%d,%d)", j_first, j_first + j_count, v.j0, v.j1);
    if (int rc = ensure_stage(c, (size_t)v.nic*v.njl)) return rc;
    CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*v.njl, f + (v.j0 - j_first), sizeof(double)*j_count, sizeof(double)*v.njl, v.nic,
                            cudaMemcpyHostToDevice, c->stream));
    field_to_plane_kernel<<<dim3((v.njl + 31)/32, (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, dst, JOFF, v.njl);
    CKL(c); c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    return SGPU_OK;
}
int sgpu_set_field(sgpu_ctx* c, const char* name, const double* f) {
    if (!c) return SGPU_ERR_ARG;
    return sgpu_set_field_window(c, name, f, 0, c->v.njc);
}

int sgpu_get_field(sgpu_ctx* c, const char* name, double* out) {
    if (!c || !name || !out) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const double* src = !strcmp(name, "wall_distance") ? c->wdist : (!strcmp(name, "beta") ? c->beta : nullptr);
    if (!src) FAIL(c, SGPU_ERR_ARG, "unknown field '%s' (wall_distance | beta)", name);
    const View& v = c->v;                                          // cold path: whole plane to the host, scatter there
    std::vector<double> h(v.plane);
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(h.data(), src, sizeof(double)*v.plane, cudaMemcpyDeviceToHost));
    for (int i = 0; i < v.nic; i++)
        for (int j = v.j0; j < v.j1; j++) out[(size_t)i*v.njc + j] = h[v.at(j - v.j0 + JOFF, i + IOFF)];
    return SGPU_OK;
}

// Wall distance on the device (SA extension).  segments: [nseg][4] = x0 y0 x1 y1 of the wall edges.
int sgpu_compute_wall_distance(sgpu_ctx* c, const double* segments, int nseg) {
    if (!c || (!segments && nseg > 0) || nseg < 0) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    if (nseg == 0) {                                               // no wall anywhere: the destruction term vanishes as d -> inf
        fill_kernel<<<296, 256, 0, c->stream>>>(c->wdist, v.plane, 1e30);
        CKL(c); c->launches++;
        return SGPU_OK;
    }
    std::vector<double> h((size_t)nseg*5);
    for (int k = 0; k < nseg; k++) {
        const double ax = segments[4*k], ay = segments[4*k + 1], bx = segments[4*k + 2] - ax, by = segments[4*k + 3] - ay;
        const double l2 = bx*bx + by*by;
        h[5*k] = ax; h[5*k + 1] = ay; h[5*k + 2] = bx; h[5*k + 3] = by; h[5*k + 4] = l2 > 0.0 ? 1.0/l2 : 0.0;
    }
    if (int rc = ensure_stage(c, h.size())) return rc;
    CK(c, cudaMemcpyAsync(c->stage, h.data(), sizeof(double)*h.size(), cudaMemcpyHostToDevice, c->stream));
    wall_distance_kernel<<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, c->xv, c->yv, c->stage, nseg, c->wdist);
    CKL(c); c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));                       // h goes out of scope
    return SGPU_OK;
}

// The same with the segments derived from the [[boundary]] tables: every edge of the grid boundary covered by a `wall`
// or `isothermalwall` table.  xv, yv: GLOBAL host vertex arrays [ni][nj] (Mesh::xv.data()).
int sgpu_wall_distance_from_bcs(sgpu_ctx* c, const double* xv, const double* yv) {
    if (!c || !xv || !yv) return SGPU_ERR_ARG;
    const View& v = c->v;
    std::vector<double> seg;
    auto X = [&](const double* a, int i, int j) { return a[(size_t)i*v.nj + j]; };
    for (const sgpu_bc& b : c->bcs) {
        if (b.type != SGPU_BC_WALL && b.type != SGPU_BC_ISOTHERMALWALL) continue;
        const bool horiz = b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP;
        const int n = horiz ? v.nic : v.njc;
        for (int p = std::max(b.start, 1); p <= std::min(b.end, n); p++) {     // padded index p <-> cell p - 1
            int i0, j0, i1, j1;
            if (horiz) { i0 = p - 1; i1 = p; j0 = j1 = (b.face == SGPU_FACE_BOTTOM ? 0 : v.nj - 1); }
            else { j0 = p - 1; j1 = p; i0 = i1 = (b.face == SGPU_FACE_LEFT ? 0 : v.ni - 1); }
            seg.push_back(X(xv, i0, j0)); seg.push_back(X(yv, i0, j0)); seg.push_back(X(xv, i1, j1)); seg.push_back(X(yv, i1, j1));
        }
    }
    return sgpu_compute_wall_distance(c, seg.data(), (int)(seg.size()/4));
}

// download a cell plane group to a GLOBAL host AoS array (owned rows only); window = host holds owned rows only
static int download_planes(sgpu_ctx* c, const double* planes, int nvp, double* host, bool window = false) {
    const View& v = c->v;
    const size_t M = (size_t)v.njl*v.nv;
    if (int rc = ensure_stage(c, (size_t)v.nic*M)) return rc;
    planes_to_aos_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, planes, JOFF, v.njl, nvp);
    CKL(c); c->launches++;
    if (window) CK(c, cudaMemcpyAsync(host, c->stage, sizeof(double)*M*v.nic, cudaMemcpyDeviceToHost, c->stream));
    else CK(c, cudaMemcpy2DAsync(host + (size_t)v.j0*v.nv, sizeof(double)*v.njc*v.nv, c->stage, sizeof(double)*M, sizeof(double)*M, v.nic,
                                 cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return SGPU_OK;
}

int sgpu_get_metrics(sgpu_ctx* c, double* normal_chi, double* normal_eta, double* volume) {
    if (!c) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    // small, cold path: copy the planes to the host and scatter
    const View& v = c->v;
    std::vector<double> h(v.plane*5);
    CK(c, cudaMemcpy(h.data(), c->met, sizeof(double)*v.plane*5, cudaMemcpyDeviceToHost));
    const size_t pl = v.plane;
    for (int j = v.j0; j <= v.j1; j++) {
        const int r = j - v.j0 + JOFF;
        for (int i = 0; i < v.ni; i++) {
            const size_t o = v.at(r, i + IOFF);
            if (normal_chi && j < v.j1) { normal_chi[((size_t)i*v.njc + j)*2] = h[o]; normal_chi[((size_t)i*v.njc + j)*2 + 1] = h[pl + o]; }
            if (normal_eta && i < v.nic) { normal_eta[((size_t)i*v.nj + j)*2] = h[2*pl + o]; normal_eta[((size_t)i*v.nj + j)*2 + 1] = h[3*pl + o]; }
            if (volume && i < v.nic && j < v.j1) volume[(size_t)i*v.njc + j] = h[4*pl + o];
        }
    }
    return SGPU_OK;
}

// ---------------------------------------------------------------------------------------------- state
int sgpu_set_state_window(sgpu_ctx* c, int which, const double* q, int j_first, int j_count) {
    if (!c || !q || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    if (j_first > v.j0 || j_first + j_count < v.j1) FAIL(c, SGPU_ERR_ARG, "state window [%d,%d) does not cover the owned rows [%d,%d)", j_first, j_first + j_count, v.j0, v.j1);
    const int ja = std::max(std::max(v.j0 - 2, 0), j_first), jb = std::min(std::min(v.j1 + 2, v.njc), j_first + j_count);   // cell rows [ja, jb)
    const int nrows = jb - ja, r0 = ja - v.j0 + JOFF;
    const size_t M = (size_t)nrows*v.nv;
    if (int rc = ensure_stage(c, (size_t)v.nic*M)) return rc;
    CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*M, q + (size_t)(ja - j_first)*v.nv, sizeof(double)*j_count*v.nv, sizeof(double)*M, v.nic,
                            cudaMemcpyHostToDevice, c->stream));
    aos_to_planes_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, c->q[which], r0, nrows);
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_set_state(sgpu_ctx* c, int which, const double* q) {
    if (!c) return SGPU_ERR_ARG;
    return sgpu_set_state_window(c, which, q, 0, c->v.njc);
}
int sgpu_get_state(sgpu_ctx* c, int which, double* q) {
    if (!c || !q || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    return download_planes(c, c->q[which], c->v.nv, q);
}
int sgpu_get_rhs(sgpu_ctx* c, double* rhs) {
    if (!c || !rhs) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    return download_planes(c, c->rhs, c->v.nv, rhs);
}
int sgpu_get_rhs_window(sgpu_ctx* c, double* rhs) {
    if (!c || !rhs) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    return download_planes(c, c->rhs, c->v.nv, rhs, true);
}
int sgpu_get_dt(sgpu_ctx* c, double* dt) {
    if (!c || !dt) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    return download_planes(c, c->dt, 1, dt);
}
int sgpu_copy_state(sgpu_ctx* c, int dst, int src) {
    if (!c || dst < 0 || dst > 1 || src < 0 || src > 1) return SGPU_ERR_ARG;
    if (dst == src) return SGPU_OK;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaMemcpyAsync(c->q[dst], c->q[src], sizeof(double)*c->v.plane*c->v.nv, cudaMemcpyDeviceToDevice, c->stream));
    return SGPU_OK;
}

} // extern "C"
// ---------------------------------------------------------------------------------------------- hot path
// Applies the [[boundary]] tables in file order.  [jlo, jhi] (padded row indices, inclusive) restricts the work to
// the ghost cells a row range needs (pipelined host path); the default covers everything this slab holds.
static int apply_bcs(sgpu_ctx* c, int which, int jlo = -(1 << 30), int jhi = (1 << 30), int ilo = -(1 << 30), int ihi = (1 << 30)) {
    const View& v = c->v;
    Metrics m = metrics_of(c);
    for (const sgpu_bc& b : c->bcs) {
        BcArgs a; a.type = b.type; a.face = b.face; a.u = b.u; a.v = b.v; a.T = b.T;
        const bool horiz = b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP;
        if (horiz) {
            if (b.face == SGPU_FACE_BOTTOM && (v.j0 != 0 || jlo > 0)) continue;
            if (b.face == SGPU_FACE_TOP && (v.j1 != v.njc || jhi < v.njc + 1)) continue;
            a.lo = std::max(b.start, ilo); a.hi = std::min(b.end, ihi);          // [ilo, ihi]: padded column range
        } else {                                                   // rows this slab holds: padded j in [j0, j1+1]
            if (b.face == SGPU_FACE_LEFT && ilo > 0) continue;
            if (b.face == SGPU_FACE_RIGHT && ihi < v.nic + 1) continue;
            a.lo = std::max(std::max(b.start, v.j0), jlo); a.hi = std::min(std::min(b.end, v.j1 + 1), jhi);
        }
        if (a.hi < a.lo) continue;
        const int n = a.hi - a.lo + 1;
        bc_kernel<<<(n + 127)/128, 128, 0, c->stream>>>(v, c->g, m, c->q[which], a);
        CKL(c); c->launches++;
    }
    return SGPU_OK;
}

// Grid shape: strips of RCELLS columns x chunks of rows.  The chunk count is chosen so that the grid is an
// integer number of full waves of (SM count x resident CTAs per SM) -- a partial last wave would idle most
// SMs for a whole chunk -- while chunks stay tall enough to amortise the 4-row prologue.
static void shape_grid(const View& v, int ctas_per_sm, int sms, ResParams& p) {
    const int nrows = p.row1 - p.row0;
    if (p.nstrips <= 0) p.nstrips = (v.nic + RCELLS - 1)/RCELLS;
    const int wave = std::max(1, ctas_per_sm*sms);
    int best = 1; double best_cost = 1e300;
    const int max_chunks = std::max(1, nrows/8);
    for (int nch = 1; nch <= max_chunks; nch++) {
        const int rpc = (nrows + nch - 1)/nch;
        const int nchunks = (nrows + rpc - 1)/rpc;
        const long long ctas = (long long)p.nstrips*nchunks;
        const long long waves = (ctas + wave - 1)/wave;
        const double cost = (double)waves*(rpc + 5.0);          // time ~ waves x (rows + prologue) per CTA
        if (cost < best_cost - 1e-9) { best_cost = cost; best = nchunks; }
    }
    p.rpc = (nrows + best - 1)/best;
    p.nchunks = (nrows + p.rpc - 1)/p.rpc;
}

template <int NV, int ORDER, int FLUX, bool VISC, bool UPD>
static int launch_residual_t(sgpu_ctx* c, ResParams& p, int* grid_out) {
    using Cfg = ResCfg<NV, VISC>;
    static int occ_dev[64] = {0}, sms_dev[64] = {0};           // function attributes are per device
    auto kern = residual_kernel<NV, ORDER, FLUX, VISC, UPD>;
    int& occ = occ_dev[c->device & 63]; int& sms = sms_dev[c->device & 63];
    if (!occ) {
        CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes));
        CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RW, Cfg::smem_bytes));
        CK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (occ < 1) occ = 1;
    }
    shape_grid(c->v, occ, sms, p);
    const int grid = p.nstrips*p.nchunks;
    if (p.partial) {
        const size_t need = (size_t)grid*NV;
        if (need > c->partial_cap) {
            if (c->partial) CK(c, cudaFree(c->partial));
            CK(c, cudaMalloc(&c->partial, need*sizeof(double))); c->partial_cap = need;
        }
        p.partial = c->partial;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        if (c->ev_used == c->ev.size()) {
            cudaEvent_t a, b; CK(c, cudaEventCreate(&a)); CK(c, cudaEventCreate(&b)); c->ev.emplace_back(a, b);
        }
        e0 = c->ev[c->ev_used].first; e1 = c->ev[c->ev_used].second; c->ev_used++;
        CK(c, cudaEventRecord(e0, c->stream));
    }
    kern<<<grid, RW, Cfg::smem_bytes, c->stream>>>(p);
    CKL(c);
    if (c->timing) CK(c, cudaEventRecord(e1, c->stream));
    *grid_out = grid;
    return SGPU_OK;
}

// fused stage update handed to the residual kernel's epilogue: dst = q + rhs*dt/div on the owned cells (dst must not be the state evaluated)
struct StageUpdate { const double* q; const double* dt; double* dst; double div; double* dst2 = nullptr; };

static int launch_residual(sgpu_ctx* c, int which, int lhs, bool want_norms, int row0 = 0, int row1 = -1, int strip0 = 0, int nstrips = 0,
                           const StageUpdate* upd = nullptr) {
    const View& v = c->v;
    ResParams p;
    if (upd) { p.uq = upd->q; p.udt = upd->dt; p.udst = upd->dst; p.udst2 = upd->dst2; p.udiv = upd->div; p.uzinv = 1.0/upd->div; }
    p.row0 = row0; p.row1 = row1 < 0 ? v.njl : row1;
    p.strip0 = strip0; p.nstrips = nstrips;
    p.v = v; p.g = c->g; p.m = metrics_of(c);
    p.q = c->q[which]; p.rhs = c->rhs; p.wdist = c->wdist; p.beta = c->beta;
    p.eps_chi = c->eps_chi; p.eps_eta = c->eps_eta; p.dpdx = c->d.dpdx; p.dpdy = c->d.dpdy;
    p.eps12_chi = c->eps_chi/12.0; p.epsh_chi = 0.5*c->eps_chi; p.eps12_eta = c->eps_eta/12.0; p.epsh_eta = 0.5*c->eps_eta;
    p.partial = want_norms ? (double*)1 : nullptr;              // placeholder: sized once the grid is known
    int grid = 0;
    const int order = lhs ? c->d.lhs_order : c->d.order;           // eulerequation.cpp:203-208
    int rc = SGPU_ERR_ARG;
    const bool roe = c->d.flux == SGPU_FLUX_ROE;
#define RES_CASE(NV_, ORD_, FL_, VI_) rc = upd ? launch_residual_t<NV_, ORD_, FL_, VI_, true>(c, p, &grid) : launch_residual_t<NV_, ORD_, FL_, VI_, false>(c, p, &grid)
    if (v.nv == 5) {
        if (order == 2) { if (roe) RES_CASE(5, 2, SGPU_FLUX_ROE, true); else RES_CASE(5, 2, SGPU_FLUX_AUSM, true); }
        else            { if (roe) RES_CASE(5, 1, SGPU_FLUX_ROE, true); else RES_CASE(5, 1, SGPU_FLUX_AUSM, true); }
    } else if (c->viscous) {
        if (order == 2) { if (roe) RES_CASE(4, 2, SGPU_FLUX_ROE, true); else RES_CASE(4, 2, SGPU_FLUX_AUSM, true); }
        else            { if (roe) RES_CASE(4, 1, SGPU_FLUX_ROE, true); else RES_CASE(4, 1, SGPU_FLUX_AUSM, true); }
    } else {
        if (order == 2) { if (roe) RES_CASE(4, 2, SGPU_FLUX_ROE, false); else RES_CASE(4, 2, SGPU_FLUX_AUSM, false); }
        else            { if (roe) RES_CASE(4, 1, SGPU_FLUX_ROE, false); else RES_CASE(4, 1, SGPU_FLUX_AUSM, false); }
    }
#undef RES_CASE
    if (rc) return rc;
    c->launches++;
    if (want_norms) {
        reduce_partials_kernel<<<v.nv, 256, 0, c->stream>>>(c->partial, grid, v.nv, c->l2sq_dev);
        CKL(c); c->launches++;
    }
    return SGPU_OK;
}

extern "C" {
int sgpu_residual(sgpu_ctx* c, int which, int lhs, double* l2sq) {
    if (!c || which < 0 || which > 1) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    if (int rc = apply_bcs(c, which)) return rc;
    if (int rc = launch_residual(c, which, lhs, l2sq != nullptr)) return rc;
    if (c->wall_track && c->v.j0 == 0 && c->v.njl >= 2) {         // keep what this evaluation would have left in grad_{u,v}_eta[i][0]
        if (!c->wall_last) CK(c, cudaMalloc(&c->wall_last, 9*(size_t)c->v.nic*sizeof(double)));
        wall_data_kernel<<<(c->v.nic + 127)/128, 128, 0, c->stream>>>(c->v, metrics_of(c), c->q[which], c->q[which], c->xv, c->yv, c->wall_last);
        CKL(c); c->launches++;
        c->wall_valid = true;
    }
    if (l2sq) {
        CK(c, cudaMemcpyAsync(l2sq, c->l2sq_dev, sizeof(double)*c->v.nv, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
    }
    return SGPU_OK;
}

} // extern "C"

// streams and events of the pipelined host paths (created on first use)
static int pipe_setup(sgpu_ctx* c) {
    if (c->pipe_init) return SGPU_OK;
    for (int k = 0; k < 5; k++) CK(c, cudaStreamCreateWithFlags(&c->pipe_stream[k], cudaStreamNonBlocking));
    for (int k = 0; k < 64; k++)
        for (cudaEvent_t* e : {&c->pipe_up[k], &c->pipe_cmp[k], &c->pipe_h2d[k], &c->pipe_tout[k], &c->pipe_d2h[k]}) CK(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->pipe_start, cudaEventDisableTiming));
    c->pipe_init = true;
    return SGPU_OK;
}

// Column-chunk pipeline (preferred): the host arrays are [i][j][k], so a range of i is ONE contiguous block -- the
// H2D and D2H copies are plain 1-D transfers (measured 98 GB/s aggregate full duplex on this box vs 75 GB/s for the
// strided 2-D copies a row-chunk pipeline needs).  Chunks are whole strips of the residual kernel.
static int residual_host_pipelined_cols(sgpu_ctx* c, const double* q, int qj0, int qjn, double* rhs, int rj0, int rjn, int lhs) {
    const View& v = c->v;
    const int nstrips_all = (v.nic + RCELLS - 1)/RCELLS;
    int nch = std::min(14, nstrips_all/2);                      // measured at 4096^2 (69 strips), ramped sizes, copy streams of their own: 10 chunks 1120, 14 1128, 18 1121 Mcell/s
    if (const char* e = getenv("SGPU_PIPE_CHUNKS")) nch = std::max(1, std::min(atoi(e), nstrips_all));
    nch = std::min(nch, 62);                                    // the event arrays hold 64 entries (one is the wrap pre-upload's)
    const int jlo = std::max(std::max(v.j0 - 2, 0), qj0), jhi = std::min(std::min(v.j1 + 2, v.njc), qj0 + qjn);   // rows uploaded
    const int nrows_up = jhi - jlo, r0_up = jlo - v.j0 + JOFF;
    bool vert_periodic = false;
    for (const sgpu_bc& b : c->bcs) if (b.type == SGPU_BC_PERIODIC && (b.face == SGPU_FACE_LEFT || b.face == SGPU_FACE_RIGHT)) vert_periodic = true;
    if (int rc = pipe_setup(c)) return rc;
    const int strips_per = (nstrips_all + nch - 1)/nch;
    nch = (nstrips_all + strips_per - 1)/strips_per;
    // chunk boundaries in strips.  Uniform chunks leave the link half idle while the pipeline fills (only H2D runs until the
    // first chunk is up) and drains (only D2H after the last kernel): the chunk sizes RAMP 1, 2, .. strips_per at the head and
    // back down at the tail, so fill and drain cost one strip's copy each instead of a full chunk's (SGPU_PIPE_RAMP=0: uniform)
    std::vector<int> bound(1, 0);
    {
        const char* e = getenv("SGPU_PIPE_RAMP");
        const int ramp = strips_per*(strips_per - 1);                   // strips in the two ramps
        if ((e && atoi(e) == 0) || ramp + strips_per > nstrips_all || nch + 2*(strips_per - 1) > 62) {
            for (int ch = 0; ch < nch; ch++) bound.push_back(std::min((ch + 1)*strips_per, nstrips_all));
        } else {
            for (int k = 1; k < strips_per; k++) bound.push_back(bound.back() + k);
            int mid = nstrips_all - ramp;
            if (mid % strips_per) { bound.push_back(bound.back() + mid % strips_per); mid -= mid % strips_per; }
            for (; mid > 0; mid -= strips_per) bound.push_back(bound.back() + strips_per);
            for (int k = strips_per - 1; k >= 1; k--) bound.push_back(bound.back() + k);
        }
        nch = (int)bound.size() - 1;
    }
    const size_t cols_max = (size_t)strips_per*RCELLS + 4;
    const size_t stage_dbl = cols_max*(size_t)std::max(nrows_up, v.njl)*v.nv;
    if (stage_dbl > c->pipe_stage_cap) {
        for (int k = 0; k < 4; k++) { if (c->pipe_stage[k]) CK(c, cudaFree(c->pipe_stage[k])); c->pipe_stage[k] = nullptr; }
        for (int k = 0; k < 4; k++) CK(c, cudaMalloc(&c->pipe_stage[k], stage_dbl*sizeof(double)));
        c->pipe_stage_cap = stage_dbl;
    }
    // five streams: the two copy engines get streams of their OWN (s_h2d, s_d2h) so that the next chunk's copy runs while the
    // previous chunk's transpose kernel does (on one stream the link idled for every transpose); events order them per chunk and
    // guard the two staging buffers of each direction
    cudaStream_t s_in = c->pipe_stream[0], s_cmp = c->pipe_stream[1], s_out = c->pipe_stream[2], s_h2d = c->pipe_stream[3], s_d2h = c->pipe_stream[4];
    cudaStream_t user = c->stream;
    CK(c, cudaEventRecord(c->pipe_start, user));
    for (int k = 0; k < 5; k++) CK(c, cudaStreamWaitEvent(c->pipe_stream[k], c->pipe_start, 0));
    const size_t Mq = (size_t)nrows_up*v.nv;                       // doubles per column in the upload staging
    const bool q_contig = (jlo == qj0 && nrows_up == qjn);         // the window IS the uploaded row range: 1-D copies
    auto upload_cols = [&](int ia, int ib, double* st, cudaEvent_t copied) -> int {    // columns [ia, ib)
        if (ib <= ia) return SGPU_OK;
        const int ni = ib - ia;
        if (q_contig) CK(c, cudaMemcpyAsync(st, q + (size_t)ia*qjn*v.nv, sizeof(double)*Mq*ni, cudaMemcpyHostToDevice, s_h2d));
        else CK(c, cudaMemcpy2DAsync(st, sizeof(double)*Mq, q + ((size_t)ia*qjn + (jlo - qj0))*v.nv, sizeof(double)*qjn*v.nv, sizeof(double)*Mq, ni, cudaMemcpyHostToDevice, s_h2d));
        CK(c, cudaEventRecord(copied, s_h2d));
        CK(c, cudaStreamWaitEvent(s_in, copied, 0));
        aos_to_planes_kernel<<<dim3((unsigned)((Mq + 31)/32), (ni + 31)/32), dim3(32, 8), 0, s_in>>>(v, st, c->q[0], r0_up, nrows_up, ia, ni);
        CKL(c); c->launches++;
        return SGPU_OK;
    };
    int rc = SGPU_OK;
    if (vert_periodic) {                                                               // wrap-around ghost source of chunk 0
        rc = upload_cols(std::max(v.nic - 2, 0), v.nic, c->pipe_stage[0], c->pipe_h2d[63]);
        CK(c, cudaStreamSynchronize(s_in));                                            // staging buffer 0 is reused right away (tiny copy)
    }
    for (int ch = 0; ch < nch && rc == SGPU_OK; ch++) {
        const int s0 = bound[ch], s1 = bound[ch + 1];
        const int a = s0*RCELLS, b = std::min(s1*RCELLS, v.nic);                       // cell columns of this chunk
        const int ul = ch == 0 ? 0 : std::min(a + 2, v.nic), uh = ch == nch - 1 ? v.nic : std::min(b + 2, v.nic);
        if (ch >= 2) CK(c, cudaStreamWaitEvent(s_h2d, c->pipe_up[ch - 2], 0));         // staging buffer ch & 1 has been transposed
        rc = upload_cols(ul, uh, c->pipe_stage[ch & 1], c->pipe_h2d[ch]);
        if (rc != SGPU_OK) break;
        CK(c, cudaEventRecord(c->pipe_up[ch], s_in));
        CK(c, cudaStreamWaitEvent(s_cmp, c->pipe_up[ch], 0));
        c->stream = s_cmp;
        // the left ghost column copies cells nic-1 INCLUDING their bottom/top ghosts (the corner ghosts of cell 0 are
        // "whatever the last BC wrote", src/model/bc.cpp:430-433): give the wrap source its own ghosts first
        if (vert_periodic && ch == 0) rc = apply_bcs(c, SGPU_STATE_Q, -(1 << 30), (1 << 30), std::max(v.nic - 1, 1), v.nic);
        if (rc == SGPU_OK) rc = apply_bcs(c, SGPU_STATE_Q, -(1 << 30), (1 << 30), a, b + 1);   // padded columns a .. b+1 (cells a-1 .. b)
        if (rc == SGPU_OK) rc = launch_residual(c, SGPU_STATE_Q, lhs, false, 0, -1, s0, s1 - s0);
        c->stream = user;
        if (rc != SGPU_OK) break;
        CK(c, cudaEventRecord(c->pipe_cmp[ch], s_cmp));
        CK(c, cudaStreamWaitEvent(s_out, c->pipe_cmp[ch], 0));
        if (ch >= 2) CK(c, cudaStreamWaitEvent(s_out, c->pipe_d2h[ch - 2], 0));        // staging buffer 2 + (ch & 1) has been downloaded
        {
            const int ni = b - a; const size_t M = (size_t)v.njl*v.nv;
            double* st = c->pipe_stage[2 + (ch & 1)];
            planes_to_aos_kernel<<<dim3((unsigned)((M + 31)/32), (ni + 31)/32), dim3(32, 8), 0, s_out>>>(v, st, c->rhs, JOFF, v.njl, v.nv, a, ni);
            CKL(c); c->launches++;
            CK(c, cudaEventRecord(c->pipe_tout[ch], s_out));
            CK(c, cudaStreamWaitEvent(s_d2h, c->pipe_tout[ch], 0));
            if (rj0 == v.j0 && rjn == v.njl) CK(c, cudaMemcpyAsync(rhs + (size_t)a*rjn*v.nv, st, sizeof(double)*M*ni, cudaMemcpyDeviceToHost, s_d2h));
            else CK(c, cudaMemcpy2DAsync(rhs + ((size_t)a*rjn + (v.j0 - rj0))*v.nv, sizeof(double)*rjn*v.nv, st, sizeof(double)*M, sizeof(double)*M, ni, cudaMemcpyDeviceToHost, s_d2h));
            CK(c, cudaEventRecord(c->pipe_d2h[ch], s_d2h));
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_d2h), e2 = cudaStreamSynchronize(s_out), e3 = cudaStreamSynchronize(s_cmp), e4 = cudaStreamSynchronize(s_in), e5 = cudaStreamSynchronize(s_h2d);
    if (rc != SGPU_OK) return rc;
    CK(c, e1); CK(c, e2); CK(c, e3); CK(c, e4); CK(c, e5);
    return SGPU_OK;
}

// Host-buffer form of calc_residual, software pipelined over j-chunks on three streams so that the PCIe link runs
// full duplex: H2D + AoS->SoA of chunk c+1  ||  boundary conditions + residual of chunk c  ||  SoA->AoS + D2H of
// chunk c-1.  q holds rows [qj0, qj0+qjn) with qjn*nv doubles per i-row; rhs holds rows [rj0, rj0+rjn).
static int residual_host_pipelined(sgpu_ctx* c, const double* q, int qj0, int qjn, double* rhs, int rj0, int rjn, int lhs) {
    const View& v = c->v;
    const int ulo = std::max(std::max(v.j0 - 2, 0), qj0), uhi = std::min(std::min(v.j1 + 2, v.njc), qj0 + qjn);
    bool horiz_periodic = false;
    for (const sgpu_bc& b : c->bcs) if (b.type == SGPU_BC_PERIODIC && (b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP)) horiz_periodic = true;
    bool wake = false;
    for (const sgpu_bc& b : c->bcs) if (b.type == SGPU_BC_WAKE) wake = true;          // mirrored columns: no column chunks
    if (!wake && (v.nic + RCELLS - 1)/RCELLS >= 4 && !getenv("SGPU_PIPE_ROWS")) return residual_host_pipelined_cols(c, q, qj0, qjn, rhs, rj0, rjn, lhs);
    int nch = std::min(8, v.njl/64);
    if (const char* e = getenv("SGPU_PIPE_CHUNKS")) nch = std::max(1, std::min(16, std::min(atoi(e), v.njl/8)));
    if (nch < 2 || horiz_periodic) {                             // small grids / wrap-around ghosts: plain sequence
        if (int rc = sgpu_set_state_window(c, SGPU_STATE_Q, q, qj0, qjn)) return rc;
        if (int rc = sgpu_residual(c, SGPU_STATE_Q, lhs, nullptr)) return rc;
        return download_planes(c, c->rhs, v.nv, rhs, !(rj0 == 0 && rjn == v.njc));
    }
    if (int rc = pipe_setup(c)) return rc;
    const int rows_per = (v.njl + nch - 1)/nch;
    const size_t stage_rows = (size_t)rows_per + 4;
    const size_t stage_dbl = (size_t)v.nic*stage_rows*v.nv;
    if (stage_dbl > c->pipe_stage_cap) {
        for (int k = 0; k < 4; k++) { if (c->pipe_stage[k]) CK(c, cudaFree(c->pipe_stage[k])); c->pipe_stage[k] = nullptr; }
        for (int k = 0; k < 4; k++) CK(c, cudaMalloc(&c->pipe_stage[k], stage_dbl*sizeof(double)));
        c->pipe_stage_cap = stage_dbl;
    }
    cudaStream_t s_in = c->pipe_stream[0], s_cmp = c->pipe_stream[1], s_out = c->pipe_stream[2];
    cudaStream_t user = c->stream;
    CK(c, cudaEventRecord(c->pipe_start, user));                 // order after whatever the caller enqueued before
    for (int k = 0; k < 3; k++) CK(c, cudaStreamWaitEvent(c->pipe_stream[k], c->pipe_start, 0));
    int rc = SGPU_OK;
    for (int ch = 0; ch < nch && rc == SGPU_OK; ch++) {
        const int a = v.j0 + ch*rows_per, b = std::min(a + rows_per, v.j1);          // owned rows of this chunk
        const int ul = ch == 0 ? ulo : std::min(a + 2, uhi), uh = ch == nch - 1 ? uhi : std::min(b + 2, uhi);
        // ---- upload rows [ul, uh)
        if (uh > ul) {
            const int nrows = uh - ul; const size_t M = (size_t)nrows*v.nv;
            double* st = c->pipe_stage[ch & 1];
            CK(c, cudaMemcpy2DAsync(st, sizeof(double)*M, q + (size_t)(ul - qj0)*v.nv, sizeof(double)*qjn*v.nv, sizeof(double)*M, v.nic, cudaMemcpyHostToDevice, s_in));
            aos_to_planes_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, s_in>>>(v, st, c->q[0], ul - v.j0 + JOFF, nrows);
            CKL(c); c->launches++;
        }
        CK(c, cudaEventRecord(c->pipe_up[ch], s_in));
        // ---- boundary conditions for the ghost cells this chunk reads + residual of its rows
        CK(c, cudaStreamWaitEvent(s_cmp, c->pipe_up[ch], 0));
        c->stream = s_cmp;
        rc = apply_bcs(c, SGPU_STATE_Q, a, b + 1);               // padded rows a .. b+1  (cells a-1 .. b)
        if (rc == SGPU_OK) rc = launch_residual(c, SGPU_STATE_Q, lhs, false, a - v.j0, b - v.j0);
        c->stream = user;
        if (rc != SGPU_OK) break;
        CK(c, cudaEventRecord(c->pipe_cmp[ch], s_cmp));
        // ---- download the chunk's rhs rows
        CK(c, cudaStreamWaitEvent(s_out, c->pipe_cmp[ch], 0));
        {
            const int nrows = b - a; const size_t M = (size_t)nrows*v.nv;
            double* st = c->pipe_stage[2 + (ch & 1)];
            planes_to_aos_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, s_out>>>(v, st, c->rhs, a - v.j0 + JOFF, nrows, v.nv);
            CKL(c); c->launches++;
            CK(c, cudaMemcpy2DAsync(rhs + (size_t)(a - rj0)*v.nv, sizeof(double)*rjn*v.nv, st, sizeof(double)*M, sizeof(double)*M, v.nic, cudaMemcpyDeviceToHost, s_out));
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_out), e2 = cudaStreamSynchronize(s_cmp), e3 = cudaStreamSynchronize(s_in);
    if (rc != SGPU_OK) return rc;
    CK(c, e1); CK(c, e2); CK(c, e3);
    return SGPU_OK;
}

extern "C" {

int sgpu_residual_host(sgpu_ctx* c, const double* q, double* rhs, int lhs) {
    if (!c || !q || !rhs) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    return residual_host_pipelined(c, q, 0, c->v.njc, rhs, 0, c->v.njc, lhs);
}
int sgpu_residual_host_window(sgpu_ctx* c, const double* q, int j_first, int j_count, double* rhs_owned, int lhs) {
    if (!c || !q || !rhs_owned) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    if (j_first > c->v.j0 || j_first + j_count < c->v.j1) FAIL(c, SGPU_ERR_ARG, "state window does not cover the owned rows");
    CK(c, cudaSetDevice(c->device));
    return residual_host_pipelined(c, q, j_first, j_count, rhs_owned, c->v.j0, c->v.njl, lhs);
}

int sgpu_calc_dt(sgpu_ctx* c, double cfl) {
    if (!c) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    dt_kernel<<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, metrics_of(c), c->q[0], c->dt, cfl, c->d.mu_inf);
    CKL(c); c->launches++;
    c->have_dt = true;
    return SGPU_OK;
}

int sgpu_rk_stage(sgpu_ctx* c, int order) {
    if (!c || order < 0 || order > 3) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    axpy_dt_div_kernel<<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, c->q[1], c->q[0], c->rhs, c->dt, 4.0 - order);
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_forward_euler(sgpu_ctx* c) {
    if (!c) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    axpy_dt_div_kernel<<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, c->q[0], c->q[0], c->rhs, c->dt, 1.0);
    CKL(c); c->launches++;
    return SGPU_OK;
}

int sgpu_explicit_step(sgpu_ctx* c, int scheme, double cfl, double* l2sq) {
    if (!c || (scheme != 0 && scheme != 1)) { if (c) c->err = "scheme not defined."; return SGPU_ERR_ARG; }   // solver.cpp:119
    // On a j-slab every residual evaluation is preceded by the ghost-row exchange over peer memory (sgpu_halo_push / pull: the
    // neighbours' receive buffers must be registered); l2sq then holds THIS slab's sums, which the caller adds over the ranks.
    const bool slab = c->v.j0 != 0 || c->v.j1 != c->v.njc;
    auto exchange = [&](int which) -> int {
        if (!slab) return SGPU_OK;
        if (int rc = sgpu_halo_push(c, which)) return rc;
        return sgpu_halo_pull(c, which);
    };
    if (int rc = sgpu_calc_dt(c, cfl)) return rc;                  // solver.cpp:66
    if (scheme == 0) {                                             // solver.cpp:103-106
        if (int rc = exchange(SGPU_STATE_Q)) return rc;
        if (int rc = sgpu_residual(c, SGPU_STATE_Q, 0, l2sq)) return rc;
        return sgpu_forward_euler(c);
    }
    const char* fe = getenv("SGPU_RK_FUSED");
    if (fe && atoi(fe) == 0) {                                     // the two-kernel form (A/B, and what a slab-partitioned caller runs)
        for (int order = 0; order < 4; order++) {                  // solver.cpp:109-112
            if (int rc = exchange(SGPU_STATE_Q_TMP)) return rc;
            if (int rc = sgpu_residual(c, SGPU_STATE_Q_TMP, 0, order == 3 ? l2sq : nullptr)) return rc;
            if (int rc = sgpu_rk_stage(c, order)) return rc;
        }
        return sgpu_copy_state(c, SGPU_STATE_Q, SGPU_STATE_Q_TMP); // solver.cpp:114
    }
    // Fused stages: the residual kernel's epilogue writes q_tmp' = q + rhs*dt/(4 - order) (update_rk4, solver.cpp:4-13) -- the
    // stage update as a separate pass moved 128 B per cell for 15 flops, 0.42 ms beside a 1.16 ms residual at 4096^2.  The
    // neighbours of a cell still read the OLD q_tmp while its new value is written, so the stages ping-pong between q_tmp and a
    // second buffer (a full copy of q_tmp at first use: ghost cells no boundary condition writes keep their fill value in both);
    // four stages leave the result in q_tmp itself.  Same arithmetic per cell as the two-kernel form, bit for bit.
    const View& v = c->v;
    CK(c, cudaSetDevice(c->device));
    if (!c->q_scratch) {
        CK(c, cudaMalloc(&c->q_scratch, sizeof(double)*v.plane*v.nv));
        CK(c, cudaMemcpyAsync(c->q_scratch, c->q[1], sizeof(double)*v.plane*v.nv, cudaMemcpyDeviceToDevice, c->stream));
    }
    int rc = SGPU_OK;
    int swaps = 0;
    for (int order = 0; order < 4 && rc == SGPU_OK; order++) {
        rc = exchange(SGPU_STATE_Q_TMP);
        if (rc != SGPU_OK) break;
        rc = apply_bcs(c, SGPU_STATE_Q_TMP);
        if (rc != SGPU_OK) break;
        StageUpdate upd{c->q[0], c->dt, c->q_scratch, 4.0 - order};
        // the last stage also writes q itself (q <- q_tmp, solver.cpp:114: each thread reads its own cell of q before it writes it;
        // q's ghost cells are re-made by the boundary-condition pass every consumer of them runs first)
        if (order == 3) upd.dst2 = c->q[0];
        rc = launch_residual(c, SGPU_STATE_Q_TMP, 0, order == 3 && l2sq != nullptr, 0, -1, 0, 0, &upd);
        if (rc != SGPU_OK) break;
        if (c->wall_track && v.j0 == 0 && v.njl >= 2) {            // as in sgpu_residual: the wall rows of this evaluation
            if (!c->wall_last) { cudaError_t e = cudaMalloc(&c->wall_last, 9*(size_t)v.nic*sizeof(double)); if (e != cudaSuccess) { rc = SGPU_ERR_CUDA; break; } }
            wall_data_kernel<<<(v.nic + 127)/128, 128, 0, c->stream>>>(v, metrics_of(c), c->q[1], c->q[1], c->xv, c->yv, c->wall_last);
            c->launches++;
            c->wall_valid = true;
        }
        ghost_frame_copy_kernel<<<dim3((v.pitch + 127)/128, 2*JOFF), 128, 0, c->stream>>>(v, c->q_scratch, c->q[1], 0);
        ghost_frame_copy_kernel<<<dim3((v.njl + 127)/128, v.pitch - v.nic), 128, 0, c->stream>>>(v, c->q_scratch, c->q[1], 1);
        c->launches += 2;
        std::swap(c->q[1], c->q_scratch); swaps++;                 // q_tmp now names the buffer just written
    }
    if (swaps & 1) std::swap(c->q[1], c->q_scratch);               // only after an error: keep the pointers where they were
    if (rc != SGPU_OK) return rc;
    if (l2sq) {
        CK(c, cudaMemcpyAsync(l2sq, c->l2sq_dev, sizeof(double)*v.nv, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
    }
    return SGPU_OK;
}

// ---------------------------------------------------------------------------------------------- surface output
// device part of sgpu_wall_data / sgpu_surface: BCs on which_res, one small kernel, one D2H of 9 nic doubles
static int wall_data_host(sgpu_ctx* c, int which_res, int which_q, std::vector<double>& h) {
    if (!c || which_res < SGPU_STATE_LAST_RESIDUAL || which_res > 1 || which_q < 0 || which_q > 1) return SGPU_ERR_ARG;
    const bool last = which_res == SGPU_STATE_LAST_RESIDUAL;
    if (last && !c->wall_valid) FAIL(c, SGPU_ERR_STATE, "no tracked residual evaluation: call sgpu_track_wall(ctx, 1) before the evaluation whose wall gradients are wanted");
    if (last) which_res = which_q;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    const View& v = c->v;
    if (v.j0 != 0 || v.njl < 2) FAIL(c, SGPU_ERR_STATE, "the wall data live on the slab that owns j = 0 (with at least two cell rows)");
    CK(c, cudaSetDevice(c->device));
    if (!last) if (int rc = apply_bcs(c, which_res)) return rc;
    const size_t n = 9*(size_t)v.nic;
    if (int rc = ensure_stage(c, n)) return rc;
    wall_data_kernel<<<(v.nic + 127)/128, 128, 0, c->stream>>>(v, metrics_of(c), c->q[which_res], c->q[which_q], c->xv, c->yv, c->stage);
    CKL(c); c->launches++;
    if (last) CK(c, cudaMemcpyAsync(c->stage, c->wall_last, 4*(size_t)v.nic*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));   // gu | gv of the tracked evaluation
    h.resize(n);
    CK(c, cudaMemcpyAsync(h.data(), c->stage, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return SGPU_OK;
}

extern "C" int sgpu_track_wall(sgpu_ctx* c, int on) {
    if (!c) return SGPU_ERR_ARG;
    c->wall_track = on != 0;
    if (!on) c->wall_valid = false;
    return SGPU_OK;
}

extern "C" int sgpu_wall_data(sgpu_ctx* c, int which_res, int which_q, double* grad_u, double* grad_v, double* p_row0, double* p_row1) {
    std::vector<double> h;
    if (int rc = wall_data_host(c, which_res, which_q, h)) return rc;
    const size_t n = (size_t)c->v.nic;
    if (grad_u) memcpy(grad_u, h.data(), 2*n*sizeof(double));
    if (grad_v) memcpy(grad_v, h.data() + 2*n, 2*n*sizeof(double));
    if (p_row0) memcpy(p_row0, h.data() + 4*n, n*sizeof(double));
    if (p_row1) memcpy(p_row1, h.data() + 5*n, n*sizeof(double));
    return SGPU_OK;
}

// IOManager::write_surface, src/utils/io.cpp:182-255: the per-column loop (:219-238) and the force rotation (:240-249),
// same statements in the same order on the host (O(nic) work; the field reads happened on the device above)
extern "C" int sgpu_surface(sgpu_ctx* c, int which_res, int which_q, int i_first, int count, double aoa,
                            double* xw, double* cp_out, double* cf_out, double* coeffs) {
    if (!c) return SGPU_ERR_ARG;
    if (i_first < 0 || count < 0 || i_first + count > c->v.nic) FAIL(c, SGPU_ERR_ARG, "surface range [%d, %d) outside the %d cell columns", i_first, i_first + count, c->v.nic);
    std::vector<double> h;
    if (int rc = wall_data_host(c, which_res, which_q, h)) return rc;
    const size_t n = (size_t)c->v.nic;
    const double* gu = h.data(); const double* gv = gu + 2*n; const double* p0 = gv + 2*n; const double* p1 = p0 + n;
    const double* xc0 = p1 + n; const double* dxs = xc0 + n; const double* dys = dxs + n;
    const sgpu_desc& d = c->d;
    double Fn_pressure = 0.0, Fc_pressure = 0.0, Fn_viscous = 0.0, Fc_viscous = 0.0;
    for (int i = i_first; i < i_first + count; i++) {
        const double qinf = 0.5*d.rho_inf*(d.u_inf*d.u_inf + d.v_inf*d.v_inf);
        const double cp = (0.5*(p0[i] + p1[i]) - d.p_inf)/qinf;
        const double tau = d.mu_inf*(gu[2*i + 1] - gv[2*i])/qinf;
        if (xw) xw[i - i_first] = xc0[i];
        if (cp_out) cp_out[i - i_first] = cp;
        if (cf_out) cf_out[i - i_first] = tau;
        const double dx = dxs[i], dy = dys[i];
        Fn_pressure = Fn_pressure - cp*dx;
        Fc_pressure = Fc_pressure + cp*dy;
        const double sfdiv = 2.0/3.0*(gu[2*i] + gv[2*i + 1]);
        const double sxx = d.mu_inf*(2.0*gu[2*i] - sfdiv)/qinf;
        const double syy = d.mu_inf*(2.0*gv[2*i + 1] - sfdiv)/qinf;
        Fn_viscous = Fn_viscous - tau*dy + syy*dx;
        Fc_viscous = Fc_viscous + tau*dx - sxx*dy;
    }
    if (coeffs) {
        const double ca = cos(aoa), sa = sin(aoa);
        coeffs[0] = -Fc_pressure*sa + Fn_pressure*ca; coeffs[1] = Fc_pressure*ca + Fn_pressure*sa;
        coeffs[2] = -Fc_viscous*sa + Fn_viscous*ca; coeffs[3] = Fc_viscous*ca + Fn_viscous*sa;
        coeffs[4] = coeffs[2] + coeffs[0]; coeffs[5] = coeffs[3] + coeffs[1];
    }
    return SGPU_OK;
}

// ---------------------------------------------------------------------------------------------- halos
int sgpu_halo_count(const sgpu_ctx* c) { return c ? 2*c->v.nv*c->v.nic : 0; }
static int halo_rows(const sgpu_ctx* c, int side, bool ghost) {
    // first of the two rows: own boundary rows (pack) or ghost rows (unpack)
    if (side == 0) return ghost ? 0 : JOFF;
    return ghost ? c->v.njl + JOFF : c->v.njl + JOFF - 2;
}
int sgpu_halo_pack(sgpu_ctx* c, int which, int side, double* buf) {
    if (!c || !buf || side < 0 || side > 1 || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_pack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, c->q[which], buf, halo_rows(c, side, false));
    CKL(c); c->launches++;
    return SGPU_OK;
}
// the two GHOST rows of a side in the layout of sgpu_halo_pack: lets a caller verify what an exchange delivered
int sgpu_halo_pack_ghost(sgpu_ctx* c, int which, int side, double* buf) {
    if (!c || !buf || side < 0 || side > 1 || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_pack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, c->q[which], buf, halo_rows(c, side, true));
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_halo_unpack(sgpu_ctx* c, int which, int side, const double* buf) {
    if (!c || !buf || side < 0 || side > 1 || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_unpack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, c->q[which], buf, halo_rows(c, side, true));
    CKL(c); c->launches++;
    return SGPU_OK;
}
// receive buffer layout per side: [2 slots][halo_count doubles] + one 64-bit sequence flag
static size_t halo_slot_doubles(const sgpu_ctx* c) { return (size_t)2*c->v.nv*c->v.nic; }
static unsigned long long* halo_flag(const sgpu_ctx* c, double* buf) { return (unsigned long long*)(buf + 2*halo_slot_doubles(c)); }

int sgpu_halo_recv_buffer(sgpu_ctx* c, int side, double** p) {
    if (!c || !p || side < 0 || side > 1) return SGPU_ERR_ARG;
    *p = c->halo_recv[side];
    return SGPU_OK;
}
int sgpu_halo_enable_peer(sgpu_ctx* c, int peer_device) {
    if (!c) return SGPU_ERR_ARG;
    if (peer_device == c->device) return SGPU_OK;
    CK(c, cudaSetDevice(c->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SGPU_OK; }
    CK(c, e);
    return SGPU_OK;
}
int sgpu_halo_set_peer(sgpu_ctx* c, int side, double* peer) {
    if (!c || side < 0 || side > 1) return SGPU_ERR_ARG;
    c->halo_peer[side] = peer;
    return SGPU_OK;
}
int sgpu_halo_ipc_handle(sgpu_ctx* c, int side, void* handle64) {
    if (!c || !handle64 || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, c->halo_recv[side]));
    memcpy(handle64, &h, 64);
    return SGPU_OK;
}
int sgpu_halo_open_peer(sgpu_ctx* c, int side, const void* handle64) {
    if (!c || !handle64 || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->halo_peer[side] = (double*)p;
    c->halo_peer_ipc[side] = true;
    return SGPU_OK;
}
int sgpu_halo_push(sgpu_ctx* c, int which) {
    if (!c || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    for (int side = 0; side < 2; side++) {                          // a neighbour without a registered buffer would leave its
        const bool has_nb = side == 0 ? v.j0 > 0 : v.j1 < v.njc;    // halo_wait_kernel spinning forever: refuse instead
        if (has_nb && !c->halo_peer[side]) FAIL(c, SGPU_ERR_STATE, "halo push: side %d has a neighbour slab but no peer receive buffer (sgpu_halo_set_peer / sgpu_halo_open_peer)", side);
    }
    c->halo_seq++;
    const size_t slot = (size_t)(c->halo_seq & 1ull)*halo_slot_doubles(c);
    for (int side = 0; side < 2; side++) {
        if (!c->halo_peer[side]) continue;
        // the two boundary rows go straight into the neighbour's receive slot (stores over NVLink), then the flag
        halo_pack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, c->q[which], c->halo_peer[side] + slot, halo_rows(c, side, false));
        CKL(c);
        halo_signal_kernel<<<1, 1, 0, c->stream>>>(halo_flag(c, c->halo_peer[side]), c->halo_seq);
        CKL(c); c->launches += 2;
    }
    return SGPU_OK;
}
int sgpu_halo_pull(sgpu_ctx* c, int which) {
    if (!c || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    if (c->halo_seq == 0) FAIL(c, SGPU_ERR_STATE, "halo pull before the first halo push");
    const size_t slot = (size_t)(c->halo_seq & 1ull)*halo_slot_doubles(c);
    for (int side = 0; side < 2; side++) {
        const bool has_nb = side == 0 ? v.j0 > 0 : v.j1 < v.njc;
        if (!has_nb) continue;
        if (!c->halo_peer[side]) FAIL(c, SGPU_ERR_STATE, "halo pull: side %d has a neighbour slab but no registered peer", side);
        halo_wait_kernel<<<1, 1, 0, c->stream>>>(halo_flag(c, c->halo_recv[side]), c->halo_seq, 20000000000ull, c->halo_err);
        CKL(c);
        halo_unpack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, c->q[which], c->halo_recv[side] + slot, halo_rows(c, side, true));
        CKL(c); c->launches += 2;
    }
    return SGPU_OK;
}

// ---------------------------------------------------------------------------------------------- timing
int sgpu_enable_kernel_timing(sgpu_ctx* c, int on) { if (!c) return SGPU_ERR_ARG; c->timing = on != 0; c->ev_used = 0; return SGPU_OK; }
int sgpu_kernel_times(sgpu_ctx* c, float* ms, int n) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    int cnt = 0;
    for (size_t k = 0; k < c->ev_used && cnt < n; k++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->ev[k].first, c->ev[k].second) == cudaSuccess) ms[cnt++] = t;
    }
    c->ev_used = 0;
    return cnt;
}

// ---------------------------------------------------------------------------------------------- Jacobian
#include "jacobian_api.inl"
// ---------------------------------------------------------------------------------------------- linear solve
#include "linsolve_api.inl"
