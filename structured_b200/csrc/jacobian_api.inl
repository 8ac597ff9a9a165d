// Host side of the Jacobian entry points (included inside extern "C" of sgpu_api.cu).

} // extern "C"
#include <cub/device/device_scan.cuh>

// Replay the [[boundary]] tables in file order and record, for every ghost cell, the LAST condition that
// writes it together with its source cells (BoundaryContainer::apply, src/model/bc.cpp:430-433).
static int build_ghost_table(sgpu_ctx* c) {
    const int nic = c->v.nic, njc = c->v.njc;
    const size_t n = (size_t)2*(nic + 2) + 2*(njc + 2);
    std::vector<GhostDesc> tab(n);
    for (auto& g : tab) { g.type = -1; g.a_ip = g.a_jp = g.b_ip = g.b_jp = 0; g.face = 0; g.u = g.v = g.T = 0; }
    auto id = [&](int ip, int jp) -> size_t {
        if (jp <= 0) return ip; if (jp >= njc + 1) return (size_t)(nic + 2) + ip;
        if (ip <= 0) return (size_t)2*(nic + 2) + jp; return (size_t)2*(nic + 2) + (njc + 2) + jp;
    };
    auto set = [&](int ip, int jp, const sgpu_bc& b, int aip, int ajp, int bip, int bjp) {
        if (ip < 0 || ip > nic + 1 || jp < 0 || jp > njc + 1) return;
        GhostDesc& g = tab[id(ip, jp)];
        g.type = b.type; g.face = b.face; g.u = b.u; g.v = b.v; g.T = b.T;
        g.a_ip = aip; g.a_jp = ajp; g.b_ip = bip; g.b_jp = bjp;
    };
    for (const sgpu_bc& b : c->bcs) {
        const bool horiz = b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP;
        const bool bot = b.face == SGPU_FACE_BOTTOM, left = b.face == SGPU_FACE_LEFT;
        for (int s = b.start; s <= b.end; s++) {
            if (horiz) {
                const int jg = bot ? 0 : njc + 1, j1 = bot ? 1 : njc, j2 = bot ? 2 : njc - 1;
                switch (b.type) {
                case SGPU_BC_WAKE: set(s, jg, b, nic + 1 - s, 1, 0, 0); set(nic + 1 - s, jg, b, s, 1, 0, 0); break;
                case SGPU_BC_PERIODIC: set(s, 0, b, s, njc, 0, 0); set(s, njc + 1, b, s, 1, 0, 0); break;
                default: set(s, jg, b, s, j1, s, j2); break;
                }
            } else {
                const int ig = left ? 0 : nic + 1, i1 = left ? 1 : nic, i2 = left ? 2 : nic - 1;
                switch (b.type) {
                case SGPU_BC_PERIODIC: set(0, s, b, nic, s, 0, 0); set(nic + 1, s, b, 1, s, 0, 0); break;
                case SGPU_BC_OUTFLOW: set(ig, s, b, ig - 1, s, 0, 0); break;
                default: set(ig, s, b, i1, s, i2, s); break;
                }
            }
        }
    }
    if (!c->ghost_tab) CK(c, cudaMalloc(&c->ghost_tab, n*sizeof(GhostDesc)));
    CK(c, cudaMemcpy(c->ghost_tab, tab.data(), n*sizeof(GhostDesc), cudaMemcpyHostToDevice));
    if (!c->jac_err) CK(c, cudaMalloc(&c->jac_err, sizeof(int)));
    return SGPU_OK;
}

static GhostTable ghost_table_of(const sgpu_ctx* c) {
    GhostTable gt; gt.d = (const GhostDesc*)c->ghost_tab; gt.nic = c->v.nic; gt.njc = c->v.njc; return gt;
}

template <int NV, int ORDER, int FLUX, bool VISC>
static int launch_jacobian_t(sgpu_ctx* c, const JacParams& p) {
    const View& v = c->v;
    // stage 1: every face differentiated once (chi faces i = 0..nic of the owned rows; eta faces j0..j1)
    jac_face_kernel<NV, ORDER, FLUX, VISC, 0><<<dim3((v.nic + 1 + 127)/128, v.njl), 128, 0, c->stream>>>(p);
    CKL(c);
    jac_face_kernel<NV, ORDER, FLUX, VISC, 1><<<dim3((v.nic + 127)/128, v.njl + 1), 128, 0, c->stream>>>(p);
    CKL(c);
    // stage 2: gather per row cell, SA source row, ghost fold
    jac_gather_kernel<NV, ORDER, VISC><<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(p);
    CKL(c);
    c->launches += 2;
    return SGPU_OK;
}

// single-pass marching build (jacobian_march.cuh): strips of 31 columns x row chunks, then the boundary-band fold
template <int NV, int ORDER, int FLUX, bool VISC>
static int launch_jacobian_march_t(sgpu_ctx* c, const JacParams& p) {
    using Cfg = JmCfg<NV, ORDER, VISC>;
    const View& v = c->v;
    auto kern = jac_march_kernel<NV, ORDER, FLUX, VISC>;
    static int occ_dev[64] = {0}, sms_dev[64] = {0};
    int& occ = occ_dev[c->device & 63]; int& sms = sms_dev[c->device & 63];
    if (!occ) {
        CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes));
        CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32*JM_WARPS, Cfg::smem_bytes));
        CK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (occ < 1) occ = 1;
    }
    if (!c->jgeo_valid) {                                           // static per-face geometry weights, once per grid
        if (!c->jgeo) CK(c, cudaMalloc(&c->jgeo, (size_t)2*JG_N*v.plane*sizeof(double)));
        CK(c, cudaMemsetAsync(c->jgeo, 0, (size_t)2*JG_N*v.plane*sizeof(double), c->stream));
        jac_geom_kernel<VISC><<<dim3((v.pitch + 127)/128, v.rows), 128, 0, c->stream>>>(v, p.m, c->jgeo, c->jgeo + (size_t)JG_N*v.plane);
        CKL(c); c->launches++;
        c->jgeo_valid = true;
    }
    JmParams jp;
    jp.gchi = c->jgeo; jp.geta = c->jgeo + (size_t)JG_N*v.plane;
    jp.v = v; jp.g = p.g; jp.m = p.m; jp.q = p.q; jp.J = p.J; jp.wdist = p.wdist; jp.beta = p.beta;
    jp.eps_chi = p.eps_chi; jp.eps_eta = p.eps_eta;
    jp.nstrips = (v.nic + JM_CELLS - 1)/JM_CELLS;
    // chunk count: an integer number of full waves of (SMs x resident CTAs) where possible, chunks tall enough to amortise
    // the prologue (ring fill + one extra eta core ~ 2 rows)
    const int wave = std::max(1, occ*sms);
    int best = 1; double best_cost = 1e300;
    for (int nch = 1; nch <= std::max(1, v.njl/8); nch++) {
        const int rpc = (v.njl + nch - 1)/nch, nchunks = (v.njl + rpc - 1)/rpc;
        const long long waves = ((long long)jp.nstrips*nchunks + wave - 1)/wave;
        const double cost = (double)waves*(rpc + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = nchunks; }
    }
    jp.rpc = (v.njl + best - 1)/best; jp.nchunks = (v.njl + jp.rpc - 1)/jp.rpc;
    kern<<<jp.nstrips*jp.nchunks, 32*JM_WARPS, Cfg::smem_bytes, c->stream>>>(jp);
    CKL(c);
    if (v.nic < 8 || v.njl < 8) jac_fold_kernel<NV><<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, p.g, p.m, p.gt, p.q, p.J, p.nslots, p.err, 0);
    else {                                                          // the boundary band only: two thin launches
        jac_fold_kernel<NV><<<dim3((v.nic + 127)/128, 4), 128, 0, c->stream>>>(v, p.g, p.m, p.gt, p.q, p.J, p.nslots, p.err, 1);
        jac_fold_kernel<NV><<<dim3((v.njl + 127)/128, 4), 128, 0, c->stream>>>(v, p.g, p.m, p.gt, p.q, p.J, p.nslots, p.err, 2);
    }
    CKL(c);
    c->launches += 1;
    return SGPU_OK;
}

static int jacobian_build(sgpu_ctx* c, float* build_ms) {
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    if (!c->ghost_tab) if (int rc = build_ghost_table(c)) return rc;
    const int nslots = c->d.lhs_order == 2 ? 13 : 9;
    const size_t need = (size_t)nslots*v.nv*v.nv*v.plane;
    if (need > c->jac.cap) {
        if (c->jac.blocks) CK(c, cudaFree(c->jac.blocks));
        c->jac.blocks = nullptr; c->jac.cap = 0;
        CK(c, cudaMalloc(&c->jac.blocks, need*sizeof(double)));
        c->jac.cap = need;
    }
    const char* jm = getenv("SGPU_JAC");
    const bool two_stage = jm && !strcmp(jm, "two_stage");          // the round-1 build (per-face scratch in HBM), kept for A/B runs
    const int stiles = (v.pitch + STILE - 1)/STILE;
    const size_t dir_s = two_stage ? (size_t)(v.rows + 1)*stiles*STILE*8*v.nv*v.nv : 0;   // per-face scratch of one direction: 8 stencil cells each
    const size_t need_s = 2*dir_s;
    if (need_s > c->jac_scratch_cap) {
        if (c->jac_scratch) CK(c, cudaFree(c->jac_scratch));
        c->jac_scratch = nullptr; c->jac_scratch_cap = 0;
        CK(c, cudaMalloc(&c->jac_scratch, need_s*sizeof(double)));
        c->jac_scratch_cap = need_s;
    }
    c->jac.slots = nslots; c->jac.valid = false;
    if (int rc = apply_bcs(c, SGPU_STATE_Q)) return rc;          // ghost values of the state the tape would have seen
    CK(c, cudaMemsetAsync(c->jac_err, 0, sizeof(int), c->stream));
    JacParams p;
    p.v = v; p.g = c->g; p.m = metrics_of(c); p.gt = ghost_table_of(c);
    p.q = c->q[0]; p.J = c->jac.blocks; p.wdist = c->wdist; p.beta = c->beta;
    p.Schi = c->jac_scratch; p.Seta = c->jac_scratch + dir_s; p.stiles = stiles;
    p.eps_chi = c->eps_chi; p.eps_eta = c->eps_eta; p.nslots = nslots; p.err = c->jac_err;
    struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } ev0, ev1;
    CK(c, cudaEventCreate(&ev0.e)); CK(c, cudaEventCreate(&ev1.e));
    const cudaEvent_t e0 = ev0.e, e1 = ev1.e;
    CK(c, cudaEventRecord(e0, c->stream));
    int rc = SGPU_ERR_ARG;
    const bool roe = c->d.flux == SGPU_FLUX_ROE;
    const int order = c->d.lhs_order;                              // calc_residual(..., lhs = true), src/solver/solver.cpp:80
#define JAC_CASE(NV_, ORD_, FL_, VI_) rc = two_stage ? launch_jacobian_t<NV_, ORD_, FL_, VI_>(c, p) : launch_jacobian_march_t<NV_, ORD_, FL_, VI_>(c, p)
    if (v.nv == 5) {
        if (order == 2) { if (roe) JAC_CASE(5, 2, SGPU_FLUX_ROE, true); else JAC_CASE(5, 2, SGPU_FLUX_AUSM, true); }
        else            { if (roe) JAC_CASE(5, 1, SGPU_FLUX_ROE, true); else JAC_CASE(5, 1, SGPU_FLUX_AUSM, true); }
    } else if (c->viscous) {
        if (order == 2) { if (roe) JAC_CASE(4, 2, SGPU_FLUX_ROE, true); else JAC_CASE(4, 2, SGPU_FLUX_AUSM, true); }
        else            { if (roe) JAC_CASE(4, 1, SGPU_FLUX_ROE, true); else JAC_CASE(4, 1, SGPU_FLUX_AUSM, true); }
    } else {
        if (order == 2) { if (roe) JAC_CASE(4, 2, SGPU_FLUX_ROE, false); else JAC_CASE(4, 2, SGPU_FLUX_AUSM, false); }
        else            { if (roe) JAC_CASE(4, 1, SGPU_FLUX_ROE, false); else JAC_CASE(4, 1, SGPU_FLUX_AUSM, false); }
    }
#undef JAC_CASE
    if (rc) return rc;
    c->launches++;
    CK(c, cudaEventRecord(e1, c->stream));
    int herr = 0;
    CK(c, cudaMemcpyAsync(&herr, c->jac_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f; CK(c, cudaEventElapsedTime(&ms, e0, e1));
    if (build_ms) *build_ms = ms;
    if (herr) FAIL(c, SGPU_ERR_ARG, "%d boundary couplings could not be represented in the block-stencil Jacobian", herr);
    c->jac.valid = true;
    return SGPU_OK;
}

// device / host temporaries that are released on every exit path (CK / FAIL return early)
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() const { return (T*)p; }
};
struct HostCoo {                                                   // the caller's arrays: handed over only on success
    unsigned int* r = nullptr; unsigned int* c = nullptr; double* v = nullptr;
    ~HostCoo() { free(r); free(c); free(v); }
};

// COO export of the cell rows jl in [jl0, jl1) (local rows of this slab) of the resident block-stencil Jacobian
template <int NV>
static int jacobian_export_t(sgpu_ctx* c, int jl0, int jl1, int* nnz, unsigned int** rind, unsigned int** cind, double** values, int lhs_transform) {
    const View& v = c->v;
    *nnz = 0; *rind = nullptr; *cind = nullptr; *values = nullptr;
    const int nrw = jl1 - jl0;
    const size_t nrows = (size_t)v.nic*nrw*NV;
    const GhostTable gt = ghost_table_of(c);
    const bool order2 = c->d.lhs_order == 2;
    DevBuf counts, offs, tmp, dr, dc, dv; size_t tmp_bytes = 0;
    CK(c, counts.alloc((nrows + 1)*sizeof(int)));
    CK(c, offs.alloc((nrows + 1)*sizeof(long long)));
    CK(c, cudaMemsetAsync(counts.p, 0, (nrows + 1)*sizeof(int), c->stream));
    const dim3 grd((v.nic + 127)/128, nrw);
    jac_count_kernel<NV><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, jl0, nrw, counts.as<int>());
    CKL(c); c->launches++;
    CK(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.as<int>(), offs.as<long long>(), (int)(nrows + 1), c->stream));
    CK(c, tmp.alloc(tmp_bytes));
    CK(c, cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.as<int>(), offs.as<long long>(), (int)(nrows + 1), c->stream));
    c->launches++;
    long long total = 0;
    CK(c, cudaMemcpyAsync(&total, offs.as<long long>() + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (total > 2147483647LL) FAIL(c, SGPU_ERR_OVERFLOW, "nnz = %lld does not fit the reference's `int nnz` (src/solver/solution.h:16); use sgpu_jacobian_device", total);
    const size_t n = (size_t)std::max<long long>(total, 1);
    CK(c, dr.alloc(n*sizeof(unsigned int))); CK(c, dc.alloc(n*sizeof(unsigned int))); CK(c, dv.alloc(n*sizeof(double)));
    jac_fill_kernel<NV><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, offs.as<long long>(), c->dt, lhs_transform,
                                                    jl0, nrw, dr.as<unsigned int>(), dc.as<unsigned int>(), dv.as<double>());
    CKL(c); c->launches++;
    // malloc: ownership passes to the caller, who free()s the three arrays as after sparse_jac (src/solver/solver.cpp:181-183)
    HostCoo h;
    h.r = (unsigned int*)malloc(n*sizeof(unsigned int)); h.c = (unsigned int*)malloc(n*sizeof(unsigned int)); h.v = (double*)malloc(n*sizeof(double));
    if (!h.r || !h.c || !h.v) FAIL(c, SGPU_ERR_ARG, "malloc of the COO arrays failed");
    CK(c, cudaMemcpyAsync(h.r, dr.p, (size_t)total*sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(h.c, dc.p, (size_t)total*sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(h.v, dv.p, (size_t)total*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *rind = h.r; *cind = h.c; *values = h.v; h.r = nullptr; h.c = nullptr; h.v = nullptr;
    *nnz = (int)total;
    return SGPU_OK;
}

extern "C" {

int sgpu_jacobian_device(sgpu_ctx* c, int* slots, float* build_ms) {
    if (!c) return SGPU_ERR_ARG;
    if (int rc = jacobian_build(c, build_ms)) return rc;
    if (slots) *slots = c->jac.slots;
    return SGPU_OK;
}

int sgpu_jacobian_coo(sgpu_ctx* c, int* nnz, unsigned int** rind, unsigned int** cind, double** values, int apply_lhs_transform) {
    if (!c || !nnz || !rind || !cind || !values) return SGPU_ERR_ARG;
    if (apply_lhs_transform && !c->have_dt) FAIL(c, SGPU_ERR_STATE, "the LHS transform needs dt: call sgpu_calc_dt first (src/solver/solver.cpp:66,167-170)");
    if (int rc = jacobian_build(c, nullptr)) return rc;
    return c->v.nv == 5 ? jacobian_export_t<5>(c, 0, c->v.njl, nnz, rind, cind, values, apply_lhs_transform)
                        : jacobian_export_t<4>(c, 0, c->v.njl, nnz, rind, cind, values, apply_lhs_transform);
}

int sgpu_jacobian_coo_rows(sgpu_ctx* c, int j_first, int j_count, int* nnz, unsigned int** rind, unsigned int** cind, double** values,
                           int apply_lhs_transform) {
    if (!c || !nnz || !rind || !cind || !values) return SGPU_ERR_ARG;
    if (!c->jac.valid) FAIL(c, SGPU_ERR_STATE, "no device Jacobian: call sgpu_jacobian_device first");
    if (apply_lhs_transform && !c->have_dt) FAIL(c, SGPU_ERR_STATE, "the LHS transform needs dt: call sgpu_calc_dt first (src/solver/solver.cpp:66,167-170)");
    const int ja = std::max(j_first, c->v.j0), jb = std::min(j_first + j_count, c->v.j1);
    if (jb <= ja) FAIL(c, SGPU_ERR_ARG, "rows [%d, %d) do not intersect the owned rows [%d, %d)", j_first, j_first + j_count, c->v.j0, c->v.j1);
    CK(c, cudaSetDevice(c->device));
    return c->v.nv == 5 ? jacobian_export_t<5>(c, ja - c->v.j0, jb - c->v.j0, nnz, rind, cind, values, apply_lhs_transform)
                        : jacobian_export_t<4>(c, ja - c->v.j0, jb - c->v.j0, nnz, rind, cind, values, apply_lhs_transform);
}

int sgpu_dres_dbeta(sgpu_ctx* c, double* out) {
    if (!c || !out) return SGPU_ERR_ARG;
    if (c->v.nv != 5) FAIL(c, SGPU_ERR_ARG, "d rhs / d beta exists only with the SA extension (ntrans = 1)");
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    if (int rc = apply_bcs(c, SGPU_STATE_Q)) return rc;
    DevBuf tmpb;
    CK(c, tmpb.alloc(v.plane*sizeof(double)));
    double* tmp = tmpb.as<double>();
    sa_dbeta_kernel<true><<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, c->g, metrics_of(c), c->q[0], c->wdist, tmp);
    CKL(c); c->launches++;
    // per-cell field [nic][njc]: reuse the state download with one plane broadcast, then compact on the host side
    std::vector<double> h((size_t)v.plane);
    CK(c, cudaMemcpyAsync(h.data(), tmp, v.plane*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < v.nic; i++) for (int jl = 0; jl < v.njl; jl++) out[(size_t)i*v.njc + v.j0 + jl] = h[v.at(jl + JOFF, i + IOFF)];
    return SGPU_OK;
}

int sgpu_surface_gradient(sgpu_ctx* c, int which, int i_first, int count, double aoa, const double* weights, double* dFdq) {
    if (!c || !weights || !dFdq || which < 0 || which > 1) return SGPU_ERR_ARG;
    if (!c->have_grid) FAIL(c, SGPU_ERR_STATE, "sgpu_set_grid has not been called");
    const View& v = c->v;
    if (v.j0 != 0 || v.j1 != v.njc) FAIL(c, SGPU_ERR_STATE, "sgpu_surface_gradient needs the whole grid on one context (boundary-condition chains cross slabs)");
    if (i_first < 0 || count < 0 || i_first + count > v.nic) FAIL(c, SGPU_ERR_ARG, "surface range [%d, %d) outside the %d cell columns", i_first, i_first + count, v.nic);
    CK(c, cudaSetDevice(c->device));
    if (!c->ghost_tab) if (int rc = build_ghost_table(c)) return rc;
    if (int rc = apply_bcs(c, which)) return rc;
    DevBuf gb;
    CK(c, gb.alloc(v.plane*v.nv*sizeof(double)));
    double* g = gb.as<double>();
    CK(c, cudaMemsetAsync(g, 0, v.plane*v.nv*sizeof(double), c->stream));
    // cl = -Fc sin(aoa) + Fn cos(aoa), cd = Fc cos(aoa) + Fn sin(aoa)   (io.cpp:240-246)
    const double ca = cos(aoa), sa = sin(aoa);
    const double a_np = weights[0]*ca + weights[1]*sa, a_cp = -weights[0]*sa + weights[1]*ca;
    const double a_nv = weights[2]*ca + weights[3]*sa, a_cv = -weights[2]*sa + weights[3]*ca;
    const double qinf = 0.5*c->d.rho_inf*(c->d.u_inf*c->d.u_inf + c->d.v_inf*c->d.v_inf);
    if (count > 0) {
        const GhostTable gt = ghost_table_of(c);
        const Metrics m = metrics_of(c);
        if (v.nv == 5) surface_grad_kernel<5><<<(count + 63)/64, 64, 0, c->stream>>>(v, c->g, m, gt, c->q[which], c->xv, c->yv, i_first, count, a_np, a_cp, a_nv, a_cv, c->d.mu_inf, qinf, g);
        else surface_grad_kernel<4><<<(count + 63)/64, 64, 0, c->stream>>>(v, c->g, m, gt, c->q[which], c->xv, c->yv, i_first, count, a_np, a_cp, a_nv, a_cv, c->d.mu_inf, qinf, g);
        CKL(c); c->launches++;
    }
    return download_planes(c, g, v.nv, dFdq);
}

int sgpu_jacobian_apply(sgpu_ctx* c, int transpose, const double* x, double* y) {
    if (!c || !x || !y) return SGPU_ERR_ARG;
    if (!c->jac.valid) FAIL(c, SGPU_ERR_STATE, "no device Jacobian: call sgpu_jacobian_device first");
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    DevBuf xb, yb;
    CK(c, xb.alloc(v.plane*v.nv*sizeof(double))); CK(c, yb.alloc(v.plane*v.nv*sizeof(double)));
    double *xp = xb.as<double>(), *yp = yb.as<double>();
    CK(c, cudaMemsetAsync(xp, 0, v.plane*v.nv*sizeof(double), c->stream));
    CK(c, cudaMemsetAsync(yp, 0, v.plane*v.nv*sizeof(double), c->stream));
    // x: GLOBAL host AoS -> planes (owned rows + the slab's ghost rows)
    {
        const int ja = std::max(v.j0 - 2, 0), jb = std::min(v.j1 + 2, v.njc);
        const int nrows = jb - ja, r0 = ja - v.j0 + JOFF;
        const size_t M = (size_t)nrows*v.nv;
        if (int rc = ensure_stage(c, (size_t)v.nic*M)) return rc;
        CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*M, x + (size_t)ja*v.nv, sizeof(double)*v.njc*v.nv, sizeof(double)*M, v.nic, cudaMemcpyHostToDevice, c->stream));
        aos_to_planes_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, xp, r0, nrows);
        CKL(c); c->launches++;
    }
    const dim3 grd((v.nic + 127)/128, v.njl);
    const GhostTable gt = ghost_table_of(c);
    const bool order2 = c->d.lhs_order == 2;
    if (v.nv == 5) jac_apply_kernel<5><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, xp, yp, transpose);
    else jac_apply_kernel<4><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, xp, yp, transpose);
    CKL(c); c->launches++;
    return download_planes(c, yp, v.nv, y);
}

} // extern "C"
