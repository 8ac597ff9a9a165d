// Host side of the device linear solve and of the device-resident implicit step (included by sgpu_api.cu).
// Replaces, for a Jacobian that stays on the GPU, the sequence
//     linearsolver->set_lhs(nnz, rind, cind, values); set_rhs(rhs); solve_and_update(q, UNDER_RELAXATION)
// of src/solver/solver.cpp:172-175 (src/linearsolver/ls_eigen.cpp:32-70: SparseLU; ls_petsc.cpp: GMRES + LU).
// Method: restarted GMRES, right-preconditioned (so the monitored residual is the true one), classical
// Gram-Schmidt with one optional re-orthogonalisation pass, Givens rotations on the host (m x m doubles).

struct LinWork {
    size_t n = 0;                  // doubles per vector
    int m = 0;                     // basis capacity (restart length)
    double* V = nullptr;           // (m+1) basis vectors
    double *w = nullptr, *z = nullptr, *x = nullptr, *b = nullptr, *u = nullptr;
    double *psi = nullptr, *g = nullptr;   // adjoint continuation: current psi, objective gradient (allocated on first use)
    double* Dinv = nullptr;        // block-Jacobi inverse (nv*nv planes) or the line factors Dinv, DA, DC (3*nv*nv planes)
    double* partial = nullptr;     // [blocks][ldp]
    double* hdev = nullptr;        // ldp doubles: projections, then norm^2 in the last used column
    double* ydev = nullptr;        // m doubles
    double* hhost = nullptr;       // pinned, ldp doubles
    int* err = nullptr;
    unsigned* ticket = nullptr;    // arrival counter of the in-kernel final reductions
    int blocks = 0, ldp = 0;       // blocks: grid of the flat vector kernels (and the capacity of `partial`)
    int blocks_dots = 0, blocks_upd = 0;   // grids of the two Gram-Schmidt kernels = exactly ONE resident wave each
    bool twisted = false;          // the line factors in Dinv are those of the twisted factorisation (line_apply_twisted_kernel reads them)
};

static void lin_free(LinWork*& L) {
    if (!L) return;
    cudaFree(L->V); cudaFree(L->w); cudaFree(L->z); cudaFree(L->x); cudaFree(L->b); cudaFree(L->u);
    cudaFree(L->psi); cudaFree(L->g);
    cudaFree(L->Dinv); cudaFree(L->partial); cudaFree(L->hdev); cudaFree(L->ydev); cudaFree(L->err); cudaFree(L->ticket);
    if (L->hhost) cudaFreeHost(L->hhost);
    delete L; L = nullptr;
}

static int lin_prepare(sgpu_ctx* c, int m) {
    const View& v = c->v;
    const size_t n = v.plane*v.nv;
    LinWork*& L = c->lin;
    if (L && (L->n != n || L->m < m)) lin_free(L);
    if (L) return SGPU_OK;
    L = new LinWork();
    L->n = n; L->m = m; L->ldp = SGPU_GMRES_MAX + 2;       // reduction scratch sized for the longest basis: the building blocks below never reallocate
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    // Grids = what is RESIDENT at once.  The Gram-Schmidt kernels are grid-stride loops that end in a last-block reduction: with
    // 4 x SMs blocks and 76 registers only 3 blocks fit an SM, so a quarter of the tiles ran in a second wave of one block per SM
    // at a third of the bandwidth (ncu launch list: 0.38 ms of every dots launch).  The flat kernels get 8 x SMs blocks.
    int occ_d = 1, occ_u = 1;
    CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_d, dots_kernel, DOT_THREADS, 0));
    CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, gs_update_kernel, DOT_THREADS, 0));
    L->blocks_dots = std::max(1, occ_d)*sms; L->blocks_upd = std::max(1, occ_u)*sms;
    L->blocks = std::max(8*sms, std::max(L->blocks_dots, L->blocks_upd));
    const size_t vb = n*sizeof(double);
    CK(c, cudaMalloc(&L->V, vb*(m + 1)));
    CK(c, cudaMalloc(&L->w, vb)); CK(c, cudaMalloc(&L->z, vb)); CK(c, cudaMalloc(&L->x, vb)); CK(c, cudaMalloc(&L->b, vb)); CK(c, cudaMalloc(&L->u, vb));
    CK(c, cudaMalloc(&L->Dinv, 3*v.plane*v.nv*v.nv*sizeof(double)));
    CK(c, cudaMalloc(&L->partial, (size_t)L->blocks*L->ldp*sizeof(double)));
    CK(c, cudaMalloc(&L->hdev, L->ldp*sizeof(double))); CK(c, cudaMalloc(&L->ydev, (m + 1)*sizeof(double)));
    CK(c, cudaMalloc(&L->err, sizeof(int)));
    CK(c, cudaMalloc(&L->ticket, sizeof(unsigned))); CK(c, cudaMemsetAsync(L->ticket, 0, sizeof(unsigned), c->stream));
    CK(c, cudaMallocHost(&L->hhost, L->ldp*sizeof(double)));
    // ghosts / padding of every vector stay 0 from here on: the kernels either write owned cells only or are flat combinations
    CK(c, cudaMemsetAsync(L->V, 0, vb*(m + 1), c->stream));
    for (double* p : {L->w, L->z, L->x, L->b, L->u}) CK(c, cudaMemsetAsync(p, 0, vb, c->stream));
    CK(c, cudaMemsetAsync(L->Dinv, 0, 3*v.plane*v.nv*v.nv*sizeof(double), c->stream));
    return SGPU_OK;
}

static inline bool mat_transposed(int matrix) { return matrix == SGPU_MAT_JT || matrix == SGPU_MAT_LHS_T; }
static inline int mat_op(int matrix) { return (matrix == SGPU_MAT_LHS || matrix == SGPU_MAT_LHS_T) ? OP_LHS : OP_J; }

// y = A x for matrix in {SGPU_MAT_LHS, SGPU_MAT_J, SGPU_MAT_JT, SGPU_MAT_LHS_T}
static int lin_apply_op(sgpu_ctx* c, int matrix, const double* x, double* y) {
    const View& v = c->v;
    const dim3 grd((v.nic + 127)/128, v.njl);
    const GhostTable gt = ghost_table_of(c);
    const bool order2 = c->d.lhs_order == 2;
    const int op = mat_op(matrix);
    if (mat_transposed(matrix)) {
        // gather form for the inner row cells (writes every owned y once), then the boundary band's atomic scatter;
        // interior slab edges: the ghost rows of y receive this slab's share of the neighbour's cells
        const int jl0 = v.j0 > 0 ? -JOFF : 0, jl1 = v.j1 < v.njc ? v.njl + JOFF : v.njl;
        const dim3 grdt((v.nic + 127)/128, jl1 - jl0);
        if (v.nv == 5) op_apply_t_kernel<5><<<grdt, 128, 0, c->stream>>>(v, c->jac.slots, c->viscous, order2, c->jac.blocks, x, y, jl0);
        else op_apply_t_kernel<4><<<grdt, 128, 0, c->stream>>>(v, c->jac.slots, c->viscous, order2, c->jac.blocks, x, y, jl0);
        CKL(c); c->launches++;
        if (v.nv == 5) op_apply_t_band_kernel<5><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, x, y);
        else op_apply_t_band_kernel<4><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, x, y);
        if (op == OP_LHS) {
            CKL(c); c->launches++;
            if (v.nv == 5) lhs_fixup_kernel<5><<<grdt, 128, 0, c->stream>>>(v, c->dt, x, y, jl0);
            else lhs_fixup_kernel<4><<<grdt, 128, 0, c->stream>>>(v, c->dt, x, y, jl0);
        }
    } else if (v.nv == 5) op_apply_kernel<5><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, c->dt, op, x, y);
    else op_apply_kernel<4><<<grd, 128, 0, c->stream>>>(v, gt, c->jac.slots, c->viscous, order2, c->jac.blocks, c->dt, op, x, y);
    CKL(c); c->launches++;
    return SGPU_OK;
}

static int lin_factor(sgpu_ctx* c, int matrix, int precond) {
    const View& v = c->v; LinWork* L = c->lin;
    const int op = mat_op(matrix);
    CK(c, cudaMemsetAsync(L->err, 0, sizeof(int), c->stream));
    if (precond == SGPU_PC_LINE_J) {
        // one lane per line (default) or NV lanes per line (SGPU_LINE_FACTOR=rows: 5x the warps and a shorter chain per lane, but the
        // Gauss-Jordan pivot search / row broadcasts become dependent shuffles -- measured 29.2 vs 22.7 ms at 4096^2, identical factors)
        const char* lf = getenv("SGPU_LINE_FACTOR");
        const bool serial = !(lf && !strcmp(lf, "rows"));
        // default: the TWISTED factorisation (two warps per 32 lines, eliminating from both ends towards the middle row: half the
        // sequential rows per warp); SGPU_LINE_TWISTED=0 or SGPU_LINE_FACTOR=rows select the one-directional forms
        const char* tw = getenv("SGPU_LINE_TWISTED");
        L->twisted = serial && !(tw && atoi(tw) == 0);
#define LINE_FACTOR(NV_) do { \
        if (L->twisted) { \
            CK(c, cudaFuncSetAttribute(line_factor_twisted_kernel<NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fact_twisted_bytes<NV_>())); \
            line_factor_twisted_kernel<NV_><<<(v.nic + 31)/32, 64, fact_twisted_bytes<NV_>(), c->stream>>>(v, c->jac.blocks, c->dt, op, c->jac.slots, L->Dinv, L->err); \
        } else if (serial) { \
            CK(c, cudaFuncSetAttribute(line_factor_kernel<NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fact_ring_bytes<NV_>())); \
            line_factor_kernel<NV_><<<(v.nic + 31)/32, 32, fact_ring_bytes<NV_>(), c->stream>>>(v, c->jac.blocks, c->dt, op, c->jac.slots, L->Dinv, L->err); \
        } else { \
            const int lpw = factr_lines<NV_>(); \
            CK(c, cudaFuncSetAttribute(line_factor_rows_kernel<NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)factr_ring_bytes<NV_>())); \
            line_factor_rows_kernel<NV_><<<(v.nic + lpw - 1)/lpw, 32, factr_ring_bytes<NV_>(), c->stream>>>(v, c->jac.blocks, c->dt, op, c->jac.slots, L->Dinv, L->err); \
        } } while (0)
        if (v.nv == 5) LINE_FACTOR(5); else LINE_FACTOR(4);
#undef LINE_FACTOR
    } else {
        const dim3 grd((v.nic + 127)/128, v.njl);
        if (v.nv == 5) bj_factor_kernel<5><<<grd, 128, 0, c->stream>>>(v, c->jac.blocks, c->dt, op, L->Dinv, L->err);
        else bj_factor_kernel<4><<<grd, 128, 0, c->stream>>>(v, c->jac.blocks, c->dt, op, L->Dinv, L->err);
    }
    CKL(c); c->launches++;
    int e = 0;
    CK(c, cudaMemcpyAsync(&e, L->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (e) FAIL(c, SGPU_ERR_STATE, "singular diagonal block in the preconditioner");
    return SGPU_OK;
}

static int lin_apply_pc(sgpu_ctx* c, int matrix, int precond, const double* r, double* z) {
    const View& v = c->v; LinWork* L = c->lin;
    if (precond == SGPU_PC_LINE_J) {
        const int nb = (v.nic + 31)/32;
#define LINE_APPLY(NV_, TR_) do { \
        CK(c, cudaFuncSetAttribute(line_apply_kernel<NV_, TR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)line_ring_bytes<NV_>())); \
        line_apply_kernel<NV_, TR_><<<nb, 32, line_ring_bytes<NV_>(), c->stream>>>(v, L->Dinv, r, z); } while (0)
#define LINE_APPLY_TW(NV_, TR_) do { \
        CK(c, cudaFuncSetAttribute(line_apply_twisted_kernel<NV_, TR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)line_twisted_bytes<NV_>())); \
        line_apply_twisted_kernel<NV_, TR_><<<nb, 64, line_twisted_bytes<NV_>(), c->stream>>>(v, L->Dinv, r, z); } while (0)
        if (L->twisted) {
            if (mat_transposed(matrix)) { if (v.nv == 5) LINE_APPLY_TW(5, true); else LINE_APPLY_TW(4, true); }
            else { if (v.nv == 5) LINE_APPLY_TW(5, false); else LINE_APPLY_TW(4, false); }
        } else if (mat_transposed(matrix)) { if (v.nv == 5) LINE_APPLY(5, true); else LINE_APPLY(4, true); }
        else { if (v.nv == 5) LINE_APPLY(5, false); else LINE_APPLY(4, false); }
#undef LINE_APPLY_TW
#undef LINE_APPLY
    } else {
        const dim3 grd((v.nic + 127)/128, v.njl);
        const int tr = mat_transposed(matrix);
        if (v.nv == 5) bj_apply_kernel<5><<<grd, 128, 0, c->stream>>>(v, L->Dinv, r, z, tr);
        else bj_apply_kernel<4><<<grd, 128, 0, c->stream>>>(v, L->Dinv, r, z, tr);
    }
    CKL(c); c->launches++;
    return SGPU_OK;
}

// L->hdev[j0 .. j0+cnt) = w . V_j on the device (final reduction inside the kernel); no synchronisation
static int lin_dots(sgpu_ctx* c, const double* w, const double* V, int cnt, int j0 = 0) {
    LinWork* L = c->lin;
    for (int g = 0; g < cnt; g += DOT_GROUP) {
        const int k = std::min(DOT_GROUP, cnt - g);
        dots_kernel<<<L->blocks_dots, DOT_THREADS, 0, c->stream>>>(w, V + (size_t)g*L->n, L->n, k, L->partial, L->ldp, j0 + g, L->hdev, L->ticket);
        CKL(c); c->launches++;
    }
    return SGPU_OK;
}
// w -= V h (h = L->hdev[0 .. cnt)) and L->hdev[cnt] = |w|^2 of the result, one pass
static int lin_gs_update(sgpu_ctx* c, double* w, const double* V, int cnt) {
    LinWork* L = c->lin;
    gs_update_kernel<<<L->blocks_upd, DOT_THREADS, 0, c->stream>>>(w, V, L->n, cnt, L->hdev, L->partial, L->ldp, cnt, L->hdev, L->ticket);
    CKL(c); c->launches++;
    return SGPU_OK;
}
static int lin_fetch(sgpu_ctx* c, int cnt) {
    LinWork* L = c->lin;
    CK(c, cudaMemcpyAsync(L->hhost, L->hdev, cnt*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return SGPU_OK;
}
static int lin_norm(sgpu_ctx* c, const double* a, double* out) {
    if (int rc = lin_dots(c, a, a, 1)) return rc;
    if (int rc = lin_fetch(c, 1)) return rc;
    *out = std::sqrt(c->lin->hhost[0]);
    return SGPU_OK;
}

// Solves A x = b with b in L->b; the solution is left in L->x.
static int lin_gmres(sgpu_ctx* c, int matrix, sgpu_linsolve* io, bool refactor = true) {
    LinWork* L = c->lin;
    const size_t n = L->n, vb = n*sizeof(double);
    const int m = std::min(io->restart > 0 ? io->restart : 30, L->m), G = L->blocks;   // L->m is the allocated capacity
    int precond = io->precond;
    const int max_iter = io->max_iter > 0 ? io->max_iter : 500;
    const double rtol = io->rtol > 0 ? io->rtol : 1e-10;
    struct Events { cudaEvent_t e[6] = {}; ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); } } evs;   // released on every exit path
    for (auto& x : evs.e) CK(c, cudaEventCreate(&x));
    const cudaEvent_t e0 = evs.e[0], e1 = evs.e[1], e2 = evs.e[2], t0 = evs.e[3], t1 = evs.e[4], t2 = evs.e[5];
    bool timed = false;
    io->matvec_ms = io->precond_ms = 0.0f;
    CK(c, cudaEventRecord(e0, c->stream));
    if (refactor) { if (int rc = lin_factor(c, matrix, precond)) return rc; }
    CK(c, cudaEventRecord(e1, c->stream));
    CK(c, cudaMemsetAsync(L->x, 0, vb, c->stream));
    double bnorm = 0.0;
    if (int rc = lin_norm(c, L->b, &bnorm)) return rc;
    io->iterations = 0; io->converged = 1; io->rel_residual = 0.0;
    std::vector<double> H((size_t)(m + 1)*m, 0.0), cs(m, 0.0), sn(m, 0.0), g(m + 1, 0.0), y(m, 0.0);
    double beta = bnorm;
    int iters = 0;
    bool first = true;
    while (bnorm > 0.0) {
        // r = b - A x  (x = 0 on the first cycle) into w
        if (first) { CK(c, cudaMemcpyAsync(L->w, L->b, vb, cudaMemcpyDeviceToDevice, c->stream)); first = false; }
        else {
            if (int rc = lin_apply_op(c, matrix, L->x, L->w)) return rc;
            axpby_kernel<<<G, 256, 0, c->stream>>>(L->w, L->b, n, 1.0, -1.0); CKL(c); c->launches++;
            if (int rc = lin_norm(c, L->w, &beta)) return rc;
        }
        io->rel_residual = beta/bnorm;
        if (beta <= rtol*bnorm || iters >= max_iter) break;
        // V0 = r/beta
        if (int rc = lin_dots(c, L->w, L->w, 1)) return rc;
        scale_rsqrt_kernel<<<G, 256, 0, c->stream>>>(L->V, L->w, n, L->hdev); CKL(c); c->launches++;
        std::fill(g.begin(), g.end(), 0.0); g[0] = beta;
        int k = 0;
        bool done = false;
        for (; k < m && iters < max_iter && !done; k++, iters++) {
            double* Vk = L->V + (size_t)k*n;
            const bool tm = !timed && iters == 1;                  // the second application: caches and clocks are warm
            if (tm) CK(c, cudaEventRecord(t0, c->stream));
            if (int rc = lin_apply_pc(c, matrix, precond, Vk, L->z)) return rc;
            if (tm) CK(c, cudaEventRecord(t1, c->stream));
            if (int rc = lin_apply_op(c, matrix, L->z, L->w)) return rc;
            if (tm) { CK(c, cudaEventRecord(t2, c->stream)); timed = true; }
            // classical Gram-Schmidt against V_0..V_k: projections (one pass over w per 16 basis vectors), then update + |w|^2
            // in one pass; everything stays on the device until the single fetch below
            if (int rc = lin_dots(c, L->w, L->V, k + 1)) return rc;
            if (int rc = lin_gs_update(c, L->w, L->V, k + 1)) return rc;
            std::vector<double> h(k + 2, 0.0);
            if (io->reorthogonalize) {
                if (int rc = lin_fetch(c, k + 1)) return rc;
                for (int j = 0; j <= k; j++) h[j] = L->hhost[j];
                if (int rc = lin_dots(c, L->w, L->V, k + 1)) return rc;
                if (int rc = lin_gs_update(c, L->w, L->V, k + 1)) return rc;
            }
            // next basis vector scaled from the device value of |w|^2 (column k+1)
            scale_rsqrt_kernel<<<G, 256, 0, c->stream>>>(L->V + (size_t)(k + 1)*n, L->w, n, L->hdev + (k + 1)); CKL(c); c->launches++;
            if (int rc = lin_fetch(c, k + 2)) return rc;
            for (int j = 0; j <= k; j++) h[j] += L->hhost[j];
            h[k + 1] = std::sqrt(L->hhost[k + 1]);
            // Givens rotations (host, O(m))
            for (int j = 0; j < k; j++) { const double t = cs[j]*h[j] + sn[j]*h[j + 1]; h[j + 1] = -sn[j]*h[j] + cs[j]*h[j + 1]; h[j] = t; }
            const double den = std::hypot(h[k], h[k + 1]);
            if (den == 0.0) { cs[k] = 1.0; sn[k] = 0.0; } else { cs[k] = h[k]/den; sn[k] = h[k + 1]/den; }
            h[k] = cs[k]*h[k] + sn[k]*h[k + 1];
            g[k + 1] = -sn[k]*g[k]; g[k] = cs[k]*g[k];
            for (int j = 0; j <= k; j++) H[(size_t)j*m + k] = h[j];
            if (std::fabs(g[k + 1]) <= rtol*bnorm || h[k + 1] == 0.0 || !std::isfinite(den)) done = true;
        }
        // y from the triangular system, x += M^-1 (V y)
        for (int j = k - 1; j >= 0; j--) {
            double s = g[j];
            for (int l = j + 1; l < k; l++) s -= H[(size_t)j*m + l]*y[l];
            y[j] = H[(size_t)j*m + j] != 0.0 ? s/H[(size_t)j*m + j] : 0.0;
        }
        if (k > 0) {
            CK(c, cudaMemcpyAsync(L->ydev, y.data(), k*sizeof(double), cudaMemcpyHostToDevice, c->stream));
            combine_kernel<<<G, 256, 0, c->stream>>>(L->u, L->V, n, k, L->ydev); CKL(c); c->launches++;
            if (int rc = lin_apply_pc(c, matrix, precond, L->u, L->z)) return rc;
            axpby_kernel<<<G, 256, 0, c->stream>>>(L->x, L->z, n, 1.0, 1.0); CKL(c); c->launches++;
            CK(c, cudaStreamSynchronize(c->stream));             // y is a host vector reused by the next cycle
        }
        if (k == 0) break;
    }
    CK(c, cudaEventRecord(e2, c->stream));
    CK(c, cudaEventSynchronize(e2));
    cudaEventElapsedTime(&io->setup_ms, e0, e1); cudaEventElapsedTime(&io->solve_ms, e1, e2);
    if (timed) { cudaEventElapsedTime(&io->precond_ms, t0, t1); cudaEventElapsedTime(&io->matvec_ms, t1, t2); }
    io->iterations = iters;
    io->converged = bnorm == 0.0 || io->rel_residual <= rtol;
    return SGPU_OK;
}

static int lin_check(sgpu_ctx* c, int matrix, sgpu_linsolve* io) {
    if (matrix < SGPU_MAT_LHS || matrix > SGPU_MAT_LHS_T) FAIL(c, SGPU_ERR_ARG, "matrix must be SGPU_MAT_LHS, SGPU_MAT_J, SGPU_MAT_JT or SGPU_MAT_LHS_T");
    if (io->precond != SGPU_PC_BLOCK_JACOBI && io->precond != SGPU_PC_LINE_J) FAIL(c, SGPU_ERR_ARG, "unknown preconditioner");
    if (io->restart < 0 || io->restart > SGPU_GMRES_MAX) FAIL(c, SGPU_ERR_ARG, "restart must be in 1..%d", SGPU_GMRES_MAX);
    if (!c->jac.valid) FAIL(c, SGPU_ERR_STATE, "no device Jacobian: call sgpu_jacobian_device first");
    if (mat_op(matrix) == OP_LHS && !c->have_dt) FAIL(c, SGPU_ERR_STATE, "the LHS matrix needs dt: call sgpu_calc_dt first (src/solver/solver.cpp:66,167-170)");
    if (c->v.j0 != 0 || c->v.j1 != c->v.njc) FAIL(c, SGPU_ERR_STATE, "the device linear solve drives a whole grid (one GPU); slab-partitioned solves are not built yet");
    return SGPU_OK;
}

// host AoS [nic][njc][nv] -> zero-padded planes
static int upload_planes(sgpu_ctx* c, const double* host, double* planes) {
    const View& v = c->v;
    const size_t M = (size_t)v.njl*v.nv;
    if (int rc = ensure_stage(c, (size_t)v.nic*M)) return rc;
    CK(c, cudaMemsetAsync(planes, 0, v.plane*v.nv*sizeof(double), c->stream));
    CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*M, host + (size_t)v.j0*v.nv, sizeof(double)*v.njc*v.nv, sizeof(double)*M, v.nic, cudaMemcpyHostToDevice, c->stream));
    aos_to_planes_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, planes, JOFF, v.njl);
    CKL(c); c->launches++;
    return SGPU_OK;
}

extern "C" {

int sgpu_linear_solve(sgpu_ctx* c, int matrix, const double* b, double* x, sgpu_linsolve* io) {
    if (!c || !io) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = lin_check(c, matrix, io)) return rc;
    if (int rc = lin_prepare(c, io->restart > 0 ? io->restart : 30)) return rc;
    const View& v = c->v;
    if (b) { if (int rc = upload_planes(c, b, c->lin->b)) return rc; }
    else {
        masked_copy_kernel<<<dim3((v.pitch + 127)/128, v.rows), 128, 0, c->stream>>>(v, c->rhs, c->lin->b);
        CKL(c); c->launches++;
    }
    if (int rc = lin_gmres(c, matrix, io)) return rc;
    if (x) return download_planes(c, c->lin->x, v.nv, x);
    return SGPU_OK;
}

int sgpu_adjoint_solve(sgpu_ctx* c, const double* g, double* psi, double cfl, int max_steps, double tol, sgpu_linsolve* io,
                       int* steps_out, double* rel_out) {
    return sgpu_adjoint_solve_ramp(c, g, psi, cfl, 1.0, cfl, max_steps, tol, io, steps_out, rel_out);
}

// The same continuation with the pseudo-time step RAMPED like the forward solver's CFL ramp (Solver::solve,
// src/solver/solver.cpp:211-214): step k uses CFL_k = min(cfl0 * growth^k, cfl_max) -- dt and the line factors are rebuilt
// when the CFL changes.  The SA adjoint operator is strongly non-normal: at a fixed moderate CFL the transient dominates for
// tens of steps; the ramp damps it first and then approaches Newton on J^T psi = -g.
int sgpu_adjoint_solve_ramp(sgpu_ctx* c, const double* g, double* psi, double cfl, double cfl_growth, double cfl_max, int max_steps, double tol,
                            sgpu_linsolve* io, int* steps_out, double* rel_out) {
    if (!c || !g || !psi || !io || !(cfl > 0.0) || !(cfl_growth >= 1.0)) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = sgpu_calc_dt(c, cfl)) return rc;
    if (int rc = lin_check(c, SGPU_MAT_LHS_T, io)) return rc;
    if (int rc = lin_prepare(c, io->restart > 0 ? io->restart : 30)) return rc;
    LinWork* L = c->lin;
    const size_t n = L->n, vb = n*sizeof(double);
    if (!L->psi) { CK(c, cudaMalloc(&L->psi, vb)); CK(c, cudaMalloc(&L->g, vb)); }
    if (int rc = upload_planes(c, g, L->g)) return rc;
    CK(c, cudaMemsetAsync(L->psi, 0, vb, c->stream));
    double gnorm = 0.0;
    if (int rc = lin_norm(c, L->g, &gnorm)) return rc;
    int steps = 0, total_iters = 0; double rel = gnorm > 0.0 ? 1.0 : 0.0;
    float setup = 0.0f, solve = 0.0f;
    const int nmax = max_steps > 0 ? max_steps : 50;
    const double stop = tol > 0.0 ? tol : 1e-8;
    while (gnorm > 0.0) {
        // rho = g + J^T psi (the negated adjoint residual) into L->b
        if (steps == 0) CK(c, cudaMemcpyAsync(L->b, L->g, vb, cudaMemcpyDeviceToDevice, c->stream));
        else {
            if (int rc = lin_apply_op(c, SGPU_MAT_JT, L->psi, L->b)) return rc;
            axpby_kernel<<<L->blocks, 256, 0, c->stream>>>(L->b, L->g, n, 1.0, 1.0); CKL(c); c->launches++;
            double rn = 0.0;
            if (int rc = lin_norm(c, L->b, &rn)) return rc;
            rel = rn/gnorm;
        }
        if (rel <= stop || steps >= nmax) break;
        // (delta/dt - J^T) dpsi = rho, psi += dpsi: backward Euler on d psi/d tau = J^T psi + g
        bool refactor = steps == 0;
        if (steps > 0 && cfl_growth > 1.0 && cfl < cfl_max) {
            cfl = std::min(cfl*cfl_growth, cfl_max);
            if (int rc = sgpu_calc_dt(c, cfl)) return rc;
            refactor = true;
        }
        if (int rc = lin_gmres(c, SGPU_MAT_LHS_T, io, refactor)) return rc;
        total_iters += io->iterations; setup += io->setup_ms; solve += io->solve_ms;
        axpby_kernel<<<L->blocks, 256, 0, c->stream>>>(L->psi, L->x, n, 1.0, 1.0); CKL(c); c->launches++;
        steps++;
    }
    io->iterations = total_iters; io->setup_ms = setup; io->solve_ms = solve; io->rel_residual = rel; io->converged = rel <= stop;
    if (steps_out) *steps_out = steps;
    if (rel_out) *rel_out = rel;
    return download_planes(c, L->psi, c->v.nv, psi);
}

// ---- building blocks on DEVICE vectors for a slab-partitioned Krylov solve (one process per GPU; the iteration is
//      driven by the host over torch.distributed: structured_b200/slab.py).  A vector is nv state planes of this slab
//      (sgpu_vec_size doubles); owned cells carry the data, the two ghost rows per interior slab edge are filled by
//      sgpu_vec_halo_pack -> transport -> sgpu_vec_halo_unpack before every operator application.
static int vec_check(sgpu_ctx* c, int matrix) {
    if (matrix < SGPU_MAT_LHS || matrix > SGPU_MAT_LHS_T) FAIL(c, SGPU_ERR_ARG, "matrix must be SGPU_MAT_LHS, SGPU_MAT_J, SGPU_MAT_JT or SGPU_MAT_LHS_T");
    if (!c->jac.valid) FAIL(c, SGPU_ERR_STATE, "no device Jacobian: call sgpu_jacobian_device first");
    if ((matrix == SGPU_MAT_LHS || matrix == SGPU_MAT_LHS_T) && !c->have_dt) FAIL(c, SGPU_ERR_STATE, "the LHS matrix needs dt: call sgpu_calc_dt first (src/solver/solver.cpp:66,167-170)");
    return SGPU_OK;
}
int sgpu_vec_size(const sgpu_ctx* c, long long* n) {
    if (!c || !n) return SGPU_ERR_ARG;
    *n = (long long)(c->v.plane*c->v.nv);
    return SGPU_OK;
}
int sgpu_vec_from_rhs(sgpu_ctx* c, double* vec) {
    if (!c || !vec) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    masked_copy_kernel<<<dim3((v.pitch + 127)/128, v.rows), 128, 0, c->stream>>>(v, c->rhs, vec);
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_vec_add_to_state(sgpu_ctx* c, int which, const double* vec, double omega) {
    if (!c || !vec || which < 0 || which > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    masked_axpy_kernel<<<dim3((v.nic + 127)/128, v.njl), 128, 0, c->stream>>>(v, c->q[which], vec, omega);
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_vec_halo_pack(sgpu_ctx* c, const double* vec, int side, double* buf) {
    if (!c || !vec || !buf || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_pack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, vec, buf, halo_rows(c, side, false));
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_vec_halo_unpack(sgpu_ctx* c, double* vec, int side, const double* buf) {
    if (!c || !vec || !buf || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_unpack_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, vec, buf, halo_rows(c, side, true));
    CKL(c); c->launches++;
    return SGPU_OK;
}
// transposed products on a slab partition: the ghost rows of y = A^T x hold this slab's contribution to the neighbour's
// boundary cells.  pack_ghost copies them out (layout of sgpu_halo_pack) and clears them; add accumulates the neighbour's
// buffer into this slab's two boundary rows on that side.
int sgpu_vec_halo_pack_ghost(sgpu_ctx* c, double* vec, int side, double* buf) {
    if (!c || !vec || !buf || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    const dim3 grd((v.nic + 255)/256, 2*v.nv);
    halo_pack_kernel<<<grd, 256, 0, c->stream>>>(v, vec, buf, halo_rows(c, side, true));
    CKL(c);
    halo_zero_kernel<<<grd, 256, 0, c->stream>>>(v, vec, halo_rows(c, side, true));
    CKL(c); c->launches += 2;
    return SGPU_OK;
}
int sgpu_vec_halo_add(sgpu_ctx* c, double* vec, int side, const double* buf) {
    if (!c || !vec || !buf || side < 0 || side > 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    halo_add_kernel<<<dim3((v.nic + 255)/256, 2*v.nv), 256, 0, c->stream>>>(v, vec, buf, halo_rows(c, side, false));
    CKL(c); c->launches++;
    return SGPU_OK;
}
// host <-> device vector (GLOBAL host arrays [nic][njc][nv]; owned rows; ghost rows / padding of the vector are zeroed)
int sgpu_vec_from_host(sgpu_ctx* c, const double* host, double* vec) {
    if (!c || !host || !vec) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    const View& v = c->v;
    CK(c, cudaMemsetAsync(vec, 0, v.plane*v.nv*sizeof(double), c->stream));
    const size_t M = (size_t)v.njl*v.nv;
    if (int rc = ensure_stage(c, (size_t)v.nic*M)) return rc;
    CK(c, cudaMemcpy2DAsync(c->stage, sizeof(double)*M, host + (size_t)v.j0*v.nv, sizeof(double)*v.njc*v.nv, sizeof(double)*M, v.nic, cudaMemcpyHostToDevice, c->stream));
    aos_to_planes_kernel<<<dim3((unsigned)((M + 31)/32), (v.nic + 31)/32), dim3(32, 8), 0, c->stream>>>(v, c->stage, vec, JOFF, v.njl);
    CKL(c); c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    return SGPU_OK;
}
int sgpu_vec_to_host(sgpu_ctx* c, const double* vec, double* host) {
    if (!c || !host || !vec) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    return download_planes(c, vec, c->v.nv, host);
}
int sgpu_op_apply(sgpu_ctx* c, int matrix, const double* x, double* y) {
    if (!c || !x || !y) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = vec_check(c, matrix)) return rc;
    return lin_apply_op(c, matrix, x, y);
}
int sgpu_precond_setup(sgpu_ctx* c, int matrix, int precond) {
    if (!c) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = vec_check(c, matrix)) return rc;
    if (precond != SGPU_PC_BLOCK_JACOBI && precond != SGPU_PC_LINE_J) FAIL(c, SGPU_ERR_ARG, "unknown preconditioner");
    if (int rc = lin_prepare(c, 1)) return rc;
    return lin_factor(c, matrix, precond);
}
int sgpu_precond_apply(sgpu_ctx* c, int matrix, int precond, const double* r, double* z) {
    if (!c || !r || !z) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (!c->lin) FAIL(c, SGPU_ERR_STATE, "sgpu_precond_setup has not been called");
    return lin_apply_pc(c, matrix, precond, r, z);
}

// Gram-Schmidt building blocks of a slab-partitioned Krylov solve: the SAME kernels sgpu_linear_solve runs, with the results left
// on the device so that the caller can all-reduce them between the projection and the update.
int sgpu_vec_dots(sgpu_ctx* c, const double* w, const double* V, int cnt, double* out) {
    if (!c || !w || !V || !out || cnt < 1 || cnt > SGPU_GMRES_MAX + 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = lin_prepare(c, 1)) return rc;
    LinWork* L = c->lin;
    for (int g = 0; g < cnt; g += DOT_GROUP) {
        const int k = std::min(DOT_GROUP, cnt - g);
        dots_kernel<<<L->blocks_dots, DOT_THREADS, 0, c->stream>>>(w, V + (size_t)g*L->n, L->n, k, L->partial, L->ldp, g, out, L->ticket);
        CKL(c); c->launches++;
    }
    return SGPU_OK;
}
int sgpu_vec_gs_update(sgpu_ctx* c, double* w, const double* V, int cnt, const double* h, double* normsq) {
    if (!c || !w || !V || !h || !normsq || cnt < 1 || cnt > SGPU_GMRES_MAX + 1) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = lin_prepare(c, 1)) return rc;
    LinWork* L = c->lin;
    // the kernel writes its reduction to out[ncol]: column 0 of the caller's scalar
    gs_update_kernel<<<L->blocks_upd, DOT_THREADS, 0, c->stream>>>(w, V, L->n, cnt, h, L->partial, L->ldp, 0, normsq, L->ticket);
    CKL(c); c->launches++;
    return SGPU_OK;
}
int sgpu_vec_scale_rsqrt(sgpu_ctx* c, double* dst, const double* src, const double* normsq) {
    if (!c || !dst || !src || !normsq) return SGPU_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (int rc = lin_prepare(c, 1)) return rc;
    scale_rsqrt_kernel<<<c->lin->blocks, 256, 0, c->stream>>>(dst, src, c->lin->n, normsq); CKL(c); c->launches++;
    return SGPU_OK;
}

int sgpu_implicit_step(sgpu_ctx* c, double cfl, double under_relaxation, sgpu_linsolve* io, double* l2sq) {
    if (!c || !io) return SGPU_ERR_ARG;
    if (int rc = sgpu_calc_dt(c, cfl)) return rc;                                   // solver.cpp:66
    if (int rc = sgpu_residual(c, SGPU_STATE_Q, 0, l2sq)) return rc;                // rhs with solver.order (solver.cpp:80-101)
    if (int rc = jacobian_build(c, nullptr)) return rc;                             // d rhs(lhs_order) / d q  (solver.cpp:80,156)
    if (int rc = sgpu_linear_solve(c, SGPU_MAT_LHS, nullptr, nullptr, io)) return rc;   // solver.cpp:162-175
    const size_t n = c->lin->n;
    axpby_kernel<<<c->lin->blocks, 256, 0, c->stream>>>(c->q[0], c->lin->x, n, under_relaxation, 1.0);   // ls_eigen.cpp:66-70
    CKL(c); c->launches++;
    return SGPU_OK;
}

} // extern "C"
