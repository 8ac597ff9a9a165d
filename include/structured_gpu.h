/*
 * structured_gpu.h -- C ABI of the B200 (sm_100a) residual + Jacobian path.
 *
 * Drop-in boundary for the hot path of anandpratap/structured.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree).
 * Plain pointers and sizes only; all functions return 0 (SGPU_OK) or a negative error code and
 * leave a message retrievable with sgpu_last_error().  One context per Mesh; calls on one context
 * are issued from one host thread (the reference is single threaded, src/main.cpp:8-44).
 *
 * Host array layouts are exactly the reference's rarray layouts (row-major, last index fastest):
 *   vertices  xv, yv      [ni][nj]              Mesh::xv.data()          src/utils/mesh.cpp:354-355
 *   state     q, rhs, dt  [nic][njc][nv]        Solution::q.data()       src/solver/solution.cpp:28,49-51
 *   fields    per cell    [nic][njc]
 * with nic = ni-1, njc = nj-1, nv = 4 + ntrans.  Flat index (i*njc + j)*nv + k is the row/column
 * numbering of the Jacobian (src/solver/solver.cpp:73-88,164-166).
 *
 * There is NO CPU fallback: every compute entry point runs CUDA kernels on the context's device and
 * fails with SGPU_ERR_CUDA when that is impossible.
 */
#ifndef STRUCTURED_GPU_H
#define STRUCTURED_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define SGPU_OK            0
#define SGPU_ERR_ARG      -1   /* bad argument / unsupported configuration */
#define SGPU_ERR_CUDA     -2   /* CUDA runtime error (message has the CUDA string) */
#define SGPU_ERR_STATE    -3   /* call order (e.g. residual before set_grid) */
#define SGPU_ERR_OVERFLOW -4   /* COO export would overflow `int nnz` (src/solver/solution.h:16) */

/* [[boundary]] type strings of src/model/bc.cpp:494-524 */
enum sgpu_bc_type {
    SGPU_BC_FREESTREAM = 0,      /* "freestream"      bc.cpp:26-60   */
    SGPU_BC_SLIPWALL = 1,        /* "slipwall"        bc.cpp:78-128  */
    SGPU_BC_WALL = 2,            /* "wall" adiabatic  bc.cpp:150-204 */
    SGPU_BC_ISOTHERMALWALL = 3,  /* "isothermalwall"  bc.cpp:388-427 */
    SGPU_BC_WAKE = 4,            /* "wake"            bc.cpp:224-263 */
    SGPU_BC_OUTFLOW = 5,         /* "outflow"         bc.cpp:284-311 */
    SGPU_BC_PERIODIC = 6         /* "periodic"        bc.cpp:329-365 */
};
/* face ids of src/model/bc.h:8-11 */
enum sgpu_face { SGPU_FACE_BOTTOM = 0, SGPU_FACE_RIGHT = 1, SGPU_FACE_TOP = 2, SGPU_FACE_LEFT = 3 };
/* solver.flux strings, src/model/eulerequation.cpp:123-128 */
enum sgpu_flux { SGPU_FLUX_ROE = 0, SGPU_FLUX_AUSM = 1 };
/* which state array a residual is evaluated on / an update writes (src/solver/solver.cpp:104,110) */
enum sgpu_state { SGPU_STATE_Q = 0, SGPU_STATE_Q_TMP = 1,
                  SGPU_STATE_LAST_RESIDUAL = -1 /* sgpu_wall_data / sgpu_surface only: the state the last tracked sgpu_residual saw */ };

/* One [[boundary]] table (src/model/bc.cpp:472-488). start/end index the PADDED arrays, inclusive;
 * a negative `end` is resolved as in BoundaryContainer::get_index (bc.cpp:436-457). */
typedef struct sgpu_bc {
    int type;      /* sgpu_bc_type */
    int face;      /* sgpu_face */
    int start;
    int end;
    double u, v, T;
} sgpu_bc;

/* Everything EulerEquation's constructor reads from Config (src/model/eulerequation.cpp:31-133,
 * src/utils/config.cpp:32-87).  gamma is the reference's constant 1.4 (src/common.h:40). */
typedef struct sgpu_desc {
    int ni, nj;                 /* geometry.ni / geometry.nj : GLOBAL vertex counts */
    int ntrans;                 /* 0 = laminar (the reference), 1 = Spalart-Allmaras extension */
    int order, lhs_order;       /* solver.order / solver.lhs_order : 1 or 2 */
    int flux;                   /* sgpu_flux */
    double rho_inf, u_inf, v_inf, p_inf, T_inf, mu_inf, pr_inf;   /* [freestream] */
    double dpdx, dpdy;          /* [source] */
    int n_bc;
    const sgpu_bc* bc;          /* in config-file order (BoundaryContainer::apply, bc.cpp:430-433) */
    int device;                 /* CUDA device ordinal */
    /* j-slab partition for multi-GPU runs: this context owns global cell rows [j_begin, j_end).
     * j_begin = j_end = 0 means the whole grid. */
    int j_begin, j_end;
} sgpu_desc;

typedef struct sgpu_ctx sgpu_ctx;

/* ---- life cycle ---------------------------------------------------------------------------- */
/* replaces Mesh::setup -> Solution / EulerEquation construction (src/utils/mesh.cpp:392-414) */
int sgpu_create(const sgpu_desc* desc, sgpu_ctx** out);
int sgpu_destroy(sgpu_ctx* ctx);
const char* sgpu_last_error(const sgpu_ctx* ctx);       /* ctx may be NULL: last create() failure */
int sgpu_set_stream(sgpu_ctx* ctx, void* cuda_stream);  /* all later work is enqueued on this stream */
int sgpu_synchronize(sgpu_ctx* ctx);
int sgpu_dims(const sgpu_ctx* ctx, int* nic, int* njc, int* nv, int* j_begin, int* j_end);

/* ---- static inputs ------------------------------------------------------------------------- */
/* replaces Mesh::simple_loader / plot3d_loader output + Mesh::calc_metrics (src/utils/mesh.cpp:134-279):
 * GLOBAL vertex arrays [ni][nj]; the metrics are evaluated on the device from these. */
int sgpu_set_grid(sgpu_ctx* ctx, const double* xv, const double* yv);
/* Window forms for slab runs: the host arrays hold only vertex rows [jv_first, jv_first + jv_count) (shape
 * [ni][jv_count]) resp. cell rows [j_first, j_first + j_count) (shape [nic][j_count]); the window must cover the
 * slab's rows plus its two ghost rows on each interior side. */
int sgpu_set_grid_window(sgpu_ctx* ctx, const double* xv, const double* yv, int jv_first, int jv_count);
int sgpu_set_field_window(sgpu_ctx* ctx, const char* name, const double* field, int j_first, int j_count);
/* Large grids: the vertices straight from a BINARY file -- "SGRIDF64", int32 ni, int32 nj, x[nj][ni],
 * y[nj][ni] float64, the order of Mesh::plot3d_loader (src/utils/mesh.cpp:146-169) without the ASCII parse.  Only the
 * window of rows this slab needs is read; a file row is a row of the device vertex planes (no transpose). */
int sgpu_set_grid_file(sgpu_ctx* ctx, const char* path);
/* SA extension inputs, GLOBAL [nic][njc]: name = "wall_distance" | "beta" (no reference counterpart) */
int sgpu_set_field(sgpu_ctx* ctx, const char* name, const double* field);
int sgpu_get_field(sgpu_ctx* ctx, const char* name, double* field);   /* GLOBAL [nic][njc]; owned rows written */
/* Wall distance of the SA extension evaluated on the device (no reference counterpart; the reference has no turbulence
 * model, src/solver/solution.cpp:9): distance from every cell centre (Mesh::xc, yc, src/utils/mesh.cpp:199-200) to the
 * nearest point of the wall edges.  segments: host [nseg][4] = x0 y0 x1 y1; nseg = 0 sets d = 1e30.
 * _from_bcs derives the edges from the `wall` / `isothermalwall` [[boundary]] tables (any face, ranges as in
 * src/model/bc.cpp:436-457) and the GLOBAL host vertex arrays [ni][nj]. */
int sgpu_compute_wall_distance(sgpu_ctx* ctx, const double* segments, int nseg);
int sgpu_wall_distance_from_bcs(sgpu_ctx* ctx, const double* xv, const double* yv);
/* metrics as Mesh::calc_metrics leaves them, for parity checks: normal_chi [ni][njc][2],
 * normal_eta [nic][nj][2], volume [nic][njc] (GLOBAL shapes; only owned rows are written) */
int sgpu_get_metrics(sgpu_ctx* ctx, double* normal_chi, double* normal_eta, double* volume);

/* ---- state --------------------------------------------------------------------------------- */
/* Solution::q / q_tmp  (src/solver/solution.cpp:28,51), GLOBAL [nic][njc][nv] host arrays.
 * set: owned rows plus the slab's ghost rows are read.  get: only owned rows are written. */
int sgpu_set_state(sgpu_ctx* ctx, int which, const double* q);
int sgpu_get_state(sgpu_ctx* ctx, int which, double* q);
/* Same for a host array that only holds a WINDOW of rows: q is [nic][j_count][nv] = global rows
 * [j_first, j_first + j_count), which must cover the owned rows (ghost rows outside the window are left
 * to the halo exchange).  Lets each rank of a slab run keep only its own rows on the host. */
int sgpu_set_state_window(sgpu_ctx* ctx, int which, const double* q, int j_first, int j_count);
int sgpu_copy_state(sgpu_ctx* ctx, int dst, int src);   /* set_rarray, src/solver/solver.cpp:4-10,114 */
int sgpu_get_rhs(sgpu_ctx* ctx, double* rhs);            /* Solution::rhs */
int sgpu_get_rhs_window(sgpu_ctx* ctx, double* rhs);     /* rhs is [nic][j_end - j_begin][nv]: the owned rows only */
int sgpu_get_dt(sgpu_ctx* ctx, double* dt);              /* Solution::dt (all nv entries of a cell equal) */

/* ---- the hot path -------------------------------------------------------------------------- */
/* EulerEquation::calc_dt(cfl) on state q  (src/model/eulerequation.cpp:237-259) */
int sgpu_calc_dt(sgpu_ctx* ctx, double cfl);
/* EulerEquation::calc_residual(state, rhs, lhs)  (src/model/eulerequation.cpp:202-232): applies the
 * boundary conditions to `which`, evaluates rhs on the device.  If l2sq != NULL it receives, for each
 * of the nv equations, sum over owned cells of rhs^2 (the sums of src/solver/solver.cpp:125-134 before
 * the sqrt, so that slabs can be added) and the call synchronizes. */
int sgpu_residual(sgpu_ctx* ctx, int which, int lhs, double* l2sq);
/* Same through HOST buffers: q -> device, residual, rhs -> host.  This is the call-site form of
 * `equation->calc_residual(solution->q, solution->rhs)` (src/solver/solver.cpp:104). */
int sgpu_residual_host(sgpu_ctx* ctx, const double* q, double* rhs, int lhs);
/* slab form: q holds rows [j_first, j_first + j_count) (owned rows + ghost rows), rhs_owned is [nic][j_end-j_begin][nv].
 * Both host forms are software pipelined over row chunks (H2D || kernel || D2H on three streams). */
int sgpu_residual_host_window(sgpu_ctx* ctx, const double* q, int j_first, int j_count, double* rhs_owned, int lhs);
/* update_rk4: q_tmp = q + rhs*dt/(4-order)   (src/solver/solver.cpp:20-26,111) */
int sgpu_rk_stage(sgpu_ctx* ctx, int order);
/* update_forward_euler: q = q + rhs*dt       (src/solver/solver.cpp:12-18,105) */
int sgpu_forward_euler(sgpu_ctx* ctx);
/* The explicit branch of Solver::step (src/solver/solver.cpp:66,103-134) entirely on the device:
 * calc_dt, 1 (forward_euler) or 4 (rk4_jameson) residual evaluations + updates, then l2sq[nv]
 * of the last rhs.  scheme: 0 = forward_euler, 1 = rk4_jameson.  The Runge-Kutta stage update is written from the residual
 * kernel's epilogue (bit-identical to the residual + sgpu_rk_stage sequence).  On a j-slab (j_begin/j_end) every evaluation is
 * preceded by the ghost-row exchange over the registered peer buffers (sgpu_halo_set_peer / sgpu_halo_open_peer) and l2sq holds
 * THIS slab's sums: all ranks call it once per step and add their l2sq. */
int sgpu_explicit_step(sgpu_ctx* ctx, int scheme, double cfl, double* l2sq);

/* ---- Jacobian ------------------------------------------------------------------------------ */
/* Replaces trace_on .. trace_off + sparse_jac(1, nt, nt, repeat, q, &nnz, &rind, &cind, &values, options)
 * (src/solver/solver.cpp:72-90,156): d rhs / d q of calc_residual(q, lhs = true) at state q.
 * Output arrays are malloc()ed and owned by the caller, who free()s them exactly as after sparse_jac
 * (src/solver/solver.cpp:181-183).  Entries are sorted by (row, col); duplicates are summed.
 * apply_lhs_transform != 0 additionally performs the loop of src/solver/solver.cpp:162-171:
 * values = -values, diagonal += 1/dt[row] (sgpu_calc_dt must have been called). */
int sgpu_jacobian_coo(sgpu_ctx* ctx, int* nnz, unsigned int** rind, unsigned int** cind, double** values,
                      int apply_lhs_transform);
/* The same export for the cell rows j in [j_first, j_first + j_count) only, from the Jacobian a previous
 * sgpu_jacobian_device left on the device (no rebuild): large grids whose full COO would overflow `int nnz`, sampled
 * parity checks.  Row / column numbers stay GLOBAL. */
int sgpu_jacobian_coo_rows(sgpu_ctx* ctx, int j_first, int j_count, int* nnz, unsigned int** rind, unsigned int** cind,
                           double** values, int apply_lhs_transform);
/* Device-resident block-stencil form (no size limit): evaluates the Jacobian and leaves it on the device.
 * *slots = number of stencil slots per cell, block layout J[slot][r][c][cell]. build_ms (may be NULL)
 * receives the device time of the build. */
int sgpu_jacobian_device(sgpu_ctx* ctx, int* slots, float* build_ms);
/* y = J x  and  y = J^T x with the device-resident Jacobian (adjoint building block); x, y host [nic][njc][nv] */
int sgpu_jacobian_apply(sgpu_ctx* ctx, int transpose, const double* x, double* y);
/* SA extension / field inversion: d rhs[i][j][4] / d beta[i][j] at state q (diagonal), GLOBAL host [nic][njc]
 * (owned rows written).  Gradient of an objective w.r.t. the correction field = psi^T dR/dbeta. */
int sgpu_dres_dbeta(sgpu_ctx* ctx, double* out);

/* ---- surface output ------------------------------------------------------- */
/* What IOManager::write_surface (src/utils/io.cpp:182-255) reads: grad_u_eta[i][0][0..1], grad_v_eta[i][0][0..1] as
 * the last residual evaluation left them in EulerEquation's work arrays (Mesh::calc_gradient on the j = 0 faces,
 * src/utils/mesh.cpp:88-128) and Solution::p[i][0], p[i][1] of the final state (IOManager::write, io.cpp:41).
 * which_res = the state the last calc_residual saw (BCs are applied to it here); which_q = the final state.
 * Only the slab that owns j = 0 may call this.  Host outputs: grad_u, grad_v [nic][2]; p_row0, p_row1 [nic];
 * any may be NULL. */
int sgpu_wall_data(sgpu_ctx* ctx, int which_res, int which_q, double* grad_u, double* grad_v, double* p_row0, double* p_row1);
/* The reference's grad arrays are simply whatever its LAST calc_residual left behind -- in an RK4 step the stage-3 state,
 * which no longer exists once the step has updated q.  While tracking is on, every sgpu_residual (also inside
 * sgpu_explicit_step / sgpu_implicit_step) keeps its wall rows (one O(nic) kernel), and which_res =
 * SGPU_STATE_LAST_RESIDUAL selects them: the `.surface` file then equals the stock binary's. */
int sgpu_track_wall(sgpu_ctx* ctx, int on);
/* The loop of IOManager::write_surface over cell columns i in [i_first, i_first + count) (the reference uses
 * i_first = mesh->j1 - 1, count = mesh->nb, src/utils/io.cpp:219, src/utils/mesh.cpp:349-350): per column
 * xw = xc[i][0], cp, cf (each may be NULL), and coeffs[6] = {cl_pressure, cd_pressure, cl_viscous, cd_viscous, cl, cd}
 * with the angle of attack aoa in radians (config->freestream->aoa), summed in the reference's order. */
int sgpu_surface(sgpu_ctx* ctx, int which_res, int which_q, int i_first, int count, double aoa,
                 double* xw, double* cp, double* cf, double* coeffs);

/* d F / d q of the surface functional F = sum_k weights[k] coeffs[k], k = cl_pressure, cd_pressure, cl_viscous, cd_viscous
 * (the first four coefficients of sgpu_surface, evaluated with which_res = which_q = which): the right-hand side g of the
 * adjoint system (d rhs/d q)^T psi = -g for sgpu_adjoint_solve (no reference counterpart -- the
 * reference never differentiates its surface loop).  Ghost cells are chained through the boundary conditions that wrote
 * them.  dFdq: GLOBAL host [nic][njc][nv], fully written (zeros away from the wall). */
int sgpu_surface_gradient(sgpu_ctx* ctx, int which, int i_first, int count, double aoa, const double* weights, double* dFdq);

/* ---- device linear solve ---------------------------------------------------- */
/* Replaces, for a Jacobian that stays on the device, the reference's linear-solver plug-in
 *     linearsolver->set_lhs(nnz, rind, cind, values); set_rhs(rhs); solve_and_update(q, UNDER_RELAXATION)
 * (src/solver/solver.cpp:172-175; src/linearsolver/ls_eigen.h:26-31, ls_eigen.cpp:32-70; ls_petsc.cpp GMRES).
 * No COO is exported, so neither the `int nnz` limit nor the D2H of the matrix applies.
 * Method: restarted right-preconditioned GMRES on the block-stencil planes. */
enum sgpu_matrix {
    SGPU_MAT_LHS = 0,          /* A = -J + delta/dt: the matrix of src/solver/solver.cpp:162-171 (needs sgpu_calc_dt) */
    SGPU_MAT_J   = 1,          /* A = J  = d rhs / d q                                              */
    SGPU_MAT_JT  = 2,          /* A = J^T (adjoint systems)                          */
    SGPU_MAT_LHS_T = 3         /* A = (-J + delta/dt)^T: one pseudo-time step of the adjoint        */
};
enum sgpu_precond {
    SGPU_PC_BLOCK_JACOBI = 0,  /* inverse of the nv x nv diagonal block of every cell               */
    SGPU_PC_LINE_J       = 1   /* exact block-tridiagonal solve along every grid line i = const     */
};
#define SGPU_GMRES_MAX 200
typedef struct sgpu_linsolve {
    /* in (0 selects the default) */
    int    precond;            /* enum sgpu_precond                                                 */
    int    restart;            /* GMRES restart length m (default 30, at most SGPU_GMRES_MAX)       */
    int    max_iter;           /* cap on Krylov iterations over all restarts (default 500)          */
    int    reorthogonalize;    /* != 0: second Gram-Schmidt pass per iteration                      */
    double rtol;               /* stop when |b - A x| <= rtol |b| (default 1e-10)                   */
    /* out */
    int    iterations;
    int    converged;
    double rel_residual;       /* true |b - A x| / |b| of the returned x                            */
    float  setup_ms, solve_ms; /* device time: preconditioner factorisation, Krylov iterations      */
    float  matvec_ms, precond_ms; /* device time of ONE operator / preconditioner application (first iteration) */
} sgpu_linsolve;
/* Solves A x = b with the Jacobian of the last sgpu_jacobian_device / sgpu_jacobian_coo / sgpu_implicit_step.
 * b: host [nic][njc][nv], or NULL for the device rhs of the last sgpu_residual (what set_rhs receives).
 * x: host [nic][njc][nv] or NULL (the solution also stays on the device for sgpu_implicit_step). */
int sgpu_linear_solve(sgpu_ctx* ctx, int matrix, const double* b, double* x, sgpu_linsolve* io);
/* Steady adjoint of the discrete residual (no reference code -- "parity unpinned", checked against a
 * sparse LU of the transposed COO matrix):  J^T psi = -g  with J = d rhs / d q of the last Jacobian build and
 * g = d objective / d q, host [nic][njc][nv].  Solved by pseudo-time continuation, each step one GMRES solve with the
 * transposed LHS matrix:  (delta/dt - J^T) dpsi = g + J^T psi,  psi += dpsi,  dt = calc_dt(cfl), until
 * |g + J^T psi| <= tol |g| or max_steps.  io: GMRES controls for the inner solves; on return io->iterations is the
 * total over all steps and io->rel_residual the outer residual.  The field-inversion gradient is then
 * d objective / d beta = psi[..., 4] * sgpu_dres_dbeta (+ the explicit part). */
int sgpu_adjoint_solve(sgpu_ctx* ctx, const double* g, double* psi, double cfl, int max_steps, double tol, sgpu_linsolve* io,
                       int* steps, double* rel_residual);
/* The same with the pseudo-time step ramped like the forward solver's CFL ramp (Solver::solve, src/solver/solver.cpp:211-214):
 * step k uses CFL_k = min(cfl0 * cfl_growth^k, cfl_max); dt and the preconditioner are rebuilt when the CFL changes. */
int sgpu_adjoint_solve_ramp(sgpu_ctx* ctx, const double* g, double* psi, double cfl0, double cfl_growth, double cfl_max, int max_steps,
                            double tol, sgpu_linsolve* io, int* steps, double* rel_residual);
/* The whole ENABLE_ADOLC branch of Solver::step on the device (src/solver/solver.cpp:66-101,154-175):
 * calc_dt(cfl); rhs = residual(q) with solver.order; J = d residual(lhs_order)/dq; solve (-J + 1/dt) dq = rhs;
 * q += under_relaxation * dq (ls_eigen.cpp:66-70).  l2sq as in sgpu_residual. */
int sgpu_implicit_step(sgpu_ctx* ctx, double cfl, double under_relaxation, sgpu_linsolve* io, double* l2sq);

/* Building blocks on DEVICE vectors for a slab-partitioned Krylov solve (one process per GPU; the iteration itself is
 * host logic over torch.distributed, structured_b200/slab.py: dot products are all-reduced, the two ghost rows of the
 * operand are exchanged before every product).  A vector = nv state planes of this slab, sgpu_vec_size doubles, zero
 * outside the owned cells except for halo rows filled by sgpu_vec_halo_unpack. */
int sgpu_vec_size(const sgpu_ctx* ctx, long long* n);
int sgpu_vec_from_rhs(sgpu_ctx* ctx, double* vec_dev);                                  /* what set_rhs receives */
int sgpu_vec_add_to_state(sgpu_ctx* ctx, int which, const double* vec_dev, double omega);  /* q += omega*x, ls_eigen.cpp:66-70 */
int sgpu_vec_halo_pack(sgpu_ctx* ctx, const double* vec_dev, int side, double* buf_dev);   /* layout of sgpu_halo_pack */
int sgpu_vec_halo_unpack(sgpu_ctx* ctx, double* vec_dev, int side, const double* buf_dev);
int sgpu_op_apply(sgpu_ctx* ctx, int matrix, const double* x_dev, double* y_dev);       /* y = A x on the owned rows */
/* Transposed matrices (SGPU_MAT_JT, SGPU_MAT_LHS_T) on a slab: sgpu_op_apply needs NO operand halo, but leaves in the two
 * ghost rows of y what this slab's rows contribute to the neighbour's boundary cells.  pack_ghost copies those rows out
 * (layout of sgpu_halo_pack) and clears them; after the transport the receiver ADDS the buffer to its two boundary rows
 * on that side: the one extra ghost-row exchange of two block rows of SURVEY.md section 8(e). */
int sgpu_vec_halo_pack_ghost(sgpu_ctx* ctx, double* vec_dev, int side, double* buf_dev);
int sgpu_vec_halo_add(sgpu_ctx* ctx, double* vec_dev, int side, const double* buf_dev);
/* device vector <-> GLOBAL host array [nic][njc][nv] (owned rows) */
int sgpu_vec_from_host(sgpu_ctx* ctx, const double* host, double* vec_dev);
int sgpu_vec_to_host(sgpu_ctx* ctx, const double* vec_dev, double* host);
int sgpu_precond_setup(sgpu_ctx* ctx, int matrix, int precond);                         /* slab-local factors */
int sgpu_precond_apply(sgpu_ctx* ctx, int matrix, int precond, const double* r_dev, double* z_dev);
/* Gram-Schmidt building blocks (device pointers throughout; V = cnt basis vectors of sgpu_vec_size doubles, contiguous):
 *   out[0 .. cnt) = w . V_j            -- all-reduce out across the slabs, then
 *   w -= sum_j h[j] V_j, *normsq = |w|^2 of this slab's rows -- all-reduce normsq, then
 *   dst = src / sqrt(*normsq)
 * The kernels of sgpu_linear_solve (deterministic in-kernel reductions); partition the inner products of ls_petsc.cpp's GMRES. */
int sgpu_vec_dots(sgpu_ctx* ctx, const double* w_dev, const double* V_dev, int cnt, double* out_dev);
int sgpu_vec_gs_update(sgpu_ctx* ctx, double* w_dev, const double* V_dev, int cnt, const double* h_dev, double* normsq_dev);
int sgpu_vec_scale_rsqrt(sgpu_ctx* ctx, double* dst_dev, const double* src_dev, const double* normsq_dev);

/* ---- multi-GPU j-slabs ---------------------------------------------------------------------- */
/* Two ghost rows of q per interior slab edge.  side: 0 = low-j neighbour, 1 = high-j neighbour.
 * pack writes this slab's two boundary rows into a contiguous DEVICE buffer of sgpu_halo_count()
 * doubles (layout [nv][2][nic]); unpack reads the neighbour's packed rows into this slab's ghost rows. */
int sgpu_halo_count(const sgpu_ctx* ctx);
int sgpu_halo_pack(sgpu_ctx* ctx, int which, int side, double* dev_buf);
int sgpu_halo_unpack(sgpu_ctx* ctx, int which, int side, const double* dev_buf);
/* verification aid: the two GHOST rows of a side, packed like sgpu_halo_pack (what the last exchange delivered) */
int sgpu_halo_pack_ghost(sgpu_ctx* ctx, int which, int side, double* dev_buf);
/* Peer-memory variant (NVLink P2P): each slab owns, per side, a receive buffer (two slots + a sequence flag).
 * After the neighbour's buffer has been registered -- same process: sgpu_halo_recv_buffer + sgpu_halo_enable_peer +
 * sgpu_halo_set_peer; one process per GPU: sgpu_halo_ipc_handle on the owner, sgpu_halo_open_peer on the writer --
 * sgpu_halo_push packs the two boundary rows with a kernel whose stores go STRAIGHT into the neighbour's memory and
 * then publishes a sequence number; sgpu_halo_pull waits (on the device) for the neighbour's sequence number and
 * unpacks into the ghost rows.  Call push on every slab, then pull: no host synchronisation, no NCCL. */
int sgpu_halo_recv_buffer(sgpu_ctx* ctx, int side, double** dev_ptr);
int sgpu_halo_enable_peer(sgpu_ctx* ctx, int peer_device);
int sgpu_halo_set_peer(sgpu_ctx* ctx, int side, double* peer_recv_buf);
int sgpu_halo_ipc_handle(sgpu_ctx* ctx, int side, void* handle64 /* 64 bytes out */);
int sgpu_halo_open_peer(sgpu_ctx* ctx, int side, const void* handle64);
int sgpu_halo_push(sgpu_ctx* ctx, int which);
int sgpu_halo_pull(sgpu_ctx* ctx, int which);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* number of kernels this library launched on the context since creation */
long long sgpu_launch_count(const sgpu_ctx* ctx);
/* device time (ms, CUDA events on the context's stream) of the most recent residual kernel launches:
 * fills at most n entries, returns how many are available */
int sgpu_kernel_times(sgpu_ctx* ctx, float* ms, int n);
int sgpu_enable_kernel_timing(sgpu_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif
