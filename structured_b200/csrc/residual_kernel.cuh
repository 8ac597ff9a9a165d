// The residual kernel: EulerEquation::calc_residual (src/model/eulerequation.cpp:202-232) as ONE fused
// sm_100a kernel -- primitives, MUSCL/first-order reconstruction, Roe/AUSM flux, Green-Gauss viscous
// flux, SA transport + source, accumulation, /V and the partial sums of the residual norms.
//
// Decomposition ("j-marching strips"): a CTA of RW = 64 threads owns a strip of 64 plane
// columns (60 cells + 2 halo columns each side) and walks a chunk of rows upward; six CTAs are resident per SM
// (168 registers, 37 KB shared memory each).  Measured at 4096^2 SA: RW = 128 x 3 CTAs 1.53 ms, 96 x 4 1.45 ms,
// 64 x 6 1.41 ms -- the same 12 warps per SM, but more independent CTAs de-synchronise the row barriers and the
// fp64-dense / memory-dense phases, which pays for the 3 % more halo columns.  Thread t owns column
// i0-2+t.  Per row it
//   phase 1  converts the prefetched q of row jl+2 to primitives and stores them and the row's metrics into
//            shared-memory rings (every HBM byte is loaded once, coalesced, one iteration ahead of its use);
//            evaluates the MUSCL limiter of its own cell along i ONCE (both face values) and the vertex
//            average of its upper-left vertex ONCE, publishing what the right neighbour needs; evaluates the
//            limiter of cell (i, jl+1) along j once (own column only: no barrier needed; one value used in
//            phase 2, one carried in registers).  Both limiters run branch-free for all four variables;
//   phase 2  evaluates the eta face on top of its cell and the chi face on its left, unconditionally and in one
//            basic block -- each face flux is computed exactly once; the bottom eta flux, the lower vertex
//            average and the left state of the next eta face are carried in registers, the right chi flux
//            comes from thread t+1 via smem;
//   phase 3  accumulates, adds sources, divides by V, stores rhs (coalesced) and accumulates rhs^2.
// Two __syncthreads per row.  Arithmetic per cell is independent of the strip/chunk/slab decomposition, so
// one GPU and N slabs agree bit for bit.
#pragma once
#include "common.cuh"

namespace sg {

#ifndef SG_RW
#define SG_RW 64
#endif
#ifndef SG_RES_MINB
#define SG_RES_MINB 6
#endif
constexpr int RW = SG_RW;               // threads per CTA = plane columns per strip
constexpr int RCELLS = RW - 4;          // cells per strip (threads 2 .. RW-3)

struct ResParams {
    View v; Gas g; Metrics m;
    const double* q; double* rhs;
    const double* wdist; const double* beta;
    double* partial;                    // [items][nv] sums of rhs^2, or nullptr
    double eps_chi, eps_eta, dpdx, dpdy;
    double eps12_chi, epsh_chi, eps12_eta, epsh_eta;      // eps/12, eps/2 of the halved limiter form (muscl_cell2)
    int nstrips, nchunks, rpc;
    int row0, row1;                     // local cell rows [row0, row1) this launch covers (whole slab: 0, njl)
    int strip0;                         // first strip of this launch (column-chunked host pipeline), normally 0
    // fused Runge-Kutta stage update (update_rk4, src/solver/solver.cpp:4-13): udst[cell] = uq[cell] + rhs*udt/udiv written from the
    // epilogue (udst = nullptr: off; udst2: a second destination for the last stage, or nullptr).  udst must not alias q.
    const double* uq = nullptr; const double* udt = nullptr; double* udst = nullptr; double* udst2 = nullptr; double udiv = 1.0, uzinv = 1.0;
};

// vertex-average / dual-cell variable set
enum { VU = 0, VV = 1, VT = 2, VM = 3, VN = 4, VMT = 5, VRN = 6 };
// metric ring planes
enum { MCX = 0, MCY = 1, MEX = 2, MEY = 3, MVOL = 4 };

template <int NV, bool VISC> struct ResCfg {
    static constexpr bool SA = NV > 4;
    static constexpr int NB = VISC ? (SA ? 4 : 2) : (SA ? 1 : 0);      // ring B: T, mu [, nut, mut]  (inviscid SA: nut)
    static constexpr int NVA = VISC ? (SA ? 7 : 4) : 0;
    static constexpr int NFC = NV + (SA ? 3 : 0);
    static constexpr int A_DBL = 4*4*RW;                               // ring A: 4 rows x (rho,u,v,p)
    static constexpr int B_DBL = 3*(NB ? NB : 1)*RW;                   // ring B: 3 rows
    static constexpr int M_DBL = 4*5*RW;                               // metric ring: 4 rows x (ncx,ncy,nex,ney,vol)
    static constexpr int VX_DBL = (NVA ? NVA : 1)*RW;
    static constexpr int FC_DBL = NFC*RW;
    static constexpr int MX_DBL = 4*RW;
    static constexpr int Q_DBL = (NV + (SA ? 2 : 0))*RW;                // cp.async staging: raw q row (+ wall distance, beta)
    static constexpr size_t smem_bytes = sizeof(double)*(size_t)(A_DBL + B_DBL + M_DBL + VX_DBL + FC_DBL + MX_DBL + Q_DBL);
};

struct FaceGeom {
    double nx, ny;                       // face normal (un-normalised, |n| = face length)
    double tx, ty, bx, by, rx, ry, lx, ly;   // doubled dual-cell normals: top, bottom, right, left
    double ivol2;                        // 1 / (doubled dual-cell volume)
};

// Net face flux D = G(viscous) - F(inviscid) for one face.
//   ql/qr   reconstructed (rho,u,v,p) left/right states
//   qt/qb/qrr/qll  the four dual-cell values (top, bottom, right, left) of each vertex-average variable
//                  (src/utils/mesh.cpp:44-53,93-98) for the Green-Gauss gradient and the face average
template <int NV, int FLUX, bool VISC>
__device__ __forceinline__ void face_net_flux(const Gas& g, const FaceGeom& fg, const double* ql, const double* qr,
                                              double nutL, double nutR, const double* qt, const double* qb,
                                              const double* qrr, const double* qll, double* D, double* bars) {
    constexpr bool SA = NV > 4;
    double F[4];
    if (FLUX == SGPU_FLUX_ROE) roe_flux<double>(fg.nx, fg.ny, ql[0], ql[1], ql[2], ql[3], qr[0], qr[1], qr[2], qr[3], F);
    else ausm_flux<double>(fg.nx, fg.ny, ql[0], ql[1], ql[2], ql[3], qr[0], qr[1], qr[2], qr[3], F);
    D[0] = -F[0]; D[1] = -F[1]; D[2] = -F[2]; D[3] = -F[3];
    if (SA) D[4] = -((F[0] >= 0.0) ? F[0]*nutL : F[0]*nutR);      // first-order upwind on the face mass flux
    if (VISC) {
        // Green-Gauss sums and face averages stay UN-NORMALISED: G(k) = 2 V_dual grad q_k (mesh.cpp:83-84,127-128) and
        // S(k) = 4 bar q_k (mesh.cpp:18,29); the factors 1/(2 V_dual) and 1/4 go into the viscosities once per face (every
        // gradient enters the flux multiplied by mu, k or mu_sa), the 1/4 of ubar, vbar into one fma of the work term, and
        // the SA source's face averages leave as 4 bar (phase 3 folds the 1/4 into 1/V).  Same formulas as viscous_flux<>
        // (flux.cpp:12-48); differs by rounding only.  -10 fp64 instructions per face.
        auto Gx = [&](int k) { return fg.tx*qt[k] - fg.bx*qb[k] + fg.rx*qrr[k] - fg.lx*qll[k]; };
        auto Gy = [&](int k) { return fg.ty*qt[k] - fg.by*qb[k] + fg.ry*qrr[k] - fg.ly*qll[k]; };
        auto S4 = [&](int k) { return (qll[k] + qrr[k]) + (qt[k] + qb[k]); };
        const double Su = S4(VU), Sv = S4(VV), Sm = S4(VM);
        const double iv4 = 0.25*fg.ivol2;
        double mu, kk;
        if (SA) { const double Smt = S4(VMT); kk = (Sm*g.cp_over_pr + Smt*g.cp_over_prt)*iv4; mu = (Sm + Smt)*iv4; }
        else { kk = Sm*g.cp_over_pr*iv4; mu = Sm*iv4; }
        const double dudx = Gx(VU), dudy = Gy(VU), dvdx = Gx(VV), dvdy = Gy(VV);
        const double div = dudx + dvdy;
        const double tau_xy = mu*(dudy + dvdx);
        const double tau_xx = mu*(2.0*dudx - K23*div);
        const double tau_yy = mu*(2.0*dvdy - K23*div);
        const double ex = fma(0.25, Su*tau_xx + Sv*tau_xy, kk*Gx(VT));      // ubar tau_xx + vbar tau_xy - q_x
        const double ey = fma(0.25, Su*tau_xy + Sv*tau_yy, kk*Gy(VT));
        D[1] += tau_xx*fg.nx + tau_xy*fg.ny;
        D[2] += tau_xy*fg.nx + tau_yy*fg.ny;
        D[3] += ex*fg.nx + ey*fg.ny;
        if (SA) {
            const double musa = (Sm + S4(VRN))*(iv4*(1.0/SA_SIGMA));
            D[4] += musa*(Gx(VN)*fg.nx + Gy(VN)*fg.ny);
            bars[0] = Su; bars[1] = Sv; bars[2] = S4(VN);                  // 4 x the face averages
        }
    } else if (SA) { bars[0] = bars[1] = bars[2] = 0.0; }
}

// x/d for a divisor d known to the host, correctly rounded like the IEEE division the reference's stage update performs
// (q + rhs*dt/(4.0 - order), src/solver/solver.cpp:4-13) but without its slow path: q0 = RN(x z), z = RN(1/d); the residual
// r = x - q0 d is exact in one fma; RN(q0 + r z) is the correctly rounded quotient (d = 1, 2, 4: z exact, r = 0; d = 3: the
// classical division-by-constant sequence).  The bitwise equality with axpy_dt_div_kernel is what the fused-stage test checks.
__device__ __forceinline__ double div_const(double x, double d, double z) {
    const double q0 = x*z;
    return fma(fma(-d, q0, x), z, q0);
}

template <int NV, int ORDER, int FLUX, bool VISC, bool UPD = false>
__global__ void __launch_bounds__(RW, SG_RES_MINB) residual_kernel(const ResParams prm) {
    using Cfg = ResCfg<NV, VISC>;
    constexpr bool SA = Cfg::SA;
    constexpr int NB = Cfg::NB, NVA = Cfg::NVA;
    constexpr int BT = 0, BM = 1, BN = VISC ? 2 : 0, BMT = 3;   // ring B variable indices
    extern __shared__ double smem[];
    double* sA = smem;                                  // [4][4][RW]   rho,u,v,p      rows jl..jl+2 live
    double* sB = sA + Cfg::A_DBL;                       // [3][NB][RW]  T,mu,...       rows jl..jl+1 live
    double* sM = sB + Cfg::B_DBL;                       // [4][5][RW]   metrics        rows jl..jl+2 live
    double* sVX = sM + Cfg::M_DBL;                      // [NVA][RW]    vertex averages of vertex row jl+1
    double* sFC = sVX + Cfg::VX_DBL;                    // [NFC][RW]    chi net flux (+ SA face averages)
    double* sMX = sFC + Cfg::FC_DBL;                    // [4][RW]      MUSCL value at the cell's high-i face
    double* sQ = sMX + Cfg::MX_DBL;                     // [NV(+2)][RW] cp.async staging of the next raw q row (+ SA fields)

    const View& v = prm.v; const Gas& g = prm.g;
    const int t = threadIdx.x;
    const int strip = prm.strip0 + blockIdx.x % prm.nstrips, chunk = blockIdx.x / prm.nstrips;
    const int i0 = strip*RCELLS;
    const int ra = prm.row0 + chunk*prm.rpc, rb = imin(ra + prm.rpc, prm.row1);
    const int i = i0 - 2 + t;                            // global index of own cell column / own chi face
    const bool face_ok = t >= 2 && t <= RW - 2 && i <= v.nic;
    const bool cell_ok = t >= 2 && t <= RW - 3 && i < v.nic;
    const int c = imin(i + IOFF, v.pitch - 1);           // plane column (clamped: the last strip may overhang the pitch)
    const size_t pl = v.plane;
    const bool col_int = i >= 0 && i <= v.nic - 1;       // own column is an interior cell column

    auto Arow = [&](int jl) { return sA + ((jl + 4) & 3)*4*RW; };
    auto Brow = [&](int jl) { return sB + ((jl + 3) % 3)*(NB ? NB : 1)*RW; };
    auto Mrow = [&](int jl) { return sM + ((jl + 4) & 3)*5*RW; };

    // ---- HBM -> on-chip, one iteration ahead of use.  q (+ wall distance, beta) goes through registers because it
    //      is transformed (primitives) on the way into the ring; the metric planes are copied untransformed,
    //      so they go global -> shared directly with cp.async (no registers, no LDS/STS issue slots).
    auto cp8 = [&](unsigned dst, const double* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src)); };
    // one commit group per row: raw q of row jl -> staging, metrics of row jl -> their ring row, SA fields of row jf
    auto fetch_async = [&](int jl, int jf) {
        const size_t o = v.at(jl + JOFF, c);
        const unsigned dq = (unsigned)__cvta_generic_to_shared(sQ + t);
#pragma unroll
        for (int k = 0; k < NV; k++) cp8(dq + k*RW*8, prm.q + k*pl + o);
        if (SA) { const size_t of = v.at(jf + JOFF, c); cp8(dq + NV*RW*8, prm.wdist + of); cp8(dq + (NV + 1)*RW*8, prm.beta + of); }
        const unsigned dm = (unsigned)__cvta_generic_to_shared(Mrow(jl) + t);
        cp8(dm + MCX*RW*8, prm.m.ncx + o); cp8(dm + MCY*RW*8, prm.m.ncy + o);
        cp8(dm + MEX*RW*8, prm.m.nex + o); cp8(dm + MEY*RW*8, prm.m.ney + o); cp8(dm + MVOL*RW*8, prm.m.vol + o);
        asm volatile("cp.async.commit_group;");
    };
    auto wait_async = [&]() { asm volatile("cp.async.wait_group 0;" ::: "memory"); };
    auto store_prims = [&](int jl, const double* pq, double* prim) {   // FluidModel::primvars + mu loop (eulerequation.cpp:158,183-188)
        double* A = Arow(jl); double* B = Brow(jl);
        double rho, u, vv, p, T;
        cons_to_prim<double>(g, pq[0], pq[1], pq[2], pq[3], rho, u, vv, p, T);
        A[0*RW + t] = rho; A[1*RW + t] = u; A[2*RW + t] = vv; A[3*RW + t] = p;
        prim[0] = rho; prim[1] = u; prim[2] = vv; prim[3] = p;
        double mul = 0.0;
        if (VISC) { B[BT*RW + t] = T; mul = laminar_viscosity<double>(g, T); B[BM*RW + t] = mul; }
        if (SA) {
            const double rn = pq[NV - 1];
            B[BN*RW + t] = rn*rcp_fast(rho);
            if (VISC) {                                  // mu_t = rho nu~ f_v1, f_v1 = X^3/(X^3 + c_v1^3), X = rho nu~/mu: one reciprocal
                const double r3 = rn*rn*rn, m3 = mul*mul*mul;
                B[BMT*RW + t] = rn*(r3*rcp_fast(fma(SA_CV1*SA_CV1*SA_CV1, m3, r3)));
            }
        }
    };
    auto store_row_direct = [&](int jl) {                // prologue: straight from HBM
        double pq[NV];
        const size_t o = v.at(jl + JOFF, c);
#pragma unroll
        for (int k = 0; k < NV; k++) pq[k] = __ldg(prm.q + k*pl + o);
        double prim[4];
        store_prims(jl, pq, prim);
    };
    auto store_row_staged = [&](int jl, double* prim) {  // main loop: from the cp.async staging row (own column only)
        double pq[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) pq[k] = sQ[k*RW + t];
        store_prims(jl, pq, prim);
    };
    // the dual-cell variable set of a cell (row jl, ring column k)
    auto cell_vars = [&](int jl, int k, double* out) {
        const double* A = Arow(jl); const double* B = Brow(jl);
        out[VU] = A[1*RW + k]; out[VV] = A[2*RW + k]; out[VT] = B[BT*RW + k]; out[VM] = B[BM*RW + k];
        if (SA) { out[VN] = B[BN*RW + k]; out[VMT] = B[BMT*RW + k]; out[VRN] = A[0*RW + k]*out[VN]; }
    };
    // vertex (i, vr) = 1/4 of the four cells around it (mesh.cpp:16-17,27-28,49-52,96-97)
    auto vertex_avg = [&](int vr, double* out) {
        if (!VISC) return;
        const int tm = imax(t - 1, 0);
        double a[NVA ? NVA : 1], b[NVA ? NVA : 1], cc_[NVA ? NVA : 1], d[NVA ? NVA : 1];
        cell_vars(vr - 1, tm, a); cell_vars(vr - 1, t, b); cell_vars(vr, tm, cc_); cell_vars(vr, t, d);
#pragma unroll
        for (int n = 0; n < NVA; n++) out[n] = 0.25*(a[n] + b[n] + cc_[n] + d[n]);
    };
    // MUSCL limiter of the own cell along i at row jl: value at its low face (kept) and high face (published)
    auto chi_limiter = [&](int jl, double* to_low) {
        const double* A = Arow(jl);
        const int tm = imax(t - 1, 0), tp = imin(t + 1, RW - 1);
        // all four variables in ONE region, no branch: the limiter runs for every lane (a ghost column's result is discarded by
        // a select), two variables share a reciprocal.  Four separate `if (interior)` regions per limiter kept the scheduler
        // from interleaving the four reciprocal chains (A/B: 1.258 -> 1.176 ms for both limiters)
        double qm_[4], q0_[4], qp_[4], hi_[4], lo_[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { q0_[k] = A[k*RW + t]; qm_[k] = A[k*RW + tm]; qp_[k] = A[k*RW + tp]; hi_[k] = lo_[k] = q0_[k]; }   // ghost / first order: the cell value (reconstruction.cpp:29-35)
        if (ORDER == 2) {
            double h2[4], l2[4];
            muscl_cell2(qm_, q0_, qp_, prm.eps12_chi, prm.epsh_chi, h2, l2); muscl_cell2(qm_ + 2, q0_ + 2, qp_ + 2, prm.eps12_chi, prm.epsh_chi, h2 + 2, l2 + 2);
#pragma unroll
            for (int k = 0; k < 4; k++) { hi_[k] = col_int ? h2[k] : q0_[k]; lo_[k] = col_int ? l2[k] : q0_[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { to_low[k] = lo_[k]; sMX[k*RW + t] = hi_[k]; }
    };
    // MUSCL limiter of cell (i, jl) along j: to_high (left state of face jl+1), to_low (right state of face jl)
    // qp_reg: the primitives of row jl+1 in registers (main loop), or nullptr to read them from the ring
    auto eta_limiter = [&](int jl, double* to_high, double* to_low, const double* qp_reg) {
        const double* Am = Arow(jl - 1); const double* A0 = Arow(jl); const double* Ap = Arow(jl + 1);
        const int gjc = v.j0 + jl;
        const bool row_int = gjc >= 0 && gjc <= v.njc - 1;
        double qm_[4], q0_[4], qp_[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { q0_[k] = A0[k*RW + t]; qm_[k] = Am[k*RW + t]; qp_[k] = qp_reg ? qp_reg[k] : Ap[k*RW + t]; to_high[k] = to_low[k] = q0_[k]; }
        if (ORDER == 2) {
            double h2[4], l2[4];
            muscl_cell2(qm_, q0_, qp_, prm.eps12_eta, prm.epsh_eta, h2, l2); muscl_cell2(qm_ + 2, q0_ + 2, qp_ + 2, prm.eps12_eta, prm.epsh_eta, h2 + 2, l2 + 2);
#pragma unroll
            for (int k = 0; k < 4; k++) { to_high[k] = row_int ? h2[k] : q0_[k]; to_low[k] = row_int ? l2[k] : q0_[k]; }
        }
    };

    // ---- eta face between local cell rows jl and jl+1 (global face row gj = j0+jl+1), column i
    // own = cell_vars(jl, t): loaded once per row by the caller and shared with the chi face
    auto eta_face = [&](int jl, const double* ql, const double* qr, const double* vown, const double* own, double* D, double* bars) {
        const int gj = v.j0 + jl + 1;
        FaceGeom fg;
        const double* Mf = Mrow(jl + 1);
        fg.nx = Mf[MEX*RW + t]; fg.ny = Mf[MEY*RW + t];
        double qt[NVA ? NVA : 1], qrr[NVA ? NVA : 1];
        if (VISC) {                                      // mesh.cpp:99-126 with clamped indices
            const int la = imax(gj - 1, 0) - v.j0, lb = imin(gj, v.njc - 1) - v.j0;     // local rows of the two cells
            const int lT = imin(gj + 1, v.nj - 1) - v.j0, lBo = imax(gj - 1, 0) - v.j0;
            const double* MA = Mrow(la); const double* MB = Mrow(lb); const double* MT = Mrow(lT); const double* MBo = Mrow(lBo);
            fg.tx = fg.nx + MT[MEX*RW + t]; fg.ty = fg.ny + MT[MEY*RW + t];
            fg.bx = fg.nx + MBo[MEX*RW + t]; fg.by = fg.ny + MBo[MEY*RW + t];
            fg.lx = MA[MCX*RW + t] + MB[MCX*RW + t]; fg.ly = MA[MCY*RW + t] + MB[MCY*RW + t];
            fg.rx = MA[MCX*RW + t + 1] + MB[MCX*RW + t + 1]; fg.ry = MA[MCY*RW + t + 1] + MB[MCY*RW + t + 1];
            fg.ivol2 = rcp_fast(MA[MVOL*RW + t] + MB[MVOL*RW + t]);
            cell_vars(jl + 1, t, qt);
#pragma unroll
            for (int n = 0; n < NVA; n++) qrr[n] = sVX[n*RW + t + 1];
        }
        const double nutL = SA ? (VISC ? own[VN] : Brow(jl)[BN*RW + t]) : 0.0, nutR = SA ? (VISC ? qt[VN] : Brow(jl + 1)[BN*RW + t]) : 0.0;
        face_net_flux<NV, FLUX, VISC>(g, fg, ql, qr, nutL, nutR, qt, own, qrr, vown, D, bars);
    };
    // ---- chi face between cells (i-1, jl) and (i, jl)
    auto chi_face = [&](int jl, const double* qr, const double* vtop, const double* vbot, const double* own, double* D, double* bars) {
        double ql[4];
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k] = sMX[k*RW + t - 1];
        FaceGeom fg;
        const double* M0 = Mrow(jl); const double* M1 = Mrow(jl + 1);
        fg.nx = M0[MCX*RW + t]; fg.ny = M0[MCY*RW + t];
        double qll[NVA ? NVA : 1];
        if (VISC) {                                      // mesh.cpp:54-82 with clamped indices
            const int ta = t - 1 + (i == 0 ? 1 : 0), tb = t - (i == v.nic ? 1 : 0);
            const int tR = t + (i + 1 <= v.ni - 1 ? 1 : 0), tL = t - (i - 1 >= 0 ? 1 : 0);
            fg.tx = M1[MEX*RW + ta] + M1[MEX*RW + tb]; fg.ty = M1[MEY*RW + ta] + M1[MEY*RW + tb];
            fg.bx = M0[MEX*RW + ta] + M0[MEX*RW + tb]; fg.by = M0[MEY*RW + ta] + M0[MEY*RW + tb];
            fg.rx = fg.nx + M0[MCX*RW + tR]; fg.ry = fg.ny + M0[MCY*RW + tR];
            fg.lx = fg.nx + M0[MCX*RW + tL]; fg.ly = fg.ny + M0[MCY*RW + tL];
            fg.ivol2 = rcp_fast(M0[MVOL*RW + ta] + M0[MVOL*RW + tb]);
            cell_vars(jl, t - 1, qll);
        }
        const double nutL = SA ? (VISC ? qll[VN] : Brow(jl)[BN*RW + t - 1]) : 0.0, nutR = SA ? (VISC ? own[VN] : Brow(jl)[BN*RW + t]) : 0.0;
        face_net_flux<NV, FLUX, VISC>(g, fg, ql, qr, nutL, nutR, vtop, vbot, own, qll, D, bars);
    };

    // ---- prologue: rows ra-2 .. ra+1 into the rings; limiter of cells ra-1 and ra along j; vertex row ra;
    //      bottom eta face of row ra
    for (int jl = ra - 1; jl <= ra + 1; jl++) {          // metric rows ra-1 .. ra+1 straight into their ring rows
        const size_t o = v.at(jl + JOFF, c);
        double* M = Mrow(jl);
        M[MCX*RW + t] = __ldg(prm.m.ncx + o); M[MCY*RW + t] = __ldg(prm.m.ncy + o);
        M[MEX*RW + t] = __ldg(prm.m.nex + o); M[MEY*RW + t] = __ldg(prm.m.ney + o); M[MVOL*RW + t] = __ldg(prm.m.vol + o);
    }
    for (int jl = ra - 2; jl <= ra + 1; jl++) store_row_direct(jl);
    __syncthreads();
    double vbot[NVA ? NVA : 1], ehi[4];
    double Dbot[NV], bbot[3] = {0, 0, 0};
    {
        double qlo[4], elo[4], dummy[4];
        vertex_avg(ra, vbot);
#pragma unroll
        for (int n = 0; n < NVA; n++) sVX[n*RW + t] = vbot[n];
        eta_limiter(ra - 1, qlo /*to_high of cell ra-1*/, dummy, nullptr);
        eta_limiter(ra, ehi, elo, nullptr);
        __syncthreads();
        double own0[NVA ? NVA : 1];
        if (VISC) cell_vars(ra - 1, t, own0);
        if (cell_ok) eta_face(ra - 1, qlo, elo, vbot, own0, Dbot, bbot);
        else {
#pragma unroll
            for (int k = 0; k < NV; k++) Dbot[k] = 0.0;
        }
    }
    __syncthreads();                                     // all prologue reads of ring rows / sVX done before they are overwritten
    fetch_async(ra + 2, ra);                             // consumed in phase 1 of the first iteration
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;

    for (int jl = ra; jl < rb; jl++) {
        // ---- phase 1
        double cqr[4], vtop[NVA ? NVA : 1];
        wait_async();                                    // own cp.async of row jl+2 (issued one iteration ago) has landed
        double prim2[4];
        store_row_staged(jl + 2, prim2);
        const double wd = SA ? sQ[NV*RW + t] : 1.0, beta = SA ? sQ[(NV + 1)*RW + t] : 1.0;   // SA inputs of row jl
        chi_limiter(jl, cqr);
        vertex_avg(jl + 1, vtop);
#pragma unroll
        for (int n = 0; n < NVA; n++) sVX[n*RW + t] = vtop[n];
        double ehi_next[4], elo[4];
        // limiter of cell (i, jl+1) along j: own column only (ring rows jl, jl+1 were written by this thread, row jl+2 is still
        // in registers), so it needs no barrier and its shared-memory latency overlaps the rest of phase 1 (A/B: 1.286 -> 1.277 ms)
        eta_limiter(jl + 1, ehi_next, elo, prim2);
        __syncthreads();                                 // ring rows jl+2 (and everybody's metric copies) are visible
        // ---- phase 2
        // next row's HBM traffic overlaps this row's flux arithmetic; the metric ring slot is that of row jl-1
        // and the staging row was consumed above: no reader is left after the barrier
        if (jl + 1 < rb) fetch_async(jl + 3, jl + 1);
        if (UPD && jl + 1 < rb) {                          // the stage update's operands of the NEXT row: into L2, so that phase 3's loads find them there
            const size_t on = v.at(jl + 1 + JOFF, c);
#pragma unroll
            for (int k = 0; k < NV; k++) asm volatile("prefetch.global.L2 [%0];" :: "l"(prm.uq + k*pl + on));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(prm.udt + on));
        }
        double Dtop[NV], btop[3] = {0, 0, 0}, Dchi[NV], bchi[3] = {0, 0, 0};
        // both faces unconditionally, in one basic block: the halo lanes of a warp execute the face code anyway, and
        // without the two divergent regions the scheduler interleaves the independent eta / chi chains (A/B: 1.365 ->
        // 1.358 ms).  Halo lanes compute on in-range shared memory and never store.
        double own[NVA ? NVA : 1];
        if (VISC) cell_vars(jl, t, own);
        eta_face(jl, ehi, elo, vtop, own, Dtop, btop);
        chi_face(jl, cqr, vtop, vbot, own, Dchi, bchi);
        if (face_ok) {
#pragma unroll
            for (int k = 0; k < NV; k++) sFC[k*RW + t] = Dchi[k];
            if (SA) { sFC[(NV + 0)*RW + t] = bchi[0]; sFC[(NV + 1)*RW + t] = bchi[1]; sFC[(NV + 2)*RW + t] = bchi[2]; }
        }
        __syncthreads();
        // ---- phase 3
        if (cell_ok) {
            const size_t o = v.at(jl + JOFF, c);
            // operands of the fused stage update: requested here, used ~200 instructions later
            double uq_[NV], udt_ = 0.0;
            if (UPD) {
#pragma unroll
                for (int k = 0; k < NV; k++) uq_[k] = prm.uq[k*pl + o];      // plain loads: the last stage writes these cells of q itself (udst2)
                udt_ = __ldg(prm.udt + o);
            }
            const double* M0 = Mrow(jl); const double* M1 = Mrow(jl + 1);
            const double V = M0[MVOL*RW + t], Vi = rcp_fast(V);
            double res[NV];
#pragma unroll
            for (int k = 0; k < NV; k++) res[k] = (Dtop[k] - Dbot[k]) + (sFC[k*RW + t + 1] - Dchi[k]);
            res[1] += -prm.dpdx*V;                       // calc_source_residual, eulerequation.cpp:22-29
            res[2] += -prm.dpdy*V;
            if (SA) {
                double om = 0.0, dndx = 0.0, dndy = 0.0;
                if (VISC) {                              // Green-Gauss over the cell's own four faces
                    const double cxr = M0[MCX*RW + t + 1], cyr = M0[MCY*RW + t + 1], cxl = M0[MCX*RW + t], cyl = M0[MCY*RW + t];
                    const double ext = M1[MEX*RW + t], eyt = M1[MEY*RW + t], exb = M0[MEX*RW + t], eyb = M0[MEY*RW + t];
                    const double ur = sFC[(NV + 0)*RW + t + 1], vr = sFC[(NV + 1)*RW + t + 1], nr = sFC[(NV + 2)*RW + t + 1];
                    const double Vi4 = 0.25*Vi;             // the faces hand over 4 x their averages
                    const double dvdx = (vr*cxr - bchi[1]*cxl + btop[1]*ext - bbot[1]*exb)*Vi4;
                    const double dudy = (ur*cyr - bchi[0]*cyl + btop[0]*eyt - bbot[0]*eyb)*Vi4;
                    dndx = (nr*cxr - bchi[2]*cxl + btop[2]*ext - bbot[2]*exb)*Vi4;
                    dndy = (nr*cyr - bchi[2]*cyl + btop[2]*eyt - bbot[2]*eyb)*Vi4;
                    om = fabs(dvdx - dudy);
                }
                const double mul = VISC ? Brow(jl)[BM*RW + t] : g.mu_ref;
                // the double-only SA source reuses mu_t from the ring (one reciprocal fewer, division-free sixth root)
                const double src = VISC ? sa_source_mut(Arow(jl)[t], Brow(jl)[BN*RW + t], mul, Brow(jl)[BMT*RW + t], om, dndx, dndy, wd, beta)
                                        : sa_source<double>(Arow(jl)[t], Brow(jl)[BN*RW + t], mul, om, dndx, dndy, wd, beta);
                res[NV - 1] += src*V;
            }
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double rk = res[k]*Vi;             // eulerequation.cpp:224-230
                prm.rhs[k*pl + o] = rk;
                acc[k] += rk*rk;
                if (UPD) {                               // q + rhs*dt/(4.0 - order): the value axpy_dt_div_kernel computes, bit for bit
                    const double qn = uq_[k] + div_const(rk*udt_, prm.udiv, prm.uzinv);
                    prm.udst[k*pl + o] = qn;
                    if (prm.udst2) prm.udst2[k*pl + o] = qn;
                }
            }
#pragma unroll
            for (int k = 0; k < NV; k++) Dbot[k] = Dtop[k];
            if (SA) { bbot[0] = btop[0]; bbot[1] = btop[1]; bbot[2] = btop[2]; }
        }
#pragma unroll
        for (int n = 0; n < NVA; n++) vbot[n] = vtop[n];
#pragma unroll
        for (int k = 0; k < 4; k++) ehi[k] = ehi_next[k];
    }

    // ---- partial sums of rhs^2 for the L2 norms (src/solver/solver.cpp:125-134): warp shuffle + smem
    if (prm.partial) {
        __syncthreads();
        double* red = sFC;                               // reuse: [NV][RW/32]
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = acc[k];
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if ((t & 31) == 0) red[k*(RW/32) + (t >> 5)] = s;
        }
        __syncthreads();
        if (t < NV) {
            double s = 0.0;
            for (int w = 0; w < RW/32; w++) s += red[t*(RW/32) + w];
            prm.partial[(size_t)blockIdx.x*NV + t] = s;
        }
    }
}

} // namespace sg
