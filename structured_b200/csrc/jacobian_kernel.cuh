// K-JAC: the sparse block Jacobian d rhs / d q of calc_residual(q, lhs = true) that the reference gets from
// ADOL-C's tape + sparse_jac (src/solver/solver.cpp:72-90,156), built directly on the device.
//
// Formulation (DESIGN.md "Jacobian"): one thread per ROW cell.  The row's 13 (second-order lhs) or 9
// (first-order lhs) stencil cells are the static distance-2 colouring of the structured grid -- each is a
// "slot" holding an nv x nv block -- so no run-time colouring or sparsity detection exists.  Inside a
// thread the chain rule is applied by blocks:
//     dD_face/dq_s = dF/d(WL,WR) . d(WL,WR)/dW_s . dW_s/dq_s          (inviscid: forward-mode dual numbers
//                  + dG/d(aggregates) . weights_s . dz_s/dq_s           through the SAME flux functions the
//                                                                       residual kernel runs; viscous: the
//                                                                       Green-Gauss aggregates are linear in
//                                                                       the six cells, src/utils/mesh.cpp:10-131)
// Ghost cells are first treated as independent slots, then folded into the interior cells they are built
// from with the boundary condition's own Jacobian (dual numbers through the BC formulas of
// src/model/bc.cpp), copy-type ghosts (periodic, wake) keep their slot with a remapped column.
#pragma once
#include "common.cuh"
#include "dual.cuh"
#include "aux_kernels.cuh"

namespace sg {

// ---------------------------------------------------------------------------------------------------
// device-resident block-stencil storage: J[slot][r][c][plane cell]
// ---------------------------------------------------------------------------------------------------
struct JacStore {
    double* blocks = nullptr;
    int slots = 0;
    size_t cap = 0;
    bool valid = false;
};
inline void jac_free(JacStore& j) { if (j.blocks) cudaFree(j.blocks); j = JacStore(); }

constexpr int NSLOT_MAX = 13;
// slot -> (dx, dy); the first 9 are the 3x3 block (first-order lhs), 9..12 the radius-2 cross arms
__constant__ int c_slot_dx[NSLOT_MAX] = {0, -1, 1, 0, 0, -1, 1, -1, 1, -2, 2, 0, 0};
__constant__ int c_slot_dy[NSLOT_MAX] = {0, 0, 0, -1, 1, -1, -1, 1, 1, 0, 0, -2, 2};
// (dy+2)*5 + (dx+2) -> slot or -1
__constant__ int c_slot_of[25] = {-1, -1, 11, -1, -1,
                                  -1, 5, 3, 6, -1,
                                  9, 1, 0, 2, 10,
                                  -1, 7, 4, 8, -1,
                                  -1, -1, 12, -1, -1};
__device__ __forceinline__ int slot_of(int dx, int dy) { return c_slot_of[(dy + 2)*5 + (dx + 2)]; }

// ---------------------------------------------------------------------------------------------------
// ghost-cell descriptors: who wrote each ghost cell last (BoundaryContainer::apply order, bc.cpp:430-433)
// ---------------------------------------------------------------------------------------------------
struct GhostDesc {
    int type;              // sgpu_bc_type, or -1 if no BC covers this ghost cell
    int a_ip, a_jp;        // first source cell (padded coordinates)
    int b_ip, b_jp;        // second source cell
    int face;
    double u, v, T;
};
struct GhostTable {
    const GhostDesc* d;    // bottom row [nic+2], top row [nic+2], left col [njc+2], right col [njc+2]
    int nic, njc;
    __device__ __forceinline__ bool is_ghost(int ip, int jp) const { return ip < 1 || ip > nic || jp < 1 || jp > njc; }
    __device__ __forceinline__ const GhostDesc& at(int ip, int jp) const {
        int id;
        if (jp <= 0) id = ip; else if (jp >= njc + 1) id = (nic + 2) + ip;
        else if (ip <= 0) id = 2*(nic + 2) + jp; else id = 2*(nic + 2) + (njc + 2) + jp;
        return d[id];
    }
    // follow copy-type ghosts (periodic, wake) to the cell they duplicate
    __device__ __forceinline__ void resolve(int& ip, int& jp) const {
        for (int it = 0; it < 4; it++) {
            if (!is_ghost(ip, jp)) return;
            if (ip < 0 || ip > nic + 1 || jp < 0 || jp > njc + 1) return;      // beyond the ghost layer: never read
            const GhostDesc& g = at(ip, jp);
            if (g.type == SGPU_BC_PERIODIC || g.type == SGPU_BC_WAKE) { ip = g.a_ip; jp = g.a_jp; }
            else return;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// per-cell quantities and their derivatives with respect to the cell's own conservative variables
// ---------------------------------------------------------------------------------------------------
template <int NV>
struct CellD {
    double r, u, v, p, T, mu, nut, mut, rn, ri;
    double dT[4], dmu[4], dmut[5];
};

template <int NV, bool VISC>
__device__ __forceinline__ void cell_from_q(const Gas& g, double q0, double q1, double q2, double q3, double q4, CellD<NV>& w) {
    cons_to_prim<double>(g, q0, q1, q2, q3, w.r, w.u, w.v, w.p, w.T);
    w.ri = rcp_fast(w.r);
    const double ke = 0.5*(w.u*w.u + w.v*w.v), iR = 1.0/g.R;
    // T = p/(rho R):  dT = (dp - p/rho drho)/(rho R)
    w.dT[0] = (GM1*ke - w.p*w.ri)*w.ri*iR; w.dT[1] = -GM1*w.u*w.ri*iR; w.dT[2] = -GM1*w.v*w.ri*iR; w.dT[3] = GM1*w.ri*iR;
    w.mu = 0; w.nut = 0; w.mut = 0; w.rn = 0;
    if (VISC) {
        // libdevice cbrt (a call, not inlined) keeps the register pressure of these spill-bound kernels down; the value
        // agrees with the residual kernel's s_pow23 to 1 ulp
        { const double cb = cbrt(w.T*g.iT_ref); w.mu = g.mu_ref*cb*cb; }
        const double dmudT = (2.0/3.0)*w.mu*rcp_fast(w.T);
#pragma unroll
        for (int k = 0; k < 4; k++) w.dmu[k] = dmudT*w.dT[k];
    }
    if (NV > 4) {
        w.rn = q4;
        w.nut = w.rn*w.ri;
        const double chi = w.rn*rcp_fast(w.mu), c3 = SA_CV1*SA_CV1*SA_CV1, x3 = chi*chi*chi, den = rcp_fast(x3 + c3);
        const double fv1 = x3*den, dfv1 = 3.0*chi*chi*c3*den*den;
        w.mut = w.rn*fv1;
        w.dmut[4] = fv1 + chi*dfv1;
#pragma unroll
        for (int k = 0; k < 4; k++) w.dmut[k] = -chi*chi*dfv1*w.dmu[k];
    }
}

template <int NV, bool VISC>
__device__ __forceinline__ void load_cell(const View& v, const Gas& g, const double* __restrict__ q, int r, int c, CellD<NV>& w) {
    const size_t o = v.at(r, c);
    cell_from_q<NV, VISC>(g, q[o], q[v.plane + o], q[2*v.plane + o], q[3*v.plane + o], NV > 4 ? q[4*v.plane + o] : 0.0, w);
}
// the same from the face kernel's shared-memory staging: sq[(n*NV + k)*FACE_THREADS] = q_k of stencil cell n (own thread's column)
constexpr int FACE_THREADS = 128;
template <int NV, bool VISC>
__device__ __forceinline__ void load_cell_staged(const Gas& g, const double* sq, int n, CellD<NV>& w) {
    const double* p = sq + (size_t)n*NV*FACE_THREADS;
    cell_from_q<NV, VISC>(g, p[0], p[FACE_THREADS], p[2*FACE_THREADS], p[3*FACE_THREADS], NV > 4 ? p[4*FACE_THREADS] : 0.0, w);
}

// rows of dW/dq (W = rho,u,v,p) and of dz/dq (z = u,v,T,mu,mut,nut,rn) applied to a coefficient vector:
// out[c] += sum_k coefW[k]*dW_k/dq_c + sum_z coefZ[z]*dz_z/dq_c
template <int NV>
__device__ __forceinline__ void chain_W(const CellD<NV>& w, const double* cw /*[4]: rho,u,v,p*/, double* out /*[NV]*/) {
    const double ke = 0.5*(w.u*w.u + w.v*w.v);
    out[0] += cw[0] - (cw[1]*w.u + cw[2]*w.v)*w.ri + cw[3]*GM1*ke;
    out[1] += cw[1]*w.ri - cw[3]*GM1*w.u;
    out[2] += cw[2]*w.ri - cw[3]*GM1*w.v;
    out[3] += cw[3]*GM1;
}
template <int NV>
__device__ __forceinline__ void chain_Z(const CellD<NV>& w, double cu, double cv, double cT, double cmu, double cmut, double cnut, double crn, double* out) {
    out[0] += -(cu*w.u + cv*w.v)*w.ri + cT*w.dT[0] + cmu*w.dmu[0];
    out[1] += cu*w.ri + cT*w.dT[1] + cmu*w.dmu[1];
    out[2] += cv*w.ri + cT*w.dT[2] + cmu*w.dmu[2];
    out[3] += cT*w.dT[3] + cmu*w.dmu[3];
    if (NV > 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] += cmut*w.dmut[k];
        out[0] += -cnut*w.nut*w.ri;
        out[4] += cmut*w.dmut[4] + cnut*w.ri + crn;
    }
}

struct JacParams {
    View v; Gas g; Metrics m; GhostTable gt;
    const double* q; double* J;
    double* Schi; double* Seta;          // per-face scratch, tiled: S[row][tile of 32 faces][cell(8)*nv*nv + e][32]
    int stiles;                          // tiles per row
    const double* wdist; const double* beta;
    double eps_chi, eps_eta;
    int nslots;
    int* err;
};

// Stage 1 of the build: ONE thread per FACE computes d(net face flux D = G - F)/dq of the face's eight stencil
// cells -- every face is differentiated exactly once -- and streams the eight nv x nv blocks to scratch planes.
//   line cells   0:LL 1:L | 2:R 3:RR   (reconstruction, src/model/reconstruction.cpp)
//   dual cell    D0 = L, D1 = R (direct), 4:P0 5:P1 / 6:M0 7:M1 the cells completing the "plus" / "minus"
//                vertex averages                          (src/utils/mesh.cpp:44-53, 93-98)
// fg.t*/b* are the doubled normals of the plus/minus sides, fg.r*/l* those of D1/D0.
struct CellRef { int r, c; };

// Scratch layout: the whole record set of 32 adjacent faces of a row (8 blocks x nv*nv entries) is ONE contiguous
// 8*nv*nv*256-byte run, so a warp of the face kernel writes -- and a warp of the gather kernel reads -- long
// sequential DRAM bursts instead of 200 chunks of 256 bytes that lie a whole plane apart.
constexpr int STILE = 32;
template <int NV>
__device__ __forceinline__ size_t scratch_off(int stiles, int row, int col) {
    return ((size_t)row*stiles + (col >> 5))*(size_t)(8*NV*NV*STILE) + (col & 31);
}

template <int NV, int ORDER, int FLUX, bool VISC, int NL>
__device__ __forceinline__ void face_blocks(const View& v, const Gas& g, const double* __restrict__ q, const FaceGeom& fg, double eps,
                                            const CellRef* cr, bool Lint, bool Rint, double* __restrict__ S, size_t fo, double* sq) {
    constexpr bool SA = NV > 4;
    const size_t stride = v.plane;
    // the eight stencil cells are each used twice below (aggregates, then blocks) inside rolled loops: all 8*nv words
    // are requested at once with cp.async into this thread's shared-memory column (no registers, full memory-level
    // parallelism), so the dependent loads further down cost a shared-memory hit instead of an L2/HBM round trip
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const size_t o = v.at(cr[n].r, cr[n].c);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sq + (size_t)n*NV*FACE_THREADS);
#pragma unroll
        for (int k = 0; k < NV; k++) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + k*FACE_THREADS*8), "l"(q + k*stride + o));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- 1. primitives of the line cells, reconstruction and its derivative scalars
    double W[4][4];                               // [LL,L,R,RR][rho,u,v,p]
#pragma unroll
    for (int n = 0; n < 4; n++) {
        const bool need = (n == 1 || n == 2) || (ORDER == 2 && ((n == 0 && Lint) || (n == 3 && Rint)));
        if (need) {
            const double* p = sq + (size_t)n*NV*FACE_THREADS;
            double T;
            cons_to_prim<double>(g, p[0], p[FACE_THREADS], p[2*FACE_THREADS], p[3*FACE_THREADS], W[n][0], W[n][1], W[n][2], W[n][3], T);
        } else { W[n][0] = W[n][1] = W[n][2] = W[n][3] = 1.0; }
    }
    double ql[4], qr[4], dl[4][3], dr[4][3];      // dl[k] = d ql_k / d(LL_k, L_k, R_k); dr[k] = d qr_k / d(L_k, R_k, RR_k)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ql[k] = W[1][k]; qr[k] = W[2][k];
        dl[k][0] = 0; dl[k][1] = 1; dl[k][2] = 0; dr[k][0] = 0; dr[k][1] = 1; dr[k][2] = 0;
    }
    if (ORDER == 2) {
        typedef Dual<3> D3;
        if (Lint) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                D3 a(W[0][k]), b(W[1][k]), c(W[2][k]), hi, lo; a.d[0] = 1; b.d[1] = 1; c.d[2] = 1;
                muscl_cell<D3>(a, b, c, eps, hi, lo);
                ql[k] = hi.v; dl[k][0] = hi.d[0]; dl[k][1] = hi.d[1]; dl[k][2] = hi.d[2];
            }
        }
        if (Rint) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                D3 a(W[1][k]), b(W[2][k]), c(W[3][k]), hi, lo; a.d[0] = 1; b.d[1] = 1; c.d[2] = 1;
                muscl_cell<D3>(a, b, c, eps, hi, lo);
                qr[k] = lo.v; dr[k][0] = lo.d[0]; dr[k][1] = lo.d[1]; dr[k][2] = lo.d[2];
            }
        }
    }
    // ---- 2. dF/d(ql, qr): forward-mode passes of NL lanes through the flux function the residual kernel runs
    double Fd[4][8], F0 = 0.0;
    typedef Dual<NL> DN;
#pragma unroll
    for (int pass = 0; pass < 8/NL; pass++) {
        DN a[8];
#pragma unroll
        for (int k = 0; k < 4; k++) { a[k] = DN(ql[k]); a[4 + k] = DN(qr[k]); }
#pragma unroll
        for (int l = 0; l < NL; l++) a[pass*NL + l].d[l] = 1.0;
        DN F[4];
        if (FLUX == SGPU_FLUX_ROE) roe_flux<DN>(fg.nx, fg.ny, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], F);
        else ausm_flux<DN>(fg.nx, fg.ny, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], F);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int l = 0; l < NL; l++) Fd[r][pass*NL + l] = F[r].d[l];
        F0 = F[0].v;
    }
    // ---- 3. viscous: the 13 face aggregates (linear in the six cells) and dG/d(aggregate)
    const double iv = VISC ? fg.ivol2 : 0.0;
    double G_ux[4], G_uy[4], G_vx[4], G_vy[4], G_mu[4];           // rows 1..3
    double G_Tx3 = 0, G_Ty3 = 0, G_ub3 = 0, G_vb3 = 0, G_k3 = 0, gn = 0, musa_s = 0;
    const double nxf = fg.nx, nyf = fg.ny;
    double nutL = 0.0, nutR = 0.0;
    if (VISC || SA) {
        double sD0[7], sD1[7], sP[7], sM[7];
#pragma unroll
        for (int n = 0; n < 7; n++) { sD0[n] = sD1[n] = sP[n] = sM[n] = 0.0; }
#pragma unroll 1
        for (int n = 1; n < 8; n++) {
            if (n == 3) continue;
            if (!VISC && n > 2) break;
            CellD<NV> w; load_cell_staged<NV, VISC>(g, sq, n, w);
            const double z[7] = {w.u, w.v, w.T, w.mu, w.mut, w.nut, w.rn};
            double* dst = n == 1 ? sD0 : (n == 2 ? sD1 : (n < 6 ? sP : sM));
#pragma unroll
            for (int k = 0; k < 7; k++) dst[k] += z[k];
        }
        nutL = sD0[5]; nutR = sD1[5];
        if (VISC) {
            auto agg = [&](int k, double& gx, double& gy, double& bar) {
                const double qp = 0.25*(sD0[k] + sD1[k] + sP[k]), qm = 0.25*(sD0[k] + sD1[k] + sM[k]);
                gx = (fg.tx*qp - fg.bx*qm + fg.rx*sD1[k] - fg.lx*sD0[k])*iv;
                gy = (fg.ty*qp - fg.by*qm + fg.ry*sD1[k] - fg.ly*sD0[k])*iv;
                bar = 0.25*(sD0[k] + sD1[k] + qp + qm);
            };
            double ux, uy, ub, vx, vy, vb, Tx, Ty, Tb, mub, d1, d2;
            agg(0, ux, uy, ub); agg(1, vx, vy, vb); agg(2, Tx, Ty, Tb); agg(3, d1, d2, mub);
            double mutb = 0, rnb = 0, nx_ = 0, ny_ = 0, nb_ = 0;
            if (SA) { agg(4, d1, d2, mutb); agg(6, d1, d2, rnb); agg(5, nx_, ny_, nb_); }
            (void)Tb; (void)nb_;
            const double mu = mub + mutb;
            const double kk = SA ? (mub*g.cp_over_pr + mutb*g.cp_over_prt) : mub*g.cp_over_pr;
            const double div = ux + vy;
            const double txx_h = 2.0*ux - (2.0/3.0)*div, tyy_h = 2.0*vy - (2.0/3.0)*div, txy_h = uy + vx;   // tau / mu
            const double txx = mu*txx_h, tyy = mu*tyy_h, txy = mu*txy_h;
            const double c43 = 4.0/3.0*mu, c23 = 2.0/3.0*mu;       // flux.cpp:36-45
            G_ux[1] = c43*nxf;  G_uy[1] = mu*nyf; G_vx[1] = mu*nyf; G_vy[1] = -c23*nxf; G_mu[1] = txx_h*nxf + txy_h*nyf;
            G_ux[2] = -c23*nyf; G_uy[2] = mu*nxf; G_vx[2] = mu*nxf; G_vy[2] = c43*nyf;  G_mu[2] = txy_h*nxf + tyy_h*nyf;
            G_ux[3] = nxf*ub*c43 - nyf*vb*c23; G_vy[3] = -nxf*ub*c23 + nyf*vb*c43;
            G_uy[3] = mu*(nxf*vb + nyf*ub); G_vx[3] = G_uy[3];
            G_Tx3 = kk*nxf; G_Ty3 = kk*nyf;
            G_ub3 = nxf*txx + nyf*txy; G_vb3 = nxf*txy + nyf*tyy;
            G_mu[3] = nxf*(ub*txx_h + vb*txy_h) + nyf*(ub*txy_h + vb*tyy_h);
            G_k3 = nxf*Tx + nyf*Ty;
            gn = (nx_*nxf + ny_*nyf)*(1.0/SA_SIGMA); musa_s = (mub + rnb)*(1.0/SA_SIGMA);
        }
    }
    const double dk_dmu = g.cp_over_pr, dk_dmut = g.cp_over_prt;
    const bool upL = F0 >= 0.0;
    const double nut_up = SA ? (upL ? nutL : nutR) : 0.0;
    // ---- 4. one block per stencil cell: inviscid part (line cells) + viscous part (dual-cell cells)
    const double qx = 0.25*(fg.tx - fg.bx), qy = 0.25*(fg.ty - fg.by);
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const bool line = n < 4, dual = VISC && (n == 1 || n == 2 || n >= 4);
        const bool used = (line && ((n == 1 || n == 2) || (ORDER == 2 && ((n == 0 && Lint) || (n == 3 && Rint))))) || dual;
        double blk[NV*NV];
#pragma unroll
        for (int e = 0; e < NV*NV; e++) blk[e] = 0.0;
        if (used) {
            CellD<NV> cs; load_cell_staged<NV, VISC>(g, sq, n, cs);
            const int il = n == 0 ? 0 : (n == 1 ? 1 : (n == 2 ? 2 : -1));      // which dl[k][.] belongs to this cell
            const int ir = n == 1 ? 0 : (n == 2 ? 1 : (n == 3 ? 2 : -1));      // which dr[k][.]
            double wx = 0, wy = 0, wb = 0;
            if (dual) {
                if (n == 2) { wx = (qx + fg.rx)*iv; wy = (qy + fg.ry)*iv; wb = 0.375; }
                else if (n == 1) { wx = (qx - fg.lx)*iv; wy = (qy - fg.ly)*iv; wb = 0.375; }
                else if (n < 6) { wx = 0.25*fg.tx*iv; wy = 0.25*fg.ty*iv; wb = 0.0625; }
                else { wx = -0.25*fg.bx*iv; wy = -0.25*fg.by*iv; wb = 0.0625; }
            }
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double out[NV];
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
                if (line) {                                                     // D = -F
                    double cw[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const double fl = (r < 4) ? Fd[r < 4 ? r : 0][k] : nut_up*Fd[0][k];
                        const double fr = (r < 4) ? Fd[r < 4 ? r : 0][4 + k] : nut_up*Fd[0][4 + k];
                        double cwk = 0.0;
                        if (il == 0) cwk += fl*dl[k][0]; else if (il == 1) cwk += fl*dl[k][1]; else if (il == 2) cwk += fl*dl[k][2];
                        if (ir == 0) cwk += fr*dr[k][0]; else if (ir == 1) cwk += fr*dr[k][1]; else if (ir == 2) cwk += fr*dr[k][2];
                        cw[k] = -cwk;
                    }
                    chain_W<NV>(cs, cw, out);
                    if (SA && r == 4 && ((n == 1 && upL) || (n == 2 && !upL))) {    // d(F0 nut_up)/d nut_up
                        out[0] += F0*cs.nut*cs.ri;
                        out[4] += -F0*cs.ri;
                    }
                }
                if (dual && r >= 1) {
                    if (r < 4) {
                        const double cu = G_ux[r]*wx + G_uy[r]*wy + (r == 3 ? G_ub3*wb : 0.0);
                        const double cv = G_vx[r]*wx + G_vy[r]*wy + (r == 3 ? G_vb3*wb : 0.0);
                        const double cT = r == 3 ? (G_Tx3*wx + G_Ty3*wy) : 0.0;
                        const double gk = r == 3 ? G_k3 : 0.0;
                        const double cmu = (G_mu[r] + gk*dk_dmu)*wb;
                        const double cmut = SA ? (G_mu[r] + gk*dk_dmut)*wb : 0.0;
                        chain_Z<NV>(cs, cu, cv, cT, cmu, cmut, 0.0, 0.0, out);
                    } else {                                                    // G4 = (mub + rnb)/sigma (grad nut . n)
                        chain_Z<NV>(cs, 0.0, 0.0, 0.0, gn*wb, 0.0, musa_s*(wx*nxf + wy*nyf), gn*wb, out);
                    }
                }
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) blk[r*NV + c2] = out[c2];
            }
        }
        double* p = S + (size_t)n*NV*NV*STILE + fo;
        // streaming stores: the record is read back once by the gather kernel long after it has left L2; evict-first
        // keeps the face kernel's spill lines and q resident instead (A/B: build 37.15 -> 36.45 ms)
#pragma unroll
        for (int e = 0; e < NV*NV; e++) __stcs(p + e*STILE, blk[e]);
    }
}

// DIR = 0: chi faces (i in [0, nic], owned rows); DIR = 1: eta faces (rows j0 .. j1)
template <int NV, int ORDER, int FLUX, bool VISC, int DIR>
__global__ void __launch_bounds__(FACE_THREADS) jac_face_kernel(const JacParams prm) {
    constexpr int NL = 2;
    __shared__ double s_stage[8*NV*FACE_THREADS];                 // q of the eight stencil cells, one column per thread
    double* sq = s_stage + threadIdx.x;
    const View& v = prm.v; const Metrics& m = prm.m;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    FaceGeom fg; CellRef cr[8];
    if (DIR == 0) {
        if (i > v.nic) return;
        const int r = jl + JOFF, cf = i + IOFF;                    // chi face i lies left of cell i: L = cell i-1, R = cell i
        fg.nx = m.ncx[v.at(r, cf)]; fg.ny = m.ncy[v.at(r, cf)];
        if (VISC) {
            const int ca = imax(i - 1, 0) + IOFF, cb = imin(i, v.nic - 1) + IOFF;
            const int cR = imin(i + 1, v.ni - 1) + IOFF, cL = imax(i - 1, 0) + IOFF;
            fg.tx = m.nex[v.at(r + 1, ca)] + m.nex[v.at(r + 1, cb)]; fg.ty = m.ney[v.at(r + 1, ca)] + m.ney[v.at(r + 1, cb)];
            fg.bx = m.nex[v.at(r, ca)] + m.nex[v.at(r, cb)]; fg.by = m.ney[v.at(r, ca)] + m.ney[v.at(r, cb)];
            fg.rx = fg.nx + m.ncx[v.at(r, cR)]; fg.ry = fg.ny + m.ncy[v.at(r, cR)];
            fg.lx = fg.nx + m.ncx[v.at(r, cL)]; fg.ly = fg.ny + m.ncy[v.at(r, cL)];
            fg.ivol2 = rcp_fast(m.vol[v.at(r, ca)] + m.vol[v.at(r, cb)]);
        }
        const int cL0 = cf - 1;                                    // plane column of the L cell
        cr[0] = {r, imax(cL0 - 1, 0)}; cr[1] = {r, cL0}; cr[2] = {r, cL0 + 1}; cr[3] = {r, imin(cL0 + 2, v.pitch - 1)};
        cr[4] = {r + 1, cL0}; cr[5] = {r + 1, cL0 + 1}; cr[6] = {r - 1, cL0}; cr[7] = {r - 1, cL0 + 1};
        const bool Lint = i - 1 >= 0, Rint = i <= v.nic - 1;
        face_blocks<NV, ORDER, FLUX, VISC, NL>(v, prm.g, prm.q, fg, prm.eps_chi, cr, Lint, Rint, prm.Schi, scratch_off<NV>(prm.stiles, r, cf), sq);
    } else {
        if (i >= v.nic || jl > v.njl) return;
        const int fj = v.j0 + jl;                                  // global eta-face index; L = cell row fj-1, R = cell row fj
        const int rf = jl + JOFF, c = i + IOFF;
        fg.nx = m.nex[v.at(rf, c)]; fg.ny = m.ney[v.at(rf, c)];
        if (VISC) {
            const int a = imax(fj - 1, 0), b = imin(fj, v.njc - 1);
            const int rA = a - v.j0 + JOFF, rB = b - v.j0 + JOFF;
            const int rT = imin(fj + 1, v.nj - 1) - v.j0 + JOFF, rBo = imax(fj - 1, 0) - v.j0 + JOFF;
            // generic roles: t* = plus side (right), b* = minus side (left), r* = D1 (top), l* = D0 (bottom)
            fg.rx = fg.nx + m.nex[v.at(rT, c)]; fg.ry = fg.ny + m.ney[v.at(rT, c)];
            fg.lx = fg.nx + m.nex[v.at(rBo, c)]; fg.ly = fg.ny + m.ney[v.at(rBo, c)];
            fg.bx = m.ncx[v.at(rA, c)] + m.ncx[v.at(rB, c)]; fg.by = m.ncy[v.at(rA, c)] + m.ncy[v.at(rB, c)];
            fg.tx = m.ncx[v.at(rA, c + 1)] + m.ncx[v.at(rB, c + 1)]; fg.ty = m.ncy[v.at(rA, c + 1)] + m.ncy[v.at(rB, c + 1)];
            fg.ivol2 = rcp_fast(m.vol[v.at(rA, c)] + m.vol[v.at(rB, c)]);
        }
        const int rL = rf - 1;                                     // plane row of the L cell
        cr[0] = {imax(rL - 1, 0), c}; cr[1] = {rL, c}; cr[2] = {rL + 1, c}; cr[3] = {imin(rL + 2, v.rows - 1), c};
        cr[4] = {rL, c + 1}; cr[5] = {rL + 1, c + 1}; cr[6] = {rL, c - 1}; cr[7] = {rL + 1, c - 1};
        const bool Lint = fj - 1 >= 0, Rint = fj <= v.njc - 1;
        face_blocks<NV, ORDER, FLUX, VISC, NL>(v, prm.g, prm.q, fg, prm.eps_eta, cr, Lint, Rint, prm.Seta, scratch_off<NV>(prm.stiles, rf, c), sq);
    }
}

// Boundary-condition Jacobians dq_ghost/dq_a, dq_ghost/dq_b by dual numbers through the BC formulas (bc.cpp).
template <int NV>
__device__ __forceinline__ void bc_ghost_jacobian(const Gas& g, const GhostDesc& gd, double nx, double ny,
                                                  const double* qa, const double* qb, double* Ma, double* Mb) {
    typedef Dual<NV> D;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        D a[NV], b[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) { a[k] = D(qa[k]); b[k] = D(qb[k]); }
#pragma unroll
        for (int k = 0; k < NV; k++) { if (pass == 0) a[k].d[k] = 1.0; else b[k].d[k] = 1.0; }
        D ar, au, av, ap, aT, br, bu, bv, bp, bT;
        cons_to_prim<D>(g, a[0], a[1], a[2], a[3], ar, au, av, ap, aT);
        cons_to_prim<D>(g, b[0], b[1], b[2], b[3], br, bu, bv, bp, bT);
        D an(0.0), bn(0.0);
        if (NV > 4) { an = a[4]*s_rcp(ar); bn = b[4]*s_rcp(br); }
        D wr(0.0), wu(0.0), wv(0.0), wp(0.0), wn(0.0);
        switch (gd.type) {
        case SGPU_BC_SLIPWALL: {
            const double ds = nx*nx + ny*ny;
            wp = 1.5*ap - 0.5*bp; wr = 1.5*ar - 0.5*br;
            D un = au*nx + av*ny;
            wu = au - un*(2.0*nx/ds); wv = av - un*(2.0*ny/ds);
            wn = 1.5*an - 0.5*bn;
        } break;
        case SGPU_BC_WALL: {
            D T = 1.5*aT - 0.5*bT;
            wr = 1.5*ar - 0.5*br;
            wu = 2.0*gd.u - (1.5*au - 0.5*bu); wv = 2.0*gd.v - (1.5*av - 0.5*bv);
            wp = wr*g.R*T;
            wn = -(1.5*an - 0.5*bn);
        } break;
        case SGPU_BC_ISOTHERMALWALL: {
            wp = 1.5*ap - 0.5*bp;
            wu = 2.0*gd.u - (1.5*au - 0.5*bu); wv = 2.0*gd.v - (1.5*av - 0.5*bv);
            wr = wp*(1.0/(gd.T*g.R));
            wn = -(1.5*an - 0.5*bn);
        } break;
        case SGPU_BC_OUTFLOW: {
            wr = ar; wu = au; wv = av; wp = D(g.p_inf); wn = an;
        } break;
        default: break;
        }
        D qg[NV];
        qg[0] = wr; qg[1] = wr*wu; qg[2] = wr*wv; qg[3] = wp*OGM1 + 0.5*wr*(wu*wu + wv*wv);
        if (NV > 4) qg[4] = wr*wn;
        double* M = pass == 0 ? Ma : Mb;
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) M[r*NV + c2] = qg[r].d[c2];
    }
}

// Stage 2: one thread per ROW cell gathers, for each stencil slot, the blocks of the (at most four) faces whose
// stencil contains that cell: J[slot] = sum_f sign_f/V * S_f[cell index within f].  Pure streaming; every scratch
// block is read by exactly the two cells its face separates.  Then the SA source row, the ghost fold.
//   face 0: chi face i   (-)   face 1: chi face i+1 (+)   face 2: eta face j (-)   face 3: eta face j+1 (+)
__device__ __forceinline__ constexpr int gather_dx(int f, int n) {
    return f < 2 ? ((n < 4 ? n - 2 : ((n & 1) ? 0 : -1)) + (f == 1 ? 1 : 0))
                 : (n < 4 ? 0 : (n < 6 ? 1 : -1));
}
__device__ __forceinline__ constexpr int gather_dy(int f, int n) {
    return f < 2 ? (n < 4 ? 0 : (n < 6 ? 1 : -1))
                 : ((n < 4 ? n - 2 : ((n & 1) ? 0 : -1)) + (f == 3 ? 1 : 0));
}
__device__ __forceinline__ constexpr int gather_slot(int dx, int dy) {
    return (dx == 0 && dy == 0) ? 0 : (dy == 0 ? (dx == -1 ? 1 : dx == 1 ? 2 : dx == -2 ? 9 : 10)
         : dx == 0 ? (dy == -1 ? 3 : dy == 1 ? 4 : dy == -2 ? 11 : 12)
         : (dy == -1 ? (dx == -1 ? 5 : 6) : (dx == -1 ? 7 : 8)));
}

template <int NV, int ORDER, bool VISC>
__global__ void __launch_bounds__(128) jac_gather_kernel(const JacParams prm) {
    constexpr bool SA = NV > 4;
    constexpr int NS = ORDER == 2 ? 13 : 9;
    const View& v = prm.v; const Gas& g = prm.g; const Metrics& m = prm.m;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int r = jl + JOFF, c = i + IOFF;
    const int gj = v.j0 + jl;
    const size_t o = v.at(r, c), pl = v.plane;
    const double V = m.vol[o], Vi = 1.0/V;
    const size_t fo[4] = {scratch_off<NV>(prm.stiles, r, c), scratch_off<NV>(prm.stiles, r, c + 1),
                          scratch_off<NV>(prm.stiles, r, c), scratch_off<NV>(prm.stiles, r + 1, c)};

    // SA source sensitivities (3x3 block, row 4 only)
    double Wx[3][3], Wy[3][3], Sd[6] = {0, 0, 0, 0, 0, 0}, sgn = 1.0;
    if (SA) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) { Wx[a][b] = 0.0; Wy[a][b] = 0.0; }
        const double cxr = m.ncx[v.at(r, c + 1)], cyr = m.ncy[v.at(r, c + 1)], cxl = m.ncx[o], cyl = m.ncy[o];
        const double ext = m.nex[v.at(r + 1, c)], eyt = m.ney[v.at(r + 1, c)], exb = m.nex[o], eyb = m.ney[o];
        auto addw = [&](int dx, int dy, double w, double nxx, double nyy) { Wx[dx + 1][dy + 1] += w*nxx*Vi; Wy[dx + 1][dy + 1] += w*nyy*Vi; };
#pragma unroll
        for (int s = 0; s < 2; s++) {                              // chi faces i (-), i+1 (+): 3/8 direct cells, 1/16 neighbours
            const double sg_ = s ? 1.0 : -1.0, nxx = sg_*(s ? cxr : cxl), nyy = sg_*(s ? cyr : cyl);
            const int x0 = s ? 0 : -1;
            addw(x0, 0, 0.375, nxx, nyy); addw(x0 + 1, 0, 0.375, nxx, nyy);
            addw(x0, 1, 0.0625, nxx, nyy); addw(x0 + 1, 1, 0.0625, nxx, nyy); addw(x0, -1, 0.0625, nxx, nyy); addw(x0 + 1, -1, 0.0625, nxx, nyy);
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const double sg_ = s ? 1.0 : -1.0, nxx = sg_*(s ? ext : exb), nyy = sg_*(s ? eyt : eyb);
            const int y0 = s ? 0 : -1;
            addw(0, y0, 0.375, nxx, nyy); addw(0, y0 + 1, 0.375, nxx, nyy);
            addw(1, y0, 0.0625, nxx, nyy); addw(1, y0 + 1, 0.0625, nxx, nyy); addw(-1, y0, 0.0625, nxx, nyy); addw(-1, y0 + 1, 0.0625, nxx, nyy);
        }
        double dvdx = 0, dudy = 0, dndx = 0, dndy = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const size_t oc = v.at(r + dy, c + dx);
                const double ri = 1.0/prm.q[oc], uu = prm.q[pl + oc]*ri, vv = prm.q[2*pl + oc]*ri, nn = prm.q[4*pl + oc]*ri;
                dvdx += Wx[dx + 1][dy + 1]*vv; dudy += Wy[dx + 1][dy + 1]*uu;
                dndx += Wx[dx + 1][dy + 1]*nn; dndy += Wy[dx + 1][dy + 1]*nn;
            }
        const double aa = dvdx - dudy;
        sgn = aa < 0.0 ? -1.0 : 1.0;
        CellD<NV> w0; load_cell<NV, VISC>(v, g, prm.q, r, c, w0);
        typedef Dual<6> D6;
        D6 a_rho(w0.r), a_nut(w0.nut), a_mu(w0.mu), a_om(fabs(aa)), a_nx(dndx), a_ny(dndy);
        a_rho.d[0] = 1; a_nut.d[1] = 1; a_mu.d[2] = 1; a_om.d[3] = 1; a_nx.d[4] = 1; a_ny.d[5] = 1;
        const D6 S = sa_source<D6>(a_rho, a_nut, a_mu, a_om, a_nx, a_ny, prm.wdist[o], prm.beta[o]);
#pragma unroll
        for (int k = 0; k < 6; k++) Sd[k] = S.d[k];
    }

    // slot order: runs of slots along i first -- the scratch plane (n, e) of chi face i+1 feeds slot A of this cell
    // and slot A+1 of the right neighbour, so walking the i-runs makes the two reads of every chi plane line
    // adjacent in time (L1/L2 hits instead of a second HBM read)
    constexpr int ORD[13] = {9, 1, 0, 2, 10, 7, 4, 8, 5, 3, 6, 11, 12};
#pragma unroll
    for (int si = 0; si < 13; si++) {
        constexpr int dummy = 0; (void)dummy;
        const int s = ORD[si];
        if (s >= NS) continue;
        if (!VISC && s >= 5 && s <= 8) {                           // corners exist through the viscous stencil only
            continue;
        }
        double blk[NV*NV];
#pragma unroll
        for (int e = 0; e < NV*NV; e++) blk[e] = 0.0;
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const double* S = (f < 2 ? prm.Schi : prm.Seta) + fo[f];
            const double sc = (f & 1) ? Vi : -Vi;
#pragma unroll
            for (int n = 0; n < 8; n++) {
                if (gather_slot(gather_dx(f, n), gather_dy(f, n)) != s) continue;
                if (ORDER == 1 && (n == 0 || n == 3)) continue;
                if (!VISC && n >= 4) continue;
#pragma unroll
                for (int e = 0; e < NV*NV; e++) blk[e] += sc*S[(n*NV*NV + e)*STILE];
            }
        }
        if (SA && s < 9) {                                         // rhs[4] += S*V then /V  ->  d rhs4 = dS
            const int dx = (s == 1 || s == 5 || s == 7) ? -1 : ((s == 2 || s == 6 || s == 8) ? 1 : 0);
            const int dy = (s == 3 || s == 5 || s == 6) ? -1 : ((s == 4 || s == 7 || s == 8) ? 1 : 0);
            CellD<NV> w; load_cell<NV, VISC>(v, g, prm.q, r + dy, c + dx, w);
            double out[NV];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
            const double wx = Wx[dx + 1][dy + 1], wy = Wy[dx + 1][dy + 1];
            chain_Z<NV>(w, -Sd[3]*sgn*wy, Sd[3]*sgn*wx, 0.0, 0.0, 0.0, Sd[4]*wx + Sd[5]*wy, 0.0, out);
            if (s == 0) {
                out[0] += Sd[0];
                chain_Z<NV>(w, 0.0, 0.0, 0.0, Sd[2], 0.0, Sd[1], 0.0, out);
            }
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) blk[4*NV + c2] += out[c2];
        }
        double* p = prm.J + ((size_t)s*NV*NV)*pl + o;
#pragma unroll
        for (int e = 0; e < NV*NV; e++) p[e*pl] = blk[e];
    }
    if (!VISC) {                                                   // corner slots of an inviscid stencil hold zeros
        for (int s = 5; s <= 8 && s < NS; s++) {
            double* p = prm.J + ((size_t)s*NV*NV)*pl + o;
            for (int e = 0; e < NV*NV; e++) p[e*pl] = 0.0;
        }
    }

    struct { unsigned touched; double* J; size_t stride, cell;
             __device__ void add(int slot, const double* blk) { double* p = J + ((size_t)slot*NV*NV)*stride + cell; for (int e = 0; e < NV*NV; e++) p[e*stride] += blk[e]; } } acc;
    acc.touched = (1u << NS) - 1u; acc.J = prm.J; acc.stride = pl; acc.cell = o;
    // ---- fold ghost slots into the interior cells they are functions of (corners first, then arms, then edges)
    const int ip0 = i + 1, jp0 = gj + 1;                           // padded coordinates of the row cell
    const bool near_boundary = (i < 2) || (i > v.nic - 3) || (gj < 2) || (gj > v.njc - 3);
    if (near_boundary) {
        const GhostTable& gt = prm.gt;
        const int order_list[12] = {5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 3, 4};
        // three sweeps: a fold may deposit into a ghost slot that was already visited (chains of two BC maps)
#pragma unroll 1
        for (int n3 = 0; n3 < 36; n3++) {
            const int s = order_list[n3 % 12];
            if (s >= prm.nslots || !(acc.touched & (1u << s))) continue;
            int ip = ip0 + c_slot_dx[s], jp = jp0 + c_slot_dy[s];
            if (!gt.is_ghost(ip, jp)) continue;
            if (ip < 0 || ip > gt.nic + 1 || jp < 0 || jp > gt.njc + 1) continue;   // beyond the ghost layer: never read
            gt.resolve(ip, jp);
            if (!gt.is_ghost(ip, jp)) continue;                    // copy-type ghost: keeps its slot, column remapped at export
            const GhostDesc& gd = gt.at(ip, jp);
            // read the raw block, clear the slot
            double B[NV*NV];
            double* p = prm.J + ((size_t)s*NV*NV)*v.plane + o;
#pragma unroll
            for (int e = 0; e < NV*NV; e++) { B[e] = p[e*v.plane]; p[e*v.plane] = 0.0; }
            if (gd.type == SGPU_BC_FREESTREAM || gd.type < 0) continue;   // constants: no dependency
            // source cells in padded coordinates -> local plane coordinates relative to the row cell
            int aip = gd.a_ip, ajp = gd.a_jp, bip = gd.b_ip, bjp = gd.b_jp;
            const bool has_b = gd.type != SGPU_BC_OUTFLOW;
            double qa[NV], qb[NV];
            // values: the planes hold every ghost/interior value at its own padded position
            auto ldq = [&](int pip, int pjp, double* dst) {
                const int rr = pjp - 1 - v.j0 + JOFF, cc2 = pip - 1 + IOFF;
#pragma unroll
                for (int k = 0; k < NV; k++) dst[k] = prm.q[k*v.plane + v.at(rr, cc2)];
            };
            ldq(aip, ajp, qa);
            if (has_b) ldq(bip, bjp, qb); else {
#pragma unroll
                for (int k = 0; k < NV; k++) qb[k] = qa[k];
            }
            double nx = 0.0, ny = 0.0;
            if (gd.type == SGPU_BC_SLIPWALL) {
                const int rf = (gd.face == SGPU_FACE_BOTTOM ? 0 : v.njc) - v.j0 + JOFF, cf = ip - 1 + IOFF;
                nx = m.nex[v.at(rf, cf)]; ny = m.ney[v.at(rf, cf)];
            }
            double Ma[NV*NV], Mb[NV*NV];
            bc_ghost_jacobian<NV>(g, gd, nx, ny, qa, qb, Ma, Mb);
            // targets: find the slot whose resolved identity equals the resolved source cell
#pragma unroll 1
            for (int t = 0; t < (has_b ? 2 : 1); t++) {
                int tip = t ? bip : aip, tjp = t ? bjp : ajp;
                gt.resolve(tip, tjp);
                int ts = -1;
                for (int s2 = 0; s2 < prm.nslots; s2++) {
                    int sip = ip0 + c_slot_dx[s2], sjp = jp0 + c_slot_dy[s2];
                    gt.resolve(sip, sjp);
                    if (sip == tip && sjp == tjp) { ts = s2; break; }
                }
                if (ts < 0) { atomicAdd(prm.err, 1); continue; }
                const double* M = t ? Mb : Ma;
                double blk[NV*NV];
#pragma unroll
                for (int rr = 0; rr < NV; rr++)
#pragma unroll
                    for (int cc2 = 0; cc2 < NV; cc2++) {
                        double sacc = 0.0;
#pragma unroll
                        for (int k = 0; k < NV; k++) sacc += B[rr*NV + k]*M[k*NV + cc2];
                        blk[rr*NV + cc2] = sacc;
                    }
                acc.add(ts, blk);
            }
        }
    }
}

// d rhs4 / d beta per cell (field inversion: the gradient of an objective w.r.t. the correction field is
// psi^T dR/dbeta, and dR/dbeta is diagonal): the SA source is linear in beta, so the derivative is S(beta=1) - S(beta=0)
// evaluated with the same cell-centred Green-Gauss gradients as the residual kernel's epilogue.
template <bool VISC>
__global__ void sa_dbeta_kernel(View v, Gas g, Metrics m, const double* __restrict__ q, const double* __restrict__ wdist, double* __restrict__ out) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int r = jl + JOFF, c = i + IOFF;
    const size_t o = v.at(r, c), pl = v.plane;
    const double V = m.vol[o], Vi = 1.0/V;
    double Wx[3][3], Wy[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { Wx[a][b] = 0.0; Wy[a][b] = 0.0; }
    const double cxr = m.ncx[v.at(r, c + 1)], cyr = m.ncy[v.at(r, c + 1)], cxl = m.ncx[o], cyl = m.ncy[o];
    const double ext = m.nex[v.at(r + 1, c)], eyt = m.ney[v.at(r + 1, c)], exb = m.nex[o], eyb = m.ney[o];
    auto addw = [&](int dx, int dy, double w, double nxx, double nyy) { Wx[dx + 1][dy + 1] += w*nxx*Vi; Wy[dx + 1][dy + 1] += w*nyy*Vi; };
    for (int s = 0; s < 2; s++) {
        const double sg_ = s ? 1.0 : -1.0, nxx = sg_*(s ? cxr : cxl), nyy = sg_*(s ? cyr : cyl);
        const int x0 = s ? 0 : -1;
        addw(x0, 0, 0.375, nxx, nyy); addw(x0 + 1, 0, 0.375, nxx, nyy);
        addw(x0, 1, 0.0625, nxx, nyy); addw(x0 + 1, 1, 0.0625, nxx, nyy); addw(x0, -1, 0.0625, nxx, nyy); addw(x0 + 1, -1, 0.0625, nxx, nyy);
    }
    for (int s = 0; s < 2; s++) {
        const double sg_ = s ? 1.0 : -1.0, nxx = sg_*(s ? ext : exb), nyy = sg_*(s ? eyt : eyb);
        const int y0 = s ? 0 : -1;
        addw(0, y0, 0.375, nxx, nyy); addw(0, y0 + 1, 0.375, nxx, nyy);
        addw(1, y0, 0.0625, nxx, nyy); addw(1, y0 + 1, 0.0625, nxx, nyy); addw(-1, y0, 0.0625, nxx, nyy); addw(-1, y0 + 1, 0.0625, nxx, nyy);
    }
    double dvdx = 0, dudy = 0, dndx = 0, dndy = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const size_t oc = v.at(r + dy, c + dx);
            const double ri = 1.0/q[oc], uu = q[pl + oc]*ri, vv = q[2*pl + oc]*ri, nn = q[4*pl + oc]*ri;
            dvdx += Wx[dx + 1][dy + 1]*vv; dudy += Wy[dx + 1][dy + 1]*uu;
            dndx += Wx[dx + 1][dy + 1]*nn; dndy += Wy[dx + 1][dy + 1]*nn;
        }
    CellD<5> w0; load_cell<5, VISC>(v, g, q, r, c, w0);
    const double om = fabs(dvdx - dudy);
    out[o] = sa_source<double>(w0.r, w0.nut, w0.mu, om, dndx, dndy, wdist[o], 1.0) - sa_source<double>(w0.r, w0.nut, w0.mu, om, dndx, dndy, wdist[o], 0.0);
}

// ---------------------------------------------------------------------------------------------------
// d(surface functional)/dq: the adjoint right-hand side of an objective built from IOManager::write_surface's
// force coefficients (src/utils/io.cpp:219-249), F = a_np Fn_pressure + a_cp Fc_pressure + a_nv Fn_viscous + a_cv Fc_viscous.
// One thread per wall column i: F's partial derivatives with respect to p(i,0), p(i,1) and the four wall-face gradient
// components are constants of the column's geometry; the gradients are linear in (u, v) of the six cells of the j = 0
// dual cell (mesh.cpp:88-128), three of which are ghosts.  Each cell's share is chained to its conservative variables and
// scattered: interior cells directly, ghost cells through the BC that wrote them last (copy-type ghosts to the cell they
// duplicate, functional ghosts through bc_ghost_jacobian -- the same maps the Jacobian fold uses), to depth 3.
// out: nv planes, zeroed by the caller; columns overlap, so contributions are added atomically.
// ---------------------------------------------------------------------------------------------------
template <int NV>
__global__ void surface_grad_kernel(View v, Gas g, Metrics m, GhostTable gt, const double* __restrict__ q,
                                    const double* __restrict__ xv, const double* __restrict__ yv, int i_first, int count,
                                    double a_np, double a_cp, double a_nv, double a_cv, double mu_inf, double qinf,
                                    double* __restrict__ out) {
    const int n = blockIdx.x*blockDim.x + threadIdx.x;
    if (n >= count) return;
    const int i = i_first + n, c = i + IOFF, r0 = JOFF;
    const size_t pl = v.plane;
    const size_t o = v.at(r0, c), o1 = v.at(r0 + 1, c), oc1 = v.at(r0, c + 1);
    const double dx = xv[oc1] - xv[o], dy = yv[oc1] - yv[o];
    const double ivol = 1.0/m.vol[o];
    const double tx = 0.5*(m.nex[o] + m.nex[o1]), ty = 0.5*(m.ney[o] + m.ney[o1]);
    const double bx = m.nex[o], by = m.ney[o], lx = m.ncx[o], ly = m.ncy[o], rx = m.ncx[oc1], ry = m.ncy[oc1];
    // weights of the six cells in d/dx and d/dy of the face gradient: [column i-1, i, i+1][row: 0 = cell row j = 0, 1 = ghost row]
    double wx[3][2], wy[3][2];
    wx[1][0] = (tx + 0.25*(rx - lx))*ivol; wx[1][1] = (-bx + 0.25*(rx - lx))*ivol;
    wx[2][0] = wx[2][1] = 0.25*rx*ivol;     wx[0][0] = wx[0][1] = -0.25*lx*ivol;
    wy[1][0] = (ty + 0.25*(ry - ly))*ivol; wy[1][1] = (-by + 0.25*(ry - ly))*ivol;
    wy[2][0] = wy[2][1] = 0.25*ry*ivol;     wy[0][0] = wy[0][1] = -0.25*ly*ivol;
    // F_i = Cp cp_i + T tau_i + Sxx sxx_i + Syy syy_i  (io.cpp:223-237)
    const double k = mu_inf/qinf;
    const double T = -a_nv*dy + a_cv*dx, Sxx = -a_cv*dy, Syy = a_nv*dx;
    const double Gux = k*((4.0/3.0)*Sxx - (2.0/3.0)*Syy), Gvy = k*((4.0/3.0)*Syy - (2.0/3.0)*Sxx);
    const double Guy = k*T, Gvx = -k*T;
    const double Cp = (-a_np*dx + a_cp*dy)*0.5/qinf;                  // dF/dp(i,0) = dF/dp(i,1)

    struct Item { int ip, jp, depth; double w[NV]; };
    Item st[8]; int sp = 0;
    auto push_cell = [&](int ip, int jp, double du, double dv, double dp) {   // chain (u, v, p) -> q of that cell
        const size_t oc = v.at(jp - 1 + JOFF, ip - 1 + IOFF);
        const double rho = q[oc], ri = 1.0/rho, uu = q[pl + oc]*ri, vv = q[2*pl + oc]*ri;
        Item& it = st[sp++];
        it.ip = ip; it.jp = jp; it.depth = 0;
        const double gp = dp*(GAMMA - 1.0);
        it.w[0] = -(uu*du + vv*dv)*ri + gp*0.5*(uu*uu + vv*vv);
        it.w[1] = du*ri - gp*uu; it.w[2] = dv*ri - gp*vv; it.w[3] = gp;
        if (NV > 4) it.w[4] = 0.0;
    };
    auto drain = [&]() {
        while (sp > 0) {
            const Item it = st[--sp];
            const int ip = it.ip, jp = it.jp;
            if (ip < 0 || ip > gt.nic + 1 || jp < 0 || jp > gt.njc + 1) continue;
            if (!gt.is_ghost(ip, jp)) {
                const size_t oc = v.at(jp - 1 + JOFF, ip - 1 + IOFF);
#pragma unroll
                for (int e = 0; e < NV; e++) if (it.w[e] != 0.0) atomicAdd(out + e*pl + oc, it.w[e]);
                continue;
            }
            if (it.depth >= 3) continue;
            const GhostDesc& gd = gt.at(ip, jp);
            if (gd.type == SGPU_BC_PERIODIC || gd.type == SGPU_BC_WAKE) {
                Item& nx_ = st[sp++]; nx_ = it; nx_.ip = gd.a_ip; nx_.jp = gd.a_jp; nx_.depth = it.depth + 1;
                continue;
            }
            if (gd.type == SGPU_BC_FREESTREAM || gd.type < 0) continue;      // constants
            const bool has_b = gd.type != SGPU_BC_OUTFLOW;
            double qa[NV], qb[NV];
            auto ldq = [&](int pip, int pjp, double* dst) {
                const size_t oc = v.at(pjp - 1 + JOFF, pip - 1 + IOFF);
#pragma unroll
                for (int e = 0; e < NV; e++) dst[e] = q[e*pl + oc];
            };
            ldq(gd.a_ip, gd.a_jp, qa);
            if (has_b) ldq(gd.b_ip, gd.b_jp, qb); else {
#pragma unroll
                for (int e = 0; e < NV; e++) qb[e] = qa[e];
            }
            double nx = 0.0, ny = 0.0;
            if (gd.type == SGPU_BC_SLIPWALL) {
                const size_t of = v.at((gd.face == SGPU_FACE_BOTTOM ? 0 : v.njc) - v.j0 + JOFF, ip - 1 + IOFF);
                nx = m.nex[of]; ny = m.ney[of];
            }
            double Ma[NV*NV], Mb[NV*NV];
            bc_ghost_jacobian<NV>(g, gd, nx, ny, qa, qb, Ma, Mb);
            for (int tt = 0; tt < (has_b ? 2 : 1); tt++) {
                const double* M = tt ? Mb : Ma;
                Item& nx_ = st[sp++];
                nx_.ip = tt ? gd.b_ip : gd.a_ip; nx_.jp = tt ? gd.b_jp : gd.a_jp; nx_.depth = it.depth + 1;
#pragma unroll
                for (int cc = 0; cc < NV; cc++) {
                    double acc = 0.0;
#pragma unroll
                    for (int rr = 0; rr < NV; rr++) acc += it.w[rr]*M[rr*NV + cc];
                    nx_.w[cc] = acc;
                }
            }
        }
    };
    for (int dc = -1; dc <= 1; dc++)
        for (int row = 0; row < 2; row++) {
            const double du = Gux*wx[dc + 1][row] + Guy*wy[dc + 1][row], dv = Gvx*wx[dc + 1][row] + Gvy*wy[dc + 1][row];
            push_cell(i + 1 + dc, row == 0 ? 1 : 0, du, dv, (dc == 0 && row == 0) ? Cp : 0.0);
            drain();
        }
    push_cell(i + 1, 2, 0.0, 0.0, Cp);                                  // p(i, 1)
    drain();
}

// ---------------------------------------------------------------------------------------------------
// Slot -> column cell, shared by the COO export and the matrix-vector products
// ---------------------------------------------------------------------------------------------------
struct SlotCols {
    int col[NSLOT_MAX];        // flat GLOBAL cell index i*njc + j of each slot's column cell, -1 if the slot is empty
};
__device__ __forceinline__ void resolve_slots(const GhostTable& gt, int nslots, int i, int gj, bool viscous, bool order2, SlotCols& sc) {
    for (int s = 0; s < NSLOT_MAX; s++) sc.col[s] = -1;
    for (int s = 0; s < nslots; s++) {
        if (!viscous && s >= 5 && s <= 8) continue;                 // corners exist through the viscous stencil only
        if (!order2 && s >= 9) continue;
        int ip = i + 1 + c_slot_dx[s], jp = gj + 1 + c_slot_dy[s];
        if (ip < 0 || ip > gt.nic + 1 || jp < 0 || jp > gt.njc + 1) continue;
        gt.resolve(ip, jp);
        if (gt.is_ghost(ip, jp)) continue;                          // functional ghost: folded away
        sc.col[s] = (ip - 1)*gt.njc + (jp - 1);
    }
}

// Structural presence of entry (r, c) of a slot (the rules ADOL-C's index-domain propagation yields for this
// operator; DESIGN.md "COO export"): the mass row sees no viscous terms (flux.cpp:42), nothing but the SA
// row's own 3x3 viscous/source terms sees q4, the radius-2 arms enter through (rho,u,v,p) only.
__device__ __forceinline__ bool entry_present(int nv, int s, int r, int c) {
    const bool corner = s >= 5 && s <= 8, arm = s >= 9;
    if (r == 0 && corner) return false;
    if (nv > 4) {
        if (c == 4 && (r == 0 || arm)) return false;
    }
    return true;
}

template <int NV>
__global__ void jac_count_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, int jl0, int nrw, int* __restrict__ counts) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jw = blockIdx.y, jl = jl0 + jw;                      // row window [jl0, jl0 + nrw) of the slab
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    SlotCols sc; resolve_slots(gt, nslots, i, gj, viscous, order2, sc);
    for (int r = 0; r < NV; r++) {
        int cnt = 0;
        for (int s = 0; s < nslots; s++) {
            if (sc.col[s] < 0) continue;
            bool dup = false;                                      // duplicates (tiny periodic grids) merge into the first slot
            for (int s2 = 0; s2 < s; s2++) if (sc.col[s2] == sc.col[s]) dup = true;
            if (dup) continue;
            for (int c2 = 0; c2 < NV; c2++) {
                bool present = false;
                for (int s3 = s; s3 < nslots; s3++) if (sc.col[s3] == sc.col[s] && entry_present(NV, s3, r, c2)) present = true;
                cnt += present ? 1 : 0;
            }
        }
        counts[((size_t)i*nrw + jw)*NV + r] = cnt;
    }
}

template <int NV>
__global__ void jac_fill_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, const double* __restrict__ J,
                                const long long* __restrict__ offsets, const double* __restrict__ dt, int lhs_transform,
                                int jl0, int nrw, unsigned int* __restrict__ rind, unsigned int* __restrict__ cind, double* __restrict__ values) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jw = blockIdx.y, jl = jl0 + jw;
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    SlotCols sc; resolve_slots(gt, nslots, i, gj, viscous, order2, sc);
    // order the distinct column cells ascending (insertion sort over <= 13 entries)
    int ord[NSLOT_MAX], n = 0;
    for (int s = 0; s < nslots; s++) {
        if (sc.col[s] < 0) continue;
        bool dup = false;
        for (int s2 = 0; s2 < s; s2++) if (sc.col[s2] == sc.col[s]) dup = true;
        if (dup) continue;
        int k = n++;
        while (k > 0 && sc.col[ord[k - 1]] > sc.col[s]) { ord[k] = ord[k - 1]; k--; }
        ord[k] = s;
    }
    const unsigned int rowcell = (unsigned int)i*(unsigned int)v.njc + (unsigned int)gj;
    for (int r = 0; r < NV; r++) {
        long long pos = offsets[((size_t)i*nrw + jw)*NV + r];
        const unsigned int row = rowcell*NV + r;
        for (int k = 0; k < n; k++) {
            const int s = ord[k];
            for (int c2 = 0; c2 < NV; c2++) {
                bool present = false; double val = 0.0;
                for (int s3 = s; s3 < nslots; s3++) if (sc.col[s3] == sc.col[s]) {
                    if (entry_present(NV, s3, r, c2)) present = true;
                    val += J[((size_t)s3*NV*NV + r*NV + c2)*v.plane + o];
                }
                if (!present) continue;
                const unsigned int col = (unsigned int)sc.col[s]*NV + c2;
                if (lhs_transform) {                               // src/solver/solver.cpp:162-171
                    val = -val;
                    if (row == col) val += 1.0/dt[o];
                }
                rind[pos] = row; cind[pos] = col; values[pos] = val;
                pos++;
            }
        }
    }
}

// y = J x (transpose = 0) or y += J^T x scattered with atomics (transpose = 1); x, y are state planes
template <int NV>
__global__ void jac_apply_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, const double* __restrict__ J,
                                 const double* __restrict__ x, double* __restrict__ y, int transpose) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    SlotCols sc; resolve_slots(gt, nslots, i, gj, viscous, order2, sc);
    double yr[NV];
    for (int r = 0; r < NV; r++) yr[r] = 0.0;
    for (int s = 0; s < nslots; s++) {
        if (sc.col[s] < 0) continue;
        const int ci = sc.col[s]/v.njc, cj = sc.col[s] - ci*v.njc;
        const int rr = cj - v.j0 + JOFF;
        if (rr < 0 || rr >= v.rows) continue;                      // column outside this slab's planes
        const size_t oc = v.at(rr, ci + IOFF);
        for (int r = 0; r < NV; r++)
            for (int c2 = 0; c2 < NV; c2++) {
                const double a = J[((size_t)s*NV*NV + r*NV + c2)*v.plane + o];
                if (!transpose) yr[r] += a*x[c2*v.plane + oc];
                else atomicAdd(&y[c2*v.plane + oc], a*x[r*v.plane + o]);
            }
    }
    if (!transpose) for (int r = 0; r < NV; r++) y[r*v.plane + o] = yr[r];
}

} // namespace sg
