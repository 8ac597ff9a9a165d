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
__device__ __forceinline__ void load_cell(const View& v, const Gas& g, const double* __restrict__ q, int r, int c, CellD<NV>& w) {
    const size_t o = v.at(r, c);
    cons_to_prim<double>(g, q[o], q[v.plane + o], q[2*v.plane + o], q[3*v.plane + o], w.r, w.u, w.v, w.p, w.T);
    w.ri = 1.0/w.r;
    const double ke = 0.5*(w.u*w.u + w.v*w.v), iR = 1.0/g.R;
    // T = p/(rho R):  dT = (dp - p/rho drho)/(rho R)
    w.dT[0] = (GM1*ke - w.p*w.ri)*w.ri*iR; w.dT[1] = -GM1*w.u*w.ri*iR; w.dT[2] = -GM1*w.v*w.ri*iR; w.dT[3] = GM1*w.ri*iR;
    w.mu = 0; w.nut = 0; w.mut = 0; w.rn = 0;
    if (VISC) {
        w.mu = laminar_viscosity<double>(g, w.T);
        const double dmudT = (2.0/3.0)*w.mu/w.T;
#pragma unroll
        for (int k = 0; k < 4; k++) w.dmu[k] = dmudT*w.dT[k];
    }
    if (NV > 4) {
        w.rn = q[4*v.plane + o];
        w.nut = w.rn*w.ri;
        const double chi = w.rn/w.mu, c3 = SA_CV1*SA_CV1*SA_CV1, x3 = chi*chi*chi, den = 1.0/(x3 + c3);
        const double fv1 = x3*den, dfv1 = 3.0*chi*chi*c3*den*den;
        w.mut = w.rn*fv1;
        w.dmut[4] = fv1 + chi*dfv1;
#pragma unroll
        for (int k = 0; k < 4; k++) w.dmut[k] = -chi*chi*dfv1*w.dmu[k];
    }
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
    const double* wdist; const double* beta;
    double eps_chi, eps_eta;
    int nslots;
    int* err;
};

// accumulate a block into slot storage; first touch stores, later touches read-modify-write
template <int NV>
struct SlotAcc {
    double* J; size_t stride; size_t cell; unsigned touched;
    __device__ __forceinline__ void add(int slot, const double* blk /*[NV*NV]*/) {
        double* p = J + ((size_t)slot*NV*NV)*stride + cell;
        if (touched & (1u << slot)) {
#pragma unroll
            for (int e = 0; e < NV*NV; e++) p[e*stride] += blk[e];
        } else {
#pragma unroll
            for (int e = 0; e < NV*NV; e++) p[e*stride] = blk[e];
            touched |= (1u << slot);
        }
    }
};

// d(net face flux D = G - F)/dq of every stencil cell of ONE face, scaled by `scale` (= +-1/V), accumulated
// into the row cell's slots.  Generic over chi / eta faces:
//   line cells   LL, L | R, RR     (reconstruction, src/model/reconstruction.cpp)
//   dual cell    D0 = L, D1 = R (direct), vertex averages over {D0, D1, P0, P1} ("plus" side) and
//                {D0, D1, M0, M1} ("minus" side)          (src/utils/mesh.cpp:44-53, 93-98)
template <int NV, int ORDER, int FLUX, bool VISC, int NL>
__device__ __forceinline__ void face_jacobian(const Gas& g, const FaceGeom& fg, double scale, double eps,
                                              const CellD<NV>* LL, const CellD<NV>& L, const CellD<NV>& R, const CellD<NV>* RR,
                                              bool Lint, bool Rint,
                                              const CellD<NV>* P0, const CellD<NV>* P1, const CellD<NV>* M0, const CellD<NV>* M1,
                                              int sLL, int sL, int sR, int sRR, int sP0, int sP1, int sM0, int sM1,
                                              SlotAcc<NV>& acc) {
    constexpr bool SA = NV > 4;
    // ---- 1. reconstruction and its derivative scalars
    double ql[4], qr[4], dl[4][3], dr[4][3];     // dl[k] = d ql_k / d(LL_k, L_k, R_k); dr[k] = d qr_k / d(L_k, R_k, RR_k)
    const double Lw[4] = {L.r, L.u, L.v, L.p}, Rw[4] = {R.r, R.u, R.v, R.p};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ql[k] = Lw[k]; qr[k] = Rw[k];
        dl[k][0] = 0; dl[k][1] = 1; dl[k][2] = 0; dr[k][0] = 0; dr[k][1] = 1; dr[k][2] = 0;
    }
    if (ORDER == 2) {
        typedef Dual<3> D3;
        if (Lint) {
            const double LLw[4] = {LL->r, LL->u, LL->v, LL->p};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                D3 a(LLw[k]), b(Lw[k]), c(Rw[k]), hi, lo; a.d[0] = 1; b.d[1] = 1; c.d[2] = 1;
                muscl_cell<D3>(a, b, c, eps, hi, lo);
                ql[k] = hi.v; dl[k][0] = hi.d[0]; dl[k][1] = hi.d[1]; dl[k][2] = hi.d[2];
            }
        }
        if (Rint) {
            const double RRw[4] = {RR->r, RR->u, RR->v, RR->p};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                D3 a(Lw[k]), b(Rw[k]), c(RRw[k]), hi, lo; a.d[0] = 1; b.d[1] = 1; c.d[2] = 1;
                muscl_cell<D3>(a, b, c, eps, hi, lo);
                qr[k] = lo.v; dr[k][0] = lo.d[0]; dr[k][1] = lo.d[1]; dr[k][2] = lo.d[2];
            }
        }
    }
    // ---- 2. dF/d(ql, qr) by forward-mode passes of NL lanes through the flux function itself
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
    // ---- 3. inviscid blocks of the four line cells: -scale * (FdL diag(dl_s) + FdR diag(dr_s)) dW_s/dq_s
    const bool upL = F0 >= 0.0;
    const double nut_up = SA ? (upL ? L.nut : R.nut) : 0.0;
    auto line_cell = [&](const CellD<NV>& cs, int slot, int il, int ir, bool isL, bool isR) {
        // il / ir: which of dl[k][.] / dr[k][.] belongs to this cell (-1: none)
        double blk[NV*NV];
#pragma unroll
        for (int r = 0; r < NV; r++) {
            double cw[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double fr_l = (r < 4) ? Fd[r][k] : nut_up*Fd[0][k];
                const double fr_r = (r < 4) ? Fd[r][4 + k] : nut_up*Fd[0][4 + k];
                cw[k] = (il >= 0 ? fr_l*dl[k][il] : 0.0) + (ir >= 0 ? fr_r*dr[k][ir] : 0.0);
                cw[k] *= -scale;
            }
            double out[NV];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
            chain_W<NV>(cs, cw, out);
            if (SA && r == 4 && ((isL && upL) || (isR && !upL))) {       // d(F0 nut_up)/d nut_up
                out[0] += -scale*F0*(-cs.nut*cs.ri);
                out[4] += -scale*F0*cs.ri;
            }
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) blk[r*NV + c2] = out[c2];
        }
        acc.add(slot, blk);
    };
    if (ORDER == 2 && Lint) line_cell(*LL, sLL, 0, -1, false, false);
    line_cell(L, sL, 1, 0, true, false);
    line_cell(R, sR, 2, 1, false, true);
    if (ORDER == 2 && Rint) line_cell(*RR, sRR, -1, 2, false, false);

    // ---- 4. viscous blocks of the six dual-cell cells
    if (VISC) {
        const double iv = fg.ivol2;
        // aggregates
        auto agg = [&](double d0, double d1, double p0, double p1, double m0, double m1, double& gx, double& gy, double& bar) {
            const double qp = 0.25*(d0 + d1 + p0 + p1), qm = 0.25*(d0 + d1 + m0 + m1);
            gx = (fg.tx*qp - fg.bx*qm + fg.rx*d1 - fg.lx*d0)*iv;
            gy = (fg.ty*qp - fg.by*qm + fg.ry*d1 - fg.ly*d0)*iv;
            bar = 0.25*(d0 + d1 + qp + qm);
        };
        // NOTE on naming: fg.t* / fg.b* are the normals of the "plus" / "minus" vertex-average sides and
        // fg.r* / fg.l* those of the direct cells D1 / D0 (the caller fills them accordingly for eta faces).
        double ux, uy, ub, vx, vy, vb, Tx, Ty, Tb, mub, dum1, dum2;
        agg(L.u, R.u, P0->u, P1->u, M0->u, M1->u, ux, uy, ub);
        agg(L.v, R.v, P0->v, P1->v, M0->v, M1->v, vx, vy, vb);
        agg(L.T, R.T, P0->T, P1->T, M0->T, M1->T, Tx, Ty, Tb);
        agg(L.mu, R.mu, P0->mu, P1->mu, M0->mu, M1->mu, dum1, dum2, mub);
        double mutb = 0, rnb = 0, nx_ = 0, ny_ = 0, nb_ = 0;
        if (SA) {
            agg(L.mut, R.mut, P0->mut, P1->mut, M0->mut, M1->mut, dum1, dum2, mutb);
            agg(L.rn, R.rn, P0->rn, P1->rn, M0->rn, M1->rn, dum1, dum2, rnb);
            agg(L.nut, R.nut, P0->nut, P1->nut, M0->nut, M1->nut, nx_, ny_, nb_);
        }
        (void)Tb; (void)nb_;
        const double mu = mub + mutb;
        const double kk = SA ? g.cp*(mub/g.pr + mutb/SA_PRT) : mub*g.cp_over_pr;
        const double nxf = fg.nx, nyf = fg.ny;
        const double div = ux + vy;
        const double txx_h = 2.0*ux - (2.0/3.0)*div, tyy_h = 2.0*vy - (2.0/3.0)*div, txy_h = uy + vx;   // tau / mu
        const double txx = mu*txx_h, tyy = mu*tyy_h, txy = mu*txy_h;
        // dG_r / d aggregate, r = 1..3 (flux.cpp:36-45)
        double G_ux[4], G_uy[4], G_vx[4], G_vy[4], G_Tx[4], G_Ty[4], G_ub[4], G_vb[4], G_mu[4], G_k[4];
        const double c43 = 4.0/3.0*mu, c23 = 2.0/3.0*mu;
        G_ux[1] = c43*nxf;  G_uy[1] = mu*nyf; G_vx[1] = mu*nyf; G_vy[1] = -c23*nxf; G_mu[1] = txx_h*nxf + txy_h*nyf;
        G_ux[2] = -c23*nyf; G_uy[2] = mu*nxf; G_vx[2] = mu*nxf; G_vy[2] = c43*nyf;  G_mu[2] = txy_h*nxf + tyy_h*nyf;
        G_ux[3] = nxf*ub*c43 - nyf*vb*c23; G_vy[3] = -nxf*ub*c23 + nyf*vb*c43;
        G_uy[3] = mu*(nxf*vb + nyf*ub); G_vx[3] = G_uy[3];
        G_Tx[3] = kk*nxf; G_Ty[3] = kk*nyf; G_Tx[1] = G_Tx[2] = G_Ty[1] = G_Ty[2] = 0;
        G_ub[3] = nxf*txx + nyf*txy; G_vb[3] = nxf*txy + nyf*tyy; G_ub[1] = G_ub[2] = G_vb[1] = G_vb[2] = 0;
        G_mu[3] = nxf*(ub*txx_h + vb*txy_h) + nyf*(ub*txy_h + vb*tyy_h);
        G_k[3] = nxf*Tx + nyf*Ty; G_k[1] = G_k[2] = 0;
        const double dk_dmu = SA ? g.cp/g.pr : g.cp_over_pr, dk_dmut = g.cp/SA_PRT;
        const double gn = (nx_*nxf + ny_*nyf)*(1.0/SA_SIGMA), musa_s = (mub + rnb)*(1.0/SA_SIGMA);
        auto visc_cell = [&](const CellD<NV>& cs, int slot, double wx, double wy, double wb) {
            double blk[NV*NV];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) blk[c2] = 0.0;                 // mass row: viscous flux[0] is the constant 0 (flux.cpp:42)
#pragma unroll
            for (int r = 1; r < 4; r++) {
                const double cu = G_ux[r]*wx + G_uy[r]*wy + G_ub[r]*wb;
                const double cv = G_vx[r]*wx + G_vy[r]*wy + G_vb[r]*wb;
                const double cT = G_Tx[r]*wx + G_Ty[r]*wy;
                const double cmu = (G_mu[r] + G_k[r]*dk_dmu)*wb;
                const double cmut = SA ? (G_mu[r] + G_k[r]*dk_dmut)*wb : 0.0;
                double out[NV];
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
                chain_Z<NV>(cs, scale*cu, scale*cv, scale*cT, scale*cmu, scale*cmut, 0.0, 0.0, out);
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) blk[r*NV + c2] = out[c2];
            }
            if (SA) {                                                    // G4 = (mub + rnb)/sigma (grad nut . n)
                double out[NV];
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
                chain_Z<NV>(cs, 0.0, 0.0, 0.0, scale*gn*wb, 0.0, scale*musa_s*(wx*nxf + wy*nyf), scale*gn*wb, out);
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) blk[4*NV + c2] = out[c2];
            }
            acc.add(slot, blk);
        };
        const double qx = 0.25*(fg.tx - fg.bx), qy = 0.25*(fg.ty - fg.by);
        visc_cell(R, sR, (qx + fg.rx)*iv, (qy + fg.ry)*iv, 0.375);
        visc_cell(L, sL, (qx - fg.lx)*iv, (qy - fg.ly)*iv, 0.375);
        visc_cell(*P0, sP0, 0.25*fg.tx*iv, 0.25*fg.ty*iv, 0.0625);
        visc_cell(*P1, sP1, 0.25*fg.tx*iv, 0.25*fg.ty*iv, 0.0625);
        visc_cell(*M0, sM0, -0.25*fg.bx*iv, -0.25*fg.by*iv, 0.0625);
        visc_cell(*M1, sM1, -0.25*fg.bx*iv, -0.25*fg.by*iv, 0.0625);
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

template <int NV, int ORDER, int FLUX, bool VISC>
__global__ void __launch_bounds__(128) jacobian_kernel(const JacParams prm) {
    constexpr bool SA = NV > 4;
    constexpr int NL = 2;
    const View& v = prm.v; const Gas& g = prm.g; const Metrics& m = prm.m;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int r = jl + JOFF, c = i + IOFF;
    const int gj = v.j0 + jl;
    const size_t o = v.at(r, c);
    const double V = m.vol[o], Vi = 1.0/V;

    SlotAcc<NV> acc; acc.J = prm.J; acc.stride = v.plane; acc.cell = o; acc.touched = 0u;

    // the cells of the 3x3 block + cross arms, loaded on demand per face (registers are the scarce resource)
    auto cell = [&](int dx, int dy, CellD<NV>& w) { load_cell<NV, VISC>(v, g, prm.q, r + dy, c + dx, w); };

    // ---- chi faces: left (face i, sign -) and right (face i+1, sign +)
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int fi = i + side;                                   // global chi-face index
        const int cf = fi + IOFF;
        const int ox = side - 1;                                   // dx of the face's L cell relative to the row cell
        CellD<NV> LL, L, R, RR, P0, P1, M0, M1;
        cell(ox, 0, L); cell(ox + 1, 0, R);
        const bool Lint = fi - 1 >= 0, Rint = fi <= v.nic - 1;
        if (ORDER == 2) { if (Lint) cell(ox - 1, 0, LL); if (Rint) cell(ox + 2, 0, RR); }
        FaceGeom fg;
        fg.nx = m.ncx[v.at(r, cf)]; fg.ny = m.ncy[v.at(r, cf)];
        if (VISC) {
            cell(ox, 1, P0); cell(ox + 1, 1, P1); cell(ox, -1, M0); cell(ox + 1, -1, M1);
            const int ca = imax(fi - 1, 0) + IOFF, cb = imin(fi, v.nic - 1) + IOFF;
            const int cR = imin(fi + 1, v.ni - 1) + IOFF, cL = imax(fi - 1, 0) + IOFF;
            fg.tx = m.nex[v.at(r + 1, ca)] + m.nex[v.at(r + 1, cb)]; fg.ty = m.ney[v.at(r + 1, ca)] + m.ney[v.at(r + 1, cb)];
            fg.bx = m.nex[v.at(r, ca)] + m.nex[v.at(r, cb)]; fg.by = m.ney[v.at(r, ca)] + m.ney[v.at(r, cb)];
            fg.rx = fg.nx + m.ncx[v.at(r, cR)]; fg.ry = fg.ny + m.ncy[v.at(r, cR)];
            fg.lx = fg.nx + m.ncx[v.at(r, cL)]; fg.ly = fg.ny + m.ncy[v.at(r, cL)];
            fg.ivol2 = 1.0/(m.vol[v.at(r, ca)] + m.vol[v.at(r, cb)]);
        }
        const double scale = (side == 0 ? -1.0 : 1.0)*Vi;
        face_jacobian<NV, ORDER, FLUX, VISC, NL>(g, fg, scale, prm.eps_chi, &LL, L, R, &RR, Lint, Rint, &P0, &P1, &M0, &M1,
            slot_of(ox - 1, 0), slot_of(ox, 0), slot_of(ox + 1, 0), slot_of(ox + 2, 0),
            slot_of(ox, 1), slot_of(ox + 1, 1), slot_of(ox, -1), slot_of(ox + 1, -1), acc);
    }
    // ---- eta faces: bottom (face gj, sign -) and top (face gj+1, sign +)
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int fj = gj + side;                                  // global eta-face index
        const int rf = fj - v.j0 + JOFF;
        const int oy = side - 1;
        CellD<NV> LL, L, R, RR, P0, P1, M0, M1;
        cell(0, oy, L); cell(0, oy + 1, R);
        const bool Lint = fj - 1 >= 0, Rint = fj <= v.njc - 1;
        if (ORDER == 2) { if (Lint) cell(0, oy - 1, LL); if (Rint) cell(0, oy + 2, RR); }
        FaceGeom fg;
        fg.nx = m.nex[v.at(rf, c)]; fg.ny = m.ney[v.at(rf, c)];
        if (VISC) {
            // "plus" side = right vertex average (cells i+1), "minus" side = left (cells i-1); direct: D1 = top, D0 = bottom
            cell(1, oy, P0); cell(1, oy + 1, P1); cell(-1, oy, M0); cell(-1, oy + 1, M1);
            const int a = imax(fj - 1, 0), b = imin(fj, v.njc - 1);
            const int rA = a - v.j0 + JOFF, rB = b - v.j0 + JOFF;
            const int rT = imin(fj + 1, v.nj - 1) - v.j0 + JOFF, rBo = imax(fj - 1, 0) - v.j0 + JOFF;
            // generic roles: t* = plus side (right), b* = minus side (left), r* = D1 (top), l* = D0 (bottom)
            fg.rx = fg.nx + m.nex[v.at(rT, c)]; fg.ry = fg.ny + m.ney[v.at(rT, c)];
            fg.lx = fg.nx + m.nex[v.at(rBo, c)]; fg.ly = fg.ny + m.ney[v.at(rBo, c)];
            fg.bx = m.ncx[v.at(rA, c)] + m.ncx[v.at(rB, c)]; fg.by = m.ncy[v.at(rA, c)] + m.ncy[v.at(rB, c)];
            fg.tx = m.ncx[v.at(rA, c + 1)] + m.ncx[v.at(rB, c + 1)]; fg.ty = m.ncy[v.at(rA, c + 1)] + m.ncy[v.at(rB, c + 1)];
            fg.ivol2 = 1.0/(m.vol[v.at(rA, c)] + m.vol[v.at(rB, c)]);
        }
        const double scale = (side == 0 ? -1.0 : 1.0)*Vi;
        face_jacobian<NV, ORDER, FLUX, VISC, NL>(g, fg, scale, prm.eps_eta, &LL, L, R, &RR, Lint, Rint, &P0, &P1, &M0, &M1,
            slot_of(0, oy - 1), slot_of(0, oy), slot_of(0, oy + 1), slot_of(0, oy + 2),
            slot_of(1, oy), slot_of(1, oy + 1), slot_of(-1, oy), slot_of(-1, oy + 1), acc);
    }

    // ---- SA source: depends on the own cell and, through the cell-centred Green-Gauss gradients of the face
    //      averages, on the 3x3 block
    if (SA) {
        double Wx[3][3], Wy[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) { Wx[a][b] = 0.0; Wy[a][b] = 0.0; }
        const double cxr = m.ncx[v.at(r, c + 1)], cyr = m.ncy[v.at(r, c + 1)], cxl = m.ncx[o], cyl = m.ncy[o];
        const double ext = m.nex[v.at(r + 1, c)], eyt = m.ney[v.at(r + 1, c)], exb = m.nex[o], eyb = m.ney[o];
        auto addw = [&](int dx, int dy, double w, double nxx, double nyy) { Wx[dx + 1][dy + 1] += w*nxx*Vi; Wy[dx + 1][dy + 1] += w*nyy*Vi; };
        // chi face i+1 (+), chi face i (-): bar = 3/8 (two direct cells) + 1/16 (four neighbours)
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
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll 1
            for (int dx = -1; dx <= 1; dx++) {
                CellD<NV> w; cell(dx, dy, w);
                dvdx += Wx[dx + 1][dy + 1]*w.v; dudy += Wy[dx + 1][dy + 1]*w.u;
                dndx += Wx[dx + 1][dy + 1]*w.nut; dndy += Wy[dx + 1][dy + 1]*w.nut;
            }
        const double aa = dvdx - dudy, sgn = aa < 0.0 ? -1.0 : 1.0;
        CellD<NV> w0; cell(0, 0, w0);
        typedef Dual<6> D6;
        D6 a_rho(w0.r), a_nut(w0.nut), a_mu(w0.mu), a_om(fabs(aa)), a_nx(dndx), a_ny(dndy);
        a_rho.d[0] = 1; a_nut.d[1] = 1; a_mu.d[2] = 1; a_om.d[3] = 1; a_nx.d[4] = 1; a_ny.d[5] = 1;
        const D6 S = sa_source<D6>(a_rho, a_nut, a_mu, a_om, a_nx, a_ny, prm.wdist[o], prm.beta[o]);
        // rhs[4] += S*V then /V  ->  d rhs4 = dS
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll 1
            for (int dx = -1; dx <= 1; dx++) {
                CellD<NV> w; cell(dx, dy, w);
                double blk[NV*NV];
#pragma unroll
                for (int e = 0; e < NV*NV; e++) blk[e] = 0.0;
                double out[NV];
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
                const double wx = Wx[dx + 1][dy + 1], wy = Wy[dx + 1][dy + 1];
                chain_Z<NV>(w, -S.d[3]*sgn*wy, S.d[3]*sgn*wx, 0.0, 0.0, 0.0, S.d[4]*wx + S.d[5]*wy, 0.0, out);
                if (dx == 0 && dy == 0) {
                    out[0] += S.d[0];
                    chain_Z<NV>(w, 0.0, 0.0, 0.0, S.d[2], 0.0, S.d[1], 0.0, out);
                }
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) blk[4*NV + c2] = out[c2];
                acc.add(slot_of(dx, dy), blk);
            }
    }

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
    // ---- untouched slots hold zeros
#pragma unroll 1
    for (int s = 0; s < prm.nslots; s++) {
        if (acc.touched & (1u << s)) continue;
        double* p = prm.J + ((size_t)s*NV*NV)*v.plane + o;
#pragma unroll
        for (int e = 0; e < NV*NV; e++) p[e*v.plane] = 0.0;
    }
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
__global__ void jac_count_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, int* __restrict__ counts) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
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
        counts[((size_t)i*v.njl + jl)*NV + r] = cnt;
    }
}

template <int NV>
__global__ void jac_fill_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, const double* __restrict__ J,
                                const long long* __restrict__ offsets, const double* __restrict__ dt, int lhs_transform,
                                unsigned int* __restrict__ rind, unsigned int* __restrict__ cind, double* __restrict__ values) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
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
        long long pos = offsets[((size_t)i*v.njl + jl)*NV + r];
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
