// The residual kernel: EulerEquation::calc_residual (src/model/eulerequation.cpp:202-232) as ONE fused
// sm_100a kernel -- primitives, MUSCL/first-order reconstruction, Roe/AUSM flux, Green-Gauss viscous
// flux, SA transport + source, accumulation, /V and the partial sums of the residual norms.
//
// Decomposition ("j-marching strips"): a CTA of RW threads owns a strip of RW-1 cell
// columns and walks a chunk of rows upward.  Thread t owns column i0+t: it converts that column's q to
// primitives once per row (shared-memory ring of 5 rows), computes the chi face on the LEFT of its cell
// and the eta face on TOP of it, keeps the bottom eta flux of the next row in registers, and gets its
// right chi flux from thread t+1 through shared memory.  Every face flux is therefore evaluated exactly
// once (plus one face column per strip seam and one face row per chunk seam), and q is read from HBM
// once (plus 3 halo columns per strip and 4 halo rows per chunk).
#pragma once
#include "common.cuh"

namespace sg {

constexpr int RW = 128;                 // threads per CTA = chi faces per strip; cells per strip = RW-1
constexpr int PW = RW + 3;              // primitive columns held per row: i0-2 .. i0+RW

struct ResParams {
    View v; Gas g; Metrics m;
    const double* q; double* rhs;
    const double* wdist; const double* beta;
    double* partial;                    // [items][nv] sums of rhs^2, or nullptr
    double eps_chi, eps_eta, dpdx, dpdy;
    int nstrips, nchunks, rpc;
};

// primitive ring variables / vertex-average variables
enum { PR = 0, PU = 1, PV = 2, PP = 3, PT = 4, PM = 5, PN = 6, PMT = 7, PRN = 8 };
enum { VU = 0, VV = 1, VT = 2, VM = 3, VN = 4, VMT = 5, VRN = 6 };

template <int NV, bool VISC> struct ResCfg {
    static constexpr bool SA = NV > 4;
    static constexpr int NPV = VISC ? (SA ? 9 : 6) : (SA ? 7 : 4);
    static constexpr int NVA = VISC ? (SA ? 7 : 4) : 0;
    static constexpr int NFC = NV + (SA ? 3 : 0);
    static constexpr size_t smem_bytes = sizeof(double)*((size_t)5*NPV*PW + (size_t)2*(NVA ? NVA : 1)*RW + (size_t)NFC*RW);
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
        auto gradx = [&](int k) { return (fg.tx*qt[k] - fg.bx*qb[k] + fg.rx*qrr[k] - fg.lx*qll[k])*fg.ivol2; };   // mesh.cpp:83,127
        auto grady = [&](int k) { return (fg.ty*qt[k] - fg.by*qb[k] + fg.ry*qrr[k] - fg.ly*qll[k])*fg.ivol2; };   // mesh.cpp:84,128
        auto bar = [&](int k) { return 0.25*(qll[k] + qrr[k] + qt[k] + qb[k]); };                                 // mesh.cpp:18,29
        const double ubar = bar(VU), vbar = bar(VV);
        double mu = bar(VM), kk;
        if (SA) { const double mut = bar(VMT); kk = g.cp*(mu/g.pr + mut/SA_PRT); mu = mu + mut; }
        else kk = mu*g.cp_over_pr;
        double G[4];
        viscous_flux<double>(fg.nx, fg.ny, gradx(VU), grady(VU), gradx(VV), grady(VV), gradx(VT), grady(VT), ubar, vbar, mu, kk, G);
        D[1] += G[1]; D[2] += G[2]; D[3] += G[3];
        if (SA) {
            const double musa = bar(VM) + bar(VRN);
            D[4] += musa*(1.0/SA_SIGMA)*(gradx(VN)*fg.nx + grady(VN)*fg.ny);
            bars[0] = ubar; bars[1] = vbar; bars[2] = bar(VN);
        }
    } else if (SA) { bars[0] = bars[1] = bars[2] = 0.0; }
}

template <int NV, int ORDER, int FLUX, bool VISC>
__global__ void __launch_bounds__(RW) residual_kernel(const ResParams prm) {
    using Cfg = ResCfg<NV, VISC>;
    constexpr bool SA = Cfg::SA;
    constexpr int NPV = Cfg::NPV, NVA = Cfg::NVA, NFC = Cfg::NFC;
    extern __shared__ double smem[];
    double* sP = smem;                                  // [5][NPV][PW]
    double* sVA = sP + 5*NPV*PW;                        // [2][NVA][RW]
    double* sFC = sVA + 2*(NVA ? NVA : 1)*RW;           // [NFC][RW]

    const View& v = prm.v; const Gas& g = prm.g; const Metrics& m = prm.m;
    const int t = threadIdx.x;
    const int strip = blockIdx.x % prm.nstrips, chunk = blockIdx.x / prm.nstrips;
    const int i0 = strip*(RW - 1);
    const int ra = chunk*prm.rpc, rb = imin(ra + prm.rpc, v.njl);
    const int i = i0 + t;                                // global index of own cell column / own chi face
    const int cc = t + 2;                                // own column inside the primitive ring
    const bool face_ok = i <= v.nic;
    const bool cell_ok = (t < RW - 1) && (i < v.nic);
    const int c = i + IOFF;                              // plane column of own cell / face
    const size_t pl = v.plane;

    auto slot = [&](int jl) { return sP + ((jl + 10) % 5)*NPV*PW; };
    auto vaslot = [&](int vr) { return sVA + ((vr + 2) & 1)*NVA*RW; };

    // ---- R1: q row -> primitive ring (FluidModel::primvars + mu loop, eulerequation.cpp:158,183-188)
    auto load_row = [&](int jl) {
        double* S = slot(jl);
        const int r = jl + JOFF;
        for (int k = t; k < PW; k += RW) {
            const int col = i0 + k;                      // plane column = (i0 - 2 + k) + IOFF
            if (col < v.pitch) {
                const size_t o = v.at(r, col);
                double rho, u, vv, p, T;
                cons_to_prim<double>(g, prm.q[o], prm.q[pl + o], prm.q[2*pl + o], prm.q[3*pl + o], rho, u, vv, p, T);
                S[PR*PW + k] = rho; S[PU*PW + k] = u; S[PV*PW + k] = vv; S[PP*PW + k] = p;
                double mul = 0.0;
                if (VISC) { S[PT*PW + k] = T; mul = laminar_viscosity<double>(g, T); S[PM*PW + k] = mul; }
                if (SA) {
                    const double rn = prm.q[4*pl + o];
                    const double nut = rn/rho;
                    S[(VISC ? PN : 4)*PW + k] = nut;
                    if (VISC) { S[PMT*PW + k] = rn*sa_fv1<double>(rn/mul); S[PRN*PW + k] = rn; }
                }
            }
        }
    };
    constexpr int PNI = VISC ? PN : 4;                   // ring index of nu~

    // ---- R2: vertex averages of row vr: vertex (i, vr) = 1/4 of the four cells around it
    auto vertex_row = [&](int vr) {
        if (!VISC) return;
        const double* A = slot(vr - 1); const double* B = slot(vr);
        double* V = vaslot(vr);
        auto avg = [&](int pv) { return 0.25*(A[pv*PW + cc - 1] + A[pv*PW + cc] + B[pv*PW + cc - 1] + B[pv*PW + cc]); };
        V[VU*RW + t] = avg(PU); V[VV*RW + t] = avg(PV); V[VT*RW + t] = avg(PT); V[VM*RW + t] = avg(PM);
        if (SA) { V[VN*RW + t] = avg(PN); V[VMT*RW + t] = avg(PMT); V[VRN*RW + t] = avg(PRN); }
    };
    // gather the vertex-average variable set of a cell / of a vertex into a small register array
    auto cell_vars = [&](const double* S, int k, double* out) {
        out[VU] = S[PU*PW + k]; out[VV] = S[PV*PW + k]; out[VT] = S[PT*PW + k]; out[VM] = S[PM*PW + k];
        if (SA) { out[VN] = S[PN*PW + k]; out[VMT] = S[PMT*PW + k]; out[VRN] = S[PRN*PW + k]; }
    };
    auto vert_vars = [&](const double* V, int k, double* out) {
#pragma unroll
        for (int n = 0; n < NVA; n++) out[n] = V[n*RW + k];
    };

    // ---- eta face between local cell rows jl (left state) and jl+1 (right state), column i
    auto eta_face = [&](int jl, double* D, double* bars) {
        const double* LL = slot(jl - 1); const double* L = slot(jl); const double* R = slot(jl + 1); const double* RR = slot(jl + 2);
        const int gj = v.j0 + jl + 1;                    // global eta-face index (face row gj lies below cell row gj)
        double ql[4], qr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { ql[k] = L[k*PW + cc]; qr[k] = R[k*PW + cc]; }
        if (ORDER == 2) {                                // reconstruction.cpp:133-150; ghost side stays first order
            const bool Lint = gj - 1 >= 0, Rint = gj <= v.njc - 1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double hi, lo;
                if (Lint) { muscl_cell<double>(LL[k*PW + cc], L[k*PW + cc], R[k*PW + cc], prm.eps_eta, hi, lo); ql[k] = hi; }
                if (Rint) { muscl_cell<double>(L[k*PW + cc], R[k*PW + cc], RR[k*PW + cc], prm.eps_eta, hi, lo); qr[k] = lo; }
            }
        }
        FaceGeom fg;
        const int rf = gj - v.j0 + JOFF;
        fg.nx = m.nex[v.at(rf, c)]; fg.ny = m.ney[v.at(rf, c)];
        double qt[NVA ? NVA : 1], qb[NVA ? NVA : 1], qrr[NVA ? NVA : 1], qll[NVA ? NVA : 1];
        if (VISC) {                                      // mesh.cpp:99-126 with clamped indices
            const int a = imax(gj - 1, 0), b = imin(gj, v.njc - 1);
            const int rA = a - v.j0 + JOFF, rB = b - v.j0 + JOFF;
            const int rT = imin(gj + 1, v.nj - 1) - v.j0 + JOFF, rBo = imax(gj - 1, 0) - v.j0 + JOFF;
            fg.tx = fg.nx + m.nex[v.at(rT, c)]; fg.ty = fg.ny + m.ney[v.at(rT, c)];
            fg.bx = fg.nx + m.nex[v.at(rBo, c)]; fg.by = fg.ny + m.ney[v.at(rBo, c)];
            fg.lx = m.ncx[v.at(rA, c)] + m.ncx[v.at(rB, c)]; fg.ly = m.ncy[v.at(rA, c)] + m.ncy[v.at(rB, c)];
            fg.rx = m.ncx[v.at(rA, c + 1)] + m.ncx[v.at(rB, c + 1)]; fg.ry = m.ncy[v.at(rA, c + 1)] + m.ncy[v.at(rB, c + 1)];
            fg.ivol2 = 1.0/(m.vol[v.at(rA, c)] + m.vol[v.at(rB, c)]);
            const double* V = vaslot(jl + 1);
            cell_vars(R, cc, qt); cell_vars(L, cc, qb);
            vert_vars(V, t + 1, qrr); vert_vars(V, t, qll);
        }
        const double nutL = SA ? L[PNI*PW + cc] : 0.0, nutR = SA ? R[PNI*PW + cc] : 0.0;
        face_net_flux<NV, FLUX, VISC>(g, fg, ql, qr, nutL, nutR, qt, qb, qrr, qll, D, bars);
    };

    // ---- chi face between cells (i-1, jl) and (i, jl)
    auto chi_face = [&](int jl, double* D, double* bars) {
        const double* S = slot(jl);
        const int gj = v.j0 + jl;
        double ql[4], qr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { ql[k] = S[k*PW + cc - 1]; qr[k] = S[k*PW + cc]; }
        if (ORDER == 2) {                                // reconstruction.cpp:94-111
            const bool Lint = i - 1 >= 0, Rint = i <= v.nic - 1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double hi, lo;
                if (Lint) { muscl_cell<double>(S[k*PW + cc - 2], S[k*PW + cc - 1], S[k*PW + cc], prm.eps_chi, hi, lo); ql[k] = hi; }
                if (Rint) { muscl_cell<double>(S[k*PW + cc - 1], S[k*PW + cc], S[k*PW + cc + 1], prm.eps_chi, hi, lo); qr[k] = lo; }
            }
        }
        FaceGeom fg;
        const int r = jl + JOFF;
        fg.nx = m.ncx[v.at(r, c)]; fg.ny = m.ncy[v.at(r, c)];
        double qt[NVA ? NVA : 1], qb[NVA ? NVA : 1], qrr[NVA ? NVA : 1], qll[NVA ? NVA : 1];
        if (VISC) {                                      // mesh.cpp:54-82 with clamped indices
            const int ca = imax(i - 1, 0) + IOFF, cb = imin(i, v.nic - 1) + IOFF;
            const int cR = imin(i + 1, v.ni - 1) + IOFF, cL = imax(i - 1, 0) + IOFF;
            fg.tx = m.nex[v.at(r + 1, ca)] + m.nex[v.at(r + 1, cb)]; fg.ty = m.ney[v.at(r + 1, ca)] + m.ney[v.at(r + 1, cb)];
            fg.bx = m.nex[v.at(r, ca)] + m.nex[v.at(r, cb)]; fg.by = m.ney[v.at(r, ca)] + m.ney[v.at(r, cb)];
            fg.rx = fg.nx + m.ncx[v.at(r, cR)]; fg.ry = fg.ny + m.ncy[v.at(r, cR)];
            fg.lx = fg.nx + m.ncx[v.at(r, cL)]; fg.ly = fg.ny + m.ncy[v.at(r, cL)];
            fg.ivol2 = 1.0/(m.vol[v.at(r, ca)] + m.vol[v.at(r, cb)]);
            vert_vars(vaslot(jl + 1), t, qt); vert_vars(vaslot(jl), t, qb);
            cell_vars(S, cc, qrr); cell_vars(S, cc - 1, qll);
        }
        (void)gj;
        const double nutL = SA ? S[PNI*PW + cc - 1] : 0.0, nutR = SA ? S[PNI*PW + cc] : 0.0;
        face_net_flux<NV, FLUX, VISC>(g, fg, ql, qr, nutL, nutR, qt, qb, qrr, qll, D, bars);
    };

    // ---- prologue: rows ra-2 .. ra+1, vertex row ra, bottom eta face of row ra
    load_row(ra - 2); load_row(ra - 1); load_row(ra); load_row(ra + 1);
    __syncthreads();
    vertex_row(ra);
    __syncthreads();
    double Dbot[NV], bbot[3] = {0, 0, 0};
    if (cell_ok) eta_face(ra - 1, Dbot, bbot);
    else {
#pragma unroll
        for (int k = 0; k < NV; k++) Dbot[k] = 0.0;
    }
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;

    for (int jl = ra; jl < rb; jl++) {
        load_row(jl + 2);
        vertex_row(jl + 1);
        __syncthreads();
        double Dtop[NV], btop[3] = {0, 0, 0}, Dchi[NV], bchi[3] = {0, 0, 0};
        if (cell_ok) eta_face(jl, Dtop, btop);
        if (face_ok) {
            chi_face(jl, Dchi, bchi);
#pragma unroll
            for (int k = 0; k < NV; k++) sFC[k*RW + t] = Dchi[k];
            if (SA) { sFC[(NV + 0)*RW + t] = bchi[0]; sFC[(NV + 1)*RW + t] = bchi[1]; sFC[(NV + 2)*RW + t] = bchi[2]; }
        }
        __syncthreads();
        if (cell_ok) {
            const int r = jl + JOFF;
            const size_t o = v.at(r, c);
            const double V = m.vol[o], Vi = 1.0/V;
            double res[NV];
#pragma unroll
            for (int k = 0; k < NV; k++) res[k] = (Dtop[k] - Dbot[k]) + (sFC[k*RW + t + 1] - Dchi[k]);
            res[1] += -prm.dpdx*V;                       // calc_source_residual, eulerequation.cpp:22-29
            res[2] += -prm.dpdy*V;
            if (SA) {
                double om = 0.0, dndx = 0.0, dndy = 0.0;
                if (VISC) {                              // Green-Gauss over the cell's own four faces
                    const double cxr = m.ncx[v.at(r, c + 1)], cyr = m.ncy[v.at(r, c + 1)], cxl = m.ncx[o], cyl = m.ncy[o];
                    const double ext = m.nex[v.at(r + 1, c)], eyt = m.ney[v.at(r + 1, c)], exb = m.nex[o], eyb = m.ney[o];
                    const double ur = sFC[(NV + 0)*RW + t + 1], vr = sFC[(NV + 1)*RW + t + 1], nr = sFC[(NV + 2)*RW + t + 1];
                    const double dvdx = (vr*cxr - bchi[1]*cxl + btop[1]*ext - bbot[1]*exb)*Vi;
                    const double dudy = (ur*cyr - bchi[0]*cyl + btop[0]*eyt - bbot[0]*eyb)*Vi;
                    dndx = (nr*cxr - bchi[2]*cxl + btop[2]*ext - bbot[2]*exb)*Vi;
                    dndy = (nr*cyr - bchi[2]*cyl + btop[2]*eyt - bbot[2]*eyb)*Vi;
                    om = fabs(dvdx - dudy);
                }
                const double* S = slot(jl);
                const double mul = VISC ? S[PM*PW + cc] : g.mu_ref;
                const double src = sa_source<double>(S[PR*PW + cc], S[PNI*PW + cc], mul, om, dndx, dndy, prm.wdist[o], prm.beta[o]);
                res[4] += src*V;
            }
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double rk = res[k]*Vi;             // eulerequation.cpp:224-230
                prm.rhs[k*pl + o] = rk;
                acc[k] += rk*rk;
            }
#pragma unroll
            for (int k = 0; k < NV; k++) Dbot[k] = Dtop[k];
            if (SA) { bbot[0] = btop[0]; bbot[1] = btop[1]; bbot[2] = btop[2]; }
        }
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
