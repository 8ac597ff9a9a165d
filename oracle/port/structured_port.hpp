// TEST INFRASTRUCTURE ONLY -- the CPU oracle ("port").  Nothing under structured_b200/ may use it.
//
// A plain restatement of the reference's residual path, templated on the scalar type T so that
// T = double gives the residual and T = oad::Dual<N> / oad::DepSet give the Jacobian through
// oracle/jacdriver.hpp.  Each function cites the reference lines it follows (paths relative to
// /root/reference).  The laminar path (ntrans = 0) is pinned against the reference itself
// (oracle/_ref, tests/test_oracle_pin.py and the fixtures in tests/golden/).
//
// The Spalart-Allmaras extension (ntrans = 1) has NO reference implementation (the snapshot only has
// the ntrans hooks: src/solver/solution.cpp:9,43-45, src/model/eulerequation.cpp:98,226): that part
// is "parity unpinned"; its specification is DESIGN.md section "SA extension" and this file IS its
// normative CPU statement.
#ifndef STRUCTURED_PORT_HPP
#define STRUCTURED_PORT_HPP
#include <vector>
#include <cmath>
#include <cstddef>
#include <algorithm>
#include "../../include/structured_gpu.h"
#include "../adtypes.hpp"

namespace sport {

using std::sqrt; using std::fabs; using std::pow;
using oad::sqrt; using oad::fabs; using oad::pow;

constexpr double GAMMA = 1.4;                       // src/common.h:40

// SA constants (NASA TMR "SA"), see DESIGN.md
constexpr double SA_CB1 = 0.1355, SA_CB2 = 0.622, SA_SIGMA = 2.0/3.0, SA_KAPPA = 0.41;
constexpr double SA_CW2 = 0.3, SA_CW3 = 2.0, SA_CV1 = 7.1, SA_PRT = 0.9;

struct Case {
    int ni = 0, nj = 0, nic = 0, njc = 0, nq = 4, ntrans = 0, nv = 4;
    int order = 1, lhs_order = 1, flux = SGPU_FLUX_ROE;
    double rho_inf = 1, u_inf = 0, v_inf = 0, p_inf = 1/1.4, T_inf = 1/1.4, mu_inf = 0, pr_inf = 0.7;
    double dpdx = 0, dpdy = 0;
    bool viscous = false;
    double R = 1, cp = 1;                           // FluidModel ctor, src/model/fluid.cpp:5-15
    std::vector<sgpu_bc> bcs;
    std::vector<double> xv, yv;                     // [ni][nj]
    std::vector<double> nchi, neta, vol, ds_chi, ds_eta;   // [ni][njc][2], [nic][nj][2], [nic][njc]
    std::vector<double> wall_dist, beta;            // [nic][njc]  (SA)
    int eps_nic = 0, eps_njc = 0;                   // cell counts the limiter constants are taken from (= nic, njc)

    void init(const sgpu_desc& d) {
        ni = d.ni; nj = d.nj; nic = ni - 1; njc = nj - 1; eps_nic = nic; eps_njc = njc;
        ntrans = d.ntrans; nv = nq + ntrans;
        order = d.order; lhs_order = d.lhs_order; flux = d.flux;
        rho_inf = d.rho_inf; u_inf = d.u_inf; v_inf = d.v_inf; p_inf = d.p_inf; T_inf = d.T_inf;
        mu_inf = d.mu_inf; pr_inf = d.pr_inf; dpdx = d.dpdx; dpdy = d.dpdy;
        viscous = mu_inf > 1e-15;                   // src/utils/config.cpp:41
        R = p_inf/rho_inf/T_inf;                    // src/model/fluid.cpp:12
        cp = GAMMA*R/(GAMMA - 1.0);                 // src/model/fluid.cpp:14
        bcs.assign(d.bc, d.bc + d.n_bc);
        for (auto& b : bcs) {                       // BoundaryContainer::get_index, src/model/bc.cpp:436-457
            if (b.end < 0) b.end = ((b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP) ? nic : njc) + 2 + b.end;
        }
        wall_dist.assign((size_t)nic*njc, 1.0);
        beta.assign((size_t)nic*njc, 1.0);
    }

    // Mesh::calc_metrics, src/utils/mesh.cpp:172-205
    void set_grid(const double* x, const double* y) {
        xv.assign(x, x + (size_t)ni*nj); yv.assign(y, y + (size_t)ni*nj);
        nchi.assign((size_t)ni*njc*2, 0); neta.assign((size_t)nic*nj*2, 0);
        vol.assign((size_t)nic*njc, 0); ds_chi.assign((size_t)nic*njc, 0); ds_eta.assign((size_t)nic*njc, 0);
        std::vector<double> reta((size_t)nic*nj*2), rchi((size_t)ni*njc*2);
        auto V = [&](const std::vector<double>& a, int i, int j) { return a[(size_t)i*nj + j]; };
        for (int i = 0; i < nic; i++) for (int j = 0; j < nj; j++) {          // :176-183
            double rx = V(xv, i+1, j) - V(xv, i, j), ry = V(yv, i+1, j) - V(yv, i, j);
            reta[((size_t)i*nj + j)*2] = rx; reta[((size_t)i*nj + j)*2 + 1] = ry;
            neta[((size_t)i*nj + j)*2] = -ry; neta[((size_t)i*nj + j)*2 + 1] = rx;
        }
        for (int i = 0; i < ni; i++) for (int j = 0; j < njc; j++) {          // :185-192
            double rx = V(xv, i, j+1) - V(xv, i, j), ry = V(yv, i, j+1) - V(yv, i, j);
            rchi[((size_t)i*njc + j)*2] = rx; rchi[((size_t)i*njc + j)*2 + 1] = ry;
            nchi[((size_t)i*njc + j)*2] = ry; nchi[((size_t)i*njc + j)*2 + 1] = -rx;
        }
        auto RE = [&](int i, int j, int k) { return reta[((size_t)i*nj + j)*2 + k]; };
        auto RC = [&](int i, int j, int k) { return rchi[((size_t)i*njc + j)*2 + k]; };
        for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) {         // :194-204
            vol[(size_t)i*njc + j] = 0.5*(RE(i,j,0)*RC(i,j,1) - RC(i,j,0)*RE(i,j,1)
                                          + RE(i,j+1,0)*RC(i+1,j,1) - RC(i+1,j,0)*RE(i,j+1,1));
            ds_eta[(size_t)i*njc + j] = std::sqrt(std::pow(neta[((size_t)i*nj + j)*2], 2) + std::pow(neta[((size_t)i*nj + j)*2 + 1], 2));
            ds_chi[(size_t)i*njc + j] = std::sqrt(std::pow(nchi[((size_t)i*njc + j)*2], 2) + std::pow(nchi[((size_t)i*njc + j)*2 + 1], 2));
        }
    }
    double NC(int i, int j, int k) const { return nchi[((size_t)i*njc + j)*2 + k]; }
    double NE(int i, int j, int k) const { return neta[((size_t)i*nj + j)*2 + k]; }
    double VOL(int i, int j) const { return vol[(size_t)i*njc + j]; }
};

// Work arrays of EulerEquation (src/model/eulerequation.cpp:41-96), all for one scalar type.
template <class T>
struct Work {
    int nic, njc, ni, nj;
    std::vector<T> rho, u, v, p, Tm, mu, k;                 // padded (nic+2)x(njc+2)
    std::vector<T> nut, musa;                               // SA: nu~, mu + rho nu~
    // face arrays: chi [ni][njc], eta [nic][nj]
    std::vector<T> l_chi[4], r_chi[4], l_eta[4], r_eta[4];
    std::vector<T> f_chi, f_eta, g_chi, g_eta;              // inviscid / viscous flux [..][..][5]
    std::vector<T> gu_chi, gv_chi, gT_chi, gn_chi, gu_eta, gv_eta, gT_eta, gn_eta;   // gradients [..][..][2]
    std::vector<T> ub_chi, vb_chi, mub_chi, kb_chi, nb_chi, msb_chi;
    std::vector<T> ub_eta, vb_eta, mub_eta, kb_eta, nb_eta, msb_eta;
    void init(const Case& c) {
        nic = c.nic; njc = c.njc; ni = c.ni; nj = c.nj;
        size_t np = (size_t)(nic + 2)*(njc + 2), nc = (size_t)ni*njc, ne = (size_t)nic*nj;
        for (auto* a : {&rho, &u, &v, &p, &Tm, &mu, &k, &nut, &musa}) a->assign(np, T(0.0));
        for (int q = 0; q < 4; q++) { l_chi[q].assign(nc, T(0.0)); r_chi[q].assign(nc, T(0.0)); l_eta[q].assign(ne, T(0.0)); r_eta[q].assign(ne, T(0.0)); }
        f_chi.assign(nc*5, T(0.0)); g_chi.assign(nc*5, T(0.0)); f_eta.assign(ne*5, T(0.0)); g_eta.assign(ne*5, T(0.0));
        for (auto* a : {&gu_chi, &gv_chi, &gT_chi, &gn_chi}) a->assign(nc*2, T(0.0));
        for (auto* a : {&gu_eta, &gv_eta, &gT_eta, &gn_eta}) a->assign(ne*2, T(0.0));
        for (auto* a : {&ub_chi, &vb_chi, &mub_chi, &kb_chi, &nb_chi, &msb_chi}) a->assign(nc, T(0.0));
        for (auto* a : {&ub_eta, &vb_eta, &mub_eta, &kb_eta, &nb_eta, &msb_eta}) a->assign(ne, T(0.0));
    }
    size_t P(int ip, int jp) const { return (size_t)ip*(njc + 2) + jp; }      // padded
    size_t C(int i, int j) const { return (size_t)i*njc + j; }                // chi faces / cells
    size_t E(int i, int j) const { return (size_t)i*nj + j; }                 // eta faces
};

// ---------------------------------------------------------------- fluid model, src/model/fluid.cpp
template <class T> inline T get_T_prho(const Case& c, const T& p, const T& rho) { return p/rho/c.R; }       // :17-21
template <class T> inline T get_rho_pT(const Case& c, const T& p, const T& Tm) { return p/Tm/c.R; }        // :24-27
template <class T> inline T get_p_rhoT(const Case& c, const T& rho, const T& Tm) { return rho*c.R*Tm; }    // :31-34
template <class T> inline T laminar_viscosity(const Case& c, const T& Tm) { return c.mu_inf*pow(Tm/c.T_inf, 2.0/3.0); }   // :38-40
template <class T> inline T thermal_conductivity(const Case& c, const T& Tm) { return laminar_viscosity(c, Tm)*c.cp/c.pr_inf; } // :44-46

// FluidModel::primvars with shifti = shiftj = 1, src/model/fluid.cpp:50-67, src/model/eulerequation.cpp:156-159
template <class T>
void primvars(const Case& c, Work<T>& w, const T* q) {
    for (int i = 0; i < c.nic; i++) for (int j = 0; j < c.njc; j++) {
        const T* Q = q + ((size_t)i*c.njc + j)*c.nv;
        T r = Q[0], uu = Q[1]/r, vv = Q[2]/r;
        size_t o = w.P(i+1, j+1);
        w.rho[o] = r; w.u[o] = uu; w.v[o] = vv;
        w.p[o] = (Q[3] - 0.5*r*(uu*uu + vv*vv))*(GAMMA - 1.0);
        w.Tm[o] = get_T_prho(c, w.p[o], w.rho[o]);
        if (c.ntrans) w.nut[o] = Q[4]/r;                                        // SA: nu~ = (rho nu~)/rho
    }
}

// ---------------------------------------------------------------- boundary conditions, src/model/bc.cpp
// SA ghost rules: freestream nu~ = 3 mu_inf/rho_inf; solid walls nu~_g = -(1.5a - 0.5b);
// slipwall nu~_g = 1.5a - 0.5b; wake/periodic/outflow copy like rho.
template <class T>
void apply_bc(const Case& c, Work<T>& w, const sgpu_bc& b) {
    const int nic = c.nic, njc = c.njc;
    const bool sa = c.ntrans > 0;
    auto P = [&](int ip, int jp) { return w.P(ip, jp); };
    const bool horiz = (b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP);
    switch (b.type) {
    case SGPU_BC_FREESTREAM: {                                                  // bc.cpp:26-60
        for (int s = b.start; s <= b.end; s++) {
            size_t o = horiz ? P(s, b.face == SGPU_FACE_BOTTOM ? 0 : njc + 1) : P(b.face == SGPU_FACE_LEFT ? 0 : nic + 1, s);
            w.rho[o] = c.rho_inf; w.u[o] = c.u_inf; w.v[o] = c.v_inf; w.p[o] = c.p_inf;
            w.Tm[o] = get_T_prho(c, w.p[o], w.rho[o]);
            if (sa) w.nut[o] = 3.0*c.mu_inf/c.rho_inf;
        }
    } break;
    case SGPU_BC_SLIPWALL: {                                                    // bc.cpp:78-128 (bottom/top only)
        if (!horiz) break;
        const bool bot = b.face == SGPU_FACE_BOTTOM;
        const int jend = bot ? 0 : njc + 1, j1 = bot ? 1 : njc, j2 = bot ? 2 : njc - 1;
        for (int i = b.start; i <= b.end; i++) {
            const double nx = c.NE(i-1, bot ? 0 : njc, 0), ny = c.NE(i-1, bot ? 0 : njc, 1);   // :90-91,102-103
            const double ds = nx*nx + ny*ny;
            size_t o = P(i, jend), a = P(i, j1), bb = P(i, j2);
            w.p[o] = 1.5*w.p[a] - 0.5*w.p[bb];
            w.rho[o] = 1.5*w.rho[a] - 0.5*w.rho[bb];
            T un = w.u[a]*nx + w.v[a]*ny;
            w.u[o] = w.u[a] - 2.0*un*nx/ds;
            w.v[o] = w.v[a] - 2.0*un*ny/ds;
            w.Tm[o] = get_T_prho(c, w.p[o], w.rho[o]);
            if (sa) w.nut[o] = 1.5*w.nut[a] - 0.5*w.nut[bb];
        }
    } break;
    case SGPU_BC_WALL: {                                                        // bc.cpp:150-204 (all four faces)
        for (int s = b.start; s <= b.end; s++) {
            size_t o, a, bb;
            if (b.face == SGPU_FACE_BOTTOM) { o = P(s, 0); a = P(s, 1); bb = P(s, 2); }
            else if (b.face == SGPU_FACE_TOP) { o = P(s, njc + 1); a = P(s, njc); bb = P(s, njc - 1); }
            else if (b.face == SGPU_FACE_LEFT) { o = P(0, s); a = P(1, s); bb = P(2, s); }
            else { o = P(nic + 1, s); a = P(nic, s); bb = P(nic - 1, s); }
            w.Tm[o] = 1.5*w.Tm[a] - 0.5*w.Tm[bb];
            w.rho[o] = 1.5*w.rho[a] - 0.5*w.rho[bb];
            w.u[o] = 2.0*b.u - (1.5*w.u[a] - 0.5*w.u[bb]);
            w.v[o] = 2.0*b.v - (1.5*w.v[a] - 0.5*w.v[bb]);
            w.p[o] = get_p_rhoT(c, w.rho[o], w.Tm[o]);
            if (sa) w.nut[o] = -(1.5*w.nut[a] - 0.5*w.nut[bb]);
        }
    } break;
    case SGPU_BC_ISOTHERMALWALL: {                                              // bc.cpp:388-427 (bottom/top only)
        if (!horiz) break;
        const bool bot = b.face == SGPU_FACE_BOTTOM;
        const int jend = bot ? 0 : njc + 1, j1 = bot ? 1 : njc, j2 = bot ? 2 : njc - 1;
        for (int i = b.start; i <= b.end; i++) {
            size_t o = P(i, jend), a = P(i, j1), bb = P(i, j2);
            w.p[o] = 1.5*w.p[a] - 0.5*w.p[bb];
            w.u[o] = 2.0*b.u - (1.5*w.u[a] - 0.5*w.u[bb]);
            w.v[o] = 2.0*b.v - (1.5*w.v[a] - 0.5*w.v[bb]);
            w.Tm[o] = b.T;
            w.rho[o] = get_rho_pT(c, w.p[o], w.Tm[o]);
            if (sa) w.nut[o] = -(1.5*w.nut[a] - 0.5*w.nut[bb]);
        }
    } break;
    case SGPU_BC_WAKE: {                                                        // bc.cpp:224-263 (bottom/top)
        if (!horiz) break;
        const int jend = b.face == SGPU_FACE_BOTTOM ? 0 : njc + 1;
        auto cp = [&](size_t o, size_t s) { w.rho[o] = w.rho[s]; w.u[o] = w.u[s]; w.v[o] = w.v[s]; w.p[o] = w.p[s]; w.Tm[o] = w.Tm[s]; if (sa) w.nut[o] = w.nut[s]; };
        for (int i = b.start; i <= b.end; i++) cp(P(i, jend), P(nic + 1 - i, 1));      // :233-239 (source row 1 for both faces)
        for (int i = b.start; i <= b.end; i++) cp(P(nic + 1 - i, jend), P(i, 1));      // :242-248
    } break;
    case SGPU_BC_OUTFLOW: {                                                     // bc.cpp:284-311 (left/right)
        if (horiz) break;
        const int iend = b.face == SGPU_FACE_LEFT ? 0 : nic + 1;
        for (int j = b.start; j <= b.end; j++) {
            size_t o = P(iend, j), a = P(iend - 1, j);                          // :303 (iend-1 also for "left": reference quirk)
            w.rho[o] = w.rho[a]; w.u[o] = w.u[a]; w.v[o] = w.v[a];
            w.p[o] = c.p_inf;
            w.Tm[o] = get_T_prho(c, w.p[o], w.rho[o]);
            if (sa) w.nut[o] = w.nut[a];
        }
    } break;
    case SGPU_BC_PERIODIC: {                                                    // bc.cpp:329-365
        auto cp = [&](size_t o, size_t s) { w.rho[o] = w.rho[s]; w.u[o] = w.u[s]; w.v[o] = w.v[s]; w.p[o] = w.p[s]; w.Tm[o] = w.Tm[s]; if (sa) w.nut[o] = w.nut[s]; };
        if (horiz) for (int i = b.start; i <= b.end; i++) { cp(P(i, 0), P(i, njc)); cp(P(i, njc + 1), P(i, 1)); }
        else for (int j = b.start; j <= b.end; j++) { cp(P(0, j), P(nic, j)); cp(P(nic + 1, j), P(1, j)); }
    } break;
    default: break;
    }
}

// ---------------------------------------------------------------- reconstruction, src/model/reconstruction.cpp
// first order :29-48 ; second order :78-154 (kappa = 1/3 MUSCL, thm = 2/3, thp = 4/3, reconstruction.h:55-56)
template <class T>
void reconstruct(const Case& c, Work<T>& w, const std::vector<T>& q, int var, int order) {
    const int ni = c.ni, nj = c.nj, nic = c.nic, njc = c.njc;
    std::vector<T>& lc = w.l_chi[var]; std::vector<T>& rc = w.r_chi[var];
    std::vector<T>& le = w.l_eta[var]; std::vector<T>& re = w.r_eta[var];
    for (int i = 0; i < ni; i++) for (int j = 0; j < njc; j++) { lc[w.C(i,j)] = q[w.P(i, j+1)]; rc[w.C(i,j)] = q[w.P(i+1, j+1)]; }
    for (int i = 0; i < nic; i++) for (int j = 0; j < nj; j++) { le[w.E(i,j)] = q[w.P(i+1, j)]; re[w.E(i,j)] = q[w.P(i+1, j+1)]; }
    if (order != 2) return;
    const double thm = 2.0/3.0, thp = 4.0/3.0;
    // :62-63.  eps_nic / eps_njc = nic / njc, except for a CROPPED window of a larger grid (tests/helpers.py::crop_case),
    // which must use the limiter constants of the grid it was cut from
    const double eps_chi = std::pow(10.0/c.eps_nic, 3), eps_eta = std::pow(10.0/c.eps_njc, 3);
    for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) {                   // chi :94-111
        T f2a = q[w.P(i+1, j+1)] - q[w.P(i, j+1)];        // f2[i][j]
        T f2b = q[w.P(i+2, j+1)] - q[w.P(i+1, j+1)];      // f2[i+1][j]
        T a1 = 3.0*f2b*f2a;
        T a2 = 2.0*(f2b - f2a)*(f2b - f2a) + a1;
        T f3qt = 0.25*(a1 + eps_chi)/(a2 + eps_chi);
        lc[w.C(i+1, j)] = lc[w.C(i+1, j)] + f3qt*(thm*f2a + thp*f2b);
        rc[w.C(i, j)] = rc[w.C(i, j)] - f3qt*(thp*f2a + thm*f2b);
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) {                   // eta :133-150
        T f2a = q[w.P(i+1, j+1)] - q[w.P(i+1, j)];
        T f2b = q[w.P(i+1, j+2)] - q[w.P(i+1, j+1)];
        T a1 = 3.0*f2b*f2a;
        T a2 = 2.0*(f2b - f2a)*(f2b - f2a) + a1;
        T f3qt = 0.25*(a1 + eps_eta)/(a2 + eps_eta);
        le[w.E(i, j+1)] = le[w.E(i, j+1)] + f3qt*(thm*f2a + thp*f2b);
        re[w.E(i, j)] = re[w.E(i, j)] - f3qt*(thp*f2a + thm*f2b);
    }
}

// ---------------------------------------------------------------- inviscid fluxes, src/model/flux.cpp
template <class T>
inline void roe_flux(double nx, double ny, const T& rlft, const T& ulft, const T& vlft, const T& plft,
                     const T& rrht, const T& urht, const T& vrht, const T& prht, T* f) {   // :51-146
    const double gm1 = GAMMA - 1.0, ogm1 = 1.0/gm1;
    T rlfti = 1.0/rlft, rulft = rlft*ulft, rvlft = rlft*vlft;
    T uvl = 0.5*(ulft*ulft + vlft*vlft), elft = plft*ogm1 + rlft*uvl, hlft = (elft + plft)*rlfti;
    T rrhti = 1.0/rrht, rurht = rrht*urht, rvrht = rrht*vrht;
    T uvr = 0.5*(urht*urht + vrht*vrht), erht = prht*ogm1 + rrht*uvr, hrht = (erht + prht)*rrhti;
    T rat = sqrt(rrht*rlfti), rati = 1.0/(rat + 1.0), rav = rat*rlft;
    T uav = (rat*urht + ulft)*rati, vav = (rat*vrht + vlft)*rati, hav = (rat*hrht + hlft)*rati;
    T uv = 0.5*(uav*uav + vav*vav), cav = sqrt(gm1*(hav - uv));
    T aq1 = rrht - rlft, aq2 = urht - ulft, aq3 = vrht - vlft, aq4 = prht - plft;
    const double dr = std::sqrt(nx*nx + ny*ny), r1 = nx/dr, r2 = ny/dr;
    T uu = r1*uav + r2*vav, c2 = cav*cav, c2i = 1.0/c2;
    T auu = fabs(uu), aupc = fabs(uu + cav), aumc = fabs(uu - cav);
    T uulft = r1*ulft + r2*vlft, uurht = r1*urht + r2*vrht, rcav = rav*cav, aquu = uurht - uulft;
    T c2ih = 0.5*c2i, ruuav = auu*rav;
    T b1 = auu*(aq1 - c2i*aq4), b2 = c2ih*aupc*(aq4 + rcav*aquu), b3 = c2ih*aumc*(aq4 - rcav*aquu);
    T b4 = b1 + b2 + b3, b5 = cav*(b2 - b3), b6 = ruuav*(aq2 - r1*aquu), b7 = ruuav*(aq3 - r2*aquu);
    aq1 = b4; aq2 = uav*b4 + r1*b5 + b6; aq3 = vav*b4 + r2*b5 + b7;
    aq4 = hav*b4 + uu*b5 + uav*b6 + vav*b7 - c2*b1*ogm1;
    const double aj = 0.5*dr;
    T plar = plft + prht, eplft = elft + plft, eprht = erht + prht;
    f[0] = aj*(rlft*uulft + rrht*uurht - aq1);
    f[1] = aj*(rulft*uulft + rurht*uurht + r1*plar - aq2);
    f[2] = aj*(rvlft*uulft + rvrht*uurht + r2*plar - aq3);
    f[3] = aj*(eplft*uulft + eprht*uurht - aq4);
}

template <class T> inline T mach_p(const T& M) { return fabs(M) <= 1.0 ? 0.25*(M + 1.0)*(M + 1.0) : 0.5*(M + fabs(M)); }   // :150
template <class T> inline T mach_m(const T& M) { return fabs(M) <= 1.0 ? -0.25*(M - 1.0)*(M - 1.0) : 0.5*(M - fabs(M)); }  // :152
template <class T> inline T pres_p(const T& M, const T& p) { return fabs(M) <= 1.0 ? 0.25*p*(M + 1.0)*(M + 1.0)*(2.0 - M) : 0.5*p*(M + fabs(M))/M; } // :154
template <class T> inline T pres_m(const T& M, const T& p) { return fabs(M) <= 1.0 ? 0.25*p*(M - 1.0)*(M - 1.0)*(2.0 + M) : 0.5*p*(M - fabs(M))/M; } // :156

template <class T>
inline void ausm_flux(double nx, double ny, const T& rlft, const T& ulft, const T& vlft, const T& plft,
                      const T& rrht, const T& urht, const T& vrht, const T& prht, T* f) {  // :159-224
    const double gm1 = GAMMA - 1.0, ogm1 = 1.0/gm1;
    const double ds = std::sqrt(nx*nx + ny*ny);
    T uln = (ulft*nx + vlft*ny)/ds, urn = (urht*nx + vrht*ny)/ds;
    T alft = sqrt(GAMMA*plft/rlft), arht = sqrt(GAMMA*prht/rrht);
    T machlft = uln/alft, machrht = urn/arht;
    T rlfti = 1.0/rlft, uvl = 0.5*(ulft*ulft + vlft*vlft), elft = plft*ogm1 + rlft*uvl, hlft = (elft + plft)*rlfti;
    T rrhti = 1.0/rrht, uvr = 0.5*(urht*urht + vrht*vrht), erht = prht*ogm1 + rrht*uvr, hrht = (erht + prht)*rrhti;
    T mach_half = mach_p(machlft) + mach_m(machrht);
    T p_half = pres_p(machlft, plft) + pres_m(machrht, prht);
    if (mach_half >= 0.0) {
        f[0] = rlft*alft*mach_half*ds;
        f[1] = rlft*alft*ulft*mach_half*ds + p_half*nx;
        f[2] = rlft*alft*vlft*mach_half*ds + p_half*ny;
        f[3] = rlft*alft*hlft*mach_half*ds;
    } else {
        f[0] = rrht*arht*mach_half*ds;
        f[1] = rrht*arht*urht*mach_half*ds + p_half*nx;
        f[2] = rrht*arht*vrht*mach_half*ds + p_half*ny;
        f[3] = rrht*arht*hrht*mach_half*ds;
    }
}

// ---------------------------------------------------------------- face operators, src/utils/mesh.cpp
// Mesh::calc_face :10-32
template <class T>
void calc_face(const Case& c, const Work<T>& w, const std::vector<T>& q, std::vector<T>& q_chi, std::vector<T>& q_eta) {
    for (int i = 0; i < c.ni; i++) for (int j = 0; j < c.njc; j++) {
        T ql = q[w.P(i, j+1)], qr = q[w.P(i+1, j+1)];
        T qt = 0.25*(ql + qr + q[w.P(i, j+2)] + q[w.P(i+1, j+2)]);
        T qb = 0.25*(ql + qr + q[w.P(i, j)] + q[w.P(i+1, j)]);
        q_chi[w.C(i, j)] = 0.25*(ql + qr + qt + qb);
    }
    for (int i = 0; i < c.nic; i++) for (int j = 0; j < c.nj; j++) {
        T qt = q[w.P(i+1, j+1)], qb = q[w.P(i+1, j)];
        T ql = 0.25*(qt + qb + q[w.P(i, j+1)] + q[w.P(i, j)]);
        T qr = 0.25*(qt + qb + q[w.P(i+2, j+1)] + q[w.P(i+2, j)]);
        q_eta[w.E(i, j)] = 0.25*(ql + qr + qt + qb);
    }
}

// Mesh::calc_gradient(q, grad_chi, grad_eta) :36-131
template <class T>
void calc_gradient(const Case& c, const Work<T>& w, const std::vector<T>& q, std::vector<T>& g_chi, std::vector<T>& g_eta) {
    const int ni = c.ni, nj = c.nj, nic = c.nic, njc = c.njc;
    double nt[2], nb[2], nl[2], nr[2], vol;
    for (int i = 0; i < ni; i++) for (int j = 0; j < njc; j++) {
        T ql = q[w.P(i, j+1)], qr = q[w.P(i+1, j+1)];
        T qt = 0.25*(ql + qr + q[w.P(i, j+2)] + q[w.P(i+1, j+2)]);
        T qb = 0.25*(ql + qr + q[w.P(i, j)] + q[w.P(i+1, j)]);
        if (i == 0) {                                                            // :54-63
            vol = c.VOL(i, j);
            for (int k = 0; k < 2; k++) { nt[k] = c.NE(i, j+1, k); nb[k] = c.NE(i, j, k); nr[k] = 0.5*(c.NC(i, j, k) + c.NC(i+1, j, k)); nl[k] = c.NC(i, j, k); }
        } else if (i == ni - 1) {                                                // :65-73
            vol = c.VOL(i-1, j);
            for (int k = 0; k < 2; k++) { nt[k] = c.NE(i-1, j+1, k); nb[k] = c.NE(i-1, j, k); nr[k] = c.NC(i, j, k); nl[k] = 0.5*(c.NC(i, j, k) + c.NC(i-1, j, k)); }
        } else {                                                                 // :74-82
            vol = 0.5*(c.VOL(i, j) + c.VOL(i-1, j));
            for (int k = 0; k < 2; k++) {
                nt[k] = 0.5*(c.NE(i, j+1, k) + c.NE(i-1, j+1, k)); nb[k] = 0.5*(c.NE(i, j, k) + c.NE(i-1, j, k));
                nr[k] = 0.5*(c.NC(i, j, k) + c.NC(i+1, j, k)); nl[k] = 0.5*(c.NC(i, j, k) + c.NC(i-1, j, k));
            }
        }
        for (int k = 0; k < 2; k++) g_chi[w.C(i, j)*2 + k] = (nt[k]*qt - nb[k]*qb + nr[k]*qr - nl[k]*ql)/vol;   // :83-84
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < nj; j++) {
        T qt = q[w.P(i+1, j+1)], qb = q[w.P(i+1, j)];
        T ql = 0.25*(qt + qb + q[w.P(i, j+1)] + q[w.P(i, j)]);
        T qr = 0.25*(qt + qb + q[w.P(i+2, j+1)] + q[w.P(i+2, j)]);
        if (j == 0) {                                                            // :99-107
            vol = c.VOL(i, j);
            for (int k = 0; k < 2; k++) { nt[k] = 0.5*(c.NE(i, j, k) + c.NE(i, j+1, k)); nb[k] = c.NE(i, j, k); nl[k] = c.NC(i, j, k); nr[k] = c.NC(i+1, j, k); }
        } else if (j == nj - 1) {                                                // :108-116
            for (int k = 0; k < 2; k++) { nt[k] = c.NE(i, j, k); nb[k] = 0.5*(c.NE(i, j, k) + c.NE(i, j-1, k)); nl[k] = c.NC(i, j-1, k); nr[k] = c.NC(i+1, j-1, k); }
            vol = c.VOL(i, j-1);
        } else {                                                                 // :117-126
            for (int k = 0; k < 2; k++) {
                nt[k] = 0.5*(c.NE(i, j, k) + c.NE(i, j+1, k)); nb[k] = 0.5*(c.NE(i, j, k) + c.NE(i, j-1, k));
                nl[k] = 0.5*(c.NC(i, j, k) + c.NC(i, j-1, k)); nr[k] = 0.5*(c.NC(i+1, j, k) + c.NC(i+1, j-1, k));
            }
            vol = 0.5*(c.VOL(i, j) + c.VOL(i, j-1));
        }
        for (int k = 0; k < 2; k++) g_eta[w.E(i, j)*2 + k] = (nt[k]*qt - nb[k]*qb + nr[k]*qr - nl[k]*ql)/vol;   // :127-128
    }
}

// DiffusiveFluxGreenGauss::evaluate, src/model/flux.cpp:12-48 ; component 4 (SA) is the extension
template <class T>
inline void viscous_flux(double nx, double ny, const T* gu, const T* gv, const T* gT, const T& ubar, const T& vbar,
                         const T& mu, const T& k, T* f) {
    T dudx = gu[0], dudy = gu[1], dvdx = gv[0], dvdy = gv[1], dTdx = gT[0], dTdy = gT[1];
    T tau_xy = mu*(dudy + dvdx);
    T tau_xx = mu*(2.0*dudx - 2.0/3.0*(dudx + dvdy));
    T tau_yy = mu*(2.0*dvdy - 2.0/3.0*(dudx + dvdy));
    T q_x = -k*dTdx, q_y = -k*dTdy;
    f[0] = 0.0;
    f[1] = tau_xx*nx + tau_xy*ny;
    f[2] = tau_xy*nx + tau_yy*ny;
    f[3] = (ubar*tau_xx + vbar*tau_xy - q_x)*nx + (ubar*tau_xy + vbar*tau_yy - q_y)*ny;
}

// ---------------------------------------------------------------- EulerEquation::calc_residual
// src/model/eulerequation.cpp:202-232 with calc_intermediates :168-199, calc_convective_residual :136-154,
// calc_viscous_residual :5-20, calc_source_residual :22-29.
template <class T>
void calc_residual(const Case& c, Work<T>& w, const T* q, T* rhs, bool lhs) {
    const int ni = c.ni, nj = c.nj, nic = c.nic, njc = c.njc, nv = c.nv;
    const bool sa = c.ntrans > 0;
    const int order = lhs ? c.lhs_order : c.order;                               // :203-208
    for (size_t n = 0; n < (size_t)nic*njc*nv; n++) rhs[n] = T(0.0);             // :210
    primvars(c, w, q);                                                           // :158
    for (const auto& b : c.bcs) apply_bc(c, w, b);                               // :164, bc.cpp:430-433
    reconstruct(c, w, w.rho, 0, order); reconstruct(c, w, w.u, 1, order);        // :172-180
    reconstruct(c, w, w.v, 2, order); reconstruct(c, w, w.p, 3, order);
    if (c.viscous) {                                                             // :182-198
        for (int ip = 0; ip < nic + 2; ip++) for (int jp = 0; jp < njc + 2; jp++) {
            size_t o = w.P(ip, jp);
            w.mu[o] = laminar_viscosity(c, w.Tm[o]);
            w.k[o] = thermal_conductivity(c, w.Tm[o]);
            if (sa) {                                                            // SA coupling
                T chi = w.rho[o]*w.nut[o]/w.mu[o];
                T chi3 = chi*chi*chi;
                T fv1 = chi3/(chi3 + SA_CV1*SA_CV1*SA_CV1);
                T mut = w.rho[o]*w.nut[o]*fv1;
                w.musa[o] = w.mu[o] + w.rho[o]*w.nut[o];
                w.k[o] = c.cp*(w.mu[o]/c.pr_inf + mut/SA_PRT);
                w.mu[o] = w.mu[o] + mut;
            }
        }
        calc_gradient(c, w, w.u, w.gu_chi, w.gu_eta);
        calc_gradient(c, w, w.v, w.gv_chi, w.gv_eta);
        calc_gradient(c, w, w.Tm, w.gT_chi, w.gT_eta);
        calc_face(c, w, w.u, w.ub_chi, w.ub_eta);
        calc_face(c, w, w.v, w.vb_chi, w.vb_eta);
        calc_face(c, w, w.mu, w.mub_chi, w.mub_eta);
        calc_face(c, w, w.k, w.kb_chi, w.kb_eta);
        if (sa) {
            calc_gradient(c, w, w.nut, w.gn_chi, w.gn_eta);
            calc_face(c, w, w.nut, w.nb_chi, w.nb_eta);
            calc_face(c, w, w.musa, w.msb_chi, w.msb_eta);
        }
    }
    // convective fluxes :138-145
    for (int i = 0; i < ni; i++) for (int j = 0; j < njc; j++) {
        size_t o = w.C(i, j); T* f = &w.f_chi[o*5];
        if (c.flux == SGPU_FLUX_ROE) roe_flux(c.NC(i,j,0), c.NC(i,j,1), w.l_chi[0][o], w.l_chi[1][o], w.l_chi[2][o], w.l_chi[3][o], w.r_chi[0][o], w.r_chi[1][o], w.r_chi[2][o], w.r_chi[3][o], f);
        else ausm_flux(c.NC(i,j,0), c.NC(i,j,1), w.l_chi[0][o], w.l_chi[1][o], w.l_chi[2][o], w.l_chi[3][o], w.r_chi[0][o], w.r_chi[1][o], w.r_chi[2][o], w.r_chi[3][o], f);
        if (sa) { T nl = w.nut[w.P(i, j+1)], nr = w.nut[w.P(i+1, j+1)]; f[4] = (f[0] >= 0.0) ? f[0]*nl : f[0]*nr; }
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < nj; j++) {
        size_t o = w.E(i, j); T* f = &w.f_eta[o*5];
        if (c.flux == SGPU_FLUX_ROE) roe_flux(c.NE(i,j,0), c.NE(i,j,1), w.l_eta[0][o], w.l_eta[1][o], w.l_eta[2][o], w.l_eta[3][o], w.r_eta[0][o], w.r_eta[1][o], w.r_eta[2][o], w.r_eta[3][o], f);
        else ausm_flux(c.NE(i,j,0), c.NE(i,j,1), w.l_eta[0][o], w.l_eta[1][o], w.l_eta[2][o], w.l_eta[3][o], w.r_eta[0][o], w.r_eta[1][o], w.r_eta[2][o], w.r_eta[3][o], f);
        if (sa) { T nl = w.nut[w.P(i+1, j)], nr = w.nut[w.P(i+1, j+1)]; f[4] = (f[0] >= 0.0) ? f[0]*nl : f[0]*nr; }
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) for (int k = 0; k < nv; k++) {      // :146-153
        T& r = rhs[((size_t)i*njc + j)*nv + k];
        r -= (w.f_eta[w.E(i, j+1)*5 + k] - w.f_eta[w.E(i, j)*5 + k]);
        r -= (w.f_chi[w.C(i+1, j)*5 + k] - w.f_chi[w.C(i, j)*5 + k]);
    }
    if (c.viscous) {                                                             // :5-20
        for (int i = 0; i < ni; i++) for (int j = 0; j < njc; j++) {
            size_t o = w.C(i, j); T* g = &w.g_chi[o*5];
            viscous_flux(c.NC(i,j,0), c.NC(i,j,1), &w.gu_chi[o*2], &w.gv_chi[o*2], &w.gT_chi[o*2], w.ub_chi[o], w.vb_chi[o], w.mub_chi[o], w.kb_chi[o], g);
            if (sa) g[4] = w.msb_chi[o]/SA_SIGMA*(w.gn_chi[o*2]*c.NC(i,j,0) + w.gn_chi[o*2 + 1]*c.NC(i,j,1));
        }
        for (int i = 0; i < nic; i++) for (int j = 0; j < nj; j++) {
            size_t o = w.E(i, j); T* g = &w.g_eta[o*5];
            viscous_flux(c.NE(i,j,0), c.NE(i,j,1), &w.gu_eta[o*2], &w.gv_eta[o*2], &w.gT_eta[o*2], w.ub_eta[o], w.vb_eta[o], w.mub_eta[o], w.kb_eta[o], g);
            if (sa) g[4] = w.msb_eta[o]/SA_SIGMA*(w.gn_eta[o*2]*c.NE(i,j,0) + w.gn_eta[o*2 + 1]*c.NE(i,j,1));
        }
        for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) for (int k = 1; k < nv; k++) {  // :11-18 (k starts at 1)
            T& r = rhs[((size_t)i*njc + j)*nv + k];
            r += (w.g_eta[w.E(i, j+1)*5 + k] - w.g_eta[w.E(i, j)*5 + k]);
            r += (w.g_chi[w.C(i+1, j)*5 + k] - w.g_chi[w.C(i, j)*5 + k]);
        }
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) {                // :22-29
        rhs[((size_t)i*njc + j)*nv + 1] += -c.dpdx*c.VOL(i, j);
        rhs[((size_t)i*njc + j)*nv + 2] += -c.dpdy*c.VOL(i, j);
    }
    if (sa) {                                                                    // SA source
        const double k2 = SA_KAPPA*SA_KAPPA;
        const double cw1 = SA_CB1/k2 + (1.0 + SA_CB2)/SA_SIGMA;
        const double cw36 = std::pow(SA_CW3, 6);
        for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) {
            const double V = c.VOL(i, j);
            auto gg = [&](const std::vector<T>& bc, const std::vector<T>& be, int k) {   // Green-Gauss over the cell's own faces
                return (bc[w.C(i+1, j)]*c.NC(i+1, j, k) - bc[w.C(i, j)]*c.NC(i, j, k)
                        + be[w.E(i, j+1)]*c.NE(i, j+1, k) - be[w.E(i, j)]*c.NE(i, j, k))/V;
            };
            T dvdx = gg(w.vb_chi, w.vb_eta, 0), dudy = gg(w.ub_chi, w.ub_eta, 1);
            T dndx = gg(w.nb_chi, w.nb_eta, 0), dndy = gg(w.nb_chi, w.nb_eta, 1);
            T om = fabs(dvdx - dudy);
            size_t o = w.P(i+1, j+1);
            T rho = w.rho[o], nut = w.nut[o];
            T mul = laminar_viscosity(c, w.Tm[o]);
            T chi = rho*nut/mul, chi3 = chi*chi*chi;
            T fv1 = chi3/(chi3 + SA_CV1*SA_CV1*SA_CV1);
            T fv2 = 1.0 - chi/(1.0 + chi*fv1);
            const double d = c.wall_dist[(size_t)i*njc + j], k2d2 = k2*d*d;
            T sbar = nut*fv2/k2d2;
            T st = om + sbar, st_min = 0.3*om;
            if (st < st_min) st = st_min;
            T den = st*k2d2;
            if (den < 1e-30) den = T(1e-30);
            T r = nut/den;
            if (r > 10.0) r = T(10.0);
            T r2 = r*r, r6 = r2*r2*r2;
            T g = r + SA_CW2*(r6 - r);
            T g2 = g*g, g6 = g2*g2*g2;
            T fw = g*pow((1.0 + cw36)/(g6 + cw36), 1.0/6.0);
            T nd = nut/d;
            T src = rho*(c.beta[(size_t)i*njc + j]*SA_CB1*st*nut - cw1*fw*nd*nd) + SA_CB2/SA_SIGMA*rho*(dndx*dndx + dndy*dndy);
            rhs[((size_t)i*njc + j)*nv + 4] += src*V;
        }
    }
    for (int i = 0; i < nic; i++) for (int j = 0; j < njc; j++) for (int k = 0; k < nv; k++)        // :224-230
        rhs[((size_t)i*njc + j)*nv + k] /= c.VOL(i, j);
}

// EulerEquation::calc_dt, src/model/eulerequation.cpp:237-259 ; the SA row takes the same dt
inline void calc_dt(const Case& c, const double* q, double cfl, double* dt) {
    for (int i = 0; i < c.nic; i++) for (int j = 0; j < c.njc; j++) {
        const double* Q = q + ((size_t)i*c.njc + j)*c.nv;
        double rho = Q[0], u = Q[1]/rho, v = Q[2]/rho, rhoE = Q[3];
        double p = (rhoE - 0.5*rho*(u*u + v*v))*(GAMMA - 1.0);
        double lambda = std::sqrt(GAMMA*p/rho) + std::fabs(u) + std::fabs(v);
        double len_min = std::min(c.ds_eta[(size_t)i*c.njc + j], c.ds_chi[(size_t)i*c.njc + j]);
        double mu = c.mu_inf;
        for (int k = 0; k < c.nv; k++) dt[((size_t)i*c.njc + j)*c.nv + k] = cfl/(lambda/len_min + 2.0*mu/len_min/len_min);
    }
}

} // namespace sport
#endif
