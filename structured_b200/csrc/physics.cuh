// Device-side physics of the residual path (cited per function, paths relative to the reference tree).  Templated on
// the scalar type S so that the same statements run on `double` (residual kernel) and on small forward-mode dual numbers
// (Jacobian kernel).
// roe_flux and ausm_flux below FOLLOW THE REFERENCE'S OPERATION SEQUENCE (src/model/flux.cpp:51-146,150-224) statement by
// statement and keep its identifiers (rlft, uav, b1..b7, aq1..aq4, plar, eplft): the 1e-12 parity bar on a flux whose terms
// cancel leaves no freedom in the order of the dissipation sums.  What differs is the reciprocal / square-root structure (one
// rsqrt of rho_L rho_R serves both 1/rho and the Roe weight, one rsqrt gives c and 1/c^2, MUFU seeds + third-order steps
// instead of IEEE divisions) -- see the notes at each site.  Everything else in this file is written from the formulas.
#pragma once
#include <cuda_runtime.h>

namespace sg {

constexpr double GAMMA = 1.4;                 // src/common.h:40
constexpr double GM1_C = GAMMA - 1.0;
constexpr double OGM1 = 1.0/GM1_C;
// Doubles whose low mantissa word is non-zero cannot be encoded as SASS immediates: as literals each use costs two
// UMOV issue slots.  Read from the constant bank they are free operands of DFMA/DMUL.
__constant__ double c_kc[4] = {GAMMA - 1.0, 2.0/3.0, 4.0/3.0, GAMMA};
#define GM1 (c_kc[0])
#define K23 (c_kc[1])
#define K43 (c_kc[2])

// Spalart-Allmaras constants (extension, DESIGN.md "SA extension")
constexpr double SA_CB1 = 0.1355, SA_CB2 = 0.622, SA_SIGMA = 2.0/3.0, SA_KAPPA = 0.41;
constexpr double SA_CW2 = 0.3, SA_CW3 = 2.0, SA_CV1 = 7.1, SA_PRT = 0.9;

// ---- scalar helpers (double overloads; Dual overloads live in dual.cuh) ---------------------------
// fp64 division / sqrt cost 10-20 issue slots each as IEEE sequences; the kernels are fp64-issue bound, so
// reciprocals and reciprocal square roots are MUFU seeds (RCP64H / RSQ64H, ~20 bits) + two Newton steps:
// measured max relative error 1.1e-16 / 2.2e-16 on B200, i.e. 1 ulp -- far inside the 1e-12 parity bar.
__device__ __forceinline__ double rcp_fast(double x) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // one third-order step: e = 1 - x r0 (|e| <= 2^-20), r = r0 (1 + e + e^2): error e^3 ~ 2^-60, 3 dependent DFMAs
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // third-order step: e = 1 - x y0^2, y = y0 (1 + e/2 + 3 e^2/8)
    const double t = x*y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5)*e;
    return fma(y, p, y);
}
__device__ __forceinline__ double s_rcp(double x) { return rcp_fast(x); }
__device__ __forceinline__ double s_rsqrt(double x) { return rsqrt_fast(x); }
__device__ __forceinline__ double s_sqrt(double x) { return x > 0.0 ? x*rsqrt_fast(x) : 0.0; }
__device__ __forceinline__ double s_abs(double x) { return fabs(x); }
__device__ __forceinline__ double s_val(double x) { return x; }
// x^(2/3): FluidModel::get_laminar_viscosity uses pow(T/T_ref, 2.0/3.0) (src/model/fluid.cpp:38-40).  libdevice's
// cbrt costs ~80 issue slots; here x^(2/3) = x y with y = x^(-1/3) from a float seed (MUFU.LG2/EX2, ~1e-6) and ONE
// division-free third-order step, y = y0 (1 + e/3 + 2 e^2/9), e = 1 - x y0^3 (truncation 0.17 e^3 ~ 5e-18): 7 fp64
// instructions, no reciprocal in the dependency chain; measured max relative error 4.4e-16 on [1e-3, 1e3].
// float seed x^y for the root refinements below: bare MUFU.LG2 / MUFU.EX2 (the arguments are normal floats well inside the
// range, so the denormal scaling and range checks __powf wraps around them -- 6 extra instructions with long fixed stalls --
// are not needed); relative error ~ 2^-22 |y log2 x|
__device__ __forceinline__ double pow_seed(double x, float y) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l*y));
    return (double)r;
}
__device__ __forceinline__ double s_pow23(double x) {
    const double y0 = pow_seed(x, -0.333333333f);
    const double e = fma(-x, y0*y0*y0, 1.0);
    const double pe = fma(2.0/9.0, e, 1.0/3.0)*e;
    return x*fma(y0, pe, y0);
}
// x^(1/6) for the SA f_w function: float seed + two Newton steps on y^6 = x (measured 8.9e-16 on [1e-33, 1e3])
__device__ __forceinline__ double s_pow16(double x) {
    double y = (double)__powf((float)x, 0.166666667f);
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double y2 = y*y, y6 = y2*y2*y2;
        y = y*fma(x*rcp_fast(y6), 1.0/6.0, 5.0/6.0);
    }
    return y;
}

struct Gas {            // FluidModel, src/model/fluid.cpp:5-15
    double R, cp, pr, mu_ref, T_ref;
    double rho_inf, u_inf, v_inf, p_inf;
    double cp_over_pr;
    double iR, iT_ref, cp_over_prt;      // 1/R, 1/T_ref, cp/Pr_t: divisions by constants hoisted to the host
};

// FluidModel::primvars, src/model/fluid.cpp:50-67 (T = p/rho/R, :19-21)
template <class S>
__device__ __forceinline__ void cons_to_prim(const Gas& g, S q0, S q1, S q2, S q3, S& r, S& u, S& v, S& p, S& T) {
    r = q0;
    S ri = s_rcp(q0);
    u = q1*ri; v = q2*ri;
    p = (q3 - 0.5*r*(u*u + v*v))*GM1;
    T = p*ri*g.iR;
}
__device__ __forceinline__ void prim_to_cons(double r, double u, double v, double p, double& q0, double& q1, double& q2, double& q3) {
    q0 = r; q1 = r*u; q2 = r*v; q3 = p*OGM1 + 0.5*r*(u*u + v*v);
}
template <class S>
__device__ __forceinline__ S laminar_viscosity(const Gas& g, S T) { return g.mu_ref*s_pow23(T*g.iT_ref); }   // fluid.cpp:38-40

// ---- reconstruction: ReconstructionSecondOrder, src/model/reconstruction.cpp:94-111,133-150 -------
// For one interior cell with neighbours (qm, q0, qp) along a direction returns the value extrapolated to
// its high face (becomes that face's LEFT state) and to its low face (that face's RIGHT state).
template <class S>
__device__ __forceinline__ void muscl_cell(S qm, S q0, S qp, double eps, S& to_high, S& to_low) {
    const double thm = K23, thp = K43;                  // reconstruction.h:55-56
    S f2a = q0 - qm, f2b = qp - q0;
    S a1 = 3.0*f2b*f2a;
    S d = f2b - f2a;
    S a2 = 2.0*d*d + a1;
    S f3qt = 0.25*(a1 + eps)*s_rcp(a2 + eps);
    to_high = q0 + f3qt*(thm*f2a + thp*f2b);
    to_low = q0 - f3qt*(thp*f2a + thm*f2b);
}

// double form for the residual kernel: the same limiter with the products contracted by hand -- m = f2a f2b once -- and
// numerator and denominator both HALVED with thm = 2/3 folded into the numerator (the ratio is what matters):
//   thm f3qt = (m/4 + eps/12) / (d^2 + 3/2 m + eps/2),   thp = 2 thm,
// and two variables sharing ONE reciprocal: 1/den_a = den_b/(den_a den_b) (den >= eps/2 ~ 1e-8, products stay far inside the
// fp64 range).  15 fp64 instructions per variable and direction.  Differs from the template above by rounding only
// (<= 4 ulp of q0 measured).  eps12 = eps/12, epsh = eps/2 are formed once per thread by the caller.
__device__ __forceinline__ void muscl_cell2(const double* qm, const double* q0, const double* qp, double eps12, double epsh, double* to_high, double* to_low) {
    double f2a[2], f2b[2], num[2], den[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        f2a[k] = q0[k] - qm[k]; f2b[k] = qp[k] - q0[k];
        const double m = f2b[k]*f2a[k], d = f2b[k] - f2a[k];
        num[k] = fma(0.25, m, eps12);
        den[k] = fma(d, d, fma(1.5, m, epsh));
    }
    const double r = rcp_fast(den[0]*den[1]);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double g = num[k]*(den[1 - k]*r);
        to_high[k] = fma(g, fma(2.0, f2b[k], f2a[k]), q0[k]);
        to_low[k] = fma(-g, fma(2.0, f2a[k], f2b[k]), q0[k]);
    }
}

// ---- ConvectiveFluxRoe::evaluate, src/model/flux.cpp:51-146 ------------------------------------------
template <class S>
__device__ __forceinline__ void roe_flux(double nx, double ny, S rlft, S ulft, S vlft, S plft,
                                         S rrht, S urht, S vrht, S prht, S* f) {
    // ONE reciprocal square root of rho_L rho_R gives 1/(rho_L rho_R) (both 1/rho) and sqrt(rho_L rho_R) (the Roe density and,
    // times 1/rho_L, the Roe weight): dependency depth rsqrt -> rcp instead of rcp -> rsqrt -> rcp
    S rr_ = rlft*rrht, yrr = s_rsqrt(rr_), pi = yrr*yrr;
    S rlfti = rrht*pi, rrhti = rlft*pi;
    S rulft = rlft*ulft, rvlft = rlft*vlft;
    S uvl = 0.5*(ulft*ulft + vlft*vlft), elft = plft*OGM1 + rlft*uvl, hlft = (elft + plft)*rlfti;
    S rurht = rrht*urht, rvrht = rrht*vrht;
    S uvr = 0.5*(urht*urht + vrht*vrht), erht = prht*OGM1 + rrht*uvr, hrht = (erht + prht)*rrhti;
    S rav = rr_*yrr, rat = rav*rlfti, rati = s_rcp(rat + 1.0);
    S uav = (rat*urht + ulft)*rati, vav = (rat*vrht + vlft)*rati, hav = (rat*hrht + hlft)*rati;
    S uv = 0.5*(uav*uav + vav*vav);
    S c2 = GM1*(hav - uv);
    S yc = s_rsqrt(c2);
    S cav = c2*yc, c2i = yc*yc;
    S aq1 = rrht - rlft, aq2 = urht - ulft, aq3 = vrht - vlft, aq4 = prht - plft;
    const double nn = nx*nx + ny*ny, dri = rsqrt_fast(nn), dr = nn*dri, r1 = nx*dri, r2 = ny*dri;
    S uu = r1*uav + r2*vav;
    S auu = s_abs(uu), aupc = s_abs(uu + cav), aumc = s_abs(uu - cav);
    S uulft = r1*ulft + r2*vlft, uurht = r1*urht + r2*vrht, rcav = rav*cav, aquu = uurht - uulft;
    S c2ih = 0.5*c2i, ruuav = auu*rav;
    S b1 = auu*(aq1 - c2i*aq4), b2 = c2ih*aupc*(aq4 + rcav*aquu), b3 = c2ih*aumc*(aq4 - rcav*aquu);
    S b4 = b1 + b2 + b3, b5 = cav*(b2 - b3), b6 = ruuav*(aq2 - r1*aquu), b7 = ruuav*(aq3 - r2*aquu);
    aq1 = b4; aq2 = uav*b4 + r1*b5 + b6; aq3 = vav*b4 + r2*b5 + b7;
    aq4 = hav*b4 + uu*b5 + uav*b6 + vav*b7 - c2*b1*OGM1;
    const double aj = 0.5*dr;
    S plar = plft + prht, eplft = elft + plft, eprht = erht + prht;
    f[0] = aj*(rlft*uulft + rrht*uurht - aq1);
    f[1] = aj*(rulft*uulft + rurht*uurht + r1*plar - aq2);
    f[2] = aj*(rvlft*uulft + rvrht*uurht + r2*plar - aq3);
    f[3] = aj*(eplft*uulft + eprht*uurht - aq4);
}

// ---- ConvectiveFluxAUSM, src/model/flux.cpp:150-224 --------------------------------------------------
template <class S> __device__ __forceinline__ S mach_p(S M) { return s_val(s_abs(M)) <= 1.0 ? 0.25*(M + 1.0)*(M + 1.0) : 0.5*(M + s_abs(M)); }
template <class S> __device__ __forceinline__ S mach_m(S M) { return s_val(s_abs(M)) <= 1.0 ? -0.25*(M - 1.0)*(M - 1.0) : 0.5*(M - s_abs(M)); }
template <class S> __device__ __forceinline__ S pres_p(S M, S p) { return s_val(s_abs(M)) <= 1.0 ? 0.25*p*(M + 1.0)*(M + 1.0)*(2.0 - M) : 0.5*p*(M + s_abs(M))*s_rcp(M); }
template <class S> __device__ __forceinline__ S pres_m(S M, S p) { return s_val(s_abs(M)) <= 1.0 ? 0.25*p*(M - 1.0)*(M - 1.0)*(2.0 + M) : 0.5*p*(M - s_abs(M))*s_rcp(M); }

template <class S>
__device__ __forceinline__ void ausm_flux(double nx, double ny, S rlft, S ulft, S vlft, S plft,
                                          S rrht, S urht, S vrht, S prht, S* f) {
    const double nn = nx*nx + ny*ny, dsi = rsqrt_fast(nn), ds = nn*dsi;
    S uln = (ulft*nx + vlft*ny)*dsi, urn = (urht*nx + vrht*ny)*dsi;
    S rlfti = s_rcp(rlft), rrhti = s_rcp(rrht);
    S alft = s_sqrt(GAMMA*plft*rlfti), arht = s_sqrt(GAMMA*prht*rrhti);
    S machlft = uln*s_rcp(alft), machrht = urn*s_rcp(arht);
    S uvl = 0.5*(ulft*ulft + vlft*vlft), elft = plft*OGM1 + rlft*uvl, hlft = (elft + plft)*rlfti;
    S uvr = 0.5*(urht*urht + vrht*vrht), erht = prht*OGM1 + rrht*uvr, hrht = (erht + prht)*rrhti;
    S mach_half = mach_p(machlft) + mach_m(machrht);
    S p_half = pres_p(machlft, plft) + pres_m(machrht, prht);
    if (s_val(mach_half) >= 0.0) {
        S m = rlft*alft*mach_half*ds;
        f[0] = m; f[1] = m*ulft + p_half*nx; f[2] = m*vlft + p_half*ny; f[3] = m*hlft;
    } else {
        S m = rrht*arht*mach_half*ds;
        f[0] = m; f[1] = m*urht + p_half*nx; f[2] = m*vrht + p_half*ny; f[3] = m*hrht;
    }
}

// ---- DiffusiveFluxGreenGauss::evaluate, src/model/flux.cpp:12-48 ----------------------------------------
// components 1..3 (component 0 is the constant 0.0, :42)
template <class S>
__device__ __forceinline__ void viscous_flux(double nx, double ny, S dudx, S dudy, S dvdx, S dvdy, S dTdx, S dTdy,
                                             S ubar, S vbar, S mu, S k, S* f) {
    S div = dudx + dvdy;
    S tau_xy = mu*(dudy + dvdx);
    S tau_xx = mu*(2.0*dudx - K23*div);
    S tau_yy = mu*(2.0*dvdy - K23*div);
    S q_x = -(k*dTdx), q_y = -(k*dTdy);
    f[1] = tau_xx*nx + tau_xy*ny;
    f[2] = tau_xy*nx + tau_yy*ny;
    f[3] = (ubar*tau_xx + vbar*tau_xy - q_x)*nx + (ubar*tau_xy + vbar*tau_yy - q_y)*ny;
}

// ---- SA per-cell closures (extension; normative CPU statement: oracle/port/structured_port.hpp) ------
template <class S>
__device__ __forceinline__ S sa_fv1(S chi) { S c3 = chi*chi*chi; return c3*s_rcp(c3 + SA_CV1*SA_CV1*SA_CV1); }

template <class S>
__device__ __forceinline__ S sa_source(S rho, S nut, S mul, S om, S dndx, S dndy, double d, double beta) {
    const double k2 = SA_KAPPA*SA_KAPPA;
    const double cw1 = SA_CB1/k2 + (1.0 + SA_CB2)/SA_SIGMA;
    const double cw36 = 64.0;   // cw3^6
    S chi = rho*nut*s_rcp(mul);
    S fv1 = sa_fv1(chi);
    S fv2 = 1.0 - chi*s_rcp(1.0 + chi*fv1);
    const double k2d2 = k2*d*d;
    const double id = rcp_fast(d);       // one reciprocal of the wall distance serves 1/d and 1/(kappa d)^2
    S sbar = nut*fv2*(id*id*(1.0/k2));
    S st = om + sbar, st_min = 0.3*om;
    if (s_val(st) < s_val(st_min)) st = st_min;
    S den = st*k2d2;
    if (s_val(den) < 1e-30) den = S(1e-30);
    S r = nut*s_rcp(den);
    if (s_val(r) > 10.0) r = S(10.0);
    S r2 = r*r, r6 = r2*r2*r2;
    S gg = r + SA_CW2*(r6 - r);
    S g2 = gg*gg, g6 = g2*g2*g2;
    S fw = gg*s_pow16((1.0 + cw36)*s_rcp(g6 + cw36));
    S nd = nut*id;
    return rho*(beta*SA_CB1*st*nut - cw1*fw*nd*nd) + (SA_CB2/SA_SIGMA)*rho*(dndx*dndx + dndy*dndy);
}

// double-only form of sa_source for the residual kernel, given mu_t = rho nu~ f_v1 (already in the kernel's ring):
// X f_v1 = mu_t/mu, so f_v2 = 1 - X/(1 + X f_v1) = 1 - rho nu~/(mu + mu_t) needs one reciprocal instead of three, and
// the sixth root is taken division-free (see below).
// Same formulas as the template above; differs from it by rounding only.
__device__ __forceinline__ double sa_source_mut(double rho, double nut, double mul, double mut, double om, double dndx, double dndy,
                                                double d, double beta) {
    const double k2 = SA_KAPPA*SA_KAPPA;
    const double cw1 = SA_CB1/k2 + (1.0 + SA_CB2)/SA_SIGMA;
    const double cw36 = 64.0;   // cw3^6
    const double fv2 = fma(-rho*nut, rcp_fast(mul + mut), 1.0);     // X/(1 + X f_v1) = rho nu~/(mu + mu_t)
    const double k2d2 = k2*d*d;
    const double id = rcp_fast(d);
    const double sbar = nut*fv2*(id*id*(1.0/k2));
    double st = om + sbar;
    const double st_min = 0.3*om;
    if (st < st_min) st = st_min;
    double den = st*k2d2;
    if (den < 1e-30) den = 1e-30;
    double r = nut*rcp_fast(den);
    if (r > 10.0) r = 10.0;
    const double r2 = r*r, r6 = r2*r2*r2;
    const double gg = r + SA_CW2*(r6 - r);
    const double g2 = gg*gg, g6 = g2*g2*g2;
    // fw = g ((1 + cw3^6)/z)^(1/6), z = g^6 + cw3^6 in [64, 1e33]: y = z^(-1/6) from a float seed and one division-free
    // third-order step y0 (1 + e/6 + 7 e^2/72), e = 1 - z y0^6 (truncation 0.07 e^3 ~ 1e-17) -- no reciprocal in the chain
    const double z = g6 + cw36;
    const double y0 = pow_seed(z, -0.166666667f);
    const double y02 = y0*y0;
    const double e = fma(-z, y02*y02*y02, 1.0);
    const double pe = fma(7.0/72.0, e, 1.0/6.0)*e;
    const double y = 2.0051747451504216*fma(y0, pe, y0);       // 65^(1/6)
    const double fw = gg*y;
    const double nd = nut*id;
    return rho*(beta*SA_CB1*st*nut - cw1*fw*nd*nd) + (SA_CB2/SA_SIGMA)*rho*(dndx*dndx + dndy*dndy);
}

} // namespace sg
