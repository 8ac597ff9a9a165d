// K-JAC, single pass: the block-stencil Jacobian d rhs / d q of calc_residual(q, lhs = true)
// (the reference's trace_on .. sparse_jac, src/solver/solver.cpp:72-90,156) written ONCE, complete, with no per-face
// scratch in HBM.
//
// A CTA owns a strip of 31 cell columns (+ the chi face that closes it: 32 lanes) and marches a chunk of rows upward.
// What a face contributes to the 13 stencil blocks of its two cells factors as
//     d D_face / d q_s  =  [ face "core" ]  x  [ per-cell chain d(W, z)_s / d q_s ]
// where the core -- dF/d(ql, qr) of the flux function (forward-mode duals through the SAME roe_flux / ausm_flux the
// residual kernel runs), the limiter derivatives, the viscous-flux coefficients -- is 68 doubles per face.  Cores live in
// SHARED MEMORY only: the chi core of a row is exchanged between neighbouring lanes, the eta core of the row below is kept
// in a two-row ring.  The on-chip state is per CELL (3 cores + a ring of per-cell primitives, ~2.3 KB), so occupancy is
// bought with THREADS PER CELL: the four warps of a CTA work on the same 31 cells, lanes = columns (coalesced), warps =
// tasks --
//   phase A  (cores)     warp 0: chi limiter derivatives + flux passes 0,1     warp 1: chi flux passes 2,3 + viscous coefficients
//                        warp 2: eta limiter derivatives + flux passes 0,1     warp 3: eta flux passes 2,3 + viscous coefficients
//   phase B  (assembly)  the 13 slots are split over the warps; per slot and equation row the coefficient vectors of the
//                        <= 4 contributing faces are summed first and the per-cell chain is applied once; every J entry
//                        is stored exactly once (streaming stores).  The warps also convert the next rows entering the
//                        rings and prepare the SA source sensitivities of the next row.
// Two __syncthreads per row; three CTAs (12 warps) per SM.  Ghost cells are independent slots here; jac_fold_kernel
// folds them into the interior cells they are functions of afterwards (boundary band only).
#pragma once
#include "jacobian_kernel.cuh"

namespace sg {

constexpr int JM_CELLS = 31;            // cells per strip; lane 31 only owns the chi face that closes the strip
#ifndef JM_NWARPS
#define JM_NWARPS 4
#endif
constexpr int JM_WARPS = JM_NWARPS;     // 4 face-core warps; 5 = one more that prepares the rings / SA sensitivities during phase A (measured slower: 128 registers, 33.9 vs 24.5 ms)
constexpr int JM_RC = 36;               // ring columns: cells i0-2 .. i0+33

// viscous part of a core.  JM_RICH: the face-only products of the viscous-flux coefficients and the geometry weights of the four
// cell classes are formed ONCE per face in phase A (26 values) instead of per (slot, face) pair in phase B (11 raw values + 4 plane
// loads): 83 doubles per core, 2 CTAs per SM (measured equal to 3 for this kernel).
#ifndef JM_RICH
#define JM_RICH 1
#endif
#if JM_RICH
enum { JV_UB = 0, JV_VB, JV_GMU1, JV_GMU2, JV_GMU3, JV_WX0, JV_WY0, JV_WX1, JV_WY1, JV_WXP, JV_WYP, JV_WXM, JV_WYM,           // half 0
       JV_MU, JV_A1, JV_A2, JV_A3, JV_A4, JV_A5, JV_A6, JV_K1, JV_K2, JV_GK3, JV_GN, JV_MSX, JV_MSY, JV_COUNT };              // half 1
#else
enum { JV_MU = 0, JV_KK, JV_UB, JV_VB, JV_TXX, JV_TYY, JV_TXY, JV_TX, JV_TY, JV_GN, JV_MUSA, JV_COUNT };
#endif

template <int V> struct IC { static constexpr int value = V; };   // compile-time int passed through generic lambdas

template <int NV, int ORDER, bool VISC> struct JmCfg {
    static constexpr bool SA = NV > 4;
    static constexpr int NWV = 6;                                   // ring W: rho, u, v, p, 1/rho, rho nu~
    static constexpr int NZV = VISC ? (SA ? 6 : 3) : 1;             // ring Z: T, mu, dmu/dT [, mu_t, c_mt, dmu_t/dq4]
    static constexpr int C_FD = 0, C_F0 = 32, C_DL = 33, C_DR = C_DL + (ORDER == 2 ? 12 : 0), C_V = C_DR + (ORDER == 2 ? 12 : 0);
    static constexpr int CORE = C_V + (VISC ? JV_COUNT : 0);
    static constexpr int W_DBL = 6*NWV*JM_RC, Z_DBL = 4*NZV*JM_RC, C_DBL = CORE*32, S_DBL = SA ? 2*7*32 : 0;
    static constexpr int G_DBL = 2*10*32, M8_DBL = 9*32;                // staged face-geometry weights (chi, eta) and cell metrics of a row
    static constexpr size_t smem_bytes = sizeof(double)*(size_t)(W_DBL + Z_DBL + 3*C_DBL + S_DBL + G_DBL + M8_DBL + 2);
};
enum { JW_R = 0, JW_U, JW_V, JW_P, JW_RI, JW_RN };
enum { JZ_T = 0, JZ_MU, JZ_DMUDT, JZ_MUT, JZ_CMT, JZ_DMUT4 };

enum { JF_C0 = 0, JF_C1 = 1, JF_E0 = 2, JF_E1 = 3 };   // chi face i (-), chi face i+1 (+), eta face j (-), eta face j+1 (+)
enum { JC_NONE = -1, JC_D0 = 0, JC_D1 = 1, JC_P = 2, JC_M = 3 };

// role of the stencil cell at offset (dx, dy) from the row cell in each of its four faces
__host__ __device__ constexpr int jm_line_role(int f, int dx, int dy) {       // 0:LL 1:L 2:R 3:RR, -1: not on the face's line
    return f == JF_C0 ? ((dy == 0 && dx >= -2 && dx <= 1) ? dx + 2 : -1)
         : f == JF_C1 ? ((dy == 0 && dx >= -1 && dx <= 2) ? dx + 1 : -1)
         : f == JF_E0 ? ((dx == 0 && dy >= -2 && dy <= 1) ? dy + 2 : -1)
         :              ((dx == 0 && dy >= -1 && dy <= 2) ? dy + 1 : -1);
}
__host__ __device__ constexpr int jm_dual_class(int f, int dx, int dy) {
    // chi faces: D0/D1 = the two cells of the face's row, P = row above, M = row below (src/utils/mesh.cpp:44-53)
    // eta faces: D0/D1 = the two cells of the face's column, P = column to the right, M = column to the left (:93-98)
    const int a = (f == JF_C0 || f == JF_C1) ? dx - (f == JF_C1 ? 1 : 0) : dy - (f == JF_E1 ? 1 : 0);   // -1: D0 side, 0: D1 side
    const int b = (f == JF_C0 || f == JF_C1) ? dy : dx;                                                  // 0: direct, +1: P, -1: M
    if (a != -1 && a != 0) return JC_NONE;
    return b == 0 ? (a == -1 ? JC_D0 : JC_D1) : (b == 1 ? JC_P : (b == -1 ? JC_M : JC_NONE));
}

// Static face geometry of the viscous dual cells (src/utils/mesh.cpp:54-82,99-126), evaluated ONCE per grid: the Green-Gauss
// gradient and the face average are linear in the six cells of the dual cell, so everything the Jacobian needs from the
// metrics is, per face, its normal and the weight (wx, wy) with which a cell of each class enters d/dx, d/dy:
//   D0 / D1 (the face's own two cells), P / M (the two cells completing each vertex average).
enum { JG_NX = 0, JG_NY, JG_XD0, JG_YD0, JG_XD1, JG_YD1, JG_XP, JG_YP, JG_XM, JG_YM, JG_N };

// packed roles of slot s: bits 0-2 dx+2, 3-5 dy+2, 6-8 number of contributing faces, then one byte per CONTRIBUTING face from
// bit 16: face id (2 bits), line role + 1 (3 bits), dual class + 1 (3 bits)
__host__ __device__ constexpr unsigned long long jm_desc(int s, bool visc) {
    const int dxs[13] = {0, -1, 1, 0, 0, -1, 1, -1, 1, -2, 2, 0, 0}, dys[13] = {0, 0, 0, -1, 1, -1, -1, 1, 1, 0, 0, -2, 2};   // = c_slot_dx / c_slot_dy
    if (s < 0 || s > 12) return 0ull;
    unsigned long long d = (unsigned long long)(dxs[s] + 2) | ((unsigned long long)(dys[s] + 2) << 3);
    int n = 0;
    for (int f = 0; f < 4; f++) {
        const int lr = jm_line_role(f, dxs[s], dys[s]), dc = visc ? jm_dual_class(f, dxs[s], dys[s]) : JC_NONE;
        if (lr < 0 && dc == JC_NONE) continue;
        d |= ((unsigned long long)f | ((unsigned long long)(lr + 1) << 2) | ((unsigned long long)(dc + 1) << 5)) << (16 + 8*n);
        n++;
    }
    return d | ((unsigned long long)n << 6);
}
// phase B task order (most expensive first; the warps pull tasks from a shared counter): slot 0, the four edges, the SA
// preparation of the next row (13), the corners, the ring rows entering (14: W, 15: Z), the arms
constexpr unsigned long long JM_TASKS = JM_NWARPS > 4 ? 0x000CBA9876543210ull : 0xCBA9FE8765D43210ull;      // nibble n = task n
constexpr int JM_NTASK = JM_NWARPS > 4 ? 13 : 16;                  // with a fifth warp tasks 13..15 are its phase-A work

struct JmParams {
    View v; Gas g; Metrics m;
    const double* q; double* J;
    const double* wdist; const double* beta;
    const double* gchi; const double* geta;      // [JG_N][plane]: chi face i at (row(j), col(i)), eta face at (row(face j), col(i))
    double eps_chi, eps_eta;
    int nstrips, nchunks, rpc;
};

template <bool VISC>
__global__ void jac_geom_kernel(View v, Metrics m, double* __restrict__ gchi, double* __restrict__ geta) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= v.pitch || r >= v.rows) return;
    const int i = c - IOFF, jl = r - JOFF;
    const size_t o = v.at(r, c), pl = v.plane;
    auto put = [&](double* G, const FaceGeom& fg) {
        G[JG_NX*pl + o] = fg.nx; G[JG_NY*pl + o] = fg.ny;
        if (VISC) {
            const double iv = fg.ivol2, qx = 0.25*(fg.tx - fg.bx), qy = 0.25*(fg.ty - fg.by);
            G[JG_XD0*pl + o] = (qx - fg.lx)*iv; G[JG_YD0*pl + o] = (qy - fg.ly)*iv;
            G[JG_XD1*pl + o] = (qx + fg.rx)*iv; G[JG_YD1*pl + o] = (qy + fg.ry)*iv;
            G[JG_XP*pl + o] = 0.25*fg.tx*iv; G[JG_YP*pl + o] = 0.25*fg.ty*iv;
            G[JG_XM*pl + o] = -0.25*fg.bx*iv; G[JG_YM*pl + o] = -0.25*fg.by*iv;
        }
    };
    FaceGeom fg;
    if (i >= 0 && i <= v.nic && jl >= 0 && jl < v.njl) {               // chi face i of cell row jl (same clamped variants as the residual kernel)
        fg.nx = m.ncx[o]; fg.ny = m.ncy[o];
        if (VISC) {
            const int ca = imax(i - 1, 0) + IOFF, cb = imin(i, v.nic - 1) + IOFF;
            const int cR = imin(i + 1, v.ni - 1) + IOFF, cL = imax(i - 1, 0) + IOFF;
            fg.tx = m.nex[v.at(r + 1, ca)] + m.nex[v.at(r + 1, cb)]; fg.ty = m.ney[v.at(r + 1, ca)] + m.ney[v.at(r + 1, cb)];
            fg.bx = m.nex[v.at(r, ca)] + m.nex[v.at(r, cb)]; fg.by = m.ney[v.at(r, ca)] + m.ney[v.at(r, cb)];
            fg.rx = fg.nx + m.ncx[v.at(r, cR)]; fg.ry = fg.ny + m.ncy[v.at(r, cR)];
            fg.lx = fg.nx + m.ncx[v.at(r, cL)]; fg.ly = fg.ny + m.ncy[v.at(r, cL)];
            fg.ivol2 = rcp_fast(m.vol[v.at(r, ca)] + m.vol[v.at(r, cb)]);
        }
        put(gchi, fg);
    }
    if (i >= 0 && i < v.nic && jl >= 0 && jl <= v.njl) {               // eta face with local face row jl
        const int fj = v.j0 + jl;
        fg.nx = m.nex[o]; fg.ny = m.ney[o];
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
        put(geta, fg);
    }
}

template <int NV, int ORDER, int FLUX, bool VISC>
#ifndef JM_MINB
#define JM_MINB (JM_RICH ? 2 : 3)
#endif
__global__ void __launch_bounds__(32*JM_WARPS, JM_MINB) jac_march_kernel(const JmParams prm) {
    using Cfg = JmCfg<NV, ORDER, VISC>;
    constexpr bool SA = Cfg::SA;
    constexpr int NWV = Cfg::NWV, NZV = Cfg::NZV;
    constexpr int NS = ORDER == 2 ? 13 : 9;
    extern __shared__ double smem[];
    double* sW = smem;                               // [6 rows][NWV][JM_RC]
    double* sZ = sW + Cfg::W_DBL;                    // [4 rows][NZV][JM_RC]
    double* sC = sZ + Cfg::Z_DBL;                    // chi cores of the current row   [CORE][32]
    double* sE = sC + Cfg::C_DBL;                    // eta cores, ring of two rows    [2][CORE][32]
    double* sS = sE + 2*Cfg::C_DBL;                  // SA source sensitivities        [2][7][32]
    double* sG = sS + Cfg::S_DBL;                    // face-geometry weights of the row's chi / eta face [2][JG_N][32] (cp.async, one phase ahead)
    double* sM8 = sG + Cfg::G_DBL;                   // the row's cell metrics                            [9][32]     (cp.async, one phase ahead)
    int* sTask = (int*)(sM8 + Cfg::M8_DBL);          // phase B task counter

    const View& v = prm.v; const Gas& g = prm.g; const Metrics& m = prm.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x % prm.nstrips, chunk = blockIdx.x / prm.nstrips;
    const int i0 = strip*JM_CELLS;
    const int ra = chunk*prm.rpc, rb = imin(ra + prm.rpc, v.njl);
    const int i = i0 + lane;                         // own cell column / own chi face
    const size_t pl = v.plane;
    const int cc0 = lane + 2;                        // ring column of the own cell
    const int ic = imin(i, v.nic - 1) + IOFF;        // plane column of the own cell (clamped for the lanes past the grid)

    auto wrow = [&](int jl) { return sW + (((jl % 6) + 6) % 6)*NWV*JM_RC; };
    auto zrow = [&](int jl) { return sZ + ((jl + 8) & 3)*NZV*JM_RC; };

    // ---- ring maintenance -------------------------------------------------------------------------------------
    // primitives of cell row jl, ring column cc, straight from the state planes (FluidModel::primvars, fluid.cpp:50-67)
    auto convert_w = [&](int jl, int cc) {
        const int c = imax(imin(i0 - 2 + cc + IOFF, v.pitch - 1), 0), r = imax(imin(jl + JOFF, v.rows - 1), 0);
        const size_t o = v.at(r, c);
        double rho, u, vv, p, T;
        cons_to_prim<double>(g, __ldg(prm.q + o), __ldg(prm.q + pl + o), __ldg(prm.q + 2*pl + o), __ldg(prm.q + 3*pl + o), rho, u, vv, p, T);
        double* W = wrow(jl);
        W[JW_R*JM_RC + cc] = rho; W[JW_U*JM_RC + cc] = u; W[JW_V*JM_RC + cc] = vv; W[JW_P*JM_RC + cc] = p;
        W[JW_RI*JM_RC + cc] = rcp_fast(rho);
        W[JW_RN*JM_RC + cc] = SA ? __ldg(prm.q + 4*pl + o) : 0.0;
    };
    // T, mu, mu_t and the derivative factors of cell row jl (from ring W)
    auto convert_z = [&](int jl, int cc) {
        if (!VISC) return;
        const double* W = wrow(jl); double* Z = zrow(jl);
        const double ri = W[JW_RI*JM_RC + cc], T = W[JW_P*JM_RC + cc]*ri*g.iR;
        const double cb = cbrt(T*g.iT_ref), mu = g.mu_ref*cb*cb;
        const double dmudT = (2.0/3.0)*mu*rcp_fast(T);
        Z[JZ_T*JM_RC + cc] = T; Z[JZ_MU*JM_RC + cc] = mu; Z[JZ_DMUDT*JM_RC + cc] = dmudT;
        if (SA) {
            const double rn = W[JW_RN*JM_RC + cc];
            const double chi = rn*rcp_fast(mu), c3 = SA_CV1*SA_CV1*SA_CV1, x3 = chi*chi*chi, den = rcp_fast(x3 + c3);
            const double fv1 = x3*den, dfv1 = 3.0*chi*chi*c3*den*den;
            Z[JZ_MUT*JM_RC + cc] = rn*fv1;
            Z[JZ_CMT*JM_RC + cc] = -chi*chi*dfv1*dmudT;
            Z[JZ_DMUT4*JM_RC + cc] = fv1 + chi*dfv1;
        }
    };
    auto convert_w_row = [&](int jl) {
#pragma unroll 1
        for (int cc = lane; cc < JM_RC; cc += 32) convert_w(jl, cc);
    };
    auto convert_z_row = [&](int jl) {
#pragma unroll 1
        for (int cc = lane; cc < JM_RC; cc += 32) convert_z(jl, cc);
    };

    // the CellD of the chain rule (jacobian_kernel.cuh) rebuilt from the rings; full = with the viscous / SA derivative rows
    auto load_cell = [&](int jl, int cc, bool full, CellD<NV>& w) {
        const double* W = wrow(jl);
        w.r = W[JW_R*JM_RC + cc]; w.u = W[JW_U*JM_RC + cc]; w.v = W[JW_V*JM_RC + cc]; w.p = W[JW_P*JM_RC + cc]; w.ri = W[JW_RI*JM_RC + cc];
        w.rn = W[JW_RN*JM_RC + cc]; w.nut = w.rn*w.ri;
        w.T = 0; w.mu = 0; w.mut = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { w.dT[k] = 0.0; w.dmu[k] = 0.0; w.dmut[k] = 0.0; }
        w.dmut[4] = 0.0;
        if (VISC && full) {                                        // (the arms only go through chain_W: u, v, 1/rho)
            const double ke = 0.5*(w.u*w.u + w.v*w.v), s = w.ri*g.iR;
            w.dT[0] = (GM1*ke - w.p*w.ri)*s; w.dT[1] = -GM1*w.u*s; w.dT[2] = -GM1*w.v*s; w.dT[3] = GM1*s;
            const double* Z = zrow(jl);
            w.T = Z[JZ_T*JM_RC + cc]; w.mu = Z[JZ_MU*JM_RC + cc];
            const double dmudT = Z[JZ_DMUDT*JM_RC + cc];
#pragma unroll
            for (int k = 0; k < 4; k++) w.dmu[k] = dmudT*w.dT[k];
            if (SA) {
                w.mut = Z[JZ_MUT*JM_RC + cc];
                const double cmt = Z[JZ_CMT*JM_RC + cc];
#pragma unroll
                for (int k = 0; k < 4; k++) w.dmut[k] = cmt*w.dT[k];
                w.dmut[4] = Z[JZ_DMUT4*JM_RC + cc];
            }
        }
    };

    // ---- phase A: one half of the core of one face -- ONE code path for all four warps (dir and half are warp-uniform
    //      run-time values: four specialised copies starved the instruction cache, ncu no_instruction 7 cycles per issue) ----
    //   dir  = 0: chi face i of cell row jA            -> sC          line cells (jA, i-2 .. i+1)
    //   dir  = 1: eta face with local face row jA      -> sE[jA & 1]  line cells (jA-2 .. jA+1, i)
    //   half = 0: stores the limiter derivatives, flux passes 0,1 (d/d ql), F0;  half = 1: flux passes 2,3 (d/d qr), viscous coefficients
    // ---- staging one phase ahead (round 2b).  ncu attributes 14 % of the kernel's stall samples to long_scoreboard at the FIRST
    //      USE of plain global loads -- the face-geometry weights in phase A, the cell metrics at the head of phase B, the state row
    //      entering the rings -- and with two warps per scheduler nothing hides them.  The weights of the NEXT row's two faces
    //      are copied into shared memory with cp.async during phase B, the row's cell metrics during phase A (each warp a share
    //      of the planes, every thread waits for its own copies ahead of the phase barrier); the state row is prefetched into L1.
    auto cp8 = [&](double* dst, const double* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory"); };
    auto cp_wait = [&]() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); };
    auto pf_l1 = [&](const double* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); };
    auto stage_G = [&](int jl) {                                    // weights of chi face row jl and eta face row jl + 1
        const double* Gc = prm.gchi + v.at(jl + JOFF, imin(i, v.nic) + IOFF);
        const double* Ge = prm.geta + v.at(imin(jl + 1 + JOFF, v.rows - 1), ic);
        for (int k = warp; k < 2*JG_N; k += JM_WARPS) cp8(sG + k*32 + lane, k < JG_N ? Gc + (size_t)k*pl : Ge + (size_t)(k - JG_N)*pl);
    };
    auto stage_M8 = [&](int jl, bool more) {                        // issued at the head of phase A of row jl
        const size_t o = v.at(jl + JOFF, ic), o1 = v.at(jl + JOFF + 1, ic);
        if (warp == 0) { cp8(sM8 + 0*32 + lane, m.ncx + o); cp8(sM8 + 4*32 + lane, m.nex + o); cp8(sM8 + 8*32 + lane, m.vol + o); }
        else if (warp == 1) { cp8(sM8 + 1*32 + lane, m.ncy + o); cp8(sM8 + 5*32 + lane, m.ney + o); }
        else if (warp == 2) { cp8(sM8 + 2*32 + lane, m.ncx + o + 1); cp8(sM8 + 6*32 + lane, m.nex + o1); }
        else if (warp == 3) { cp8(sM8 + 3*32 + lane, m.ncy + o + 1); cp8(sM8 + 7*32 + lane, m.ney + o1); }
        if (more) {                                                 // the state row convert_w_row(jl + 3) reads: into L1
            const int r = imax(imin(jl + 3 + JOFF, v.rows - 1), 0);
            for (int cc = lane; cc < JM_RC; cc += 32) {
                const size_t oq = v.at(r, imax(imin(i0 - 2 + cc + IOFF, v.pitch - 1), 0));
                for (int k = warp; k < NV; k += JM_WARPS) pf_l1(prm.q + (size_t)k*pl + oq);
            }
        }
    };
    auto face_core = [&](int dir, int half, int jA) {
        const int di = dir ? 0 : 1, dj = dir ? 1 : 0;
        const int rLL = dir ? jA - 2 : jA, cLL = dir ? cc0 : cc0 - 2;
        bool Lint, Rint;
        if (dir) { const int fj = v.j0 + jA; Lint = fj - 1 >= 0; Rint = fj <= v.njc - 1; }
        else { Lint = i - 1 >= 0; Rint = i <= v.nic - 1; }
        const double eps = dir ? prm.eps_eta : prm.eps_chi;
        const double* G = sG + dir*JG_N*32 + lane;                 // staged by stage_G one phase earlier
        double* core = dir ? sE + (jA & 1)*Cfg::C_DBL + lane : sC + lane;
        const double nx = G[JG_NX*32], ny = G[JG_NY*32];
        double ql[4], qr[4];
        {
            const double* W0 = wrow(rLL); const double* W1 = wrow(rLL + dj); const double* W2 = wrow(rLL + 2*dj); const double* W3 = wrow(rLL + 3*dj);
#pragma unroll
            for (int k = 0; k < 4; k++) {                              // src/model/reconstruction.cpp:94-111,133-150 by 3-lane duals
                const double wLL = W0[k*JM_RC + cLL], wL = W1[k*JM_RC + cLL + di], wR = W2[k*JM_RC + cLL + 2*di], wRR = W3[k*JM_RC + cLL + 3*di];
                double d[6] = {0, 1, 0, 0, 1, 0};
                ql[k] = wL; qr[k] = wR;
                if (ORDER == 2) {
                    if (half == 0) {                                   // values + derivatives (stored); the other half needs the values only
                        typedef Dual<3> D3;
                        D3 hi, lo, hi2, lo2;
                        { D3 a(wLL), b(wL), c(wR); a.d[0] = 1; b.d[1] = 1; c.d[2] = 1; muscl_cell<D3>(a, b, c, eps, hi, lo); }
                        { D3 a(wL), b(wR), c(wRR); a.d[0] = 1; b.d[1] = 1; c.d[2] = 1; muscl_cell<D3>(a, b, c, eps, hi2, lo2); }
                        if (Lint) { ql[k] = hi.v; d[0] = hi.d[0]; d[1] = hi.d[1]; d[2] = hi.d[2]; }
                        if (Rint) { qr[k] = lo2.v; d[3] = lo2.d[0]; d[4] = lo2.d[1]; d[5] = lo2.d[2]; }
#pragma unroll
                        for (int n = 0; n < 3; n++) { core[(Cfg::C_DL + k*3 + n)*32] = d[n]; core[(Cfg::C_DR + k*3 + n)*32] = d[3 + n]; }
                    } else {
                        double hi, lo, hi2, lo2;
                        muscl_cell<double>(wLL, wL, wR, eps, hi, lo); muscl_cell<double>(wL, wR, wRR, eps, hi2, lo2);
                        if (Lint) ql[k] = hi;
                        if (Rint) qr[k] = lo2;
                    }
                }
            }
        }
        // dF/d(ql, qr): forward-mode passes of two lanes through the flux function the residual kernel runs
        // (one pass of four lanes per half: 696 instructions against 2 x 568 for two passes of two lanes)
        {
            typedef Dual<4> D4;
            const double sl = half ? 0.0 : 1.0, sr = half ? 1.0 : 0.0; // seeds: d/d ql_0..3 (half 0) or d/d qr_0..3 (half 1)
            D4 a[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a[k] = D4(ql[k]); a[4 + k] = D4(qr[k]);
                a[k].d[k] = sl; a[4 + k].d[k] = sr;
            }
            D4 F[4];
            if (FLUX == SGPU_FLUX_ROE) roe_flux<D4>(nx, ny, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], F);
            else ausm_flux<D4>(nx, ny, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], F);
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int l = 0; l < 4; l++) core[(Cfg::C_FD + r*8 + half*4 + l)*32] = F[r].d[l];
            if (half == 0) core[Cfg::C_F0*32] = F[0].v;
        }
        if (VISC) {                                                    // the face aggregates, linear in the six cells (mesh.cpp:10-131)
            // half 0: the velocity part (stress / mu, face velocities); half 1: temperature, viscosities, nu~ (balances the phase)
            const int rD = rLL + dj, cD = cLL + di;                    // D0; D1 = D0 + (dj, di); P / M = one step across the line
            const int pj = dir ? 0 : 1, pi = dir ? 1 : 0;
            const double xD0 = G[JG_XD0*32], yD0 = G[JG_YD0*32], xD1 = G[JG_XD1*32], yD1 = G[JG_YD1*32];
            const double xP = G[JG_XP*32], yP = G[JG_YP*32], xM = G[JG_XM*32], yM = G[JG_YM*32];
            constexpr int NZ = 5;
            double sD0[NZ], sD1[NZ], sP[NZ], sM[NZ];
#pragma unroll
            for (int n = 0; n < NZ; n++) { sP[n] = sM[n] = 0.0; }
#pragma unroll
            for (int n = 0; n < 6; n++) {
                const int rr = rD + (n & 1)*dj + (n < 2 ? 0 : (n < 4 ? pj : -pj)), cc = cD + (n & 1)*di + (n < 2 ? 0 : (n < 4 ? pi : -pi));
                const double* W = wrow(rr); const double* Z = zrow(rr);
                double z[NZ];
                if (half == 0) { z[0] = W[JW_U*JM_RC + cc]; z[1] = W[JW_V*JM_RC + cc]; z[2] = z[3] = z[4] = 0.0; }
                else {
                    const double rn = W[JW_RN*JM_RC + cc];
                    z[0] = Z[JZ_T*JM_RC + cc]; z[1] = Z[JZ_MU*JM_RC + cc]; z[2] = SA ? Z[JZ_MUT*JM_RC + cc] : 0.0;
                    z[3] = SA ? rn*W[JW_RI*JM_RC + cc] : 0.0; z[4] = rn;
                }
#pragma unroll
                for (int k = 0; k < NZ; k++) {
                    if (n == 0) sD0[k] = z[k]; else if (n == 1) sD1[k] = z[k]; else if (n < 4) sP[k] += z[k]; else sM[k] += z[k];
                }
            }
            auto gx = [&](int k) { return xD0*sD0[k] + xD1*sD1[k] + xP*sP[k] + xM*sM[k]; };
            auto gy = [&](int k) { return yD0*sD0[k] + yD1*sD1[k] + yP*sP[k] + yM*sM[k]; };
            auto bar = [&](int k) { return 0.375*(sD0[k] + sD1[k]) + 0.0625*(sP[k] + sM[k]); };
#if JM_RICH
            if (half == 0) {
                const double ux = gx(0), uy = gy(0), vx = gx(1), vy = gy(1);
                const double div = ux + vy;
                const double ub = bar(0), vb = bar(1);
                const double txx_h = 2.0*ux - (2.0/3.0)*div, tyy_h = 2.0*vy - (2.0/3.0)*div, txy_h = uy + vx;   // tau / mu (flux.cpp:36-45)
                core[(Cfg::C_V + JV_UB)*32] = ub; core[(Cfg::C_V + JV_VB)*32] = vb;
                core[(Cfg::C_V + JV_GMU1)*32] = txx_h*nx + txy_h*ny; core[(Cfg::C_V + JV_GMU2)*32] = txy_h*nx + tyy_h*ny;
                core[(Cfg::C_V + JV_GMU3)*32] = nx*(ub*txx_h + vb*txy_h) + ny*(ub*txy_h + vb*tyy_h);
                core[(Cfg::C_V + JV_WX0)*32] = xD0; core[(Cfg::C_V + JV_WY0)*32] = yD0; core[(Cfg::C_V + JV_WX1)*32] = xD1; core[(Cfg::C_V + JV_WY1)*32] = yD1;
                core[(Cfg::C_V + JV_WXP)*32] = xP; core[(Cfg::C_V + JV_WYP)*32] = yP; core[(Cfg::C_V + JV_WXM)*32] = xM; core[(Cfg::C_V + JV_WYM)*32] = yM;
            } else {
                const double mub = bar(1), mutb = SA ? bar(2) : 0.0;
                const double mu = mub + mutb, kk = SA ? (mub*g.cp_over_pr + mutb*g.cp_over_prt) : mub*g.cp_over_pr;
                const double c43 = 4.0/3.0*mu, c23 = 2.0/3.0*mu;
                core[(Cfg::C_V + JV_MU)*32] = mu;
                core[(Cfg::C_V + JV_A1)*32] = c43*nx; core[(Cfg::C_V + JV_A2)*32] = mu*ny; core[(Cfg::C_V + JV_A3)*32] = -c23*nx;
                core[(Cfg::C_V + JV_A4)*32] = -c23*ny; core[(Cfg::C_V + JV_A5)*32] = mu*nx; core[(Cfg::C_V + JV_A6)*32] = c43*ny;
                core[(Cfg::C_V + JV_K1)*32] = kk*nx; core[(Cfg::C_V + JV_K2)*32] = kk*ny;
                core[(Cfg::C_V + JV_GK3)*32] = nx*gx(0) + ny*gy(0);
                const double musa_s = SA ? (mub + bar(4))*(1.0/SA_SIGMA) : 0.0;
                core[(Cfg::C_V + JV_GN)*32] = SA ? (gx(3)*nx + gy(3)*ny)*(1.0/SA_SIGMA) : 0.0;
                core[(Cfg::C_V + JV_MSX)*32] = musa_s*nx; core[(Cfg::C_V + JV_MSY)*32] = musa_s*ny;
            }
#else
            if (half == 0) {
                const double ux = gx(0), uy = gy(0), vx = gx(1), vy = gy(1);
                const double div = ux + vy;
                core[(Cfg::C_V + JV_UB)*32] = bar(0); core[(Cfg::C_V + JV_VB)*32] = bar(1);
                core[(Cfg::C_V + JV_TXX)*32] = 2.0*ux - (2.0/3.0)*div; core[(Cfg::C_V + JV_TYY)*32] = 2.0*vy - (2.0/3.0)*div;   // tau / mu
                core[(Cfg::C_V + JV_TXY)*32] = uy + vx;
            } else {
                const double mub = bar(1), mutb = SA ? bar(2) : 0.0;
                core[(Cfg::C_V + JV_MU)*32] = mub + mutb;
                core[(Cfg::C_V + JV_KK)*32] = SA ? (mub*g.cp_over_pr + mutb*g.cp_over_prt) : mub*g.cp_over_pr;
                core[(Cfg::C_V + JV_TX)*32] = gx(0); core[(Cfg::C_V + JV_TY)*32] = gy(0);
                core[(Cfg::C_V + JV_GN)*32] = SA ? (gx(3)*nx + gy(3)*ny)*(1.0/SA_SIGMA) : 0.0;
                core[(Cfg::C_V + JV_MUSA)*32] = SA ? (mub + bar(4))*(1.0/SA_SIGMA) : 0.0;
            }
#endif
        }
    };

    // ---- SA source sensitivities of cell (i, jl) -> sS[jl & 1]: dS/d(rho, nut, mu, om, dndx, dndy) and sign(dvdx - dudy) ----
    // weight of the stencil cell (dx, dy) in the Green-Gauss gradient over the cell's own four faces with face values
    // 3/8 (the two cells of the face) + 1/16 (the four cells completing its vertex averages)
    struct Met8 { double cxl, cyl, cxr, cyr, exb, eyb, ext, eyt, Vi; };
    auto load_met8 = [&](int jl, Met8& M) {
        const int r = jl + JOFF;
        const size_t o = v.at(r, ic);
        M.cxl = __ldg(m.ncx + o); M.cyl = __ldg(m.ncy + o); M.cxr = __ldg(m.ncx + o + 1); M.cyr = __ldg(m.ncy + o + 1);
        M.exb = __ldg(m.nex + o); M.eyb = __ldg(m.ney + o); M.ext = __ldg(m.nex + v.at(r + 1, ic)); M.eyt = __ldg(m.ney + v.at(r + 1, ic));
        M.Vi = 1.0/__ldg(m.vol + o);
    };
    auto sa_w = [&](const Met8& M, int dx, int dy, double& wx, double& wy) {
        const double wc = dy == 0 ? 0.375 : 0.0625, we = dx == 0 ? 0.375 : 0.0625;     // chi faces touch dx in {-1,0} / {0,1}; eta faces dy likewise
        const double ax = wc*((dx >= 0 ? M.cxr : 0.0) - (dx <= 0 ? M.cxl : 0.0)) + we*((dy >= 0 ? M.ext : 0.0) - (dy <= 0 ? M.exb : 0.0));
        const double ay = wc*((dx >= 0 ? M.cyr : 0.0) - (dx <= 0 ? M.cyl : 0.0)) + we*((dy >= 0 ? M.eyt : 0.0) - (dy <= 0 ? M.eyb : 0.0));
        wx = ax*M.Vi; wy = ay*M.Vi;
    };
    auto sa_prep = [&](int jl) {
        if (!SA) return;
        Met8 M; load_met8(jl, M);
        double dvdx = 0, dudy = 0, dndx = 0, dndy = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                double wx, wy; sa_w(M, dx, dy, wx, wy);
                const double* W = wrow(jl + dy); const int cc = cc0 + dx;
                const double uu = W[JW_U*JM_RC + cc], vv = W[JW_V*JM_RC + cc], nn = W[JW_RN*JM_RC + cc]*W[JW_RI*JM_RC + cc];
                dvdx += wx*vv; dudy += wy*uu; dndx += wx*nn; dndy += wy*nn;
            }
        const double aa = dvdx - dudy;
        const double* W = wrow(jl);
        const double rho = W[JW_R*JM_RC + cc0], nut = W[JW_RN*JM_RC + cc0]*W[JW_RI*JM_RC + cc0];
        const double mu = VISC ? zrow(jl)[JZ_MU*JM_RC + cc0] : g.mu_ref;
        typedef Dual<6> D6;
        D6 a_rho(rho), a_nut(nut), a_mu(mu), a_om(fabs(aa)), a_nx(dndx), a_ny(dndy);
        a_rho.d[0] = 1; a_nut.d[1] = 1; a_mu.d[2] = 1; a_om.d[3] = 1; a_nx.d[4] = 1; a_ny.d[5] = 1;
        const size_t o = v.at(jl + JOFF, ic);
        const D6 S = sa_source<D6>(a_rho, a_nut, a_mu, a_om, a_nx, a_ny, __ldg(prm.wdist + o), __ldg(prm.beta + o));
        double* out = sS + (jl & 1)*7*32 + lane;
#pragma unroll
        for (int k = 0; k < 6; k++) out[k*32] = S.d[k];
        out[6*32] = aa < 0.0 ? -1.0 : 1.0;
    };

    // ---- phase B: one stencil slot of the row cell (i, jl); s, and with it every branch below, is warp-uniform --------
    const bool cell_ok = lane < JM_CELLS && i < v.nic;
    const unsigned long long my_desc = jm_desc(lane, VISC);   // lane s holds the role descriptor of slot s: fetched by shuffle, no table load
    auto assemble_slot = [&](int s, int jl, const Met8& M8) {
        const unsigned long long desc = __shfl_sync(0xffffffffu, my_desc, s);
        const int DX = (int)(desc & 7ull) - 2, DY = (int)((desc >> 3) & 7ull) - 2, nface = (int)((desc >> 6) & 7ull);
        const bool inner = DX >= -1 && DX <= 1 && DY >= -1 && DY <= 1, corner = DX != 0 && DY != 0;
        const size_t o = v.at(jl + JOFF, ic);
        double* Jp = prm.J + ((size_t)s*NV*NV)*pl + o;
        if (!VISC && corner) {                                         // corners exist through the viscous stencil only
            if (cell_ok) for (int e = 0; e < NV*NV; e++) __stcs(Jp + e*pl, 0.0);
            return;
        }
        const double Vi = M8.Vi;
        if (!inner) {
            // radius-2 arm: ONE face, the cell is its LL or RR line cell -- only the reconstruction chain of (rho, u, v, p), so
            // the block is d(-F)/d(ql or qr) . diag(limiter derivative) . dW/dq with an empty q4 column (3.1 k -> cycles of the
            // generic path spent mostly on its bookkeeping)
            const unsigned ent = (unsigned)(desc >> 16) & 255u;
            const int f = (int)(ent & 3u), lr = (int)((ent >> 2) & 7u) - 1;
            const double* core = (f < 2) ? sC + lane + f : sE + ((jl + f) & 1)*Cfg::C_DBL + lane;
            const double sc = (f & 1) ? Vi : -Vi;
            const double* Fd = core + (Cfg::C_FD + (lr == 0 ? 0 : 4))*32;          // LL: d/d ql, RR: d/d qr
            const double* dp = core + (lr == 0 ? Cfg::C_DL : Cfg::C_DR + 2)*32;    // LL: dl[k][0], RR: dr[k][2]
            double nut_up = 0.0;
            if (SA) {
                const bool upL = core[Cfg::C_F0*32] >= 0.0;
                const int st = upL ? 0 : 1;
                const int udx = (f < 2) ? (f == JF_C0 ? -1 : 0) + st : 0, udy = (f < 2) ? 0 : (f == JF_E0 ? -1 : 0) + st;
                const double* W = wrow(jl + udy);
                nut_up = W[JW_RN*JM_RC + cc0 + udx]*W[JW_RI*JM_RC + cc0 + udx];
            }
            const double* W = wrow(jl + DY); const int cc = cc0 + DX;
            const double u = W[JW_U*JM_RC + cc], vv = W[JW_V*JM_RC + cc], ri = W[JW_RI*JM_RC + cc];
            const double ke = 0.5*(u*u + vv*vv);
            double dk[4];
#pragma unroll
            for (int k = 0; k < 4; k++) dk[k] = ORDER == 2 ? -sc*dp[k*3*32] : 0.0;
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double cwr[4];
#pragma unroll
                for (int k = 0; k < 4; k++) cwr[k] = Fd[((r < 4 ? r : 0)*8 + k)*32]*dk[k]*(r < 4 ? 1.0 : nut_up);
                // chain_W (jacobian_kernel.cuh): d(rho, u, v, p)/dq of the arm cell
                const double o0 = cwr[0] - (cwr[1]*u + cwr[2]*vv)*ri + cwr[3]*GM1*ke;
                const double o1 = cwr[1]*ri - cwr[3]*GM1*u, o2 = cwr[2]*ri - cwr[3]*GM1*vv, o3 = cwr[3]*GM1;
                if (cell_ok) {
                    double* Jr = Jp + (size_t)(r*NV)*pl;
                    __stcs(Jr, o0); __stcs(Jr + pl, o1); __stcs(Jr + 2*pl, o2); __stcs(Jr + 3*pl, o3);
                    if (NV > 4) __stcs(Jr + 4*pl, 0.0);
                }
            }
            return;
        }
        double cw[5][4], cu[4], cv[4], cT3 = 0.0, cmu[5], cmut[4], cnut4 = 0.0, ex0 = 0.0, ex4 = 0.0;
#pragma unroll
        for (int r = 0; r < 5; r++) {
#pragma unroll
            for (int k = 0; k < 4; k++) cw[r][k] = 0.0;
            cmu[r] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) { cu[r] = cv[r] = cmut[r] = 0.0; }
        const double nut_s = wrow(jl + DY)[JW_RN*JM_RC + cc0 + DX]*wrow(jl + DY)[JW_RI*JM_RC + cc0 + DX], ri_s = wrow(jl + DY)[JW_RI*JM_RC + cc0 + DX];
        // geometry of the (up to four) viscous contributions first: 16 independent loads in flight instead of an L2 round trip
        // at the head of every face (the shared-memory carve-out leaves these planes almost no L1)
        // (the face loop stays ROLLED: unrolled, the kernel outgrows the instruction cache -- 7.5 k instructions, ncu
        //  no_instruction 2.8 cycles per issue against 0.3 at 5.7 k; it runs over the CONTRIBUTING faces only, and the geometry
        //  of the next one is requested while the current one is processed: the shared-memory carve-out leaves these planes
        //  almost no L1)
        double n_nx = 0.0, n_ny = 0.0, n_wx = 0.0, n_wy = 0.0;
        auto geom_req = [&](unsigned ent) {
            const int f = (int)(ent & 3u), dcf = (int)((ent >> 5) & 7u) - 1;
            if (dcf < 0) return;
            const double* G = (f < 2) ? prm.gchi + v.at(jl + JOFF, imin(i + f, v.nic) + IOFF) : prm.geta + v.at(jl + (f - 2) + JOFF, ic);
            n_nx = __ldg(G + JG_NX*pl); n_ny = __ldg(G + JG_NY*pl);
            n_wx = __ldg(G + (JG_XD0 + 2*dcf)*pl); n_wy = __ldg(G + (JG_YD0 + 2*dcf)*pl);
        };
        unsigned ents = (unsigned)(desc >> 16);
        if (VISC && !JM_RICH) geom_req(ents & 255u);
#pragma unroll 1
        for (int e = 0; e < nface; e++) {
            const unsigned ent = ents & 255u; ents >>= 8;
            const int f = (int)(ent & 3u), lr = (int)((ent >> 2) & 7u) - 1, dc = (int)((ent >> 5) & 7u) - 1;
            const double g_nx = n_nx, g_ny = n_ny, g_wx = n_wx, g_wy = n_wy;
            if (VISC && !JM_RICH && e + 1 < nface) geom_req(ents & 255u);
            const double* core = (f < 2) ? sC + lane + f : sE + ((jl + f) & 1)*Cfg::C_DBL + lane;      // C0, C1 | E0 (face row jl), E1 (jl + 1)
            const double sc = (f & 1) ? Vi : -Vi;
            if (lr >= 0) {                                             // D = -F: reconstruction chain, line cells LL L | R RR
                const int il = lr <= 2 ? lr : -1, ir = lr >= 1 ? lr - 1 : -1;      // which dl[k][.] / dr[k][.] belongs to this cell
                double nut_up = 0.0;
                if (SA) {
                    const double F0 = core[Cfg::C_F0*32]; const bool upL = F0 >= 0.0;
                    // the upwind cell of the face relative to the row cell: L = (-1,0) C0, (0,0) C1/E1, (0,-1) E0; R = L + one step
                    const int st = upL ? 0 : 1;
                    const int udx = (f < 2) ? (f == JF_C0 ? -1 : 0) + st : 0, udy = (f < 2) ? 0 : (f == JF_E0 ? -1 : 0) + st;
                    const double* W = wrow(jl + udy);
                    nut_up = W[JW_RN*JM_RC + cc0 + udx]*W[JW_RI*JM_RC + cc0 + udx];
                    if ((lr == 1 && upL) || (lr == 2 && !upL)) {       // d(-F0 nut_up)/d q of the upwind cell: (+F0 nut/rho, ., ., ., -F0/rho)
                        ex0 += sc*(F0*nut_s*ri_s);
                        ex4 -= sc*(F0*ri_s);
                    }
                }
                const double* dlp = core + (Cfg::C_DL + imax(il, 0))*32; const double* drp = core + (Cfg::C_DR + imax(ir, 0))*32;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double dlk = (il == 1) ? 1.0 : 0.0, drk = (ir == 1) ? 1.0 : 0.0;
                    if (ORDER == 2) {
                        dlk = il >= 0 ? dlp[k*3*32] : 0.0;
                        drk = ir >= 0 ? drp[k*3*32] : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const double c = core[(Cfg::C_FD + r*8 + k)*32]*dlk + core[(Cfg::C_FD + r*8 + 4 + k)*32]*drk;
                        cw[r][k] -= sc*c;
                        if (SA && r == 0) cw[4][k] -= sc*nut_up*c;
                    }
                }
            }
#if JM_RICH
            if (dc != JC_NONE) {                                       // viscous flux: the dual cell's six cells (flux.cpp:12-48 through mesh.cpp:10-131)
                const double* cv_ = core + Cfg::C_V*32;
                const double wx = sc*cv_[(JV_WX0 + 2*dc)*32], wy = sc*cv_[(JV_WY0 + 2*dc)*32], wb = sc*(dc <= JC_D1 ? 0.375 : 0.0625);
                const double mu = cv_[JV_MU*32], ub = cv_[JV_UB*32], vb = cv_[JV_VB*32];
                const double A1 = cv_[JV_A1*32], A2 = cv_[JV_A2*32], A3 = cv_[JV_A3*32], A4 = cv_[JV_A4*32], A5 = cv_[JV_A5*32], A6 = cv_[JV_A6*32];
                const double gmu1 = cv_[JV_GMU1*32], gmu2 = cv_[JV_GMU2*32], gmu3 = cv_[JV_GMU3*32], gk3 = cv_[JV_GK3*32];
                cu[1] += A1*wx + A2*wy;                    cv[1] += A2*wx + A3*wy;
                cu[2] += A4*wx + A5*wy;                    cv[2] += A5*wx + A6*wy;
                const double b2 = A5*vb + A2*ub;                       // mu (nx vb + ny ub)
                cu[3] += (A1*ub + A4*vb)*wx + b2*wy + (mu*gmu1)*wb;
                cv[3] += b2*wx + (A3*ub + A6*vb)*wy + (mu*gmu2)*wb;
                cT3 += cv_[JV_K1*32]*wx + cv_[JV_K2*32]*wy;
                cmu[1] += gmu1*wb; cmu[2] += gmu2*wb; cmu[3] += (gmu3 + gk3*g.cp_over_pr)*wb;
                if (SA) {
                    cmut[1] += gmu1*wb; cmut[2] += gmu2*wb; cmut[3] += (gmu3 + gk3*g.cp_over_prt)*wb;
                    cmu[4] += cv_[JV_GN*32]*wb;                        // G4 = (mub + rnb)/sigma (grad nut . n): d/d mu and d/d rn share gn wb
                    cnut4 += cv_[JV_MSX*32]*wx + cv_[JV_MSY*32]*wy;
                }
            }
#else
            if (dc != JC_NONE) {                                       // viscous flux: the dual cell's six cells (flux.cpp:12-48 through mesh.cpp:10-131)
                const double nxf = g_nx, nyf = g_ny;
                const double wx = sc*g_wx, wy = sc*g_wy, wb = sc*(dc <= JC_D1 ? 0.375 : 0.0625);
                const double mu = core[(Cfg::C_V + JV_MU)*32], kk = core[(Cfg::C_V + JV_KK)*32];
                const double ub = core[(Cfg::C_V + JV_UB)*32], vb = core[(Cfg::C_V + JV_VB)*32];
                const double txx_h = core[(Cfg::C_V + JV_TXX)*32], tyy_h = core[(Cfg::C_V + JV_TYY)*32], txy_h = core[(Cfg::C_V + JV_TXY)*32];
                const double Tx = core[(Cfg::C_V + JV_TX)*32], Ty = core[(Cfg::C_V + JV_TY)*32];
                const double c43 = 4.0/3.0*mu, c23 = 2.0/3.0*mu;       // flux.cpp:36-45
                const double txx = mu*txx_h, tyy = mu*tyy_h, txy = mu*txy_h;
                // d G_r / d(ux, uy, vx, vy) . (wx, wy) etc., rows 1..3
                cu[1] += (c43*nxf)*wx + (mu*nyf)*wy;       cv[1] += (mu*nyf)*wx + (-c23*nxf)*wy;
                cu[2] += (-c23*nyf)*wx + (mu*nxf)*wy;      cv[2] += (mu*nxf)*wx + (c43*nyf)*wy;
                const double g_uy3 = mu*(nxf*vb + nyf*ub);
                cu[3] += (nxf*ub*c43 - nyf*vb*c23)*wx + g_uy3*wy + (nxf*txx + nyf*txy)*wb;
                cv[3] += g_uy3*wx + (-nxf*ub*c23 + nyf*vb*c43)*wy + (nxf*txy + nyf*tyy)*wb;
                cT3 += (kk*nxf)*wx + (kk*nyf)*wy;
                const double gmu1 = txx_h*nxf + txy_h*nyf, gmu2 = txy_h*nxf + tyy_h*nyf;
                const double gmu3 = nxf*(ub*txx_h + vb*txy_h) + nyf*(ub*txy_h + vb*tyy_h), gk3 = nxf*Tx + nyf*Ty;
                cmu[1] += gmu1*wb; cmu[2] += gmu2*wb; cmu[3] += (gmu3 + gk3*g.cp_over_pr)*wb;
                if (SA) {
                    cmut[1] += gmu1*wb; cmut[2] += gmu2*wb; cmut[3] += (gmu3 + gk3*g.cp_over_prt)*wb;
                    const double gn = core[(Cfg::C_V + JV_GN)*32], musa_s = core[(Cfg::C_V + JV_MUSA)*32];
                    cmu[4] += gn*wb;                                   // G4 = (mub + rnb)/sigma (grad nut . n): d/d mu and d/d rn share gn wb
                    cnut4 += musa_s*(wx*nxf + wy*nyf);
                }
            }
#endif
        }
        // SA source row (rhs[4] += S V then / V): weights of this cell in the cell-centred gradients
        double s_cu = 0.0, s_cv = 0.0, s_cn = 0.0, s_rho = 0.0, s_mu = 0.0;
        if (SA && inner) {
            const double* Sd = sS + (jl & 1)*7*32 + lane;
            double wx, wy; sa_w(M8, DX, DY, wx, wy);
            const double sgn = Sd[6*32], S3 = Sd[3*32]*sgn;
            s_cu = -S3*wy; s_cv = S3*wx; s_cn = Sd[4*32]*wx + Sd[5*32]*wy;
            if (DX == 0 && DY == 0) { s_rho = Sd[0]; s_mu = Sd[2*32]; s_cn += Sd[1*32]; }
        }
        CellD<NV> cs;
        load_cell(jl + DY, cc0 + DX, inner, cs);
#pragma unroll
        for (int r = 0; r < NV; r++) {
            double out[NV];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) out[c2] = 0.0;
            chain_W<NV>(cs, cw[r], out);
            if (VISC && r >= 1 && inner) {                             // all-zero coefficients outside the 3x3 block
                if (r < 4) chain_Z<NV>(cs, cu[r], cv[r], r == 3 ? cT3 : 0.0, cmu[r], SA ? cmut[r] : 0.0, 0.0, 0.0, out);
                else chain_Z<NV>(cs, s_cu, s_cv, 0.0, cmu[4] + s_mu, 0.0, cnut4 + s_cn, cmu[4], out);
            }
            if (SA && r == 4) { out[0] += ex0 + s_rho; out[4] += ex4; }
            if (cell_ok) {
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) __stcs(Jp + (size_t)(r*NV + c2)*pl, out[c2]);
            }
        }
    };

    // ---- ring rows ra-2 .. ra+2 (W), ra-1 .. ra+1 (Z) ----
#pragma unroll 1
    for (int jl = ra - 2 + warp; jl <= ra + 2; jl += JM_WARPS) convert_w_row(jl);
    __syncthreads();
#pragma unroll 1
    for (int jl = ra - 1 + warp; jl <= ra + 1; jl += JM_WARPS) convert_z_row(jl);
    stage_G(ra - 1);                                 // the lead iteration's eta core (face row ra)
    cp_wait();
    __syncthreads();

    // The loop starts one row early: iteration ra - 1 only produces what row ra inherits from "the row below" -- the eta
    // core of face ra and the SA sensitivities of row ra -- through the SAME call sites as every other row (one inlined
    // copy of each routine: smaller code, and chunk seams cannot change a bit of the result).
#ifdef JM_TIMING
    __shared__ unsigned long long s_ttask[16];
    if (threadIdx.x < 16) s_ttask[threadIdx.x] = 0;
    long long tA = 0, tB = 0, tW = 0, t0, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
#define JM_T(acc) { asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory"); acc += t1 - t0; t0 = t1; }
#else
#define JM_T(acc)
#endif
#pragma unroll 1
    for (int jl = ra - 1; jl < rb; jl++) {
        const bool lead = jl < ra;
        // ---- phase A: cores of chi face (i, jl) [warps 0, 1] and eta face (i, jl+1) [warps 2, 3]
        if (threadIdx.x == 0) *sTask = 0;               // nobody pulls tasks between the barrier behind us and the one ahead
        if (!lead) stage_M8(jl, jl + 1 < rb);
        if (warp < 4) { if (!lead || warp >= 2) face_core(warp >> 1, warp & 1, jl + (warp >> 1)); }
        else if (!lead) {                               // fifth warp: SA sensitivities of THIS row, ring rows of the next rows
            sa_prep(jl);
            if (jl + 1 < rb) { convert_w_row(jl + 3); convert_z_row(jl + 2); }
        }
        cp_wait();                                      // own copies of the row's cell metrics have landed
        JM_T(tA)
        __syncthreads();
#ifdef JM_TIMING
        if (((volatile double*)sW)[lane] == 1.2345e300) tW = 0;    // a barrier-protected read: the clock below is read after the RELEASE
#endif
        JM_T(tW)
        // ---- phase B: the 13 slots, the ring rows entering and the SA preparation of the next row are TASKS pulled from a
        //      shared counter, most expensive first -- the warps finish within one cheap task of each other whatever the
        //      template configuration
        if (jl + 1 < rb) stage_G(jl + 1);               // the next row's cores read these after the barrier below
        if (!lead) {
            Met8 M8;
            M8.cxl = sM8[0*32 + lane]; M8.cyl = sM8[1*32 + lane]; M8.cxr = sM8[2*32 + lane]; M8.cyr = sM8[3*32 + lane];
            M8.exb = sM8[4*32 + lane]; M8.eyb = sM8[5*32 + lane]; M8.ext = sM8[6*32 + lane]; M8.eyt = sM8[7*32 + lane];
            M8.Vi = 1.0/sM8[8*32 + lane];
            const bool more = jl + 1 < rb;
            while (true) {
                int n = 0;
                if (lane == 0) n = atomicAdd(sTask, 1);
                n = __shfl_sync(0xffffffffu, n, 0);
                if (n >= JM_NTASK) break;
                const int task = (int)((JM_TASKS >> (4*n)) & 15ull);
#ifdef JM_TIMING
                long long tt0, tt1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt0) :: "memory");
#endif
                if (task < 13) { if (task < NS) assemble_slot(task, jl, M8); }
                else if (more) {
                    if (task == 13) sa_prep(jl + 1);
                    else if (task == 14) convert_w_row(jl + 3);
                    else convert_z_row(jl + 2);
                }
#ifdef JM_TIMING
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt1) :: "memory");
                if (lane == 0) atomicAdd((unsigned long long*)&s_ttask[task], (unsigned long long)(tt1 - tt0));
#endif
            }
        } else if (warp == 3 && JM_WARPS == 4) sa_prep(jl + 1);
        cp_wait();
        JM_T(tB)
        __syncthreads();
#ifdef JM_TIMING
        if (((volatile double*)sW)[lane] == 1.2345e300) tW = 0;
#endif
        JM_T(tW)
    }
#ifdef JM_TIMING
    if (lane == 0 && blockIdx.x % 997 == 5) printf("cta %d warp %d: phase A %lld  phase B %lld  barrier wait %lld cycles over %d rows\n", blockIdx.x, warp, tA, tB, tW, rb - ra);
    if (threadIdx.x == 0 && blockIdx.x % 997 == 5) {
        printf("cta %d task cycles per row:", blockIdx.x);
        for (int k = 0; k < 16; k++) printf(" t%d=%lld", k, (long long)(s_ttask[k]/(unsigned long long)(rb - ra)));
        printf("\n");
    }
#endif
#undef JM_T
}

// Fold ghost slots into the interior cells they are functions of (boundary band only): the chain d ghost / d interior of
// the boundary condition that wrote the ghost cell last (src/model/bc.cpp), corners first, then arms, then edges.
template <int NV>
__global__ void __launch_bounds__(128) jac_fold_kernel(View v, Gas g, Metrics m, GhostTable gt, const double* __restrict__ q, double* __restrict__ J,
                                                       int nslots, int* __restrict__ err, int mode) {
    // mode 0: one thread per cell of the slab, the band test decides (small grids); mode 1: the two bottom and two top rows of the
    // slab, grid (nic/128, 4); mode 2: the two left and two right columns, grid (njl/128, 4), without the cells mode 1 covers --
    // the full-grid launch spent 0.5 ms at 4096^2 on 16.7 M threads that return at once
    int i, jl;
    if (mode == 2) { jl = blockIdx.x*blockDim.x + threadIdx.x; i = (int)blockIdx.y < 2 ? (int)blockIdx.y : v.nic - 4 + (int)blockIdx.y; if (jl >= v.njl) return; }
    else { i = blockIdx.x*blockDim.x + threadIdx.x; jl = mode == 1 ? ((int)blockIdx.y < 2 ? (int)blockIdx.y : v.njl - 4 + (int)blockIdx.y) : (int)blockIdx.y; }
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    if (!((i < 2) || (i > v.nic - 3) || (gj < 2) || (gj > v.njc - 3))) return;
    if (mode == 2 && (jl < 2 || jl > v.njl - 3)) return;            // the slab's own row bands belong to mode 1
    const size_t o = v.at(jl + JOFF, i + IOFF);
    unsigned touched = (1u << nslots) - 1u;
    const int ip0 = i + 1, jp0 = gj + 1;                           // padded coordinates of the row cell
    const int order_list[12] = {5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 3, 4};
    // three sweeps: a fold may deposit into a ghost slot that was already visited (chains of two BC maps)
#pragma unroll 1
    for (int n3 = 0; n3 < 36; n3++) {
        const int s = order_list[n3 % 12];
        if (s >= nslots || !(touched & (1u << s))) continue;
        int ip = ip0 + c_slot_dx[s], jp = jp0 + c_slot_dy[s];
        if (!gt.is_ghost(ip, jp)) continue;
        if (ip < 0 || ip > gt.nic + 1 || jp < 0 || jp > gt.njc + 1) continue;   // beyond the ghost layer: never read
        gt.resolve(ip, jp);
        if (!gt.is_ghost(ip, jp)) continue;                        // copy-type ghost: keeps its slot, column remapped at export
        const GhostDesc& gd = gt.at(ip, jp);
        double B[NV*NV];
        double* p = J + ((size_t)s*NV*NV)*v.plane + o;
#pragma unroll
        for (int e = 0; e < NV*NV; e++) { B[e] = p[e*v.plane]; p[e*v.plane] = 0.0; }
        if (gd.type == SGPU_BC_FREESTREAM || gd.type < 0) continue;   // constants: no dependency
        const int aip = gd.a_ip, ajp = gd.a_jp, bip = gd.b_ip, bjp = gd.b_jp;
        const bool has_b = gd.type != SGPU_BC_OUTFLOW;
        double qa[NV], qb[NV];
        auto ldq = [&](int pip, int pjp, double* dst) {
            const int rr = pjp - 1 - v.j0 + JOFF, cc2 = pip - 1 + IOFF;
#pragma unroll
            for (int k = 0; k < NV; k++) dst[k] = q[k*v.plane + v.at(rr, cc2)];
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
#pragma unroll 1
        for (int t = 0; t < (has_b ? 2 : 1); t++) {
            int tip = t ? bip : aip, tjp = t ? bjp : ajp;
            gt.resolve(tip, tjp);
            int ts = -1;
            for (int s2 = 0; s2 < nslots; s2++) {
                int sip = ip0 + c_slot_dx[s2], sjp = jp0 + c_slot_dy[s2];
                gt.resolve(sip, sjp);
                if (sip == tip && sjp == tjp) { ts = s2; break; }
            }
            if (ts < 0) { atomicAdd(err, 1); continue; }
            const double* M = t ? Mb : Ma;
            double* pt = J + ((size_t)ts*NV*NV)*v.plane + o;
#pragma unroll
            for (int rr = 0; rr < NV; rr++)
#pragma unroll
                for (int cc2 = 0; cc2 < NV; cc2++) {
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < NV; k++) sacc += B[rr*NV + k]*M[k*NV + cc2];
                    pt[(rr*NV + cc2)*v.plane] += sacc;
                }
        }
    }
}

} // namespace sg
