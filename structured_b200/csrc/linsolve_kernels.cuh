// Device kernels of the matrix-free-storage Krylov solve that replaces src/linearsolver/* on the GPU
// (reference API: src/linearsolver/ls_eigen.h:26-31, call site src/solver/solver.cpp:172-175).
//
// The matrix is never assembled into COO/CSR: the operator works on the block-stencil Jacobian planes
// J[slot][r][c][cell] that jac_gather_kernel wrote (13 or 9 slots of nv x nv blocks per row cell), and the
// LHS transform of src/solver/solver.cpp:162-171 (A = -J + delta/dt) is applied on the fly.
//
// Vectors are state-plane arrays (nv planes of rows x pitch doubles) whose ghost rows/columns and padding are
// kept at exactly 0, so the BLAS-1 style kernels below run flat over plane*nv elements, fully coalesced.
//
// Roofline: op_apply_kernel streams every Jacobian block once -> (slots*nv*nv + 2*nv + 1)*8 B per cell
// (2648 B at 13 slots, nv = 5); everything else is O(nv) per cell.  All of it is HBM-bound.
#pragma once
#include "jacobian_kernel.cuh"

namespace sg {

enum { OP_J = 0, OP_LHS = 1 };     // A = J   |   A = -J + delta/dt  (what linearsolver->set_lhs receives)

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ------------------------------------------------------------------------------------------------
// y = A x, one thread per row cell.  Cells whose whole stencil is interior take the natural-neighbour
// fast path; the boundary band resolves each slot's column through the ghost table (periodic / wake
// remaps, folded functional ghosts) exactly like the COO export does.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128) op_apply_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2,
                                                       const double* __restrict__ J, const double* __restrict__ dt, int op,
                                                       const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    const size_t pl = v.plane;
    double yr[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) yr[r] = 0.0;
    const bool inner = i >= 2 && i < v.nic - 2 && gj >= 2 && gj < v.njc - 2;
    if (inner) {
        // Software pipelined over the slots: the 25 block entries and 5 operand values of the NEXT slot are loaded while the
        // current slot's products run (ncu on the plain loop: latency bound -- long_scoreboard 62 stall cycles per issue at 37
        // warps per SM -- because the compiler sinks every load to its first use and a warp issues in order, so the first
        // dependent DFMA blocks the loads behind it: ~6 loads in flight per warp.  Here a warp keeps 30.)
        auto slot_on = [&](int s) { return !((!viscous && s >= 5 && s <= 8) || (!order2 && s >= 9)); };
        auto fetch = [&](int s, double (&a)[NV*NV], double (&xs)[NV]) {
            const size_t oc = o + (long long)c_slot_dy[s]*v.pitch + c_slot_dx[s];
            const double* Js = J + (size_t)s*NV*NV*pl + o;
            // the 45 of 325 entries that are STRUCTURALLY zero (entry_present: mass row x corner cells, q4 column of the mass row
            // and of the radius-2 arms) are not read: these planes hold exact zeros for inner cells (no ghost folding there)
            const bool corner = s >= 5 && s <= 8, arm = s >= 9;
#pragma unroll
            for (int e = 0; e < NV*NV; e++) {
                const int r = e/NV, c2 = e - r*NV;
                const bool zero = (r == 0 && corner) || (NV > 4 && c2 == 4 && (r == 0 || arm));
                a[e] = zero ? 0.0 : ld_stream(Js + (size_t)e*pl);
            }
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) xs[c2] = x[c2*pl + oc];
        };
        double a[NV*NV], xs[NV], an[NV*NV], xn[NV];
        int s = 0;
        fetch(0, a, xs);                                           // slot 0 (the cell itself) is always present
        while (s < nslots) {
            int sn = s + 1;
            while (sn < nslots && !slot_on(sn)) sn++;
            if (sn < nslots) fetch(sn, an, xn);
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) yr[r] += a[r*NV + c2]*xs[c2];
#pragma unroll
            for (int e = 0; e < NV*NV; e++) a[e] = an[e];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) xs[c2] = xn[c2];
            s = sn;
        }
    } else {
        SlotCols sc; resolve_slots(gt, nslots, i, gj, viscous, order2, sc);
#pragma unroll 1
        for (int s = 0; s < nslots; s++) {
            if (sc.col[s] < 0) continue;
            const int ci = sc.col[s]/v.njc, cj = sc.col[s] - ci*v.njc;
            const int rr = cj - v.j0 + JOFF;
            if (rr < 0 || rr >= v.rows) continue;                  // column outside this slab's planes
            const size_t oc = v.at(rr, ci + IOFF);
            double xs[NV];
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) xs[c2] = x[c2*pl + oc];
            const double* Js = J + (size_t)s*NV*NV*pl + o;
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c2 = 0; c2 < NV; c2++) yr[r] += ld_stream(Js + (size_t)(r*NV + c2)*pl)*xs[c2];
        }
    }
    if (op == OP_LHS) {
        const double idt = 1.0/dt[o];                              // src/solver/solver.cpp:167-170
#pragma unroll
        for (int r = 0; r < NV; r++) yr[r] = x[r*pl + o]*idt - yr[r];
    }
#pragma unroll
    for (int r = 0; r < NV; r++) y[r*pl + o] = yr[r];
}

// y = J^T x in GATHER form for the contributions of "inner" row cells (whole stencil interior, natural columns):
//     (J^T x)[c] = sum_s J_s[c - off_s]^T x[c - off_s]
// -- the blocks of the neighbouring ROW cells are read at a shifted offset, still coalesced, and every y is written
// once (deterministic, no atomics).  Row cells in the boundary band (remapped / folded columns) are added afterwards by
// jac_apply_kernel's atomic scatter restricted to the band (op_apply_t_band_kernel below).
// On a j-slab the launch also covers the two ghost rows of every interior edge (jl0 = -2 / rows up to njl + 2): what this
// slab's row cells contribute to the neighbour's cells is left in the GHOST rows of y, which the caller sends over and
// adds there (sgpu_vec_halo_pack_ghost -> transport -> sgpu_vec_halo_add: the transpose of the operand halo exchange).
template <int NV>
__global__ void __launch_bounds__(128) op_apply_t_kernel(View v, int nslots, bool viscous, bool order2,
                                                         const double* __restrict__ J, const double* __restrict__ x, double* __restrict__ y, int jl0) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = jl0 + (int)blockIdx.y;
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    const size_t pl = v.plane;
    double yr[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) yr[r] = 0.0;
    // software pipelined like op_apply_kernel: the next contributing slot's loads are in flight under this slot's products, and
    // the structurally-zero entries (exact zeros in inner row cells) are not read
    auto slot_row = [&](int s, size_t& orow) -> bool {              // the row cell whose slot s points at this cell, if it contributes here
        if ((!viscous && s >= 5 && s <= 8) || (!order2 && s >= 9)) return false;
        const int ri = i - c_slot_dx[s], rj = gj - c_slot_dy[s];
        if (!(ri >= 2 && ri < v.nic - 2 && rj >= 2 && rj < v.njc - 2)) return false;    // band rows: scattered separately
        const int rl = rj - v.j0;
        if (rl < 0 || rl >= v.njl) return false;                   // row cell owned by another slab
        orow = v.at(rl + JOFF, ri + IOFF);
        return true;
    };
    auto fetch = [&](int s, size_t orow, double (&a)[NV*NV], double (&xs)[NV]) {
        const double* Js = J + (size_t)s*NV*NV*pl + orow;
        const bool corner = s >= 5 && s <= 8, arm = s >= 9;
#pragma unroll
        for (int e = 0; e < NV*NV; e++) {
            const int r = e/NV, c2 = e - r*NV;
            const bool zero = (r == 0 && corner) || (NV > 4 && c2 == 4 && (r == 0 || arm));
            a[e] = zero ? 0.0 : ld_stream(Js + (size_t)e*pl);
        }
#pragma unroll
        for (int r = 0; r < NV; r++) xs[r] = x[r*pl + orow];
    };
    double a[NV*NV], xs[NV], an[NV*NV], xn[NV];
    size_t orow = 0, onext = 0;
    int s = 0;
    while (s < nslots && !slot_row(s, orow)) s++;
    if (s < nslots) fetch(s, orow, a, xs);
    while (s < nslots) {
        int sn = s + 1;
        while (sn < nslots && !slot_row(sn, onext)) sn++;
        if (sn < nslots) fetch(sn, onext, an, xn);
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c2 = 0; c2 < NV; c2++) yr[c2] += a[r*NV + c2]*xs[r];
#pragma unroll
        for (int e = 0; e < NV*NV; e++) a[e] = an[e];
#pragma unroll
        for (int r = 0; r < NV; r++) xs[r] = xn[r];
        s = sn;
    }
#pragma unroll
    for (int r = 0; r < NV; r++) y[r*pl + o] = yr[r];
}

// the band rows' share of J^T x: atomic scatter, but only from row cells within 2 of a physical boundary
template <int NV>
__global__ void op_apply_t_band_kernel(View v, GhostTable gt, int nslots, bool viscous, bool order2, const double* __restrict__ J,
                                       const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const int gj = v.j0 + jl;
    if (i >= 2 && i < v.nic - 2 && gj >= 2 && gj < v.njc - 2) return;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    SlotCols sc; resolve_slots(gt, nslots, i, gj, viscous, order2, sc);
    for (int s = 0; s < nslots; s++) {
        if (sc.col[s] < 0) continue;
        const int ci = sc.col[s]/v.njc, cj = sc.col[s] - ci*v.njc;
        const int rr = cj - v.j0 + JOFF;
        if (rr < 0 || rr >= v.rows) continue;
        const size_t oc = v.at(rr, ci + IOFF);
        for (int r = 0; r < NV; r++) {
            const double xr = x[r*v.plane + o];
            for (int c2 = 0; c2 < NV; c2++) atomicAdd(&y[c2*v.plane + oc], J[((size_t)s*NV*NV + r*NV + c2)*v.plane + o]*xr);
        }
    }
}

// ghost rows of a vector: dst rows += packed buffer [nv][2][nic] (reverse halo exchange of a transposed product)
__global__ void halo_add_kernel(View v, double* __restrict__ vec, const double* __restrict__ buf, int r_first) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= v.nic) return;
    const int k = blockIdx.y >> 1, rr = blockIdx.y & 1;
    vec[k*v.plane + v.at(r_first + rr, i + IOFF)] += buf[((size_t)k*2 + rr)*v.nic + i];
}
__global__ void halo_zero_kernel(View v, double* __restrict__ vec, int r_first) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= v.nic) return;
    const int k = blockIdx.y >> 1, rr = blockIdx.y & 1;
    vec[k*v.plane + v.at(r_first + rr, i + IOFF)] = 0.0;
}

// after the scattered J^T x accumulation: y = -y + x/dt on the owned cells
template <int NV>
// (rows from jl0: on a slab the ghost rows of y carry the neighbour's share of J^T x, which only changes sign)
__global__ void lhs_fixup_kernel(View v, const double* __restrict__ dt, const double* __restrict__ x, double* __restrict__ y, int jl0) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = jl0 + (int)blockIdx.y;
    if (i >= v.nic) return;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    if (jl < 0 || jl >= v.njl) {
        for (int r = 0; r < NV; r++) y[r*v.plane + o] = -y[r*v.plane + o];
        return;
    }
    const double idt = 1.0/dt[o];
    for (int r = 0; r < NV; r++) y[r*v.plane + o] = x[r*v.plane + o]*idt - y[r*v.plane + o];
}

// dst = src on the owned cells, 0 on ghosts and padding
__global__ void masked_copy_kernel(View v, const double* __restrict__ src, double* __restrict__ dst) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= v.pitch) return;
    const bool own = c >= IOFF && c < IOFF + v.nic && r >= JOFF && r < JOFF + v.njl;
    const size_t o = v.at(r, c);
    for (int k = 0; k < v.nv; k++) dst[k*v.plane + o] = own ? src[k*v.plane + o] : 0.0;
}

// q += omega * x on the owned cells (the update of src/linearsolver/ls_eigen.cpp:66-70 for a slab vector whose ghost
// rows hold halo data)
__global__ void masked_axpy_kernel(View v, double* __restrict__ q, const double* __restrict__ x, double omega) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    for (int k = 0; k < v.nv; k++) q[k*v.plane + o] += omega*x[k*v.plane + o];
}

// ------------------------------------------------------------------------------------------------
// nv x nv inverse in registers: Gauss-Jordan with partial pivoting, every index a compile-time constant.
// Returns false when a pivot is exactly zero.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ bool invert_block(double (&a)[NV][NV], double (&b)[NV][NV]) {
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) b[r][c] = r == c ? 1.0 : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        int p = k; double big = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < NV; r++) { const double t = fabs(a[r][k]); if (t > big) { big = t; p = r; } }
#pragma unroll
        for (int r = k + 1; r < NV; r++)
            if (r == p) {
#pragma unroll
                for (int c = 0; c < NV; c++) { double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; t = b[k][c]; b[k][c] = b[r][c]; b[r][c] = t; }
            }
        if (a[k][k] == 0.0) { ok = false; a[k][k] = 1.0; }
        const double ip = 1.0/a[k][k];
#pragma unroll
        for (int c = 0; c < NV; c++) { a[k][c] *= ip; b[k][c] *= ip; }
#pragma unroll
        for (int r = 0; r < NV; r++) {
            if (r == k) continue;
            const double f = a[r][k];
#pragma unroll
            for (int c = 0; c < NV; c++) { a[r][c] -= f*a[k][c]; b[r][c] -= f*b[k][c]; }
        }
    }
    return ok;
}

// The same Gauss-Jordan elimination for the line factorisation's dependent chain: the pivot row is swapped in with SELECTS
// (32 lanes = 32 lines pivot differently: the branchy swaps above diverge into up to ten serial blocks per inversion) and the
// pivot reciprocal is a MUFU seed + one third-order step (rcp_fast, 1 ulp) instead of an IEEE division.
template <int NV>
__device__ __forceinline__ bool invert_block_fast(double (&a)[NV][NV], double (&b)[NV][NV]) {
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) b[r][c] = r == c ? 1.0 : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        int p = k; double big = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < NV; r++) { const double t = fabs(a[r][k]); const bool g = t > big; big = g ? t : big; p = g ? r : p; }
#pragma unroll
        for (int r = k + 1; r < NV; r++) {
            const bool sw = r == p;
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const double x = a[k][c], y = a[r][c]; a[k][c] = sw ? y : x; a[r][c] = sw ? x : y;
                const double u = b[k][c], w = b[r][c]; b[k][c] = sw ? w : u; b[r][c] = sw ? u : w;
            }
        }
        if (a[k][k] == 0.0) { ok = false; a[k][k] = 1.0; }
        const double ip = rcp_fast(a[k][k]);
#pragma unroll
        for (int c = 0; c < NV; c++) { a[k][c] *= ip; b[k][c] *= ip; }
#pragma unroll
        for (int r = 0; r < NV; r++) {
            if (r == k) continue;
            const double f = a[r][k];
#pragma unroll
            for (int c = 0; c < NV; c++) { a[r][c] = fma(-f, a[k][c], a[r][c]); b[r][c] = fma(-f, b[k][c], b[r][c]); }
        }
    }
    return ok;
}

template <int NV>
__device__ __forceinline__ void load_block(const double* __restrict__ J, size_t pl, size_t o, int s, int op, double idt, bool diag, double (&a)[NV][NV]) {
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const double vj = J[((size_t)s*NV*NV + r*NV + c)*pl + o];
            a[r][c] = op == OP_LHS ? ((diag && r == c) ? idt - vj : -vj) : vj;
        }
}

// ------------------------------------------------------------------------------------------------
// Block-Jacobi preconditioner: Dinv = (diagonal block of A)^-1 per cell, stored as nv*nv planes
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void bj_factor_kernel(View v, const double* __restrict__ J, const double* __restrict__ dt, int op, double* __restrict__ Dinv, int* __restrict__ err) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    double a[NV][NV], b[NV][NV];
    load_block<NV>(J, v.plane, o, 0, op, op == OP_LHS ? 1.0/dt[o] : 0.0, true, a);
    if (!invert_block<NV>(a, b)) atomicExch(err, 1);
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) Dinv[(size_t)(r*NV + c)*v.plane + o] = b[r][c];
}

// z = Dinv r (transpose = 0) or Dinv^T r (transpose = 1) on the owned cells
template <int NV>
__global__ void bj_apply_kernel(View v, const double* __restrict__ Dinv, const double* __restrict__ rv, double* __restrict__ z, int transpose) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    if (i >= v.nic) return;
    const size_t o = v.at(jl + JOFF, i + IOFF);
    double xs[NV], zr[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) { xs[c] = rv[c*v.plane + o]; zr[c] = 0.0; }
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const double a = ld_stream(Dinv + (size_t)(r*NV + c)*v.plane + o);
            if (!transpose) zr[r] += a*xs[c]; else zr[c] += a*xs[r];
        }
#pragma unroll
    for (int r = 0; r < NV; r++) z[r*v.plane + o] = zr[r];
}

// ------------------------------------------------------------------------------------------------
// j-line preconditioner: block-tridiagonal solve along each grid line i = const (the wall-normal, strongly
// coupled direction of a boundary-layer grid).  One thread per line, block Thomas algorithm.
// Line blocks: sub-diagonal A' = slot 3 (0,-1) + slot 11 (0,-2), diagonal slot 0, super-diagonal
// C' = slot 4 (0,+1) + slot 12 (0,+2).  LUMPING the radius-2 arms of a second-order Jacobian onto their
// radius-1 neighbours matters: the bare tridiagonal part of the kappa = 1/3 MUSCL stencil is not a usable
// approximation of the line operator (the preconditioned spectrum reaches into the left half plane),
// the lumped one clusters it in [0.27, 2.2] like the exact pentadiagonal line solve would.
// Couplings that leave the line's natural neighbours (bottom/top ghosts -- already folded into the interior
// slots -- or periodic/wake remaps) are not part of the preconditioner.
// Stored factors (3 x nv*nv planes): Dinv = D'^-1, DA = D'^-1 A', DC = D'^-1 C' with
//     D'_j = D_j - A'_j DC_{j-1}
// so that both sweeps have ONE dependent nv x nv product per row on the critical path.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void matmul_block(const double (&a)[NV][NV], const double (&b)[NV][NV], double (&c)[NV][NV]) {
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int cc = 0; cc < NV; cc++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NV; k++) s += a[r][k]*b[k][cc];
            c[r][cc] = s;
        }
}

// The recurrence D'_j = D_j - A'_j DC_{j-1} is sequential in j and only nic/32 warps exist, so the kernel is pure latency:
// with the row's five Jacobian blocks loaded when the loop reaches them, every row paid 3-4 dependent DRAM round trips
// (35 ms at 4096^2).  As in line_apply_kernel below, each lane streams the words it will read itself -- the blocks of slots
// 0, 3, 11, 4, 12 and 1/dt of the rows ahead -- through a FACT_STAGES-deep shared-memory ring with cp.async: no barrier,
// coalesced 256-byte row segments per plane, and the loop body is arithmetic only.
constexpr int FACT_STAGES = 6;
template <int NV> constexpr int fact_ring_planes() { return 5*NV*NV + 1; }
template <int NV> constexpr size_t fact_ring_bytes() { return (size_t)FACT_STAGES*fact_ring_planes<NV>()*32*sizeof(double); }

__device__ __forceinline__ void fact_cp_async8(double* smem, const double* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(g) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(32) line_factor_kernel(View v, const double* __restrict__ J, const double* __restrict__ dt, int op, int nslots,
                                                         double* __restrict__ F, int* __restrict__ err) {
    extern __shared__ double fring[];
    constexpr int B = NV*NV, NP = 5*NV*NV + 1, S = FACT_STAGES;
    const int lane = threadIdx.x;
    const int i = blockIdx.x*32 + lane;
    const bool live = i < v.nic;
    const int ic = live ? i : v.nic - 1;                           // idle lanes shadow the last line and never store
    const size_t pl = v.plane;
    double* __restrict__ Dinv = F;
    double* __restrict__ DAo = F + (size_t)NV*NV*pl;
    double* __restrict__ DC = F + (size_t)2*NV*NV*pl;
    const bool arms = nslots > 9;
    auto slot = [&](int st, int p) -> double* { return fring + ((size_t)st*NP + p)*32 + lane; };
    auto issue = [&](int jl) {                                     // ring group g = 0..4 <- Jacobian slots 0, 3, 11, 4, 12
        if (jl < v.njl) {
            const size_t o = v.at(jl + JOFF, ic + IOFF);
            const int st = jl % S;
#pragma unroll
            for (int g = 0; g < 5; g++) {
                const int sl = g == 0 ? 0 : (g == 1 ? 3 : (g == 2 ? 11 : (g == 3 ? 4 : 12)));
                if ((g == 2 || g == 4) && !arms) continue;
#pragma unroll
                for (int e = 0; e < B; e++) fact_cp_async8(slot(st, g*B + e), J + ((size_t)sl*B + e)*pl + o);
            }
            if (op == OP_LHS) fact_cp_async8(slot(st, 5*B), dt + o);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the same transform load_block applies: A = -J + I/dt for OP_LHS, J otherwise
    auto ring_block = [&](int st, int g, double idt, bool diag, double (&a)[NV][NV]) {
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const double vj = *slot(st, g*B + r*NV + c);
                a[r][c] = op == OP_LHS ? ((diag && r == c) ? idt - vj : -vj) : vj;
            }
    };
    // one row of a ring block (lumped with the radius-2 arm's block when `arm`), transformed like load_block
    auto ring_row = [&](int st, int g, int ga, bool arm, int r, double (&a)[NV]) {
#pragma unroll
        for (int c = 0; c < NV; c++) {
            double vj = *slot(st, g*B + r*NV + c);
            if (arm) vj += *slot(st, ga*B + r*NV + c);
            a[c] = op == OP_LHS ? -vj : vj;
        }
    };
    for (int jl = 0; jl < S - 1; jl++) issue(jl);
    double DCp[NV][NV];                                            // DC_{j-1}
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) DCp[r][c] = 0.0;
    for (int jl = 0; jl < v.njl; jl++) {
        issue(jl + S - 1);
        asm volatile("cp.async.wait_group %0;" :: "n"(S - 1) : "memory");
        const int st = jl % S;
        const int gj = v.j0 + jl;
        const size_t o = v.at(jl + JOFF, ic + IOFF);
        const bool lo = jl > 0, hi = jl + 1 < v.njl;
        const bool arm_lo = arms && gj - 2 >= 0, arm_hi = arms && gj + 2 <= v.njc - 1;
        // D' = D - A' DC_{j-1}, formed row by row: the blocks are re-read from the ring where they are needed instead of being
        // held (five nv x nv blocks in registers put part of them into local memory: 232 registers + a 400-byte stack frame)
        double D[NV][NV], I[NV][NV];
        ring_block(st, 0, op == OP_LHS ? rcp_fast(*slot(st, 5*B)) : 0.0, true, D);
        if (lo) {
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double a[NV];
                ring_row(st, 1, 2, arm_lo, r, a);
#pragma unroll
                for (int k = 0; k < NV; k++)
#pragma unroll
                    for (int c = 0; c < NV; c++) D[r][c] = fma(-a[k], DCp[k][c], D[r][c]);
            }
        }
        if (!invert_block_fast<NV>(D, I)) atomicExch(err, 1);
        // DA = D'^-1 A' and DC = D'^-1 C': accumulated over the rows k of A' / C' as they come out of the ring
        double DA[NV][NV];
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c = 0; c < NV; c++) { DA[r][c] = 0.0; DCp[r][c] = 0.0; }
        if (lo) {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                double a[NV];
                ring_row(st, 1, 2, arm_lo, k, a);
#pragma unroll
                for (int r = 0; r < NV; r++)
#pragma unroll
                    for (int c = 0; c < NV; c++) DA[r][c] = fma(I[r][k], a[c], DA[r][c]);
            }
        }
        if (live) {
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) { Dinv[(size_t)(r*NV + c)*pl + o] = I[r][c]; DAo[(size_t)(r*NV + c)*pl + o] = DA[r][c]; }
        }
        if (hi) {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                double a[NV];
                ring_row(st, 3, 4, arm_hi, k, a);
#pragma unroll
                for (int r = 0; r < NV; r++)
#pragma unroll
                    for (int c = 0; c < NV; c++) DCp[r][c] = fma(I[r][k], a[c], DCp[r][c]);
            }
        }
        if (live) {
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) DC[(size_t)(r*NV + c)*pl + o] = DCp[r][c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TWISTED factorisation (round 2b; the solve is line_apply_twisted_kernel): two warps per 32 lines eliminate from both ends of
// the line towards the middle row m = njl/2.  Below m: D'_j = D_j - A'_j DC_{j-1} as above.  Above m, descending:
// D''_j = D_j - C'_j DA_{j+1}.  Middle row: D*_m = D_m - A'_m DC_{m-1} - C'_m DA_{m+1}.  Every row stores Dinv = (its eliminated
// diagonal)^-1, DA = Dinv A', DC = Dinv C' as before -- which of the two the neighbouring row's elimination uses depends on the
// side.  Half the dependent rows per warp: the setup halves.
// ------------------------------------------------------------------------------------------------
constexpr int TWF_STAGES = 3;
template <int NV> constexpr int fact_twisted_planes() { return 5*(NV*NV + (NV*NV & 1)) + 2; }      // five block groups padded to even plane counts + the dt pair
template <int NV> constexpr size_t fact_twisted_bytes() { return (size_t)2*TWF_STAGES*fact_twisted_planes<NV>()*32*sizeof(double) + 2*NV*NV*32*sizeof(double); }

template <int NV>
__global__ void __launch_bounds__(64) line_factor_twisted_kernel(View v, const double* __restrict__ J, const double* __restrict__ dt, int op, int nslots,
                                                                 double* __restrict__ F, int* __restrict__ err) {
    extern __shared__ double tw_fring[];
    constexpr int B = NV*NV, BP = B + (B & 1), NP = 5*BP + 2, S = TWF_STAGES;
    const int lane = threadIdx.x & 31, up = threadIdx.x >> 5;
    double* fring = tw_fring + (size_t)up*S*NP*32;
    double* ex = tw_fring + (size_t)2*S*NP*32;                     // [2][B][32]: the product each half hands to the middle row
    const int i = blockIdx.x*32 + lane;
    const bool live = i < v.nic;
    const int ic = live ? i : v.nic - 1;                           // idle lanes shadow the last line and never store
    const size_t pl = v.plane;
    double* __restrict__ Dinv = F;
    double* __restrict__ DAo = F + (size_t)B*pl;
    double* __restrict__ DCo = F + (size_t)2*B*pl;
    const bool arms = nslots > 9;
    const int n = v.njl, m = n/2;
    const int row0 = up ? n - 1 : 0, step = up ? -1 : 1, count = up ? n - 1 - m : m;
    auto slot = [&](int st, int p) -> double* { return fring + ((size_t)st*NP + p)*32 + lane; };
    // sixteen-byte copies, two planes of one block per instruction (lanes 0-15 the first, 16-31 the second), as in the sweeps
    const int half = lane >> 4, word = (lane & 15)*2;
    const int c0 = blockIdx.x*32 + IOFF;
    const bool in_row = c0 + word < v.pitch;
    const size_t lane_src = (size_t)half*pl + word;
    const unsigned lane_dst = (unsigned)__cvta_generic_to_shared(fring) + (unsigned)((half*32 + word)*sizeof(double));
    for (int k = lane; k < S*NP*32; k += 32) fring[k] = 1.0;      // words past a short last segment are never copied
    __syncwarp();
    auto issue = [&](int k) {                                      // k-th row of this half; ring group g = 0..4 <- Jacobian slots 0, 3, 11, 4, 12
        if (k < count && in_row) {
            const size_t o = v.at(row0 + k*step + JOFF, c0) + lane_src;
            const unsigned d = lane_dst + (unsigned)((k % S)*NP*32*sizeof(double));
#pragma unroll
            for (int g = 0; g < 5; g++) {
                const int sl = g == 0 ? 0 : (g == 1 ? 3 : (g == 2 ? 11 : (g == 3 ? 4 : 12)));
                if ((g == 2 || g == 4) && !arms) continue;
                const double* src = J + (size_t)sl*B*pl + o;
#pragma unroll
                for (int h = 0; h < (B + 1)/2; h++)
                    if (2*h + 1 < B || half == 0)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)((g*BP + 2*h)*32*sizeof(double))), "l"(src + (size_t)(2*h)*pl) : "memory");
            }
            if (op == OP_LHS && half == 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)(5*BP*32*sizeof(double))), "l"(dt + v.at(row0 + k*step + JOFF, c0) + word) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto ring_row = [&](int st, int g, int ga, bool arm, int r, double (&a)[NV]) {
#pragma unroll
        for (int c = 0; c < NV; c++) {
            double vj = *slot(st, g*BP + r*NV + c);
            if (arm) vj += *slot(st, ga*BP + r*NV + c);
            a[c] = op == OP_LHS ? -vj : vj;
        }
    };
    for (int k = 0; k < S - 1; k++) issue(k);
    double P[NV][NV];                                              // Dinv x (the block towards the middle) of the previous row of this half
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) P[r][c] = 0.0;
    // X = the block that couples a row to the previous row of its half (A' below the middle, C' above), Y = the other one
    const int gx = up ? 3 : 1, gxa = up ? 4 : 2, gy = up ? 1 : 3, gya = up ? 2 : 4;
    double* __restrict__ DX = up ? DCo : DAo;                      // Dinv X is stored where Dinv A' / Dinv C' belong
    double* __restrict__ DY = up ? DAo : DCo;
    for (int k = 0; k < count; k++) {
        __syncwarp();                                              // every lane is done with the stage the next copy overwrites
        issue(k + S - 1);
        asm volatile("cp.async.wait_group %0;" :: "n"(S - 1) : "memory");
        __syncwarp();                                              // the other lanes' copies of this row have landed too
        const int st = k % S;
        const int jl = row0 + k*step, gj = v.j0 + jl;
        const size_t o = v.at(jl + JOFF, ic + IOFF);
        const bool has_a = jl > 0, has_c = jl + 1 < n;
        const bool arm_a = arms && gj - 2 >= 0, arm_c = arms && gj + 2 <= v.njc - 1;
        const bool has_x = up ? has_c : has_a, has_y = up ? has_a : has_c, arm_x = up ? arm_c : arm_a, arm_y = up ? arm_a : arm_c;
        double D[NV][NV], I[NV][NV];
        {
            const double idt = op == OP_LHS ? rcp_fast(*slot(st, 5*BP)) : 0.0;
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    const double vj = *slot(st, r*NV + c);
                    D[r][c] = op == OP_LHS ? (r == c ? idt - vj : -vj) : vj;
                }
        }
        if (has_x && k > 0) {
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double a[NV];
                ring_row(st, gx, gxa, arm_x, r, a);
#pragma unroll
                for (int kk = 0; kk < NV; kk++)
#pragma unroll
                    for (int c = 0; c < NV; c++) D[r][c] = fma(-a[kk], P[kk][c], D[r][c]);
            }
        }
        if (!invert_block_fast<NV>(D, I) && live) atomicExch(err, 1);
        double DXv[NV][NV];
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c = 0; c < NV; c++) { DXv[r][c] = 0.0; P[r][c] = 0.0; }
        if (has_x) {
#pragma unroll
            for (int kk = 0; kk < NV; kk++) {
                double a[NV];
                ring_row(st, gx, gxa, arm_x, kk, a);
#pragma unroll
                for (int r = 0; r < NV; r++)
#pragma unroll
                    for (int c = 0; c < NV; c++) DXv[r][c] = fma(I[r][kk], a[c], DXv[r][c]);
            }
        }
        if (has_y) {
#pragma unroll
            for (int kk = 0; kk < NV; kk++) {
                double a[NV];
                ring_row(st, gy, gya, arm_y, kk, a);
#pragma unroll
                for (int r = 0; r < NV; r++)
#pragma unroll
                    for (int c = 0; c < NV; c++) P[r][c] = fma(I[r][kk], a[c], P[r][c]);
            }
        }
        if (live) {
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    Dinv[(size_t)(r*NV + c)*pl + o] = I[r][c]; DX[(size_t)(r*NV + c)*pl + o] = DXv[r][c]; DY[(size_t)(r*NV + c)*pl + o] = P[r][c];
                }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- the middle row: D* = D - A' DC_{m-1} - C' DA_{m+1}
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
        for (int c = 0; c < NV; c++) ex[((size_t)up*B + r*NV + c)*32 + lane] = P[r][c];
    __syncthreads();
    if (up) return;
    {
        const int jl = m, gj = v.j0 + jl;
        const size_t o = v.at(jl + JOFF, ic + IOFF);
        const bool has_a = jl > 0, has_c = jl + 1 < n;
        const bool arm_a = arms && gj - 2 >= 0, arm_c = arms && gj + 2 <= v.njc - 1;
        auto gblock = [&](int sl, int sla, bool arm, bool diag, double (&a)[NV][NV]) {     // straight from the Jacobian planes
            const double idt = (diag && op == OP_LHS) ? 1.0/dt[o] : 0.0;
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    double vj = J[((size_t)sl*B + r*NV + c)*pl + o];
                    if (arm) vj += J[((size_t)sla*B + r*NV + c)*pl + o];
                    a[r][c] = op == OP_LHS ? ((diag && r == c) ? idt - vj : -vj) : vj;
                }
        };
        double D[NV][NV], A[NV][NV], C[NV][NV], I[NV][NV];
        gblock(0, 0, false, true, D);
#pragma unroll
        for (int r = 0; r < NV; r++)
#pragma unroll
            for (int c = 0; c < NV; c++) { A[r][c] = 0.0; C[r][c] = 0.0; }
        if (has_a) {
            gblock(3, 11, arm_a, false, A);
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int kk = 0; kk < NV; kk++)
#pragma unroll
                    for (int c = 0; c < NV; c++) D[r][c] = fma(-A[r][kk], ex[((size_t)kk*NV + c)*32 + lane], D[r][c]);
        }
        if (has_c) {
            gblock(4, 12, arm_c, false, C);
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int kk = 0; kk < NV; kk++)
#pragma unroll
                    for (int c = 0; c < NV; c++) D[r][c] = fma(-C[r][kk], ex[((size_t)B + kk*NV + c)*32 + lane], D[r][c]);
        }
        if (!invert_block_fast<NV>(D, I) && live) atomicExch(err, 1);
        if (live) {
#pragma unroll
            for (int r = 0; r < NV; r++)
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    double sa = 0.0, sc = 0.0;
#pragma unroll
                    for (int kk = 0; kk < NV; kk++) { sa = fma(I[r][kk], A[kk][c], sa); sc = fma(I[r][kk], C[kk][c], sc); }
                    Dinv[(size_t)(r*NV + c)*pl + o] = I[r][c]; DAo[(size_t)(r*NV + c)*pl + o] = sa; DCo[(size_t)(r*NV + c)*pl + o] = sc;
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row-parallel form of line_factor_kernel (round 2).  The recurrence is sequential in j and the kernel above gives the
// whole GPU nic/32 warps (128 at 4096^2: less than one per SM), each executing ~1 500 dependent instructions per row.
// Here a line is worked on by NV lanes -- lane (l, r) holds ROW r of every block of line l -- so a warp covers 32/NV lines,
// the machine gets NV x the warps, and the per-row dependency chain shrinks to a row's share:
//   products   T = A B      : lane r forms row r of T; the rows of B come from the line's other lanes by shuffle
//   inversion  Gauss-Jordan : per column the pivot lane (largest |entry| among the rows not yet used, the choice partial
//                             pivoting makes) scales and broadcasts its row, all other lanes eliminate; no row swaps --
//                             the lane that pivoted on column k ends up holding row k of the inverse, one final gather
// Same arithmetic per row as invert_block / matmul_block.  Rows ahead are streamed through a cp.async ring exactly like
// above, each lane copying the words of ITS row of the five blocks (+ 1/dt).
// ------------------------------------------------------------------------------------------------
constexpr int FACTR_STAGES = 6;
template <int NV> constexpr int factr_lines() { return 32/NV; }
template <int NV> constexpr int factr_words() { return 5*NV + 1; }                 // per lane and stage
template <int NV> constexpr size_t factr_ring_bytes() { return (size_t)FACTR_STAGES*factr_words<NV>()*32*sizeof(double); }

template <int NV>
__global__ void __launch_bounds__(32) line_factor_rows_kernel(View v, const double* __restrict__ J, const double* __restrict__ dt, int op, int nslots,
                                                              double* __restrict__ F, int* __restrict__ err) {
    extern __shared__ double fring[];
    constexpr int B = NV*NV, LPW = 32/NV, NW = 5*NV + 1, S = FACTR_STAGES;
    const int lane = threadIdx.x;
    const int l = lane/NV, r = lane - l*NV;                        // line within the warp, block row
    const bool lane_ok = l < LPW;
    const int i = blockIdx.x*LPW + l;
    const bool live = lane_ok && i < v.nic;
    const int ic = (lane_ok && i < v.nic) ? i : v.nic - 1;          // idle lanes shadow the last line and never store
    const int base = (lane_ok ? l : 0)*NV;                          // first lane of this line's group (idle lanes mirror group 0)
    const unsigned full = 0xffffffffu;
    const size_t pl = v.plane;
    double* __restrict__ Dinv = F;
    double* __restrict__ DA = F + (size_t)NV*NV*pl;
    double* __restrict__ DC = F + (size_t)2*NV*NV*pl;
    const bool arms = nslots > 9;
    auto word = [&](int st, int w) -> double* { return fring + ((size_t)st*NW + w)*32 + lane; };
    auto issue = [&](int jl) {                                     // ring group g = 0..4 <- Jacobian slots 0, 3, 11, 4, 12 (row r of each)
        if (jl < v.njl) {
            const size_t o = v.at(jl + JOFF, ic + IOFF);
            const int st = jl % S;
#pragma unroll
            for (int g = 0; g < 5; g++) {
                const int sl = g == 0 ? 0 : (g == 1 ? 3 : (g == 2 ? 11 : (g == 3 ? 4 : 12)));
                if ((g == 2 || g == 4) && !arms) continue;
#pragma unroll
                for (int c = 0; c < NV; c++) fact_cp_async8(word(st, g*NV + c), J + ((size_t)sl*B + r*NV + c)*pl + o);
            }
            if (op == OP_LHS) fact_cp_async8(word(st, 5*NV), dt + o);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto ring_row = [&](int st, int g, double idt, bool diag, double (&a)[NV]) {   // A = -J + I/dt for OP_LHS, J otherwise
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const double vj = *word(st, g*NV + c);
            a[c] = op == OP_LHS ? ((diag && r == c) ? idt - vj : -vj) : vj;
        }
    };
    // row r of X Y, the rows of Y held by the line's lanes
    auto rowmul = [&](const double (&x)[NV], const double (&y)[NV], double (&t)[NV]) {
#pragma unroll
        for (int c = 0; c < NV; c++) t[c] = 0.0;
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
            for (int c = 0; c < NV; c++) t[c] += x[k]*__shfl_sync(full, y[c], base + k);
    };
    for (int jl = 0; jl < S - 1; jl++) issue(jl);
    double DCp[NV];                                                // row r of DC_{j-1}
#pragma unroll
    for (int c = 0; c < NV; c++) DCp[c] = 0.0;
    for (int jl = 0; jl < v.njl; jl++) {
        issue(jl + S - 1);
        asm volatile("cp.async.wait_group %0;" :: "n"(S - 1) : "memory");
        const int st = jl % S;
        const int gj = v.j0 + jl;
        const size_t o = v.at(jl + JOFF, ic + IOFF);
        double D[NV], A[NV], T[NV], I[NV];
        ring_row(st, 0, op == OP_LHS ? 1.0/(*word(st, 5*NV)) : 0.0, true, D);
        const bool lo = jl > 0, hi = jl + 1 < v.njl;
        if (lo) {
            ring_row(st, 1, 0.0, false, A);
            if (arms && gj - 2 >= 0) {
                ring_row(st, 2, 0.0, false, T);
#pragma unroll
                for (int c = 0; c < NV; c++) A[c] += T[c];
            }
            rowmul(A, DCp, T);
#pragma unroll
            for (int c = 0; c < NV; c++) D[c] -= T[c];
        } else {
#pragma unroll
            for (int c = 0; c < NV; c++) A[c] = 0.0;
        }
        // ---- Gauss-Jordan across the NV lanes of the line
#pragma unroll
        for (int c = 0; c < NV; c++) I[c] = r == c ? 1.0 : 0.0;
        bool used = false; int pcol = -1;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double mine = used ? -1.0 : fabs(D[k]);
            double big = -1.0; int p = 0;
#pragma unroll
            for (int q = 0; q < NV; q++) {                         // first maximum in lane order = the row partial pivoting swaps in
                const double t = __shfl_sync(full, mine, base + q);
                if (t > big) { big = t; p = q; }
            }
            const double piv = __shfl_sync(full, D[k], base + p);
            if (piv == 0.0 && lane_ok) atomicExch(err, 1);
            const double ip = 1.0/(piv == 0.0 ? 1.0 : piv);
            const bool me = r == p;
            if (me) {
#pragma unroll
                for (int c = 0; c < NV; c++) { D[c] *= ip; I[c] *= ip; }
                used = true; pcol = k;
            }
            const double f = me ? 0.0 : D[k];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const double pd = __shfl_sync(full, D[c], base + p), pi = __shfl_sync(full, I[c], base + p);
                if (!me) { D[c] -= f*pd; I[c] -= f*pi; }
            }
        }
        {   // the lane that pivoted on column k holds row k of the inverse: bring row r to lane r
            int src = 0;
#pragma unroll
            for (int q = 0; q < NV; q++) { const int pc = __shfl_sync(full, pcol, base + q); if (pc == r) src = q; }
#pragma unroll
            for (int c = 0; c < NV; c++) I[c] = __shfl_sync(full, I[c], base + src);
        }
        rowmul(I, A, T);                                           // DA = D'^-1 A'
        if (live) {
#pragma unroll
            for (int c = 0; c < NV; c++) { Dinv[(size_t)(r*NV + c)*pl + o] = I[c]; DA[(size_t)(r*NV + c)*pl + o] = T[c]; }
        }
        if (hi) {
            ring_row(st, 3, 0.0, false, A);
            if (arms && gj + 2 <= v.njc - 1) {
                ring_row(st, 4, 0.0, false, T);
#pragma unroll
                for (int c = 0; c < NV; c++) A[c] += T[c];
            }
            rowmul(I, A, DCp);                                     // DC = D'^-1 C'
        } else {
#pragma unroll
            for (int c = 0; c < NV; c++) DCp[c] = 0.0;
        }
        if (live) {
#pragma unroll
            for (int c = 0; c < NV; c++) DC[(size_t)(r*NV + c)*pl + o] = DCp[c];
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// z = M^-1 r:  forward  t_j = Dinv_j r_j - DA_j t_{j-1},  backward  z_j = t_j - DC_j z_{j+1}.
// transpose (M = L U with L = blockdiag(D') + lower(A'), U = I + upper(DC), so M^T = U^T L^T):
//           forward  y_j = r_j - DC_{j-1}^T y_{j-1},  backward  w_j = y_j - DA_{j+1}^T w_{j+1},  z_j = Dinv_j^T w_j.
//
// The recurrence is sequential in j, so the kernel is latency- not bandwidth-limited unless the rows ahead are
// already on chip: one warp owns 32 adjacent lines and streams the rows it is about to need through a
// LINE_STAGES-deep shared-memory ring with cp.async (LDGSTS), each lane copying exactly the 8-byte words it
// will read itself -- so no barrier is needed, only cp.async.wait_group -- and every global access is a
// coalesced 256-byte row segment per plane.
constexpr int LINE_STAGES = 12;

__device__ __forceinline__ void cp_async8(double* smem, const double* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int NV> constexpr int line_ring_planes() { return 2*(NV*NV + (NV*NV & 1)) + NV + (NV & 1); }   // each group padded to an even plane count
template <int NV> constexpr size_t line_ring_bytes() { return (size_t)LINE_STAGES*line_ring_planes<NV>()*32*sizeof(double); }
__device__ __forceinline__ void cp_async16(double* smem, const double* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
}

// One warp owns 32 adjacent lines (a CTA is one warp) and walks them row by row; the sweep is sequential in j, i.e. pure latency,
// and only nic/32 warps exist, so what bounds it is the warp's own INSTRUCTION count per row -- and of the ~300 instructions per
// row and sweep, 55 were eight-byte cp.async (LDGSTS costs ~8 cycles of the load/store unit each, whatever its size).  Round 2b:
// the ring is filled with SIXTEEN-byte copies: one instruction moves the 256-byte row segments of TWO planes (lanes 0-15 the
// first, lanes 16-31 the second), 28 instead of 55 per row; the lanes then read words other lanes copied, so the wait is
// followed by a __syncwarp.  Same arithmetic in the same order (results bit-identical).  Measured and rejected: one TMA bulk copy
// (cp.async.bulk + mbarrier) per plane and row -- 256-byte bulk copies complete at ~1 per 50 cycles and SM: 3.1 -> 11.9 ms.
// Lanes past the last line compute on padding and never store.
template <int NV, bool TR>
__global__ void __launch_bounds__(32) line_apply_kernel(View v, const double* __restrict__ F, const double* __restrict__ rv, double* __restrict__ z) {
    extern __shared__ __align__(128) double ring[];
    // ring planes of a stage: [0, B) first block, [BP, BP + B) second block, [2 BP, 2 BP + NV) the operand; BP = B rounded up to even,
    // so that every sixteen-byte copy instruction serves two planes of the SAME source array (no per-lane source selection)
    constexpr int B = NV*NV, BP = B + (B & 1), NP = 2*BP + NV + (NV & 1), S = LINE_STAGES;
    const int lane = threadIdx.x;
    const int i = blockIdx.x*32 + lane;
    const bool live = i < v.nic;
    const int c0 = blockIdx.x*32 + IOFF;                           // first plane column of the warp's segment (16-byte aligned)
    const size_t pl = v.plane;
    const double* __restrict__ Dinv = F;
    const double* __restrict__ DA = F + (size_t)B*pl;
    const double* __restrict__ DC = F + (size_t)2*B*pl;
    auto slot = [&](int st, int p) -> double* { return ring + ((size_t)st*NP + p)*32 + lane; };
    const int half = lane >> 4, word = (lane & 15)*2;               // this lane copies words word, word+1 of plane 2h + half
    const bool in_row = c0 + word < v.pitch;                        // the last warp's segment ends with the row
    const size_t lane_src = (size_t)half*pl + word;                 // this lane's offset inside a plane pair
    const unsigned lane_dst = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)((half*32 + word)*sizeof(double));
    for (int k = lane; k < S*NP*32; k += 32) ring[k] = 0.0;        // words past a short last segment are never written by the copies
    __syncwarp();
    // one commit group per row: blocks P0 (and P1 unless null) and the operand's nv planes, two planes per instruction
    auto issue_row = [&](int st, int jl, bool valid, const double* P0, const double* P1, const double* vec) {
        if (valid && in_row) {
            const size_t o = v.at(jl + JOFF, c0) + lane_src;
            const unsigned d = lane_dst + (unsigned)(st*NP*32*sizeof(double));
            auto group = [&](const double* src, int first, int count) {
#pragma unroll
                for (int h = 0; h < (count + 1)/2; h++)
                    if (2*h + 1 < count || half == 0)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)((first + 2*h)*32*sizeof(double))), "l"(src + o + (size_t)(2*h)*pl) : "memory");
            };
            group(P0, 0, B);
            if (P1) group(P1, BP, B);
            group(vec, 2*BP, NV);
        }
        cp_async_commit();
    };
    double t[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) t[r] = 0.0;

    // ---- sweep 1: rows ascending
    {
        const double* __restrict__ P0 = TR ? DC : Dinv;            // first nv*nv planes of a stage
        const double* __restrict__ P1 = TR ? nullptr : DA;
        for (int jl = 0; jl < S - 1; jl++) issue_row(jl % S, jl, jl < v.njl, P0, P1, rv);
        for (int jl = 0; jl < v.njl; jl++) {
            __syncwarp();                                          // every lane is done with the stage row jl + S - 1 overwrites
            issue_row((jl + S - 1) % S, jl + S - 1, jl + S - 1 < v.njl, P0, P1, rv);
            const int st = jl % S;
            cp_async_wait<S - 1>();                                // this lane's copies of row jl have landed ...
            __syncwarp();                                          // ... and so have the other lanes'
            const size_t o = v.at(jl + JOFF, c0 + lane);
            double g[NV];
            if (!TR) {
                double rr[NV];
#pragma unroll
                for (int r = 0; r < NV; r++) rr[r] = *slot(st, 2*BP + r);
#pragma unroll
                for (int r = 0; r < NV; r++) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < NV; c++) s += *slot(st, r*NV + c)*rr[c];
                    g[r] = s;
                }
#pragma unroll
                for (int r = 0; r < NV; r++) {
                    double s = g[r];
#pragma unroll
                    for (int c = 0; c < NV; c++) s -= *slot(st, BP + r*NV + c)*t[c];
                    g[r] = s;
                }
#pragma unroll
                for (int r = 0; r < NV; r++) t[r] = g[r];
                if (live) {
#pragma unroll
                    for (int r = 0; r < NV; r++) z[r*pl + o] = g[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < NV; r++) g[r] = *slot(st, 2*BP + r) - t[r];
                if (live) {
#pragma unroll
                    for (int r = 0; r < NV; r++) z[r*pl + o] = g[r];
                }
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NV; r++) s += *slot(st, r*NV + c)*g[r];
                    t[c] = s;
                }
            }
        }
    }
    cp_async_wait<0>();
    __threadfence();                                               // the z rows written above are copied in below by OTHER lanes
    __syncwarp();

    // ---- sweep 2: rows descending
    {
        const double* __restrict__ P0 = TR ? DA : DC;
        const double* __restrict__ P1 = TR ? Dinv : nullptr;
        const int jtop = TR ? v.njl - 1 : v.njl - 2;               // the untransposed last row is already final (t holds it)
        if (TR) {
#pragma unroll
            for (int r = 0; r < NV; r++) t[r] = 0.0;
        }
        // n-th row of this sweep = row jtop - n
        for (int n = 0; n < S - 1; n++) issue_row(n % S, jtop - n, n <= jtop, P0, P1, z);
        for (int n = 0; n <= jtop; n++) {
            __syncwarp();
            issue_row((n + S - 1) % S, jtop - (n + S - 1), n + S - 1 <= jtop, P0, P1, z);
            const int st = n % S;
            cp_async_wait<S - 1>();
            __syncwarp();
            const int jl = jtop - n;
            const size_t o = v.at(jl + JOFF, c0 + lane);
            double g[NV];
            if (!TR) {
#pragma unroll
                for (int r = 0; r < NV; r++) {
                    double s = *slot(st, 2*BP + r);
#pragma unroll
                    for (int c = 0; c < NV; c++) s -= *slot(st, r*NV + c)*t[c];
                    g[r] = s;
                }
#pragma unroll
                for (int r = 0; r < NV; r++) t[r] = g[r];
                if (live) {
#pragma unroll
                    for (int r = 0; r < NV; r++) z[r*pl + o] = g[r];
                }
            } else {
                double w[NV];
#pragma unroll
                for (int r = 0; r < NV; r++) w[r] = *slot(st, 2*BP + r) - t[r];
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NV; r++) s += *slot(st, r*NV + c)*w[r];
                    t[c] = s;
                }
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NV; r++) s += *slot(st, BP + r*NV + c)*w[r];
                    g[c] = s;
                }
                if (live) {
#pragma unroll
                    for (int r = 0; r < NV; r++) z[r*pl + o] = g[r];
                }
            }
        }
        cp_async_wait<0>();
    }
}

// ------------------------------------------------------------------------------------------------
// TWISTED line solve (round 2b).  The sweeps above are sequential in j and what bounds them is one warp's instruction stream per
// row, so the lever is the LENGTH of the sequence: the line is eliminated from BOTH ends towards its middle row m = njl/2 by two
// warps of one CTA -- rows 0 .. m-1 ascending with (Dinv, DA) exactly as before, rows njl-1 .. m+1 descending with the roles of the
// sub- and super-diagonal blocks exchanged (factors from line_factor_twisted_kernel: D''_j = D_j - C_j DA_{j+1}) -- the two halves
// meet in the middle row, x_m = Dinv_m r_m - DA_m y_{m-1} - DC_m y_{m+1}, and both substitute back outwards.  The same
// block-tridiagonal system, hence the same preconditioner up to rounding; half the dependent rows per warp.  The transposed
// solve is the adjoint of that sequence: inward sweeps with the transposed back-substitution blocks, the middle row, outward
// sweeps with the transposed elimination blocks.
//   mode 0  inward,  A x = r :  g = P0 r - P1 t;  t = g;                 z = g      (P0 = Dinv, P1 = DA below / DC above)
//   mode 1  outward, A x = r :  g = y - P0 t;     t = g;                 z = g      (P0 = DC below / DA above)
//   mode 2  inward,  A^T     :  g = r - t;        t = P0^T g;            z = g      (P0 = DC below / DA above)
//   mode 3  outward, A^T     :  w = y - t;        t = P0^T w;            z = P1^T w (P0 = DA below / DC above, P1 = Dinv)
// ------------------------------------------------------------------------------------------------
constexpr int TW_STAGES = 6;
template <int NV> constexpr size_t line_twisted_bytes() { return (size_t)2*TW_STAGES*line_ring_planes<NV>()*32*sizeof(double) + 2*NV*32*sizeof(double); }

template <int NV, int MODE>
__device__ __forceinline__ void twisted_sweep(const View& v, double* __restrict__ ring, int lane, int c0, bool live, int row0, int step, int count,
                                              const double* __restrict__ P0, const double* __restrict__ P1, const double* __restrict__ vec,
                                              double* __restrict__ z, double (&t)[NV]) {
    constexpr int B = NV*NV, BP = B + (B & 1), NP = 2*BP + NV + (NV & 1), S = TW_STAGES;
    const size_t pl = v.plane;
    const int half = lane >> 4, word = (lane & 15)*2;
    const bool in_row = c0 + word < v.pitch;
    const size_t lane_src = (size_t)half*pl + word;
    const unsigned lane_dst = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)((half*32 + word)*sizeof(double));
    auto slot = [&](int st, int p) -> double* { return ring + ((size_t)st*NP + p)*32 + lane; };
    auto issue_row = [&](int n) {                                  // n-th row of this sweep = row row0 + n step
        if (n < count && in_row) {
            const size_t o = v.at(row0 + n*step + JOFF, c0) + lane_src;
            const unsigned d = lane_dst + (unsigned)((n % S)*NP*32*sizeof(double));
            auto group = [&](const double* src, int first, int cnt) {
#pragma unroll
                for (int h = 0; h < (cnt + 1)/2; h++)
                    if (2*h + 1 < cnt || half == 0)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)((first + 2*h)*32*sizeof(double))), "l"(src + o + (size_t)(2*h)*pl) : "memory");
            };
            group(P0, 0, B);
            if (MODE == 0 || MODE == 3) group(P1, BP, B);
            group(vec, 2*BP, NV);
        }
        cp_async_commit();
    };
    for (int n = 0; n < S - 1; n++) issue_row(n);
    for (int n = 0; n < count; n++) {
        __syncwarp();                                              // every lane is done with the stage the next copy overwrites
        issue_row(n + S - 1);
        const int st = n % S;
        cp_async_wait<S - 1>();
        __syncwarp();
        const size_t o = v.at(row0 + n*step + JOFF, c0 + lane);
        double g[NV];
        if (MODE == 0) {
            double rr[NV];
#pragma unroll
            for (int r = 0; r < NV; r++) rr[r] = *slot(st, 2*BP + r);
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NV; c++) s += *slot(st, r*NV + c)*rr[c];
                g[r] = s;
            }
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double s = g[r];
#pragma unroll
                for (int c = 0; c < NV; c++) s -= *slot(st, BP + r*NV + c)*t[c];
                g[r] = s;
            }
#pragma unroll
            for (int r = 0; r < NV; r++) t[r] = g[r];
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double s = *slot(st, 2*BP + r);
#pragma unroll
                for (int c = 0; c < NV; c++) s -= *slot(st, r*NV + c)*t[c];
                g[r] = s;
            }
#pragma unroll
            for (int r = 0; r < NV; r++) t[r] = g[r];
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < NV; r++) g[r] = *slot(st, 2*BP + r) - t[r];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < NV; r++) s += *slot(st, r*NV + c)*g[r];
                t[c] = s;
            }
        } else {
            double w[NV];
#pragma unroll
            for (int r = 0; r < NV; r++) w[r] = *slot(st, 2*BP + r) - t[r];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < NV; r++) s += *slot(st, r*NV + c)*w[r];
                t[c] = s;
            }
#pragma unroll
            for (int c = 0; c < NV; c++) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < NV; r++) s += *slot(st, BP + r*NV + c)*w[r];
                g[c] = s;
            }
        }
        if (live) {
#pragma unroll
            for (int r = 0; r < NV; r++) z[r*pl + o] = g[r];
        }
    }
    cp_async_wait<0>();
}

template <int NV, bool TR>
__global__ void __launch_bounds__(64) line_apply_twisted_kernel(View v, const double* __restrict__ F, const double* __restrict__ rv, double* __restrict__ z) {
    extern __shared__ __align__(128) double tw_smem[];
    constexpr int B = NV*NV, BP = B + (B & 1), NP = 2*BP + NV + (NV & 1), S = TW_STAGES;
    const int lane = threadIdx.x & 31, up = threadIdx.x >> 5;      // warp 0: the rows below the middle row, warp 1: the rows above it
    double* ring = tw_smem + (size_t)up*S*NP*32;
    double* ex = tw_smem + (size_t)2*S*NP*32;                      // [2][NV][32]: what each half hands to the middle row
    const int i = blockIdx.x*32 + lane;
    const bool live = i < v.nic;
    const int c0 = blockIdx.x*32 + IOFF;
    const size_t pl = v.plane;
    const double* __restrict__ Dinv = F;
    const double* __restrict__ DA = F + (size_t)B*pl;
    const double* __restrict__ DC = F + (size_t)2*B*pl;
    for (int k = lane; k < S*NP*32; k += 32) ring[k] = 0.0;        // words past a short last segment are never written by the copies
    __syncwarp();
    const int n = v.njl, m = n/2;
    const int row0 = up ? n - 1 : 0, step = up ? -1 : 1, count = up ? n - 1 - m : m;
    double t[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) t[r] = 0.0;
    // ---- inward
    if (!TR) twisted_sweep<NV, 0>(v, ring, lane, c0, live, row0, step, count, Dinv, up ? DC : DA, rv, z, t);
    else twisted_sweep<NV, 2>(v, ring, lane, c0, live, row0, step, count, up ? DA : DC, nullptr, rv, z, t);
#pragma unroll
    for (int r = 0; r < NV; r++) ex[(up*NV + r)*32 + lane] = t[r];
    __threadfence();                                               // the z rows written above are copied in below by OTHER lanes
    __syncthreads();
    // ---- the middle row (both warps evaluate it; warp 0 stores it)
    {
        const int ic = imin(c0 + lane, v.pitch - 1);               // lanes past the last line shadow a valid column and never store
        const size_t o = v.at(m + JOFF, ic);
        double tl[NV], tu[NV], rm[NV], xm[NV];
#pragma unroll
        for (int r = 0; r < NV; r++) { tl[r] = ex[r*32 + lane]; tu[r] = ex[(NV + r)*32 + lane]; rm[r] = rv[r*pl + o]; }
        if (!TR) {
#pragma unroll
            for (int r = 0; r < NV; r++) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NV; c++) s += Dinv[(size_t)(r*NV + c)*pl + o]*rm[c] - DA[(size_t)(r*NV + c)*pl + o]*tl[c] - DC[(size_t)(r*NV + c)*pl + o]*tu[c];
                xm[r] = s;
            }
#pragma unroll
            for (int r = 0; r < NV; r++) t[r] = xm[r];
            if (live && !up) {
#pragma unroll
                for (int r = 0; r < NV; r++) z[r*pl + o] = xm[r];
            }
        } else {
            double w[NV];
#pragma unroll
            for (int r = 0; r < NV; r++) w[r] = rm[r] - tl[r] - tu[r];
            const double* __restrict__ Pt = up ? DC : DA;          // what the outward sweep of this half subtracts first
#pragma unroll
            for (int c = 0; c < NV; c++) {
                double s = 0.0, so = 0.0;
#pragma unroll
                for (int r = 0; r < NV; r++) { s += Pt[(size_t)(r*NV + c)*pl + o]*w[r]; so += Dinv[(size_t)(r*NV + c)*pl + o]*w[r]; }
                t[c] = s; xm[c] = so;
            }
            if (live && !up) {
#pragma unroll
                for (int r = 0; r < NV; r++) z[r*pl + o] = xm[r];
            }
        }
    }
    // ---- outward: the rows of this half in the opposite order
    const int orow0 = up ? m + 1 : m - 1, ostep = -step;
    if (!TR) twisted_sweep<NV, 1>(v, ring, lane, c0, live, orow0, ostep, count, up ? DA : DC, nullptr, z, z, t);
    else twisted_sweep<NV, 3>(v, ring, lane, c0, live, orow0, ostep, count, up ? DC : DA, Dinv, z, z, t);
}

// ------------------------------------------------------------------------------------------------
// Flat vector kernels (n = plane*nv elements; ghosts/padding are zero in every operand)
// ------------------------------------------------------------------------------------------------
constexpr int DOT_GROUP = 16;
constexpr int DOT_THREADS = 256;

// The last block to arrive (atomic ticket) sums the per-block partials of `cnt` columns in block order -- deterministic,
// and no separate reduction launch (one Gram-Schmidt sweep was 3 launches + 2 reductions per 8 basis vectors).
__device__ __forceinline__ void last_block_reduce(double* __restrict__ partial, int ldp, int col0, int cnt, double* __restrict__ out, unsigned* ticket) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = wid; c < cnt; c += blockDim.x >> 5) {             // one warp per column, lanes stride over the blocks: fixed order
        double s = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) s += ((volatile double*)partial)[(size_t)b*ldp + col0 + c];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) out[col0 + c] = s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// Both Gram-Schmidt kernels walk the vectors TILE-major: a block holds GS_PER consecutive elements per thread of w in
// registers and visits the basis vectors one after the other, so at any moment it streams ONE 16 KB-contiguous run of
// one vector (an element-major loop reads k + 2 vectors at once per warp: ~20 concurrently open DRAM pages per block,
// measured 3.5 TB/s; tile-major restores long sequential bursts).
constexpr int GS_PER = 8;                                          // elements per thread and tile
constexpr int GS_TILE = DOT_THREADS*GS_PER;

// out[j0 .. j0+cnt) = w . V_j  (cnt <= DOT_GROUP basis vectors per pass over w)
__global__ void __launch_bounds__(DOT_THREADS) dots_kernel(const double* __restrict__ w, const double* __restrict__ V, size_t n, int cnt,
                                                           double* __restrict__ partial, int ldp, int j0, double* __restrict__ out, unsigned* ticket) {
    double acc[DOT_GROUP];
#pragma unroll
    for (int j = 0; j < DOT_GROUP; j++) acc[j] = 0.0;
    for (size_t t0 = (size_t)blockIdx.x*GS_TILE; t0 < n; t0 += (size_t)gridDim.x*GS_TILE) {
        double we[GS_PER];
#pragma unroll
        for (int p = 0; p < GS_PER; p++) { const size_t e = t0 + (size_t)p*DOT_THREADS + threadIdx.x; we[p] = e < n ? w[e] : 0.0; }
#pragma unroll
        for (int j = 0; j < DOT_GROUP; j++) {
            if (j < cnt) {
                const double* Vj = V + (size_t)j*n;
                double a = 0.0;
#pragma unroll
                for (int p = 0; p < GS_PER; p++) { const size_t e = t0 + (size_t)p*DOT_THREADS + threadIdx.x; a += we[p]*(e < n ? Vj[e] : 0.0); }
                acc[j] += a;
            }
        }
    }
    __shared__ double ws[DOT_GROUP][DOT_THREADS/32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < DOT_GROUP; j++) {
        double s = acc[j];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) ws[j][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
        double s = 0.0;
        for (int k = 0; k < DOT_THREADS/32; k++) s += ws[threadIdx.x][k];
        partial[(size_t)blockIdx.x*ldp + j0 + threadIdx.x] = s;
    }
    last_block_reduce(partial, ldp, j0, cnt, out, ticket);
}

// w -= sum_j h[j] V_j  and, in the same pass, out[ncol] = |w_new|^2  (h on the device: no host round trip between the
// projection and the update, and no extra pass over w for the norm of the next basis vector)
__global__ void __launch_bounds__(DOT_THREADS) gs_update_kernel(double* __restrict__ w, const double* __restrict__ V, size_t n, int cnt, const double* __restrict__ h,
                                                                double* __restrict__ partial, int ldp, int ncol, double* __restrict__ out, unsigned* ticket) {
    double acc = 0.0;
    for (size_t t0 = (size_t)blockIdx.x*GS_TILE; t0 < n; t0 += (size_t)gridDim.x*GS_TILE) {
        double s[GS_PER];
#pragma unroll
        for (int p = 0; p < GS_PER; p++) { const size_t e = t0 + (size_t)p*DOT_THREADS + threadIdx.x; s[p] = e < n ? w[e] : 0.0; }
        for (int j = 0; j < cnt; j++) {
            const double hj = h[j];
            const double* Vj = V + (size_t)j*n;
#pragma unroll
            for (int p = 0; p < GS_PER; p++) { const size_t e = t0 + (size_t)p*DOT_THREADS + threadIdx.x; s[p] -= hj*(e < n ? Vj[e] : 0.0); }
        }
#pragma unroll
        for (int p = 0; p < GS_PER; p++) { const size_t e = t0 + (size_t)p*DOT_THREADS + threadIdx.x; if (e < n) { w[e] = s[p]; acc += s[p]*s[p]; } }
    }
    __shared__ double ws[DOT_THREADS/32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) ws[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < DOT_THREADS/32; k++) s += ws[k];
        partial[(size_t)blockIdx.x*ldp + ncol] = s;
    }
    last_block_reduce(partial, ldp, ncol, 1, out, ticket);
}

// The flat kernels move two doubles per access and keep four accesses per thread in flight (n is a multiple of 16: planes are
// rows x pitch with pitch a multiple of 16); scalar loops at 4 x SMs blocks ran at 4-5 TB/s.
constexpr int FLAT_U = 4;
// dst = src * (1/sqrt(*normsq))
__global__ void scale_rsqrt_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n, const double* __restrict__ normsq) {
    const double sc = 1.0/sqrt(*normsq);
    const size_t n2 = n/2, stride = (size_t)gridDim.x*blockDim.x;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
    for (size_t e = (size_t)blockIdx.x*blockDim.x + threadIdx.x; e < n2; e += FLAT_U*stride) {
        double2 v[FLAT_U];
#pragma unroll
        for (int u = 0; u < FLAT_U; u++) if (e + u*stride < n2) v[u] = s2[e + u*stride];
#pragma unroll
        for (int u = 0; u < FLAT_U; u++) if (e + u*stride < n2) d2[e + u*stride] = make_double2(v[u].x*sc, v[u].y*sc);
    }
}

// dst = sum_j y[j] V_j  (y passed by value through a small device array)
__global__ void combine_kernel(double* __restrict__ dst, const double* __restrict__ V, size_t n, int cnt, const double* __restrict__ y) {
    const size_t n2 = n/2, stride = (size_t)gridDim.x*blockDim.x;
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
    for (size_t e = (size_t)blockIdx.x*blockDim.x + threadIdx.x; e < n2; e += stride) {
        double2 s = make_double2(0.0, 0.0);
        int j = 0;
        for (; j + FLAT_U <= cnt; j += FLAT_U) {
            double2 v[FLAT_U];
#pragma unroll
            for (int u = 0; u < FLAT_U; u++) v[u] = reinterpret_cast<const double2*>(V + (size_t)(j + u)*n)[e];
#pragma unroll
            for (int u = 0; u < FLAT_U; u++) { s.x += y[j + u]*v[u].x; s.y += y[j + u]*v[u].y; }
        }
        for (; j < cnt; j++) { const double2 v = reinterpret_cast<const double2*>(V + (size_t)j*n)[e]; s.x += y[j]*v.x; s.y += y[j]*v.y; }
        d2[e] = s;
    }
}

// y = a*x + b*y
__global__ void axpby_kernel(double* __restrict__ y, const double* __restrict__ x, size_t n, double a, double b) {
    const size_t n2 = n/2, stride = (size_t)gridDim.x*blockDim.x;
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
    double2* __restrict__ y2 = reinterpret_cast<double2*>(y);
    for (size_t e = (size_t)blockIdx.x*blockDim.x + threadIdx.x; e < n2; e += FLAT_U*stride) {
        double2 vx[FLAT_U], vy[FLAT_U];
#pragma unroll
        for (int u = 0; u < FLAT_U; u++) if (e + u*stride < n2) { vx[u] = x2[e + u*stride]; vy[u] = y2[e + u*stride]; }
#pragma unroll
        for (int u = 0; u < FLAT_U; u++) if (e + u*stride < n2) y2[e + u*stride] = make_double2(a*vx[u].x + b*vy[u].x, a*vx[u].y + b*vy[u].y);
    }
}

} // namespace sg
