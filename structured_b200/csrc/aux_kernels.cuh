// Small kernels around the residual: metrics, boundary conditions, local time step, stage updates,
// AoS<->SoA transposes at the host boundary, halo packing and the final norm reduction.
#pragma once
#include "common.cuh"
#include "../../include/structured_gpu.h"

namespace sg {

// ------------------------------------------------------------------------------------------------
// Mesh::calc_metrics (src/utils/mesh.cpp:172-205) from the vertex planes.
// vertex (i, j) lives at (r = j - j0 + JOFF, c = i + IOFF) of xv / yv (planes of rows+1 rows).
// ------------------------------------------------------------------------------------------------
__global__ void metrics_kernel(View v, const double* __restrict__ xv, const double* __restrict__ yv,
                               double* __restrict__ ncx, double* __restrict__ ncy,
                               double* __restrict__ nex, double* __restrict__ ney, double* __restrict__ vol) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= v.pitch || r >= v.rows) return;
    const int i = c - IOFF, j = r - JOFF + v.j0;
    const size_t o = v.at(r, c);
    double a_ncx = 0, a_ncy = 0, a_nex = 0, a_ney = 0, a_vol = 1.0;
    const bool vi = i >= 0 && i < v.ni, vi1 = i + 1 >= 0 && i + 1 < v.ni;
    const bool vj = j >= 0 && j < v.nj, vj1 = j + 1 >= 0 && j + 1 < v.nj;
    double retax = 0, retay = 0, rchix = 0, rchiy = 0;
    if (vi && vi1 && vj) {            // eta face (i, j): edge (i,j)->(i+1,j)            :176-183
        retax = xv[v.at(r, c + 1)] - xv[o]; retay = yv[v.at(r, c + 1)] - yv[o];
        a_nex = -retay; a_ney = retax;
    }
    if (vi && vj && vj1) {            // chi face (i, j): edge (i,j)->(i,j+1)            :185-192
        rchix = xv[v.at(r + 1, c)] - xv[o]; rchiy = yv[v.at(r + 1, c)] - yv[o];
        a_ncx = rchiy; a_ncy = -rchix;
    }
    if (vi && vi1 && vj && vj1) {     // volume                                          :196-197
        const double retax1 = xv[v.at(r + 1, c + 1)] - xv[v.at(r + 1, c)], retay1 = yv[v.at(r + 1, c + 1)] - yv[v.at(r + 1, c)];
        const double rchix1 = xv[v.at(r + 1, c + 1)] - xv[v.at(r, c + 1)], rchiy1 = yv[v.at(r + 1, c + 1)] - yv[v.at(r, c + 1)];
        // same left-to-right, unfused evaluation order as the reference expression (bit-identical volumes)
        const double p1 = __dmul_rn(retax, rchiy), p2 = __dmul_rn(rchix, retay), p3 = __dmul_rn(retax1, rchiy1), p4 = __dmul_rn(rchix1, retay1);
        a_vol = __dmul_rn(0.5, __dsub_rn(__dadd_rn(__dsub_rn(p1, p2), p3), p4));
    }
    ncx[o] = a_ncx; ncy[o] = a_ncy; nex[o] = a_nex; ney[o] = a_ney; vol[o] = a_vol;
}

// ------------------------------------------------------------------------------------------------
// Boundary conditions (src/model/bc.cpp).  Ghost cells are stored in the padded state planes as
// CONSERVATIVE values; each BC computes its ghost primitives exactly as the reference does and converts.
// One launch per [[boundary]] table, in file order (BoundaryContainer::apply, bc.cpp:430-433).
// ------------------------------------------------------------------------------------------------
struct BcArgs {
    int type, face, lo, hi;     // padded index range [lo, hi] already clipped to this slab
    double u, v, T;
};

struct PrimSA { double r, u, v, p, T, nut; };

__device__ __forceinline__ PrimSA load_prim(const View& v, const Gas& g, const double* __restrict__ q, int r, int c) {
    const size_t o = v.at(r, c);
    PrimSA w;
    cons_to_prim(g, q[o], q[v.plane + o], q[2*v.plane + o], q[3*v.plane + o], w.r, w.u, w.v, w.p, w.T);
    w.nut = (v.nv > 4) ? q[4*v.plane + o]/w.r : 0.0;
    return w;
}
__device__ __forceinline__ void store_prim(const View& v, double* __restrict__ q, int r, int c, const PrimSA& w) {
    const size_t o = v.at(r, c);
    double q0, q1, q2, q3;
    prim_to_cons(w.r, w.u, w.v, w.p, q0, q1, q2, q3);
    q[o] = q0; q[v.plane + o] = q1; q[2*v.plane + o] = q2; q[3*v.plane + o] = q3;
    if (v.nv > 4) q[4*v.plane + o] = w.r*w.nut;
}
__device__ __forceinline__ void copy_cell(const View& v, double* __restrict__ q, int rd, int cd, int rs, int cs) {
    const size_t d = v.at(rd, cd), s = v.at(rs, cs);
    for (int k = 0; k < v.nv; k++) q[k*v.plane + d] = q[k*v.plane + s];
}

// local row of padded index jp, local column of padded index ip
__device__ __forceinline__ int row_of(const View& v, int jp) { return jp - 1 - v.j0 + JOFF; }
__device__ __forceinline__ int col_of(int ip) { return ip - 1 + IOFF; }

__global__ void bc_kernel(View v, Gas g, Metrics m, double* __restrict__ q, BcArgs b) {
    const int s = b.lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (s > b.hi) return;
    const bool horiz = (b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP);
    const bool bot = b.face == SGPU_FACE_BOTTOM, left = b.face == SGPU_FACE_LEFT;
    // ghost cell and the two interior cells normal to the boundary
    int rg, cg, ra, ca, rb, cb;
    if (horiz) {
        cg = ca = cb = col_of(s);
        rg = row_of(v, bot ? 0 : v.njc + 1); ra = row_of(v, bot ? 1 : v.njc); rb = row_of(v, bot ? 2 : v.njc - 1);
    } else {
        rg = ra = rb = row_of(v, s);
        cg = col_of(left ? 0 : v.nic + 1); ca = col_of(left ? 1 : v.nic); cb = col_of(left ? 2 : v.nic - 1);
    }
    PrimSA w;
    switch (b.type) {
    case SGPU_BC_FREESTREAM: {                                   // bc.cpp:26-60
        w.r = g.rho_inf; w.u = g.u_inf; w.v = g.v_inf; w.p = g.p_inf; w.T = 0; w.nut = 3.0*g.mu_ref/g.rho_inf;
        store_prim(v, q, rg, cg, w);
    } break;
    case SGPU_BC_SLIPWALL: {                                     // bc.cpp:78-114
        const PrimSA a = load_prim(v, g, q, ra, ca), c2 = load_prim(v, g, q, rb, cb);
        const int rf = row_of(v, bot ? 1 : v.njc + 1);          // eta face j = 0 or njc sits on the row of cell j
        const double nx = m.nex[v.at(rf, cg)], ny = m.ney[v.at(rf, cg)];
        const double ds = nx*nx + ny*ny;
        w.p = 1.5*a.p - 0.5*c2.p;
        w.r = 1.5*a.r - 0.5*c2.r;
        const double un = a.u*nx + a.v*ny;
        w.u = a.u - 2.0*un*nx/ds;
        w.v = a.v - 2.0*un*ny/ds;
        w.nut = 1.5*a.nut - 0.5*c2.nut;
        store_prim(v, q, rg, cg, w);
    } break;
    case SGPU_BC_WALL: {                                         // bc.cpp:150-204
        const PrimSA a = load_prim(v, g, q, ra, ca), c2 = load_prim(v, g, q, rb, cb);
        const double T = 1.5*a.T - 0.5*c2.T;
        w.r = 1.5*a.r - 0.5*c2.r;
        w.u = 2.0*b.u - (1.5*a.u - 0.5*c2.u);
        w.v = 2.0*b.v - (1.5*a.v - 0.5*c2.v);
        w.p = w.r*g.R*T;
        w.nut = -(1.5*a.nut - 0.5*c2.nut);
        store_prim(v, q, rg, cg, w);
    } break;
    case SGPU_BC_ISOTHERMALWALL: {                               // bc.cpp:388-413
        const PrimSA a = load_prim(v, g, q, ra, ca), c2 = load_prim(v, g, q, rb, cb);
        w.p = 1.5*a.p - 0.5*c2.p;
        w.u = 2.0*b.u - (1.5*a.u - 0.5*c2.u);
        w.v = 2.0*b.v - (1.5*a.v - 0.5*c2.v);
        w.r = w.p/b.T/g.R;
        w.nut = -(1.5*a.nut - 0.5*c2.nut);
        store_prim(v, q, rg, cg, w);
    } break;
    case SGPU_BC_WAKE: {                                         // bc.cpp:224-249 (source row is padded j = 1 for both faces)
        const int r1 = row_of(v, 1);
        const int cm = col_of(v.nic + 1 - s);
        copy_cell(v, q, rg, cg, r1, cm);
        copy_cell(v, q, rg, cm, r1, cg);
    } break;
    case SGPU_BC_OUTFLOW: {                                      // bc.cpp:295-309 (right face)
        w = load_prim(v, g, q, rg, cg - 1);
        w.p = g.p_inf;
        store_prim(v, q, rg, cg, w);
    } break;
    case SGPU_BC_PERIODIC: {                                     // bc.cpp:329-365
        if (horiz) {
            copy_cell(v, q, row_of(v, 0), cg, row_of(v, v.njc), cg);
            copy_cell(v, q, row_of(v, v.njc + 1), cg, row_of(v, 1), cg);
        } else {
            copy_cell(v, q, rg, col_of(0), rg, col_of(v.nic));
            copy_cell(v, q, rg, col_of(v.nic + 1), rg, col_of(1));
        }
    } break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// EulerEquation::calc_dt (src/model/eulerequation.cpp:237-259); one value per cell (the reference stores
// the same value nq times).  ds_eta/ds_chi = lengths of the cell's LOW faces (src/utils/mesh.cpp:202-203).
// ------------------------------------------------------------------------------------------------
__global__ void dt_kernel(View v, Metrics m, const double* __restrict__ q, double* __restrict__ dt, double cfl, double mu_inf) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x + IOFF;
    const int r = blockIdx.y + JOFF;
    if (c >= v.nic + IOFF) return;
    const size_t o = v.at(r, c);
    const double rho = q[o], u = q[v.plane + o]/rho, vv = q[2*v.plane + o]/rho, rhoE = q[3*v.plane + o];
    const double p = (rhoE - 0.5*rho*(u*u + vv*vv))*GM1;
    const double lambda = sqrt(GAMMA*p/rho) + fabs(u) + fabs(vv);
    const double ds_eta = sqrt(m.nex[o]*m.nex[o] + m.ney[o]*m.ney[o]);
    const double ds_chi = sqrt(m.ncx[o]*m.ncx[o] + m.ncy[o]*m.ncy[o]);
    const double len_min = fmin(ds_eta, ds_chi);
    dt[o] = cfl/(lambda/len_min + 2.0*mu_inf/len_min/len_min);
}

// update_rk4 / update_forward_euler (src/solver/solver.cpp:12-26): dst = q + rhs*dt*scale on owned cells
__global__ void axpy_dt_kernel(View v, double* __restrict__ dst, const double* __restrict__ q, const double* __restrict__ rhs,
                               const double* __restrict__ dt, double inv_div) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x + IOFF;
    const int r = blockIdx.y + JOFF;
    if (c >= v.nic + IOFF) return;
    const size_t o = v.at(r, c);
    const double d = dt[o];
    for (int k = 0; k < v.nv; k++) {
        const size_t ok = k*v.plane + o;
        dst[ok] = q[ok] + rhs[ok]*d*inv_div;      // inv_div = 1/(4-order) (exact for 1, 1/2, 1/4; 1/3 differs by <= 1 ulp)
    }
}
// ghost frame of a state (the two padded rows below / above the owned rows and the padded columns left / right of the cells) copied
// from src to dst, all nv planes: the fused Runge-Kutta stages write their result into the OTHER q_tmp buffer, whose ghost cells
// must look to the next boundary-condition pass exactly as q_tmp's own would have (a BC may read a ghost cell a later table writes)
__global__ void ghost_frame_copy_kernel(View v, double* __restrict__ dst, const double* __restrict__ src, int cols) {
    int r, c;
    if (!cols) {                                   // the 2 JOFF ghost rows, every column: grid (pitch/128, 2 JOFF)
        c = blockIdx.x*blockDim.x + threadIdx.x;
        r = (int)blockIdx.y < JOFF ? (int)blockIdx.y : v.njl + (int)blockIdx.y;
        if (c >= v.pitch) return;
    } else {                                       // the ghost / padding columns of the owned rows: grid (rows/128, pitch - nic)
        r = JOFF + blockIdx.x*blockDim.x + threadIdx.x;
        c = (int)blockIdx.y < IOFF ? (int)blockIdx.y : v.nic + (int)blockIdx.y;
        if (r >= JOFF + v.njl) return;
    }
    const size_t o = v.at(r, c);
    for (int k = 0; k < v.nv; k++) dst[k*v.plane + o] = src[k*v.plane + o];
}
// the reference divides: q + rhs*dt/(4.0-order).  Keep that form for bit-fidelity of the stage update.
__global__ void axpy_dt_div_kernel(View v, double* __restrict__ dst, const double* __restrict__ q, const double* __restrict__ rhs,
                                   const double* __restrict__ dt, double div) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x + IOFF;
    const int r = blockIdx.y + JOFF;
    if (c >= v.nic + IOFF) return;
    const size_t o = v.at(r, c);
    const double d = dt[o];
    for (int k = 0; k < v.nv; k++) {
        const size_t ok = k*v.plane + o;
        dst[ok] = q[ok] + rhs[ok]*d/div;
    }
}

// ------------------------------------------------------------------------------------------------
// Host boundary: AoS [i][jr][k] staging  <->  SoA planes [k][r][c].   jr counts rows from local row r0.
// 32x32 tiles through shared memory so that both sides are coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void aos_to_planes_kernel(View v, const double* __restrict__ stage, double* __restrict__ planes, int r0, int nrows,
                                     int ia = 0, int ni = -1 /* stage holds columns [ia, ia+ni) ; default: all */) {
    __shared__ double tile[32][33];
    if (ni < 0) ni = v.nic;
    const int M = nrows*v.nv;                      // contiguous length per i in the staging buffer
    const int m0 = blockIdx.x*32, i0 = blockIdx.y*32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int i = i0 + dy, mm = m0 + threadIdx.x;
        if (i < ni && mm < M) tile[dy][threadIdx.x] = stage[(size_t)i*M + mm];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int mm = m0 + dy, i = i0 + threadIdx.x;
        if (i < ni && mm < M) {
            const int jr = mm/v.nv, k = mm - jr*v.nv;
            planes[k*v.plane + v.at(r0 + jr, ia + i + IOFF)] = tile[threadIdx.x][dy];
        }
    }
}
__global__ void planes_to_aos_kernel(View v, double* __restrict__ stage, const double* __restrict__ planes, int r0, int nrows,
                                     int nvp /* planes available: nv, or 1 to broadcast a per-cell plane */, int ia = 0, int ni = -1) {
    __shared__ double tile[32][33];
    if (ni < 0) ni = v.nic;
    const int M = nrows*v.nv;
    const int m0 = blockIdx.x*32, i0 = blockIdx.y*32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int mm = m0 + dy, i = i0 + threadIdx.x;
        if (i < ni && mm < M) {
            const int jr = mm/v.nv, k = mm - jr*v.nv;
            tile[threadIdx.x][dy] = planes[(nvp == 1 ? 0 : k)*v.plane + v.at(r0 + jr, ia + i + IOFF)];
        }
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int i = i0 + dy, mm = m0 + threadIdx.x;
        if (i < ni && mm < M) stage[(size_t)i*M + mm] = tile[dy][threadIdx.x];
    }
}
// per-cell field [i][jr] -> one plane
__global__ void field_to_plane_kernel(View v, const double* __restrict__ stage, double* __restrict__ plane, int r0, int nrows) {
    __shared__ double tile[32][33];
    const int m0 = blockIdx.x*32, i0 = blockIdx.y*32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int i = i0 + dy, mm = m0 + threadIdx.x;
        if (i < v.nic && mm < nrows) tile[dy][threadIdx.x] = stage[(size_t)i*nrows + mm];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int mm = m0 + dy, i = i0 + threadIdx.x;
        if (i < v.nic && mm < nrows) plane[v.at(r0 + mm, i + IOFF)] = tile[threadIdx.x][dy];
    }
}
// vertex arrays [iv][jr] -> vertex plane (ni columns)
__global__ void vertex_to_plane_kernel(View v, const double* __restrict__ stage, double* __restrict__ plane, int r0, int nrows) {
    __shared__ double tile[32][33];
    const int m0 = blockIdx.x*32, i0 = blockIdx.y*32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int i = i0 + dy, mm = m0 + threadIdx.x;
        if (i < v.ni && mm < nrows) tile[dy][threadIdx.x] = stage[(size_t)i*nrows + mm];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int mm = m0 + dy, i = i0 + threadIdx.x;
        if (i < v.ni && mm < nrows) plane[v.at(r0 + mm, i + IOFF)] = tile[threadIdx.x][dy];
    }
}

// ------------------------------------------------------------------------------------------------
// j-slab halos: two boundary cell rows of q <-> contiguous buffer [nv][2][nic]
// ------------------------------------------------------------------------------------------------
__global__ void halo_pack_kernel(View v, const double* __restrict__ q, double* __restrict__ buf, int r_first) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= v.nic) return;
    const int k = blockIdx.y >> 1, rr = blockIdx.y & 1;
    buf[((size_t)k*2 + rr)*v.nic + i] = q[k*v.plane + v.at(r_first + rr, i + IOFF)];
}
__global__ void halo_unpack_kernel(View v, double* __restrict__ q, const double* __restrict__ buf, int r_first) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= v.nic) return;
    const int k = blockIdx.y >> 1, rr = blockIdx.y & 1;
    q[k*v.plane + v.at(r_first + rr, i + IOFF)] = buf[((size_t)k*2 + rr)*v.nic + i];
}

// sequence flags of the peer-memory halo exchange (system scope: the flag lives in another GPU's memory)
__global__ void halo_signal_kernel(unsigned long long* flag, unsigned long long seq) {
    __threadfence_system();
    *(volatile unsigned long long*)flag = seq;
    __threadfence_system();
}
// bounded spin: a neighbour that never pushes (mismatched call counts, a dead rank) must not hang the GPU -- after
// `timeout_ns` the kernel gives up and raises *err, which the host reads at its next synchronising call
__global__ void halo_wait_kernel(const unsigned long long* flag, unsigned long long seq, unsigned long long timeout_ns, int* err) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*(volatile const unsigned long long*)flag < seq) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { atomicAdd(err, 1); break; }
    }
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// Final stage of the residual-norm reduction (src/solver/solver.cpp:125-134): partial[item][nv] -> out[nv]
// ------------------------------------------------------------------------------------------------
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nitems, int nv, double* __restrict__ out) {
    const int k = blockIdx.x;
    double s = 0.0;
    for (int n = threadIdx.x; n < nitems; n += blockDim.x) s += partial[(size_t)n*nv + k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    __shared__ double ws[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) ws[w] = s;
    __syncthreads();
    if (w == 0) {
        s = (lane < (blockDim.x >> 5)) ? ws[lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) out[k] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Wall data of IOManager::write_surface (src/utils/io.cpp:182-255): the eta-face gradients of u and v on the
// j = 0 faces as the last residual evaluation left them in EulerEquation's work arrays (Mesh::calc_gradient
// with its j == 0 variant, src/utils/mesh.cpp:88-107,127-128) and the pressure of the two lowest cell rows of
// the final state (Solution::p, unpadded, filled by IOManager::write, io.cpp:41).  One thread per cell column.
//   q_res: state whose ghosts/gradients are used (BCs already applied); q_fin: the final state
// out = gu[nic][2] | gv[nic][2] | p0[nic] | p1[nic] | xc0[nic] | dx[nic] | dy[nic]
// ------------------------------------------------------------------------------------------------
__global__ void wall_data_kernel(View v, Metrics m, const double* __restrict__ q_res, const double* __restrict__ q_fin,
                                 const double* __restrict__ xv, const double* __restrict__ yv, double* __restrict__ out) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= v.nic) return;
    const int nic = v.nic;
    const size_t pl = v.plane;
    const int r0 = JOFF, rg = JOFF - 1;                 // cell row j = 0 and the bottom ghost row
    auto prim = [&](const double* q, int r, int c, double& u, double& vv, double& p) {      // fluid.cpp:50-67
        const size_t o = v.at(r, c);
        const double rho = q[o];
        u = q[pl + o]/rho; vv = q[2*pl + o]/rho;
        p = (q[3*pl + o] - 0.5*rho*(u*u + vv*vv))*(GAMMA - 1.0);
    };
    double* gu = out; double* gv = gu + 2*(size_t)nic; double* p0 = gv + 2*(size_t)nic; double* p1 = p0 + nic;
    double* xc0 = p1 + nic; double* dx = xc0 + nic; double* dy = dx + nic;
    const int c = i + IOFF;
    double ut, vt, ub, vb, ua, va, ud, vd, pp;
    prim(q_fin, r0, c, ua, va, pp); p0[i] = pp;
    prim(q_fin, r0 + 1, c, ua, va, pp); p1[i] = pp;
    prim(q_res, r0, c, ut, vt, pp); prim(q_res, rg, c, ub, vb, pp);
    prim(q_res, r0, c - 1, ua, va, pp); prim(q_res, rg, c - 1, ud, vd, pp);
    const double ul = 0.25*(ut + ub + ua + ud), vl = 0.25*(vt + vb + va + vd);           // mesh.cpp:96
    prim(q_res, r0, c + 1, ua, va, pp); prim(q_res, rg, c + 1, ud, vd, pp);
    const double ur = 0.25*(ut + ub + ua + ud), vr = 0.25*(vt + vb + va + vd);           // mesh.cpp:97
    const size_t o = v.at(r0, c), o1 = v.at(r0 + 1, c), oc1 = v.at(r0, c + 1);
    const double vol = m.vol[o];
    const double tx = 0.5*(m.nex[o] + m.nex[o1]), ty = 0.5*(m.ney[o] + m.ney[o1]);       // mesh.cpp:99-107
    const double bx = m.nex[o], by = m.ney[o], lx = m.ncx[o], ly = m.ncy[o], rx = m.ncx[oc1], ry = m.ncy[oc1];
    gu[2*i] = (tx*ut - bx*ub + rx*ur - lx*ul)/vol; gu[2*i + 1] = (ty*ut - by*ub + ry*ur - ly*ul)/vol;   // :127-128
    gv[2*i] = (tx*vt - bx*vb + rx*vr - lx*vl)/vol; gv[2*i + 1] = (ty*vt - by*vb + ry*vr - ly*vl)/vol;
    xc0[i] = 0.25*(xv[o] + xv[oc1] + xv[o1] + xv[v.at(r0 + 1, c + 1)]);                  // mesh.cpp:199
    dx[i] = xv[oc1] - xv[o]; dy[i] = yv[oc1] - yv[o];
}

// ------------------------------------------------------------------------------------------------
// Wall distance of the SA extension (no reference counterpart; SURVEY.md Appendix C "wall distance d"): for every cell
// the distance from its centre (xc, yc = 1/4 of its four vertices, src/utils/mesh.cpp:199-200) to the nearest point of
// the wall -- the union of the boundary edges covered by a `wall` / `isothermalwall` table.  Brute force, exact
// point-to-segment distance: segments are staged through shared memory in tiles, one thread per cell.
// seg: [nseg][5] = ax, ay, abx, aby, 1/|ab|^2
// ------------------------------------------------------------------------------------------------
constexpr int WD_TILE = 256;
__global__ void __launch_bounds__(128) wall_distance_kernel(View v, const double* __restrict__ xv, const double* __restrict__ yv,
                                                            const double* __restrict__ seg, int nseg, double* __restrict__ wdist) {
    __shared__ double s_seg[WD_TILE*5];
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int r = blockIdx.y + JOFF;
    const bool ok = i < v.nic;
    const int c = (ok ? i : 0) + IOFF;
    const size_t o = v.at(r, c), o1 = v.at(r + 1, c);
    const double px = 0.25*(xv[o] + xv[o + 1] + xv[o1] + xv[o1 + 1]), py = 0.25*(yv[o] + yv[o + 1] + yv[o1] + yv[o1 + 1]);
    double best = 1e300;
    for (int s0 = 0; s0 < nseg; s0 += WD_TILE) {
        const int n = min(WD_TILE, nseg - s0);
        __syncthreads();
        for (int k = threadIdx.x; k < n*5; k += blockDim.x) s_seg[k] = seg[(size_t)s0*5 + k];
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const double ax = s_seg[5*k], ay = s_seg[5*k + 1], bx = s_seg[5*k + 2], by = s_seg[5*k + 3], il2 = s_seg[5*k + 4];
            const double dx = px - ax, dy = py - ay;
            double t = (dx*bx + dy*by)*il2;
            t = fmin(fmax(t, 0.0), 1.0);
            const double ex = dx - t*bx, ey = dy - t*by;
            best = fmin(best, ex*ex + ey*ey);
        }
    }
    if (ok) wdist[o] = sqrt(best);
}

__global__ void fill_kernel(double* __restrict__ p, size_t n, double val) {
    size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x*blockDim.x;
    for (; i < n; i += stride) p[i] = val;
}

} // namespace sg
