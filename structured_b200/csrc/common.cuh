// Shared device-side view of a slab and small utilities.
//
// HBM layout (DESIGN.md "data layout"): every per-cell / per-face / per-vertex quantity is a PLANE of
// `rows x pitch` doubles, i fastest (coalesced across a warp), j-slabs are contiguous row ranges.
//   column c = i + IOFF            (IOFF = 2: one BC ghost column each side + one never-read pad column)
//   row    r = j - j_begin + JOFF  (JOFF = 2: two ghost rows each side -- the BC ghost row at a physical
//                                   boundary, the neighbour slab's two cell rows at an interior slab edge)
// State arrays are nv consecutive planes (SoA): q[k][r][c].
#pragma once
#include <cuda_runtime.h>
#include "physics.cuh"

namespace sg {

constexpr int IOFF = 2;
constexpr int JOFF = 2;

struct View {
    int nic, njc;          // GLOBAL cell counts
    int ni, nj;            // GLOBAL vertex counts
    int j0, j1, njl;       // owned global cell rows [j0, j1), njl = j1 - j0
    int nv;
    int pitch, rows;       // plane geometry: rows = njl + 2*JOFF (+1 for vertex planes, allocated with rows+1)
    size_t plane;          // rows*pitch (cell planes); vertex planes use (rows+1)*pitch
    __host__ __device__ size_t at(int r, int c) const { return (size_t)r*pitch + c; }
};

struct Metrics {           // Mesh::calc_metrics outputs (src/utils/mesh.cpp:172-205) as planes
    const double* ncx; const double* ncy;      // normal_chi[i][j][0..1]  at (r(j), c(i))
    const double* nex; const double* ney;      // normal_eta[i][j][0..1]
    const double* vol;                         // volume[i][j]
};

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

} // namespace sg
