// Does the LAYOUT of the Jacobian blocks bound the matvec?  Reads NE = 325 doubles per cell for 4096 x 4096 cells (43.6 GB, the
// block-stencil Jacobian of the bench workload) and reduces them, (a) plane-major as stored today: entry e of cell o at e*plane + o
// -- 325 concurrent DRAM streams -- and (b) tile-major: [tile of 32 cells][entry][32] -- one sequential 83 KB run per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/stream_layout_bench tools/stream_layout_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NE = 325;
__device__ __forceinline__ double ld_cs(const double* p) { double v; asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
template <bool TILE, int SKIP>
__global__ void __launch_bounds__(128) rd(const double* __restrict__ J, size_t plane, int pitch, int nic, double* __restrict__ y) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nic) return;
    const size_t o = (size_t)j*pitch + i;
    double s[5] = {0, 0, 0, 0, 0};
#pragma unroll 1
    for (int sl = 0; sl < 13; sl++) {
        double v[25];
#pragma unroll
        for (int e = 0; e < 25; e++) {
            if (SKIP && (e % 7 == 3)) { v[e] = 0.0; continue; }          // skip ~14 % of the entries
            const size_t idx = TILE ? ((o >> 5)*NE + sl*25 + e)*32 + (o & 31) : (size_t)(sl*25 + e)*plane + o;
            v[e] = ld_cs(J + idx);
        }
#pragma unroll
        for (int e = 0; e < 25; e++) s[e % 5] += v[e];
    }
    y[o] = s[0] + s[1] + s[2] + s[3] + s[4];
}
int main() {
    const int nic = 4096, njc = 4096, pitch = 4096;
    const size_t plane = (size_t)pitch*njc, n = plane*NE;
    double *J, *y; cudaMalloc(&J, n*8); cudaMalloc(&y, plane*8); cudaMemset(J, 0, n*8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char* name, auto kern, double frac) {
        dim3 g(nic/128, njc);
        kern<<<g, 128>>>(J, plane, pitch, nic, y); cudaDeviceSynchronize();
        cudaEventRecord(a); for (int r = 0; r < 3; r++) kern<<<g, 128>>>(J, plane, pitch, nic, y); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
        printf("%-34s %.3f ms  %.0f GB/s\n", name, ms, frac*n*8/ms/1e6);
    };
    run("plane-major, all entries", rd<false, 0>, 1.0);
    run("tile-major,  all entries", rd<true, 0>, 1.0);
    run("plane-major, 14 % skipped", rd<false, 1>, 21.0/25.0);
    run("tile-major,  14 % skipped", rd<true, 1>, 21.0/25.0);
    return 0;
}
