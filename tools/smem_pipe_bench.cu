// Micro-benchmark of the sm_100a shared-memory pipe: cycles per conflict-free warp LDS of 4 / 8 / 16 bytes per lane, alone and
// mixed with DFMAs (the residual kernel issues 127 LDS.64 + 27 STS.64 + 12 LDGSTS per 733 fp64 instructions and thread-row).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/smem_pipe_bench tools/smem_pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int BYTES, int NFMA>
__global__ void __launch_bounds__(128) k(double* out, int iters, long long* cyc) {
    __shared__ __align__(16) double sm[128*2*8];
    for (int i = threadIdx.x; i < 128*2*8; i += 128) sm[i] = 1e-9*i;
    __syncthreads();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double f[8]; for (int i = 0; i < 8; i++) f[i] = 1.0 + i*1e-9 + threadIdx.x*1e-12;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x*BYTES;
    unsigned off = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const unsigned a = base + ((off + u*128*BYTES) & (128*2*8*8 - 1 - (128*BYTES - 1)));
            if (BYTES == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc[u] += (double)v; }
            if (BYTES == 8) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); acc[u] += v; }
            if (BYTES == 16) { double v, w; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(w) : "r"(a)); acc[u] += v; acc[(u + 1) & 7] += w; }
#pragma unroll
            for (int m = 0; m < NFMA; m++) f[(u + m) & 7] = fma(f[(u + m) & 7], 1.0000001, 1e-9);
        }
        off += 128*BYTES;
    }
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += acc[i] + f[i];
    out[blockIdx.x*128 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int BYTES, int NFMA>
void run(const char* name, int sms, double* out, long long* cyc) {
    const int iters = 2048;
    for (int per = 1; per <= 6; per += (per < 3 ? 1 : 3)) {
        const int grid = sms*per;
        for (int rep = 0; rep < 2; rep++) { k<BYTES, NFMA><<<grid, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); }
        static long long h[148*8]; cudaMemcpy(h, cyc, sizeof(long long)*grid, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        const double lds_sm = 4.0*per*iters*8;          // warp LDS instructions per SM
        const double fp64 = per*(double)iters*8*(NFMA + (BYTES == 16 ? 2 : 1));
        printf("%-34s warps/scheduler %d: %.2f clk per warp-LDS per SM (%.0f B/clk/SM), fp64 pipe %5.1f %%\n", name, per, mx/lds_sm,
               32.0*BYTES*lds_sm/mx, 100.0*fp64/mx/0.5);
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double)*sms*8*128); cudaMalloc(&cyc, sizeof(long long)*sms*8);
    run<4, 0>("LDS.32 + 1 DADD", sms, out, cyc);
    run<8, 0>("LDS.64 + 1 DADD", sms, out, cyc);
    run<16, 0>("LDS.128 + 2 DADD", sms, out, cyc);
    run<8, 3>("LDS.64 + 1 DADD + 3 DFMA", sms, out, cyc);
    run<8, 5>("LDS.64 + 1 DADD + 5 DFMA", sms, out, cyc);
    run<8, 7>("LDS.64 + 1 DADD + 7 DFMA", sms, out, cyc);
    run<16, 10>("LDS.128 + 2 DADD + 10 DFMA", sms, out, cyc);
    run<16, 14>("LDS.128 + 2 DADD + 14 DFMA", sms, out, cyc);
    return 0;
}
