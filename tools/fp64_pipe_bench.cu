// Micro-benchmark of the sm_100a fp64 pipe: what fraction of the nominal 64 DFMA lanes / clk / SM a warp mix can sustain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/fp64_pipe_bench tools/fp64_pipe_bench.cu
// Each CTA = 128 threads (one warp per SM sub-partition); `cta_per_sm` CTAs resident per SM -> that many warps per scheduler.
// Variants: ILP independent chains per thread; operand pattern (accumulate-only vs three distinct sources); an ALU / LDS
// instruction mixed in after every `mix` fp64 instructions (the residual kernel issues 0.65 non-fp64 per fp64 instruction).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP, int PATTERN, int MIX>
__global__ void __launch_bounds__(128) k(double* out, int iters, long long* cyc) {
    __shared__ double sm[128*4];
    double a[ILP], x[ILP], y[ILP];
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x*1e-3 + i; x[i] = 1.0 + 1e-9*(i + threadIdx.x); y[i] = 1e-7*(i + 1); }
    sm[threadIdx.x] = 1.0; sm[threadIdx.x + 128] = 2.0; sm[threadIdx.x + 256] = 3.0; sm[threadIdx.x + 384] = 4.0;
    int z = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (PATTERN == 0) a[i] = fma(a[i], 1.0000001, 1e-9);                 // one register source + immediates
            else if (PATTERN == 1) a[i] = fma(x[i], y[i], a[i]);                   // three distinct register sources
            else if (PATTERN == 2) {                                               // DFMA / DMUL / DADD round robin, distinct sources
                if (i % 3 == 0) a[i] = fma(x[i], y[i], a[i]);
                else if (i % 3 == 1) a[i] = a[i]*x[i];
                else a[i] = a[i] + y[i];
            }
            if (MIX == 1) { z = z*3 + i; }                                         // one IMAD per fp64
            if (MIX == 2 && (i & 1)) { z = z*3 + i; }                              // one IMAD per two fp64
            if (MIX == 3 && (i & 3) == 0) { a[i] += sm[(z + i*32) & 511]; z += 1; }   // an LDS-fed DADD per four fp64
        }
    }
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x*128 + threadIdx.x] = s + z;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int PATTERN, int MIX>
void run(const char* name, int sms, double* out, long long* cyc) {
    const int iters = 4096;
    for (int per = 1; per <= 6; per++) {
        const int grid = sms*per;
        k<ILP, PATTERN, MIX><<<grid, 128>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        k<ILP, PATTERN, MIX><<<grid, 128>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long h[148*8]; cudaMemcpy(h, cyc, sizeof(long long)*grid, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        double fp64_per_warp = (double)iters*ILP + ((MIX == 3) ? iters*(ILP/4) : 0);
        // per scheduler: `per` warps, each fp64_per_warp instructions; nominal peak = 1 warp instruction per 2 cycles
        const double rate = per*fp64_per_warp/mx;
        printf("%-46s warps/scheduler %d: %.3f fp64 warp-instr/clk/scheduler = %5.1f %% of nominal (16 lanes)\n", name, per, rate, 100.0*rate/0.5);
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double)*sms*8*128); cudaMalloc(&cyc, sizeof(long long)*sms*8);
    run<8, 0, 0>("DFMA acc-only, ILP 8", sms, out, cyc);
    run<2, 1, 0>("DFMA 3 reg sources, ILP 2", sms, out, cyc);
    run<4, 1, 0>("DFMA 3 reg sources, ILP 4", sms, out, cyc);
    run<8, 1, 0>("DFMA 3 reg sources, ILP 8", sms, out, cyc);
    run<9, 2, 0>("DFMA/DMUL/DADD, ILP 9", sms, out, cyc);
    run<3, 2, 0>("DFMA/DMUL/DADD, ILP 3", sms, out, cyc);
    run<8, 1, 1>("DFMA 3 src ILP 8 + 1 IMAD per fp64", sms, out, cyc);
    run<8, 1, 2>("DFMA 3 src ILP 8 + 1 IMAD per 2 fp64", sms, out, cyc);
    run<4, 1, 2>("DFMA 3 src ILP 4 + 1 IMAD per 2 fp64", sms, out, cyc);
    run<8, 1, 3>("DFMA 3 src ILP 8 + LDS-fed DADD per 4 fp64", sms, out, cyc);
    return 0;
}
