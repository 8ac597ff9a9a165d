// HBM read-only bandwidth on B200 (the matvec of the device GMRES is a pure read stream; MEASURED_PEAKS.json's figure is a COPY).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/hbm_read_bench tools/hbm_read_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC, int UNROLL>
__global__ void rd(const double* __restrict__ p, size_t n, double* out) {
    double s = 0.0;
    const size_t stride = (size_t)gridDim.x*blockDim.x*VEC;
    size_t i = ((size_t)blockIdx.x*blockDim.x + threadIdx.x)*VEC;
    for (; i + (UNROLL - 1)*stride < n; i += UNROLL*stride) {
        double v[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (VEC == 1) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v[u][0]) : "l"(p + i + u*stride));
            else asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v[u][0]), "=d"(v[u][VEC - 1]) : "l"(p + i + u*stride));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < VEC; k++) s += v[u][k];
    }
    if (s == 1.2345e-300) out[0] = s;
}
__global__ void cp(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) b[i] = a[i];
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < 5; r++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms/5;
}
int main() {
    const size_t n = (size_t)1 << 30;     // 8 GiB of doubles
    double *p, *q, *out; cudaMalloc(&p, n*8); cudaMalloc(&q, n*8); cudaMalloc(&out, 8);
    cudaMemset(p, 0, n*8); cudaMemset(q, 0, n*8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mult : {8, 16, 32}) {
        float t1 = timeit([&] { rd<1, 8><<<sms*mult, 256>>>(p, n, out); });
        float t2 = timeit([&] { rd<2, 4><<<sms*mult, 256>>>(p, n, out); });
        float t3 = timeit([&] { rd<2, 8><<<sms*mult, 256>>>(p, n, out); });
        printf("grid %2d x SMs: read 8B x8 %.0f GB/s, 16B x4 %.0f GB/s, 16B x8 %.0f GB/s\n", mult, n*8/t1/1e6, n*8/t2/1e6, n*8/t3/1e6);
    }
    float tc = timeit([&] { cp<<<sms*32, 256>>>((const double2*)p, (double2*)q, n/2); });
    printf("copy (read + write counted): %.0f GB/s\n", 2.0*n*8/tc/1e6);
    return 0;
}
