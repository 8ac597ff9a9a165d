// Small forward-mode dual numbers for device code: value + N tangent lanes in registers.
// Used by the Jacobian kernel to differentiate the SAME device functions the residual kernel runs
// (physics.cuh is templated on the scalar type), which is what replaces ADOL-C's tape
// (reference call sites: src/solver/solver.cpp:72-90,156).
#pragma once
#include <cuda_runtime.h>

namespace sg {

template <int N>
struct Dual {
    double v;
    double d[N];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double x) : v(x) {
#pragma unroll
        for (int k = 0; k < N; k++) d[k] = 0.0;
    }
};

#define SG_DUAL_LOOP _Pragma("unroll") for (int k = 0; k < N; k++)

template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; SG_DUAL_LOOP r.d[k] = a.d[k] + b.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; SG_DUAL_LOOP r.d[k] = a.d[k] - b.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v*b.v; SG_DUAL_LOOP r.d[k] = a.d[k]*b.v + a.v*b.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r; r.v = b - a.v; SG_DUAL_LOOP r.d[k] = -a.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v*b; SG_DUAL_LOOP r.d[k] = a.d[k]*b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(double b, const Dual<N>& a) { Dual<N> r; r.v = a.v*b; SG_DUAL_LOOP r.d[k] = a.d[k]*b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; SG_DUAL_LOOP r.d[k] = -a.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, double b) { return a*(1.0/b); }
__device__ __forceinline__ double rcp_fast(double x);
__device__ __forceinline__ double rsqrt_fast(double x);
__device__ __forceinline__ double s_pow23(double x);
__device__ __forceinline__ double s_pow16(double x);
template <int N> __device__ __forceinline__ Dual<N> s_rcp(const Dual<N>& a) { Dual<N> r; r.v = rcp_fast(a.v); const double s = -r.v*r.v; SG_DUAL_LOOP r.d[k] = s*a.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> s_rsqrt(const Dual<N>& a) { Dual<N> r; r.v = rsqrt_fast(a.v); const double s = -0.5*r.v*r.v*r.v; SG_DUAL_LOOP r.d[k] = s*a.d[k]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { return a*s_rcp(b); }
template <int N> __device__ __forceinline__ Dual<N> s_sqrt(const Dual<N>& a) { Dual<N> r; const double y = rsqrt_fast(a.v); r.v = a.v*y; const double s = 0.5*y; SG_DUAL_LOOP r.d[k] = s*a.d[k]; return r; }
// derivative of the taken branch (a.v < 0 ? -a : a), the same convention as oracle/adtypes.hpp
template <int N> __device__ __forceinline__ Dual<N> s_abs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
template <int N> __device__ __forceinline__ double s_val(const Dual<N>& a) { return a.v; }
template <int N> __device__ __forceinline__ Dual<N> s_pow23(const Dual<N>& a) {
    Dual<N> r; r.v = s_pow23(a.v); const double s = (2.0/3.0)*r.v*rcp_fast(a.v); SG_DUAL_LOOP r.d[k] = s*a.d[k]; return r;
}
template <int N> __device__ __forceinline__ Dual<N> s_pow16(const Dual<N>& a) {
    Dual<N> r; r.v = s_pow16(a.v); const double s = (1.0/6.0)*r.v*rcp_fast(a.v); SG_DUAL_LOOP r.d[k] = s*a.d[k]; return r;
}
#undef SG_DUAL_LOOP

} // namespace sg
