// TEST INFRASTRUCTURE ONLY -- not part of the shipped GPU path.
//
// Scalar types used to differentiate CPU residual code (the reference's own templates in
// oracle/ref/, and the restatement in oracle/port/) in place of ADOL-C's `adouble`
// (reference: src/common.h:55-58, src/solver/solver.cpp:72-90,156; ADOL-C itself is an
// un-vendored, un-pinned third-party dependency that is not installable here):
//
//   Dual<N>  forward-mode dual number carrying N tangent lanes  -> Jacobian columns
//   DepSet   index-domain propagation (what ADOL-C's sparse_jac does for options[0]=0)
//            -> the STRUCTURAL sparsity pattern, including entries that are numerically zero
//
// Nothing under structured_b200/ may include this file.
#ifndef ORACLE_ADTYPES_HPP
#define ORACLE_ADTYPES_HPP
#include <cmath>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <ostream>

namespace oad {

// ---------------------------------------------------------------- Dual<N>
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0.0) { for (int k = 0; k < N; k++) d[k] = 0.0; }
    Dual(double x) : v(x) { for (int k = 0; k < N; k++) d[k] = 0.0; }
    Dual(int x) : v(x) { for (int k = 0; k < N; k++) d[k] = 0.0; }
    Dual& operator+=(const Dual& o) { v += o.v; for (int k = 0; k < N; k++) d[k] += o.d[k]; return *this; }
    Dual& operator-=(const Dual& o) { v -= o.v; for (int k = 0; k < N; k++) d[k] -= o.d[k]; return *this; }
    Dual& operator*=(const Dual& o) { for (int k = 0; k < N; k++) d[k] = d[k]*o.v + v*o.d[k]; v *= o.v; return *this; }
    Dual& operator/=(const Dual& o) {
        const double q = v/o.v;
        for (int k = 0; k < N; k++) d[k] = (d[k] - q*o.d[k])/o.v;
        v = q; return *this;
    }
};
template <int N> inline Dual<N> operator+(Dual<N> a, const Dual<N>& b) { a += b; return a; }
template <int N> inline Dual<N> operator-(Dual<N> a, const Dual<N>& b) { a -= b; return a; }
template <int N> inline Dual<N> operator*(Dual<N> a, const Dual<N>& b) { a *= b; return a; }
template <int N> inline Dual<N> operator/(Dual<N> a, const Dual<N>& b) { a /= b; return a; }
template <int N> inline Dual<N> operator+(Dual<N> a, double b) { a.v += b; return a; }
template <int N> inline Dual<N> operator+(double b, Dual<N> a) { a.v += b; return a; }
template <int N> inline Dual<N> operator-(Dual<N> a, double b) { a.v -= b; return a; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r; r.v = b - a.v; for (int k = 0; k < N; k++) r.d[k] = -a.d[k]; return r; }
template <int N> inline Dual<N> operator*(Dual<N> a, double b) { a.v *= b; for (int k = 0; k < N; k++) a.d[k] *= b; return a; }
template <int N> inline Dual<N> operator*(double b, Dual<N> a) { a.v *= b; for (int k = 0; k < N; k++) a.d[k] *= b; return a; }
template <int N> inline Dual<N> operator/(Dual<N> a, double b) { a.v /= b; for (int k = 0; k < N; k++) a.d[k] /= b; return a; }
template <int N> inline Dual<N> operator/(double b, const Dual<N>& a) { Dual<N> r(b); r /= a; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int k = 0; k < N; k++) r.d[k] = -a.d[k]; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a) { return a; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) {
    Dual<N> r; r.v = std::sqrt(a.v); const double s = 0.5/r.v;
    for (int k = 0; k < N; k++) r.d[k] = s*a.d[k]; return r;
}
// derivative of the taken branch, as ADOL-C does away from the kink
template <int N> inline Dual<N> fabs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
template <int N> inline Dual<N> abs(const Dual<N>& a) { return fabs(a); }
template <int N> inline Dual<N> pow(const Dual<N>& a, double e) {
    Dual<N> r; r.v = std::pow(a.v, e); const double s = e*std::pow(a.v, e - 1.0);
    for (int k = 0; k < N; k++) r.d[k] = s*a.d[k]; return r;
}
template <int N> inline Dual<N> pow(const Dual<N>& a, int e) { return pow(a, (double)e); }
template <int N> inline Dual<N> exp(const Dual<N>& a) {
    Dual<N> r; r.v = std::exp(a.v); for (int k = 0; k < N; k++) r.d[k] = r.v*a.d[k]; return r;
}
template <int N> inline Dual<N> log(const Dual<N>& a) {
    Dual<N> r; r.v = std::log(a.v); for (int k = 0; k < N; k++) r.d[k] = a.d[k]/a.v; return r;
}
#define OAD_CMP(op) \
template <int N> inline bool operator op(const Dual<N>& a, const Dual<N>& b) { return a.v op b.v; } \
template <int N> inline bool operator op(const Dual<N>& a, double b) { return a.v op b; } \
template <int N> inline bool operator op(double a, const Dual<N>& b) { return a op b.v; }
OAD_CMP(<) OAD_CMP(>) OAD_CMP(<=) OAD_CMP(>=) OAD_CMP(==) OAD_CMP(!=)
#undef OAD_CMP
template <int N> inline Dual<N> min(const Dual<N>& a, const Dual<N>& b) { return b.v < a.v ? b : a; }
template <int N> inline Dual<N> max(const Dual<N>& a, const Dual<N>& b) { return a.v < b.v ? b : a; }
template <int N> inline std::ostream& operator<<(std::ostream& os, const Dual<N>& a) { return os << a.v; }

// ---------------------------------------------------------------- DepSet
// Value + sorted set of independent-variable indices the value structurally depends on.
// Every arithmetic operation unions the index domains of its operands; constants have an
// empty domain.  `0.0 * x` still depends on x, `flux[0] = 0.0` (src/model/flux.cpp:42)
// depends on nothing: exactly ADOL-C's index-domain rules for its sparsity detection.
struct DepSet {
    double v;
    std::vector<uint32_t> s;
    DepSet() : v(0.0) {}
    DepSet(double x) : v(x) {}
    DepSet(int x) : v(x) {}
    void absorb(const DepSet& o) {
        if (o.s.empty()) return;
        if (s.empty()) { s = o.s; return; }
        std::vector<uint32_t> r; r.reserve(s.size() + o.s.size());
        std::set_union(s.begin(), s.end(), o.s.begin(), o.s.end(), std::back_inserter(r));
        s.swap(r);
    }
    DepSet& operator+=(const DepSet& o) { v += o.v; absorb(o); return *this; }
    DepSet& operator-=(const DepSet& o) { v -= o.v; absorb(o); return *this; }
    DepSet& operator*=(const DepSet& o) { v *= o.v; absorb(o); return *this; }
    DepSet& operator/=(const DepSet& o) { v /= o.v; absorb(o); return *this; }
};
inline DepSet operator+(DepSet a, const DepSet& b) { a += b; return a; }
inline DepSet operator-(DepSet a, const DepSet& b) { a -= b; return a; }
inline DepSet operator*(DepSet a, const DepSet& b) { a *= b; return a; }
inline DepSet operator/(DepSet a, const DepSet& b) { a /= b; return a; }
inline DepSet operator+(DepSet a, double b) { a.v += b; return a; }
inline DepSet operator+(double b, DepSet a) { a.v += b; return a; }
inline DepSet operator-(DepSet a, double b) { a.v -= b; return a; }
inline DepSet operator-(double b, DepSet a) { a.v = b - a.v; return a; }
inline DepSet operator*(DepSet a, double b) { a.v *= b; return a; }
inline DepSet operator*(double b, DepSet a) { a.v *= b; return a; }
inline DepSet operator/(DepSet a, double b) { a.v /= b; return a; }
inline DepSet operator/(double b, DepSet a) { a.v = b/a.v; return a; }
inline DepSet operator-(DepSet a) { a.v = -a.v; return a; }
inline DepSet operator+(const DepSet& a) { return a; }
inline DepSet sqrt(DepSet a) { a.v = std::sqrt(a.v); return a; }
inline DepSet fabs(DepSet a) { a.v = std::fabs(a.v); return a; }
inline DepSet abs(DepSet a) { a.v = std::fabs(a.v); return a; }
inline DepSet pow(DepSet a, double e) { a.v = std::pow(a.v, e); return a; }
inline DepSet pow(DepSet a, int e) { a.v = std::pow(a.v, (double)e); return a; }
inline DepSet exp(DepSet a) { a.v = std::exp(a.v); return a; }
inline DepSet log(DepSet a) { a.v = std::log(a.v); return a; }
#define OAD_CMP(op) \
inline bool operator op(const DepSet& a, const DepSet& b) { return a.v op b.v; } \
inline bool operator op(const DepSet& a, double b) { return a.v op b; } \
inline bool operator op(double a, const DepSet& b) { return a op b.v; }
OAD_CMP(<) OAD_CMP(>) OAD_CMP(<=) OAD_CMP(>=) OAD_CMP(==) OAD_CMP(!=)
#undef OAD_CMP
inline DepSet min(const DepSet& a, const DepSet& b) { DepSet r = b.v < a.v ? b : a; r.absorb(a); r.absorb(b); return r; }
inline DepSet max(const DepSet& a, const DepSet& b) { DepSet r = a.v < b.v ? b : a; r.absorb(a); r.absorb(b); return r; }
inline std::ostream& operator<<(std::ostream& os, const DepSet& a) { return os << a.v; }

inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Dual<N>& x) { return x.v; }
inline double value_of(const DepSet& x) { return x.v; }

} // namespace oad
#endif
