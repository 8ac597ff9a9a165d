// TEST INFRASTRUCTURE ONLY -- not part of the shipped GPU path.
//
// CPU stand-in for ADOL-C's `sparse_jac(tag, m, n, repeat, x, &nnz, &rind, &cind, &values, options)`
// as called by the reference at src/solver/solver.cpp:156 with options {0,0,0,0}
// (src/solver/solution.h:21): index-domain sparsity detection -> column colouring ->
// forward vector sweeps -> recovery into row-major COO.  ADOL-C/ColPack are absent from
// /root/reference (un-vendored, un-pinned: config.py:1-2) and from this image, so the same
// four stages are restated here on top of oracle/adtypes.hpp:
//   1. pattern : one residual evaluation with T = DepSet (structural pattern incl. numerical zeros)
//   2. colour  : greedy distance-1 colouring of the column intersection graph
//   3. sweeps  : ceil(ncolours / N) residual evaluations with T = Dual<N>
//   4. recover : values[nz] = d(rhs[row]) / d(colour(col)), emitted sorted by (row, col)
// The evaluator is any callable `void eval(const T* q, T* rhs)` over flat arrays of length n
// in the reference's order (i*njc + j)*nv + k  (src/solver/solver.cpp:73-88,164-166).
#ifndef ORACLE_JACDRIVER_HPP
#define ORACLE_JACDRIVER_HPP
#include "adtypes.hpp"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace oad {

struct Coo {
    int nnz = 0;
    unsigned int* rind = nullptr;
    unsigned int* cind = nullptr;
    double* values = nullptr;
    int ncolors = 0;
};

// rows[r] = sorted list of structural column indices of row r
template <class EvalDep>
inline void detect_pattern(size_t n, const double* q, EvalDep&& eval, std::vector<std::vector<uint32_t>>& rows) {
    std::vector<DepSet> a_q(n), a_rhs(n);
    for (size_t c = 0; c < n; c++) { a_q[c].v = q[c]; a_q[c].s.assign(1, (uint32_t)c); }
    eval(a_q.data(), a_rhs.data());
    rows.resize(n);
    for (size_t r = 0; r < n; r++) rows[r] = a_rhs[r].s;
}

// Greedy colouring: two columns may share a colour iff no row contains both.
inline int color_columns(size_t n, const std::vector<std::vector<uint32_t>>& rows, std::vector<int>& color) {
    std::vector<std::vector<uint32_t>> cols(n);           // rows touching each column
    for (size_t r = 0; r < n; r++) for (uint32_t c : rows[r]) cols[c].push_back((uint32_t)r);
    color.assign(n, -1);
    std::vector<int> mark;                                 // mark[colour] = last column that forbade it
    int ncolors = 0;
    for (size_t c = 0; c < n; c++) {
        for (uint32_t r : cols[c]) for (uint32_t c2 : rows[r]) {
            int k = color[c2];
            if (k >= 0) { if ((size_t)k >= mark.size()) mark.resize(k + 1, -1); mark[k] = (int)c; }
        }
        int k = 0;
        while ((size_t)k < mark.size() && mark[k] == (int)c) k++;
        color[c] = k;
        if (k + 1 > ncolors) ncolors = k + 1;
        if ((size_t)k >= mark.size()) mark.resize(k + 1, -1);
    }
    return ncolors;
}

template <int N, class EvalDep, class EvalDual>
inline Coo sparse_jacobian(size_t n, const double* q, EvalDep&& eval_dep, EvalDual&& eval_dual) {
    std::vector<std::vector<uint32_t>> rows;
    detect_pattern(n, q, eval_dep, rows);
    std::vector<int> color;
    const int ncolors = color_columns(n, rows, color);
    size_t nnz = 0;
    std::vector<size_t> rowptr(n + 1, 0);
    for (size_t r = 0; r < n; r++) { rowptr[r] = nnz; nnz += rows[r].size(); }
    rowptr[n] = nnz;
    Coo out;
    out.nnz = (int)nnz;
    out.ncolors = ncolors;
    // malloc: the reference frees these with free() (src/solver/solver.cpp:181-183)
    out.rind = (unsigned int*)std::malloc(sizeof(unsigned int)*(nnz ? nnz : 1));
    out.cind = (unsigned int*)std::malloc(sizeof(unsigned int)*(nnz ? nnz : 1));
    out.values = (double*)std::malloc(sizeof(double)*(nnz ? nnz : 1));
    for (size_t r = 0; r < n; r++) for (size_t k = 0; k < rows[r].size(); k++) {
        out.rind[rowptr[r] + k] = (unsigned int)r;
        out.cind[rowptr[r] + k] = rows[r][k];
    }
    std::vector<Dual<N>> a_q(n), a_rhs(n);
    for (int c0 = 0; c0 < ncolors; c0 += N) {
        for (size_t c = 0; c < n; c++) {
            a_q[c] = Dual<N>(q[c]);
            const int l = color[c] - c0;
            if (l >= 0 && l < N) a_q[c].d[l] = 1.0;
        }
        eval_dual(a_q.data(), a_rhs.data());
        for (size_t r = 0; r < n; r++) for (size_t k = 0; k < rows[r].size(); k++) {
            const int l = color[rows[r][k]] - c0;
            if (l >= 0 && l < N) out.values[rowptr[r] + k] = a_rhs[r].d[l];
        }
    }
    return out;
}

} // namespace oad
#endif
