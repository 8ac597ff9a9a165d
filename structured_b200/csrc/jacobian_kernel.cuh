// Jacobian kernels (K-JAC).  Placeholder storage type until the kernels land.
#pragma once
#include "common.cuh"
namespace sg {
struct JacStore { double* blocks = nullptr; int slots = 0; size_t cap = 0; bool valid = false; };
inline void jac_free(JacStore& j) { if (j.blocks) cudaFree(j.blocks); j = JacStore(); }
}
