int sgpu_jacobian_coo(sgpu_ctx* c, int* nnz, unsigned int** rind, unsigned int** cind, double** values, int apply_lhs_transform) {
    (void)nnz; (void)rind; (void)cind; (void)values; (void)apply_lhs_transform;
    if (!c) return SGPU_ERR_ARG;
    FAIL(c, SGPU_ERR_STATE, "Jacobian kernels not built yet");
}
int sgpu_jacobian_device(sgpu_ctx* c, int* slots, float* build_ms) {
    (void)slots; (void)build_ms;
    if (!c) return SGPU_ERR_ARG;
    FAIL(c, SGPU_ERR_STATE, "Jacobian kernels not built yet");
}
int sgpu_jacobian_apply(sgpu_ctx* c, int transpose, const double* x, double* y) {
    (void)transpose; (void)x; (void)y;
    if (!c) return SGPU_ERR_ARG;
    FAIL(c, SGPU_ERR_STATE, "Jacobian kernels not built yet");
}
