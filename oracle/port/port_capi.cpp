// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (oracle/port/structured_port.hpp),
// built into oracle/liboracle_port.so by oracle/Makefile.  Used by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline leg as the checker; never by the GPU product path.
#include "structured_port.hpp"
#include "../jacdriver.hpp"
#include <chrono>
#include <cstring>

#ifndef PORT_DUAL_LANES
#define PORT_DUAL_LANES 8
#endif
typedef oad::Dual<PORT_DUAL_LANES> PDual;

namespace {
struct PortHandle {
    sport::Case c;
    sport::Work<double> wd;
    sport::Work<PDual> wdual;
    sport::Work<oad::DepSet> wdep;
    bool have_ad = false;
};
}

extern "C" {

void* port_create(const sgpu_desc* d) {
    auto h = new PortHandle();
    h->c.init(*d);
    return h;
}
void port_destroy(void* hv) { delete (PortHandle*)hv; }

void port_set_grid(void* hv, const double* xv, const double* yv) {
    auto h = (PortHandle*)hv;
    h->c.set_grid(xv, yv);
    h->wd.init(h->c);
    h->have_ad = false;
}

int port_set_field(void* hv, const char* name, const double* f) {
    auto h = (PortHandle*)hv;
    size_t n = (size_t)h->c.nic*h->c.njc;
    if (!std::strcmp(name, "wall_distance")) h->c.wall_dist.assign(f, f + n);
    else if (!std::strcmp(name, "beta")) h->c.beta.assign(f, f + n);
    else return -1;
    return 0;
}

void port_get_metrics(void* hv, double* nchi, double* neta, double* vol) {
    auto h = (PortHandle*)hv;
    std::memcpy(nchi, h->c.nchi.data(), sizeof(double)*h->c.nchi.size());
    std::memcpy(neta, h->c.neta.data(), sizeof(double)*h->c.neta.size());
    std::memcpy(vol, h->c.vol.data(), sizeof(double)*h->c.vol.size());
}

void port_residual(void* hv, const double* q, double* rhs, int lhs) {
    auto h = (PortHandle*)hv;
    sport::calc_residual<double>(h->c, h->wd, q, rhs, lhs != 0);
}

// padded primitives of the last port_residual: [nic+2][njc+2]
void port_get_primitives(void* hv, double* rho, double* u, double* v, double* p, double* T, double* nut) {
    auto h = (PortHandle*)hv;
    size_t n = h->wd.rho.size()*sizeof(double);
    std::memcpy(rho, h->wd.rho.data(), n); std::memcpy(u, h->wd.u.data(), n); std::memcpy(v, h->wd.v.data(), n);
    std::memcpy(p, h->wd.p.data(), n); std::memcpy(T, h->wd.Tm.data(), n);
    if (nut) std::memcpy(nut, h->wd.nut.data(), n);
}

void port_calc_dt(void* hv, const double* q, double cfl, double* dt) {
    auto h = (PortHandle*)hv;
    sport::calc_dt(h->c, q, cfl, dt);
}

double port_time_residual(void* hv, const double* q, double* rhs, int reps, int lhs) {
    auto h = (PortHandle*)hv;
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int r = 0; r < reps; r++) sport::calc_residual<double>(h->c, h->wd, q, rhs, lhs != 0);
    std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
    return dt.count();
}

int port_jacobian(void* hv, const double* q, int lhs, int* nnz, unsigned int** rind, unsigned int** cind,
                  double** values, int* ncolors) {
    auto h = (PortHandle*)hv;
    if (!h->have_ad) { h->wdual.init(h->c); h->wdep.init(h->c); h->have_ad = true; }
    const size_t n = (size_t)h->c.nic*h->c.njc*h->c.nv;
    const bool l = lhs != 0;
    oad::Coo coo = oad::sparse_jacobian<PORT_DUAL_LANES>(n, q,
        [&](const oad::DepSet* a, oad::DepSet* r) { sport::calc_residual<oad::DepSet>(h->c, h->wdep, a, r, l); },
        [&](const PDual* a, PDual* r) { sport::calc_residual<PDual>(h->c, h->wdual, a, r, l); });
    *nnz = coo.nnz; *rind = coo.rind; *cind = coo.cind; *values = coo.values;
    if (ncolors) *ncolors = coo.ncolors;
    return 0;
}

double port_time_jacobian(void* hv, const double* q, int lhs, int* nnz_out) {
    int nnz; unsigned int *r, *c; double* v; int nc;
    auto t0 = std::chrono::high_resolution_clock::now();
    port_jacobian(hv, q, lhs, &nnz, &r, &c, &v, &nc);
    std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
    if (nnz_out) *nnz_out = nnz;
    free(r); free(c); free(v);
    return dt.count();
}

void port_free(void* p) { free(p); }

} // extern "C"
