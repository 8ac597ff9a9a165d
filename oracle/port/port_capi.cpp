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

// IOManager::write_surface, src/utils/io.cpp:182-255.  The reference reads grad_{u,v}_eta[i][0] as the last
// calc_residual left them (state q_res) and Solution::p[i][0..1] recomputed from the final q (io.cpp:41, unpadded,
// shift 0): per cell column i in [i_first, i_first + count) (reference: j1 - 1 .. j1 - 1 + nb) xw, cp, cf and
// coeffs[6] = {cl_p, cd_p, cl_v, cd_v, cl, cd}.  wall[6][nic] (may be NULL) = gu_x, gu_y, gv_x, gv_y, p0, p1.
void port_surface(void* hv, const double* q_res, const double* q_fin, int i_first, int count, double aoa,
                  double* xw, double* cp_out, double* cf_out, double* coeffs, double* wall) {
    auto h = (PortHandle*)hv;
    const sport::Case& c = h->c;
    std::vector<double> rhs((size_t)c.nic*c.njc*c.nv);
    sport::calc_residual<double>(c, h->wd, q_res, rhs.data(), false);
    auto pcell = [&](int i, int j) {                                             // fluid.cpp:50-67 with shift 0
        const double* Q = q_fin + ((size_t)i*c.njc + j)*c.nv;
        double r = Q[0], uu = Q[1]/r, vv = Q[2]/r;
        return (Q[3] - 0.5*r*(uu*uu + vv*vv))*(sport::GAMMA - 1.0);
    };
    auto X = [&](const std::vector<double>& a, int i, int j) { return a[(size_t)i*c.nj + j]; };
    const sport::Work<double>& w = h->wd;
    if (wall) for (int i = 0; i < c.nic; i++) {
        wall[i] = w.gu_eta[w.E(i, 0)*2]; wall[c.nic + i] = w.gu_eta[w.E(i, 0)*2 + 1];
        wall[2*c.nic + i] = w.gv_eta[w.E(i, 0)*2]; wall[3*c.nic + i] = w.gv_eta[w.E(i, 0)*2 + 1];
        wall[4*c.nic + i] = pcell(i, 0); wall[5*c.nic + i] = pcell(i, 1);
    }
    double Fn_pressure = 0.0, Fc_pressure = 0.0, Fn_viscous = 0.0, Fc_viscous = 0.0;
    for (int i = i_first; i < i_first + count; i++) {                            // :219-238
        const double gux = w.gu_eta[w.E(i, 0)*2], guy = w.gu_eta[w.E(i, 0)*2 + 1];
        const double gvx = w.gv_eta[w.E(i, 0)*2], gvy = w.gv_eta[w.E(i, 0)*2 + 1];
        double qinf = 0.5*c.rho_inf*(c.u_inf*c.u_inf + c.v_inf*c.v_inf);
        double cp = (0.5*(pcell(i, 0) + pcell(i, 1)) - c.p_inf)/qinf;
        double tau = c.mu_inf*(guy - gvx)/qinf;
        if (xw) xw[i - i_first] = 0.25*(X(c.xv, i, 0) + X(c.xv, i+1, 0) + X(c.xv, i, 1) + X(c.xv, i+1, 1));   // mesh.cpp:199
        if (cp_out) cp_out[i - i_first] = cp;
        if (cf_out) cf_out[i - i_first] = tau;
        double dx = X(c.xv, i+1, 0) - X(c.xv, i, 0), dy = X(c.yv, i+1, 0) - X(c.yv, i, 0);
        Fn_pressure = Fn_pressure - cp*dx;
        Fc_pressure = Fc_pressure + cp*dy;
        double sfdiv = 2.0/3.0*(gux + gvy);
        double sxx = c.mu_inf*(2.0*gux - sfdiv)/qinf;
        double syy = c.mu_inf*(2.0*gvy - sfdiv)/qinf;
        Fn_viscous = Fn_viscous - tau*dy + syy*dx;
        Fc_viscous = Fc_viscous + tau*dx - sxx*dy;
    }
    if (coeffs) {                                                                // :240-249
        double ca = std::cos(aoa), sa = std::sin(aoa);
        coeffs[0] = -Fc_pressure*sa + Fn_pressure*ca; coeffs[1] = Fc_pressure*ca + Fn_pressure*sa;
        coeffs[2] = -Fc_viscous*sa + Fn_viscous*ca; coeffs[3] = Fc_viscous*ca + Fn_viscous*sa;
        coeffs[4] = coeffs[2] + coeffs[0]; coeffs[5] = coeffs[3] + coeffs[1];
    }
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
