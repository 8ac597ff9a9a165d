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

// a case that is a cropped window of a larger grid keeps that grid's limiter constants (reconstruction.cpp:62-63)
void port_set_global_counts(void* hv, int nic_global, int njc_global) {
    auto h = (PortHandle*)hv;
    h->c.eps_nic = nic_global; h->c.eps_njc = njc_global;
}

int port_set_field(void* hv, const char* name, const double* f) {
    auto h = (PortHandle*)hv;
    size_t n = (size_t)h->c.nic*h->c.njc;
    if (!std::strcmp(name, "wall_distance")) h->c.wall_dist.assign(f, f + n);
    else if (!std::strcmp(name, "beta")) h->c.beta.assign(f, f + n);
    else return -1;
    return 0;
}

// SA wall distance (extension; no reference counterpart): for every cell the distance from its centre (Mesh::xc, yc,
// src/utils/mesh.cpp:199-200) to the nearest point of the boundary edges covered by a `wall` / `isothermalwall` table.
// Brute force over all edges; stores the field in the case and copies it to out [nic][njc] (may be NULL).
int port_wall_distance(void* hv, double* out) {
    auto h = (PortHandle*)hv;
    sport::Case& c = h->c;
    struct Seg { double ax, ay, bx, by, il2; };
    std::vector<Seg> segs;
    auto X = [&](const std::vector<double>& a, int i, int j) { return a[(size_t)i*c.nj + j]; };
    for (const sgpu_bc& b : c.bcs) {
        if (b.type != SGPU_BC_WALL && b.type != SGPU_BC_ISOTHERMALWALL) continue;
        const bool horiz = b.face == SGPU_FACE_BOTTOM || b.face == SGPU_FACE_TOP;
        const int n = horiz ? c.nic : c.njc;
        for (int p = std::max(b.start, 1); p <= std::min(b.end, n); p++) {
            int i0, j0, i1, j1;
            if (horiz) { i0 = p - 1; i1 = p; j0 = j1 = (b.face == SGPU_FACE_BOTTOM ? 0 : c.nj - 1); }
            else { j0 = p - 1; j1 = p; i0 = i1 = (b.face == SGPU_FACE_LEFT ? 0 : c.ni - 1); }
            Seg s; s.ax = X(c.xv, i0, j0); s.ay = X(c.yv, i0, j0); s.bx = X(c.xv, i1, j1) - s.ax; s.by = X(c.yv, i1, j1) - s.ay;
            const double l2 = s.bx*s.bx + s.by*s.by; s.il2 = l2 > 0.0 ? 1.0/l2 : 0.0;
            segs.push_back(s);
        }
    }
    for (int i = 0; i < c.nic; i++) for (int j = 0; j < c.njc; j++) {
        const double px = 0.25*(X(c.xv, i, j) + X(c.xv, i+1, j) + X(c.xv, i, j+1) + X(c.xv, i+1, j+1));
        const double py = 0.25*(X(c.yv, i, j) + X(c.yv, i+1, j) + X(c.yv, i, j+1) + X(c.yv, i+1, j+1));
        double best = 1e300;
        for (const Seg& s : segs) {
            const double dx = px - s.ax, dy = py - s.ay;
            double t = (dx*s.bx + dy*s.by)*s.il2;
            t = std::min(std::max(t, 0.0), 1.0);
            const double ex = dx - t*s.bx, ey = dy - t*s.by;
            best = std::min(best, ex*ex + ey*ey);
        }
        c.wall_dist[(size_t)i*c.njc + j] = segs.empty() ? 1e30 : std::sqrt(best);
    }
    if (out) std::memcpy(out, c.wall_dist.data(), sizeof(double)*c.wall_dist.size());
    return (int)segs.size();
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

// Jacobian ROWS of sampled cells on grids too large for the full pattern + colouring pipeline above: static colouring
// colour(i, j, k) = ((i mod 5)*5 + (j mod 5))*nv + k -- two cells of one colour are >= 5 apart in i or j, so they never
// share a row (row stencil: |di|, |dj| <= 2; SURVEY.md Appendix B "Colouring").  ceil(25 nv / lanes) dual sweeps over the
// whole grid, then for each sampled row cell the derivative w.r.t. the 25 cells of its 5x5 window is read off.
// Valid for rows whose ghost cells depend on cells INSIDE that window only (walls, slipwall, outflow, freestream);
// rows within 2 cells of a periodic or wake boundary must not be sampled (their wrap-around columns alias).
// cells: [ncells][2] = (i, j);  out: [ncells][nv][25][nv] = d rhs[i][j][r] / d q[i+di][j+dj][c], window index (di+2)*5 + (dj+2),
// zero for window cells outside the grid.
int port_jacobian_rows(void* hv, const double* q, int lhs, const int* cells, int ncells, double* out) {
    auto h = (PortHandle*)hv;
    if (!h->have_ad) { h->wdual.init(h->c); h->wdep.init(h->c); h->have_ad = true; }
    const sport::Case& c = h->c;
    const int nv = c.nv, N = PORT_DUAL_LANES;
    const size_t n = (size_t)c.nic*c.njc*nv;
    const int ncolors = 25*nv;
    std::vector<PDual> a_q(n), a_rhs(n);
    std::memset(out, 0, sizeof(double)*(size_t)ncells*nv*25*nv);
    auto color = [&](int i, int j, int k) { return ((i % 5)*5 + (j % 5))*nv + k; };
    for (int c0 = 0; c0 < ncolors; c0 += N) {
        for (int i = 0; i < c.nic; i++) for (int j = 0; j < c.njc; j++) for (int k = 0; k < nv; k++) {
            const size_t id = ((size_t)i*c.njc + j)*nv + k;
            a_q[id] = PDual(q[id]);
            const int l = color(i, j, k) - c0;
            if (l >= 0 && l < N) a_q[id].d[l] = 1.0;
        }
        sport::calc_residual<PDual>(c, h->wdual, a_q.data(), a_rhs.data(), lhs != 0);
        for (int m = 0; m < ncells; m++) {
            const int i = cells[2*m], j = cells[2*m + 1];
            for (int di = -2; di <= 2; di++) for (int dj = -2; dj <= 2; dj++) {
                const int ci = i + di, cj = j + dj;
                if (ci < 0 || ci >= c.nic || cj < 0 || cj >= c.njc) continue;
                for (int k = 0; k < nv; k++) {
                    const int l = color(ci, cj, k) - c0;
                    if (l < 0 || l >= N) continue;
                    for (int r = 0; r < nv; r++)
                        out[(((size_t)m*nv + r)*25 + (di + 2)*5 + (dj + 2))*nv + k] = a_rhs[((size_t)i*c.njc + j)*nv + r].d[l];
                }
            }
        }
    }
    return 0;
}

void port_free(void* p) { free(p); }

} // extern "C"
