// Drop-in replacement of the reference's src/solver/solver.cpp (same class `Solver<Tx,Tad>` declared in the
// reference's own src/solver/solver.h): the explicit branch of Solver::step runs on the B200 through the C ABI of
// include/structured_gpu.h.  main.cpp, Config, Mesh, IOManager, the .inp format and the grid loaders are the
// reference's unmodified sources, compiled from /root/reference by integration/Makefile.
//   reference lines replaced: calc_dt + residual + stage updates + norms   src/solver/solver.cpp:66,103-134
#include "solver.h"
#include "structured_gpu.h"
#include <map>
#include <cstdlib>

#if defined(STRUCTURED_GPU_IMPLICIT) && !defined(STRUCTURED_GPU_DEVICE_SOLVE)
// The implicit branch of Solver::step (src/solver/solver.cpp:66-101,154-183) WITHOUT ADOL-C: the sparse Jacobian comes
// from sgpu_jacobian_coo and is consumed by the reference's own, unmodified LinearSolverEigen (Eigen SparseLU,
// src/linearsolver/ls_eigen.cpp).  That file instantiates the class for <double, adouble>; with Tad = double there
// is no adouble, so the name is mapped onto double for this one include.
#define adouble double
#include "ls_eigen.cpp"
#undef adouble
#endif

template<class Tx> void set_rarray(size_t size, Tx* __restrict__ dest, Tx* __restrict__ src) { for (size_t i = 0; i < size; i++) dest[i] = src[i]; }
template<class Tx> void update_forward_euler(size_t size, Tx* __restrict__ q, Tx* __restrict__ rhs, Tx* __restrict__ dt) { for (size_t i = 0; i < size; i++) q[i] = q[i] + rhs[i]*dt[i]; }
template<class Tx, class To> void update_rk4(size_t size, Tx* __restrict__ q_i, Tx* __restrict__ q, Tx* __restrict__ rhs, Tx* __restrict__ dt, To order) { for (size_t i = 0; i < size; i++) q_i[i] = q[i] + rhs[i]*dt[i]/(4.0 - order); }

static void gpu_check(int rc, sgpu_ctx* ctx, const char* what) {
    if (rc == 0) return;
    spdlog::get("console")->critical("{}: {}", what, sgpu_last_error(ctx));
    std::abort();
}

// ---- SA extension through the drop-in -----------------------------------------------------------------------------
// The reference's Solution is laminar (ntrans = 0 hard-coded, src/solver/solution.cpp:9): its arrays hold four variables
// per cell.  An OPTIONAL `[turbulence]` table in the .inp file (stock files have none and behave exactly as before) turns
// the Spalart-Allmaras extension on:
//     [turbulence]
//     ntrans = 1
//     wall_distance = "compute"      # or the name of a file of nic*njc raw doubles, [i][j]
//     beta_file = "beta.bin"         # optional correction field beta(x), nic*njc raw doubles, default 1
// The five-variable state then lives on the device and in `g_q5` (host, [nic][njc][5]); Solution::q receives the four
// mean-flow variables before IOManager::write, and the complete state is written next to the reference's files as
// `<label>.sa.out` (raw doubles, the layout of the reference's `.out` restart with nv = 5).
struct TurbulenceCfg { int ntrans = 0; std::string wall_distance = "compute", beta_file; };
static TurbulenceCfg g_turb;
static std::vector<double> g_q5;

static bool read_raw(const std::string& fn, std::vector<double>& a, size_t n) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return false;
    a.resize(n);
    const size_t got = fread(a.data(), sizeof(double), n, f);
    fclose(f);
    return got == n;
}

// builds the context from what Mesh/Config/BoundaryContainer hold (src/model/bc.cpp:470-488, src/utils/config.cpp:32-87)
template <class Tx, class Tad>
static sgpu_ctx* make_ctx(std::shared_ptr<Mesh<Tx, Tad>> mesh, std::shared_ptr<Config<Tx>> cfg) {
    static const std::map<std::string, int> T{{"freestream", 0}, {"slipwall", 1}, {"wall", 2}, {"isothermalwall", 3}, {"wake", 4}, {"outflow", 5}, {"periodic", 6}};
    static const std::map<std::string, int> F{{"bottom", 0}, {"right", 1}, {"top", 2}, {"left", 3}};
    std::vector<sgpu_bc> bcs;
    auto toml = cpptoml::parse_file(cfg->filename);
    for (const auto& b : *toml->get_table_array("boundary")) {
        const std::string type = b->template get_qualified_as<std::string>("type").value_or("");
        const std::string face = b->template get_qualified_as<std::string>("face").value_or("");
        if (!T.count(type)) { spdlog::get("console")->info("Wrong type of BC."); continue; }
        sgpu_bc e{};
        e.type = T.at(type); e.face = F.at(face);
        e.start = (int)b->template get_qualified_as<int64_t>("start").value_or(0);
        e.end = (int)b->template get_qualified_as<int64_t>("end").value_or(0);
        e.u = b->template get_qualified_as<double>("u").value_or(0.0);
        e.v = b->template get_qualified_as<double>("v").value_or(0.0);
        e.T = b->template get_qualified_as<double>("T").value_or(0.0);
        bcs.push_back(e);
    }
    g_turb.ntrans = (int)toml->template get_qualified_as<int64_t>("turbulence.ntrans").value_or(0);
    g_turb.wall_distance = toml->template get_qualified_as<std::string>("turbulence.wall_distance").value_or("compute");
    g_turb.beta_file = toml->template get_qualified_as<std::string>("turbulence.beta_file").value_or("");
#if defined(STRUCTURED_GPU_IMPLICIT) && !defined(STRUCTURED_GPU_DEVICE_SOLVE)
    if (g_turb.ntrans) { spdlog::get("console")->critical("[turbulence] needs the device-resident binaries: LinearSolverEigen is sized for the reference's four variables"); std::abort(); }
#endif
    sgpu_desc d{};
    d.ni = (int)mesh->ni; d.nj = (int)mesh->nj; d.ntrans = g_turb.ntrans ? g_turb.ntrans : (int)mesh->solution->ntrans;
    d.order = (int)cfg->solver->order; d.lhs_order = (int)cfg->solver->lhs_order;
    d.flux = cfg->solver->flux == "roe" ? SGPU_FLUX_ROE : SGPU_FLUX_AUSM;
    d.rho_inf = cfg->freestream->rho_inf; d.u_inf = cfg->freestream->u_inf; d.v_inf = cfg->freestream->v_inf;
    d.p_inf = cfg->freestream->p_inf; d.T_inf = cfg->freestream->T_inf; d.mu_inf = cfg->freestream->mu_inf;
    d.pr_inf = cfg->freestream->pr_inf; d.dpdx = cfg->solver->dpdx; d.dpdy = cfg->solver->dpdy;
    d.n_bc = (int)bcs.size(); d.bc = bcs.data(); d.device = 0;
    sgpu_ctx* ctx = nullptr;
    if (sgpu_create(&d, &ctx)) { spdlog::get("console")->critical("sgpu_create: {}", sgpu_last_error(nullptr)); std::abort(); }
    gpu_check(sgpu_set_grid(ctx, mesh->xv.data(), mesh->yv.data()), ctx, "sgpu_set_grid");
    if (!g_turb.ntrans) {
        gpu_check(sgpu_set_state(ctx, SGPU_STATE_Q, mesh->solution->q.data()), ctx, "sgpu_set_state");
        gpu_check(sgpu_set_state(ctx, SGPU_STATE_Q_TMP, mesh->solution->q_tmp.data()), ctx, "sgpu_set_state");
        return ctx;
    }
    const size_t nc = (size_t)mesh->nic*mesh->njc;
    std::vector<double> field;
    if (g_turb.wall_distance == "compute") gpu_check(sgpu_wall_distance_from_bcs(ctx, mesh->xv.data(), mesh->yv.data()), ctx, "sgpu_wall_distance_from_bcs");
    else {
        if (!read_raw(g_turb.wall_distance, field, nc)) { spdlog::get("console")->critical("turbulence.wall_distance: cannot read {} doubles from {}", nc, g_turb.wall_distance); std::abort(); }
        gpu_check(sgpu_set_field(ctx, "wall_distance", field.data()), ctx, "sgpu_set_field");
    }
    if (!g_turb.beta_file.empty()) {
        if (!read_raw(g_turb.beta_file, field, nc)) { spdlog::get("console")->critical("turbulence.beta_file: cannot read {} doubles from {}", nc, g_turb.beta_file); std::abort(); }
        gpu_check(sgpu_set_field(ctx, "beta", field.data()), ctx, "sgpu_set_field");
    }
    // state: the reference's initial / restarted mean flow + rho nu~ = 3 mu_inf (the freestream value of the SA ghost rule),
    // or the complete five-variable state of an earlier run when io.restart is set and <label>.sa.out exists
    g_q5.resize(nc*5);
    if (!(cfg->io->restart && read_raw(cfg->io->label + ".sa.out", g_q5, nc*5))) {
        const double* q4 = mesh->solution->q.data();
        for (size_t c = 0; c < nc; c++) {
            for (int k = 0; k < 4; k++) g_q5[5*c + k] = q4[4*c + k];
            g_q5[5*c + 4] = 3.0*cfg->freestream->mu_inf;
        }
    }
    gpu_check(sgpu_set_state(ctx, SGPU_STATE_Q, g_q5.data()), ctx, "sgpu_set_state");
    gpu_check(sgpu_set_state(ctx, SGPU_STATE_Q_TMP, g_q5.data()), ctx, "sgpu_set_state");
    return ctx;
}

template<class Tx, class Tad>
void Solver<Tx, Tad>::add_mesh(std::shared_ptr<Mesh<Tx,Tad>> mesh) { mesh_list.push_back(mesh); }

template <class Tx, class Tad>
Solver<Tx, Tad>::Solver(std::shared_ptr<Config<Tx>> val_config) {
    config = val_config;
    label = config->io->label;
    logger_convergence = spdlog::basic_logger_mt("convergence", label + ".history", true);
    logger_convergence->info(" ");
    logger = spdlog::get("console");
    CFL = config->solver->cfl;
    UNDER_RELAXATION = config->solver->under_relaxation;
}
template <class Tx, class Tad> Solver<Tx, Tad>::~Solver() {}

template <class Tx, class Tad>
bool Solver<Tx, Tad>::step(std::shared_ptr<Mesh<Tx,Tad>> mesh, size_t counter, Tx t) {
    static sgpu_ctx* ctx = make_ctx<Tx, Tad>(mesh, config);        // one context per Mesh (main adds exactly one, src/main.cpp:40)
    auto solution = mesh->solution;
    const size_t nv = solution->nq + (g_turb.ntrans ? g_turb.ntrans : solution->ntrans);
    double l2sq[8] = {0}, l2norm[8] = {0};
    // steps that end in IOManager::write (src/solver/solver.cpp:136-141,192-194) keep the wall rows of their last residual evaluation
    const bool will_write = counter > config->solver->iteration_max || counter % config->io->fileout_frequency == 0;
    gpu_check(sgpu_track_wall(ctx, will_write ? 1 : 0), ctx, "sgpu_track_wall");
    config->profiler->reset_time_residual();
#if defined(STRUCTURED_GPU_DEVICE_SOLVE)
    // The whole implicit branch on the device (src/solver/solver.cpp:66-101,154-175): no COO export, no host linear solver.
    // The reference solves its LHS exactly (SparseLU); GMRES is driven to round-off so the iteration history matches it.
    if (!(counter > config->solver->iteration_max)) {
        sgpu_linsolve ls{};
        ls.precond = SGPU_PC_LINE_J; ls.restart = 60; ls.max_iter = 2000; ls.rtol = 1e-13;
        gpu_check(sgpu_implicit_step(ctx, CFL, UNDER_RELAXATION, &ls, l2sq), ctx, "sgpu_implicit_step");
        if (!ls.converged) logger->warn("GMRES stopped at |r|/|b| = {:.2e} after {} iterations", ls.rel_residual, ls.iterations);
        logger->debug("GMRES iterations = {}", ls.iterations);
    } else {
        gpu_check(sgpu_calc_dt(ctx, CFL), ctx, "sgpu_calc_dt");
        gpu_check(sgpu_residual(ctx, SGPU_STATE_Q, 0, l2sq), ctx, "sgpu_residual");
    }
    config->profiler->update_time_residual();
#elif !defined(STRUCTURED_GPU_IMPLICIT)
    int scheme = -1;
    if (config->solver->scheme == "forward_euler") scheme = 0;
    else if (config->solver->scheme == "rk4_jameson") scheme = 1;
    else logger->critical("scheme not defined.");
    if (scheme >= 0) gpu_check(sgpu_explicit_step(ctx, scheme, CFL, l2sq), ctx, "sgpu_explicit_step");
    config->profiler->update_time_residual();
#else
    static LinearSolverEigen<Tx, Tad>* linearsolver = new LinearSolverEigen<Tx, Tad>(mesh, config);
    gpu_check(sgpu_calc_dt(ctx, CFL), ctx, "sgpu_calc_dt");                                   // solver.cpp:66
    gpu_check(sgpu_residual(ctx, SGPU_STATE_Q, /*lhs=*/0, l2sq), ctx, "sgpu_residual");     // rhs with solver.order (solver.cpp:92-101)
    gpu_check(sgpu_get_rhs(ctx, solution->rhs.data()), ctx, "sgpu_get_rhs");
    config->profiler->update_time_residual();
    if (!(counter > config->solver->iteration_max)) {
        config->profiler->reset_time_jacobian();
        // replaces trace_on .. trace_off + sparse_jac (solver.cpp:72-90,156) and the LHS loop (solver.cpp:162-171)
        gpu_check(sgpu_jacobian_coo(ctx, &solution->nnz, &solution->rind, &solution->cind, &solution->values, 1), ctx, "sgpu_jacobian_coo");
        config->profiler->update_time_jacobian();
        logger->debug("NNZ = {}", solution->nnz);
        if (counter == 0) linearsolver->preallocate(solution->nnz);
        config->profiler->reset_time_linearsolver();
        gpu_check(sgpu_get_state(ctx, SGPU_STATE_Q, solution->q.data()), ctx, "sgpu_get_state");
        linearsolver->set_lhs(solution->nnz, solution->rind, solution->cind, solution->values);
        linearsolver->set_rhs(solution->rhs.data());
        linearsolver->solve_and_update(solution->q.data(), UNDER_RELAXATION);
        auto dt_perf = config->profiler->update_time_linearsolver();
        logger->info("Linear algebra time = {:03.2f}", dt_perf);
        gpu_check(sgpu_set_state(ctx, SGPU_STATE_Q, solution->q.data()), ctx, "sgpu_set_state");
        free(solution->rind); solution->rind = nullptr;                                       // unchanged ownership (solver.cpp:181-183)
        free(solution->cind); solution->cind = nullptr;
        free(solution->values); solution->values = nullptr;
    }
#endif
    for (size_t k = 0; k < nv; k++) l2norm[k] = sqrt(l2sq[k]);
    // what IOManager::write reads: Solution::q, and -- in write_surface (src/utils/io.cpp:215-216,224-234) -- the wall-face
    // rows grad_u_eta[i][0], grad_v_eta[i][0] of EulerEquation's work arrays as the LAST calc_residual of this step left them
    // (tracked on the device for the steps that write, see below)
    auto sync_host = [&]() {
        if (!g_turb.ntrans) gpu_check(sgpu_get_state(ctx, SGPU_STATE_Q, solution->q.data()), ctx, "sgpu_get_state");
        else {                                                      // four mean-flow variables for the reference's writers + the full state
            gpu_check(sgpu_get_state(ctx, SGPU_STATE_Q, g_q5.data()), ctx, "sgpu_get_state");
            const size_t nc = (size_t)mesh->nic*mesh->njc;
            double* q4 = solution->q.data();
            for (size_t c = 0; c < nc; c++) for (int k = 0; k < 4; k++) q4[4*c + k] = g_q5[5*c + k];
            FILE* f = fopen((label + ".sa.out").c_str(), "wb");
            if (f) { fwrite(g_q5.data(), sizeof(double), g_q5.size(), f); fclose(f); }
            logger->info("SA: |rhs(rho nu~)| = {:.3e}, full state written to {}.sa.out", l2norm[4], label);
        }
        const size_t nic = mesh->nic;
        std::vector<double> gu(2*nic), gv(2*nic);
        gpu_check(sgpu_wall_data(ctx, SGPU_STATE_LAST_RESIDUAL, SGPU_STATE_Q, gu.data(), gv.data(), nullptr, nullptr), ctx, "sgpu_wall_data");
        auto eq = mesh->equation;
        for (size_t i = 0; i < nic; i++)
            for (size_t k = 0; k < 2; k++) { eq->grad_u_eta[i][0][k] = gu[2*i + k]; eq->grad_v_eta[i][0][k] = gv[2*i + k]; }
    };
    if (counter > config->solver->iteration_max) {
        logger->info("Max iteration reached!");
        sync_host();
        mesh->iomanager->write(counter);
        auto dt_main = config->profiler->current_time();
        logger->info("Final:: Step: {:08d} Time: {:.2e} Wall Time: {:.2e} CFL: {:.2e} Density Norm: {:.2e}", counter, t, dt_main, CFL, l2norm[0]);
        logger_convergence->info("{:08d} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e}", counter, t, dt_main, CFL, l2norm[0], l2norm[1], l2norm[2], l2norm[3]);
        return true;
    }
    if (counter % config->io->stdout_frequency == 0) {
        auto dt_main = config->profiler->current_time();
        logger->info("Step: {:08d} Time: {:.2e} Wall Time: {:.2e} CFL: {:.2e} Density Norm: {:.2e}", counter, t, dt_main, CFL, l2norm[0]);
        logger_convergence->info("{:08d} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e} {:.2e}", counter, t, dt_main, CFL, l2norm[0], l2norm[1], l2norm[2], l2norm[3]);
    }
    if (counter % config->io->fileout_frequency == 0) { sync_host(); mesh->iomanager->write(counter); }
    return false;
}

template <class Tx, class Tad>
void Solver<Tx, Tad>::solve() {
    size_t counter = 0;
    Tx t = 0.0;
    logger->info("Welcome to structured! (residual path on the GPU)");
    config->profiler->timer_main->reset();
    while (1) {
        bool if_break = true;
        for (auto&& mesh : mesh_list) if_break = step(mesh, counter, t) && if_break;
        if (if_break) break;
        counter += 1;
        if (config->solver->cfl_ramp && counter > config->solver->cfl_ramp_iteration) {
            CFL = pow(CFL, config->solver->cfl_ramp_exponent);
            CFL = std::min(CFL, static_cast<Tx>(1e12));
        }
        if (config->solver->under_relaxation_ramp && counter > config->solver->under_relaxation_ramp_iteration) {
            UNDER_RELAXATION = pow(UNDER_RELAXATION, config->solver->under_relaxation_ramp_exponent);
            UNDER_RELAXATION = std::min(UNDER_RELAXATION, static_cast<Tx>(10.0));
        }
    }
}

template class Solver<double, double>;
