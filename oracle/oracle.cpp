// oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of eagle-mpc's SbFDDP hot path (single OCP, scalar).
//
// PARITY UNPINNED: the reference ships no tests / golden vectors, and its arithmetic (PepMS Crocoddyl fork,
// Pinocchio, example-robot-data URDFs) is absent from /root/reference and cannot be built or imported here
// (SURVEY.md §0, §8c).  This restates the algorithm from in-tree src/sbfddp.cpp plus the published Crocoddyl 1.x
// SolverDDP/SolverFDDP algorithm; the fork-only stop rules are *inferred*  and isolated in
// stopping_criteria()/stopping_test*() below.  It is validated in tests/ by finite differences, algebraic identities
// and an LQR closed form, and is the checker for the CUDA path — never the thing measured or shipped.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this library.
//
// Control flow followed line by line:
//   SolverSbFDDP::solve            src/sbfddp.cpp:192-226
//   SolverSbFDDP::solveFDDP        src/sbfddp.cpp:228-315
//   SolverSbFDDP::solveDDP         src/sbfddp.cpp:317-393
//   expectedImprovementDDP         src/sbfddp.cpp:395-408
//   tryStepDDP / forwardPassDDP    src/sbfddp.cpp:410-460
//   squashingUpdate/barrierUpdate  src/sbfddp.cpp:462-477
//   fillSquashedOutputs            src/sbfddp.cpp:479-486
// Upstream (restated): SolverAbstract::setCandidate, SolverDDP::{calcDiff,backwardPass,computeGains,
//   increase/decreaseRegularization}, SolverFDDP::{forwardPass,updateExpectedImprovement,expectedImprovement}.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "oracle_model.hpp"

namespace orc {

struct Solver {
  // deep copy of the problem
  empc_problem_desc_t desc;
  std::vector<int32_t> costset_begin, node_costset, costset_contact;
  std::vector<empc_cost_t> costs;
  std::vector<empc_contact_t> contacts;
  std::vector<double> pool;
  Model m;
  empc_solver_params_t P;
  int T = 0, node_map = 0;

  std::vector<double> x0, xs, us, xs_try, us_try, fs, dx, K, k, Vx, Vxx, Qu, Quuk, tiles, us_squash;
  std::vector<Work> work;
  std::vector<std::vector<CostEval>> evals;
  bool is_feasible = false, was_feasible = false;
  double cost = 0, cost_prev = 0, cost_try = 0, xreg = 0, ureg = 0, steplength = 1, dV = 0, dVexp = 0;
  double dg = 0, dq = 0, dv = 0, d0 = 0, d1 = 0, stop = 0, th_stop = 0, convergence = 0;
  double smooth = 0.1;        // SolverSbFDDP::smooth_ (the schedule variable)
  double smooth_model = 0.1;  // what squashingUpdate()/barrierUpdate() last pushed into the models
  int iter = 0;
  long total_iters = 0;
  double alphas[EMPC_N_ALPHAS];

  void init(const empc_problem_desc_t* d) {
    desc = *d;
    costset_begin.assign(d->costset_begin, d->costset_begin + d->n_costsets + 1);
    costs.assign(d->costs, d->costs + d->n_costs);
    pool.assign(d->pool, d->pool + d->n_pool);
    node_costset.assign(d->node_costset, d->node_costset + (size_t)d->n_node_maps * (d->T + 1));
    desc.costset_begin = costset_begin.data(); desc.costs = costs.data();
    desc.pool = pool.data(); desc.node_costset = node_costset.data();
    if (d->n_contacts > 0 && d->contacts && d->costset_contact) {
      contacts.assign(d->contacts, d->contacts + d->n_contacts);
      costset_contact.assign(d->costset_contact, d->costset_contact + d->n_costsets);
      desc.contacts = contacts.data(); desc.costset_contact = costset_contact.data();
    } else {
      desc.n_contacts = 0; desc.contacts = nullptr; desc.costset_contact = nullptr;
    }
    m.init(&desc);
    T = d->T;
    const int nx = m.nx, ndx = m.ndx, nu = m.nu;
    x0.assign(nx, 0); state_zero(m, x0.data());
    xs.assign((size_t)(T + 1) * nx, 0); xs_try = xs;
    us.assign((size_t)T * nu, 0); us_try = us; us_squash = us;
    fs.assign((size_t)(T + 1) * ndx, 0); dx = fs; Vx = fs;
    K.assign((size_t)T * nu * ndx, 0); k.assign((size_t)T * nu, 0); Qu = k; Quuk = k;
    Vxx.assign((size_t)(T + 1) * ndx * ndx, 0);
    tiles.assign((size_t)(T + 1) * m.tile, 0);
    work.resize(T + 1); evals.resize(T + 1);
    for (int n = 0; n < EMPC_N_ALPHAS; ++n) alphas[n] = 1.0 / std::pow(2.0, (double)n);
    set_candidate(nullptr, nullptr, false);
  }
  int costset_of(int t) const { return node_costset[(size_t)node_map * (T + 1) + t]; }
  SolverCtx ctx() const { return SolverCtx{smooth_model, P.barrier_weight}; }

  // SolverAbstract::setCandidate (crocoddyl/core/solver-base.cpp)
  void set_candidate(const double* xs_in, const double* us_in, bool feasible) {
    if (xs_in) std::memcpy(xs.data(), xs_in, sizeof(double) * xs.size());
    else for (int t = 0; t <= T; ++t) state_zero(m, &xs[(size_t)t * m.nx]);
    if (us_in) std::memcpy(us.data(), us_in, sizeof(double) * us.size());
    else std::fill(us.begin(), us.end(), 0.0);
    is_feasible = feasible;
  }

  // ShootingProblem::calc
  void problem_calc() {
    for (int t = 0; t <= T; ++t)
      node_calc(m, ctx(), costset_of(t), &xs[(size_t)t * m.nx], t < T ? &us[(size_t)t * m.nu] : nullptr, work[t], &evals[t]);
  }
  // ShootingProblem::calcDiff (returns the summed cost left by the last calc on each data)
  double problem_calc_diff() {
    double c = 0;
    for (int t = 0; t <= T; ++t) {
      node_calc_diff(m, ctx(), costset_of(t), work[t], evals[t], &tiles[(size_t)t * m.tile]);
      c += work[t].cost;
    }
    return c;
  }
  // SolverDDP::calcDiff
  void calc_diff() {
    if (iter == 0) problem_calc();
    cost = problem_calc_diff();
    const int ndx = m.ndx;
    if (!is_feasible) {
      state_diff(m, &xs[0], x0.data(), &fs[0]);
      bool could = true;
      for (int i = 0; i < ndx; ++i) if (std::fabs(fs[i]) >= P.th_gaptol) could = false;
      for (int t = 0; t < T; ++t) {
        state_diff(m, &xs[(size_t)(t + 1) * m.nx], work[t].xnext, &fs[(size_t)(t + 1) * ndx]);
        if (could)
          for (int i = 0; i < ndx; ++i) if (std::fabs(fs[(size_t)(t + 1) * ndx + i]) >= P.th_gaptol) could = false;
      }
      is_feasible = could;
    } else if (!was_feasible) {
      std::fill(fs.begin(), fs.end(), 0.0);
    }
  }

  // crocoddyl::BoxQP::solve (crocoddyl/core/solvers/box-qp.cpp; Tassa's projected Newton): minimise 1/2 x'Hx + q'x over
  // lb <= x <= ub from the clamped warm start.  Per iteration: gradient g = q + Hx; a coordinate is CLAMPED when it sits on a
  // bound with the gradient pushing outwards, FREE otherwise; converged when ||g||_inf <= th_grad or nothing is free; Newton
  // step on the free block (LLT of Hff [+ reg]), dxf = Hff^-1 (-qf - Hfc xc) - xf; projected line search over alpha = 1, 1/2,
  // ... 1/512 with the Armijo test  f(x) - f(xnew) > th_acceptstep g'(x - xnew).  Leaves x, the free / clamped index sets of
  // the last iteration and Hff^-1 (of the last factorised free block) in qp_*; false = LLT failure ("backward_error").
  std::vector<int> qp_free, qp_clamped;
  std::vector<double> qp_Hff_inv, qp_x;
  int qp_iters = 0;
  // statistics of the box QPs since the last reset (test diagnostics): calls, calls whose LAST allowed iteration still moved
  // the iterate by more than 1e-12 (the projected Newton iteration had not settled within maxiter), iterations until it
  // settled summed over the calls, calls that clamped something
  double qp_stats[4] = {0, 0, 0, 0};
  bool box_qp(const double* H, const double* q, const double* lb, const double* ub, const double* xinit, int nx) {
    std::vector<double> x(nx), g(nx), dx(nx), xnew(nx), Hff, Hfc, qf, xf, xc, dxf, Lf;
    for (int i = 0; i < nx; ++i) x[i] = std::max(std::min(xinit[i], ub[i]), lb[i]);
    auto value = [&](const std::vector<double>& v) {
      double a = 0, b = 0;
      for (int i = 0; i < nx; ++i) {
        double r = 0;
        for (int j = 0; j < nx; ++j) r += H[(size_t)i * nx + j] * v[j];
        a += v[i] * r; b += q[i] * v[i];
      }
      return 0.5 * a + b;
    };
    auto factor_free = [&]() {  // Hff (+ reg) = L L', Hff_inv = (L L')^-1
      const int nf = (int)qp_free.size();
      Hff.assign((size_t)nf * nf, 0.0);
      for (int i = 0; i < nf; ++i)
        for (int j = 0; j < nf; ++j) Hff[(size_t)i * nf + j] = H[(size_t)qp_free[i] * nx + qp_free[j]];
      if (P.boxqp_reg != 0.0) for (int i = 0; i < nf; ++i) Hff[(size_t)i * nf + i] += P.boxqp_reg;
      Lf = Hff;
      if (nf > 0 && !llt_inplace(Lf.data(), nf)) return false;
      qp_Hff_inv.assign((size_t)nf * nf, 0.0);
      for (int i = 0; i < nf; ++i) qp_Hff_inv[(size_t)i * nf + i] = 1.0;
      if (nf > 0) llt_solve(Lf.data(), nf, qp_Hff_inv.data(), nf);
      return true;
    };
    qp_iters = 0;
    qp_stats[0] += 1;
    int settled_at = -1; bool last_moved = false;
    struct Tail { double* st; int* settled; bool* moved; std::vector<int>* cl; int* its;
                  ~Tail() { st[2] += (*settled >= 0 ? *settled : *its); if (*moved) st[1] += 1; if (!cl->empty()) st[3] += 1; } } tail{qp_stats, &settled_at, &last_moved, &qp_clamped, &qp_iters};
    for (int k = 0; k < P.boxqp_maxiter; ++k) {
      qp_iters = k + 1;
      last_moved = false;
      qp_free.clear(); qp_clamped.clear();
      double gmax = 0;
      for (int i = 0; i < nx; ++i) {
        double r = q[i];
        for (int j = 0; j < nx; ++j) r += H[(size_t)i * nx + j] * x[j];
        g[i] = r;
        if (std::fabs(r) > gmax) gmax = std::fabs(r);
      }
      for (int j = 0; j < nx; ++j) {
        if ((x[j] == lb[j] && g[j] > 0.0) || (x[j] == ub[j] && g[j] < 0.0)) qp_clamped.push_back(j);
        else qp_free.push_back(j);
      }
      const int nf = (int)qp_free.size(), nc = (int)qp_clamped.size();
      if (gmax <= P.boxqp_th_grad || nf == 0) {
        // the inverse of the free Hessian is still needed by the caller (upstream computes it here only for k = 0 and otherwise
        // hands out the previous iteration's; when the free set has changed size since, that would index a matrix of another
        // shape, so it is recomputed)
        if ((k == 0 || qp_Hff_inv.size() != (size_t)nf * nf) && !factor_free()) return false;
        qp_x = x;
        return true;
      }
      if (!factor_free()) return false;
      dxf.assign(nf, 0.0);
      for (int i = 0; i < nf; ++i) {
        double r = -q[qp_free[i]];
        for (int j = 0; j < nc; ++j) r -= H[(size_t)qp_free[i] * nx + qp_clamped[j]] * x[qp_clamped[j]];
        dxf[i] = r;
      }
      llt_solve(Lf.data(), nf, dxf.data(), 1);
      std::fill(dx.begin(), dx.end(), 0.0);
      for (int i = 0; i < nf; ++i) dx[qp_free[i]] = dxf[i] - x[qp_free[i]];
      const double fold = value(x);
      for (int n = 0; n < EMPC_N_ALPHAS; ++n) {
        const double a = alphas[n];
        double gd = 0;
        for (int i = 0; i < nx; ++i) { xnew[i] = std::max(std::min(x[i] + a * dx[i], ub[i]), lb[i]); gd += g[i] * (x[i] - xnew[i]); }
        const double fnew = value(xnew);
        if (fold - fnew > P.boxqp_th_acceptstep * gd) {
          double mv = 0;
          for (int i = 0; i < nx; ++i) mv = std::max(mv, std::fabs(xnew[i] - x[i]) / std::max(1.0, std::fabs(x[i])));
          if (mv > 1e-12) { last_moved = true; settled_at = -1; } else if (settled_at < 0) settled_at = k;
          x = xnew; break;
        }
      }
      if (!last_moved && settled_at < 0) settled_at = k;
    }
    qp_x = x;
    return true;
  }

  // SolverDDP::backwardPass + computeGains; returns false on "backward_error"
  bool backward_pass() {
    const int ndx = m.ndx, nu = m.nu;
    double* VxxT = &Vxx[(size_t)T * ndx * ndx];
    double* VxT = &Vx[(size_t)T * ndx];
    const double* tl = &tiles[(size_t)T * m.tile];
    std::memcpy(VxxT, tl + m.oLxx, sizeof(double) * ndx * ndx);
    std::memcpy(VxT, tl + m.oLx, sizeof(double) * ndx);
    for (int i = 0; i < ndx; ++i) VxxT[i * ndx + i] += xreg;
    if (!is_feasible)
      for (int i = 0; i < ndx; ++i) {
        double s = 0;
        for (int j = 0; j < ndx; ++j) s += VxxT[i * ndx + j] * fs[(size_t)T * ndx + j];
        VxT[i] += s;
      }
    std::vector<double> FxTV(ndx * ndx), FuTV(nu * ndx), Qxx(ndx * ndx), Qxu(ndx * nu), Quu(nu * nu), Qx(ndx), L(nu * nu);
    for (int t = T - 1; t >= 0; --t) {
      const double* tile = &tiles[(size_t)t * m.tile];
      const double *Fx = tile + m.oFx, *Fu = tile + m.oFu, *Lxx = tile + m.oLxx, *Lxu = tile + m.oLxu,
                   *Luu = tile + m.oLuu, *Lx = tile + m.oLx, *Lu = tile + m.oLu;
      const double* Vxx_p = &Vxx[(size_t)(t + 1) * ndx * ndx];
      const double* Vx_p = &Vx[(size_t)(t + 1) * ndx];
      double* Qu_t = &Qu[(size_t)t * nu];
      // FxTVxx_p = Fx^T Vxx_p ; Qxx = Lxx + FxTVxx_p Fx ; Qx = Lx + Fx^T Vx_p
      for (int i = 0; i < ndx; ++i)
        for (int j = 0; j < ndx; ++j) {
          double s = 0;
          for (int l = 0; l < ndx; ++l) s += Fx[l * ndx + i] * Vxx_p[l * ndx + j];
          FxTV[i * ndx + j] = s;
        }
      for (int i = 0; i < ndx; ++i) {
        for (int j = 0; j < ndx; ++j) {
          double s = 0;
          for (int l = 0; l < ndx; ++l) s += FxTV[i * ndx + l] * Fx[l * ndx + j];
          Qxx[i * ndx + j] = Lxx[i * ndx + j] + s;
        }
        double s = 0;
        for (int l = 0; l < ndx; ++l) s += Fx[l * ndx + i] * Vx_p[l];
        Qx[i] = Lx[i] + s;
      }
      // FuTVxx_p = Fu^T Vxx_p ; Qxu = Lxu + FxTVxx_p Fu ; Quu = Luu + FuTVxx_p Fu ; Qu = Lu + Fu^T Vx_p
      for (int i = 0; i < nu; ++i)
        for (int j = 0; j < ndx; ++j) {
          double s = 0;
          for (int l = 0; l < ndx; ++l) s += Fu[l * nu + i] * Vxx_p[l * ndx + j];
          FuTV[i * ndx + j] = s;
        }
      for (int i = 0; i < ndx; ++i)
        for (int j = 0; j < nu; ++j) {
          double s = 0;
          for (int l = 0; l < ndx; ++l) s += FxTV[i * ndx + l] * Fu[l * nu + j];
          Qxu[i * nu + j] = Lxu[i * nu + j] + s;
        }
      for (int i = 0; i < nu; ++i) {
        for (int j = 0; j < nu; ++j) {
          double s = 0;
          for (int l = 0; l < ndx; ++l) s += FuTV[i * ndx + l] * Fu[l * nu + j];
          Quu[i * nu + j] = Luu[i * nu + j] + s;
        }
        double s = 0;
        for (int l = 0; l < ndx; ++l) s += Fu[l * nu + i] * Vx_p[l];
        Qu_t[i] = Lu[i] + s;
        Quu[i * nu + i] += ureg;
      }
      double* K_t = &K[(size_t)t * nu * ndx];
      double* k_t = &k[(size_t)t * nu];
      if (P.solver_type != EMPC_SOLVER_SBFDDP && is_feasible) {
        // crocoddyl SolverBoxFDDP / SolverBoxDDP::computeGains (every model of the reference's problems has control limits:
        // src/trajectory.cpp:131-132, src/mpc-controllers/carrot-mpc.cpp:220-221; an infeasible candidate takes the plain
        // gains below): du = argmin 1/2 du' Quu du + Qu' du, u_lb - us <= du <= u_ub - us, warm-started at k_[t];
        // K = Quu_inv Qxu^T with Quu_inv the inverse of the free block (zero rows / columns for the clamped controls);
        // k = -du; the clamped entries of Qu are zeroed ("important for accounting the algorithm advancement")
        std::vector<double> du_lb(nu), du_ub(nu);
        for (int i = 0; i < nu; ++i) { du_lb[i] = desc.u_lb[i] - us[(size_t)t * nu + i]; du_ub[i] = desc.u_ub[i] - us[(size_t)t * nu + i]; }
        if (!box_qp(Quu.data(), Qu_t, du_lb.data(), du_ub.data(), k_t, nu)) return false;
        std::vector<double> Quu_inv((size_t)nu * nu, 0.0);
        const int nf = (int)qp_free.size();
        for (int i = 0; i < nf; ++i)
          for (int j = 0; j < nf; ++j) Quu_inv[(size_t)qp_free[i] * nu + qp_free[j]] = qp_Hff_inv[(size_t)i * nf + j];
        for (int i = 0; i < nu; ++i)
          for (int j = 0; j < ndx; ++j) {
            double sK = 0;
            for (int l = 0; l < nu; ++l) sK += Quu_inv[(size_t)i * nu + l] * Qxu[j * nu + l];
            K_t[i * ndx + j] = sK;
          }
        for (int i = 0; i < nu; ++i) k_t[i] = -qp_x[i];
        for (int c : qp_clamped) Qu_t[c] = 0.0;
      } else {
      // computeGains: LLT(Quu); K = Quu^-1 Qxu^T ; k = Quu^-1 Qu
      L = Quu;
      if (!llt_inplace(L.data(), nu)) return false;
      for (int i = 0; i < nu; ++i) {
        for (int j = 0; j < ndx; ++j) K_t[i * ndx + j] = Qxu[j * nu + i];
        k_t[i] = Qu_t[i];
      }
      llt_solve(L.data(), nu, K_t, ndx);
      llt_solve(L.data(), nu, k_t, 1);
      }
      // value function
      double* Vx_t = &Vx[(size_t)t * ndx];
      double* Vxx_t = &Vxx[(size_t)t * ndx * ndx];
      double* Quuk_t = &Quuk[(size_t)t * nu];
      for (int i = 0; i < nu; ++i) {
        double s = 0;
        for (int j = 0; j < nu; ++j) s += Quu[i * nu + j] * k_t[j];
        Quuk_t[i] = s;
      }
      for (int i = 0; i < ndx; ++i) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < nu; ++j) { s1 += K_t[j * ndx + i] * Quuk_t[j]; s2 += K_t[j * ndx + i] * Qu_t[j]; }
        Vx_t[i] = Qx[i] + s1 - 2 * s2;
        for (int j = 0; j < ndx; ++j) {
          double s = 0;
          for (int l = 0; l < nu; ++l) s += Qxu[i * nu + l] * K_t[l * ndx + j];
          Vxx_t[i * ndx + j] = Qxx[i * ndx + j] - s;
        }
      }
      for (int i = 0; i < ndx; ++i)
        for (int j = i; j < ndx; ++j) {
          const double a = 0.5 * (Vxx_t[i * ndx + j] + Vxx_t[j * ndx + i]);
          Vxx_t[i * ndx + j] = a; Vxx_t[j * ndx + i] = a;
        }
      for (int i = 0; i < ndx; ++i) Vxx_t[i * ndx + i] += xreg;
      if (!is_feasible) {
        const double* f = &fs[(size_t)t * ndx];
        for (int i = 0; i < ndx; ++i) {
          double s = 0;
          for (int j = 0; j < ndx; ++j) s += Vxx_t[i * ndx + j] * f[j];
          Qx[i] = s;
        }
        for (int i = 0; i < ndx; ++i) Vx_t[i] += Qx[i];
      }
      for (int i = 0; i < ndx; ++i) if (!(std::fabs(Vx_t[i]) < 1e30)) return false;  // raiseIfNaN(lpNorm<inf>)
      for (int i = 0; i < ndx * ndx; ++i) if (!(std::fabs(Vxx_t[i]) < 1e30)) return false;
    }
    return true;
  }
  void increase_reg() { xreg *= P.reg_factor; if (xreg > P.reg_max) xreg = P.reg_max; ureg = xreg; }
  void decrease_reg() { xreg /= P.reg_factor; if (xreg < P.reg_min) xreg = P.reg_min; ureg = xreg; }
  // SolverDDP::computeDirection
  bool compute_direction(bool recalc) {
    if (recalc) calc_diff();
    return backward_pass();
  }
  // SolverFDDP::updateExpectedImprovement
  void update_expected_improvement() {
    const int ndx = m.ndx, nu = m.nu;
    dg = 0; dq = 0;
    auto gapterms = [&](int t) {
      const double* f = &fs[(size_t)t * ndx];
      const double* V = &Vxx[(size_t)t * ndx * ndx];
      double s0 = 0;
      for (int i = 0; i < ndx; ++i) s0 += Vx[(size_t)t * ndx + i] * f[i];
      dg -= s0;
      double s1 = 0;
      for (int i = 0; i < ndx; ++i) {
        double r = 0;
        for (int j = 0; j < ndx; ++j) r += V[i * ndx + j] * f[j];
        s1 += f[i] * r;
      }
      dq += s1;
    };
    if (!is_feasible) gapterms(T);
    for (int t = 0; t < T; ++t) {
      double a = 0, b = 0;
      for (int i = 0; i < nu; ++i) { a += Qu[(size_t)t * nu + i] * k[(size_t)t * nu + i]; b += k[(size_t)t * nu + i] * Quuk[(size_t)t * nu + i]; }
      dg += a; dq -= b;
      if (!is_feasible) gapterms(t);
    }
  }
  // SolverFDDP::expectedImprovement
  void expected_improvement() {
    const int ndx = m.ndx;
    dv = 0;
    if (!is_feasible) {
      std::vector<double> dxl(ndx);
      auto term = [&](int t) {
        state_diff(m, &xs_try[(size_t)t * m.nx], &xs[(size_t)t * m.nx], dxl.data());
        const double* V = &Vxx[(size_t)t * ndx * ndx];
        const double* f = &fs[(size_t)t * ndx];
        double s = 0;
        for (int i = 0; i < ndx; ++i) {
          double r = 0;
          for (int j = 0; j < ndx; ++j) r += V[i * ndx + j] * dxl[j];
          s += f[i] * r;
        }
        dv -= s;
      };
      term(T);
      for (int t = 0; t < T; ++t) term(t);
    }
    d0 = dg + dv; d1 = dq - 2 * dv;
  }
  // src/sbfddp.cpp:395-408
  void expected_improvement_ddp() {
    const int nu = m.nu;
    d0 = 0; d1 = 0;
    for (int t = 0; t < T; ++t) {
      double a = 0, b = 0;
      for (int i = 0; i < nu; ++i) { a += Qu[(size_t)t * nu + i] * k[(size_t)t * nu + i]; b += k[(size_t)t * nu + i] * Quuk[(size_t)t * nu + i]; }
      d0 += a; d1 -= b;
    }
  }
  // crocoddyl::raiseIfNaN(value): isnan || isinf || value >= 1e30; has_bad is that test on the infinity norm of v
  static bool raise_if_nan(double v) { return std::isnan(v) || std::isinf(v) || v >= 1e30; }
  static bool has_bad(const double* v, int n) { for (int i = 0; i < n; ++i) if (!(std::fabs(v[i]) < 1e30)) return true; return false; }

  // one node of a rollout: us_try = us - alpha k - K dx ; calc ; returns false on NaN ("forward_error")
  bool rollout_node(int t, double alpha, double* xnext) {
    const int nx = m.nx, ndx = m.ndx, nu = m.nu;
    double* dxt = &dx[(size_t)t * ndx];
    state_diff(m, &xs[(size_t)t * nx], &xs_try[(size_t)t * nx], dxt);
    for (int i = 0; i < nu; ++i) {
      double kd = 0;
      for (int j = 0; j < ndx; ++j) kd += K[((size_t)t * nu + i) * ndx + j] * dxt[j];
      double ut = us[(size_t)t * nu + i] - k[(size_t)t * nu + i] * alpha - kd;
      // SolverBoxFDDP / SolverBoxDDP::forwardPass: us_try = us_try.cwiseMax(u_lb).cwiseMin(u_ub)
      if (P.solver_type != EMPC_SOLVER_SBFDDP) ut = std::min(std::max(ut, desc.u_lb[i]), desc.u_ub[i]);
      us_try[(size_t)t * nu + i] = ut;
    }
    node_calc(m, ctx(), costset_of(t), &xs_try[(size_t)t * nx], &us_try[(size_t)t * nu], work[t], &evals[t]);
    std::memcpy(xnext, work[t].xnext, sizeof(double) * nx);
    cost_try += work[t].cost;
    if (raise_if_nan(cost_try)) return false;
    if (has_bad(xnext, nx)) return false;
    return true;
  }
  // SolverFDDP::forwardPass
  bool forward_pass(double alpha) {
    const int nx = m.nx, ndx = m.ndx;
    cost_try = 0;
    std::vector<double> xnext(x0), gap(ndx);
    const bool plain = is_feasible || alpha == 1;
    for (int t = 0; t < T; ++t) {
      if (plain) std::memcpy(&xs_try[(size_t)t * nx], xnext.data(), sizeof(double) * nx);
      else {
        for (int i = 0; i < ndx; ++i) gap[i] = fs[(size_t)t * ndx + i] * (alpha - 1);
        state_integrate(m, xnext.data(), gap.data(), &xs_try[(size_t)t * nx]);
      }
      if (!rollout_node(t, alpha, xnext.data())) return false;
    }
    if (plain) std::memcpy(&xs_try[(size_t)T * nx], xnext.data(), sizeof(double) * nx);
    else {
      for (int i = 0; i < ndx; ++i) gap[i] = fs[(size_t)T * ndx + i] * (alpha - 1);
      state_integrate(m, xnext.data(), gap.data(), &xs_try[(size_t)T * nx]);
    }
    node_calc(m, ctx(), costset_of(T), &xs_try[(size_t)T * nx], nullptr, work[T], &evals[T]);
    cost_try += work[T].cost;
    return !raise_if_nan(cost_try);
  }
  // SolverSbFDDP::forwardPassDDP (src/sbfddp.cpp:416-460): classical rollout from xs_try[0] as left by earlier calls
  bool forward_pass_ddp(double alpha) {
    const int nx = m.nx;
    cost_try = 0;
    std::vector<double> xnext(nx);
    for (int t = 0; t < T; ++t) {
      if (!rollout_node(t, alpha, xnext.data())) return false;
      std::memcpy(&xs_try[(size_t)(t + 1) * nx], xnext.data(), sizeof(double) * nx);
    }
    node_calc(m, ctx(), costset_of(T), &xs_try[(size_t)T * nx], nullptr, work[T], &evals[T]);
    cost_try += work[T].cost;
    return !raise_if_nan(cost_try);
  }
  void accept_candidate(bool feasible) {
    xs = xs_try; us = us_try; is_feasible = feasible;
  }

  // ---- fork-defined stop rules: INFERRED, kept in one place ----
  double gap_norm() const {
    if (is_feasible) return 0.0;  // gaps are closed (and zeroed at the next calcDiff)
    double n = 0;
    for (size_t i = 0; i < fs.size(); ++i) {
      const double a = std::fabs(fs[i]);
      if (P.stop_gap_norm == 0) { if (a > n) n = a; } else n += a;
    }
    return n;
  }
  void stopping_criteria() {
    if (P.stop_criteria == EMPC_STOP_CRITERIA_QU_NORM) {  // upstream SolverDDP::stoppingCriteria: sum_t ||Qu_t||^2
      stop = 0;
      for (int t = 0; t < T; ++t) {
        double s = 0;
        for (int i = 0; i < m.nu; ++i) s += Qu[(size_t)t * m.nu + i] * Qu[(size_t)t * m.nu + i];
        stop += s;
      }
    } else {
      stop = std::fabs(cost_prev - cost);  // StopCriteriaCostReduction (fork, inferred)
    }
  }
  bool stopping_test_feasible() const { return was_feasible && stop < th_stop; }
  bool stopping_test() const {
    if (P.stop_test == EMPC_STOP_TEST_FEASIBLE) return stopping_test_feasible();  // upstream SolverFDDP::solve
    return stop < th_stop && gap_norm() < P.th_stop_gaps;                         // StopTestGaps (fork, inferred)
  }

  // the point where the reference runs its callbacks (src/sbfddp.cpp:303-307, :381-385): one record per iteration
  std::vector<empc_iter_record_t> log;
  int accepted_index = -1;
  void record(int phase) {
    empc_iter_record_t r;
    r.iter = iter; r.total_iter = (int)total_iters + iter; r.phase = phase; r.accepted = accepted_index;
    r.is_feasible = is_feasible ? 1 : 0; r.reserved = 0;
    r.cost = cost; r.stop = stop; r.steplength = steplength; r.xreg = xreg; r.d0 = d0; r.d1 = d1; r.smooth = smooth_model;
    log.push_back(r);
  }
  void trace(const char* ph) const {
    static const char* on = std::getenv("ORC_TRACE");
    if (!on) return;
    std::printf("[orc %s] it=%d cost=%.15e prev=%.15e step=%g xreg=%g feas=%d wasf=%d stop=%.6e gap=%.6e dV=%.6e dVexp=%.6e d0=%.6e d1=%.6e smooth=%g\n",
                ph, iter, cost, cost_prev, steplength, xreg, (int)is_feasible, (int)was_feasible, stop, gap_norm(), dV, dVexp, d0, d1, smooth_model);
  }
  // src/sbfddp.cpp:228-315
  bool solve_fddp(int maxiter, bool feasible_arg, double reginit) {
    is_feasible = feasible_arg;
    xreg = std::isnan(reginit) ? P.reg_min : reginit; ureg = xreg;
    was_feasible = false;
    bool recalc = true;
    for (iter = 0; iter < maxiter; ++iter) {
      while (true) {
        if (!compute_direction(recalc)) {
          recalc = false;
          increase_reg();
          if (xreg == P.reg_max) return false;
          continue;
        }
        break;
      }
      update_expected_improvement();
      recalc = false;
      accepted_index = -1;
      for (int n = 0; n < EMPC_N_ALPHAS; ++n) {
        steplength = alphas[n];
        if (!forward_pass(steplength)) continue;
        dV = cost - cost_try;
        expected_improvement();
        dVexp = steplength * (d0 + 0.5 * steplength * d1);
        if (dVexp >= 0) {
          if (d0 < P.th_grad || dV > P.th_acceptstep * dVexp) {
            was_feasible = is_feasible;
            accept_candidate(was_feasible || steplength == 1);
            cost_prev = cost; cost = cost_try; recalc = true; accepted_index = n;
            break;
          }
        } else {
          if (dV > P.th_acceptnegstep * dVexp) {
            was_feasible = is_feasible;
            accept_candidate(was_feasible || steplength == 1);
            cost_prev = cost; cost = cost_try; recalc = true; accepted_index = n;
            break;
          }
        }
      }
      if (steplength > P.th_stepdec) decrease_reg();
      if (steplength <= P.th_stepinc) {
        increase_reg();
        if (xreg == P.reg_max) return false;
      }
      stopping_criteria();
      record(0);
      trace("fddp");
      if (stopping_test()) return true;
    }
    iter = iter >= maxiter ? maxiter - 1 : iter;
    return false;
  }
  // src/sbfddp.cpp:317-393
  bool solve_ddp(int maxiter, double reginit) {
    xreg = std::isnan(reginit) ? P.reg_min : reginit; ureg = xreg;
    was_feasible = false;
    bool recalc = true;
    for (iter = 0; iter < maxiter; ++iter) {
      while (true) {
        if (!compute_direction(recalc)) {
          recalc = false;
          increase_reg();
          if (xreg == P.reg_max) return false;
          continue;
        }
        break;
      }
      expected_improvement_ddp();
      recalc = false;
      accepted_index = -1;
      for (int n = 0; n < EMPC_N_ALPHAS; ++n) {
        steplength = alphas[n];
        if (!forward_pass_ddp(steplength)) continue;
        dV = cost - cost_try;
        dVexp = steplength * (d0 + 0.5 * steplength * d1);
        if (dVexp >= 0) {
          if (d0 < P.th_grad || !is_feasible || dV > P.th_acceptstep * dVexp) {
            was_feasible = is_feasible;
            accept_candidate(true);
            cost_prev = cost; cost = cost_try; recalc = true; accepted_index = n;
            break;
          }
        }
      }
      if (steplength > P.th_stepdec) decrease_reg();
      if (steplength <= P.th_stepinc) {
        increase_reg();
        if (xreg == P.reg_max) return false;
      }
      stopping_criteria();
      record(1);
      trace("ddp");
      if (stopping_test_feasible()) return true;
    }
    iter = iter >= maxiter ? maxiter - 1 : iter;
    return false;
  }
  // src/sbfddp.cpp:192-226
  void solve(const double* xs_in, const double* us_in, int maxiter, bool feasible_arg) {
    std::memcpy(&xs_try[0], x0.data(), sizeof(double) * m.nx);
    set_candidate(xs_in, us_in, feasible_arg);
    if (P.solver_type != EMPC_SOLVER_SBFDDP) {
      // crocoddyl::SolverBoxFDDP / SolverBoxDDP (src/mpc-controllers/carrot-mpc.cpp:236-241): one upstream SolverFDDP::solve or
      // SolverDDP::solve with th_stop_ = 5e-5, the caller's feasibility flag, box gains and clamped rollouts
      smooth = smooth_model = P.smooth_init;
      th_stop = P.th_stop;
      total_iters = 0;
      log.clear();
      if (P.solver_type == EMPC_SOLVER_BOXFDDP) solve_fddp(maxiter, feasible_arg, P.reg_init);
      else solve_ddp(maxiter, P.reg_init);
      total_iters = iter + 1;
      us_squash = us;  // no squashing function on this path
      return;
    }
    smooth = P.smooth_init;
    convergence = P.convergence_init;
    total_iters = 0;
    log.clear();
    while (convergence >= P.convergence_stop) {
      smooth_model = smooth;  // squashingUpdate() + barrierUpdate() (src/sbfddp.cpp:206-207,462-477)
      th_stop = convergence;
      solve_fddp(maxiter, false, P.reg_init);
      smooth *= P.smooth_mult;
      convergence *= P.convergence_mult;
      total_iters += iter + 1;
    }
    if (!is_feasible) {
      solve_ddp(maxiter, P.reg_init);
      total_iters += iter + 1;
    }
    iter = (int)total_iters - 1;
    // fillSquashedOutputs.  The node data were last evaluated with the smoothing of the final pass (set_smooth is not
    // called again after the loop; the DDP clean-up phase therefore also runs with it).  Default: s(us[t]);
    // squash_quirk: whatever the last calc on node t left.
    const double last_smooth = smooth_model;
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < m.nu; ++i) {
        if (P.squash_quirk) { us_squash[(size_t)t * m.nu + i] = work[t].s[i]; continue; }
        const double lb = desc.u_lb[i], ub = desc.u_ub[i], dd = (ub - lb) * last_smooth, a = dd * dd;
        const double u = us[(size_t)t * m.nu + i], l = u - lb, h = u - ub;
        us_squash[(size_t)t * m.nu + i] = 0.5 * (std::sqrt(l * l + a) - std::sqrt(h * h + a) + lb + ub);
      }
  }
};

}  // namespace orc

using orc::Solver;

extern "C" {

void orc_default_params(empc_solver_params_t* p) {
  std::memset(p, 0, sizeof(*p));
  p->maxiter = 100; p->stop_gap_norm = 0; p->squash_quirk = 0;
  p->convergence_init = 1e-2; p->convergence_stop = 1e-3; p->convergence_mult = 1e-1;
  p->smooth_init = 0.1; p->smooth_mult = 0.5; p->barrier_weight = 1e-3;
  p->reg_init = 1e-9; p->reg_min = 1e-9; p->reg_max = 1e9; p->reg_factor = 10;
  p->th_acceptstep = 0.1; p->th_acceptnegstep = 2; p->th_grad = 1e-12; p->th_gaptol = 1e-16;
  p->th_stepdec = 0.5; p->th_stepinc = 0.01; p->th_stop_gaps = 1.0;
  p->solver_type = EMPC_SOLVER_SBFDDP;
  p->th_stop = 5e-5; p->boxqp_maxiter = 100; p->boxqp_th_acceptstep = 0.1; p->boxqp_th_grad = 1e-5; p->boxqp_reg = 0.0;
}
void orc_box_params(empc_solver_params_t* p, int32_t solver_type) {
  orc_default_params(p);
  p->solver_type = solver_type;
  p->stop_criteria = EMPC_STOP_CRITERIA_QU_NORM; p->stop_test = EMPC_STOP_TEST_FEASIBLE;
}

void* orc_create(const empc_problem_desc_t* d) {
  Solver* s = new Solver();
  orc_default_params(&s->P);
  s->init(d);
  return s;
}
void orc_destroy(void* h) { delete (Solver*)h; }
void orc_set_params(void* h, const empc_solver_params_t* p) { ((Solver*)h)->P = *p; }
void orc_set_node_map(void* h, int map) { ((Solver*)h)->node_map = map; }
void orc_set_x0(void* h, const double* x0) { Solver* s = (Solver*)h; std::memcpy(s->x0.data(), x0, sizeof(double) * s->m.nx); }
void orc_set_candidate(void* h, const double* xs, const double* us, int feasible) { ((Solver*)h)->set_candidate(xs, us, feasible != 0); }
void orc_update_costs(void* h, int first, int n, const empc_cost_t* costs, int pool_off, int n_pool, const double* pool) {
  Solver* s = (Solver*)h;
  for (int i = 0; i < n; ++i) s->costs[first + i] = costs[i];
  for (int i = 0; i < n_pool; ++i) s->pool[pool_off + i] = pool[i];
}
void orc_update_node_costsets(void* h, const int32_t* nc) {
  Solver* s = (Solver*)h;
  std::memcpy(s->node_costset.data(), nc, sizeof(int32_t) * s->node_costset.size());
}
void orc_dims(void* h, int32_t* out /* nq nv nx ndx nu T tile */) {
  Solver* s = (Solver*)h;
  out[0] = s->m.nq; out[1] = s->m.nv; out[2] = s->m.nx; out[3] = s->m.ndx; out[4] = s->m.nu; out[5] = s->T; out[6] = s->m.tile;
}

// full solve; outputs may be NULL
void orc_solve(void* h, const double* xs_in, const double* us_in, int feasible) {
  Solver* s = (Solver*)h;
  s->solve(xs_in, us_in, s->P.maxiter, feasible != 0);
}

// ---- phase hooks mirroring empc_phase_* ----
void orc_phase_calc_diff(void* h, double smooth) {
  Solver* s = (Solver*)h;
  s->smooth = s->smooth_model = smooth; s->iter = 0;
  s->calc_diff();
}
int orc_phase_backward(void* h, double xreg, int feasible) {
  Solver* s = (Solver*)h;
  s->xreg = s->ureg = xreg; s->is_feasible = feasible != 0;
  const bool ok = s->backward_pass();
  if (ok) s->update_expected_improvement();
  return ok ? 1 : 0;
}
int orc_phase_rollout(void* h, double smooth, int feasible, int ddp, int alpha_index) {
  Solver* s = (Solver*)h;
  s->smooth = s->smooth_model = smooth; s->is_feasible = feasible != 0;
  const double a = s->alphas[alpha_index];
  bool ok;
  if (ddp) ok = s->forward_pass_ddp(a);
  else { ok = s->forward_pass(a); if (ok) s->expected_improvement(); }
  return ok ? 1 : 0;
}
void orc_set_xs_try0(void* h, const double* x) { Solver* s = (Solver*)h; std::memcpy(&s->xs_try[0], x, sizeof(double) * s->m.nx); }

// iteration records of the last solve (what a CallbackVerbose would have seen), oldest first
int orc_get_iteration_log(void* h, empc_iter_record_t* out, int max_records) {
  Solver* s = (Solver*)h;
  const int n = std::min((int)s->log.size(), max_records);
  for (int i = 0; i < n; ++i) out[i] = s->log[i];
  return (int)s->log.size();
}

static void copy_out(const std::vector<double>& v, double* out) { std::memcpy(out, v.data(), sizeof(double) * v.size()); }
int orc_get(void* h, const char* name, double* out) {
  Solver* s = (Solver*)h;
  const std::string n(name);
  if (n == "xs") copy_out(s->xs, out);
  else if (n == "us") copy_out(s->us, out);
  else if (n == "xs_try") copy_out(s->xs_try, out);
  else if (n == "us_try") copy_out(s->us_try, out);
  else if (n == "us_squash") copy_out(s->us_squash, out);
  else if (n == "K") copy_out(s->K, out);
  else if (n == "k") copy_out(s->k, out);
  else if (n == "Qu") copy_out(s->Qu, out);
  else if (n == "qp_stats") { std::memcpy(out, s->qp_stats, sizeof(s->qp_stats)); std::fill(s->qp_stats, s->qp_stats + 4, 0.0); }
  else if (n == "Vx") copy_out(s->Vx, out);
  else if (n == "Vxx") copy_out(s->Vxx, out);
  else if (n == "fs") copy_out(s->fs, out);
  else if (n == "tiles") copy_out(s->tiles, out);
  else if (n == "xnext") { for (int t = 0; t <= s->T; ++t) std::memcpy(out + (size_t)t * s->m.nx, s->work[t].xnext, sizeof(double) * s->m.nx); }
  else if (n == "node_cost") { for (int t = 0; t <= s->T; ++t) out[t] = s->work[t].cost; }
  else if (n == "cost") out[0] = s->cost;
  else if (n == "cost_try") out[0] = s->cost_try;
  else if (n == "stop") out[0] = s->stop;
  else if (n == "xreg") out[0] = s->xreg;
  else if (n == "dgdq") { out[0] = s->dg; out[1] = s->dq; }
  else if (n == "dv") out[0] = s->dv;
  else if (n == "iter") out[0] = (double)s->iter;
  else if (n == "feasible") out[0] = s->is_feasible ? 1.0 : 0.0;
  else return 1;
  return 0;
}

// Solves `n` OCPs of the same problem (different x0, zero initial guess) on `nthreads` host threads, one OCP per
// thread at a time — the execution model of the reference (single-threaded solver) replicated across cores.
// Returns wall seconds; iters_out[i] = iterations executed by OCP i (iter_+1), cost_out[i] its final cost.
double orc_solve_batch(const empc_problem_desc_t* d, const empc_solver_params_t* p, const double* x0, int n, int nthreads,
                       int32_t* iters_out, double* cost_out) {
  if (nthreads < 1) nthreads = 1;
  std::vector<Solver*> solvers(nthreads);
  for (int t = 0; t < nthreads; ++t) { solvers[t] = new Solver(); solvers[t]->P = *p; solvers[t]->init(d); }
  const int nx = solvers[0]->m.nx;
  std::atomic<int> next(0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      Solver* s = solvers[t];
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        std::memcpy(s->x0.data(), x0 + (size_t)i * nx, sizeof(double) * nx);
        s->solve(nullptr, nullptr, s->P.maxiter, false);
        if (iters_out) iters_out[i] = s->iter + 1;
        if (cost_out) cost_out[i] = s->cost;
      }
    });
  for (auto& x : th) x.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (Solver* s : solvers) delete s;
  return sec;
}

// RK4 plant (bindings/python/eagle_mpc/utils/simulator.py:7-29; crocoddyl 1.x IntegratedActionModelRK4::calc) with the
// plain multicopter actuation: tau = [tau_f u_rotors ; u_arm]
void orc_plant_step(void* h, const double* x, const double* u, double dt, double* xnext) {
  Solver* s = (Solver*)h;
  const orc::Model& m = s->m;
  double tau[orc::MAXV];
  for (int i = 0; i < 6; ++i) {
    double t = 0;
    for (int j = 0; j < m.nr; ++j) t += s->desc.tau_f[i * m.nr + j] * u[j];
    tau[i] = t;
  }
  for (int i = 0; i < m.na; ++i) tau[6 + i] = u[m.nr + i];
  const double c[4] = {0.0, 0.5, 0.5, 1.0}, wgt[4] = {1.0, 2.0, 2.0, 1.0};
  double ksum[orc::MAXDX] = {0}, kprev[orc::MAXDX] = {0}, y[orc::MAXX], dxs[orc::MAXDX];
  for (int st = 0; st < 4; ++st) {
    for (int i = 0; i < m.ndx; ++i) dxs[i] = dt * c[st] * kprev[i];
    if (st == 0) std::memcpy(y, x, sizeof(double) * m.nx);
    else orc::state_integrate(m, x, dxs, y);
    orc::Work w;
    orc::aba(m, y, y + m.nq, tau, w);
    for (int i = 0; i < m.nv; ++i) { kprev[i] = y[m.nq + i]; kprev[m.nv + i] = w.a[i]; }
    for (int i = 0; i < m.ndx; ++i) ksum[i] += wgt[st] * kprev[i];
  }
  double dx[orc::MAXDX];
  for (int i = 0; i < m.ndx; ++i) dx[i] = ksum[i] * (dt / 6.0);
  orc::state_integrate(m, x, dx, xnext);
}

// ---- math unit-test exports ----
void orc_exp6(const double* nu, double* R, double* p) { orc::SE3 M; orc::exp6(nu, M); std::memcpy(R, M.R, 72); std::memcpy(p, M.p, 24); }
void orc_log6(const double* R, const double* p, double* nu) { orc::SE3 M; std::memcpy(M.R, R, 72); std::memcpy(M.p, p, 24); orc::log6(M, nu); }
void orc_Jlog6(const double* R, const double* p, double* J) { orc::SE3 M; std::memcpy(M.R, R, 72); std::memcpy(M.p, p, 24); orc::Jlog6(M, J); }
void orc_Jexp6(const double* nu, double* J) { orc::Jexp6(nu, J); }
void orc_exp3(const double* w, double* R) { orc::exp3(w, R); }
void orc_log3(const double* R, double* w) { double t; orc::log3(R, w, t); }
void orc_Jlog3(const double* R, double* J) { double w[3], t; orc::log3(R, w, t); orc::Jlog3(t, w, J); }
void orc_quat_to_R(const double* q, double* R) { orc::quat_to_R(q, R); }
void orc_R_to_quat(const double* R, double* q) { orc::R_to_quat(R, q); }
void orc_state_integrate(void* h, const double* x, const double* dx, double* out) { orc::state_integrate(((Solver*)h)->m, x, dx, out); }
void orc_state_diff(void* h, const double* x0, const double* x1, double* dx) { orc::state_diff(((Solver*)h)->m, x0, x1, dx); }
void orc_aba(void* h, const double* q, const double* v, const double* tau, double* a) {
  Solver* s = (Solver*)h; orc::Work w; orc::aba(s->m, q, v, tau, w); std::memcpy(a, w.a, sizeof(double) * s->m.nv);
}
void orc_rnea(void* h, const double* q, const double* v, const double* a, double* tau) { orc::rnea(((Solver*)h)->m, q, v, a, tau); }
// single-node evaluation: out_tile may be NULL (calc only).  u may be NULL (terminal convention u = 0).
void orc_node_eval(void* h, int costset, double smooth, const double* x, const double* u, double* xnext, double* cost,
                   double* s_out, double* tile) {
  Solver* s = (Solver*)h;
  orc::Work w; std::vector<orc::CostEval> ev;
  orc::SolverCtx ctx{smooth, s->P.barrier_weight};
  orc::node_calc(s->m, ctx, costset, x, u, w, &ev);
  std::memcpy(xnext, w.xnext, sizeof(double) * s->m.nx);
  *cost = w.cost;
  if (s_out) std::memcpy(s_out, w.s, sizeof(double) * s->m.nu);
  if (tile) orc::node_calc_diff(s->m, ctx, costset, w, ev, tile);
}
// acceleration and its partials (nv, nv*nv, nv*nv, nv*nv)
void orc_aba_derivatives(void* h, const double* q, const double* v, const double* tau, double* a, double* a_q, double* a_v, double* Minv) {
  Solver* s = (Solver*)h; orc::Work w;
  orc::aba(s->m, q, v, tau, w);
  orc::aba_derivatives(s->m, w, a_q, a_v);
  std::memcpy(a, w.a, sizeof(double) * s->m.nv);
  std::memcpy(Minv, w.Minv, sizeof(double) * s->m.nv * s->m.nv);
}

}  // extern "C"
