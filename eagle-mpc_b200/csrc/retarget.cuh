// retarget.cuh — batched MPC instances: one private copy of the cost tables per instance and device-side retargeting of
// the rail reference.  The reference retargets one problem on the host before every solve
// (src/mpc-controllers/rail-mpc.cpp:154-200: a loop over the knots with string-keyed map lookups); with thousands of
// instances at different controller times that loop becomes the bottleneck, so here a kernel writes the per-node state
// references of all instances at once.
#pragma once
#include "kernels.cuh"

namespace empc {

// n private copies of the (single node map) cost tables: cost records with their pool offsets shifted, the pool, the
// cost-set boundaries and the node -> cost set map.
__global__ void replicate_tables_kernel(const empc_cost_t* costs, int n_costs, const double* pool, int n_pool, const int* begin,
                                        int n_sets, const int* node_set, int T1, int n, empc_cost_t* o_costs, double* o_pool,
                                        int* o_begin, int* o_node_set) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = gridDim.x * (size_t)blockDim.x;
  for (size_t e = tid; e < (size_t)n * n_costs; e += nth) {
    const int m = (int)(e / n_costs);
    empc_cost_t cs = costs[e - (size_t)m * n_costs];
    const int sh = m * n_pool;
    if (cs.ref_off >= 0) cs.ref_off += sh;
    if (cs.w_off >= 0) cs.w_off += sh;
    if (cs.lb_off >= 0) cs.lb_off += sh;
    if (cs.ub_off >= 0) cs.ub_off += sh;
    o_costs[e] = cs;
  }
  for (size_t e = tid; e < (size_t)n * n_pool; e += nth) o_pool[e] = pool[e % n_pool];
  for (size_t e = tid; e < (size_t)n * n_sets; e += nth) { const int m = (int)(e / n_sets); o_begin[e] = begin[e - (size_t)m * n_sets] + m * n_costs; }
  if (tid == 0) o_begin[(size_t)n * n_sets] = n * n_costs;
  for (size_t e = tid; e < (size_t)n * T1; e += nth) { const int m = (int)(e / T1); o_node_set[e] = node_set[e - (size_t)m * T1] + m * n_sets; }
}

// RailMpc::updateProblem for every instance (rail-mpc.cpp:154-200), thread per (instance, knot): node_time = time_m +
// i dt; idx_state = upper_bound(t_ref, node_time) with t_ref[j] = j dt_ref, i.e. node_time / dt_ref + 1; past the end of
// the reference the knot tracks the hover state (row n_ref of the table, built by empc_set_reference_trajectory),
// otherwise state_ref[idx_state - 1] (the interpolation factor is evaluated in integer arithmetic and is always 0,
// :188-189).  The reference of the knot's state cost ("rail_state", the only CostModelState of a rail knot) is rewritten.
__global__ void rail_retarget_kernel(const empc_cost_t* costs, double* pool, const int* begin, const int* node_set, int T1,
                                     int n_maps, const long long* times, int dt_node_ms, const double* ref_table, int n_ref,
                                     int dt_ref_ms, int nx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_maps * T1) return;
  const int m = idx / T1, i = idx - m * T1;
  const long long node_time = times[m] + (long long)i * dt_node_ms;
  const long long q = node_time / dt_ref_ms;
  const int row = (q + 1 >= n_ref) ? n_ref : (int)q;
  const int set = node_set[idx];
  for (int c = begin[set]; c < begin[set + 1]; ++c) {
    const empc_cost_t cs = costs[c];
    if (cs.type != EMPC_COST_STATE || cs.ref_off < 0) continue;
    for (int k = 0; k < nx; ++k) pool[cs.ref_off + k] = ref_table[(size_t)row * nx + k];
    break;
  }
}

// CarrotMpc::updateProblem for every instance (carrot-mpc.cpp:298-401), thread per (instance, knot).  The knot's cost set
// in name order is [barrier,] carrot_state, carrot_tail, control_reg, state_limits, state_reg: the first two
// CostModelState records are the carrot and its tail.  Row n_ref + 1 of the reference table is the tail state (last
// configuration, zero velocity, :376-388).
__global__ void carrot_retarget_kernel(empc_cost_t* costs, double* pool, const int* begin, const int* node_set, int T1, int n_maps,
                                       const long long* times, int dt_node_ms, const double* ref_table, int n_ref, int dt_ref_ms,
                                       int nx, int n_stages, const long long* t_stages, const unsigned char* is_transition) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_maps * T1) return;
  const int m = idx / T1, i = idx - m * T1;
  const long long node_time = times[m] + (long long)i * dt_node_ms;
  int ub = 0;  // first entry of t_stages (n_stages + 1 entries) beyond node_time
  while (ub <= n_stages && t_stages[ub] <= node_time) ++ub;
  const int st = ub - 1;
  const long long q = node_time / dt_ref_ms;
  const double* ref = ref_table + (size_t)((q + 1 >= n_ref) ? n_ref + 1 : (int)q) * nx;
  const int set = node_set[idx];
  int c_state = -1, c_tail = -1;
  for (int c = begin[set]; c < begin[set + 1]; ++c) {
    if (costs[c].type != EMPC_COST_STATE) continue;
    if (c_state < 0) c_state = c; else { c_tail = c; break; }
  }
  if (c_state < 0 || c_tail < 0) return;
  if (st < n_stages) {
    if (!is_transition[st] || i == T1 - 1) {
      costs[c_state].active = 1;
      const int off = costs[c_state].ref_off;
      for (int k = 0; k < nx; ++k) pool[off + k] = ref[k];
    } else {
      costs[c_state].active = 0;
    }
  } else {
    costs[c_state].active = 0;
    costs[c_tail].active = 1;
    const int off = costs[c_tail].ref_off;
    for (int k = 0; k < nx; ++k) pool[off + k] = ref[k];
  }
}

// WeightedMpc::updateProblem for every instance (weighted-mpc.cpp:173-245), block per instance.  The active stage of a
// knot depends on the one of the knot before it (a stage of zero duration is not skipped, :186-192): thread 0 walks that
// chain into shared memory (stage = upper_bound(t_ini, node_time) - 1, pulled back by one when it jumped two stages), then
// the knots are rewritten in parallel: costs of the knot's stage on, its task costs re-weighted, all other costs off; the
// barrier is left alone.
struct WeightedScheduleDev {
  int n_stages, n_slots;
  const long long* t_ini; const long long* t_end;
  long long duration;
  double alpha, beta;
  const unsigned char* match; const unsigned char* task;
  const double* base;
};

__device__ inline int weighted_stage_of(const WeightedScheduleDev& s, long long t) {
  int ub = 0;  // first stage with t_ini > t
  while (ub < s.n_stages && s.t_ini[ub] <= t) ++ub;
  return ub - 1;
}

__global__ void weighted_retarget_kernel(empc_cost_t* costs, const int* begin, const int* node_set, int T1, int n_maps,
                                         const long long* times, int dt_node_ms, WeightedScheduleDev s) {
  extern __shared__ int wr_stage[];  // T1 entries
  const int m = blockIdx.x;
  const long long t0 = times[m];
  if (threadIdx.x == 0) {
    int last = weighted_stage_of(s, t0);
    for (int i = 0; i < T1; ++i) {
      int st = weighted_stage_of(s, t0 + (long long)i * dt_node_ms);
      if (st == last + 2) st -= 1;
      wr_stage[i] = st;
      last = st;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T1; i += blockDim.x) {
    const long long node_time = t0 + (long long)i * dt_node_ms;
    const int st = wr_stage[i];
    const double wt = (node_time > s.duration) ? 0.0 : (double)((int)node_time - (int)s.t_end[st]) / 1000.0;
    const double w = exp(s.alpha * wt);
    const int set = node_set[m * T1 + i];
    int slot = 0;
    for (int c = begin[set]; c < begin[set + 1]; ++c) {
      if (costs[c].type == EMPC_COST_SQUASH_BARRIER) continue;
      const int e = st * s.n_slots + slot;
      if (s.match[e]) {
        costs[c].active = 1;
        if (s.task[e]) costs[c].weight = s.base[e] * w * s.beta;
      } else {
        costs[c].active = 0;
      }
      ++slot;
    }
  }
}

}  // namespace empc
