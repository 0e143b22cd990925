// Shared device-side definitions of the adaptive-RK path: controller state,
// tableau constants, block primitives.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/eno.h"

namespace eno {

constexpr int kThreads = 512;    // threads per CTA of every row kernel
constexpr int kMaxStages = 10;   // tanyam
constexpr int kRedFloats = 2048; // per-row-of-tile scratch of the split-K gemv

// Butcher tableau in the working precision (rounded on the host exactly like
// the reference rounds its Python doubles, lib/ode.py:108-120 etc.).
struct Tableau {
  int s;             // stages
  int order;         // exponent for optimal_step_size
  int init_order;    // argument of initial_step_size
  int nfe_per_iter;  // what the reference adds to nfe per loop iteration
  int clip;          // 0: dopri5 flavour (overshoot + c_mid dense output); 1: clip flavour
  float alpha[kMaxStages];
  float beta[kMaxStages - 1][kMaxStages];
  float c_sol[kMaxStages];
  float c_err[kMaxStages];
  float c_mid[kMaxStages];
};

// Device-resident controller of one solve (lib/ode.py:827-851 carry + loop state).
struct Ctrl {
  float t, dt;          // solver time and the step size the next iteration will try
  float last_t;         // start time of the last accepted step
  float step_dt;        // dt of the last accepted step (for the interpolant)
  float ratio;          // last mean squared error ratio
  float h0;             // initial_step_size intermediate
  float d0, d1;
  float rtol, atol;
  int cur, prev, cand;  // ring indices into the 3 state buffers
  int mid_acc;          // which of the 2 mid-point buffers belongs to the accepted step
  int iters_tgt;        // i of the while loop for the current target
  int tgt;              // index into ts of the current target
  int T;
  int mxstep;           // INT_MAX when unbounded
  int done;
  int status;
  int n_iters, n_accepted;
  int nfe;
  int out_begin, out_end;  // pending dense outputs [begin, end)
  unsigned int ticket;     // last-block-done counter
  unsigned int ticket2;
  int trace_cap;
  float n_state;           // N of the mean in error_ratio (as float)
};

// The host mirror (ode.py: CTRL_DONE_I32, CTRL_NITERS_I32) reads these two words straight out of a workspace.
static_assert(offsetof(Ctrl, done) == 72 && offsetof(Ctrl, n_iters) == 80, "Ctrl layout is part of the host contract");

// Optional per-kernel device timing (bench.py's roofline leg): event pairs around the launches of
// one kernel kind, resolved after the solve's final synchronize.
enum ProfKind { PROF_FWD_STEP = 0, PROF_ADJ_ROWS = 1, PROF_ADJ_PGRAD = 2, PROF_EXCHANGE = 3, PROF_ADJ_PGEMM = 4, PROF_KINDS = 5 };
struct Profiler {
  bool enabled = false;
  static constexpr int kMax = 512;
  cudaEvent_t ev[kMax][2];
  int kind[kMax];
  int n = 0;
  bool created = false;
  double ms[PROF_KINDS] = {0, 0, 0, 0, 0};
  long long count[PROF_KINDS] = {0, 0, 0, 0, 0};
  int begin(int k, cudaStream_t st) {
    if (!enabled || n >= kMax) return -1;
    if (!created) {
      for (int i = 0; i < kMax; ++i) { cudaEventCreate(&ev[i][0]); cudaEventCreate(&ev[i][1]); }
      created = true;
    }
    kind[n] = k;
    cudaEventRecord(ev[n][0], st);
    return n++;
  }
  void end(int slot, cudaStream_t st) { if (slot >= 0) cudaEventRecord(ev[slot][1], st); }
  // call after the stream has been synchronized; drops launches that exited early (done flag)
  void resolve(int valid_per_kind[PROF_KINDS]) {
    int seen[PROF_KINDS] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const int k = kind[i];
      if (seen[k]++ < valid_per_kind[k]) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev[i][0], ev[i][1]) == cudaSuccess) { ms[k] += t; count[k] += 1; }
      }
    }
    n = 0;
  }
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum of n doubles living in global memory: fixed
// contiguous chunk per thread, fixed-shape tree.  Result valid in thread 0.
__device__ __forceinline__ double block_sum_global(const double* __restrict__ v, int n, double* sm /*[blockDim]*/) {
  const int nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int b = threadIdx.x * per;
  const int e = min(n, b + per);
  double acc = 0.0;
  for (int i = b; i < e; ++i) acc += v[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  int o = 1;
  while (o < nt) o <<= 1;
  for (o >>= 1; o > 0; o >>= 1) {
    if (threadIdx.x < o && threadIdx.x + o < nt) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  return sm[0];
}

// The same sum over n logical entries stored as records: entry i lives at v[(i / n_loc) * rec + (i % n_loc)]
// (one record per rank, see adj_host.cuh).  Chunking and tree are those of block_sum_global, so a packed
// layout gives bit-identical sums.
template <typename T>
__device__ __forceinline__ double block_sum_records(const T* __restrict__ v, int n, int n_loc, int rec, double* sm) {
  const int nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int b = threadIdx.x * per;
  const int e = min(n, b + per);
  double acc = 0.0;
  for (int i = b; i < e; ++i) {
    const int r = i / n_loc;
    acc += (double)v[(size_t)r * rec + (i - r * n_loc)];
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  int o = 1;
  while (o < nt) o <<= 1;
  for (o >>= 1; o > 0; o >>= 1) {
    if (threadIdx.x < o && threadIdx.x + o < nt) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  return sm[0];
}

// One WARP sums n logical entries stored as records (entry i at v[(i / n_loc) * rec + i % n_loc]; n_loc = n, rec = 0
// for a contiguous array) in double, in an order that depends on n alone: lane-strided accumulation (loads issued
// eight at a time), then an xor tree.  Several warps of a finalising CTA run different sums side by side, which
// replaces a chain of block-wide reductions (8 x ~1.2k cycles per adjoint iteration) by one pass.
template <typename T>
__device__ __forceinline__ double warp_sum_records(const T* __restrict__ v, int n, int n_loc, int rec) {
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int i0 = lane; i0 < n; i0 += 32 * 8) {
    T x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + 32 * u;
      const int r = i / n_loc;
      x[u] = (i < n) ? v[(size_t)r * rec + (i - r * n_loc)] : (T)0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (double)x[u];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// optimal_step_size, lib/ode.py:529-538, in fp32 like the reference's traced scalars.
__device__ __forceinline__ float next_dt(float last, float ratio, int order) {
  const float dfactor = (ratio < 1.0f) ? 1.0f : 0.2f;
  const float err = sqrtf(ratio);
  float factor = fminf(powf(err, 1.0f / (float)order) / 0.9f, 1.0f / dfactor);
  factor = fmaxf(1.0f / 10.0f, factor);
  // np.maximum / np.minimum propagate NaN; fminf/fmaxf do not:
  if (isnan(ratio)) factor = ratio;
  return (ratio == 0.0f) ? last * 10.0f : last / factor;
}

// Loop bookkeeping after an iteration (or after init): leave every target whose
// while-condition (lib/ode.py:829-831) is false, queueing its dense output.
// Works on a register copy of the controller: through the pointer every field access is a dependent global round trip
// (the compiler must assume ts / trace alias it), which made this single-thread tail ~20 k cycles per iteration.
__device__ __forceinline__ void advance_targets_local(Ctrl& c, const float* __restrict__ ts) {
  while (c.tgt < c.T) {
    const float target = ts[c.tgt];
    const bool go = (c.t < target) && (c.iters_tgt < c.mxstep) && (c.dt > 0.0f);
    if (go) break;
    if (c.t < target) {  // left early: record why (first cause wins)
      if (c.status == ENO_OK) c.status = (c.iters_tgt >= c.mxstep) ? ENO_MXSTEP : ENO_DT_UNDERFLOW;
    }
    c.out_end = c.tgt + 1;
    c.tgt += 1;
    c.iters_tgt = 0;
  }
  if (c.tgt >= c.T) c.done = 1;
}
__device__ __forceinline__ void advance_targets(Ctrl* cp, const float* __restrict__ ts) {
  Ctrl c = *cp;
  advance_targets_local(c, ts);
  *cp = c;
}

// Controller update of one iteration (lib/ode.py:836-847).  Single thread.
__device__ __forceinline__ void finish_iteration(Ctrl* cp, const float* __restrict__ ts, float ratio, float dt_used,
                                                 const Tableau& tab, eno_trace_row* trace) {
  Ctrl c = *cp;
  const bool ok = (ratio <= 1.0f);
  if (trace && c.n_iters < c.trace_cap) {
    eno_trace_row r;
    r.t = c.t; r.dt = dt_used; r.ratio = ratio; r.accepted = ok ? 1.0f : 0.0f;
    trace[c.n_iters] = r;
  }
  if (!(ratio == ratio) || isinf(ratio)) { if (c.status == ENO_OK) c.status = ENO_NONFINITE; }
  const float dtn = next_dt(dt_used, ratio, tab.order);
  c.ratio = ratio;
  c.out_begin = c.out_end;
  if (ok) {
    const int old_prev = c.prev;
    c.prev = c.cur;
    c.cur = c.cand;
    c.cand = old_prev;
    c.mid_acc ^= 1;
    c.last_t = c.t;
    c.t = c.t + dt_used;
    c.step_dt = dt_used;
    c.n_accepted += 1;
  }
  c.dt = dtn;
  c.iters_tgt += 1;
  c.n_iters += 1;
  c.nfe += tab.nfe_per_iter;
  advance_targets_local(c, ts);
  *cp = c;
}

}  // namespace eno
