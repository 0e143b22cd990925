// Forward adaptive-RK solve on the plain state y [B, D] with MLP dynamics.
//
// One launch per loop iteration (k_fwd_step): every CTA owns R samples and runs all
// stages of the tableau for them out of shared memory, forms the candidate state,
// the embedded error estimate and per-sample partial sums of (err/tol)^2; the last
// CTA to finish reduces those in a canonical order and runs the accept/reject
// controller in device memory (Ctrl).  The host never decides anything per step.
//
// Reference behaviour restated (not ported): lib/ode.py:823-862 (dopri5 loop),
// 1048-1390 (clip-flavour loops), 86-104 (initial step), 518-538 (norm, controller),
// 56-84 + 849-850 (dense output).
#pragma once
#include "dist.cuh"
#include "mlp_device.cuh"

namespace eno {

__constant__ Tableau c_tab;

struct FwdArgs {
  Ctrl* ctrl;
  MlpWeights w;
  int B;
  size_t N;             // B * D
  const float* y0;      // [B, D]
  const float* ts;      // [T]
  float* Y;             // [3][N] ring of states
  float* F;             // [3][N] ring of derivatives at those states
  float* MID;           // [2][N] mid-point of candidate / accepted step
  double* rowpart0;     // per-sample partial sums, THIS rank's rows (view into rp0_all)
  double* rowpart1;
  const double* rp0_all;  // all ranks' partials in canonical global row order, n_rp entries
  const double* rp1_all;
  int n_rp;             // entries summed by the finaliser: stride * global batch
  int dist;             // 1: multi-GPU - the step kernel leaves the reduction + controller to k_fwd_finish
  float* ys_out;        // [T][N]
  eno_trace_row* trace; // or nullptr
};

// last-block-done election; returns true in every thread of the last CTA.
__device__ __forceinline__ bool elect_last_block(unsigned int* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int v = atomicAdd(ticket, 1u);
    s_last = (v == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Sum of q[r][0..D) per sample in double, fixed order for a given D (lane-strided,
// then xor tree), so the value does not depend on R or on the grid.
template <int R>
__device__ __forceinline__ void row_sums_to_global(const float* q, int D, int row0, int B, double* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) acc += (double)q[r * D + d];
    acc = warp_sum(acc);
    if (lane == 0 && row0 + r < B) out[row0 + r] = acc;
  }
}

struct FwdSmem {
  float *y, *k, *xs, *q;
  MlpScratch sc;
  double* dred;
};

template <int R>
__device__ __forceinline__ FwdSmem carve_fwd(float* sm, int D, int H, int s) {
  FwdSmem m;
  m.dred = reinterpret_cast<double*>(sm);
  sm += 2 * kThreads;
  m.y = sm; sm += R * D;
  m.k = sm; sm += (size_t)s * R * D;
  m.xs = sm; sm += R * D;
  m.q = sm; sm += R * D;
  m.sc.s = sm; sm += R * D;
  m.sc.u = nullptr;
  m.sc.h = sm; sm += R * H;
  m.sc.a1 = m.sc.h1 = m.sc.a2 = m.sc.h2 = nullptr;
  m.sc.red = sm;
  return m;
}

inline size_t fwd_smem_bytes(int R, int D, int H, int s) {
  return sizeof(float) * ((size_t)2 * kThreads + (size_t)R * D * (4 + s) + (size_t)R * H + (size_t)kRedFloats * R);
}

// ---- initial step size, two phases (lib/ode.py:86-104, 853-859) -------------------
template <int R, int PHASE>
__global__ void __launch_bounds__(kThreads, 1) k_fwd_init(FwdArgs a) {
  extern __shared__ float4 smem4[];
  Ctrl* c = a.ctrl;
  const int D = a.w.D, H = a.w.H;
  FwdSmem m = carve_fwd<R>(reinterpret_cast<float*>(smem4), D, H, 2);
  const int row0 = blockIdx.x * R;
  const float t0 = a.ts[0];
  const float rtol = c->rtol, atol = c->atol;
  const float h0 = c->h0;
  float* f0 = m.k;            // [R][D]
  float* f1 = m.k + R * D;    // [R][D]
  for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
    const int r = i / D;
    const bool ok = row0 + r < a.B;
    const size_t g = (size_t)row0 * D + i;
    m.y[i] = ok ? a.y0[g] : 0.f;
    if (PHASE == 1) {
      f0[i] = ok ? a.F[g] : 0.f;
      m.xs[i] = fmaf(h0, f0[i], m.y[i]);
    }
  }
  __syncthreads();
  if (PHASE == 0) {
    mlp_primal<R>(a.w, t0, m.y, f0, m.sc);
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const int r = i / D;
      const size_t g = (size_t)row0 * D + i;
      const float y = m.y[i], f = f0[i];
      const float scale = atol + fabsf(y) * rtol;
      const float q0 = y / scale, q1 = f / scale;
      m.q[i] = q0 * q0;
      m.xs[i] = q1 * q1;
      if (row0 + r < a.B) {
        a.F[g] = f;
        a.Y[g] = y;
        a.ys_out[g] = y;
      }
    }
    __syncthreads();
    row_sums_to_global<R>(m.q, D, row0, a.B, a.rowpart0);
    row_sums_to_global<R>(m.xs, D, row0, a.B, a.rowpart1);
  } else {
    mlp_primal<R>(a.w, t0 + h0, m.xs, f1, m.sc);
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const float scale = atol + fabsf(m.y[i]) * rtol;
      const float q = (f1[i] - f0[i]) / scale;
      m.q[i] = q * q;
    }
    __syncthreads();
    row_sums_to_global<R>(m.q, D, row0, a.B, a.rowpart0);
  }
  if (a.dist) return;
  if (!elect_last_block(&c->ticket)) return;
  const double s0 = block_sum_global(a.rp0_all, a.n_rp, m.dred);
  __syncthreads();
  double s1 = 0.0;
  if (PHASE == 0) s1 = block_sum_global(a.rp1_all, a.n_rp, m.dred);
  if (threadIdx.x == 0) {
    c->ticket = 0;
    if (PHASE == 0) {
      const float d0 = sqrtf((float)s0), d1 = sqrtf((float)s1);
      c->d0 = d0;
      c->d1 = d1;
      c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
    } else {
      const float d1 = c->d1;
      const float d2 = sqrtf((float)s0) / h0;
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
      else h1 = powf(0.01f / (d1 + d2), 1.0f / ((float)c_tab.init_order + 1.0f));
      c->dt = fminf(100.f * h0, h1);
      if (isnan(h0) || isnan(h1)) c->dt = nanf("");
      c->t = t0;
      c->last_t = t0;
      c->step_dt = 0.f;
      c->cur = 0; c->prev = 2; c->cand = 1; c->mid_acc = 0;
      c->iters_tgt = 0; c->tgt = 1;
      c->n_iters = 0; c->n_accepted = 0; c->nfe = 2;
      c->out_begin = 1; c->out_end = 1;
      c->done = 0; c->status = ENO_OK;
      advance_targets(c, a.ts);
    }
  }
}

// ---- one loop iteration ---------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 1) k_fwd_step(FwdArgs a) {
  Ctrl* c = a.ctrl;
  if (c->done) return;
  extern __shared__ float4 smem4[];
  const int D = a.w.D, H = a.w.H;
  const int s = c_tab.s;
  FwdSmem m = carve_fwd<R>(reinterpret_cast<float*>(smem4), D, H, s);
  const int RD = R * D;
  const float t = c->t;
  float dt = c->dt;
  const float rtol = c->rtol, atol = c->atol;
  const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
  const float target = a.ts[c->tgt];
  if (c_tab.clip && (t + dt > target)) dt = target - t;
  const int row0 = blockIdx.x * R;
  const float* Yc = a.Y + (size_t)cur * a.N;
  const float* Fc = a.F + (size_t)cur * a.N;
  float* Yn = a.Y + (size_t)cand * a.N;
  float* Fn = a.F + (size_t)cand * a.N;
  float* Mn = a.MID + (size_t)midw * a.N;

  for (int i = threadIdx.x; i < RD; i += blockDim.x) {
    const bool ok = row0 + i / D < a.B;
    const size_t g = (size_t)row0 * D + i;
    m.y[i] = ok ? Yc[g] : 0.f;
    m.k[i] = ok ? Fc[g] : 0.f;
  }
  __syncthreads();

  for (int st = 1; st < s; ++st) {
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      float acc = 0.f;
      for (int j = 0; j < st; ++j) acc = fmaf(c_tab.beta[st - 1][j], m.k[j * RD + i], acc);
      m.xs[i] = fmaf(dt, acc, m.y[i]);
    }
    __syncthreads();
    mlp_primal<R>(a.w, t + dt * c_tab.alpha[st - 1], m.xs, m.k + (size_t)st * RD, m.sc);
  }

  // candidate state, error estimate, (dopri5) mid-point
  for (int i = threadIdx.x; i < RD; i += blockDim.x) {
    float as = 0.f, ae = 0.f, am = 0.f;
    for (int j = 0; j < s; ++j) {
      const float kj = m.k[j * RD + i];
      as = fmaf(c_tab.c_sol[j], kj, as);
      ae = fmaf(c_tab.c_err[j], kj, ae);
      am = fmaf(c_tab.c_mid[j], kj, am);
    }
    const float y = m.y[i];
    const float y1 = fmaf(dt, as, y);
    const float err = dt * ae;
    const float tol = atol + rtol * fmaxf(fabsf(y), fabsf(y1));
    const float q = err / tol;
    m.q[i] = q * q;
    if (row0 + i / D < a.B) {
      const size_t g = (size_t)row0 * D + i;
      Yn[g] = y1;
      Fn[g] = m.k[(size_t)(s - 1) * RD + i];
      if (!c_tab.clip) Mn[g] = fmaf(dt, am, y);
    }
  }
  __syncthreads();
  row_sums_to_global<R>(m.q, D, row0, a.B, a.rowpart0);

  if (c_tab.clip) {
    // mid-point from a second step of dt/2 (lib/ode.py:1063 and siblings); k[0] is reused.
    const float dt2 = dt / 2;
    for (int st = 1; st < s; ++st) {
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < st; ++j) acc = fmaf(c_tab.beta[st - 1][j], m.k[j * RD + i], acc);
        m.xs[i] = fmaf(dt2, acc, m.y[i]);
      }
      __syncthreads();
      mlp_primal<R>(a.w, t + dt2 * c_tab.alpha[st - 1], m.xs, m.k + (size_t)st * RD, m.sc);
    }
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      float as = 0.f;
      for (int j = 0; j < s; ++j) as = fmaf(c_tab.c_sol[j], m.k[j * RD + i], as);
      if (row0 + i / D < a.B) Mn[(size_t)row0 * D + i] = fmaf(dt2, as, m.y[i]);
    }
  }

  if (a.dist) return;
  if (!elect_last_block(&c->ticket)) return;
  const double sum = block_sum_global(a.rp0_all, a.n_rp, m.dred);
  if (threadIdx.x == 0) {
    c->ticket = 0;
    const float ratio = (float)(sum / (double)c->n_state);
    finish_iteration(c, a.ts, ratio, dt, c_tab, a.trace);
  }
}

// Multi-GPU: reduction + controller as a separate one-CTA launch, after the per-sample partials of
// every rank have been all-gathered.  mode 0 iteration, 1 / 2 initial-step phases.  Every rank runs it on
// identical inputs and therefore takes identical decisions.
__device__ __forceinline__ void fwd_finish_body(const FwdArgs& a, int mode, double* dred, const double* rp0_all, const double* rp1_all) {
  Ctrl* c = a.ctrl;
  const double s0 = block_sum_global(rp0_all, a.n_rp, dred);
  __syncthreads();
  double s1 = 0.0;
  if (mode == 1) s1 = block_sum_global(rp1_all, a.n_rp, dred);
  if (threadIdx.x != 0) return;
  if (mode == 1) {
    const float d0 = sqrtf((float)s0), d1 = sqrtf((float)s1);
    c->d0 = d0;
    c->d1 = d1;
    c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
  } else if (mode == 2) {
    const float h0 = c->h0, d1 = c->d1;
    const float d2 = sqrtf((float)s0) / h0;
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
    else h1 = powf(0.01f / (d1 + d2), 1.0f / ((float)c_tab.init_order + 1.0f));
    c->dt = fminf(100.f * h0, h1);
    if (isnan(h0) || isnan(h1)) c->dt = nanf("");
    c->t = a.ts[0];
    c->last_t = c->t;
    c->step_dt = 0.f;
    c->cur = 0; c->prev = 2; c->cand = 1; c->mid_acc = 0;
    c->iters_tgt = 0; c->tgt = 1;
    c->n_iters = 0; c->n_accepted = 0; c->nfe = 2;
    c->out_begin = 1; c->out_end = 1;
    c->done = 0; c->status = ENO_OK;
    advance_targets(c, a.ts);
  } else {
    float dt = c->dt;
    const float target = a.ts[c->tgt];
    if (c_tab.clip && (c->t + dt > target)) dt = target - c->t;
    const float ratio = (float)(s0 / (double)c->n_state);
    finish_iteration(c, a.ts, ratio, dt, c_tab, a.trace);
  }
}

__global__ void __launch_bounds__(kThreads) k_fwd_finish(FwdArgs a, int mode) {
  __shared__ double dred[kThreads];
  if (mode == 0 && a.ctrl->done) return;
  // the next iteration kernel (launched with programmatic serialization) may load its weights under this exchange
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  fwd_finish_body(a, mode, dred, a.rp0_all, a.rp1_all);
}

// The same with the exchange inside the kernel (dist.cuh): push this rank's per-sample partials into every rank's
// gathered array over NVLink, signal, wait for all ranks, then reduce + run the controller.  No NCCL call.
__global__ void __launch_bounds__(kThreads) k_fwd_finish_peer(FwdArgs a, int mode, unsigned char* peer_base, int n_loc) {
  __shared__ double dred[kThreads];
  if (mode == 0 && a.ctrl->done) return;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // as in k_fwd_finish
  const PeerView pv = *reinterpret_cast<const PeerView*>(peer_base + kPeerOffView);
  unsigned long long* seqp = reinterpret_cast<unsigned long long*>(peer_base + kPeerOffSeq);
  const unsigned long long seq = *seqp + 1;
  const size_t par_off = (size_t)(seq & 1) * 2 * a.n_rp;      // doubles
  for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
    const double v0 = a.rowpart0[i], v1 = (mode == 1) ? a.rowpart1[i] : 0.0;
    for (int r = 0; r < pv.world; ++r) {
      double* dst = reinterpret_cast<double*>(pv.base[r] + kPeerOffFwd) + par_off + (size_t)pv.rank * n_loc;
      dst[i] = v0;
      if (mode == 1) dst[a.n_rp + i] = v1;
    }
  }
  __syncthreads();
  peer_signal(pv, 0, seq);
  peer_wait(pv, 0, seq);
  const double* mine = reinterpret_cast<const double*>(peer_base + kPeerOffFwd) + par_off;
  fwd_finish_body(a, mode, dred, mine, mine + a.n_rp);
  if (threadIdx.x == 0) *seqp = seq;
}

// ---- dense output for the targets the last iteration left behind ----------------------
__global__ void k_fwd_out(FwdArgs a) {
  const Ctrl* c = a.ctrl;
  const int ob = c->out_begin, oe = c->out_end;
  if (oe <= ob) return;
  const float t = c->t, last_t = c->last_t, dt = c->step_dt;
  const bool degenerate = (c->n_accepted == 0);
  const float* y_init = a.ys_out;
  const float* Yp = degenerate ? y_init : (c_tab.clip ? y_init : a.Y + (size_t)c->prev * a.N);
  const float* Yc = a.Y + (size_t)c->cur * a.N;
  const float* Mc = a.MID + (size_t)c->mid_acc * a.N;
  const float* Fp = a.F + (size_t)c->prev * a.N;
  const float* Fc = a.F + (size_t)c->cur * a.N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (size_t)gridDim.x * blockDim.x) {
    float ca, cb, cc, cd, ce;
    if (degenerate) {
      ca = cb = cc = cd = ce = y_init[i];
    } else {
      const float y0 = Yp[i], y1 = Yc[i], ym = Mc[i], dy0 = Fp[i], dy1 = Fc[i];
      ca = -2.f * dt * dy0 + 2.f * dt * dy1 - 8.f * y0 - 8.f * y1 + 16.f * ym;
      cb = 5.f * dt * dy0 - 3.f * dt * dy1 + 18.f * y0 + 14.f * y1 - 32.f * ym;
      cc = -4.f * dt * dy0 + dt * dy1 - 11.f * y0 - 5.f * y1 + 16.f * ym;
      cd = dt * dy0;
      ce = y0;
    }
    for (int o = ob; o < oe; ++o) {
      const float x = (a.ts[o] - last_t) / (t - last_t);
      float v = 0.f;
      v = v * x + ca;
      v = v * x + cb;
      v = v * x + cc;
      v = v * x + cd;
      v = v * x + ce;
      a.ys_out[(size_t)o * a.N + i] = v;
    }
  }
}

// ---- single evaluation (eno_rhs mode 0) -------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 1) k_rhs_plain(MlpWeights w, int B, const float* y, float t, float* out) {
  extern __shared__ float4 smem4[];
  const int D = w.D, H = w.H;
  FwdSmem m = carve_fwd<R>(reinterpret_cast<float*>(smem4), D, H, 1);
  const int row0 = blockIdx.x * R;
  for (int i = threadIdx.x; i < R * D; i += blockDim.x)
    m.y[i] = (row0 + i / D < B) ? y[(size_t)row0 * D + i] : 0.f;
  __syncthreads();
  mlp_primal<R>(w, t, m.y, m.k, m.sc);
  for (int i = threadIdx.x; i < R * D; i += blockDim.x)
    if (row0 + i / D < B) out[(size_t)row0 * D + i] = m.k[i];
}

}  // namespace eno
