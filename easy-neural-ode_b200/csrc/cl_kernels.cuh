// Cluster-resident step kernels (the fast path when the weight slices fit in shared memory):
// same controller protocol, workspace and outputs as fwd_kernels.cuh / adj_kernels.cuh, but a
// 5-CTA cluster owns a tile of R samples, keeps its weight slices in shared memory for the whole
// launch, loops persistently over tiles, and exchanges only R x H partial sums over DSMEM.
#pragma once
#include <cstdlib>

#include "adj_kernels.cuh"
#include "cl_device.cuh"

namespace eno {

// ------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------
struct ClSmem {
  float *W1, *W2, *vec;
  ClScratch m;
  ClComm cm;
  float *xz, *xzb;        // stage inputs [R][Dlp]
  float *y, *k;           // forward only: state and stage derivatives [R][Dlp], [s][R][Dlp]
  float *r, *rb, *kr;     // per-sample scalars [R], [R], [7][R]
  float* aux;             // [36 R] scalar state / stage derivatives of the augmented forward solve
  double* dred;           // [kThreads] (forward finaliser)
};

template <int R>
__device__ __forceinline__ ClSmem cl_carve(float* sm, int D, int H, const ClGeom& g, int n_dvec_extra, bool adjoint) {
  ClSmem s;
  const int RD = R * g.Dlp, RH = R * H;
  const ClGeom g0 = cl_geom(D, H, 0);   // widest slice
  s.dred = reinterpret_cast<double*>(sm);
  if (!adjoint) sm += 2 * kThreads;        // only the forward kernel hosts a finaliser
  s.W1 = sm; sm += (size_t)g0.Dl * g.ldw1;
  s.W2 = sm; sm += (size_t)H * g.Dlp;
  s.vec = sm; sm += 2 * H + 2 * g.Dlp;
  s.m.s = sm; sm += RD;
  s.m.u = sm; sm += RD;
  s.m.f = sm; sm += RD;
  s.xz = sm; sm += RD;
  s.m.h = sm; sm += RH;
  if (adjoint) {
    s.xzb = sm; sm += RD;
    s.m.y1 = sm; sm += RD;
    s.m.y2 = sm; sm += RD;
    // three reverse-sweep vectors reuse storage whose forward-sweep contents are dead by then
    // (cl_aug_reverse): y2 is last read when y2_bar is formed, u after the product that consumes it,
    // the stage input after the first sigmoid
    s.m.gf = s.m.y2;
    s.m.gy1 = s.m.u;
    s.m.zacc = s.xz;
    s.m.a1 = sm; sm += RH; s.m.h1 = sm; sm += RH; s.m.a2 = sm; sm += RH; s.m.h2 = sm; sm += RH;
    s.m.ga = sm; sm += RH; s.m.ga1 = sm; sm += RH; s.m.ga2 = sm; sm += RH;
    s.y = s.k = nullptr;
  } else {
    s.xzb = s.m.y1 = s.m.y2 = s.m.gf = s.m.gy1 = s.m.zacc = nullptr;
    s.m.a1 = s.m.h1 = s.m.a2 = s.m.h2 = s.m.ga = s.m.ga1 = s.m.ga2 = nullptr;
    s.y = sm; sm += RD;
    s.k = sm; sm += (size_t)n_dvec_extra * RD;
  }
  const int red_n = (kKG * R * H > kKGw * R * g.Dlp) ? kKG * R * H : kKGw * R * g.Dlp;
  s.m.red = sm; sm += red_n;
  s.cm.part[0] = sm; sm += R * g.Hs;
  s.cm.part[1] = sm; sm += R * g.Hs;
  s.cm.phase = 0;
  s.m.scal = sm; sm += 4 * R;
  s.m.rdot = sm; sm += 4 * ((R + 3) / 4);
  s.r = sm; sm += 4 * ((R + 3) / 4);
  s.rb = sm; sm += 4 * ((R + 3) / 4);
  s.kr = sm; sm += 4 * ((7 * R + 3) / 4);
  s.aux = sm; sm += 36 * R;
  return s;
}

inline size_t cl_smem_bytes(int R, int D, int H, int stages, bool adjoint) {
  const ClGeom g = cl_geom(D, H, 0);
  const size_t RD = (size_t)R * g.Dlp, RH = (size_t)R * H;
  size_t f = (adjoint ? 0 : 2 * kThreads) + (size_t)g.Dl * g.ldw1 + (size_t)H * g.Dlp + 2 * H + 2 * g.Dlp + 4 * RD + RH;
  f += adjoint ? (3 * RD + 7 * RH) : ((size_t)(1 + stages) * RD);
  const size_t red_n = std::max((size_t)kKG * R * H, (size_t)kKGw * R * g.Dlp);
  f += red_n + 2 * (size_t)R * g.Hs + 4 * R + 3 * 4 * ((R + 3) / 4) + 4 * ((7 * R + 3) / 4) + 36 * R + 16;
  return f * sizeof(float);
}

// Can the cluster path run this problem?  (thread-count and divisibility limits of mv_n / mv_k)
inline bool cl_supported(int D, int H, int R, int stages, bool adjoint) {
  if (D % 4 || H % 4 || D < 4) return false;   // D/4 < kC: the trailing CTAs of a cluster own empty slices
  const ClGeom g = cl_geom(D, H, 0);
  if (adjoint && R * g.Dlp > 2 * kThreads) return false;   // EPT = 2 register-resident elements per thread
  if ((H / 4) * kKG * kRG > kThreads) return false;
  if ((g.Dl / 4) * kKGw * kRG > kThreads) return false;
  return cl_smem_bytes(R, D, H, stages, adjoint) <= 227 * 1024;
}

// per-sample sums of q over the local slice -> out[rank * B + row]  (+ extra[r] on rank 0)
template <int R>
__device__ __forceinline__ void cl_row_sums(const float* q, const ClGeom& g, int row0, int B, double* out,
                                            const float* extra) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    double acc = 0.0;
    for (int d = lane; d < g.Dl; d += 32) acc += (double)q[r * g.Dlp + d];
    acc = warp_sum(acc);
    if (lane == 0 && row0 + r < B) out[(size_t)(row0 + r) * kC + g.rank] = acc + ((extra && g.rank == 0) ? (double)extra[r] : 0.0);
  }
}

// Programmatic dependent launch (single-GPU iteration kernels): the weight prologue of an iteration kernel does
// not depend on the kernel before it, so it may run while that kernel's one-CTA finaliser is still busy.
// Producers call pdl_trigger() when their per-CTA work is done (just before the ticket election); a consumer
// launched with the programmatic-serialization attribute calls pdl_wait() after its prologue and only then reads
// anything the producer wrote.  `done` goes 0 -> 1 once per solve and is reset kernels earlier, so a (possibly
// stale) read of 1 BEFORE the wait is always true.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// forward: MODE 0 = iteration, 1 = initial-step phase 0, 2 = initial-step phase 1
// ------------------------------------------------------------------------------------------
template <int R, int MODE, int DT, int HT>
__global__ void __launch_bounds__(kThreads, 1) k_cl_fwd(FwdArgs a) {
  Ctrl* c = a.ctrl;
  if (MODE == 0 && *reinterpret_cast<volatile int*>(&c->done)) return;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float4 smem4[];
  const int D = DT > 0 ? DT : a.w.D, H = HT > 0 ? HT : a.w.H;   // compile-time sizes in the specialised build
  const int s = (MODE == 0) ? c_tab.s : 2;
  const ClGeom g = cl_geom(D, H, (int)cluster.block_rank());
  ClSmem S = cl_carve<R>(reinterpret_cast<float*>(smem4), D, H, g, s, false);
  cl_load_weights(a.w, g, S.W1, S.W2, S.vec);
  if (MODE == 0) {
    pdl_wait();
    if (*reinterpret_cast<volatile int*>(&c->done)) return;   // uniform over the grid
  }
  ClWeights w{S.W1, S.W2, S.vec, S.vec + H, S.vec + 2 * H, S.vec + 2 * H + g.Dlp, D, H};
  const int Dl = g.Dl, Dlp = g.Dlp, RD = R * Dlp;
  const float rtol = c->rtol, atol = c->atol;
  const float t = (MODE == 0) ? c->t : a.ts[0];
  float dt = (MODE == 0) ? c->dt : c->h0;
  const int cur = (MODE == 0) ? c->cur : 0, cand = c->cand, midw = c->mid_acc ^ 1;
  if (MODE == 0) {
    const float target = a.ts[c->tgt];
    if (c_tab.clip && (t + dt > target)) dt = target - t;
  }
  const float* Yc = a.Y + (size_t)cur * a.N;
  const float* Fc = a.F + (size_t)cur * a.N;
  float* Yn = a.Y + (size_t)cand * a.N;
  float* Fn = a.F + (size_t)cand * a.N;
  float* Mn = a.MID + (size_t)midw * a.N;
  __syncthreads();

  const int n_tiles = (a.B + R - 1) / R;
  const int n_clusters = gridDim.x / kC;
  for (int tile = blockIdx.x / kC; tile < n_tiles; tile += n_clusters) {
    const int row0 = tile * R;
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const int r = i / Dlp, d = i - r * Dlp;
      const bool ok = d < Dl && row0 + r < a.B;
      const size_t gi = (size_t)(row0 + r) * D + g.d0 + d;
      if (MODE == 1) {
        S.y[i] = ok ? a.y0[gi] : 0.f;
      } else {
        S.y[i] = ok ? Yc[gi] : 0.f;
        S.k[i] = ok ? Fc[gi] : 0.f;
      }
    }
    __syncthreads();

    if (MODE == 1) {
      cl_primal<R>(cluster, S.cm, g, w, t, S.y, S.k, S.m);
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        const int r = i / Dlp, d = i - r * Dlp;
        const float y = S.y[i], f = S.k[i];
        const float scale = atol + fabsf(y) * rtol;
        const float q0 = y / scale, q1 = f / scale;
        S.m.u[i] = q0 * q0;
        S.xz[i] = q1 * q1;
        if (d < Dl && row0 + r < a.B) {
          const size_t gi = (size_t)(row0 + r) * D + g.d0 + d;
          a.F[gi] = f;
          a.Y[gi] = y;
          a.ys_out[gi] = y;
        }
      }
      __syncthreads();
      cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, nullptr);
      cl_row_sums<R>(S.xz, g, row0, a.B, a.rowpart1, nullptr);
      __syncthreads();
      continue;
    }
    if (MODE == 2) {
      for (int i = threadIdx.x; i < RD; i += blockDim.x) S.xz[i] = fmaf(dt, S.k[i], S.y[i]);
      __syncthreads();
      cl_primal<R>(cluster, S.cm, g, w, t + dt, S.xz, S.k + RD, S.m);
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        const float scale = atol + fabsf(S.y[i]) * rtol;
        const float q = (S.k[RD + i] - S.k[i]) / scale;
        S.m.u[i] = q * q;
      }
      __syncthreads();
      cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, nullptr);
      __syncthreads();
      continue;
    }

    for (int st = 1; st < s; ++st) {
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < st; ++j) acc = fmaf(c_tab.beta[st - 1][j], S.k[j * RD + i], acc);
        S.xz[i] = fmaf(dt, acc, S.y[i]);
      }
      __syncthreads();
      cl_primal<R>(cluster, S.cm, g, w, t + dt * c_tab.alpha[st - 1], S.xz, S.k + (size_t)st * RD, S.m);
    }
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const int r = i / Dlp, d = i - r * Dlp;
      float as = 0.f, ae = 0.f, am = 0.f;
      for (int j = 0; j < s; ++j) {
        const float kj = S.k[j * RD + i];
        as = fmaf(c_tab.c_sol[j], kj, as);
        ae = fmaf(c_tab.c_err[j], kj, ae);
        am = fmaf(c_tab.c_mid[j], kj, am);
      }
      const float y = S.y[i];
      const float y1 = fmaf(dt, as, y);
      const float err = dt * ae;
      const float tol = atol + rtol * fmaxf(fabsf(y), fabsf(y1));
      const float q = err / tol;
      S.m.u[i] = q * q;
      if (d < Dl && row0 + r < a.B) {
        const size_t gi = (size_t)(row0 + r) * D + g.d0 + d;
        Yn[gi] = y1;
        Fn[gi] = S.k[(size_t)(s - 1) * RD + i];
        if (!c_tab.clip) Mn[gi] = fmaf(dt, am, y);
      }
    }
    __syncthreads();
    cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, nullptr);
    __syncthreads();
    if (c_tab.clip) {
      const float dt2 = dt / 2;
      for (int st = 1; st < s; ++st) {
        for (int i = threadIdx.x; i < RD; i += blockDim.x) {
          float acc = 0.f;
          for (int j = 0; j < st; ++j) acc = fmaf(c_tab.beta[st - 1][j], S.k[j * RD + i], acc);
          S.xz[i] = fmaf(dt2, acc, S.y[i]);
        }
        __syncthreads();
        cl_primal<R>(cluster, S.cm, g, w, t + dt2 * c_tab.alpha[st - 1], S.xz, S.k + (size_t)st * RD, S.m);
      }
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        const int r = i / Dlp, d = i - r * Dlp;
        float as = 0.f;
        for (int j = 0; j < s; ++j) as = fmaf(c_tab.c_sol[j], S.k[j * RD + i], as);
        if (d < Dl && row0 + r < a.B) Mn[(size_t)(row0 + r) * D + g.d0 + d] = fmaf(dt2, as, S.y[i]);
      }
      __syncthreads();
    }
  }
  cluster.sync();   // no CTA may exit while a peer can still read its shared memory

  if (a.dist) return;
  if (MODE == 0) pdl_trigger();   // the next iteration's launch may run its prologue under the finaliser below
  if (!elect_last_block(&c->ticket)) return;
  const int nrp = a.n_rp;
  const double s0 = block_sum_global(a.rp0_all, nrp, S.dred);
  __syncthreads();
  double s1 = 0.0;
  if (MODE == 1) s1 = block_sum_global(a.rp1_all, nrp, S.dred);
  if (threadIdx.x != 0) return;
  c->ticket = 0;
  if (MODE == 1) {
    const float d0 = sqrtf((float)s0), d1 = sqrtf((float)s1);
    c->d0 = d0;
    c->d1 = d1;
    c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
  } else if (MODE == 2) {
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
    const float ratio = (float)(s0 / (double)c->n_state);
    finish_iteration(c, a.ts, ratio, dt, c_tab, a.trace);
  }
}

// ------------------------------------------------------------------------------------------
// adjoint per-sample blocks.  Stage derivatives live in REGISTERS (each thread owns EPT elements of
// the tile's local slice for the whole iteration); shared memory only holds the RHS working set.
// ------------------------------------------------------------------------------------------
// FIN = the "fin" augmented dynamics (cl_fin_forward / cl_fin_reverse): two scalar states per sample and
// the eps_bar block, which never feeds back into the RHS and is therefore carried as running stage
// combinations (solution, error, mid-point) instead of stage buffers.
// TC (ADJ_STEP only): the stage factors go to the tensor-core sinks (single GPU) instead of the row-major sinks that
// k_adj_pgrad contracts (collective solves: one launch fewer per iteration, measured faster there).
template <int R, int K, int MODE, int DT, int HT, bool FIN = false, bool TC = false>
__global__ void __launch_bounds__(kThreads, 1) k_cl_adj_rows(AdjArgs a, float rhs_tau) {
  Ctrl* c = a.ctrl;
  if (MODE == ADJ_STEP && *reinterpret_cast<volatile int*>(&c->done)) return;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float4 smem4[];
  const int D = DT > 0 ? DT : a.w.D, H = HT > 0 ? HT : a.w.H;
  const ClGeom g = cl_geom(D, H, (int)cluster.block_rank());
  ClSmem S = cl_carve<R>(reinterpret_cast<float*>(smem4), D, H, g, 0, true);
  ENO_TICK_START();
  cl_load_weights(a.w, g, S.W1, S.W2, S.vec);
  if (MODE == ADJ_STEP) {
    pdl_wait();   // the prologue above overlapped the parameter-gradient kernel's finaliser
    if (*reinterpret_cast<volatile int*>(&c->done)) return;   // uniform over the grid
  }
  ClWeights w{S.W1, S.W2, S.vec, S.vec + H, S.vec + 2 * H, S.vec + 2 * H + g.Dlp, D, H};
  const int Dl = g.Dl, Dlp = g.Dlp, RD = R * Dlp;
  constexpr int EPT = 2;   // elements per thread: R * Dlp <= 2 * kThreads (checked on the host)
  const float rtol = c->rtol, atol = c->atol;
  const int cur = (MODE == ADJ_RHS) ? 0 : c->cur;
  const int cand = c->cand, midw = c->mid_acc ^ 1;
  const float tau = (MODE == ADJ_STEP) ? c->t : (MODE == ADJ_RHS ? rhs_tau : a.sc->seg_ts[0]);
  const float dt = (MODE == ADJ_STEP) ? c->dt : c->h0;
  const size_t B = a.B;
  constexpr int NA = FIN ? 2 : 1;   // scalar states per sample
  float* const sr = FIN ? S.aux : S.r;                 // [NA][R] state
  float* const srb = FIN ? S.aux + 2 * R : S.rb;       // [NA][R] cotangent (constant during a segment)
  float* const skr = FIN ? S.aux + 4 * R : S.kr;       // [7][NA][R] stage derivatives
  float* const srd = FIN ? S.aux + 18 * R : S.m.rdot;  // [NA][R] RHS output
  __syncthreads();
  ENO_TICK(PH_PROLOGUE);

  const int n_tiles = (a.B + R - 1) / R;
  const int n_clusters = gridDim.x / kC;
  for (int tile = blockIdx.x / kC; tile < n_tiles; tile += n_clusters) {
    const int row0 = tile * R;
    float z[EPT], zb[EPT], kz[kAdjStages][EPT], kzb[kAdjStages][EPT];
    float ze[EPT], ke0[EPT], kel[EPT], se[EPT], ee[EPT], me[EPT];   // eps_bar block (FIN only)
    size_t gi[EPT];
    bool ok[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int e = threadIdx.x + j * kThreads;
      const int r = e / Dlp, d = e - r * Dlp;
      ok[j] = e < RD && d < Dl && row0 + r < a.B;
      gi[j] = ok[j] ? (size_t)(row0 + r) * D + g.d0 + d : 0;
      z[j] = ok[j] ? a.Z[(size_t)cur * a.N + gi[j]] : 0.f;
      zb[j] = ok[j] ? a.ZB[(size_t)cur * a.N + gi[j]] : 0.f;
      kz[0][j] = kzb[0][j] = 0.f;
      if (MODE == ADJ_STEP || MODE == ADJ_INIT1) {
        kz[0][j] = ok[j] ? a.FZ[(size_t)cur * a.N + gi[j]] : 0.f;
        kzb[0][j] = ok[j] ? a.FZB[(size_t)cur * a.N + gi[j]] : 0.f;
      }
#pragma unroll
      for (int st = 1; st < kAdjStages; ++st) kz[st][j] = kzb[st][j] = 0.f;
      ze[j] = ke0[j] = kel[j] = se[j] = ee[j] = me[j] = 0.f;
      if (FIN) {
        ze[j] = ok[j] ? a.ZE[(size_t)cur * a.N + gi[j]] : 0.f;
        if (MODE == ADJ_STEP || MODE == ADJ_INIT1) ke0[j] = ok[j] ? a.FZE[(size_t)cur * a.N + gi[j]] : 0.f;
        se[j] = c_tab.c_sol[0] * ke0[j];
        ee[j] = c_tab.c_err[0] * ke0[j];
        me[j] = c_tab.c_mid[0] * ke0[j];
      }
    }
    if (threadIdx.x < NA * R) {
      const int q = threadIdx.x / R, r = threadIdx.x - q * R;
      const bool okr = row0 + r < a.B;
      sr[threadIdx.x] = okr ? a.RS[((size_t)cur * NA + q) * B + row0 + r] : 0.f;
      srb[threadIdx.x] = okr ? a.RB[(size_t)q * B + row0 + r] : 0.f;
      skr[threadIdx.x] = (okr && (MODE == ADJ_STEP || MODE == ADJ_INIT1)) ? a.FR[((size_t)cur * NA + q) * B + row0 + r] : 0.f;
    }

    constexpr int NEVAL = (MODE == ADJ_STEP) ? kAdjStages - 1 : 1;
    for (int ev = 1; ev <= NEVAL; ++ev) {
      // stage input
      float tev;
      if (MODE == ADJ_STEP) tev = tau + dt * c_tab.alpha[ev - 1];
      else if (MODE == ADJ_INIT1) tev = tau + dt;
      else tev = tau;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          float vz = z[j], vb = zb[j];
          if (MODE == ADJ_STEP) {
            float az = 0.f, ab = 0.f;
#pragma unroll
            for (int q = 0; q < kAdjStages - 1; ++q) {
              const float bq = c_tab.beta[ev - 1][q];   // zero above the diagonal
              az = fmaf(bq, kz[q][j], az);
              ab = fmaf(bq, kzb[q][j], ab);
            }
            vz = fmaf(dt, az, vz);
            vb = fmaf(dt, ab, vb);
          } else if (MODE == ADJ_INIT1) {
            vz = fmaf(dt, kz[0][j], vz);
            vb = fmaf(dt, kzb[0][j], vb);
          }
          S.xz[e] = vz;
          S.xzb[e] = vb;
        }
      }
      __syncthreads();
      ENO_TICK(PH_ELEM);
      // the ADJ_STEP factors feed the tensor-core contraction (transposed hi / lo planes), the others k_adj_pgrad
      const FactorDstT<TC> slot(a.slot[(MODE == ADJ_STEP) ? ev - 1 : 0], a.tc, (MODE == ADJ_STEP) ? ev - 1 : 0);
      if (FIN) {
        cl_fin_forward<R>(cluster, S.cm, g, w, -tev, S.xz, a.eps, S.m, &slot, row0, a.B);
        cl_fin_reverse<R>(cluster, S.cm, g, w, S.xzb, srb, a.eps, S.m, srd, slot, row0, a.B);
      } else {
        cl_aug_forward<R, K>(cluster, S.cm, g, w, -tev, S.xz, S.m, &slot, row0, a.B);
        cl_aug_reverse<R, K>(cluster, S.cm, g, w, S.xzb, S.rb, S.m, slot, row0, a.B);
      }
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        const float vf = (e < RD) ? -S.m.f[e] : 0.f;
        const float vb = (e < RD) ? S.m.zacc[e] : 0.f;
#pragma unroll
        for (int st = 1; st < kAdjStages; ++st)
          if (st == ev) { kz[st][j] = vf; kzb[st][j] = vb; }
        if (FIN) {
          kel[j] = (e < RD) ? S.m.gy1[e] : 0.f;   // eps_bar' (cl_fin_reverse leaves it in gy1's storage)
          if (MODE == ADJ_STEP) {
            se[j] = fmaf(c_tab.c_sol[ev], kel[j], se[j]);
            ee[j] = fmaf(c_tab.c_err[ev], kel[j], ee[j]);
            me[j] = fmaf(c_tab.c_mid[ev], kel[j], me[j]);
          }
        }
      }
      if (threadIdx.x < NA * R) skr[ev * NA * R + threadIdx.x] = -srd[threadIdx.x];
      __syncthreads();
      ENO_TICK(PH_SCALAR);
    }

    if (MODE == ADJ_RHS) {
#pragma unroll
      for (int j = 0; j < EPT; ++j)
        if (ok[j]) {
          a.rhs_dy[gi[j]] = -kz[1][j];
          if (a.rhs_zb) a.rhs_zb[gi[j]] = kzb[1][j];
          if (FIN && a.rhs_eb) a.rhs_eb[gi[j]] = kel[j];
        }
      if (g.rank == 0 && threadIdx.x < NA * R && a.rhs_r) {
        const int q = threadIdx.x / R, r = threadIdx.x - q * R;
        if (row0 + r < a.B) a.rhs_r[(size_t)q * B + row0 + r] = -skr[NA * R + threadIdx.x];
      }
      continue;
    }

    if (MODE == ADJ_INIT0) {
      // f0 -> derivative ring; partial 2-norms; t_bar pieces <func(ys[i], ts[i]), g[i]> (lib/ode.py:1513)
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          const float kzv = kz[1][j], kbv = kzb[1][j];
          const float sz = atol + fabsf(z[j]) * rtol, szb = atol + fabsf(zb[j]) * rtol;
          const float q0a = z[j] / sz, q0b = zb[j] / szb, q1a = kzv / sz, q1b = kbv / szb;
          float p0 = q0a * q0a + q0b * q0b, p1 = q1a * q1a + q1b * q1b;
          if (FIN) {
            const float sze = atol + fabsf(ze[j]) * rtol;
            const float q0e = ze[j] / sze, q1e = kel[j] / sze;
            p0 = fmaf(q0e, q0e, p0);
            p1 = fmaf(q1e, q1e, p1);
            if (ok[j]) a.FZE[(size_t)cur * a.N + gi[j]] = kel[j];
          }
          S.m.u[e] = ok[j] ? p0 : 0.f;
          S.xz[e] = ok[j] ? p1 : 0.f;
          S.xzb[e] = ok[j] ? (-kzv) * a.g_y[gi[j]] : 0.f;
          if (ok[j]) {
            a.FZ[(size_t)cur * a.N + gi[j]] = kzv;
            a.FZB[(size_t)cur * a.N + gi[j]] = kbv;
          }
        }
      }
      if (threadIdx.x < R) {
        const int r = threadIdx.x;
        const bool okr = row0 + r < a.B;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          const float rr = sr[q * R + r], rbv = srb[q * R + r], kr = skr[NA * R + q * R + r];
          const float tr = atol + fabsf(rr) * rtol, trb = atol + fabsf(rbv) * rtol;
          const float q0 = rr / tr, q0b = rbv / trb, q1 = kr / tr;
          a0 += q0 * q0 + q0b * q0b;
          a1 += q1 * q1;
          if (okr) {
            a2 += (-kr) * a.g_r[(size_t)q * B + row0 + r];
            if (g.rank == 0) a.FR[((size_t)cur * NA + q) * B + row0 + r] = kr;
          }
        }
        S.m.scal[r * 4 + 0] = a0;
        S.m.scal[r * 4 + 1] = a1;
        S.m.scal[r * 4 + 2] = a2;
      }
      __syncthreads();
      {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int r = warp; r < R; r += nw) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
          for (int d = lane; d < Dl; d += 32) {
            s0 += (double)S.m.u[r * Dlp + d];
            s1 += (double)S.xz[r * Dlp + d];
            s2 += (double)S.xzb[r * Dlp + d];
          }
          s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
          if (lane == 0 && row0 + r < a.B) {
            const bool lead = g.rank == 0;
            a.rowpart0[(size_t)(row0 + r) * kC + g.rank] = s0 + (lead ? (double)S.m.scal[r * 4 + 0] : 0.0);
            a.rowpart1[(size_t)(row0 + r) * kC + g.rank] = s1 + (lead ? (double)S.m.scal[r * 4 + 1] : 0.0);
            a.rowpart2[(size_t)(row0 + r) * kC + g.rank] = s2 + (lead ? (double)S.m.scal[r * 4 + 2] : 0.0);
          }
        }
      }
      __syncthreads();
      continue;
    }

    if (MODE == ADJ_INIT1) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          const float sz = atol + fabsf(z[j]) * rtol, szb = atol + fabsf(zb[j]) * rtol;
          const float qa = (kz[1][j] - kz[0][j]) / sz, qb = (kzb[1][j] - kzb[0][j]) / szb;
          float p = qa * qa + qb * qb;
          if (FIN) {
            const float qe = (kel[j] - ke0[j]) / (atol + fabsf(ze[j]) * rtol);
            p = fmaf(qe, qe, p);
          }
          S.m.u[e] = ok[j] ? p : 0.f;
        }
      }
      if (threadIdx.x < R) {
        const int r = threadIdx.x;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          const float tr = atol + fabsf(sr[q * R + r]) * rtol;
          const float d = (skr[NA * R + q * R + r] - skr[q * R + r]) / tr;
          acc += d * d;
        }
        S.m.rdot[r] = acc;
      }
      __syncthreads();
      cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, S.m.rdot);
      __syncthreads();
      continue;
    }

    // ADJ_STEP: candidate state, error estimate, y_bar mid-point
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int e = threadIdx.x + j * kThreads;
      if (e < RD) {
        float sz = 0.f, ez = 0.f, sb = 0.f, eb = 0.f, mb = 0.f;
#pragma unroll
        for (int q = 0; q < kAdjStages; ++q) {
          sz = fmaf(c_tab.c_sol[q], kz[q][j], sz);
          ez = fmaf(c_tab.c_err[q], kz[q][j], ez);
          sb = fmaf(c_tab.c_sol[q], kzb[q][j], sb);
          eb = fmaf(c_tab.c_err[q], kzb[q][j], eb);
          mb = fmaf(c_tab.c_mid[q], kzb[q][j], mb);
        }
        const float z1 = fmaf(dt, sz, z[j]), zb1 = fmaf(dt, sb, zb[j]);
        const float tz = atol + rtol * fmaxf(fabsf(z[j]), fabsf(z1));
        const float tb = atol + rtol * fmaxf(fabsf(zb[j]), fabsf(zb1));
        const float qz = dt * ez / tz, qb = dt * eb / tb;
        float p = qz * qz + qb * qb;
        if (FIN) {
          const float ze1 = fmaf(dt, se[j], ze[j]);
          const float te = atol + rtol * fmaxf(fabsf(ze[j]), fabsf(ze1));
          const float qe = dt * ee[j] / te;
          p = fmaf(qe, qe, p);
          if (ok[j]) {
            a.ZE[(size_t)cand * a.N + gi[j]] = ze1;
            a.FZE[(size_t)cand * a.N + gi[j]] = kel[j];
            a.MIDZE[(size_t)midw * a.N + gi[j]] = fmaf(dt, me[j], ze[j]);
          }
        }
        S.m.u[e] = ok[j] ? p : 0.f;
        if (ok[j]) {
          a.Z[(size_t)cand * a.N + gi[j]] = z1;
          a.ZB[(size_t)cand * a.N + gi[j]] = zb1;
          a.FZ[(size_t)cand * a.N + gi[j]] = kz[kAdjStages - 1][j];
          a.FZB[(size_t)cand * a.N + gi[j]] = kzb[kAdjStages - 1][j];
          a.MIDZB[(size_t)midw * a.N + gi[j]] = fmaf(dt, mb, zb[j]);
        }
      }
    }
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      float acc = 0.f;
#pragma unroll
      for (int qa = 0; qa < NA; ++qa) {
        float cs = 0.f, er = 0.f;
        for (int q = 0; q < kAdjStages; ++q) {
          cs = fmaf(c_tab.c_sol[q], skr[q * NA * R + qa * R + r], cs);
          er = fmaf(c_tab.c_err[q], skr[q * NA * R + qa * R + r], er);
        }
        const float r0 = sr[qa * R + r];
        const float r1 = fmaf(dt, cs, r0);
        const float tr = atol + rtol * fmaxf(fabsf(r0), fabsf(r1));
        const float q = dt * er / tr;
        acc += q * q;
        if (g.rank == 0 && row0 + r < a.B) {
          a.RS[((size_t)cand * NA + qa) * B + row0 + r] = r1;
          a.FR[((size_t)cand * NA + qa) * B + row0 + r] = skr[(kAdjStages - 1) * NA * R + qa * R + r];
        }
      }
      S.m.scal[r * 4 + 0] = acc;
    }
    __syncthreads();
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int r = warp; r < R; r += nw) {
        double s0 = 0.0;
        for (int d = lane; d < Dl; d += 32) s0 += (double)S.m.u[r * Dlp + d];
        s0 = warp_sum(s0);
        if (lane == 0 && row0 + r < a.B)
          a.rowpart0[(size_t)(row0 + r) * kC + g.rank] = s0 + (g.rank == 0 ? (double)S.m.scal[r * 4 + 0] : 0.0);
      }
    }
    __syncthreads();
    ENO_TICK(PH_COMBINE);
  }
  ENO_TICK_FLUSH();
  cluster.sync();
}

template <bool PDL = false, typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t st, Args... args) {
  static thread_local const void* s_last = nullptr;   // per kernel instantiation and device: set the attribute once
  static thread_local size_t s_smem = 0;
  static thread_local int s_dev = -1;
  int dev = -1;
  cudaGetDevice(&dev);
  if (s_last != (const void*)kernel || s_smem < smem || s_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    s_last = (const void*)kernel;
    s_smem = smem;
    s_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PDL ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ENO_PDL=0 switches the programmatic dependent launches off (A/B measurements)
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("ENO_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

constexpr int kClRows = 4;           // samples per cluster tile
constexpr int kClMaxClusters = 26;   // 5-CTA clusters co-resident on 148 SMs (measured with cudaOccupancyMaxActiveClusters)

inline int cl_grid(int B) {
  const int tiles = (B + kClRows - 1) / kClRows;
  return kC * (tiles < kClMaxClusters ? tiles : kClMaxClusters);
}

// The MNIST shape (D = 784, H = 100: BASELINE configs[0]) gets a fully specialised instantiation (all
// index arithmetic folds, the product loops unroll); any other shape runs the generic one.
inline bool cl_is_mnist(int D, int H) { return D == 784 && H == 100; }

template <int K, int DT, int HT, bool FIN = false>
inline cudaError_t cl_adj_rows_launch(int mode, const AdjArgs& a, int grid, size_t smem, float rhs_tau, cudaStream_t st) {
  switch (mode) {
    case ADJ_STEP:
      if (a.tc.X) {   // tensor-core factor sinks (adj_host.cuh decides)
        if (!a.dist && pdl_enabled()) return launch_cluster<true>(k_cl_adj_rows<kClRows, K, ADJ_STEP, DT, HT, FIN, true>, grid, smem, st, a, rhs_tau);
        return launch_cluster(k_cl_adj_rows<kClRows, K, ADJ_STEP, DT, HT, FIN, true>, grid, smem, st, a, rhs_tau);
      }
      if (!a.dist && pdl_enabled()) return launch_cluster<true>(k_cl_adj_rows<kClRows, K, ADJ_STEP, DT, HT, FIN>, grid, smem, st, a, rhs_tau);
      return launch_cluster(k_cl_adj_rows<kClRows, K, ADJ_STEP, DT, HT, FIN>, grid, smem, st, a, rhs_tau);
    case ADJ_INIT0: return launch_cluster(k_cl_adj_rows<kClRows, K, ADJ_INIT0, DT, HT, FIN>, grid, smem, st, a, rhs_tau);
    case ADJ_INIT1: return launch_cluster(k_cl_adj_rows<kClRows, K, ADJ_INIT1, DT, HT, FIN>, grid, smem, st, a, rhs_tau);
    default: return launch_cluster(k_cl_adj_rows<kClRows, K, ADJ_RHS, DT, HT, FIN>, grid, smem, st, a, rhs_tau);
  }
}

inline cudaError_t cl_adj_rows(int K, int mode, const AdjArgs& a, int grid, size_t smem, float rhs_tau, cudaStream_t st) {
  if (a.fin) {
    if (cl_is_mnist(a.w.D, a.w.H)) return cl_adj_rows_launch<2, 784, 100, true>(mode, a, grid, smem, rhs_tau, st);
    return cl_adj_rows_launch<2, 0, 0, true>(mode, a, grid, smem, rhs_tau, st);
  }
  if (cl_is_mnist(a.w.D, a.w.H)) {
    if (K == 0) return cl_adj_rows_launch<0, 784, 100>(mode, a, grid, smem, rhs_tau, st);
    if (K == 2) return cl_adj_rows_launch<2, 784, 100>(mode, a, grid, smem, rhs_tau, st);
    return cl_adj_rows_launch<3, 784, 100>(mode, a, grid, smem, rhs_tau, st);
  }
  if (K == 0) return cl_adj_rows_launch<0, 0, 0>(mode, a, grid, smem, rhs_tau, st);
  if (K == 2) return cl_adj_rows_launch<2, 0, 0>(mode, a, grid, smem, rhs_tau, st);
  return cl_adj_rows_launch<3, 0, 0>(mode, a, grid, smem, rhs_tau, st);
}

// forward: mode 0 iteration, 1 / 2 initial-step phases
inline cudaError_t cl_fwd_launch(int mode, const FwdArgs& a, int grid, size_t smem, cudaStream_t st) {
  const bool pdl = !a.dist && pdl_enabled();   // multi-GPU: measured slower with it (875 vs 944 steps/s at N = 2)
  if (cl_is_mnist(a.w.D, a.w.H)) {
    if (mode == 0 && pdl) return launch_cluster<true>(k_cl_fwd<kClRows, 0, 784, 100>, grid, smem, st, a);
    if (mode == 0) return launch_cluster(k_cl_fwd<kClRows, 0, 784, 100>, grid, smem, st, a);
    if (mode == 1) return launch_cluster(k_cl_fwd<kClRows, 1, 784, 100>, grid, smem, st, a);
    return launch_cluster(k_cl_fwd<kClRows, 2, 784, 100>, grid, smem, st, a);
  }
  if (mode == 0 && pdl) return launch_cluster<true>(k_cl_fwd<kClRows, 0, 0, 0>, grid, smem, st, a);
  if (mode == 0) return launch_cluster(k_cl_fwd<kClRows, 0, 0, 0>, grid, smem, st, a);
  if (mode == 1) return launch_cluster(k_cl_fwd<kClRows, 1, 0, 0>, grid, smem, st, a);
  return launch_cluster(k_cl_fwd<kClRows, 2, 0, 0>, grid, smem, st, a);
}

}  // namespace eno
