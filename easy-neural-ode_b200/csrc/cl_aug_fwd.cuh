// Augmented FORWARD solve (evaluation path of the reference, mnist.py:333-334, 343-348, 561-575):
//     odeint(aug_dynamics,     (y, r),           ts, eps, params)   ENO_AUG_REG
//     odeint(all_aug_dynamics, (y, r, fro, kin), ts, eps, params)   ENO_AUG_ALL
// This is where the regulariser VALUE exists (on the training path the forward solve integrates the
// plain state only, lib/ode.py:993-1004).  dopri5 like every odeint() call (lib/ode.py:746).
//
// Per RHS: primal pass, Taylor passes (z'', z''' -> r' = mean_d((z^(K))^2), mnist.py:287-292), and for
// ENO_AUG_ALL one forward-mode pass with the Rademacher vector eps (fin_dynamics, mnist.py:294-300:
// eps_dy = J_f eps; fro' = mean(eps_dy^2), kin' = mean(dy^2)) - the Hutchinson-style estimator of
// ||J_f||_F^2 fused into the same kernel.  Cluster-resident weights, stage derivatives in registers,
// same controller protocol as k_cl_fwd.
#pragma once
#include "cl_kernels.cuh"

namespace eno {

struct AugFwdArgs {
  FwdArgs f;            // plain-state part: rings, partial sums, ys_out, controller
  const float* eps;     // [B][D] (ENO_AUG_ALL) or nullptr
  float* RSa;           // [3][3][B]  ring x aux x sample
  float* FRa;           // [3][3][B]
  float* MIDa;          // [2][3][B]
  float* aux_out;       // [T][n_aux][B]
  int n_aux;            // 1 (r) or 3 (r, fro, kin)
};

// cluster all-reduce of up to 4 per-sample scalars (partials in scal[r*4 + j]) -> out[r*4 + j]
template <int R>
__device__ __forceinline__ void allreduce_scal(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g, int H,
                                               const float* scal, float* out, int nscal) {
  float* mine = cm.part[cm.phase];
  __syncthreads();
  if (threadIdx.x < R * nscal) {
    const int r = threadIdx.x / nscal, j = threadIdx.x - r * nscal;
    mine[r * g.Hs + H + j] = scal[r * 4 + j];
  }
  cluster.sync();
  if (threadIdx.x < R * nscal) {
    const int r = threadIdx.x / nscal, j = threadIdx.x - r * nscal;
    const int o = r * g.Hs + H + j;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) s += cluster.map_shared_rank(mine, c)[o];
    out[r * 4 + j] = s;
  }
  cm.phase ^= 1;
  __syncthreads();
}

template <int R, int K, int AUG, int MODE, int DT, int HT>
__global__ void __launch_bounds__(kThreads, 1) k_cl_aug_fwd(AugFwdArgs A) {
  FwdArgs& a = A.f;
  Ctrl* c = a.ctrl;
  if (MODE == 0 && c->done) return;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float4 smem4[];
  const int D = DT > 0 ? DT : a.w.D, H = HT > 0 ? HT : a.w.H;
  const ClGeom g = cl_geom(D, H, (int)cluster.block_rank());
  ClSmem S = cl_carve<R>(reinterpret_cast<float*>(smem4), D, H, g, 0, true);
  cl_load_weights(a.w, g, S.W1, S.W2, S.vec);
  ClWeights w{S.W1, S.W2, S.vec, S.vec + H, S.vec + 2 * H, S.vec + 2 * H + g.Dlp, D, H};
  const int Dl = g.Dl, Dlp = g.Dlp, RD = R * Dlp;
  constexpr int EPT = 2;
  constexpr int NA = (AUG == ENO_AUG_ALL) ? 3 : 1;
  constexpr int S7 = kAdjStages;   // dopri5
  const float rtol = c->rtol, atol = c->atol;
  const float t = (MODE == 0) ? c->t : a.ts[0];
  const float dt = (MODE == 0) ? c->dt : c->h0;   // dopri5: never clipped
  const int cur = (MODE == 0) ? c->cur : 0, cand = c->cand, midw = c->mid_acc ^ 1;
  const size_t B = a.B;
  float* sums = S.aux;              // [R][4] all-reduced scalar derivatives of the current stage
  float* sst = S.aux + 4 * R;       // [R][4] scalar state
  float* ksc = S.aux + 8 * R;       // [S7][R][4] scalar stage derivatives
  __syncthreads();

  const int n_tiles = (a.B + R - 1) / R;
  const int n_clusters = gridDim.x / kC;
  for (int tile = blockIdx.x / kC; tile < n_tiles; tile += n_clusters) {
    const int row0 = tile * R;
    float y[EPT], ky[S7][EPT];
    size_t gi[EPT];
    bool ok[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int e = threadIdx.x + j * kThreads;
      const int r = e / Dlp, d = e - r * Dlp;
      ok[j] = e < RD && d < Dl && row0 + r < a.B;
      gi[j] = ok[j] ? (size_t)(row0 + r) * D + g.d0 + d : 0;
      y[j] = ok[j] ? (MODE == 1 ? a.y0[gi[j]] : a.Y[(size_t)cur * a.N + gi[j]]) : 0.f;
#pragma unroll
      for (int st = 0; st < S7; ++st) ky[st][j] = 0.f;
      if (MODE != 1) ky[0][j] = ok[j] ? a.F[(size_t)cur * a.N + gi[j]] : 0.f;
    }
    if (threadIdx.x < R * 4) {
      const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
      const bool okr = row0 + r < a.B && q < NA;
      sst[threadIdx.x] = (okr && MODE != 1) ? A.RSa[((size_t)cur * 3 + q) * B + row0 + r] : 0.f;   // initial aux state is 0
      ksc[threadIdx.x] = (okr && MODE != 1) ? A.FRa[((size_t)cur * 3 + q) * B + row0 + r] : 0.f;
    }

    constexpr int NEVAL = (MODE == 0) ? S7 - 1 : 1;
    for (int ev = 1; ev <= NEVAL; ++ev) {
      float tev;
      if (MODE == 0) tev = t + dt * c_tab.alpha[ev - 1];
      else if (MODE == 2) tev = t + dt;
      else tev = t;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          float v = y[j];
          if (MODE == 0) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < S7 - 1; ++q) acc = fmaf(c_tab.beta[ev - 1][q], ky[q][j], acc);
            v = fmaf(dt, acc, v);
          } else if (MODE == 2) {
            v = fmaf(dt, ky[0][j], v);
          }
          S.xz[e] = v;
        }
      }
      __syncthreads();
      cl_aug_forward<R, K>(cluster, S.cm, g, w, tev, S.xz, S.m, (const FactorDstT<false>*)nullptr, row0, a.B);   // scal[.][0] = sum_d yK^2
      if (AUG == ENO_AUG_ALL) {
        // forward-mode pass with tangent eps: eps_dy = J_f eps  (fin_dynamics, mnist.py:294-300)
        for (int i = threadIdx.x; i < RD; i += blockDim.x) {
          const int r = i / Dlp, d = i - r * Dlp;
          const float s = S.m.s[i];
          const bool in = d < Dl && row0 + r < a.B;
          S.m.u[i] = in ? s * (1.f - s) * __ldg(A.eps + (size_t)(row0 + r) * D + g.d0 + d) : 0.f;
        }
        __syncthreads();
        mv_n<R>(w.W1, g.ldw1, Dl, H, S.m.u, Dlp, S.m.red, kKG);
        allreduce_H<R>(cluster, S.cm, g, S.m.red, kKG, H, nullptr, 0,
                       [&](int r, int n, float sum) {
                         const float h = S.m.h[r * H + n];
                         S.m.h2[r * H + n] = h * (1.f - h) * sum;
                       },
                       [&](int, int, float) {});
        mv_n<R>(w.W2, Dlp, H, Dl, S.m.h2, H, S.m.red, kKGw);
        reduce_local<R>(S.m.red, kKGw, Dl, [&](int r, int d, float sum) { S.m.y2[r * Dlp + d] = sum; });
        cl_sumsq_local<R>(S.m.y2, g, S.m.scal, 1);
        cl_sumsq_local<R>(S.m.f, g, S.m.scal, 2);
        __syncthreads();
      }
      allreduce_scal<R>(cluster, S.cm, g, H, S.m.scal, sums, NA);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        const float vf = (e < RD) ? S.m.f[e] : 0.f;
#pragma unroll
        for (int st = 1; st < S7; ++st)
          if (st == ev) ky[st][j] = vf;
      }
      if (threadIdx.x < R * 4) {
        const int q = threadIdx.x & 3;
        float v = (q < NA) ? sums[threadIdx.x] / (float)D : 0.f;   // mean over d
        if (K < 2 && q == 0) v = 0.f;                               // --reg none: r' = 0
        ksc[ev * R * 4 + threadIdx.x] = v;
      }
      __syncthreads();
    }

    // ---- mode-specific tails ------------------------------------------------------------
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          const float f = ky[1][j];
          const float scale = atol + fabsf(y[j]) * rtol;
          const float q0 = y[j] / scale, q1 = f / scale;
          S.m.u[e] = ok[j] ? q0 * q0 : 0.f;
          S.xz[e] = ok[j] ? q1 * q1 : 0.f;
          if (ok[j]) {
            a.F[gi[j]] = f;
            a.Y[gi[j]] = y[j];
            a.ys_out[gi[j]] = y[j];
          }
        }
      }
      if (threadIdx.x < R) {
        // aux state starts at 0: scale = atol; d0 part 0, d1 part sum_q (k/atol)^2
        float s1 = 0.f;
        for (int q = 0; q < NA; ++q) {
          const float k = ksc[R * 4 + threadIdx.x * 4 + q];
          const float v = k / atol;
          s1 += v * v;
          if (g.rank == 0 && row0 + threadIdx.x < a.B) {
            A.FRa[(size_t)q * B + row0 + threadIdx.x] = k;
            A.RSa[(size_t)q * B + row0 + threadIdx.x] = 0.f;
            A.aux_out[(size_t)q * B + row0 + threadIdx.x] = 0.f;
          }
        }
        S.m.rdot[threadIdx.x] = s1;
      }
      __syncthreads();
      cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, nullptr);
      cl_row_sums<R>(S.xz, g, row0, a.B, a.rowpart1, S.m.rdot);
      __syncthreads();
      continue;
    }
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int e = threadIdx.x + j * kThreads;
        if (e < RD) {
          const float scale = atol + fabsf(y[j]) * rtol;
          const float q = (ky[1][j] - ky[0][j]) / scale;
          S.m.u[e] = ok[j] ? q * q : 0.f;
        }
      }
      if (threadIdx.x < R) {
        float s1 = 0.f;
        for (int q = 0; q < NA; ++q) {
          const float sc = atol + fabsf(sst[threadIdx.x * 4 + q]) * rtol;
          const float v = (ksc[R * 4 + threadIdx.x * 4 + q] - ksc[threadIdx.x * 4 + q]) / sc;
          s1 += v * v;
        }
        S.m.rdot[threadIdx.x] = s1;
      }
      __syncthreads();
      cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, S.m.rdot);
      __syncthreads();
      continue;
    }
    // MODE 0: candidate, error, mid-point
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int e = threadIdx.x + j * kThreads;
      if (e < RD) {
        float as = 0.f, ae = 0.f, am = 0.f;
#pragma unroll
        for (int q = 0; q < S7; ++q) {
          as = fmaf(c_tab.c_sol[q], ky[q][j], as);
          ae = fmaf(c_tab.c_err[q], ky[q][j], ae);
          am = fmaf(c_tab.c_mid[q], ky[q][j], am);
        }
        const float y1 = fmaf(dt, as, y[j]);
        const float tol = atol + rtol * fmaxf(fabsf(y[j]), fabsf(y1));
        const float q = dt * ae / tol;
        S.m.u[e] = ok[j] ? q * q : 0.f;
        if (ok[j]) {
          a.Y[(size_t)cand * a.N + gi[j]] = y1;
          a.F[(size_t)cand * a.N + gi[j]] = ky[S7 - 1][j];
          a.MID[(size_t)midw * a.N + gi[j]] = fmaf(dt, am, y[j]);
        }
      }
    }
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      float s1 = 0.f;
      for (int q = 0; q < NA; ++q) {
        float as = 0.f, ae = 0.f, am = 0.f;
        for (int st = 0; st < S7; ++st) {
          const float k = ksc[st * R * 4 + r * 4 + q];
          as = fmaf(c_tab.c_sol[st], k, as);
          ae = fmaf(c_tab.c_err[st], k, ae);
          am = fmaf(c_tab.c_mid[st], k, am);
        }
        const float v0 = sst[r * 4 + q];
        const float v1 = fmaf(dt, as, v0);
        const float tol = atol + rtol * fmaxf(fabsf(v0), fabsf(v1));
        const float qq = dt * ae / tol;
        s1 += qq * qq;
        if (g.rank == 0 && row0 + r < a.B) {
          A.RSa[((size_t)cand * 3 + q) * B + row0 + r] = v1;
          A.FRa[((size_t)cand * 3 + q) * B + row0 + r] = ksc[(S7 - 1) * R * 4 + r * 4 + q];
          A.MIDa[((size_t)midw * 3 + q) * B + row0 + r] = fmaf(dt, am, v0);
        }
      }
      S.m.rdot[r] = s1;
    }
    __syncthreads();
    cl_row_sums<R>(S.m.u, g, row0, a.B, a.rowpart0, S.m.rdot);
    __syncthreads();
  }
  cluster.sync();

  if (a.dist) return;
  if (!elect_last_block(&c->ticket)) return;
  __shared__ double dred[kThreads];
  const double s0 = block_sum_global(a.rp0_all, a.n_rp, dred);
  __syncthreads();
  double s1 = 0.0;
  if (MODE == 1) s1 = block_sum_global(a.rp1_all, a.n_rp, dred);
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

// dense output of the auxiliary scalars (same quartic as k_fwd_out)
__global__ void k_aux_out(AugFwdArgs A) {
  const FwdArgs& a = A.f;
  const Ctrl* c = a.ctrl;
  const int ob = c->out_begin, oe = c->out_end;
  if (oe <= ob) return;
  const float t = c->t, last_t = c->last_t, dt = c->step_dt;
  const bool degenerate = (c->n_accepted == 0);
  const size_t B = a.B;
  const int n = A.n_aux * a.B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int q = i / a.B, r = i - q * a.B;
    float ca, cb, cc, cd, ce;
    if (degenerate) {
      ca = cb = cc = cd = ce = 0.f;
    } else {
      const float y0 = A.RSa[((size_t)c->prev * 3 + q) * B + r], y1 = A.RSa[((size_t)c->cur * 3 + q) * B + r];
      const float ym = A.MIDa[((size_t)c->mid_acc * 3 + q) * B + r];
      const float dy0 = A.FRa[((size_t)c->prev * 3 + q) * B + r], dy1 = A.FRa[((size_t)c->cur * 3 + q) * B + r];
      ca = -2.f * dt * dy0 + 2.f * dt * dy1 - 8.f * y0 - 8.f * y1 + 16.f * ym;
      cb = 5.f * dt * dy0 - 3.f * dt * dy1 + 18.f * y0 + 14.f * y1 - 32.f * ym;
      cc = -4.f * dt * dy0 + dt * dy1 - 11.f * y0 - 5.f * y1 + 16.f * ym;
      cd = dt * dy0;
      ce = y0;
    }
    for (int o = ob; o < oe; ++o) {
      const float x = (a.ts[o] - last_t) / (t - last_t);
      float v = 0.f;
      v = v * x + ca; v = v * x + cb; v = v * x + cc; v = v * x + cd; v = v * x + ce;
      A.aux_out[((size_t)o * A.n_aux + q) * B + r] = v;
    }
  }
}

template <int K, int AUG, int DT, int HT>
inline cudaError_t cl_aug_fwd_launch3(int mode, const AugFwdArgs& A, int grid, size_t smem, cudaStream_t st) {
  if (mode == 0) return launch_cluster(k_cl_aug_fwd<kClRows, K, AUG, 0, DT, HT>, grid, smem, st, A);
  if (mode == 1) return launch_cluster(k_cl_aug_fwd<kClRows, K, AUG, 1, DT, HT>, grid, smem, st, A);
  return launch_cluster(k_cl_aug_fwd<kClRows, K, AUG, 2, DT, HT>, grid, smem, st, A);
}

template <int DT, int HT>
inline cudaError_t cl_aug_fwd_launch2(int K, int aug, int mode, const AugFwdArgs& A, int grid, size_t smem, cudaStream_t st) {
  if (aug == ENO_AUG_ALL) {
    if (K == 2) return cl_aug_fwd_launch3<2, ENO_AUG_ALL, DT, HT>(mode, A, grid, smem, st);
    if (K == 3) return cl_aug_fwd_launch3<3, ENO_AUG_ALL, DT, HT>(mode, A, grid, smem, st);
    return cl_aug_fwd_launch3<0, ENO_AUG_ALL, DT, HT>(mode, A, grid, smem, st);
  }
  if (K == 2) return cl_aug_fwd_launch3<2, ENO_AUG_REG, DT, HT>(mode, A, grid, smem, st);
  if (K == 3) return cl_aug_fwd_launch3<3, ENO_AUG_REG, DT, HT>(mode, A, grid, smem, st);
  return cl_aug_fwd_launch3<0, ENO_AUG_REG, DT, HT>(mode, A, grid, smem, st);
}

inline cudaError_t cl_aug_fwd_launch(int K, int aug, int mode, const AugFwdArgs& A, int grid, size_t smem, cudaStream_t st) {
  if (cl_is_mnist(A.f.w.D, A.f.w.H)) return cl_aug_fwd_launch2<784, 100>(K, aug, mode, A, grid, smem, st);
  return cl_aug_fwd_launch2<0, 0>(K, aug, mode, A, grid, smem, st);
}

}  // namespace eno
