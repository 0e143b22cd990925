// Per-example adaptive solves in ONE launch: the `--vmap` NFE path of mnist.py:350-353,
//     jax.vmap(lambda y0, t, params: odeint(dynamics_wrap, y0, t, params)[1], (0, None, None)),
// i.e. every sample is its own dopri5 problem (state [D], error norm and initial step over its D elements,
// lib/ode.py:86-104, 518-538, 823-862) with its own t, dt, accept / reject sequence and NFE.
//
// Nothing couples the samples, so there is no grid-wide event at all: a 5-CTA cluster keeps its weight
// slices in shared memory (cl_device.cuh), takes a tile of R samples and runs their complete solves -
// initial step, every iteration, dense output at t1 - before moving to the next tile.  The per-sample
// norms are sums of per-CTA partials exchanged over distributed shared memory in rank order, so all
// CTAs hold bit-identical controller state and take the same branches.  Rows of a tile that finish early
// idle until the slowest row of the tile is done.
#pragma once
#include "cl_kernels.cuh"

namespace eno {

struct VmapArgs {
  MlpWeights w;
  int B;
  const float* y0;     // [B][D]
  const float* ts;     // [2] device
  float rtol, atol;
  int mxstep;          // > 0
  float* ys_out;       // [2][B][D]
  float* nfe_out;      // [B]
  int* iters_out;      // [B][2]  iterations, accepted steps  (or null)
  int* status_out;     // [B]     ENO_OK / ENO_MXSTEP / ENO_DT_UNDERFLOW (or null)
};

// f = MLP(z, t_r) with one time per row (cl_primal with a per-row t)
template <int R>
__device__ __forceinline__ void cl_primal_rows(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g, const ClWeights& w,
                                               const float* trow, const float* z, float* f, const ClScratch& sc) {
  const int H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
    const int d = i % Dlp;
    sc.s[i] = (d < Dl) ? sigmoidf_(z[i]) : 0.f;
  }
  __syncthreads();
  mv_n<R>(w.W1, g.ldw1, Dl, H, sc.s, Dlp, sc.red, kKG);
  allreduce_H<R>(cluster, cm, g, sc.red, kKG, H, nullptr, 0,
                 [&](int r, int n, float sum) { sc.h[r * H + n] = sigmoidf_(sum + trow[r] * w.w1t[n] + w.b1[n]); },
                 [&](int, int, float) {});
  mv_n<R>(w.W2, Dlp, H, Dl, sc.h, H, sc.red, kKGw);
  reduce_local<R>(sc.red, kKGw, Dl, [&](int r, int d, float sum) {
    f[r * Dlp + d] = sum + trow[r] * w.w2t[d] + w.b2[d];
  });
}

// out[r] = sum over the cluster (rank order) of sum_d q[r][d]; identical on every CTA.
// `stage` is a double[2][8] staging area in this CTA's shared memory, `phase` alternates per call.
template <int R>
__device__ __forceinline__ void cl_rows_allsum(cg::cluster_group& cluster, const ClGeom& g, const float* q, double* stage,
                                               int& phase, double* out) {
  static_assert(R <= 8, "staging area holds 8 rows");
  double* mine = stage + phase * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    double acc = 0.0;
    for (int d = lane; d < g.Dl; d += 32) acc += (double)q[r * g.Dlp + d];
    acc = warp_sum(acc);
    if (lane == 0) mine[r] = acc;
  }
  cluster.sync();
  if (threadIdx.x < R) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kC; ++k) s += cluster.map_shared_rank(mine, k)[threadIdx.x];
    out[threadIdx.x] = s;
  }
  phase ^= 1;
  __syncthreads();
}

template <int R, int DT, int HT>
__global__ void __launch_bounds__(kThreads, 1) k_cl_fwd_vmap(VmapArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float4 smem4[];
  const int D = DT > 0 ? DT : a.w.D, H = HT > 0 ? HT : a.w.H;
  constexpr int s = 7;   // dopri5 (odeint is always dopri5, lib/ode.py:746)
  const ClGeom g = cl_geom(D, H, (int)cluster.block_rank());
  ClSmem S = cl_carve<R>(reinterpret_cast<float*>(smem4), D, H, g, s, false);
  cl_load_weights(a.w, g, S.W1, S.W2, S.vec);
  ClWeights w{S.W1, S.W2, S.vec, S.vec + H, S.vec + 2 * H, S.vec + 2 * H + g.Dlp, D, H};
  const int Dl = g.Dl, Dlp = g.Dlp, RD = R * Dlp;
  const float rtol = a.rtol, atol = a.atol;
  const float t0 = a.ts[0], t1 = a.ts[1];
  const size_t N = (size_t)a.B * D;
  // per-row controller state, replicated (and bit-identical) in every CTA of the cluster
  float* t = S.aux;             // [R]
  float* dt = S.aux + R;        // [R]
  float* tst = S.aux + 2 * R;   // [R] stage time
  float* h0 = S.aux + 3 * R;
  float* d1 = S.aux + 4 * R;
  float* told = S.aux + 5 * R;  // start of the step just taken
  float* dtu = S.aux + 6 * R;   // step size just used
  int* iters = reinterpret_cast<int*>(S.aux + 7 * R);
  int* nacc = reinterpret_cast<int*>(S.aux + 8 * R);
  int* active = reinterpret_cast<int*>(S.aux + 9 * R);
  int* acc = reinterpret_cast<int*>(S.aux + 10 * R);
  int* any = reinterpret_cast<int*>(S.aux + 11 * R);
  double* sum0 = S.dred + 32;   // [8]   (S.dred[0..16) is the staging area of cl_rows_allsum)
  double* sum1 = S.dred + 40;
  int phase = 0;
  float* y1b = S.m.f;    // candidate state       (buffers cl_primal_rows leaves free or dead)
  float* ymb = S.m.u;    // mid-point
  float* qb = S.xz;      // squared scaled quantities for the norms
  __syncthreads();

  const int n_tiles = (a.B + R - 1) / R;
  const int n_clusters = gridDim.x / kC;
  for (int tile = blockIdx.x / kC; tile < n_tiles; tile += n_clusters) {
    const int row0 = tile * R;
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const int r = i / Dlp, d = i - r * Dlp;
      const bool ok = d < Dl && row0 + r < a.B;
      const size_t gi = (size_t)(row0 + r) * D + g.d0 + d;
      const float v = ok ? a.y0[gi] : 0.f;
      S.y[i] = v;
      if (ok) { a.ys_out[gi] = v; a.ys_out[N + gi] = v; }
    }
    if (threadIdx.x < R) tst[threadIdx.x] = t0;
    __syncthreads();

    // ---- initial step size, lib/ode.py:86-104 with order = 4, per sample ----
    cl_primal_rows<R>(cluster, S.cm, g, w, tst, S.y, S.k, S.m);
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const float y = S.y[i], f = S.k[i];
      const float scale = atol + fabsf(y) * rtol;
      const float q0 = y / scale, q1 = f / scale;
      qb[i] = q0 * q0;
      ymb[i] = q1 * q1;
    }
    __syncthreads();
    cl_rows_allsum<R>(cluster, g, qb, S.dred, phase, sum0);
    cl_rows_allsum<R>(cluster, g, ymb, S.dred, phase, sum1);
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      const float e0 = sqrtf((float)sum0[r]), e1 = sqrtf((float)sum1[r]);
      d1[r] = e1;
      h0[r] = (e0 < 1e-5f || e1 < 1e-5f) ? 1e-6f : 0.01f * e0 / e1;
      tst[r] = t0 + h0[r];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RD; i += blockDim.x) y1b[i] = fmaf(h0[i / Dlp], S.k[i], S.y[i]);
    __syncthreads();
    cl_primal_rows<R>(cluster, S.cm, g, w, tst, y1b, S.k + RD, S.m);
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const float scale = atol + fabsf(S.y[i]) * rtol;
      const float q = (S.k[RD + i] - S.k[i]) / scale;
      qb[i] = q * q;
    }
    __syncthreads();
    cl_rows_allsum<R>(cluster, g, qb, S.dred, phase, sum0);
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      const float d2 = sqrtf((float)sum0[r]) / h0[r];
      float h1;
      if (d1[r] <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0[r] * 1e-3f);
      else h1 = powf(0.01f / (d1[r] + d2), 1.0f / 5.0f);
      float v = fminf(100.f * h0[r], h1);
      if (isnan(h0[r]) || isnan(h1)) v = nanf("");
      dt[r] = v;
      t[r] = t0;
      iters[r] = 0;
      nacc[r] = 0;
    }
    __syncthreads();

    // ---- the samples' own adaptive loops, lib/ode.py:829-850 ----
    for (;;) {
      if (threadIdx.x == 0) {
        int n = 0;
        for (int r = 0; r < R; ++r) {
          const int on = (row0 + r < a.B) && (t[r] < t1) && (iters[r] < a.mxstep) && (dt[r] > 0.f);
          active[r] = on;
          n |= on;
        }
        any[0] = n;
      }
      __syncthreads();
      if (!any[0]) break;   // identical on all CTAs of the cluster: their controller state is bit-identical
      for (int st = 1; st < s; ++st) {
        for (int i = threadIdx.x; i < RD; i += blockDim.x) {
          float accv = 0.f;
          for (int j = 0; j < st; ++j) accv = fmaf(c_tab.beta[st - 1][j], S.k[j * RD + i], accv);
          S.xz[i] = fmaf(dt[i / Dlp], accv, S.y[i]);
        }
        if (threadIdx.x < R) tst[threadIdx.x] = t[threadIdx.x] + dt[threadIdx.x] * c_tab.alpha[st - 1];
        __syncthreads();
        cl_primal_rows<R>(cluster, S.cm, g, w, tst, S.xz, S.k + (size_t)st * RD, S.m);
      }
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        const float h = dt[i / Dlp];
        float as = 0.f, ae = 0.f, am = 0.f;
        for (int j = 0; j < s; ++j) {
          const float kj = S.k[j * RD + i];
          as = fmaf(c_tab.c_sol[j], kj, as);
          ae = fmaf(c_tab.c_err[j], kj, ae);
          am = fmaf(c_tab.c_mid[j], kj, am);
        }
        const float y = S.y[i];
        const float y1 = fmaf(h, as, y);
        const float err = h * ae;
        const float tol = atol + rtol * fmaxf(fabsf(y), fabsf(y1));
        const float q = err / tol;
        qb[i] = q * q;
        y1b[i] = y1;
        ymb[i] = fmaf(h, am, y);
      }
      __syncthreads();
      cl_rows_allsum<R>(cluster, g, qb, S.dred, phase, sum0);
      if (threadIdx.x < R) {
        const int r = threadIdx.x;
        int ok = 0;
        if (active[r]) {
          const float ratio = (float)(sum0[r] / (double)D);
          const float h = dt[r];
          ok = ratio <= 1.0f;
          dt[r] = next_dt(h, ratio, 5);
          iters[r] += 1;
          if (ok) {
            nacc[r] += 1;
            told[r] = t[r];
            dtu[r] = h;
            t[r] = t[r] + h;
          }
        }
        acc[r] = ok;
      }
      __syncthreads();
      // accepted rows: dense output at t1 from this step's quartic (lib/ode.py:56-84, 848-850; an
      // extrapolation until the step that crosses t1 overwrites it), then advance y and the FSAL derivative
      for (int i = threadIdx.x; i < RD; i += blockDim.x) {
        const int r = i / Dlp, d = i - r * Dlp;
        if (!acc[r]) continue;
        const float ya = S.y[i], yb = y1b[i], ym = ymb[i], dy0 = S.k[i], dy1 = S.k[(size_t)(s - 1) * RD + i];
        const float h = dtu[r];
        const float x = (t1 - told[r]) / (t[r] - told[r]);
        const float ca = -2.f * h * dy0 + 2.f * h * dy1 - 8.f * ya - 8.f * yb + 16.f * ym;
        const float cb = 5.f * h * dy0 - 3.f * h * dy1 + 18.f * ya + 14.f * yb - 32.f * ym;
        const float cc = -4.f * h * dy0 + h * dy1 - 11.f * ya - 5.f * yb + 16.f * ym;
        const float cd = h * dy0;
        float v = 0.f;
        v = v * x + ca; v = v * x + cb; v = v * x + cc; v = v * x + cd; v = v * x + ya;
        if (d < Dl && row0 + r < a.B) a.ys_out[N + (size_t)(row0 + r) * D + g.d0 + d] = v;
        S.y[i] = yb;
        S.k[i] = dy1;
      }
      __syncthreads();
    }
    if (g.rank == 0 && threadIdx.x < R && row0 + threadIdx.x < a.B) {
      const int r = threadIdx.x, row = row0 + r;
      a.nfe_out[row] = 2.0f + 6.0f * (float)iters[r];
      if (a.iters_out) { a.iters_out[2 * row] = iters[r]; a.iters_out[2 * row + 1] = nacc[r]; }
      if (a.status_out) a.status_out[row] = (t[r] >= t1) ? ENO_OK : (iters[r] >= a.mxstep ? ENO_MXSTEP : ENO_DT_UNDERFLOW);
    }
    __syncthreads();
  }
  cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

inline cudaError_t cl_fwd_vmap_launch(const VmapArgs& a, int grid, size_t smem, cudaStream_t st) {
  if (cl_is_mnist(a.w.D, a.w.H)) return launch_cluster(k_cl_fwd_vmap<kClRows, 784, 100>, grid, smem, st, a);
  return launch_cluster(k_cl_fwd_vmap<kClRows, 0, 0>, grid, smem, st, a);
}

}  // namespace eno
