// Per-sample part of one right-hand side of the adjoint ODE (lib/ode.py:1498-1504) for
// rev_func = aug_dynamics('our') of mnist.py:302-308, i.e.
//     (y, r) -> (f(y, t), mean_d((d^K y/dt^K)^2))
// pulled back with cotangent (y_bar, r_bar).  Hand-derived reverse sweep of the three
// Taylor passes, pass 3 -> 2 -> 1.
//
// Everything that is per-sample is finished here; what is a sum over the batch
// (params_bar and t0_bar derivatives) leaves this function as FACTORS written to global
// memory, and is contracted over the batch by the parameter-gradient kernel:
//     dW1x = S^T  A      with S = [s; s1; s2]   (3B x D),  A = [a_bar; a1_bar; a2_bar] (3B x H)
//     dW2x = Hh^T G      with Hh = [h; h1; h2]  (3B x H),  G = [f_bar; y1_bar; y2_bar] (3B x D)
//     db1 = colsum(a_bar), dw1t = colsum(t*a_bar + a1_bar), db2 = colsum(f_bar),
//     dw2t = colsum(t*f_bar + y1_bar),  t_bar = sum_rows(f_bar . w2t + a_bar . w1t)
#pragma once
#include "mlp_device.cuh"

namespace eno {

// Global sinks of one RHS evaluation (one "slot"); each is [3][B][*] with the pass index
// leading so that [3B][*] is a contiguous K-major GEMM operand.
struct FactorSlot {
  float* S;     // [3][B][D]
  float* A;     // [3][B][H]
  float* Hh;    // [3][B][H]
  float* G;     // [3][B][D]
  float* tbar;  // [B]
};

// Tensor-core sinks (cluster path, ADJ_STEP): the same factors written TRANSPOSED and pre-split into tf32 hi / lo
// planes, i.e. directly as the K-major operands of the batched tcgen05 contraction (k_gemm_tf32x3, gemm_tc.cuh):
//   X [2 planes][6 stages][2][dpad][ldk] : which 0 = S^T,  which 1 = G^T      (rows = state columns d)
//   Y [2 planes][6 stages][2][hpad][ldk] : which 0 = A^T,  which 1 = Hh^T     (rows = hidden columns h)
// with the contraction index k = pass * Bk + sample along the row (Bk = B rounded up to 4; the padding holds zeros).
struct FactorTc {
  float* X;     // null: the ADJ_STEP launches use the row-major sinks (streamed path)
  float* Y;
  unsigned xplane, yplane;   // elements between the hi and the lo plane
  int ldk, dpad, hpad, Bk;
};

// What the cluster kernels write one stage's factors through: the row-major sinks (TC = false; the set-up modes and
// eno_rhs, contracted by k_adj_pgrad) or the tensor-core sinks (TC = true; ADJ_STEP).  `rm` and `tc` point into the
// kernel's parameter block (constant bank: their fields cost no registers).
template <bool TC>
struct FactorDstT;
template <>
struct FactorDstT<false> {
  static constexpr bool kTc = false;
  const FactorSlot* rm;
  __device__ __forceinline__ FactorDstT(const FactorSlot& s, const FactorTc&, int) : rm(&s) {}
};
template <>
struct FactorDstT<true> {
  static constexpr bool kTc = true;
  const FactorSlot* rm;      // only for the per-sample t_bar pieces
  const FactorTc* tc;
  float *tX, *tY;            // this stage's [2][dpad][ldk] / [2][hpad][ldk] hi-plane blocks
  __device__ __forceinline__ FactorDstT(const FactorSlot& s, const FactorTc& t, int stage) : rm(&s), tc(&t) {
    tX = t.X + (size_t)stage * 2 * t.dpad * t.ldk;
    tY = t.Y + (size_t)stage * 2 * t.hpad * t.ldk;
  }
};

enum FactorMat { FM_S = 0, FM_G = 1, FM_A = 0, FM_HH = 1 };   // `which` index of the D-type / H-type factor

struct AdjScratch {
  MlpScratch sc;   // s, u, h, a1, h1, a2, h2, red
  float *f, *y1, *y2;      // [R][D]
  float *gf, *gy1, *zacc;  // [R][D]
  float *ga, *ga1, *ga2;   // [R][H]
  float* rdot;             // [R]
};

template <int R>
__device__ __forceinline__ void store_rows(float* dst /*[B][W] slice*/, const float* src /*smem [R][W]*/, int W,
                                           int row0, int B) {
  for (int i = threadIdx.x; i < R * W; i += blockDim.x)
    if (row0 + i / W < B) dst[(size_t)row0 * W + i] = src[i];
}

// Forward half: f, (y1, y2), r_dot.  Factor sinks S/Hh are written when slot != nullptr.
template <int R, int K>
__device__ __forceinline__ void aug_forward(const MlpWeights& w, float t, const float* z, const AdjScratch& m,
                                            const FactorSlot* slot, int row0, int B) {
  const int D = w.D, H = w.H;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  mlp_primal<R>(w, t, z, m.f, m.sc);
  if (slot) {
    store_rows<R>(slot->S, m.sc.s, D, row0, B);
    store_rows<R>(slot->Hh, m.sc.h, H, row0, B);
  }
  if (K >= 2) {
    // pass 2 (first-order coefficients); u = s1
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const float s = m.sc.s[i];
      m.sc.u[i] = s * (1.f - s) * m.f[i];
    }
    __syncthreads();
    if (slot) store_rows<R>(slot->S + BD, m.sc.u, D, row0, B);
    gemv<R>(w.W1x, D, H, m.sc.u, D, m.sc.red, [&](int r, int n, float sum) {
      const float h = m.sc.h[r * H + n];
      const float a1 = sum + __ldg(w.w1t + n);
      m.sc.a1[r * H + n] = a1;
      m.sc.h1[r * H + n] = h * (1.f - h) * a1;
    });
    if (slot) store_rows<R>(slot->Hh + BH, m.sc.h1, H, row0, B);
    gemv<R>(w.W2x, H, D, m.sc.h1, H, m.sc.red, [&](int r, int n, float sum) { m.y1[r * D + n] = sum + __ldg(w.w2t + n); });
  }
  if (K >= 3) {
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const float s = m.sc.s[i];
      const float sp = s * (1.f - s);
      const float fi = m.f[i];
      m.sc.u[i] = sp * (1.f - 2.f * s) * fi * fi + sp * m.y1[i];
    }
    __syncthreads();
    if (slot) store_rows<R>(slot->S + 2 * BD, m.sc.u, D, row0, B);
    gemv<R>(w.W1x, D, H, m.sc.u, D, m.sc.red, [&](int r, int n, float sum) {
      const float h = m.sc.h[r * H + n];
      const float hp = h * (1.f - h);
      const float a1 = m.sc.a1[r * H + n];
      m.sc.a2[r * H + n] = sum;
      m.sc.h2[r * H + n] = hp * (1.f - 2.f * h) * a1 * a1 + hp * sum;
    });
    if (slot) store_rows<R>(slot->Hh + 2 * BH, m.sc.h2, H, row0, B);
    gemv<R>(w.W2x, H, D, m.sc.h2, H, m.sc.red, [&](int r, int n, float sum) { m.y2[r * D + n] = sum; });
  }
  // r_dot[r] = mean_d(yK^2)
  {
    const float* yk = (K >= 3) ? m.y2 : m.y1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
      double acc = 0.0;
      if (K >= 2)
        for (int d = lane; d < D; d += 32) {
          const float v = yk[r * D + d];
          acc += (double)(v * v);
        }
      acc = warp_sum(acc);
      if (lane == 0) m.rdot[r] = (float)(acc / (double)D);
    }
    __syncthreads();
  }
}

// Reverse half.  zb: cotangent of f [R][D] (shared); rb[r]: cotangent of r_dot.
// Leaves: zacc = vjp wrt z; factors A, G and the per-sample t_bar in the slot.
template <int R, int K>
__device__ __forceinline__ void aug_reverse(const MlpWeights& w, const float* zb, const float* rb /*smem [R]*/,
                                            const AdjScratch& m, const FactorSlot& slot, int row0, int B) {
  const int D = w.D, H = w.H;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  const float invD2 = 2.0f / (float)D;

  if (K >= 3) {
    // y2_bar = r_bar * 2*y2/D   (kept in u, which the forward no longer needs)
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) m.sc.u[i] = rb[i / D] * invD2 * m.y2[i];
    __syncthreads();
    store_rows<R>(slot.G + 2 * BD, m.sc.u, D, row0, B);
    gemv<R>(w.W2xT, D, H, m.sc.u, D, m.sc.red, [&](int r, int n, float hb2) {
      const float h = m.sc.h[r * H + n];
      const float hp = h * (1.f - h);
      const float hpp = hp * (1.f - 2.f * h);
      const float hppp = hp * (1.f - 6.f * h + 6.f * h * h);
      const float a1 = m.sc.a1[r * H + n], a2 = m.sc.a2[r * H + n];
      m.ga2[r * H + n] = hb2 * hp;
      m.ga1[r * H + n] = hb2 * 2.f * hpp * a1;
      m.ga[r * H + n] = hb2 * (hppp * a1 * a1 + hpp * a2);
    });
    store_rows<R>(slot.A + 2 * BH, m.ga2, H, row0, B);
    gemv<R>(w.W1xT, H, D, m.ga2, H, m.sc.red, [&](int r, int n, float sb2) {
      const int i = r * D + n;
      const float s = m.sc.s[i];
      const float sp = s * (1.f - s);
      const float spp = sp * (1.f - 2.f * s);
      const float sppp = sp * (1.f - 6.f * s + 6.f * s * s);
      const float fi = m.f[i];
      m.zacc[i] = sb2 * (sppp * fi * fi + spp * m.y1[i]);
      m.gf[i] = sb2 * 2.f * spp * fi;
      m.gy1[i] = sb2 * sp;
    });
  } else if (K == 2) {
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      m.gy1[i] = rb[i / D] * invD2 * m.y1[i];
      m.gf[i] = 0.f;
      m.zacc[i] = 0.f;
    }
    for (int i = threadIdx.x; i < R * H; i += blockDim.x) { m.ga[i] = 0.f; m.ga1[i] = 0.f; }
    __syncthreads();
  }
  if (K >= 2) {
    store_rows<R>(slot.G + BD, m.gy1, D, row0, B);
    gemv<R>(w.W2xT, D, H, m.gy1, D, m.sc.red, [&](int r, int n, float hb1) {
      const float h = m.sc.h[r * H + n];
      const float hp = h * (1.f - h);
      const float hpp = hp * (1.f - 2.f * h);
      m.ga[r * H + n] += hb1 * hpp * m.sc.a1[r * H + n];
      m.ga1[r * H + n] += hb1 * hp;
    });
    store_rows<R>(slot.A + BH, m.ga1, H, row0, B);
    gemv<R>(w.W1xT, H, D, m.ga1, H, m.sc.red, [&](int r, int n, float sb1) {
      const int i = r * D + n;
      const float s = m.sc.s[i];
      const float sp = s * (1.f - s);
      const float spp = sp * (1.f - 2.f * s);
      m.zacc[i] += sb1 * spp * m.f[i];
      m.gf[i] += sb1 * sp;
    });
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) m.gf[i] += zb[i];
  } else {
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) { m.gf[i] = zb[i]; m.zacc[i] = 0.f; }
    for (int i = threadIdx.x; i < R * H; i += blockDim.x) m.ga[i] = 0.f;
  }
  __syncthreads();
  store_rows<R>(slot.G, m.gf, D, row0, B);
  gemv<R>(w.W2xT, D, H, m.gf, D, m.sc.red, [&](int r, int n, float hb) {
    const float h = m.sc.h[r * H + n];
    m.ga[r * H + n] += hb * h * (1.f - h);
  });
  store_rows<R>(slot.A, m.ga, H, row0, B);
  gemv<R>(w.W1xT, H, D, m.ga, H, m.sc.red, [&](int r, int n, float sb) {
    const int i = r * D + n;
    const float s = m.sc.s[i];
    m.zacc[i] += sb * s * (1.f - s);
  });
  // per-sample contribution to vjp wrt t: f_bar . w2t + a_bar . w1t
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
      double acc = 0.0;
      for (int d = lane; d < D; d += 32) acc += (double)(m.gf[r * D + d] * __ldg(w.w2t + d));
      for (int n = lane; n < H; n += 32) acc += (double)(m.ga[r * H + n] * __ldg(w.w1t + n));
      acc = warp_sum(acc);
      if (lane == 0 && row0 + r < B) slot.tbar[row0 + r] = (float)acc;
    }
  }
  __syncthreads();
}

}  // namespace eno
