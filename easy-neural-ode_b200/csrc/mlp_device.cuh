// Device functions for the MLPDynamics family (reference: mnist.py:195-222) and its
// Taylor-mode coefficients (sol_recursive, mnist.py:105-131; closed forms in
// SURVEY.md Appendix B).  A CTA owns R consecutive samples; every vector below is a
// shared-memory array [R][D] or [R][H]; weights are streamed from L2 with coalesced
// 128-bit loads and split over K inside the CTA.
#pragma once
#include "eno_common.cuh"

namespace eno {

struct MlpWeights {
  const float* b1;    // [H]
  const float* w1t;   // [H]      row 0 of W1 (time input)
  const float* W1x;   // [D][H]   rows 1.. of W1
  const float* b2;    // [D]
  const float* w2t;   // [D]      row 0 of W2
  const float* W2x;   // [H][D]
  const float* W1xT;  // [H][D]   transpose of W1x (adjoint only)
  const float* W2xT;  // [D][H]   transpose of W2x (adjoint only)
  int D, H;
};

__host__ __device__ inline MlpWeights mlp_weights_from_flat(const float* p, int D, int H, const float* W1xT,
                                                             const float* W2xT) {
  MlpWeights w;
  w.b1 = p;
  w.w1t = p + H;
  w.W1x = p + H + H;
  w.b2 = p + H + (size_t)(D + 1) * H;
  w.w2t = w.b2 + D;
  w.W2x = w.b2 + D + D;
  w.W1xT = W1xT;
  w.W2xT = W2xT;
  w.D = D;
  w.H = H;
  return w;
}

// out[r][n] = epi(r, n, sum_k x[r][k] * W[k][n]).  W: global [K][N] row-major, N % 4 == 0.
// Threads are laid out as (k-group, 4-column group); partial sums meet in `red`
// ([KG][R][N] floats) and are combined in a fixed order.
template <int R, typename Epi>
__device__ __forceinline__ void gemv(const float* __restrict__ W, int K, int N, const float* __restrict__ x, int ldx,
                                     float* __restrict__ red, Epi epi) {
  const int NC4 = N >> 2;
  int KG = blockDim.x / NC4;
  if (KG > 32) KG = 32;
  const int kg = threadIdx.x / NC4;
  const int c4 = threadIdx.x - kg * NC4;
  if (kg < KG) {
    const int Kper = (K + KG - 1) / KG;
    const int k0 = kg * Kper;
    const int k1 = min(K, k0 + Kper);
    float4 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* __restrict__ Wp = reinterpret_cast<const float4*>(W) + c4;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      const float4 w0 = __ldg(Wp + (size_t)(k + 0) * NC4);
      const float4 w1 = __ldg(Wp + (size_t)(k + 1) * NC4);
      const float4 w2 = __ldg(Wp + (size_t)(k + 2) * NC4);
      const float4 w3 = __ldg(Wp + (size_t)(k + 3) * NC4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float x0 = x[r * ldx + k], x1 = x[r * ldx + k + 1], x2 = x[r * ldx + k + 2], x3 = x[r * ldx + k + 3];
        acc[r].x = fmaf(x0, w0.x, acc[r].x); acc[r].y = fmaf(x0, w0.y, acc[r].y);
        acc[r].z = fmaf(x0, w0.z, acc[r].z); acc[r].w = fmaf(x0, w0.w, acc[r].w);
        acc[r].x = fmaf(x1, w1.x, acc[r].x); acc[r].y = fmaf(x1, w1.y, acc[r].y);
        acc[r].z = fmaf(x1, w1.z, acc[r].z); acc[r].w = fmaf(x1, w1.w, acc[r].w);
        acc[r].x = fmaf(x2, w2.x, acc[r].x); acc[r].y = fmaf(x2, w2.y, acc[r].y);
        acc[r].z = fmaf(x2, w2.z, acc[r].z); acc[r].w = fmaf(x2, w2.w, acc[r].w);
        acc[r].x = fmaf(x3, w3.x, acc[r].x); acc[r].y = fmaf(x3, w3.y, acc[r].y);
        acc[r].z = fmaf(x3, w3.z, acc[r].z); acc[r].w = fmaf(x3, w3.w, acc[r].w);
      }
    }
    for (; k < k1; ++k) {
      const float4 w0 = __ldg(Wp + (size_t)k * NC4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float x0 = x[r * ldx + k];
        acc[r].x = fmaf(x0, w0.x, acc[r].x); acc[r].y = fmaf(x0, w0.y, acc[r].y);
        acc[r].z = fmaf(x0, w0.z, acc[r].z); acc[r].w = fmaf(x0, w0.w, acc[r].w);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) reinterpret_cast<float4*>(red)[(size_t)(kg * R + r) * NC4 + c4] = acc[r];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * N; idx += blockDim.x) {
    const int r = idx / N, n = idx - r * N;
    float sum = 0.f;
    for (int g = 0; g < KG; ++g) sum += red[(size_t)(g * R + r) * N + n];
    epi(r, n, sum);
  }
  __syncthreads();
}

// Shared-memory working set of one CTA for the Taylor passes.
struct MlpScratch {
  float* s;    // [R][D] sigmoid(z)
  float* u;    // [R][D] layer-1 input of the current pass (s, s1, s2)
  float* h;    // [R][H] sigmoid(a)
  float* a1;   // [R][H]
  float* h1;   // [R][H]
  float* a2;   // [R][H]
  float* h2;   // [R][H]
  float* red;  // [kRedFloats * R]
};

// f = MLP(z, t): z, f in shared memory [R][D].  Leaves s = sigmoid(z), h = sigmoid(a).
template <int R>
__device__ __forceinline__ void mlp_primal(const MlpWeights& w, float t, const float* z, float* f, const MlpScratch& sc) {
  const int D = w.D, H = w.H;
  for (int i = threadIdx.x; i < R * D; i += blockDim.x) sc.s[i] = sigmoidf_(z[i]);
  __syncthreads();
  gemv<R>(w.W1x, D, H, sc.s, D, sc.red, [&](int r, int n, float sum) {
    sc.h[r * H + n] = sigmoidf_(sum + t * __ldg(w.w1t + n) + __ldg(w.b1 + n));
  });
  gemv<R>(w.W2x, H, D, sc.h, H, sc.red, [&](int r, int n, float sum) {
    f[r * D + n] = sum + t * __ldg(w.w2t + n) + __ldg(w.b2 + n);
  });
}

// Second and third time-derivative of the flow (y1 = z'', y2 = z''') given the primal
// pass has run (s, h, f valid).  Optional global sinks receive the layer inputs that the
// parameter-gradient GEMM of the adjoint needs.
template <int R, int K>
__device__ __forceinline__ void mlp_taylor(const MlpWeights& w, const float* f, float* y1, float* y2,
                                           const MlpScratch& sc) {
  const int D = w.D, H = w.H;
  if (K >= 2) {
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const float s = sc.s[i];
      sc.u[i] = s * (1.f - s) * f[i];
    }
    __syncthreads();
    gemv<R>(w.W1x, D, H, sc.u, D, sc.red, [&](int r, int n, float sum) {
      const float h = sc.h[r * H + n];
      const float a1 = sum + __ldg(w.w1t + n);
      sc.a1[r * H + n] = a1;
      sc.h1[r * H + n] = h * (1.f - h) * a1;
    });
    gemv<R>(w.W2x, H, D, sc.h1, H, sc.red, [&](int r, int n, float sum) { y1[r * D + n] = sum + __ldg(w.w2t + n); });
  }
  if (K >= 3) {
    for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
      const float s = sc.s[i];
      const float sp = s * (1.f - s);
      const float fi = f[i];
      sc.u[i] = sp * (1.f - 2.f * s) * fi * fi + sp * y1[i];
    }
    __syncthreads();
    gemv<R>(w.W1x, D, H, sc.u, D, sc.red, [&](int r, int n, float sum) {
      const float h = sc.h[r * H + n];
      const float hp = h * (1.f - h);
      const float a1 = sc.a1[r * H + n];
      sc.a2[r * H + n] = sum;
      sc.h2[r * H + n] = hp * (1.f - 2.f * h) * a1 * a1 + hp * sum;
    });
    gemv<R>(w.W2x, H, D, sc.h2, H, sc.red, [&](int r, int n, float sum) { y2[r * D + n] = sum; });
  }
}

}  // namespace eno
