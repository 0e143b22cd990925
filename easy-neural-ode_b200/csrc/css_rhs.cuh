// Right-hand sides of the FFJORD tabular path (BASELINE config C2) as a pipeline of tensor-core GEMMs with fused
// epilogues.  Restates, for the CUDA path (nothing here is a port):
//   ConcatSquashLinear / NN_Dynamics ...... ffjord_tabular.py:140-193   u = (x W + b) * sigmoid(w_g t + b_g) + w_b t
//   softplus via _logaddexp ............... ffjord_tabular.py:92-103
//   ffjord_dynamics / ffjord2_dynamics .... ffjord_tabular.py:250-268   forward-mode J eps, div = sum eps * (J eps)
//   sol_recursive (fixed K = 2) ........... ffjord_tabular.py:116-136   d^2 z/dt^2 = J f + df/dt
//   aug_dynamics / all_aug_dynamics ....... ffjord_tabular.py:270-297
//   the adjoint's vjp of aug_dynamics ..... lib/ode.py:1498-1504
//
// Three row streams run through the SAME weights: 0 = primal x, 1 = tangent with eps ("e", no time tangent),
// 2 = tangent with f and time tangent 1 ("d", the Taylor term).  Forward: primal GEMMs (M = B), then the two
// tangent streams stacked (M = 2 Bp).  Reverse: tangent streams first (their cotangent reaches f through d_0 = f),
// then the primal stream; weight gradients contract all three streams over the batch (K = 3 Bp) in one GEMM per
// layer, reading the activations and cotangents as they lie in memory (MN-major operand descriptors, gemm_tc.cuh).
// Every operand a GEMM reads is stored pre-split into tf32 hi / lo planes by the epilogue that produced it.
#pragma once
#include "eno_common.cuh"
#include "gemm_tc.cuh"

namespace eno {
namespace css {

constexpr int kMaxLayers = 6;

enum Mode { MODE_PLAIN = 0, MODE_AUG = 1, MODE_ADJ = 2 };

// geometry + every device buffer of one RHS pipeline (all carved from the caller's workspace)
struct Net {
  int B, Bp, D, H, L;       // Bp: B rounded up to 128 (stream stride in the stacked operands)
  int nq;                   // tangent streams in use: 0 (plain f), 1 (e), 2 (e, d)
  int mode;
  int fin;                  // 1: aux outputs are (dfro, dkin) instead of r  (--reg_type fin)
  int din[kMaxLayers], dout[kMaxLayers], ldi[kMaxLayers], ldo[kMaxLayers];
  // parameters (flat vector, Haiku ravel order per layer: b | W | w_bias | b_gate | w_gate)
  const float *b[kMaxLayers], *W[kMaxLayers], *wb[kMaxLayers], *bg[kMaxLayers], *wg[kMaxLayers];
  size_t poff[kMaxLayers];  // offset of the layer's block in the flat vector
  // pre-split weights: W2 [2][din][ld(dout)] (reverse: N = din, K = dout), WT2 [2][dout][ld(din)] (forward)
  float *W2[kMaxLayers], *WT2[kMaxLayers];
  // per-evaluation gate vectors
  float *G[kMaxLayers], *G1[kMaxLayers], *CC[kMaxLayers];
  float* tval;              // [1] the time this evaluation runs at
  // activations
  float *X2[kMaxLayers];    // [2][3 Bp][ldi]    inputs of layer l: slot 0 x, 1 e, 2 d
  float *LIN[kMaxLayers];   // [B][ldo]
  float *S[kMaxLayers];     // [B][ldo]          sigmoid(u_l) = softplus'(u_l)
  float *A1[kMaxLayers];    // [2][B][ldo]       x1 W per tangent stream
  float *V[kMaxLayers];     // [2][B][ldo]       cotangent of u1_l per stream
  float *WQ[kMaxLayers];    // [2][B][ldo]       tangent streams' contribution to the cotangent of u_l
  float *WW[kMaxLayers];    // [B][ldo]          cotangent of u_l
  float *AB2[kMaxLayers];   // [2][3 Bp][ldo]    slot 0 lin_bar, 1 a1_bar(e), 2 a1_bar(d)
  float *WP[kMaxLayers];    // [3][din][dout]    split-K partials of the weight gradient
  float *CR;                // [layers][chunks][4][Hmax] column partial sums
  float *TC;                // [layers][Hmax] per-column contributions to t_bar
  int cr_chunks, hmax;
  float *F, *JE, *Y1;       // [B][ldD] f, J eps, d^2z/dt^2
  float *XB1;               // [2][B][ldD] cotangent of the tangent streams' inputs (e_0 = eps, d_0 = f)
  float *EBD;               // [B][ldD] direct eps cotangent  -p_bar * J eps
  float *YBDOT;             // [B][ldD] vjp_y
  float *DIV, *RR, *FRO, *KIN;   // [B]
  const float* eps;         // [B][D]
  int ldD;
  __host__ __device__ size_t xplane(int l) const { return (size_t)3 * Bp * ldi[l]; }
  __host__ __device__ size_t abplane(int l) const { return (size_t)3 * Bp * ldo[l]; }
};

constexpr int kW = tc::kEpiW;   // columns one epilogue call covers

__device__ __forceinline__ void softplus_sigmoid(float u, float& sp, float& sg) {
  // _logaddexp(u, 0) = max(u, 0) + log1p(exp(-|u|))  (ffjord_tabular.py:92-103); sigmoid from the same exponential.
  // The epilogues are instruction-issue bound, so the transcendental part uses the SFU approximations: exp(-|u|) in
  // (0, 1] to ~2 ulp, log(1 + e) in (0, log 2] to 2^-22 absolute, the reciprocal to 1 ulp - all at the level of the fp32
  // round-off of the surrounding arithmetic (single evaluations are checked to 3e-6 against fp64 in
  // tests/test_gpu_ffjord.py).  -DENO_CSS_ACCURATE_MATH restores expf / log1pf / IEEE division.
#ifdef ENO_CSS_ACCURATE_MATH
  const float e = expf(-fabsf(u));
  sp = fmaxf(u, 0.f) + log1pf(e);
  const float r = 1.0f / (1.0f + e);
#else
  const float e = __expf(-fabsf(u));
  sp = fmaxf(u, 0.f) + __logf(1.0f + e);
  const float r = __fdividef(1.0f, 1.0f + e);
#endif
  sg = (u >= 0.f) ? r : e * r;
}

// row-major hi / lo planes: kW consecutive elements of one row
__device__ __forceinline__ void put_split_row(float* base, size_t plane, size_t row_off, int n0, int ncols, const float (&x)[kW]) {
  float* hi = base + row_off + n0;
  float* lo = hi + plane;
  if (n0 + kW <= ncols) {
#pragma unroll
    for (int j = 0; j < kW; j += 4) {
      float4 h, l;
      tc::split_tf32(x[j], h.x, l.x); tc::split_tf32(x[j + 1], h.y, l.y);
      tc::split_tf32(x[j + 2], h.z, l.z); tc::split_tf32(x[j + 3], h.w, l.w);
      *reinterpret_cast<float4*>(hi + j) = h;
      *reinterpret_cast<float4*>(lo + j) = l;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kW; ++j)
      if (n0 + j < ncols) { float h, l; tc::split_tf32(x[j], h, l); hi[j] = h; lo[j] = l; }
  }
}
// transposed hi / lo planes [ncols][ldt]: element (n0 + j, col); the lanes of a warp hold consecutive `col`
__device__ __forceinline__ void put_split_col(float* base, size_t plane, int ldt, int col, int n0, int ncols, const float (&x)[kW]) {
#pragma unroll
  for (int j = 0; j < kW; ++j)
    if (n0 + j < ncols) {
      float h, l;
      tc::split_tf32(x[j], h, l);
      const size_t o = (size_t)(n0 + j) * ldt + col;
      base[o] = h;
      base[plane + o] = l;
    }
}
__device__ __forceinline__ void put_row(float* row, int n0, int ncols, const float (&x)[kW]) {
  if (n0 + kW <= ncols) {
#pragma unroll
    for (int j = 0; j < kW; j += 4) *reinterpret_cast<float4*>(row + n0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < kW; ++j)
      if (n0 + j < ncols) row[n0 + j] = x[j];
  }
}
__device__ __forceinline__ void get_row(const float* row, int n0, int ncols, float (&x)[kW]) {
  if (n0 + kW <= ncols) {
#pragma unroll
    for (int j = 0; j < kW; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(row + n0 + j);
      x[j] = t.x; x[j + 1] = t.y; x[j + 2] = t.z; x[j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kW; ++j) x[j] = (n0 + j < ncols) ? row[n0 + j] : 0.f;
  }
}

// ---- forward, primal stream: layer l ------------------------------------------------------------------------
struct EpiPrimal {
  Net n;
  int l;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    const int dout = n.dout[l], ldo = n.ldo[l];
    if (m >= n.B || n0 >= dout) return;
    const bool last = (l == n.L - 1);
    float lin[kW], s[kW], x[kW];
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      const int c = min(n0 + j, dout - 1);
      lin[j] = v[j] + n.b[l][c];
      const float u = lin[j] * n.G[l][c] + n.CC[l][c];
      if (last) { x[j] = u; s[j] = 0.f; }
      else softplus_sigmoid(u, x[j], s[j]);
    }
    if (n.nq == 2 || n.mode == MODE_ADJ) put_row(n.LIN[l] + (size_t)m * ldo, n0, dout, lin);
    if (!last) {
      if (n.nq > 0) put_row(n.S[l] + (size_t)m * ldo, n0, dout, s);
      put_split_row(n.X2[l + 1], n.xplane(l + 1), (size_t)m * n.ldi[l + 1], n0, dout, x);
    } else {
      put_row(n.F + (size_t)m * n.ldD, n0, dout, x);
      if (n.nq == 2) {   // d_0 = f feeds the Taylor stream
        put_split_row(n.X2[0], n.xplane(0), (size_t)(2 * n.Bp + m) * n.ldi[0], n0, dout, x);
      }
    }
  }
};

// ---- forward, tangent streams (rows: stream q = m / Bp, sample bm = m % Bp) ------------------------------------
struct EpiTangent {
  Net n;
  int l;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    const int dout = n.dout[l], ldo = n.ldo[l];
    const int q = m / n.Bp, bm = m - q * n.Bp;
    if (q >= n.nq || bm >= n.B || n0 >= dout) return;
    const bool last = (l == n.L - 1);
    float u1[kW];
    if (q == 1) {
      float lin[kW];
      get_row(n.LIN[l] + (size_t)bm * ldo, n0, dout, lin);
#pragma unroll
      for (int j = 0; j < kW; ++j) {
        const int c = min(n0 + j, dout - 1);
        u1[j] = v[j] * n.G[l][c] + (lin[j] * n.G1[l][c] + n.wb[l][c]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kW; ++j) u1[j] = v[j] * n.G[l][min(n0 + j, dout - 1)];
    }
    if (n.mode == MODE_ADJ) put_row(n.A1[l] + ((size_t)q * n.B + bm) * ldo, n0, dout, v);
    if (!last) {
      float s[kW];
      get_row(n.S[l] + (size_t)bm * ldo, n0, dout, s);
#pragma unroll
      for (int j = 0; j < kW; ++j) u1[j] *= s[j];
      put_split_row(n.X2[l + 1], n.xplane(l + 1), (size_t)((1 + q) * n.Bp + bm) * n.ldi[l + 1], n0, dout, u1);
    } else {
      put_row((q == 0 ? n.JE : n.Y1) + (size_t)bm * n.ldD, n0, dout, u1);
    }
  }
};

// ---- reverse, tangent streams: the GEMM of layer l produced x1_bar_l = a1_bar_l W_l^T; fold layer l-1 -----------
struct EpiTanRev {
  Net n;
  int l;   // >= 1
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    const int lm = l - 1;
    const int dout = n.dout[lm], ldo = n.ldo[lm];
    const int q = m / n.Bp, bm = m - q * n.Bp;
    if (q >= n.nq || bm >= n.B || n0 >= dout) return;
    float s[kW], a1[kW], ab[kW], wq[kW], vv[kW];
    get_row(n.S[lm] + (size_t)bm * ldo, n0, dout, s);
    get_row(n.A1[lm] + ((size_t)q * n.B + bm) * ldo, n0, dout, a1);
    if (q == 1) {
      float lin[kW];
      get_row(n.LIN[lm] + (size_t)bm * ldo, n0, dout, lin);
#pragma unroll
      for (int j = 0; j < kW; ++j) {
        const int c = min(n0 + j, dout - 1);
        a1[j] = a1[j] * n.G[lm][c] + (lin[j] * n.G1[lm][c] + n.wb[lm][c]);   // u1_{l-1}
      }
    } else {
#pragma unroll
      for (int j = 0; j < kW; ++j) a1[j] *= n.G[lm][min(n0 + j, dout - 1)];
    }
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      vv[j] = v[j] * s[j];                              // cotangent of u1_{l-1}
      wq[j] = v[j] * a1[j] * s[j] * (1.f - s[j]);       // -> cotangent of u_{l-1} through softplus''
      ab[j] = vv[j] * n.G[lm][min(n0 + j, dout - 1)];   // cotangent of a1_{l-1}
    }
    const size_t ro = ((size_t)q * n.B + bm) * ldo;
    put_row(n.V[lm] + ro, n0, dout, vv);
    put_row(n.WQ[lm] + ro, n0, dout, wq);
    put_split_row(n.AB2[lm], n.abplane(lm), (size_t)((1 + q) * n.Bp + bm) * ldo, n0, dout, ab);
  }
};

struct EpiTanRev0 {   // layer 0: cotangent of the stream inputs e_0 = eps, d_0 = f
  Net n;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    const int q = m / n.Bp, bm = m - q * n.Bp;
    if (q >= n.nq || bm >= n.B || n0 >= n.D) return;
    put_row(n.XB1 + ((size_t)q * n.B + bm) * n.ldD, n0, n.D, v);
  }
};

// ---- reverse, primal stream ---------------------------------------------------------------------------------
struct EpiPrimRev {
  Net n;
  int l;   // >= 1: x_bar_l arrives, fold layer l-1
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    const int lm = l - 1;
    const int dout = n.dout[lm], ldo = n.ldo[lm];
    if (m >= n.B || n0 >= dout) return;
    float s[kW], w[kW], lb[kW], t[kW];
    get_row(n.S[lm] + (size_t)m * ldo, n0, dout, s);
#pragma unroll
    for (int j = 0; j < kW; ++j) w[j] = v[j] * s[j];
    if (n.nq >= 1) {
      get_row(n.WQ[lm] + (size_t)m * ldo, n0, dout, t);
#pragma unroll
      for (int j = 0; j < kW; ++j) w[j] += t[j];
    }
    if (n.nq == 2) {
      get_row(n.WQ[lm] + ((size_t)n.B + m) * ldo, n0, dout, t);
#pragma unroll
      for (int j = 0; j < kW; ++j) w[j] += t[j];
      get_row(n.V[lm] + ((size_t)n.B + m) * ldo, n0, dout, t);
    }
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      const int c = min(n0 + j, dout - 1);
      lb[j] = w[j] * n.G[lm][c] + (n.nq == 2 ? t[j] * n.G1[lm][c] : 0.f);
    }
    put_row(n.WW[lm] + (size_t)m * ldo, n0, dout, w);
    put_split_row(n.AB2[lm], n.abplane(lm), (size_t)m * ldo, n0, dout, lb);
  }
};

struct EpiPrimRev0 {
  Net n;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int) const {
    if (m >= n.B || n0 >= n.D) return;
    put_row(n.YBDOT + (size_t)m * n.ldD, n0, n.D, v);
  }
};

// weight-gradient partial of one stream (split = stream slot)
struct EpiWgrad {
  float* WP;
  int din, dout;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kW], int split) const {
    if (m >= din || n0 >= dout) return;
    float* row = WP + ((size_t)split * din + m) * dout;
#pragma unroll
    for (int j = 0; j < kW; ++j)
      if (n0 + j < dout) row[n0 + j] = v[j];
  }
};

}  // namespace css
}  // namespace eno
