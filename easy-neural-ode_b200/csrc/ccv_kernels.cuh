// Right-hand sides of the FFJORD image path (BASELINE config C4, ffjord_mnist.py) - FORWARD closures.  Restates, for the
// CUDA path (nothing here is a port):
//   ConcatConv2D / NN_Dynamics ............ ffjord_mnist.py:123-135, 186-232   3x3 convolutions (stride 1, zero padding 1)
//                                           on concat([t, x]) along the channel axis, softplus between the layers
//   ffjord_dynamics / ffjord2_dynamics .... ffjord_mnist.py:291-309            forward-mode J eps, div = sum eps * (J eps)
//   sol_recursive (fixed K = 2) ........... ffjord_mnist.py:99-119             d^2 z/dt^2 = J f + df/dt
//   aug / all_aug closures (values) ....... ffjord_mnist.py:311-338
//
// Layout: every activation lives on a ZERO-PADDED position grid, (H + 2) x (W + 2) positions per image, C channels per
// position, row-major [position][channel] - so tap (dy, dx) of a 3x3 convolution is the constant row shift
// dy * (W + 2) + dx, halo rows hold zeros (they are never written), and the two 64 -> 64 layers are 9-tap K loops of the
// tcgen05 3xTF32 kernel (gemm_tc.cuh: launch_conv3x3) over tf32 hi / lo planes written by the previous layer's epilogue.
// The time channel of ConcatConv2D is a constant inside the zero padding: its contribution is t * T_l[class][c_out],
// T = the sum of the time-channel weights over the taps that fall inside the image (9 border classes), added in the
// epilogue; in the Taylor tangent stream (dt = 1) it is T itself.  The 1 -> C first layer and the C -> 1 last layer are
// direct kernels.  Three row streams run through the same weights (css_rhs.cuh): 0 primal, 1 eps tangent, 2 Taylor
// tangent (direction f).  The per-sample scalars and the slot assembly reuse k_css_seed / k_css_assemble through a
// css::Net that only names the buffers those two kernels read in the forward modes.
#pragma once
#include "css_kernels.cuh"

namespace eno {
namespace ccv {

constexpr int kLayers = 4;   // hidden_dims (C, C, C) + the image channel (ffjord_mnist.py:190-192)

struct CNet {
  int B, Hh, Ww, C;          // images, image height / width, hidden channels (32 or 64)
  int HP, WP, PP;            // padded grid: Hh + 2, Ww + 2, positions per image
  int R, Rp;                 // positions of one stream (B * PP), rounded up to 128 (stream stride in the stacked operands)
  int D, ldD;                // Hh * Ww (one image channel), row pitch of the [B][D] outputs
  int nq, mode;
  const float *b[kLayers], *W[kLayers];   // Haiku layout per layer: b [c_out], w [3][3][c_in + 1][c_out] (time = channel 0)
  size_t poff[kLayers];
  float* WT2[2];             // layers 1, 2: [2 planes][C][9 C]  weights without the time channel, k = tap * C + c_in
  float* TM[kLayers];        // [9 classes][c_out]  time maps
  float* X2[2];              // inputs of layers 1, 2: [2 planes][3 Rp][C]
  float* X3;                 // input of layer 3: [3 Rp][C] fp32
  float* S[3];               // sigmoid(a_l) of the primal stream, l = 0, 1, 2: [Rp][C]
  float *F, *JE, *Y1;        // [B][ldD]
  const float* eps;          // [B][D]
  const float* ys;           // stage input y [B][D] (the flat driver's YS)
  float* tval;
  // ---- adjoint (mode == MODE_ADJ): what the reverse sweep needs from the forward passes, and its own buffers --------
  int fin;
  float* A1[3];              // tangent pre-activations a1_l of layers 0..2: [2 streams][Rp][C]
  float* W2F[2];             // layers 1, 2: [2 planes][C (c_in)][9 C], k = tap' * C + c_out, tap' = 8 - tap (backward data)
  float* AB3;                // cotangent of the last layer's output per stream: [3 Rp] on the padded grid
  float* AB2[2];             // cotangents of a_l / a1_l, l = 1, 2: [2 planes][3 Rp][C]
  float* AB0;                // cotangents of a_0 / a1_0: [3 Rp][C] fp32
  float* WQ[3];              // tangent streams' contribution to the cotangent of a_l: [2][Rp][C]
  float* XB1;                // [2][B][ldD] cotangent of the stream inputs (e_0 = eps, d_0 = f)
  float* EBD;                // [B][ldD] direct eps cotangent
  float* YBDOT;              // [B][ldD] vjp_y
  float* WPART[2];           // split-K partials of the layer 1, 2 weight gradients: [splits][9][C][C]
  int wsplits;
  float* PART;               // [4 layers][chunks][27][C] partial sums: 9 class sums of ab_x, 9 of ab_d, 9 direct taps
  int chunks;
  __host__ __device__ size_t xplane() const { return (size_t)3 * Rp * C; }
};

constexpr int kRedChunks = 592;   // 4 per SM; every block is 4 pixel groups x 64 channels

// border class of image pixel (y, x): 3 * {0 top, 1 inside, 2 bottom} + {0 left, 1 inside, 2 right}
__device__ __forceinline__ int border_class(int y, int x, int Hh, int Ww) {
  const int ry = (y == 0) ? 0 : (y == Hh - 1 ? 2 : 1), rx = (x == 0) ? 0 : (x == Ww - 1 ? 2 : 1);
  return 3 * ry + rx;
}
__device__ __forceinline__ bool tap_valid(int cls, int tap) {
  const int ry = cls / 3, rx = cls % 3, dy = tap / 3 - 1, dx = tap % 3 - 1;
  return !(ry == 0 && dy < 0) && !(ry == 2 && dy > 0) && !(rx == 0 && dx < 0) && !(rx == 2 && dx > 0);
}

// stage input (the generic part of k_css_stage) + the time of this evaluation
__global__ void k_ccv_stage(css::Flat a, float* tval, int stage, int init_mode) {
  Ctrl* c = a.ctrl;
  if (init_mode == css::ST_STAGE && c->done) return;
  const float dt = c->dt;
  float tau;
  if (init_mode == css::ST_INIT0 || init_mode == css::ST_PROBE) tau = a.ts[0];
  else if (init_mode == css::ST_INIT1) tau = a.ts[0] + c->h0;
  else tau = c->t + dt * css::c_tab.alpha[stage - 1];
  const float* Yc = a.Y + (size_t)c->cur * a.N;
  const float* Fc = a.F + (size_t)c->cur * a.N;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gid; i < a.n_feed; i += gsz) {
    float ys;
    if (init_mode == css::ST_INIT0 || init_mode == css::ST_PROBE) ys = Yc[i];
    else if (init_mode == css::ST_INIT1) ys = fmaf(c->h0, Fc[i], Yc[i]);
    else {
      float acc = css::c_tab.beta[stage - 1][0] * Fc[i];
      for (int j = 1; j < stage; ++j) acc = fmaf(css::c_tab.beta[stage - 1][j], a.KS[(size_t)(j - 1) * a.N + i], acc);
      ys = fmaf(dt, acc, Yc[i]);
    }
    a.YS[i] = ys;
  }
  if (gid == 0) *tval = a.tsign * tau;
}

// ---- once per solve: time maps and the split weights of the two tensor-core layers -------------------------------
__global__ void k_ccv_prepare(CNet n) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (int l = 0; l < kLayers; ++l) {
    const int cin = (l == 0) ? 1 : n.C, cout = (l == kLayers - 1) ? 1 : n.C;
    for (int i = gid; i < 9 * cout; i += gsz) {
      const int cls = i / cout, co = i - cls * cout;
      float acc = 0.f;
      for (int tap = 0; tap < 9; ++tap)
        if (tap_valid(cls, tap)) acc += n.W[l][((size_t)tap * (cin + 1)) * cout + co];
      n.TM[l][i] = acc;
    }
  }
  for (int l = 1; l <= 2; ++l) {
    const int C = n.C;
    const size_t plane = (size_t)C * 9 * C;
    for (int i = gid; i < C * 9 * C; i += gsz) {
      const int co = i / (9 * C), k = i - co * 9 * C, tap = k / C, ci = k - tap * C;
      float hi, lo;
      tc::split_tf32(n.W[l][((size_t)tap * (C + 1) + 1 + ci) * C + co], hi, lo);
      n.WT2[l - 1][i] = hi;
      n.WT2[l - 1][plane + i] = lo;
    }
    if (n.mode == css::MODE_ADJ) {   // x_bar[p][c_in] = sum_tap sum_c_out ab[p - shift(tap)][c_out] w[tap][c_in][c_out]
      for (int i = gid; i < C * 9 * C; i += gsz) {
        const int ci = i / (9 * C), k = i - ci * 9 * C, tp = k / C, co = k - tp * C;
        float hi, lo;
        tc::split_tf32(n.W[l][((size_t)(8 - tp) * (C + 1) + 1 + ci) * C + co], hi, lo);
        n.W2F[l - 1][i] = hi;
        n.W2F[l - 1][plane + i] = lo;
      }
    }
  }
}

// ---- layer 0 (1 -> C), direct.  TANGENT = false: primal stream from the stage input; true: the nq tangent streams
// (0: eps, 1: f with dt = 1).  One thread = one pixel x 4 output channels.
template <bool TANGENT>
__global__ void k_ccv_l0(CNet n, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, cq = C / 4;
  const size_t total = (size_t)(TANGENT ? n.nq : 1) * n.B * n.D * cq;
  const float t = *n.tval;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cq) * 4;
    size_t r = i / cq;
    const int pix = (int)(r % n.D);
    r /= n.D;
    const int b = (int)(r % n.B), q = (int)(r / n.B);
    const int y = pix / n.Ww, x = pix - y * n.Ww;
    const float* in = TANGENT ? (q == 0 ? n.eps : n.F) : n.ys;
    const int ldin = (TANGENT && q == 1) ? n.ldD : n.D;
    const int cls = border_class(y, x, n.Hh, n.Ww);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= n.Hh || xx < 0 || xx >= n.Ww) continue;
      const float v = in[(size_t)b * ldin + yy * n.Ww + xx];
      const float* w = n.W[0] + ((size_t)tap * 2 + 1) * C + c4;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(v, w[j], acc[j]);
    }
    const size_t m = (size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1);   // padded position
    float out[4];
    if (!TANGENT) {
      float s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = acc[j] + n.b[0][c4 + j] + t * n.TM[0][cls * C + c4 + j];
        css::softplus_sigmoid(a, out[j], s[j]);
      }
      if (n.nq > 0 || n.mode == css::MODE_ADJ) *reinterpret_cast<float4*>(n.S[0] + m * C + c4) = make_float4(s[0], s[1], s[2], s[3]);
      css::put_split_row(n.X2[0], n.xplane(), m * C, c4, C, out);
    } else {
      const float4 s = *reinterpret_cast<const float4*>(n.S[0] + m * C + c4);
      const float sv[4] = {s.x, s.y, s.z, s.w};
      float a1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a1[j] = acc[j] + (q == 1 ? n.TM[0][cls * C + c4 + j] : 0.f);
        out[j] = sv[j] * a1[j];
      }
      const size_t row = (size_t)(1 + q) * n.Rp + m;
      css::put_split_row(n.X2[0], n.xplane(), row * C, c4, C, out);
      if (n.mode == css::MODE_ADJ) {
        *reinterpret_cast<float4*>(n.A1[0] + ((size_t)q * n.Rp + m) * C + c4) = make_float4(a1[0], a1[1], a1[2], a1[3]);
      }
    }
  }
}

// ---- layers 1, 2 (C -> C) on the tensor cores: epilogues ---------------------------------------------------------
// decode a padded position; false for halo / padding rows (they keep their zeros)
__device__ __forceinline__ bool interior(const CNet& n, int m, int& cls) {
  if (m >= n.R) return false;
  const int pos = m % n.PP, py = pos / n.WP, px = pos - py * n.WP;
  if (py == 0 || py > n.Hh || px == 0 || px > n.Ww) return false;
  cls = border_class(py - 1, px - 1, n.Hh, n.Ww);
  return true;
}

struct EpiConvPrimal {
  CNet n;
  int l;   // 1 or 2
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[css::kW], int) const {
    const int C = n.C;
    int cls;
    if (n0 >= C || !interior(n, m, cls)) return;
    const float t = *n.tval;
    float x[css::kW], s[css::kW];
#pragma unroll
    for (int j = 0; j < css::kW; ++j) {
      const int c = min(n0 + j, C - 1);
      css::softplus_sigmoid(v[j] + n.b[l][c] + t * n.TM[l][cls * C + c], x[j], s[j]);
    }
    if (n.nq > 0 || n.mode == css::MODE_ADJ) css::put_row(n.S[l] + (size_t)m * C, n0, C, s);
    if (l == 1) {
      css::put_split_row(n.X2[1], n.xplane(), (size_t)m * C, n0, C, x);
    } else css::put_row(n.X3 + (size_t)m * C, n0, C, x);
  }
};

struct EpiConvTangent {   // rows: stream q = m / Rp (0 eps, 1 Taylor), position mm = m % Rp
  CNet n;
  int l;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[css::kW], int) const {
    const int C = n.C;
    const int q = m / n.Rp, mm = m - q * n.Rp;
    int cls;
    if (q >= n.nq || n0 >= C || !interior(n, mm, cls)) return;
    float s[css::kW], x1[css::kW], a1[css::kW];
    css::get_row(n.S[l] + (size_t)mm * C, n0, C, s);
#pragma unroll
    for (int j = 0; j < css::kW; ++j) {
      const int c = min(n0 + j, C - 1);
      a1[j] = v[j] + (q == 1 ? n.TM[l][cls * C + c] : 0.f);
      x1[j] = s[j] * a1[j];
    }
    const size_t row = (size_t)(1 + q) * n.Rp + mm;
    if (n.mode == css::MODE_ADJ) css::put_row(n.A1[l] + ((size_t)q * n.Rp + mm) * C, n0, C, a1);
    if (l == 1) {
      css::put_split_row(n.X2[1], n.xplane(), row * C, n0, C, x1);
    } else css::put_row(n.X3 + row * C, n0, C, x1);
  }
};

// ---- layer 3 (C -> 1), direct: one warp = one pixel of one stream (0 primal, 1 + q tangent) -------------------------
__global__ void k_ccv_l3(CNet n, int first_stream, int n_streams, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = (size_t)n_streams * n.B * n.D;
  const float t = *n.tval;
  for (size_t i = warp; i < total; i += nwarps) {
    const int pix = (int)(i % n.D);
    const size_t r = i / n.D;
    const int b = (int)(r % n.B), st = first_stream + (int)(r / n.B);
    const int y = pix / n.Ww, x = pix - y * n.Ww;
    const size_t m = (size_t)st * n.Rp + (size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1);
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* xr = n.X3 + (m + (tap / 3 - 1) * n.WP + (tap % 3 - 1)) * C;   // halo rows are zero
      const float* w = n.W[3] + ((size_t)tap * (C + 1) + 1);                     // c_out = 1
      for (int c = lane; c < C; c += 32) acc = fmaf(xr[c], w[c], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const int cls = border_class(y, x, n.Hh, n.Ww);
      if (st == 0) n.F[(size_t)b * n.ldD + pix] = acc + n.b[3][0] + t * n.TM[3][cls];
      else if (st == 1) n.JE[(size_t)b * n.ldD + pix] = acc;
      else n.Y1[(size_t)b * n.ldD + pix] = acc + n.TM[3][cls];
    }
  }
}

// The same layer, one warp = one image ROW of one stream (image width <= 32): the three input rows are read once each
// (3 (W + 2) positions for W outputs instead of 9 W), a lane holds channels (lane, lane + 32) and W running outputs.
// KIND 0: layer 3 forward (reads X3, writes F / JE / Y1); KIND 1: backward data of layer 0 (reads the cotangents AB0 with the
// taps mirrored, writes vjp_y / the cotangents of the tangent streams' inputs) - both are C -> 1 convolutions.
template <int KIND>
__global__ void k_ccv_l3_rows(CNet n, int first_stream, int n_streams, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = (size_t)n_streams * n.B * n.Hh;
  const float t = *n.tval;
  const bool two = C > 32;
  float w0[9], w1[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    if (KIND == 0) {
      w0[tap] = n.W[3][(size_t)tap * (C + 1) + 1 + lane];
      w1[tap] = two ? n.W[3][(size_t)tap * (C + 1) + 1 + lane + 32] : 0.f;
    } else {   // out[p] = sum_tap g[p - shift(tap)] w0[tap][c_out] = sum_tap' g[p + shift(tap')] w0[8 - tap'][c_out]
      w0[tap] = n.W[0][((size_t)(8 - tap) * 2 + 1) * C + lane];
      w1[tap] = two ? n.W[0][((size_t)(8 - tap) * 2 + 1) * C + lane + 32] : 0.f;
    }
  }
  const float* src = (KIND == 0) ? n.X3 : n.AB0;
  for (size_t i = warp; i < total; i += nwarps) {
    const int y = (int)(i % n.Hh);
    const size_t r = i / n.Hh;
    const int b = (int)(r % n.B), st = first_stream + (int)(r / n.B);
    float acc[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) acc[x] = 0.f;
#pragma unroll
    for (int ry = 0; ry < 3; ++ry) {
      const float* row = src + ((size_t)st * n.Rp + (size_t)b * n.PP + (size_t)(y + ry) * n.WP) * C;   // padded row y + ry
#pragma unroll
      for (int xx = 0; xx < 34; ++xx) {
        if (xx < n.Ww + 2) {
          const float v0 = row[(size_t)xx * C + lane];
          const float v1 = two ? row[(size_t)xx * C + lane + 32] : 0.f;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int x = xx - dx;   // the output column this position feeds through tap (ry, dx)
            if (x >= 0 && x < 32) acc[x] = fmaf(v0, w0[ry * 3 + dx], fmaf(v1, w1[ry * 3 + dx], acc[x]));
          }
        }
      }
    }
#pragma unroll
    for (int x = 0; x < 32; ++x) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[x] += __shfl_xor_sync(0xffffffffu, acc[x], o);
    }
    // lane x finishes output column x
    float mine = 0.f;
#pragma unroll
    for (int x = 0; x < 32; ++x)
      if (lane == x) mine = acc[x];
    if (lane < n.Ww) {
      const int pix = y * n.Ww + lane;
      if (KIND == 0) {
        const int cls = border_class(y, lane, n.Hh, n.Ww);
        if (st == 0) n.F[(size_t)b * n.ldD + pix] = mine + n.b[3][0] + t * n.TM[3][cls];
        else if (st == 1) n.JE[(size_t)b * n.ldD + pix] = mine;
        else n.Y1[(size_t)b * n.ldD + pix] = mine + n.TM[3][cls];
      } else {
        if (st == 0) n.YBDOT[(size_t)b * n.ldD + pix] = mine;
        else n.XB1[((size_t)(st - 1) * n.B + b) * n.ldD + pix] = mine;
      }
    }
  }
}

// =====================================================================================================================
// Adjoint: the vjp of aug_dynamics (ffjord_mnist.py:311-324) wrt (y, t, eps, params), lib/ode.py:1498-1504.  Same order
// as css_rhs.cuh: tangent streams first (their cotangent reaches f through d_0 = f), then the primal stream; weight
// gradients contract all streams over the positions.  Cotangents live on the padded grid as well (halo rows zero), so
// the backward-data convolution is the same 9-tap loop with the taps mirrored (W2F).
// =====================================================================================================================

// per-sample scalars (div, r, fro, kin) and - adjoint - the cotangents of the last layer's tangent outputs
__global__ void k_ccv_seed(css::Flat a, CNet n, float* DIV, float* RR, float* FRO, float* KIN) {
  if (a.ctrl->done) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n.B) return;
  const int b = warp, D = n.D;
  const float* je = n.JE + (size_t)b * n.ldD;
  const float* y1 = n.Y1 + (size_t)b * n.ldD;
  const float* f = n.F + (size_t)b * n.ldD;
  const float* ep = n.eps + (size_t)b * D;
  float div = 0.f, rr = 0.f, fro = 0.f, kin = 0.f;
  for (int d = lane; d < D; d += 32) {
    if (n.nq >= 1) { div = fmaf(je[d], ep[d], div); fro = fmaf(je[d], je[d], fro); }
    if (n.nq == 2) rr = fmaf(y1[d], y1[d], rr);
    kin = fmaf(f[d], f[d], kin);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    div += __shfl_xor_sync(0xffffffffu, div, o);
    rr += __shfl_xor_sync(0xffffffffu, rr, o);
    fro += __shfl_xor_sync(0xffffffffu, fro, o);
    kin += __shfl_xor_sync(0xffffffffu, kin, o);
  }
  if (lane == 0) { DIV[b] = div; RR[b] = rr / D; FRO[b] = fro / D; KIN[b] = kin / D; }
  if (n.mode != css::MODE_ADJ) return;
  const float* cur = a.Y + (size_t)a.ctrl->cur * a.N;
  const float pb = cur[a.oAUXB + b];
  const float rb = n.fin ? 0.f : cur[a.oAUXB + (size_t)a.B + b];
  const float frob = n.fin ? cur[a.oAUXB + (size_t)a.B + b] : 0.f;
  for (int d = lane; d < D; d += 32) {
    const int y = d / n.Ww, x = d - y * n.Ww;
    const size_t m = (size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1);
    // dp = -div: cotangent of J eps is -p_bar eps (+ fro_bar 2 J eps / D); direct eps cotangent -p_bar J eps
    n.AB3[(size_t)n.Rp + m] = -pb * ep[d] + frob * 2.f * je[d] / D;
    n.EBD[(size_t)b * n.ldD + d] = -pb * je[d];
    if (n.nq == 2) n.AB3[(size_t)2 * n.Rp + m] = rb * 2.f * y1[d] / D;
  }
}

// cotangent of f: y_bar (+ cotangent of d_0 = f) (+ kin term)
__global__ void k_ccv_seed2(css::Flat a, CNet n) {
  if (a.ctrl->done) return;
  const size_t BD = (size_t)n.B * n.D;
  const float* cur = a.Y + (size_t)a.ctrl->cur * a.N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < BD; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    float w = a.YS[a.oYB + i];
    if (n.nq == 2) w += n.XB1[((size_t)n.B + b) * n.ldD + d];
    if (n.fin) w += cur[a.oAUXB + 2 * (size_t)a.B + b] * 2.f * n.F[(size_t)b * n.ldD + d] / n.D;
    const int y = d / n.Ww, x = d - y * n.Ww;
    n.AB3[(size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1)] = w;
  }
}

// Fold the activation of layer lm (0..2): xb = cotangent of that layer's output (x_{lm+1} of stream s at position mm,
// channels c0..c0+3).  Tangent stream q: a1_bar = xb S, and xb a1 S (1 - S) flows to the primal pre-activation (WQ);
// primal stream: a_bar = xb S + sum_q WQ.  The result is the operand of the layer's own backward pass.
__device__ __forceinline__ void fold(const CNet& n, int lm, int s, int mm, int c0, const float (&xb)[css::kW]) {
  const int C = n.C;
  float sg[css::kW], ab[css::kW];
  css::get_row(n.S[lm] + (size_t)mm * C, c0, C, sg);
  if (s >= 1) {
    const size_t o = ((size_t)(s - 1) * n.Rp + mm) * C;
    float a1[css::kW], wq[css::kW];
    css::get_row(n.A1[lm] + o, c0, C, a1);
#pragma unroll
    for (int j = 0; j < css::kW; ++j) {
      ab[j] = xb[j] * sg[j];
      wq[j] = xb[j] * a1[j] * sg[j] * (1.f - sg[j]);
    }
    css::put_row(n.WQ[lm] + o, c0, C, wq);
  } else {
#pragma unroll
    for (int j = 0; j < css::kW; ++j) ab[j] = xb[j] * sg[j];
    for (int q = 0; q < n.nq; ++q) {
      float wq[css::kW];
      css::get_row(n.WQ[lm] + ((size_t)q * n.Rp + mm) * C, c0, C, wq);
#pragma unroll
      for (int j = 0; j < css::kW; ++j) ab[j] += wq[j];
    }
  }
  const size_t row = (size_t)s * n.Rp + mm;
  if (lm == 0) css::put_row(n.AB0 + row * C, c0, C, ab);
  else {
    css::put_split_row(n.AB2[lm - 1], n.xplane(), row * C, c0, C, ab);
  }
}

// backward data of layer 3 (1 -> C "broadcast") + fold of layer 2; streams [first, first + count)
__global__ void k_ccv_l3_bwd(CNet n, int first_stream, int n_streams, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, cq = C / 4;
  const size_t total = (size_t)n_streams * n.B * n.D * cq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cq) * 4;
    size_t r = i / cq;
    const int pix = (int)(r % n.D);
    r /= n.D;
    const int b = (int)(r % n.B), s = first_stream + (int)(r / n.B);
    const int y = pix / n.Ww, x = pix - y * n.Ww;
    const int mm = b * n.PP + (y + 1) * n.WP + (x + 1);
    const float* ab3 = n.AB3 + (size_t)s * n.Rp + mm;
    float xb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float g = ab3[-((tap / 3 - 1) * n.WP + (tap % 3 - 1))];     // halo entries are zero
      const float* w = n.W[3] + ((size_t)tap * (C + 1) + 1 + c4);
#pragma unroll
      for (int j = 0; j < 4; ++j) xb[j] = fmaf(g, w[j], xb[j]);
    }
    fold(n, 2, s, mm, c4, xb);
  }
}

// epilogue of the backward-data convolution of layer l (2 or 1): folds layer l - 1
struct EpiConvRev {
  CNet n;
  int l, primal;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[css::kW], int) const {
    const int s = primal ? 0 : 1 + m / n.Rp;
    const int mm = primal ? m : m - (s - 1) * n.Rp;
    int cls;
    if (s > n.nq || n0 >= n.C || !interior(n, mm, cls)) return;
    fold(n, l - 1, s, mm, n0, v);
  }
};

// backward data of layer 0 (C -> 1): one warp = one pixel of one stream
__global__ void k_ccv_l0_bwd(CNet n, int first_stream, int n_streams, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = (size_t)n_streams * n.B * n.D;
  for (size_t i = warp; i < total; i += nwarps) {
    const int pix = (int)(i % n.D);
    const size_t r = i / n.D;
    const int b = (int)(r % n.B), s = first_stream + (int)(r / n.B);
    const int y = pix / n.Ww, x = pix - y * n.Ww;
    const size_t m = (size_t)s * n.Rp + (size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1);
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* g = n.AB0 + (m - ((tap / 3 - 1) * n.WP + (tap % 3 - 1))) * C;
      const float* w = n.W[0] + ((size_t)tap * 2 + 1) * C;
      for (int c = lane; c < C; c += 32) acc = fmaf(g[c], w[c], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (s == 0) n.YBDOT[(size_t)b * n.ldD + pix] = acc;
      else n.XB1[((size_t)(s - 1) * n.B + b) * n.ldD + pix] = acc;
    }
  }
}

// Weight gradients of the two C -> C layers: dW[tap][c_in][c_out] = sum_p x[p + shift(tap)][c_in] ab[p][c_out] over the
// positions of all streams.  On the tensor cores the contraction index is the position, i.e. the SLOW index of both operands
// as they lie in memory ([position][channel]): MN-major operand descriptors (gemm_tc.cuh: k_gemm_tf32x3<.., MN = true>,
// SWIZZLE_128B_BASE32B), the tap shift on the TMA row coordinate, two taps per 128-row tile (EpiConvWgradMN).  With K-major
// (transposed) planes a tap would be a shift along the CONTIGUOUS dimension of a TMA box - not a multiple of 16 bytes,
// which TMA rejects (measured: illegal instruction).
struct EpiConvWgradMN {
  float* WP;
  int C;
  // tile rows: 64 input channels of tap 2 x, then of tap 2 x + 1; tile columns: x * 64 + c_out
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[css::kW], int split) const {
    const int pair = n0 >> 6, co = n0 & 63, tap = 2 * pair + (m >> 6), ci = m & 63;
    if (tap > 8 || ci >= C || co >= C) return;
    float* row = WP + (((size_t)split * 9 + tap) * C + ci) * C + co;
#pragma unroll
    for (int j = 0; j < css::kW; ++j)
      if (co + j < C) row[j] = v[j];
  }
};

// The FFMA formulation of the same contraction (ENO_CCV_WGRAD_FFMA=1; 43 TFLOP/s = 60 % of the FFMA peak at batch 200):
// grid (9 taps, splits): one block = one tap x one chunk of positions, 64 x 64 outputs, 4 x 4 per thread.
__global__ void __launch_bounds__(256) k_ccv_wgrad(CNet n, int l, const int* __restrict__ done) {
  if (done && *done) return;
  const int C = n.C, tap = blockIdx.x, split = blockIdx.y;
  const int K = (1 + n.nq) * n.Rp;
  const int per = ((K + gridDim.y - 1) / gridDim.y + 15) / 16 * 16;
  const int k0 = split * per, k1 = min(K, k0 + per);
  const int sh = (tap / 3 - 1) * n.WP + (tap % 3 - 1);
  __shared__ __align__(16) float xs[16][64], gs[16][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;     // this thread's piece of a 16 x 64 tile
  const float* xh = n.X2[l];
  const float* gh = n.AB2[l];
  const size_t plane = n.xplane();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int kk = k0; kk < k1; kk += 16) {
    const int p = kk + lr, px = p + sh;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
    if (p < k1 && lc < C) {
      const float4 h = *reinterpret_cast<const float4*>(gh + (size_t)p * C + lc);
      const float4 lo = *reinterpret_cast<const float4*>(gh + plane + (size_t)p * C + lc);
      gv = make_float4(h.x + lo.x, h.y + lo.y, h.z + lo.z, h.w + lo.w);
      if (px >= 0 && px < 3 * n.Rp) {
        const float4 a = *reinterpret_cast<const float4*>(xh + (size_t)px * C + lc);
        const float4 b = *reinterpret_cast<const float4*>(xh + plane + (size_t)px * C + lc);
        xv = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&xs[lr][lc]) = xv;
    *reinterpret_cast<float4*>(&gs[lr][lc]) = gv;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 a4 = *reinterpret_cast<const float4*>(&xs[r][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&gs[r][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ty * 4 + i;
    if (ci >= C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = tx * 4 + j;
      if (co < C) n.WPART[l][(((size_t)split * 9 + tap) * C + ci) * C + co] = acc[i][j];
    }
  }
}

// Everything that is a plain sum over the positions, per layer, in kRedChunks fixed chunks (deterministic):
//   [0, 9)   class sums of the primal cotangent ab_x          -> bias, time weights (x t), t_bar
//   [9, 18)  class sums of the Taylor stream's cotangent ab_d -> time weights
//   [18, 27) layer 0: dW0[tap][c_out] = sum x_s[p + shift] ab_s[p][c_out];  layer 3: dW3[tap][c_in] = sum x3_s[p + shift][c_in] ab_s[p]
// thread = (pixel group g of 4, channel c): c_out for layers 0..2, c_in for layer 3; the four groups take every fourth pixel of
// the block's chunk and are summed in order through shared memory
__global__ void __launch_bounds__(256) k_ccv_reduce(CNet n, const int* __restrict__ done) {
  if (done && *done) return;
  const int l = blockIdx.y, ch = blockIdx.x, c = threadIdx.x & 63, grp = threadIdx.x >> 6, C = n.C;
  __shared__ float red[3][27][64];
  float acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.f;
  const int total = n.B * n.D;
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int i0 = ch * per, i1 = min(total, i0 + per);
  for (int i = i0 + grp; i < i1 && c < C; i += 4) {
    const int b = i / n.D, pix = i - b * n.D;
    const int y = pix / n.Ww, x = pix - y * n.Ww;
    const int cls = border_class(y, x, n.Hh, n.Ww);
    const size_t m = (size_t)b * n.PP + (size_t)(y + 1) * n.WP + (x + 1);
    float ab[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      // the class sums need the primal and the Taylor stream; only layer 0's direct weight gradient reads all three
      if (s > n.nq || (s == 1 && l != 0) || l == 3) { ab[s] = 0.f; continue; }
      const size_t row = (size_t)s * n.Rp + m;
      if (l == 0) ab[s] = n.AB0[row * C + c];
      else ab[s] = n.AB2[l - 1][row * C + c] + n.AB2[l - 1][n.xplane() + row * C + c];
    }
    if (l == 3 && c == 0) { ab[0] = n.AB3[m]; ab[2] = (n.nq == 2) ? n.AB3[(size_t)2 * n.Rp + m] : 0.f; }
    if (l < 3 || c == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (k == cls) { acc[k] += ab[0]; acc[9 + k] += ab[2]; }
      }
    }
    if (l == 0) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= n.Hh || xx < 0 || xx >= n.Ww) continue;
        const int o = yy * n.Ww + xx;
        float v = n.ys[(size_t)b * n.D + o] * ab[0];
        if (n.nq >= 1) v = fmaf(n.eps[(size_t)b * n.D + o], ab[1], v);
        if (n.nq == 2) v = fmaf(n.F[(size_t)b * n.ldD + o], ab[2], v);
        acc[18 + tap] += v;
      }
    } else if (l == 3) {
      // dW3[tap][c_in] = sum_p x3[p + shift][c_in] ab3[p] = sum_q x3[q][c_in] ab3[q - shift]: every x3 row is read ONCE (it is
      // zero on the halo, and so is ab3), the nine cotangents around it are scalars
      for (int s = 0; s <= n.nq; ++s) {
        const float xv = n.X3[((size_t)s * n.Rp + m) * C + c];
        const float* g3 = n.AB3 + (size_t)s * n.Rp + m;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) acc[18 + tap] = fmaf(xv, g3[-((tap / 3 - 1) * n.WP + (tap % 3 - 1))], acc[18 + tap]);
      }
    }
  }
  if (grp > 0) {
#pragma unroll
    for (int k = 0; k < 27; ++k) red[grp - 1][k][c] = acc[k];
  }
  __syncthreads();
  if (grp == 0 && c < C) {
    float* out = n.PART + (((size_t)l * gridDim.x + ch) * 27) * C + c;
#pragma unroll
    for (int k = 0; k < 27; ++k) out[(size_t)k * C] = ((acc[k] + red[0][k][c]) + red[1][k][c]) + red[2][k][c];
  }
}

// second stage of k_ccv_reduce: PART[l][chunk][k][c] summed over the chunks in order -> PART[l][0][k][c]
__global__ void k_ccv_reduce2(CNet n, const int* __restrict__ done) {
  if (done && *done) return;
  const int l = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 27 * n.C; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int ch = 0; ch < n.chunks; ++ch) acc += n.PART[(((size_t)l * n.chunks + ch) * 27) * n.C + i];
    n.PART[((size_t)l * n.chunks * 27) * n.C + i] = acc;
  }
}

// the slot of the adjoint state: [ -f | -aux dots | vjp_y | 0 | t_bar | eps_bar | params_bar ]
__global__ void k_ccv_assemble_adj(css::Flat a, CNet n, const float* DIV, const float* RR, const float* FRO, const float* KIN,
                                   int stage, int init_mode, float* k_override) {
  if (init_mode == css::ST_STAGE && a.ctrl->done) return;
  float* k = k_override ? k_override : css::k_slot(a, stage, init_mode);
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t BD = (size_t)n.B * n.D, Bz = (size_t)n.B;
  const int C = n.C;
  const float t = *n.tval;
  for (size_t i = gid; i < BD; i += gsz) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    k[i] = -n.F[(size_t)b * n.ldD + d];
    k[a.oYB + i] = n.YBDOT[(size_t)b * n.ldD + d];
    k[a.oEB + i] = (n.nq >= 1) ? n.XB1[(size_t)b * n.ldD + d] + n.EBD[(size_t)b * n.ldD + d] : 0.f;
  }
  for (size_t i = gid; i < (size_t)a.n_aux * Bz; i += gsz) {
    const int q = (int)(i / Bz), b = (int)(i - (size_t)q * Bz);
    float v;
    if (q == 0) v = -DIV[b];
    else if (n.fin) v = (q == 1) ? FRO[b] : KIN[b];
    else v = (n.nq == 2) ? RR[b] : 0.f;
    k[a.oAUX + i] = -v;
    k[a.oAUXB + i] = 0.f;
  }
  // params_bar, per layer [ b | w [3][3][c_in + 1][c_out] ]
  for (int l = 0; l < kLayers; ++l) {
    const int cin = (l == 0) ? 1 : C, cout = (l == kLayers - 1) ? 1 : C;
    const float* P = n.PART + ((size_t)l * n.chunks * 27) * C;      // reduced sums [27][C]
    float* pb = k + a.oPR + n.poff[l];
    float* pw = pb + cout;
    for (size_t i = gid; i < (size_t)cout; i += gsz) {
      float acc = 0.f;
      for (int cls = 0; cls < 9; ++cls) acc += P[(size_t)cls * C + i];
      pb[i] = acc;
    }
    const size_t nw = (size_t)9 * (cin + 1) * cout;
    for (size_t i = gid; i < nw; i += gsz) {
      const int co = (int)(i % cout);
      const size_t r = i / cout;
      const int ci1 = (int)(r % (cin + 1)), tap = (int)(r / (cin + 1));
      float v;
      if (ci1 == 0) {          // time channel: the taps that fall inside the image
        v = 0.f;
        for (int cls = 0; cls < 9; ++cls)
          if (tap_valid(cls, tap)) v += fmaf(t, P[(size_t)cls * C + co], P[(size_t)(9 + cls) * C + co]);
      } else if (l == 0) v = P[(size_t)(18 + tap) * C + co];
      else if (l == 3) v = P[(size_t)(18 + tap) * C + (ci1 - 1)];
      else {
        v = 0.f;
        for (int sp = 0; sp < n.wsplits; ++sp) v += n.WPART[l - 1][(((size_t)sp * 9 + tap) * C + (ci1 - 1)) * C + co];
      }
      pw[i] = v;
    }
  }
  if (blockIdx.x == 0) {   // t_bar = sum over layers, classes, channels of (class sum of ab_x) * T
    __shared__ double sm[32];
    double acc = 0.0;
    for (int l = 0; l < kLayers; ++l) {
      const int cout = (l == kLayers - 1) ? 1 : C;
      const float* P = n.PART + ((size_t)l * n.chunks * 27) * C;
      for (int i = threadIdx.x; i < 9 * cout; i += blockDim.x) {
        const int cls = i / cout, co = i - cls * cout;
        acc += (double)P[(size_t)cls * C + co] * (double)n.TM[l][i];
      }
    }
    const double tot = css::block_reduce(acc, sm);
    if (threadIdx.x == 0) k[a.oT0] = (float)tot;
  }
}

}  // namespace ccv
}  // namespace eno
