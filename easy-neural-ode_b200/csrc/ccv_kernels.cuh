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
  __host__ __device__ size_t xplane() const { return (size_t)3 * Rp * C; }
};

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
      if (n.nq > 0) *reinterpret_cast<float4*>(n.S[0] + m * C + c4) = make_float4(s[0], s[1], s[2], s[3]);
      css::put_split_row(n.X2[0], n.xplane(), m * C, c4, C, out);
    } else {
      const float4 s = *reinterpret_cast<const float4*>(n.S[0] + m * C + c4);
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = sv[j] * (acc[j] + (q == 1 ? n.TM[0][cls * C + c4 + j] : 0.f));
      css::put_split_row(n.X2[0], n.xplane(), ((size_t)(1 + q) * n.Rp + m) * C, c4, C, out);
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
    if (n.nq > 0) css::put_row(n.S[l] + (size_t)m * C, n0, C, s);
    if (l == 1) css::put_split_row(n.X2[1], n.xplane(), (size_t)m * C, n0, C, x);
    else css::put_row(n.X3 + (size_t)m * C, n0, C, x);
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
    float s[css::kW], x1[css::kW];
    css::get_row(n.S[l] + (size_t)mm * C, n0, C, s);
#pragma unroll
    for (int j = 0; j < css::kW; ++j) {
      const int c = min(n0 + j, C - 1);
      x1[j] = s[j] * (v[j] + (q == 1 ? n.TM[l][cls * C + c] : 0.f));
    }
    const size_t row = (size_t)(1 + q) * n.Rp + mm;
    if (l == 1) css::put_split_row(n.X2[1], n.xplane(), row * C, n0, C, x1);
    else css::put_row(n.X3 + row * C, n0, C, x1);
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

}  // namespace ccv
}  // namespace eno
