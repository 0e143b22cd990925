// Second translation unit of libeno.so: the GEMM-pipeline engine for wide dynamics networks - the ConcatSquash
// stacks of ffjord_tabular.py (BASELINE config C2).  Stage contractions run on tcgen05 as 3xTF32 (gemm_tc.cuh) with
// fused epilogues (css_rhs.cuh); the adaptive dopri5 loop is a flat-state driver with a device-resident controller
// (css_kernels.cuh); one loop iteration (6 right-hand sides + controller) is captured once into a CUDA graph and
// replayed, so the host only polls the controller once per chunk of iterations.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "css_api.h"
#include "css_kernels.cuh"
#include "ccv_kernels.cuh"
#include "eno_ctx.cuh"
#include "tableaux.cuh"

using namespace eno;

namespace {

thread_local std::string g_css_err;

int css_fail(eno_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_css_err = msg;
  return code;
}

#define CSS_CUDA(ctx, expr)                                                                                       \
  do {                                                                                                            \
    cudaError_t e_ = (expr);                                                                                      \
    if (e_ != cudaSuccess) return css_fail(ctx, ENO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

inline size_t up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int ld4(int w) { return (w + 3) / 4 * 4; }

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += up(n * sizeof(T));
    return p;
  }
};

struct Plan {
  css::Net net;
  css::Flat flat;
  bool conv = false;         // ENO_ARCH_CONCATCONV: the right-hand side is the convolutional pipeline of ccv_kernels.cuh
  ccv::CNet cnet;
  size_t P = 0;
  float* params = nullptr;   // copies inside the workspace, so that a captured iteration graph only names the workspace
  float* eps = nullptr;
  float* seg_ts = nullptr;
  size_t bytes = 0;
  size_t zero_from = 0, zero_bytes = 0;   // transposed operands: their padding columns must read as zero
};

int aux_count(int aug) {
  switch (aug) {
    case ENO_AUG_NONE: return 0;
    case ENO_AUG_FFJORD: return 1;
    case ENO_AUG_REG: return 2;
    case ENO_AUG_FIN: return 3;
    case ENO_AUG_ALL: return 4;
  }
  return -1;
}

// ConcatConv2D dynamics (forward closures only): the flat driver's buffers + the padded-grid activations
void plan_build_conv(const eno_dynamics_desc* d, int B, int T, int mode, void* ws, Plan* p) {
  p->conv = true;
  ccv::CNet& c = p->cnet;
  css::Net& n = p->net;
  css::Flat& f = p->flat;
  memset(&c, 0, sizeof(c));
  memset(&n, 0, sizeof(n));
  memset(&f, 0, sizeof(f));
  const int C = d->H, D = d->D;
  c.B = B; c.Hh = d->img_h; c.Ww = d->img_w; c.C = C;
  c.HP = c.Hh + 2; c.WP = c.Ww + 2; c.PP = c.HP * c.WP;
  c.R = B * c.PP; c.Rp = (c.R + 127) / 128 * 128;
  c.D = D; c.ldD = ld4(D);
  switch (d->aug) {
    case ENO_AUG_NONE: c.nq = 0; break;
    case ENO_AUG_FFJORD: case ENO_AUG_FIN: c.nq = 1; break;
    default: c.nq = d->reg_order ? 2 : 1;
  }
  c.mode = mode ? css::MODE_ADJ : (d->aug == ENO_AUG_NONE ? css::MODE_PLAIN : css::MODE_AUG);
  c.fin = (d->aug == ENO_AUG_FIN);
  // what k_css_assemble / k_css_tbar read
  n.B = B; n.Bp = (B + 127) / 128 * 128; n.D = D; n.H = C; n.L = 1; n.nq = c.nq; n.mode = c.mode;
  n.fin = (d->aug == ENO_AUG_FIN); n.ldD = c.ldD;
  size_t po = 0;
  for (int l = 0; l < ccv::kLayers; ++l) {
    const size_t cin = l == 0 ? 1 : C, cout = l == ccv::kLayers - 1 ? 1 : C;
    c.poff[l] = po;
    po += cout + 9 * (cin + 1) * cout;
  }
  p->P = po;
  const size_t BD = (size_t)B * D;
  const int na = aux_count(d->aug);
  f.B = B; f.D = D; f.n_aux = na;
  f.oAUX = BD;
  if (mode) {
    f.oYB = BD + (size_t)na * B;
    f.oAUXB = f.oYB + BD;
    f.oT0 = f.oAUXB + (size_t)na * B;
    f.oEB = f.oT0 + 1;
    f.oPR = f.oEB + BD;
    f.N = f.oPR + p->P;
    f.n_feed = f.oAUXB;
    f.tsign = -1.f;
  } else {
    f.N = BD + (size_t)na * B;
    f.n_feed = BD;
    f.tsign = 1.f;
  }
  Carver cv{static_cast<char*>(ws)};
  f.ctrl = cv.take<Ctrl>(1);
  f.Y = cv.take<float>(3 * f.N);
  f.F = cv.take<float>(3 * f.N);
  f.KS = cv.take<float>(5 * f.N);
  f.MID = cv.take<float>(2 * f.N);
  f.YS = cv.take<float>(f.n_feed);
  f.n_part = 2048;
  f.part = cv.take<double>(4 * (size_t)f.n_part);
  f.out = cv.take<float>((size_t)std::max(T, 2) * f.N);
  f.world = 1; f.rank = 0;
  f.gbuf = cv.take<double>(32);
  f.n_x = mode ? 1 + p->P : 0;
  f.xbuf = cv.take<float>(std::max<size_t>(4 * f.n_x, 4));
  p->params = cv.take<float>(p->P);
  p->eps = cv.take<float>(BD);
  p->seg_ts = cv.take<float>(4);
  n.eps = p->eps;
  n.tval = cv.take<float>(4);
  c.eps = p->eps; c.ys = f.YS; c.tval = n.tval;
  for (int l = 0; l < ccv::kLayers; ++l) {
    const size_t cout = l == ccv::kLayers - 1 ? 1 : C;
    c.b[l] = p->params ? p->params + c.poff[l] : nullptr;
    c.W[l] = p->params ? p->params + c.poff[l] + cout : nullptr;
    c.TM[l] = cv.take<float>(9 * cout);
  }
  for (int l = 0; l < 2; ++l) c.WT2[l] = cv.take<float>((size_t)2 * C * 9 * C);
  for (int l = 0; l < 3; ++l) c.S[l] = cv.take<float>((size_t)c.Rp * C);
  c.F = n.F = cv.take<float>((size_t)B * c.ldD);
  c.JE = n.JE = cv.take<float>((size_t)B * c.ldD);
  c.Y1 = n.Y1 = cv.take<float>((size_t)B * c.ldD);
  n.DIV = cv.take<float>(B); n.RR = cv.take<float>(B); n.FRO = cv.take<float>(B); n.KIN = cv.take<float>(B);
  if (mode) {
    for (int l = 0; l < 3; ++l) {
      c.A1[l] = cv.take<float>((size_t)2 * c.Rp * C);
      c.WQ[l] = cv.take<float>((size_t)2 * c.Rp * C);
    }
    for (int l = 0; l < 2; ++l) c.W2F[l] = cv.take<float>((size_t)2 * C * 9 * C);
    c.XB1 = cv.take<float>((size_t)2 * B * c.ldD);
    c.EBD = cv.take<float>((size_t)B * c.ldD);
    c.YBDOT = cv.take<float>((size_t)B * c.ldD);
    c.wsplits = 29;   // 5 tap pairs x 29 position chunks = 145 CTAs
    for (int l = 0; l < 2; ++l) c.WPART[l] = cv.take<float>((size_t)c.wsplits * 9 * C * C);
    c.chunks = ccv::kRedChunks;
    c.PART = cv.take<float>((size_t)ccv::kLayers * c.chunks * 27 * C);
  }
  p->zero_from = cv.off;   // halo / padding rows of the activations and of the cotangents must read as zero
  for (int l = 0; l < 2; ++l) c.X2[l] = cv.take<float>(2 * c.xplane());
  c.X3 = cv.take<float>((size_t)3 * c.Rp * C);
  if (mode) {
    for (int l = 0; l < 2; ++l) c.AB2[l] = cv.take<float>(2 * c.xplane());
    c.AB0 = cv.take<float>((size_t)3 * c.Rp * C);
    c.AB3 = cv.take<float>((size_t)3 * c.Rp);
  }
  p->zero_bytes = cv.off - p->zero_from;
  p->bytes = cv.off + 1024;
}

// mode: 0 forward solve, 1 adjoint solve
void plan_build(const eno_dynamics_desc* d, int B, int T, int mode, void* ws, Plan* p) {
  if (d->arch == ENO_ARCH_CONCATCONV) { plan_build_conv(d, B, T, mode, ws, p); return; }
  css::Net& n = p->net;
  css::Flat& f = p->flat;
  memset(&n, 0, sizeof(n));
  memset(&f, 0, sizeof(f));
  const int D = d->D, H = d->H, L = d->n_hidden + 1;
  n.B = B; n.Bp = (B + 127) / 128 * 128; n.D = D; n.H = H; n.L = L;
  n.mode = mode ? css::MODE_ADJ : (d->aug == ENO_AUG_NONE ? css::MODE_PLAIN : css::MODE_AUG);
  n.fin = (d->aug == ENO_AUG_FIN);
  switch (d->aug) {
    case ENO_AUG_NONE: n.nq = 0; break;
    case ENO_AUG_FFJORD: case ENO_AUG_FIN: n.nq = 1; break;
    default: n.nq = d->reg_order ? 2 : 1;
  }
  n.ldD = ld4(D);
  n.hmax = std::max(H, D);
  size_t po = 0;
  for (int l = 0; l < L; ++l) {
    n.din[l] = l == 0 ? D : H;
    n.dout[l] = l == L - 1 ? D : H;
    n.ldi[l] = ld4(n.din[l]);
    n.ldo[l] = ld4(n.dout[l]);
    n.poff[l] = po;
    po += (size_t)n.din[l] * n.dout[l] + 4 * (size_t)n.dout[l];
  }
  p->P = po;
  const size_t BD = (size_t)B * D;
  const int na = aux_count(d->aug);
  f.B = B; f.D = D; f.n_aux = na;
  f.oAUX = BD;
  if (mode) {
    f.oYB = BD + (size_t)na * B;
    f.oAUXB = f.oYB + BD;
    f.oT0 = f.oAUXB + (size_t)na * B;
    f.oEB = f.oT0 + 1;
    f.oPR = f.oEB + BD;
    f.N = f.oPR + p->P;
    f.n_feed = f.oAUXB;    // y, aux, y_bar (the aux_bar blocks are constants read from the state ring)
    f.tsign = -1.f;
  } else {
    f.N = BD + (size_t)na * B;
    f.n_feed = BD;
    f.tsign = 1.f;
  }
  const bool adj = mode != 0;
  Carver cv{static_cast<char*>(ws)};
  f.ctrl = cv.take<Ctrl>(1);
  f.Y = cv.take<float>(3 * f.N);
  f.F = cv.take<float>(3 * f.N);
  f.KS = cv.take<float>(5 * f.N);
  f.MID = cv.take<float>(2 * f.N);
  f.YS = cv.take<float>(f.n_feed);
  f.n_part = 2048;
  f.part = cv.take<double>(4 * (size_t)f.n_part);
  f.out = cv.take<float>((size_t)std::max(T, 2) * f.N);
  f.world = 1; f.rank = 0;
  f.gbuf = cv.take<double>(32);
  f.n_x = mode ? 1 + p->P : 0;
  f.xbuf = cv.take<float>(std::max<size_t>(4 * f.n_x, 4));
  p->params = cv.take<float>(p->P);
  p->eps = cv.take<float>(BD);
  p->seg_ts = cv.take<float>(4);
  n.eps = p->eps;
  n.tval = cv.take<float>(4);
  // adjoint: the weight-gradient GEMMs contract X2 / AB2 over their ROWS (MN-major operands): rows B .. Bp of every stream
  // must read as zero, so everything from here on is cleared once per solve
  const size_t zero_all_from = cv.off;
  for (int l = 0; l < L; ++l) {
    const float* base = p->params ? p->params + n.poff[l] : nullptr;
    const size_t dd = (size_t)n.din[l] * n.dout[l];
    n.b[l] = base;
    n.W[l] = base ? base + n.dout[l] : nullptr;
    n.wb[l] = base ? base + n.dout[l] + dd : nullptr;
    n.bg[l] = base ? base + 2 * (size_t)n.dout[l] + dd : nullptr;
    n.wg[l] = base ? base + 3 * (size_t)n.dout[l] + dd : nullptr;
    n.WT2[l] = cv.take<float>(2 * (size_t)n.dout[l] * n.ldi[l]);
    if (adj) n.W2[l] = cv.take<float>(2 * (size_t)n.din[l] * n.ldo[l]);
    n.G[l] = cv.take<float>(n.dout[l]);
    n.G1[l] = cv.take<float>(n.dout[l]);
    n.CC[l] = cv.take<float>(n.dout[l]);
    n.X2[l] = cv.take<float>(2 * n.xplane(l));
    n.LIN[l] = cv.take<float>((size_t)B * n.ldo[l]);
    if (l < L - 1) n.S[l] = cv.take<float>((size_t)B * n.ldo[l]);
    if (adj) {
      n.A1[l] = cv.take<float>(2 * (size_t)B * n.ldo[l]);
      n.V[l] = cv.take<float>(2 * (size_t)B * n.ldo[l]);
      n.WQ[l] = cv.take<float>(2 * (size_t)B * n.ldo[l]);
      n.WW[l] = cv.take<float>((size_t)B * n.ldo[l]);
      n.AB2[l] = cv.take<float>(2 * n.abplane(l));
      n.WP[l] = cv.take<float>(3 * (size_t)n.din[l] * n.dout[l]);
    }
  }
  n.cr_chunks = css::kCrChunks;
  if (adj) {
    n.CR = cv.take<float>((size_t)L * n.cr_chunks * 4 * n.hmax);
    n.TC = cv.take<float>((size_t)L * n.hmax);
    n.XB1 = cv.take<float>(2 * (size_t)B * n.ldD);
    n.EBD = cv.take<float>((size_t)B * n.ldD);
    n.YBDOT = cv.take<float>((size_t)B * n.ldD);
  }
  n.F = cv.take<float>((size_t)B * n.ldD);
  n.JE = cv.take<float>((size_t)B * n.ldD);
  n.Y1 = cv.take<float>((size_t)B * n.ldD);
  n.DIV = cv.take<float>(B); n.RR = cv.take<float>(B); n.FRO = cv.take<float>(B); n.KIN = cv.take<float>(B);
  p->zero_from = adj ? zero_all_from : cv.off;
  p->zero_bytes = cv.off - p->zero_from;
  p->bytes = cv.off + 1024;
}

int pick_bn(int M, int N) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("ENO_CSS_BN"); forced = e ? atoi(e) : 0; }
  if (N < 128) return 64;
  if (forced == 64 || forced == 128) return forced;
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  return tiles < 100 ? 64 : 128;
}

// bench.py's roofline leg: CUDA-event pairs around every GEMM launch of a profiled solve (direct launches, one
// iteration per poll so that no launch runs past the end of the solve), with the ALGORITHMIC flops 2 M N K of each.
struct GemmProf {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<double> flop;
  size_t n = 0;
  double ms = 0.0, flops = 0.0;
  long long count = 0;
  void begin(cudaStream_t st, double f) {
    if (ev.size() < 2 * (n + 1)) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev.push_back(a); ev.push_back(b); }
    if (flop.size() < n + 1) flop.push_back(f); else flop[n] = f;
    cudaEventRecord(ev[2 * n], st);
  }
  void end(cudaStream_t st) { cudaEventRecord(ev[2 * n + 1], st); ++n; }
  void resolve() {   // after the stream has been synchronised
    for (size_t i = 0; i < n; ++i) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]) == cudaSuccess) { ms += t; flops += flop[i]; ++count; }
    }
    n = 0;
  }
};
thread_local GemmProf* g_prof = nullptr;

template <class Epi>
cudaError_t gemm(const tc::GemmOperand& A, const tc::GemmOperand& Bm, int M, int N, int splits, const int* done, const Epi& epi,
                 cudaStream_t st, std::string* err) {
  const bool prof = g_prof && g_prof->on;
  if (prof) g_prof->begin(st, 2.0 * M * N * (double)A.k);
  cudaError_t e = pick_bn(M, N) == 128 ? tc::launch_gemm<128>(A, Bm, M, N, splits, done, epi, st, err)
                                       : tc::launch_gemm<64>(A, Bm, M, N, splits, done, epi, st, err);
  if (prof) g_prof->end(st);
  return e;
}

#define CSS_G(expr)                                                                   \
  do {                                                                                \
    cudaError_t e_ = (expr);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      if (err->empty()) *err = std::string(#expr) + ": " + cudaGetErrorString(e_);    \
      return ENO_E_CUDA;                                                              \
    }                                                                                 \
  } while (0)

// ConcatConv2D dynamics: direct first / last layers, the two C -> C layers as 9-tap loops of the tensor-core kernel
int ccv_rhs_launch(const Plan& p, int stage, int init_mode, float* k_override, int sms, cudaStream_t st, std::string* err) {
  const ccv::CNet& c = p.cnet;
  const css::Flat& f = p.flat;
  const int C = c.C, B = c.B;
  const int* done = (init_mode == css::ST_STAGE) ? &f.ctrl->done : nullptr;
  const int egrid = std::min<int>((int)((f.n_feed + 255) / 256), 4 * sms);
  ccv::k_ccv_stage<<<egrid, 256, 0, st>>>(f, c.tval, stage, init_mode);
  CUtensorMap ta[2], tb[2];
  for (int l = 0; l < 2; ++l) {
    if (!tc::make_operand_map(&ta[l], c.X2[l], 3 * c.Rp, C, C, c.xplane(), tc::kBM, err) ||
        !tc::make_operand_map(&tb[l], c.WT2[l], C, 9 * C, 9 * C, (size_t)C * 9 * C, 64, err))
      return ENO_E_CUDA;
  }
  const size_t pix = (size_t)B * c.D;
  auto l0grid = [&](size_t streams) { return (int)std::min<size_t>((streams * pix * (C / 4) + 255) / 256, (size_t)16 * sms); };
  auto l3grid = [&](size_t streams) { return (int)std::min<size_t>((streams * pix * 32 + 255) / 256, (size_t)16 * sms); };
  // bench.py's roofline leg: CUDA-event pairs around the tensor-core launches, ALGORITHMIC flops (interior positions,
  // 2 * C * 9 C per position and stream - not the halo rows, not the three tf32 passes)
  const bool prof = g_prof && g_prof->on;
  const double conv_flop = 2.0 * (double)pix * C * 9.0 * C;
  auto pb = [&](double streams) { if (prof) g_prof->begin(st, streams * conv_flop); };
  auto pe = [&]() { if (prof) g_prof->end(st); };
  ccv::k_ccv_l0<false><<<l0grid(1), 256, 0, st>>>(c, done);
  pb(1); CSS_G(tc::launch_conv3x3<64>(ta[0], tb[0], 0, c.R, C, C, c.WP, done, ccv::EpiConvPrimal{c, 1}, st)); pe();
  pb(1); CSS_G(tc::launch_conv3x3<64>(ta[1], tb[1], 0, c.R, C, C, c.WP, done, ccv::EpiConvPrimal{c, 2}, st)); pe();
  const bool l3rows = c.Ww <= 32;   // one warp per image row: every input row is read once
  auto l3rgrid = [&](size_t streams) { return (int)std::min<size_t>((streams * B * c.Hh * 32 + 255) / 256, (size_t)16 * sms); };
  if (l3rows) ccv::k_ccv_l3_rows<0><<<l3rgrid(1), 256, 0, st>>>(c, 0, 1, done);
  else ccv::k_ccv_l3<<<l3grid(1), 256, 0, st>>>(c, 0, 1, done);
  if (c.nq > 0) {   // the tangent streams, stacked (the Taylor stream's direction is f: after the primal pass)
    ccv::k_ccv_l0<true><<<l0grid(c.nq), 256, 0, st>>>(c, done);
    pb(c.nq); CSS_G(tc::launch_conv3x3<64>(ta[0], tb[0], c.Rp, c.nq * c.Rp, C, C, c.WP, done, ccv::EpiConvTangent{c, 1}, st)); pe();
    pb(c.nq); CSS_G(tc::launch_conv3x3<64>(ta[1], tb[1], c.Rp, c.nq * c.Rp, C, C, c.WP, done, ccv::EpiConvTangent{c, 2}, st)); pe();
    if (l3rows) ccv::k_ccv_l3_rows<0><<<l3rgrid(c.nq), 256, 0, st>>>(c, 1, c.nq, done);
    else ccv::k_ccv_l3<<<l3grid(c.nq), 256, 0, st>>>(c, 1, c.nq, done);
  }
  const css::Net& n = p.net;
  if (c.mode != css::MODE_PLAIN) ccv::k_ccv_seed<<<(B * 32 + 255) / 256, 256, 0, st>>>(f, c, n.DIV, n.RR, n.FRO, n.KIN);
  if (c.mode != css::MODE_ADJ) {
    css::k_css_assemble<<<std::min((int)((pix + 255) / 256), 8 * sms), 256, 0, st>>>(f, n, stage, init_mode, k_override);
    CSS_G(cudaGetLastError());
    return 0;
  }
  // ---- reverse sweep: backward-data convolutions (taps mirrored), tangent streams first, then the primal stream
  CUtensorMap tg[2], tw[2];
  for (int l = 0; l < 2; ++l) {
    if (!tc::make_operand_map(&tg[l], c.AB2[l], 3 * c.Rp, C, C, c.xplane(), tc::kBM, err) ||
        !tc::make_operand_map(&tw[l], c.W2F[l], C, 9 * C, 9 * C, (size_t)C * 9 * C, 64, err))
      return ENO_E_CUDA;
  }
  auto reverse = [&](int first, int count, int primal) -> int {
    ccv::k_ccv_l3_bwd<<<l0grid(count), 256, 0, st>>>(c, first, count, done);
    pb(count);
    CSS_G(tc::launch_conv3x3<64>(tg[1], tw[1], first * c.Rp, primal ? c.R : count * c.Rp, C, C, c.WP, done,
                                 ccv::EpiConvRev{c, 2, primal}, st));
    pe();
    pb(count);
    CSS_G(tc::launch_conv3x3<64>(tg[0], tw[0], first * c.Rp, primal ? c.R : count * c.Rp, C, C, c.WP, done,
                                 ccv::EpiConvRev{c, 1, primal}, st));
    pe();
    if (l3rows) ccv::k_ccv_l3_rows<1><<<l3rgrid(count), 256, 0, st>>>(c, first, count, done);
    else ccv::k_ccv_l0_bwd<<<l3grid(count), 256, 0, st>>>(c, first, count, done);
    return 0;
  };
  int rc;
  if (c.nq > 0 && (rc = reverse(1, c.nq, 0))) return rc;
  ccv::k_ccv_seed2<<<std::min((int)((pix + 255) / 256), 4 * sms), 256, 0, st>>>(f, c);
  if ((rc = reverse(0, 1, 1))) return rc;
  // ---- weight gradients: the two C -> C layers on the tensor cores (MN-major operands, 5 tap pairs x position chunks),
  // everything else by chunked sums
  static const bool wgrad_ffma = getenv("ENO_CCV_WGRAD_FFMA") != nullptr;
  for (int l = 0; l < 2; ++l) {
    if (wgrad_ffma) { ccv::k_ccv_wgrad<<<dim3(9, c.wsplits), 256, 0, st>>>(c, l, done); continue; }
    CUtensorMap txa, txb;   // the operands as they lie in memory ([position][channel] planes), MN-major
    const int K = (1 + c.nq) * c.Rp;
    if (!tc::make_operand_map(&txa, c.X2[l], K, C, C, c.xplane(), 32, err, true) ||
        !tc::make_operand_map(&txb, c.AB2[l], K, C, C, c.xplane(), 32, err, true))
      return ENO_E_CUDA;
    pb(1 + c.nq);
    CSS_G(tc::launch_gemm_mn<64>(txa, txb, 128, 64, K, c.wsplits, c.WP, done, ccv::EpiConvWgradMN{c.WPART[l], C}, st));
    pe();
  }
  ccv::k_ccv_reduce<<<dim3(c.chunks, ccv::kLayers), 256, 0, st>>>(c, done);
  ccv::k_ccv_reduce2<<<dim3(8, ccv::kLayers), 256, 0, st>>>(c, done);
  const size_t work = std::max(pix, (size_t)9 * (C + 1) * C);
  ccv::k_ccv_assemble_adj<<<std::min((int)((work + 255) / 256), 8 * sms), 256, 0, st>>>(f, c, n.DIV, n.RR, n.FRO, n.KIN, stage, init_mode,
                                                                                        k_override);
  CSS_G(cudaGetLastError());
  return 0;
}

// One evaluation of the integrated function into the stage slot (or k_override).
int css_rhs_launch(const Plan& p, int stage, int init_mode, float* k_override, int sms, cudaStream_t st, std::string* err) {
  if (p.conv) return ccv_rhs_launch(p, stage, init_mode, k_override, sms, st, err);
  const css::Net& n = p.net;
  const css::Flat& f = p.flat;
  const int L = n.L, B = n.B, Bp = n.Bp;
  const int* done = (init_mode == css::ST_STAGE) ? &f.ctrl->done : nullptr;
  const int egrid = std::min<int>((int)((std::max(f.n_feed, (size_t)n.hmax) + 255) / 256), 4 * sms);
  CSS_G(tc::launch_pdl(css::k_css_stage, dim3(egrid), dim3(256), 0, st, f, n, stage, init_mode));
  // forward: primal stream, then the tangent streams stacked
  for (int l = 0; l < L; ++l) {
    tc::GemmOperand A{n.X2[l], 3 * Bp, n.din[l], n.ldi[l], n.xplane(l), 0};
    tc::GemmOperand W{n.WT2[l], n.dout[l], n.din[l], n.ldi[l], (size_t)n.dout[l] * n.ldi[l], 0};
    CSS_G(gemm(A, W, B, n.dout[l], 1, done, css::EpiPrimal{n, l}, st, err));
  }
  if (n.nq > 0) {
    for (int l = 0; l < L; ++l) {
      tc::GemmOperand A{n.X2[l], 3 * Bp, n.din[l], n.ldi[l], n.xplane(l), Bp};
      tc::GemmOperand W{n.WT2[l], n.dout[l], n.din[l], n.ldi[l], (size_t)n.dout[l] * n.ldi[l], 0};
      CSS_G(gemm(A, W, n.nq * Bp, n.dout[l], 1, done, css::EpiTangent{n, l}, st, err));
    }
  }
  if (n.mode != css::MODE_PLAIN) CSS_G(tc::launch_pdl(css::k_css_seed, dim3((B * 32 + 255) / 256), dim3(256), 0, st, f, n));
  if (n.mode == css::MODE_ADJ) {
    // reverse: tangent streams
    if (n.nq > 0) {
      for (int l = L - 1; l >= 0; --l) {
        tc::GemmOperand A{n.AB2[l], 3 * Bp, n.dout[l], n.ldo[l], n.abplane(l), Bp};
        tc::GemmOperand W{n.W2[l], n.din[l], n.dout[l], n.ldo[l], (size_t)n.din[l] * n.ldo[l], 0};
        if (l > 0) CSS_G(gemm(A, W, n.nq * Bp, n.din[l], 1, done, css::EpiTanRev{n, l}, st, err));
        else CSS_G(gemm(A, W, n.nq * Bp, n.din[l], 1, done, css::EpiTanRev0{n}, st, err));
      }
    }
    CSS_G(tc::launch_pdl(css::k_css_seed2, dim3(std::min((int)(((size_t)B * n.D + 255) / 256), 4 * sms)), dim3(256), 0, st, f, n));
    for (int l = L - 1; l >= 0; --l) {
      tc::GemmOperand A{n.AB2[l], 3 * Bp, n.dout[l], n.ldo[l], n.abplane(l), 0};
      tc::GemmOperand W{n.W2[l], n.din[l], n.dout[l], n.ldo[l], (size_t)n.din[l] * n.ldo[l], 0};
      if (l > 0) CSS_G(gemm(A, W, B, n.din[l], 1, done, css::EpiPrimRev{n, l}, st, err));
      else CSS_G(gemm(A, W, B, n.din[l], 1, done, css::EpiPrimRev0{n}, st, err));
    }
    // weight gradients: contraction over the batch, one K-split per row stream
    // (MN-major operands: the activations and cotangents as they lie in memory, [row][column] planes - no transposed copies)
    for (int l = 0; l < L; ++l) {
      const int K = (1 + n.nq) * Bp;
      CUtensorMap txa, txb;
      if (!tc::make_operand_map(&txa, n.X2[l], K, n.din[l], n.ldi[l], n.xplane(l), 32, err, true) ||
          !tc::make_operand_map(&txb, n.AB2[l], K, n.dout[l], n.ldo[l], n.abplane(l), 32, err, true))
        return ENO_E_CUDA;
      const bool prof = g_prof && g_prof->on;
      if (prof) g_prof->begin(st, 2.0 * n.din[l] * n.dout[l] * (double)((1 + n.nq) * B));
      const css::EpiWgrad epi{n.WP[l], n.din[l], n.dout[l]};
      if (n.dout[l] >= 128) CSS_G(tc::launch_gemm_mn<128>(txa, txb, n.din[l], n.dout[l], K, 1 + n.nq, 0, done, epi, st));
      else CSS_G(tc::launch_gemm_mn<64>(txa, txb, n.din[l], n.dout[l], K, 1 + n.nq, 0, done, epi, st));
      if (prof) g_prof->end(st);
    }
    CSS_G(tc::launch_pdl(css::k_css_colred, dim3((n.hmax + 31) / 32, n.cr_chunks, L), dim3(32, 8), 0, st, f, n));
    CSS_G(tc::launch_pdl(css::k_css_colfin, dim3((n.hmax + 63) / 64, L), dim3(64), 0, st, f, n, stage, init_mode,
                         k_override ? k_override + f.oPR : (float*)nullptr));
  }
  const size_t work = std::max((size_t)B * n.D, n.mode == css::MODE_ADJ ? (size_t)n.H * n.H : (size_t)1);
  CSS_G(tc::launch_pdl(css::k_css_assemble, dim3(std::min((int)((work + 255) / 256), 8 * sms)), dim3(256), 0, st, f, n, stage, init_mode,
                       k_override));
  CSS_G(cudaGetLastError());
  return 0;
}

int css_prepare(eno_ctx* ctx, Plan& p, const float* params, const float* eps, cudaStream_t st) {
  const css::Net& n = p.net;
  Tableau* tab = &ctx->tabs[p.flat.grid_step > 0.f ? ENO_NUM_TABLEAUX : ENO_TABLEAU_DOPRI5];
  CSS_CUDA(ctx, cudaMemcpyToSymbolAsync(css::c_tab, tab, sizeof(Tableau), 0, cudaMemcpyHostToDevice, st));
  CSS_CUDA(ctx, cudaMemcpyAsync(p.params, params, p.P * 4, cudaMemcpyDeviceToDevice, st));
  if (eps) CSS_CUDA(ctx, cudaMemcpyAsync(p.eps, eps, (size_t)n.B * n.D * 4, cudaMemcpyDeviceToDevice, st));
  else CSS_CUDA(ctx, cudaMemsetAsync(p.eps, 0, (size_t)n.B * n.D * 4, st));
  if (p.zero_bytes) CSS_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(p.flat.ctrl) + p.zero_from, 0, p.zero_bytes, st));
  if (p.conv) {
    ccv::k_ccv_prepare<<<256, 256, 0, st>>>(p.cnet);
    CSS_CUDA(ctx, cudaGetLastError());
    return 0;
  }
  for (int l = 0; l < n.L; ++l) {
    // forward operand: W^T [dout][din];  reverse operand: W [din][dout]
    tc::k_split_planes<<<512, 256, 0, st>>>(n.W[l], n.din[l], n.dout[l], n.dout[l], n.WT2[l], n.ldi[l], (size_t)n.dout[l] * n.ldi[l], 1);
    if (n.W2[l])
      tc::k_split_planes<<<512, 256, 0, st>>>(n.W[l], n.din[l], n.dout[l], n.dout[l], n.W2[l], n.ldo[l], (size_t)n.din[l] * n.ldo[l], 0);
  }
  if (n.nq > 0) css::k_css_eps<<<std::min((int)(((size_t)n.B * n.D + 255) / 256), 1024), 256, 0, st>>>(n);
  CSS_CUDA(ctx, cudaGetLastError());
  return 0;
}

// cached graph of one loop iteration, keyed by everything its kernel arguments depend on
struct GraphKey {
  void* ws; int B, D, H, L, aug, reg, mode, T, world, rank, grid, spare;
  bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
};
struct CssState {
  std::map<GraphKey, cudaGraphExec_t> graphs;
  int use_graph = 1;
  int hint[2] = {8, 8};
  GemmProf prof;
};

CssState* state_of(eno_ctx* ctx) {
  if (!ctx->css) {
    CssState* s = new CssState();
    if (const char* e = getenv("ENO_CSS_GRAPH")) s->use_graph = atoi(e);
    ctx->css = s;
  }
  return static_cast<CssState*>(ctx->css);
}

// Collective solves (eno_comm_init has run): the batch rows are sharded over the ranks, the solve is ONE global solve of the
// concatenated batch.  Returns the element count of the global state (the mean in the error norm, lib/ode.py:518-538).
double set_collective(eno_ctx* ctx, Plan* p, int B) {
  css::Flat& f = p->flat;
  if (!ctx->dist.active()) return (double)f.N;
  f.world = ctx->dist.world;
  f.rank = ctx->dist.rank;
  const double per_row = (double)(f.N - f.n_x) / (double)B;
  return per_row * (double)ctx->dist.rows(B) + (double)f.n_x;
}

// NCCL calls of the collective solves (dist.cuh); every rank enqueues the same sequence
#define CSS_NCCL(expr, what)                                                                      \
  do {                                                                                            \
    int n_ = (expr);                                                                              \
    if (n_ != 0) { if (err) *err = std::string(what) + ": " + dc->api.GetErrorString(n_); return ENO_E_CUDA; } \
  } while (0)

// sums the replicated elements ([t0_bar | params_bar]) of one flat vector over the ranks, in place
int allreduce_replicated(const Plan& p, const DistCtx* dc, float* vec, cudaStream_t st, std::string* err) {
  if (!dc || p.flat.n_x == 0) return 0;
  CSS_NCCL(dc->api.GroupStart(), "ncclGroupStart");
  CSS_NCCL(dc->allreduce(vec + p.flat.oT0, 1, st), "ncclAllReduce");
  CSS_NCCL(dc->allreduce(vec + p.flat.oPR, p.flat.n_x - 1, st), "ncclAllReduce");
  CSS_NCCL(dc->api.GroupEnd(), "ncclGroupEnd");
  return 0;
}

int enqueue_iteration(const Plan& p, const DistCtx* dc, int sms, cudaStream_t st, std::string* err) {
  for (int s = 1; s <= 6; ++s) {
    if (p.flat.grid_step > 0.f && (s == 4 || s == 5)) continue;   // RK4 in the 7-stage frame (tableaux.cuh:fill_rk4_grid)
    int rc = css_rhs_launch(p, s, css::ST_STAGE, nullptr, sms, st, err);
    if (rc) return rc;
  }
  const int fgrid = std::min<int>((int)((p.flat.N + 255) / 256), p.flat.n_part);
  if (dc && p.flat.n_x) {
    css::k_flat_xcomb<<<std::min<int>((int)((p.flat.n_x + 255) / 256), 4 * sms), 256, 0, st>>>(p.flat);
    CSS_NCCL(dc->allreduce(p.flat.xbuf, 4 * p.flat.n_x, st), "ncclAllReduce");
  }
  CSS_G(tc::launch_pdl(css::k_flat_finish, dim3(fgrid), dim3(256), 0, st, p.flat));
  if (dc) {
    CSS_NCCL(dc->gather(p.flat.gbuf, 2, st), "ncclAllGather");
    css::k_flat_control<<<1, 1, 0, st>>>(p.flat);
  }
  CSS_G(tc::launch_pdl(css::k_flat_out, dim3(std::min<int>((int)((p.flat.N + 255) / 256), 4 * sms)), dim3(256), 0, st, p.flat));
  return 0;
}

// Runs the adaptive loop to completion on an initialised state (Y[0] / out[0] set, controller reset).
int run_loop(eno_ctx* ctx, const Plan& p, const GraphKey& key, int which, eno_stats* stats, int* launches, cudaStream_t st) {
  CssState* S = state_of(ctx);
  std::string err;
  const int sms = ctx->sm_count;
  const DistCtx* dc = p.flat.world > 1 ? &ctx->dist : nullptr;
  int rc;
  // initial step: f0, norms, f(y0 + h0 f0), norms.  Collective solves: the replicated elements of both derivatives are
  // summed over the ranks (the controller ring is at slot 0 here: F[cur] = F), the norms gathered in rank order.
  if ((rc = css_rhs_launch(p, 0, css::ST_INIT0, nullptr, sms, st, &err))) return css_fail(ctx, rc, err);
  const int igrid = std::min<int>((int)((p.flat.N + 255) / 256), p.flat.n_part);
  const bool grid = p.flat.grid_step > 0.f;
  if (grid) {   // fixed grid: no initial step-size selection; stage buffers 4 and 5 are never written, clear all of them
    if ((rc = allreduce_replicated(p, dc, p.flat.F, st, &err))) return css_fail(ctx, rc, err);
    CSS_CUDA(ctx, cudaMemsetAsync(p.flat.KS, 0, 5 * p.flat.N * sizeof(float), st));
    css::k_flat_grid_begin<<<1, 1, 0, st>>>(p.flat);
  }
  for (int phase = 0; phase < 2 && !grid; ++phase) {
    if (phase == 1 && (rc = css_rhs_launch(p, 0, css::ST_INIT1, nullptr, sms, st, &err))) return css_fail(ctx, rc, err);
    if ((rc = allreduce_replicated(p, dc, phase == 0 ? p.flat.F : p.flat.KS, st, &err))) return css_fail(ctx, rc, err);
    css::k_flat_init<<<igrid, 256, 0, st>>>(p.flat, phase);
    if (dc) {
      if (dc->gather(p.flat.gbuf, 2, st) != 0) return css_fail(ctx, ENO_E_CUDA, "ncclAllGather failed (initial step)");
      css::k_flat_init_control<<<1, 1, 0, st>>>(p.flat, phase);
    }
  }
  cudaGraphExec_t exec = nullptr;
  const bool profiling = ctx->prof.enabled;
  S->prof.on = false;
  if (profiling) { S->prof.on = true; g_prof = &S->prof; }   // the initial-step evaluations above are not timed
  if (S->use_graph && !profiling) {
    auto it = S->graphs.find(key);
    if (it != S->graphs.end()) exec = it->second;
    else {
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = enqueue_iteration(p, dc, sms, st, &err);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc == 0 && e == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) S->graphs[key] = exec;
        else { exec = nullptr; cudaGetLastError(); }
        if (g) cudaGraphDestroy(g);
      } else cudaGetLastError();
    }
  }
  int chunk = profiling ? 1 : (grid ? 8 : S->hint[which]);
  int n_launched_iters = 0;
  for (;;) {
    for (int k = 0; k < chunk; ++k) {
      if (exec) CSS_CUDA(ctx, cudaGraphLaunch(exec, st));
      else if ((rc = enqueue_iteration(p, dc, sms, st, &err))) return css_fail(ctx, rc, err);
      ++n_launched_iters;
    }
    CSS_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, p.flat.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    CSS_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->mailbox->done) break;
    chunk = profiling ? 1 : 2;
  }
  if (profiling) { S->prof.resolve(); S->prof.on = false; }
  const Ctrl& c = *ctx->mailbox;
  if (!grid) S->hint[which] = std::min(64, c.n_iters + 1);
  if (stats) {
    stats->nfe = (float)c.nfe; stats->n_iters = c.n_iters; stats->n_accepted = c.n_accepted; stats->status = c.status;
    stats->last_dt = c.dt; stats->last_t = c.t; stats->n_rhs = grid ? 1 + 4 * c.n_iters : 2 + 6 * c.n_iters;
    stats->n_launches = n_launched_iters;   // iteration graphs (or direct iteration sequences) enqueued
  }
  if (launches) *launches += n_launched_iters;
  return c.status;
}

}  // namespace

namespace eno {
namespace css {

bool conv_desc_ok(const eno_dynamics_desc* d) {
  return d && d->arch == ENO_ARCH_CONCATCONV && (d->H == 32 || d->H == 64) && d->n_hidden == 3 && d->img_h >= 2 && d->img_w >= 2 &&
         d->D == d->img_h * d->img_w && d->D % 4 == 0 && aux_count(d->aug) >= 0 && (d->reg_order == 0 || d->reg_order == 2);
}

bool desc_ok(const eno_dynamics_desc* d) {
  if (d && d->arch == ENO_ARCH_CONCATCONV) return conv_desc_ok(d);
  return d && d->arch == ENO_ARCH_CONCATSQUASH && d->D > 0 && d->H > 0 && d->n_hidden >= 1 && d->n_hidden + 1 <= kMaxLayers &&
         aux_count(d->aug) >= 0 && (d->reg_order == 0 || d->reg_order == 2);
}

int64_t param_count(const eno_dynamics_desc* d) {
  int64_t P = 0;
  if (d->arch == ENO_ARCH_CONCATCONV) {   // per layer b [c_out] + w [3][3][c_in + 1][c_out]
    const int64_t C = d->H;
    return (C + 18 * C) + 2 * (C + 9 * (C + 1) * C) + (1 + 9 * (C + 1));
  }
  const int L = d->n_hidden + 1;
  for (int l = 0; l < L; ++l) {
    const int64_t din = l == 0 ? d->D : d->H, dout = l == L - 1 ? d->D : d->H;
    P += din * dout + 4 * dout;
  }
  return P;
}

int n_aux(const eno_dynamics_desc* d) { return aux_count(d->aug); }

int64_t adjoint_state_size(const eno_dynamics_desc* d, int B) {
  const int64_t BD = (int64_t)B * d->D, na = aux_count(d->aug);
  return 2 * (BD + na * B) + 1 + BD + param_count(d);
}

size_t workspace_bytes(const eno_dynamics_desc* d, int B, int T, int mode) {
  if (!desc_ok(d) || B <= 0) return 0;
  Plan p;
  plan_build(d, B, T, mode, nullptr, &p);
  return p.bytes;
}

int odeint_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int B, const float* y0, const float* ts, int T, const float* params,
               const float* eps, float rtol, float atol, int mxstep, void* ws, size_t ws_bytes, float* ys_out, float* aux_out,
               eno_stats* stats, eno_trace_row* trace, int trace_cap, cudaStream_t st, float grid_step) {
  if (!desc_ok(desc)) return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: unsupported ConcatSquash descriptor");
  if (grid_step > 0.f && T != 2) return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_grid_fwd: ts must be [t0, t1] (lib/ode.py:884)");
  if (B <= 0 || T < 1 || !y0 || !ts || !params || !ws || !ys_out) return css_fail(ctx, ENO_E_ARG, "eno_odeint_fwd: null pointer or empty batch");
  const int na = aux_count(desc->aug);
  if (na > 0 && !eps) return css_fail(ctx, ENO_E_ARG, "eno_odeint_fwd: the FFJORD closures read eps");
  if (na > 0 && !aux_out) return css_fail(ctx, ENO_E_ARG, "eno_odeint_fwd: aux_out is required for augmented dynamics");
  Plan p;
  plan_build(desc, B, T, 0, ws, &p);
  if (ws_bytes < p.bytes) return css_fail(ctx, ENO_E_WORKSPACE, "eno_odeint_fwd: workspace too small");
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  p.flat.ts = ts;
  p.flat.trace = trace;
  p.flat.grid_step = grid_step;
  int rc;
  if ((rc = css_prepare(ctx, p, params, eps, st))) return rc;
  const size_t N = p.flat.N, BD = (size_t)B * desc->D;
  const double n_state = set_collective(ctx, &p, B);
  css::k_flat_reset<<<1, 1, 0, st>>>(p.flat.ctrl, rtol, atol, T, mxstep <= 0 ? INT_MAX : mxstep, trace ? trace_cap : 0, (float)n_state);
  // the auxiliary states start from aux_out[0] (zeros from aug_init; delta_logp of the logit pre-flow in
  // ffjord_mnist.py:409-417): their magnitude enters the error norm through tol = atol + rtol max(|y0|, |y1|)
  css::k_css_fwd_begin<<<std::min<int>((int)((N + 255) / 256), 1024), 256, 0, st>>>(p.flat, y0, na > 0 ? aux_out : nullptr);
  int launches = 0;
  int status = ENO_OK;
  if (T > 1) {
    GraphKey key{ws, B, desc->D, desc->H, desc->n_hidden, desc->aug, desc->reg_order, 0, T, p.flat.world, p.flat.rank, grid_step > 0.f ? 1 : 0, 0};
    status = run_loop(ctx, p, key, 0, stats, &launches, st);
    if (status < 0) return status;
  } else if (stats) memset(stats, 0, sizeof(*stats));
  // un-flatten the trajectory: out [T][N] -> ys [T][B][D], aux [T][n_aux][B]
  CSS_CUDA(ctx, cudaMemcpy2DAsync(ys_out, BD * 4, p.flat.out, N * 4, BD * 4, T, cudaMemcpyDeviceToDevice, st));
  if (na > 0)
    CSS_CUDA(ctx, cudaMemcpy2DAsync(aux_out, (size_t)na * B * 4, p.flat.out + BD, N * 4, (size_t)na * B * 4, T, cudaMemcpyDeviceToDevice, st));
  CSS_CUDA(ctx, cudaStreamSynchronize(st));
  return status;
}

int odeint_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int B, const float* ys, const float* ts, int T, const float* params,
               const float* eps, const float* g_y, const float* g_aux, float rtol, float atol, int mxstep, void* ws,
               size_t ws_bytes, float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out, float* ts_bar_out,
               float* params_bar_out, eno_stats* stats, eno_trace_row* trace, int trace_cap, cudaStream_t st, float grid_step) {
  if (!desc_ok(desc)) return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_bwd: unsupported ConcatSquash descriptor");
  if (grid_step > 0.f && T != 2) return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_grid_bwd: ts must be [t0, t1] (lib/ode.py:884)");
  const int na = aux_count(desc->aug);
  if (desc->aug != ENO_AUG_REG && desc->aug != ENO_AUG_FIN)
    return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_bwd: the FFJORD adjoint integrates aug_dynamics ('our' or 'fin')");
  if (B <= 0 || T < 1 || !ys || !ts || !params || !eps || !g_y || !g_aux || !ws || !y0_bar_out || !aux0_bar_out || !ts_bar_out ||
      !params_bar_out)
    return css_fail(ctx, ENO_E_ARG, "eno_odeint_bwd: null pointer or empty batch");
  Plan p;
  plan_build(desc, B, 2, 1, ws, &p);
  if (ws_bytes < p.bytes) return css_fail(ctx, ENO_E_WORKSPACE, "eno_odeint_bwd: workspace too small");
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t N = p.flat.N, BD = (size_t)B * desc->D, AB = (size_t)na * B;
  if (T == 1) {
    CSS_CUDA(ctx, cudaMemcpyAsync(y0_bar_out, g_y, BD * 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(ctx, cudaMemcpyAsync(aux0_bar_out, g_aux, AB * 4, cudaMemcpyDeviceToDevice, st));
    if (eps_bar_out) CSS_CUDA(ctx, cudaMemsetAsync(eps_bar_out, 0, BD * 4, st));
    CSS_CUDA(ctx, cudaMemsetAsync(ts_bar_out, 0, 4, st));
    CSS_CUDA(ctx, cudaMemsetAsync(params_bar_out, 0, p.P * 4, st));
    CSS_CUDA(ctx, cudaStreamSynchronize(st));
    return ENO_OK;
  }
  int rc;
  p.flat.grid_step = grid_step;
  if ((rc = css_prepare(ctx, p, params, eps, st))) return rc;
  p.flat.ts = p.seg_ts;
  p.flat.trace = trace;
  const double n_state = set_collective(ctx, &p, B);
  const DistCtx* dc = p.flat.world > 1 ? &ctx->dist : nullptr;
  const int egrid = std::min<int>((int)((N + 255) / 256), 2048);
  int worst = ENO_OK, launches = 0;
  std::string err;
  for (int i = T - 1; i >= 1; --i) {
    const float* gy_i = g_y + (size_t)i * BD;
    const float* ga_i = g_aux + (size_t)i * AB;
    css::k_css_seg_ts<<<1, 1, 0, st>>>(ts, i, p.seg_ts);
    css::k_flat_reset<<<1, 1, 0, st>>>(p.flat.ctrl, rtol, atol, 2, mxstep <= 0 ? INT_MAX : mxstep, trace ? trace_cap : 0, (float)n_state);
    css::k_css_adj_begin<<<egrid, 256, 0, st>>>(p.flat, ys + (size_t)i * BD, nullptr, gy_i, ga_i, i == T - 1);
    // t_bar needs func(ys[i], ts[i]): the forward half of one evaluation (scratch slot: stage buffer 1)
    if ((rc = css_rhs_launch(p, 0, css::ST_PROBE, p.flat.KS + N, ctx->sm_count, st, &err))) return css_fail(ctx, rc, err);
    css::k_css_tbar<<<1, 512, 0, st>>>(p.flat, p.net, gy_i, ga_i, ts_bar_out + i);
    if (dc) {   // t_bar is a sum over the rows of the concatenated batch
      if (dc->allreduce(p.flat.xbuf, 1, st) != 0) return css_fail(ctx, ENO_E_CUDA, "ncclAllReduce failed (t_bar)");
      css::k_css_tbar_apply<<<1, 1, 0, st>>>(p.flat, ts_bar_out + i);
    }
    GraphKey key{ws, B, desc->D, desc->H, desc->n_hidden, desc->aug, desc->reg_order, 1, 2, p.flat.world, p.flat.rank, grid_step > 0.f ? 1 : 0, 0};
    int status = run_loop(ctx, p, key, 1, stats ? stats + (T - 1 - i) : nullptr, &launches, st);
    if (status < 0) return status;
    if (status != ENO_OK && worst == ENO_OK) worst = status;
    p.flat.trace = nullptr;   // only the first segment is traced
  }
  css::k_css_adj_finish<<<egrid, 256, 0, st>>>(p.flat, g_y, g_aux, y0_bar_out, aux0_bar_out, eps_bar_out, ts_bar_out, params_bar_out, p.P);
  CSS_CUDA(ctx, cudaGetLastError());
  CSS_CUDA(ctx, cudaStreamSynchronize(st));
  return worst;
}

// eno_rhs for ConcatSquash dynamics.  mode 0: f;  1: the augmented closure named by desc->aug;  2: its vjp.
int rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int mode, int B, const float* y, float t, const float* params,
        const float* eps, const float* y_bar, const float* aux_bar, void* ws, size_t ws_bytes, float* dy_out, float* aux_out,
        float* ybar_dot_out, float* tbar_out, float* params_bar_out, float* eps_bar_dot_out, cudaStream_t st) {
  if (!desc_ok(desc)) return css_fail(ctx, ENO_E_UNSUPPORTED, "eno_rhs: unsupported ConcatSquash descriptor");
  if (!y || !params || !ws || !dy_out || B <= 0) return css_fail(ctx, ENO_E_ARG, "eno_rhs: null pointer or empty batch");
  eno_dynamics_desc d = *desc;
  if (mode == 0) d.aug = ENO_AUG_NONE;
  const int na = aux_count(d.aug);
  if (mode > 0 && (!eps || !aux_out)) return css_fail(ctx, ENO_E_ARG, "eno_rhs: augmented modes need eps and aux_out");
  if (mode == 2 && (!y_bar || !aux_bar || !ybar_dot_out || !tbar_out || !params_bar_out))
    return css_fail(ctx, ENO_E_ARG, "eno_rhs: mode 2 needs y_bar, aux_bar and all outputs");
  Plan p;
  plan_build(&d, B, 2, mode == 2 ? 1 : 0, ws, &p);
  if (ws_bytes < p.bytes) return css_fail(ctx, ENO_E_WORKSPACE, "eno_rhs: workspace too small");
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = css_prepare(ctx, p, params, eps, st))) return rc;
  const size_t N = p.flat.N, BD = (size_t)B * d.D, AB = (size_t)na * B;
  // solver time of the evaluation: ts[0] = tsign * t
  const float tau = p.flat.tsign * t;
  CSS_CUDA(ctx, cudaMemcpyAsync(p.seg_ts, &tau, 4, cudaMemcpyHostToDevice, st));
  p.flat.ts = p.seg_ts;
  css::k_flat_reset<<<1, 1, 0, st>>>(p.flat.ctrl, 1.f, 1.f, 2, INT_MAX, 0, (float)N);
  CSS_CUDA(ctx, cudaMemsetAsync(p.flat.Y, 0, N * 4, st));
  CSS_CUDA(ctx, cudaMemcpyAsync(p.flat.Y, y, BD * 4, cudaMemcpyDeviceToDevice, st));
  if (mode == 2) {
    CSS_CUDA(ctx, cudaMemcpyAsync(p.flat.Y + p.flat.oYB, y_bar, BD * 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(ctx, cudaMemcpyAsync(p.flat.Y + p.flat.oAUXB, aux_bar, AB * 4, cudaMemcpyDeviceToDevice, st));
  }
  std::string err;
  float* k = p.flat.KS;
  if ((rc = css_rhs_launch(p, 0, css::ST_PROBE, k, ctx->sm_count, st, &err))) return css_fail(ctx, rc, err);
  const float sgn = (mode == 2) ? -1.f : 1.f;   // the adjoint slot holds -f, +div, -r
  css::k_css_scale_copy<<<256, 256, 0, st>>>(k, sgn, BD, dy_out);
  if (na > 0 && aux_out) css::k_css_scale_copy<<<64, 256, 0, st>>>(k + p.flat.oAUX, sgn, AB, aux_out);
  if (mode == 2) {
    CSS_CUDA(ctx, cudaMemcpyAsync(ybar_dot_out, k + p.flat.oYB, BD * 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(ctx, cudaMemcpyAsync(tbar_out, k + p.flat.oT0, 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(ctx, cudaMemcpyAsync(params_bar_out, k + p.flat.oPR, p.P * 4, cudaMemcpyDeviceToDevice, st));
    if (eps_bar_dot_out) CSS_CUDA(ctx, cudaMemcpyAsync(eps_bar_dot_out, k + p.flat.oEB, BD * 4, cudaMemcpyDeviceToDevice, st));
  }
  CSS_CUDA(ctx, cudaGetLastError());
  CSS_CUDA(ctx, cudaStreamSynchronize(st));
  return ENO_OK;
}

// The iteration graphs of collective solves contain NCCL kernels bound to one communicator: they must not outlive it
// (eno_comm_init / eno_comm_destroy call this).
void drop_collective_graphs(eno_ctx* ctx) {
  if (!ctx || !ctx->css) return;
  CssState* s = static_cast<CssState*>(ctx->css);
  for (auto it = s->graphs.begin(); it != s->graphs.end();) {
    if (it->first.world > 1) { cudaGraphExecDestroy(it->second); it = s->graphs.erase(it); }
    else ++it;
  }
}

void destroy(eno_ctx* ctx) {
  if (!ctx || !ctx->css) return;
  CssState* s = static_cast<CssState*>(ctx->css);
  for (auto& kv : s->graphs) cudaGraphExecDestroy(kv.second);
  delete s;
  ctx->css = nullptr;
}

}  // namespace css
}  // namespace eno

namespace {
// jax.experimental.optimizers.adam (ffjord_tabular.py:527-529) with the weight-decay term of loss_fn (lam_w * 0.5 |theta|^2,
// ffjord_tabular.py:417-432) folded into the gradient.
__global__ void k_adam(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, size_t n,
                       float lr, float b1, float b2, float eps, float c1, float c2, float lam_w) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] + lam_w * x[i];
    const float mi = (1.f - b1) * gi + b1 * m[i];
    const float vi = (1.f - b2) * gi * gi + b2 * v[i];
    m[i] = mi;
    v[i] = vi;
    x[i] = x[i] - lr * (mi / c1) / (sqrtf(vi / c2) + eps);
  }
}
}  // namespace

extern "C" {

#ifdef ENO_PHASES
// development aid: the clock stamps of the last `n` (<= 64) GEMM launches, 8 words each (gemm_tc.cuh)
int32_t eno_debug_gemm_stamps(long long* out, unsigned int* n_launches) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, tc::g_gemm_stamp, sizeof(long long) * 64 * 8);
  cudaMemcpyFromSymbol(n_launches, tc::g_gemm_launch, sizeof(unsigned int));
  return 0;
}
#endif

int32_t eno_profile_read_gemm(eno_ctx* ctx, double* ms_out, double* flop_out, int64_t* count_out) {
  if (!ctx || !ms_out || !flop_out || !count_out) return ENO_E_ARG;
  CssState* S = state_of(ctx);
  *ms_out = S->prof.ms; *flop_out = S->prof.flops; *count_out = S->prof.count;
  S->prof.ms = S->prof.flops = 0.0; S->prof.count = 0;
  return ENO_OK;
}

int32_t eno_adam_update(eno_ctx* ctx, float* params, float* m, float* v, const float* grads, int64_t n, float lr, float b1,
                        float b2, float eps, int32_t step, float lam_w, void* stream) {
  if (!ctx || !params || !m || !v || !grads || n <= 0 || step < 0) return css_fail(ctx, ENO_E_ARG, "eno_adam_update: bad argument");
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  const float c1 = 1.f - powf(b1, (float)(step + 1)), c2 = 1.f - powf(b2, (float)(step + 1));
  k_adam<<<(int)std::min<int64_t>((n + 255) / 256, 2048), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      params, m, v, grads, (size_t)n, lr, b1, b2, eps, c1, c2, lam_w);
  CSS_CUDA(ctx, cudaGetLastError());
  return ENO_OK;
}

// Unit-test hook of the tensor-core contraction (declared in include/eno.h).
int32_t eno_gemm_tf32x3(eno_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B,
                        int32_t ldb, float* C, int32_t ldc, int32_t bn, int32_t splits, void* stream) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0 || splits < 1 || (bn != 64 && bn != 128))
    return css_fail(ctx, ENO_E_ARG, "eno_gemm_tf32x3: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ka = (K + 3) / 4 * 4;
  float *a2 = nullptr, *b2 = nullptr, *part = nullptr;
  CSS_CUDA(ctx, cudaMalloc(&a2, (size_t)2 * M * ka * 4));
  CSS_CUDA(ctx, cudaMalloc(&b2, (size_t)2 * N * ka * 4));
  CSS_CUDA(ctx, cudaMemsetAsync(a2, 0, (size_t)2 * M * ka * 4, st));
  CSS_CUDA(ctx, cudaMemsetAsync(b2, 0, (size_t)2 * N * ka * 4, st));
  tc::k_split_planes<<<1024, 256, 0, st>>>(A, M, K, lda, a2, ka, (size_t)M * ka, 0);
  tc::k_split_planes<<<1024, 256, 0, st>>>(B, N, K, ldb, b2, ka, (size_t)N * ka, 0);
  tc::GemmOperand oa{a2, M, K, ka, (size_t)M * ka, 0}, ob{b2, N, K, ka, (size_t)N * ka, 0};
  tc::EpiStore epi;
  if (splits > 1) {
    CSS_CUDA(ctx, cudaMalloc(&part, (size_t)splits * M * N * 4));
    epi = tc::EpiStore{part, M, N, N, (size_t)M * N};
  } else {
    epi = tc::EpiStore{C, M, N, ldc, 0};
  }
  std::string err;
  cudaError_t e = (bn == 128) ? tc::launch_gemm<128>(oa, ob, M, N, splits, nullptr, epi, st, &err)
                              : tc::launch_gemm<64>(oa, ob, M, N, splits, nullptr, epi, st, &err);
  if (e != cudaSuccess) {
    cudaFree(a2); cudaFree(b2); if (part) cudaFree(part);
    return css_fail(ctx, ENO_E_CUDA, "eno_gemm_tf32x3: " + (err.empty() ? std::string(cudaGetErrorString(e)) : err));
  }
  if (splits > 1) tc::k_sum_splits<<<256, 256, 0, st>>>(part, splits, (size_t)M * N, N, C, ldc);
  cudaError_t es = cudaStreamSynchronize(st);
  cudaFree(a2); cudaFree(b2); if (part) cudaFree(part);
  if (es != cudaSuccess) return css_fail(ctx, ENO_E_CUDA, std::string("eno_gemm_tf32x3: ") + cudaGetErrorString(es));
  return ENO_OK;
}

// Unit-test hook of the MN-major operand mode: C[m, n] = sum_k A[k, m] B[k, n], A [K][lda], B [K][ldb] row-major.
int32_t eno_gemm_tf32x3_mn(eno_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B,
                           int32_t ldb, float* C, int32_t ldc, int32_t bn, int32_t splits, void* stream) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0 || splits < 1 || (bn != 64 && bn != 128) || lda % 4 || ldb % 4)
    return css_fail(ctx, ENO_E_ARG, "eno_gemm_tf32x3_mn: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CSS_CUDA(ctx, cudaSetDevice(ctx->device));
  float *a2 = nullptr, *b2 = nullptr, *part = nullptr;
  CSS_CUDA(ctx, cudaMalloc(&a2, (size_t)2 * K * lda * 4));
  CSS_CUDA(ctx, cudaMalloc(&b2, (size_t)2 * K * ldb * 4));
  tc::k_split_planes<<<1024, 256, 0, st>>>(A, K, M, lda, a2, lda, (size_t)K * lda, 0);
  tc::k_split_planes<<<1024, 256, 0, st>>>(B, K, N, ldb, b2, ldb, (size_t)K * ldb, 0);
  std::string err;
  CUtensorMap ta, tb;
  if (!tc::make_operand_map(&ta, a2, K, M, lda, (size_t)K * lda, 32, &err, true) ||
      !tc::make_operand_map(&tb, b2, K, N, ldb, (size_t)K * ldb, 32, &err, true)) {
    cudaFree(a2); cudaFree(b2);
    return css_fail(ctx, ENO_E_CUDA, "eno_gemm_tf32x3_mn: " + err);
  }
  tc::EpiStore epi;
  if (splits > 1) {
    CSS_CUDA(ctx, cudaMalloc(&part, (size_t)splits * M * N * 4));
    epi = tc::EpiStore{part, M, N, N, (size_t)M * N};
  } else {
    epi = tc::EpiStore{C, M, N, ldc, 0};
  }
  cudaError_t e = (bn == 128) ? tc::launch_gemm_mn<128>(ta, tb, M, N, K, splits, 0, nullptr, epi, st)
                              : tc::launch_gemm_mn<64>(ta, tb, M, N, K, splits, 0, nullptr, epi, st);
  if (e == cudaSuccess && splits > 1) tc::k_sum_splits<<<256, 256, 0, st>>>(part, splits, (size_t)M * N, N, C, ldc);
  cudaError_t es = cudaStreamSynchronize(st);
  cudaFree(a2); cudaFree(b2); if (part) cudaFree(part);
  if (e != cudaSuccess || es != cudaSuccess)
    return css_fail(ctx, ENO_E_CUDA, std::string("eno_gemm_tf32x3_mn: ") + cudaGetErrorString(e != cudaSuccess ? e : es));
  return ENO_OK;
}

}  // extern "C"
