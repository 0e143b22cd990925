// Flat-state dopri5 driver (device-resident controller) and the element-wise / reduction kernels around the GEMM
// pipeline of css_rhs.cuh.  The solver state is ONE flat vector, as in the reference (ravel_pytree, lib/ode.py:742-747):
//   forward  : [ y (B D) | aux (n_aux B) ]
//   adjoint  : [ y | aux | y_bar | aux_bar | t0_bar (1) | eps_bar (B D) | params_bar (P) ]      lib/ode.py:1516-1519
// Reference behaviour restated: lib/ode.py:823-862 (loop), 86-104 (initial step), 518-538 (norm, controller),
// 56-84 + 849-850 (dense output), 1494-1529 (adjoint).
#pragma once
#include "css_rhs.cuh"

namespace eno {
namespace css {

__constant__ Tableau c_tab;   // dopri5 (the only tableau odeint() runs, lib/ode.py:746)

struct Flat {
  Ctrl* ctrl;
  const float* ts;      // solver times (the adjoint integrates over [-t1, -t0])
  size_t N, n_feed;     // state length; leading part the RHS reads
  float *Y, *F;         // [3][N] rings: states and their derivatives
  float *KS;            // [5][N] stage derivatives 1..5
  float *MID;           // [2][N]
  float *YS;            // [n_feed] stage input
  double* part;         // [4][n_part] per-block partial sums (lanes: part + l * n_part)
  int n_part;
  // collective solves (batch rows sharded over `world` ranks, css_engine.cu): the [t0_bar | params_bar] elements of the
  // adjoint state are REPLICATED (every rank holds the global value), everything else is this rank's rows
  int world, rank;
  double* gbuf;         // [world][2] per-rank sums of the sharded elements (all-gathered), then [2] replicated-block sums
  float* xbuf;          // [4][n_x] stage combinations (sol, err, mid, k6) of the replicated elements, all-reduced
  size_t n_x;           // 1 + P
  // fixed grid (odeint_grid, lib/ode.py:561-577, 864-903): > 0 = the RK4 step size; no error control
  float grid_step;
  float* out;           // [T][N]
  eno_trace_row* trace;
  float tsign;          // +1 forward, -1 adjoint (RHS time = tsign * solver time)
  // layout
  int B, D, n_aux;
  size_t oAUX, oYB, oAUXB, oT0, oEB, oPR;
};

enum { ST_STAGE = 0, ST_INIT0 = 1, ST_INIT1 = 2, ST_PROBE = 3 };

__device__ __forceinline__ float* k_slot(const Flat& a, int stage, int init_mode) {
  const Ctrl* c = a.ctrl;
  if (init_mode == ST_INIT1) return a.KS;                           // f(y0 + h0 f0)
  if (init_mode == ST_INIT0 || init_mode == ST_PROBE || stage == 0) return a.F + (size_t)c->cur * a.N;
  if (stage == 6) return a.F + (size_t)c->cand * a.N;
  return a.KS + (size_t)(stage - 1) * a.N;
}

__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__device__ __forceinline__ double block_reduce(double v, double* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sm[i];
  return r;   // valid in thread 0
}

// collective adjoint solves: is state element i replicated?  *j = its index in xbuf (0: t0_bar, 1 + p: params_bar[p])
__device__ __forceinline__ bool replicated(const Flat& a, size_t i, size_t* j) {
  if (a.world <= 1 || a.oPR == 0) return false;
  if (i >= a.oPR) { *j = 1 + (i - a.oPR); return true; }
  if (i == a.oT0) { *j = 0; return true; }
  return false;
}
__device__ __forceinline__ size_t replicated_at(const Flat& a, size_t j) { return j == 0 ? a.oT0 : a.oPR + (j - 1); }

// Programmatic dependent launch (gemm_tc.cuh:launch_pdl): every kernel of the iteration that is launched with the attribute
// starts with this - wait until the kernel before it has completed (nothing is read earlier), then let the kernel after it
// start its own prologue.  Both instructions are no-ops in a plain launch.
#define CSS_PDL_ENTRY()                                            \
  do {                                                             \
    asm volatile("griddepcontrol.wait;" ::: "memory");             \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

// ---- stage input + network input + gate vectors, one launch ---------------------------------------------------
__global__ void k_css_stage(Flat a, Net n, int stage, int init_mode) {
  CSS_PDL_ENTRY();
  Ctrl* c = a.ctrl;
  if (init_mode == ST_STAGE && c->done) return;
  const float dt = c->dt;
  float tau;
  if (init_mode == ST_INIT0 || init_mode == ST_PROBE) tau = a.ts[0];
  else if (init_mode == ST_INIT1) tau = a.ts[0] + c->h0;
  else tau = c->t + dt * c_tab.alpha[stage - 1];
  const float t = a.tsign * tau;
  const float* Yc = a.Y + (size_t)c->cur * a.N;
  const float* Fc = a.F + (size_t)c->cur * a.N;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t BD = (size_t)a.B * a.D;
  for (size_t i = gid; i < a.n_feed; i += gsz) {
    float ys;
    if (init_mode == ST_INIT0 || init_mode == ST_PROBE) ys = Yc[i];
    else if (init_mode == ST_INIT1) ys = fmaf(c->h0, Fc[i], Yc[i]);
    else {
      float acc = c_tab.beta[stage - 1][0] * Fc[i];
      for (int j = 1; j < stage; ++j) acc = fmaf(c_tab.beta[stage - 1][j], a.KS[(size_t)(j - 1) * a.N + i], acc);
      ys = fmaf(dt, acc, Yc[i]);
    }
    a.YS[i] = ys;
    if (i < BD) {
      const int b = (int)(i / a.D), d = (int)(i - (size_t)b * a.D);
      float hi, lo;
      tc::split_tf32(ys, hi, lo);
      const size_t o = (size_t)b * n.ldi[0] + d;
      n.X2[0][o] = hi;
      n.X2[0][n.xplane(0) + o] = lo;
    }
  }
  for (int l = 0; l < n.L; ++l) {
    for (size_t i = gid; i < (size_t)n.dout[l]; i += gsz) {
      const float wg = n.wg[l][i];
      const float g = sigmoidf_(wg * t + n.bg[l][i]);
      n.G[l][i] = g;
      n.G1[l][i] = g * (1.f - g) * wg;
      n.CC[l][i] = n.wb[l][i] * t;
    }
  }
  if (gid == 0) *n.tval = t;
}

// e_0 = eps into the stacked operands (constant along a solve)
__global__ void k_css_eps(Net n) {
  const size_t BD = (size_t)n.B * n.D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < BD; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    float hi, lo;
    tc::split_tf32(n.eps[i], hi, lo);
    const size_t o = (size_t)(n.Bp + b) * n.ldi[0] + d;
    n.X2[0][o] = hi;
    n.X2[0][n.xplane(0) + o] = lo;
  }
}

// ---- after the forward passes: per-sample scalars, and (adjoint) the seeds of the reverse sweep ------------------
// one warp per sample.  Cotangents: aux_bar rows of the stage input (constant blocks of the adjoint state).
__global__ void k_css_seed(Flat a, Net n) {
  CSS_PDL_ENTRY();
  if (a.ctrl->done) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n.B) return;
  const int b = warp, D = n.D, Ll = n.L - 1;
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
  if (lane == 0) { n.DIV[b] = div; n.RR[b] = rr / D; n.FRO[b] = fro / D; n.KIN[b] = kin / D; }
  if (n.mode != MODE_ADJ) return;
  // aux cotangents live behind the primal blocks: aux order (p, r) or (p, fro, kin)
  const float* cur = a.Y + (size_t)a.ctrl->cur * a.N;
  const float pb = cur[a.oAUXB + b];
  const float rb = n.fin ? 0.f : cur[a.oAUXB + (size_t)a.B + b];
  const float frob = n.fin ? cur[a.oAUXB + (size_t)a.B + b] : 0.f;
  const int ldo = n.ldo[Ll];
  for (int d = lane; d < D; d += 32) {
    const float g = n.G[Ll][d];
    // dp = -div: cotangent of J eps is -p_bar eps (+ fro_bar 2 J eps / D); direct eps cotangent -p_bar J eps
    const float ve = -pb * ep[d] + frob * 2.f * je[d] / D;
    n.V[Ll][(size_t)b * ldo + d] = ve;
    n.EBD[(size_t)b * n.ldD + d] = -pb * je[d];
    float hi, lo;
    tc::split_tf32(ve * g, hi, lo);
    size_t o = (size_t)(n.Bp + b) * ldo + d;
    n.AB2[Ll][o] = hi; n.AB2[Ll][n.abplane(Ll) + o] = lo;
    if (n.nq == 2) {
      const float vd = rb * 2.f * y1[d] / D;
      n.V[Ll][((size_t)n.B + b) * ldo + d] = vd;
      tc::split_tf32(vd * g, hi, lo);
      o = (size_t)(2 * n.Bp + b) * ldo + d;
      n.AB2[Ll][o] = hi; n.AB2[Ll][n.abplane(Ll) + o] = lo;
      }
  }
}

// seed of the primal reverse sweep: u_bar of the last layer = y_bar (+ cotangent of d_0 = f) (+ kin term)
__global__ void k_css_seed2(Flat a, Net n) {
  CSS_PDL_ENTRY();
  if (a.ctrl->done) return;
  const size_t BD = (size_t)n.B * n.D;
  const int Ll = n.L - 1, ldo = n.ldo[Ll];
  const float* cur = a.Y + (size_t)a.ctrl->cur * a.N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < BD; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    float w = a.YS[a.oYB + i];
    if (n.nq == 2) w += n.XB1[((size_t)n.B + b) * n.ldD + d];
    if (n.fin) w += cur[a.oAUXB + 2 * (size_t)a.B + b] * 2.f * n.F[(size_t)b * n.ldD + d] / n.D;
    n.WW[Ll][(size_t)b * ldo + d] = w;
    float lb = w * n.G[Ll][d];
    if (n.nq == 2) lb += n.V[Ll][((size_t)n.B + b) * ldo + d] * n.G1[Ll][d];
    float hi, lo;
    tc::split_tf32(lb, hi, lo);
    const size_t o = (size_t)b * ldo + d;
    n.AB2[Ll][o] = hi; n.AB2[Ll][n.abplane(Ll) + o] = lo;
  }
}

// ---- column sums over the batch feeding the gate / bias gradients ----------------------------------------------
// grid (col blocks of 32, row chunks, layers), block (32, 8).  CR[l][chunk][q][col]:
//   q 0: sum_b (w lin + v_e a1_e + v_d a1_d)   -> G_bar        q 1: sum_b v_d lin   -> G1_bar
//   q 2: sum_b w                                               q 3: sum_b v_d
__global__ void k_css_colred(Flat a, Net n) {
  CSS_PDL_ENTRY();
  if (a.ctrl->done) return;
  const int l = blockIdx.z, dout = n.dout[l], ldo = n.ldo[l];
  const int col = blockIdx.x * 32 + threadIdx.x;
  __shared__ float sm[4][8][33];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < dout) {
    const int per = (n.B + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(n.B, r0 + per);
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const size_t o = (size_t)r * ldo + col;
      const float w = n.WW[l][o], lin = n.LIN[l][o];
      s0 = fmaf(w, lin, s0);
      s2 += w;
      if (n.nq >= 1) s0 = fmaf(n.V[l][o], n.A1[l][o], s0);
      if (n.nq == 2) {
        const size_t o2 = o + (size_t)n.B * ldo;
        const float vd = n.V[l][o2];
        s0 = fmaf(vd, n.A1[l][o2], s0);
        s1 = fmaf(vd, lin, s1);
        s3 += vd;
      }
    }
  }
  sm[0][threadIdx.y][threadIdx.x] = s0; sm[1][threadIdx.y][threadIdx.x] = s1;
  sm[2][threadIdx.y][threadIdx.x] = s2; sm[3][threadIdx.y][threadIdx.x] = s3;
  __syncthreads();
  if (threadIdx.y < 4 && col < dout) {
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += sm[threadIdx.y][i][threadIdx.x];
    n.CR[(((size_t)l * n.cr_chunks + blockIdx.y) * 4 + threadIdx.y) * n.hmax + col] = acc;
  }
}

// gate / bias gradients of every layer into the stage derivative's params block; per-column t_bar pieces
constexpr int kCrChunks = 16;   // row chunks of k_css_colred (Net::cr_chunks)

__global__ void k_css_colfin(Flat a, Net n, int stage, int init_mode, float* pr_override) {
  CSS_PDL_ENTRY();
  if (init_mode == ST_STAGE && a.ctrl->done) return;
  float* pr = pr_override ? pr_override : k_slot(a, stage, init_mode) + a.oPR;
  const float t = *n.tval;
  const int l = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n.dout[l]) return;
  float q[4] = {0.f, 0.f, 0.f, 0.f};
  float v[kCrChunks][4];
#pragma unroll
  for (int ch = 0; ch < kCrChunks; ++ch)     // all loads in flight before the first add
#pragma unroll
    for (int k = 0; k < 4; ++k) v[ch][k] = n.CR[(((size_t)l * kCrChunks + ch) * 4 + k) * n.hmax + col];
#pragma unroll
  for (int ch = 0; ch < kCrChunks; ++ch)
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] += v[ch][k];
  const float G = n.G[l][col], wg = n.wg[l][col], wb = n.wb[l][col];
  const float dG = G * (1.f - G);
  const float gbar = q[0] * dG + q[1] * dG * (1.f - 2.f * G) * wg;   // cotangent of the gate pre-activation
  float* p = pr + n.poff[l];
  const size_t din_dout = (size_t)n.din[l] * n.dout[l];
  p[col] = G * q[2] + n.G1[l][col] * q[3];                                  // b_bar = column sum of lin_bar
  p[(size_t)n.dout[l] + din_dout + col] = t * q[2] + q[3];                  // w_bias_bar
  p[(size_t)2 * n.dout[l] + din_dout + col] = gbar;                          // b_gate_bar
  p[(size_t)3 * n.dout[l] + din_dout + col] = gbar * t + q[1] * dG;          // w_gate_bar
  n.TC[(size_t)l * n.hmax + col] = gbar * wg + wb * q[2];
}

// ---- stage derivative of every block --------------------------------------------------------------------------
// plain: k = f.   aug: k = (f, -div [, r, fro, kin]).   adjoint: (-f, +div, -r | vjp_y, 0 | vjp_t | vjp_eps | vjp_params)
__global__ void k_css_assemble(Flat a, Net n, int stage, int init_mode, float* k_override) {
  CSS_PDL_ENTRY();
  if (init_mode == ST_STAGE && a.ctrl->done) return;
  float* k = k_override ? k_override : k_slot(a, stage, init_mode);
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t BD = (size_t)n.B * n.D, Bz = (size_t)n.B;
  const float sgn = (n.mode == MODE_ADJ) ? -1.f : 1.f;
  for (size_t i = gid; i < BD; i += gsz) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    k[i] = sgn * n.F[(size_t)b * n.ldD + d];
    if (n.mode == MODE_ADJ) {
      k[a.oYB + i] = n.YBDOT[(size_t)b * n.ldD + d];
      k[a.oEB + i] = (n.nq >= 1) ? n.XB1[(size_t)b * n.ldD + d] + n.EBD[(size_t)b * n.ldD + d] : 0.f;
    }
  }
  for (size_t i = gid; i < (size_t)a.n_aux * Bz; i += gsz) {
    const int q = (int)(i / Bz), b = (int)(i - (size_t)q * Bz);
    float v;
    if (q == 0) v = -n.DIV[b];
    else if (a.n_aux == 4) v = (q == 1) ? n.RR[b] : (q == 2 ? n.FRO[b] : n.KIN[b]);   // all_aug_dynamics: (p, r, fro, kin)
    else if (n.fin) v = (q == 1) ? n.FRO[b] : n.KIN[b];
    else v = (n.nq == 2) ? n.RR[b] : 0.f;
    k[a.oAUX + i] = sgn * v;
    if (n.mode == MODE_ADJ) k[a.oAUXB + i] = 0.f;
  }
  if (n.mode != MODE_ADJ) return;
  // weight gradients: fixed-order sum of the per-stream partials
  for (int l = 0; l < n.L; ++l) {
    const size_t cnt = (size_t)n.din[l] * n.dout[l];
    float* dst = k + a.oPR + n.poff[l] + n.dout[l];
    for (size_t i = gid; i < cnt; i += gsz) {
      float acc = n.WP[l][i];
      for (int s = 1; s <= n.nq; ++s) acc += n.WP[l][(size_t)s * cnt + i];
      dst[i] = acc;
    }
  }
  if (blockIdx.x == 0) {   // t_bar: fixed-shape block sum of the per-column pieces
    __shared__ double sm[32];
    double acc = 0.0;
    for (int l = 0; l < n.L; ++l)
      for (int c = threadIdx.x; c < n.dout[l]; c += blockDim.x) acc += (double)n.TC[(size_t)l * n.hmax + c];
    const double tot = block_reduce(acc, sm);
    if (threadIdx.x == 0) k[a.oT0] = (float)tot;
  }
}

// ---- initial step size (lib/ode.py:86-104), two phases ----------------------------------------------------------
__device__ __forceinline__ void init_control(const Flat& a, Ctrl* c, int phase, double t0s, double t1s) {
  if (phase == 0) {
    const float d0 = sqrtf((float)t0s), d1 = sqrtf((float)t1s);
    c->d0 = d0; c->d1 = d1;
    c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
  } else {
    const float h0 = c->h0, d1 = c->d1;
    const float d2 = sqrtf((float)t0s) / h0;
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
    else h1 = powf(0.01f / (d1 + d2), 1.0f / ((float)c_tab.init_order + 1.0f));
    c->dt = fminf(100.f * h0, h1);
    if (isnan(h0) || isnan(h1)) c->dt = nanf("");
    c->t = a.ts[0];
    c->last_t = c->t;
    c->step_dt = 0.f;
    c->iters_tgt = 0; c->tgt = 1;
    c->n_iters = 0; c->n_accepted = 0; c->nfe = 2;
    c->out_begin = 1; c->out_end = 1;
    c->done = 0; c->status = ENO_OK;
    advance_targets(c, a.ts);
  }
}

// sums over ranks in rank order (every rank holds the same gathered values, so every controller takes the same decision)
__device__ __forceinline__ double gathered_sum(const Flat& a, int lane) {
  double s = 0.0;
  for (int r = 0; r < a.world; ++r) s += a.gbuf[2 * r + lane];
  return s + a.gbuf[2 * a.world + lane];
}

__global__ void k_flat_init(Flat a, int phase) {
  __shared__ double sm[32];
  __shared__ double dred[512];
  Ctrl* c = a.ctrl;
  const float rtol = c->rtol, atol = c->atol;
  const float* y = a.Y + (size_t)c->cur * a.N;
  const float* f0 = a.F + (size_t)c->cur * a.N;
  double s0 = 0.0, s1 = 0.0, r0 = 0.0, r1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (size_t)gridDim.x * blockDim.x) {
    const float scale = atol + fabsf(y[i]) * rtol;
    size_t j;
    const bool rep = replicated(a, i, &j);
    if (phase == 0) {
      const float q0 = y[i] / scale, q1 = f0[i] / scale;
      if (rep) { r0 += (double)(q0 * q0); r1 += (double)(q1 * q1); }
      else { s0 += (double)(q0 * q0); s1 += (double)(q1 * q1); }
    } else {
      const float q = (a.KS[i] - f0[i]) / scale;
      if (rep) r0 += (double)(q * q);
      else s0 += (double)(q * q);
    }
  }
  const double b0 = block_reduce(s0, sm);
  const double b1 = block_reduce(s1, sm);
  const double b2 = block_reduce(r0, sm);
  const double b3 = block_reduce(r1, sm);
  if (threadIdx.x == 0) {
    a.part[blockIdx.x] = b0; a.part[a.n_part + blockIdx.x] = b1;
    a.part[2 * a.n_part + blockIdx.x] = b2; a.part[3 * a.n_part + blockIdx.x] = b3;
  }
  if (!last_block(&c->ticket)) return;
  const double t0s = block_sum_global(a.part, gridDim.x, dred);
  __syncthreads();
  const double t1s = block_sum_global(a.part + a.n_part, gridDim.x, dred);
  __syncthreads();
  const double t2s = block_sum_global(a.part + 2 * a.n_part, gridDim.x, dred);
  __syncthreads();
  const double t3s = block_sum_global(a.part + 3 * a.n_part, gridDim.x, dred);
  if (threadIdx.x != 0) return;
  c->ticket = 0;
  if (a.world > 1) {   // k_flat_init_control finishes after the all-gather
    a.gbuf[2 * a.rank] = t0s; a.gbuf[2 * a.rank + 1] = t1s;
    a.gbuf[2 * a.world] = t2s; a.gbuf[2 * a.world + 1] = t3s;
    return;
  }
  init_control(a, c, phase, t0s + t2s, t1s + t3s);
}

__global__ void k_flat_init_control(Flat a, int phase) {
  init_control(a, a.ctrl, phase, gathered_sum(a, 0), gathered_sum(a, 1));
}

// ---- end of an iteration: candidate, error estimate, mid-point, norm, controller ------------------------------
// Collective solves: the stage derivatives of the replicated elements are per-rank PARTIAL sums over this rank's rows
// (stages 1..6; the first stage is the global derivative the previous iteration left in the F ring), and everything
// downstream of the batch sum is linear: k_flat_xcomb forms the three stage combinations + k6, ONE all-reduce sums
// them, k_flat_finish completes the replicated elements from the sums and stores the global k6 into the F ring.
// Fixed-grid controller (lib/ode.py:877-897): every step is taken; next_t = min(t + step_size, ts[-1]), dt = next_t - t,
// nfe += 4.  Single thread.
__device__ __forceinline__ void finish_grid_step(Ctrl* cp, const float* __restrict__ ts, float step, eno_trace_row* trace) {
  Ctrl c = *cp;
  const float t_end = ts[c.T - 1];
  if (trace && c.n_iters < c.trace_cap) {
    eno_trace_row r;
    r.t = c.t; r.dt = c.dt; r.ratio = 0.f; r.accepted = 1.0f;
    trace[c.n_iters] = r;
  }
  c.ratio = 0.f;
  c.out_begin = c.out_end;
  const int old_prev = c.prev;
  c.prev = c.cur;
  c.cur = c.cand;
  c.cand = old_prev;
  c.mid_acc ^= 1;
  c.last_t = c.t;
  c.step_dt = c.dt;
  c.t = fminf(c.t + step, t_end);
  c.dt = fminf(c.t + step, t_end) - c.t;
  c.n_accepted += 1;
  c.iters_tgt += 1;
  c.n_iters += 1;
  c.nfe += 4;
  advance_targets_local(c, ts);
  *cp = c;
}

// start of a fixed-grid solve: F[cur] = f(y0, t0) has been evaluated (the first stage of the first step)
__global__ void k_flat_grid_begin(Flat a) {
  Ctrl* c = a.ctrl;
  const float t0 = a.ts[0], t_end = a.ts[c->T - 1];
  c->t = t0;
  c->last_t = t0;
  c->step_dt = 0.f;
  c->dt = fminf(t0 + a.grid_step, t_end) - t0;
  c->iters_tgt = 0; c->tgt = 1;
  c->n_iters = 0; c->n_accepted = 0; c->nfe = 0;
  c->out_begin = 1; c->out_end = 1;
  c->done = 0; c->status = ENO_OK;
  advance_targets(c, a.ts);
}

__global__ void k_flat_xcomb(Flat a) {
  const Ctrl* c = a.ctrl;
  if (c->done) return;
  const size_t N = a.N;
  const float* K6 = a.F + (size_t)c->cand * N;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n_x; j += (size_t)gridDim.x * blockDim.x) {
    const size_t i = replicated_at(a, j);
    float as = 0.f, ae = 0.f, am = 0.f;
#pragma unroll
    for (int s = 1; s < 6; ++s) {
      const float k = a.KS[(size_t)(s - 1) * N + i];
      as = fmaf(c_tab.c_sol[s], k, as);
      ae = fmaf(c_tab.c_err[s], k, ae);
      am = fmaf(c_tab.c_mid[s], k, am);
    }
    const float k6 = K6[i];
    a.xbuf[j] = fmaf(c_tab.c_sol[6], k6, as);
    a.xbuf[a.n_x + j] = fmaf(c_tab.c_err[6], k6, ae);
    a.xbuf[2 * a.n_x + j] = fmaf(c_tab.c_mid[6], k6, am);
    a.xbuf[3 * a.n_x + j] = k6;
  }
}

__global__ void k_flat_finish(Flat a) {
  __shared__ double sm[32];
  __shared__ double dred[512];
  CSS_PDL_ENTRY();
  Ctrl* c = a.ctrl;
  if (c->done) return;
  const float dt = c->dt, rtol = c->rtol, atol = c->atol;
  const size_t N = a.N;
  const float* Yc = a.Y + (size_t)c->cur * N;
  const float* K0 = a.F + (size_t)c->cur * N;
  float* Yn = a.Y + (size_t)c->cand * N;
  float* K6 = a.F + (size_t)c->cand * N;
  float* Mn = a.MID + (size_t)(c->mid_acc ^ 1) * N;
  double acc = 0.0, accr = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
    float as = 0.f, ae = 0.f, am = 0.f;
    size_t j;
    const bool rep = replicated(a, i, &j);
    if (rep) {
      const float k0 = K0[i];
      as = fmaf(c_tab.c_sol[0], k0, a.xbuf[j]);
      ae = fmaf(c_tab.c_err[0], k0, a.xbuf[a.n_x + j]);
      am = fmaf(c_tab.c_mid[0], k0, a.xbuf[2 * a.n_x + j]);
      K6[i] = a.xbuf[3 * a.n_x + j];
    } else {
      float kk[7];
      kk[0] = K0[i];
#pragma unroll
      for (int s = 1; s < 6; ++s) kk[s] = a.KS[(size_t)(s - 1) * N + i];
      kk[6] = K6[i];
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        as = fmaf(c_tab.c_sol[s], kk[s], as);
        ae = fmaf(c_tab.c_err[s], kk[s], ae);
        am = fmaf(c_tab.c_mid[s], kk[s], am);
      }
    }
    const float y = Yc[i];
    const float y1 = fmaf(dt, as, y);
    const float err = dt * ae;
    const float tol = atol + rtol * fmaxf(fabsf(y), fabsf(y1));
    const float q = err / tol;
    if (rep) accr += (double)(q * q);
    else acc += (double)(q * q);
    Yn[i] = y1;
    Mn[i] = fmaf(dt, am, y);
  }
  const double b0 = block_reduce(acc, sm);
  if (threadIdx.x == 0) a.part[blockIdx.x] = b0;
  if (a.world > 1) {
    const double b1 = block_reduce(accr, sm);
    if (threadIdx.x == 0) a.part[a.n_part + blockIdx.x] = b1;
  }
  if (!last_block(&c->ticket)) return;
  const double sum = block_sum_global(a.part, gridDim.x, dred);
  if (a.world > 1) {   // k_flat_control finishes after the all-gather
    __syncthreads();
    const double sumr = block_sum_global(a.part + a.n_part, gridDim.x, dred);
    if (threadIdx.x == 0) {
      c->ticket = 0;
      a.gbuf[2 * a.rank] = sum; a.gbuf[2 * a.rank + 1] = 0.0;
      a.gbuf[2 * a.world] = sumr; a.gbuf[2 * a.world + 1] = 0.0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    c->ticket = 0;
    const float ratio = (float)(sum / (double)c->n_state);
    if (a.grid_step > 0.f) finish_grid_step(c, a.ts, a.grid_step, a.trace);
    else finish_iteration(c, a.ts, ratio, dt, c_tab, a.trace);
  }
}

__global__ void k_flat_control(Flat a) {
  Ctrl* c = a.ctrl;
  if (c->done) return;
  const float ratio = (float)(gathered_sum(a, 0) / (double)c->n_state);
  if (a.grid_step > 0.f) finish_grid_step(c, a.ts, a.grid_step, a.trace);
  else finish_iteration(c, a.ts, ratio, c->dt, c_tab, a.trace);
}

// ---- dense output for the targets the last iteration left behind (lib/ode.py:56-84, 849-850) -------------------
__global__ void k_flat_out(Flat a) {
  CSS_PDL_ENTRY();
  const Ctrl* c = a.ctrl;
  const int ob = c->out_begin, oe = c->out_end;
  if (oe <= ob) return;
  const float t = c->t, last_t = c->last_t, dt = c->step_dt;
  const bool degenerate = (c->n_accepted == 0);
  const size_t N = a.N;
  const float* Yp = a.Y + (size_t)c->prev * N;
  const float* Yc = a.Y + (size_t)c->cur * N;
  const float* Mc = a.MID + (size_t)c->mid_acc * N;
  const float* Fp = a.F + (size_t)c->prev * N;
  const float* Fc = a.F + (size_t)c->cur * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
    float ca, cb, cc, cd, ce;
    if (degenerate) {
      ca = cb = cc = cd = ce = a.out[i];
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
      if (a.grid_step > 0.f && !degenerate) v = Yc[i];   // the last grid step lands on ts[-1]: the state itself
      a.out[(size_t)o * N + i] = v;
    }
  }
}

__global__ void k_flat_reset(Ctrl* c, float rtol, float atol, int T, int mxstep, int trace_cap, float n_state) {
  memset(c, 0, sizeof(Ctrl));
  c->rtol = rtol; c->atol = atol; c->T = T; c->mxstep = mxstep; c->trace_cap = trace_cap; c->n_state = n_state;
  c->cur = 0; c->prev = 2; c->cand = 1; c->mid_acc = 0;
}

// ---- initial states --------------------------------------------------------------------------------------------
// forward: Y[0] = out[0] = (y0, aux0)
__global__ void k_css_fwd_begin(Flat a, const float* y0, const float* aux0) {
  const size_t BD = (size_t)a.B * a.D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (i < BD) ? y0[i] : (aux0 ? aux0[i - BD] : 0.f);
    a.Y[i] = v;
    a.out[i] = v;
  }
}

// adjoint segment i: state (ys[i], aux[i] | y_bar, aux_bar | t0_bar | eps_bar | params_bar), lib/ode.py:1516-1519.
// first segment: cotangents start from g[i]; later ones continue from the previous segment's result (carry in out[1])
// plus g[i] (lib/ode.py:1522).  t0_bar is completed by k_css_tbar.
__global__ void k_css_adj_begin(Flat a, const float* ys_i, const float* aux_i, const float* gy_i, const float* gaux_i, int first) {
  const size_t BD = (size_t)a.B * a.D, AB = (size_t)a.n_aux * a.B;
  const float* carry = a.out + a.N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (size_t)gridDim.x * blockDim.x) {
    float v;
    if (i < BD) v = ys_i[i];
    else if (i < a.oYB) v = aux_i ? aux_i[i - BD] : 0.f;
    else if (i < a.oAUXB) v = gy_i[i - a.oYB] + (first ? 0.f : carry[i]);
    else if (i < a.oT0) v = gaux_i[i - a.oAUXB] + (first ? 0.f : carry[i]);
    else v = first ? 0.f : carry[i];
    (void)AB;
    a.Y[i] = v;
    a.out[i] = v;
  }
}

// t_bar = <func(ys[i], ts[i]), g[i]> (lib/ode.py:1513); t0_bar -= t_bar.  One block.
__global__ void k_css_tbar(Flat a, Net n, const float* gy_i, const float* gaux_i, float* tbar_out) {
  __shared__ double sm[32];
  const size_t BD = (size_t)n.B * n.D;
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < BD; i += blockDim.x) {
    const int b = (int)(i / n.D), d = (int)(i - (size_t)b * n.D);
    acc += (double)(n.F[(size_t)b * n.ldD + d] * gy_i[i]);
  }
  for (int b = threadIdx.x; b < n.B; b += blockDim.x) {
    acc += (double)(-n.DIV[b] * gaux_i[b]);
    if (n.fin) acc += (double)(n.FRO[b] * gaux_i[n.B + b]) + (double)(n.KIN[b] * gaux_i[2 * n.B + b]);
    else if (n.nq == 2) acc += (double)(n.RR[b] * gaux_i[n.B + b]);
  }
  const double tot = block_reduce(acc, sm);
  if (threadIdx.x == 0) {
    const float tb = (float)tot;
    if (a.world > 1) { a.xbuf[0] = tb; return; }   // this rank's rows only: all-reduced, then k_css_tbar_apply
    *tbar_out = tb;
    a.Y[a.oT0] -= tb;
    a.out[a.oT0] = a.Y[a.oT0];
  }
}

__global__ void k_css_tbar_apply(Flat a, float* tbar_out) {
  const float tb = a.xbuf[0];
  *tbar_out = tb;
  a.Y[a.oT0] -= tb;
  a.out[a.oT0] = a.Y[a.oT0];
}

// results of the whole backward pass from the last segment's final state (out row 1)
__global__ void k_css_adj_finish(Flat a, const float* gy0, const float* gaux0, float* y0_bar, float* aux0_bar, float* eps_bar,
                                 float* t0_bar, float* params_bar, size_t P) {
  const float* fin = a.out + a.N;
  const size_t BD = (size_t)a.B * a.D, AB = (size_t)a.n_aux * a.B;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gid; i < BD; i += gsz) {
    y0_bar[i] = fin[a.oYB + i] + gy0[i];
    if (eps_bar) eps_bar[i] = fin[a.oEB + i];
  }
  for (size_t i = gid; i < AB; i += gsz) aux0_bar[i] = fin[a.oAUXB + i] + gaux0[i];
  for (size_t i = gid; i < P; i += gsz) params_bar[i] = fin[a.oPR + i];
  if (gid == 0) *t0_bar = fin[a.oT0];
}

// solver times of adjoint segment i: [-ts[i], -ts[i-1]] (lib/ode.py:1518)
__global__ void k_css_seg_ts(const float* ts, int i, float* seg) {
  seg[0] = -ts[i];
  seg[1] = -ts[i - 1];
}

__global__ void k_css_scale_copy(const float* src, float s, size_t n, float* dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = s * src[i];
}

}  // namespace css
}  // namespace eno
