// Adjoint (backward) solve: dopri5 on the augmented flat state
//     [ y (B*D) | r (B) | y_bar (B*D) | r_bar (B) | t0_bar (1) | eps_bar (B*D, == 0) | params_bar (P) ]
// of lib/ode.py:1516-1519 for rev_func = aug_dynamics('our') (mnist.py:302-308).
//
// Two launches per loop iteration:
//   k_adj_rows  : per-sample blocks (y, r, y_bar, r_bar).  A CTA owns R samples and runs
//                 all 6 new stages out of shared memory; the batch-summed blocks cannot be
//                 finished here, so each stage leaves GEMM factors in global memory.
//   k_adj_pgrad : contracts the factors over the batch into the 6 stage derivatives of
//                 params_bar (one CTA = one 32x32 output tile, 6 thread groups = 6 stages),
//                 forms candidate / error / mid-point of that block, and its last CTA
//                 reduces the error norm over ALL blocks and runs the controller.
// The shared blocks never feed back into the RHS, so no stage buffers are kept for them.
#pragma once
#include "dist.cuh"
#include <cuda_pipeline.h>

#include <string>

#include "adj_device.cuh"
#include "gemm_tc.cuh"
#include "tableaux.cuh"
#include "fwd_kernels.cuh"

namespace eno {

constexpr int kAdjStages = 7;    // dopri5 (odeint() is always dopri5, lib/ode.py:746)
constexpr int kSlots = 6;        // new RHS evaluations per iteration
constexpr int kTile = 32;        // pgrad output tile edge
constexpr int kPStages = 4;      // cp.async stages in flight per pgrad thread group (8 x 4 measured best of 4x6, 8x4, 16x3)
constexpr int kPRows = 8;        // contraction rows per stage
constexpr int kGroup = 64;       // threads per stage group in pgrad (8x8 threads x 4x4 outputs)

enum AdjMode { ADJ_STEP = 0, ADJ_INIT0 = 1, ADJ_INIT1 = 2, ADJ_RHS = 3 };

struct AdjScalars {   // ring state of the single t0_bar element + segment bookkeeping
  float t0[3];
  float ft0[3];
  float midt0[2];
  float t0_out;       // t0_bar after the segment
  float seg_ts[2];    // [-ts[i], -ts[i-1]]
};

struct AdjArgs {
  Ctrl* ctrl;
  AdjScalars* sc;
  MlpWeights w;
  int B, P, np;            // np = Taylor passes whose factors exist (1, 2 or 3)
  int n_rp;                // entries of the per-sample partial-sum arrays over ALL ranks: world * n_loc
  int n_loc;               // ... of this rank: B (streamed) or kC*B (cluster)
  double* xrec;            // [world] records of 6*n_loc doubles: rp0 | t_bar pieces of the 6 slots (floats) | rp1 | rp2
                           // - everything a rank contributes to the controller besides params_bar, so that the
                           //   multi-GPU exchange is ONE all-gather
  size_t N;                // B * D
  float n_state;
  // per-sample rings
  float *Z, *ZB;           // [3][N]
  float *FZ, *FZB;         // [3][N]
  float *RS, *FR;          // [3][na][B]
  float* RB;               // [na][B]  r_bar (constant during a segment)
  float *MIDZB;            // [2][N]   (only y_bar's interpolant is ever read)
  // "fin" dynamics (ENO_AUG_FIN): two scalar states per sample and the eps_bar block
  int na;                  // scalar states per sample: 1, or 2 (fro, kin)
  int fin;
  const float* eps;        // [N]
  float *ZE, *FZE;         // [3][N]   eps_bar ring and its derivative
  float* MIDZE;            // [2][N]
  float* EBAR;             // [N]      eps_bar after the segment
  // shared block
  float *PB, *FP;          // [3][P]
  float* MIDP;             // [2][P]
  // factors
  FactorSlot slot[kSlots];
  FactorTc tc;             // tc.X non-null: ADJ_STEP factors go to the tensor-core sinks (adj_device.cuh), the stage
  float* KP;               // contractions [kSlots][P] come from the batched tcgen05 GEMM and k_adj_pcombine finishes
  float *tcX_ws, *tcY_ws;  // workspace blocks behind tc.X / tc.Y (host side bookkeeping)
  // partial sums
  double *rowpart0, *rowpart1, *rowpart2;  // per-sample partials of THIS rank's rows (views into rp*_all)
  const double *rp0_all, *rp1_all, *rp2_all;   // rank 0's entries; rank r's are 6*n_loc doubles further (block_sum_records)
  const float* tbar_all[kSlots];           // same for the per-sample t_bar pieces of every slot (12*n_loc floats per rank)
  double *ppart0, *ppart1;                 // [pgrad grid]
  float* pcomb;                            // [4][P] multi-GPU: this rank's share of the stage combinations
  int dist;                                // 1: multi-GPU split (local contraction | exchange | finish)
  unsigned char* peer_base;                // non-null: the exchange runs inside k_adj_pfinish_peer over NVLink peer memory
                                           // (dist.cuh); this rank's exchange buffer (its PeerView lives at kPeerOffView)
  size_t peer_pcomb_off;                   // byte offset of pcomb [2 parities][4 P] inside every rank's exchange buffer
  int p_lo, p_cnt;                         // params_bar slice this rank owns in peer mode
  // segment inputs / outputs
  const float* g_y;        // [N]   g[i] (for t_bar)
  const float* g_r;        // [na][B]
  float* YBAR;             // [N]   y_bar after the segment (+ g[i-1])
  const float* g_y_prev;   // [N]   g[i-1]
  float* POUT;             // [P]
  float* tbar_out;         // scalar: ts_bar[i]
  // eno_rhs outputs
  float *rhs_dy, *rhs_r, *rhs_zb, *rhs_eb;
  eno_trace_row* trace;
};

struct AdjSmem {
  float *z, *zb, *kz, *kzb, *xz, *xzb, *q;
  float *r, *rb, *kr, *xr;
  AdjScratch m;
  double* dred;
};

template <int R>
__device__ __forceinline__ AdjSmem carve_adj(float* sm, int D, int H, int stages) {
  AdjSmem a;
  a.dred = reinterpret_cast<double*>(sm); sm += 2 * kThreads;
  a.z = sm; sm += R * D;
  a.zb = sm; sm += R * D;
  a.kz = sm; sm += (size_t)stages * R * D;
  a.kzb = sm; sm += (size_t)stages * R * D;
  a.xz = sm; sm += R * D;
  a.xzb = sm; sm += R * D;
  a.m.sc.s = sm; sm += R * D;
  a.m.sc.u = sm; sm += R * D;
  a.m.f = sm; sm += R * D;
  a.m.y1 = sm; sm += R * D;
  a.m.y2 = sm; sm += R * D;
  a.m.gf = sm; sm += R * D;
  a.m.gy1 = sm; sm += R * D;
  a.m.zacc = sm; sm += R * D;
  a.q = a.m.sc.u;  // free once a stage has finished
  a.m.sc.h = sm; sm += R * H;
  a.m.sc.a1 = sm; sm += R * H;
  a.m.sc.h1 = sm; sm += R * H;
  a.m.sc.a2 = sm; sm += R * H;
  a.m.sc.h2 = sm; sm += R * H;
  a.m.ga = sm; sm += R * H;
  a.m.ga1 = sm; sm += R * H;
  a.m.ga2 = sm; sm += R * H;
  a.m.sc.red = sm; sm += (size_t)kRedFloats * R;
  a.r = sm; sm += 4 * ((R + 3) / 4);
  a.rb = sm; sm += 4 * ((R + 3) / 4);
  a.xr = sm; sm += 4 * ((R + 3) / 4);
  a.m.rdot = sm; sm += 4 * ((R + 3) / 4);
  a.kr = sm; sm += 4 * ((stages * R + 3) / 4);
  return a;
}

inline size_t adj_smem_bytes(int R, int D, int H, int stages) {
  return sizeof(float) * ((size_t)2 * kThreads + (size_t)R * D * (12 + 2 * stages) + (size_t)R * H * 8 +
                          (size_t)kRedFloats * R + 16 * ((R + 3) / 4) + 4 * ((stages * R + 3) / 4) + 16);
}

// One RHS of the adjoint ODE for the CTA's samples at solver time tau (t = -tau):
//   kz = -f, kr = -r_dot, kzb = vjp_z, (r_bar' = 0); factors + per-sample t_bar -> slot.
template <int R, int K>
__device__ __forceinline__ void adj_rhs_rows(const MlpWeights& w, float tau, const float* z, const float* zb,
                                             const float* rb, float* kz, float* kzb, float* kr, const AdjScratch& m,
                                             const FactorSlot& slot, int row0, int B) {
  const int D = w.D;
  aug_forward<R, K>(w, -tau, z, m, &slot, row0, B);
  aug_reverse<R, K>(w, zb, rb, m, slot, row0, B);
  for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
    kz[i] = -m.f[i];
    kzb[i] = m.zacc[i];
  }
  if (threadIdx.x < R) kr[threadIdx.x] = -m.rdot[threadIdx.x];
  __syncthreads();
}

template <int R, int K, int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_adj_rows(AdjArgs a, float rhs_tau) {
  Ctrl* c = a.ctrl;
  if (MODE == ADJ_STEP && c->done) return;
  extern __shared__ float4 smem4[];
  const int D = a.w.D, H = a.w.H;
  constexpr int NST = (MODE == ADJ_STEP) ? kAdjStages : 2;
  AdjSmem m = carve_adj<R>(reinterpret_cast<float*>(smem4), D, H, NST);
  const int RD = R * D;
  const int row0 = blockIdx.x * R;
  const float rtol = c->rtol, atol = c->atol;
  const int cur = (MODE == ADJ_RHS) ? 0 : c->cur;
  const float* Zc = a.Z + (size_t)cur * a.N;
  const float* ZBc = a.ZB + (size_t)cur * a.N;

  for (int i = threadIdx.x; i < RD; i += blockDim.x) {
    const bool ok = row0 + i / D < a.B;
    const size_t g = (size_t)row0 * D + i;
    m.z[i] = ok ? Zc[g] : 0.f;
    m.zb[i] = ok ? ZBc[g] : 0.f;
  }
  if (threadIdx.x < R) {
    const bool ok = row0 + threadIdx.x < a.B;
    m.r[threadIdx.x] = ok ? a.RS[(size_t)cur * a.B + row0 + threadIdx.x] : 0.f;
    m.rb[threadIdx.x] = ok ? a.RB[row0 + threadIdx.x] : 0.f;
  }
  __syncthreads();

  if (MODE == ADJ_RHS) {
    adj_rhs_rows<R, K>(a.w, rhs_tau, m.z, m.zb, m.rb, m.kz, m.kzb, m.kr, m.m, a.slot[0], row0, a.B);
    for (int i = threadIdx.x; i < RD; i += blockDim.x)
      if (row0 + i / D < a.B) {
        const size_t g = (size_t)row0 * D + i;
        a.rhs_dy[g] = -m.kz[i];
        if (a.rhs_zb) a.rhs_zb[g] = m.kzb[i];
      }
    if (threadIdx.x < R && row0 + threadIdx.x < a.B && a.rhs_r) a.rhs_r[row0 + threadIdx.x] = -m.kr[threadIdx.x];
    return;
  }

  if (MODE == ADJ_INIT0) {
    // f0 at the segment start; also t_bar = <func(ys[i], ts[i]), g[i]> (lib/ode.py:1513)
    const float tau0 = a.sc->seg_ts[0];
    adj_rhs_rows<R, K>(a.w, tau0, m.z, m.zb, m.rb, m.kz, m.kzb, m.kr, m.m, a.slot[0], row0, a.B);
    float* FZc = a.FZ + (size_t)cur * a.N;
    float* FZBc = a.FZB + (size_t)cur * a.N;
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const int r = i / D;
      const bool ok = row0 + r < a.B;
      const size_t g = (size_t)row0 * D + i;
      const float z = m.z[i], zb = m.zb[i], kz = m.kz[i], kzb = m.kzb[i];
      const float sz = atol + fabsf(z) * rtol, szb = atol + fabsf(zb) * rtol;
      const float q0a = z / sz, q0b = zb / szb, q1a = kz / sz, q1b = kzb / szb;
      m.q[i] = q0a * q0a + q0b * q0b;
      m.xz[i] = q1a * q1a + q1b * q1b;
      m.xzb[i] = ok ? (-kz) * a.g_y[g] : 0.f;  // f . g_y
      if (ok) { FZc[g] = kz; FZBc[g] = kzb; }
    }
    __syncthreads();
    // scalars r, r_bar of each sample join the per-sample sums
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      const float rr = m.r[r], rb = m.rb[r], kr = m.kr[r];
      const float sr = atol + fabsf(rr) * rtol, srb = atol + fabsf(rb) * rtol;
      const float q0 = rr / sr, q0b = rb / srb, q1 = kr / sr;
      m.xr[r] = q0 * q0 + q0b * q0b;   // d0 part
      m.m.rdot[r] = q1 * q1;          // d1 part (r_bar' = 0)
      if (row0 + r < a.B) a.FR[(size_t)cur * a.B + row0 + r] = kr;
    }
    __syncthreads();
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int r = warp; r < R; r += nw) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int d = lane; d < D; d += 32) {
          s0 += (double)m.q[r * D + d];
          s1 += (double)m.xz[r * D + d];
          s2 += (double)m.xzb[r * D + d];
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0 && row0 + r < a.B) {
          a.rowpart0[row0 + r] = s0 + (double)m.xr[r];
          a.rowpart1[row0 + r] = s1 + (double)m.m.rdot[r];
          a.rowpart2[row0 + r] = s2 + (double)(-m.kr[r]) * (double)a.g_r[row0 + r];
        }
      }
    }
    return;
  }

  if (MODE == ADJ_INIT1) {
    const float h0 = c->h0;
    const float tau0 = a.sc->seg_ts[0];
    const float* FZc = a.FZ + (size_t)cur * a.N;
    const float* FZBc = a.FZB + (size_t)cur * a.N;
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const bool ok = row0 + i / D < a.B;
      const size_t g = (size_t)row0 * D + i;
      m.kz[i] = ok ? FZc[g] : 0.f;
      m.kzb[i] = ok ? FZBc[g] : 0.f;
      m.xz[i] = fmaf(h0, m.kz[i], m.z[i]);
      m.xzb[i] = fmaf(h0, m.kzb[i], m.zb[i]);
    }
    if (threadIdx.x < R) m.kr[threadIdx.x] = (row0 + threadIdx.x < a.B) ? a.FR[(size_t)cur * a.B + row0 + threadIdx.x] : 0.f;
    __syncthreads();
    adj_rhs_rows<R, K>(a.w, tau0 + h0, m.xz, m.xzb, m.rb, m.kz + RD, m.kzb + RD, m.kr + R, m.m, a.slot[0], row0, a.B);
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const float sz = atol + fabsf(m.z[i]) * rtol, szb = atol + fabsf(m.zb[i]) * rtol;
      const float qa = (m.kz[RD + i] - m.kz[i]) / sz, qb = (m.kzb[RD + i] - m.kzb[i]) / szb;
      m.q[i] = qa * qa + qb * qb;
    }
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      const float sr = atol + fabsf(m.r[r]) * rtol;
      const float q = (m.kr[R + r] - m.kr[r]) / sr;
      m.xr[r] = q * q;
    }
    __syncthreads();
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int r = warp; r < R; r += nw) {
        double s0 = 0.0;
        for (int d = lane; d < D; d += 32) s0 += (double)m.q[r * D + d];
        s0 = warp_sum(s0);
        if (lane == 0 && row0 + r < a.B) a.rowpart0[row0 + r] = s0 + (double)m.xr[r];
      }
    }
    return;
  }

  // ---- ADJ_STEP: all stages of one dopri5 iteration -------------------------------------
  const float tau = c->t;
  const float dt = c->dt;   // dopri5 flavour: never clipped (lib/ode.py:833-839)
  const int cand = c->cand, midw = c->mid_acc ^ 1;
  {
    const float* FZc = a.FZ + (size_t)cur * a.N;
    const float* FZBc = a.FZB + (size_t)cur * a.N;
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      const bool ok = row0 + i / D < a.B;
      const size_t g = (size_t)row0 * D + i;
      m.kz[i] = ok ? FZc[g] : 0.f;
      m.kzb[i] = ok ? FZBc[g] : 0.f;
    }
    if (threadIdx.x < R) m.kr[threadIdx.x] = (row0 + threadIdx.x < a.B) ? a.FR[(size_t)cur * a.B + row0 + threadIdx.x] : 0.f;
  }
  __syncthreads();
  for (int st = 1; st < kAdjStages; ++st) {
    for (int i = threadIdx.x; i < RD; i += blockDim.x) {
      float az = 0.f, ab = 0.f;
      for (int j = 0; j < st; ++j) {
        const float bj = c_tab.beta[st - 1][j];
        az = fmaf(bj, m.kz[j * RD + i], az);
        ab = fmaf(bj, m.kzb[j * RD + i], ab);
      }
      m.xz[i] = fmaf(dt, az, m.z[i]);
      m.xzb[i] = fmaf(dt, ab, m.zb[i]);
    }
    // (r does not enter the RHS; its stage value is not needed.)
    __syncthreads();
    adj_rhs_rows<R, K>(a.w, tau + dt * c_tab.alpha[st - 1], m.xz, m.xzb, m.rb, m.kz + (size_t)st * RD,
                       m.kzb + (size_t)st * RD, m.kr + st * R, m.m, a.slot[st - 1], row0, a.B);
  }
  float* Zn = a.Z + (size_t)cand * a.N;
  float* ZBn = a.ZB + (size_t)cand * a.N;
  float* FZn = a.FZ + (size_t)cand * a.N;
  float* FZBn = a.FZB + (size_t)cand * a.N;
  float* Mn = a.MIDZB + (size_t)midw * a.N;
  for (int i = threadIdx.x; i < RD; i += blockDim.x) {
    float sz = 0.f, ez = 0.f, sb = 0.f, eb = 0.f, mb = 0.f;
    for (int j = 0; j < kAdjStages; ++j) {
      const float kz = m.kz[j * RD + i], kb = m.kzb[j * RD + i];
      sz = fmaf(c_tab.c_sol[j], kz, sz);
      ez = fmaf(c_tab.c_err[j], kz, ez);
      sb = fmaf(c_tab.c_sol[j], kb, sb);
      eb = fmaf(c_tab.c_err[j], kb, eb);
      mb = fmaf(c_tab.c_mid[j], kb, mb);
    }
    const float z = m.z[i], zb = m.zb[i];
    const float z1 = fmaf(dt, sz, z), zb1 = fmaf(dt, sb, zb);
    const float tz = atol + rtol * fmaxf(fabsf(z), fabsf(z1));
    const float tb = atol + rtol * fmaxf(fabsf(zb), fabsf(zb1));
    const float qz = dt * ez / tz, qb = dt * eb / tb;
    m.q[i] = qz * qz + qb * qb;
    if (row0 + i / D < a.B) {
      const size_t g = (size_t)row0 * D + i;
      Zn[g] = z1;
      ZBn[g] = zb1;
      FZn[g] = m.kz[(size_t)(kAdjStages - 1) * RD + i];
      FZBn[g] = m.kzb[(size_t)(kAdjStages - 1) * RD + i];
      Mn[g] = fmaf(dt, mb, zb);
    }
  }
  if (threadIdx.x < R) {
    const int r = threadIdx.x;
    float sr = 0.f, er = 0.f;
    for (int j = 0; j < kAdjStages; ++j) {
      sr = fmaf(c_tab.c_sol[j], m.kr[j * R + r], sr);
      er = fmaf(c_tab.c_err[j], m.kr[j * R + r], er);
    }
    const float r0 = m.r[r];
    const float r1 = fmaf(dt, sr, r0);
    const float tr = atol + rtol * fmaxf(fabsf(r0), fabsf(r1));
    const float q = dt * er / tr;
    m.xr[r] = q * q;   // r_bar: derivative 0, error 0
    if (row0 + r < a.B) {
      a.RS[(size_t)cand * a.B + row0 + r] = r1;
      a.FR[(size_t)cand * a.B + row0 + r] = m.kr[(kAdjStages - 1) * R + r];
    }
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
      double s0 = 0.0;
      for (int d = lane; d < D; d += 32) s0 += (double)m.q[r * D + d];
      s0 = warp_sum(s0);
      if (lane == 0 && row0 + r < a.B) a.rowpart0[row0 + r] = s0 + (double)m.xr[r];
    }
  }
}

// Norm over ALL blocks of the adjoint state, the t0_bar element, and the controller
// (lib/ode.py:86-104 for the init modes, 836-847 for an iteration).  Runs in one CTA; `n_pp` is the
// number of parameter-block partial sums in ppart0/ppart1.  Inputs are the per-sample partials
// (rowpart*, slot[].tbar: n_rp entries each, already gathered from every rank in the multi-GPU case).
template <int MODE>
__device__ __forceinline__ void adj_finalize(AdjArgs& a, Ctrl* c, double* dred, int n_pp, float dt,
                                             float* rhs_tbar, const double* gx = nullptr) {
  // gx: every rank's record in one gathered array somewhere else than a.xrec (peer-memory exchange)
  const double* rp0_all = gx ? gx : a.rp0_all;
  const double* rp1_all = gx ? gx + 4 * a.n_loc : a.rp1_all;
  const double* rp2_all = gx ? gx + 5 * a.n_loc : a.rp2_all;
  constexpr bool SPLITK = (MODE != ADJ_STEP);
  constexpr int NS = kSlots;
  const int B = a.B;
  (void)B;
  const float rtol = c->rtol, atol = c->atol;
  const float h0 = c->h0;
  __shared__ double fin[12];
  if (MODE == ADJ_STEP) {
    // the eight sums of an iteration, one warp each (the CTA has at least eight warps)
    const int w = threadIdx.x >> 5;
    double v = 0.0;
    if (w == 0) v = warp_sum_records(rp0_all, a.n_rp, a.n_loc, 6 * a.n_loc);
    else if (w == 1) v = warp_sum_records(a.ppart0, n_pp, n_pp, 0);
    else if (w < 2 + NS) {
      const int j = w - 2;
      const float* tb = gx ? reinterpret_cast<const float*>(gx) + 2 * a.n_loc + (size_t)j * a.n_loc : a.tbar_all[j];
      v = warp_sum_records(tb, a.n_rp, a.n_loc, 12 * a.n_loc);
    }
    if ((threadIdx.x & 31) == 0 && w < 2 + NS) fin[w < 2 ? w : 3 + w] = v;   // tbar_j -> fin[5 + j]
    __syncthreads();
  } else {
    double v;
    v = block_sum_records(rp0_all, a.n_rp, a.n_loc, 6 * a.n_loc, dred); if (threadIdx.x == 0) fin[0] = v; __syncthreads();
    v = block_sum_global(a.ppart0, n_pp, dred); if (threadIdx.x == 0) fin[1] = v; __syncthreads();
    if (MODE == ADJ_INIT0) {
      v = block_sum_records(rp1_all, a.n_rp, a.n_loc, 6 * a.n_loc, dred); if (threadIdx.x == 0) fin[2] = v; __syncthreads();
      v = block_sum_global(a.ppart1, n_pp, dred); if (threadIdx.x == 0) fin[3] = v; __syncthreads();
      v = block_sum_records(rp2_all, a.n_rp, a.n_loc, 6 * a.n_loc, dred); if (threadIdx.x == 0) fin[4] = v; __syncthreads();
    }
    // stage derivatives of t0_bar: sums of the per-sample t_bar contributions
    for (int j = 0; j < (SPLITK ? 1 : NS); ++j) {
      const float* tb = gx ? reinterpret_cast<const float*>(gx) + 2 * a.n_loc + (size_t)j * a.n_loc : a.tbar_all[j];
      const int per = (a.n_rp + blockDim.x - 1) / blockDim.x;
      const int b0 = threadIdx.x * per, b1 = min(a.n_rp, b0 + per);
      double acc = 0.0;
      for (int i = b0; i < b1; ++i) {
        const int r = i / a.n_loc;
        acc += (double)tb[(size_t)r * 12 * a.n_loc + (i - r * a.n_loc)];
      }
      dred[threadIdx.x] = acc;
      __syncthreads();
      for (int o = 256; o > 0; o >>= 1) {
        if (threadIdx.x < o && threadIdx.x + o < blockDim.x) dred[threadIdx.x] += dred[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) fin[5 + j] = dred[0];
      __syncthreads();
    }
  }
  if (threadIdx.x != 0) return;
  c->ticket = 0;
  AdjScalars* sc = a.sc;
  if (MODE == ADJ_RHS) {
    if (rhs_tbar) *rhs_tbar = (float)fin[5];
    return;
  }
  if (MODE == ADJ_INIT0) {
    // t_bar of this output time, then the state's t0_bar element (lib/ode.py:1513-1514)
    const float t_bar = (float)fin[4];
    *a.tbar_out = t_bar;
    const float t0b = sc->t0[0] - t_bar;
    sc->t0[0] = t0b;
    const float f0 = (float)fin[5];
    sc->ft0[0] = f0;
    const float s = atol + fabsf(t0b) * rtol;
    const double q0 = (double)(t0b / s) * (double)(t0b / s), q1 = (double)(f0 / s) * (double)(f0 / s);
    const float d0 = sqrtf((float)(fin[0] + fin[1] + q0));
    const float d1 = sqrtf((float)(fin[2] + fin[3] + q1));
    c->d0 = d0; c->d1 = d1;
    c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
    return;
  }
  if (MODE == ADJ_INIT1) {
    const float s = atol + fabsf(sc->t0[0]) * rtol;
    const float qt = ((float)fin[5] - sc->ft0[0]) / s;
    const float d1 = c->d1;
    const float d2 = sqrtf((float)(fin[0] + fin[1] + (double)(qt * qt))) / h0;
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
    else h1 = powf(0.01f / (d1 + d2), 1.0f / ((float)c_tab.init_order + 1.0f));
    c->dt = fminf(100.f * h0, h1);
    if (isnan(h0) || isnan(h1)) c->dt = nanf("");
    c->t = sc->seg_ts[0];
    c->last_t = c->t;
    c->step_dt = 0.f;
    c->cur = 0; c->prev = 2; c->cand = 1; c->mid_acc = 0;
    c->iters_tgt = 0; c->tgt = 1;
    c->n_iters = 0; c->n_accepted = 0; c->nfe = 2;
    c->out_begin = 1; c->out_end = 1;
    c->done = 0; c->status = ENO_OK;
    advance_targets(c, sc->seg_ts);
    return;
  }
  // ADJ_STEP
  {
    const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
    const float k0 = sc->ft0[cur];
    float s = c_tab.c_sol[0] * k0, er = c_tab.c_err[0] * k0, md = c_tab.c_mid[0] * k0, k6 = 0.f;
    for (int j = 0; j < kSlots; ++j) {
      const float kj = (float)fin[5 + j];
      s = fmaf(c_tab.c_sol[j + 1], kj, s);
      er = fmaf(c_tab.c_err[j + 1], kj, er);
      md = fmaf(c_tab.c_mid[j + 1], kj, md);
      k6 = kj;
    }
    const float p0 = sc->t0[cur];
    const float p1 = fmaf(dt, s, p0);
    const float tol = atol + rtol * fmaxf(fabsf(p0), fabsf(p1));
    const float q = dt * er / tol;
    sc->t0[cand] = p1;
    sc->ft0[cand] = k6;
    sc->midt0[midw] = fmaf(dt, md, p0);
    const double total = fin[0] + fin[1] + (double)(q * q);
    const float ratio = (float)(total / (double)a.n_state);
    finish_iteration(c, sc->seg_ts, ratio, dt, c_tab, a.trace);
  }
}

// --------------------------------------------------------------------------------------
// Parameter-gradient contraction.  Block kinds by blockIdx.x:
//   [0, nt1)            : 32x32 tiles of dW1x  (M = D rows, N = H cols),  A = S,  Bm = A-factors
//   [nt1, nt1+nt2)      : 32x32 tiles of dW2x  (M = H, N = D),            A = Hh, Bm = G
//   [nt1+nt2, +nv)      : 64-wide pieces of the four vectors b1, w1t, b2, w2t (column sums)
// Thread group g (64 threads) handles RHS slot g (stage g+1); NS = 1 outside ADJ_STEP.
// --------------------------------------------------------------------------------------
struct PgradGeom {
  int tm1, tn1, nt1, tm2, tn2, nt2, nv1, nv2, nv, grid;
};

__host__ __device__ inline PgradGeom pgrad_geom(int D, int H) {
  PgradGeom g;
  g.tm1 = (D + kTile - 1) / kTile; g.tn1 = (H + kTile - 1) / kTile; g.nt1 = g.tm1 * g.tn1;
  g.tm2 = (H + kTile - 1) / kTile; g.tn2 = (D + kTile - 1) / kTile; g.nt2 = g.tm2 * g.tn2;
  g.nv1 = (2 * H + kGroup - 1) / kGroup;
  g.nv2 = (2 * D + kGroup - 1) / kGroup;
  g.nv = g.nv1 + g.nv2;
  g.grid = g.nt1 + g.nt2 + g.nv;
  return g;
}

template <int MODE>
__global__ void __launch_bounds__(kSlots * kGroup) k_adj_pgrad(AdjArgs a, float rhs_tau, float* rhs_pbar,
                                                               float* rhs_tbar) {
  Ctrl* c = a.ctrl;
  if (MODE == ADJ_STEP && c->done) return;
  // ADJ_STEP: thread group g contracts RHS slot g (stage g+1) over all rows.  Other modes have a single
  // slot; the six groups then split its rows (K) between them and the tiles are summed.
  constexpr bool SPLITK = (MODE != ADJ_STEP);
  constexpr int NS = kSlots;
  __shared__ float tile[NS][kTile * kTile];
  __shared__ int pidx[kTile * kTile];   // flat parameter index of each tile element (-1 = padding)
  __shared__ double dred[kSlots * kGroup];
  extern __shared__ __align__(16) float ring[];   // [kSlots][kPStages][kPRows][64] operand staging (dynamic)
  const int D = a.w.D, H = a.w.H, B = a.B;
  const PgradGeom geo = pgrad_geom(D, H);
  const int grp = threadIdx.x / kGroup;
  const int tig = threadIdx.x - grp * kGroup;
  const int KB = a.np * B;   // contraction length per slot
  const int b = blockIdx.x;
  const float tau = (MODE == ADJ_STEP) ? c->t : (MODE == ADJ_RHS ? rhs_tau : a.sc->seg_ts[0]);
  const float dt = (MODE == ADJ_STEP) ? c->dt : 0.f;
  const float h0 = c->h0;
  const int offW1 = 2 * H, offb2 = H * (D + 2), offW2 = H * (D + 2) + 2 * D;

  int n_out;  // outputs in this block's tile
  if (b < geo.nt1 + geo.nt2) {
    const bool first = b < geo.nt1;
    const int tb = first ? b : b - geo.nt1;
    const int tn = first ? geo.tn1 : geo.tn2;
    const int M = first ? D : H, N = first ? H : D;
    const int m0 = (tb / tn) * kTile, n0 = (tb % tn) * kTile;
    const int tx = tig & 7, ty = tig >> 3;          // 8 x 8 threads
    const int mm = m0 + ty * 4, nn = n0 + tx * 4;   // 4 x 4 outputs per thread
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // Operand rows are staged through a per-group ring in shared memory with cp.async (LDGSTS): kPStages
    // stages of 4 contraction rows are in flight while the FMAs consume the oldest one, which takes the
    // L2 latency off the critical path without spending registers on prefetch buffers.
    {
      const bool active = grp < NS;
      const FactorSlot& sl = a.slot[(SPLITK || !active) ? 0 : grp];
      const float* Am = first ? sl.S : sl.Hh;   // [KB][M]   (rows live as [3][B][*]; the first np*B are valid)
      const float* Bm = first ? sl.A : sl.G;    // [KB][N]
      const int kchunk = (KB + kSlots - 1) / kSlots;
      const int kbeg = SPLITK ? grp * kchunk : 0;
      const int kend = SPLITK ? min(KB, kbeg + kchunk) : KB;
      float* rg = ring + (size_t)grp * kPStages * kPRows * 64;     // [kPStages][kPRows][A 32 | B 32]
      // this thread's 16-byte piece of a stage: operand (A/B), row 0..3, float4 column 0..7
      const int which = tig >> 5, lrow = (tig & 31) >> 3, lc4 = tig & 7;
      const float* src_base = which ? Bm : Am;
      const int ld = which ? N : M;
      const int col = (which ? n0 : m0) + 4 * lc4;
      const bool col_ok = col < ld;
      const int nst = active ? (kend - kbeg + kPRows - 1) / kPRows : 0;
      auto issue = [&](int st) {
#pragma unroll
        for (int h = 0; h < kPRows / 4; ++h) {
          const int r = lrow + 4 * h;
          const int k = kbeg + kPRows * st + r;
          float* dst = rg + ((size_t)(st % kPStages) * kPRows + r) * 64 + which * 32 + 4 * lc4;
          const bool ok = col_ok && k < kend;
          const float* src = ok ? src_base + (size_t)k * ld + col : src_base;
          __pipeline_memcpy_async(dst, src, 16, ok ? 0 : 16);    // out-of-range pieces are zero-filled
        }
      };
      for (int st = 0; st < kPStages - 1; ++st) {
        if (st < nst) issue(st);
        __pipeline_commit();
      }
      for (int st = 0; st < nst; ++st) {
        __pipeline_wait_prior(kPStages - 2);
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kGroup) : "memory");   // stage st visible to the group
        const float* sb = rg + (size_t)(st % kPStages) * kPRows * 64;
#pragma unroll
        for (int u = 0; u < kPRows; ++u) {
          const float4 a4 = *reinterpret_cast<const float4*>(sb + u * 64 + ty * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(sb + u * 64 + 32 + tx * 4);
          const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
          const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kGroup) : "memory");   // everyone done with the slot being refilled
        if (st + kPStages - 1 < nst) issue(st + kPStages - 1);
        __pipeline_commit();
      }
      __pipeline_wait_prior(0);
    }
    if (grp < NS) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[grp][(ty * 4 + i) * kTile + tx * 4 + j] = acc[i][j];
    }
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mi = mm + i, nj = nn + j;
          pidx[(ty * 4 + i) * kTile + tx * 4 + j] =
              (mi < M && nj < N) ? (first ? offW1 + mi * H + nj : offW2 + mi * D + nj) : -1;
        }
    }
    n_out = kTile * kTile;
  } else {
    // column sums: b1 = sum a_bar, w1t = sum (t*a_bar + a1_bar), b2 = sum f_bar, w2t = sum (t*f_bar + y1_bar)
    const int vb = b - geo.nt1 - geo.nt2;
    const bool first = vb < geo.nv1;
    const int W = first ? H : D;
    const int o = (first ? vb : vb - geo.nv1) * kGroup + tig;   // in [0, 2W)
    float acc = 0.f;
    if (grp < NS && o < 2 * W) {
      const FactorSlot& sl = a.slot[SPLITK ? 0 : grp];
      const float* F0 = first ? sl.A : sl.G;             // pass-0 factor [B][W]
      const float* F1 = F0 + (size_t)B * W;              // pass-1 factor
      const float t_stage = -(tau + ((MODE == ADJ_STEP) ? dt * c_tab.alpha[grp] : (MODE == ADJ_INIT1 ? h0 : 0.f)));
      const int rchunk = (B + kSlots - 1) / kSlots;
      const int rb = SPLITK ? grp * rchunk : 0, re = SPLITK ? min(B, rb + rchunk) : B;
      if (o < W) {
        for (int r = rb; r < re; ++r) acc += F0[(size_t)r * W + o];
      } else {
        const int n = o - W;
        // the jvp of the fin dynamics is taken wrt y only: its tangent pass has no w1t / w2t terms
        if (a.np >= 2 && !a.fin) for (int r = rb; r < re; ++r) acc += fmaf(t_stage, F0[(size_t)r * W + n], F1[(size_t)r * W + n]);
        else for (int r = rb; r < re; ++r) acc += t_stage * F0[(size_t)r * W + n];
      }
    }
    if (grp < NS) tile[grp][tig] = acc;
    if (grp == 0) pidx[tig] = (o < 2 * W) ? ((first ? 0 : offb2) + o) : -1;
    n_out = kGroup;
  }
  __syncthreads();

  if (a.dist) {
    // multi-GPU: this rank only holds the sums over ITS rows.  Everything downstream is linear in the
    // stage derivatives, so ship the four stage combinations (without the replicated k0 term) and let
    // k_adj_pfinish continue after the all-reduce.
    const size_t P = (size_t)a.P;
    float* pcomb = a.pcomb;
    if (a.peer_base) {   // the slot of the exchange that the finish kernel after this launch will run
      const unsigned long long seq = *reinterpret_cast<const unsigned long long*>(a.peer_base + kPeerOffSeq) + 1;
      pcomb = reinterpret_cast<float*>(a.peer_base + a.peer_pcomb_off) + (size_t)(seq & 1) * 4 * P;
    }
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p < 0) continue;
      if (MODE == ADJ_STEP) {
        float s = 0.f, er = 0.f, md = 0.f;
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          const float kj = tile[j][e];
          s = fmaf(c_tab.c_sol[j + 1], kj, s);
          er = fmaf(c_tab.c_err[j + 1], kj, er);
          md = fmaf(c_tab.c_mid[j + 1], kj, md);
        }
        pcomb[p] = s;
        pcomb[P + p] = er;
        pcomb[2 * P + p] = md;
        pcomb[3 * P + p] = tile[kSlots - 1][e];
      } else {
        pcomb[p] = ((((tile[0][e] + tile[1][e]) + tile[2][e]) + tile[3][e]) + tile[4][e]) + tile[5][e];
      }
    }
    return;
  }

  // ---- combine the stage derivatives of this tile ----------------------------------------
  const float rtol = c->rtol, atol = c->atol;
  double part0 = 0.0, part1 = 0.0;
  if (MODE == ADJ_STEP) {
    const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
    const float* PBc = a.PB + (size_t)cur * a.P;
    const float* FPc = a.FP + (size_t)cur * a.P;
    float* PBn = a.PB + (size_t)cand * a.P;
    float* FPn = a.FP + (size_t)cand * a.P;
    float* Mn = a.MIDP + (size_t)midw * a.P;
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p < 0) continue;
      const float k0 = FPc[p];
      float s = c_tab.c_sol[0] * k0, er = c_tab.c_err[0] * k0, md = c_tab.c_mid[0] * k0;
      float k6 = 0.f;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        const float kj = tile[j][e];
        s = fmaf(c_tab.c_sol[j + 1], kj, s);
        er = fmaf(c_tab.c_err[j + 1], kj, er);
        md = fmaf(c_tab.c_mid[j + 1], kj, md);
        k6 = kj;
      }
      const float p0 = PBc[p];
      const float p1 = fmaf(dt, s, p0);
      const float tol = atol + rtol * fmaxf(fabsf(p0), fabsf(p1));
      const float q = dt * er / tol;
      part0 += (double)(q * q);
      PBn[p] = p1;
      FPn[p] = k6;
      Mn[p] = fmaf(dt, md, p0);
    }
  } else if (MODE == ADJ_INIT0) {
    const float* PBc = a.PB;   // ring slot 0
    float* FPc = a.FP;
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p < 0) continue;
      const float f0 = ((((tile[0][e] + tile[1][e]) + tile[2][e]) + tile[3][e]) + tile[4][e]) + tile[5][e];
      const float p0 = PBc[p];
      const float sc = atol + fabsf(p0) * rtol;
      const float q0 = p0 / sc, q1 = f0 / sc;
      part0 += (double)(q0 * q0);
      part1 += (double)(q1 * q1);
      FPc[p] = f0;
    }
  } else if (MODE == ADJ_INIT1) {
    const float* PBc = a.PB;
    const float* FPc = a.FP;
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p < 0) continue;
      const float sc = atol + fabsf(PBc[p]) * rtol;
      const float f1 = ((((tile[0][e] + tile[1][e]) + tile[2][e]) + tile[3][e]) + tile[4][e]) + tile[5][e];
      const float q = (f1 - FPc[p]) / sc;
      part0 += (double)(q * q);
    }
  } else {  // ADJ_RHS
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p >= 0) rhs_pbar[p] = ((((tile[0][e] + tile[1][e]) + tile[2][e]) + tile[3][e]) + tile[4][e]) + tile[5][e];
    }
  }
  // deterministic block reduction of the partial sums
  dred[threadIdx.x] = part0;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o && threadIdx.x + o < blockDim.x) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.ppart0[b] = dred[0];
  __syncthreads();
  if (MODE == ADJ_INIT0) {
    dred[threadIdx.x] = part1;
    __syncthreads();
    for (int o = 256; o > 0; o >>= 1) {
      if (threadIdx.x < o && threadIdx.x + o < blockDim.x) dred[threadIdx.x] += dred[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) a.ppart1[b] = dred[0];
    __syncthreads();
  }

  // ---- last CTA: norm over all blocks, t0_bar element, controller -----------------------
  // the next per-sample kernel may run its weight prologue under the finaliser (cl_kernels.cuh: pdl_wait)
  if (MODE == ADJ_STEP || MODE == ADJ_INIT1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!elect_last_block(&c->ticket)) return;
  adj_finalize<MODE>(a, c, dred, geo.grid, dt, rhs_tbar);
}

// --------------------------------------------------------------------------------------
// Tensor-core formulation of the ADJ_STEP parameter-gradient contraction (cluster path).  k_cl_adj_rows leaves
// the factors of its six stages as K-major tf32 hi / lo planes (FactorTc); ONE batched launch of the tcgen05
// 3xTF32 GEMM (gemm_tc.cuh, grid.z = stage x {dW1x, dW2x^T}, K = np * B) writes the twelve products into
// KP [kSlots][P] through EpiPgradStage; k_adj_pcombine adds the bias / time-column sums, combines the stages
// into candidate / error / mid-point of params_bar and runs the finaliser - the second half of k_adj_pgrad.
// --------------------------------------------------------------------------------------
struct EpiPgradStage {
  float* KP;
  int D, H, P, offW1, offW2;
  // tile row m = state column d, tile columns n0.. = hidden columns; z = 2 * stage + which
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[tc::kEpiW], int z) const {
    if (m >= D || n0 >= H) return;
    float* kp = KP + (size_t)(z >> 1) * P;
    if ((z & 1) == 0) {            // dW1x[d][h]
      float* row = kp + offW1 + (size_t)m * H + n0;
      if (n0 + tc::kEpiW <= H && ((offW1 + (size_t)m * H + n0) & 3) == 0) {
        *reinterpret_cast<float4*>(row) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < tc::kEpiW; ++j)
          if (n0 + j < H) row[j] = v[j];
      }
    } else {                        // dW2x[h][d] (the GEMM produced its transpose)
#pragma unroll
      for (int j = 0; j < tc::kEpiW; ++j)
        if (n0 + j < H) kp[offW2 + (size_t)(n0 + j) * D + m] = v[j];
    }
  }
};

constexpr int kPcGrid = 256;   // fixed grid: the order of the norm partials does not depend on the device

__global__ void __launch_bounds__(256, 2) k_adj_pcombine(AdjArgs a) {
  Ctrl* c = a.ctrl;
  if (c->done) return;
  __shared__ double dred[256];
  const int D = a.w.D, H = a.w.H;
  const size_t P = (size_t)a.P;
  const float tau = c->t, dt = c->dt;
  const float rtol = c->rtol, atol = c->atol;
  const int offW1 = 2 * H, offb2 = H * (D + 2), offW2 = H * (D + 2) + 2 * D;
  const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
  const float* PBc = a.PB + (size_t)cur * P;
  const float* FPc = a.FP + (size_t)cur * P;
  float* PBn = a.PB + (size_t)cand * P;
  float* FPn = a.FP + (size_t)cand * P;
  float* Mn = a.MIDP + (size_t)midw * P;
  float* pcomb = a.pcomb;
  if (a.dist && a.peer_base) {   // the slot of the exchange that the finish kernel after this launch will run
    const unsigned long long seq = *reinterpret_cast<const unsigned long long*>(a.peer_base + kPeerOffSeq) + 1;
    pcomb = reinterpret_cast<float*>(a.peer_base + a.peer_pcomb_off) + (size_t)(seq & 1) * 4 * P;
  }
  double part0 = 0.0;
  auto combine = [&](int p, const float (&kj)[kSlots]) {
    if (a.dist) {   // this rank's share of the four stage combinations (k_adj_pfinish* continue after the exchange)
      float s = 0.f, er = 0.f, md = 0.f;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        s = fmaf(c_tab.c_sol[j + 1], kj[j], s);
        er = fmaf(c_tab.c_err[j + 1], kj[j], er);
        md = fmaf(c_tab.c_mid[j + 1], kj[j], md);
      }
      pcomb[p] = s;
      pcomb[P + p] = er;
      pcomb[2 * P + p] = md;
      pcomb[3 * P + p] = kj[kSlots - 1];
      return;
    }
    const float k0 = FPc[p];
    float s = c_tab.c_sol[0] * k0, er = c_tab.c_err[0] * k0, md = c_tab.c_mid[0] * k0;
#pragma unroll
    for (int j = 0; j < kSlots; ++j) {
      s = fmaf(c_tab.c_sol[j + 1], kj[j], s);
      er = fmaf(c_tab.c_err[j + 1], kj[j], er);
      md = fmaf(c_tab.c_mid[j + 1], kj[j], md);
    }
    const float p0 = PBc[p];
    const float p1 = fmaf(dt, s, p0);
    const float tol = atol + rtol * fmaxf(fabsf(p0), fabsf(p1));
    const float q = dt * er / tol;
    part0 += (double)(q * q);
    PBn[p] = p1;
    FPn[p] = kj[kSlots - 1];
    Mn[p] = fmaf(dt, md, p0);
  };
  // bias / time-column blocks: b1 = sum a_bar, w1t = sum (t a_bar + a1_bar), b2 = sum f_bar, w2t = sum (t f_bar + y1_bar)
  // - row sums of the transposed factor planes (hi + lo restores the value exactly); one warp per output
  {
    const FactorTc& t = a.tc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = 2 * H + 2 * D;
    const bool two = a.np >= 2 && !a.fin;   // the jvp of the fin dynamics has no time tangent
    const int Bk = t.Bk;                    // samples per pass along a row, padding (zeros) included
    for (int e = blockIdx.x * 8 + warp; e < nvec; e += gridDim.x * 8) {
      const bool first = e < 2 * H;
      const int W = first ? H : D;
      const int o = first ? e : e - 2 * H;
      const int col = o < W ? o : o - W;
      const unsigned plane = first ? t.yplane : t.xplane;
      const float* row0p = first ? t.Y + (size_t)col * t.ldk : t.X + ((size_t)t.dpad + col) * t.ldk;
      const size_t sstride = (size_t)2 * (first ? t.hpad : t.dpad) * t.ldk;   // stage to stage
      const bool second_pass = (o >= W) && two;
      float kj[kSlots];
#pragma unroll
      for (int j = 0; j < kSlots; ++j) kj[j] = 0.f;
      // a lane owns 4 consecutive samples; the loads of all six stages are issued before the first use
      for (int k0 = 4 * lane; k0 < Bk; k0 += 128) {
        float4 h0[kSlots], l0[kSlots], h1[kSlots], l1[kSlots];
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          const float* hp = row0p + (size_t)j * sstride + k0;
          h0[j] = *reinterpret_cast<const float4*>(hp);
          l0[j] = *reinterpret_cast<const float4*>(hp + plane);
          if (second_pass) {
            h1[j] = *reinterpret_cast<const float4*>(hp + Bk);
            l1[j] = *reinterpret_cast<const float4*>(hp + Bk + plane);
          }
        }
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          const float f0[4] = {h0[j].x + l0[j].x, h0[j].y + l0[j].y, h0[j].z + l0[j].z, h0[j].w + l0[j].w};
          if (o < W) {
            kj[j] += (f0[0] + f0[1]) + (f0[2] + f0[3]);
          } else {
            const float t_stage = -(tau + dt * c_tab.alpha[j]);
            if (second_pass) {
              const float f1[4] = {h1[j].x + l1[j].x, h1[j].y + l1[j].y, h1[j].z + l1[j].z, h1[j].w + l1[j].w};
              kj[j] += (fmaf(t_stage, f0[0], f1[0]) + fmaf(t_stage, f0[1], f1[1])) +
                       (fmaf(t_stage, f0[2], f1[2]) + fmaf(t_stage, f0[3], f1[3]));
            } else {
              kj[j] += t_stage * ((f0[0] + f0[1]) + (f0[2] + f0[3]));
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kSlots; ++j)
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) kj[j] += __shfl_xor_sync(0xffffffffu, kj[j], sft);
      if (lane == 0) combine((first ? 0 : offb2) + o, kj);
    }
  }
  // weight blocks: the GEMM's stage products
  {
    const int DH = D * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * DH; i += gridDim.x * blockDim.x) {
      const int p = i < DH ? offW1 + i : offW2 + (i - DH);
      float kj[kSlots];
#pragma unroll
      for (int j = 0; j < kSlots; ++j) kj[j] = a.KP[(size_t)j * P + p];
      combine(p, kj);
    }
  }
  if (a.dist) return;
  dred[threadIdx.x] = part0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.ppart0[blockIdx.x] = dred[0];
  __syncthreads();
  // the next per-sample kernel may run its weight prologue under the finaliser (cl_kernels.cuh: pdl_wait)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!elect_last_block(&c->ticket)) return;
  adj_finalize<ADJ_STEP>(a, c, dred, (int)gridDim.x, dt, nullptr);
}

// Multi-GPU continuation of k_adj_pgrad: pcomb now holds the all-reduced stage combinations of the
// params_bar block.  One thread per parameter forms candidate / error / mid-point (STEP) or the
// initial-step partial norms (INIT*), the last CTA runs the same finaliser as the single-GPU kernel.
// Every rank executes it on identical inputs.
template <int MODE>
__global__ void __launch_bounds__(256) k_adj_pfinish(AdjArgs a) {
  Ctrl* c = a.ctrl;
  if (MODE == ADJ_STEP && c->done) return;
  // every CTA of this grid is resident by the time all have passed this point: the next per-sample kernel
  // (programmatic serialization) may load its weights while the exchange below runs
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ double dred[256];
  const size_t P = (size_t)a.P;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const float rtol = c->rtol, atol = c->atol;
  const float dt = (MODE == ADJ_STEP) ? c->dt : 0.f;
  double part0 = 0.0, part1 = 0.0;
  if (p < a.P) {
    if (MODE == ADJ_STEP) {
      const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
      const float k0 = a.FP[(size_t)cur * P + p];
      const float s = fmaf(c_tab.c_sol[0], k0, a.pcomb[p]);
      const float er = fmaf(c_tab.c_err[0], k0, a.pcomb[P + p]);
      const float md = fmaf(c_tab.c_mid[0], k0, a.pcomb[2 * P + p]);
      const float p0 = a.PB[(size_t)cur * P + p];
      const float p1 = fmaf(dt, s, p0);
      const float tol = atol + rtol * fmaxf(fabsf(p0), fabsf(p1));
      const float q = dt * er / tol;
      part0 = (double)(q * q);
      a.PB[(size_t)cand * P + p] = p1;
      a.FP[(size_t)cand * P + p] = a.pcomb[3 * P + p];
      a.MIDP[(size_t)midw * P + p] = fmaf(dt, md, p0);
    } else if (MODE == ADJ_INIT0) {
      const float f0 = a.pcomb[p];
      const float p0 = a.PB[p];
      const float sc = atol + fabsf(p0) * rtol;
      const float q0 = p0 / sc, q1 = f0 / sc;
      part0 = (double)(q0 * q0);
      part1 = (double)(q1 * q1);
      a.FP[p] = f0;
    } else if (MODE == ADJ_INIT1) {
      const float sc = atol + fabsf(a.PB[p]) * rtol;
      const float q = (a.pcomb[p] - a.FP[p]) / sc;
      part0 = (double)(q * q);
    }
  }
  dred[threadIdx.x] = part0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.ppart0[blockIdx.x] = dred[0];
  __syncthreads();
  if (MODE == ADJ_INIT0) {
    dred[threadIdx.x] = part1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) a.ppart1[blockIdx.x] = dred[0];
    __syncthreads();
  }
  if (!elect_last_block(&c->ticket)) return;
  adj_finalize<MODE>(a, c, dred, (int)gridDim.x, dt, nullptr);
}

// Peer-memory variant of k_adj_pfinish (dist.cuh): no NCCL call.  Round 1: CTA 0 pushes this rank's record (norm
// partials + t_bar pieces, 6 n_loc doubles) into every rank's gathered array and signals; every CTA waits for all ranks,
// then PULLS, for the params_bar slice this rank owns, the stage combinations of every rank in rank order and forms
// candidate / error / mid-point and the slice's norm partial.  Round 2 (last CTA): the per-rank norm partials are
// exchanged, summed in rank order, and the usual finaliser runs on identical inputs on every rank.  The params_bar
// rings therefore hold valid data for the owned slice only; the host gathers POUT once per segment.
#ifdef ENO_PHASES
// development aid: clock stamps of the peer exchange (CTA 0: entry, pushed, all ranks arrived, slice done; last CTA:
// elected, round 2 done, finalised), most recent ADJ_STEP launch
__device__ long long g_peer_stamp[8];
#define PEER_STAMP(cond, slot) do { if ((cond) && threadIdx.x == 0 && MODE == ADJ_STEP) g_peer_stamp[slot] = clock64(); } while (0)
#else
#define PEER_STAMP(cond, slot) do { } while (0)
#endif

template <int MODE>
__global__ void __launch_bounds__(256) k_adj_pfinish_peer(AdjArgs a) {
  Ctrl* c = a.ctrl;
  if (MODE == ADJ_STEP && c->done) return;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // as in k_adj_pfinish
  __shared__ double dred[256];
  PEER_STAMP(blockIdx.x == 0, 0);
  const PeerView pv = *reinterpret_cast<const PeerView*>(a.peer_base + kPeerOffView);
  unsigned long long* seqp = reinterpret_cast<unsigned long long*>(a.peer_base + kPeerOffSeq);
  const unsigned long long seq = *seqp + 1;
  const int par = (int)(seq & 1);
  const size_t P = (size_t)a.P;
  const size_t rec = (size_t)6 * a.n_loc;                       // doubles per rank record
  const size_t xoff = kPeerOffAdj + (size_t)par * pv.world * rec * 8;
  if (blockIdx.x == 0) {
    const double* mine = a.xrec + (size_t)pv.rank * rec;
    // all loads in flight before the first remote store (a load -> store chain per element cost 15 us here)
    constexpr int kPer = 16;
    for (int base = 0; base < (int)rec; base += kPer * blockDim.x) {
      double v[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int i = base + k * blockDim.x + threadIdx.x;
        v[k] = (i < (int)rec) ? mine[i] : 0.0;
      }
      for (int r = 0; r < pv.world; ++r) {
        double* dst = reinterpret_cast<double*>(pv.base[r] + xoff) + (size_t)pv.rank * rec;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int i = base + k * blockDim.x + threadIdx.x;
          if (i < (int)rec) dst[i] = v[k];
        }
      }
    }
    __syncthreads();
    peer_signal(pv, 0, seq);
    PEER_STAMP(true, 1);
  }
  peer_wait(pv, 0, seq);
  PEER_STAMP(blockIdx.x == 0, 2);
  const int p = a.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  const float rtol = c->rtol, atol = c->atol;
  const float dt = (MODE == ADJ_STEP) ? c->dt : 0.f;
  double part0 = 0.0, part1 = 0.0;
  if (p < a.p_lo + a.p_cnt && p < a.P) {
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    for (int r = 0; r < pv.world; ++r) {
      const float* pc = reinterpret_cast<const float*>(pv.base[r] + a.peer_pcomb_off) + (size_t)par * 4 * P;
      v0 += __ldcv(pc + p);
      if (MODE == ADJ_STEP) { v1 += __ldcv(pc + P + p); v2 += __ldcv(pc + 2 * P + p); v3 += __ldcv(pc + 3 * P + p); }
    }
    if (MODE == ADJ_STEP) {
      const int cur = c->cur, cand = c->cand, midw = c->mid_acc ^ 1;
      const float k0 = a.FP[(size_t)cur * P + p];
      const float s = fmaf(c_tab.c_sol[0], k0, v0);
      const float er = fmaf(c_tab.c_err[0], k0, v1);
      const float md = fmaf(c_tab.c_mid[0], k0, v2);
      const float p0 = a.PB[(size_t)cur * P + p];
      const float p1 = fmaf(dt, s, p0);
      const float tol = atol + rtol * fmaxf(fabsf(p0), fabsf(p1));
      const float q = dt * er / tol;
      part0 = (double)(q * q);
      a.PB[(size_t)cand * P + p] = p1;
      a.FP[(size_t)cand * P + p] = v3;
      a.MIDP[(size_t)midw * P + p] = fmaf(dt, md, p0);
    } else if (MODE == ADJ_INIT0) {
      const float f0 = v0;
      const float p0 = a.PB[p];
      const float sc = atol + fabsf(p0) * rtol;
      const float q0 = p0 / sc, q1 = f0 / sc;
      part0 = (double)(q0 * q0);
      part1 = (double)(q1 * q1);
      a.FP[p] = f0;
    } else if (MODE == ADJ_INIT1) {
      const float sc = atol + fabsf(a.PB[p]) * rtol;
      const float q = (v0 - a.FP[p]) / sc;
      part0 = (double)(q * q);
    }
  }
  dred[threadIdx.x] = part0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.ppart0[blockIdx.x] = dred[0];
  __syncthreads();
  if (MODE == ADJ_INIT0) {
    dred[threadIdx.x] = part1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) a.ppart1[blockIdx.x] = dred[0];
    __syncthreads();
  }
  PEER_STAMP(blockIdx.x == 0, 3);
  if (!elect_last_block(&c->ticket)) return;
  PEER_STAMP(true, 4);
  // ---- round 2: this rank's norm partial(s) to every rank, all ranks' summed in rank order
  const double my0 = block_sum_global(a.ppart0, (int)gridDim.x, dred);
  __syncthreads();
  double my1 = 0.0;
  if (MODE == ADJ_INIT0) { my1 = block_sum_global(a.ppart1, (int)gridDim.x, dred); __syncthreads(); }
  if ((int)threadIdx.x < pv.world) {
    double* dst = reinterpret_cast<double*>(pv.base[threadIdx.x] + kPeerOffPart) + ((size_t)par * kPeerMax + pv.rank) * 2;
    dst[0] = my0;
    dst[1] = my1;
  }
  __syncthreads();
  peer_signal(pv, 1, seq);
  peer_wait(pv, 1, seq);
  PEER_STAMP(true, 5);
  if (threadIdx.x == 0) {
    const volatile double* src = reinterpret_cast<const volatile double*>(pv.base[pv.rank] + kPeerOffPart) + (size_t)par * kPeerMax * 2;
    double t0 = 0.0, t1 = 0.0;
    for (int r = 0; r < pv.world; ++r) { t0 += src[2 * r]; t1 += src[2 * r + 1]; }
    a.ppart0[0] = t0;
    a.ppart1[0] = t1;
  }
  __syncthreads();
  // the finaliser reads every rank's record from the gathered array in this rank's exchange buffer
  const double* gx = reinterpret_cast<const double*>(a.peer_base + xoff);
  if (threadIdx.x == 0) *seqp = seq;   // (thread 0 is the one that survives the finaliser's early return)
  adj_finalize<MODE>(a, c, dred, 1, dt, nullptr, gx);
  PEER_STAMP(true, 6);
}

// ---- segment set-up: state rings at index 0 (lib/ode.py:1516-1518) -------------------------
__global__ void k_adj_seg_begin(AdjArgs a, const float* ys_i, const float* g_y_i, const float* g_r_i,
                                const float* ts, int i, int first, float rtol, float atol, int mxstep, int trace_cap) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < a.N; k += nth) {
    a.Z[k] = ys_i[k];
    a.ZB[k] = first ? g_y_i[k] : a.YBAR[k];
  }
  for (size_t k = tid; k < (size_t)a.na * a.B; k += nth) {
    a.RS[k] = 0.f;   // the split wrappers' residual aux state is zeros (lib/ode.py:1003, 1679)
    a.RB[k] = first ? g_r_i[k] : a.RB[k] + g_r_i[k];
  }
  if (a.fin)
    for (size_t k = tid; k < a.N; k += nth) a.ZE[k] = first ? 0.f : a.EBAR[k];
  for (size_t k = tid; k < (size_t)a.P; k += nth) a.PB[k] = first ? 0.f : a.POUT[k];
  if (tid == 0) {
    Ctrl* c = a.ctrl;
    memset(c, 0, sizeof(Ctrl));
    c->rtol = rtol; c->atol = atol; c->T = 2; c->mxstep = mxstep; c->trace_cap = trace_cap;
    c->n_state = a.n_state;
    AdjScalars* sc = a.sc;
    const float keep = first ? 0.f : sc->t0_out;
    memset(sc, 0, sizeof(AdjScalars));
    sc->t0[0] = keep;
    sc->seg_ts[0] = -ts[i];
    sc->seg_ts[1] = -ts[i - 1];
  }
}

// ---- dense output of the blocks the caller needs: y_bar, params_bar, t0_bar ------------------
__global__ void k_adj_out(AdjArgs a) {
  const Ctrl* c = a.ctrl;
  if (c->out_end <= c->out_begin) return;
  const float t = c->t, last_t = c->last_t, dt = c->step_dt;
  const float x = (a.sc->seg_ts[1] - last_t) / (t - last_t);
  const bool degenerate = (c->n_accepted == 0);
  const int prev = degenerate ? c->cur : c->prev, cur = c->cur, mid = c->mid_acc;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  auto interp = [&](float y0, float y1, float ym, float dy0, float dy1) {
    float ca, cb, cc, cd, ce;
    if (degenerate) {
      ca = cb = cc = cd = ce = y0;
    } else {
      ca = -2.f * dt * dy0 + 2.f * dt * dy1 - 8.f * y0 - 8.f * y1 + 16.f * ym;
      cb = 5.f * dt * dy0 - 3.f * dt * dy1 + 18.f * y0 + 14.f * y1 - 32.f * ym;
      cc = -4.f * dt * dy0 + dt * dy1 - 11.f * y0 - 5.f * y1 + 16.f * ym;
      cd = dt * dy0;
      ce = y0;
    }
    float v = 0.f;
    v = v * x + ca; v = v * x + cb; v = v * x + cc; v = v * x + cd; v = v * x + ce;
    return v;
  };
  for (size_t k = tid; k < a.N; k += nth) {
    const float v = interp(a.ZB[(size_t)prev * a.N + k], a.ZB[(size_t)cur * a.N + k], a.MIDZB[(size_t)mid * a.N + k],
                           a.FZB[(size_t)prev * a.N + k], a.FZB[(size_t)cur * a.N + k]);
    a.YBAR[k] = v + a.g_y_prev[k];
  }
  if (a.fin)
    for (size_t k = tid; k < a.N; k += nth)
      a.EBAR[k] = interp(a.ZE[(size_t)prev * a.N + k], a.ZE[(size_t)cur * a.N + k], a.MIDZE[(size_t)mid * a.N + k],
                         a.FZE[(size_t)prev * a.N + k], a.FZE[(size_t)cur * a.N + k]);
  for (size_t k = tid; k < (size_t)a.P; k += nth)
    a.POUT[k] = interp(a.PB[(size_t)prev * a.P + k], a.PB[(size_t)cur * a.P + k], a.MIDP[(size_t)mid * a.P + k],
                       a.FP[(size_t)prev * a.P + k], a.FP[(size_t)cur * a.P + k]);
  if (tid == 0)
    a.sc->t0_out = interp(a.sc->t0[prev], a.sc->t0[cur], a.sc->midt0[mid], a.sc->ft0[prev], a.sc->ft0[cur]);
}

__global__ void k_adj_finish(AdjArgs a, const float* g_r0, float* r0_bar_out, float* ts_bar0) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < (size_t)a.na * a.B; k += nth) r0_bar_out[k] = a.RB[k] + g_r0[k];
  if (tid == 0) *ts_bar0 = a.sc->t0_out;
}

// W [K][N] -> WT [N][K]
__global__ void k_transpose(const float* __restrict__ W, float* __restrict__ WT, int K, int N) {
  __shared__ float t[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (k0 + j < K && n0 + threadIdx.x < N) t[j][threadIdx.x] = W[(size_t)(k0 + j) * N + n0 + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (n0 + j < N && k0 + threadIdx.x < K) WT[(size_t)(n0 + j) * K + k0 + threadIdx.x] = t[threadIdx.x][j];
}

}  // namespace eno
