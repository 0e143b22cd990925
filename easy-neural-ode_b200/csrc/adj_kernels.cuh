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
#include <string>

#include "adj_device.cuh"
#include "tableaux.cuh"
#include "fwd_kernels.cuh"

namespace eno {

constexpr int kAdjStages = 7;    // dopri5 (odeint() is always dopri5, lib/ode.py:746)
constexpr int kSlots = 6;        // new RHS evaluations per iteration
constexpr int kTile = 32;        // pgrad output tile edge
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
  size_t N;                // B * D
  float n_state;
  // per-sample rings
  float *Z, *ZB;           // [3][N]
  float *FZ, *FZB;         // [3][N]
  float *RS, *FR;          // [3][B]
  float* RB;               // [B]      r_bar (constant during a segment)
  float *MIDZB;            // [2][N]   (only y_bar's interpolant is ever read)
  // shared block
  float *PB, *FP;          // [3][P]
  float* MIDP;             // [2][P]
  // factors
  FactorSlot slot[kSlots];
  // partial sums
  double *rowpart0, *rowpart1, *rowpart2;  // [B]
  double *ppart0, *ppart1;                 // [pgrad grid]
  // segment inputs / outputs
  const float* g_y;        // [N]   g[i] (for t_bar)
  const float* g_r;        // [B]
  float* YBAR;             // [N]   y_bar after the segment (+ g[i-1])
  const float* g_y_prev;   // [N]   g[i-1]
  float* POUT;             // [P]
  float* tbar_out;         // scalar: ts_bar[i]
  // eno_rhs outputs
  float *rhs_dy, *rhs_r, *rhs_zb;
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
  constexpr int NS = (MODE == ADJ_STEP) ? kSlots : 1;
  __shared__ float tile[NS][kTile * kTile];
  __shared__ int pidx[kTile * kTile];   // flat parameter index of each tile element (-1 = padding)
  __shared__ double dred[kSlots * kGroup];
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
    if (grp < NS && mm < M && nn < N) {
      const FactorSlot& sl = a.slot[grp];
      const float* Am = first ? sl.S : sl.Hh;   // [KB][M]
      const float* Bm = first ? sl.A : sl.G;    // [KB][N]
      // factor rows live as [3][B][*]; only the first np*B rows are valid and contiguous
      const float4* Ap = reinterpret_cast<const float4*>(Am + mm);
      const float4* Bp = reinterpret_cast<const float4*>(Bm + nn);
      const size_t lda = (size_t)M / 4, ldb = (size_t)N / 4;
      int k = 0;
      for (; k + 4 <= KB; k += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          av[u] = Ap[(size_t)(k + u) * lda];
          bv[u] = Bp[(size_t)(k + u) * ldb];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float ar[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
          const float br[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
      }
      for (; k < KB; ++k) {
        const float4 a4 = Ap[(size_t)k * lda], b4 = Bp[(size_t)k * ldb];
        const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
        const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
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
      const FactorSlot& sl = a.slot[grp];
      const float* F0 = first ? sl.A : sl.G;             // pass-0 factor [B][W]
      const float* F1 = F0 + (size_t)B * W;              // pass-1 factor
      const float t_stage = -(tau + ((MODE == ADJ_STEP) ? dt * c_tab.alpha[grp] : (MODE == ADJ_INIT1 ? h0 : 0.f)));
      if (o < W) {
        for (int r = 0; r < B; ++r) acc += F0[(size_t)r * W + o];
      } else {
        const int n = o - W;
        if (a.np >= 2) for (int r = 0; r < B; ++r) acc += fmaf(t_stage, F0[(size_t)r * W + n], F1[(size_t)r * W + n]);
        else for (int r = 0; r < B; ++r) acc += t_stage * F0[(size_t)r * W + n];
      }
    }
    if (grp < NS) tile[grp][tig] = acc;
    if (grp == 0) pidx[tig] = (o < 2 * W) ? ((first ? 0 : offb2) + o) : -1;
    n_out = kGroup;
  }
  __syncthreads();

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
      const float f0 = tile[0][e];
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
      const float q = (tile[0][e] - FPc[p]) / sc;
      part0 += (double)(q * q);
    }
  } else {  // ADJ_RHS
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      const int p = pidx[e];
      if (p >= 0) rhs_pbar[p] = tile[0][e];
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
  if (!elect_last_block(&c->ticket)) return;
  __shared__ double fin[12];
  {
    double v;
    v = block_sum_global(a.rowpart0, B, dred); if (threadIdx.x == 0) fin[0] = v; __syncthreads();
    v = block_sum_global(a.ppart0, geo.grid, dred); if (threadIdx.x == 0) fin[1] = v; __syncthreads();
    if (MODE == ADJ_INIT0) {
      v = block_sum_global(a.rowpart1, B, dred); if (threadIdx.x == 0) fin[2] = v; __syncthreads();
      v = block_sum_global(a.ppart1, geo.grid, dred); if (threadIdx.x == 0) fin[3] = v; __syncthreads();
      v = block_sum_global(a.rowpart2, B, dred); if (threadIdx.x == 0) fin[4] = v; __syncthreads();
    }
    // stage derivatives of t0_bar: sums of the per-sample t_bar contributions
    for (int j = 0; j < NS; ++j) {
      const float* tb = a.slot[j].tbar;
      const int per = (B + blockDim.x - 1) / blockDim.x;
      const int b0 = threadIdx.x * per, b1 = min(B, b0 + per);
      double acc = 0.0;
      for (int i = b0; i < b1; ++i) acc += (double)tb[i];
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

// ---- segment set-up: state rings at index 0 (lib/ode.py:1516-1518) -------------------------
__global__ void k_adj_seg_begin(AdjArgs a, const float* ys_i, const float* g_y_i, const float* g_r_i,
                                const float* ts, int i, int first, float rtol, float atol, int mxstep, int trace_cap) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < a.N; k += nth) {
    a.Z[k] = ys_i[k];
    a.ZB[k] = first ? g_y_i[k] : a.YBAR[k];
  }
  for (size_t k = tid; k < (size_t)a.B; k += nth) {
    a.RS[k] = 0.f;   // the split wrappers' residual aux state is zeros (lib/ode.py:1003, 1679)
    a.RB[k] = first ? g_r_i[k] : a.RB[k] + g_r_i[k];
  }
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
  for (size_t k = tid; k < (size_t)a.P; k += nth)
    a.POUT[k] = interp(a.PB[(size_t)prev * a.P + k], a.PB[(size_t)cur * a.P + k], a.MIDP[(size_t)mid * a.P + k],
                       a.FP[(size_t)prev * a.P + k], a.FP[(size_t)cur * a.P + k]);
  if (tid == 0)
    a.sc->t0_out = interp(a.sc->t0[prev], a.sc->t0[cur], a.sc->midt0[mid], a.sc->ft0[prev], a.sc->ft0[cur]);
}

__global__ void k_adj_finish(AdjArgs a, const float* g_r0, float* r0_bar_out, float* ts_bar0) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < (size_t)a.B; k += nth) r0_bar_out[k] = a.RB[k] + g_r0[k];
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

// ======================================================================================
// host side
// ======================================================================================
inline size_t adj_align(size_t x) { return (x + 255) / 256 * 256; }

inline size_t adj_ws_bytes(int B, int D, int H) {
  const size_t N = (size_t)B * D, P = (size_t)(D + 1) * H + H + (size_t)(H + 1) * D + D;
  const PgradGeom g = pgrad_geom(D, H);
  size_t s = 0;
  s += adj_align(sizeof(Ctrl)) + adj_align(sizeof(AdjScalars));
  s += 4 * adj_align(3 * N * 4);                    // Z, ZB, FZ, FZB
  s += 2 * adj_align(3 * (size_t)B * 4) + adj_align((size_t)B * 4);  // RS, FR, RB
  s += adj_align(2 * N * 4);                        // MIDZB
  s += 2 * adj_align(3 * P * 4) + adj_align(2 * P * 4);             // PB, FP, MIDP
  s += kSlots * (2 * adj_align(3 * N * 4) + 2 * adj_align(3 * (size_t)B * H * 4) + adj_align((size_t)B * 4));
  s += 3 * adj_align((size_t)B * 8) + 2 * adj_align((size_t)g.grid * 8);
  s += adj_align(N * 4) + adj_align(P * 4);         // YBAR, POUT
  s += 2 * adj_align((size_t)D * H * 4);            // transposes
  return s + 4096;
}

struct AdjCarver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off += adj_align(n * sizeof(T));
    return p;
  }
};

inline void adj_carve(void* ws, int B, int D, int H, const float* params, AdjArgs* a) {
  const size_t N = (size_t)B * D, P = (size_t)(D + 1) * H + H + (size_t)(H + 1) * D + D;
  const PgradGeom g = pgrad_geom(D, H);
  AdjCarver cv{static_cast<char*>(ws)};
  a->ctrl = cv.take<Ctrl>(1);
  a->sc = cv.take<AdjScalars>(1);
  a->Z = cv.take<float>(3 * N); a->ZB = cv.take<float>(3 * N);
  a->FZ = cv.take<float>(3 * N); a->FZB = cv.take<float>(3 * N);
  a->RS = cv.take<float>(3 * (size_t)B); a->FR = cv.take<float>(3 * (size_t)B); a->RB = cv.take<float>(B);
  a->MIDZB = cv.take<float>(2 * N);
  a->PB = cv.take<float>(3 * P); a->FP = cv.take<float>(3 * P); a->MIDP = cv.take<float>(2 * P);
  for (int j = 0; j < kSlots; ++j) {
    a->slot[j].S = cv.take<float>(3 * N);
    a->slot[j].G = cv.take<float>(3 * N);
    a->slot[j].A = cv.take<float>(3 * (size_t)B * H);
    a->slot[j].Hh = cv.take<float>(3 * (size_t)B * H);
    a->slot[j].tbar = cv.take<float>(B);
  }
  a->rowpart0 = cv.take<double>(B); a->rowpart1 = cv.take<double>(B); a->rowpart2 = cv.take<double>(B);
  a->ppart0 = cv.take<double>(g.grid); a->ppart1 = cv.take<double>(g.grid);
  a->YBAR = cv.take<float>(N);
  a->POUT = cv.take<float>(P);
  float* W1xT = cv.take<float>((size_t)D * H);
  float* W2xT = cv.take<float>((size_t)D * H);
  a->w = mlp_weights_from_flat(params, D, H, W1xT, W2xT);
  a->B = B; a->P = (int)P; a->N = N;
  a->n_state = (float)(2 * (N + (size_t)B) + 1 + N + P);
}

inline int adj_pick_rows(int B, int D, int H, int sms) {
  int R = 1;
  if (B > sms) R = 2;
  while (R > 1 && adj_smem_bytes(R, D, H, kAdjStages) > 220 * 1024) R >>= 1;
  return R;
}

#define ENO_ADJ_CHECK(expr)                                                                 \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(e_);                   \
      return ENO_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

template <int R, int K>
int adj_launch_rows(int mode, const AdjArgs& a, int grid, size_t smem, float rhs_tau, cudaStream_t st,
                    std::string* err) {
  switch (mode) {
    case ADJ_STEP:
      ENO_ADJ_CHECK(cudaFuncSetAttribute(k_adj_rows<R, K, ADJ_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_adj_rows<R, K, ADJ_STEP><<<grid, kThreads, smem, st>>>(a, rhs_tau);
      break;
    case ADJ_INIT0:
      ENO_ADJ_CHECK(cudaFuncSetAttribute(k_adj_rows<R, K, ADJ_INIT0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_adj_rows<R, K, ADJ_INIT0><<<grid, kThreads, smem, st>>>(a, rhs_tau);
      break;
    case ADJ_INIT1:
      ENO_ADJ_CHECK(cudaFuncSetAttribute(k_adj_rows<R, K, ADJ_INIT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_adj_rows<R, K, ADJ_INIT1><<<grid, kThreads, smem, st>>>(a, rhs_tau);
      break;
    default:
      ENO_ADJ_CHECK(cudaFuncSetAttribute(k_adj_rows<R, K, ADJ_RHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_adj_rows<R, K, ADJ_RHS><<<grid, kThreads, smem, st>>>(a, rhs_tau);
  }
  return 0;
}

inline int adj_rows(int R, int K, int mode, const AdjArgs& a, int grid, size_t smem, float rhs_tau, cudaStream_t st,
                    std::string* err) {
#define ENO_RK(RR, KK) if (R == RR && K == KK) return adj_launch_rows<RR, KK>(mode, a, grid, smem, rhs_tau, st, err);
  ENO_RK(1, 0) ENO_RK(1, 2) ENO_RK(1, 3) ENO_RK(2, 0) ENO_RK(2, 2) ENO_RK(2, 3)
#undef ENO_RK
  if (err) *err = "adjoint: unsupported (rows per CTA, reg_order) combination";
  return ENO_E_UNSUPPORTED;
}

inline void adj_transposes(const AdjArgs& a, cudaStream_t st) {
  const int D = a.w.D, H = a.w.H;
  dim3 blk(32, 8);
  k_transpose<<<dim3((H + 31) / 32, (D + 31) / 32), blk, 0, st>>>(a.w.W1x, const_cast<float*>(a.w.W1xT), D, H);
  k_transpose<<<dim3((D + 31) / 32, (H + 31) / 32), blk, 0, st>>>(a.w.W2x, const_cast<float*>(a.w.W2xT), H, D);
}

inline int adj_solve(Profiler* prof, std::string* err, Ctrl* mailbox, int* hint, int sms, const eno_dynamics_desc* desc, int B,
                     const float* ys, const float* ts, int T, const float* params, const float* g_y, const float* g_r,
                     float rtol, float atol, int mxstep, void* ws, size_t ws_bytes, float* y0_bar_out,
                     float* r0_bar_out, float* ts_bar_out, float* params_bar_out, eno_stats* stats,
                     eno_trace_row* trace, int trace_cap, cudaStream_t st) {
  const int D = desc->D, H = desc->H, K = desc->reg_order;
  if (!(K == 0 || K == 2 || K == 3)) { if (err) *err = "eno_odeint_bwd: reg_order must be 0, 2 or 3"; return ENO_E_UNSUPPORTED; }
  if (B <= 0 || T < 1 || !ys || !ts || !params || !g_y || !g_r || !ws || !y0_bar_out || !r0_bar_out || !ts_bar_out ||
      !params_bar_out) { if (err) *err = "eno_odeint_bwd: null pointer or empty batch"; return ENO_E_ARG; }
  if (ws_bytes < adj_ws_bytes(B, D, H)) { if (err) *err = "eno_odeint_bwd: workspace too small"; return ENO_E_WORKSPACE; }
  const size_t N = (size_t)B * D;
  AdjArgs a = {};
  adj_carve(ws, B, D, H, params, &a);
  a.np = (K == 0) ? 1 : K;
  {
    const size_t n_aux = (desc->aug == ENO_AUG_NONE) ? 0 : 1;
    a.n_state = (float)(2 * (N + n_aux * (size_t)B) + 1 + (desc->eps_in_args ? N : 0) + (size_t)a.P);
  }
  a.trace = trace;
  Tableau tab;
  fill_tableau(ENO_TABLEAU_DOPRI5, &tab);
  ENO_ADJ_CHECK(cudaMemcpyToSymbolAsync(c_tab, &tab, sizeof(Tableau), 0, cudaMemcpyHostToDevice, st));
  adj_transposes(a, st);
  const int R = adj_pick_rows(B, D, H, sms);
  const size_t smem = adj_smem_bytes(R, D, H, kAdjStages);
  const int grid = (B + R - 1) / R;
  const PgradGeom geo = pgrad_geom(D, H);
  const int egrid = (int)std::min<size_t>((std::max(N, (size_t)a.P) + 255) / 256, (size_t)4 * sms);
  const int mx = mxstep <= 0 ? INT_MAX : mxstep;
  int worst = ENO_OK;
  if (T == 1) {   // nothing to integrate: y0_bar = g[0], everything else zero (empty scan)
    ENO_ADJ_CHECK(cudaMemcpyAsync(y0_bar_out, g_y, N * 4, cudaMemcpyDeviceToDevice, st));
    ENO_ADJ_CHECK(cudaMemcpyAsync(r0_bar_out, g_r, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(ts_bar_out, 0, 4, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(params_bar_out, 0, (size_t)a.P * 4, st));
    ENO_ADJ_CHECK(cudaStreamSynchronize(st));
    return ENO_OK;
  }
  for (int i = T - 1; i >= 1; --i) {
    int launches = 0;
    const int first = (i == T - 1);
    a.g_y = g_y + (size_t)i * N;
    a.g_r = g_r + (size_t)i * B;
    a.g_y_prev = g_y + (size_t)(i - 1) * N;
    a.tbar_out = ts_bar_out + i;
    k_adj_seg_begin<<<egrid, 256, 0, st>>>(a, ys + (size_t)i * N, a.g_y, a.g_r, ts, i, first, rtol, atol, mx,
                                           trace ? trace_cap : 0);
    int rc;
    if ((rc = adj_rows(R, K, ADJ_INIT0, a, grid, smem, 0.f, st, err))) return rc;
    k_adj_pgrad<ADJ_INIT0><<<geo.grid, kGroup, 0, st>>>(a, 0.f, nullptr, nullptr);
    if ((rc = adj_rows(R, K, ADJ_INIT1, a, grid, smem, 0.f, st, err))) return rc;
    k_adj_pgrad<ADJ_INIT1><<<geo.grid, kGroup, 0, st>>>(a, 0.f, nullptr, nullptr);
    launches += 5;
    int chunk = *hint;
    for (;;) {
      for (int k = 0; k < chunk; ++k) {
        int ps = prof->begin(PROF_ADJ_ROWS, st);
        if ((rc = adj_rows(R, K, ADJ_STEP, a, grid, smem, 0.f, st, err))) return rc;
        prof->end(ps, st);
        ps = prof->begin(PROF_ADJ_PGRAD, st);
        k_adj_pgrad<ADJ_STEP><<<geo.grid, kSlots * kGroup, 0, st>>>(a, 0.f, nullptr, nullptr);
        prof->end(ps, st);
        k_adj_out<<<egrid, 256, 0, st>>>(a);
        launches += 3;
      }
      ENO_ADJ_CHECK(cudaMemcpyAsync(mailbox, a.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
      ENO_ADJ_CHECK(cudaStreamSynchronize(st));
      if (mailbox->done) break;
      chunk = 4;
    }
    {
      int valid[PROF_KINDS] = {0, mailbox->n_iters, mailbox->n_iters, 0};
      prof->resolve(valid);
    }
    *hint = mailbox->n_iters + 1 > 64 ? 64 : mailbox->n_iters + 1;
    if (stats) {
      eno_stats* s = stats + (T - 1 - i);
      s->nfe = (float)mailbox->nfe; s->n_iters = mailbox->n_iters; s->n_accepted = mailbox->n_accepted;
      s->status = mailbox->status; s->last_dt = mailbox->dt; s->last_t = mailbox->t; s->n_launches = launches;
      s->n_rhs = 2 + 6 * mailbox->n_iters;
    }
    if (mailbox->status != ENO_OK && worst == ENO_OK) worst = mailbox->status;
    trace = nullptr;  // only the first segment is traced
    a.trace = nullptr;
  }
  k_adj_finish<<<(B + 255) / 256, 256, 0, st>>>(a, g_r, r0_bar_out, ts_bar_out);
  ENO_ADJ_CHECK(cudaMemcpyAsync(y0_bar_out, a.YBAR, N * 4, cudaMemcpyDeviceToDevice, st));
  ENO_ADJ_CHECK(cudaMemcpyAsync(params_bar_out, a.POUT, (size_t)a.P * 4, cudaMemcpyDeviceToDevice, st));
  ENO_ADJ_CHECK(cudaStreamSynchronize(st));
  ENO_ADJ_CHECK(cudaGetLastError());
  return worst;
}

inline int adj_rhs(std::string* err, int sms, const eno_dynamics_desc* desc, int mode, int B, const float* y, float t,
                   const float* params, const float* y_bar, const float* r_bar, void* ws, size_t ws_bytes, float* dy_out,
                   float* aux_out, float* ybar_dot_out, float* tbar_out, float* params_bar_out, cudaStream_t st) {
  (void)sms;
  const int D = desc->D, H = desc->H, K = desc->reg_order;
  if (!(K == 0 || K == 2 || K == 3)) { if (err) *err = "eno_rhs: reg_order must be 0, 2 or 3"; return ENO_E_UNSUPPORTED; }
  if (!ws || ws_bytes < adj_ws_bytes(B, D, H)) { if (err) *err = "eno_rhs: workspace too small"; return ENO_E_WORKSPACE; }
  if (mode == 2 && (!y_bar || !r_bar || !ybar_dot_out || !tbar_out || !params_bar_out || !aux_out)) {
    if (err) *err = "eno_rhs: mode 2 needs y_bar, r_bar and all outputs"; return ENO_E_ARG;
  }
  const size_t N = (size_t)B * D;
  AdjArgs a = {};
  adj_carve(ws, B, D, H, params, &a);
  a.np = (K == 0) ? 1 : K;
  adj_transposes(a, st);
  ENO_ADJ_CHECK(cudaMemcpyAsync(a.Z, y, N * 4, cudaMemcpyDeviceToDevice, st));
  if (mode == 2) {
    ENO_ADJ_CHECK(cudaMemcpyAsync(a.ZB, y_bar, N * 4, cudaMemcpyDeviceToDevice, st));
    ENO_ADJ_CHECK(cudaMemcpyAsync(a.RB, r_bar, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    ENO_ADJ_CHECK(cudaMemsetAsync(a.ZB, 0, N * 4, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(a.RB, 0, (size_t)B * 4, st));
  }
  ENO_ADJ_CHECK(cudaMemsetAsync(a.RS, 0, (size_t)B * 4, st));
  ENO_ADJ_CHECK(cudaMemsetAsync(a.ctrl, 0, sizeof(Ctrl), st));
  a.rhs_dy = dy_out;
  a.rhs_r = aux_out;
  a.rhs_zb = ybar_dot_out;
  const size_t smem = adj_smem_bytes(1, D, H, 2);
  int rc;
  // solver time tau = -t, so that the kernel's t = -tau is the caller's t
  if ((rc = adj_rows(1, K, ADJ_RHS, a, B, smem, -t, st, err))) return rc;
  if (mode == 2) {
    const PgradGeom geo = pgrad_geom(D, H);
    k_adj_pgrad<ADJ_RHS><<<geo.grid, kGroup, 0, st>>>(a, -t, params_bar_out, tbar_out);
  }
  ENO_ADJ_CHECK(cudaGetLastError());
  ENO_ADJ_CHECK(cudaStreamSynchronize(st));
  return ENO_OK;
}

}  // namespace eno
