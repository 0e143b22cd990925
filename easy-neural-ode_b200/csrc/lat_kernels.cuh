// Per-trajectory differentiable solves for small time-independent tanh stacks: the generative ODE of latent_ode.py
// (BASELINE config C3).  Replaces, per trajectory,
//     odeint(augment_dynamics(gen_dynamics_wrap), aug_init(z), timesteps, params)      latent_ode.py:344-350
// as mapped by jax.vmap over 3 x 50 trajectories: under vmap every trajectory has its own adaptive step sequence
// (lib/ode.py:823-862) and, in the backward pass, its own adjoint state INCLUDING its own params_bar block
// (lib/ode.py:1494-1529, 9,763 elements at the script's sizes); the per-trajectory params_bar are summed at the end.
//
// B200 design: one CTA of 128 threads = one trajectory, two CTAs per SM (150 trajectories are co-resident on 148 SMs),
// no grid-wide event, no host round-trip: the WHOLE forward solve (all output times) or the WHOLE adjoint (all
// segments) of a trajectory runs inside one launch.  The weights (77.8 KB in fp64) stay in shared memory for the
// launch; the three Taylor streams of the regulariser (primal, first and second total derivative, closed form of
// jax.experimental.jet for a tanh stack) and their hand-derived reverse sweep run out of shared
// memory; the params_bar block of the adjoint state lives in global memory (L2-resident) and its seven stage
// derivatives are never materialised: per stage only the factors of the three outer products per layer
// (a_c, g_c; 1,320 numbers at the script's sizes) are kept, and the candidate / error / norm of the 9,720 parameter
// elements is formed from the factors in one pass per iteration.
//
// Real = double reproduces the script (jax_enable_x64, latent_ode.py:23); Real = float is the fp32 variant.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/eno.h"

namespace eno {
namespace lat {

constexpr int kMaxL = 8;       // linear layers
constexpr int kMaxW = 64;      // widest layer (2 threads per output unit)
constexpr int kThreads = 128;
constexpr int kStages = 7;     // dopri5
constexpr int kNC = 3 * kStages;   // (stage, stream) pairs of one layer's params_bar derivative

struct Net {
  int L;                // linear layers: x_{l+1} = tanh(x_l) W_l + b_l, f = x_L       latent_ode.py:196-208
  int K;                // order of the regulariser: 0 none, 2, 3 (sol_recursive latent_ode.py:82-108)
  int aug;              // 1: state (z, r), 0: state z
  int d[kMaxL + 1];     // widths, d[0] = d[L] = latent dimension
  int ld[kMaxL];        // row pitch of W_l in shared memory (odd)
  int wo[kMaxL];        // offset of W_l in the shared weight block
  int bo[kMaxL];        // offset of b_l in the shared weight block
  int po[kMaxL];        // offset of layer l in the flat parameter vector: [ b_l (out) | W_l (in x out) ]
  int io[kMaxL + 1];    // prefix sums of the layer input widths
  int oo[kMaxL + 1];    // prefix sums of the layer output widths
  int IN, OUT;          // totals of the two
  int P;                // parameters
  int wtot;             // elements of the shared weight block
  int FS;               // factor record of one stage: 3 IN + 3 OUT
  int scratch;          // elements of the work area (also holds one layer's factors of all stages)
};

// dopri5, spelled like lib/ode.py:108-120 (+ the mid-point row 58-61); double constants, rounded to Real at use
__device__ __constant__ double c_beta[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
__device__ __constant__ double c_sol[7] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0};
__device__ __constant__ double c_err[7] = {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085,
                                           125.0 / 192 - 451.0 / 720, -2187.0 / 6784 - -12231.0 / 42400,
                                           11.0 / 84 - 649.0 / 6300, -1. / 60.};
__device__ __constant__ double c_mid[7] = {6025192743.0 / 30085553152.0 / 2, 0, 51252292925.0 / 65400821598.0 / 2,
                                           -2691868925.0 / 45128329728.0 / 2, 187940372067.0 / 1594534317056.0 / 2,
                                           -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2};

__device__ __forceinline__ double tanhR(double x) { return tanh(x); }
__device__ __forceinline__ float tanhR(float x) { return tanhf(x); }
__device__ __forceinline__ double sqrtR(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrtR(float x) { return sqrtf(x); }
__device__ __forceinline__ double powR(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float powR(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double fmaR(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fmaR(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double absR(double x) { return fabs(x); }
__device__ __forceinline__ float absR(float x) { return fabsf(x); }

// optimal_step_size, lib/ode.py:529-538 (np.maximum / np.minimum propagate NaN, so plain comparisons)
template <typename R>
__device__ __forceinline__ R next_dt(R last, R ratio) {
  const R dfactor = (ratio < R(1)) ? R(1) : R(0.2);
  R factor = powR(sqrtR(ratio), R(1.0 / 5.0)) / R(0.9);
  const R hi = R(1) / dfactor;
  if (factor > hi) factor = hi;
  if (factor < R(0.1)) factor = R(0.1);
  if (ratio != ratio) factor = ratio;
  return (ratio == R(0)) ? last * R(10) : last / factor;
}

// initial_step_size tail, lib/ode.py:93-104 (order = 4 for dopri5)
template <typename R>
__device__ __forceinline__ R init_h0(R d0, R d1) {
  return (d0 < R(1e-5) || d1 < R(1e-5)) ? R(1e-6) : R(0.01) * d0 / d1;
}
template <typename R>
__device__ __forceinline__ R init_dt(R h0, R d1, R d2) {
  R h1;
  if (d1 <= R(1e-15) && d2 <= R(1e-15)) {
    h1 = h0 * R(1e-3);
    if (h1 < R(1e-6)) h1 = R(1e-6);
  } else {
    h1 = powR(R(0.01) / (d1 + d2), R(1.0 / 5.0));
  }
  const R a = R(100) * h0;
  if (a != a || h1 != h1) return a + h1;   // NaN propagates like np.minimum
  return a < h1 ? a : h1;
}

// quartic through (y0, y_mid, y1) with end slopes, evaluated at x (lib/ode.py:56-63, 78-84, 849-850)
template <typename R>
__device__ __forceinline__ R quartic_at(R y0, R y1, R ym, R dy0, R dy1, R dt, R x) {
  const R a = R(-2) * dt * dy0 + R(2) * dt * dy1 - R(8) * y0 - R(8) * y1 + R(16) * ym;
  const R b = R(5) * dt * dy0 - R(3) * dt * dy1 + R(18) * y0 + R(14) * y1 - R(32) * ym;
  const R c = R(-4) * dt * dy0 + dt * dy1 - R(11) * y0 - R(5) * y1 + R(16) * ym;
  const R d = dt * dy0;
  R y = R(0) * x;
  y = y * x + a;
  y = y * x + b;
  y = y * x + c;
  y = y * x + d;
  y = y * x + y0;
  return y;
}

template <typename R>
struct Sm {
  R* W;                          // weights (padded rows) and biases
  R *A, *AU, *AV, *GX, *GU, *GV; // the factor record of the current right-hand side (contiguous, FS elements)
  R *P1, *P2, *U, *V, *P1B, *P2B, *UB;
  R *F, *Y1, *Y2, *FB;           // stream outputs f, x'', x''' and the cotangent that reaches f through the seeds
  R *zin, *zbin, *zbdot;         // stage inputs / vjp wrt z
  R* red;                        // 16 scalars
  R* st;                         // solver vectors (see the kernels)
};

template <typename R>
__host__ __device__ inline size_t smem_elems(const Net& n, int solver_elems) {
  return (size_t)n.wtot + n.scratch + 7 * (size_t)n.d[0] + 16 + solver_elems;
}

template <typename R>
__device__ __forceinline__ void carve(const Net& n, unsigned char* raw, Sm<R>& s) {
  R* p = reinterpret_cast<R*>(raw);
  s.W = p; p += n.wtot;
  R* work = p; p += n.scratch;
  s.A = work; s.AU = s.A + n.IN; s.AV = s.AU + n.IN;
  s.GX = s.AV + n.IN; s.GU = s.GX + n.OUT; s.GV = s.GU + n.OUT;
  s.P1 = s.GV + n.OUT; s.P2 = s.P1 + n.IN; s.U = s.P2 + n.IN; s.V = s.U + n.IN;
  s.P1B = s.V + n.IN; s.P2B = s.P1B + n.IN; s.UB = s.P2B + n.IN;
  const int d0 = n.d[0];
  s.F = p; s.Y1 = s.F + d0; s.Y2 = s.Y1 + d0; s.FB = s.Y2 + d0;
  s.zin = s.FB + d0; s.zbin = s.zin + d0; s.zbdot = s.zbin + d0;
  p += 7 * d0;
  s.red = p; p += 16;
  s.st = p;
}

// flat parameters [ b_l | W_l ] (Haiku ravel order, sorted keys) -> shared memory with padded rows
template <typename R>
__device__ __forceinline__ void load_weights(const Net& n, const Sm<R>& s, const R* __restrict__ params) {
  for (int l = 0; l < n.L; ++l) {
    const int din = n.d[l], dout = n.d[l + 1];
    const R* src = params + n.po[l];
    for (int k = threadIdx.x; k < dout; k += kThreads) s.W[n.bo[l] + k] = src[k];
    for (int k = threadIdx.x; k < din * dout; k += kThreads) {
      const int i = k / dout, o = k - i * dout;
      s.W[n.wo[l] + i * n.ld[l] + o] = src[dout + k];
    }
  }
}

// sum over rows i = half, half + 2, ... of in[i] * W[i][o]
template <typename R>
__device__ __forceinline__ R dot_rows(const R* __restrict__ W, int ld, int din, const R* __restrict__ in, int o, int half) {
  R a0 = R(0), a1 = R(0);
  int i = half;
  for (; i + 2 < din; i += 4) {
    a0 = fmaR(in[i], W[i * ld + o], a0);
    a1 = fmaR(in[i + 2], W[(i + 2) * ld + o], a1);
  }
  if (i < din) a0 = fmaR(in[i], W[i * ld + o], a0);
  return a0 + a1;
}
// sum over columns o = half, half + 2, ... of W[i][o] * g[o]
template <typename R>
__device__ __forceinline__ R dot_cols(const R* __restrict__ W, int ld, int dout, const R* __restrict__ g, int i, int half) {
  const R* row = W + i * ld;
  R a0 = R(0), a1 = R(0);
  int o = half;
  for (; o + 2 < dout; o += 4) {
    a0 = fmaR(row[o], g[o], a0);
    a1 = fmaR(row[o + 2], g[o + 2], a1);
  }
  if (o < dout) a0 = fmaR(row[o], g[o], a0);
  return a0 + a1;
}

// The three streams through the stack (closed form of the Taylor-mode recursion for x' = f(x), f time-independent):
//   S0: a_l = tanh(x_l), x_{l+1} = a_l W_l + b_l                                   f   = x_L
//   S1: u_0 = f,   u_{l+1} = (p1_l u_l) W_l                  p1 = 1 - a^2          x'' = u_L
//   S2: v_0 = x'', v_{l+1} = (p2_l u_l^2 + p1_l v_l) W_l     p2 = -2 a p1          x''' = v_L
// Leaves A/AU/AV (layer inputs of the three streams) and P1, P2, U, V in shared memory for the reverse sweep.
// Returns mean(r^2), r = x'' (K = 2) or x''' (K = 3), in s.red[0] (valid after the trailing barrier).
template <typename R>
__device__ __forceinline__ void rhs_forward(const Net& n, const Sm<R>& s, const R* __restrict__ zin) {
  const int tid = threadIdx.x, o = tid >> 1, half = tid & 1;
  const int d0 = n.d[0];
  if (tid < d0) {
    const R a = tanhR(zin[tid]);
    const R p1 = R(1) - a * a;
    s.A[tid] = a; s.P1[tid] = p1; s.P2[tid] = R(-2) * a * p1;
  }
  __syncthreads();
  for (int l = 0; l < n.L; ++l) {
    const int din = n.d[l], dout = n.d[l + 1];
    R acc = (o < dout) ? dot_rows(s.W + n.wo[l], n.ld[l], din, s.A + n.io[l], o, half) : R(0);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0 && o < dout) {
      const R x = acc + s.W[n.bo[l] + o];
      if (l + 1 < n.L) {
        const R a = tanhR(x);
        const R p1 = R(1) - a * a;
        const int k = n.io[l + 1] + o;
        s.A[k] = a; s.P1[k] = p1; s.P2[k] = R(-2) * a * p1;
      } else {
        s.F[o] = x;
      }
    }
    __syncthreads();
  }
  if (n.K >= 2) {
    if (tid < d0) { const R u = s.F[tid]; s.U[tid] = u; s.AU[tid] = s.P1[tid] * u; }
    __syncthreads();
    for (int l = 0; l < n.L; ++l) {
      const int din = n.d[l], dout = n.d[l + 1];
      R acc = (o < dout) ? dot_rows(s.W + n.wo[l], n.ld[l], din, s.AU + n.io[l], o, half) : R(0);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0 && o < dout) {
        if (l + 1 < n.L) { const int k = n.io[l + 1] + o; s.U[k] = acc; s.AU[k] = s.P1[k] * acc; }
        else s.Y1[o] = acc;
      }
      __syncthreads();
    }
  }
  if (n.K >= 3) {
    if (tid < d0) { const R v = s.Y1[tid], u = s.U[tid]; s.V[tid] = v; s.AV[tid] = s.P2[tid] * u * u + s.P1[tid] * v; }
    __syncthreads();
    for (int l = 0; l < n.L; ++l) {
      const int din = n.d[l], dout = n.d[l + 1];
      R acc = (o < dout) ? dot_rows(s.W + n.wo[l], n.ld[l], din, s.AV + n.io[l], o, half) : R(0);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0 && o < dout) {
        if (l + 1 < n.L) { const int k = n.io[l + 1] + o; const R u = s.U[k]; s.V[k] = acc; s.AV[k] = s.P2[k] * u * u + s.P1[k] * acc; }
        else s.Y2[o] = acc;
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    R m = R(0);
    if (n.K >= 2) {
      const R* r = (n.K >= 3) ? s.Y2 : s.Y1;
      for (int e = 0; e < d0; ++e) m = fmaR(r[e], r[e], m);
      m = m / R(d0);
    }
    s.red[0] = m;
  }
  __syncthreads();
}

// Reverse sweep of  z_bar . f + r_bar * mean(r^2)  over the three streams (S2, then S1, then S0: the top cotangent of
// a stream is the seed cotangent the stream above it leaves).  Afterwards GX/GU/GV hold the cotangents of the layer
// OUTPUTS of the three streams, so that  W_l_bar = A_l (x) GX_l + AU_l (x) GU_l + AV_l (x) GV_l,  b_l_bar = GX_l,
// and s.zbdot the vjp wrt z.
template <typename R>
__device__ __forceinline__ void rhs_reverse(const Net& n, const Sm<R>& s, const R* __restrict__ zbin, R rb) {
  const int tid = threadIdx.x, i = tid >> 1, half = tid & 1;
  const int L = n.L, d0 = n.d[0];
  for (int k = tid; k < n.IN; k += kThreads) { s.P1B[k] = R(0); s.P2B[k] = R(0); s.UB[k] = R(0); }
  for (int k = tid; k < n.OUT; k += kThreads) {
    if (n.K < 2) s.GU[k] = R(0);
    if (n.K < 3) s.GV[k] = R(0);
  }
  if (n.K < 2) for (int k = tid; k < n.IN; k += kThreads) s.AU[k] = R(0);
  if (n.K < 3) for (int k = tid; k < n.IN; k += kThreads) s.AV[k] = R(0);
  __syncthreads();
  const R sc = R(2) * rb / R(d0);
  if (tid < d0) {
    if (n.K >= 3) s.GV[n.oo[L - 1] + tid] = sc * s.Y2[tid];
    else if (n.K == 2) s.GU[n.oo[L - 1] + tid] = sc * s.Y1[tid];
  }
  __syncthreads();
  if (n.K >= 3) {
    for (int l = L - 1; l >= 0; --l) {
      const int din = n.d[l], dout = n.d[l + 1];
      R acc = (i < din) ? dot_cols(s.W + n.wo[l], n.ld[l], dout, s.GV + n.oo[l], i, half) : R(0);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0 && i < din) {
        const int k = n.io[l] + i;
        const R u = s.U[k];
        s.P2B[k] += u * u * acc;
        s.P1B[k] += s.V[k] * acc;
        s.UB[k] = R(2) * s.P2[k] * u * acc;
        const R gv = s.P1[k] * acc;
        if (l > 0) s.GV[n.oo[l - 1] + i] = gv;
        else s.GU[n.oo[L - 1] + i] = gv;   // cotangent of x'' = top of S1
      }
      __syncthreads();
    }
  }
  if (n.K >= 2) {
    for (int l = L - 1; l >= 0; --l) {
      const int din = n.d[l], dout = n.d[l + 1];
      R acc = (i < din) ? dot_cols(s.W + n.wo[l], n.ld[l], dout, s.GU + n.oo[l], i, half) : R(0);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0 && i < din) {
        const int k = n.io[l] + i;
        s.P1B[k] += s.U[k] * acc;
        const R gu = s.P1[k] * acc + s.UB[k];
        if (l > 0) s.GU[n.oo[l - 1] + i] = gu;
        else s.FB[i] = gu;                 // reaches f through the seed u_0 = f
      }
      __syncthreads();
    }
  }
  if (tid < d0) s.GX[n.oo[L - 1] + tid] = zbin[tid] + (n.K >= 2 ? s.FB[tid] : R(0));
  __syncthreads();
  for (int l = L - 1; l >= 0; --l) {
    const int din = n.d[l], dout = n.d[l + 1];
    R acc = (i < din) ? dot_cols(s.W + n.wo[l], n.ld[l], dout, s.GX + n.oo[l], i, half) : R(0);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0 && i < din) {
      const int k = n.io[l] + i;
      const R a = s.A[k];
      const R tot = acc + s.P1B[k] * (R(-2) * a) + s.P2B[k] * (R(-2) + R(6) * a * a);
      const R gx = s.P1[k] * tot;
      if (l > 0) s.GX[n.oo[l - 1] + i] = gx;
      else s.zbdot[i] = gx;
    }
    __syncthreads();
  }
}

template <typename R>
__device__ __forceinline__ R block_sum(R v, R* red /*[>= 4]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  R t = red[0];
  for (int w = 1; w < kThreads / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------------------------
// Forward: one CTA = one trajectory, all output times.  State e = 0..d0-1: z, e = d0: r (when n.aug).
//   z0 [n_traj, d0]   ts [T]   zs_out [n_traj, T, d0]   rs_out [n_traj, T] (aug)   nfe_out [n_traj] (Real)
//   iters_out [n_traj, 2] (iterations, accepted) or NULL   status_out [n_traj] or NULL
// ------------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kThreads, 2)
k_traj_fwd(const __grid_constant__ Net n, const R* __restrict__ z0, const R* __restrict__ ts, int T, const R* __restrict__ params,
           R rtol, R atol, int mxstep, R* __restrict__ zs_out, R* __restrict__ rs_out, R* __restrict__ nfe_out,
           int* __restrict__ iters_out, int* __restrict__ status_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm<R> s;
  carve(n, smem_raw, s);
  const int tid = threadIdx.x, traj = blockIdx.x;
  const int d0 = n.d[0], ns = d0 + (n.aug ? 1 : 0);
  // solver vectors: y[ns], k[7][ns], coef[5][ns], y1[ns]
  R* y = s.st;
  R* k = y + ns;
  R* coef = k + kStages * ns;
  R* y1v = coef + 5 * ns;
  load_weights(n, s, params);
  if (tid < ns) y[tid] = (tid < d0) ? z0[(size_t)traj * d0 + tid] : R(0);
  __syncthreads();
  // f0 and the starting step (lib/ode.py:853-858, 86-104)
  rhs_forward(n, s, y);
  if (tid < ns) k[tid] = (tid < d0) ? s.F[tid] : s.red[0];
  __syncthreads();
  R d0n, d1n;
  {
    R q0 = R(0), q1 = R(0);
    if (tid < ns) { const R sc = atol + absR(y[tid]) * rtol; q0 = y[tid] / sc; q1 = k[tid] / sc; q0 *= q0; q1 *= q1; }
    d0n = sqrtR(block_sum(q0, s.red + 4));
    d1n = sqrtR(block_sum(q1, s.red + 4));
  }
  const R h0 = init_h0(d0n, d1n);
  if (tid < d0) s.zin[tid] = y[tid] + h0 * k[tid];
  __syncthreads();
  rhs_forward(n, s, s.zin);
  R dt;
  {
    R q = R(0);
    if (tid < ns) { const R sc = atol + absR(y[tid]) * rtol; const R f1 = (tid < d0) ? s.F[tid] : s.red[0]; q = (f1 - k[tid]) / sc; q *= q; }
    const R d2 = sqrtR(block_sum(q, s.red + 4)) / h0;
    dt = init_dt(h0, d1n, d2);
  }
  if (tid < ns) {
    for (int c = 0; c < 5; ++c) coef[c * ns + tid] = y[tid];
    if (tid < d0) zs_out[((size_t)traj * T) * d0 + tid] = y[tid];
    else if (rs_out) rs_out[(size_t)traj * T] = y[tid];
  }
  R t = ts[0], last_t = ts[0];
  int n_iters = 0, n_acc = 0, status = ENO_OK;
  for (int it = 1; it < T; ++it) {
    const R target = ts[it];
    int i = 0;
    while (t < target && (mxstep <= 0 || i < mxstep) && dt > R(0)) {
      for (int j = 1; j < kStages; ++j) {
        if (tid < d0) {
          R acc = R(0);
          for (int m = 0; m < j; ++m) acc = fmaR(R(c_beta[j - 1][m]), k[m * ns + tid], acc);
          s.zin[tid] = y[tid] + dt * acc;
        }
        __syncthreads();
        rhs_forward(n, s, s.zin);
        if (tid < ns) k[j * ns + tid] = (tid < d0) ? s.F[tid] : s.red[0];
        __syncthreads();
      }
      R q = R(0), ynew = R(0), ymid = R(0);
      if (tid < ns) {
        R sa = R(0), ea = R(0), ma = R(0);
        for (int m = 0; m < kStages; ++m) {
          const R km = k[m * ns + tid];
          sa = fmaR(R(c_sol[m]), km, sa);
          ea = fmaR(R(c_err[m]), km, ea);
          ma = fmaR(R(c_mid[m]), km, ma);
        }
        ynew = dt * sa + y[tid];
        ymid = y[tid] + dt * ma;
        const R tol = atol + rtol * fmax(absR(y[tid]), absR(ynew));
        q = dt * ea / tol;
        q *= q;
      }
      const R ratio = block_sum(q, s.red + 4) / R(ns);
      const bool ok = ratio <= R(1);
      if (!(ratio == ratio) || isinf(ratio)) { if (status == ENO_OK) status = ENO_NONFINITE; }
      if (ok) {
        if (tid < ns) {
          const R y0e = y[tid], dy0 = k[tid], dy1 = k[6 * ns + tid];
          coef[0 * ns + tid] = R(-2) * dt * dy0 + R(2) * dt * dy1 - R(8) * y0e - R(8) * ynew + R(16) * ymid;
          coef[1 * ns + tid] = R(5) * dt * dy0 - R(3) * dt * dy1 + R(18) * y0e + R(14) * ynew - R(32) * ymid;
          coef[2 * ns + tid] = R(-4) * dt * dy0 + dt * dy1 - R(11) * y0e - R(5) * ynew + R(16) * ymid;
          coef[3 * ns + tid] = dt * dy0;
          coef[4 * ns + tid] = y0e;
          y[tid] = ynew;
          k[tid] = dy1;   // f of the next step (lib/ode.py:133: f1 = k[-1])
        }
        last_t = t;
        t = t + dt;
        ++n_acc;
      }
      dt = next_dt(dt, ratio);
      ++i;
      __syncthreads();
    }
    if (t < target && status == ENO_OK) status = (mxstep > 0 && i >= mxstep) ? ENO_MXSTEP : ENO_DT_UNDERFLOW;
    n_iters += i;
    const R x = (target - last_t) / (t - last_t);
    if (tid < ns) {
      R v = R(0) * x;
      for (int c = 0; c < 5; ++c) v = v * x + coef[c * ns + tid];
      if (tid < d0) zs_out[((size_t)traj * T + it) * d0 + tid] = v;
      else if (rs_out) rs_out[(size_t)traj * T + it] = v;
    }
  }
  (void)y1v;
  if (tid == 0) {
    nfe_out[traj] = R(2) + R(6) * R(n_iters);
    if (iters_out) { iters_out[2 * traj] = n_iters; iters_out[2 * traj + 1] = n_acc; }
    if (status_out) status_out[traj] = status;
  }
}

// ------------------------------------------------------------------------------------------------------------
// One pass over the params_bar block of a trajectory, formed from stage factors:
//   k_j[e] = sum_c A_jc[i] * G_jc[o]   (e = element (i, o) of a layer; bias row: i = din, A_j0 = 1)
// MODE 0 (iteration): cand = p0 + dt * sum_j c_sol[j] k_j;  q = dt * sum_j c_err[j] k_j / tol;  returns sum q^2
// MODE 1 (init 0):    returns sum (p0 / scale)^2, and in *aux sum (k_slot0 / scale)^2
// MODE 2 (init 1):    returns sum ((k_slot1 - k_slot0) / scale)^2
// MODE 3 (output):    dst = p0 + dt * sum_j w[j] k_j   (dense output at the segment end, linear form of the quartic)
// ------------------------------------------------------------------------------------------------------------
template <typename R, int MODE>
__device__ __forceinline__ R pb_pass(const Net& n, const Sm<R>& s, const R* __restrict__ fact /*[7][FS] of this trajectory*/,
                                     const int* slot, const R* __restrict__ p0, R* __restrict__ dst, R dt, R rtol, R atol,
                                     const R* __restrict__ w /*MODE 3: [7] in shared memory*/, R* aux) {
  constexpr int NJ = (MODE == 0 || MODE == 3) ? kStages : (MODE == 1 ? 1 : 2);
  constexpr int NCc = 3 * NJ;
  R* SA = s.A;   // the work area is free between right-hand sides: [NCc][din + 1] then [NCc][dout]
  const int tid = threadIdx.x, o = tid & 63, grp = tid >> 6;
  R part = R(0), part2 = R(0);
  for (int l = 0; l < n.L; ++l) {
    const int din = n.d[l], dout = n.d[l + 1];
    const int pa = din + 1;
    R* SG = SA + NCc * pa;
    __syncthreads();
    for (int e = tid; e < NCc * pa; e += kThreads) {
      const int jc = e / pa, i = e - jc * pa, j = jc / 3, c = jc - 3 * j;
      SA[e] = (i < din) ? fact[(size_t)slot[j] * n.FS + c * n.IN + n.io[l] + i] : (c == 0 ? R(1) : R(0));
    }
    for (int e = tid; e < NCc * dout; e += kThreads) {
      const int jc = e / dout, oo_ = e - jc * dout, j = jc / 3, c = jc - 3 * j;
      SG[e] = fact[(size_t)slot[j] * n.FS + 3 * n.IN + c * n.OUT + n.oo[l] + oo_];
    }
    __syncthreads();
    if (o < dout) {
      R g1[NCc], g2[NCc];
#pragma unroll
      for (int jc = 0; jc < NCc; ++jc) {
        const R g = SG[jc * dout + o];
        const int j = jc / 3;
        if (MODE == 0) { g1[jc] = R(c_sol[j]) * g; g2[jc] = R(c_err[j]) * g; }
        else if (MODE == 1) { g1[jc] = g; g2[jc] = R(0); }
        else if (MODE == 2) { g1[jc] = (j == 1) ? g : -g; g2[jc] = R(0); }
        else { g1[jc] = w[j] * g; g2[jc] = R(0); }
      }
      const R* pl = p0 + n.po[l];
      R* dl = dst ? dst + n.po[l] : nullptr;
      for (int i = grp; i <= din; i += 2) {
        const int pe = (i == din) ? o : dout + i * dout + o;
        R v1 = R(0), v2 = R(0);
#pragma unroll
        for (int jc = 0; jc < NCc; ++jc) {
          const R a = SA[jc * pa + i];
          v1 = fmaR(a, g1[jc], v1);
          if (MODE == 0) v2 = fmaR(a, g2[jc], v2);
        }
        const R p = pl[pe];
        if (MODE == 0) {
          const R pn = dt * v1 + p;
          const R tol = atol + rtol * fmax(absR(p), absR(pn));
          const R q = dt * v2 / tol;
          part = fmaR(q, q, part);
          dl[pe] = pn;
        } else if (MODE == 1) {
          const R sc = atol + absR(p) * rtol;
          const R q0 = p / sc, q1 = v1 / sc;
          part = fmaR(q0, q0, part);
          part2 = fmaR(q1, q1, part2);
        } else if (MODE == 2) {
          const R sc = atol + absR(p) * rtol;
          const R q = v1 / sc;
          part = fmaR(q, q, part);
        } else {
          dl[pe] = p + dt * v1;
        }
      }
    }
  }
  __syncthreads();
  if (MODE == 3) return R(0);
  const R tot = block_sum(part, s.red + 4);
  if (MODE == 1) *aux = block_sum(part2, s.red + 4);
  return tot;
}

// ------------------------------------------------------------------------------------------------------------
// Backward: one CTA = the whole adjoint of one trajectory (lib/ode.py:1494-1529), segments T-1 .. 1.
// Small state v[e]: [ z (d0) | r | z_bar (d0) | r_bar | t0_bar ]  (nsm = 2 d0 + 3; without aug: [ z | z_bar | t0_bar ]).
// params_bar: ring of two buffers in the workspace; stage factors: 7 records of FS elements in the workspace.
//   zs [n_traj, T, d0]  rs [n_traj, T]  g_z [n_traj, T, d0]  g_r [n_traj, T]
//   z0_bar_out [n_traj, d0]  r0_bar_out [n_traj]  ts_bar_out [n_traj, T]  pb_out [n_traj, P]
//   iters_out [n_traj, T-1, 2] or NULL (per segment, in the order the segments run: i = T-1 .. 1)
// ------------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kThreads, 2)
k_traj_bwd(const __grid_constant__ Net n, const R* __restrict__ zs, const R* __restrict__ rs, const R* __restrict__ ts, int T,
           const R* __restrict__ params, const R* __restrict__ g_z, const R* __restrict__ g_r, R rtol, R atol, int mxstep,
           R* __restrict__ ws_pb /*[n_traj][2][P]*/, R* __restrict__ ws_fact /*[n_traj][7][FS]*/, R* __restrict__ z0_bar_out,
           R* __restrict__ r0_bar_out, R* __restrict__ ts_bar_out, R* __restrict__ pb_out, int* __restrict__ iters_out,
           int* __restrict__ status_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm<R> s;
  carve(n, smem_raw, s);
  const int tid = threadIdx.x, traj = blockIdx.x;
  const int d0 = n.d[0], na = n.aug ? 1 : 0;
  const int nsm = 2 * (d0 + na) + 1;
  const int oZB = d0 + na, oRB = 2 * d0 + na, oT0 = 2 * (d0 + na);   // offsets of z_bar, r_bar, t0_bar in v
  const R n_state = R(nsm + n.P);
  // solver vectors: v[nsm], k[7][nsm], wq[8]
  R* v = s.st;
  R* k = v + nsm;
  R* wq = k + kStages * nsm;
  __shared__ int slot[kStages];
  R* pb[2] = {ws_pb + (size_t)traj * 2 * n.P, ws_pb + (size_t)traj * 2 * n.P + n.P};
  R* fact = ws_fact + (size_t)traj * kStages * n.FS;
  load_weights(n, s, params);
  for (int e = tid; e < n.P; e += kThreads) pb[0][e] = R(0);
  if (tid < kStages) slot[tid] = tid;
  if (tid < nsm) {
    R x = R(0);
    if (tid >= oZB && tid < oZB + d0) x = g_z[((size_t)traj * T + (T - 1)) * d0 + (tid - oZB)];
    else if (na && tid == oRB) x = g_r[(size_t)traj * T + (T - 1)];
    v[tid] = x;
  }
  int cur = 0, status = ENO_OK;
  __syncthreads();

  // one right-hand side of the adjoint system at (zin, zbin, r_bar): stage derivative j of the small state and the
  // factor record of the params_bar block
  auto rhs = [&](int j, R rb) {
    rhs_forward(n, s, s.zin);
    const R rdot = s.red[0];
    rhs_reverse(n, s, s.zbin, rb);
    if (tid < nsm) {
      R x = R(0);
      if (tid < d0) x = -s.F[tid];
      else if (na && tid == d0) x = -rdot;
      else if (tid >= oZB && tid < oZB + d0) x = s.zbdot[tid - oZB];
      k[j * nsm + tid] = x;
    }
    R* dstf = fact + (size_t)slot[j] * n.FS;
    for (int e = tid; e < n.FS; e += kThreads) dstf[e] = s.A[e];
    __syncthreads();
  };

  for (int seg = T - 1; seg >= 1; --seg) {
    // state at the output time: y from the forward trajectory (lib/ode.py:1519)
    if (tid < d0) v[tid] = zs[((size_t)traj * T + seg) * d0 + tid];
    else if (na && tid == d0) v[tid] = rs[(size_t)traj * T + seg];
    __syncthreads();
    const R rb = na ? v[oRB] : R(0);
    if (tid < d0) { s.zin[tid] = v[tid]; s.zbin[tid] = v[oZB + tid]; }
    __syncthreads();
    rhs(0, rb);
    // t_bar = func(ys[i], ts[i]) . g[i];  t0_bar -= t_bar   (lib/ode.py:1513-1514)
    {
      R q = R(0);
      if (tid < d0) q = -k[tid] * g_z[((size_t)traj * T + seg) * d0 + tid];
      else if (na && tid == d0) q = -k[tid] * g_r[(size_t)traj * T + seg];
      const R t_bar = block_sum(q, s.red + 4);
      if (tid == 0) { ts_bar_out[(size_t)traj * T + seg] = t_bar; v[oT0] = v[oT0] - t_bar; }
      __syncthreads();
    }
    // initial_step_size over the whole adjoint state (lib/ode.py:86-104)
    R dt;
    {
      R q0 = R(0), q1 = R(0);
      if (tid < nsm) { const R sc = atol + absR(v[tid]) * rtol; q0 = v[tid] / sc; q1 = k[tid] / sc; q0 *= q0; q1 *= q1; }
      R s0 = block_sum(q0, s.red + 4), s1 = block_sum(q1, s.red + 4);
      R p1s = R(0);
      s0 += pb_pass<R, 1>(n, s, fact, slot, pb[cur], nullptr, R(0), rtol, atol, nullptr, &p1s);
      s1 += p1s;
      const R d0n = sqrtR(s0), d1n = sqrtR(s1);
      const R h0 = init_h0(d0n, d1n);
      if (tid < d0) { s.zin[tid] = v[tid] + h0 * k[tid]; s.zbin[tid] = v[oZB + tid] + h0 * k[oZB + tid]; }
      __syncthreads();
      rhs(1, rb);
      R q = R(0);
      if (tid < nsm) { const R sc = atol + absR(v[tid]) * rtol; q = (k[nsm + tid] - k[tid]) / sc; q *= q; }
      R s2 = block_sum(q, s.red + 4);
      s2 += pb_pass<R, 2>(n, s, fact, slot, pb[cur], nullptr, R(0), rtol, atol, nullptr, nullptr);
      dt = init_dt(h0, d1n, sqrtR(s2) / h0);
    }
    R t = -ts[seg], last_t = t;
    const R target = -ts[seg - 1];
    int i = 0, acc_n = 0;
    bool have_out = false;
    while (t < target && (mxstep <= 0 || i < mxstep) && dt > R(0)) {
      for (int j = 1; j < kStages; ++j) {
        if (tid < d0) {
          R az = R(0), ab = R(0);
          for (int m = 0; m < j; ++m) {
            const R b = R(c_beta[j - 1][m]);
            az = fmaR(b, k[m * nsm + tid], az);
            ab = fmaR(b, k[m * nsm + oZB + tid], ab);
          }
          s.zin[tid] = v[tid] + dt * az;
          s.zbin[tid] = v[oZB + tid] + dt * ab;
        }
        __syncthreads();
        rhs(j, rb);
      }
      R q = R(0), ynew = R(0), ymid = R(0);
      if (tid < nsm) {
        R sa = R(0), ea = R(0), ma = R(0);
        for (int m = 0; m < kStages; ++m) {
          const R km = k[m * nsm + tid];
          sa = fmaR(R(c_sol[m]), km, sa);
          ea = fmaR(R(c_err[m]), km, ea);
          ma = fmaR(R(c_mid[m]), km, ma);
        }
        ynew = dt * sa + v[tid];
        ymid = v[tid] + dt * ma;
        const R tol = atol + rtol * fmax(absR(v[tid]), absR(ynew));
        q = dt * ea / tol;
        q *= q;
      }
      R tot = block_sum(q, s.red + 4);
      tot += pb_pass<R, 0>(n, s, fact, slot, pb[cur], pb[cur ^ 1], dt, rtol, atol, nullptr, nullptr);
      const R ratio = tot / n_state;
      const bool ok = ratio <= R(1);
      if (!(ratio == ratio) || isinf(ratio)) { if (status == ENO_OK) status = ENO_NONFINITE; }
      if (ok) {
        const R t1 = t + dt;
        if (!(t1 < target)) {
          // the step crosses the segment end: the result of this odeint call is the dense output at the target
          // (lib/ode.py:849-850); nothing after it is used, so only the output is formed
          const R x = (target - t) / (t1 - t);
          if (tid < nsm) v[tid] = quartic_at(v[tid], ynew, ymid, k[tid], k[6 * nsm + tid], dt, x);
          if (tid == 0) {
            const R x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
            const R w1 = R(-8) * x4 + R(14) * x3 - R(5) * x2, wm = R(16) * x4 - R(32) * x3 + R(16) * x2;
            const R wd0 = R(-2) * x4 + R(5) * x3 - R(4) * x2 + x, wd1 = R(2) * x4 - R(3) * x3 + x2;
            for (int m = 0; m < kStages; ++m) wq[m] = w1 * R(c_sol[m]) + wm * R(c_mid[m]) + (m == 0 ? wd0 : R(0)) + (m == 6 ? wd1 : R(0));
          }
          __syncthreads();
          R* dst = (seg == 1) ? pb_out + (size_t)traj * n.P : pb[cur ^ 1];
          pb_pass<R, 3>(n, s, fact, slot, pb[cur], dst, dt, rtol, atol, wq, nullptr);
          have_out = true;
        } else {
          if (tid < nsm) { v[tid] = ynew; k[tid] = k[6 * nsm + tid]; }
          if (tid == 0) { const int s0_ = slot[0]; slot[0] = slot[6]; slot[6] = s0_; }
        }
        cur ^= 1;
        last_t = t;
        t = t1;
        ++acc_n;
      }
      dt = next_dt(dt, ratio);
      ++i;
      __syncthreads();
    }
    if (!have_out) {
      // the loop ended without reaching the target (mxstep / dt underflow): the state reached so far is handed on
      if (status == ENO_OK) status = (mxstep > 0 && i >= mxstep) ? ENO_MXSTEP : ENO_DT_UNDERFLOW;
      if (seg == 1) for (int e = tid; e < n.P; e += kThreads) pb_out[(size_t)traj * n.P + e] = pb[cur][e];
    }
    (void)last_t;
    if (iters_out && tid == 0) {
      iters_out[((size_t)traj * (T - 1) + (T - 1 - seg)) * 2] = i;
      iters_out[((size_t)traj * (T - 1) + (T - 1 - seg)) * 2 + 1] = acc_n;
    }
    // y_bar += g[i - 1]   (lib/ode.py:1524)
    __syncthreads();
    if (tid >= oZB && tid < oZB + d0) v[tid] += g_z[((size_t)traj * T + seg - 1) * d0 + (tid - oZB)];
    else if (na && tid == oRB) v[tid] += g_r[(size_t)traj * T + seg - 1];
    __syncthreads();
  }
  if (T < 2) for (int e = tid; e < n.P; e += kThreads) pb_out[(size_t)traj * n.P + e] = R(0);
  if (tid < d0) z0_bar_out[(size_t)traj * d0 + tid] = v[oZB + tid];
  if (tid == 0) {
    if (na && r0_bar_out) r0_bar_out[traj] = v[oRB];
    ts_bar_out[(size_t)traj * T] = v[oT0];
    if (status_out) status_out[traj] = status;
  }
}

// params_bar_out[p] = sum over trajectories in index order (vmap's transpose sums the per-trajectory cotangents)
template <typename R>
__global__ void k_traj_sum(const R* __restrict__ pb, int n_traj, int P, R* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  R acc = R(0);
  for (int t = 0; t < n_traj; ++t) acc += pb[(size_t)t * P + p];
  out[p] = acc;
}

// Unit-test hook: one evaluation of the augmented dynamics and (with_vjp) its vjp per trajectory.
//   z [n, d0]  z_bar [n, d0]  r_bar [n]  ->  f_out [n, d0]  rdot_out [n]  zbdot_out [n, d0]  pbar_out [n, P]
template <typename R>
__global__ void __launch_bounds__(kThreads, 2)
k_traj_rhs(const __grid_constant__ Net n, const R* __restrict__ z, const R* __restrict__ params, const R* __restrict__ z_bar,
           const R* __restrict__ r_bar, int with_vjp, R* __restrict__ f_out, R* __restrict__ rdot_out, R* __restrict__ zbdot_out,
           R* __restrict__ pbar_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm<R> s;
  carve(n, smem_raw, s);
  const int tid = threadIdx.x, traj = blockIdx.x, d0 = n.d[0];
  load_weights(n, s, params);
  if (tid < d0) { s.zin[tid] = z[(size_t)traj * d0 + tid]; s.zbin[tid] = with_vjp ? z_bar[(size_t)traj * d0 + tid] : R(0); }
  __syncthreads();
  rhs_forward(n, s, s.zin);
  if (tid < d0) f_out[(size_t)traj * d0 + tid] = s.F[tid];
  if (tid == 0 && rdot_out) rdot_out[traj] = s.red[0];
  if (!with_vjp) return;
  rhs_reverse(n, s, s.zbin, r_bar ? r_bar[traj] : R(0));
  if (tid < d0) zbdot_out[(size_t)traj * d0 + tid] = s.zbdot[tid];
  R* pb = pbar_out + (size_t)traj * n.P;
  for (int l = 0; l < n.L; ++l) {
    const int din = n.d[l], dout = n.d[l + 1];
    for (int e = tid; e < (din + 1) * dout; e += kThreads) {
      const int i = e / dout, o = e - i * dout;
      R x;
      if (i == din) x = s.GX[n.oo[l] + o];
      else x = s.A[n.io[l] + i] * s.GX[n.oo[l] + o] + s.AU[n.io[l] + i] * s.GU[n.oo[l] + o] + s.AV[n.io[l] + i] * s.GV[n.oo[l] + o];
      pb[n.po[l] + (i == din ? o : dout + i * dout + o)] = x;
    }
  }
}

}  // namespace lat
}  // namespace eno
