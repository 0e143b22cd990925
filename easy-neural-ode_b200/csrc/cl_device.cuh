// "Cluster-resident" formulation of the MLP dynamics and its Taylor / adjoint sweeps.
//
// A thread-block cluster of kC = 5 CTAs owns a tile of R samples.  The state width D is split
// into kC column slices; CTA `rank` keeps, for the whole launch, its slice of BOTH weight
// matrices in shared memory:
//     W1 rows  d0..d0+Dl  ([Dl][H])   - used as  a[n]  = sum_{d local} x[d] W1[d][n]   (partial over d)
//                                       and      sb[d] = sum_n ab[n] W1[d][n]           (local d)
//     W2 cols  d0..d0+Dl  ([H][Dl])   - used as  f[d]  = sum_n h[n] W2[n][d]            (local d)
//                                       and      hb[n] = sum_{d local} g[d] W2[n][d]    (partial over d)
// Every D-vector is distributed (each CTA holds its Dl columns), every H-vector is replicated;
// the only cluster traffic is an all-reduce of R x H partial sums per layer-1 product, done with
// distributed-shared-memory loads after one cluster barrier.  Compared with the streamed
// formulation (mlp_device.cuh) no weight byte crosses L2 after the prologue.
//
// Reference arithmetic: mnist.py:195-222 (MLPDynamics), 105-131 (sol_recursive), lib/ode.py:1498-1504.
#pragma once
#include "gemm_tc.cuh"
#include <cooperative_groups.h>

#include "adj_device.cuh"

namespace eno {
namespace cg = cooperative_groups;

// Optional phase clocks (build with -DENO_PHASES): thread 0 of CTA 0 accumulates clock64() deltas by
// phase kind.  Development aid only; compiled out of the product library.
enum PhaseKind { PH_PROLOGUE = 0, PH_ELEM, PH_MV_DL, PH_MV_H, PH_ALLRED, PH_REDLOC, PH_STORE, PH_SCALAR, PH_COMBINE,
                 PH_WAIT_MV, PH_AR_LOCAL, PH_AR_CSYNC, PH_AR_REMOTE, PH_NKINDS };
#ifdef ENO_PHASES
__device__ long long g_phase_cycles[32];
__device__ long long g_phase_count[32];
// thread 0 of CTA 0 keeps the clocks in SHARED memory (a tick costs a few cycles) and flushes once
__device__ __forceinline__ long long* ph_slots() { __shared__ long long s[72]; return s; }
#define ENO_TICK(kind)                                                        \
  do {                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                \
      long long* s_ = ph_slots();                                             \
      const long long now_ = clock64();                                       \
      s_[kind] += now_ - s_[64];                                              \
      s_[32 + kind] += 1;                                                     \
      s_[64] = clock64();                                                     \
    }                                                                         \
  } while (0)
#define ENO_TICK_START()                                                      \
  do {                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                \
      long long* s_ = ph_slots();                                             \
      for (int i_ = 0; i_ < 64; ++i_) s_[i_] = 0;                             \
      s_[64] = clock64();                                                     \
    }                                                                         \
  } while (0)
#define ENO_TICK_FLUSH()                                                      \
  do {                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                \
      long long* s_ = ph_slots();                                             \
      for (int i_ = 0; i_ < 32; ++i_) { g_phase_cycles[i_] += s_[i_]; g_phase_count[i_] += s_[32 + i_]; } \
    }                                                                         \
  } while (0)
#else
#define ENO_TICK(kind) do {} while (0)
#define ENO_TICK_START() do {} while (0)
#define ENO_TICK_FLUSH() do {} while (0)
#endif

constexpr int kC = 5;          // CTAs per cluster: 25 clusters x 5 = 125 of the 148 SMs at B = 100 (26 such clusters can be
                               // co-resident, tools/probe/cluster_occ.cu; 4-CTA clusters used 100 SMs and were 5 % slower per step)
constexpr int kKG = 14;        // k-groups of a product contracting over the local slice (K = Dl = 49 float4 rows at D = 784)
constexpr int kKGw = 7;        // k-groups of a product contracting over H (25 float4 rows at H = 100)
constexpr int kRG = 1;         // row groups: 1 = every weight word is read from shared memory ONCE per product and
                               // feeds all R samples (the products are shared-memory-wavefront bound)

struct ClGeom {
  int rank;   // CTA rank in cluster
  int d0;     // first global column of this CTA's slice
  int Dl;     // columns in the slice (multiple of 4)
  int Dlp;    // padded slice width: stride of D-vectors and of W2 rows ((Dlp/4) odd -> conflict-free)
  int ldw1;   // stride of W1 rows ((ldw1/4) odd)
  int Hs;     // stride of the all-reduce staging rows (H + 4 scalar slots)
};

__host__ __device__ inline ClGeom cl_geom(int D, int H, int rank) {
  ClGeom g;
  const int c4 = D / 4, q = c4 / kC, rem = c4 % kC;
  const int dl4 = q + (rank < rem ? 1 : 0);
  const int d04 = rank * q + (rank < rem ? rank : rem);
  int dmax4 = q + (rem > 0 ? 1 : 0);
  if ((dmax4 & 1) == 0) dmax4 += 1;
  int h4 = H / 4;
  if ((h4 & 1) == 0) h4 += 1;
  g.rank = rank;
  g.d0 = 4 * d04;
  g.Dl = 4 * dl4;
  g.Dlp = 4 * dmax4;
  g.ldw1 = 4 * h4;
  g.Hs = H + 4;
  return g;
}

struct ClWeights {
  const float* W1;   // smem [Dl][ldw1]
  const float* W2;   // smem [H][Dlp]
  const float* b1;   // smem [H]
  const float* w1t;  // smem [H]
  const float* b2;   // smem [Dl]   this CTA's slice
  const float* w2t;  // smem [Dl]
  int D, H;
};

// ---- bulk asynchronous copies (TMA engine, cp.async.bulk) of the weight slices ------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(mbar)) : "memory");
}

// Copy of this CTA's weight slices into shared memory, once per launch.  The two matrices (157 KB at
// D = 784, H = 100) travel as bulk asynchronous copies issued by one thread and tracked by an mbarrier -
// no register staging, full L2 -> shared-memory bandwidth; the four small vectors use ordinary loads and
// overlap with them.  W1's slice is contiguous in global memory (rows d0..d0+Dl of W1x) when the row stride
// needs no padding; W2's slice is one 4*Dl-byte segment per row.
__device__ __forceinline__ void cl_load_weights(const MlpWeights& w, const ClGeom& g, float* W1s, float* W2s,
                                                float* vec /*[2H + 2Dlp]: b1, w1t, b2 slice, w2t slice*/) {
  const int H = w.H, D = w.D;
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && g.Dl > 0) {
    const unsigned total = 2u * (unsigned)g.Dl * (unsigned)H * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar)), "r"(total) : "memory");
    if (g.ldw1 == H) {
      // one contiguous block; split so that every copy stays well below the 1 MB bulk limit
      const unsigned bytes = (unsigned)g.Dl * (unsigned)H * 4u;
      const unsigned half = (bytes / 2u) & ~15u;
      bulk_g2s(W1s, w.W1x + (size_t)g.d0 * H, half, &mbar);
      bulk_g2s(reinterpret_cast<char*>(W1s) + half, reinterpret_cast<const char*>(w.W1x + (size_t)g.d0 * H) + half,
               bytes - half, &mbar);
    } else {
      for (int d = 0; d < g.Dl; ++d) bulk_g2s(W1s + (size_t)d * g.ldw1, w.W1x + (size_t)(g.d0 + d) * H, (unsigned)H * 4u, &mbar);
    }
    for (int n = 0; n < H; ++n) bulk_g2s(W2s + (size_t)n * g.Dlp, w.W2x + (size_t)n * D + g.d0, (unsigned)g.Dl * 4u, &mbar);
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    vec[i] = __ldg(w.b1 + i);
    vec[H + i] = __ldg(w.w1t + i);
  }
  for (int i = threadIdx.x; i < g.Dlp; i += blockDim.x) {
    vec[2 * H + i] = (i < g.Dl) ? __ldg(w.b2 + g.d0 + i) : 0.f;
    vec[2 * H + g.Dlp + i] = (i < g.Dl) ? __ldg(w.w2t + g.d0 + i) : 0.f;
  }
  if (g.Dl > 0) {
    // every thread waits for the transaction count to drain (phase parity 0: the barrier is used once);
    // the spin is bounded so that a lost copy traps instead of hanging the device
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_addr(&mbar)) : "memory");
    }
    if (!done) __trap();
  }
}

// partial[kg][r][n] = sum_{k in chunk kg} x[r][k] * W[k][n]      (outputs along W's contiguous dim)
// R must be a multiple of kRG.  All sizes are compile-time constants in the specialised kernels, so the
// k-loop unrolls completely and every address is an immediate offset.
template <int R>
__device__ __forceinline__ void mv_n(const float* __restrict__ W, int ldw, int K, int N, const float* __restrict__ x,
                                     int ldx, float* __restrict__ red, int KG) {
  constexpr int TR = R / kRG;
  static_assert(R % kRG == 0, "rows per tile must be a multiple of the row groups");
  const int NC4 = N >> 2;
  const int per = NC4 * KG;
  const int rg = threadIdx.x / per;
  if (rg < kRG) {
    const int rem = threadIdx.x - rg * per;
    const int kg = rem / NC4, c4 = rem - kg * NC4;
    const int K4 = K >> 2;
    const int chunk = (K4 + KG - 1) / KG;
    const bool exact = (chunk * KG == K4);   // folds to true in the specialised build: no tail predicate
    const int k4b = kg * chunk;
    const int r0 = rg * TR;
    const int l4 = ldw >> 2;
    float4 acc[TR];
#pragma unroll
    for (int t = 0; t < TR; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* Wp = reinterpret_cast<const float4*>(W) + (size_t)(4 * k4b) * l4 + c4;
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)r0 * ldx) + k4b;
    const int lx4 = ldx >> 2;
#pragma unroll
    for (int i = 0; i < chunk; ++i) {
      if (exact || k4b + i < K4) {
        const float4 w0 = Wp[(4 * i) * l4], w1 = Wp[(4 * i + 1) * l4], w2 = Wp[(4 * i + 2) * l4], w3 = Wp[(4 * i + 3) * l4];
#pragma unroll
        for (int t = 0; t < TR; ++t) {
          const float4 xv = xp[t * lx4 + i];
          acc[t].x = fmaf(xv.x, w0.x, acc[t].x); acc[t].y = fmaf(xv.x, w0.y, acc[t].y);
          acc[t].z = fmaf(xv.x, w0.z, acc[t].z); acc[t].w = fmaf(xv.x, w0.w, acc[t].w);
          acc[t].x = fmaf(xv.y, w1.x, acc[t].x); acc[t].y = fmaf(xv.y, w1.y, acc[t].y);
          acc[t].z = fmaf(xv.y, w1.z, acc[t].z); acc[t].w = fmaf(xv.y, w1.w, acc[t].w);
          acc[t].x = fmaf(xv.z, w2.x, acc[t].x); acc[t].y = fmaf(xv.z, w2.y, acc[t].y);
          acc[t].z = fmaf(xv.z, w2.z, acc[t].z); acc[t].w = fmaf(xv.z, w2.w, acc[t].w);
          acc[t].x = fmaf(xv.w, w3.x, acc[t].x); acc[t].y = fmaf(xv.w, w3.y, acc[t].y);
          acc[t].z = fmaf(xv.w, w3.z, acc[t].z); acc[t].w = fmaf(xv.w, w3.w, acc[t].w);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TR; ++t) reinterpret_cast<float4*>(red + (size_t)(kg * R + r0 + t) * N)[c4] = acc[t];
  }
}

// partial[kg][r][n] = sum_{k in chunk kg} x[r][k] * W[n][k]      (contraction along W's contiguous dim)
// thread q owns outputs n = q, q + N/4, q + 2N/4, q + 3N/4 (row stride with (ldw/4) odd: conflict-free)
template <int R>
__device__ __forceinline__ void mv_k(const float* __restrict__ W, int ldw, int N, int K, const float* __restrict__ x,
                                     int ldx, float* __restrict__ red, int KG) {
  constexpr int TR = R / kRG;
  const int NQ = N >> 2;
  if (NQ == 0) return;   // a CTA whose column slice is empty (D/4 < kC) has nothing to produce
  const int per = NQ * KG;
  const int rg = threadIdx.x / per;
  if (rg < kRG) {
    const int rem = threadIdx.x - rg * per;
    const int kg = rem / NQ, q = rem - kg * NQ;
    const int K4 = K >> 2;
    const int chunk = (K4 + KG - 1) / KG;
    const bool exact = (chunk * KG == K4);
    const int k4b = kg * chunk;
    const int r0 = rg * TR;
    float acc[TR][4];
#pragma unroll
    for (int t = 0; t < TR; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
    const int rs = NQ * (ldw >> 2);
    const float4* W0 = reinterpret_cast<const float4*>(W + (size_t)q * ldw) + k4b;
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)r0 * ldx) + k4b;
    const int lx4 = ldx >> 2;
#pragma unroll
    for (int j = 0; j < chunk; ++j) {
      if (exact || k4b + j < K4) {
        float4 wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = W0[i * rs + j];
#pragma unroll
        for (int t = 0; t < TR; ++t) {
          const float4 xv = xp[t * lx4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[t][i] = fmaf(xv.x, wv[i].x, acc[t][i]);
            acc[t][i] = fmaf(xv.y, wv[i].y, acc[t][i]);
            acc[t][i] = fmaf(xv.z, wv[i].z, acc[t][i]);
            acc[t][i] = fmaf(xv.w, wv[i].w, acc[t][i]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TR; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) red[(size_t)(kg * R + r0 + t) * N + q + i * NQ] = acc[t][i];
  }
}

// local (distributed-D) product: reduce the k-group partials and hand (r, n, sum) to the epilogue
template <int R, typename Epi>
__device__ __forceinline__ void reduce_local(const float* red, int KG, int N, Epi epi) {
  __syncthreads();
  ENO_TICK(PH_WAIT_MV);
  for (int idx = threadIdx.x; idx < R * N; idx += blockDim.x) {
    const int r = idx / N, n = idx - r * N;
    float s = 0.f;
    for (int g = 0; g < KG; ++g) s += red[(size_t)(g * R + r) * N + n];
    epi(r, n, s);
  }
  __syncthreads();
  ENO_TICK(PH_REDLOC);
}

// Per-CTA staging area of the cluster all-reduce: two buffers used alternately so that a CTA can
// start filling the next one while slower peers still read the previous one.
struct ClComm {
  float* part[2];   // [R][Hs]   columns [0,H): layer-1 partial sums; [H, H+4): per-sample scalars
  int phase;
};

// All-reduce over the cluster of R x (H + nscal) partial sums.  `red` holds the k-group partials of
// this CTA's layer-1 product; `scal[r*4 + j]` (may be null) holds up to 4 per-sample scalar partials.
// Sum order over CTAs is fixed (rank 0..kC-1), so every CTA obtains bit-identical results.
template <int R, typename Epi, typename EpiS>
__device__ __forceinline__ void allreduce_H(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g, const float* red,
                                            int KG, int H, const float* scal, int nscal, Epi epi, EpiS epis) {
  float* mine = cm.part[cm.phase];
  __syncthreads();
  ENO_TICK(PH_WAIT_MV);
  for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
    const int r = idx / H, n = idx - r * H;
    float s = 0.f;
    for (int k = 0; k < KG; ++k) s += red[(size_t)(k * R + r) * H + n];
    mine[r * g.Hs + n] = s;
  }
  if (scal && threadIdx.x < R * nscal) {
    const int r = threadIdx.x / nscal, j = threadIdx.x - r * nscal;
    mine[r * g.Hs + H + j] = scal[r * 4 + j];
  }
  ENO_TICK(PH_AR_LOCAL);
  cluster.sync();
  ENO_TICK(PH_AR_CSYNC);
  const float* pk[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) pk[k] = cluster.map_shared_rank(mine, k);
  for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
    const int r = idx / H, n = idx - r * H;
    const int o = r * g.Hs + n;
    float v[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) v[k] = pk[k][o];   // all remote loads in flight before the adds
    float s = v[0];
#pragma unroll
    for (int k = 1; k < kC; ++k) s += v[k];
    epi(r, n, s);
  }
  if (scal && threadIdx.x < R * nscal) {
    const int r = threadIdx.x / nscal, j = threadIdx.x - r * nscal;
    const int o = r * g.Hs + H + j;
    float s = pk[0][o];
#pragma unroll
    for (int k = 1; k < kC; ++k) s += pk[k][o];
    epis(r, j, s);
  }
  cm.phase ^= 1;
  ENO_TICK(PH_AR_REMOTE);
  __syncthreads();
  ENO_TICK(PH_ALLRED);
}

// shared-memory working set (distributed D-vectors have stride Dlp, H-vectors stride H)
struct ClScratch {
  float *s, *u;                    // [R][Dlp]
  float *f, *y1, *y2;              // [R][Dlp]
  float *gf, *gy1, *zacc;          // [R][Dlp]
  float *h, *a1, *h1, *a2, *h2;    // [R][H]
  float *ga, *ga1, *ga2;           // [R][H]
  float* red;
  float* scal;                     // [R][4]
  float* rdot;                     // [R]
};

// f = MLP(z, t) for the CTA's column slice; leaves s (local slice) and h (replicated).
template <int R>
__device__ __forceinline__ void cl_primal(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g, const ClWeights& w,
                                          float t, const float* z, float* f, const ClScratch& sc) {
  const int H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
    const int d = i % Dlp;
    sc.s[i] = (d < Dl) ? sigmoidf_(z[i]) : 0.f;
  }
  __syncthreads();
  ENO_TICK(PH_ELEM);
  mv_n<R>(w.W1, g.ldw1, Dl, H, sc.s, Dlp, sc.red, kKG);
  ENO_TICK(PH_MV_DL);
  allreduce_H<R>(cluster, cm, g, sc.red, kKG, H, nullptr, 0,
                 [&](int r, int n, float sum) { sc.h[r * H + n] = sigmoidf_(sum + t * w.w1t[n] + w.b1[n]); },
                 [&](int, int, float) {});
  mv_n<R>(w.W2, Dlp, H, Dl, sc.h, H, sc.red, kKGw);
  ENO_TICK(PH_MV_H);
  reduce_local<R>(sc.red, kKGw, Dl, [&](int r, int d, float sum) {
    f[r * Dlp + d] = sum + t * w.w2t[d] + w.b2[d];
  });
}

}  // namespace eno

namespace eno {

// tf32 split of gemm_tc.cuh (round-to-nearest pieces)
__device__ __forceinline__ void cl_split_tf32(float x, float& hi, float& lo) { tc::split_tf32(x, hi, lo); }

// store this CTA's column slice of R distributed rows of factor `which` (FM_S / FM_G), Taylor pass `pass`:
// row-major into the global [3][B][D] block, or - tensor-core sinks - transposed and hi / lo split: the R samples of
// the tile are R consecutive k of row (d0 + d), one 16-byte store per plane (R == 4; rows past B store zeros)
template <int R, class FD>
__device__ __forceinline__ void cl_store_D(const FD& sl, int which, int pass, const float* src, const ClGeom& g, int D,
                                           int row0, int B) {
  if constexpr (FD::kTc) {
    static_assert(R == 4, "tensor-core factor sinks store one float4 of samples per row");
    const int ldk = sl.tc->ldk;
    const unsigned xplane = sl.tc->xplane;
    float* base = sl.tX + ((size_t)which * sl.tc->dpad + g.d0) * ldk + pass * sl.tc->Bk + row0;
    const int nv = B - row0;   // valid rows of this tile
    for (int d = threadIdx.x; d < g.Dl; d += blockDim.x) {
      float4 h, l;
      cl_split_tf32(src[d], h.x, l.x);
      cl_split_tf32(nv > 1 ? src[g.Dlp + d] : 0.f, h.y, l.y);
      cl_split_tf32(nv > 2 ? src[2 * g.Dlp + d] : 0.f, h.z, l.z);
      cl_split_tf32(nv > 3 ? src[3 * g.Dlp + d] : 0.f, h.w, l.w);
      float* dst = base + (size_t)d * ldk;
      *reinterpret_cast<float4*>(dst) = h;
      *reinterpret_cast<float4*>(dst + xplane) = l;
    }
  } else {
    float* dst = (which == FM_S ? sl.rm->S : sl.rm->G) + (size_t)pass * B * D;
    for (int i = threadIdx.x; i < R * g.Dlp; i += blockDim.x) {
      const int r = i / g.Dlp, d = i - r * g.Dlp;
      if (d < g.Dl && row0 + r < B) dst[(size_t)(row0 + r) * D + g.d0 + d] = src[i];
    }
  }
}

// store replicated H-rows of factor `which` (FM_A / FM_HH) (rank 0 only)
template <int R, class FD>
__device__ __forceinline__ void cl_store_H(const FD& sl, int which, int pass, const float* src, const ClGeom& g, int H,
                                           int row0, int B) {
  if (g.rank != 0) return;
  if constexpr (FD::kTc) {
    const int ldk = sl.tc->ldk;
    const unsigned yplane = sl.tc->yplane;
    float* base = sl.tY + (size_t)which * sl.tc->hpad * ldk + pass * sl.tc->Bk + row0;
    const int nv = B - row0;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      float4 hi, lo;
      cl_split_tf32(src[h], hi.x, lo.x);
      cl_split_tf32(nv > 1 ? src[H + h] : 0.f, hi.y, lo.y);
      cl_split_tf32(nv > 2 ? src[2 * H + h] : 0.f, hi.z, lo.z);
      cl_split_tf32(nv > 3 ? src[3 * H + h] : 0.f, hi.w, lo.w);
      float* dst = base + (size_t)h * ldk;
      *reinterpret_cast<float4*>(dst) = hi;
      *reinterpret_cast<float4*>(dst + yplane) = lo;
    }
  } else {
    float* dst = (which == FM_A ? sl.rm->A : sl.rm->Hh) + (size_t)pass * B * H;
    for (int i = threadIdx.x; i < R * H; i += blockDim.x)
      if (row0 + i / H < B) dst[(size_t)row0 * H + i] = src[i];
  }
}

// per-sample partial of sum_d v[d]^2 over the local slice -> scal[r*4 + slot]
template <int R>
__device__ __forceinline__ void cl_sumsq_local(const float* v, const ClGeom& g, float* scal, int slot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    double acc = 0.0;
    for (int d = lane; d < g.Dl; d += 32) {
      const float x = v[r * g.Dlp + d];
      acc += (double)(x * x);
    }
    acc = warp_sum(acc);
    if (lane == 0) scal[r * 4 + slot] = (float)acc;
  }
}

// Forward Taylor passes (aug_forward of adj_device.cuh, distributed).  On return: f, y1, y2 local slices,
// h, a1, h1, a2, h2 replicated, scal[r*4+0] = local partial of sum_d yK^2 (K >= 2).
template <int R, int K, class FD>
__device__ __forceinline__ void cl_aug_forward(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g,
                                               const ClWeights& w, float t, const float* z, const ClScratch& m,
                                               const FD* slot, int row0, int B) {
  const int D = w.D, H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  cl_primal<R>(cluster, cm, g, w, t, z, m.f, m);
  if (slot) {
    cl_store_D<R>(*slot, FM_S, 0, m.s, g, D, row0, B);
    ENO_TICK(PH_STORE);
    cl_store_H<R>(*slot, FM_HH, 0, m.h, g, H, row0, B);
    ENO_TICK(PH_STORE);
  }
  if (K >= 2) {
    for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
      const float s = m.s[i];
      m.u[i] = ((i % Dlp) < Dl) ? s * (1.f - s) * m.f[i] : 0.f;
    }
    __syncthreads();
    if (slot) cl_store_D<R>(*slot, FM_S, 1, m.u, g, D, row0, B);
    ENO_TICK(PH_STORE);
    mv_n<R>(w.W1, g.ldw1, Dl, H, m.u, Dlp, m.red, kKG);
    ENO_TICK(PH_MV_DL);
    allreduce_H<R>(cluster, cm, g, m.red, kKG, H, nullptr, 0,
                   [&](int r, int n, float sum) {
                     const float h = m.h[r * H + n];
                     const float a1 = sum + w.w1t[n];
                     m.a1[r * H + n] = a1;
                     m.h1[r * H + n] = h * (1.f - h) * a1;
                   },
                   [&](int, int, float) {});
    if (slot) cl_store_H<R>(*slot, FM_HH, 1, m.h1, g, H, row0, B);
    ENO_TICK(PH_STORE);
    mv_n<R>(w.W2, Dlp, H, Dl, m.h1, H, m.red, kKGw);
    ENO_TICK(PH_MV_H);
    reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sum) { m.y1[r * Dlp + d] = sum + w.w2t[d]; });
  }
  if (K >= 3) {
    for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
      const float s = m.s[i];
      const float sp = s * (1.f - s);
      const float fi = m.f[i];
      m.u[i] = ((i % Dlp) < Dl) ? sp * (1.f - 2.f * s) * fi * fi + sp * m.y1[i] : 0.f;
    }
    __syncthreads();
    if (slot) cl_store_D<R>(*slot, FM_S, 2, m.u, g, D, row0, B);
    ENO_TICK(PH_STORE);
    mv_n<R>(w.W1, g.ldw1, Dl, H, m.u, Dlp, m.red, kKG);
    ENO_TICK(PH_MV_DL);
    allreduce_H<R>(cluster, cm, g, m.red, kKG, H, nullptr, 0,
                   [&](int r, int n, float sum) {
                     const float h = m.h[r * H + n];
                     const float hp = h * (1.f - h);
                     const float a1 = m.a1[r * H + n];
                     m.a2[r * H + n] = sum;
                     m.h2[r * H + n] = hp * (1.f - 2.f * h) * a1 * a1 + hp * sum;
                   },
                   [&](int, int, float) {});
    if (slot) cl_store_H<R>(*slot, FM_HH, 2, m.h2, g, H, row0, B);
    ENO_TICK(PH_STORE);
    mv_n<R>(w.W2, Dlp, H, Dl, m.h2, H, m.red, kKGw);
    ENO_TICK(PH_MV_H);
    reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sum) { m.y2[r * Dlp + d] = sum; });
  }
  if (K >= 2) {
    cl_sumsq_local<R>((K >= 3) ? m.y2 : m.y1, g, m.scal, 0);
    __syncthreads();
  }
}

// Reverse sweep (aug_reverse, distributed).  zb: cotangent of f (local slice), rb[r]: cotangent of r_dot.
// On return: zacc = vjp wrt z (local slice), rdot[r] = r_dot (all ranks), factors + per-CTA t_bar pieces stored.
template <int R, int K, class FD>
__device__ __forceinline__ void cl_aug_reverse(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g,
                                               const ClWeights& w, const float* zb, const float* rb,
                                               const ClScratch& m, const FD& slot, int row0, int B) {
  const int D = w.D, H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  const float invD2 = 2.0f / (float)D;
  const float invD = 1.0f / (float)D;
  auto take_rdot = [&](int r, int, float v) { m.rdot[r] = v * invD; };

  if (K >= 3) {
    for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) m.u[i] = rb[i / Dlp] * invD2 * m.y2[i];
    __syncthreads();
    cl_store_D<R>(slot, FM_G, 2, m.u, g, D, row0, B);
    ENO_TICK(PH_STORE);
    mv_k<R>(w.W2, Dlp, H, Dl, m.u, Dlp, m.red, kKG);
    ENO_TICK(PH_MV_DL);
    allreduce_H<R>(cluster, cm, g, m.red, kKG, H, m.scal, 1,
                   [&](int r, int n, float hb2) {
                     const float h = m.h[r * H + n];
                     const float hp = h * (1.f - h);
                     const float hpp = hp * (1.f - 2.f * h);
                     const float hppp = hp * (1.f - 6.f * h + 6.f * h * h);
                     const float a1 = m.a1[r * H + n], a2 = m.a2[r * H + n];
                     m.ga2[r * H + n] = hb2 * hp;
                     m.ga1[r * H + n] = hb2 * 2.f * hpp * a1;
                     m.ga[r * H + n] = hb2 * (hppp * a1 * a1 + hpp * a2);
                   },
                   take_rdot);
    cl_store_H<R>(slot, FM_A, 2, m.ga2, g, H, row0, B);
    ENO_TICK(PH_STORE);
    mv_k<R>(w.W1, g.ldw1, Dl, H, m.ga2, H, m.red, kKGw);
    ENO_TICK(PH_MV_H);
    reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sb2) {
      const int i = r * Dlp + d;
      const float s = m.s[i];
      const float sp = s * (1.f - s);
      const float spp = sp * (1.f - 2.f * s);
      const float sppp = sp * (1.f - 6.f * s + 6.f * s * s);
      const float fi = m.f[i];
      m.zacc[i] = sb2 * (sppp * fi * fi + spp * m.y1[i]);
      m.gf[i] = sb2 * 2.f * spp * fi;
      m.gy1[i] = sb2 * sp;
    });
  } else if (K == 2) {
    for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
      m.gy1[i] = rb[i / Dlp] * invD2 * m.y1[i];
      m.gf[i] = 0.f;
      m.zacc[i] = 0.f;
    }
    for (int i = threadIdx.x; i < R * H; i += blockDim.x) { m.ga[i] = 0.f; m.ga1[i] = 0.f; }
    __syncthreads();
  }
  if (K >= 2) {
    cl_store_D<R>(slot, FM_G, 1, m.gy1, g, D, row0, B);
    ENO_TICK(PH_STORE);
    mv_k<R>(w.W2, Dlp, H, Dl, m.gy1, Dlp, m.red, kKG);
    ENO_TICK(PH_MV_DL);
    allreduce_H<R>(cluster, cm, g, m.red, kKG, H, (K == 2) ? m.scal : nullptr, 1,
                   [&](int r, int n, float hb1) {
                     const float h = m.h[r * H + n];
                     const float hp = h * (1.f - h);
                     const float hpp = hp * (1.f - 2.f * h);
                     m.ga[r * H + n] += hb1 * hpp * m.a1[r * H + n];
                     m.ga1[r * H + n] += hb1 * hp;
                   },
                   take_rdot);
    cl_store_H<R>(slot, FM_A, 1, m.ga1, g, H, row0, B);
    ENO_TICK(PH_STORE);
    mv_k<R>(w.W1, g.ldw1, Dl, H, m.ga1, H, m.red, kKGw);
    ENO_TICK(PH_MV_H);
    reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sb1) {
      const int i = r * Dlp + d;
      const float s = m.s[i];
      const float sp = s * (1.f - s);
      const float spp = sp * (1.f - 2.f * s);
      m.zacc[i] += sb1 * spp * m.f[i];
      m.gf[i] += sb1 * sp + zb[i];
    });
  } else {
    for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) { m.gf[i] = zb[i]; m.zacc[i] = 0.f; }
    for (int i = threadIdx.x; i < R * H; i += blockDim.x) m.ga[i] = 0.f;
    if (threadIdx.x < R) m.rdot[threadIdx.x] = 0.f;
    __syncthreads();
  }
  cl_store_D<R>(slot, FM_G, 0, m.gf, g, D, row0, B);
    ENO_TICK(PH_STORE);
  mv_k<R>(w.W2, Dlp, H, Dl, m.gf, Dlp, m.red, kKG);
    ENO_TICK(PH_MV_DL);
  allreduce_H<R>(cluster, cm, g, m.red, kKG, H, nullptr, 0,
                 [&](int r, int n, float hb) {
                   const float h = m.h[r * H + n];
                   m.ga[r * H + n] += hb * h * (1.f - h);
                 },
                 [&](int, int, float) {});
  cl_store_H<R>(slot, FM_A, 0, m.ga, g, H, row0, B);
    ENO_TICK(PH_STORE);
  mv_k<R>(w.W1, g.ldw1, Dl, H, m.ga, H, m.red, kKGw);
    ENO_TICK(PH_MV_H);
  reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sb) {
    const int i = r * Dlp + d;
    const float s = m.s[i];
    m.zacc[i] += sb * s * (1.f - s);
  });
  // per-CTA piece of vjp wrt t: f_bar . w2t over the local slice (+ a_bar . w1t on rank 0)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
      double acc = 0.0;
      for (int d = lane; d < Dl; d += 32) acc += (double)(m.gf[r * Dlp + d] * w.w2t[d]);
      if (g.rank == 0)
        for (int n = lane; n < H; n += 32) acc += (double)(m.ga[r * H + n] * w.w1t[n]);
      acc = warp_sum(acc);
      if (lane == 0 && row0 + r < B) slot.rm->tbar[(size_t)(row0 + r) * kC + g.rank] = (float)acc;
    }
  }
  __syncthreads();
  ENO_TICK(PH_SCALAR);
}


// ------------------------------------------------------------------------------------------------
// "fin" augmented dynamics (mnist.py:294-313 with --reg_type fin; the same structure as the FFJORD
// Hutchinson term of ffjord_tabular.py:250-268): next to f the RHS returns
//     dfro = mean_d (J eps)^2        (J = df/dz, a forward-mode product with the fixed probe eps)
//     dkin = mean_d f^2
// The tangent pass is the first-order Taylor pass above with z' = eps instead of f and WITHOUT the time
// tangent (jvp wrt y only: no w1t / w2t terms).  Its reverse sweep additionally yields the cotangent of
// eps, which is a block of the adjoint state (lib/ode.py:1498-1504 differentiates wrt every arg).
// On return of cl_fin_forward: f, y1 = J eps (local slices), h, a1, h1 replicated,
// scal[r*4+0] = local partial of sum_d (J eps)^2, scal[r*4+1] = of sum_d f^2.
template <int R, class FD>
__device__ __forceinline__ void cl_fin_forward(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g,
                                               const ClWeights& w, float t, const float* z, const float* eps_g,
                                               const ClScratch& m, const FD* slot, int row0, int B) {
  const int D = w.D, H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  cl_primal<R>(cluster, cm, g, w, t, z, m.f, m);
  if (slot) {
    cl_store_D<R>(*slot, FM_S, 0, m.s, g, D, row0, B);
    cl_store_H<R>(*slot, FM_HH, 0, m.h, g, H, row0, B);
  }
  for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
    const int r = i / Dlp, d = i - r * Dlp;
    const float s = m.s[i];
    const bool ok = d < Dl && row0 + r < B;
    m.u[i] = ok ? s * (1.f - s) * __ldg(eps_g + (size_t)(row0 + r) * D + g.d0 + d) : 0.f;
  }
  __syncthreads();
  if (slot) cl_store_D<R>(*slot, FM_S, 1, m.u, g, D, row0, B);
  mv_n<R>(w.W1, g.ldw1, Dl, H, m.u, Dlp, m.red, kKG);
  allreduce_H<R>(cluster, cm, g, m.red, kKG, H, nullptr, 0,
                 [&](int r, int n, float sum) {
                   const float h = m.h[r * H + n];
                   m.a1[r * H + n] = sum;
                   m.h1[r * H + n] = h * (1.f - h) * sum;
                 },
                 [&](int, int, float) {});
  if (slot) cl_store_H<R>(*slot, FM_HH, 1, m.h1, g, H, row0, B);
  mv_n<R>(w.W2, Dlp, H, Dl, m.h1, H, m.red, kKGw);
  reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sum) { m.y1[r * Dlp + d] = sum; });
  cl_sumsq_local<R>(m.y1, g, m.scal, 0);
  cl_sumsq_local<R>(m.f, g, m.scal, 1);
  __syncthreads();
}

// Reverse sweep of the fin dynamics.  zb: cotangent of f (local slice); rb2[r], rb2[R + r]: cotangents of
// dfro, dkin.  On return: zacc = vjp wrt z, ebar = vjp wrt eps (local slices; ebar aliases u),
// rdot2[r] = dfro, rdot2[R + r] = dkin (all ranks), factors + per-CTA t_bar pieces stored.
template <int R, class FD>
__device__ __forceinline__ void cl_fin_reverse(cg::cluster_group& cluster, ClComm& cm, const ClGeom& g,
                                               const ClWeights& w, const float* zb, const float* rb2,
                                               const float* eps_g, const ClScratch& m, float* rdot2,
                                               const FD& slot, int row0, int B) {
  const int D = w.D, H = w.H, Dl = g.Dl, Dlp = g.Dlp;
  const size_t BD = (size_t)B * D, BH = (size_t)B * H;
  const float invD2 = 2.0f / (float)D;
  const float invD = 1.0f / (float)D;
  // cotangent of e = J eps; cotangent of f from its own adjoint and from dkin
  for (int i = threadIdx.x; i < R * Dlp; i += blockDim.x) {
    const int r = i / Dlp;
    m.gy1[i] = rb2[r] * invD2 * m.y1[i];
    m.gf[i] = zb[i] + rb2[R + r] * invD2 * m.f[i];
  }
  __syncthreads();
  cl_store_D<R>(slot, FM_G, 1, m.gy1, g, D, row0, B);
  mv_k<R>(w.W2, Dlp, H, Dl, m.gy1, Dlp, m.red, kKG);
  allreduce_H<R>(cluster, cm, g, m.red, kKG, H, m.scal, 2,
                 [&](int r, int n, float hb1) {
                   const float h = m.h[r * H + n];
                   const float hp = h * (1.f - h);
                   m.ga[r * H + n] = hb1 * hp * (1.f - 2.f * h) * m.a1[r * H + n];
                   m.ga1[r * H + n] = hb1 * hp;
                 },
                 [&](int r, int j, float v) { rdot2[j * R + r] = v * invD; });
  cl_store_H<R>(slot, FM_A, 1, m.ga1, g, H, row0, B);
  mv_k<R>(w.W1, g.ldw1, Dl, H, m.ga1, H, m.red, kKGw);
  // gy1 (= u) is dead once its product and factor store are done: its storage takes eps_bar
  float* ebar = m.gy1;
  reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sb1) {
    const int i = r * Dlp + d;
    const float s = m.s[i];
    const float sp = s * (1.f - s);
    const float ev = (row0 + r < B) ? __ldg(eps_g + (size_t)(row0 + r) * D + g.d0 + d) : 0.f;
    m.zacc[i] = sb1 * sp * (1.f - 2.f * s) * ev;
    ebar[i] = sb1 * sp;
  });
  cl_store_D<R>(slot, FM_G, 0, m.gf, g, D, row0, B);
  mv_k<R>(w.W2, Dlp, H, Dl, m.gf, Dlp, m.red, kKG);
  allreduce_H<R>(cluster, cm, g, m.red, kKG, H, nullptr, 0,
                 [&](int r, int n, float hb) {
                   const float h = m.h[r * H + n];
                   m.ga[r * H + n] += hb * h * (1.f - h);
                 },
                 [&](int, int, float) {});
  cl_store_H<R>(slot, FM_A, 0, m.ga, g, H, row0, B);
  mv_k<R>(w.W1, g.ldw1, Dl, H, m.ga, H, m.red, kKGw);
  reduce_local<R>(m.red, kKGw, Dl, [&](int r, int d, float sb) {
    const int i = r * Dlp + d;
    const float s = m.s[i];
    m.zacc[i] += sb * s * (1.f - s);
  });
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
      double acc = 0.0;
      for (int d = lane; d < Dl; d += 32) acc += (double)(m.gf[r * Dlp + d] * w.w2t[d]);
      if (g.rank == 0)
        for (int n = lane; n < H; n += 32) acc += (double)(m.ga[r * H + n] * w.w1t[n]);
      acc = warp_sum(acc);
      if (lane == 0 && row0 + r < B) slot.rm->tbar[(size_t)(row0 + r) * kC + g.rank] = (float)acc;
    }
  }
  __syncthreads();
}

}  // namespace eno
