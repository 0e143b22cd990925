// fp32-accurate GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
// 3-pass split  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,  x_hi = x rounded to the nearest tf32, x_lo = the
// remainder rounded to the nearest tf32 (split_tf32 below).  Accumulators live in TMEM, operand
// tiles are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a ring of shared-memory stages,
// one elected thread issues the MMAs, four warps drain TMEM with tcgen05.ld and run a fused epilogue functor.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]          both operands K-major ("TN"), given as [2 planes][rows][ld]
//
// Used where the hidden width makes the stage contractions dense GEMMs (BASELINE config C2, ffjord_tabular:
// [1000 x 860] x [860 x 860], ffjord_tabular.py:140-193); the MNIST-sized per-sample products stay on FFMA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace eno {
namespace tc {

constexpr int kBM = 128;        // rows of the accumulator tile = TMEM lanes
constexpr int kBK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;       // K of one tcgen05.mma kind::tf32
constexpr int kAccs = 4;        // TMEM accumulators per tile: 3 x (hi*hi by k-block) + 1 x cross terms
constexpr int kEpiWarps = 16;   // epilogue warps: 4 per TMEM lane quadrant, each owns BN / 4 columns of its 32 rows
constexpr int kEpiChunk = 16;   // columns drained from TMEM per step (per warp: 32 rows x 16 columns)
constexpr int kEpiW = 4;        // columns per epilogue-functor call (one float4 of one row)
constexpr int kEpiPitch = kEpiChunk + 1;   // padded row pitch of the per-warp transposition tile
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..: epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// one box {kBK, rows, 1} of a [plane][row][k] tensor -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int k, int row, int plane) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(k), "r"(row), "r"(plane)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, one 128 x BN x 8 tf32 MMA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive columns: thread i of the warp gets lane (base + i), v[j] = column (col + j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 consecutive columns, no wait: the caller batches loads and waits once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, 128-byte swizzled (what TMA
// wrote): 8-row groups 1024 bytes apart; version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major), in 16-byte units
  d |= (uint64_t)(1024u >> 4) << 32;      // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// The same for an MN-major tile (the contraction index is the SLOW one: rows = k, 32 contiguous m or n per 128-byte row,
// 128-byte swizzled as TMA wrote it).  Canonical layout (CUTLASS make_umma_desc<Major::MN>, SWIZZLE_128B, in 16-byte
// units): for 32-bit elements the atom is the 128-byte swizzle with 32-BYTE granularity (LayoutType SWIZZLE_128B_BASE32B,
// Swizzle<2,5,2>: Layout_MN_SW128_32B_Atom = 128 bytes along M / N x 4 k-rows; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -
// with the ordinary 16-byte-granular SWIZZLE_128B the products come out wrong (measured).  Atoms along M / N (chunks of 32
// elements) are LBO apart, groups of 4 k-rows SBO = 512 bytes apart; one tf32 MMA (K = 8) consumes two atoms per chunk.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((chunk_bytes >> 4) & 0x3FFFu) << 16;   // leading byte offset: next chunk of 32 elements along M / N
  d |= (uint64_t)(512u >> 4) << 32;                      // stride byte offset: next group of 4 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                // SWIZZLE_128B_BASE32B
  return d;
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
// ... both operands MN-major ("transposed": bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int bn) { return idesc_tf32(bn) | (1u << 15) | (1u << 16); }

#ifdef ENO_PHASES
// development aid (libeno_phases.so): clock stamps of CTA (0,0,0) of the most recent launches, 8 per launch:
// entry, set-up done, first operand stage landed, last MMA issued, accumulators complete, epilogue done, exit
static __device__ long long g_gemm_stamp[64 * 8];
static __device__ unsigned int g_gemm_launch;
#define GEMM_STAMP(slot) do { if (stamp) stamp[slot] = clock64(); } while (0)
#else
#define GEMM_STAMP(slot) do { } while (0)
#endif

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN >= 128) ? 3 : 4;
  static constexpr int kABytes = kBM * kBK * 4;   // one plane of the A tile
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// Epi: struct with  __device__ void operator()(int m, int n0, float (&v)[kEpiW], int split) const
// called once per (row, 4-column piece) of the tile; the functor
// masks rows / columns outside its matrices itself.  `done` (may be null): device word; non-zero = the solve has
// finished, the launch exits at once (iterations enqueued past the end of an adaptive solve).
// blockIdx.z is a K-split (a_zrows == 0: z takes k-blocks [z * kb_per_split, ...)) or a BATCH index (a_zrows != 0:
// problem z reads operand rows a_row0 + z * a_zrows / b_row0 + z * b_zrows over the whole K; the functor gets z).
// tap_kb != 0: 3x3 CONVOLUTION as a 9-tap K loop (ccv_kernels.cuh).  The A rows are the positions of a zero-padded
// image grid with tap_w positions per grid row and tap_kb * 32 channels per position, so tap (dy, dx) of output row
// m is input row m + dy * tap_w + dx: k-block i reads A at row shift(i / tap_kb), channel block i % tap_kb, and B (the
// weights, [n][tap][channel]) at k-block i.  Rows shifted out of the tensor read as zero (TMA out-of-bounds fill).
// (Only ROW coordinates are shifted: a shift along the contiguous dimension of a box is not a multiple of 16 bytes,
// which TMA rejects - so the weight gradient of the convolution, whose contraction runs over the positions, cannot use
// transposed planes this way; ccv_kernels.cuh: k_ccv_wgrad.)
// MN = true: D[m, n] = sum_k A[k, m] * B[k, n], both operands MN-major ([2 planes][k rows][ld], m / n contiguous): TMA boxes
// of 32 columns x 32 k-rows; a_row0 / b_row0 are COLUMN offsets then, and tap_w != 0 selects the convolution weight-gradient
// mode: tile column block x = tap pair, tile rows 0..63 / 64..127 = the two taps' input channels, A read shift(tap) rows
// further along K (a row coordinate: any shift is legal), B from column block 0.
template <int BN, class Epi, bool MN = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int a_row0, int b_row0,
              int kb_per_split, int kb_total, const int* __restrict__ done, Epi epi, int a_zrows = 0, int b_zrows = 0,
              int tap_kb = 0, int tap_w = 0) {
#ifdef ENO_PHASES
  __shared__ long long* stamp_s;
  if (threadIdx.x == 0) {
    stamp_s = nullptr;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      const unsigned int id = atomicAdd(&g_gemm_launch, 1u);
      stamp_s = g_gemm_stamp + (id % 64) * 8;
      stamp_s[0] = clock64();
      stamp_s[7] = ((long long)BN << 40) | ((long long)kb_total << 24) | ((long long)gridDim.x << 12) | gridDim.y;
    }
  }
  __syncthreads();
  long long* stamp = stamp_s;
#endif
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kBM, split = blockIdx.z;
  const bool batched = a_zrows != 0;
  const int kb0 = batched ? 0 : split * kb_per_split;
  const int kb1 = batched ? kb_total : min(kb_total, kb0 + kb_per_split);
  const int nkb = kb1 - kb0;
  int a_krow0 = 0, b_krow0 = 0;     // MN-major batched: the problems are stacked along the ROWS (= k) of the operands
  if (MN) { a_krow0 = split * a_zrows; b_krow0 = split * b_zrows; }
  else { a_row0 += split * a_zrows; b_row0 += split * b_zrows; }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kAccs * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GEMM_STAMP(1);
  // Programmatic dependent launch (launch_pdl below): everything above ran under the tail of the previous kernel; nothing
  // that kernel wrote (operands, the controller's `done` word) is read before this wait.  Both are no-ops in a plain launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool skip = done && *done;   // the solve has finished: a replayed iteration must not touch its results

  if (skip) {
  } else if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::kStages;
        const uint32_t ph = (uint32_t)(i / Cfg::kStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = smem + (size_t)s * Cfg::kStageBytes;
        mbar_expect_tx(&full[s], Cfg::kStageBytes);
        const int k = (kb0 + i) * kBK;
        if constexpr (MN) {
          constexpr int kChunk = kBK * 128;   // one box: 32 k-rows x 128 bytes
#pragma unroll
          for (int c = 0; c < kBM / 32; ++c) {
            int col = a_row0 + m0 + 32 * c, row = a_krow0 + k;
            if (tap_w) {
              const int tap = min(2 * (int)blockIdx.x + (c >> 1), 8);
              col = a_row0 + 32 * (c & 1);
              row += (tap / 3 - 1) * tap_w + (tap % 3 - 1);
            }
            tma_load_3d(st + c * kChunk, &tmA, &full[s], col, row, 0);
            tma_load_3d(st + Cfg::kABytes + c * kChunk, &tmA, &full[s], col, row, 1);
          }
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            const int col = b_row0 + (tap_w ? 0 : n0) + 32 * c;
            tma_load_3d(st + 2 * Cfg::kABytes + c * kChunk, &tmB, &full[s], col, b_krow0 + k, 0);
            tma_load_3d(st + 2 * Cfg::kABytes + Cfg::kBBytes + c * kChunk, &tmB, &full[s], col, b_krow0 + k, 1);
          }
          continue;
        }
        int ka = k, ra = a_row0 + m0;
        if (tap_kb) {
          const int tap = (kb0 + i) / tap_kb;
          ka = ((kb0 + i) - tap * tap_kb) * kBK;
          ra += (tap / 3 - 1) * tap_w + (tap % 3 - 1);
        }
        const int rb = b_row0 + n0;
        tma_load_3d(st, &tmA, &full[s], ka, ra, 0);
        tma_load_3d(st + Cfg::kABytes, &tmA, &full[s], ka, ra, 1);
        tma_load_3d(st + 2 * Cfg::kABytes, &tmB, &full[s], k, rb, 0);
        tma_load_3d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full[s], k, rb, 1);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = MN ? idesc_tf32_mn(BN) : idesc_tf32(BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::kStages;
        const uint32_t ph = (uint32_t)(i / Cfg::kStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (i == 0) GEMM_STAMP(2);
        const uint32_t sa = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
        constexpr uint32_t kChunk = kBK * 128;
        const uint64_t a_hi = MN ? smem_desc_mn_sw128(sa, kChunk) : smem_desc_k_sw128(sa);
        const uint64_t a_lo = MN ? smem_desc_mn_sw128(sa + Cfg::kABytes, kChunk) : smem_desc_k_sw128(sa + Cfg::kABytes);
        const uint64_t b_hi = MN ? smem_desc_mn_sw128(sa + 2 * Cfg::kABytes, kChunk) : smem_desc_k_sw128(sa + 2 * Cfg::kABytes);
        const uint64_t b_lo = MN ? smem_desc_mn_sw128(sa + 2 * Cfg::kABytes + Cfg::kBBytes, kChunk)
                                 : smem_desc_k_sw128(sa + 2 * Cfg::kABytes + Cfg::kBBytes);
        // The tensor core adds into its fp32 accumulator with truncation; a long chain of adds into ONE
        // accumulator shows up as a bias (measured: 2.1e-6 of sum|a||b| at K = 860 against 3.3e-7 for an fp32
        // FFMA GEMM).  So the hi*hi products rotate over three accumulators by k-block and the two small cross
        // terms go to a fourth; the epilogue adds the four with round-to-nearest.
        const uint32_t acc_main = tmem_base + (uint32_t)((i % (kAccs - 1)) * BN);
        const uint32_t acc_cross = tmem_base + (uint32_t)((kAccs - 1) * BN);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          // start-address field counts 16-byte units: K-major: 8 elements along the 128-byte row; MN-major: 8 k-rows = 1024 B
          const uint64_t adv = MN ? (uint64_t)((k * 1024) >> 4) : (uint64_t)((k * kUmmaK * 4) >> 4);
          umma_tf32(acc_cross, a_lo + adv, b_hi + adv, idesc, (i | k) != 0);
          umma_tf32(acc_cross, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_tf32(acc_main, a_hi + adv, b_hi + adv, idesc, (i >= kAccs - 1 || k != 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);
      GEMM_STAMP(3);
    }
    __syncwarp();
  } else {
    // Epilogue: 16 warps so that the element-wise work of the fused functors (exponentials, tf32 splits, several
    // output streams) has four warps per scheduler to hide its latencies behind - with four warps in all it took
    // longer than the main loop (tools/gemm_stamps.py).
    const int q = warp & 3;                // TMEM lane quadrant this warp may read
    const int cg = (warp - 2) >> 2;        // column group: BN / 4 columns
    constexpr int kCols = BN / 4;
    if (nkb > 0) {
      while (!mbar_try_wait(tmem_full, 0)) __nanosleep(128);
      tc_fence_after();
    }
    if (threadIdx.x == 64) GEMM_STAMP(4);
    const int used = min(nkb, kAccs - 1);
    // TMEM hands every thread one ROW of the tile.  Global accesses in that shape touch 32 different lines per
    // instruction and the epilogue became L1-wavefront bound; so the chunk is transposed through shared memory (the
    // operand stages are free once tmem_full has fired) into a quad layout: a thread owns 4 consecutive columns of 4
    // rows, 4 lanes cover the 16 columns of a row, a warp instruction touches 8 rows.
    float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kEpiPitch);
    const int cq = (lane & 3) * kEpiW, r0 = lane >> 2;
#pragma unroll 1
    for (int c = cg * kCols; c < (cg + 1) * kCols; c += kEpiChunk) {
      float v[kEpiChunk];
      if (nkb > 0) {
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
        float w[kEpiChunk];
        tmem_ld16_nowait(lane_base, v);
        tmem_ld16_nowait(lane_base + (uint32_t)((kAccs - 1) * BN), w);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < kEpiChunk; ++e) v[e] += w[e];
        if (used > 1) {   // accumulators the MMA warp never wrote hold garbage: they are not read
          float w2[kEpiChunk];
          tmem_ld16_nowait(lane_base + (uint32_t)BN, w);
          if (used > 2) tmem_ld16_nowait(lane_base + (uint32_t)(2 * BN), w2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < kEpiChunk; ++e) v[e] += w[e];
          if (used > 2) {
#pragma unroll
            for (int e = 0; e < kEpiChunk; ++e) v[e] += w2[e];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kEpiChunk; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < kEpiChunk; ++j) tile[lane * kEpiPitch + j] = v[j];
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = r0 + 8 * it;
        float x[kEpiW];
#pragma unroll
        for (int j = 0; j < kEpiW; ++j) x[j] = tile[row * kEpiPitch + cq + j];
        epi(m0 + q * 32 + row, n0 + c + cq, x, split);
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 64) GEMM_STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kAccs * BN);
  if (threadIdx.x == 64) GEMM_STAMP(6);
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map of an operand stored as [2 planes (hi, lo)][rows][ld] fp32, K contiguous; boxes of box_rows x 32
// elements.  Rows >= `rows` and columns >= `k` read as zero (out-of-bounds fill), so neither has to be padded.
inline bool make_operand_map(CUtensorMap* tm, const float* base, int rows, int k, int ld, size_t plane_stride, int box_rows,
                             std::string* err, bool mn_major = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled is not available from the driver"; return false; }
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)plane_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + "): rows " + std::to_string(rows) + " k " +
                    std::to_string(k) + " ld " + std::to_string(ld);
    return false;
  }
  return true;
}

struct GemmOperand {
  const float* base;    // [2][rows_alloc][ld]
  int rows, k, ld;      // logical extent
  size_t plane_stride;  // elements between the hi and the lo plane
  int row0;             // first row of this problem inside the tensor (stream offset)
};

// Launches of the flat-state engine carry the programmatic-stream-serialization attribute, so that a GEMM's prologue (barrier
// init, TMEM allocation, descriptor prefetch) overlaps the tail of the kernel before it: ffjord_tabular 45.1 -> 49.3 train
// steps/s, 51.1 with the element-wise kernels of the iteration launched the same way (css_kernels.cuh:CSS_PDL_ENTRY);
// ffjord_mnist 8.59 -> 8.68 (ENO_GEMM_PDL=0 restores plain launches for A/B runs).
inline bool gemm_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ENO_GEMM_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}
template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  if (!gemm_pdl_enabled()) {
    kern<<<grid, block, smem, st>>>(args...);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// D tile grid over (M, N) with `splits` K-partitions.  The functor sees global (m, n0) of the problem.
template <int BN, class Epi>
inline cudaError_t launch_gemm(const GemmOperand& A, const GemmOperand& Bm, int M, int N, int splits, const int* done,
                               const Epi& epi, cudaStream_t st, std::string* err) {
  CUtensorMap ta, tb;
  if (!make_operand_map(&ta, A.base, A.rows, A.k, A.ld, A.plane_stride, kBM, err)) return cudaErrorInvalidValue;
  if (!make_operand_map(&tb, Bm.base, Bm.rows, Bm.k, Bm.ld, Bm.plane_stride, BN, err)) return cudaErrorInvalidValue;
  const int kb_total = (A.k + kBK - 1) / kBK;
  const int kb_per = (kb_total + splits - 1) / splits;
  {   // per device, so set it on every launch (cheap)
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN>::kSmem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM, splits);
  return launch_pdl(k_gemm_tf32x3<BN, Epi>, grid, kGemmThreads, GemmCfg<BN>::kSmem, st, ta, tb, A.row0, Bm.row0, kb_per, kb_total, done,
                    epi, 0, 0, 0, 0);
}

// Batched launch over prebuilt tensor maps: grid.z problems, problem z at operand rows row0 + z * zrows (all K).
template <int BN, class Epi>
inline cudaError_t launch_gemm_batched(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int batch, int a_zrows,
                                       int b_zrows, const int* done, const Epi& epi, cudaStream_t st) {
  const int kb_total = (K + kBK - 1) / kBK;
  cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN>::kSmem);
  if (e != cudaSuccess) return e;
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM, batch);
  k_gemm_tf32x3<BN, Epi><<<grid, kGemmThreads, GemmCfg<BN>::kSmem, st>>>(ta, tb, 0, 0, kb_total, kb_total, done, epi, a_zrows, b_zrows);
  return cudaGetLastError();
}

// 3x3 convolution over a padded position grid (see k_gemm_tf32x3): M output positions starting at row a_row0 of the
// activation planes, N output channels, `cin` input channels per position (a multiple of 32), weights [N][9 * cin].
template <int BN, class Epi>
inline cudaError_t launch_conv3x3(const CUtensorMap& ta, const CUtensorMap& tb, int a_row0, int M, int N, int cin, int grid_w,
                                  const int* done, const Epi& epi, cudaStream_t st) {
  const int tap_kb = cin / kBK;
  const int kb_total = 9 * tap_kb;
  cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN>::kSmem);
  if (e != cudaSuccess) return e;
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM, 1);
  return launch_pdl(k_gemm_tf32x3<BN, Epi>, grid, kGemmThreads, GemmCfg<BN>::kSmem, st, ta, tb, a_row0, 0, kb_total, kb_total, done, epi,
                    0, 0, tap_kb, grid_w);
}

// MN-major operands: D[m, n] = sum_k A[k, m] B[k, n].  Maps from make_operand_map(tm, base, K rows, columns, ld, plane, 32).
// tap_w == 0: plain GEMM (grid = N tiles x M tiles x K-splits).  tap_w != 0: weight gradient of a 3x3 convolution over a
// padded position grid with tap_w positions per grid row: grid.x = 5 tap pairs, 64 input channels per tap (see the kernel).
template <int BN, class Epi>
inline cudaError_t launch_gemm_mn(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int splits, int tap_w,
                                  const int* done, const Epi& epi, cudaStream_t st) {
  const int kb_total = (K + kBK - 1) / kBK;
  const int kb_per = (kb_total + splits - 1) / splits;
  cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3<BN, Epi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN>::kSmem);
  if (e != cudaSuccess) return e;
  dim3 grid(tap_w ? 5 : (N + BN - 1) / BN, tap_w ? 1 : (M + kBM - 1) / kBM, splits);
  return launch_pdl(k_gemm_tf32x3<BN, Epi, true>, grid, kGemmThreads, GemmCfg<BN>::kSmem, st, ta, tb, 0, 0, kb_per, kb_total, done, epi,
                    0, 0, 0, tap_w);
}

// ---- operand preparation ----------------------------------------------------------------------------------
// hi = x rounded to NEAREST tf32, lo = (x - hi) rounded to nearest tf32: both pieces are exactly what the tensor core
// reads (it would otherwise TRUNCATE the 13-bit remainder to 11 bits: a one-sided error of up to 2^-20 |x| per operand;
// this way |x - hi - lo| <= 2^-22 |x|, unbiased, and every product hi*hi, hi*lo, lo*hi is exact in fp32)
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rn_tf32(x);
  lo = rn_tf32(x - hi);
}

// src [rows][cols] (ld_src) -> dst planes [2][rows_alloc][ld_dst] (optionally transposed: dst row = src col)
static __global__ void k_split_planes(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ dst, int ld_dst,
                               size_t plane_stride, int transpose) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r, c;
    if (transpose) { c = (int)(i % rows); r = (int)(i / rows); }   // consecutive threads walk a dst row
    else { r = (int)(i / cols); c = (int)(i % cols); }
    // (r, c): destination coordinates
    const float x = transpose ? src[(size_t)c * ld_src + r] : src[(size_t)r * ld_src + c];
    float hi, lo;
    split_tf32(x, hi, lo);
    dst[(size_t)r * ld_dst + c] = hi;
    dst[plane_stride + (size_t)r * ld_dst + c] = lo;
  }
}

// plain epilogue: C[m, n] = acc (row-major, ld), partial sums of split s at C + s * split_stride
struct EpiStore {
  float* C;
  int M, N, ld;
  size_t split_stride;
  __device__ __forceinline__ void operator()(int m, int n0, float (&v)[kEpiW], int split) const {
    if (m >= M) return;
    float* row = C + (size_t)split * split_stride + (size_t)m * ld;
#pragma unroll
    for (int j = 0; j < kEpiW; ++j)
      if (n0 + j < N) row[n0 + j] = v[j];
  }
};

// C[r, c] = sum_s part[s][r * n_cols + c], fixed order (deterministic split-K)
static __global__ void k_sum_splits(const float* __restrict__ part, int splits, size_t n, int n_cols, float* __restrict__ C, int ldc) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[(size_t)s * n + i];
    C[(i / n_cols) * (size_t)ldc + (i % n_cols)] = acc;
  }
}

}  // namespace tc
}  // namespace eno
