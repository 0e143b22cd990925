// Host-side driver of the adjoint solve: workspace carving, path selection (cluster-resident vs
// streamed weights), launch sequence of lib/ode.py:1494-1529 per output segment.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>

#include "adj_kernels.cuh"
#include "cl_kernels.cuh"
#include <cstdlib>

#include "dist.cuh"
#include "tableaux.cuh"

namespace eno {

// ======================================================================================
// host side
// ======================================================================================
inline size_t adj_align(size_t x) { return (x + 255) / 256 * 256; }

// geometry of the tensor-core factor sinks (FactorTc, adj_device.cuh)
struct PgradTcGeom {
  int ldk, dpad, hpad, Bk;
  size_t xplane, yplane, x_elems, y_elems;
};
inline PgradTcGeom pgrad_tc_geom(int B, int D, int H) {
  PgradTcGeom g;
  g.Bk = (B + 3) / 4 * 4;
  g.ldk = 3 * g.Bk;
  g.dpad = (D + tc::kBM - 1) / tc::kBM * tc::kBM;
  g.hpad = (H + 127) / 128 * 128;
  g.xplane = (size_t)kSlots * 2 * g.dpad * g.ldk;
  g.yplane = (size_t)kSlots * 2 * g.hpad * g.ldk;
  g.x_elems = 2 * g.xplane;
  g.y_elems = 2 * g.yplane;
  return g;
}

inline size_t adj_ws_bytes(int B, int D, int H, int world = 1, bool fin = false, int slot = 0) {
  slot = std::max(slot, B);   // rows of the largest shard (uneven shards: every rank's record is laid out for it)
  const size_t N = (size_t)B * D, P = (size_t)(D + 1) * H + H + (size_t)(H + 1) * D + D;
  const PgradGeom g = pgrad_geom(D, H);
  size_t s = 0;
  s += adj_align(sizeof(Ctrl)) + adj_align(sizeof(AdjScalars));
  s += 4 * adj_align(3 * N * 4);                    // Z, ZB, FZ, FZB
  s += 2 * adj_align(3 * 2 * (size_t)B * 4) + adj_align(2 * (size_t)B * 4);  // RS, FR, RB (up to 2 scalars per sample)
  if (fin) s += 2 * adj_align(3 * N * 4) + adj_align(2 * N * 4) + adj_align(N * 4);   // ZE, FZE, MIDZE, EBAR
  s += adj_align(2 * N * 4);                        // MIDZB
  s += 2 * adj_align(3 * P * 4) + adj_align(2 * P * 4);             // PB, FP, MIDP
  const size_t npp = std::max(std::max((size_t)g.grid, (P + 255) / 256), (size_t)kPcGrid);
  s += kSlots * (2 * adj_align(3 * N * 4) + 2 * adj_align(3 * (size_t)B * H * 4));
  s += adj_align((size_t)6 * kC * slot * world * 8) + 2 * adj_align(npp * 8);   // xrec (one record per rank), ppart0/1
  if (world > 1) s += adj_align(4 * P * 4);          // pcomb
  s += adj_align(N * 4) + adj_align((P + 256 * (size_t)(world + 1)) * 4);   // YBAR, POUT (padded to world equal 256-multiples)
  s += 2 * adj_align((size_t)D * H * 4);            // transposes
  const PgradTcGeom tg = pgrad_tc_geom(B, D, H);   // tensor-core factor sinks + stage products
  s += adj_align(tg.x_elems * 4) + adj_align(tg.y_elems * 4) + adj_align((size_t)kSlots * P * 4);
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

inline void adj_carve(void* ws, int B, int D, int H, const float* params, AdjArgs* a, int world = 1, bool fin = false, int slot = 0) {
  slot = std::max(slot, B);
  const size_t N = (size_t)B * D, P = (size_t)(D + 1) * H + H + (size_t)(H + 1) * D + D;
  const PgradGeom g = pgrad_geom(D, H);
  AdjCarver cv{static_cast<char*>(ws)};
  a->ctrl = cv.take<Ctrl>(1);
  a->sc = cv.take<AdjScalars>(1);
  a->Z = cv.take<float>(3 * N); a->ZB = cv.take<float>(3 * N);
  a->FZ = cv.take<float>(3 * N); a->FZB = cv.take<float>(3 * N);
  a->RS = cv.take<float>(3 * 2 * (size_t)B); a->FR = cv.take<float>(3 * 2 * (size_t)B); a->RB = cv.take<float>(2 * (size_t)B);
  a->na = fin ? 2 : 1;
  a->fin = fin ? 1 : 0;
  if (fin) {
    a->ZE = cv.take<float>(3 * N); a->FZE = cv.take<float>(3 * N);
    a->MIDZE = cv.take<float>(2 * N); a->EBAR = cv.take<float>(N);
  }
  a->MIDZB = cv.take<float>(2 * N);
  a->PB = cv.take<float>(3 * P); a->FP = cv.take<float>(3 * P); a->MIDP = cv.take<float>(2 * P);
  for (int j = 0; j < kSlots; ++j) {
    a->slot[j].S = cv.take<float>(3 * N);
    a->slot[j].G = cv.take<float>(3 * N);
    a->slot[j].A = cv.take<float>(3 * (size_t)B * H);
    a->slot[j].Hh = cv.take<float>(3 * (size_t)B * H);
    a->slot[j].tbar = nullptr;   // adj_set_views
  }
  a->xrec = cv.take<double>((size_t)6 * kC * slot * world);
  a->rowpart0 = a->rowpart1 = a->rowpart2 = nullptr;
  a->n_rp = a->n_loc = B;
  const size_t npp = std::max(std::max((size_t)g.grid, (P + 255) / 256), (size_t)kPcGrid);
  a->ppart0 = cv.take<double>(npp); a->ppart1 = cv.take<double>(npp);
  a->pcomb = (world > 1) ? cv.take<float>(4 * P) : nullptr;
  a->dist = 0;
  a->peer_base = nullptr;
  a->YBAR = cv.take<float>(N);
  a->POUT = cv.take<float>(P + 256 * (size_t)(world + 1));
  float* W1xT = cv.take<float>((size_t)D * H);
  float* W2xT = cv.take<float>((size_t)D * H);
  a->w = mlp_weights_from_flat(params, D, H, W1xT, W2xT);
  {
    const PgradTcGeom tg = pgrad_tc_geom(B, D, H);
    a->tc.X = nullptr;   // adj_solve switches the sinks on when the path and the shape allow it
    a->tc.Y = nullptr;
    a->tc.xplane = (unsigned)tg.xplane; a->tc.yplane = (unsigned)tg.yplane;
    a->tc.ldk = tg.ldk; a->tc.dpad = tg.dpad; a->tc.hpad = tg.hpad; a->tc.Bk = tg.Bk;
    a->tcX_ws = cv.take<float>(tg.x_elems);
    a->tcY_ws = cv.take<float>(tg.y_elems);
    a->KP = cv.take<float>((size_t)kSlots * P);
  }
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

// Everything a rank contributes to the controller besides the params_bar block lives in ONE record of
// 6 * n_loc doubles, so that the multi-GPU exchange is a single all-gather (plus the params_bar all-reduce):
//   [ rp0: n_loc doubles | t_bar pieces of slots 0..5: 6 * n_loc floats | rp1: n_loc doubles | rp2: n_loc doubles ]
// Point the kernels' write views at this rank's record and the readers at rank 0's (block_sum_records).
inline void adj_set_views(AdjArgs* a, int stride, int rank, int world, int slot = 0) {
  const size_t n_loc = (size_t)stride * std::max(slot, a->B);   // slot rows per rank (>= this rank's B; the tail stays zero)
  a->n_loc = (int)n_loc;
  a->n_rp = (int)(n_loc * world);
  a->rp0_all = a->xrec;
  a->rp1_all = a->xrec + 4 * n_loc;
  a->rp2_all = a->xrec + 5 * n_loc;
  double* mine = a->xrec + (size_t)rank * 6 * n_loc;
  a->rowpart0 = mine;
  a->rowpart1 = mine + 4 * n_loc;
  a->rowpart2 = mine + 5 * n_loc;
  float* f0 = reinterpret_cast<float*>(a->xrec) + 2 * n_loc;   // rank 0's t_bar pieces
  for (int j = 0; j < kSlots; ++j) {
    a->tbar_all[j] = f0 + (size_t)j * n_loc;
    a->slot[j].tbar = f0 + (size_t)rank * 12 * n_loc + (size_t)j * n_loc;
  }
}

constexpr size_t kPgradSmem = (size_t)kSlots * kPStages * kPRows * 64 * sizeof(float);   // operand rings of k_adj_pgrad

// The opt-in is a per-DEVICE attribute: set it on every solve (four cheap calls) rather than once per thread.
inline cudaError_t pgrad_prepare() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_adj_pgrad<ADJ_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPgradSmem))) return e;
  if ((e = cudaFuncSetAttribute(k_adj_pgrad<ADJ_INIT0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPgradSmem))) return e;
  if ((e = cudaFuncSetAttribute(k_adj_pgrad<ADJ_INIT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPgradSmem))) return e;
  if ((e = cudaFuncSetAttribute(k_adj_pgrad<ADJ_RHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPgradSmem))) return e;
  return cudaSuccess;
}

// One launch of the per-sample kernel on the chosen path.
struct AdjPlan {
  bool cluster;
  int R, grid;
  size_t smem;
};

inline AdjPlan adj_plan(int B, int D, int H, int sms, int path_mode, bool single_rhs) {
  AdjPlan p;
  p.cluster = path_mode != 1 && cl_supported(D, H, kClRows, kAdjStages, true);
  if (p.cluster) {
    p.R = kClRows;
    p.grid = cl_grid(B);
    p.smem = cl_smem_bytes(kClRows, D, H, kAdjStages, true);
  } else {
    p.R = single_rhs ? 1 : adj_pick_rows(B, D, H, sms);
    p.grid = (B + p.R - 1) / p.R;
    p.smem = adj_smem_bytes(p.R, D, H, single_rhs ? 2 : kAdjStages);
  }
  return p;
}

inline int adj_rows_any(const AdjPlan& p, int K, int mode, const AdjArgs& a, float rhs_tau, cudaStream_t st,
                        std::string* err) {
  if (p.cluster) {
    cudaError_t e = cl_adj_rows(K, mode, a, p.grid, p.smem, rhs_tau, st);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cluster launch of k_cl_adj_rows: ") + cudaGetErrorString(e);
      return ENO_E_CUDA;
    }
    return 0;
  }
  return adj_rows(p.R, K, mode, a, p.grid, p.smem, rhs_tau, st, err);
}

inline void adj_transposes(const AdjArgs& a, cudaStream_t st) {
  const int D = a.w.D, H = a.w.H;
  dim3 blk(32, 8);
  k_transpose<<<dim3((H + 31) / 32, (D + 31) / 32), blk, 0, st>>>(a.w.W1x, const_cast<float*>(a.w.W1xT), D, H);
  k_transpose<<<dim3((D + 31) / 32, (H + 31) / 32), blk, 0, st>>>(a.w.W2x, const_cast<float*>(a.w.W2xT), H, D);
}

inline int adj_solve(Profiler* prof, int path_mode, const DistCtx* dc, int deferred_iters, const Tableau* tab_host,
                     std::string* err, Ctrl* mailbox, int* hint, int sms, const eno_dynamics_desc* desc, int B,
                     const float* ys, const float* ts, int T, const float* params, const float* eps, const float* g_y,
                     const float* g_r, float rtol, float atol, int mxstep, void* ws, size_t ws_bytes, float* y0_bar_out,
                     float* r0_bar_out, float* eps_bar_out, float* ts_bar_out, float* params_bar_out, eno_stats* stats,
                     eno_trace_row* trace, int trace_cap, cudaStream_t st) {
  const bool fin = desc->aug == ENO_AUG_FIN;
  const int D = desc->D, H = desc->H, K = fin ? 2 : desc->reg_order;
  if (!(K == 0 || K == 2 || K == 3)) { if (err) *err = "eno_odeint_bwd: reg_order must be 0, 2 or 3"; return ENO_E_UNSUPPORTED; }
  if (fin && !eps) { if (err) *err = "eno_odeint_bwd: ENO_AUG_FIN needs eps"; return ENO_E_ARG; }
  if (B <= 0 || T < 1 || !ys || !ts || !params || !g_y || !g_r || !ws || !y0_bar_out || !r0_bar_out || !ts_bar_out ||
      !params_bar_out) { if (err) *err = "eno_odeint_bwd: null pointer or empty batch"; return ENO_E_ARG; }
  const bool dist = dc && dc->active();
  const int world = dist ? dc->world : 1, rank = dist ? dc->rank : 0;
  const int slot = dist ? dc->slot(B) : B;
  if (B > slot) { if (err) *err = "eno_odeint_bwd: this rank's batch exceeds the shard slot (eno_comm_set_shards)"; return ENO_E_ARG; }
  if (ws_bytes < adj_ws_bytes(B, D, H, world, fin, slot)) { if (err) *err = "eno_odeint_bwd: workspace too small"; return ENO_E_WORKSPACE; }
  const size_t N = (size_t)B * D;
  AdjArgs a = {};
  adj_carve(ws, B, D, H, params, &a, world, fin, slot);
  a.eps = eps;
  const size_t na = (size_t)a.na;
  a.dist = dist ? 1 : 0;
  a.np = (K == 0) ? 1 : K;
  {
    const size_t n_aux = (desc->aug == ENO_AUG_NONE) ? 0 : na;
    const size_t Bg = (size_t)(dist ? dc->rows(B) : B), Ng = Bg * D;   // the ODE state spans the GLOBAL batch
    a.n_state = (float)(2 * (Ng + n_aux * Bg) + 1 + (desc->eps_in_args ? Ng : 0) + (size_t)a.P);
  }
  a.trace = trace;
  // tab_host lives in pinned memory owned by the context, so the copy stays valid if this call is being
  // captured into a CUDA graph (deferred mode)
  ENO_ADJ_CHECK(cudaMemcpyToSymbolAsync(c_tab, tab_host, sizeof(Tableau), 0, cudaMemcpyHostToDevice, st));
  const bool deferred = deferred_iters > 0;
  if (deferred && T != 2) { if (err) *err = "eno_odeint_bwd: deferred mode supports T == 2 only"; return ENO_E_UNSUPPORTED; }
  const AdjPlan plan = adj_plan(B, D, H, sms, path_mode, false);
  if (fin && !plan.cluster) {
    if (err) *err = "eno_odeint_bwd: ENO_AUG_FIN runs on the cluster-resident path only (shape or ENO_PATH excludes it)";
    return ENO_E_UNSUPPORTED;
  }
  ENO_ADJ_CHECK(pgrad_prepare());
  if (!plan.cluster) adj_transposes(a, st);
  adj_set_views(&a, plan.cluster ? kC : 1, rank, world, slot);
  const PgradGeom geo = pgrad_geom(D, H);
  // ADJ_STEP parameter gradients of the cluster path run on the tensor cores (batched tcgen05 3xTF32 GEMM +
  // k_adj_pcombine): k_cl_adj_rows<ADJ_STEP> writes its factors as K-major hi / lo planes
  // (single GPU and collective solves alike; ENO_PGRAD_FFMA=1 keeps the FFMA kernel k_adj_pgrad for A/B runs - measured
  // at 2 GPUs: 931 steps/s with it against 944 with the tensor-core path)
  static const bool pg_ffma = getenv("ENO_PGRAD_FFMA") != nullptr;
  const bool pg_tc = plan.cluster && !pg_ffma;
  if (pg_tc && (size_t)kSlots * 2 * a.tc.dpad * a.tc.ldk >= (1ull << 31)) {
    if (err) *err = "eno_odeint_bwd: batch too large for the tensor-core factor planes";
    return ENO_E_UNSUPPORTED;
  }
  CUtensorMap tmX, tmY;
  if (pg_tc) {
    a.tc.X = a.tcX_ws;
    a.tc.Y = a.tcY_ws;
    std::string terr;
    if (!tc::make_operand_map(&tmX, a.tc.X, kSlots * 2 * a.tc.dpad, a.np * a.tc.Bk, a.tc.ldk, a.tc.xplane, tc::kBM, &terr) ||
        !tc::make_operand_map(&tmY, a.tc.Y, kSlots * 2 * a.tc.hpad, a.np * a.tc.Bk, a.tc.ldk, a.tc.yplane, 128, &terr)) {
      if (err) *err = "eno_odeint_bwd: " + terr;
      return ENO_E_CUDA;
    }
  }
  const EpiPgradStage pg_epi{a.KP, D, H, a.P, 2 * H, H * (D + 2) + 2 * D};
  const size_t n_loc = (size_t)(plan.cluster ? kC : 1) * slot;
  const int pf_grid = (a.P + 255) / 256;
  // multi-GPU continuation of one pgrad launch: exchange, then the replicated finish kernel
  // peer-memory exchange (dist.cuh) when the ranks have mapped each other's exchange buffers
  const int slice = (int)((((size_t)a.P + world - 1) / world + 255) / 256 * 256);   // params_bar entries per rank
  const size_t peer_xrec_bytes = (size_t)2 * world * 6 * n_loc * 8;
  const bool peer = dist && dc->peer_ready() && !getenv("ENO_EXCHANGE_NCCL") &&
                    kPeerOffAdj + peer_xrec_bytes + 256 + (size_t)2 * 4 * a.P * 4 <= dc->peer.bytes;
  if (peer) {
    a.peer_base = dc->peer.base[rank];
    a.peer_pcomb_off = kPeerOffAdj + ((peer_xrec_bytes + 255) / 256) * 256;
    a.p_lo = rank * slice;
    a.p_cnt = slice;
  }
  auto exchange = [&](int mode) -> int {
    if (!dist) return 0;
    if (peer) {
      const int g = slice / 256;
      if (mode == ADJ_STEP) k_adj_pfinish_peer<ADJ_STEP><<<g, 256, 0, st>>>(a);
      else if (mode == ADJ_INIT0) k_adj_pfinish_peer<ADJ_INIT0><<<g, 256, 0, st>>>(a);
      else k_adj_pfinish_peer<ADJ_INIT1><<<g, 256, 0, st>>>(a);
      return 0;
    }
    int rc = dc->api.GroupStart();
    rc |= dc->allreduce(a.pcomb, (mode == ADJ_STEP ? 4 : 1) * (size_t)a.P, st);
    rc |= dc->gather(a.xrec, 6 * n_loc, st);   // every rank's record: norm partials + t_bar pieces
    rc |= dc->api.GroupEnd();
    if (rc) { if (err) *err = "eno_odeint_bwd: NCCL exchange failed"; return ENO_E_CUDA; }
    if (mode == ADJ_STEP) k_adj_pfinish<ADJ_STEP><<<pf_grid, 256, 0, st>>>(a);
    else if (mode == ADJ_INIT0) k_adj_pfinish<ADJ_INIT0><<<pf_grid, 256, 0, st>>>(a);
    else k_adj_pfinish<ADJ_INIT1><<<pf_grid, 256, 0, st>>>(a);
    return 0;
  };
  const int egrid = (int)std::min<size_t>((std::max(N, (size_t)a.P) + 255) / 256, (size_t)4 * sms);
  const int mx = mxstep <= 0 ? INT_MAX : mxstep;
  int worst = ENO_OK;
  if (T == 1) {   // nothing to integrate: y0_bar = g[0], everything else zero (empty scan)
    ENO_ADJ_CHECK(cudaMemcpyAsync(y0_bar_out, g_y, N * 4, cudaMemcpyDeviceToDevice, st));
    ENO_ADJ_CHECK(cudaMemcpyAsync(r0_bar_out, g_r, na * B * 4, cudaMemcpyDeviceToDevice, st));
    if (eps_bar_out) ENO_ADJ_CHECK(cudaMemsetAsync(eps_bar_out, 0, N * 4, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(ts_bar_out, 0, 4, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(params_bar_out, 0, (size_t)a.P * 4, st));
    ENO_ADJ_CHECK(cudaStreamSynchronize(st));
    return ENO_OK;
  }
  for (int i = T - 1; i >= 1; --i) {
    int launches = 0;
    const int first = (i == T - 1);
    a.g_y = g_y + (size_t)i * N;
    a.g_r = g_r + (size_t)i * na * B;
    a.g_y_prev = g_y + (size_t)(i - 1) * N;
    a.tbar_out = ts_bar_out + i;
    if (slot != B)   // uneven shards: the tail of this rank's record (rows B .. slot) must read as zero
      ENO_ADJ_CHECK(cudaMemsetAsync(a.xrec + (size_t)rank * 6 * n_loc, 0, 6 * n_loc * 8, st));
    k_adj_seg_begin<<<egrid, 256, 0, st>>>(a, ys + (size_t)i * N, a.g_y, a.g_r, ts, i, first, rtol, atol, mx,
                                           trace ? trace_cap : 0);
    int rc;
    if ((rc = adj_rows_any(plan, K, ADJ_INIT0, a, 0.f, st, err))) return rc;
    k_adj_pgrad<ADJ_INIT0><<<geo.grid, kSlots * kGroup, kPgradSmem, st>>>(a, 0.f, nullptr, nullptr);
    if ((rc = exchange(ADJ_INIT0))) return rc;
    if ((rc = adj_rows_any(plan, K, ADJ_INIT1, a, 0.f, st, err))) return rc;
    k_adj_pgrad<ADJ_INIT1><<<geo.grid, kSlots * kGroup, kPgradSmem, st>>>(a, 0.f, nullptr, nullptr);
    if ((rc = exchange(ADJ_INIT1))) return rc;
    launches += dist ? 7 : 5;
    int chunk = deferred ? deferred_iters : *hint;
    for (;;) {
      for (int k = 0; k < chunk; ++k) {
        int ps = prof->begin(PROF_ADJ_ROWS, st);
        if ((rc = adj_rows_any(plan, K, ADJ_STEP, a, 0.f, st, err))) return rc;
        prof->end(ps, st);
        ps = prof->begin(PROF_ADJ_PGRAD, st);
        if (pg_tc) {
          const int pg = prof->begin(PROF_ADJ_PGEMM, st);
          ENO_ADJ_CHECK(tc::launch_gemm_batched<128>(tmX, tmY, D, H, a.np * a.tc.Bk, kSlots * 2, a.tc.dpad, a.tc.hpad, &a.ctrl->done,
                                                     pg_epi, st));
          prof->end(pg, st);
          k_adj_pcombine<<<kPcGrid, 256, 0, st>>>(a);
          ++launches;
        } else {
          k_adj_pgrad<ADJ_STEP><<<geo.grid, kSlots * kGroup, kPgradSmem, st>>>(a, 0.f, nullptr, nullptr);
        }
        prof->end(ps, st);
        ps = dist ? prof->begin(PROF_EXCHANGE, st) : -1;
        if ((rc = exchange(ADJ_STEP))) return rc;
        prof->end(ps, st);
        launches += dist ? 3 : 2;
      }
      // A segment has one target, so a dense output can only become pending on the iteration that ends
      // it; later launches exit at their first instruction and leave the rings intact.  One output
      // launch per chunk (a no-op unless the segment finished) is therefore enough.
      k_adj_out<<<egrid, 256, 0, st>>>(a);
      ++launches;
      if (deferred) {   // no host round-trip: the caller checks Ctrl.done later (eno_read_stats)
        if (peer && dc->gather(a.POUT, (size_t)slice, st)) { if (err) *err = "eno_odeint_bwd: NCCL all-gather of params_bar failed"; return ENO_E_CUDA; }
        break;
      }
      ENO_ADJ_CHECK(cudaMemcpyAsync(mailbox, a.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
      ENO_ADJ_CHECK(cudaStreamSynchronize(st));
      if (mailbox->done) {
        // peer mode: every rank holds its own slice of params_bar; gather them once per segment
        if (peer && dc->gather(a.POUT, (size_t)slice, st)) { if (err) *err = "eno_odeint_bwd: NCCL all-gather of params_bar failed"; return ENO_E_CUDA; }
        break;
      }
      chunk = 4;
    }
    if (deferred) {
      if (stats) { memset(stats, 0, sizeof(eno_stats)); stats->status = -100; stats->n_launches = launches; }
      continue;
    }
    {
      int valid[PROF_KINDS] = {0, mailbox->n_iters, mailbox->n_iters, dist ? mailbox->n_iters : 0, pg_tc ? mailbox->n_iters : 0};
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
  if (eps_bar_out) {
    // reg_type 'our' never reads eps: its cotangent stays identically zero (it only dilutes the error norm)
    if (fin) ENO_ADJ_CHECK(cudaMemcpyAsync(eps_bar_out, a.EBAR, N * 4, cudaMemcpyDeviceToDevice, st));
    else ENO_ADJ_CHECK(cudaMemsetAsync(eps_bar_out, 0, N * 4, st));
  }
  ENO_ADJ_CHECK(cudaMemcpyAsync(params_bar_out, a.POUT, (size_t)a.P * 4, cudaMemcpyDeviceToDevice, st));
  if (!deferred) ENO_ADJ_CHECK(cudaStreamSynchronize(st));
  ENO_ADJ_CHECK(cudaGetLastError());
  return worst;
}

inline int adj_rhs(int path_mode, std::string* err, int sms, const eno_dynamics_desc* desc, int mode, int B, const float* y, float t,
                   const float* params, const float* eps, const float* y_bar, const float* r_bar, void* ws, size_t ws_bytes,
                   float* dy_out, float* aux_out, float* ybar_dot_out, float* tbar_out, float* params_bar_out,
                   float* eps_bar_dot_out, cudaStream_t st) {
  const bool fin = desc->aug == ENO_AUG_FIN;
  const int D = desc->D, H = desc->H, K = fin ? 2 : desc->reg_order;
  if (!(K == 0 || K == 2 || K == 3)) { if (err) *err = "eno_rhs: reg_order must be 0, 2 or 3"; return ENO_E_UNSUPPORTED; }
  if (fin && !eps) { if (err) *err = "eno_rhs: ENO_AUG_FIN needs eps"; return ENO_E_ARG; }
  if (!ws || ws_bytes < adj_ws_bytes(B, D, H, 1, fin)) { if (err) *err = "eno_rhs: workspace too small"; return ENO_E_WORKSPACE; }
  if (mode == 2 && (!y_bar || !r_bar || !ybar_dot_out || !tbar_out || !params_bar_out || !aux_out)) {
    if (err) *err = "eno_rhs: mode 2 needs y_bar, r_bar and all outputs"; return ENO_E_ARG;
  }
  const size_t N = (size_t)B * D;
  AdjArgs a = {};
  adj_carve(ws, B, D, H, params, &a, 1, fin);
  a.np = (K == 0) ? 1 : K;
  a.eps = eps;
  const size_t na = (size_t)a.na;
  const AdjPlan plan = adj_plan(B, D, H, sms, path_mode, true);
  if (fin && !plan.cluster) {
    if (err) *err = "eno_rhs: ENO_AUG_FIN runs on the cluster-resident path only";
    return ENO_E_UNSUPPORTED;
  }
  ENO_ADJ_CHECK(pgrad_prepare());
  if (!plan.cluster) adj_transposes(a, st);
  adj_set_views(&a, plan.cluster ? kC : 1, 0, 1);
  ENO_ADJ_CHECK(cudaMemcpyAsync(a.Z, y, N * 4, cudaMemcpyDeviceToDevice, st));
  if (mode == 2) {
    ENO_ADJ_CHECK(cudaMemcpyAsync(a.ZB, y_bar, N * 4, cudaMemcpyDeviceToDevice, st));
    ENO_ADJ_CHECK(cudaMemcpyAsync(a.RB, r_bar, na * B * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    ENO_ADJ_CHECK(cudaMemsetAsync(a.ZB, 0, N * 4, st));
    ENO_ADJ_CHECK(cudaMemsetAsync(a.RB, 0, na * B * 4, st));
  }
  ENO_ADJ_CHECK(cudaMemsetAsync(a.RS, 0, na * B * 4, st));
  if (fin) ENO_ADJ_CHECK(cudaMemsetAsync(a.ZE, 0, N * 4, st));
  ENO_ADJ_CHECK(cudaMemsetAsync(a.ctrl, 0, sizeof(Ctrl), st));
  a.rhs_dy = dy_out;
  a.rhs_r = aux_out;
  a.rhs_zb = ybar_dot_out;
  a.rhs_eb = eps_bar_dot_out;
  int rc;
  // solver time tau = -t, so that the kernel's t = -tau is the caller's t
  if ((rc = adj_rows_any(plan, K, ADJ_RHS, a, -t, st, err))) return rc;
  if (mode == 2) {
    const PgradGeom geo = pgrad_geom(D, H);
    k_adj_pgrad<ADJ_RHS><<<geo.grid, kSlots * kGroup, kPgradSmem, st>>>(a, -t, params_bar_out, tbar_out);
  }
  ENO_ADJ_CHECK(cudaGetLastError());
  ENO_ADJ_CHECK(cudaStreamSynchronize(st));
  return ENO_OK;
}

}  // namespace eno
