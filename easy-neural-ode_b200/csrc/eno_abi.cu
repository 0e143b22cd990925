// C ABI (include/eno.h) and host-side launch logic of libeno.so.
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>
#include <cstring>
#include <string>

#include "eno_ctx.cuh"
#include "css_api.h"
#include "adj_host.cuh"
#include "cl_aug_fwd.cuh"
#include "cl_kernels.cuh"
#include "cl_vmap.cuh"
#include "fwd_kernels.cuh"
#include "head_kernels.cuh"
#include "tableaux.cuh"

using namespace eno;


namespace {

thread_local std::string g_err;

int fail(eno_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

#define ENO_CUDA(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(ctx, ENO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));             \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T));
    return p;
  }
};

bool desc_ok(const eno_dynamics_desc* d) {
  return d && d->arch == ENO_ARCH_MLP_SIGMOID && d->D > 0 && d->H > 0 && d->D % 4 == 0 && d->H % 4 == 0 &&
         d->D / 4 <= kThreads && d->H / 4 <= kThreads;
}

int pick_rows_fwd(int B, int D, int H, int s, int sms) {
  int R = 1;
  if (B > sms) R = 2;
  if (B > 2 * sms) R = 4;
  while (R > 1 && fwd_smem_bytes(R, D, H, s) > 220 * 1024) R >>= 1;
  return R;
}

__global__ void k_ctrl_reset(Ctrl* c, float rtol, float atol, int T, int mxstep, int trace_cap, float n_state) {
  memset(c, 0, sizeof(Ctrl));
  c->rtol = rtol;
  c->atol = atol;
  c->T = T;
  c->mxstep = mxstep;
  c->trace_cap = trace_cap;
  c->n_state = n_state;
}

// 16 independent FMA chains per thread: the non-tensor fp32 peak this chip actually reaches.
__global__ void k_ffma_peak(float* sink, int iters, float m) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, 0.5f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) *sink = s;
}

template <typename KernelT>
cudaError_t set_smem(KernelT k, size_t bytes) {
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int g_world = 1;   // ranks of this process group (set by eno_comm_init); sizes the gathered partial-sum arrays
int g_slot = 0;    // rows of the largest shard (eno_comm_set_shards); 0 = equal shards
inline int slot_of(int B) { return std::max(B, g_world > 1 ? g_slot : 0); }

size_t fwd_ws_bytes(int B, int D) {
  const size_t N = (size_t)B * D;
  return align_up(sizeof(Ctrl)) + align_up(3 * N * 4) * 2 + align_up(2 * N * 4) +
         2 * align_up((size_t)kC * slot_of(B) * g_world * 8) + 3 * align_up((size_t)9 * B * 4) + 1024;
}

template <int R>
cudaError_t fwd_stream_launch(int mode, const FwdArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e;
  if (mode == 1) {
    if ((e = set_smem(k_fwd_init<R, 0>, smem)) != cudaSuccess) return e;
    k_fwd_init<R, 0><<<grid, kThreads, smem, st>>>(a);
  } else if (mode == 2) {
    if ((e = set_smem(k_fwd_init<R, 1>, smem)) != cudaSuccess) return e;
    k_fwd_init<R, 1><<<grid, kThreads, smem, st>>>(a);
  } else {
    if ((e = set_smem(k_fwd_step<R>, smem)) != cudaSuccess) return e;
    k_fwd_step<R><<<grid, kThreads, smem, st>>>(a);
  }
  return cudaGetLastError();
}


void fill_stats(eno_stats* st, const Ctrl& c, int launches, int n_rhs) {
  if (!st) return;
  st->nfe = (float)c.nfe;
  st->n_iters = c.n_iters;
  st->n_accepted = c.n_accepted;
  st->status = c.status;
  st->last_dt = c.dt;
  st->last_t = c.t;
  st->n_launches = launches;
  st->n_rhs = n_rhs;
}

}  // namespace

extern "C" {

int32_t eno_abi_version(void) { return ENO_ABI_VERSION; }

const char* eno_last_error_string(const eno_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int32_t eno_create(eno_ctx** out, int32_t device) {
  if (!out) return fail(nullptr, ENO_E_ARG, "eno_create: out is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(nullptr, ENO_E_CUDA, std::string("eno_create: no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, ENO_E_ARG, "eno_create: bad device index");
  eno_ctx* c = new eno_ctx();
  c->device = device;
  ENO_CUDA(c, cudaSetDevice(device));
  ENO_CUDA(c, cudaMallocHost(&c->mailbox, sizeof(Ctrl)));
  ENO_CUDA(c, cudaMallocHost(&c->tabs, sizeof(Tableau) * (ENO_NUM_TABLEAUX + 1)));
  for (int i = 0; i < ENO_NUM_TABLEAUX; ++i) fill_tableau(i, &c->tabs[i]);
  fill_rk4_grid(&c->tabs[ENO_NUM_TABLEAUX]);   // odeint_grid (css_engine.cu)
  cudaDeviceProp prop;
  ENO_CUDA(c, cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    std::string m = "eno_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    ", this library is built for sm_100a only";
    cudaFreeHost(c->mailbox);
    delete c;
    return fail(nullptr, ENO_E_UNSUPPORTED, m);
  }
  if (const char* pm = getenv("ENO_PATH")) {
    if (!strcmp(pm, "stream")) c->path_mode = 1;
    if (!strcmp(pm, "cluster")) c->path_mode = 2;
  }
  *out = c;
  return ENO_OK;
}

void eno_destroy(eno_ctx* ctx) {
  if (!ctx) return;
  css::destroy(ctx);
  if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
  if (ctx->tabs) cudaFreeHost(ctx->tabs);
  delete ctx;
}

int64_t eno_param_count(const eno_dynamics_desc* d) {
  if (!d) return -1;
  if ((d->arch == ENO_ARCH_CONCATSQUASH || d->arch == ENO_ARCH_CONCATCONV)) return css::desc_ok(d) ? css::param_count(d) : -1;
  return (int64_t)(d->D + 1) * d->H + d->H + (int64_t)(d->H + 1) * d->D + d->D;
}

int64_t eno_adjoint_state_size(const eno_dynamics_desc* d, int32_t B) {
  if (!d) return -1;
  if ((d->arch == ENO_ARCH_CONCATSQUASH || d->arch == ENO_ARCH_CONCATCONV)) return css::desc_ok(d) ? css::adjoint_state_size(d, B) : -1;
  const int64_t n_aux = (d->aug == ENO_AUG_NONE) ? 0 : (d->aug == ENO_AUG_FIN ? 2 : 1);
  return 2 * ((int64_t)B * d->D + n_aux * B) + 1 + (d->eps_in_args ? (int64_t)B * d->D : 0) + eno_param_count(d);
}

size_t eno_workspace_bytes(const eno_dynamics_desc* d, int32_t B, int32_t T, int32_t tableau, int32_t mode) {
  (void)tableau;
  if (d && (d->arch == ENO_ARCH_CONCATSQUASH || d->arch == ENO_ARCH_CONCATCONV)) return css::workspace_bytes(d, B, T, mode);
  if (!desc_ok(d) || B <= 0) return 0;
  if (mode == 0) return fwd_ws_bytes(B, d->D);
  return adj_ws_bytes(B, d->D, d->H, g_world, d->aug == ENO_AUG_FIN, slot_of(B));
}

int32_t eno_odeint_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t tableau, int32_t B, const float* y0,
                       const float* ts, int32_t T, const float* params, const float* eps, float rtol, float atol,
                       int32_t mxstep, void* workspace, size_t workspace_bytes, float* ys_out, float* aux_out,
                       eno_stats* stats_host, eno_trace_row* trace_out, int32_t trace_cap, void* stream_) {
  if (ctx && desc && (desc->arch == ENO_ARCH_CONCATSQUASH || desc->arch == ENO_ARCH_CONCATCONV)) {
    if (tableau != ENO_TABLEAU_DOPRI5)
      return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: the ConcatSquash dynamics run dopri5 only (the tableau sweep is defined on the MNIST dynamics)");
    return css::odeint_fwd(ctx, desc, B, y0, ts, T, params, eps, rtol, atol, mxstep, workspace, workspace_bytes, ys_out, aux_out,
                           stats_host, trace_out, trace_cap, static_cast<cudaStream_t>(stream_));
  }
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_odeint_fwd: ctx is null");
  if (!desc_ok(desc)) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: unsupported dynamics desc (MLP sigmoid, D%4==0, H%4==0)");
  const int aug = desc->aug;
  const int n_aux = aug == ENO_AUG_ALL ? 3 : (aug == ENO_AUG_REG ? 1 : 0);
  if (aug != ENO_AUG_NONE) {
    if (aug != ENO_AUG_REG && aug != ENO_AUG_ALL) return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: bad desc->aug");
    if (tableau != ENO_TABLEAU_DOPRI5) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: augmented states integrate with dopri5 only (odeint is always dopri5, lib/ode.py:746)");
    if (!(desc->reg_order == 0 || desc->reg_order == 2 || desc->reg_order == 3)) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: reg_order must be 0, 2 or 3");
    if (!aux_out || (aug == ENO_AUG_ALL && !eps)) return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: augmented solve needs aux_out (and eps for ENO_AUG_ALL)");
    if (ctx->path_mode == 1 || !cl_supported(desc->D, desc->H, kClRows, kAdjStages, true))
      return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd: the augmented forward solve exists on the cluster-resident path only (weights must fit the cluster's shared memory)");
  }
  if (B <= 0 || T < 1 || !y0 || !ts || !params || !ys_out || !workspace)
    return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: null pointer or empty batch");
  if (((uintptr_t)params | (uintptr_t)y0 | (uintptr_t)ys_out | (uintptr_t)workspace) & 15)
    return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: buffers must be 16-byte aligned");
  if (workspace_bytes < fwd_ws_bytes(B, desc->D)) return fail(ctx, ENO_E_WORKSPACE, "eno_odeint_fwd: workspace too small");
  Tableau tab;
  if (!fill_tableau(tableau, &tab)) return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: unknown tableau id");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));

  const int D = desc->D, H = desc->H;
  const size_t N = (size_t)B * D;
  Carver cv(workspace);
  FwdArgs a;
  a.ctrl = cv.take<Ctrl>(1);
  a.Y = cv.take<float>(3 * N);
  a.F = cv.take<float>(3 * N);
  a.MID = cv.take<float>(2 * N);
  a.rowpart0 = cv.take<double>((size_t)kC * slot_of(B) * g_world);
  a.rowpart1 = cv.take<double>((size_t)kC * slot_of(B) * g_world);
  AugFwdArgs A;
  A.RSa = cv.take<float>((size_t)9 * B);
  A.FRa = cv.take<float>((size_t)9 * B);
  A.MIDa = cv.take<float>((size_t)9 * B);
  A.eps = eps;
  A.aux_out = aux_out;
  A.n_aux = n_aux;
  a.w = mlp_weights_from_flat(params, D, H, nullptr, nullptr);
  a.B = B;
  a.N = N;
  a.y0 = y0;
  a.ts = ts;
  a.ys_out = ys_out;
  a.trace = trace_out;

  const DistCtx& dc = ctx->dist;
  const bool dist = dc.active();
  const int world = dist ? dc.world : 1;
  const bool use_cluster = aug != ENO_AUG_NONE ||
                           (ctx->path_mode != 1 && cl_supported(D, H, kClRows, tab.s < 2 ? 2 : tab.s, false));
  const int stride = use_cluster ? kC : 1;           // partial sums per sample
  if (dist && B > dc.slot(B)) return fail(ctx, ENO_E_ARG, "eno_odeint_fwd: this rank's batch exceeds the shard slot (eno_comm_set_shards)");
  const size_t n_loc = (size_t)stride * dc.slot(B);  // this rank's slot of partial sums (stride * B of them written, the rest zero)
  a.n_rp = (int)(n_loc * world);
  a.dist = dist ? 1 : 0;
  a.rp0_all = a.rowpart0;
  a.rp1_all = a.rowpart1;
  a.rowpart0 += (size_t)(dist ? dc.rank : 0) * n_loc;
  a.rowpart1 += (size_t)(dist ? dc.rank : 0) * n_loc;
  if (dist && dc.slot(B) != B) {   // uneven shards: the unused tail of this rank's slot must read as zero
    ENO_CUDA(ctx, cudaMemsetAsync(a.rowpart0, 0, n_loc * 8, static_cast<cudaStream_t>(stream_)));
    ENO_CUDA(ctx, cudaMemsetAsync(a.rowpart1, 0, n_loc * 8, static_cast<cudaStream_t>(stream_)));
  }
  double* rp0 = const_cast<double*>(a.rp0_all);
  double* rp1 = const_cast<double*>(a.rp1_all);

  const int R = pick_rows_fwd(B, D, H, tab.s, ctx->sm_count);
  const size_t smem = aug != ENO_AUG_NONE ? cl_smem_bytes(kClRows, D, H, kAdjStages, true)
                      : use_cluster       ? cl_smem_bytes(kClRows, D, H, tab.s < 2 ? 2 : tab.s, false)
                                          : fwd_smem_bytes(R, D, H, tab.s < 2 ? 2 : tab.s);
  const int grid = use_cluster ? cl_grid(B) : (B + R - 1) / R;
  int launches = 0;

  ENO_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tab, &ctx->tabs[tableau], sizeof(Tableau), 0, cudaMemcpyHostToDevice, stream));
  const bool deferred = ctx->deferred && ctx->spec_fwd > 0;
  k_ctrl_reset<<<1, 1, 0, stream>>>(a.ctrl, rtol, atol, T, mxstep <= 0 ? INT_MAX : mxstep, trace_out ? trace_cap : 0,
                                    (float)((size_t)(D + n_aux) * (size_t)dc.rows(B)));
  ++launches;

  auto launch = [&](int mode) -> cudaError_t {
    ++launches;
    if (aug != ENO_AUG_NONE) {
      A.f = a;
      return cl_aug_fwd_launch(desc->reg_order, aug, mode, A, grid, smem, stream);
    }
    if (use_cluster) return cl_fwd_launch(mode, a, grid, smem, stream);
    if (R == 1) return fwd_stream_launch<1>(mode, a, grid, smem, stream);
    if (R == 2) return fwd_stream_launch<2>(mode, a, grid, smem, stream);
    return fwd_stream_launch<4>(mode, a, grid, smem, stream);
  };
  // multi-GPU: gather every rank's per-sample partials, then one CTA per rank runs the controller
  const bool peer = dist && dc.peer_ready() && (size_t)4 * a.n_rp * 8 <= kPeerOffAdj - kPeerOffFwd;
  auto finish = [&](int mode) -> int {
    if (!dist) return 0;
    if (peer) {   // exchange inside the finish kernel, over NVLink peer memory (dist.cuh)
      k_fwd_finish_peer<<<1, kThreads, 0, stream>>>(a, mode, dc.peer.base[dc.rank], (int)n_loc);
      ++launches;
      return 0;
    }
    if (dc.api.GroupStart()) return 1;
    int rc = dc.gather(rp0, n_loc, stream);
    if (mode == 1) rc |= dc.gather(rp1, n_loc, stream);
    rc |= dc.api.GroupEnd();
    if (rc) return rc;
    k_fwd_finish<<<1, kThreads, 0, stream>>>(a, mode);
    ++launches;
    return 0;
  };
  ENO_CUDA(ctx, launch(1));
  if (finish(1)) return fail(ctx, ENO_E_CUDA, "eno_odeint_fwd: NCCL all-gather failed");
  ENO_CUDA(ctx, launch(2));
  if (finish(2)) return fail(ctx, ENO_E_CUDA, "eno_odeint_fwd: NCCL all-gather failed");
  {
    int chunk = deferred ? ctx->spec_fwd : ctx->hint_fwd;
    const int out_grid = (int)((N + 1023) / 1024) < 4 * ctx->sm_count ? (int)((N + 1023) / 1024) : 4 * ctx->sm_count;
    for (;;) {
      for (int i = 0; i < chunk; ++i) {
        const int ps_ = ctx->prof.begin(PROF_FWD_STEP, stream);
        ENO_CUDA(ctx, launch(0));
        ctx->prof.end(ps_, stream);
        const int px_ = dist ? ctx->prof.begin(PROF_EXCHANGE, stream) : -1;
        if (finish(0)) return fail(ctx, ENO_E_CUDA, "eno_odeint_fwd: NCCL all-gather failed");
        ctx->prof.end(px_, stream);
        // with several output times a step can leave targets behind mid-solve: evaluate them before
        // the next step recycles the ring; with a single target one output launch per chunk suffices
        if (T > 2) {
          k_fwd_out<<<out_grid, 256, 0, stream>>>(a); ++launches;
          if (n_aux) { A.f = a; k_aux_out<<<(n_aux * B + 255) / 256, 256, 0, stream>>>(A); ++launches; }
        }
      }
      if (T <= 2) {
        k_fwd_out<<<out_grid, 256, 0, stream>>>(a); ++launches;
        if (n_aux) { A.f = a; k_aux_out<<<(n_aux * B + 255) / 256, 256, 0, stream>>>(A); ++launches; }
      }
      if (deferred) break;   // graph-capturable: no host round-trip, Ctrl is inspected later (eno_read_stats)
      ENO_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, a.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
      ENO_CUDA(ctx, cudaStreamSynchronize(stream));
      if (ctx->mailbox->done) break;
      chunk = 4;
    }
  }

  ENO_CUDA(ctx, cudaGetLastError());
  if (deferred) {
    if (stats_host) { memset(stats_host, 0, sizeof(eno_stats)); stats_host->status = -100; stats_host->n_launches = launches; }
    return ENO_OK;
  }
  const Ctrl& c = *ctx->mailbox;
  {
    int valid[PROF_KINDS] = {c.n_iters, 0, 0, dist ? c.n_iters : 0, 0};
    ctx->prof.resolve(valid);
  }
  ctx->hint_fwd = c.n_iters + 1 > 64 ? 64 : c.n_iters + 1;
  const int per_iter = tab.clip ? 2 * (tab.s - 1) : (tab.s - 1);
  (void)eps;
  fill_stats(stats_host, c, launches, 2 + per_iter * c.n_iters);
  return c.status;
}

int32_t eno_odeint_fwd_vmap(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* y0, const float* ts,
                            int32_t T, const float* params, float rtol, float atol, int32_t mxstep, float* ys_out,
                            float* nfe_out, int32_t* iters_out, int32_t* status_out, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_odeint_fwd_vmap: ctx is null");
  if (!desc_ok(desc)) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd_vmap: unsupported dynamics desc (MLP sigmoid, D%4==0, H%4==0)");
  if (desc->aug != ENO_AUG_NONE) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd_vmap: plain dynamics only (mnist.py:350-353 vmaps odeint(dynamics_wrap, ...))");
  if (T != 2) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd_vmap: T must be 2 (ts = [t0, t1])");
  if (B <= 0 || !y0 || !ts || !params || !ys_out || !nfe_out) return fail(ctx, ENO_E_ARG, "eno_odeint_fwd_vmap: null pointer or empty batch");
  if (ctx->path_mode == 1 || !cl_supported(desc->D, desc->H, kClRows, kAdjStages, false))
    return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_fwd_vmap: runs on the cluster-resident path only (weights must fit the cluster's shared memory)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  ENO_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tab, &ctx->tabs[ENO_TABLEAU_DOPRI5], sizeof(Tableau), 0, cudaMemcpyHostToDevice, stream));
  VmapArgs a = {};
  a.w = mlp_weights_from_flat(params, desc->D, desc->H, nullptr, nullptr);
  a.B = B;
  a.y0 = y0;
  a.ts = ts;
  a.rtol = rtol;
  a.atol = atol;
  a.mxstep = mxstep <= 0 ? INT_MAX : mxstep;
  a.ys_out = ys_out;
  a.nfe_out = nfe_out;
  a.iters_out = iters_out;
  a.status_out = status_out;
  ENO_CUDA(ctx, cl_fwd_vmap_launch(a, cl_grid(B), cl_smem_bytes(kClRows, desc->D, desc->H, kAdjStages, false), stream));
  ENO_CUDA(ctx, cudaGetLastError());
  ENO_CUDA(ctx, cudaStreamSynchronize(stream));
  return ENO_OK;
}

int32_t eno_rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t mode, int32_t B, const float* y, float t,
                const float* params, const float* eps, const float* y_bar, const float* r_bar, void* workspace,
                size_t workspace_bytes, float* dy_out, float* aux_out, float* ybar_dot_out, float* tbar_out,
                float* params_bar_out, float* eps_bar_dot_out, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_rhs: ctx is null");
  if (desc && (desc->arch == ENO_ARCH_CONCATSQUASH || desc->arch == ENO_ARCH_CONCATCONV))
    return css::rhs(ctx, desc, mode, B, y, t, params, eps, y_bar, r_bar, workspace, workspace_bytes, dy_out, aux_out, ybar_dot_out,
                    tbar_out, params_bar_out, eps_bar_dot_out, static_cast<cudaStream_t>(stream_));
  if (!desc_ok(desc)) return fail(ctx, ENO_E_UNSUPPORTED, "eno_rhs: unsupported dynamics desc");
  if (B <= 0 || !y || !params || !dy_out) return fail(ctx, ENO_E_ARG, "eno_rhs: null pointer or empty batch");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int D = desc->D, H = desc->H;
  if (mode == 0) {
    const size_t smem = fwd_smem_bytes(1, D, H, 2);
    ENO_CUDA(ctx, set_smem(k_rhs_plain<1>, smem));
    k_rhs_plain<1><<<B, kThreads, smem, stream>>>(mlp_weights_from_flat(params, D, H, nullptr, nullptr), B, y, t, dy_out);
    ENO_CUDA(ctx, cudaGetLastError());
    ENO_CUDA(ctx, cudaStreamSynchronize(stream));
    return ENO_OK;
  }
  return adj_rhs(ctx->path_mode, &ctx->err, ctx->sm_count, desc, mode, B, y, t, params, eps, y_bar, r_bar, workspace,
                 workspace_bytes, dy_out, aux_out, ybar_dot_out, tbar_out, params_bar_out, eps_bar_dot_out, stream);
}

int32_t eno_odeint_grid_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* y0, const float* ts, int32_t T,
                            const float* params, const float* eps, float step_size, void* workspace, size_t workspace_bytes,
                            float* ys_out, float* aux_out, eno_stats* stats_host, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_odeint_grid_fwd: ctx is null");
  if (!desc || !(desc->arch == ENO_ARCH_CONCATSQUASH || desc->arch == ENO_ARCH_CONCATCONV))
    return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_grid_fwd: the fixed-grid solver runs on the ConcatSquash / ConcatConv2D engine");
  if (!(step_size > 0.f)) return fail(ctx, ENO_E_ARG, "eno_odeint_grid_fwd: step_size must be positive");
  return css::odeint_fwd(ctx, desc, B, y0, ts, T, params, eps, 1.f, 1.f, 0, workspace, workspace_bytes, ys_out, aux_out, stats_host,
                         nullptr, 0, static_cast<cudaStream_t>(stream_), step_size);
}

int32_t eno_odeint_grid_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts, int32_t T,
                            const float* params, const float* eps, const float* g_y, const float* g_aux, float step_size,
                            void* workspace, size_t workspace_bytes, float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out,
                            float* ts_bar_out, float* params_bar_out, eno_stats* stats_host, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_odeint_grid_bwd: ctx is null");
  if (!desc || !(desc->arch == ENO_ARCH_CONCATSQUASH || desc->arch == ENO_ARCH_CONCATCONV))
    return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_grid_bwd: the fixed-grid solver runs on the ConcatSquash / ConcatConv2D engine");
  if (!(step_size > 0.f)) return fail(ctx, ENO_E_ARG, "eno_odeint_grid_bwd: step_size must be positive");
  return css::odeint_bwd(ctx, desc, B, ys, ts, T, params, eps, g_y, g_aux, 1.f, 1.f, 0, workspace, workspace_bytes, y0_bar_out,
                         aux0_bar_out, eps_bar_out, ts_bar_out, params_bar_out, stats_host, nullptr, 0,
                         static_cast<cudaStream_t>(stream_), step_size);
}

int32_t eno_odeint_bwd_sepaux(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts,
                              int32_t T, const float* params, const float* eps, const float* g_y, const float* g_aux,
                              float rtol, float atol, int32_t mxstep, void* workspace, size_t workspace_bytes,
                              float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out, float* ts_bar_out,
                              float* params_bar_out, eno_stats* stats_host, eno_trace_row* trace_out, int32_t trace_cap,
                              void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_odeint_bwd: ctx is null");
  if (desc && (desc->arch == ENO_ARCH_CONCATSQUASH || desc->arch == ENO_ARCH_CONCATCONV))
    return css::odeint_bwd(ctx, desc, B, ys, ts, T, params, eps, g_y, g_aux, rtol, atol, mxstep, workspace, workspace_bytes,
                           y0_bar_out, aux0_bar_out, eps_bar_out, ts_bar_out, params_bar_out, stats_host, trace_out, trace_cap,
                           static_cast<cudaStream_t>(stream_));
  if (!desc_ok(desc)) return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_bwd: unsupported dynamics desc");
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (desc->aug != ENO_AUG_NONE && desc->aug != ENO_AUG_REG && desc->aug != ENO_AUG_FIN)
    return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_bwd: aug must be ENO_AUG_NONE, ENO_AUG_REG or ENO_AUG_FIN");
  return adj_solve(&ctx->prof, ctx->path_mode, &ctx->dist, (ctx->deferred ? ctx->spec_bwd : 0), &ctx->tabs[ENO_TABLEAU_DOPRI5],
                   &ctx->err, ctx->mailbox, &ctx->hint_bwd, ctx->sm_count, desc, B, ys, ts, T, params, eps, g_y, g_aux, rtol,
                   atol, mxstep, workspace, workspace_bytes, y0_bar_out, aux0_bar_out, eps_bar_out, ts_bar_out,
                   params_bar_out, stats_host, trace_out, trace_cap, static_cast<cudaStream_t>(stream_));
}

int32_t eno_odeint_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts,
                       int32_t T, const float* params, const float* g_y, const float* g_r, float rtol, float atol,
                       int32_t mxstep, void* workspace, size_t workspace_bytes, float* y0_bar_out,
                       float* r0_bar_out, float* ts_bar_out, float* params_bar_out, eno_stats* stats_host,
                       eno_trace_row* trace_out, int32_t trace_cap, void* stream_) {
  if (desc && desc->aug == ENO_AUG_FIN)
    return fail(ctx, ENO_E_UNSUPPORTED, "eno_odeint_bwd: ENO_AUG_FIN reads eps; call eno_odeint_bwd_sepaux");
  return eno_odeint_bwd_sepaux(ctx, desc, B, ys, ts, T, params, nullptr, g_y, g_r, rtol, atol, mxstep, workspace,
                               workspace_bytes, y0_bar_out, r0_bar_out, nullptr, ts_bar_out, params_bar_out, stats_host,
                               trace_out, trace_cap, stream_);
}

// ---- the rest of mnist.py's update() around the ODE block (head_kernels.cuh) ----------------------
int32_t eno_mnist_head(eno_ctx* ctx, int32_t B, int32_t D, int32_t C, const float* y1, const int64_t* labels, const float* W,
                       const float* b, float inv_count, float* scratch, float* loss_out, float* gy_out, float* gW_out,
                       float* gb_out, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_mnist_head: ctx is null");
  if (B <= 0 || D <= 0 || C <= 0 || C > kHeadMaxClasses) return fail(ctx, ENO_E_ARG, "eno_mnist_head: bad sizes (classes <= 32)");
  if (!y1 || !labels || !W || !b || !scratch || !loss_out || !gy_out || !gW_out || !gb_out)
    return fail(ctx, ENO_E_ARG, "eno_mnist_head: null pointer");
  if ((size_t)D * 4 > 48 * 1024) return fail(ctx, ENO_E_UNSUPPORTED, "eno_mnist_head: D too large for the row kernel");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  float* s = scratch;                       // [B][D]
  float* dl = scratch + (size_t)B * D;      // [B][C]
  float* loss_rows = dl + (size_t)B * C;    // [B]
  k_head_rows<<<B, 256, (size_t)D * 4, stream>>>(y1, labels, W, b, D, C, inv_count, s, dl, loss_rows, gy_out);
  const int n = D * C + C + 1;   // one warp per output
  k_head_wgrad<<<(n + 7) / 8, 256, 0, stream>>>(s, dl, loss_rows, B, D, C, gW_out, gb_out, loss_out);
  ENO_CUDA(ctx, cudaGetLastError());
  return ENO_OK;
}

int32_t eno_momentum_update(eno_ctx* ctx, int32_t n_blocks, float* const* params, float* const* velocity,
                            const float* const* grads, const int64_t* sizes, float lr, float mass, float lam_w,
                            const float* lr_dev, const int32_t* ok_a, const int32_t* ok_b, void* stream_) {
  if (!ctx) return fail(nullptr, ENO_E_ARG, "eno_momentum_update: ctx is null");
  if (n_blocks < 1 || n_blocks > 3 || !params || !velocity || !grads || !sizes)
    return fail(ctx, ENO_E_ARG, "eno_momentum_update: 1..3 parameter blocks");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  MomentumBlocks m = {};
  long long total = 0;
  for (int k = 0; k < n_blocks; ++k) {
    m.p[k] = params[k]; m.v[k] = velocity[k]; m.g[k] = grads[k]; m.n[k] = sizes[k];
    total = sizes[k] > total ? sizes[k] : total;
  }
  const int grid = (int)((total + 255) / 256 < 4LL * ctx->sm_count ? (total + 255) / 256 : 4LL * ctx->sm_count);
  k_momentum<<<grid > 0 ? grid : 1, 256, 0, stream>>>(m, lr, mass, lam_w, lr_dev, ok_a, ok_b);
  ENO_CUDA(ctx, cudaGetLastError());
  return ENO_OK;
}

// ---- deferred launch mode (CUDA-graph capturable solves) ----------------------------------------
int32_t eno_set_deferred(eno_ctx* ctx, int32_t on, int32_t fwd_iters, int32_t bwd_iters) {
  if (!ctx) return ENO_E_ARG;
  if (on && (fwd_iters <= 0 || bwd_iters <= 0)) return fail(ctx, ENO_E_ARG, "eno_set_deferred: iteration budgets must be positive");
  ctx->deferred = on ? 1 : 0;
  ctx->spec_fwd = fwd_iters;
  ctx->spec_bwd = bwd_iters;
  return ENO_OK;
}

int32_t eno_read_stats(eno_ctx* ctx, const void* workspace, eno_stats* out, void* stream_) {
  if (!ctx || !workspace || !out) return fail(ctx, ENO_E_ARG, "eno_read_stats: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  // the controller block is the first object of both the forward and the backward workspace layout
  ENO_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, workspace, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
  ENO_CUDA(ctx, cudaStreamSynchronize(stream));
  const Ctrl& c = *ctx->mailbox;
  fill_stats(out, c, 0, 2 + 6 * c.n_iters);
  if (!c.done) out->status = -101;   // the iteration budget of the deferred launch was too small
  return ENO_OK;
}

// ---- multi-GPU (one process per GPU) -----------------------------------------------------------
int32_t eno_comm_unique_id(eno_ctx* ctx, uint8_t* id_out, const char* nccl_path) {
  if (!ctx || !id_out) return fail(ctx, ENO_E_ARG, "eno_comm_unique_id: null argument");
  if (!ctx->dist.api.load(nccl_path, &ctx->err)) return ENO_E_UNSUPPORTED;
  NcclUniqueId id;
  if (int rc = ctx->dist.api.GetUniqueId(&id))
    return fail(ctx, ENO_E_CUDA, std::string("ncclGetUniqueId: ") + ctx->dist.api.GetErrorString(rc));
  memcpy(id_out, id.internal, 128);
  return ENO_OK;
}

int32_t eno_comm_init(eno_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world, const char* nccl_path) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, ENO_E_ARG, "eno_comm_init: bad arguments");
  if (world == 1) { ctx->dist.world = 1; ctx->dist.rank = 0; return ENO_OK; }
  if (!ctx->dist.api.load(nccl_path, &ctx->err)) return ENO_E_UNSUPPORTED;
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  css::drop_collective_graphs(ctx);
  NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  if (int rc = ctx->dist.api.CommInitRank(&ctx->dist.comm, world, uid, rank))
    return fail(ctx, ENO_E_CUDA, std::string("ncclCommInitRank: ") + ctx->dist.api.GetErrorString(rc));
  ctx->dist.world = world;
  ctx->dist.rank = rank;
  g_world = world;
  return ENO_OK;
}

int32_t eno_comm_set_shards(eno_ctx* ctx, int32_t slot_rows, int64_t rows_global) {
  if (!ctx || slot_rows < 0 || rows_global < 0) return fail(ctx, ENO_E_ARG, "eno_comm_set_shards: bad arguments");
  if (slot_rows > 0 && rows_global > (int64_t)slot_rows * ctx->dist.world)
    return fail(ctx, ENO_E_ARG, "eno_comm_set_shards: rows_global exceeds world * slot_rows");
  ctx->dist.slot_rows = slot_rows;
  ctx->dist.rows_global = rows_global;
  g_slot = slot_rows;
  return ENO_OK;
}

// Exchange over NVLink peer memory (dist.cuh): every rank allocates its symmetric buffer and exports it ...
int32_t eno_peer_export(eno_ctx* ctx, uint64_t bytes, uint8_t* handle_out /*[64]*/) {
  if (!ctx || !handle_out || bytes < (kPeerOffAdj << 1)) return fail(ctx, ENO_E_ARG, "eno_peer_export: bad arguments");
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->dist.peer_mine) return fail(ctx, ENO_E_ARG, "eno_peer_export: already exported");
  void* p = nullptr;
  ENO_CUDA(ctx, cudaMalloc(&p, bytes));
  ENO_CUDA(ctx, cudaMemset(p, 0, bytes));
  ENO_CUDA(ctx, cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(ctx, ENO_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  ctx->dist.peer_mine = p;
  ctx->dist.peer.bytes = bytes;
  return ENO_OK;
}

// ... and maps every peer's (handles of all ranks in rank order, gathered by the host side).
int32_t eno_peer_attach(eno_ctx* ctx, const uint8_t* handles /*[world][64]*/) {
  if (!ctx || !handles || !ctx->dist.active() || !ctx->dist.peer_mine) return fail(ctx, ENO_E_ARG, "eno_peer_attach: call eno_comm_init and eno_peer_export first");
  if (ctx->dist.world > kPeerMax) return fail(ctx, ENO_E_UNSUPPORTED, "eno_peer_attach: at most 8 ranks");
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  PeerView pv;
  pv.world = ctx->dist.world; pv.rank = ctx->dist.rank; pv.bytes = ctx->dist.peer.bytes;
  for (int r = 0; r < pv.world; ++r) {
    if (r == pv.rank) { pv.base[r] = static_cast<unsigned char*>(ctx->dist.peer_mine); continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(ctx, ENO_E_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
    pv.base[r] = static_cast<unsigned char*>(p);
  }
  ctx->dist.peer = pv;
  // the kernels read the view from the buffer itself (keeps their parameter blocks small)
  ENO_CUDA(ctx, cudaMemcpy(pv.base[pv.rank] + kPeerOffView, &pv, sizeof(PeerView), cudaMemcpyHostToDevice));
  return ENO_OK;
}

int32_t eno_comm_destroy(eno_ctx* ctx) {
  if (!ctx) return ENO_E_ARG;
  css::drop_collective_graphs(ctx);
  if (ctx->dist.comm) ctx->dist.api.CommDestroy(ctx->dist.comm);
  ctx->dist.comm = nullptr;
  for (int r = 0; r < ctx->dist.peer.world; ++r)
    if (r != ctx->dist.peer.rank && ctx->dist.peer.base[r]) cudaIpcCloseMemHandle(ctx->dist.peer.base[r]);
  if (ctx->dist.peer_mine) cudaFree(ctx->dist.peer_mine);
  ctx->dist.peer_mine = nullptr;
  ctx->dist.peer = PeerView();
  ctx->dist.world = 1;
  ctx->dist.rank = 0;
  ctx->dist.slot_rows = 0;
  ctx->dist.rows_global = 0;
  g_world = 1;
  g_slot = 0;
  return ENO_OK;
}

// ---- measurement hooks (bench.py) ------------------------------------------------------------
int32_t eno_profile_enable(eno_ctx* ctx, int32_t on) {
  if (!ctx) return ENO_E_ARG;
  ctx->prof.enabled = on != 0;
  return ENO_OK;
}

int32_t eno_profile_read(eno_ctx* ctx, double* ms_out, int64_t* count_out) {
  if (!ctx || !ms_out || !count_out) return ENO_E_ARG;
  for (int k = 0; k < PROF_KINDS; ++k) {
    ms_out[k] = ctx->prof.ms[k];
    count_out[k] = ctx->prof.count[k];
    ctx->prof.ms[k] = 0;
    ctx->prof.count[k] = 0;
  }
  return ENO_OK;
}

int32_t eno_ffma_peak(eno_ctx* ctx, double* tflops_out, void* stream_) {
  if (!ctx || !tflops_out) return ENO_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  ENO_CUDA(ctx, cudaSetDevice(ctx->device));
  float* sink = nullptr;
  ENO_CUDA(ctx, cudaMalloc(&sink, 4));
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    k_ffma_peak<<<blocks, threads, 0, st>>>(sink, iters, 1.0001f);
    cudaEventRecord(e1, st);
    ENO_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 16 * iters * (double)blocks * threads;
    if (ms > 0 && fl / ms * 1e-9 > best) best = fl / ms * 1e-9;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops_out = best;
  return ENO_OK;
}

#ifdef ENO_PHASES
int32_t eno_debug_peer_stamps(long long* out /*[8]*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_peer_stamp, sizeof(long long) * 8);
  return 0;
}
// development aid: read and clear the phase clocks of the cluster kernels
int32_t eno_debug_phases(long long* cycles, long long* counts) {
  long long zero[32] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(cycles, g_phase_cycles, sizeof(long long) * PH_NKINDS);
  (void)sizeof(zero);
  cudaMemcpyFromSymbol(counts, g_phase_count, sizeof(long long) * PH_NKINDS);
  cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero));
  cudaMemcpyToSymbol(g_phase_count, zero, sizeof(zero));
  return PH_NKINDS;
}
#endif

}  // extern "C"
