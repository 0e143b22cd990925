// Third translation unit of libeno.so: per-trajectory solves of small tanh stacks (ENO_ARCH_TANH_STACK, the generative
// ODE of latent_ode.py, BASELINE config C3).  Host side of lat_kernels.cuh: descriptor -> network geometry, shared-memory
// sizing, launches.  Every entry point only enqueues work on the stream (no synchronisation, CUDA-graph capturable).
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "eno_ctx.cuh"
#include "lat_kernels.cuh"

using namespace eno;

namespace {

int lat_fail(eno_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define LAT_CUDA(ctx, expr)                                                                                       \
  do {                                                                                                            \
    cudaError_t e_ = (expr);                                                                                      \
    if (e_ != cudaSuccess) return lat_fail(ctx, ENO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

bool lat_desc_ok(const eno_dynamics_desc* d, std::string* why) {
  if (!d || d->arch != ENO_ARCH_TANH_STACK) { *why = "descriptor is not ENO_ARCH_TANH_STACK"; return false; }
  if (d->D < 1 || d->H < 1 || d->D > lat::kMaxW || d->H > lat::kMaxW) { *why = "widths must be in [1, 64]"; return false; }
  if (d->n_hidden < 1 || d->n_hidden + 1 > lat::kMaxL) { *why = "n_hidden must be in [1, 7]"; return false; }
  if (d->reg_order != 0 && d->reg_order != 2 && d->reg_order != 3) { *why = "reg_order must be 0, 2 or 3"; return false; }
  if (d->aug != ENO_AUG_NONE && d->aug != ENO_AUG_REG) { *why = "aug must be ENO_AUG_NONE or ENO_AUG_REG"; return false; }
  return true;
}

lat::Net make_net(const eno_dynamics_desc* d) {
  lat::Net n;
  memset(&n, 0, sizeof(n));
  n.L = d->n_hidden + 1;
  n.K = d->reg_order;
  n.aug = d->aug == ENO_AUG_REG ? 1 : 0;
  n.d[0] = d->D;
  for (int l = 1; l < n.L; ++l) n.d[l] = d->H;
  n.d[n.L] = d->D;
  int w = 0, p = 0, in = 0, out = 0, big = 0;
  for (int l = 0; l < n.L; ++l) {
    const int din = n.d[l], dout = n.d[l + 1];
    n.ld[l] = dout | 1;
    n.wo[l] = w; w += din * n.ld[l];
    n.po[l] = p; p += dout + din * dout;
    n.io[l] = in; in += din;
    n.oo[l] = out; out += dout;
    big = std::max(big, din + 1 + dout);
  }
  for (int l = 0; l < n.L; ++l) { n.bo[l] = w; w += n.d[l + 1]; }
  n.io[n.L] = in; n.oo[n.L] = out;
  n.IN = in; n.OUT = out; n.P = p; n.wtot = w;
  n.FS = 3 * in + 3 * out;
  n.scratch = std::max(10 * in + 3 * out, lat::kNC * big);
  return n;
}

template <typename R>
size_t smem_bytes(const lat::Net& n, int solver_elems) {
  return lat::smem_elems<R>(n, solver_elems) * sizeof(R) + 16;
}

int fwd_solver_elems(const lat::Net& n) { const int ns = n.d[0] + n.aug; return ns * (1 + lat::kStages + 5 + 1); }
int bwd_solver_elems(const lat::Net& n) { const int nsm = 2 * (n.d[0] + n.aug) + 1; return nsm * (1 + lat::kStages) + 8; }

template <typename R>
int fwd_t(eno_ctx* ctx, const lat::Net& n, int n_traj, const void* z0, const void* ts, int T, const void* params, double rtol,
          double atol, int mxstep, void* zs_out, void* rs_out, void* nfe_out, int* iters_out, int* status_out, cudaStream_t st) {
  const size_t sm = smem_bytes<R>(n, fwd_solver_elems(n));
  LAT_CUDA(ctx, cudaFuncSetAttribute(lat::k_traj_fwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  lat::k_traj_fwd<R><<<n_traj, lat::kThreads, sm, st>>>(n, static_cast<const R*>(z0), static_cast<const R*>(ts), T,
                                                       static_cast<const R*>(params), (R)rtol, (R)atol, mxstep,
                                                       static_cast<R*>(zs_out), static_cast<R*>(rs_out),
                                                       static_cast<R*>(nfe_out), iters_out, status_out);
  LAT_CUDA(ctx, cudaGetLastError());
  return ENO_OK;
}

template <typename R>
int bwd_t(eno_ctx* ctx, const lat::Net& n, int n_traj, const void* zs, const void* rs, const void* ts, int T, const void* params,
          const void* g_z, const void* g_r, double rtol, double atol, int mxstep, void* ws, void* z0_bar_out, void* r0_bar_out,
          void* ts_bar_out, void* params_bar_out, void* params_bar_traj_out, int* iters_out, int* status_out, cudaStream_t st) {
  const size_t sm = smem_bytes<R>(n, bwd_solver_elems(n));
  LAT_CUDA(ctx, cudaFuncSetAttribute(lat::k_traj_bwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  R* pb = static_cast<R*>(ws);
  R* fact = pb + (size_t)n_traj * 2 * n.P;
  R* pbo = params_bar_traj_out ? static_cast<R*>(params_bar_traj_out) : fact + (size_t)n_traj * lat::kStages * n.FS;
  lat::k_traj_bwd<R><<<n_traj, lat::kThreads, sm, st>>>(
      n, static_cast<const R*>(zs), static_cast<const R*>(rs), static_cast<const R*>(ts), T, static_cast<const R*>(params),
      static_cast<const R*>(g_z), static_cast<const R*>(g_r), (R)rtol, (R)atol, mxstep, pb, fact, static_cast<R*>(z0_bar_out),
      static_cast<R*>(r0_bar_out), static_cast<R*>(ts_bar_out), pbo, iters_out, status_out);
  LAT_CUDA(ctx, cudaGetLastError());
  if (params_bar_out) {
    lat::k_traj_sum<R><<<(n.P + 255) / 256, 256, 0, st>>>(pbo, n_traj, n.P, static_cast<R*>(params_bar_out));
    LAT_CUDA(ctx, cudaGetLastError());
  }
  return ENO_OK;
}

template <typename R>
int rhs_t(eno_ctx* ctx, const lat::Net& n, int n_traj, const void* z, const void* params, const void* z_bar, const void* r_bar,
          void* f_out, void* rdot_out, void* zbdot_out, void* pbar_out, cudaStream_t st) {
  const size_t sm = smem_bytes<R>(n, 0);
  LAT_CUDA(ctx, cudaFuncSetAttribute(lat::k_traj_rhs<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int with_vjp = (z_bar && zbdot_out && pbar_out) ? 1 : 0;
  lat::k_traj_rhs<R><<<n_traj, lat::kThreads, sm, st>>>(n, static_cast<const R*>(z), static_cast<const R*>(params),
                                                       static_cast<const R*>(z_bar), static_cast<const R*>(r_bar), with_vjp,
                                                       static_cast<R*>(f_out), static_cast<R*>(rdot_out),
                                                       static_cast<R*>(zbdot_out), static_cast<R*>(pbar_out));
  LAT_CUDA(ctx, cudaGetLastError());
  return ENO_OK;
}

}  // namespace

extern "C" {

int64_t eno_traj_param_count(const eno_dynamics_desc* desc) {
  std::string why;
  if (!lat_desc_ok(desc, &why)) return -1;
  return make_net(desc).P;
}

size_t eno_traj_workspace_bytes(const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj) {
  std::string why;
  if (!lat_desc_ok(desc, &why) || n_traj < 1) return 0;
  const lat::Net n = make_net(desc);
  const size_t el = dtype == ENO_F64 ? 8 : 4;
  return (size_t)n_traj * (3 * (size_t)n.P + (size_t)lat::kStages * n.FS) * el + 256;
}

int32_t eno_traj_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* z0, const void* ts,
                     int32_t T, const void* params, double rtol, double atol, int32_t mxstep, void* zs_out, void* rs_out,
                     void* nfe_out, int32_t* iters_out, int32_t* status_out, void* stream) {
  if (!ctx) return ENO_E_ARG;
  std::string why;
  if (!lat_desc_ok(desc, &why)) return lat_fail(ctx, ENO_E_UNSUPPORTED, why);
  if (n_traj < 1 || T < 1 || !z0 || !ts || !params || !zs_out || !nfe_out) return lat_fail(ctx, ENO_E_ARG, "eno_traj_fwd: null or empty argument");
  if (dtype != ENO_F32 && dtype != ENO_F64) return lat_fail(ctx, ENO_E_ARG, "dtype must be ENO_F32 or ENO_F64");
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const lat::Net n = make_net(desc);
  if (n.aug && !rs_out) return lat_fail(ctx, ENO_E_ARG, "eno_traj_fwd: rs_out is required with ENO_AUG_REG");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == ENO_F64 ? fwd_t<double>(ctx, n, n_traj, z0, ts, T, params, rtol, atol, mxstep, zs_out, rs_out, nfe_out, iters_out, status_out, st)
                          : fwd_t<float>(ctx, n, n_traj, z0, ts, T, params, rtol, atol, mxstep, zs_out, rs_out, nfe_out, iters_out, status_out, st);
}

int32_t eno_traj_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* zs, const void* rs,
                     const void* ts, int32_t T, const void* params, const void* g_z, const void* g_r, double rtol, double atol,
                     int32_t mxstep, void* workspace, size_t workspace_bytes, void* z0_bar_out, void* r0_bar_out, void* ts_bar_out,
                     void* params_bar_out, void* params_bar_traj_out, int32_t* iters_out, int32_t* status_out, void* stream) {
  if (!ctx) return ENO_E_ARG;
  std::string why;
  if (!lat_desc_ok(desc, &why)) return lat_fail(ctx, ENO_E_UNSUPPORTED, why);
  if (n_traj < 1 || T < 1 || !zs || !ts || !params || !g_z || !z0_bar_out || !ts_bar_out)
    return lat_fail(ctx, ENO_E_ARG, "eno_traj_bwd: null or empty argument");
  if (dtype != ENO_F32 && dtype != ENO_F64) return lat_fail(ctx, ENO_E_ARG, "dtype must be ENO_F32 or ENO_F64");
  const lat::Net n = make_net(desc);
  if (n.aug && (!rs || !g_r)) return lat_fail(ctx, ENO_E_ARG, "eno_traj_bwd: rs and g_r are required with ENO_AUG_REG");
  if (!workspace || workspace_bytes < eno_traj_workspace_bytes(desc, dtype, n_traj)) return lat_fail(ctx, ENO_E_WORKSPACE, "eno_traj_bwd: workspace too small");
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == ENO_F64 ? bwd_t<double>(ctx, n, n_traj, zs, rs, ts, T, params, g_z, g_r, rtol, atol, mxstep, workspace, z0_bar_out,
                                          r0_bar_out, ts_bar_out, params_bar_out, params_bar_traj_out, iters_out, status_out, st)
                          : bwd_t<float>(ctx, n, n_traj, zs, rs, ts, T, params, g_z, g_r, rtol, atol, mxstep, workspace, z0_bar_out,
                                         r0_bar_out, ts_bar_out, params_bar_out, params_bar_traj_out, iters_out, status_out, st);
}

int32_t eno_traj_rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* z, const void* params,
                     const void* z_bar, const void* r_bar, void* f_out, void* rdot_out, void* zbdot_out, void* params_bar_traj_out,
                     void* stream) {
  if (!ctx) return ENO_E_ARG;
  std::string why;
  if (!lat_desc_ok(desc, &why)) return lat_fail(ctx, ENO_E_UNSUPPORTED, why);
  if (n_traj < 1 || !z || !params || !f_out) return lat_fail(ctx, ENO_E_ARG, "eno_traj_rhs: null or empty argument");
  if (dtype != ENO_F32 && dtype != ENO_F64) return lat_fail(ctx, ENO_E_ARG, "dtype must be ENO_F32 or ENO_F64");
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const lat::Net n = make_net(desc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == ENO_F64 ? rhs_t<double>(ctx, n, n_traj, z, params, z_bar, r_bar, f_out, rdot_out, zbdot_out, params_bar_traj_out, st)
                          : rhs_t<float>(ctx, n, n_traj, z, params, z_bar, r_bar, f_out, rdot_out, zbdot_out, params_bar_traj_out, st);
}

}  // extern "C"
