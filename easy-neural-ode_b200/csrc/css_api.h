// Internal interface between the two translation units of libeno.so: the C entry points of include/eno.h
// (eno_abi.cu) hand ConcatSquash-stack dynamics (ENO_ARCH_CONCATSQUASH) to the GEMM-pipeline engine (css_engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/eno.h"

struct eno_ctx;

namespace eno {
namespace css {

bool desc_ok(const eno_dynamics_desc* d);
int64_t param_count(const eno_dynamics_desc* d);
int n_aux(const eno_dynamics_desc* d);
int64_t adjoint_state_size(const eno_dynamics_desc* d, int B);
size_t workspace_bytes(const eno_dynamics_desc* d, int B, int T, int mode);

int odeint_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int B, const float* y0, const float* ts, int T, const float* params,
               const float* eps, float rtol, float atol, int mxstep, void* ws, size_t ws_bytes, float* ys_out, float* aux_out,
               eno_stats* stats, eno_trace_row* trace, int trace_cap, cudaStream_t st, float grid_step = 0.f);

int odeint_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int B, const float* ys, const float* ts, int T, const float* params,
               const float* eps, const float* g_y, const float* g_aux, float rtol, float atol, int mxstep, void* ws,
               size_t ws_bytes, float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out, float* ts_bar_out,
               float* params_bar_out, eno_stats* stats, eno_trace_row* trace, int trace_cap, cudaStream_t st,
               float grid_step = 0.f);   // > 0: fixed-grid RK4 (odeint_grid, lib/ode.py:561-577) instead of adaptive dopri5

int rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int mode, int B, const float* y, float t, const float* params,
        const float* eps, const float* y_bar, const float* aux_bar, void* ws, size_t ws_bytes, float* dy_out, float* aux_out,
        float* ybar_dot_out, float* tbar_out, float* params_bar_out, float* eps_bar_dot_out, cudaStream_t st);

void destroy(eno_ctx* ctx);
void drop_collective_graphs(eno_ctx* ctx);

}  // namespace css
}  // namespace eno
