/*
 * eno.h - C ABI of the B200-native adaptive ODE-solve path (libeno.so).
 *
 * Drop-in boundary for easy-neural-ode's lib/ode.py hot path.  Every entry point
 * names the reference interface it replaces (paths relative to the reference
 * checkout).  Conventions:
 *   - plain C, no torch / C++ types in any signature;
 *   - every buffer pointer is a DEVICE pointer owned by the caller unless it is
 *     called `*_host`; the library never allocates result buffers;
 *   - arrays are row-major contiguous fp32: state [B, D], trajectories [T, B, D];
 *   - the flat parameter vector of the MLP dynamics follows the reference's
 *     ravel_pytree order of the Haiku dict (sorted keys: 'b' before 'w'):
 *         [ b1 (H) | W1 ((D+1) x H) | b2 (D) | W2 ((H+1) x D) ]
 *     with time being input row 0 of W1 and W2 (mnist.py:213-220);
 *   - calls are asynchronous on `stream` until they have to read the device-side
 *     controller state; they return after the solve has finished, with `stats`
 *     filled in;
 *   - return value: 0 ok, >0 solver condition (results still written, like the
 *     reference which has no error channel, lib/ode.py:829-831), <0 usage error;
 *   - thread-compatible: one in-flight call per eno_ctx.
 */
#ifndef ENO_H_
#define ENO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ENO_ABI_VERSION 9

/* status codes */
#define ENO_OK 0
#define ENO_MXSTEP 1        /* loop left because i >= mxstep (lib/ode.py:831) */
#define ENO_DT_UNDERFLOW 2  /* loop left because !(dt > 0) (also NaN dt)      */
#define ENO_NONFINITE 3     /* non-finite error ratio seen                     */
#define ENO_E_ARG (-1)
#define ENO_E_CUDA (-2)
#define ENO_E_WORKSPACE (-3)
#define ENO_E_UNSUPPORTED (-4)

/* Butcher tableaux: odeint() is always dopri5 (lib/ode.py:746, 823-862); the others
 * are the private forward-only loops lib/ode.py:1048-1390. */
#define ENO_TABLEAU_DOPRI5 0
#define ENO_TABLEAU_HEUN 1
#define ENO_TABLEAU_FEHLBERG 2
#define ENO_TABLEAU_BOSH 3
#define ENO_TABLEAU_RK_FEHLBERG 4
#define ENO_TABLEAU_CASH_KARP 5
#define ENO_TABLEAU_OWRENZEN 6
#define ENO_TABLEAU_OWRENZEN5 7
#define ENO_TABLEAU_TANYAM 8
#define ENO_NUM_TABLEAUX 9

/* dynamics families (the reference's Python `func` cannot be called from a kernel;
 * the spec names the closure instead) */
#define ENO_ARCH_MLP_SIGMOID 0 /* MLPDynamics, mnist.py:195-222 */
#define ENO_ARCH_CONCATSQUASH 1 /* NN_Dynamics of ConcatSquashLinear layers, softplus, ffjord_tabular.py:140-193:
                                   D -> H (x n_hidden) -> D; flat parameters per layer in Haiku ravel order
                                   [ b (out) | W (in x out) | w_hyper_bias (out) | b_gate (out) | w_gate (out) ] */

#define ENO_ARCH_TANH_STACK 2   /* GenDynamics, latent_ode.py:191-208: x <- tanh(x) W_l + b_l for n_hidden layers of width H,
                                   then tanh and Linear(D); no time input; flat parameters per layer in Haiku ravel order
                                   [ b (out) | W (in x out) ]; served by the eno_traj_* entry points only */

#define ENO_ARCH_CONCATCONV 3   /* NN_Dynamics of ConcatConv2D layers, softplus, ffjord_mnist.py:123-135, 186-232: four 3x3
                                   convolutions (stride 1, zero padding 1) 1 -> H -> H -> H -> 1 channels on an img_h x img_w
                                   image, time concatenated as input channel 0 of every layer; H = 32 or 64; flat parameters per
                                   layer in Haiku ravel order [ b (c_out) | w (3 x 3 x (c_in + 1) x c_out) ]; state NHWC with one
                                   channel, i.e. [B, img_h * img_w].  FORWARD closures only (eno_odeint_fwd with ENO_AUG_NONE /
                                   FFJORD / REG / FIN / ALL values, eno_rhs modes 0 and 1): the adjoint is not built */

/* element type of the eno_traj_* entry points */
#define ENO_F32 0
#define ENO_F64 1

/* what the integrated function returns besides dy/dt (the reference's closures) */
#define ENO_AUG_NONE 0 /* dynamics_wrap            mnist.py:285            */
#define ENO_AUG_REG 1  /* aug_dynamics 'our'       mnist.py:302-308: (dy, r_K) */
#define ENO_AUG_ALL 2  /* all_aug_dynamics         mnist.py:315-324: (dy, r_K, fro, kin) */
#define ENO_AUG_FIN 3  /* aug_dynamics 'fin'       mnist.py:294-313: (dy, dfro, dkin), reads eps (backward only) */
#define ENO_AUG_FFJORD 4 /* ffjord_dynamics       ffjord_tabular.py:250-258: (dy, -div), div = sum eps * (J eps)          */
/* For ENO_ARCH_CONCATSQUASH the closures are those of ffjord_tabular.py:250-297, all reading eps:
 *   ENO_AUG_FFJORD (dy, -div)                      ffjord_dynamics, the NFE count of nfe_fn (forward only)
 *   ENO_AUG_REG    (dy, -div, r_2)                 aug_dynamics 'our' (r = 0 when reg_order == 0); the adjoint of odeint_sepaux
 *   ENO_AUG_FIN    (dy, -div, dfro, dkin)          aug_dynamics 'fin'; the adjoint of odeint_fin_sepaux
 *   ENO_AUG_ALL    (dy, -div, r_2, dfro, dkin)     all_aug_dynamics (evaluation, forward only)
 * reg_order is 0 or 2: sol_recursive of the FFJORD scripts stops at K = 2 whatever --reg says (ffjord_tabular.py:133-136). */

typedef struct eno_ctx eno_ctx;

typedef struct eno_dynamics_desc {
  int32_t arch;      /* ENO_ARCH_*                                              */
  int32_t D;         /* state width per sample (784 for MNIST); D % 4 == 0     */
  int32_t H;         /* hidden width (100);                     H % 4 == 0     */
  int32_t reg_order; /* K of the Taylor regulariser: 0 none, 2 (--reg r2), 3 (--reg r3),
                        sol_recursive mnist.py:105-131 */
  int32_t aug;       /* ENO_AUG_*                                               */
  int32_t eps_in_args; /* 1 when the reference call passes eps among *args (mnist.py:330-332):
                        its cotangent eps_bar (identically 0 for reg_type 'our') is then
                        part of the adjoint state and dilutes the error mean            */
  int32_t n_hidden;  /* ENO_ARCH_CONCATSQUASH: hidden layers (--num_layers, ffjord_tabular.py:56); ENO_ARCH_CONCATCONV: 3 */
  int32_t img_h;     /* ENO_ARCH_CONCATCONV: image height and width (28 x 28, one channel: D = img_h * img_w); else ignored */
  int32_t img_w;
} eno_dynamics_desc;

typedef struct eno_stats {
  float nfe;          /* the reference's counter: 2 + nfe_per_iter * iterations
                         (lib/ode.py:845-847, 856)                              */
  int32_t n_iters;    /* loop iterations, accepted + rejected                   */
  int32_t n_accepted; /* accepted steps                                         */
  int32_t status;     /* ENO_OK / ENO_MXSTEP / ENO_DT_UNDERFLOW / ENO_NONFINITE */
  float last_dt;      /* controller's dt after the last iteration               */
  float last_t;       /* solver time reached                                    */
  int32_t n_launches; /* kernels launched by this call                          */
  int32_t n_rhs;      /* dynamics evaluations actually performed per sample     */
} eno_stats;

/* One row per loop iteration when a trace buffer is supplied. */
typedef struct eno_trace_row {
  float t, dt, ratio, accepted;
} eno_trace_row;

int32_t eno_abi_version(void);
const char* eno_last_error_string(const eno_ctx* ctx);

/* Context = device id + pinned mailbox + cached kernel attributes. */
int32_t eno_create(eno_ctx** out, int32_t device);
void eno_destroy(eno_ctx* ctx);

/* Number of parameters of a dynamics family: (D+1)*H + H + (H+1)*D + D. */
int64_t eno_param_count(const eno_dynamics_desc* desc);
/* Length of the flat adjoint ODE state, lib/ode.py:1516-1519:
 * 2*(B*D + n_aux*B) + 1 + (eps_in_args ? B*D : 0) + P,  n_aux = 1 for ENO_AUG_REG, 2 for ENO_AUG_FIN, 0 for NONE. */
int64_t eno_adjoint_state_size(const eno_dynamics_desc* desc, int32_t B);

/* Scratch the solve needs (device bytes). mode: 0 forward, 1 backward. */
size_t eno_workspace_bytes(const eno_dynamics_desc* desc, int32_t B, int32_t T, int32_t tableau, int32_t mode);

/*
 * Forward solve.  Replaces
 *   odeint(func, y0, t, *args, rtol, atol, mxstep)            lib/ode.py:540-559, 742-747
 *   -> _dopri5_odeint                                        lib/ode.py:823-862
 *   and _X_odeint(func, rtol, atol, mxstep, y0, ts, *args)    lib/ode.py:1048-1390
 * for func = desc (aug == ENO_AUG_NONE: y [B,D];  ENO_AUG_REG: (y, r[B]);
 * ENO_AUG_ALL: (y, r, fro, kin)).  The forward half of odeint_aux_one /
 * odeint_sepaux (lib/ode.py:993-1018) is this call with ENO_AUG_NONE: the aux
 * outputs are zeros and are produced by the host wrapper.
 *
 *   y0      [B, D]            ts  [T] strictly increasing (device)
 *   params  [P]               eps [B, D] or NULL (only ENO_AUG_ALL reads it)
 *   ys_out  [T, B, D]         aux_out [T, n_aux, B] or NULL (n_aux = 1 REG, 3 ALL)
 *   ENO_ARCH_CONCATSQUASH / ENO_ARCH_CONCATCONV: aux_out[0] is READ as the initial auxiliary state (zeros after
 *   aug_init; delta_logp of the logit pre-flow in ffjord_mnist.py:409-417) - its magnitude enters the error norm.
 *   mxstep <= 0 means unbounded (np.inf).
 *   trace_out: NULL or device array of trace_cap rows.
 */
int32_t eno_odeint_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t tableau, int32_t B, const float* y0,
                       const float* ts, int32_t T, const float* params, const float* eps, float rtol, float atol,
                       int32_t mxstep, void* workspace, size_t workspace_bytes, float* ys_out, float* aux_out,
                       eno_stats* stats_host, eno_trace_row* trace_out, int32_t trace_cap, void* stream);

/*
 * Per-example forward solves: every sample is its own dopri5 problem with its own step sequence.  Replaces
 *   jax.vmap(lambda y0, t, params: odeint(dynamics_wrap, y0, t, params)[1], (0, None, None))   mnist.py:350-353
 * (the `--vmap` NFE count of the evaluation path; each lane runs lib/ode.py:823-862 on a state of D elements).
 * One launch runs all B solves to completion; there is no workspace and no host round-trip per step.
 *
 *   y0 [B, D]   ts [2] (device)   ys_out [2, B, D]   nfe_out [B] (2 + 6 * iterations of that sample)
 *   iters_out [B, 2] or NULL: iterations and accepted steps;  status_out [B] or NULL: ENO_OK / ENO_MXSTEP / ...
 * Runs on the cluster-resident path only (ENO_E_UNSUPPORTED otherwise).
 */
int32_t eno_odeint_fwd_vmap(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* y0, const float* ts,
                            int32_t T, const float* params, float rtol, float atol, int32_t mxstep, float* ys_out,
                            float* nfe_out, int32_t* iters_out, int32_t* status_out, void* stream);

/*
 * Backward (adjoint) solve.  Replaces
 *   _odeint_rev(func, rtol, atol, mxstep, res, g)             lib/ode.py:1494-1529
 *   as reached through _odeint_aux_one_rev                    lib/ode.py:1677-1684
 * for rev_func = aug_dynamics(reg_type 'our') (mnist.py:302-308), i.e. the
 * augmented state per sample is (y [D], r) and the adjoint ODE state is
 * (y, r, y_bar, r_bar, t0_bar, eps_bar = 0, params_bar).
 *
 *   ys   [T, B, D]  forward trajectory (residuals, lib/ode.py:1444)
 *   g_y  [T, B, D]  cotangent of ys;   g_r [T, B] cotangent of the aux output r
 *   y0_bar_out [B, D], r0_bar_out [B], ts_bar_out [T], params_bar_out [P]
 *   stats_host: array of T-1 entries, one per backward segment (i = T-1 .. 1).
 */
int32_t eno_odeint_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts,
                       int32_t T, const float* params, const float* g_y, const float* g_r, float rtol, float atol,
                       int32_t mxstep, void* workspace, size_t workspace_bytes, float* y0_bar_out,
                       float* r0_bar_out, float* ts_bar_out, float* params_bar_out, eno_stats* stats_host,
                       eno_trace_row* trace_out, int32_t trace_cap, void* stream);

/*
 * Backward solve for rev_funcs that read eps.  Replaces
 *   _odeint_rev as reached through _odeint_sepaux_rev        lib/ode.py:1686-1693
 * (odeint_sepaux, lib/ode.py:678-697, 1006-1018) for rev_func = aug_dynamics(reg_type 'fin')
 * (mnist.py:294-313): per sample the augmented state is (y [D], fro, kin), the RHS is
 * (f, mean_d (J eps)^2, mean_d f^2) with J eps = jvp(f wrt y, eps), and eps_bar is a live block of the
 * adjoint state.  With ENO_AUG_NONE / ENO_AUG_REG it is eno_odeint_bwd (eps may be NULL).
 *
 *   eps   [B, D]               g_aux [T, n_aux, B]  cotangents of the aux outputs (n_aux = 2: fro, kin)
 *   aux0_bar_out [n_aux, B]    eps_bar_out [B, D] or NULL
 * Runs on the cluster-resident path only (ENO_E_UNSUPPORTED otherwise).
 */
int32_t eno_odeint_bwd_sepaux(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts,
                              int32_t T, const float* params, const float* eps, const float* g_y, const float* g_aux,
                              float rtol, float atol, int32_t mxstep, void* workspace, size_t workspace_bytes,
                              float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out, float* ts_bar_out,
                              float* params_bar_out, eno_stats* stats_host, eno_trace_row* trace_out, int32_t trace_cap,
                              void* stream);

/*
 * Fixed-grid RK4 solves.  Replace
 *   odeint_grid(func, y0, t, *args, step_size)                                  lib/ode.py:561-577 -> _rk4_odeint 864-903
 *   odeint_grid_sepaux_one / odeint_grid_sepaux(fwd_func, rev_func, ...)        lib/ode.py:579-615, 906-946
 *   _rk4_odeint_rev (the adjoint: the same RK4 grid on the augmented system)    lib/ode.py:1531-1566
 * as ffjord_tabular.py:75-78 / ffjord_mnist.py:68-71 select them: steps of step_size from ts[0], the last one clipped so
 * that it lands on ts[-1]; the result has the two rows (y0, y(ts[-1])) - T must be 2, like the reference (lib/ode.py:884) -
 * and nfe = 4 per step.  ENO_ARCH_CONCATSQUASH / ENO_ARCH_CONCATCONV only (ENO_E_UNSUPPORTED otherwise).  Arguments as in
 * eno_odeint_fwd / eno_odeint_bwd_sepaux, step_size in place of (rtol, atol, mxstep); collective after eno_comm_init.
 */
int32_t eno_odeint_grid_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* y0, const float* ts, int32_t T,
                            const float* params, const float* eps, float step_size, void* workspace, size_t workspace_bytes,
                            float* ys_out, float* aux_out, eno_stats* stats_host, void* stream);
int32_t eno_odeint_grid_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t B, const float* ys, const float* ts, int32_t T,
                            const float* params, const float* eps, const float* g_y, const float* g_aux, float step_size,
                            void* workspace, size_t workspace_bytes, float* y0_bar_out, float* aux0_bar_out, float* eps_bar_out,
                            float* ts_bar_out, float* params_bar_out, eno_stats* stats_host, void* stream);

/*
 * One evaluation of the integrated function (unit-test hook), the closures of
 * mnist.py:285-324.
 *   mode 0: dy = f(y, t)                                       -> dy_out [B,D]
 *   mode 1: (dy, r_K [, fro, kin]) per desc->aug               -> dy_out, aux_out [n_aux, B]
 *   mode 2: vjp of aug_dynamics('our') at (y, t) with cotangent (y_bar [B,D], r_bar [B])
 *           -> dy_out, aux_out [1,B], ybar_dot_out [B,D], tbar_out [1], params_bar_out [P]
 *              (lib/ode.py:1498-1504)
 *           With ENO_AUG_FIN (eps required): aux_out [2,B] = (dfro, dkin), r_bar [2,B], and
 *           eps_bar_dot_out [B,D] (or NULL) receives the cotangent of eps.
 */
int32_t eno_rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t mode, int32_t B, const float* y, float t,
                const float* params, const float* eps, const float* y_bar, const float* r_bar, void* workspace,
                size_t workspace_bytes, float* dy_out, float* aux_out, float* ybar_dot_out, float* tbar_out,
                float* params_bar_out, float* eps_bar_dot_out, void* stream);

/*
 * The rest of mnist.py's update() around the ODE block, so that a train step is library launches only.
 * eno_mnist_head replaces PostODE + the loss and its gradients (mnist.py:225-238, 90-94, 449-452, 528):
 *   logits = sigmoid(y1) @ W + b;  loss = mean softmax cross-entropy;  inv_count = 1 / (global batch)
 *   y1 [B, D]  labels [B] int64  W [D, C]  b [C]  scratch [B*D + B*C + B] floats
 *   -> loss_out [1] (mean over this call's B rows), gy_out [B, D] = d(sum_r loss_r * inv_count)/dy1,
 *      gW_out [D, C], gb_out [C] (sums over this call's rows; all-reduce them across ranks)
 * eno_momentum_update replaces optimizers.momentum (mnist.py:530-540) and the lam_w * 0.5 |theta|^2 term of loss_fn
 * (mnist.py:456-470) for up to three parameter blocks in one launch:
 *   g <- g + lam_w * x;  v <- mass * v + g;  x <- x - lr * v
 * applied only if both guard words (device int32, may be NULL) are non-zero - in deferred mode they are the solves'
 * `done` flags (first word pairs of the workspaces).  The step size is the value of the script's lr_schedule
 * (mnist.py:530-538) at this iteration, evaluated by the host: `lr`, or - when `lr_dev` is not NULL - the device float
 * it points to (a captured CUDA graph keeps working when the schedule moves).
 * Both calls only enqueue work (CUDA-graph capturable).
 */
int32_t eno_mnist_head(eno_ctx* ctx, int32_t B, int32_t D, int32_t C, const float* y1, const int64_t* labels, const float* W,
                       const float* b, float inv_count, float* scratch, float* loss_out, float* gy_out, float* gW_out,
                       float* gb_out, void* stream);
int32_t eno_momentum_update(eno_ctx* ctx, int32_t n_blocks, float* const* params, float* const* velocity,
                            const float* const* grads, const int64_t* sizes, float lr, float mass, float lam_w,
                            const float* lr_dev, const int32_t* ok_a, const int32_t* ok_b, void* stream);

/*
 * Deferred launch mode (no reference counterpart).  By default a solve polls its device-side controller
 * once per chunk of iterations.  With eno_set_deferred(ctx, 1, fwd_iters, bwd_iters) the calls enqueue the
 * two initial-step launches plus EXACTLY that many iteration launches (iterations after termination exit
 * at their first instruction) and return without synchronising: they can be captured into a CUDA graph,
 * which removes every launch gap of a training step.  stats_host then only reports status -100
 * ("deferred"); call eno_read_stats on the same workspace afterwards: status -101 means the budget was
 * too small (the outputs are then meaningless and the step must be repeated with a larger budget).
 * The backward call supports T == 2 in this mode.
 */
int32_t eno_set_deferred(eno_ctx* ctx, int32_t on, int32_t fwd_iters, int32_t bwd_iters);
int32_t eno_read_stats(eno_ctx* ctx, const void* workspace, eno_stats* out, void* stream);

/*
 * Multi-GPU, one process per GPU (no reference counterpart: the reference is single-device).  The batch
 * is sharded in contiguous row blocks of equal size B per rank, parameters are replicated.  After
 * eno_comm_init every eno_odeint_fwd / eno_odeint_bwd call on this ctx is COLLECTIVE: all ranks must
 * call it with their shard; the error norm (lib/ode.py:518-527), the initial-step norms (86-104) and,
 * in the adjoint, the batch-summed blocks params_bar / t0_bar are exchanged over NCCL every iteration so
 * that all ranks follow the step sequence of the single-device solve on the concatenated batch.
 * params_bar_out and ts_bar_out are then already global sums, identical on every rank.  This holds for every
 * architecture: the ConcatSquash / ConcatConv2D engine all-reduces the stage combinations of its replicated
 * [t0_bar | params_bar] block once per adjoint iteration and gathers the per-rank error-norm sums (NCCL, part of the
 * captured iteration graph; the peer-memory exchange below serves the MLP dynamics).  eno_rhs stays local.
 * Rank 0 obtains `id` (128 bytes) from eno_comm_unique_id and ships it to the other ranks (e.g. with
 * torch.distributed); `nccl_path` may name libnccl.so.2 explicitly (NULL: the copy already loaded).
 */
int32_t eno_comm_unique_id(eno_ctx* ctx, uint8_t* id_out /*[128]*/, const char* nccl_path);
int32_t eno_comm_init(eno_ctx* ctx, const uint8_t* id /*[128]*/, int32_t rank, int32_t world, const char* nccl_path);
int32_t eno_comm_destroy(eno_ctx* ctx);
/* Uneven shards (global batch 100 on 8 GPUs = 13,13,13,13,12,12,12,12 rows).  By default every rank
 * must pass the same B.  After this call a rank may pass any B <= slot_rows; every rank's exchange records are laid out for
 * slot_rows rows (the unused tail reads as zero, which leaves the rank-ordered sums unchanged) and the mean of the error
 * norm runs over rows_global rows (the concatenated batch).  (0, 0) restores equal shards.  Same values on every rank. */
int32_t eno_comm_set_shards(eno_ctx* ctx, int32_t slot_rows, int64_t rows_global);
/*
 * Optional, after eno_comm_init: move the per-iteration exchange of the collective solves from NCCL calls into the
 * solver's own finish kernels, over NVLink peer memory (one node, at most 8 ranks).  Every rank calls eno_peer_export
 * (allocates its exchange buffer of `bytes` >= 2 MiB + what the solves need - 64 MiB is plenty for the MNIST size - and
 * returns its 64-byte CUDA IPC handle), the host side gathers the handles of all ranks in rank order, and every rank
 * calls eno_peer_attach with them.  What travels: the per-sample norm partials are pushed to every rank (forward and
 * adjoint); of the four params_bar stage combinations every rank pulls only the 1/world slice it owns, forms candidate
 * / error / norm partial for that slice, and one more scalar per rank is exchanged; params_bar is gathered once per
 * segment.  Without these calls the solves use the NCCL exchange.
 */
int32_t eno_peer_export(eno_ctx* ctx, uint64_t bytes, uint8_t* handle_out /*[64]*/);
int32_t eno_peer_attach(eno_ctx* ctx, const uint8_t* handles /*[world][64], rank order*/);

/*
 * Measurement hooks used by bench.py (no reference counterpart; the reference only wall-clocks the
 * whole jitted update, mnist.py:632-640).  With profiling on, the step kernels are bracketed by CUDA
 * events on the launch stream; eno_profile_read returns and clears the accumulated milliseconds and
 * launch counts per kernel kind: [0] forward step, [1] adjoint per-sample step, [2] adjoint
 * parameter-gradient step (cluster path: the batched tcgen05 contraction + k_adj_pcombine), [3] multi-GPU exchange
 * (finish kernel incl. the peer-memory or NCCL exchange), [4] the batched tcgen05 contraction of [2] alone.
 * eno_ffma_peak measures the non-tensor fp32 FMA peak (TFLOP/s).
 */
int32_t eno_profile_enable(eno_ctx* ctx, int32_t on);
int32_t eno_profile_read(eno_ctx* ctx, double* ms_out /*[5]*/, int64_t* count_out /*[5]*/);
int32_t eno_ffma_peak(eno_ctx* ctx, double* tflops_out, void* stream);
/* With profiling on, solves of ENO_ARCH_CONCATSQUASH dynamics bracket every tensor-core GEMM launch of their loop
 * iterations with CUDA events (direct launches instead of the iteration graph).  Returns and clears the accumulated
 * milliseconds, ALGORITHMIC flops (2 M N K per contraction, not the three tf32 passes, not the tile padding) and launches. */
int32_t eno_profile_read_gemm(eno_ctx* ctx, double* ms_out, double* flop_out, int64_t* count_out);

/*
 * Optimiser step of the FFJORD scripts, one launch (only enqueues work).  Replaces
 *   optimizers.adam(step_size=lr_schedule) + the lam_w * 0.5 |theta|^2 term of loss_fn     ffjord_tabular.py:417-432, 527-529
 *   g <- g + lam_w * x;  m <- (1-b1) g + b1 m;  v <- (1-b2) g^2 + b2 v;
 *   x <- x - lr * (m / (1 - b1^(step+1))) / (sqrt(v / (1 - b2^(step+1))) + eps)            (jax.experimental.optimizers.adam)
 * `lr` is the schedule's value at `step` (piecewise_constant, ffjord_tabular.py:525-526; evaluated by the host).
 */
int32_t eno_adam_update(eno_ctx* ctx, float* params, float* m, float* v, const float* grads, int64_t n, float lr, float b1,
                        float b2, float eps, int32_t step, float lam_w, void* stream);

/*
 * Unit-test hook of the tensor-core contraction used by the wide-network path (no reference counterpart; the
 * reference's contractions are XLA dots inside ffjord_tabular.py:151-152): C[M, N] = A[M, K] * B[N, K]^T in fp32
 * accuracy, computed as the 3-pass tf32 split on tcgen05 with TMEM accumulators and TMA operand loads.
 * lda / ldb / ldc in elements, multiples of 4; bn: accumulator tile width, 64 or 128; splits >= 1: K partitions
 * summed in a fixed order.  Allocates its own scratch; synchronises the stream.
 */
int32_t eno_gemm_tf32x3(eno_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B,
                        int32_t ldb, float* C, int32_t ldc, int32_t bn, int32_t splits, void* stream);
/* The same contraction with MN-major operands (tcgen05 "transposed" operand descriptors): C[m, n] = sum_k A[k, m] B[k, n],
 * A [K][lda], B [K][ldb] row-major, lda % 4 == ldb % 4 == 0.  Unit-test hook of the mode the convolution weight gradients use. */
int32_t eno_gemm_tf32x3_mn(eno_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B,
                           int32_t ldb, float* C, int32_t ldc, int32_t bn, int32_t splits, void* stream);

/*
 * Per-trajectory differentiable solves (BASELINE config C3, latent_ode.py).  Replace
 *   jax.vmap(lambda z_, t_: odeint(dynamics, init_fn(z_), t_, params["gen_dynamics"], **gen_ode_kwargs),
 *            in_axes=(0, None))(z0_, timesteps)                                       latent_ode.py:337-350
 * (itself under a second vmap over the 3 latent samples) and its reverse pass: under vmap every trajectory is its own
 * odeint problem - own dopri5 step sequence over all T output times (lib/ode.py:823-862), own adjoint over T-1
 * segments whose state (y, y_bar, t0_bar, params_bar) carries a full params_bar block per trajectory
 * (lib/ode.py:1494-1529); vmap's transpose then sums the per-trajectory params_bar.
 * desc: arch ENO_ARCH_TANH_STACK, D latent width, H hidden width, n_hidden hidden layers (the script's
 * gen_layers + 1), reg_order 0 / 2 / 3 (sol_recursive latent_ode.py:82-108), aug ENO_AUG_REG for
 * dynamics = augment_dynamics(gen_dynamics_wrap) with state (z, r) (latent_ode.py:238-256, 352-355: dr/dt =
 * mean(r_K^2)), ENO_AUG_NONE for the plain count_nfe dynamics.  dtype ENO_F64 is the script's precision
 * (jax_enable_x64, latent_ode.py:23); every array below has that element type.
 * One launch runs ALL trajectories to completion (one CTA per trajectory, weights resident in shared memory);
 * the calls only enqueue work on `stream` (CUDA-graph capturable) and never synchronise.
 *
 * eno_traj_fwd:   z0 [n_traj, D]   ts [T] strictly increasing   params [P]
 *   -> zs_out [n_traj, T, D]   rs_out [n_traj, T] (ENO_AUG_REG; r starts at 0)   nfe_out [n_traj] (2 + 6 iterations)
 *      iters_out [n_traj, 2] or NULL (iterations, accepted)   status_out [n_traj] or NULL (ENO_OK / ENO_MXSTEP / ...)
 * eno_traj_bwd:   zs, rs as produced by the forward call;  g_z [n_traj, T, D], g_r [n_traj, T] cotangents of the outputs
 *   -> z0_bar_out [n_traj, D]   r0_bar_out [n_traj] or NULL   ts_bar_out [n_traj, T]
 *      params_bar_out [P] or NULL: sum over the trajectories in index order
 *      params_bar_traj_out [n_traj, P] or NULL: the per-trajectory blocks
 *      iters_out [n_traj, T-1, 2] or NULL: iterations / accepted steps of the segments in the order they run (i = T-1 .. 1)
 *   workspace: eno_traj_workspace_bytes(desc, dtype, n_traj) device bytes.
 * eno_traj_rhs (unit-test hook): one evaluation of the dynamics per trajectory, f_out [n_traj, D], rdot_out [n_traj] or
 *   NULL (mean(r_K^2)), and - when z_bar, zbdot_out and params_bar_traj_out are given - the vjp of (f, rdot) with
 *   cotangents (z_bar [n_traj, D], r_bar [n_traj] or NULL = 0) wrt z and the parameters (lib/ode.py:1498-1504).
 */
int64_t eno_traj_param_count(const eno_dynamics_desc* desc);
size_t eno_traj_workspace_bytes(const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj);
int32_t eno_traj_fwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* z0, const void* ts,
                     int32_t T, const void* params, double rtol, double atol, int32_t mxstep, void* zs_out, void* rs_out,
                     void* nfe_out, int32_t* iters_out, int32_t* status_out, void* stream);
int32_t eno_traj_bwd(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* zs, const void* rs,
                     const void* ts, int32_t T, const void* params, const void* g_z, const void* g_r, double rtol, double atol,
                     int32_t mxstep, void* workspace, size_t workspace_bytes, void* z0_bar_out, void* r0_bar_out, void* ts_bar_out,
                     void* params_bar_out, void* params_bar_traj_out, int32_t* iters_out, int32_t* status_out, void* stream);
int32_t eno_traj_rhs(eno_ctx* ctx, const eno_dynamics_desc* desc, int32_t dtype, int32_t n_traj, const void* z, const void* params,
                     const void* z_bar, const void* r_bar, void* f_out, void* rdot_out, void* zbdot_out, void* params_bar_traj_out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ENO_H_ */
