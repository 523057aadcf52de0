/*
 * otk.h - C ABI of libotk.so: the B200 (sm_100a) kernels behind the latent optimal-transport path of
 * theoad/ot-vae-lightning.
 *
 * The reference has no FFI layer: its boundary is the Python API of `ot_vae_lightning/ot/` and
 * `ot_vae_lightning/metrics/`.  Each entry point below names the reference interface it replaces
 * (file:line relative to /root/reference/ot_vae_lightning).  The host-side Python mirror of that API
 * (`ot-vae-lightning_b200/`) binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C, no exceptions, no ownership transfer.  Every function returns an otk_status (0 = ok).
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - matrices are row-major and dense: a `[L, d, d]` batch is L contiguous d*d blocks.
 *   - `dtype` arguments use otk_dtype.  Latents are always fp32; statistics / matrix results are fp32 or fp64.
 *   - work is enqueued on `stream` (a cudaStream_t / CUstream); nothing synchronises unless stated.
 *   - no hidden allocation: scratch memory is caller-provided, sized by the matching `*_workspace_bytes`.
 *   - re-entrant per stream; no global mutable state besides a per-process cache of function attributes.
 */
#ifndef OTK_H_
#define OTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTK_ABI_VERSION 1

typedef void* otk_stream_t; /* cudaStream_t */

typedef enum {
  OTK_OK = 0,
  OTK_ERR_INVALID_ARGUMENT = -1, /* bad shape / null pointer / unsupported dtype      -> ValueError          */
  OTK_ERR_WORKSPACE = -2,        /* workspace too small                                -> ValueError          */
  OTK_ERR_CUDA = -3,             /* a CUDA runtime / driver call failed                -> RuntimeError        */
  OTK_ERR_UNSUPPORTED_DEVICE = -4, /* not an sm_100 device                             -> RuntimeError        */
  OTK_ERR_NOT_CONVERGED = -5     /* Newton-Schulz diverged: the matrix is indefinite         -> NotConverged      */
} otk_status;

typedef enum { OTK_F32 = 0, OTK_F64 = 1 } otk_dtype;

/* cost kinds for the point-cloud Sinkhorn (K8/K9) */
typedef enum {
  OTK_COST_SQEUCLIDEAN = 0, /* |x-y|^2            : w2_utils.py:121-125 (mean part), SURVEY 8d cfg3 */
  OTK_COST_INV_EUCLIDEAN = 1 /* 1/(|x-y|_2 + 1e-8) : CodebookModel.energy, codebook_model.py:155-160 */
} otk_cost_kind;

int otk_abi_version(void);
const char* otk_status_string(int status);
/* last CUDA error text seen by this thread inside libotk (empty string if none) */
const char* otk_last_error(void);
/* 1 if the current device is sm_100 (B200); entry points return OTK_ERR_UNSUPPORTED_DEVICE otherwise */
int otk_device_supported(void);
/* number of libotk kernels launched by this process so far (bench.py reports the delta as gpu_launches) */
unsigned long long otk_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  streaming sufficient statistics.
 * Replaces GaussianModel._stats + update (ot/distribution_models/gaussian_model.py:99-108, 144-157),
 * utils.ema (utils/__init__.py:204-206) and FrechetInceptionDistance._extract_features/update
 * (metrics/fid.py:99-122).
 *   x        [L, rows, dim] fp32 latents; row stride `row_stride`, batch stride `batch_stride` (elements)
 *   n_obs    [L]            running count      (dtype n_dtype)
 *   sum      [L, dim]       running sum x      (dtype buf_dtype)
 *   sum_cov  [L, dim, dim]  running sum x x^T  (dtype buf_dtype)
 * running <- running + new            if decay < 0   (update_decay=None)
 * running <- running*decay + new*(1-decay) otherwise (update_decay=decay, applied to n_obs too)
 * `new` is accumulated in fp32 over short row chunks on the tensor cores (3xTF32) or FFMA and across
 * chunks in fp64.  Workspace is an fp64 staging area (zeroed internally).
 * ---------------------------------------------------------------------------------------------- */
size_t otk_stats_update_workspace_bytes(int64_t L, int64_t rows, int64_t dim);
int otk_stats_update(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                     int64_t batch_stride, double decay, void* n_obs, int n_dtype, void* sum,
                     void* sum_cov, int buf_dtype, void* workspace, size_t workspace_bytes,
                     otk_stream_t stream);

/* Same contract for fp64 latents (the reference casts `samples.type_as(buffer)`, gaussian_model.py:103; FID accumulates
 * `features.double()`, metrics/fid.py:101-104): fp64 products and accumulation on the DFMA engine, so that sums of outer
 * products stay positive semi-definite to fp64 round-off. */
int otk_stats_update_f64(const double* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                         int64_t batch_stride, double decay, void* n_obs, int n_dtype, void* sum,
                         void* sum_cov, int buf_dtype, void* workspace, size_t workspace_bytes,
                         otk_stream_t stream);

/* TransportOperator.update(source_samples, target_samples) (ot/transport/base.py: the two GaussianModel.update calls of
 * one batch, gaussian_model.py:99-108) as one call: x_a -> buffers a, x_b -> buffers b, both [rows, dim] fp32 with the same
 * row stride, one model each (no leading shape), same decay and buffer dtypes.  Batches in the latency regime (rows <= 256)
 * take ONE launch for both models; larger ones run the two updates back to back.  workspace as otk_stats_update (L = 1). */
int otk_stats_update_pair(const float* x_a, const float* x_b, int64_t rows, int64_t dim, int64_t row_stride,
                          double decay, void* n_obs_a, void* sum_a, void* sum_cov_a, void* n_obs_b, void* sum_b,
                          void* sum_cov_b, int n_dtype, int buf_dtype, void* workspace, size_t workspace_bytes,
                          otk_stream_t stream);

/* K2  mean = sum/n ; cov = sum_cov/n - mean mean^T (biased).  Replaces mean_cov, ot/matrix_utils.py:145-158
 * (full-matrix branch).  n is a device vector [L] (a python int crashes the reference: utils/__init__.py:322). */
int otk_mean_cov(const void* sum, const void* sum_cov, const void* n_obs, int n_dtype, int64_t L,
                 int64_t dim, void* mean, void* cov, int dtype, otk_stream_t stream);

/* GaussianModel.fit in one launch (gaussian_model.py:110-183: _compute_mean_cov -> _update_mean / _update_cov): for every
 * leading index with n_obs > 1e-8, mean [L,d] = sum / n and cov_raw [L,d,d] = sum_cov / n - mean mean^T are overwritten;
 * indices never observed keep their values.  cov_sym (may be NULL) additionally receives triu-mirror(cov_raw) + shift I,
 * the operand of otk_transport_operator.  sum / sum_cov in buf_dtype, mean / cov_raw / cov_sym in dtype. */
int otk_gaussian_fit(const void* sum, const void* sum_cov, int buf_dtype, const void* n_obs, int n_dtype, int64_t L,
                     int64_t dim, void* mean, void* cov_raw, void* cov_sym, double shift, int dtype,
                     otk_stream_t stream);

/* The `Symmetric` + `MakePositiveDefinite` parametrizations read on every `.cov` access
 * (gaussian_model.py:204-229): out = triu(a) + triu(a,1)^T + shift*I, shift [L] (may be NULL). */
int otk_symmetrize_shift(const void* a, const void* shift, int64_t L, int64_t dim, void* out,
                         int dtype, otk_stream_t stream);

/* sum((A - A^T)^2) per matrix -> asym [L] fp64.  Replaces is_symmetric, ot/matrix_utils.py:79-88
 * (the `< 1e-8` comparison is done by the caller). */
int otk_asymmetry(const void* a, int64_t L, int64_t dim, int dtype, double* asym, otk_stream_t stream);

/* K4  smallest eigenvalue per matrix -> lam_min [L] fp64.  Replaces min_eig, ot/matrix_utils.py:91-98
 * (used by is_pd :109 and make_psd :132).  Lanczos with full re-orthogonalisation + Sturm multisection.
 * steps <= 0 (default): runs until the smallest Ritz pair has converged (residual bound 1e-13 ||A||) or for `dim` steps -
 * a complete tridiagonalisation, i.e. the exact lambda_min.  steps > 0: exactly min(steps, dim) steps (an upper bound). */
size_t otk_min_eig_workspace_bytes(int64_t L, int64_t dim, int steps);
int otk_min_eig(const void* a, int64_t L, int64_t dim, int dtype, int steps, double* lam_min,
                void* workspace, size_t workspace_bytes, otk_stream_t stream);

/* K3  matrix square root and inverse square root of SPD matrices by coupled Newton-Schulz in fp32-accurate
 * arithmetic (3xTF32 on tcgen05 / FFMA), optionally followed by `polish` fp64 correction steps.
 * Replaces sqrtm / invsqrtm / _matrix_operator, ot/matrix_utils.py:37-76.
 *   a [L,d,d] (dtype) -> root [L,d,d], iroot [L,d,d] (either may be NULL), same dtype.
 *   `ridge` is added to the diagonal before the iteration (the reference adds 1e-8 for the inverse
 *   root, w2_utils.py:766).  iters<=0 selects the default (adaptive, <= 40). */
size_t otk_sqrtm_workspace_bytes(int64_t L, int64_t dim);
int otk_sqrtm(const void* a, int64_t L, int64_t dim, int dtype, double ridge, int iters, int polish,
              void* root, void* iroot, void* workspace, size_t workspace_bytes, otk_stream_t stream);

/* K5  W2^2 = |ms-mt|^2 + tr(Cs + Ct - 2 (Ct^1/2 Cs Ct^1/2)^1/2) -> w2 [L] fp64.
 * Replaces w2_gaussian, ot/w2_utils.py:40-80 (validation stays in the Python host). */
size_t otk_w2_gaussian_workspace_bytes(int64_t L, int64_t dim);
int otk_w2_gaussian(const void* mean_s, const void* mean_t, const void* cov_s, const void* cov_t,
                    int64_t L, int64_t dim, int dtype, int iters, int polish, double* w2,
                    void* workspace, size_t workspace_bytes, otk_stream_t stream);

/* K6  T = (1-p) Cs^-1/2 (Cs^1/2 Ct Cs^1/2)^1/2 Cs^-1/2 + p I -> T [L,d,d] (dtype).
 * Replaces _compute_transport_full_mat, ot/w2_utils.py:756-768 (deterministic full-matrix branch of
 * compute_transport_operators :391-458).  If w2 != NULL also writes W2^2 [L] fp64 reusing the same roots
 * (tr((Cs^1/2 Ct Cs^1/2)^1/2) == tr((Ct^1/2 Cs Ct^1/2)^1/2)); means may then not be NULL. */
size_t otk_transport_operator_workspace_bytes(int64_t L, int64_t dim);
int otk_transport_operator(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype,
                           double pg_star, int iters, int polish, void* T, const void* mean_s,
                           const void* mean_t, double* w2, void* workspace, size_t workspace_bytes,
                           otk_stream_t stream);

/* K6, stochastic variant (eq. 19): T = (1-p) Ct^1/2 (Ct^1/2 Cs Ct^1/2)^1/2 (Ct + 1e-8 I)^-1/2 Cs^+ + p I and the noise
 * covariance Cw = sqrt(1-p) Ct^1/2 (I - Ct^1/2 T* Cs^+ T* Ct^1/2) Ct^1/2, T* = T_{t->s} of eq. 17.
 * Replaces _compute_transport_full_mat_stochastic, ot/w2_utils.py:774-793 (torch.linalg.pinv + 5 eigh + 14 matmuls):
 * three Newton-Schulz solves and the products on the device.  Cs^+ is taken as Cs^-1 (positive definite source). */
size_t otk_transport_operator_stochastic_workspace_bytes(int64_t L, int64_t dim);
int otk_transport_operator_stochastic(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype,
                                      double pg_star, int iters, int polish, void* T, void* Cw, void* workspace,
                                      size_t workspace_bytes, otk_stream_t stream);

/* K7  y[l,b,:] = T[l] (x[l,b,:] - mean_s[l]) + mean_t[l].  Replaces apply_transport, ot/w2_utils.py:464-527
 * (deterministic, full-matrix; as called by W2Mixin.apply_transport :581-597 and
 * GaussianTransport.transport, transport/gaussian_transport.py:80-95).
 *   x, y [L, rows, dim] fp32 ; mean_s, mean_t [L, dim] and T [L, dim, dim] in `dtype`.
 * Workspace holds the fp32 hi/lo split of T and the folded bias. */
size_t otk_apply_transport_workspace_bytes(int64_t L, int64_t rows, int64_t dim);
int otk_apply_transport(const float* x, int64_t L, int64_t rows, int64_t dim, const void* mean_s,
                        const void* mean_t, const void* T, int dtype, float* y, void* workspace,
                        size_t workspace_bytes, otk_stream_t stream);
/* K7, prepared form: everything that depends only on the operator (fp32 casts, TF32 and scaled-FP16 hi/lo planes of T,
 * per-feature input scales taken from the source variances var_s = diag(cov_source), folded biases) is built ONCE into
 * a caller-owned `state` buffer (GaussianTransport.compute, transport/gaussian_transport.py:64-78), and every later
 * transport() call (transport/gaussian_transport.py:80-95) is a single kernel launch plus its device-gated fallback.
 * `state` must stay untouched between prepare and the applies, and serves one stream at a time. */
size_t otk_transport_prepared_bytes(int64_t L, int64_t dim);
int otk_transport_prepare(const void* mean_s, const void* mean_t, const void* T, const void* var_s, int dtype,
                          int64_t L, int64_t dim, void* state, size_t state_bytes, otk_stream_t stream);
int otk_apply_transport_prepared(const float* x, int64_t L, int64_t rows, int64_t dim, const void* state,
                                 size_t state_bytes, float* y, otk_stream_t stream);

/* Same, for latents that are a strided VIEW (`utils.permute_and_flatten` hands over [B, T, D] token / channel views,
 * utils/__init__.py:233-311; the reference makes them contiguous first, :260-261): x[l, b, :] starts at
 * x + l * batch_stride + b * row_stride (elements), unit feature stride; y is dense [L, rows, dim].  Strides that are
 * multiples of 4 elements are read in place by TMA, anything else by the FFMA engine. */
int otk_apply_transport_prepared_strided(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                                         int64_t batch_stride, const void* state, size_t state_bytes, float* y,
                                         otk_stream_t stream);

/* K8  log-domain Sinkhorn on a materialised cost.  Replaces sinkhorn_log, ot/w2_utils.py:276-319:
 *   u = v = 0; per iteration v = log(b+1e-8) - LSE_i(u_i - C_ij/reg), then u = log(a+1e-8) - LSE_j(v_j - C_ij/reg);
 *   stop after the iteration in which min over the batch of sum|du| + sum|dv| < threshold (device-side flag,
 *   polled from the host every `poll_every` iterations; iterations after the stop are no-ops).
 *   a [L,N], b [L,M], C [L,N,M] (dtype) -> u [L,N], v [L,M]; plan [L,N,M] = exp(u_i+v_j-C_ij/reg) if not NULL.
 *   iters_done_host (may be NULL) receives the number of iterations executed (forces a stream sync). */
size_t otk_sinkhorn_dense_workspace_bytes(int64_t L, int64_t N, int64_t M);
int otk_sinkhorn_dense(const void* a, const void* b, const void* C, int64_t L, int64_t N, int64_t M,
                       int dtype, double reg, int max_iter, double threshold, int poll_every, void* u,
                       void* v, void* plan, int* iters_done_host, void* workspace,
                       size_t workspace_bytes, otk_stream_t stream);

/* K8/K9  log-domain Sinkhorn between point clouds with the cost tile recomputed on the fly (no N x M matrix
 * in HBM).  API extension (the reference only takes a materialised C); same recurrences as above with
 * C_ij = scale * cost(x_i, y_j).  Replaces the cost producers + sinkhorn_log pair in
 * DiscreteTransport.compute (transport/discrete_transport.py:55-68) and batch_ot_gmm (w2_utils.py:255-269).
 *   x [N,d], y [M,d] fp32; a [N], b [M] fp32; u [N], v [M] fp32 in/out (not reset if `warm_start`).
 *   Row-sharded multi-GPU use: call otk_sinkhorn_points_colstep on the local rows of x, combine the
 *   (max, sumexp) partials of all ranks, then otk_sinkhorn_points_rowstep. */
size_t otk_sinkhorn_points_workspace_bytes(int64_t N, int64_t M, int64_t dim, int cost_kind);
/* scale = 1/max_ij cost (w2_utils.py:265-266) if scale_inv_max != 0, else `scale`; result in *scale_out_host */
int otk_sinkhorn_points(const float* x, const float* y, int64_t N, int64_t M, int64_t dim,
                        const float* a, const float* b, int cost_kind, double scale, int scale_inv_max,
                        double reg, int max_iter, double threshold, int poll_every,
                        int precision /* 0 auto: fused tcgen05 engine when eligible; 1: exact fp32 cost tiles */,
                        float* u, float* v, double* summary /* [4]: <C,pi>, sum pi, max|row err|, max|col err| */,
                        float* row_marginal /* [N] or NULL */, float* col_marginal /* [M] or NULL */,
                        int* iters_done_host, void* workspace, size_t workspace_bytes,
                        otk_stream_t stream);
/* half-steps for the row-sharded path: partial column LSE over local rows (m,s) [2,M]; combine; row step.
 * reuse_prepared != 0: the workspace still holds the operand planes (FP16 points, squared norms) an earlier
 * colstep / rowstep call prepared for the SAME (x_local, y, sizes) - the preparation passes are skipped; the
 * caller must then hand in the same, otherwise untouched, workspace.
 * reuse_prepared == 2: in addition the previous call of this half-step on this workspace was the previous Sinkhorn
 * iteration (its biases and partial log-sum-exps are still there): the fused engine runs its bounded-shift mode, one
 * sweep per cost tile instead of an online maximum.
 * otk_lse_combine: part_max / part_sum rows are `part_stride` floats apart (e.g. 2*M for an all-gathered
 * [ranks, 2, M] buffer). */
int otk_sinkhorn_points_colstep(const float* x_local, const float* y, int64_t n_local, int64_t M,
                                int64_t dim, const float* u_local, int cost_kind, double scale, double reg,
                                int precision, int reuse_prepared, float* col_max, float* col_sum,
                                void* workspace, size_t workspace_bytes, otk_stream_t stream);
int otk_lse_combine(const float* part_max, const float* part_sum, int64_t parts, int64_t part_stride,
                    int64_t M, const float* b, float* v, float* diff /* += sum|dv| */,
                    otk_stream_t stream);
int otk_sinkhorn_points_rowstep(const float* x_local, const float* y, int64_t n_local, int64_t M,
                                int64_t dim, const float* a_local, const float* v, int cost_kind,
                                double scale, double reg, int precision, int reuse_prepared,
                                float* u_local, float* diff /* += sum|du| */, void* workspace,
                                size_t workspace_bytes, otk_stream_t stream);
/* Row-sharded path with the exchange of the column partials fused into the kernels (no collective call):
 * every rank owns an exchange buffer of otk_sinkhorn_exchange_bytes(world, M) bytes in SYMMETRIC memory (mapped into all
 * peers of the node, zero-filled once), `peer_buffers_dev` is a device array [world] with each rank's mapping of it.
 *   otk_sinkhorn_points_colstep_push: colstep whose final kernel stores this rank's (max, sumexp) partials [2, M] into every
 *       peer's buffer over NVLink and releases a per-source flag (fused tcgen05 engine only: OTK_ERR_INVALID_ARGUMENT otherwise)
 *   otk_lse_combine_wait: acquires the flags of all ranks, reduces the partials found in the LOCAL buffer,
 *       v = log(b + 1e-8) - LSE, diff += sum |dv|, and advances the iteration counter.
 * ctrl: 4 device ints per solver instance, zeroed once ([0] iteration counter, [2] != 0 after a wait timed out: a peer
 * did not arrive within ~2 s).  One push + one combine per iteration, same order on every rank. */
size_t otk_sinkhorn_exchange_bytes(int world, int64_t M);
int otk_sinkhorn_points_colstep_push(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                     const float* u_local, int cost_kind, double scale, double reg, int precision,
                                     int reuse_prepared, void* const* peer_buffers_dev, int world, int rank, int* ctrl,
                                     void* workspace, size_t workspace_bytes, otk_stream_t stream);
int otk_lse_combine_wait(void* exchange_local, int world, int64_t M, const float* b, float* v,
                         float* diff /* += sum|dv| */, int* ctrl, otk_stream_t stream);

/* One whole row-sharded Sinkhorn iteration (v-step with the peer-memory exchange, then u-step on the local rows) in five
 * launches: column pass -> finalize + push -> wait + combine -> row pass -> finish; each finishing kernel also produces the
 * operand bias, the bounded-shift bound and the partial log-sum-exps of the next pass.  `stage` 0 = first iteration of a
 * solve (u_local / v hold the initial potentials, the workspace is prepared), >= 1 = steady state (bounded-shift mode).
 * diffs [2] (device, may be NULL): {sum |du| over the local rows, sum |dv|} of this iteration.  Same exchange buffer /
 * ctrl contract as otk_sinkhorn_points_colstep_push; the dedicated workspace must not be touched between iterations. */
int otk_sinkhorn_points_sharded_step(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                     const float* a_local, const float* b, float* u_local, float* v, int cost_kind,
                                     double scale, double reg, int precision, int stage, void* const* peer_buffers_dev,
                                     int world, int rank, void* exchange_local, int* ctrl, float* diffs, void* workspace,
                                     size_t workspace_bytes, otk_stream_t stream);

/* Plan statistics of a row shard without the plan: pi_ij = exp(u_i + v_j - scale*cost(x_i, y_j)/reg) for the LOCAL rows.
 *   part [4] fp64: <C,pi> over the local rows, their mass, max_i |sum_j pi_ij - a_i|, local max_j |col_partial_j - b_j|
 *   row_marginal [n_local] (may be NULL), col_partial [M] = sum over the local rows of pi_ij.
 * Row-sharded use (the `check` of the multi-GPU bench, SURVEY 8e "<C,pi> and column marginals need one more SUM
 * allreduce"): SUM-reduce part[0..1] and col_partial over the ranks, MAX-reduce part[2]. */
int otk_sinkhorn_points_summary(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                const float* a_local, const float* b, const float* u_local, const float* v,
                                int cost_kind, double scale, double reg, int precision, int reuse_prepared,
                                double* part, float* row_marginal, float* col_partial, void* workspace,
                                size_t workspace_bytes, otk_stream_t stream);
/* The plan on request: plan [N,M] = exp(u_i + v_j - scale*cost(x_i, y_j)/reg) from the potentials of otk_sinkhorn_points
 * (DiscreteTransport.transport_matrix, transport/discrete_transport.py:60-77; the coupling batch_ot_gmm returns,
 * w2_utils.py:266-269).  workspace >= (N+M)*4 + 512 bytes. */
int otk_sinkhorn_points_plan(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, const float* u,
                             const float* v, int cost_kind, double scale, double reg, float* plan, void* workspace,
                             size_t workspace_bytes, otk_stream_t stream);
/* max_ij cost(x_i, y_j) -> *out (device fp32), for the 1/max normalisation */
int otk_cost_max(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, int cost_kind,
                 float* out, void* workspace, size_t workspace_bytes, otk_stream_t stream);
/* materialise scale*cost(x_i,y_j) -> C [N,M] fp32 (CodebookModel.energy, codebook_model.py:155-160; the `cdist`
 * contraction).  workspace >= (N+M)*4 + 512 bytes runs the 64x64 FFMA tile kernel; with otk_cost_workspace_bytes(N, M, dim)
 * (room for the TF32 hi/lo planes of both clouds) problems of >= 2^20 entries with dim % 4 == 0, M % 4 == 0 run the
 * contraction on tcgen05 (3xTF32) and finish the cost in one elementwise pass.  (otk_cost_max: (N+M)*4 + 512 bytes) */
size_t otk_cost_workspace_bytes(int64_t N, int64_t M, int64_t dim);
int otk_cost_matrix(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, int cost_kind,
                    double scale, float* C, void* workspace, size_t workspace_bytes, otk_stream_t stream);

/* Streaming k-means step of CodebookModel in 'argmax' mode with the Euclidean energy (MixtureMixin.assign +
 * kmean_iteration, ot/distribution_models/base.py:206-253; CodebookModel.energy, codebook_model.py:155-160) without the
 * [B, K] energy / one-hot matrices: index[l,b] = argmin_k |x[l,b] - codebook[l,k]|_2 (first index on ties),
 * weights_sum[l,k] = number of samples assigned to k, samples_sum[l,k,:] = their sum (both overwritten, dtype buf_dtype).
 *   x [L,B,dim], codebook [L,K,dim] fp32; index int64 [L,B] or NULL; weights_sum / samples_sum both given or both NULL.
 * otk_kmeans_assign_workspace_bytes: the minimum (64 x 64 FFMA distance tiles).  With otk_kmeans_workspace_bytes(L, B, K,
 * dim) - room for a chunk of the score matrix that stays in L2 and for the TF32 hi/lo planes - problems with B, K >= 256,
 * B * K >= 2^22, dim >= 192, dim % 4 == 0, K % 4 == 0 run the contraction x . codebook^T on tcgen05 (3xTF32) chunk by chunk and find
 * the nearest codeword with one pass over the chunk (the [B, K] matrix never exists in HBM-resident form). */
size_t otk_kmeans_assign_workspace_bytes(int64_t L, int64_t B, int64_t K);
size_t otk_kmeans_workspace_bytes(int64_t L, int64_t B, int64_t K, int64_t dim);
int otk_kmeans_assign(const float* x, int64_t L, int64_t B, int64_t K, int64_t dim, const float* codebook,
                      int64_t* index, void* weights_sum, void* samples_sum, int buf_dtype, void* workspace,
                      size_t workspace_bytes, otk_stream_t stream);

/* fp32-accurate GEMMs on the tensor cores used inside the kernels above, exported for the kernel unit tests:
 *   otk_gemm_nt: C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C      (B K-major)
 *   otk_gemm_nn: C[M,N] = alpha * A[M,K] * B[K,N]   + beta * C      (B N-major, the Newton-Schulz case)
 * row-major, leading dims lda/ldb/ldc, batched with element strides.
 * engine: 0 = auto, 1 = FFMA (any shape), 2 = tcgen05 3xTF32, 3 = tcgen05 1xTF32 (2/3 fail if the shape is not
 * eligible: dims >= 64, 16-byte aligned rows).  Workspace (otk_gemm_workspace_bytes) holds the TF32 hi/lo planes. */
size_t otk_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch);
int otk_gemm_nt(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                int64_t ldb, int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                float alpha, float beta, int engine, void* workspace, size_t workspace_bytes,
                otk_stream_t stream);
int otk_gemm_nn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                int64_t ldb, int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                float alpha, float beta, int engine, void* workspace, size_t workspace_bytes,
                otk_stream_t stream);

/* Peak probes for the roofline denominators that MEASURED_PEAKS.json does not hold ("builder must measure"):
 * kind 0 = dense tcgen05 kind::tf32, 1 = dense tcgen05 kind::f16 (TFLOP/s; CTA pairs, operands resident on chip, no loads),
 * 2 = MUFU.EX2 (1e12 ex2/s), 3 = DFMA (TFLOP/s).  Synchronous (CUDA events, best of 3 launches); used by bench.py only. */
int otk_microbench_peak(int kind, double* result_host, otk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OTK_H_ */
