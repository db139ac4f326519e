/*
 * arvae_b200.h -- C ABI of libarvae_b200.so: AR-VAE's attribute-regularization
 * hot path as hand-written CUDA for sm_100a (NVIDIA B200).
 *
 * The reference (ashispati/ar-vae) is pure Python/PyTorch and has no FFI of its
 * own; the "interface" each entry point replaces is the reference function it
 * computes.  All file:line citations are relative to the reference tree.
 *
 *   arvae_reg_loss_fwdbwd_f32      utils/trainer.py:369-403  Trainer.compute_reg_loss +
 *                                   reg_loss_sign, the per-dim caller loop
 *                                   imagevae/image_vae_trainer.py:171-180 /
 *                                   measurevae/measure_vae_trainer.py:131-142, and the autograd
 *                                   backward that utils/trainer.py:140 (loss.backward()) triggers
 *   arvae_reg_loss_scatter_bwd_f32 the select-backward / accumulate step of that autograd graph
 *   arvae_latent_head_fwd_f32      imagevae/mnist_vae.py:74-87 (reparametrize, rsample = loc+eps*scale),
 *                                   measurevae/measure_vae.py:115-123, utils/trainer.py:354-367
 *                                   (compute_kld_loss vs the unit prior)
 *   arvae_latent_head_bwd_f32      autograd backward of the two above
 *   arvae_reg_loss_host_f32        the same compute_reg_loss loop, host buffers in / out
 *                                   (what a non-torch caller would bind)
 *   arvae_reg_sign_matrix_i8       utils/trainer.py:394-395,400 (attribute sign matrix; debug/parity)
 *
 * Conventions
 *   - plain pointers and sizes, no torch types; `stream` is a cudaStream_t passed as void*.
 *   - pointers named *_dev are device pointers on the current device; *_host are host pointers.
 *   - every function returns 0 on success, a negative ARVAE_E_* code for argument errors and a
 *     positive cudaError_t for CUDA failures; arvae_last_error() returns a thread-local message.
 *   - all device work is stream-ordered; no function synchronises the device or allocates
 *     device memory except arvae_reg_loss_host_f32 (which owns its buffers and synchronises
 *     `stream` before returning).  Entry points are re-entrant.
 *   - "rows" are samples i of the batch; "columns" are the samples j they are paired with.
 *     A call covers the row block [row_begin,row_end) against ALL B_total columns, so one entry
 *     serves single-GPU (0..B) and row-block sharding across GPUs.
 */
#ifndef ARVAE_B200_H
#define ARVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARVAE_VERSION 200

#if defined(__GNUC__)
#define ARVAE_API __attribute__((visibility("default")))
#else
#define ARVAE_API
#endif

#define ARVAE_E_BADARG (-1)    /* null pointer, negative size, R out of range ... */
#define ARVAE_E_WORKSPACE (-2) /* workspace too small */
#define ARVAE_E_NODEVICE (-3)  /* no CUDA device / wrong architecture */

#define ARVAE_MAX_REG_DIMS 32

/* algorithm selector for arvae_reg_loss_fwdbwd_f32 */
#define ARVAE_ALGO_AUTO 0
#define ARVAE_ALGO_DENSE 1  /* every tile through the general pair loop (float compares)      */
#define ARVAE_ALGO_SORTED 2 /* rows/columns ordered by attribute, constant-sign tile fast path */
#define ARVAE_ALGO_TRIANGLE 3 /* sorted + each constant-sign tile evaluated once for both sides (all rows only) */

ARVAE_API int arvae_version(void);
ARVAE_API const char *arvae_last_error(void);

/* Number of SMs of the current device (cached); <0 on error. */
ARVAE_API int arvae_device_sm_count(void);

/* Bytes of scratch the fused forward+backward needs for this shape. */
ARVAE_API size_t arvae_reg_loss_workspace_bytes(int64_t B_total, int64_t n_rows, int32_t R);
/* Same for an explicit algorithm (ARVAE_ALGO_TRIANGLE needs ~B^2 R / 128 bytes more for its column sums). */
ARVAE_API size_t arvae_reg_loss_workspace_bytes_algo(int64_t B_total, int64_t n_rows, int32_t R, int32_t algo);

/*
 * Fused forward + backward of  sum_r gamma * mean_ij | tanh(factor*(z[i,d_r]-z[j,d_r])) - sign(a[i,c_r]-a[j,c_r]) |
 * restricted to rows i in [row_begin,row_end), j over all B_total samples.
 *
 *   z_dev          [B_total, *] float, element strides (z_row_stride, z_col_stride)
 *   labels_dev     [B_total, *] float, element strides (lab_row_stride, lab_col_stride)
 *   reg_dims_host  [R] latent column d_r   (0 <= d_r)
 *   label_cols_host[R] label column c_r    (the trainers use c_r == d_r)
 *   loss_out_dev   [1] double: the block's share of the loss, ALREADY scaled by gamma / B_total^2
 *                  (shares of disjoint row blocks add up to the reference's value)
 *   loss_f32_out_dev [1] float or NULL: the same value rounded to float (the reference's dtype)
 *   grad_cols_out_dev [n_rows, R] float or NULL: d loss / d z[row_begin+k, d_r]  (full-matrix
 *                  gradient of the mean loss, times gamma), NULL skips the gradient math
 *   row_loss_out_dev  [n_rows, R] double or NULL: per-row unnormalised sums  sum_j |t_ij - s_ij|
 *                  (parity/debug output)
 *   row_sign_out_dev  [n_rows, R] int32 or NULL: sum_j sign(a[i,c_r] - a[j,c_r]) accumulated by the PAIR KERNEL from
 *                  the very classification / compares it evaluated the loss with (reference utils/trainer.py:394-395,
 *                  400) -- the integer check that the attribute sign matrix of the hot path is bit-exact.  Needs
 *                  grad_cols_out_dev; not offered by ARVAE_ALGO_TRIANGLE.
 */
ARVAE_API int arvae_reg_loss_fwdbwd_f32(const float *z_dev, int64_t z_row_stride, int64_t z_col_stride,
                              const float *labels_dev, int64_t lab_row_stride,
                              int64_t lab_col_stride, const int32_t *reg_dims_host,
                              const int32_t *label_cols_host, int32_t R, int64_t row_begin,
                              int64_t row_end, int64_t B_total, float gamma, float factor,
                              int32_t algo, double *loss_out_dev, float *loss_f32_out_dev,
                              float *grad_cols_out_dev, double *row_loss_out_dev, int32_t *row_sign_out_dev,
                              void *workspace_dev, size_t workspace_bytes, void *stream);

/*
 * Which tanh form the attribute-sorted path used in the call that last wrote `workspace_dev` (same B_total / n_rows /
 * R): flags_out_host[r] = number of INLIER samples of dim r, i.e. samples with |2 f log2(e) x| <= 62.  Pairs of two
 * inliers cost one MUFU (factorised E_j/(E_i+E_j)), pairs with an outlier two (EX2 + RCP on the latent difference):
 * the MUFU count per pair of dim r is 2 - (n_in/B)^2 up to the few tiles that straddle the segment boundary.
 * Synchronises `stream`.  Returns ARVAE_E_BADARG when the shape selects the dense path (always 2 MUFU per pair).
 */
ARVAE_API int arvae_reg_loss_path_flags(int64_t B_total, int64_t n_rows, int32_t R, int32_t algo,
                                        const void *workspace_dev, int32_t *flags_out_host,
                                        void *stream);

/*
 * grad_z[k, :] = 0 ; grad_z[k, d_r] += grad_out * grad_cols[k, r]   for k in [0,n_rows)
 *   grad_out_dev [1] float (upstream gradient of the scalar loss) or NULL for 1.0
 *   grad_z_dev   [n_rows, Z] float with row stride gz_row_stride; every element is written.
 */
ARVAE_API int arvae_reg_loss_scatter_bwd_f32(const float *grad_cols_dev, const float *grad_out_dev,
                                   const int32_t *reg_dims_host, int32_t R, int64_t n_rows,
                                   int64_t Z, float *grad_z_dev, int64_t gz_row_stride,
                                   void *stream);

/*
 * Latent head forward (one pass over [B,Z]):
 *   z       = loc + eps*scale                                   (bit-identical to Normal.rsample)
 *   kld_sum = sum_b sum_d 0.5*(scale^2 + loc^2 - 1 - log(scale^2))
 *   kld_mean = kld_sum / B ;  kld_loss = beta * |kld_mean - capacity| ;
 *   kcoef    = beta * sgn(kld_mean - capacity) / B              (what the backward needs)
 *   loc/scale/eps/z  [B,Z] float contiguous
 *   kld_sum_out_dev  [1] double;  kld_mean/kld_loss/kcoef_out_dev  [1] float each, or NULL
 *   ws_dev: scratch of arvae_latent_head_workspace_bytes(B,Z) bytes.
 */
ARVAE_API size_t arvae_latent_head_workspace_bytes(int64_t B, int64_t Z);
ARVAE_API int arvae_latent_head_fwd_f32(const float *loc_dev, const float *scale_dev, const float *eps_dev,
                              int64_t B, int64_t Z, float beta, float capacity, float *z_out_dev,
                              double *kld_sum_out_dev, float *kld_mean_out_dev,
                              float *kld_loss_out_dev, float *kcoef_out_dev, void *ws_dev,
                              size_t ws_bytes, void *stream);

/*
 * Latent head backward (one pass over [B,Z]):
 *   dloc   = dz + k * loc
 *   dscale = dz * eps + k * (scale - 1/scale)
 * with dz[b,d] = dz_up[b,d] + greg * grad_cols[b,r] (summed over r with d_r == d) and
 *   k = kscale * (kcoef_dev ? *kcoef_dev : 1) * (gkld_dev ? *gkld_dev : 1).
 *   dz_up_dev [B,Z] or NULL (-> 0); grad_cols_dev [B,R] or NULL; greg_dev [1] float or NULL (-> 1)
 *   dloc_dev / dscale_dev [B,Z] float, either may be NULL.
 */
ARVAE_API int arvae_latent_head_bwd_f32(const float *loc_dev, const float *scale_dev, const float *eps_dev,
                              const float *dz_up_dev, const float *grad_cols_dev,
                              const float *greg_dev, const int32_t *reg_dims_host, int32_t R,
                              float kscale, const float *kcoef_dev, const float *gkld_dev,
                              int64_t B, int64_t Z, float *dloc_dev, float *dscale_dev,
                              void *stream);

/*
 * The whole latent-loss head in ONE launch (csrc/head_fused.cu), for 1 <= B <= 8192 -- the sizes the reference
 * trains at (B = 128 / 256: train_image_vae.py:16, train_measure_vae.py:35):
 *   scale = sd, or exp(sd) when sd_is_log_std (imagevae/mnist_vae.py:63-65, measurevae/encoder.py:120-123)
 *   z        = loc + eps*scale                                    (mnist_vae.py:79, measure_vae.py:116)
 *   kld_mean, kld_loss = beta*|kld_mean - capacity|, kcoef        (utils/trainer.py:354-367; as latent_head_fwd)
 *   reg_loss = sum_r gamma * mean_ij |tanh(factor (z_i - z_j)) - sign(a_i - a_j)| over z[:, reg_dims[r]],
 *              labels[:, label_cols[r]]                          (utils/trainer.py:369-403 via the trainers' loop)
 *   grad_cols[b, r] = d reg_loss / d z[b, reg_dims[r]]           (NULL: no gradient wanted)
 * loc/sd/eps/z_out/scale_out [B,Z] float contiguous; labels with element strides; scale_out / kld_* / kcoef may
 * be NULL.  workspace_dev: arvae_head_fused_workspace_bytes(B, R) bytes that the CALLER ZEROES ONCE (e.g. at
 * allocation); every launch leaves it zeroed again, so the same buffer serves every later call on the same stream
 * (no memset per call, CUDA-graph capturable).  Returns ARVAE_E_BADARG for B > 8192: use arvae_latent_head_fwd_f32 +
 * arvae_reg_loss_fwdbwd_f32 there.
 *
 * Backward in one launch (arvae_latent_head_bwd_f32's pass with kscale = 1): dloc, and dsd = d/d(scale), or
 * d/d(log_std) = d/d(scale) * scale when sd_is_log_std.
 */
ARVAE_API size_t arvae_head_fused_workspace_bytes(int64_t B, int32_t R);
ARVAE_API int arvae_head_fused_fwd_f32(const float *loc_dev, const float *sd_dev, int32_t sd_is_log_std,
                             const float *eps_dev, int64_t B, int64_t Z, const float *labels_dev,
                             int64_t lab_row_stride, int64_t lab_col_stride, const int32_t *reg_dims_host,
                             const int32_t *label_cols_host, int32_t R, float beta, float capacity, float gamma,
                             float factor, float *z_out_dev, float *scale_out_dev, float *kld_mean_out_dev,
                             float *kld_loss_out_dev, float *kcoef_out_dev, float *reg_loss_out_dev,
                             float *grad_cols_out_dev, void *workspace_dev, size_t workspace_bytes, void *stream);
ARVAE_API int arvae_head_fused_bwd_f32(const float *loc_dev, const float *sd_dev, int32_t sd_is_log_std,
                             const float *eps_dev, const float *dz_up_dev, const float *grad_cols_dev,
                             const float *greg_dev, const int32_t *reg_dims_host, int32_t R,
                             const float *kcoef_dev, const float *gkld_dev, int64_t B, int64_t Z,
                             float *dloc_dev, float *dsd_dev, void *stream);

/*
 * Host-buffer form of the compute_reg_loss loop (the end-to-end call): copies z and labels to the
 * device, runs the fused forward+backward over all rows, copies the loss and dLoss/dz back.
 *   z_host [B,Z] float row-major, labels_host [B,A] float row-major
 *   loss_out_host [1] float, grad_z_out_host [B,Z] float or NULL
 * Device buffers are cached per thread between calls of the same shape; arvae_host_release() frees them.
 */
ARVAE_API int arvae_reg_loss_host_f32(const float *z_host, int64_t B, int64_t Z, const float *labels_host,
                            int64_t A, const int32_t *reg_dims_host,
                            const int32_t *label_cols_host, int32_t R, float gamma, float factor,
                            int32_t algo, float *loss_out_host, float *grad_z_out_host,
                            void *stream);
ARVAE_API void arvae_host_release(void);

/*
 * Row-block sharding helper: out[k, 0:R] = z[k, d_r], out[k, R:2R] = labels[k, c_r] for k in [0,n_rows)
 * -- the packed [n_rows, 2R] float slice each rank contributes to the all-gather of columns.
 */
ARVAE_API int arvae_pack_columns_f32(const float *z_dev, int64_t z_row_stride, int64_t z_col_stride,
                                     const float *labels_dev, int64_t lab_row_stride,
                                     int64_t lab_col_stride, const int32_t *reg_dims_host,
                                     const int32_t *label_cols_host, int32_t R, int64_t n_rows,
                                     float *out_dev, void *stream);


/*
 * The same step sharded over the GPUs of one NVSwitch box, one process per GPU, with the exchange done by the kernels
 * themselves over NVLink peer memory (csrc/reg_shard.cuh; BASELINE.json north_star's multi-GPU path):
 * every rank argsorts its own rows and stores the sorted run into every peer's buffer (the column all-gather), the
 * runs are merged into the single-GPU sorted order, each rank sweeps its 1/world share of that order's row blocks,
 * then pulls the row sums of its own samples and every rank's loss partial from the peers (gradient return +
 * all-reduce).  Loss and gradients are BITWISE those of arvae_reg_loss_fwdbwd_f32 on the concatenated batch.
 *
 *   arvae_shard_create        allocates this rank's communication buffer and workspace for up to n_cap rows per
 *                             rank and R_cap regularised dims (world <= 16)
 *   arvae_shard_ipc_handle    64-byte CUDA IPC handle of the communication buffer, to be all-gathered by the caller
 *                             (torch.distributed / MPI / a pipe: plumbing, done once)
 *   arvae_shard_open_peers    maps every peer's buffer from the gathered handles [world][64]
 *   arvae_shard_set_peer      alternative for peers that live in the same process or are mapped by other means
 *   arvae_shard_reg_loss_f32  one step.  z_local_dev / labels_local_dev hold THIS rank's n_all_host[rank] rows;
 *                             loss_out_dev [1] double (+ optional float) receives the GLOBAL loss on every rank,
 *                             grad_cols_out_dev [n_local, R] the gradient columns of this rank's rows (NULL: loss only).
 *                             `phases` = 0 runs the whole step; bits 1 | 2 | 4 | 8 run only sort+publish / rank own
 *                             runs / apply+plan+pairs / finalize (tests drive several ranks of one process through
 *                             the phases in lockstep).
 *                             Stream-ordered, no host sync, no NCCL.  A peer that never shows up makes the loss NaN
 *                             after a bounded wait (30 s; environment variable ARVAE_SHARD_WAIT_MS, read by
 *                             arvae_shard_create, overrides it; arvae_shard_status reports it) instead of hanging
 *                             the GPU.  A communicator whose wait gave up stays broken (every later step is NaN):
 *                             destroy it and create a new one.
 *   arvae_shard_reg_loss_host_f32  the same with HOST buffers in and out (H2D, step, dL/dz scatter, D2H, sync).
 */
#define ARVAE_SHARD_HANDLE_BYTES 64
ARVAE_API size_t arvae_shard_comm_bytes(int64_t n_cap, int32_t R_cap, int32_t world);
ARVAE_API int arvae_shard_create(int32_t rank, int32_t world, int64_t n_cap, int32_t R_cap, void **ctx_out);
ARVAE_API int arvae_shard_ipc_handle(void *ctx, void *handle_out);
ARVAE_API int arvae_shard_open_peers(void *ctx, const void *handles);
ARVAE_API int arvae_shard_set_peer(void *ctx, int32_t rank, void *comm_dev);
ARVAE_API void *arvae_shard_comm_ptr(void *ctx);
ARVAE_API int arvae_shard_reg_loss_f32(void *ctx, const float *z_local_dev, int64_t z_row_stride, int64_t z_col_stride,
                                       const float *labels_local_dev, int64_t lab_row_stride, int64_t lab_col_stride,
                                       const int32_t *reg_dims_host, const int32_t *label_cols_host, int32_t R,
                                       const int64_t *n_all_host, float gamma, float factor, double *loss_out_dev,
                                       float *loss_f32_out_dev, float *grad_cols_out_dev, int32_t phases, void *stream);
ARVAE_API int arvae_shard_reg_loss_host_f32(void *ctx, const float *z_local_host, int64_t Z,
                                            const float *labels_local_host, int64_t A, const int32_t *reg_dims_host,
                                            const int32_t *label_cols_host, int32_t R, const int64_t *n_all_host,
                                            float gamma, float factor, float *loss_out_host, float *grad_z_out_host,
                                            void *stream);
ARVAE_API int arvae_shard_status(void *ctx, int32_t *status_out, uint64_t *epoch_out, void *stream);
ARVAE_API int arvae_shard_destroy(void *ctx);

/*
 * perm_out_dev[k] = index of the k-th smallest attribute (ties by index, NaN last): the order the
 * attribute-sorted path uses.  workspace: arvae_attr_argsort_workspace_bytes(B) bytes.  (parity tests)
 */
ARVAE_API size_t arvae_attr_argsort_workspace_bytes(int64_t B);
ARVAE_API int arvae_attr_argsort_f32(const float *labels_dev, int64_t lab_stride, int64_t B,
                                     int32_t *perm_out_dev, void *workspace_dev,
                                     size_t workspace_bytes, void *stream);

/*
 * The four musical attributes the MeasureVAE trainer regularises (reference
 * data/dataloaders/bar_dataset.py:338-500, order of measurevae/measure_vae_trainer.py:15-20):
 * out[b] = { rhy_complexity, pitch_range, note_density, contour } for measure b.
 *   measures_dev [B, T] int64 note indices (row stride in elements); lut_dev [V] int32: MIDI pitch (>= 0) of a
 *   note symbol, or -1 slur, -2 rest, -3 None, -4 START, -5 END; rhy_weights_dev [T] float (RHY_COMPLEXITY_COEFFS);
 *   out_dev [B, 4] float.
 */
ARVAE_API int arvae_measure_attributes_i64(const int64_t *measures_dev, int64_t B, int64_t T,
                                           int64_t row_stride, const int32_t *lut_dev, int64_t V,
                                           const float *rhy_weights_dev, float *out_dev, void *stream);

/*
 * Pairwise-rank evaluation metrics of the reference (utils/evaluation.py): for latent codes [B, Z] and
 * attributes [B, A] (float32, element strides), every (code i, attribute j) pair gets
 *   rho_out_dev[i*A+j], pval_out_dev[i*A+j]  Spearman rho and its two-sided p-value, as scipy.stats.spearmanr
 *                                            returns them at utils/evaluation.py:166 (average ranks for ties;
 *                                            NaN for constant or NaN-holding columns and for B < 3);
 *   corr_out_dev[i*A+j]                      _compute_correlation_matrix (:157-173): |rho| if p <= 0.05 else 0;
 *   sap_out_dev[i*A+j]                       _compute_score_matrix (:194-214): cov^2/(var_mu var_y), ddof = 1,
 *                                            0 where var_mu <= 1e-12 (IEEE NaN/inf where var_y = 0, as numpy);
 *   scores_out_dev[0..1]                     { Corr_score (:146-155), SAP_score (:176-191, :217-219) }.
 * All outputs are device doubles; all arithmetic after the argsort is float64 like the reference's.
 * 1 <= Z <= 1024, 1 <= A <= 64, B >= 1.  workspace: arvae_eval_metrics_workspace_bytes(B, Z, A) bytes.
 * Stream-ordered, no host sync; bitwise reproducible.
 */
ARVAE_API size_t arvae_eval_metrics_workspace_bytes(int64_t B, int32_t Z, int32_t A);
ARVAE_API int arvae_eval_metrics_f32(const float *codes_dev, int64_t codes_row_stride, int64_t codes_col_stride,
                                     const float *attrs_dev, int64_t attrs_row_stride, int64_t attrs_col_stride,
                                     int64_t B, int32_t Z, int32_t A, double *rho_out_dev, double *pval_out_dev,
                                     double *corr_out_dev, double *sap_out_dev, double *scores_out_dev,
                                     void *workspace_dev, size_t workspace_bytes, void *stream);

/* s[i*B+j] = sign(a_i - a_j) as int8, for parity tests at small B (same compare the kernels use). */
ARVAE_API int arvae_reg_sign_matrix_i8(const float *labels_dev, int64_t lab_stride, int64_t B,
                             int8_t *out_dev, void *stream);

/*
 * Optional timing of the dominant (pair) kernel alone, for roofline reporting: while enabled, each
 * fused forward+backward brackets its pair-kernel launch with CUDA events on the caller's stream
 * (a ring of the last 64 launches per thread). arvae_profile_pair_kernel_ms synchronises those
 * events, returns the summed milliseconds and launch count since the last call, and clears the ring.
 */
ARVAE_API void arvae_profile_enable(int on);
ARVAE_API int arvae_profile_pair_kernel_ms(float *sum_ms_out, int *n_out);

/*
 * Step timeline for experiments: while enabled, the library records a CUDA event after each group of launches of a step
 * (sort, gather / merge, plan, pair kernel, epilogue ...) on the caller's stream.  arvae_timeline_report synchronises
 * them and writes "name:milliseconds;..." (the time between consecutive marks) into buf, then clears the list.
 */
ARVAE_API void arvae_timeline_enable(int on);
ARVAE_API int arvae_timeline_report(char *buf, int32_t buf_bytes);

/* Number of kernel launches issued through the library (all threads) since the last reset. */
ARVAE_API int64_t arvae_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* ARVAE_B200_H */
