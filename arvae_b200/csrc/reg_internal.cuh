// reg_internal.cuh -- types shared between the translation units of libarvae_b200.
#pragma once

#include "common.cuh"

namespace arvae {

constexpr int kDenseThreads = 128;     // threads per CTA of the pair kernels
constexpr int kDenseTileCols = 2048;   // columns staged in shared memory at a time (2 x 8 KiB)
constexpr int kSubCols = 256;          // columns between flushes of the fp32 accumulators to fp64

struct RegDims {
    int32_t zcol[ARVAE_MAX_REG_DIMS];  // latent column per regularised dim
    int32_t lcol[ARVAE_MAX_REG_DIMS];  // label column per regularised dim
};

// One call of the fused forward+backward (device pointers, element strides).
struct RegProblem {
    const float *z;
    int64_t zrs, zcs;
    const float *lab;
    int64_t lrs, lcs;
    RegDims dims;
    int R;
    int64_t B;  // total samples = number of columns
    int64_t row_begin, row_end;
    float gamma, factor;
    double *loss_out;      // [1]
    float *loss_f32_out;   // [1] or null
    float *grad_cols_out;  // [n_rows, R] or null
    double *row_loss_out;  // [n_rows, R] or null
    bool use_triangle = false;  // sorted path, all rows: evaluate constant-sign tiles once for both sides
};

struct DenseLayout {
    int64_t Bpad;          // columns padded to a multiple of kSubCols
    int RI;                // rows per thread
    int64_t n_row_blocks;  // CTAs along the rows
    int64_t rows_pad;
    int64_t chunk_cols;    // columns per work unit
    int n_chunks;
    int64_t n_units;
    size_t off_U, off_A, off_pgrad, off_prow, off_lossp, bytes;
};

DenseLayout dense_layout(int64_t B_total, int64_t n_rows, int R, int sm_count);
int run_reg_dense(const RegProblem &P, const DenseLayout &L, char *ws, cudaStream_t st);
int run_scatter_bwd(const float *grad_cols, const float *grad_out, const RegDims &dims, int R,
                    int64_t n_rows, int64_t Z, float *grad_z, int64_t gzrs, cudaStream_t st);
int run_pack_slice(const float *z, int64_t zrs, int64_t zcs, const float *lab, int64_t lrs, int64_t lcs,
                   const RegDims &dims, int R, int64_t n_rows, float *out, cudaStream_t st);
int run_extract_perm(const unsigned long long *keys, int64_t B, int32_t *perm, cudaStream_t st);
int run_sign_matrix(const float *a, int64_t stride, int64_t B, int8_t *out, cudaStream_t st);

// attribute-sorted path (sort.cu, reg_sorted.cu)
struct SortedLayout {
    int64_t N;      // power-of-two size of the key arrays (sort padding)
    int64_t Bpad;   // columns padded to a multiple of kSubCols
    int n_row_tiles, S;
    int64_t n_rr, F;
    int G_max, max_segs;
    size_t slot_bytes;
    size_t off_keys, off_Us, off_As, off_Es, off_perm, off_rowpos, off_flags, off_blockcnt, off_cls8, off_cost8, off_combo, off_prefix, off_pgrad, off_prow,
        off_lossp, off_dbg, off_colpart, off_eloss, bytes;
};
int64_t sort_padded_size(int64_t B);
int run_sort_keys(const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, int64_t B,
                  int64_t N, unsigned long long *keys, cudaStream_t st);
SortedLayout sorted_layout(int64_t B_total, int64_t n_rows, int R, int sm_count, bool with_triangle = false);
int run_reg_sorted(const RegProblem &P, const SortedLayout &L, char *ws, cudaStream_t st);
constexpr int64_t kSortedMinBatch = 8192;  // ARVAE_ALGO_AUTO switches to the sorted path from here

// pairwise-rank evaluation metrics (eval_metrics.cu)
constexpr int kEvalMaxCodes = 1024, kEvalMaxAttrs = 64;
size_t eval_metrics_workspace_bytes(int64_t B, int Z, int A);
int run_eval_metrics(const float *codes, int64_t crs, int64_t ccs, const float *attrs, int64_t ars, int64_t acs,
                     int64_t B, int Z, int A, double *rho, double *pval, double *corr, double *sap, double *scores,
                     char *ws, cudaStream_t st);

// latent head (latent_head.cu)
size_t latent_head_ws_bytes(int64_t B, int64_t Z);
int run_latent_head_fwd(const float *loc, const float *scale, const float *eps, int64_t B,
                        int64_t Z, float beta, float capacity, float *z_out, double *kld_sum_out,
                        float *kld_mean_out, float *kld_loss_out, float *kcoef_out, void *ws,
                        size_t ws_bytes, cudaStream_t st);
int run_latent_head_bwd(const float *loc, const float *scale, const float *eps, const float *dz_up,
                        const float *grad_cols, const float *greg, const RegDims &dims, int R,
                        float kscale, const float *kcoef, const float *gkld, int64_t B, int64_t Z,
                        float *dloc, float *dscale, cudaStream_t st);

int run_measure_attributes(const long long *measures, int64_t B, int64_t T, int64_t row_stride, const int *lut,
                           int64_t V, const float *weights, float *out, cudaStream_t st);

int sm_count();

}  // namespace arvae
