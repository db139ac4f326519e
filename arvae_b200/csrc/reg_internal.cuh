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
    int *row_sign_out = nullptr;  // [n_rows, R] or null: sum_j sign(a_i - a_j) as the pair kernels classified it
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
    size_t off_U, off_A, off_pgrad, off_prow, off_lossp, off_psign, bytes;
};

// One pair.  With xs = sgn(f) x:  d = xs_i - xs_j (exact sign of t), r = 1/(1+2^(|c| d)) = (1 - t)/2
//   =>  t = 1 - 2r,  1 - t^2 = 4 (r - r^2),  v = t - s = (1 - s) - 2r
//   loss += |v| ;  grad += sgn(v) * (r - r^2)          (x4 applied at the end)
// sgn(v) with sgn(0) = 0, as abs-backward has it: for s != 0 it is -s (or the factor (r - r^2) is
// 0); for a tie it is sgn(t) = sgn(d), which MUST come from d itself -- near d = 0 the MUFU
// approximations cannot be trusted for the sign of 1 - 2r, and (1 - t^2) is at its maximum there.
//   q = -s * 2^127 + d * 2^60 ;  sgn = clamp(q, -1, 1)
// is exact whenever |d| >= 2^-60 or d == 0 (d * 2^60 only outweighs 2^127 when tanh is saturated
// and the factor is 0 anyway).
template <bool GRAD, bool SIGNS = false>
__device__ __forceinline__ void pair_general(float xi, float ai, float xj, float aj, float cabs,
                                             float &lacc, float &gacc, float *kacc = nullptr) {
    const float d = xi - xj;
    const float e = ex2_approx(d * cabs);
    const float r = rcp_approx(e + 1.0f);
    const float gt = ai > aj ? 1.0f : 0.0f;
    const float lt = ai < aj ? 1.0f : 0.0f;
    const float k = (1.0f - gt) + lt;  // 1 - s  in {0,1,2}
    const float v = fmaf(-2.0f, r, k);
    lacc += fabsf(v);
    if (SIGNS) *kacc += k;  // small integers: exact
    if (GRAD) {
        const float w4 = fmaf(-r, r, r);
        const float q = fmaf(k - 1.0f, 1.7014118e38f, d * 1.1529215e18f);
        const float sg = fminf(fmaxf(q, -1.0f), 1.0f);
        gacc = fmaf(sg, w4, gacc);
    }
}

DenseLayout dense_layout(int64_t B_total, int64_t n_rows, int R, int sm_count);
int run_reg_dense(const RegProblem &P, const DenseLayout &L, char *ws, cudaStream_t st);
int run_scatter_bwd(const float *grad_cols, const float *grad_out, const RegDims &dims, int R,
                    int64_t n_rows, int64_t Z, float *grad_z, int64_t gzrs, cudaStream_t st);
int run_pack_slice(const float *z, int64_t zrs, int64_t zcs, const float *lab, int64_t lrs, int64_t lcs,
                   const RegDims &dims, int R, int64_t n_rows, float *out, cudaStream_t st);
size_t head_fused_ws_bytes(int64_t B, int R);
int run_head_fused_fwd(const float *loc, const float *sd, int sd_is_log, const float *eps, int64_t B, int64_t Z,
                       const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, float beta, float capacity,
                       float gamma, float factor, float *z_out, float *scale_out, float *kld_mean_out, float *kld_loss_out,
                       float *kcoef_out, float *reg_loss_out, float *grad_cols_out, char *ws, size_t ws_bytes, cudaStream_t st);
int run_extract_perm(const unsigned long long *keys, int64_t B, int32_t *perm, cudaStream_t st);
int run_sign_matrix(const float *a, int64_t stride, int64_t B, int8_t *out, cudaStream_t st);

// Layout of the small int array `flags` of a sorted-path call: [0, 32) whole-dim two-MUFU flags (unsegmented keys),
// [32] plan error, [33] some element of some dim is an outlier or outside the shared-reciprocal range (|u| > 31), [34] "last CTA" ticket of the plan kernel,
// [35, 67) per-dim non-finite latent bits (1: +-inf present, 2: NaN present), [67, 99) n_in per dim.
constexpr int kFlagError = ARVAE_MAX_REG_DIMS;
constexpr int kFlagNeedComplete = ARVAE_MAX_REG_DIMS + 1;
constexpr int kFlagTicket = ARVAE_MAX_REG_DIMS + 2;
constexpr int kFlagNonFinite = ARVAE_MAX_REG_DIMS + 3;
constexpr int kFlagNIn = 2 * ARVAE_MAX_REG_DIMS + 3;
constexpr int kFlagInts = 3 * ARVAE_MAX_REG_DIMS + 3;
constexpr int kFlagClearInts = kFlagNIn;  // n_in is always written, the rest is cleared per call

// How sort.cu builds its 64-bit keys (see make_sort_key).
struct KeySpec {
    const float *lab;
    int64_t lrs, lcs;
    const float *z;      // non-null: segmented keys [outlier:1][sortable attribute:32][index:31] of the reg path
    int64_t zrs, zcs;
    float fsign, cabs;   // u = cabs * sgn(f) z decides "outlier" (|u| > kMufu1MaxAbsU)
    int segment;         // 0: never set the outlier bit (one attribute order per dim)
    int64_t idx_offset;  // added to the local sample index (row-block shards carry GLOBAL indices)
    RegDims dims;
};
constexpr float kMufu1MaxAbsU = 62.0f;
// The pair kernel's build for the common case (reg_sorted.cu: ONLY1) shares one reciprocal between two pairs, 1 / (a b) with
// a, b <= 1 + 2^(2 kSharedMaxAbsU): it runs only when every element of every dim is within this range (flag
// kFlagNeedComplete stays 0), the complete build otherwise.
constexpr float kSharedMaxAbsU = 31.0f;
constexpr unsigned long long kKeyIdxMask = 0x7FFFFFFFull;  // segmented keys: low 31 bits = sample index
__host__ __device__ static inline bool key_is_outlier(unsigned long long k) { return (k >> 63) != 0; }
__host__ __device__ static inline unsigned int key_sortable_attr(unsigned long long k) { return (unsigned int)(k >> 31); }

// ---- sorted runs (sort.cu: chunk_sort_kernel; reg_shard.cuh: runs_merge_kernel) ----------------------------
// The batch is argsorted as T runs of <= kRunCap consecutive samples, each radix-sorted by one CTA, and the runs are
// merged by rank (binary searches) into the global order.  On several GPUs the runs of a rank are its own samples,
// stored straight into every peer's run slots over NVLink.  An element is one 16-byte store: it carries the step's
// epoch, so a reader can tell on its own whether the element has arrived -- no fences or flags on the publish side.
constexpr int kRunCap = 8192;
constexpr int kMaxRuns = 64;
constexpr int kMaxShardRanks = 16;
struct __align__(16) RunElem {
    unsigned long long key;  // [outlier:1][sortable attribute:32][global index:31]
    float xs;                // sgn(f) * latent
    unsigned int epoch;      // low 32 bits of the step's epoch (0 never occurs: buffers start zeroed)
};
// A run region holds, per (run, dim), kRunCap element slots followed by kRunPivots pivot slots: pivot j is a copy of
// element kPivotStep * j, so that a reader can bound any key's rank in the run to one bucket with a single small load.
constexpr int kPivotStep = 64;
constexpr int kRunPivots = kRunCap / kPivotStep;
constexpr int kRunSlotElems = kRunCap + kRunPivots;
struct RunDest {             // where chunk_sort_kernel stores its sorted run: slot (run, dim) of every destination buffer
    int n_dest, R_cap;
    char *base[kMaxShardRanks];  // first byte of each destination's run region
};
__host__ __device__ static inline RunElem *run_slot(char *runs_base, int R_cap, int run, int r) {
    return reinterpret_cast<RunElem *>(runs_base) + ((int64_t)run * R_cap + r) * kRunSlotElems;
}

// ---- row-block sharding over NVLink peer memory (reg_shard.cuh) -------------------------------------------

// First bytes of every rank's communication buffer.
struct ShardHeader {
    unsigned long long epoch;   // completed sharded steps; a step's kernels signal / wait for epoch + 1
    int status;                 // != 0: a wait timed out (a peer never signalled); the step's loss comes out NaN
    unsigned int done_pub, done_pair, done_fin;  // "last CTA" tickets of the publish / pair / finalize kernels
    long long loss_part[2];     // this rank's exact loss partial (hi, lo), pulled by every peer
};

// What the kernels of a sharded step need to reach their peers: passed by value.  Every rank's communication buffer
// has the same layout (offsets below); peer[h] is rank h's buffer as mapped into this process (peer[g] = its own).
struct ShardView {
    int G, g;                   // ranks, this rank; G == 0: not a sharded step
    int R_cap, Gc;              // dims the run slots are sized for; pair-kernel CTAs per rank
    int64_t n_cap;              // most rows a rank may hold
    int64_t row_off[kMaxShardRanks + 1];  // global index of every rank's first row in this step; [G] = B
    char *peer[kMaxShardRanks];
    size_t off_flagB, off_runs, off_acc;
};

// attribute-sorted path (sort.cu, reg_sorted.cu)
struct SortedLayout {
    int64_t N;      // power-of-two size of the key arrays (sort padding)
    int64_t Bpad;   // columns padded to a multiple of kSubCols
    int n_row_tiles, S;
    int64_t n_rr, F;
    int G_max;
    int n_runs;     // sorted runs of <= kRunCap samples when the whole batch is sorted as runs (0: bitonic network instead)
    size_t acc_bytes;
    size_t off_keys, off_Us, off_As, off_Es, off_perm, off_rowpos, off_flags, off_blockcnt, off_cls8, off_cost8, off_combo, off_prefix, off_runs, off_acc_g, off_acc_l,
        off_acc_s, off_lossp, off_dbg, off_colpart, off_eloss, bytes;
};
int64_t sort_padded_size(int64_t B);
int run_sort_keys(const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, int64_t B,
                  int64_t N, unsigned long long *keys, cudaStream_t st);
int run_sort_keys_spec(const KeySpec &spec, int R, int64_t B, int64_t N, unsigned long long *keys, cudaStream_t st);
// Radix-sorts the n samples of `spec` as ceil(n / kRunCap) runs (global run indices first_run ...) and stores each run
// into slot (run, dim) of every destination.  epoch_ctr: device counter of completed steps (elements carry *epoch_ctr
// + 1), or null (elements carry 1: single GPU, stream order is enough).
int run_chunk_sort(const KeySpec &spec, int R, int64_t n, int first_run, const RunDest &dest,
                   const unsigned long long *epoch_ctr, cudaStream_t st);
SortedLayout sorted_layout(int64_t B_total, int64_t n_rows, int R, int sm_count, bool with_triangle = false);
int run_reg_sorted(const RegProblem &P, const SortedLayout &L, char *ws, cudaStream_t st);
constexpr int64_t kSortedMinBatch = 8192;  // ARVAE_ALGO_AUTO switches to the sorted path from here

// One rank's state of the sharded step: its NVLink-visible communication buffer, the peers' mappings, a private workspace.
struct ShardCtx {
    int G = 0, g = 0, R_cap = 0, device = 0;
    int64_t n_cap = 0;
    char *comm = nullptr;       // cudaMalloc'ed here (exported through CUDA IPC)
    size_t comm_bytes = 0;
    char *peer[kMaxShardRanks] = {};
    bool opened[kMaxShardRanks] = {};   // mapped with cudaIpcOpenMemHandle (to be closed)
    size_t off_flagB = 0, off_runs = 0, off_pos = 0, off_acc = 0;
    int runs_cap = 0;           // run slots per buffer
    char *ws = nullptr;         // private workspace (sorted columns, plan, ...)
    size_t ws_bytes = 0, off_mypos = 0;
    float *h_z = nullptr, *h_lab = nullptr, *h_gz = nullptr;  // device staging of the host-buffer entry
    double *h_loss_host = nullptr;  // pinned, mapped: the finalize kernel writes the loss straight into host memory
    size_t h_z_bytes = 0, h_lab_bytes = 0;
    cudaStream_t h_aux = nullptr;   // second copy stream of the host-buffer entry (labels travel beside the latents)
    cudaEvent_t h_ev = nullptr;
};
struct ShardStep {
    const float *z; int64_t zrs, zcs;       // this rank's latents [n_local, *]
    const float *lab; int64_t lrs, lcs;     // this rank's attributes
    RegDims dims; int R;
    int64_t n_all[kMaxShardRanks];          // rows of every rank
    float gamma, factor;
    double *loss_out; float *loss_f32_out;  // [1] global loss (device)
    float *grad_cols_out;                   // [n_local, R] or null
    float *grad_z_out = nullptr;            // [n_local, grad_z_cols] dLoss/dz written by the finalize kernel itself (device
    int64_t grad_z_cols = 0;                // memory, or host memory mapped into the device), or null
    int phases;                             // bit 0: sort + publish, bit 1: rank own runs, bit 2: apply + plan + pair kernel, bit 3: finalize; 0 = all
};
size_t shard_comm_bytes(int64_t n_cap, int R_cap, int G, ShardCtx *fill);
size_t shard_ws_bytes(int64_t n_cap, int R_cap, int G, size_t *off_mypos);
int run_shard_step(ShardCtx &C, const ShardStep &S, cudaStream_t st);
int shard_set_wait_ms(long long ms);  // bound of every wait for a peer on the current device (default 30 s)

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
                        float *dloc, float *dscale, cudaStream_t st, int sd_is_log = 0);

int run_measure_attributes(const long long *measures, int64_t B, int64_t T, int64_t row_stride, const int *lut,
                           int64_t V, const float *weights, float *out, cudaStream_t st);

int sm_count();

}  // namespace arvae
