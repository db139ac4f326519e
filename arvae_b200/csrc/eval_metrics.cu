// eval_metrics.cu -- the reference's pairwise-rank evaluation metrics on the device.
//
// Reference: utils/evaluation.py:146-173 (Corr_score: |Spearman rho| of every (latent code, attribute) pair,
// kept only where scipy.stats.spearmanr's p-value is <= 0.05, then mean over attributes of the max over codes)
// and :176-219 (SAP_score: cov^2 / (var var) with ddof = 1, zero where var_mu <= 1e-12, then mean over
// attributes of top1 - top2 over codes).  The reference runs Z x A scipy / np.cov calls in Python loops over
// the same <= 25 728 samples; both matrices are second moments of (Z + A) columns, so here:
//
//   1. argsort of every column (sort.cu's batched bitonic network, <= 32 columns per pass),
//   2. avg_ranks_kernel: tie groups by binary search in the sorted keys -> average ranks, scattered back to
//      sample order (double: half-integers),
//   3. col_means_kernel: column means of the raw values (double, fixed-order tree),
//   4. eval_moments_kernel: centred cross moments of ranks and of raw values for every (code, attribute) pair
//      plus every column's own second moment -- tiles of rows staged in shared memory as doubles, every thread
//      accumulating up to four pairs; per-block partials (no floating-point atomics: bitwise reproducible),
//   5. eval_finish_kernel: partials -> rho, t, p (eval_math.cuh), gated |rho|, SAP matrix, both scores.
//
// All arithmetic after the sort is float64, as in the reference (np.cov / rankdata promote to float64).
// Latency-bound: a few hundred KB of input, microseconds of math.
#ifdef ARVAE_HOST_EMULATION
#include "cuda_emul.h"  // tests/native: CPU threads standing in for a CTA, so the GPU-less test tier runs this file
#else
#include <string.h>

#include "common.cuh"
#include "reg_internal.cuh"
#define ARVAE_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ARVAE_DYN_SMEM(type, name) extern __shared__ type name[]
#endif
#include "eval_math.cuh"

namespace arvae {

constexpr int kEvalThreads = 256;
constexpr int kEvalPairsPerThread = 4;
constexpr int kEvalPairsPerBlock = kEvalThreads * kEvalPairsPerThread;
constexpr int kEvalMaxBlocks = 592;          // row-tile CTAs of the moments kernel (4 x 148)
constexpr int kEvalSmemBytes = 40 * 1024;    // row tile [C][TR + 1] doubles; below the 48 KiB static limit

// inverse of sort.cu's float_to_sortable
__device__ __forceinline__ float sortable_to_float(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ float key_value(unsigned long long k) { return sortable_to_float((unsigned int)(k >> 32)); }

// ranks[c][sample] = average 1-based rank of the sample in column c (scipy.stats.rankdata 'average': tied values
// share the mean of their positions; -0.0 == +0.0).  keys[c][0..B) are sorted; NaNs sort last and only raise the
// column's flag (spearmanr propagates NaN).
__global__ void __launch_bounds__(kEvalThreads)
avg_ranks_kernel(const unsigned long long *__restrict__ keys, int64_t N, int64_t B, double *__restrict__ ranks,
                 int *__restrict__ nanflag) {
    const int c = blockIdx.y;
    const int64_t p = (int64_t)blockIdx.x * kEvalThreads + threadIdx.x;
    if (p >= B) return;
    const unsigned long long *k = keys + (int64_t)c * N;
    const unsigned long long kp = k[p];
    const float v = key_value(kp);
    const int64_t sample = (int64_t)(unsigned int)kp;
    double *out = ranks + (int64_t)c * B + sample;
    if (v != v) {
        nanflag[c] = 1;  // every writer stores the same value
        *out = 0.0;
        return;
    }
    int64_t lo = p, hi = p + 1;
    if (p > 0 && key_value(k[p - 1]) == v) {  // first position whose value is not below v
        int64_t l = 0, r = p;
        while (l < r) {
            const int64_t m = (l + r) >> 1;
            if (key_value(k[m]) < v) l = m + 1; else r = m;
        }
        lo = l;
    }
    if (p + 1 < B && key_value(k[p + 1]) == v) {  // first position whose value is above v (or NaN)
        int64_t l = p + 1, r = B;
        while (l < r) {
            const int64_t m = (l + r) >> 1;
            if (key_value(k[m]) <= v) l = m + 1; else r = m;
        }
        hi = l;
    }
    *out = 0.5 * (double)(lo + hi + 1);  // positions lo .. hi-1 hold ranks lo+1 .. hi
}

struct EvalInputs {
    const float *codes;
    int64_t crs, ccs;
    const float *attrs;
    int64_t ars, acs;
    int64_t B;
    int Z, A;
};

__device__ __forceinline__ double eval_value(const EvalInputs &in, int64_t row, int c) {
    return c < in.Z ? (double)__ldg(in.codes + row * in.crs + (int64_t)c * in.ccs)
                    : (double)__ldg(in.attrs + row * in.ars + (int64_t)(c - in.Z) * in.acs);
}

__device__ __forceinline__ double block_sum_256(double v, double *red) {
    v = warp_sum(v);  // fixed butterfly: the same bits on every run
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kEvalThreads / 32; ++w) s += red[w];
    }
    return s;  // valid in thread 0
}

__global__ void __launch_bounds__(kEvalThreads)
col_means_kernel(EvalInputs in, double *__restrict__ mean) {
    __shared__ double red[kEvalThreads / 32];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < in.B; i += kEvalThreads) s += eval_value(in, i, c);
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) mean[c] = s / (double)in.B;
}

// partial[(block * 2 + view) * NP + q]: view 0 = raw values, view 1 = ranks.  q < Z*A: pair (code q / A,
// attribute q % A); q >= Z*A: column q - Z*A with itself.  NP = Z*A + Z + A.
__global__ void __launch_bounds__(kEvalThreads)
eval_moments_kernel(EvalInputs in, const double *__restrict__ ranks, const double *__restrict__ mean, int TR,
                    int64_t n_tiles, double *__restrict__ partial) {
    ARVAE_DYN_SMEM(double, tile);  // [C][TR + 1]: odd pitch, threads of a warp read different columns
    const int Z = in.Z, A = in.A, C = Z + A, ZA = Z * A, NP = ZA + C;
    const int pitch = TR + 1;
    int ca[kEvalPairsPerThread], cb[kEvalPairsPerThread];
    bool act[kEvalPairsPerThread];
#pragma unroll
    for (int q = 0; q < kEvalPairsPerThread; ++q) {
        const int p = blockIdx.y * kEvalPairsPerBlock + q * kEvalThreads + threadIdx.x;
        act[q] = p < NP;
        ca[q] = cb[q] = 0;
        if (act[q]) {
            if (p < ZA) { ca[q] = p / A; cb[q] = Z + p % A; }
            else ca[q] = cb[q] = p - ZA;
        }
    }
    double acc[2][kEvalPairsPerThread];
#pragma unroll
    for (int q = 0; q < kEvalPairsPerThread; ++q) acc[0][q] = acc[1][q] = 0.0;
    const double rank_mean = 0.5 * ((double)in.B + 1.0);

    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t row0 = t * TR;
        const int nr = (int)min((int64_t)TR, in.B - row0);
#pragma unroll
        for (int view = 0; view < 2; ++view) {
            __syncthreads();  // previous readers of the tile are done
            if (view == 0) {
                for (int idx = threadIdx.x; idx < nr * C; idx += kEvalThreads) {  // row-major inputs: coalesced over c
                    const int r = idx / C, c = idx - r * C;
                    tile[c * pitch + r] = eval_value(in, row0 + r, c) - mean[c];
                }
            } else {
                for (int idx = threadIdx.x; idx < nr * C; idx += kEvalThreads) {  // ranks are [C][B]: coalesced over r
                    const int c = idx / nr, r = idx - c * nr;
                    tile[c * pitch + r] = ranks[(int64_t)c * in.B + row0 + r] - rank_mean;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kEvalPairsPerThread; ++q) {
                if (!act[q]) continue;
                const double *xa = tile + ca[q] * pitch, *xb = tile + cb[q] * pitch;
                double a = 0.0;
                for (int r = 0; r < nr; ++r) a = fma(xa[r], xb[r], a);
                acc[view][q] += a;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kEvalPairsPerThread; ++q) {
        const int p = blockIdx.y * kEvalPairsPerBlock + q * kEvalThreads + threadIdx.x;
        if (!act[q]) continue;
        partial[((int64_t)blockIdx.x * 2 + 0) * NP + p] = acc[0][q];
        partial[((int64_t)blockIdx.x * 2 + 1) * NP + p] = acc[1][q];
    }
}

struct EvalOutputs {
    double *rho, *pval, *corr, *sap;  // [Z][A]
    double *scores;                   // { Corr_score, SAP_score }
};

// One CTA: block partials -> totals (fixed order), then the reference's per-pair formulas and both scores.
__global__ void __launch_bounds__(kEvalThreads)
eval_finish_kernel(const double *__restrict__ partial, int n_blocks, int64_t B, int Z, int A,
                   const int *__restrict__ nanflag, double *__restrict__ tot, double *__restrict__ colstat,
                   EvalOutputs out) {
    const int C = Z + A, ZA = Z * A, NP = ZA + C;
    for (int p = threadIdx.x; p < 2 * NP; p += kEvalThreads) {
        const int view = p / NP, q = p - view * NP;
        double s = 0.0;
        for (int b = 0; b < n_blocks; ++b) s += partial[((int64_t)b * 2 + view) * NP + q];
        tot[p] = s;
    }
    __syncthreads();
    const double nm1 = (double)(B - 1), dof = (double)(B - 2);
    for (int p = threadIdx.x; p < ZA; p += kEvalThreads) {
        const int i = p / A, j = p - i * A;
        // SAP matrix (utils/evaluation.py:203-213)
        const double cov = tot[p] / nm1, var_mu = tot[ZA + i] / nm1, var_y = tot[ZA + Z + j] / nm1;
        out.sap[p] = (var_mu > 1e-12) ? (cov * cov) * 1.0 / (var_mu * var_y) : 0.0;
        // Spearman (utils/evaluation.py:166-170): Pearson correlation of the average ranks, Student-t p-value
        const double sxy = tot[NP + p], sxx = tot[NP + ZA + i], syy = tot[NP + ZA + Z + j];
        double rho = NAN, pv = NAN;
        if (B >= 3 && !nanflag[i] && !nanflag[Z + j] && sxx > 0.0 && syy > 0.0) {  // constant input: NaN, as scipy
            rho = sxy / sqrt(sxx * syy);  // exact +-1 for perfectly monotone columns (sxx = syy = |sxy|)
            rho = rho > 1.0 ? 1.0 : (rho < -1.0 ? -1.0 : rho);
            pv = student_t_two_sided(correlation_t(rho, dof), dof);
        }
        out.rho[p] = rho;
        out.pval[p] = pv;
        out.corr[p] = (pv <= 0.05) ? fabs(rho) : 0.0;  // NaN p-value fails the gate
    }
    __syncthreads();
    for (int j = threadIdx.x; j < A; j += kEvalThreads) {
        // Corr_score: max over codes (:151-153).  SAP: np.sort puts NaN last, so top1 - top2 (:217-219) is NaN as
        // soon as the column holds one.
        double cmax = -INFINITY, m1 = -INFINITY, m2 = -INFINITY;
        bool has_nan = false;
        for (int i = 0; i < Z; ++i) {
            cmax = fmax(cmax, out.corr[i * A + j]);
            const double s = out.sap[i * A + j];
            if (s != s) has_nan = true;
            else if (s > m1) { m2 = m1; m1 = s; }
            else if (s > m2) m2 = s;
        }
        colstat[j] = cmax;
        colstat[A + j] = (has_nan || Z < 2) ? NAN : m1 - m2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0.0, ss = 0.0;
        for (int j = 0; j < A; ++j) { cs += colstat[j]; ss += colstat[A + j]; }
        out.scores[0] = cs / (double)A;
        out.scores[1] = ss / (double)A;
    }
}

// ---- host side --------------------------------------------------------------------------------------------
struct EvalLayout {
    int64_t N;          // padded sort size
    int TR;             // rows per tile of the moments kernel
    int64_t n_tiles;
    int n_blocks;
    size_t off_keys, off_ranks, off_mean, off_nan, off_partial, off_tot, off_colstat, bytes;
};

static EvalLayout eval_layout(int64_t B, int Z, int A) {
    EvalLayout L;
    const int C = Z + A, NP = Z * A + C;
    L.N = sort_padded_size(B);
    int tr = kEvalSmemBytes / (8 * C) - 1;
    L.TR = tr > 128 ? 128 : (tr < 1 ? 1 : tr);
    L.n_tiles = ceil_div(B, L.TR);
    L.n_blocks = (int)(L.n_tiles < kEvalMaxBlocks ? L.n_tiles : kEvalMaxBlocks);
    const int widest = Z > A ? Z : A;
    const int sort_cols = widest < ARVAE_MAX_REG_DIMS ? widest : ARVAE_MAX_REG_DIMS;  // columns sorted per pass
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (size_t)round_up((int64_t)n, 256); return at; };
    L.off_keys = take(sizeof(unsigned long long) * (size_t)sort_cols * L.N);
    L.off_ranks = take(sizeof(double) * (size_t)C * B);
    L.off_mean = take(sizeof(double) * C);
    L.off_nan = take(sizeof(int) * C);
    L.off_partial = take(sizeof(double) * 2 * (size_t)NP * L.n_blocks);
    L.off_tot = take(sizeof(double) * 2 * NP);
    L.off_colstat = take(sizeof(double) * 2 * A);
    L.bytes = o;
    return L;
}

size_t eval_metrics_workspace_bytes(int64_t B, int Z, int A) { return eval_layout(B, Z, A).bytes; }

static int rank_columns(const float *src, int64_t rs, int64_t cs, int ncols, int64_t B, const EvalLayout &L,
                        unsigned long long *keys, double *ranks, int *nanflag, cudaStream_t st) {
    for (int c0 = 0; c0 < ncols; c0 += ARVAE_MAX_REG_DIMS) {
        const int nb = ncols - c0 < ARVAE_MAX_REG_DIMS ? ncols - c0 : ARVAE_MAX_REG_DIMS;
        RegDims d;
        memset(&d, 0, sizeof(d));
        for (int r = 0; r < nb; ++r) d.lcol[r] = c0 + r;
        int rc = run_sort_keys(src, rs, cs, d, nb, B, L.N, keys, st);
        if (rc) return rc;
        dim3 grid((unsigned)ceil_div(B, kEvalThreads), (unsigned)nb);
        ARVAE_LAUNCH(avg_ranks_kernel, grid, kEvalThreads, 0, st, keys, L.N, B, ranks + (int64_t)c0 * B, nanflag + c0);
        ARVAE_LAUNCH_CHECK("avg_ranks_kernel");
    }
    return 0;
}

int run_eval_metrics(const float *codes, int64_t crs, int64_t ccs, const float *attrs, int64_t ars, int64_t acs,
                     int64_t B, int Z, int A, double *rho, double *pval, double *corr, double *sap, double *scores,
                     char *ws, cudaStream_t st) {
    const EvalLayout L = eval_layout(B, Z, A);
    const int C = Z + A, NP = Z * A + C;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + L.off_keys);
    double *ranks = reinterpret_cast<double *>(ws + L.off_ranks);
    double *mean = reinterpret_cast<double *>(ws + L.off_mean);
    int *nanflag = reinterpret_cast<int *>(ws + L.off_nan);
    double *partial = reinterpret_cast<double *>(ws + L.off_partial);
    double *tot = reinterpret_cast<double *>(ws + L.off_tot);
    double *colstat = reinterpret_cast<double *>(ws + L.off_colstat);

    ARVAE_CUDA_TRY(cudaMemsetAsync(nanflag, 0, sizeof(int) * C, st));
    int rc = rank_columns(codes, crs, ccs, Z, B, L, keys, ranks, nanflag, st);
    if (rc) return rc;
    rc = rank_columns(attrs, ars, acs, A, B, L, keys, ranks + (int64_t)Z * B, nanflag + Z, st);
    if (rc) return rc;

    EvalInputs in{codes, crs, ccs, attrs, ars, acs, B, Z, A};
    ARVAE_LAUNCH(col_means_kernel, C, kEvalThreads, 0, st, in, mean);
    ARVAE_LAUNCH_CHECK("col_means_kernel");

    dim3 grid((unsigned)L.n_blocks, (unsigned)ceil_div(NP, kEvalPairsPerBlock));
    const size_t smem = sizeof(double) * (size_t)C * (L.TR + 1);
    ARVAE_LAUNCH(eval_moments_kernel, grid, kEvalThreads, smem, st, in, ranks, mean, L.TR, L.n_tiles, partial);
    ARVAE_LAUNCH_CHECK("eval_moments_kernel");

    EvalOutputs out{rho, pval, corr, sap, scores};
    ARVAE_LAUNCH(eval_finish_kernel, 1, kEvalThreads, 0, st, partial, L.n_blocks, B, Z, A, nanflag, tot, colstat, out);
    ARVAE_LAUNCH_CHECK("eval_finish_kernel");
    return 0;
}

}  // namespace arvae
