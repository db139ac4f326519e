// reg_sorted.cu -- attribute-sorted pair kernel (sm_100a): the fast path for large batches.
//
// Same arithmetic as reg_dense.cu (reference utils/trainer.py:390-401 and its autograd backward), reorganised so that
// almost every pair costs 0.44 MUFU + 1.7 packed FP32 instructions instead of 2 MUFU + 12 scalar ones:
//
//  * rows and columns of each regularised dim are ordered by attribute value (sort.cu).  A tile of 128 rows (one warp)
//    x 256 columns whose attribute ranges do not overlap has a CONSTANT sign s_ij, so per pair only sum(r) and sum(r^2)
//    are needed, r = (1 - t)/2:
//        s = +1:  |t - s| = 2r        g = -(1 - t^2) = -4 (r - r^2)
//        s = -1:  |t - s| = 2(1 - r)  g = +4 (r - r^2)
//    Tiles inside one tie group (all attributes equal, incl. NaN rows/columns: NaN ties with everything) have s = 0:
//    |t| = 2 |1/2 - r|, g = sgn(1/2 - r) 4 (r - r^2).  Only tiles that straddle the diagonal band / a tie-group edge run
//    the general loop with per-pair float compares of the raw attributes -- the sign is exact by construction in all
//    three classes.
//  * r = 1/(1 + 2^(u_i-u_j)) = E_j / (E_i + E_j) with u = 2 f log2(e) x and E = 2^u precomputed once per element: one
//    MUFU.RCP per pair, safe while |u| <= 62 (no overflow, full relative accuracy in both saturation directions).
//    Samples beyond that carry an outlier bit in their sort key and form a segment of their own: only tiles touching it
//    take the 2-MUFU form (EX2 + RCP on the scaled latent difference).
//  * constant-sign tiles go further (loop_const_shared): with q = 1 - r = 1 / (1 + E_j F_i), F_i = 2^-u_i, two pairs share
//    ONE reciprocal 1 / (a b) and only the sums q_a + q_b, q_a^2 + q_b^2 are formed, from column-pair sums and products
//    staged in shared memory -- valid while |u| <= 31 for EVERY element of the call, which a device flag decides: the
//    build of the kernel for that case (ONLY1) or the complete one (per-pair reciprocals) does the work.
//  * wherever a pair can be a tie (tie and general tiles), sgn(t) is taken from the exact float difference xs_i - xs_j
//    of the (sign-adjusted) latents, never from the approximated tanh: near t = 0 the factor (1 - t^2) is maximal and
//    abs-backward's sgn(0) = 0 must hold exactly for equal latents (the diagonal, duplicated samples).
//  * work is cut into units (1024-row tile, 256-column sub-chunk), visited in a permuted column order; a planner kernel
//    models each unit's cost and every CTA of a persistent grid (one per SM) gets a cost-balanced contiguous range, which
//    its two 8-warp halves consume from both ends (dynamic meeting point), so there is no wave tail and no idle half;
//    row partials are fixed-point integers added to per-row accumulators with integer atomics: the result does not
//    depend on the split.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

#ifndef ARVAE_TILE_THREADS
#define ARVAE_TILE_THREADS 256
#endif
constexpr int kTileThreads = ARVAE_TILE_THREADS;
constexpr int kTileRI = 4;
constexpr int kTileRows = kTileThreads * kTileRI;  // 1024 rows per row tile (8 warps x 128 rows; 2 CTAs/SM: best measured)
constexpr int kStageCols = 2048;                   // columns staged per __syncthreads pair
static_assert(kTileRows / (kTileThreads / 32) == kTileRI * 32, "a warp owns kTileRI x 32 consecutive rows");
static_assert(kTileThreads / 32 <= 16, "class word holds 16 warps");

// The fixed-point row sums cannot carry NaN, so non-finite latents are flagged when the columns are built and
// the outputs are patched to what the reference's float arithmetic gives: the loss is NaN, a row with a non-finite
// latent has a NaN gradient (inf - inf on its diagonal), and a NaN latent poisons every row of its dim.
__device__ __forceinline__ void note_nonfinite(float xs, int *__restrict__ flags, int r) {
    if (!(fabsf(xs) < __int_as_float(0x7f800000))) atomicOr(flags + kFlagNonFinite + r, xs != xs ? 2 : 1);
}
__device__ __forceinline__ bool row_is_poisoned(const int *__restrict__ flags, int r, float xs) {
    const int nf = flags[kFlagNonFinite + r];
    return nf != 0 && ((nf & 2) || !(fabsf(xs) < __int_as_float(0x7f800000)));
}
__device__ __forceinline__ bool any_nonfinite(const int *__restrict__ flags, int R) {
    int nf = flags[kFlagError];
    for (int r = 0; r < R; ++r) nf |= flags[kFlagNonFinite + r];
    return nf != 0;
}


// ------------------------------------------------------------------------------------------------
// gather the sorted order: Us/As/Es[r][k] for sorted position k, perm[r][k] = original index
// ------------------------------------------------------------------------------------------------
// Sorted order of a dim (sort.cu, segmented keys): [inliers by attribute | outliers by attribute | padding].
// n_in[r] = number of inliers = the first position of the outlier segment.  Exactly one thread per dim sees the
// segment boundary and writes it (no atomics, no memset).
__device__ __forceinline__ void note_segment_boundary(const unsigned long long *__restrict__ kr, int64_t k, int64_t B,
                                                      int *__restrict__ n_in_r) {
    if (k >= B) return;
    const bool out_here = key_is_outlier(kr[k]);
    if (k == 0 && out_here) *n_in_r = 0;
    if (!out_here && (k + 1 == B || key_is_outlier(kr[k + 1]))) *n_in_r = (int)(k + 1);
}

__global__ void __launch_bounds__(256)
sorted_gather_kernel(const unsigned long long *__restrict__ keys, int64_t N,
                     const float *__restrict__ z, int64_t zrs, int64_t zcs,
                     const float *__restrict__ lab, int64_t lrs, int64_t lcs, RegDims dims,
                     int64_t B, int64_t Bpad, float fsign, float cabs, float *__restrict__ Xs,
                     float *__restrict__ As, float *__restrict__ Es, int *__restrict__ perm,
                     int *__restrict__ flags, int64_t row_begin, int64_t row_end,
                     int *__restrict__ blockcnt, int unsegmented) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const unsigned long long *kr = keys + (int64_t)r * N;
    int mine = 0;
    if (k < B) {
        const int64_t idx = (int64_t)(kr[k] & kKeyIdxMask);
        mine = idx >= row_begin && idx < row_end;
    }
    const int cnt = __syncthreads_count(mine);  // rows of this call among this CTA's 256 sorted positions
    if (blockcnt && threadIdx.x == 0) blockcnt[(int64_t)r * gridDim.x + blockIdx.x] = cnt;
    if (k >= Bpad) return;
    const int64_t o = (int64_t)r * Bpad + k;
    if (k < B) {
        const unsigned long long key = kr[k];
        const int64_t idx = (int64_t)(key & kKeyIdxMask);
        const float xs = signed_latent(__ldg(z + idx * zrs + (int64_t)dims.zcol[r] * zcs), fsign);
        Xs[o] = xs;
        As[o] = __ldg(lab + idx * lrs + (int64_t)dims.lcol[r] * lcs);
        const float u = cabs * xs;
        Es[o] = key_is_outlier(key) ? 1.0f : exp2f(u);  // outliers never take the factorised form
        perm[o] = (int)idx;
        // unsegmented keys (triangle mode): one out-of-range element sends the whole dim to the two-MUFU form
        if (unsegmented && !(fabsf(u) <= kMufu1MaxAbsU)) atomicOr(flags + r, 1);
        if (!(fabsf(u) <= kSharedMaxAbsU)) atomicOr(flags + kFlagNeedComplete, 1);
        note_nonfinite(xs, flags, r);
        note_segment_boundary(kr, k, B, flags + kFlagNIn + r);
    } else {  // padding: |t - s| = 1 and zero gradient for every row, in either tanh form
        Xs[o] = ARVAE_PAD_U;
        As[o] = ARVAE_PAD_A;
        Es[o] = 8.5070592e37f;  // 2^126: E_i + E_j stays finite, r = E_j / (E_i + E_j) = 1 exactly
        perm[o] = -1;
    }
}

// rowpos[r][m] = m-th sorted position whose original index lies in [row_begin,row_end)
// (row-block sharding: this rank's rows, in attribute order).  Order-preserving compaction: the gather
// kernel left per-256-position counts; CTA b of dim r sums the counts before its 1024 positions and
// ranks its own positions with ballots.
__global__ void __launch_bounds__(1024)
row_select_kernel(const int *__restrict__ perm, const int *__restrict__ blockcnt, int n_cnt, int64_t B,
                  int64_t Bpad, int64_t row_begin, int64_t row_end, int64_t n_rows,
                  int *__restrict__ rowpos) {
    __shared__ int swarp[32];
    __shared__ int sbase;
    const int r = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int *cnt = blockcnt + (int64_t)r * n_cnt;
    int part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x * 4 && i < n_cnt; i += 1024) part += cnt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) swarp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += swarp[w];
        sbase = t;
    }
    __syncthreads();
    const int base = sbase;
    const int64_t k = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int v = k < B ? perm[(int64_t)r * Bpad + k] : -1;
    const bool f = v >= row_begin && v < row_end;
    const unsigned int b = __ballot_sync(0xffffffffu, f);
    __syncthreads();
    if (lane == 0) swarp[warp] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += swarp[w];
    if (f) rowpos[(int64_t)r * n_rows + base + before + __popc(b & ((1u << lane) - 1u))] = (int)k;
}

// ------------------------------------------------------------------------------------------------
// pair loops over one 256-column sub-chunk staged in shared memory
// ------------------------------------------------------------------------------------------------
// Row partial sums are carried in 2^-30 fixed point (int64): integer addition is associative, so a row's
// total does not depend on how its column units were split between CTAs / CTA halves at run time (the split
// is dynamic), and results stay bitwise reproducible.  |v| < 2^21 per call (a sub-chunk partial is <= 512).
typedef long long acc_t;
constexpr double kFixMagic = 6291456.0;              // 1.5 * 2^22: (v + magic) keeps round(v 2^30) in the mantissa
constexpr double kFixScale = 1.0 / 1073741824.0;     // 2^-30
// Loss partials are carried as two exact integers (v >> 20 and v & (2^20 - 1)) so that neither can overflow and their
// totals -- hence the loss -- do not depend on how the pair work was split over CTAs, halves or GPUs.
constexpr int kLossSplitBits = 20;
constexpr long long kLossLoMask = (1LL << kLossSplitBits) - 1;
constexpr double kLossHiScale = 1.0 / 1024.0;        // 2^20 * 2^-30
__device__ __forceinline__ void acc_add(acc_t &acc, float v) {
    acc += __double_as_longlong((double)v + kFixMagic) - __double_as_longlong(kFixMagic);
}

// Per-thread row operands of one row tile.
struct RowRegs {
    float e[kTileRI];  // 2^u_i          (1-MUFU form)
    float f[kTileRI];  // 2^-u_i         (one-MUFU constant-sign loops: q = 1 / (1 + E_j F_i))
    float f2[kTileRI]; // 2^-2u_i        (shared-reciprocal loop on staged column-pair sums and products)
    float x[kTileRI];  // sgn(f) x_i     (2-MUFU form, exact tie signs)
    float a[kTileRI];  // attribute
};

// Reciprocal on the FMA pipe for a fixed share of the pairs of the constant-sign loops, so that the XU pipe (MUFU.RCP,
// 16 lanes/clk/SM) and the FP32 pipe are both kept busy: magic-constant seed (relative error < 0.051), one quadratic
// Newton step (-> 2.6e-3) and one cubic step y (1 + e + e^2) (-> 1.8e-8; measured over 3e6 arguments in [1, 2^124]:
// 1.28 * 2^-24, within an ulp like MUFU.RCP itself; tests/test_pair_arithmetic.py) -- five FFMA2 and two IADD per TWO
// reciprocals (three quadratic steps, six FFMA2: 4.73 instead of 4.53 ms at C4 in the per-pair loop).  Valid for normal
// positive x below 2^126.
__device__ __forceinline__ f2_t rcp_newton2(f2_t x) {
    float x0, x1;
    unpack2(x, x0, x1);
    f2_t y = pack2(__int_as_float(0x7EF311C7 - __float_as_int(x0)), __int_as_float(0x7EF311C7 - __float_as_int(x1)));
    const f2_t one = pack2(1.0f, 1.0f), nx = pack2(-x0, -x1);
    f2_t e = fma2(nx, y, one);
    y = fma2(y, e, y);
    e = fma2(nx, y, one);
    const f2_t t = fma2(e, e, e);
    return fma2(y, t, y);
}
// Two reciprocals: on the XU pipe (two MUFU.RCP) or, where a compile-time mask says so, on the FMA pipe.
template <bool NEWTON>
__device__ __forceinline__ f2_t rcp2(f2_t x) {
    if (NEWTON) return rcp_newton2(x);
    float x0, x1;
    unpack2(x, x0, x1);
    return pack2(rcp_approx(x0), rcp_approx(x1));
}

// Compile-time choices of the one-MUFU constant-sign loops.  ptxas's schedule decides as much as the instruction counts,
// so each was an A/B sweep on the GPU (bench_tools/pair_variants.sh, profiles/r2c_pair_ab_variants.jsonl), pair kernel at
// C4 in ms:
//  * ARVAE_NR_MASK -- per-pair loop: which of the 16 (column group g, row k, column pair h) slots (bit 8 g + 2 k + h) of a
//    4 x 8 pair group take the Newton reciprocal: 0xE0E0 4.53 (kept), 0xD0D0 4.54, 0xA0E0 4.55, 0x00FC 4.60, 0xA8A8 4.62,
//    0xE00E 4.65, 0x3838 4.67, 0x0E0E 4.68, 0x5454 4.71, 0x8383 4.74, 0x7070 4.77, 0xE0F0 4.80, 0xC0E0 5.00, 0xE0C0 5.07
//    (the round's earlier form r = E_j / (E_i + E_j) with one more FMUL2 per two pairs: 5.15 at its best mask)
//  * ARVAE_NR_MASK_TP -- shared-reciprocal loop: which of the 8 (column group g, row k) quads (bit 4 g + k): none 3.51;
//    one quad 0x10 3.214 (kept), 0x20 3.218, 0x02 3.225, 0x01 3.229, 0x40 3.231, 0x04 3.252, 0x80 3.53, 0x08 3.54;
//    two quads 3.31-3.39.  (Both quotients of a quad instead of their sums: 4.09; the sums without staged T, P: 3.57)
//  * ARVAE_CONST_OUTER_UNROLL -- pair groups per branch: 1 3.90, 2 3.73, 4 3.62, 8 3.57, 16 3.61, 32 3.87 (sums-only form)
#ifndef ARVAE_NR_MASK
#define ARVAE_NR_MASK 0xE0E0
#endif
#ifndef ARVAE_NR_MASK_TP
#define ARVAE_NR_MASK_TP 0x10
#endif
#ifndef ARVAE_CONST_OUTER_UNROLL
#define ARVAE_CONST_OUTER_UNROLL 8
#endif
constexpr int kConstOuterUnroll = ARVAE_CONST_OUTER_UNROLL;

template <bool MUFU1>
__device__ __forceinline__ float pair_r(float ei, float ej, float d, float cabs) {
    if (MUFU1) return rcp_approx(ei + ej) * ej;          // E_j / (E_i + E_j)
    return rcp_approx(ex2_approx(d * cabs) + 1.0f);       // 1 / (1 + 2^(|c| (xs_i - xs_j)))
}

// Constant-sign tile: only S1 = sum_j q and S2 = sum_j q^2 per row are needed, q = 1 - r = 1 / (1 + E_j F_i) with
// F_i = 2^-u_i held per row (r - r^2 = q - q^2, so against sums of r only the two loss expressions swap):
//   s = +1:  sum |t - s| = 2 (n - S1),  sum g / 4 = S2 - S1;      s = -1:  2 S1,  S1 - S2.
// Both one-MUFU loops work on column PAIRS with the packed FP32 instructions.
// SHARED: A2[k][1] holds sum 1 / (a b) of the shared-reciprocal loop, S2 = sum w^2 - 2 sum 1 / (a b).
template <bool GRAD, bool SHARED>
__device__ __forceinline__ void const_tile_epilogue(const f2_t (&A1)[kTileRI][2], const f2_t (&A2)[kTileRI][2], bool positive,
                                                    acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) {
        float a0, a1, a2, a3, b0, b1, b2, b3;
        unpack2(A1[k][0], a0, a1);
        unpack2(A1[k][1], a2, a3);
        unpack2(A2[k][0], b0, b1);
        unpack2(A2[k][1], b2, b3);
        const float S1 = (a0 + a1) + (a2 + a3);
        const float S2 = SHARED ? fmaf(-2.0f, b2 + b3, b0 + b1) : (b0 + b1) + (b2 + b3);
        acc_add(dl[k], positive ? 2.0f * ((float)kSubCols - S1) : 2.0f * S1);
        if (GRAD) acc_add(dg[k], positive ? S2 - S1 : S1 - S2);
    }
}

// Shared-reciprocal loop (the build for the common case: every |u| <= kSharedMaxAbsU, so a b <= (1 + 2^62)^2 is finite).
// A quad = one row x the column pairs (j, j+2) and (j+1, j+3) of a group of four columns shares ONE reciprocal per packed
// lane, and only the quad's sums are formed:
//   a = 1 + E_a F_i, b = 1 + E_b F_i;  a + b - 1 = 1 + T F_i,  a b = (a + b - 1) + P F_i^2   with the STAGED column-pair
//   sums T = E_a + E_b and products P = E_a E_b;  w = q_a + q_b = (a + b) / (a b),  q_a^2 + q_b^2 = w^2 - 2 / (a b).
// Per quad: three FFMA2 (a + b - 1, a b, w), two MUFU.RCP, FADD2 (sum w), FFMA2 (sum w^2), FADD2 (sum 1/(a b)); one quad
// of eight takes the Newton reciprocal.  SASS: 580 instructions per 256 pairs (112 MUFU.RCP, 296 FFMA2, 128 FADD2,
// 17 IADD3, 16 LDS.128): XU pipe 112 and FP32 pipe 106 clk per 32 pairs and warp.
template <bool GRAD>
__device__ __forceinline__ void loop_const_shared(const RowRegs &R, const float *__restrict__ stp, bool positive,
                                                  acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
    f2_t A1[kTileRI][2], A2[kTileRI][2];  // [k][0]: sum w, sum w^2;  A2[k][1]: sum 1 / (a b);  A1[k][1] stays 0
#pragma unroll
    for (int k = 0; k < kTileRI; ++k)
#pragma unroll
        for (int h = 0; h < 2; ++h) A1[k][h] = A2[k][h] = pack2(0.0f, 0.0f);
    const f2_t one = pack2(1.0f, 1.0f);
#pragma unroll kConstOuterUnroll
    for (int q = 0; q < kSubCols; q += 8) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float4 tp = *reinterpret_cast<const float4 *>(stp + q + 4 * g);
            const f2_t T = pack2(tp.x, tp.y), P = pack2(tp.z, tp.w);
#pragma unroll
            for (int k = 0; k < kTileRI; ++k) {
                const f2_t fk = pack2(R.f[k], R.f[k]), fk2 = pack2(R.f2[k], R.f2[k]);
                const f2_t tm1 = fma2(T, fk, one);
                const f2_t p = fma2(P, fk2, tm1);
                const f2_t rp = ((ARVAE_NR_MASK_TP >> (g * 4 + k)) & 1) ? rcp2<true>(p) : rcp2<false>(p);
                const f2_t w = fma2(rp, tm1, rp);
                A1[k][0] = add2(A1[k][0], w);
                if (GRAD) {
                    A2[k][0] = fma2(w, w, A2[k][0]);
                    A2[k][1] = add2(A2[k][1], rp);
                }
            }
        }
    }
    const_tile_epilogue<GRAD, true>(A1, A2, positive, dl, dg);
}

// Per-pair loops.  One-MUFU form (complete build: valid for |u| <= 62, i.e. 1 + E_j F_i <= 1 + 2^124): per two pairs
// FFMA2 (1 + E_j F_i), two MUFU.RCP (on 6 of 16 slots the Newton reciprocal), FADD2 (sum q), FFMA2 (sum q^2) -- SASS: 914
// instructions per 256 pairs (160 MUFU.RCP, 496 FFMA2, 128 FADD2, 97 IADD3, 16 LDS.128).  Two-MUFU form (tiles touching the
// outlier segment): EX2 + RCP on the scaled latent difference.
template <bool MUFU1, bool GRAD>
__device__ __forceinline__ void loop_const(const RowRegs &R, const float *__restrict__ se,
                                           const float *__restrict__ sx, float cabs, bool positive,
                                           acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
    if (MUFU1) {
        f2_t A1[kTileRI][2], A2[kTileRI][2];
#pragma unroll
        for (int k = 0; k < kTileRI; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h) A1[k][h] = A2[k][h] = pack2(0.0f, 0.0f);
        const f2_t one = pack2(1.0f, 1.0f);
#pragma unroll kConstOuterUnroll
        for (int q = 0; q < kSubCols; q += 8) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float4 vj = *reinterpret_cast<const float4 *>(se + q + 4 * g);
                const f2_t vv[2] = {pack2(vj.x, vj.y), pack2(vj.z, vj.w)};
#pragma unroll
                for (int k = 0; k < kTileRI; ++k) {
                    const f2_t fk = pack2(R.f[k], R.f[k]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const f2_t sum = fma2(vv[h], fk, one);  // 1 + E_j F_i = (E_i + E_j) / E_i
                        const f2_t qq = ((ARVAE_NR_MASK >> (g * 8 + k * 2 + h)) & 1) ? rcp2<true>(sum) : rcp2<false>(sum);
                        A1[k][h] = add2(A1[k][h], qq);
                        if (GRAD) A2[k][h] = fma2(qq, qq, A2[k][h]);
                    }
                }
            }
        }
        const_tile_epilogue<GRAD, false>(A1, A2, positive, dl, dg);
        return;
    }
    float A1[kTileRI][4], A2[kTileRI][4];
#pragma unroll
    for (int k = 0; k < kTileRI; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) A1[k][q] = A2[k][q] = 0.0f;
#pragma unroll 2
    for (int q = 0; q < kSubCols; q += 4) {
        const float4 vj = *reinterpret_cast<const float4 *>(sx + q);
        const float vv[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float r = pair_r<false>(0.0f, 0.0f, R.x[k] - vv[e], cabs);
                A1[k][e] += r;
                if (GRAD) A2[k][e] = fmaf(r, r, A2[k][e]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) {
        const float S1 = (A1[k][0] + A1[k][1]) + (A1[k][2] + A1[k][3]);
        const float S2 = (A2[k][0] + A2[k][1]) + (A2[k][2] + A2[k][3]);
        acc_add(dl[k], positive ? 2.0f * S1 : 2.0f * ((float)kSubCols - S1));
        if (GRAD) acc_add(dg[k], positive ? S2 - S1 : S1 - S2);
    }
}

template <bool MUFU1, bool GRAD, int ncols = kSubCols>
__device__ __forceinline__ void loop_tie(const RowRegs &R, const float *__restrict__ se,
                                         const float *__restrict__ sx, float cabs,
                                         acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
    float lacc[kTileRI], gacc[kTileRI];
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) lacc[k] = gacc[k] = 0.0f;
#pragma unroll 2
    for (int q = 0; q < ncols; q += 4) {
        const float4 xj = *reinterpret_cast<const float4 *>(sx + q);
        const float xx[4] = {xj.x, xj.y, xj.z, xj.w};
        float ee[4] = {0.f, 0.f, 0.f, 0.f};
        if (MUFU1) {
            const float4 ej = *reinterpret_cast<const float4 *>(se + q);
            ee[0] = ej.x; ee[1] = ej.y; ee[2] = ej.z; ee[3] = ej.w;
        }
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = R.x[k] - xx[e];
                const float r = pair_r<MUFU1>(R.e[k], ee[e], d, cabs);
                const float h = 0.5f - r;  // t / 2
                lacc[k] += fabsf(h);
                if (GRAD) {
                    const float w4 = fmaf(-r, r, r);
                    const float sg = fminf(fmaxf(d * 8.5070592e37f, -1.0f), 1.0f);  // sgn(t) = sgn(d), exact
                    gacc[k] = fmaf(sg, w4, gacc[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) {
        acc_add(dl[k], 2.0f * lacc[k]);
        if (GRAD) acc_add(dg[k], gacc[k]);
    }
}

template <bool MUFU1, bool GRAD, int ncols = kSubCols, bool SIGNS = false>
__device__ __forceinline__ void loop_general(const RowRegs &R, const float *__restrict__ se,
                                             const float *__restrict__ sx,
                                             const float *__restrict__ sa, float cabs,
                                             acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI], int *ds = nullptr) {
    float lacc[kTileRI], gacc[kTileRI], kacc[kTileRI];
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) lacc[k] = gacc[k] = kacc[k] = 0.0f;
#pragma unroll 2
    for (int q = 0; q < ncols; q += 4) {
        const float4 xj = *reinterpret_cast<const float4 *>(sx + q);
        const float4 aj = *reinterpret_cast<const float4 *>(sa + q);
        const float xx[4] = {xj.x, xj.y, xj.z, xj.w};
        const float aa[4] = {aj.x, aj.y, aj.z, aj.w};
        float ee[4] = {0.f, 0.f, 0.f, 0.f};
        if (MUFU1) {
            const float4 ej = *reinterpret_cast<const float4 *>(se + q);
            ee[0] = ej.x; ee[1] = ej.y; ee[2] = ej.z; ee[3] = ej.w;
        }
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = R.x[k] - xx[e];
                const float r = pair_r<MUFU1>(R.e[k], ee[e], d, cabs);
                const float gt = R.a[k] > aa[e] ? 1.0f : 0.0f;
                const float lt = R.a[k] < aa[e] ? 1.0f : 0.0f;
                const float kk = (1.0f - gt) + lt;  // 1 - s
                const float v = fmaf(-2.0f, r, kk);
                lacc[k] += fabsf(v);
                if (SIGNS) kacc[k] += kk;  // small integers: exact
                if (GRAD) {
                    const float w4 = fmaf(-r, r, r);
                    // sgn(v): -s when s != 0, else sgn(d) (see reg_dense.cu: pair_general)
                    const float q2 = fmaf(kk - 1.0f, 1.7014118e38f, d * 1.1529215e18f);
                    const float sg = fminf(fmaxf(q2, -1.0f), 1.0f);
                    gacc[k] = fmaf(sg, w4, gacc[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) {
        acc_add(dl[k], lacc[k]);
        if (GRAD) acc_add(dg[k], gacc[k]);
        if (SIGNS) ds[k] += ncols - (int)kacc[k];  // sum_j s_ij = n - sum_j (1 - s_ij), from the very compares the loss used
    }
}

enum TileClass { kClassGeneral = 0, kClassPos = 1, kClassNeg = 2, kClassTie = 3 };

// Class of (row tile with attribute range [amin,amax], column sub-chunk [cmin,cmax]); both sorted
// ascending with NaN (and padding) last, so the end points are the true min / max.
__device__ __forceinline__ int classify(float amin, float amax, float cmin, float cmax) {
    const bool row_all_nan = amin != amin, row_has_nan = amax != amax;
    const bool col_all_nan = cmin != cmin, col_has_nan = cmax != cmax;
    if (row_all_nan || col_all_nan) return kClassTie;     // NaN compares false both ways: s = 0
    if (row_has_nan || col_has_nan) return kClassGeneral;
    if (cmax < amin) return kClassPos;                     // a_i >= amin > cmax >= a_j
    if (cmin > amax) return kClassNeg;
    if (amin == amax && cmin == cmax && amin == cmin) return kClassTie;
    return kClassGeneral;
}

struct TilesArgs {
    const float *Xs, *Es, *As;  // [R][Bpad] in sorted order: sgn(f) x, 2^u, attribute
    float cabs;                 // |2 f log2(e)|
    const int *rowpos;          // [R][n_rows] sorted positions of this call's rows, or null = identity
    const int *flags;           // small per-call flags (layout above): [r] non-zero = the WHOLE dim uses the 2-MUFU form
    const int *n_in;            // [R] number of inliers = first sorted position of the outlier segment
    int64_t Bpad, n_rows;
    int n_row_tiles, S;         // S = Bpad / kSubCols sub-chunks per row tile
    int P;                      // sub-chunk visiting stride (coprime to S)
    int64_t F;                  // fine units = R * n_row_tiles * S
    int G;                      // persistent CTAs the plan is cut into (all GPUs of a sharded step together)
    int c_first;                // first of those CTAs this launch runs (0 on a single GPU)
    int64_t n_rr;               // R * n_row_tiles
    int force_general;          // treat every tile as general (unsorted input / debugging)
    int dual;                   // the launch is a pair of kernels: the common-case build (ONLY1) works when every element
                                // of every dim is within the shared-reciprocal range (flags[kFlagNeedComplete] == 0), the
                                // complete one otherwise
    // plan (plan_classes_kernel / plan_scan_kernel): per fine unit in VISITING order u = rr * S + s'
    unsigned int *cls8;         // [F] per warp w: 2-bit tile class (bits 2w..2w+1) and tanh form (bit 16+w: 1 = two MUFU)
    unsigned short *cost8;      // [F] modelled cost of the unit (sum over the tile's warps)
    long long *prefix;          // [n_rr + 1] exclusive prefix of the per-row-tile cost totals; [n_rr] = T
    // fixed-point row accumulators, indexed [rr * kTileRows + row within tile]; integer atomics: order-free
    acc_t *acc_g, *acc_l;       // gradient / loss row sums (acc_l only when per-row losses are wanted)
    int *acc_s;                 // sum_j sign(a_i - a_j) per row (parity instrumentation), or null
    acc_t *lossp;               // [launch CTAs][2] per-CTA loss partials (hi, lo)
    unsigned long long *dbg_times;  // [G][2] globaltimer at CTA start / end (experiments), or null
    // triangle mode (reg_tri.cuh)
    float2 *colpart;            // per (row tile, column at or above it): column sums (sum r, sum r^2) of double-duty tiles
    int Pinv;                   // inverse of P modulo S
    int64_t B;                  // number of real columns (= rows in triangle mode)
    ShardView shard;            // sharded step (shard.G > 0): the last CTA publishes this GPU's loss partial and signals
                                // its peers (reg_shard.cuh)
};

// Work split: unit u starts at cost position p(u) (exclusive prefix of the modelled costs in visiting
// order); CTA c of G owns the units with floor(p G / T) == c, i.e. p in [ceil(cT/G), ceil((c+1)T/G)).
__device__ __forceinline__ long long owner_of_pos(long long p, long long T, long long G) {
    return (p * G) / T;
}
__device__ __forceinline__ long long ceil_share(long long c, long long T, long long G) {
    return (c * T + G - 1) / G;
}

// Modelled cost (relative time) of one warp's 128 x 256 tile by class and tanh form, in units where the one-MUFU
// constant-sign tile of the build that will run is 8.  The build is known when the plan is made (flag kFlagNeedComplete):
// the common-case build (ONLY1: shared-reciprocal constant-sign loop, 3.57 ms at C4) or the complete one (plain loop,
// 4.53 ms).  The warps of a half wait for each other at every grant, so a slow tile costs its half more than its share
// of the instructions: the constants are fitted, not counted -- pair kernel at B = 65 536 on dSprites-shaped labels
// (up to 33 % ties), ms, by (general, tie): (21, 12) 4.82, (30, 17) 4.32, (38, 21) 4.04, (42, 23) 3.78, (50, 27) 3.76,
// (50, 31) 3.87, (50, 36) 4.02; C4 (2 % general units, no ties) does not react (3.57 throughout).
#ifndef ARVAE_COST_GENERAL1
#define ARVAE_COST_GENERAL1 46
#endif
#ifndef ARVAE_COST_TIE1
#define ARVAE_COST_TIE1 25
#endif
__device__ __forceinline__ int class_cost(int cls, bool mufu1, bool shared_build) {
    // general, pos, neg, tie
    constexpr unsigned int ks = (unsigned int)ARVAE_COST_GENERAL1 | (8u << 8) | (8u << 16) | ((unsigned int)ARVAE_COST_TIE1 << 24);
    // complete build: the same loops against a constant-sign loop that takes 4.53 / 3.57 as long; two-MUFU forms
    // (EX2 + RCP per pair: XU-bound at 512 clk per 32 pairs against ~209): general 26, constant-sign 22, tie 23
    constexpr unsigned int k1 = (unsigned int)(ARVAE_COST_GENERAL1 * 357 / 453) | (8u << 8) | (8u << 16) |
                                ((unsigned int)(ARVAE_COST_TIE1 * 357 / 453) << 24);
    if (shared_build) return (ks >> (8 * cls)) & 0xFF;
    return mufu1 ? ((k1 >> (8 * cls)) & 0xFF) : ((0x1716'161Au >> (8 * cls)) & 0xFF);
}

// One CTA per row tile rr: class byte and cost of each of its S units (in visiting order), and the
// row tile's total cost.
//
// Tanh form per (warp, unit): the factorised one-MUFU form needs every element of the tile to be an inlier
// (|u| <= 62): the warp's rows must all lie before n_in[r] and so must the sub-chunk's real columns.  A warp's row
// group or a sub-chunk that straddles the inlier / outlier boundary is not attribute-sorted across it, so its end
// points say nothing about its range: such tiles run the general loop.
// The kernel also clears the row accumulators of its row tile (clear_acc bits: 1 gradient, 2 loss, 4 signs), and the
// CTA that finishes last turns the per-row-tile totals into the cost prefix the pair kernel cuts its ranges from
// (one launch instead of memset + plan + scan).
__device__ void plan_scan_tail(const int *__restrict__ combo_cost, int64_t n_rr, long long *__restrict__ prefix);

__global__ void __launch_bounds__(256)
plan_classes_kernel(TilesArgs a, int *__restrict__ combo_cost, int clear_acc, unsigned int *__restrict__ ticket) {
    __shared__ float wmin[kTileThreads / 32], wmax[kTileThreads / 32];
    __shared__ int whas[kTileThreads / 32], win[kTileThreads / 32], wmixed[kTileThreads / 32];
    __shared__ int sred[8];
    pdl_wait();     // (sharded step: launched ahead of the apply kernel's end) the sorted columns must be complete
    pdl_trigger();  // the pair kernel's CTAs may be placed as SMs free up; they wait for this grid to complete
    const int64_t rr = blockIdx.x;
    const int r = (int)(rr / a.n_row_tiles), I = (int)(rr % a.n_row_tiles);
    const float *Ar = a.As + (int64_t)r * a.Bpad;
    const int *rp = a.rowpos ? a.rowpos + (int64_t)r * a.n_rows : nullptr;
    const bool dim_mufu1 = a.flags[r] == 0;
    const bool shared_build = a.dual && a.flags[kFlagNeedComplete] == 0;  // which build of the pair kernel will work
    const int64_t nin = a.n_in[r];
    constexpr int kWarpRows = kTileRows / (kTileThreads / 32);
    if (threadIdx.x < kTileThreads / 32) {
        const int64_t m0 = (int64_t)I * kTileRows + (int64_t)threadIdx.x * kWarpRows;
        const int64_t ml = min(m0 + kWarpRows, a.n_rows) - 1;
        const bool has = m0 < a.n_rows;
        const int64_t p0 = has ? (rp ? (int64_t)rp[m0] : m0) : 0, pl = has ? (rp ? (int64_t)rp[ml] : ml) : 0;
        whas[threadIdx.x] = has;
        wmin[threadIdx.x] = has ? Ar[p0] : 0.0f;
        wmax[threadIdx.x] = has ? Ar[pl] : 0.0f;
        win[threadIdx.x] = pl < nin;                 // every row of the warp is an inlier
        wmixed[threadIdx.x] = p0 < nin && pl >= nin;  // rows from both segments
    }
    __syncthreads();
    int total = 0;
    for (int sp = threadIdx.x; sp < a.S; sp += 256) {
        const int64_t col = (((int64_t)sp * a.P) % a.S) * kSubCols;
        const float cmin = Ar[col], cmax = Ar[col + kSubCols - 1];
        const int64_t col_last_real = min(col + kSubCols, a.B) - 1;  // < col: pure padding (works in either form)
        const bool col_in = col_last_real < nin || col_last_real < col;
        const bool col_mixed = col < nin && col_last_real >= nin;
        unsigned int word = 0;
        int cost = 0;
#pragma unroll
        for (int w = 0; w < kTileThreads / 32; ++w) {
            const bool general = a.force_general || wmixed[w] || col_mixed;
            const int cls = general ? (int)kClassGeneral : classify(wmin[w], wmax[w], cmin, cmax);
            const bool mufu1 = dim_mufu1 && win[w] && col_in;
            word |= (unsigned int)cls << (2 * w);
            word |= (mufu1 ? 0u : 1u) << (16 + w);
            if (whas[w]) cost += class_cost(cls, mufu1, shared_build);
        }
        a.cls8[rr * a.S + sp] = word;
        a.cost8[rr * a.S + sp] = (unsigned short)cost;
        total += cost;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sred[w];
        combo_cost[rr] = t;
    }
    for (int i = threadIdx.x; i < kTileRows; i += 256) {
        if (clear_acc & 1) a.acc_g[rr * kTileRows + i] = 0;
        if (clear_acc & 2) a.acc_l[rr * kTileRows + i] = 0;
        if (clear_acc & 4) a.acc_s[rr * kTileRows + i] = 0;
    }
    // last CTA: exclusive prefix of the row-tile costs
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0;
    plan_scan_tail(combo_cost, a.n_rr, a.prefix);
}

// prefix[rr] = sum of combo_cost[0..rr) ; prefix[n_rr] = T, by one CTA of 256 threads (the last plan CTA).
__device__ void plan_scan_tail(const int *__restrict__ combo_cost, int64_t n_rr, long long *__restrict__ prefix) {
    __shared__ long long swarp[8];
    __shared__ long long scarry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) scarry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_rr; base += 256) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n_rr ? (long long)__ldcg(combo_cost + i) : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        long long before = scarry;
        for (int w = 0; w < warp; ++w) before += swarp[w];
        const long long excl = before + incl - v;
        if (i < n_rr) prefix[i] = excl;
        __syncthreads();
        if (threadIdx.x == 255) scarry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n_rr] = scarry;
}

// prefix[rr] = sum of combo_cost[0..rr) ; prefix[n_rr] = T.   One CTA.  (triangle variant's plan)
__global__ void __launch_bounds__(1024)
plan_scan_kernel(const int *__restrict__ combo_cost, int64_t n_rr, long long *__restrict__ prefix) {
    __shared__ long long swarp[32];
    __shared__ long long scarry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) scarry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_rr; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n_rr ? (long long)combo_cost[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = swarp[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            swarp[lane] = wi - w;
        }
        __syncthreads();
        const long long excl = scarry + swarp[warp] + incl - v;
        if (i < n_rr) prefix[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) scarry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n_rr] = scarry;
}

// First unit (rr, s') whose cost position is >= target; (n_rr, 0) when target >= T.  Cooperative over the
// CTA (kTileThreads threads): one parallel pass over the row-tile prefix, one block scan over the row tile's
// unit costs.
template <int NT>
__device__ __forceinline__ void find_unit(const TilesArgs &a, long long target, int *s_out /*smem [2]*/,
                                          int *s_scan /*smem [NT]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // lo = (number of rr in [0, n_rr] with prefix[rr] <= target) - 1   (prefix is non-decreasing, prefix[0] = 0)
    int cnt = 0;
    for (long long i = threadIdx.x; i <= a.n_rr; i += NT) cnt += a.prefix[i] <= target;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_scan[warp] = cnt;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < NT / 32; ++w) tot += s_scan[w];
    const long long lo = (long long)tot - 1;
    __syncthreads();
    if (lo >= a.n_rr) {
        if (threadIdx.x == 0) { s_out[0] = (int)a.n_rr; s_out[1] = 0; }
        __syncthreads();
        return;
    }
    const long long need = target - a.prefix[lo];  // smallest s' with W(s') >= need, W = exclusive within-tile prefix
    const unsigned short *cost = a.cost8 + lo * a.S;
    const int per = (a.S + NT - 1) / NT;
    const int q0 = min((int)threadIdx.x * per, a.S), q1 = min(q0 + per, a.S);
    int mine = 0;
    for (int q = q0; q < q1; ++q) mine += cost[q];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    if (threadIdx.x == 0) { s_out[0] = (int)lo + 1; s_out[1] = 0; }  // default: past the end of this row tile
    __syncthreads();
    int before = incl - mine;
    for (int w = 0; w < warp; ++w) before += s_scan[w];
    // the crossing lies in exactly one thread's piece: W(q0) < need <= W(q1), or need <= 0 at q0 = 0
    if (need <= before + mine && (need > before || threadIdx.x == 0) && q0 < q1) {
        int w = before, q = q0;
        while (q < q1 && w < need) w += cost[q++];
        if (w >= need) {
            s_out[0] = (int)lo;
            s_out[1] = q;
            if (q >= a.S) { s_out[0] = (int)lo + 1; s_out[1] = 0; }
        }
    }
    __syncthreads();
}

template <bool MUFU1, bool GRAD, bool SIGNS, bool SHARED = false>
__device__ __forceinline__ void sweep_subchunk(int cls, const RowRegs &R, const float *se,
                                               const float *sx, const float *sa, float cabs,
                                               acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI], int (&ds)[kTileRI],
                                               const float *stp = nullptr) {
    if (cls == kClassPos) {  // (two call sites on purpose: one copy of the loop per sign, as measured)
        if (MUFU1 && SHARED) loop_const_shared<GRAD>(R, stp, true, dl, dg);
        else loop_const<MUFU1, GRAD>(R, se, sx, cabs, true, dl, dg);
        if (SIGNS) {
#pragma unroll
            for (int k = 0; k < kTileRI; ++k) ds[k] += kSubCols;
        }
    } else if (cls == kClassNeg) {
        if (MUFU1 && SHARED) loop_const_shared<GRAD>(R, stp, false, dl, dg);
        else loop_const<MUFU1, GRAD>(R, se, sx, cabs, false, dl, dg);
        if (SIGNS) {
#pragma unroll
            for (int k = 0; k < kTileRI; ++k) ds[k] -= kSubCols;
        }
    } else if (cls == kClassTie) {
        loop_tie<MUFU1, GRAD>(R, se, sx, cabs, dl, dg);
    } else {
        loop_general<MUFU1, GRAD, kSubCols, SIGNS>(R, se, sx, sa, cabs, dl, dg, ds);
    }
}

// The pair kernel.  One CTA of 512 threads per SM, split into two independent halves of 8 warps (256 threads,
// a 1024-row tile each, own staging buffers, own named barrier).  The CTA owns a contiguous, cost-balanced
// range of units; the halves consume it from both ends -- the front half forward, the back half backward --
// claiming a few units at a time from a shared counter until they meet.  Whichever half the SM's warp
// arbitration favours simply takes more units, so both stay busy to the end (with two independent CTAs per
// SM and a static split the favoured CTA finished at ~55 % of the kernel and the other ran alone, at lower
// MUFU utilisation, for the rest).  Row partials are fixed-point integers added to per-row accumulators with
// integer atomics, so the result does not depend on where the halves meet, on the number of CTAs, or -- in a
// sharded step, where this launch runs CTAs [c_first, c_first + gridDim.x) of a plan cut into G -- on the number
// of GPUs.
constexpr int kDuoThreads = 2 * kTileThreads;
constexpr int kStageSubs = kStageCols / kSubCols;
constexpr int kStageArrays = 4;  // 2^u, sgn(f) x, attribute; column-pair sums and products of 2^u (shared-reciprocal build)
constexpr int kDuoStageBytes = 2 * kStageArrays * kStageCols * (int)sizeof(float);  // 64 KiB

__device__ __forceinline__ void half_barrier(int half) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(kTileThreads) : "memory");
}

__device__ void shard_pair_kernel_tail(const TilesArgs &a, acc_t *sh);  // reg_shard.cuh

// ONLY1: a build of the kernel without the two-MUFU loops (the common case: image configs at delta = 1 have no
// outliers).  The complete kernel carries eight inlined pair loops; dropping four of them shortens the code the
// instruction caches have to hold and was measured 2-4 % faster on the all-inlier workload, so the host launches both
// builds and each decides on the device flag whether it is the one to run (the other exits at once).  (Both builds
// as two instantiations of the body inside ONE kernel was measured too: 3.7 % slower than the pair of launches; ptxas's
// schedule of the pair loops is sensitive even to the prologue -- a variant of find_unit that located both ends of the
// range in one pass changed it and cost 2 %.  Packed-FP32 versions of the tie and general loops (263 and 374 instead of
// 362 and 524 instructions per 32 pairs) made those tiles faster but the constant-sign loop's schedule worse: 5.38 vs
// 5.15 ms at C4, -2 % only on tie-heavy labels; as non-inlined functions they forced 25 register moves into the
// constant-sign loop.  Not kept.)
template <bool GRAD, bool SIGNS, bool ONLY1>
__global__ void __launch_bounds__(kDuoThreads, 1)
reg_tiles_kernel(TilesArgs a) {
    extern __shared__ __align__(16) float stage[];  // [2 halves][kStageArrays][kStageCols] = kDuoStageBytes (dynamic)
    __shared__ acc_t sred[2][kDuoThreads / 32];
    __shared__ int s_rng[4];
    __shared__ int s_scan[kDuoThreads];
    __shared__ unsigned int s_claimed;   // units granted so far (may overshoot N)
    __shared__ int s_grant[2][2];        // per half: first unit (linear) and count of the current grant

    pdl_wait();  // launched ahead of the plan kernel's end (programmatic dependent launch): wait for the plan
    if (a.dual && (a.flags[kFlagNeedComplete] != 0) == ONLY1) return;  // the other build's turn
    const long long c = (long long)a.c_first + blockIdx.x;
    if (a.dbg_times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg_times[2 * blockIdx.x] = t;
    }
    if (threadIdx.x == 0) s_claimed = 0u;
    const long long T = a.prefix[a.n_rr];
    find_unit<kDuoThreads>(a, ceil_share(c, T, a.G), s_rng, s_scan);
    find_unit<kDuoThreads>(a, ceil_share(c + 1, T, a.G), s_rng + 2, s_scan);
    const int64_t rr0 = s_rng[0];
    const int sp0 = s_rng[1];
    const int64_t N = ((int64_t)s_rng[2] - rr0) * a.S + s_rng[3] - sp0;  // units of this CTA, linear index u:
                                                                          // (rr, s') = (rr0 + (sp0+u)/S, (sp0+u)%S)
    const int half = threadIdx.x / kTileThreads;
    const int tid = threadIdx.x % kTileThreads;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kWarpRows = kTileRows / (kTileThreads / 32);
    float *hse = stage + (half * kStageArrays + 0) * kStageCols, *hsx = stage + (half * kStageArrays + 1) * kStageCols,
          *hsa = stage + (half * kStageArrays + 2) * kStageCols, *hstp = stage + (half * kStageArrays + 3) * kStageCols;

    int64_t cursor = half == 0 ? 0 : N;  // next unit from the front / one past the next unit from the back
    int64_t cur_rr = -1;
    acc_t lhi = 0, llo = 0;
    RowRegs R;
    bool valid[kTileRI];
    acc_t dl[kTileRI], dg[kTileRI];
    int ds[kTileRI];
    bool warp_has_rows = false;
    const float *Er = nullptr, *Xr = nullptr, *Ar = nullptr;
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) { valid[k] = false; dl[k] = 0; dg[k] = 0; ds[k] = 0; R.e[k] = 1.0f; R.f[k] = 1.0f; R.f2[k] = 1.0f; R.x[k] = 0.0f; R.a[k] = 0.0f; }

    // add this half's partial sums of row tile cur_rr to the row accumulators
    auto flush = [&]() {
        if (cur_rr < 0) return;
        const int64_t base = cur_rr * kTileRows + warp * kWarpRows + lane;
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
            const int64_t o = base + k * 32;  // rr * kTileRows + (row - tile start)
            if (valid[k]) {
                lhi += dl[k] >> kLossSplitBits;
                llo += dl[k] & kLossLoMask;
                if (GRAD && dg[k] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.acc_g + o), (unsigned long long)dg[k]);
                if (a.acc_l && dl[k] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.acc_l + o), (unsigned long long)dl[k]);
                if (SIGNS && ds[k] != 0) atomicAdd(a.acc_s + o, ds[k]);
            }
            dl[k] = 0;
            dg[k] = 0;
            ds[k] = 0;
        }
    };

    while (true) {
        // ---- claim: up to 8 units, never across a row-tile boundary, smaller near the end (guided) ----
        if (tid == 0) {
            const int64_t left = N - (int64_t)min((unsigned int)N, s_claimed);
            int64_t want = left / 6;
            want = want < 1 ? 1 : (want > kStageSubs ? kStageSubs : want);
            if (half == 0) {
                const int64_t in_tile = a.S - (sp0 + cursor) % a.S;         // units up to the end of the row tile
                want = min(want, in_tile);
            } else {
                const int64_t in_tile = (sp0 + cursor - 1) % a.S + 1;        // units back to the start of the row tile
                want = min(want, in_tile);
            }
            int64_t got = 0;
            if (left > 0 && cursor >= 0) {
                const unsigned int t0 = atomicAdd(&s_claimed, (unsigned int)want);
                got = (int64_t)t0 >= N ? 0 : min(want, N - (int64_t)t0);
            }
            s_grant[half][0] = (int)(half == 0 ? cursor : cursor - got);
            s_grant[half][1] = (int)got;
        }
        half_barrier(half);  // also: everybody in this half is done with the previous staging buffers
        const int u0 = s_grant[half][0], nsub = s_grant[half][1];
        if (nsub == 0) break;
        cursor = half == 0 ? cursor + nsub : cursor - nsub;
        const int64_t lin = (int64_t)sp0 + u0;
        const int64_t rr = rr0 + lin / a.S;
        const int sp = (int)(lin % a.S);

        if (rr != cur_rr) {  // new row tile for this half: flush the old one, load the new rows
            flush();
            cur_rr = rr;
            const int r = (int)(rr / a.n_row_tiles);
            const int I = (int)(rr % a.n_row_tiles);
            Er = a.Es + (int64_t)r * a.Bpad;
            Xr = a.Xs + (int64_t)r * a.Bpad;
            Ar = a.As + (int64_t)r * a.Bpad;
            const int *rp = a.rowpos ? a.rowpos + (int64_t)r * a.n_rows : nullptr;
            // Each warp owns 128 CONSECUTIVE sorted rows of the tile (lane l, k -> row 128 w + 32 k + l) and runs
            // the class planned for ITS OWN attribute range (plan_classes_kernel).
            const int64_t m0 = (int64_t)I * kTileRows + (int64_t)warp * kWarpRows;
            warp_has_rows = m0 < a.n_rows;
#pragma unroll
            for (int k = 0; k < kTileRI; ++k) {
                const int64_t m = m0 + (int64_t)k * 32 + lane;
                valid[k] = m < a.n_rows;
                const int64_t pos = valid[k] ? (rp ? (int64_t)rp[m] : m) : 0;
                R.e[k] = valid[k] ? Er[pos] : 1.0f;
                R.x[k] = valid[k] ? Xr[pos] : 0.0f;
                R.a[k] = valid[k] ? Ar[pos] : 0.0f;
                R.f[k] = exp2f(-(a.cabs * R.x[k]));  // as Es was built: 2^(cabs xs), with the opposite sign
                R.f2[k] = R.f[k] * R.f[k];
            }
        }

        // ---- stage the granted units (visited in the permuted order s = (s' P) mod S, see plan_classes_kernel)
        for (int q = tid * 4; q < nsub * kSubCols; q += kTileThreads * 4) {
            const int w = q / kSubCols;
            const int64_t col = (((int64_t)(sp + w) * a.P) % a.S) * kSubCols + (q - w * kSubCols);
            const float4 e4 = *reinterpret_cast<const float4 *>(Er + col);
            *reinterpret_cast<float4 *>(hse + q) = e4;
            if (ONLY1)  // per group of four columns: sums and products of the column pairs (j, j+2), (j+1, j+3)
                *reinterpret_cast<float4 *>(hstp + q) = make_float4(e4.x + e4.z, e4.y + e4.w, e4.x * e4.z, e4.y * e4.w);
            *reinterpret_cast<float4 *>(hsx + q) = *reinterpret_cast<const float4 *>(Xr + col);
            *reinterpret_cast<float4 *>(hsa + q) = *reinterpret_cast<const float4 *>(Ar + col);
        }
        half_barrier(half);
        if (warp_has_rows) {
            for (int w = 0; w < nsub; ++w) {
                const int sub = w * kSubCols;
                const unsigned int word = a.cls8[rr * a.S + sp + w];
                const int cls = (word >> (2 * warp)) & 3;        // planned class of this warp's tile
                const bool mufu1 = ONLY1 || ((word >> (16 + warp)) & 1u) == 0u;  // planned tanh form
                if (mufu1) sweep_subchunk<true, GRAD, SIGNS, ONLY1>(cls, R, hse + sub, hsx + sub, hsa + sub, a.cabs, dl, dg, ds, hstp + sub);
                else sweep_subchunk<false, GRAD, SIGNS>(cls, R, hse + sub, hsx + sub, hsa + sub, a.cabs, dl, dg, ds);
            }
        }
    }
    flush();

    lhi = warp_sum(lhi);
    llo = warp_sum(llo);
    if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = lhi; sred[1][threadIdx.x >> 5] = llo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        acc_t th = 0, tl = 0;
#pragma unroll
        for (int w = 0; w < kDuoThreads / 32; ++w) { th += sred[0][w]; tl += sred[1][w]; }
        a.lossp[2 * blockIdx.x] = th;
        a.lossp[2 * blockIdx.x + 1] = tl;
        if (a.dbg_times) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            a.dbg_times[2 * blockIdx.x + 1] = tt;
        }
    }
    // a sharded step: the last CTA publishes this GPU's loss partial and signals the peers (reg_shard.cuh)
    if (a.shard.G > 0) shard_pair_kernel_tail(a, reinterpret_cast<acc_t *>(stage));
}

// Loss of a launch from its per-CTA (hi, lo) partials: exact integer totals, one rounding at the end.
__device__ __forceinline__ void sum_loss_partials(const acc_t *__restrict__ lossp, int64_t n_cta, acc_t *sh /*smem [2][256]*/,
                                                  acc_t &hi, acc_t &lo) {
    acc_t th = 0, tl = 0;
    for (int64_t u = threadIdx.x; u < n_cta; u += blockDim.x) { th += lossp[2 * u]; tl += lossp[2 * u + 1]; }
    sh[threadIdx.x] = th;
    sh[256 + threadIdx.x] = tl;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sh[threadIdx.x] += sh[threadIdx.x + o]; sh[256 + threadIdx.x] += sh[256 + threadIdx.x + o]; }
        __syncthreads();
    }
    hi = sh[0];
    lo = sh[256];
}
__device__ __forceinline__ double loss_from_hilo(acc_t hi, acc_t lo) {
    return (double)hi * kLossHiScale + (double)lo * kFixScale;
}

__global__ void __launch_bounds__(256)
reg_tiles_epilogue_kernel(TilesArgs a, const int *__restrict__ perm, int R, int64_t row_begin, int n_cta,
                          double gscale, double lscale, double pad_per_row,
                          float *__restrict__ grad_cols, double *__restrict__ row_loss, int *__restrict__ row_sign,
                          double *__restrict__ loss_out, float *__restrict__ loss_f32_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over R * n_rows, m fastest
    if (idx < (int64_t)R * a.n_rows) {
        const int r = (int)(idx / a.n_rows);
        const int64_t m = idx % a.n_rows;
        const int64_t I = m / kTileRows, lr = m % kTileRows;
        const int64_t rr = (int64_t)r * a.n_row_tiles + I;
        const int64_t pos = a.rowpos ? (int64_t)a.rowpos[(int64_t)r * a.n_rows + m] : m;
        const int64_t out = ((int64_t)perm[(int64_t)r * a.Bpad + pos] - row_begin) * R + r;
        const bool poisoned = row_is_poisoned(a.flags, r, a.Xs[(int64_t)r * a.Bpad + pos]);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (grad_cols) grad_cols[out] = poisoned ? (float)nan : (float)((double)a.acc_g[rr * kTileRows + lr] * kFixScale * gscale);
        if (row_loss) row_loss[out] = poisoned ? nan : (double)a.acc_l[rr * kTileRows + lr] * kFixScale - pad_per_row;
        if (row_sign) row_sign[out] = a.acc_s[rr * kTileRows + lr];
    }
    if (blockIdx.x == 0) {
        __shared__ acc_t sh[512];
        acc_t hi, lo;
        sum_loss_partials(a.lossp, n_cta, sh, hi, lo);
        if (threadIdx.x == 0) {
            double total = loss_from_hilo(hi, lo) - pad_per_row * (double)a.n_rows * (double)R;
            if (any_nonfinite(a.flags, R)) total = __longlong_as_double(0x7ff8000000000000LL);  // as the reference's float sum
            *loss_out = total * lscale;
            if (loss_f32_out) *loss_f32_out = (float)(total * lscale);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int golden_stride(int S) {  // integer nearest 0.618 S that is coprime to S
    if (S <= 2) return 1;
    auto gcd = [](int x, int y) { while (y) { int t = x % y; x = y; y = t; } return x; };
    int p = (int)(0.6180339887 * S + 0.5);
    if (p < 1) p = 1;
    for (int d = 0; d < S; ++d) {
        if (p + d < S && gcd(p + d, S) == 1) return p + d;
        if (p - d >= 1 && gcd(p - d, S) == 1) return p - d;
    }
    return 1;
}

static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return dev;
}

// CTAs of the pair kernel that fit one SM (1 with its 512 threads x 128 registers); also opts the kernel into
// 48 KiB of dynamic shared memory.  Function attributes are per device: cached per device.
static int tiles_ctas_per_sm() {
    static int cache[64] = {};
    int &v = cache[current_device_slot()];
    if (v == 0) {
        int n = 0;
        cudaFuncSetAttribute(reg_tiles_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoStageBytes);
        cudaFuncSetAttribute(reg_tiles_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoStageBytes);
        cudaFuncSetAttribute(reg_tiles_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoStageBytes);
        cudaFuncSetAttribute(reg_tiles_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoStageBytes);
        cudaFuncSetAttribute(reg_tiles_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDuoStageBytes);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reg_tiles_kernel<true, false, false>, kDuoThreads, kDuoStageBytes);
        if (e != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            n = 1;
        }
        v = n;
    }
    return v;
}

// a.dual (set by the caller BEFORE the plan is made: the cost model depends on it): 0 = the complete build only (parity
// instrumentation), 1 = both builds, the device flag decides which one works
static void launch_tiles(TilesArgs a, int n_cta, bool want_grad, bool want_signs, cudaStream_t st) {
    const dim3 grid((unsigned)n_cta), block(kDuoThreads);
    if (want_signs) {
        launch_kernel(reg_tiles_kernel<true, true, false>, grid, block, kDuoStageBytes, st, true, a);
        return;
    }
    if (want_grad) {
        launch_kernel(reg_tiles_kernel<true, false, true>, grid, block, kDuoStageBytes, st, true, a);
        launch_kernel(reg_tiles_kernel<true, false, false>, grid, block, kDuoStageBytes, st, true, a);
    } else {
        launch_kernel(reg_tiles_kernel<false, false, true>, grid, block, kDuoStageBytes, st, true, a);
        launch_kernel(reg_tiles_kernel<false, false, false>, grid, block, kDuoStageBytes, st, true, a);
    }
    count_launch();
}

void reg_scales(const RegProblem &P, int64_t Bpad, double &lscale, double &gscale, double &pad_per_row) {
    const double BB = (double)P.B * (double)P.B;
    lscale = (double)P.gamma / BB;
    gscale = 8.0 * (double)P.gamma * (double)P.factor / BB;
    pad_per_row = (double)(Bpad - P.B);
}

#include "reg_tri.cuh"

SortedLayout sorted_layout(int64_t B_total, int64_t n_rows, int R, int sm_count, bool with_triangle) {
    SortedLayout L;
    L.N = sort_padded_size(B_total > 0 ? B_total : 1);
    L.Bpad = round_up(B_total > 0 ? B_total : 1, kSubCols);
    L.n_row_tiles = (int)ceil_div(n_rows > 0 ? n_rows : 1, kTileRows);
    L.S = (int)(L.Bpad / kSubCols);
    L.n_rr = (int64_t)R * L.n_row_tiles;
    L.F = L.n_rr * L.S;
    const int per_sm = 2;  // the pair kernel runs 1 CTA (two halves) per SM, the triangle variant 2 CTAs per SM
    int64_t G = (int64_t)sm_count * per_sm;
    if (G > L.F / 4) G = L.F / 4;  // >= 4 units per CTA on average: every CTA owns at least one unit
    L.G_max = (int)(G > 0 ? G : 1);
    const bool tri_capable = with_triangle && (n_rows == B_total && B_total > 0);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    L.off_keys = take(sizeof(unsigned long long) * (size_t)R * L.N);
    L.off_Us = take(sizeof(float) * (size_t)R * L.Bpad);
    L.off_As = take(sizeof(float) * (size_t)R * L.Bpad);
    L.off_Es = take(sizeof(float) * (size_t)R * L.Bpad);
    L.off_perm = take(sizeof(int) * (size_t)R * L.Bpad);
    L.off_rowpos = take(sizeof(int) * (size_t)R * (n_rows > 0 ? n_rows : 1));
    L.off_flags = take(sizeof(int) * kFlagInts);
    L.off_blockcnt = take(sizeof(int) * (size_t)R * (size_t)ceil_div(L.Bpad, 256));
    L.off_cls8 = take(sizeof(unsigned int) * (size_t)L.F);
    L.off_cost8 = take(sizeof(unsigned short) * (size_t)L.F);
    L.off_combo = take(sizeof(int) * (size_t)L.n_rr);
    L.off_prefix = take(sizeof(long long) * (size_t)(L.n_rr + 1));
    // run slots of the radix chunk sort + rank merge (all rows of one GPU); larger batches keep the bitonic network
    L.n_runs = (n_rows == B_total && !with_triangle && ceil_div(B_total > 0 ? B_total : 1, kRunCap) <= kMaxRuns)
                   ? (int)ceil_div(B_total > 0 ? B_total : 1, kRunCap) : 0;
    L.off_runs = take(sizeof(RunElem) * (size_t)L.n_runs * R * kRunSlotElems);
    // row accumulators: gradient, loss, sign sums -- contiguous so that one memset clears the ones in use
    L.acc_bytes = sizeof(acc_t) * (size_t)L.n_rr * kTileRows;
    L.off_acc_g = take(L.acc_bytes);
    L.off_acc_l = take(L.acc_bytes);
    L.off_acc_s = take(sizeof(int) * (size_t)L.n_rr * kTileRows);
    L.off_lossp = take(sizeof(acc_t) * 2 * (size_t)L.G_max);
    L.off_colpart = take(tri_capable ? sizeof(float2) * (size_t)R * (size_t)colpart_size(L.n_row_tiles, L.Bpad) : 0);
    L.off_eloss = take(sizeof(double) * (size_t)(ceil_div((n_rows > 0 ? n_rows : 1) * (int64_t)R, 256)));
    L.off_dbg = take(sizeof(unsigned long long) * 2 * (size_t)L.G_max);
    L.bytes = off;
    return L;
}

#include "reg_shard.cuh"

int run_reg_sorted(const RegProblem &P, const SortedLayout &L, char *ws, cudaStream_t st) {
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + L.off_keys);
    float *Xs = reinterpret_cast<float *>(ws + L.off_Us);
    float *As = reinterpret_cast<float *>(ws + L.off_As);
    float *Es = reinterpret_cast<float *>(ws + L.off_Es);
    int *perm = reinterpret_cast<int *>(ws + L.off_perm);
    int *rowpos = reinterpret_cast<int *>(ws + L.off_rowpos);
    int *flags = reinterpret_cast<int *>(ws + L.off_flags);
    int *n_in = flags + kFlagNIn;
    int *blockcnt = reinterpret_cast<int *>(ws + L.off_blockcnt);
    const int64_t n_rows = P.row_end - P.row_begin;
    const bool want_grad = P.grad_cols_out != nullptr;
    const bool want_signs = P.row_sign_out != nullptr;
    const bool all_rows = (P.row_begin == 0 && P.row_end == P.B);
    const bool triangle = P.use_triangle && all_rows && n_rows > 0;
    if (want_signs && (triangle || !want_grad)) {
        set_error("row sign sums need the gradient pass of the dense or sorted algorithm");
        return ARVAE_E_BADARG;
    }

    const double c = 2.0 * (double)P.factor * 1.4426950408889634074;  // 2 f log2(e)
    const float fsign = P.factor > 0.f ? 1.0f : (P.factor < 0.f ? -1.0f : 0.0f);
    const float cabs = P.factor != 0.f ? (float)fabs(c) : 1.0f;  // f == 0: xs == 0, any scale works
    KeySpec spec;
    spec.lab = P.lab; spec.lrs = P.lrs; spec.lcs = P.lcs;
    spec.z = P.z; spec.zrs = P.zrs; spec.zcs = P.zcs;
    spec.fsign = fsign;
    spec.cabs = cabs;
    // triangle mode needs ONE attribute order per dim: no outlier segment there, the whole dim falls back to the
    // two-MUFU form instead (sorted_gather_kernel's dimflags)
    spec.segment = triangle ? 0 : 1;
    spec.idx_offset = 0;
    spec.dims = P.dims;
    timeline_mark(st, "begin");
    ARVAE_CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int) * kFlagClearInts, st));
    int rc = 0;
    if (L.n_runs > 0 && all_rows && !triangle && n_rows > 0) {
        // radix-sorted runs of <= kRunCap samples, merged by rank: the same pipeline a sharded step runs per GPU
        RunSet rs;
        memset(&rs, 0, sizeof(rs));
        const int64_t sizes[1] = {P.B};
        if (fill_run_set(rs, sizes, 1, nullptr) != 0) {
            set_error("internal: run set overflow");
            return ARVAE_E_BADARG;
        }
        rs.R_cap = P.R;
        rs.base = ws + L.off_runs;
        RunDest dest;
        memset(&dest, 0, sizeof(dest));
        dest.n_dest = 1; dest.R_cap = P.R; dest.base[0] = ws + L.off_runs;
        rc = run_chunk_sort(spec, P.R, P.B, 0, dest, nullptr, st);
        if (rc) return rc;
        timeline_mark(st, "sort");
        PosDest pd;
        memset(&pd, 0, sizeof(pd));  // one GPU: the merge places the elements itself
        rc = launch_runs_merge(rs, P.R, L.Bpad, cabs, Xs, As, Es, perm, flags, nullptr, 0, 0, 0, pd, 0, -1, st);
        if (rc) return rc;
        timeline_mark(st, "merge");
    } else {
        rc = run_sort_keys_spec(spec, P.R, P.B, L.N, keys, st);
        if (rc) return rc;
        timeline_mark(st, "sort");
        dim3 gg((unsigned)ceil_div(L.Bpad, 256), (unsigned)P.R);
        sorted_gather_kernel<<<gg, 256, 0, st>>>(keys, L.N, P.z, P.zrs, P.zcs, P.lab, P.lrs, P.lcs, P.dims,
                                                 P.B, L.Bpad, fsign, cabs, Xs, As, Es, perm, flags,
                                                 P.row_begin, P.row_end, all_rows ? nullptr : blockcnt, triangle ? 1 : 0);
        ARVAE_LAUNCH_CHECK("sorted_gather_kernel");
        timeline_mark(st, "gather");
    }
    if (!all_rows && n_rows > 0) {
        dim3 gs((unsigned)ceil_div(L.Bpad, 1024), (unsigned)P.R);
        row_select_kernel<<<gs, 1024, 0, st>>>(perm, blockcnt, (int)ceil_div(L.Bpad, 256), P.B, L.Bpad, P.row_begin, P.row_end,
                                               n_rows, rowpos);
        ARVAE_LAUNCH_CHECK("row_select_kernel");
    }

    TilesArgs a;
    memset(&a, 0, sizeof(a));
    a.Xs = Xs; a.Es = Es; a.As = As;
    a.cabs = cabs;
    a.rowpos = all_rows ? nullptr : rowpos;
    a.flags = flags;
    a.n_in = n_in;
    a.Bpad = L.Bpad; a.n_rows = n_rows;
    a.n_row_tiles = L.n_row_tiles; a.S = L.S; a.F = L.F; a.n_rr = L.n_rr;
    a.P = golden_stride(L.S);
    int64_t G = (int64_t)sm_count() * tiles_ctas_per_sm();
    if (G > L.G_max) G = L.G_max;
    if (G < 1) G = 1;
    a.G = (int)G;
    a.c_first = 0;
    a.force_general = 0;
    a.cls8 = reinterpret_cast<unsigned int *>(ws + L.off_cls8);
    a.cost8 = reinterpret_cast<unsigned short *>(ws + L.off_cost8);
    a.prefix = reinterpret_cast<long long *>(ws + L.off_prefix);
    int *combo_cost = reinterpret_cast<int *>(ws + L.off_combo);
    a.acc_g = reinterpret_cast<acc_t *>(ws + L.off_acc_g);
    a.acc_l = P.row_loss_out ? reinterpret_cast<acc_t *>(ws + L.off_acc_l) : nullptr;
    a.acc_s = want_signs ? reinterpret_cast<int *>(ws + L.off_acc_s) : nullptr;
    a.lossp = reinterpret_cast<acc_t *>(ws + L.off_lossp);
    a.dbg_times = getenv("ARVAE_DEBUG_TIMES") ? reinterpret_cast<unsigned long long *>(ws + L.off_dbg) : nullptr;
    a.colpart = nullptr; a.Pinv = 0; a.B = P.B;
    a.dual = want_signs ? 0 : 1;
    if (triangle) return run_reg_tri_tail(P, L, a, perm, combo_cost, ws, st);

    if (n_rows > 0) {
        plan_classes_kernel<<<(unsigned)L.n_rr, 256, 0, st>>>(a, combo_cost, (want_grad ? 1 : 0) | (a.acc_l ? 2 : 0) | (a.acc_s ? 4 : 0),
                                                             reinterpret_cast<unsigned int *>(flags + kFlagTicket));
        ARVAE_LAUNCH_CHECK("plan_classes_kernel");
        timeline_mark(st, "plan");
        profile_begin(st);
        launch_tiles(a, a.G, want_grad, want_signs, st);
        profile_end(st);
        ARVAE_LAUNCH_CHECK("reg_tiles_kernel");
        timeline_mark(st, "pairs");
    }

    double lscale, gscale, pad_per_row;
    reg_scales(P, L.Bpad, lscale, gscale, pad_per_row);
    const int64_t work = n_rows * P.R;
    reg_tiles_epilogue_kernel<<<(unsigned)(work > 0 ? ceil_div(work, 256) : 1), 256, 0, st>>>(
        a, perm, P.R, P.row_begin, n_rows > 0 ? a.G : 0, gscale, lscale, pad_per_row, P.grad_cols_out, P.row_loss_out,
        P.row_sign_out, P.loss_out, P.loss_f32_out);
    ARVAE_LAUNCH_CHECK("reg_tiles_epilogue_kernel");
    timeline_mark(st, "epilogue");
    return 0;
}

}  // namespace arvae
