// reg_tri.cuh -- triangle evaluation for the attribute-sorted path (included by reg_sorted.cu).
//
// When a call covers ALL rows (single GPU), rows and columns are the same samples in the same sorted
// order, and L_ij = L_ji, g_ji = -g_ij.  A constant-sign tile strictly above the
// diagonal is therefore evaluated ONCE and credited to both sides ("double duty"):
//   rows i of the tile (s_ij = -1, a_i < a_j):  sum_j 2(1 - r_ij)   and  +4 sum_j (r_ij - r_ij^2)
//   cols j of the tile, as rows of the mirrored pairs (s_ji = +1, r_ji = 1 - r_ij):
//                                               sum_i 2(1 - r_ij)   and  -4 sum_i (r_ij - r_ij^2)
// so besides the row sums of r and r^2 the kernel needs their COLUMN sums over the tile's rows.  Those
// are reduced across the warp in fixed point (REDUX.SUM on 2^20-scaled integers: integer addition is
// associative, the result does not depend on the order), added to per-CTA shared-memory accumulators
// (integer ATOMS across the 8 warps, again order-free) and written once per (row tile, column).  The
// mirrored constant-sign tiles below the diagonal are skipped; everything that is not constant-sign
// (the diagonal band, tie groups, NaN rows) is evaluated from both sides exactly as before, so the
// per-pair arithmetic and the sign matrix are unchanged.  Evaluated pairs ~ B^2 R / 2.
#pragma once

#ifndef ARVAE_TRI_REDUX
#define ARVAE_TRI_REDUX 0
#endif

// per half sub-chunk (128 columns) and warp: what to run
enum TriCode { kTriSkip = 0, kTriDouble = 1, kTriTie = 2, kTriGeneral = 3 };
// Fixed-point scale of the column sums: 2^20.  A thread's partial (4 rows, each r <= 1) is < 8, so
// adding 8.0f = 2^(23-20) leaves round(x * 2^20) in the mantissa bits -- a float-to-fixed conversion on the
// FP32 pipe (F2I would run on the MUFU/XU pipe, the very pipe this kernel is bound by).
constexpr float kColFix = 1048576.0f;
constexpr float kColMagic = 8.0f;
constexpr int kHalfCols = kSubCols / 2;
constexpr int kWarpsPerTile = kTileThreads / 32;
constexpr int kWarpRowsT = kTileRows / kWarpsPerTile;  // 128
static_assert(kWarpRowsT == kHalfCols, "row groups and half sub-chunks are both 128 positions");
static_assert(kWarpsPerTile * 4 <= 32, "triangle class word holds 4 bits per warp");

// first column of row tile I's strip in colpart (only columns at or above the tile are stored)
__host__ __device__ __forceinline__ int64_t colpart_base(int64_t I, int64_t Bpad) {
    return I * Bpad - (int64_t)kTileRows * (I * (I - 1) / 2);
}
__host__ __device__ __forceinline__ int64_t colpart_size(int64_t n_row_tiles, int64_t Bpad) {
    return colpart_base(n_row_tiles, Bpad);
}

__device__ __forceinline__ int tri_cost(int code, bool mufu1) {  // per half tile (128 x 128)
    return code == kTriSkip ? 0 : (code == kTriDouble ? (mufu1 ? 11 : 17) : (code == kTriTie ? (mufu1 ? 6 : 8) : 10));
}

// One CTA per row tile: the 4-bit code pair of every (warp, sub-chunk) and the unit costs, in visiting order.
__global__ void __launch_bounds__(256)
tri_plan_kernel(TilesArgs a, int *__restrict__ combo_cost) {
    __shared__ int sred[8];
    const int64_t rr = blockIdx.x;
    const int r = (int)(rr / a.n_row_tiles), I = (int)(rr % a.n_row_tiles);
    const float *Ar = a.As + (int64_t)r * a.Bpad;
    const bool mufu1 = a.flags[r] == 0;
    const int64_t n_groups = ceil_div(a.n_rows, kWarpRowsT);
    int total = 0;
    for (int sp = threadIdx.x; sp < a.S; sp += 256) {
        const int64_t J = ((int64_t)sp * a.P) % a.S;
        const float cmin = Ar[J * kSubCols], cmax = Ar[J * kSubCols + kSubCols - 1];
        unsigned int word = 0;
        int cost = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerTile; ++w) {
            const int64_t g = (int64_t)I * kWarpsPerTile + w;  // row group = positions [128 g, 128 g + 128)
            if (g >= n_groups) continue;                      // warp without rows
            int c0, c1;
            if (2 * J >= g + 1) {  // strictly above the group's rows
                const int64_t last = min((g + 1) * kWarpRowsT, a.n_rows) - 1;
                const int cls = classify(Ar[g * kWarpRowsT], Ar[last], cmin, cmax);
                c0 = c1 = (cls == kClassNeg) ? kTriDouble : (cls == kClassTie ? kTriTie : kTriGeneral);
            } else if (2 * (J + 1) <= g) {  // strictly below: evaluate only what the mirrored block did not cover
                const int64_t Jq = g / 2;   // the sub-chunk that holds this group when it acts as columns
                const float qmin = Ar[Jq * kSubCols], qmax = Ar[Jq * kSubCols + kSubCols - 1];
                int cc[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t hg = 2 * J + h;  // the half's positions, as a row group of the mirrored block
                    const bool strictly_upper = 2 * Jq >= hg + 1;
                    const int cls = classify(Ar[hg * kWarpRowsT], Ar[hg * kWarpRowsT + kWarpRowsT - 1], qmin, qmax);
                    cc[h] = (strictly_upper && cls == kClassNeg) ? kTriSkip : (cls == kClassTie ? kTriTie : kTriGeneral);
                }
                c0 = cc[0];
                c1 = cc[1];
            } else {  // overlaps the group's own rows: the diagonal
                c0 = c1 = kTriGeneral;
            }
            word |= (unsigned int)(c0 | (c1 << 2)) << (4 * w);
            cost += tri_cost(c0, mufu1) + tri_cost(c1, mufu1);
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
}

// Double-duty constant-sign (s_ij = -1) tile: row sums in registers, column sums to scol (fixed point).
template <bool MUFU1, bool GRAD, int ncols>
__device__ __forceinline__ void loop_double(const RowRegs &R, const float *__restrict__ se,
                                            const float *__restrict__ sx, float cabs,
                                            unsigned int *__restrict__ scol /* [ncols][2] for these columns */,
                                            acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
    const int lane = threadIdx.x & 31;
    float A1[kTileRI][4], A2[kTileRI][4];
#pragma unroll
    for (int k = 0; k < kTileRI; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) A1[k][q] = A2[k][q] = 0.0f;
    const float *sv = MUFU1 ? se : sx;
#pragma unroll 2
    for (int q = 0; q < ncols; q += 4) {
        const float4 vj = *reinterpret_cast<const float4 *>(sv + q);
        const float vv[4] = {vj.x, vj.y, vj.z, vj.w};
        float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // rows without a sample carry E_i = +inf (x_i = +inf): r = 0, no contribution to the column sums
                const float r = MUFU1 ? pair_r<true>(R.e[k], vv[e], 0.0f, cabs)
                                      : pair_r<false>(0.0f, 0.0f, R.x[k] - vv[e], cabs);
                A1[k][e] += r;
                c1[e] += r;
                if (GRAD) {
                    A2[k][e] = fmaf(r, r, A2[k][e]);
                    c2[e] = fmaf(r, r, c2[e]);
                }
            }
        }
        // Column sums over the warp's 128 rows.  Each lane's partial goes to 2^20 fixed point on the FP32
        // pipe (x + 8.0f leaves round(x 2^20) in the mantissa; F2I and REDUX would both run on the XU pipe this
        // kernel is bound by), then a transposing butterfly with INTEGER adds (exact, order-free): 8 values
        // per lane -> 4 -> 2 -> 1, after which every 4th lane owns one column total and adds it to the CTA's
        // shared-memory accumulator with an integer atomic.
#if ARVAE_TRI_REDUX
        // variant: warp-wide integer REDUX.SUM per value (runs on the XU pipe, which has slack in this kernel)
        unsigned int mine = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int s1 = __reduce_add_sync(0xffffffffu, __float_as_uint(c1[e] + kColMagic)) -
                                    32u * __float_as_uint(kColMagic);
            if (lane == e) mine = s1;
            const unsigned int s2 = __reduce_add_sync(0xffffffffu, __float_as_uint(c2[e] + kColMagic)) -
                                    32u * __float_as_uint(kColMagic);
            if (lane == 4 + e) mine = s2;
        }
        if (lane < 8) atomicAdd(scol + (q + (lane & 3)) * 2 + (lane >> 2), mine);
#else
        unsigned int u[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            u[e] = __float_as_uint(c1[e] + kColMagic);
            u[4 + e] = __float_as_uint(c2[e] + kColMagic);
        }
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
        unsigned int w4[4], w2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned int keep = h16 ? u[4 + i] : u[i], send = h16 ? u[i] : u[4 + i];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const unsigned int keep = h8 ? w4[2 + i] : w4[i], send = h8 ? w4[i] : w4[2 + i];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        unsigned int tot = (h4 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? w2[0] : w2[1], 4);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        tot -= 32u * __float_as_uint(kColMagic);  // remove the 32 magic offsets (unsigned wrap-around is exact)
        if ((lane & 3) == 0)  // lane bits: 16 -> sum r / sum r^2, 8 and 4 -> which of the 4 columns
            atomicAdd(scol + (q + (h8 ? 2 : 0) + (h4 ? 1 : 0)) * 2 + (h16 ? 1 : 0), tot);
#endif
    }
#pragma unroll
    for (int k = 0; k < kTileRI; ++k) {
        const float S1 = (A1[k][0] + A1[k][1]) + (A1[k][2] + A1[k][3]);
        const float S2 = (A2[k][0] + A2[k][1]) + (A2[k][2] + A2[k][3]);
        acc_add(dl[k], 2.0f * ((float)ncols - S1));  // s = -1: |t - s| = 2 (1 - r)
        if (GRAD) acc_add(dg[k], S1 - S2);            //         g / 4 = +(r - r^2)
    }
}

template <bool MUFU1, bool GRAD, int ncols>
__device__ __forceinline__ void tri_run(int code, const RowRegs &R, const float *se, const float *sx,
                                        const float *sa, float cabs, unsigned int *scol,
                                        acc_t (&dl)[kTileRI], acc_t (&dg)[kTileRI]) {
    if (code == kTriDouble) loop_double<MUFU1, GRAD, ncols>(R, se, sx, cabs, scol, dl, dg);
    else if (code == kTriTie) loop_tie<MUFU1, GRAD, ncols>(R, se, sx, cabs, dl, dg);
    else if (code == kTriGeneral) loop_general<MUFU1, GRAD, ncols>(R, se, sx, sa, cabs, dl, dg);
}

__device__ __forceinline__ bool word_has_double(unsigned int word) {
    // any 2-bit field equal to kTriDouble (01)
    return ((word & 0x55555555u) & ~((word >> 1) & 0x55555555u)) != 0;
}

template <bool GRAD>
__global__ void __launch_bounds__(kTileThreads)
reg_tri_kernel(TilesArgs a) {
    constexpr int kStageSubs = kStageCols / kSubCols;
    __shared__ __align__(16) float se[kStageCols];
    __shared__ __align__(16) float sx[kStageCols];
    __shared__ __align__(16) float sa[kStageCols];
    __shared__ unsigned int scol[kStageCols * 2];  // fixed-point column sums of the current batch (<= 1024 * 2^20 = 2^30)
    __shared__ unsigned int swords[kStageSubs];
    __shared__ int sJ[kStageSubs];
    __shared__ acc_t sred[2][kTileThreads / 32];
    __shared__ int s_rng[4];
    __shared__ int s_scan[kTileThreads];

    const long long c = blockIdx.x;
    if (a.dbg_times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg_times[2 * c] = t;
    }
    for (int q = threadIdx.x; q < kStageCols * 2; q += kTileThreads) scol[q] = 0;
    const long long T = a.prefix[a.n_rr];
    find_unit<kTileThreads>(a, ceil_share(c, T, a.G), s_rng, s_scan);
    find_unit<kTileThreads>(a, ceil_share(c + 1, T, a.G), s_rng + 2, s_scan);
    int64_t rr = s_rng[0];
    int sp0 = s_rng[1];
    const int64_t rr_end = s_rng[2];
    const int sp_end = s_rng[3];
    acc_t lhi = 0, llo = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    while (rr < rr_end || (rr == rr_end && sp0 < sp_end)) {
        const int s0 = sp0;
        const int s1 = (rr == rr_end) ? sp_end : a.S;
        const int r = (int)(rr / a.n_row_tiles);
        const int I = (int)(rr % a.n_row_tiles);
        const bool mufu1 = a.flags[r] == 0;
        const float *Er = a.Es + (int64_t)r * a.Bpad;
        const float *Xr = a.Xs + (int64_t)r * a.Bpad;
        const float *Ar = a.As + (int64_t)r * a.Bpad;
        float2 *cp = a.colpart + (int64_t)r * colpart_size(a.n_row_tiles, a.Bpad) + colpart_base(I, a.Bpad) -
                     (int64_t)I * kTileRows;  // indexed by absolute column

        const int64_t m0 = (int64_t)I * kTileRows + (int64_t)warp * kWarpRowsT;
        const bool warp_has_rows = m0 < a.n_rows;
        RowRegs R;
        bool valid[kTileRI];
        acc_t dl[kTileRI], dg[kTileRI];
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
            const int64_t m = m0 + (int64_t)k * 32 + lane;
            valid[k] = m < a.n_rows;
            const int64_t pos = valid[k] ? m : 0;
            R.e[k] = valid[k] ? Er[pos] : __int_as_float(0x7f800000);
            R.x[k] = valid[k] ? Xr[pos] : __int_as_float(0x7f800000);
            R.a[k] = valid[k] ? Ar[pos] : 0.0f;
            dl[k] = 0;
            dg[k] = 0;
        }

        for (int sp = s0; sp < s1 + kStageSubs; sp += kStageSubs) {
            // (1) everybody is done with the previous batch: flush its column sums, stage the next one
            __syncthreads();
            for (int w = 0; w < kStageSubs; ++w) {
                if (sp > s0 && word_has_double(swords[w])) {  // previous batch's unit w had double-duty tiles
                    const int64_t col = (int64_t)sJ[w] * kSubCols + threadIdx.x;  // kTileThreads == kSubCols
                    const unsigned int i1 = scol[(w * kSubCols + threadIdx.x) * 2], i2 = scol[(w * kSubCols + threadIdx.x) * 2 + 1];
                    cp[col] = make_float2((float)i1 * (1.0f / kColFix), (float)i2 * (1.0f / kColFix));
                    scol[(w * kSubCols + threadIdx.x) * 2] = 0;
                    scol[(w * kSubCols + threadIdx.x) * 2 + 1] = 0;
                }
            }
            __syncthreads();
            if (sp >= s1) break;
            const int nsub = min(kStageSubs, s1 - sp);
            if (threadIdx.x < kStageSubs) {
                const bool in = threadIdx.x < nsub;
                swords[threadIdx.x] = in ? a.cls8[rr * a.S + sp + threadIdx.x] : 0u;
                sJ[threadIdx.x] = in ? (int)(((int64_t)(sp + threadIdx.x) * a.P) % a.S) : 0;
            }
            __syncthreads();
            for (int q = threadIdx.x * 4; q < nsub * kSubCols; q += kTileThreads * 4) {
                const int w = q / kSubCols;
                if (swords[w] == 0u) continue;  // nobody works on this unit
                const int64_t col = (int64_t)sJ[w] * kSubCols + (q - w * kSubCols);
                *reinterpret_cast<float4 *>(se + q) = *reinterpret_cast<const float4 *>(Er + col);
                *reinterpret_cast<float4 *>(sx + q) = *reinterpret_cast<const float4 *>(Xr + col);
                *reinterpret_cast<float4 *>(sa + q) = *reinterpret_cast<const float4 *>(Ar + col);
            }
            __syncthreads();
            if (warp_has_rows) {
                for (int w = 0; w < nsub; ++w) {
                    const int code = (swords[w] >> (4 * warp)) & 15;
                    if (code == 0) continue;
                    const int c0 = code & 3, c1 = code >> 2;
                    const int sub = w * kSubCols;
                    if (c0 == c1) {
                        if (mufu1) tri_run<true, GRAD, kSubCols>(c0, R, se + sub, sx + sub, sa + sub, a.cabs, scol + sub * 2, dl, dg);
                        else tri_run<false, GRAD, kSubCols>(c0, R, se + sub, sx + sub, sa + sub, a.cabs, scol + sub * 2, dl, dg);
                    } else {
                        if (mufu1) {
                            tri_run<true, GRAD, kHalfCols>(c0, R, se + sub, sx + sub, sa + sub, a.cabs, scol + sub * 2, dl, dg);
                            tri_run<true, GRAD, kHalfCols>(c1, R, se + sub + kHalfCols, sx + sub + kHalfCols, sa + sub + kHalfCols, a.cabs,
                                                           scol + (sub + kHalfCols) * 2, dl, dg);
                        } else {
                            tri_run<false, GRAD, kHalfCols>(c0, R, se + sub, sx + sub, sa + sub, a.cabs, scol + sub * 2, dl, dg);
                            tri_run<false, GRAD, kHalfCols>(c1, R, se + sub + kHalfCols, sx + sub + kHalfCols, sa + sub + kHalfCols, a.cabs,
                                                            scol + (sub + kHalfCols) * 2, dl, dg);
                        }
                    }
                }
            }
        }
        // reset the batch bookkeeping for the next row tile (all flushed above)
        __syncthreads();
        if (threadIdx.x < kStageSubs) swords[threadIdx.x] = 0u;

        // this CTA's share of row tile rr goes to the fixed-point row accumulators (integer atomics: order-free)
#pragma unroll
        for (int k = 0; k < kTileRI; ++k) {
            const int64_t o = rr * kTileRows + warp * kWarpRowsT + k * 32 + lane;
            if (valid[k]) {
                lhi += dl[k] >> kLossSplitBits;
                llo += dl[k] & kLossLoMask;
                if (GRAD && dg[k] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.acc_g + o), (unsigned long long)dg[k]);
                if (a.acc_l && dl[k] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.acc_l + o), (unsigned long long)dl[k]);
            }
        }
        ++rr;
        sp0 = 0;
    }

    lhi = warp_sum(lhi);
    llo = warp_sum(llo);
    if (lane == 0) { sred[0][warp] = lhi; sred[1][warp] = llo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        acc_t th = 0, tl = 0;
#pragma unroll
        for (int w = 0; w < kTileThreads / 32; ++w) { th += sred[0][w]; tl += sred[1][w]; }
        a.lossp[2 * c] = th;
        a.lossp[2 * c + 1] = tl;
        if (a.dbg_times) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            a.dbg_times[2 * c + 1] = tt;
        }
    }
}

// Row q of dim r: row-side partial sums (segment slots) + column-side credits from every row tile at
// or below q's own that ran double-duty tiles on q's half sub-chunk.
__global__ void __launch_bounds__(256)
reg_tri_epilogue_kernel(TilesArgs a, const int *__restrict__ perm, int R, double gscale,
                        double pad_per_row, float *__restrict__ grad_cols,
                        double *__restrict__ row_loss, double *__restrict__ eloss) {
    __shared__ double sred[8];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over R * n_rows, q fastest
    double my_col_loss = 0.0;
    if (idx < (int64_t)R * a.n_rows) {
        const int r = (int)(idx / a.n_rows);
        const int64_t q = idx % a.n_rows;
        const int64_t Iq = q / kTileRows, lr = q % kTileRows;
        const int64_t rr = (int64_t)r * a.n_row_tiles + Iq;
        const acc_t gi = grad_cols ? a.acc_g[rr * kTileRows + lr] : 0, li = row_loss ? a.acc_l[rr * kTileRows + lr] : 0;
        const double g = (double)gi * kFixScale, l = (double)li * kFixScale;
        // column side
        const int64_t J = q / kSubCols;
        const int half = (int)((q / kHalfCols) & 1);
        const int64_t spJ = (J * (int64_t)a.Pinv) % a.S;
        const int64_t n_groups = ceil_div(a.n_rows, kWarpRowsT);
        double cl = 0.0, cg = 0.0;
        for (int64_t I = 0; I <= Iq; ++I) {
            const unsigned int word = a.cls8[((int64_t)r * a.n_row_tiles + I) * a.S + spJ];
            if (!word_has_double(word)) continue;
            int n_dd = 0;  // rows of tile I whose warp ran this half as a double-duty tile
#pragma unroll
            for (int w = 0; w < kWarpsPerTile; ++w) {
                const int code = (word >> (4 * w + 2 * half)) & 3;
                const int64_t gI = I * kWarpsPerTile + w;
                if (code == kTriDouble && gI < n_groups)
                    n_dd += (int)(min((gI + 1) * kWarpRowsT, a.n_rows) - gI * kWarpRowsT);
            }
            if (n_dd == 0) continue;
            const float2 cs = a.colpart[(int64_t)r * colpart_size(a.n_row_tiles, a.Bpad) + colpart_base(I, a.Bpad) +
                                        (q - I * kTileRows)];
            cl += 2.0 * ((double)n_dd - (double)cs.x);  // sum_i 2 (1 - r_iq)
            cg -= (double)cs.x - (double)cs.y;          // - sum_i (r_iq - r_iq^2)
        }
        my_col_loss = cl;
        const int64_t out = (int64_t)perm[(int64_t)r * a.Bpad + q] * R + r;
        const bool poisoned = row_is_poisoned(a.flags, r, a.Xs[(int64_t)r * a.Bpad + q]);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (grad_cols) grad_cols[out] = poisoned ? (float)nan : (float)((g + cg) * gscale);
        if (row_loss) row_loss[out] = poisoned ? nan : l + cl - pad_per_row;
    }
    my_col_loss = warp_sum(my_col_loss);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = my_col_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sred[w];
        eloss[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
reg_tri_finish_kernel(const acc_t *__restrict__ lossp, int64_t n_lossp, const double *__restrict__ eloss,
                      int64_t n_eloss, double pad_total, double lscale, const int *__restrict__ flags, int R,
                      double *__restrict__ loss_out, float *__restrict__ loss_f32_out) {
    __shared__ double sh[256];
    __shared__ acc_t shl[512];
    acc_t hi, lo;
    sum_loss_partials(lossp, n_lossp, shl, hi, lo);
    double t = threadIdx.x == 0 ? loss_from_hilo(hi, lo) : 0.0;
    for (int64_t u = threadIdx.x; u < n_eloss; u += 256) t += eloss[u];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double total = (sh[0] - pad_total) * lscale;
        if (any_nonfinite(flags, R)) total = __longlong_as_double(0x7ff8000000000000LL);
        *loss_out = total;
        if (loss_f32_out) *loss_f32_out = (float)total;
    }
}

static int tri_ctas_per_sm(bool grad) {
    static int cache[64][2] = {};
    int &v = cache[current_device_slot()][grad ? 1 : 0];
    if (v == 0) {
        int n = 0;
        cudaError_t e = grad ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reg_tri_kernel<true>, kTileThreads, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reg_tri_kernel<false>, kTileThreads, 0);
        if (e != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            n = 1;
        }
        v = n;
    }
    return v;
}

static int mod_inverse(int p, int m) {  // p^-1 mod m for gcd(p, m) = 1
    if (m <= 1) return 0;
    long long t = 0, nt = 1, rr = m, nr = p % m;
    while (nr != 0) {
        const long long qd = rr / nr;
        long long tmp = t - qd * nt; t = nt; nt = tmp;
        tmp = rr - qd * nr; rr = nr; nr = tmp;
    }
    return (int)((t % m + m) % m);
}

// Tail of run_reg_sorted for the triangle mode: plan, pair kernel, epilogue.  `a` is filled by the caller.
static int run_reg_tri_tail(const RegProblem &P, const SortedLayout &L, TilesArgs a, const int *perm,
                            int *combo_cost, char *ws, cudaStream_t st) {
    static_assert(kTileThreads == kSubCols, "the column-sum flush maps one thread to one column");
    const bool want_grad = P.grad_cols_out != nullptr;
    a.colpart = reinterpret_cast<float2 *>(ws + L.off_colpart);
    a.Pinv = mod_inverse(a.P, a.S);
    a.B = P.B;
    int64_t G = (int64_t)sm_count() * tri_ctas_per_sm(want_grad);
    if (G > L.G_max) G = L.G_max;
    if (G < 1) G = 1;
    a.G = (int)G;
    double *eloss = reinterpret_cast<double *>(ws + L.off_eloss);
    const int64_t work = P.B * P.R;
    const int64_t n_eblocks = work > 0 ? ceil_div(work, 256) : 1;

    ARVAE_CUDA_TRY(cudaMemsetAsync(a.acc_g, 0, L.acc_bytes, st));
    if (a.acc_l) ARVAE_CUDA_TRY(cudaMemsetAsync(a.acc_l, 0, L.acc_bytes, st));
    tri_plan_kernel<<<(unsigned)L.n_rr, 256, 0, st>>>(a, combo_cost);
    ARVAE_LAUNCH_CHECK("tri_plan_kernel");
    plan_scan_kernel<<<1, 1024, 0, st>>>(combo_cost, L.n_rr, a.prefix);
    ARVAE_LAUNCH_CHECK("plan_scan_kernel");
    profile_begin(st);
    if (want_grad) reg_tri_kernel<true><<<a.G, kTileThreads, 0, st>>>(a);
    else reg_tri_kernel<false><<<a.G, kTileThreads, 0, st>>>(a);
    profile_end(st);
    ARVAE_LAUNCH_CHECK("reg_tri_kernel");

    double lscale, gscale, pad_per_row;
    reg_scales(P, L.Bpad, lscale, gscale, pad_per_row);
    reg_tri_epilogue_kernel<<<(unsigned)n_eblocks, 256, 0, st>>>(a, perm, P.R, gscale, pad_per_row, P.grad_cols_out,
                                                               P.row_loss_out, eloss);
    ARVAE_LAUNCH_CHECK("reg_tri_epilogue_kernel");
    reg_tri_finish_kernel<<<1, 256, 0, st>>>(a.lossp, a.G, eloss, n_eblocks, pad_per_row * (double)P.B * (double)P.R,
                                             lscale, a.flags, P.R, P.loss_out, P.loss_f32_out);
    ARVAE_LAUNCH_CHECK("reg_tri_finish_kernel");
    return 0;
}
