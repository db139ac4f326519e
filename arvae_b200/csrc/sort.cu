// sort.cu -- batched bitonic sort of 64-bit keys (one independent sort per regularised dim).
//
// The attribute-sorted pair kernel (reg_sorted.cu) needs, per dim, the permutation that orders the
// batch by attribute value.  Keys are (order-preserving image of the float attribute) << 32 | index,
// so keys are unique, ties come out ordered by original index, NaN sorts after +inf and padding
// after NaN -- the order is a pure function of the inputs (deterministic).  B log^2 B work on
// O(B R) data: noise next to the B^2 R pair sweep, so a simple shared-memory bitonic network is enough.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

namespace cg = cooperative_groups;

#ifndef ARVAE_SORT_CHUNK
#define ARVAE_SORT_CHUNK 2048
#endif
constexpr int kSortChunk = ARVAE_SORT_CHUNK;     // keys sorted per CTA in shared memory: small enough that a
                                                 // 65536-key x 6-dim sort spreads over every SM (192 CTAs)
constexpr int kSortThreads = kSortChunk / 8;     // 8 keys per thread

__device__ __forceinline__ unsigned int float_to_sortable(float a) {
    if (a != a) return 0xFFFFFFFEu;  // every NaN: one class, after +inf (0xFF800000)
    const unsigned int b = __float_as_uint(a);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Key of sample j (local index) of dim r.
//   plain     (spec.z == null):  [sortable attribute : 32][index : 32]
//   segmented (spec.z != null):  [outlier : 1][sortable attribute : 32][index + idx_offset : 31]
// "outlier" = the factorised one-MUFU tanh cannot be used for this element (|2 f log2(e) z| > 62, NaN or inf): such
// elements sort into their own attribute-ordered segment after all inliers, so that the pair kernel can pick the tanh
// form per tile (reg_sorted.cu).  Padding (~0) sorts last in both formats.
__device__ __forceinline__ unsigned long long sort_key_from(const KeySpec &s, float a, float xs, int64_t j) {
    const unsigned long long sa = float_to_sortable(a);
    const float u = s.cabs * xs;
    const unsigned long long out = (!s.segment || fabsf(u) <= kMufu1MaxAbsU) ? 0ull : 1ull;  // NaN / inf: outlier
    return (out << 63) | (sa << 31) | (unsigned long long)(j + s.idx_offset);
}
__device__ __forceinline__ unsigned long long make_sort_key(const KeySpec &s, int r, int64_t j) {
    const float a = __ldg(s.lab + j * s.lrs + (int64_t)s.dims.lcol[r] * s.lcs);
    if (!s.z) return ((unsigned long long)float_to_sortable(a) << 32) | (unsigned long long)(unsigned int)j;
    return sort_key_from(s, a, signed_latent(__ldg(s.z + j * s.zrs + (int64_t)s.dims.zcol[r] * s.zcs), s.fsign), j);
}

__device__ __forceinline__ void cmpx(unsigned long long &a, unsigned long long &b, bool asc) {
    if ((a > b) == asc) {
        const unsigned long long t = a;
        a = b;
        b = t;
    }
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int mask) {
    const unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)v, mask);
    const unsigned int hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(v >> 32), mask);
    return ((unsigned long long)hi << 32) | lo;
}

// Shared-memory round with STEPS (1..3) partner distances j, j/2, .. >= 256: each thread owns the 2^STEPS
// keys that differ only in those bits (lanes stay consecutive in the low index bits: no bank conflicts).
template <int STEPS>
__device__ __forceinline__ void smem_round_strided(unsigned long long *s, int n, int64_t g0, int64_t k, int j) {
    constexpr int E = 1 << STEPS;
    const int jl = j >> (STEPS - 1);
    for (int t = threadIdx.x; t < (n >> STEPS); t += kSortThreads) {
        const int i0 = ((t & ~(jl - 1)) << STEPS) | (t & (jl - 1));
        const bool asc = (((g0 + i0) & k) == 0);
        unsigned long long v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = s[i0 + m * jl];
#pragma unroll
        for (int st = STEPS - 1; st >= 0; --st)
#pragma unroll
            for (int m = 0; m < E; ++m)
                if ((m & (1 << st)) == 0) cmpx(v[m], v[m | (1 << st)], asc);
#pragma unroll
        for (int m = 0; m < E; ++m) s[i0 + m * jl] = v[m];
    }
}

// Tail round: every step with partner distance <= 128 of stages k_first .. k_last, without touching
// shared memory in between.  A thread owns 8 keys at i0 + 32 m (distances 32/64/128 are register pairs),
// a warp owns 256 consecutive keys (distances 16..1 are lane exchanges by shuffle).
__device__ __forceinline__ void smem_round_tail(unsigned long long *s, int n, int64_t g0, int64_t k_first,
                                                int64_t k_last) {
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < (n >> 3); t += kSortThreads) {  // n >= 256: whole warps stay together
        const int i0 = ((t >> 5) << 8) | lane;
        unsigned long long v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = s[i0 + 32 * m];
        for (int64_t k = k_first; k <= k_last; k <<= 1) {
            const int jtop = (int)min((int64_t)128, k >> 1);
            for (int j = jtop; j >= 32; j >>= 1) {
                const int mb = j >> 5;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    if ((m & mb) == 0) {
                        const bool asc = (((g0 + i0 + 32 * m) & k) == 0);
                        // register indices must be compile-time: enumerate the three possible partners
                        if (mb == 1) cmpx(v[m], v[m | 1], asc);
                        else if (mb == 2) cmpx(v[m], v[m | 2], asc);
                        else cmpx(v[m], v[m | 4], asc);
                    }
                }
            }
            for (int j = min(jtop, 16); j >= 1; j >>= 1) {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const unsigned long long o = shfl_xor_u64(v[m], j);
                    const bool asc = (((g0 + i0 + 32 * m) & k) == 0);
                    const bool keep_min = (lower == asc);
                    v[m] = keep_min ? (v[m] < o ? v[m] : o) : (v[m] > o ? v[m] : o);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) s[i0 + 32 * m] = v[m];
    }
}

// Steps of stage k with partner distance from j_top down to 256, in strided rounds of <= 3 steps.
__device__ __forceinline__ void smem_steps_down_to_256(unsigned long long *s, int n, int64_t g0, int64_t k,
                                                       int j_top) {
    int j = j_top;
    while (j >= 256) {
        int steps = 0;
        for (int jj = j; jj >= 256 && steps < 3; jj >>= 1) ++steps;
        if (steps == 3) smem_round_strided<3>(s, n, g0, k, j);
        else if (steps == 2) smem_round_strided<2>(s, n, g0, k, j);
        else smem_round_strided<1>(s, n, g0, k, j);
        __syncthreads();
        j >>= steps;
    }
}

// Sorts each kSortChunk-sized chunk in shared memory: all stages with k <= kSortChunk when
// `k_only` == 0, or only the tail j = kSortChunk/2 .. 1 of stage `k_only` when merging.
// With `build` (first pass, k_only == 0) the keys are built on the fly from the label (and latent) column instead
// of being read back (saves a launch and a round trip through memory).
__global__ void __launch_bounds__(kSortThreads)
bitonic_local_kernel(unsigned long long *__restrict__ keys, int64_t N, int64_t k_only, int build, KeySpec spec,
                     int64_t B) {
    extern __shared__ __align__(16) unsigned long long s[];
    unsigned long long *base = keys + (int64_t)blockIdx.y * N + (int64_t)blockIdx.x * kSortChunk;
    const int64_t g0 = (int64_t)blockIdx.x * kSortChunk;  // global index of s[0] within this dim
    const int n = (int)min((int64_t)kSortChunk, N);       // N is a power of two >= 256
    if (build) {
        for (int i = threadIdx.x; i < n; i += kSortThreads) {
            const int64_t j = g0 + i;
            s[i] = j < B ? make_sort_key(spec, blockIdx.y, j) : ~0ull;  // padding sorts last
        }
    } else {
        for (int i = threadIdx.x; i < n; i += kSortThreads) s[i] = base[i];
    }
    __syncthreads();
    if (k_only == 0) {
        smem_round_tail(s, n, g0, 2, min(256, n));  // stages 2..256 entirely in registers / shuffles
        __syncthreads();
        for (int64_t k = 512; k <= n; k <<= 1) {
            smem_steps_down_to_256(s, n, g0, k, (int)(k >> 1));
            smem_round_tail(s, n, g0, k, k);
            __syncthreads();
        }
    } else {
        smem_steps_down_to_256(s, n, g0, k_only, n >> 1);
        smem_round_tail(s, n, g0, k_only, k_only);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += kSortThreads) base[i] = s[i];
}

// STEPS (1..3) consecutive compare-exchange passes of stage k with partner distances
// j, j/2, .. (all >= kSortChunk) fused in registers: each thread owns the 2^STEPS keys that
// differ only in the bits of those distances.
template <int STEPS>
__global__ void __launch_bounds__(256)
bitonic_global_kernel(unsigned long long *__restrict__ keys, int64_t N, int64_t k, int64_t j) {
    constexpr int E = 1 << STEPS;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (N >> STEPS)) return;
    unsigned long long *base = keys + (int64_t)blockIdx.y * N;
    const int64_t jl = j >> (STEPS - 1);  // smallest distance handled here
    // index with the STEPS bits jl .. j cleared: low part below jl, high part shifted up
    const int64_t i0 = ((t & ~(jl - 1)) << STEPS) | (t & (jl - 1));
    const bool asc = ((i0 & k) == 0);
    unsigned long long v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = base[i0 + (int64_t)m * jl];
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) {  // distance jl << st, i.e. register stride 1 << st
#pragma unroll
        for (int m = 0; m < E; ++m)
            if ((m & (1 << st)) == 0) cmpx(v[m], v[m | (1 << st)], asc);
    }
#pragma unroll
    for (int m = 0; m < E; ++m) base[i0 + (int64_t)m * jl] = v[m];
}

// The whole sort in ONE cooperative launch (grid-wide barriers instead of ~13 kernel boundaries) when every
// (chunk, dim) CTA can be resident at once -- true for the C4 shape (32 chunks x 6 dims = 192 CTAs of 256 threads).
// Same network, same device functions, same result as the multi-launch path below.
template <int STEPS>
__device__ __forceinline__ void coop_global_round(unsigned long long *__restrict__ base, int64_t N, int64_t k, int64_t j,
                                                  int64_t t0, int64_t nthreads) {
    constexpr int E = 1 << STEPS;
    const int64_t jl = j >> (STEPS - 1);
    for (int64_t t = t0; t < (N >> STEPS); t += nthreads) {
        const int64_t i0 = ((t & ~(jl - 1)) << STEPS) | (t & (jl - 1));
        const bool asc = ((i0 & k) == 0);
        unsigned long long v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = base[i0 + (int64_t)m * jl];
#pragma unroll
        for (int st = STEPS - 1; st >= 0; --st)
#pragma unroll
            for (int m = 0; m < E; ++m)
                if ((m & (1 << st)) == 0) cmpx(v[m], v[m | (1 << st)], asc);
#pragma unroll
        for (int m = 0; m < E; ++m) base[i0 + (int64_t)m * jl] = v[m];
    }
}

__global__ void __launch_bounds__(kSortThreads)
bitonic_coop_kernel(unsigned long long *__restrict__ keys, int64_t N, KeySpec spec, int64_t B) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned long long s[];
    unsigned long long *dim_base = keys + (int64_t)blockIdx.y * N;
    unsigned long long *base = dim_base + (int64_t)blockIdx.x * kSortChunk;
    const int64_t g0 = (int64_t)blockIdx.x * kSortChunk;
    const int n = (int)min((int64_t)kSortChunk, N);
    for (int i = threadIdx.x; i < n; i += kSortThreads) {
        const int64_t j = g0 + i;
        s[i] = j < B ? make_sort_key(spec, blockIdx.y, j) : ~0ull;  // padding sorts last
    }
    __syncthreads();
    smem_round_tail(s, n, g0, 2, min(256, n));
    __syncthreads();
    for (int64_t k = 512; k <= n; k <<= 1) {
        smem_steps_down_to_256(s, n, g0, k, (int)(k >> 1));
        smem_round_tail(s, n, g0, k, k);
        __syncthreads();
    }
    const int64_t t0 = (int64_t)blockIdx.x * kSortThreads + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kSortThreads;
    for (int64_t k = 2 * (int64_t)kSortChunk; k <= N; k <<= 1) {
        for (int i = threadIdx.x; i < n; i += kSortThreads) base[i] = s[i];
        grid.sync();  // every chunk of this stage is in global memory
        int64_t j = k >> 1;
        while (j >= kSortChunk) {
            int steps = 0;
            for (int64_t jj = j; jj >= kSortChunk && steps < 3; jj >>= 1) ++steps;
            if (steps == 3) coop_global_round<3>(dim_base, N, k, j, t0, nthreads);
            else if (steps == 2) coop_global_round<2>(dim_base, N, k, j, t0, nthreads);
            else coop_global_round<1>(dim_base, N, k, j, t0, nthreads);
            j >>= steps;
            grid.sync();
        }
        for (int i = threadIdx.x; i < n; i += kSortThreads) s[i] = base[i];
        __syncthreads();
        smem_steps_down_to_256(s, n, g0, k, n >> 1);
        smem_round_tail(s, n, g0, k, k);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += kSortThreads) base[i] = s[i];
}


// ------------------------------------------------------------------------------------------------
// radix run sort: one CTA radix-sorts up to kRunCap keys in shared memory and publishes the sorted run
// (single GPU, large batches: more runs than a wave of 8-CTA clusters holds -- see run_chunk_sort)
// ------------------------------------------------------------------------------------------------
// LSD radix sort over the 33 key bits above the index field (attribute image + outlier bit): three 8-bit passes and
// one 9-bit pass.  The sort is stable and the samples start in index order, so equal attributes come out by
// increasing index -- the same total order as sorting the full 64-bit keys.  Per pass a warp walks its contiguous
// segment 32 elements at a time: MATCH.ANY groups the lanes by digit, the group's lowest lane bumps the warp's
// private counter of that digit, and a lane's rank inside its warp segment falls out of the counter value and its
// position in the group; an exclusive scan of the (digit, warp) counters turns ranks into destinations.  ~100
// instructions per key for the whole sort, against ~1400 for the bitonic network on 64-bit keys.
constexpr int kRadixThreads = 1024;
constexpr int kRadixWarps = kRadixThreads / 32;
constexpr int kRadixMaxBins = 512;
constexpr int kRadixRounds = kRunCap / kRadixThreads;  // 32-element rounds per warp segment at full size
constexpr size_t kChunkSortSmem = 2 * sizeof(unsigned long long) * kRunCap + sizeof(unsigned short) * kRadixMaxBins * kRadixWarps +
                                  sizeof(float) * kRunCap + sizeof(int) * kRadixWarps;
static_assert(kRunCap % kRadixThreads == 0 && kRunCap <= 65535, "segment rounds; 16-bit counters");

template <int BITS>
__device__ __forceinline__ void radix_pass(const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst,
                                           unsigned short *__restrict__ cnt /* [warp][bin] */, int *__restrict__ swarp,
                                           int seg, int shift) {
    constexpr int BINS = 1 << BITS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rounds = seg >> 5;
    for (int i = threadIdx.x; i < BINS * kRadixWarps; i += kRadixThreads) cnt[i] = 0;
    __syncthreads();
    unsigned long long k[kRadixRounds];
    unsigned short pre[kRadixRounds];
    unsigned short *mycnt = cnt + warp * BINS;
#pragma unroll
    for (int q = 0; q < kRadixRounds; ++q) {
        if (q < rounds) {
            k[q] = src[warp * seg + q * 32 + lane];
            const unsigned int d = (unsigned int)(k[q] >> shift) & (unsigned int)(BINS - 1);
            // lanes holding the same digit: one ballot per digit bit (MATCH.ANY measured ~28 cycles per warp
            // instruction per SM here, which made it the bottleneck of the whole sort)
            unsigned int m = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                const bool bit = (d >> b) & 1u;
                const unsigned int bal = __ballot_sync(0xffffffffu, bit);
                m &= bit ? bal : ~bal;
            }
            const int leader = __ffs((int)m) - 1;
            unsigned int c = 0;
            if (lane == leader) c = mycnt[d];
            c = __shfl_sync(0xffffffffu, c, leader);
            pre[q] = (unsigned short)(c + __popc(m & ((1u << lane) - 1u)));
            if (lane == leader) mycnt[d] = (unsigned short)(c + __popc(m));
            __syncwarp();
        }
    }
    __syncthreads();
    // exclusive scan of the counters in (digit, warp) order; thread t owns `per` consecutive warps of one digit
    constexpr int per = BINS * kRadixWarps / kRadixThreads;  // 8 (256 bins) or 16 (512 bins)
    constexpr int groups = kRadixWarps / per;                 // threads per digit
    const int d0 = threadIdx.x / groups, w0 = (threadIdx.x % groups) * per;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < per; ++j) mine += cnt[(w0 + j) * BINS + d0];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) swarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = swarp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        swarp[lane] = wi - w;
    }
    __syncthreads();
    int run = swarp[warp] + incl - mine;
#pragma unroll
    for (int j = 0; j < per; ++j) {
        const int c = cnt[(w0 + j) * BINS + d0];
        cnt[(w0 + j) * BINS + d0] = (unsigned short)run;
        run += c;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kRadixRounds; ++q) {
        if (q < rounds) {
            const unsigned int d = (unsigned int)(k[q] >> shift) & (unsigned int)(BINS - 1);
            dst[(int)mycnt[d] + (int)pre[q]] = k[q];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRadixThreads, 1)
chunk_sort_kernel(KeySpec spec, int64_t n_total, int first_run, RunDest dest, const unsigned long long *__restrict__ epoch_ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *bufA = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *bufB = bufA + kRunCap;
    unsigned short *cnt = reinterpret_cast<unsigned short *>(bufB + kRunCap);
    float *xs_local = reinterpret_cast<float *>(cnt + kRadixMaxBins * kRadixWarps);  // sgn(f) z of the run's samples, by local index
    int *swarp = reinterpret_cast<int *>(xs_local + kRunCap);
    const int r = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * kRunCap;                 // first local sample of this run
    const int n = (int)min((int64_t)kRunCap, n_total - j0);
    if (n <= 0) return;
    const int n_pad = (n + kRadixThreads - 1) / kRadixThreads * kRadixThreads;
    const int seg = n_pad / kRadixWarps;
    {   // keys: every strided load of this thread is issued before the first one is used
        float av[kRadixRounds], zv[kRadixRounds];
#pragma unroll
        for (int q = 0; q < kRadixRounds; ++q) {
            const int i = threadIdx.x + q * kRadixThreads;
            av[q] = zv[q] = 0.0f;
            if (i < n) {
                av[q] = __ldg(spec.lab + (j0 + i) * spec.lrs + (int64_t)spec.dims.lcol[r] * spec.lcs);
                zv[q] = __ldg(spec.z + (j0 + i) * spec.zrs + (int64_t)spec.dims.zcol[r] * spec.zcs);
            }
        }
#pragma unroll
        for (int q = 0; q < kRadixRounds; ++q) {
            const int i = threadIdx.x + q * kRadixThreads;
            if (i < n_pad) {
                unsigned long long key = ~0ull;  // padding sorts last
                if (i < n) {
                    const float xs = signed_latent(zv[q], spec.fsign);
                    xs_local[i] = xs;
                    key = sort_key_from(spec, av[q], xs, j0 + i);
                }
                bufA[i] = key;
            }
        }
    }
    __syncthreads();
    radix_pass<8>(bufA, bufB, cnt, swarp, seg, 31);
    radix_pass<8>(bufB, bufA, cnt, swarp, seg, 39);
    radix_pass<8>(bufA, bufB, cnt, swarp, seg, 47);
    radix_pass<9>(bufB, bufA, cnt, swarp, seg, 55);
    // publish: one 16-byte store per element and destination
    const unsigned int epoch = epoch_ctr ? (unsigned int)(*epoch_ctr + 1ull) : 1u;
    for (int p = threadIdx.x; p < n; p += kRadixThreads) {
        const unsigned long long key = bufA[p];
        const int i = (int)((int64_t)(key & kKeyIdxMask) - spec.idx_offset - j0);
        uint4 e;
        e.x = (unsigned int)key;
        e.y = (unsigned int)(key >> 32);
        e.z = __float_as_uint(xs_local[i]);
        e.w = epoch;
        for (int h = 0; h < dest.n_dest; ++h) {
            uint4 *slot = reinterpret_cast<uint4 *>(run_slot(dest.base[h], dest.R_cap, first_run + (int)blockIdx.x, r));
            slot[p] = e;
            if (p % kPivotStep == 0) slot[kRunCap + p / kPivotStep] = e;  // pivot copy
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cluster sort: a thread-block cluster of 8 CTAs sorts one run of <= kRunCap keys and publishes it
// ------------------------------------------------------------------------------------------------
// Each CTA of the cluster takes a contiguous slice of the run's samples (M keys, M a power of two <= 1024, one key per
// thread), sorts it with a bitonic network (shuffles for partner distances < 32, shared memory above), and then every
// key finds its position in the whole run by rank: its index in its own slice plus the number of smaller keys in the
// other seven slices (copied over distributed shared memory, branch-free binary searches -- keys are unique, so the
// count of smaller keys IS the position).  Elements are then moved, again through distributed shared memory, to the
// CTA that owns their 1024 consecutive run positions, so that the publish to every destination buffer is a coalesced
// stream of 16-byte stores.  ~19 us for a full run (38k cycles: bitonic 16k, copy 3.5k, ranking 8k, publish 4k), against
// ~37 us for a one-CTA radix sort of the same 8192 keys: eight SMs share the strided key loads and each sorts an
// eighth.  (A variant with four keys per thread and 256-thread CTAs -- fewer shuffles and barriers, more register
// work -- was measured slower, 40 us per sort phase against 27: with 8 warps per SM the dependent shared-memory
// probes of the ranking are latency-bound.)
constexpr int kClusterCtas = 8;
constexpr int kSliceCap = kRunCap / kClusterCtas;  // 1024 keys per CTA, one per thread
constexpr int kCsortThreads = kSliceCap;
static_assert(kSliceCap == 1024, "one key per thread of a full CTA");
struct __align__(16) CsortSmem {
    unsigned long long own[kSliceCap];                          // this CTA's sorted slice
    unsigned long long others[kClusterCtas - 1][kSliceCap];     // the other slices' sorted keys
    uint4 out[kSliceCap];                                       // elements of run positions [1024 c, 1024 (c + 1))
    float xs[kSliceCap];                                        // sgn(f) z of this CTA's samples, by slice-local index
};

// number of keys below `key` in a sorted array of m keys, m a power of two (padding ~0 never counts: real keys are smaller)
__device__ __forceinline__ int count_below_pow2(const unsigned long long *__restrict__ arr, int m, unsigned long long key) {
    int lo = 0;
    for (int s = m >> 1; s > 0; s >>= 1) lo += arr[lo + s - 1] < key ? s : 0;
    return lo + (arr[lo] < key ? 1 : 0);
}

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kCsortThreads, 1)
cluster_sort_kernel(KeySpec spec, int64_t n_total, int first_run, RunDest dest, const unsigned long long *__restrict__ epoch_ctr) {
    pdl_trigger();  // a sharded step's ranking kernel may start now: it waits for every run element by itself
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CsortSmem &S = *reinterpret_cast<CsortSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int run = (int)blockIdx.x / kClusterCtas;
    const int r = blockIdx.y;
    const int t = threadIdx.x;
    const int64_t j0 = (int64_t)run * kRunCap;                      // first local sample of this run
    const int n = (int)min((int64_t)kRunCap, n_total - j0);        // >= 1 by construction of the grid
    int M = 32;                                                     // slice length: next power of two of ceil(n / 8)
    while (M * kClusterCtas < n) M <<= 1;
    const int i_local = c * M + t;                                  // this thread's sample within the run
    unsigned long long v = ~0ull;                                   // padding sorts last
    if (t < M && i_local < n) {
        const float a = __ldg(spec.lab + (j0 + i_local) * spec.lrs + (int64_t)spec.dims.lcol[r] * spec.lcs);
        const float xs = signed_latent(__ldg(spec.z + (j0 + i_local) * spec.zrs + (int64_t)spec.dims.zcol[r] * spec.zcs), spec.fsign);
        S.xs[t] = xs;
        v = sort_key_from(spec, a, xs, j0 + i_local);
    }
    // bitonic network over the M keys of the slice (threads >= M only keep the barriers company)
    for (int k = 2; k <= M; k <<= 1) {
        for (int j = k >> 1; j >= 1; j >>= 1) {
            unsigned long long o;
            if (j >= 32) {
                S.own[t] = v;
                __syncthreads();
                o = S.own[t ^ j];
                __syncthreads();
            } else {
                o = shfl_xor_u64(v, j);
            }
            const bool keep_min = ((t & j) == 0) == ((t & k) == 0);
            v = keep_min ? (v < o ? v : o) : (v > o ? v : o);
        }
    }
    S.own[t] = v;  // threads >= M hold padding (their partners, t ^ j with j < M, are >= M too)
    cluster.sync();
    // the other slices, in cluster-rank order with this CTA skipped
#pragma unroll
    for (int q = 0; q < kClusterCtas - 1; ++q) {
        const int oc = q + (q >= c ? 1 : 0);
        const unsigned long long *remote = cluster.map_shared_rank(S.own, oc);
        if (t < M) S.others[q][t] = remote[t];
    }
    __syncthreads();
    const bool real = v != ~0ull;
    if (real) {
        int p = t;  // rank inside the own slice
#pragma unroll
        for (int q = 0; q < kClusterCtas - 1; ++q) p += count_below_pow2(S.others[q], M, v);
        const int i = (int)((int64_t)(v & kKeyIdxMask) - spec.idx_offset - j0) - c * M;  // slice-local index of the sample
        uint4 e;
        e.x = (unsigned int)v;
        e.y = (unsigned int)(v >> 32);
        e.z = __float_as_uint(S.xs[i]);
        e.w = epoch_ctr ? (unsigned int)(*epoch_ctr + 1ull) : 1u;
        uint4 *remote_out = cluster.map_shared_rank(S.out, p / kSliceCap);
        remote_out[p % kSliceCap] = e;
    }
    cluster.sync();
    // publish run positions [1024 c, 1024 (c + 1)): one 16-byte store per element and destination
    const int p = c * kSliceCap + t;
    if (p < n) {
        const uint4 e = S.out[t];
        for (int h = 0; h < dest.n_dest; ++h) {
            uint4 *slot = reinterpret_cast<uint4 *>(run_slot(dest.base[h], dest.R_cap, first_run + run, r));
            slot[p] = e;
            if (p % kPivotStep == 0) slot[kRunCap + p / kPivotStep] = e;  // pivot copy
        }
    }
}

int run_chunk_sort(const KeySpec &spec, int R, int64_t n, int first_run, const RunDest &dest,
                   const unsigned long long *epoch_ctr, cudaStream_t st) {
    if (n <= 0) return 0;
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
        ARVAE_CUDA_TRY(cudaFuncSetAttribute(cluster_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CsortSmem)));
        ARVAE_CUDA_TRY(cudaFuncSetAttribute(chunk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChunkSortSmem));
        attr_set[dev] = true;
    }
    const int64_t runs = ceil_div(n, kRunCap);
    // The 8-CTA clusters finish a run in about half the time of the one-CTA radix sort, as long as all of them are
    // resident at once (one 1024-thread CTA per SM).  With more runs than that -- one GPU sorting a whole large batch --
    // they would queue in waves, and one CTA per run (all resident, one wave) is faster: 44 us against 57-83 us at C4.
    if (runs * kClusterCtas * R > sm_count()) {
        dim3 grid((unsigned)runs, (unsigned)R);
        chunk_sort_kernel<<<grid, kRadixThreads, kChunkSortSmem, st>>>(spec, n, first_run, dest, epoch_ctr);
        ARVAE_LAUNCH_CHECK("chunk_sort_kernel");
        return 0;
    }
    dim3 grid((unsigned)(runs * kClusterCtas), (unsigned)R);
    cluster_sort_kernel<<<grid, kCsortThreads, sizeof(CsortSmem), st>>>(spec, n, first_run, dest, epoch_ctr);
    ARVAE_LAUNCH_CHECK("cluster_sort_kernel");
    return 0;
}

// Can the cooperative single-launch sort be used for this shape on this device, on this stream right now?
static bool coop_sort_possible(int64_t chunks, int R, size_t smem, cudaStream_t st) {
    static int cache[64] = {};  // per device: max co-resident CTAs of bitonic_coop_kernel, -1 = unsupported
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (cache[dev] == 0) {
        int coop = 0, per_sm = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bitonic_coop_kernel, kSortThreads,
                                                                    kSortChunk * sizeof(unsigned long long)) != cudaSuccess) {
            (void)cudaGetLastError();
            cache[dev] = -1;
        } else {
            cache[dev] = per_sm * sms > 0 ? per_sm * sms : -1;
        }
    }
    if (cache[dev] < 0 || chunks * R > cache[dev]) return false;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return false;  // keep graph capture on the plain multi-launch path
    }
    (void)smem;
    return true;
}

int64_t sort_padded_size(int64_t B) {
    int64_t n = 256;
    while (n < B) n <<= 1;
    return n;
}

template <int STEPS>
static void launch_global(unsigned long long *keys, int64_t N, int R, int64_t k, int64_t j, cudaStream_t st) {
    dim3 gg((unsigned)ceil_div(N >> STEPS, 256), (unsigned)R);
    bitonic_global_kernel<STEPS><<<gg, 256, 0, st>>>(keys, N, k, j);
}

// keys[r][0..N) <- sorted (ascending) attribute keys of dim r; N = sort_padded_size(B).
int run_sort_keys(const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, int64_t B,
                  int64_t N, unsigned long long *keys, cudaStream_t st) {
    KeySpec spec;
    spec.lab = lab; spec.lrs = lrs; spec.lcs = lcs;
    spec.z = nullptr; spec.zrs = spec.zcs = 0;
    spec.fsign = 1.0f; spec.cabs = 1.0f; spec.idx_offset = 0; spec.segment = 0;
    spec.dims = dims;
    return run_sort_keys_spec(spec, R, B, N, keys, st);
}

int run_sort_keys_spec(const KeySpec &spec, int R, int64_t B, int64_t N, unsigned long long *keys, cudaStream_t st) {
    static_assert(kSortChunk * sizeof(unsigned long long) <= 48 * 1024, "fits the default dynamic shared memory limit");
    const int64_t chunks = N > kSortChunk ? N / kSortChunk : 1;
    const size_t smem = (size_t)(N < kSortChunk ? N : kSortChunk) * sizeof(unsigned long long);
    dim3 gl((unsigned)chunks, (unsigned)R);
    if (chunks > 1 && coop_sort_possible(chunks, R, smem, st)) {
        KeySpec sp = spec;
        void *args[] = {(void *)&keys, (void *)&N, (void *)&sp, (void *)&B};
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)bitonic_coop_kernel, gl, dim3(kSortThreads), args, smem, st);
        if (e == cudaSuccess) {
            count_launch();
            return 0;
        }
        (void)cudaGetLastError();  // e.g. co-residency not available right now: fall through to the plain path
    }
    bitonic_local_kernel<<<gl, kSortThreads, smem, st>>>(keys, N, 0, 1, spec, B);
    ARVAE_LAUNCH_CHECK("bitonic_local_kernel");
    for (int64_t k = 2 * (int64_t)kSortChunk; k <= N; k <<= 1) {
        int64_t j = k >> 1;
        while (j >= kSortChunk) {
            int steps = 0;
            for (int64_t jj = j; jj >= kSortChunk && steps < 3; jj >>= 1) ++steps;
            if (steps == 3) launch_global<3>(keys, N, R, k, j, st);
            else if (steps == 2) launch_global<2>(keys, N, R, k, j, st);
            else launch_global<1>(keys, N, R, k, j, st);
            ARVAE_LAUNCH_CHECK("bitonic_global_kernel");
            j >>= steps;
        }
        bitonic_local_kernel<<<gl, kSortThreads, smem, st>>>(keys, N, k, 0, spec, B);
        ARVAE_LAUNCH_CHECK("bitonic_local_kernel(merge)");
    }
    return 0;
}

}  // namespace arvae
