// sort.cu -- batched bitonic sort of 64-bit keys (one independent sort per regularised dim).
//
// The attribute-sorted pair kernel (reg_sorted.cu) needs, per dim, the permutation that orders the
// batch by attribute value.  Keys are (order-preserving image of the float attribute) << 32 | index,
// so keys are unique, ties come out ordered by original index, NaN sorts after +inf and padding
// after NaN -- the order is a pure function of the inputs (deterministic).  B log^2 B work on
// O(B R) data: noise next to the B^2 R pair sweep, so a simple shared-memory bitonic network is enough.
#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

constexpr int kSortChunk = 8192;    // keys sorted per CTA in (dynamic) shared memory: 64 KiB
constexpr int kSortThreads = 1024;  // each thread owns kSortChunk / kSortThreads / 2 compare-exchanges per step

__device__ __forceinline__ unsigned int float_to_sortable(float a) {
    if (a != a) return 0xFFFFFFFEu;  // every NaN: one class, after +inf (0xFF800000)
    const unsigned int b = __float_as_uint(a);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256)
make_keys_kernel(const float *__restrict__ lab, int64_t lrs, int64_t lcs, RegDims dims, int64_t B,
                 int64_t N, unsigned long long *__restrict__ keys) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= N) return;
    unsigned long long k = ~0ull;  // padding sorts last
    if (j < B) {
        const float a = __ldg(lab + j * lrs + (int64_t)dims.lcol[r] * lcs);
        k = ((unsigned long long)float_to_sortable(a) << 32) | (unsigned long long)(unsigned int)j;
    }
    keys[(int64_t)r * N + j] = k;
}

__device__ __forceinline__ void cmpx(unsigned long long &a, unsigned long long &b, bool asc) {
    if ((a > b) == asc) {
        const unsigned long long t = a;
        a = b;
        b = t;
    }
}

// Sorts each kSortChunk-sized chunk in shared memory: all stages with k <= kSortChunk when
// `k_only` == 0, or only the tail j = kSortChunk/2 .. 1 of stage `k_only` when merging.
__global__ void __launch_bounds__(kSortThreads)
bitonic_local_kernel(unsigned long long *__restrict__ keys, int64_t N, int64_t k_only) {
    extern __shared__ __align__(16) unsigned long long s[];
    unsigned long long *base = keys + (int64_t)blockIdx.y * N + (int64_t)blockIdx.x * kSortChunk;
    const int64_t g0 = (int64_t)blockIdx.x * kSortChunk;  // global index of s[0] within this dim
    const int n = (int)min((int64_t)kSortChunk, N);       // N is a power of two
    for (int i = threadIdx.x; i < n; i += kSortThreads) s[i] = base[i];
    __syncthreads();
    const int64_t k_begin = k_only ? k_only : 2;
    const int64_t k_end = k_only ? k_only : n;
    for (int64_t k = k_begin; k <= k_end; k <<= 1) {
        int j0 = (int)min((int64_t)(n >> 1), k >> 1);
        for (int j = j0; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += kSortThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
                const bool asc = (((g0 + i) & k) == 0);
                cmpx(s[i], s[i | j], asc);
            }
            __syncthreads();
        }
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
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
        ARVAE_CUDA_TRY(cudaFuncSetAttribute(bitonic_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(kSortChunk * sizeof(unsigned long long))));
        attr_set = true;
    }
    dim3 g1((unsigned)ceil_div(N, 256), (unsigned)R);
    make_keys_kernel<<<g1, 256, 0, st>>>(lab, lrs, lcs, dims, B, N, keys);
    ARVAE_LAUNCH_CHECK("make_keys_kernel");
    const int64_t chunks = N > kSortChunk ? N / kSortChunk : 1;
    const size_t smem = (size_t)(N < kSortChunk ? N : kSortChunk) * sizeof(unsigned long long);
    dim3 gl((unsigned)chunks, (unsigned)R);
    bitonic_local_kernel<<<gl, kSortThreads, smem, st>>>(keys, N, 0);
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
        bitonic_local_kernel<<<gl, kSortThreads, smem, st>>>(keys, N, k);
        ARVAE_LAUNCH_CHECK("bitonic_local_kernel(merge)");
    }
    return 0;
}

}  // namespace arvae
