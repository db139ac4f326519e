// cuda_emul.h -- just enough of the CUDA execution model on CPU threads to run a latency-class kernel source
// unchanged in the GPU-less container (TEST INFRASTRUCTURE ONLY; never part of libarvae_b200.so).
//
// One CTA at a time: blockDim.x std::threads walk the grid block by block, __syncthreads() is a std::barrier,
// __shared__ is a static (shared by the threads of the running CTA), warp shuffles go through an exchange array.
// Threads are NOT in lock-step, so a missing __syncthreads() shows up here more readily than on the device.
// Restrictions: every thread of the CTA must reach every __syncthreads() / shuffle (no early return before one).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <thread>
#include <vector>

#define ARVAE_MAX_REG_DIMS 32
#define ARVAE_E_BADARG 1

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void *cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

namespace emul {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline std::barrier<> *g_bar = nullptr;
inline char *g_dyn_smem = nullptr;
inline double g_xchg[1024];
}  // namespace emul
#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)

inline void __syncthreads() { emul::g_bar->arrive_and_wait(); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
using std::max;
using std::min;

// whole-CTA participation required (true for every use in the emulated sources)
inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
    const unsigned t = emul::t_threadIdx.x;
    emul::g_xchg[t] = v;
    __syncthreads();
    const double r = emul::g_xchg[(t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u)];
    __syncthreads();
    return r;
}

namespace arvae {

inline double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

struct RegDims {
    int32_t zcol[ARVAE_MAX_REG_DIMS];
    int32_t lcol[ARVAE_MAX_REG_DIMS];
};

inline int64_t sort_padded_size(int64_t B) {  // as sort.cu
    int64_t n = 256;
    while (n < B) n <<= 1;
    return n;
}

// stand-in for sort.cu's device sort: the same key format and order (value image << 32 | index, NaN after +inf,
// padding last), produced with std::sort
inline int run_sort_keys(const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, int64_t B, int64_t N,
                         unsigned long long *keys, cudaStream_t) {
    for (int r = 0; r < R; ++r) {
        unsigned long long *k = keys + (int64_t)r * N;
        for (int64_t j = 0; j < N; ++j) {
            k[j] = ~0ull;
            if (j < B) {
                const float a = lab[j * lrs + (int64_t)dims.lcol[r] * lcs];
                unsigned int s;
                if (a != a) s = 0xFFFFFFFEu;
                else {
                    const unsigned int b = __float_as_uint(a);
                    s = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                }
                k[j] = ((unsigned long long)s << 32) | (unsigned long long)(unsigned int)j;
            }
        }
        std::sort(k, k + N);
    }
    return 0;
}

constexpr int kEvalMaxCodes = 1024, kEvalMaxAttrs = 64;

}  // namespace arvae

namespace emul {

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F body) {
    g_gridDim = grid;
    g_blockDim = block;
    std::vector<char> dyn(smem + 64);
    g_dyn_smem = dyn.data();
    const int nt = (int)block.x;
    std::barrier<> bar(nt);
    g_bar = &bar;
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) {
        pool.emplace_back([=, &bar]() {
            t_threadIdx = dim3((unsigned)t);
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        t_blockIdx = dim3(bx, by, bz);
                        body();
                        bar.arrive_and_wait();  // CTA boundary: statics are reused by the next block
                    }
        });
    }
    for (auto &th : pool) th.join();
    g_bar = nullptr;
    g_dyn_smem = nullptr;
}

}  // namespace emul

#define ARVAE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define ARVAE_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emul::g_dyn_smem)
#define ARVAE_CUDA_TRY(expr) do { (void)(expr); } while (0)
#define ARVAE_LAUNCH_CHECK(name) do { } while (0)
inline int cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
