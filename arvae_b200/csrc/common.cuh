// common.cuh -- shared helpers for libarvae_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/arvae_b200.h"

namespace arvae {

// ---- error plumbing (thread-local message, see arvae_last_error) -------------------------------
void set_error(const char *fmt, ...);
int fail_cuda(cudaError_t e, const char *what);
void count_launch(int n = 1);
// pair-kernel timing hooks (no-ops unless arvae_profile_enable(1))
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);
// named CUDA-event marks between the launches of a step (no-ops unless arvae_timeline_enable(1)); experiments only
void timeline_mark(cudaStream_t st, const char *name);

// NVTX range around a public entry point (visible in Nsight Systems / Compute timelines; a no-op costing one
// pointer test when no tool is attached -- NVTX v3 is header-only and loads its backend lazily).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define ARVAE_CUDA_TRY(expr)                                        \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ::arvae::fail_cuda(_e, #expr); \
    } while (0)

#define ARVAE_LAUNCH_CHECK(name)                                      \
    do {                                                              \
        ::arvae::count_launch();                                      \
        cudaError_t _e = cudaGetLastError();                          \
        if (_e != cudaSuccess) return ::arvae::fail_cuda(_e, name);   \
    } while (0)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- programmatic dependent launch (sm_90+) ----------------------------------------------------
// A kernel launched with launch_kernel(..., pdl = true) may start while the previous kernel of the stream is still
// running: pdl_wait() blocks until that kernel has completed and its writes are visible (a no-op in a normally
// launched kernel); pdl_trigger() lets the NEXT kernel of the stream start early once every CTA has issued it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                        Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device math ------------------------------------------------------------------------------
// Single-instruction MUFU ops. The .ftz forms compile to one MUFU each (no range fix-up code);
// flushing only affects |2^d| < 2^-126, where tanh is already saturated to +-1 in float.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2): two IEEE float operations per issued instruction, each rounded exactly
// like its scalar form, so packing changes issue slots, not results.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
    f2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
    f2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// sign(a_i - a_j) exactly as torch.sign of the float difference: equals (a_i>a_j)-(a_i<a_j) for
// every pair of floats including NaN, +-inf, +-0 and subnormals. Compiled
// WITHOUT flush-to-zero so distinct subnormals compare unequal.
__device__ __forceinline__ float pair_sign(float ai, float aj) {
    return (ai > aj ? 1.0f : 0.0f) - (ai < aj ? 1.0f : 0.0f);
}

// xs = sgn(f) * x: the pair kernels work on xs so that sgn(tanh(f (x_i - x_j))) = sgn(xs_i - xs_j)
// exactly (f == 0 gives xs = 0: tanh(0) = 0 for every pair, as in the reference).
__device__ __forceinline__ float signed_latent(float x, float fsign) {
    return fsign > 0.0f ? x : (fsign < 0.0f ? -x : 0.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Column padding: (xs = +inf, a = NaN) makes a padded column contribute exactly |t - s| = 1 and
// gradient 0 to every row (2^-inf = 0 -> r = 1 -> t = -1; NaN compares false -> s = 0), so the
// pair loops need no bounds checks and the host subtracts n_pad per row.
#define ARVAE_PAD_U (__int_as_float(0x7f800000))
#define ARVAE_PAD_A (__int_as_float(0x7fc00000))

}  // namespace arvae
