// pipe_rates.cu -- measures the per-SM issue rates that bound the pair kernels on this B200:
// MUFU (EX2 / RCP), FP32 FMA pipe (FFMA / FADD), ALU pipe (FMNMX / FSET) and a few mixes that
// mirror the pair loops.  SURVEY.md section 8(d) takes the roofline from these rates
// (16 MUFU lanes/clk/SM, 128 FP32 lanes/clk/SM, 4 warp-instructions/clk/SM) "to be confirmed by
// a microbenchmark on the box" -- this is that microbenchmark.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Prints one JSON object; ops are counted per thread-lane ("lanes/clk/SM").
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanha(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2 -- two lanes-worth of work per issued instruction)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// reciprocal of two packed positive normal floats on the FMA / ALU pipes: magic seed + 3 Newton steps, all packed
__device__ __forceinline__ f2_t rcp_newton2(f2_t x) {
    float x0, x1;
    unpack2(x, x0, x1);
    f2_t y = pack2(__int_as_float(0x7EF311C7 - __float_as_int(x0)), __int_as_float(0x7EF311C7 - __float_as_int(x1)));
    const f2_t one = pack2(1.0f, 1.0f), nx = pack2(-x0, -x1);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const f2_t e = fma2(nx, y, one);
        y = fma2(y, e, y);
    }
    return y;
}

constexpr int ILP = 8;
constexpr int ITERS = 4096;

template <int KIND> __global__ void rate_kernel(float *out, float seed, unsigned long long *clk);
enum Kind { K_EX2, K_RCP, K_TANH, K_FFMA, K_FADD, K_FMNMX, K_FSET, K_PAIR_CONST, K_PAIR_CONST1, K_PAIR_GENERAL, K_EX2_RCP, K_REDUX, K_SHFL, K_F2I, K_ATOMS, K_FFMA2, K_FADD2, K_PAIR1_X2, K_PAIR1_X2_NR2, K_PAIR1_X2_NR3, K_COUNT };

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int KIND>
__global__ void __launch_bounds__(256) rate_kernel(float *out, float seed, unsigned long long *clk) {
    __shared__ unsigned int sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = 0;
    __syncthreads();
    const long long c0 = clock64();
    const unsigned long long t0 = gtimer();
    float v[ILP], w[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { v[k] = seed + 0.001f * (threadIdx.x + k); w[k] = 0.f; }
    const float c1 = seed * 0.5f, c2 = seed + 0.25f;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (KIND == K_EX2) v[k] = ex2a(v[k]);
            if (KIND == K_RCP) v[k] = rcpa(v[k]);
            if (KIND == K_TANH) v[k] = tanha(v[k]);
            if (KIND == K_FFMA) v[k] = fmaf(v[k], c1, c2);
            if (KIND == K_FADD) v[k] = v[k] + c1;
            if (KIND == K_FMNMX) v[k] = fminf(v[k], w[k] + 0.f), w[k] = fmaxf(w[k], c1);  // 2 FMNMX (w chain) -- counted as 2
            if (KIND == K_FSET) v[k] = (v[k] > c1 ? 1.0f : 0.0f) + 0.f, w[k] = (w[k] < c2 ? 1.0f : 0.0f);
            if (KIND == K_EX2_RCP) v[k] = rcpa(ex2a(v[k]));
            if (KIND == K_REDUX) v[k] = __uint_as_float(__reduce_add_sync(0xffffffffu, __float_as_uint(v[k])) | 0x3f000000u);
            if (KIND == K_SHFL) v[k] = __shfl_xor_sync(0xffffffffu, v[k], 1) + c1;
            if (KIND == K_F2I) v[k] = (float)__float2uint_rn(v[k]) * c1;
            if (KIND == K_ATOMS) atomicAdd(&sm[(threadIdx.x & 7) + 8 * k], (unsigned)it);
            if (KIND == K_PAIR_CONST) {      // constant-sign tile loop: FADD, EX2, FADD, RCP, FADD, FFMA
                const float r = rcpa(ex2a(c1 - v[k]) + 1.0f);
                v[k] += r;
                w[k] = fmaf(r, r, w[k]);
            }
            if (KIND == K_PAIR_CONST1) {     // 1-MUFU form: FADD, RCP, FMUL, FADD, FFMA
                const float r = c2 * rcpa(c1 + v[k]);
                v[k] += r;
                w[k] = fmaf(r, r, w[k]);
            }
            if (KIND == K_FFMA2) {  // two FMAs per issued instruction
                f2_t a = pack2(v[k], w[k]);
                a = fma2(a, pack2(c1, c1), pack2(c2, c2));
                unpack2(a, v[k], w[k]);
            }
            if (KIND == K_FADD2) {
                f2_t a = pack2(v[k], w[k]);
                a = add2(a, pack2(c1, c2));
                unpack2(a, v[k], w[k]);
            }
            if (KIND == K_PAIR1_X2 || KIND == K_PAIR1_X2_NR2 || KIND == K_PAIR1_X2_NR3) {
                // 1-MUFU constant-sign loop on TWO pairs per slot: FADD2, 2 x MUFU.RCP (or a packed Newton reciprocal
                // for NRn of the 8 slots), FMUL2, FADD2, FFMA2
                const int nr = KIND == K_PAIR1_X2_NR2 ? 2 : (KIND == K_PAIR1_X2_NR3 ? 3 : 0);
                const f2_t ej = pack2(v[k], w[k]);
                const f2_t ssum = add2(pack2(c1, c1), ej);
                f2_t q;
                if (k >= ILP - nr) {
                    q = rcp_newton2(ssum);
                } else {
                    float s0, s1;
                    unpack2(ssum, s0, s1);
                    q = pack2(rcpa(s0), rcpa(s1));
                }
                const f2_t r = mul2(q, ej);
                static_assert(sizeof(f2_t) == 8, "");
                f2_t a1 = pack2(v[k], w[k]);
                a1 = add2(a1, r);
                a1 = fma2(r, r, a1);
                unpack2(a1, v[k], w[k]);
            }
            if (KIND == K_PAIR_GENERAL) {    // general loop: 2 MUFU + 12
                const float r = rcpa(ex2a(c1 - v[k]) + 1.0f);
                const float gt = v[k] > c2 ? 1.0f : 0.0f;
                const float lt = v[k] < c1 ? 1.0f : 0.0f;
                const float kk = (1.0f - gt) + lt;
                const float vv = fmaf(-2.0f, r, kk);
                v[k] += fabsf(vv);
                const float w4 = fmaf(-r, r, r);
                const float sg = fminf(fmaxf(vv * 1099511627776.0f, -1.0f), 1.0f);
                w[k] = fmaf(sg, w4, w[k]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k] + w[k];
    if (s == 123.456f) out[0] = s + sm[threadIdx.x & 63];
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = (unsigned long long)(clock64() - c0); clk[1] = gtimer() - t0; }
}

struct Result { const char *name; double ops_per_iter; double lanes_per_clk_sm; double ms; double ghz; };
static unsigned long long *g_clk;

template <int KIND>
static Result run(const char *name, double ops_per_iter, int sms, double clk_ghz, float *out, int warps_per_sm) {
    const int threads = 256;
    const int blocks = sms * (warps_per_sm * 32 / threads);
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) rate_kernel<KIND><<<blocks, threads>>>(out, 1.0f, g_clk);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CHECK(cudaEventRecord(a));
        rate_kernel<KIND><<<blocks, threads>>>(out, 1.0f, g_clk);
        CHECK(cudaEventRecord(b));
        CHECK(cudaEventSynchronize(b));
        float ms; CHECK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    unsigned long long h[2];
    CHECK(cudaMemcpy(h, g_clk, sizeof(h), cudaMemcpyDeviceToHost));
    const double ghz = h[1] ? (double)h[0] / (double)h[1] : clk_ghz;  // SM clock seen by the last run
    const double total_ops = (double)blocks * threads * ITERS * ILP * ops_per_iter;
    const double clks = best * 1e-3 * ghz * 1e9;
    Result r{name, ops_per_iter, total_ops / clks / sms, best, ghz};
    return r;
}

int main(int argc, char **argv) {
    int dev = 0;
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, dev));
    const int sms = p.multiProcessorCount;
    int clk_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    const double clk_ghz = argc > 1 ? atof(argv[1]) : clk_khz * 1e-6;  // assume max clock unless told
    float *out; CHECK(cudaMalloc(&out, 4));
    CHECK(cudaMalloc(&g_clk, 16));
    printf("{\"device\": \"%s\", \"sms\": %d, \"assumed_clock_ghz\": %.4f, \"note\": \"lanes/clk/SM at the SM clock measured in-kernel (clock64/globaltimer); instruction-level ops per lane\", \"rates\": {", p.name, sms, clk_ghz);
    const int wps[2] = {16, 32};
    bool first = true;
    for (int wi = 0; wi < 2; ++wi) {
        const int w = wps[wi];
        Result rs[] = {
            run<K_EX2>("mufu_ex2", 1, sms, clk_ghz, out, w),
            run<K_RCP>("mufu_rcp", 1, sms, clk_ghz, out, w),
            run<K_TANH>("mufu_tanh", 1, sms, clk_ghz, out, w),
            run<K_EX2_RCP>("mufu_ex2_rcp_chain", 2, sms, clk_ghz, out, w),
            run<K_FFMA>("ffma", 1, sms, clk_ghz, out, w),
            run<K_FADD>("fadd", 1, sms, clk_ghz, out, w),
            run<K_FMNMX>("fmnmx_x2_fadd", 3, sms, clk_ghz, out, w),
            run<K_FSET>("fset_x2_fadd", 3, sms, clk_ghz, out, w),
            run<K_PAIR_CONST>("pair_const_2mufu_4fp32", 1, sms, clk_ghz, out, w),
            run<K_PAIR_CONST1>("pair_const_1mufu_4fp32", 1, sms, clk_ghz, out, w),
            run<K_PAIR_GENERAL>("pair_general_2mufu_12", 1, sms, clk_ghz, out, w),
            run<K_REDUX>("redux_sum_u32(+lop)", 1, sms, clk_ghz, out, w),
            run<K_SHFL>("shfl_xor(+fadd)", 1, sms, clk_ghz, out, w),
            run<K_F2I>("f2i+i2f+fmul", 1, sms, clk_ghz, out, w),
            run<K_ATOMS>("atoms_add_8lanes_distinct", 1, sms, clk_ghz, out, w),
            run<K_FFMA2>("ffma2_fp32_ops", 2, sms, clk_ghz, out, w),
            run<K_FADD2>("fadd2_fp32_ops", 2, sms, clk_ghz, out, w),
            run<K_PAIR1_X2>("pair_const_1mufu_packed_pairs", 2, sms, clk_ghz, out, w),
            run<K_PAIR1_X2_NR2>("pair_const_packed_nr2of8_pairs", 2, sms, clk_ghz, out, w),
            run<K_PAIR1_X2_NR3>("pair_const_packed_nr3of8_pairs", 2, sms, clk_ghz, out, w),
        };
        for (auto &r : rs) {
            printf("%s\"%s@%dw\": {\"per_clk_sm\": %.3f, \"ms\": %.4f, \"sm_ghz\": %.4f}", first ? "" : ", ", r.name, w, r.lanes_per_clk_sm, r.ms, r.ghz);
            first = false;
        }
    }
    printf("}}\n");
    return 0;
}
