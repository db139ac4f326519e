// latent_head.cu -- reparametrize + KL divergence against the unit prior, forward and backward.
//
//   z   = loc + eps * scale                         imagevae/mnist_vae.py:79 (Normal.rsample)
//   kld = sum_b sum_d 0.5 (scale^2 + loc^2 - 1 - log(scale^2))     utils/trainer.py:364-365 via
//         torch.distributions.kl._kl_normal_normal with q = N(0,1)
// Elementwise math is rounded where torch rounds it (separate multiply and add, no FMA
// contraction) so z is bit-identical to the reference's z_tilde; sums are carried in double in a
// fixed order (per-CTA partials, then one CTA), so results are run-to-run reproducible.
#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

constexpr int kHeadThreads = 256;

static int head_blocks(int64_t n) {
    int64_t b = ceil_div(n > 0 ? n : 1, (int64_t)kHeadThreads * 4);
    const int64_t cap = 4LL * sm_count();
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

__global__ void __launch_bounds__(kHeadThreads)
latent_head_fwd_kernel(const float *__restrict__ loc, const float *__restrict__ scale,
                       const float *__restrict__ eps, int64_t n, float *__restrict__ z,
                       double *__restrict__ partial) {
    __shared__ double sred[kHeadThreads / 32];
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const float m = loc[e], s = scale[e];
        z[e] = __fadd_rn(m, __fmul_rn(eps[e], s));
        const float var_ratio = __fmul_rn(s, s);
        const float t1 = __fmul_rn(m, m);
        const float v = 0.5f * (__fadd_rn(var_ratio, t1) - 1.0f - logf(var_ratio));
        acc += (double)v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) t += sred[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kHeadThreads)
latent_head_finish_kernel(const double *__restrict__ partial, int n_partial, int64_t B, float beta,
                          float capacity, double *__restrict__ kld_sum_out,
                          float *__restrict__ kld_mean_out, float *__restrict__ kld_loss_out,
                          float *__restrict__ kcoef_out) {
    __shared__ double sh[kHeadThreads];
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += kHeadThreads) t += partial[i];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = kHeadThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *kld_sum_out = sh[0];
        const float kld = (float)(sh[0] / (double)B);  // .sum(1).mean()
        const float diff = kld - capacity;
        if (kld_mean_out) *kld_mean_out = kld;
        if (kld_loss_out) *kld_loss_out = beta * fabsf(diff);
        if (kcoef_out) *kcoef_out = beta * (float)((diff > 0.0f) - (diff < 0.0f)) / (float)B;
    }
}

__global__ void __launch_bounds__(kHeadThreads)
latent_head_bwd_kernel(const float *__restrict__ loc, const float *__restrict__ scale,
                       const float *__restrict__ eps, const float *__restrict__ dz_up,
                       const float *__restrict__ grad_cols, const float *__restrict__ greg,
                       RegDims dims, int R, float kscale, const float *__restrict__ kcoef,
                       const float *__restrict__ gkld, int64_t B, int64_t Z,
                       float *__restrict__ dloc, float *__restrict__ dscale, int sd_is_log) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * Z) return;
    const int64_t b = e / Z;
    const int d = (int)(e % Z);
    float dz = dz_up ? dz_up[e] : 0.0f;
    if (grad_cols) {
        const float g = greg ? __ldg(greg) : 1.0f;
        float v = 0.0f;
        for (int r = 0; r < R; ++r)
            if (dims.zcol[r] == d) v += grad_cols[b * R + r];
        dz = fmaf(g, v, dz);
    }
    const float k = kscale * (kcoef ? __ldg(kcoef) : 1.0f) * (gkld ? __ldg(gkld) : 1.0f);
    // sd_is_log: `scale` holds log_std and `dscale` receives d/d(log_std) = d/d(scale) * scale
    const float s = sd_is_log ? expf(scale[e]) : scale[e];
    if (dloc) dloc[e] = dz + k * loc[e];
    if (dscale) {
        const float ds = dz * eps[e] + k * (s - 1.0f / s);
        dscale[e] = sd_is_log ? ds * s : ds;
    }
}

size_t latent_head_ws_bytes(int64_t B, int64_t Z) {
    return sizeof(double) * (size_t)head_blocks(B * Z);
}

int run_latent_head_fwd(const float *loc, const float *scale, const float *eps, int64_t B,
                        int64_t Z, float beta, float capacity, float *z_out, double *kld_sum_out,
                        float *kld_mean_out, float *kld_loss_out, float *kcoef_out, void *ws,
                        size_t ws_bytes, cudaStream_t st) {
    const int64_t n = B * Z;
    const int blocks = head_blocks(n);
    if (ws_bytes < sizeof(double) * (size_t)blocks) {
        set_error("latent head workspace too small: %zu < %zu", ws_bytes, sizeof(double) * (size_t)blocks);
        return ARVAE_E_WORKSPACE;
    }
    double *partial = reinterpret_cast<double *>(ws);
    latent_head_fwd_kernel<<<blocks, kHeadThreads, 0, st>>>(loc, scale, eps, n, z_out, partial);
    ARVAE_LAUNCH_CHECK("latent_head_fwd_kernel");
    latent_head_finish_kernel<<<1, kHeadThreads, 0, st>>>(partial, blocks, B, beta, capacity,
                                                          kld_sum_out, kld_mean_out, kld_loss_out,
                                                          kcoef_out);
    ARVAE_LAUNCH_CHECK("latent_head_finish_kernel");
    return 0;
}

int run_latent_head_bwd(const float *loc, const float *scale, const float *eps, const float *dz_up,
                        const float *grad_cols, const float *greg, const RegDims &dims, int R,
                        float kscale, const float *kcoef, const float *gkld, int64_t B, int64_t Z,
                        float *dloc, float *dscale, cudaStream_t st, int sd_is_log) {
    const int64_t n = B * Z;
    if (n <= 0) return 0;
    latent_head_bwd_kernel<<<(unsigned)ceil_div(n, kHeadThreads), kHeadThreads, 0, st>>>(
        loc, scale, eps, dz_up, grad_cols, greg, dims, R, kscale, kcoef, gkld, B, Z, dloc, dscale, sd_is_log);
    ARVAE_LAUNCH_CHECK("latent_head_bwd_kernel");
    return 0;
}

}  // namespace arvae
