// head_fused.cu -- the whole latent-loss head of a training step in ONE launch, for the batch sizes the reference
// actually trains at (B = 64 ... 8192).
//
//   scale    = exp(log_std)                                           imagevae/mnist_vae.py:63-65, measurevae/encoder.py:120-123
//   z        = loc + eps * scale                                      mnist_vae.py:79 / measure_vae.py:116 (Normal.rsample)
//   kld_loss = beta |mean_b sum_d 0.5 (scale^2 + loc^2 - 1 - log scale^2) - c|    utils/trainer.py:354-367
//   reg_loss = sum_dim gamma mean_ij |tanh(f (z_i - z_j)) - sign(a_i - a_j)|      utils/trainer.py:369-403 via the trainers' loop
//   and the row sums of the regularisation gradient (what the one-launch backward needs)
//
// At these sizes the pair sweep is microseconds, so the cost is launches: stock PyTorch issues ~20 per dim, the
// separate kernels of this library 7.  Here every CTA of a (row block, column chunk, dim) grid
//   * recomputes the latent of its rows and of the columns it stages from (loc, eps, scale) -- a few hundred flops
//     instead of a launch that materialises z first,
//   * also handles a slice of the elementwise work (z and scale outputs, KL partial),
//   * sweeps its pairs with a general loop: the one-MUFU form on packed FP32 where every sample of the tile is in
//     range (all of delta = 1), the two-MUFU loop of reg_dense.cu otherwise,
//   * adds its row sums into fixed-point accumulators (integer atomics: order-free, bitwise reproducible),
// and "last CTA" tickets do the epilogues in the same launch: the last chunk of a (row block, dim) converts those
// rows' sums into gradient columns, the last CTA of the grid reduces the loss and KL partials in index order.
// The accumulators and tickets live in a small caller-owned workspace that is zeroed once and left zeroed by
// every launch (CUDA-graph friendly: no memset node, no allocation).
#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

constexpr int kHeadFusedThreads = 256;     // one row per thread
constexpr int kHeadFusedMaxBatch = 8192;   // beyond: the attribute-sorted path (separate launches) is the right tool
constexpr double kHfFixMagic = 6291456.0;              // 1.5 * 2^22: (v + magic) keeps round(v 2^30) in the mantissa
constexpr double kHfFixScale = 1.0 / 1073741824.0;     // 2^-30

struct HeadFusedArgs {
    const float *loc, *sd, *eps;   // [B, Z]; sd = scale, or log_std when sd_is_log
    int sd_is_log;
    const float *lab;
    int64_t lrs, lcs;
    RegDims dims;
    int R;
    int64_t B, Z, Bpad;
    int chunk_cols, n_chunks, n_row_blocks;
    float fsign, cabs;             // sgn(factor), |2 factor log2 e|
    float beta, capacity;
    double lscale, gscale, pad_per_row;
    float *z_out, *scale_out;      // [B, Z]; scale_out may be null
    float *kld_mean_out, *kld_loss_out, *kcoef_out, *reg_loss_out;  // [1] each
    float *grad_cols_out;          // [B, R] or null (no gradient wanted)
    // persistent workspace (zero on entry, zero on exit)
    unsigned int *ticket_all, *ticket_rb;   // [1], [n_row_blocks * R]
    long long *acc_g;                       // [B * R] fixed-point row sums
    // scratch (any content)
    double *loss_part, *kld_part;           // [n_cta] each
};

__device__ __forceinline__ float head_scale(const HeadFusedArgs &a, int64_t e) {
    const float s = a.sd[e];
    return a.sd_is_log ? expf(s) : s;
}
// z of one element, rounded where torch rounds it (separate multiply and add): bit-identical to rsample
__device__ __forceinline__ float head_z(const HeadFusedArgs &a, int64_t e) {
    return __fadd_rn(a.loc[e], __fmul_rn(a.eps[e], head_scale(a, e)));
}

template <bool GRAD>
__global__ void __launch_bounds__(kHeadFusedThreads)
head_fused_kernel(HeadFusedArgs a) {
    __shared__ __align__(16) float su[kSubCols];
    __shared__ __align__(16) float se[kSubCols];
    __shared__ __align__(16) float sa[kSubCols];
    __shared__ double sred[2][kHeadFusedThreads / 32];
    __shared__ int s_last[2];
    const int tid = threadIdx.x;
    const int rb = blockIdx.x, chunk = blockIdx.y, r = blockIdx.z;
    const int64_t n_cta = (int64_t)gridDim.x * gridDim.y * gridDim.z;
    const int64_t cta = ((int64_t)r * gridDim.y + chunk) * gridDim.x + rb;
    const int zc = a.dims.zcol[r], lc = a.dims.lcol[r];

    // ---- elementwise slice: z, scale, KL partial ----------------------------------------------------------
    double kacc = 0.0;
    for (int64_t e = cta * kHeadFusedThreads + tid; e < a.B * a.Z; e += n_cta * kHeadFusedThreads) {
        const float m = a.loc[e], s = head_scale(a, e);
        a.z_out[e] = __fadd_rn(m, __fmul_rn(a.eps[e], s));
        if (a.scale_out) a.scale_out[e] = s;
        const float var_ratio = __fmul_rn(s, s);
        const float t1 = __fmul_rn(m, m);
        kacc += (double)(0.5f * (__fadd_rn(var_ratio, t1) - 1.0f - logf(var_ratio)));
    }

    // ---- pairs: this CTA's 256 rows x its column chunk, dim r ------------------------------------------------
    const int64_t row = (int64_t)rb * kHeadFusedThreads + tid;
    const bool valid = row < a.B;
    const float xi = valid ? signed_latent(head_z(a, row * a.Z + zc), a.fsign) : 0.0f;
    const float ai = valid ? __ldg(a.lab + row * a.lrs + (int64_t)lc * a.lcs) : 0.0f;
    // One-MUFU form r = E_j / (E_i + E_j), E = 2^(c x) (reg_sorted.cu), usable while |c x| <= 62 for every row and staged
    // column of a tile: decided per tile by a block-wide vote (delta = 1 configs always pass; at delta = 10 a tile with
    // an out-of-range sample runs the two-MUFU loop below).
    const float ui = a.cabs * xi;
    const bool row_ok = !valid || fabsf(ui) <= kMufu1MaxAbsU;
    const float ei = exp2f(row_ok ? ui : 0.0f);
    double dl = 0.0;
    float gsum = 0.0f;
    const int64_t c0 = (int64_t)chunk * a.chunk_cols;
    const int64_t c1 = min(c0 + (int64_t)a.chunk_cols, a.Bpad);
    for (int64_t t0 = c0; t0 < c1; t0 += kSubCols) {
        __syncthreads();
        bool col_ok = true;
        {
            const int64_t j = t0 + tid;  // kSubCols == kHeadFusedThreads: one column per thread
            const float xj = j < a.B ? signed_latent(head_z(a, j * a.Z + zc), a.fsign) : ARVAE_PAD_U;
            const float uj = a.cabs * xj;
            col_ok = j >= a.B || fabsf(uj) <= kMufu1MaxAbsU;
            su[tid] = xj;
            se[tid] = j < a.B ? exp2f(col_ok ? uj : 0.0f) : 8.5070592e37f;  // padding: 2^126 -> r = 1 exactly, as in reg_sorted.cu
            sa[tid] = j < a.B ? __ldg(a.lab + j * a.lrs + (int64_t)lc * a.lcs) : ARVAE_PAD_A;
        }
        const bool mufu1 = __syncthreads_and(row_ok && col_ok) != 0;
        float lacc = 0.0f, gacc = 0.0f;
        if (mufu1) {
            // packed FP32 on column pairs: per two pairs 11 packed instructions, two MUFU.RCP, 4 FSET, 4 FMNMX.
            // |v| is accumulated as v sgn(v); sgn(v) is the exact sign built from the attribute compares and d.
            f2_t lacc2 = pack2(0.0f, 0.0f), gacc2 = pack2(0.0f, 0.0f);
            const f2_t neg1 = pack2(-1.0f, -1.0f), one = pack2(1.0f, 1.0f), neg2 = pack2(-2.0f, -2.0f);
            const f2_t bigs = pack2(1.7014118e38f, 1.7014118e38f), bigd = pack2(1.1529215e18f, 1.1529215e18f);
            const f2_t xi2 = pack2(xi, xi), ei2 = pack2(ei, ei);
#pragma unroll 4
            for (int q = 0; q < kSubCols; q += 4) {
                const float4 xj = *reinterpret_cast<const float4 *>(su + q);
                const float4 ej = *reinterpret_cast<const float4 *>(se + q);
                const float4 aj = *reinterpret_cast<const float4 *>(sa + q);
                const float aa[4] = {aj.x, aj.y, aj.z, aj.w};
                const f2_t xv[2] = {pack2(xj.x, xj.y), pack2(xj.z, xj.w)};
                const f2_t ev[2] = {pack2(ej.x, ej.y), pack2(ej.z, ej.w)};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const f2_t d = fma2(xv[h], neg1, xi2);  // xs_i - xs_j, exact
                    float s0, s1, q0, q1;
                    unpack2(add2(ei2, ev[h]), s0, s1);
                    const f2_t r = mul2(pack2(rcp_approx(s0), rcp_approx(s1)), ev[h]);
                    const f2_t gt = pack2(ai > aa[2 * h] ? 1.0f : 0.0f, ai > aa[2 * h + 1] ? 1.0f : 0.0f);
                    const f2_t lt = pack2(ai < aa[2 * h] ? 1.0f : 0.0f, ai < aa[2 * h + 1] ? 1.0f : 0.0f);
                    const f2_t km1 = fma2(gt, neg1, lt);           // -s
                    const f2_t v = fma2(neg2, r, add2(km1, one));  // t - s = (1 - s) - 2 r
                    unpack2(fma2(km1, bigs, mul2(d, bigd)), q0, q1);
                    const f2_t sg = pack2(fminf(fmaxf(q0, -1.0f), 1.0f), fminf(fmaxf(q1, -1.0f), 1.0f));
                    lacc2 = fma2(v, sg, lacc2);
                    if (GRAD) gacc2 = fma2(sg, mul2(r, fma2(r, neg1, one)), gacc2);  // sgn(v) (r - r^2)
                }
            }
            float l0, l1, g0, g1;
            unpack2(lacc2, l0, l1);
            unpack2(gacc2, g0, g1);
            lacc = l0 + l1;
            gacc = g0 + g1;
        } else {
#pragma unroll 4
            for (int q = 0; q < kSubCols; q += 4) {
                const float4 uj = *reinterpret_cast<const float4 *>(su + q);
                const float4 aj = *reinterpret_cast<const float4 *>(sa + q);
                pair_general<GRAD>(xi, ai, uj.x, aj.x, a.cabs, lacc, gacc);
                pair_general<GRAD>(xi, ai, uj.y, aj.y, a.cabs, lacc, gacc);
                pair_general<GRAD>(xi, ai, uj.z, aj.z, a.cabs, lacc, gacc);
                pair_general<GRAD>(xi, ai, uj.w, aj.w, a.cabs, lacc, gacc);
            }
        }
        dl += (double)lacc;
        gsum += gacc;  // <= 8192 columns of |g| <= 1/4: float is ample before the fixed-point conversion below
    }
    if (!valid) dl = 0.0;
    if (GRAD && valid) {
        const long long q = __double_as_longlong((double)gsum + kHfFixMagic) - __double_as_longlong(kHfFixMagic);
        if (q != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.acc_g + row * a.R + r), (unsigned long long)q);
    }

    // ---- per-CTA partials ---------------------------------------------------------------------------------
    dl = warp_sum(dl);
    kacc = warp_sum(kacc);
    if ((tid & 31) == 0) { sred[0][tid >> 5] = dl; sred[1][tid >> 5] = kacc; }
    __threadfence();  // this thread's row-sum atomic is visible before the tickets below
    __syncthreads();
    if (tid == 0) {
        double tl = 0.0, tk = 0.0;
#pragma unroll
        for (int w = 0; w < kHeadFusedThreads / 32; ++w) { tl += sred[0][w]; tk += sred[1][w]; }
        a.loss_part[cta] = tl;
        a.kld_part[cta] = tk;
        __threadfence();
        s_last[0] = GRAD ? (atomicAdd(a.ticket_rb + (int64_t)rb * a.R + r, 1u) == (unsigned int)a.n_chunks - 1u) : 0;
        s_last[1] = atomicAdd(a.ticket_all, 1u) == (unsigned int)n_cta - 1u;
    }
    __syncthreads();

    // ---- last chunk of this (row block, dim): its rows' gradient columns; leave the accumulators zeroed ----
    if (GRAD && s_last[0]) {
        __threadfence();
        if (valid) {
            long long *p = a.acc_g + row * a.R + r;
            const long long g = __ldcg(p);
            *p = 0;
            a.grad_cols_out[row * a.R + r] = (float)((double)g * kHfFixScale * a.gscale);
        }
        if (tid == 0) a.ticket_rb[(int64_t)rb * a.R + r] = 0u;
    }
    // ---- last CTA of the grid: loss and KL, partials summed in index order -----------------------------------
    if (s_last[1]) {
        __threadfence();
        __shared__ double sh[2][kHeadFusedThreads];
        double tl = 0.0, tk = 0.0;
        for (int64_t u = tid; u < n_cta; u += kHeadFusedThreads) { tl += __ldcg(a.loss_part + u); tk += __ldcg(a.kld_part + u); }
        sh[0][tid] = tl;
        sh[1][tid] = tk;
        __syncthreads();
        for (int o = kHeadFusedThreads / 2; o > 0; o >>= 1) {
            if (tid < o) { sh[0][tid] += sh[0][tid + o]; sh[1][tid] += sh[1][tid + o]; }
            __syncthreads();
        }
        if (tid == 0) {
            const double total = sh[0][0] - a.pad_per_row * (double)a.B * (double)a.R;
            *a.reg_loss_out = (float)(total * a.lscale);
            const float kld = (float)(sh[1][0] / (double)a.B);  // .sum(1).mean()
            const float diff = kld - a.capacity;
            if (a.kld_mean_out) *a.kld_mean_out = kld;
            if (a.kld_loss_out) *a.kld_loss_out = a.beta * fabsf(diff);
            if (a.kcoef_out) *a.kcoef_out = a.beta * (float)((diff > 0.0f) - (diff < 0.0f)) / (float)a.B;
            *a.ticket_all = 0u;
        }
    }
}

struct HeadFusedLayout {
    int64_t Bpad;
    int n_row_blocks, n_chunks, chunk_cols;
    int64_t n_cta;
    size_t off_ticket_all, off_ticket_rb, off_acc, persistent_bytes, off_loss_part, off_kld_part, bytes;
};

static HeadFusedLayout head_fused_layout(int64_t B, int R) {
    HeadFusedLayout L;
    L.Bpad = round_up(B > 0 ? B : 1, kSubCols);
    L.n_row_blocks = (int)ceil_div(B > 0 ? B : 1, kHeadFusedThreads);
    // enough CTAs to fill the device a few times over, never less than one sub-chunk of columns per CTA
    const int64_t row_units = (int64_t)L.n_row_blocks * R;
    int64_t want = ceil_div(4LL * sm_count(), row_units);
    const int64_t max_chunks = L.Bpad / kSubCols;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    L.chunk_cols = (int)round_up(ceil_div(L.Bpad, want), kSubCols);
    L.n_chunks = (int)ceil_div(L.Bpad, L.chunk_cols);
    L.n_cta = (int64_t)L.n_row_blocks * L.n_chunks * R;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    L.off_ticket_all = take(sizeof(unsigned int));
    L.off_ticket_rb = take(sizeof(unsigned int) * (size_t)L.n_row_blocks * R);
    L.off_acc = take(sizeof(long long) * (size_t)(B > 0 ? B : 1) * R);
    L.persistent_bytes = off;
    L.off_loss_part = take(sizeof(double) * (size_t)L.n_cta);
    L.off_kld_part = take(sizeof(double) * (size_t)L.n_cta);
    L.bytes = off;
    return L;
}

size_t head_fused_ws_bytes(int64_t B, int R) {
    if (B < 1 || B > kHeadFusedMaxBatch || R < 1 || R > ARVAE_MAX_REG_DIMS) return 0;
    return head_fused_layout(B, R).bytes;
}

int run_head_fused_fwd(const float *loc, const float *sd, int sd_is_log, const float *eps, int64_t B, int64_t Z,
                       const float *lab, int64_t lrs, int64_t lcs, const RegDims &dims, int R, float beta, float capacity,
                       float gamma, float factor, float *z_out, float *scale_out, float *kld_mean_out, float *kld_loss_out,
                       float *kcoef_out, float *reg_loss_out, float *grad_cols_out, char *ws, size_t ws_bytes, cudaStream_t st) {
    if (B < 1 || B > kHeadFusedMaxBatch) {
        set_error("head_fused: 1 <= B <= %d (got %lld); larger batches use arvae_latent_head_fwd_f32 + arvae_reg_loss_fwdbwd_f32",
                  kHeadFusedMaxBatch, (long long)B);
        return ARVAE_E_BADARG;
    }
    const HeadFusedLayout L = head_fused_layout(B, R);
    if (ws_bytes < L.bytes) {
        set_error("head_fused: workspace too small: %zu < %zu", ws_bytes, L.bytes);
        return ARVAE_E_WORKSPACE;
    }
    HeadFusedArgs a;
    a.loc = loc; a.sd = sd; a.eps = eps; a.sd_is_log = sd_is_log;
    a.lab = lab; a.lrs = lrs; a.lcs = lcs;
    a.dims = dims; a.R = R; a.B = B; a.Z = Z; a.Bpad = L.Bpad;
    a.chunk_cols = L.chunk_cols; a.n_chunks = L.n_chunks; a.n_row_blocks = L.n_row_blocks;
    const double c = 2.0 * (double)factor * 1.4426950408889634074;  // 2 f log2(e)
    a.fsign = factor > 0.f ? 1.0f : (factor < 0.f ? -1.0f : 0.0f);
    a.cabs = factor != 0.f ? (float)fabs(c) : 1.0f;
    a.beta = beta; a.capacity = capacity;
    const double BB = (double)B * (double)B;
    a.lscale = (double)gamma / BB;
    a.gscale = 8.0 * (double)gamma * (double)factor / BB;
    a.pad_per_row = (double)(L.Bpad - B);
    a.z_out = z_out; a.scale_out = scale_out;
    a.kld_mean_out = kld_mean_out; a.kld_loss_out = kld_loss_out; a.kcoef_out = kcoef_out; a.reg_loss_out = reg_loss_out;
    a.grad_cols_out = grad_cols_out;
    a.ticket_all = reinterpret_cast<unsigned int *>(ws + L.off_ticket_all);
    a.ticket_rb = reinterpret_cast<unsigned int *>(ws + L.off_ticket_rb);
    a.acc_g = reinterpret_cast<long long *>(ws + L.off_acc);
    a.loss_part = reinterpret_cast<double *>(ws + L.off_loss_part);
    a.kld_part = reinterpret_cast<double *>(ws + L.off_kld_part);
    dim3 grid((unsigned)L.n_row_blocks, (unsigned)L.n_chunks, (unsigned)R);
    if (grad_cols_out) head_fused_kernel<true><<<grid, kHeadFusedThreads, 0, st>>>(a);
    else head_fused_kernel<false><<<grid, kHeadFusedThreads, 0, st>>>(a);
    ARVAE_LAUNCH_CHECK("head_fused_kernel");
    return 0;
}

}  // namespace arvae
