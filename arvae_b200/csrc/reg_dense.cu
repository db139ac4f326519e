// reg_dense.cu -- the general pair loop of the attribute-regularization loss (sm_100a).
//
// Computes, for a block of rows i and ALL columns j, per regularised dim r
//     L_ij = | tanh(f (x_i - x_j)) - sign(a_i - a_j) |            (reference utils/trainer.py:390-401)
//     g_ij = sgn(tanh(.) - sign(.)) * (1 - tanh(.)^2)             (its autograd backward, trainer.py:140)
// and the row sums  sum_j L_ij  and  sum_j g_ij.  Because L_ij = L_ji and g_ji = -g_ij, the full
// gradient of the mean loss w.r.t. x_i is 2 f / B^2 * sum_j g_ij -- a row sum, no column traffic.
//
// Nothing pair-sized ever touches memory: columns are staged through shared memory as packed
// (u_j, a_j) vectors (u = 2 f log2(e) x, so tanh needs one EX2 and one RCP), each thread keeps RI
// rows in registers and sweeps the columns with broadcast LDS.128 reads.  The kernel is bound by
// MUFU (2 per pair) and FP32 issue, not by HBM.
#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

// ------------------------------------------------------------------------------------------------
// pack: strided z / labels  ->  dense per-dim column vectors X[r][Bpad] = sgn(f) z, A[r][Bpad]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_columns_kernel(const float *__restrict__ z, int64_t zrs, int64_t zcs,
                    const float *__restrict__ lab, int64_t lrs, int64_t lcs, RegDims dims, int R,
                    int64_t B, int64_t Bpad, float fsign, float *__restrict__ U,
                    float *__restrict__ A) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bpad) return;
    if (j < B) {
        for (int r = 0; r < R; ++r) {
            U[(int64_t)r * Bpad + j] = signed_latent(__ldg(z + j * zrs + (int64_t)dims.zcol[r] * zcs), fsign);
            A[(int64_t)r * Bpad + j] = __ldg(lab + j * lrs + (int64_t)dims.lcol[r] * lcs);
        }
    } else {
        for (int r = 0; r < R; ++r) {
            U[(int64_t)r * Bpad + j] = ARVAE_PAD_U;
            A[(int64_t)r * Bpad + j] = ARVAE_PAD_A;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the pair loop
// ------------------------------------------------------------------------------------------------
// (the per-pair arithmetic, pair_general, lives in reg_internal.cuh: head_fused.cu sweeps with it too)
template <int RI, bool GRAD, bool SIGNS>
__global__ void __launch_bounds__(kDenseThreads)
reg_dense_kernel(const float *__restrict__ U, const float *__restrict__ A, float cabs, int64_t Bpad,
                 int64_t row_begin, int64_t row_end, int64_t rows_pad, int64_t chunk_cols,
                 double *__restrict__ pgrad, double *__restrict__ prow,
                 double *__restrict__ lossp, int *__restrict__ psign) {
    __shared__ __align__(16) float su[kDenseTileCols];
    __shared__ __align__(16) float sa[kDenseTileCols];
    __shared__ double sred[kDenseThreads / 32];

    const int r = blockIdx.z;
    const int chunk = blockIdx.y;
    const int R = gridDim.z;
    const int64_t rb0 = (int64_t)blockIdx.x * (kDenseThreads * RI);
    const float *Ur = U + (int64_t)r * Bpad;
    const float *Ar = A + (int64_t)r * Bpad;

    float ui[RI], ai[RI];
    bool valid[RI];
    double dl[RI], dg[RI];
    int ds[RI];  // sum_j sign(a_i - a_j) from the very compares the loss uses (parity instrumentation)
#pragma unroll
    for (int k = 0; k < RI; ++k) {
        const int64_t row = row_begin + rb0 + threadIdx.x + (int64_t)k * kDenseThreads;
        valid[k] = row < row_end;
        ui[k] = valid[k] ? Ur[row] : 0.0f;
        ai[k] = valid[k] ? Ar[row] : 0.0f;
        dl[k] = 0.0;
        dg[k] = 0.0;
        ds[k] = 0;
    }

    const int64_t c0 = (int64_t)chunk * chunk_cols;
    const int64_t c1 = min(c0 + chunk_cols, Bpad);
    for (int64_t t0 = c0; t0 < c1; t0 += kDenseTileCols) {
        const int n = (int)min((int64_t)kDenseTileCols, c1 - t0);  // multiple of kSubCols
        __syncthreads();
        for (int q = threadIdx.x * 4; q < n; q += kDenseThreads * 4) {
            *reinterpret_cast<float4 *>(su + q) = *reinterpret_cast<const float4 *>(Ur + t0 + q);
            *reinterpret_cast<float4 *>(sa + q) = *reinterpret_cast<const float4 *>(Ar + t0 + q);
        }
        __syncthreads();
        for (int s0 = 0; s0 < n; s0 += kSubCols) {
            float lacc[RI], gacc[RI], kacc[RI];
#pragma unroll
            for (int k = 0; k < RI; ++k) lacc[k] = gacc[k] = kacc[k] = 0.0f;
#pragma unroll 2
            for (int q = 0; q < kSubCols; q += 4) {
                const float4 uj = *reinterpret_cast<const float4 *>(su + s0 + q);
                const float4 aj = *reinterpret_cast<const float4 *>(sa + s0 + q);
#pragma unroll
                for (int k = 0; k < RI; ++k) {
                    pair_general<GRAD, SIGNS>(ui[k], ai[k], uj.x, aj.x, cabs, lacc[k], gacc[k], &kacc[k]);
                    pair_general<GRAD, SIGNS>(ui[k], ai[k], uj.y, aj.y, cabs, lacc[k], gacc[k], &kacc[k]);
                    pair_general<GRAD, SIGNS>(ui[k], ai[k], uj.z, aj.z, cabs, lacc[k], gacc[k], &kacc[k]);
                    pair_general<GRAD, SIGNS>(ui[k], ai[k], uj.w, aj.w, cabs, lacc[k], gacc[k], &kacc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < RI; ++k) {
                dl[k] += (double)lacc[k];
                dg[k] += (double)gacc[k];
                if (SIGNS) ds[k] += kSubCols - (int)kacc[k];  // sum s = n - sum (1 - s); padded columns have s = 0
            }
        }
    }

    // per-unit outputs: row partials (deterministic slot per unit) and one loss partial
    double lsum = 0.0;
    const int64_t slot = ((int64_t)chunk * R + r) * rows_pad + rb0;
#pragma unroll
    for (int k = 0; k < RI; ++k) {
        const int64_t o = slot + threadIdx.x + (int64_t)k * kDenseThreads;
        if (!valid[k]) dl[k] = 0.0;
        lsum += dl[k];
        if (GRAD) pgrad[o] = valid[k] ? dg[k] : 0.0;
        if (prow) prow[o] = dl[k];
        if (SIGNS) psign[o] = valid[k] ? ds[k] : 0;
    }
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kDenseThreads / 32; ++w) t += sred[w];
        lossp[((int64_t)chunk * R + r) * gridDim.x + blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// epilogue: fixed-order reduction of the per-unit partials
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reg_epilogue_kernel(const double *__restrict__ pgrad, const double *__restrict__ prow,
                    const double *__restrict__ lossp, int n_chunks, int R, int64_t n_rows,
                    int64_t rows_pad, int64_t n_units, double gscale, double lscale,
                    double pad_per_row, float *__restrict__ grad_cols,
                    double *__restrict__ row_loss, double *__restrict__ loss_out,
                    float *__restrict__ loss_f32_out, const int *__restrict__ psign, int *__restrict__ row_sign) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n_rows * R
    if (idx < n_rows * R) {
        const int64_t row = idx / R;
        const int r = (int)(idx % R);
        if (grad_cols) {
            double g = 0.0;
            for (int c = 0; c < n_chunks; ++c) g += pgrad[((int64_t)c * R + r) * rows_pad + row];
            grad_cols[idx] = (float)(g * gscale);
        }
        if (row_loss) {
            double l = 0.0;
            for (int c = 0; c < n_chunks; ++c) l += prow[((int64_t)c * R + r) * rows_pad + row];
            row_loss[idx] = l - pad_per_row;
        }
        if (row_sign) {
            int t = 0;
            for (int c = 0; c < n_chunks; ++c) t += psign[((int64_t)c * R + r) * rows_pad + row];
            row_sign[idx] = t;
        }
    }
    if (blockIdx.x == 0) {
        __shared__ double sh[256];
        double t = 0.0;
        for (int64_t u = threadIdx.x; u < n_units; u += 256) t += lossp[u];
        sh[threadIdx.x] = t;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double total = sh[0] - pad_per_row * (double)n_rows * (double)R;
            *loss_out = total * lscale;  // B == 0: 0 * inf = NaN, like the reference's empty mean
            if (loss_f32_out) *loss_f32_out = (float)(total * lscale);
        }
    }
}

// grad_z[k, zc] = go * sum_{r: d_r == zc} grad_cols[k, r]
__global__ void __launch_bounds__(256)
reg_scatter_bwd_kernel(const float *__restrict__ grad_cols, const float *__restrict__ grad_out,
                       RegDims dims, int R, int64_t n_rows, int64_t Z, float *__restrict__ grad_z,
                       int64_t gzrs) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * Z) return;
    const int64_t row = idx / Z;
    const int zc = (int)(idx % Z);
    const float go = grad_out ? __ldg(grad_out) : 1.0f;
    float v = 0.0f;
    for (int r = 0; r < R; ++r)
        if (dims.zcol[r] == zc) v += grad_cols[row * R + r];
    grad_z[row * gzrs + zc] = go * v;
}

__global__ void __launch_bounds__(256)
pack_slice_kernel(const float *__restrict__ z, int64_t zrs, int64_t zcs, const float *__restrict__ lab,
                  int64_t lrs, int64_t lcs, RegDims dims, int R, int64_t n_rows,
                  float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n_rows * 2R
    if (idx >= n_rows * 2 * R) return;
    const int64_t k = idx / (2 * R);
    const int c = (int)(idx % (2 * R));
    out[idx] = c < R ? __ldg(z + k * zrs + (int64_t)dims.zcol[c] * zcs)
                     : __ldg(lab + k * lrs + (int64_t)dims.lcol[c - R] * lcs);
}

__global__ void __launch_bounds__(256)
extract_perm_kernel(const unsigned long long *__restrict__ keys, int64_t B, int32_t *__restrict__ perm) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < B) perm[k] = (int32_t)(keys[k] & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(256)
sign_matrix_kernel(const float *__restrict__ a, int64_t stride, int64_t B,
                   int8_t *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * B) return;
    const int64_t i = idx / B, j = idx % B;
    out[idx] = (int8_t)pair_sign(a[i * stride], a[j * stride]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
DenseLayout dense_layout(int64_t B_total, int64_t n_rows, int R, int sm_count) {
    DenseLayout L;
    L.Bpad = round_up(B_total > 0 ? B_total : 1, kSubCols);
    // few rows: one row per thread so that more CTAs exist; many rows: 4 rows per thread
    L.RI = (n_rows * (int64_t)R >= (int64_t)sm_count * kDenseThreads * 8) ? 4 : 1;
    const int64_t TR = (int64_t)kDenseThreads * L.RI;
    L.n_row_blocks = n_rows > 0 ? ceil_div(n_rows, TR) : 0;
    L.rows_pad = L.n_row_blocks * TR;
    const int64_t target_units = 40LL * sm_count;
    const int64_t row_units = L.n_row_blocks * R > 0 ? L.n_row_blocks * R : 1;
    int64_t want = ceil_div(target_units, row_units);
    const int64_t max_chunks = L.Bpad / kSubCols;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    L.chunk_cols = round_up(ceil_div(L.Bpad, want), kSubCols);
    L.n_chunks = (int)ceil_div(L.Bpad, L.chunk_cols);
    L.n_units = (int64_t)L.n_chunks * R * L.n_row_blocks;
    // workspace carve-up (all offsets 256-byte aligned)
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    L.off_U = take(sizeof(float) * (size_t)R * L.Bpad);
    L.off_A = take(sizeof(float) * (size_t)R * L.Bpad);
    L.off_pgrad = take(sizeof(double) * (size_t)L.n_chunks * R * L.rows_pad);
    L.off_prow = take(sizeof(double) * (size_t)L.n_chunks * R * L.rows_pad);
    L.off_lossp = take(sizeof(double) * (size_t)(L.n_units > 0 ? L.n_units : 1));
    L.off_psign = take(sizeof(int) * (size_t)L.n_chunks * R * L.rows_pad);
    L.bytes = off;
    return L;
}

template <int RI>
static void launch_dense(const DenseLayout &L, int R, bool want_grad, const float *U,
                         const float *A, float cabs, int64_t row_begin, int64_t row_end, double *pgrad,
                         double *prow, double *lossp, int *psign, cudaStream_t st) {
    dim3 grid((unsigned)L.n_row_blocks, (unsigned)L.n_chunks, (unsigned)R);
    if (psign)
        reg_dense_kernel<RI, true, true><<<grid, kDenseThreads, 0, st>>>(
            U, A, cabs, L.Bpad, row_begin, row_end, L.rows_pad, L.chunk_cols, pgrad, prow, lossp, psign);
    else if (want_grad)
        reg_dense_kernel<RI, true, false><<<grid, kDenseThreads, 0, st>>>(
            U, A, cabs, L.Bpad, row_begin, row_end, L.rows_pad, L.chunk_cols, pgrad, prow, lossp, nullptr);
    else
        reg_dense_kernel<RI, false, false><<<grid, kDenseThreads, 0, st>>>(
            U, A, cabs, L.Bpad, row_begin, row_end, L.rows_pad, L.chunk_cols, pgrad, prow, lossp, nullptr);
}

int run_reg_dense(const RegProblem &P, const DenseLayout &L, char *ws, cudaStream_t st) {
    float *U = reinterpret_cast<float *>(ws + L.off_U);
    float *A = reinterpret_cast<float *>(ws + L.off_A);
    double *pgrad = reinterpret_cast<double *>(ws + L.off_pgrad);
    double *prow = P.row_loss_out ? reinterpret_cast<double *>(ws + L.off_prow) : nullptr;
    double *lossp = reinterpret_cast<double *>(ws + L.off_lossp);
    const int64_t n_rows = P.row_end - P.row_begin;
    const bool want_grad = P.grad_cols_out != nullptr;
    int *psign = P.row_sign_out ? reinterpret_cast<int *>(ws + L.off_psign) : nullptr;
    if (psign && !want_grad) {
        set_error("row sign sums need the gradient pass");
        return ARVAE_E_BADARG;
    }

    const double c = 2.0 * (double)P.factor * 1.4426950408889634074;  // 2 f log2(e)
    const float fsign = P.factor > 0.f ? 1.0f : (P.factor < 0.f ? -1.0f : 0.0f);
    const float cabs = P.factor != 0.f ? (float)fabs(c) : 1.0f;  // f == 0: xs == 0, any scale works
    pack_columns_kernel<<<(unsigned)ceil_div(L.Bpad, 256), 256, 0, st>>>(
        P.z, P.zrs, P.zcs, P.lab, P.lrs, P.lcs, P.dims, P.R, P.B, L.Bpad, fsign, U, A);
    ARVAE_LAUNCH_CHECK("pack_columns_kernel");

    if (L.n_units > 0) {
        profile_begin(st);
        if (L.RI == 4)
            launch_dense<4>(L, P.R, want_grad, U, A, cabs, P.row_begin, P.row_end, pgrad, prow, lossp, psign, st);
        else
            launch_dense<1>(L, P.R, want_grad, U, A, cabs, P.row_begin, P.row_end, pgrad, prow, lossp, psign, st);
        profile_end(st);
        ARVAE_LAUNCH_CHECK("reg_dense_kernel");
    }

    const double BB = (double)P.B * (double)P.B;
    const double lscale = (double)P.gamma / BB;
    // d/dx_i = 2 f / B^2 * sum_j g_ij, g = 4 * (accumulated sgn * (r - r^2)), times gamma
    const double gscale = 8.0 * (double)P.gamma * (double)P.factor / BB;
    const double pad_per_row = (double)(L.Bpad - P.B);
    const int64_t work = n_rows * P.R;
    reg_epilogue_kernel<<<(unsigned)(work > 0 ? ceil_div(work, 256) : 1), 256, 0, st>>>(
        pgrad, prow, lossp, L.n_chunks, P.R, n_rows, L.rows_pad, L.n_units, gscale, lscale,
        pad_per_row, P.grad_cols_out, P.row_loss_out, P.loss_out, P.loss_f32_out, psign, P.row_sign_out);
    ARVAE_LAUNCH_CHECK("reg_epilogue_kernel");
    return 0;
}

int run_scatter_bwd(const float *grad_cols, const float *grad_out, const RegDims &dims, int R,
                    int64_t n_rows, int64_t Z, float *grad_z, int64_t gzrs, cudaStream_t st) {
    const int64_t work = n_rows * Z;
    if (work <= 0) return 0;
    reg_scatter_bwd_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, st>>>(grad_cols, grad_out, dims,
                                                                         R, n_rows, Z, grad_z, gzrs);
    ARVAE_LAUNCH_CHECK("reg_scatter_bwd_kernel");
    return 0;
}

int run_pack_slice(const float *z, int64_t zrs, int64_t zcs, const float *lab, int64_t lrs, int64_t lcs,
                   const RegDims &dims, int R, int64_t n_rows, float *out, cudaStream_t st) {
    const int64_t work = n_rows * 2 * R;
    if (work <= 0) return 0;
    pack_slice_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, st>>>(z, zrs, zcs, lab, lrs, lcs, dims, R, n_rows, out);
    ARVAE_LAUNCH_CHECK("pack_slice_kernel");
    return 0;
}

int run_extract_perm(const unsigned long long *keys, int64_t B, int32_t *perm, cudaStream_t st) {
    if (B <= 0) return 0;
    extract_perm_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(keys, B, perm);
    ARVAE_LAUNCH_CHECK("extract_perm_kernel");
    return 0;
}

int run_sign_matrix(const float *a, int64_t stride, int64_t B, int8_t *out, cudaStream_t st) {
    if (B <= 0) return 0;
    sign_matrix_kernel<<<(unsigned)ceil_div(B * B, 256), 256, 0, st>>>(a, stride, B, out);
    ARVAE_LAUNCH_CHECK("sign_matrix_kernel");
    return 0;
}

}  // namespace arvae
