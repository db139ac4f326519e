// api.cu -- the extern "C" surface of libarvae_b200.so (declared in include/arvae_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <cmath>

#include "common.cuh"
#include "reg_internal.cuh"

namespace arvae {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own thread

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail_cuda(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    (void)cudaGetLastError();
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional pair-kernel timing --------------------------------------------------------------
constexpr int kProfRing = 64;
struct Profiler {
    bool on = false;
    int n = 0;  // launches recorded since last read (may exceed the ring)
    cudaEvent_t ev[kProfRing][2] = {};
    bool made = false;
};
static thread_local Profiler g_prof;

void profile_begin(cudaStream_t st) {
    Profiler &P = g_prof;
    if (!P.on) return;
    if (!P.made) {
        for (int i = 0; i < kProfRing; ++i) {
            cudaEventCreate(&P.ev[i][0]);
            cudaEventCreate(&P.ev[i][1]);
        }
        P.made = true;
    }
    cudaEventRecord(P.ev[P.n % kProfRing][0], st);
}

void profile_end(cudaStream_t st) {
    Profiler &P = g_prof;
    if (!P.on || !P.made) return;
    cudaEventRecord(P.ev[P.n % kProfRing][1], st);
    P.n++;
}

// ---- optional step timeline: an event after each group of launches ----------------------------------
constexpr int kTimelineMax = 256;
struct Timeline {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[kTimelineMax] = {};
    const char *name[kTimelineMax] = {};
};
static thread_local Timeline g_tl;

void timeline_mark(cudaStream_t st, const char *name) {
    Timeline &T = g_tl;
    if (!T.on || T.n >= kTimelineMax) return;
    if (!T.ev[T.n]) cudaEventCreate(&T.ev[T.n]);
    cudaEventRecord(T.ev[T.n], st);
    T.name[T.n] = name;
    T.n++;
}

int sm_count() {
    // immutable per-device cache (benign race: every thread computes the same value)
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

static int fill_dims(RegDims &d, const int32_t *reg_dims, const int32_t *label_cols, int R) {
    if (R < 0 || R > ARVAE_MAX_REG_DIMS) {
        set_error("R=%d out of range [0,%d]", R, ARVAE_MAX_REG_DIMS);
        return ARVAE_E_BADARG;
    }
    if (R > 0 && !reg_dims) {
        set_error("reg_dims is null");
        return ARVAE_E_BADARG;
    }
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < R; ++r) {
        d.zcol[r] = reg_dims[r];
        d.lcol[r] = label_cols ? label_cols[r] : reg_dims[r];
        if (d.zcol[r] < 0 || d.lcol[r] < 0) {
            set_error("negative column index at r=%d (resolve python-style negatives before the call)", r);
            return ARVAE_E_BADARG;
        }
    }
    return 0;
}

// per-thread device buffers of the host-buffer entry point
struct HostCache {
    int dev = -1;
    size_t z_bytes = 0, lab_bytes = 0, ws_bytes = 0, gc_bytes = 0;
    float *z = nullptr, *lab = nullptr, *gz = nullptr, *gc = nullptr;
    char *ws = nullptr;
    double *loss = nullptr;
    void release() {
        cudaFree(z); cudaFree(lab); cudaFree(gz); cudaFree(gc); cudaFree(ws); cudaFree(loss);
        z = lab = gz = gc = nullptr; ws = nullptr; loss = nullptr;
        z_bytes = lab_bytes = ws_bytes = gc_bytes = 0;
        dev = -1;
    }
};
static thread_local HostCache g_host;

}  // namespace arvae

using namespace arvae;

extern "C" {

int arvae_version(void) { return ARVAE_VERSION; }

const char *arvae_last_error(void) { return g_err; }

void arvae_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.n = 0;
}

int arvae_profile_pair_kernel_ms(float *sum_ms_out, int *n_out) {
    Profiler &P = g_prof;
    float sum = 0.f;
    const int n = P.n < kProfRing ? P.n : kProfRing;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        ARVAE_CUDA_TRY(cudaEventSynchronize(P.ev[i][1]));
        ARVAE_CUDA_TRY(cudaEventElapsedTime(&ms, P.ev[i][0], P.ev[i][1]));
        sum += ms;
    }
    if (sum_ms_out) *sum_ms_out = sum;
    if (n_out) *n_out = n;
    P.n = 0;
    return 0;
}

void arvae_timeline_enable(int on) {
    g_tl.on = on != 0;
    g_tl.n = 0;
}

// "name:ms;name:ms;..." -- milliseconds between consecutive marks recorded since arvae_timeline_enable(1); a mark named
// "begin..." starts a new interval (its own delta is not reported).  Synchronises the recorded events; clears them.
int arvae_timeline_report(char *buf, int32_t buf_bytes) {
    Timeline &T = g_tl;
    if (!buf || buf_bytes < 2) {
        set_error("bad argument to timeline_report");
        return ARVAE_E_BADARG;
    }
    int off = 0;
    buf[0] = 0;
    for (int i = 1; i < T.n; ++i) {
        if (strncmp(T.name[i], "begin", 5) == 0) continue;
        float ms = 0.f;
        ARVAE_CUDA_TRY(cudaEventSynchronize(T.ev[i]));
        ARVAE_CUDA_TRY(cudaEventElapsedTime(&ms, T.ev[i - 1], T.ev[i]));
        const int w = snprintf(buf + off, (size_t)(buf_bytes - off), "%s:%.4f;", T.name[i], ms);
        if (w < 0 || w >= buf_bytes - off) break;
        off += w;
    }
    T.n = 0;
    return 0;
}

int64_t arvae_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int arvae_device_sm_count(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        fail_cuda(e, "cudaGetDevice");
        return ARVAE_E_NODEVICE;
    }
    return sm_count();
}

size_t arvae_reg_loss_workspace_bytes_algo(int64_t B_total, int64_t n_rows, int32_t R, int32_t algo) {
    if (B_total < 0 || n_rows < 0 || R < 0 || R > ARVAE_MAX_REG_DIMS) return 0;
    const size_t d = dense_layout(B_total, n_rows, R > 0 ? R : 1, sm_count()).bytes;
    const size_t s = sorted_layout(B_total, n_rows, R > 0 ? R : 1, sm_count(), algo == ARVAE_ALGO_TRIANGLE).bytes;
    return d > s ? d : s;
}

size_t arvae_reg_loss_workspace_bytes(int64_t B_total, int64_t n_rows, int32_t R) {
    return arvae_reg_loss_workspace_bytes_algo(B_total, n_rows, R, ARVAE_ALGO_AUTO);
}

int arvae_reg_loss_fwdbwd_f32(const float *z_dev, int64_t z_row_stride, int64_t z_col_stride,
                              const float *labels_dev, int64_t lab_row_stride,
                              int64_t lab_col_stride, const int32_t *reg_dims_host,
                              const int32_t *label_cols_host, int32_t R, int64_t row_begin,
                              int64_t row_end, int64_t B_total, float gamma, float factor,
                              int32_t algo, double *loss_out_dev, float *loss_f32_out_dev,
                              float *grad_cols_out_dev, double *row_loss_out_dev, int32_t *row_sign_out_dev,
                              void *workspace_dev, size_t workspace_bytes, void *stream) {
    NvtxRange nvtx_range("arvae_reg_loss_fwdbwd_f32");
    RegProblem P;
    int rc = fill_dims(P.dims, reg_dims_host, label_cols_host, R);
    if (rc) return rc;
    if (B_total < 0 || row_begin < 0 || row_end < row_begin || row_end > B_total) {
        set_error("bad row range [%lld,%lld) for B_total=%lld", (long long)row_begin,
                  (long long)row_end, (long long)B_total);
        return ARVAE_E_BADARG;
    }
    if (!loss_out_dev || !workspace_dev || (B_total > 0 && R > 0 && (!z_dev || !labels_dev))) {
        set_error("null pointer argument");
        return ARVAE_E_BADARG;
    }
    if (algo != ARVAE_ALGO_AUTO && algo != ARVAE_ALGO_DENSE && algo != ARVAE_ALGO_SORTED && algo != ARVAE_ALGO_TRIANGLE) {
        set_error("unknown algo %d", algo);
        return ARVAE_E_BADARG;
    }
    P.z = z_dev; P.zrs = z_row_stride; P.zcs = z_col_stride;
    P.lab = labels_dev; P.lrs = lab_row_stride; P.lcs = lab_col_stride;
    P.R = R; P.B = B_total; P.row_begin = row_begin; P.row_end = row_end;
    P.gamma = gamma; P.factor = factor;
    P.loss_out = loss_out_dev; P.loss_f32_out = loss_f32_out_dev; P.grad_cols_out = grad_cols_out_dev; P.row_loss_out = row_loss_out_dev;
    P.row_sign_out = row_sign_out_dev;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    const bool sorted = algo == ARVAE_ALGO_SORTED || algo == ARVAE_ALGO_TRIANGLE ||
                        (algo == ARVAE_ALGO_AUTO && B_total >= kSortedMinBatch);
    P.use_triangle = algo == ARVAE_ALGO_TRIANGLE;
    const DenseLayout L = dense_layout(B_total, row_end - row_begin, R > 0 ? R : 1, sm_count());
    const SortedLayout LS = sorted_layout(B_total, row_end - row_begin, R > 0 ? R : 1, sm_count(), P.use_triangle);
    const size_t need = sorted ? LS.bytes : L.bytes;
    if (workspace_bytes < need) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return ARVAE_E_WORKSPACE;
    }
    if (R == 0) {  // empty dim tuple: the callers' loop adds nothing
        ARVAE_CUDA_TRY(cudaMemsetAsync(loss_out_dev, 0, sizeof(double), st));
        if (loss_f32_out_dev) ARVAE_CUDA_TRY(cudaMemsetAsync(loss_f32_out_dev, 0, sizeof(float), st));
        return 0;
    }
    if (sorted) return run_reg_sorted(P, LS, reinterpret_cast<char *>(workspace_dev), st);
    return run_reg_dense(P, L, reinterpret_cast<char *>(workspace_dev), st);
}

int arvae_reg_loss_path_flags(int64_t B_total, int64_t n_rows, int32_t R, int32_t algo,
                              const void *workspace_dev, int32_t *flags_out_host, void *stream) {
    if (B_total < 0 || n_rows < 0 || R <= 0 || R > ARVAE_MAX_REG_DIMS || !workspace_dev || !flags_out_host) {
        set_error("bad argument to path_flags");
        return ARVAE_E_BADARG;
    }
    const bool sorted = algo == ARVAE_ALGO_SORTED || algo == ARVAE_ALGO_TRIANGLE ||
                        (algo == ARVAE_ALGO_AUTO && B_total >= kSortedMinBatch);
    if (!sorted) {
        set_error("dense path selected for this shape: 2 MUFU per pair");
        return ARVAE_E_BADARG;
    }
    const SortedLayout LS = sorted_layout(B_total, n_rows, R, sm_count(), algo == ARVAE_ALGO_TRIANGLE);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int32_t tmp[3 * ARVAE_MAX_REG_DIMS + 3];
    ARVAE_CUDA_TRY(cudaMemcpyAsync(tmp, reinterpret_cast<const char *>(workspace_dev) + LS.off_flags,
                                   sizeof(tmp), cudaMemcpyDeviceToHost, st));
    ARVAE_CUDA_TRY(cudaStreamSynchronize(st));
    // whole-dim two-MUFU flag (triangle mode): report "no inliers"; else the inlier count of the segmented order
    for (int r = 0; r < R; ++r) flags_out_host[r] = tmp[r] ? 0 : tmp[2 * ARVAE_MAX_REG_DIMS + 3 + r];
    return 0;
}

// experiments only (not in the public header): byte offset of the per-CTA timestamp buffer and CTA count
extern "C" __attribute__((visibility("default"))) int64_t arvae_debug_times_offset(int64_t B_total, int64_t n_rows,
                                                                                 int32_t R, int32_t algo,
                                                                                 int32_t *g_max_out) {
    const SortedLayout LS = sorted_layout(B_total, n_rows, R, sm_count(), algo == ARVAE_ALGO_TRIANGLE);
    if (g_max_out) *g_max_out = LS.G_max;
    return (int64_t)LS.off_dbg;
}

int arvae_reg_loss_scatter_bwd_f32(const float *grad_cols_dev, const float *grad_out_dev,
                                   const int32_t *reg_dims_host, int32_t R, int64_t n_rows,
                                   int64_t Z, float *grad_z_dev, int64_t gz_row_stride,
                                   void *stream) {
    NvtxRange nvtx_range("arvae_reg_loss_scatter_bwd_f32");
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, nullptr, R);
    if (rc) return rc;
    if (n_rows < 0 || Z < 0 || (n_rows * Z > 0 && (!grad_z_dev || (R > 0 && !grad_cols_dev)))) {
        set_error("bad argument to scatter_bwd");
        return ARVAE_E_BADARG;
    }
    for (int r = 0; r < R; ++r)
        if (d.zcol[r] >= Z) {
            set_error("reg dim %d out of range for Z=%lld", d.zcol[r], (long long)Z);
            return ARVAE_E_BADARG;
        }
    return run_scatter_bwd(grad_cols_dev, grad_out_dev, d, R, n_rows, Z, grad_z_dev, gz_row_stride,
                           reinterpret_cast<cudaStream_t>(stream));
}

size_t arvae_latent_head_workspace_bytes(int64_t B, int64_t Z) {
    if (B < 0 || Z < 0) return 0;
    return latent_head_ws_bytes(B, Z);
}

int arvae_latent_head_fwd_f32(const float *loc_dev, const float *scale_dev, const float *eps_dev,
                              int64_t B, int64_t Z, float beta, float capacity, float *z_out_dev,
                              double *kld_sum_out_dev, float *kld_mean_out_dev,
                              float *kld_loss_out_dev, float *kcoef_out_dev, void *ws_dev,
                              size_t ws_bytes, void *stream) {
    NvtxRange nvtx_range("arvae_latent_head_fwd_f32");
    if (B < 0 || Z < 0 || !kld_sum_out_dev || !ws_dev ||
        (B * Z > 0 && (!loc_dev || !scale_dev || !eps_dev || !z_out_dev))) {
        set_error("bad argument to latent_head_fwd");
        return ARVAE_E_BADARG;
    }
    return run_latent_head_fwd(loc_dev, scale_dev, eps_dev, B, Z, beta, capacity, z_out_dev,
                               kld_sum_out_dev, kld_mean_out_dev, kld_loss_out_dev, kcoef_out_dev,
                               ws_dev, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int arvae_latent_head_bwd_f32(const float *loc_dev, const float *scale_dev, const float *eps_dev,
                              const float *dz_up_dev, const float *grad_cols_dev,
                              const float *greg_dev, const int32_t *reg_dims_host, int32_t R,
                              float kscale, const float *kcoef_dev, const float *gkld_dev,
                              int64_t B, int64_t Z, float *dloc_dev, float *dscale_dev,
                              void *stream) {
    NvtxRange nvtx_range("arvae_latent_head_bwd_f32");
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, nullptr, grad_cols_dev ? R : 0);
    if (rc) return rc;
    if (B < 0 || Z < 0 || (B * Z > 0 && (!loc_dev || !scale_dev || !eps_dev))) {
        set_error("bad argument to latent_head_bwd");
        return ARVAE_E_BADARG;
    }
    return run_latent_head_bwd(loc_dev, scale_dev, eps_dev, dz_up_dev, grad_cols_dev, greg_dev, d,
                               grad_cols_dev ? R : 0, kscale, kcoef_dev, gkld_dev, B, Z, dloc_dev,
                               dscale_dev, reinterpret_cast<cudaStream_t>(stream));
}

size_t arvae_head_fused_workspace_bytes(int64_t B, int32_t R) { return head_fused_ws_bytes(B, R); }

int arvae_head_fused_fwd_f32(const float *loc_dev, const float *sd_dev, int32_t sd_is_log_std, const float *eps_dev,
                             int64_t B, int64_t Z, const float *labels_dev, int64_t lab_row_stride,
                             int64_t lab_col_stride, const int32_t *reg_dims_host, const int32_t *label_cols_host,
                             int32_t R, float beta, float capacity, float gamma, float factor, float *z_out_dev,
                             float *scale_out_dev, float *kld_mean_out_dev, float *kld_loss_out_dev,
                             float *kcoef_out_dev, float *reg_loss_out_dev, float *grad_cols_out_dev,
                             void *workspace_dev, size_t workspace_bytes, void *stream) {
    NvtxRange nvtx_range("arvae_head_fused_fwd_f32");
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, label_cols_host, R);
    if (rc) return rc;
    if (R < 1 || Z < 1 || !loc_dev || !sd_dev || !eps_dev || !labels_dev || !z_out_dev || !reg_loss_out_dev ||
        !workspace_dev) {
        set_error("bad argument to head_fused_fwd (needs R >= 1 and non-null loc / sd / eps / labels / z_out / reg_loss_out / workspace)");
        return ARVAE_E_BADARG;
    }
    for (int r = 0; r < R; ++r)
        if (d.zcol[r] >= Z) {
            set_error("reg dim %d out of range for Z=%lld", d.zcol[r], (long long)Z);
            return ARVAE_E_BADARG;
        }
    return run_head_fused_fwd(loc_dev, sd_dev, sd_is_log_std, eps_dev, B, Z, labels_dev, lab_row_stride, lab_col_stride, d,
                              R, beta, capacity, gamma, factor, z_out_dev, scale_out_dev, kld_mean_out_dev,
                              kld_loss_out_dev, kcoef_out_dev, reg_loss_out_dev, grad_cols_out_dev,
                              reinterpret_cast<char *>(workspace_dev), workspace_bytes,
                              reinterpret_cast<cudaStream_t>(stream));
}

int arvae_head_fused_bwd_f32(const float *loc_dev, const float *sd_dev, int32_t sd_is_log_std, const float *eps_dev,
                             const float *dz_up_dev, const float *grad_cols_dev, const float *greg_dev,
                             const int32_t *reg_dims_host, int32_t R, const float *kcoef_dev, const float *gkld_dev,
                             int64_t B, int64_t Z, float *dloc_dev, float *dsd_dev, void *stream) {
    NvtxRange nvtx_range("arvae_head_fused_bwd_f32");
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, nullptr, grad_cols_dev ? R : 0);
    if (rc) return rc;
    if (B < 0 || Z < 0 || (B * Z > 0 && (!loc_dev || !sd_dev || !eps_dev))) {
        set_error("bad argument to head_fused_bwd");
        return ARVAE_E_BADARG;
    }
    return run_latent_head_bwd(loc_dev, sd_dev, eps_dev, dz_up_dev, grad_cols_dev, greg_dev, d, grad_cols_dev ? R : 0, 1.0f,
                               kcoef_dev, gkld_dev, B, Z, dloc_dev, dsd_dev, reinterpret_cast<cudaStream_t>(stream),
                               sd_is_log_std ? 1 : 0);
}

int arvae_reg_loss_host_f32(const float *z_host, int64_t B, int64_t Z, const float *labels_host,
                            int64_t A, const int32_t *reg_dims_host,
                            const int32_t *label_cols_host, int32_t R, float gamma, float factor,
                            int32_t algo, float *loss_out_host, float *grad_z_out_host,
                            void *stream) {
    NvtxRange nvtx_range("arvae_reg_loss_host_f32");
    if (B < 0 || Z <= 0 || A <= 0 || !loss_out_host || (B > 0 && (!z_host || !labels_host))) {
        set_error("bad argument to reg_loss_host");
        return ARVAE_E_BADARG;
    }
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, label_cols_host, R);
    if (rc) return rc;
    for (int r = 0; r < R; ++r)
        if (d.zcol[r] >= Z || d.lcol[r] >= A) {
            set_error("column index out of range at r=%d", r);
            return ARVAE_E_BADARG;
        }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0;
    ARVAE_CUDA_TRY(cudaGetDevice(&dev));
    HostCache &C = g_host;
    if (C.dev != dev) C.release();
    C.dev = dev;
    const size_t zb = sizeof(float) * (size_t)(B > 0 ? B : 1) * Z;
    const size_t lb = sizeof(float) * (size_t)(B > 0 ? B : 1) * A;
    const size_t gb = sizeof(float) * (size_t)(B > 0 ? B : 1) * (R > 0 ? R : 1);
    const size_t wb = arvae_reg_loss_workspace_bytes_algo(B, B, R, algo);
    // grow-only buffers; a failed allocation leaves the cache empty (no freed pointer or stale size survives)
    auto grow = [&](void **ptr, size_t &have, size_t need) -> cudaError_t {
        if (have >= need && *ptr) return cudaSuccess;
        cudaFree(*ptr);
        *ptr = nullptr;
        have = 0;
        cudaError_t e = cudaMalloc(ptr, need);
        if (e == cudaSuccess) have = need;
        else *ptr = nullptr;
        return e;
    };
    size_t gz_bytes = C.gz ? C.z_bytes : 0;
    cudaError_t ge = grow(reinterpret_cast<void **>(&C.gz), gz_bytes, zb);
    if (ge == cudaSuccess) ge = grow(reinterpret_cast<void **>(&C.z), C.z_bytes, zb);
    if (ge == cudaSuccess) ge = grow(reinterpret_cast<void **>(&C.lab), C.lab_bytes, lb);
    if (ge == cudaSuccess) ge = grow(reinterpret_cast<void **>(&C.gc), C.gc_bytes, gb);
    if (ge == cudaSuccess) ge = grow(reinterpret_cast<void **>(&C.ws), C.ws_bytes, wb);
    if (ge == cudaSuccess && !C.loss) ge = cudaMalloc(&C.loss, sizeof(double));
    if (ge != cudaSuccess) {
        C.release();
        return fail_cuda(ge, "cudaMalloc (host-buffer cache)");
    }

    if (B > 0) {
        ARVAE_CUDA_TRY(cudaMemcpyAsync(C.z, z_host, sizeof(float) * (size_t)B * Z, cudaMemcpyHostToDevice, st));
        ARVAE_CUDA_TRY(cudaMemcpyAsync(C.lab, labels_host, sizeof(float) * (size_t)B * A, cudaMemcpyHostToDevice, st));
    }
    rc = arvae_reg_loss_fwdbwd_f32(C.z, Z, 1, C.lab, A, 1, reg_dims_host, label_cols_host, R, 0, B, B,
                                   gamma, factor, algo, C.loss, nullptr, grad_z_out_host ? C.gc : nullptr,
                                   nullptr, nullptr, C.ws, C.ws_bytes, stream);
    if (rc) return rc;
    if (grad_z_out_host && B > 0) {
        rc = arvae_reg_loss_scatter_bwd_f32(C.gc, nullptr, reg_dims_host, R, B, Z, C.gz, Z, stream);
        if (rc) return rc;
        ARVAE_CUDA_TRY(cudaMemcpyAsync(grad_z_out_host, C.gz, sizeof(float) * (size_t)B * Z, cudaMemcpyDeviceToHost, st));
    }
    double loss = 0.0;
    ARVAE_CUDA_TRY(cudaMemcpyAsync(&loss, C.loss, sizeof(double), cudaMemcpyDeviceToHost, st));
    ARVAE_CUDA_TRY(cudaStreamSynchronize(st));
    *loss_out_host = (float)loss;
    return 0;
}

void arvae_host_release(void) { g_host.release(); }

int arvae_pack_columns_f32(const float *z_dev, int64_t z_row_stride, int64_t z_col_stride,
                           const float *labels_dev, int64_t lab_row_stride, int64_t lab_col_stride,
                           const int32_t *reg_dims_host, const int32_t *label_cols_host, int32_t R,
                           int64_t n_rows, float *out_dev, void *stream) {
    NvtxRange nvtx_range("arvae_pack_columns_f32");
    RegDims d;
    int rc = fill_dims(d, reg_dims_host, label_cols_host, R);
    if (rc) return rc;
    if (n_rows < 0 || (n_rows * R > 0 && (!z_dev || !labels_dev || !out_dev))) {
        set_error("bad argument to pack_columns");
        return ARVAE_E_BADARG;
    }
    return run_pack_slice(z_dev, z_row_stride, z_col_stride, labels_dev, lab_row_stride, lab_col_stride, d, R,
                          n_rows, out_dev, reinterpret_cast<cudaStream_t>(stream));
}

size_t arvae_attr_argsort_workspace_bytes(int64_t B) {
    if (B < 0) return 0;
    return sizeof(unsigned long long) * (size_t)sort_padded_size(B);
}

int arvae_attr_argsort_f32(const float *labels_dev, int64_t lab_stride, int64_t B, int32_t *perm_out_dev,
                           void *workspace_dev, size_t workspace_bytes, void *stream) {
    NvtxRange nvtx_range("arvae_attr_argsort_f32");
    if (B < 0 || !workspace_dev || (B > 0 && (!labels_dev || !perm_out_dev))) {
        set_error("bad argument to attr_argsort");
        return ARVAE_E_BADARG;
    }
    if (workspace_bytes < arvae_attr_argsort_workspace_bytes(B)) {
        set_error("workspace too small for attr_argsort");
        return ARVAE_E_WORKSPACE;
    }
    RegDims d;
    memset(&d, 0, sizeof(d));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(workspace_dev);
    int rc = run_sort_keys(labels_dev, lab_stride, 1, d, 1, B, sort_padded_size(B), keys, st);
    if (rc) return rc;
    return run_extract_perm(keys, B, perm_out_dev, st);
}

int arvae_measure_attributes_i64(const int64_t *measures_dev, int64_t B, int64_t T, int64_t row_stride,
                                 const int32_t *lut_dev, int64_t V, const float *rhy_weights_dev, float *out_dev,
                                 void *stream) {
    NvtxRange nvtx_range("arvae_measure_attributes_i64");
    if (B < 0 || T <= 0 || V <= 0 || (B > 0 && (!measures_dev || !lut_dev || !rhy_weights_dev || !out_dev))) {
        set_error("bad argument to measure_attributes");
        return ARVAE_E_BADARG;
    }
    return run_measure_attributes(reinterpret_cast<const long long *>(measures_dev), B, T, row_stride, lut_dev, V,
                                  rhy_weights_dev, out_dev, reinterpret_cast<cudaStream_t>(stream));
}

size_t arvae_eval_metrics_workspace_bytes(int64_t B, int32_t Z, int32_t A) {
    if (B < 1 || Z < 1 || A < 1 || Z > kEvalMaxCodes || A > kEvalMaxAttrs) return 0;
    return eval_metrics_workspace_bytes(B, Z, A);
}

int arvae_eval_metrics_f32(const float *codes_dev, int64_t codes_row_stride, int64_t codes_col_stride,
                           const float *attrs_dev, int64_t attrs_row_stride, int64_t attrs_col_stride, int64_t B,
                           int32_t Z, int32_t A, double *rho_out_dev, double *pval_out_dev, double *corr_out_dev,
                           double *sap_out_dev, double *scores_out_dev, void *workspace_dev, size_t workspace_bytes,
                           void *stream) {
    NvtxRange nvtx_range("arvae_eval_metrics_f32");
    if (B < 1 || B > 0x7fffffffLL || Z < 1 || A < 1 || Z > kEvalMaxCodes || A > kEvalMaxAttrs) {
        set_error("eval_metrics: need 1 <= B < 2^31, 1 <= Z <= %d, 1 <= A <= %d (got B=%lld Z=%d A=%d)", kEvalMaxCodes,
                  kEvalMaxAttrs, (long long)B, (int)Z, (int)A);
        return ARVAE_E_BADARG;
    }
    if (!codes_dev || !attrs_dev || !rho_out_dev || !pval_out_dev || !corr_out_dev || !sap_out_dev ||
        !scores_out_dev || !workspace_dev) {
        set_error("null pointer argument to eval_metrics");
        return ARVAE_E_BADARG;
    }
    if (workspace_bytes < eval_metrics_workspace_bytes(B, Z, A)) {
        set_error("workspace too small for eval_metrics");
        return ARVAE_E_WORKSPACE;
    }
    return run_eval_metrics(codes_dev, codes_row_stride, codes_col_stride, attrs_dev, attrs_row_stride,
                            attrs_col_stride, B, Z, A, rho_out_dev, pval_out_dev, corr_out_dev, sap_out_dev,
                            scores_out_dev, reinterpret_cast<char *>(workspace_dev),
                            reinterpret_cast<cudaStream_t>(stream));
}


// ---- sharded step over NVLink peer memory (reg_shard.cuh) --------------------------------------------------
static ShardCtx *as_ctx(void *ctx) { return reinterpret_cast<ShardCtx *>(ctx); }

size_t arvae_shard_comm_bytes(int64_t n_cap, int32_t R_cap, int32_t world) {
    if (n_cap < 1 || R_cap < 1 || R_cap > ARVAE_MAX_REG_DIMS || world < 1 || world > kMaxShardRanks) return 0;
    return shard_comm_bytes(n_cap, R_cap, world, nullptr);
}

int arvae_shard_create(int32_t rank, int32_t world, int64_t n_cap, int32_t R_cap, void **ctx_out) {
    if (!ctx_out || world < 1 || world > kMaxShardRanks || rank < 0 || rank >= world || n_cap < 1 || R_cap < 1 ||
        R_cap > ARVAE_MAX_REG_DIMS || (int64_t)world * n_cap > (int64_t)kKeyIdxMask) {
        set_error("bad argument to shard_create (world <= %d, R_cap <= %d)", kMaxShardRanks, ARVAE_MAX_REG_DIMS);
        return ARVAE_E_BADARG;
    }
    if (const char *w = getenv("ARVAE_SHARD_WAIT_MS")) {
        int rc = shard_set_wait_ms(atoll(w));
        if (rc) return rc;
    }
    ShardCtx *C = new ShardCtx();
    C->G = world; C->g = rank; C->R_cap = R_cap; C->n_cap = n_cap;
    cudaError_t e = cudaGetDevice(&C->device);
    if (e == cudaSuccess) {
        C->comm_bytes = shard_comm_bytes(n_cap, R_cap, world, C);
        C->ws_bytes = shard_ws_bytes(n_cap, R_cap, world, &C->off_mypos);
        e = cudaMalloc(&C->comm, C->comm_bytes);
    }
    if (e == cudaSuccess) e = cudaMemset(C->comm, 0, C->comm_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&C->ws, C->ws_bytes);
    if (e == cudaSuccess) e = cudaMemset(C->ws, 0, C->ws_bytes);  // the per-step flags start cleared; finalize re-clears them
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(C->comm);
        cudaFree(C->ws);
        delete C;
        return fail_cuda(e, "shard_create");
    }
    C->peer[rank] = C->comm;
    *ctx_out = C;
    return 0;
}

int arvae_shard_ipc_handle(void *ctx, void *handle_out) {
    if (!ctx || !handle_out) {
        set_error("null argument to shard_ipc_handle");
        return ARVAE_E_BADARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == ARVAE_SHARD_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    ARVAE_CUDA_TRY(cudaIpcGetMemHandle(&h, as_ctx(ctx)->comm));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int arvae_shard_open_peers(void *ctx, const void *handles) {
    if (!ctx || !handles) {
        set_error("null argument to shard_open_peers");
        return ARVAE_E_BADARG;
    }
    ShardCtx *C = as_ctx(ctx);
    for (int h = 0; h < C->G; ++h) {
        if (h == C->g || C->peer[h]) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, reinterpret_cast<const char *>(handles) + (size_t)h * ARVAE_SHARD_HANDLE_BYTES, sizeof(hd));
        void *p = nullptr;
        ARVAE_CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        C->peer[h] = reinterpret_cast<char *>(p);
        C->opened[h] = true;
    }
    return 0;
}

int arvae_shard_set_peer(void *ctx, int32_t rank, void *comm_dev) {
    ShardCtx *C = as_ctx(ctx);
    if (!C || rank < 0 || rank >= C->G || !comm_dev) {
        set_error("bad argument to shard_set_peer");
        return ARVAE_E_BADARG;
    }
    C->peer[rank] = reinterpret_cast<char *>(comm_dev);
    return 0;
}

void *arvae_shard_comm_ptr(void *ctx) { return ctx ? as_ctx(ctx)->comm : nullptr; }

static int shard_fill_step(ShardCtx *C, ShardStep &S, const int32_t *reg_dims_host, const int32_t *label_cols_host,
                           int32_t R, const int64_t *n_all_host) {
    if (!C || !n_all_host) {
        set_error("null argument to the shard step");
        return ARVAE_E_BADARG;
    }
    for (int h = 0; h < C->G; ++h)
        if (!C->peer[h]) {
            set_error("shard step: peer %d is not mapped (call arvae_shard_open_peers / arvae_shard_set_peer first)", h);
            return ARVAE_E_BADARG;
        }
    int rc = fill_dims(S.dims, reg_dims_host, label_cols_host, R);
    if (rc) return rc;
    S.R = R;
    for (int h = 0; h < kMaxShardRanks; ++h) S.n_all[h] = h < C->G ? n_all_host[h] : 0;
    return 0;
}

int arvae_shard_reg_loss_f32(void *ctx, const float *z_local_dev, int64_t z_row_stride, int64_t z_col_stride,
                             const float *labels_local_dev, int64_t lab_row_stride, int64_t lab_col_stride,
                             const int32_t *reg_dims_host, const int32_t *label_cols_host, int32_t R,
                             const int64_t *n_all_host, float gamma, float factor, double *loss_out_dev,
                             float *loss_f32_out_dev, float *grad_cols_out_dev, int32_t phases, void *stream) {
    NvtxRange nvtx_range("arvae_shard_reg_loss_f32");
    ShardCtx *C = as_ctx(ctx);
    ShardStep S;
    int rc = shard_fill_step(C, S, reg_dims_host, label_cols_host, R, n_all_host);
    if (rc) return rc;
    if (!loss_out_dev || (S.n_all[C->g] > 0 && (!z_local_dev || !labels_local_dev))) {
        set_error("null pointer argument to shard_reg_loss");
        return ARVAE_E_BADARG;
    }
    S.z = z_local_dev; S.zrs = z_row_stride; S.zcs = z_col_stride;
    S.lab = labels_local_dev; S.lrs = lab_row_stride; S.lcs = lab_col_stride;
    S.gamma = gamma; S.factor = factor;
    S.loss_out = loss_out_dev; S.loss_f32_out = loss_f32_out_dev; S.grad_cols_out = grad_cols_out_dev;
    S.phases = phases;
    return run_shard_step(*C, S, reinterpret_cast<cudaStream_t>(stream));
}

int arvae_shard_reg_loss_host_f32(void *ctx, const float *z_local_host, int64_t Z, const float *labels_local_host,
                                  int64_t A, const int32_t *reg_dims_host, const int32_t *label_cols_host, int32_t R,
                                  const int64_t *n_all_host, float gamma, float factor, float *loss_out_host,
                                  float *grad_z_out_host, void *stream) {
    NvtxRange nvtx_range("arvae_shard_reg_loss_host_f32");
    ShardCtx *C = as_ctx(ctx);
    ShardStep S;
    int rc = shard_fill_step(C, S, reg_dims_host, label_cols_host, R, n_all_host);
    if (rc) return rc;
    const int64_t n = S.n_all[C->g];
    if (Z <= 0 || A <= 0 || !loss_out_host || (n > 0 && (!z_local_host || !labels_local_host))) {
        set_error("bad argument to shard_reg_loss_host");
        return ARVAE_E_BADARG;
    }
    for (int r = 0; r < R; ++r)
        if (S.dims.zcol[r] >= Z || S.dims.lcol[r] >= A) {
            set_error("column index out of range at r=%d", r);
            return ARVAE_E_BADARG;
        }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t zb = sizeof(float) * (size_t)C->n_cap * Z, lb = sizeof(float) * (size_t)C->n_cap * A;
    if (C->h_z_bytes < zb) {
        cudaFree(C->h_z); cudaFree(C->h_gz);
        C->h_z = C->h_gz = nullptr; C->h_z_bytes = 0;
        ARVAE_CUDA_TRY(cudaMalloc(&C->h_z, zb));
        ARVAE_CUDA_TRY(cudaMalloc(&C->h_gz, zb));
        C->h_z_bytes = zb;
    }
    if (C->h_lab_bytes < lb) {
        cudaFree(C->h_lab);
        C->h_lab = nullptr; C->h_lab_bytes = 0;
        ARVAE_CUDA_TRY(cudaMalloc(&C->h_lab, lb));
        C->h_lab_bytes = lb;
    }
    if (!C->h_loss_host) ARVAE_CUDA_TRY(cudaHostAlloc(&C->h_loss_host, sizeof(double), cudaHostAllocMapped));
    if (!C->h_aux) ARVAE_CUDA_TRY(cudaStreamCreateWithFlags(&C->h_aux, cudaStreamNonBlocking));
    if (!C->h_ev) ARVAE_CUDA_TRY(cudaEventCreateWithFlags(&C->h_ev, cudaEventDisableTiming));
    if (n > 0) {
        // the two inputs travel side by side: labels on a second stream, joined before the sort kernel
        ARVAE_CUDA_TRY(cudaMemcpyAsync(C->h_lab, labels_local_host, sizeof(float) * (size_t)n * A, cudaMemcpyHostToDevice, C->h_aux));
        ARVAE_CUDA_TRY(cudaEventRecord(C->h_ev, C->h_aux));
        ARVAE_CUDA_TRY(cudaMemcpyAsync(C->h_z, z_local_host, sizeof(float) * (size_t)n * Z, cudaMemcpyHostToDevice, st));
        ARVAE_CUDA_TRY(cudaStreamWaitEvent(st, C->h_ev, 0));
    }
    // Where the finalize kernel writes the gradient: straight into the caller's buffer when that is pinned host memory
    // (mapped into the device under unified addressing: no device-to-host copy, no extra launch), else a device buffer.
    float *gz_dev = nullptr;
    bool gz_direct = false;
    if (grad_z_out_host && n > 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, grad_z_out_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
            gz_dev = reinterpret_cast<float *>(pa.devicePointer);
            gz_direct = true;
        } else {
            (void)cudaGetLastError();
            gz_dev = C->h_gz;
        }
    }
    double *loss_dev = nullptr;
    ARVAE_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&loss_dev), C->h_loss_host, 0));
    S.z = C->h_z; S.zrs = Z; S.zcs = 1;
    S.lab = C->h_lab; S.lrs = A; S.lcs = 1;
    S.gamma = gamma; S.factor = factor;
    S.loss_out = loss_dev; S.loss_f32_out = nullptr; S.grad_cols_out = nullptr;
    S.grad_z_out = gz_dev; S.grad_z_cols = Z;
    S.phases = 0;
    rc = run_shard_step(*C, S, st);
    if (rc) return rc;
    if (gz_dev && !gz_direct)
        ARVAE_CUDA_TRY(cudaMemcpyAsync(grad_z_out_host, C->h_gz, sizeof(float) * (size_t)n * Z, cudaMemcpyDeviceToHost, st));
    ARVAE_CUDA_TRY(cudaStreamSynchronize(st));
    *loss_out_host = (float)*C->h_loss_host;
    return 0;
}

// experiments only (not in the public header): per-CTA globaltimer stamps of the last pair-kernel launch of this rank
extern "C" __attribute__((visibility("default"))) int arvae_shard_debug_times(void *ctx, int64_t B_total, int32_t R,
                                                                            unsigned long long *out_host, int32_t n_cta) {
    ShardCtx *C = as_ctx(ctx);
    const SortedLayout L = sorted_layout(B_total, B_total, R, sm_count(), false);
    ARVAE_CUDA_TRY(cudaMemcpy(out_host, C->ws + L.off_dbg, sizeof(unsigned long long) * 2 * (size_t)n_cta, cudaMemcpyDeviceToHost));
    return 0;
}

int arvae_shard_status(void *ctx, int32_t *status_out, uint64_t *epoch_out, void *stream) {
    ShardCtx *C = as_ctx(ctx);
    if (!C) {
        set_error("null context");
        return ARVAE_E_BADARG;
    }
    ShardHeader h;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ARVAE_CUDA_TRY(cudaMemcpyAsync(&h, C->comm, sizeof(h), cudaMemcpyDeviceToHost, st));
    ARVAE_CUDA_TRY(cudaStreamSynchronize(st));
    if (status_out) *status_out = h.status;
    if (epoch_out) *epoch_out = h.epoch;
    return 0;
}

int arvae_shard_destroy(void *ctx) {
    ShardCtx *C = as_ctx(ctx);
    if (!C) return 0;
    for (int h = 0; h < C->G; ++h)
        if (C->opened[h]) cudaIpcCloseMemHandle(C->peer[h]);
    cudaFree(C->comm); cudaFree(C->ws);
    cudaFree(C->h_z); cudaFree(C->h_gz); cudaFree(C->h_lab);
    if (C->h_loss_host) cudaFreeHost(C->h_loss_host);
    if (C->h_aux) cudaStreamDestroy(C->h_aux);
    if (C->h_ev) cudaEventDestroy(C->h_ev);
    (void)cudaGetLastError();
    delete C;
    return 0;
}

int arvae_reg_sign_matrix_i8(const float *labels_dev, int64_t lab_stride, int64_t B,
                             int8_t *out_dev, void *stream) {
    if (B < 0 || (B > 0 && (!labels_dev || !out_dev))) {
        set_error("bad argument to sign_matrix");
        return ARVAE_E_BADARG;
    }
    return run_sign_matrix(labels_dev, lab_stride, B, out_dev, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
