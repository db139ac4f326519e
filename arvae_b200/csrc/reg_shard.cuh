// reg_shard.cuh -- the attribute-regularization step sharded over the GPUs of one NVSwitch box, with the exchange
// done by the kernels themselves over NVLink peer memory (included by reg_sorted.cu).
//
// One process per GPU; rank g holds the samples its own encoder produced (z_local [n_g, Z], labels_local [n_g, A]).
// The pair matrix is cut by ROW BLOCKS OF THE SORTED ORDER: every rank runs the same plan as a single GPU would
// (same row tiles, same tile classes, same cost model), cut into G x Gc CTA ranges of which rank g executes the g-th
// Gc.  Because row sums are fixed-point integers and the tile geometry is the single-GPU one, the loss and every
// gradient element are BITWISE identical to the single-GPU result for any G.
//
//   (A) publish   each rank argsorts ITS OWN n_g rows per dim (sort.cu; 1/G of the sort work) and stores the sorted run
//                 -- 64-bit key + latent, 12 bytes per element -- straight into the run slot g of EVERY peer's
//                 communication buffer (NVLink stores), then raises flag A at every peer.          [all-gather]
//   (B) merge     waits for the G flags, places every element of every run at its global sorted position by G - 1
//                 binary searches in the other runs (keys are unique: attribute, then global index), which rebuilds
//                 the single-GPU sorted columns on every rank; plan; pair kernel on this rank's CTA range, row sums
//                 into this rank's accumulators (integer atomics); the last CTA publishes the rank's exact loss
//                 partial and raises flag B at every peer.
//   (C) finalize  waits for the G flags B, PULLS the row sums of its own samples from the accumulators of the 1-2
//                 ranks that swept those sorted positions (NVLink loads), and all G loss partials -> grad_cols, loss.
//                                                                                [gradient return + all-reduce]
// No NCCL call and no host synchronisation in the step; flags are epoch counters, waits are bounded (a peer that
// never signals turns the loss into NaN instead of hanging the GPU).
#pragma once

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_sys_s64(const long long *p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ ShardHeader *shard_header(const ShardView &v, int h) {
    return reinterpret_cast<ShardHeader *>(v.peer[h]);
}
__device__ __forceinline__ unsigned long long *shard_flags(const ShardView &v, int h, bool second) {
    return reinterpret_cast<unsigned long long *>(v.peer[h] + (second ? v.off_flagB : v.off_flagA));
}
__device__ __forceinline__ unsigned long long *shard_run_keys(const ShardView &v, int h, int src, int r) {
    return reinterpret_cast<unsigned long long *>(v.peer[h] + v.off_keys) + ((int64_t)src * v.R_cap + r) * v.n_cap;
}
__device__ __forceinline__ float *shard_run_xs(const ShardView &v, int h, int src, int r) {
    return reinterpret_cast<float *>(v.peer[h] + v.off_xs) + ((int64_t)src * v.R_cap + r) * v.n_cap;
}
__device__ __forceinline__ acc_t *shard_acc(const ShardView &v, int h) {
    return reinterpret_cast<acc_t *>(v.peer[h] + v.off_acc);
}

constexpr unsigned long long kShardWaitNs = 4000000000ull;  // 4 s: far beyond any healthy step

// Threads 0..G-1 of the CTA each wait for one peer's flag to reach `epoch`; everybody leaves together.
__device__ __forceinline__ void shard_wait(const ShardView &v, bool second, unsigned long long epoch) {
    if ((int)threadIdx.x < v.G) {
        const unsigned long long *f = shard_flags(v, v.g, second) + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        volatile int *status = &shard_header(v, v.g)->status;
        while (ld_acquire_sys_u64(f) < epoch) {
            if (*status != 0) break;  // an earlier wait already gave up: do not stall again
            if (global_timer_ns() - t0 > kShardWaitNs) {
                atomicExch(&shard_header(v, v.g)->status, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// Tail of a kernel whose every CTA has finished writing data the peers will read: the last CTA to arrive raises this
// rank's flag at every peer.  Callers have executed __threadfence_system() after their writes.
__device__ __forceinline__ bool shard_last_cta(unsigned int *ticket, unsigned int n_cta) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(ticket, 1u) == n_cta - 1;
    }
    __syncthreads();
    if (s_last) __threadfence_system();
    return s_last != 0;
}
__device__ __forceinline__ void shard_signal(const ShardView &v, bool second, unsigned long long epoch) {
    if ((int)threadIdx.x < v.G) st_release_sys_u64(shard_flags(v, (int)threadIdx.x, second) + v.g, epoch);
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// (A) publish this rank's sorted runs to every peer
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shard_publish_kernel(ShardView v, const unsigned long long *__restrict__ keys, int64_t N, const float *__restrict__ z,
                     int64_t zrs, int64_t zcs, RegDims dims, float fsign) {
    const int r = blockIdx.y;
    const int64_t n = v.row_off[v.g + 1] - v.row_off[v.g];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long epoch = shard_header(v, v.g)->epoch + 1;
    if (p < n) {
        const unsigned long long key = keys[(int64_t)r * N + p];
        const int64_t i = (int64_t)(key & kKeyIdxMask) - v.row_off[v.g];
        const float xs = signed_latent(__ldg(z + i * zrs + (int64_t)dims.zcol[r] * zcs), fsign);
        for (int h = 0; h < v.G; ++h) {
            shard_run_keys(v, h, v.g, r)[p] = key;
            shard_run_xs(v, h, v.g, r)[p] = xs;
        }
    }
    __threadfence_system();
    if (shard_last_cta(&shard_header(v, v.g)->done_pub, gridDim.x * gridDim.y)) {
        if (threadIdx.x == 0) shard_header(v, v.g)->done_pub = 0;
        shard_signal(v, false, epoch);
    }
}

// ------------------------------------------------------------------------------------------------------------
// (B) merge the G runs into the global sorted columns
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sortable_to_float_bits(unsigned int u) {  // inverse of sort.cu's float_to_sortable
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ int64_t count_below(const unsigned long long *__restrict__ run, int64_t n, unsigned long long key) {
    int64_t lo = 0, hi = n;  // first position whose key is >= `key`
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (run[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
shard_merge_kernel(ShardView v, int64_t Bpad, float cabs, float *__restrict__ Xs, float *__restrict__ As,
                   float *__restrict__ Es, int *__restrict__ perm, int *__restrict__ flags, int *__restrict__ mypos) {
    const int r = blockIdx.y;
    const int64_t B = v.row_off[v.G];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    shard_wait(v, false, shard_header(v, v.g)->epoch + 1);
    if (t >= Bpad) return;
    const int64_t base = (int64_t)r * Bpad;
    if (t >= B) {  // padding: |t - s| = 1 and zero gradient for every row, in either tanh form
        Xs[base + t] = ARVAE_PAD_U;
        As[base + t] = ARVAE_PAD_A;
        Es[base + t] = 8.5070592e37f;
        perm[base + t] = -1;
        return;
    }
    int h = 0;
    while (t >= v.row_off[h + 1]) ++h;  // input slot t = element p of run h
    const int64_t p = t - v.row_off[h];
    const unsigned long long key = shard_run_keys(v, v.g, h, r)[p];
    int64_t pos = p;
    for (int o = 0; o < v.G; ++o)
        if (o != h) pos += count_below(shard_run_keys(v, v.g, o, r), v.row_off[o + 1] - v.row_off[o], key);
    const float xs = shard_run_xs(v, v.g, h, r)[p];
    const int64_t idx = (int64_t)(key & kKeyIdxMask);
    Xs[base + pos] = xs;
    As[base + pos] = sortable_to_float_bits(key_sortable_attr(key));  // NaNs come back as one quiet NaN: only compared
    Es[base + pos] = key_is_outlier(key) ? 1.0f : exp2f(cabs * xs);
    perm[base + pos] = (int)idx;
    note_nonfinite(xs, flags, r);
    if (h == v.g) mypos[(int64_t)r * v.n_cap + (idx - v.row_off[v.g])] = (int)pos;
    if (t == 0) {  // inliers of the dim = keys below the outlier bit, over all runs
        int64_t c = 0;
        for (int o = 0; o < v.G; ++o)
            c += count_below(shard_run_keys(v, v.g, o, r), v.row_off[o + 1] - v.row_off[o], 1ull << 63);
        flags[kFlagNIn + r] = (int)c;
    }
}

// Last CTA of the pair kernel: this rank's exact loss partial -> its header, then flag B at every peer.
__device__ void shard_pair_kernel_tail(const TilesArgs &a, acc_t *sh /* shared memory, 2 * kDuoThreads, free by now */) {
    ShardHeader *hdr = shard_header(a.shard, a.shard.g);
    const unsigned long long epoch = hdr->epoch + 1;
    if (!shard_last_cta(&hdr->done_pair, gridDim.x)) return;
    acc_t th = 0, tl = 0;
    for (unsigned int u = threadIdx.x; u < gridDim.x; u += blockDim.x) {
        th += __ldcg(a.lossp + 2 * u);
        tl += __ldcg(a.lossp + 2 * u + 1);
    }
    sh[threadIdx.x] = th;
    sh[kDuoThreads + threadIdx.x] = tl;
    __syncthreads();
    for (int o = kDuoThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[threadIdx.x] += sh[threadIdx.x + o];
            sh[kDuoThreads + threadIdx.x] += sh[kDuoThreads + threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        hdr->loss_part[0] = sh[0];
        hdr->loss_part[1] = sh[kDuoThreads];
        hdr->done_pair = 0;
        __threadfence_system();
    }
    __syncthreads();
    shard_signal(a.shard, true, epoch);
}

// ------------------------------------------------------------------------------------------------------------
// (C) finalize: pull this rank's row sums and every rank's loss partial
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shard_finalize_kernel(TilesArgs a, const int *__restrict__ mypos, int R, double gscale, double lscale,
                      double pad_per_row, float *__restrict__ grad_cols, double *__restrict__ loss_out,
                      float *__restrict__ loss_f32_out) {
    const ShardView &v = a.shard;
    ShardHeader *hdr = shard_header(v, v.g);
    const unsigned long long epoch = hdr->epoch + 1;
    shard_wait(v, true, epoch);
    const int64_t n = v.row_off[v.g + 1] - v.row_off[v.g];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n * R, dim fastest (coalesced stores)
    if (grad_cols && idx < n * R) {
        const int r = (int)(idx % R);
        const int64_t i = idx / R;
        const int64_t pos = mypos[(int64_t)r * v.n_cap + i];
        const int64_t rr = (int64_t)r * a.n_row_tiles + pos / kTileRows;
        const long long T = a.prefix[a.n_rr];
        const int h0 = (int)(owner_of_pos(a.prefix[rr], T, a.G) / v.Gc);
        const int h1 = (int)(owner_of_pos(a.prefix[rr + 1] - (long long)a.cost8[rr * a.S + a.S - 1], T, a.G) / v.Gc);
        acc_t g = 0;
        for (int h = h0; h <= h1; ++h) g += ld_relaxed_sys_s64(shard_acc(v, h) + rr * kTileRows + pos % kTileRows);
        const bool poisoned = row_is_poisoned(a.flags, r, a.Xs[(int64_t)r * a.Bpad + pos]);
        grad_cols[idx] = poisoned ? __int_as_float(0x7fc00000) : (float)((double)g * kFixScale * gscale);
    }
    if (blockIdx.x == 0) {
        __shared__ acc_t shl[2][kMaxShardRanks];
        if ((int)threadIdx.x < v.G) {
            const ShardHeader *ph = shard_header(v, (int)threadIdx.x);
            shl[0][threadIdx.x] = ld_relaxed_sys_s64(&ph->loss_part[0]);
            shl[1][threadIdx.x] = ld_relaxed_sys_s64(&ph->loss_part[1]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            acc_t hi = 0, lo = 0;
            for (int h = 0; h < v.G; ++h) { hi += shl[0][h]; lo += shl[1][h]; }
            double total = loss_from_hilo(hi, lo) - pad_per_row * (double)v.row_off[v.G] * (double)R;
            if (any_nonfinite(a.flags, R) || hdr->status) total = __longlong_as_double(0x7ff8000000000000LL);
            *loss_out = total * lscale;
            if (loss_f32_out) *loss_f32_out = (float)(total * lscale);
        }
    }
    // the step is complete on this rank once every CTA is past its wait and its pulls: advance the epoch
    if (shard_last_cta(&hdr->done_fin, gridDim.x)) {
        if (threadIdx.x == 0) {
            hdr->done_fin = 0;
            hdr->epoch = epoch;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
size_t shard_comm_bytes(int64_t n_cap, int R_cap, int G, ShardCtx *fill) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    take(sizeof(ShardHeader));
    const size_t fa = take(sizeof(unsigned long long) * kMaxShardRanks);
    const size_t fb = take(sizeof(unsigned long long) * kMaxShardRanks);
    const size_t ok = take(sizeof(unsigned long long) * (size_t)G * R_cap * n_cap);
    const size_t ox = take(sizeof(float) * (size_t)G * R_cap * n_cap);
    const int64_t tiles = ceil_div((int64_t)G * n_cap, kTileRows);
    const size_t oa = take(sizeof(acc_t) * (size_t)R_cap * tiles * kTileRows);
    if (fill) {
        fill->off_flagA = fa; fill->off_flagB = fb; fill->off_keys = ok; fill->off_xs = ox; fill->off_acc = oa;
    }
    return off;
}

size_t shard_ws_bytes(int64_t n_cap, int R_cap, int G, size_t *off_mypos) {
    const int64_t B = (int64_t)G * n_cap;
    const SortedLayout L = sorted_layout(B, B, R_cap, sm_count(), false);
    const size_t o = (L.bytes + 255) / 256 * 256;
    if (off_mypos) *off_mypos = o;
    return o + sizeof(int) * (size_t)R_cap * n_cap;
}

int run_shard_step(ShardCtx &C, const ShardStep &S, cudaStream_t st) {
    const int G = C.G, g = C.g;
    ShardView v;
    memset(&v, 0, sizeof(v));
    v.G = G; v.g = g; v.R_cap = C.R_cap; v.n_cap = C.n_cap;
    v.row_off[0] = 0;
    for (int h = 0; h < G; ++h) {
        if (S.n_all[h] < 0 || S.n_all[h] > C.n_cap) {
            set_error("shard step: rank %d has %lld rows, the communicator was sized for %lld", h, (long long)S.n_all[h],
                      (long long)C.n_cap);
            return ARVAE_E_BADARG;
        }
        v.row_off[h + 1] = v.row_off[h] + S.n_all[h];
        v.peer[h] = C.peer[h];
    }
    v.off_flagA = C.off_flagA; v.off_flagB = C.off_flagB; v.off_keys = C.off_keys; v.off_xs = C.off_xs; v.off_acc = C.off_acc;
    const int64_t B = v.row_off[G], n_local = S.n_all[g];
    if (S.R < 1 || S.R > C.R_cap || B < 1 || B > (int64_t)kKeyIdxMask) {
        set_error("shard step: R=%d (capacity %d), B=%lld out of range", S.R, C.R_cap, (long long)B);
        return ARVAE_E_BADARG;
    }
    const int phases = S.phases ? S.phases : 7;
    const SortedLayout L = sorted_layout(B, B, S.R, sm_count(), false);
    if (L.bytes > C.off_mypos) {
        set_error("shard step: workspace too small");
        return ARVAE_E_WORKSPACE;
    }
    char *ws = C.ws;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + L.off_keys);
    int *flags = reinterpret_cast<int *>(ws + L.off_flags);
    int *n_in = flags + kFlagNIn;
    int *perm = reinterpret_cast<int *>(ws + L.off_perm);
    int *mypos = reinterpret_cast<int *>(ws + C.off_mypos);
    const double c = 2.0 * (double)S.factor * 1.4426950408889634074;
    const float fsign = S.factor > 0.f ? 1.0f : (S.factor < 0.f ? -1.0f : 0.0f);
    const float cabs = S.factor != 0.f ? (float)fabs(c) : 1.0f;

    // pair-kernel CTAs per rank: one per SM, fewer when the whole plan has less than 4 units per CTA
    int64_t Gc = (int64_t)sm_count() * tiles_ctas_per_sm();
    if (Gc * G > L.F / 4) Gc = L.F / 4 / G;
    if (Gc < 1) Gc = 1;
    v.Gc = (int)Gc;

    TilesArgs a;
    memset(&a, 0, sizeof(a));
    a.Xs = reinterpret_cast<float *>(ws + L.off_Us);
    a.As = reinterpret_cast<float *>(ws + L.off_As);
    a.Es = reinterpret_cast<float *>(ws + L.off_Es);
    a.cabs = cabs;
    a.flags = flags;
    a.n_in = n_in;
    a.Bpad = L.Bpad; a.n_rows = B;
    a.n_row_tiles = L.n_row_tiles; a.S = L.S; a.F = L.F; a.n_rr = L.n_rr;
    a.P = golden_stride(L.S);
    a.G = (int)(Gc * G);
    a.c_first = (int)(Gc * g);
    a.cls8 = reinterpret_cast<unsigned int *>(ws + L.off_cls8);
    a.cost8 = reinterpret_cast<unsigned short *>(ws + L.off_cost8);
    a.prefix = reinterpret_cast<long long *>(ws + L.off_prefix);
    a.acc_g = reinterpret_cast<acc_t *>(C.comm + C.off_acc);
    a.lossp = reinterpret_cast<acc_t *>(ws + L.off_lossp);
    a.B = B;
    a.shard = v;
    const bool want_grad = S.grad_cols_out != nullptr;

    if (phases & 1) {
        const int64_t N = sort_padded_size(n_local > 0 ? n_local : 1);
        KeySpec spec;
        spec.lab = S.lab; spec.lrs = S.lrs; spec.lcs = S.lcs;
        spec.z = S.z; spec.zrs = S.zrs; spec.zcs = S.zcs;
        spec.fsign = fsign; spec.cabs = cabs; spec.segment = 1;
        spec.idx_offset = v.row_off[g];
        spec.dims = S.dims;
        int rc = run_sort_keys_spec(spec, S.R, n_local, N, keys, st);
        if (rc) return rc;
        dim3 gp((unsigned)ceil_div(n_local > 0 ? n_local : 1, 256), (unsigned)S.R);
        shard_publish_kernel<<<gp, 256, 0, st>>>(v, keys, N, S.z, S.zrs, S.zcs, S.dims, fsign);
        ARVAE_LAUNCH_CHECK("shard_publish_kernel");
    }
    if (phases & 2) {
        ARVAE_CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int) * kFlagClearInts, st));
        dim3 gm((unsigned)ceil_div(L.Bpad, 256), (unsigned)S.R);
        shard_merge_kernel<<<gm, 256, 0, st>>>(v, L.Bpad, cabs, const_cast<float *>(a.Xs), const_cast<float *>(a.As),
                                               const_cast<float *>(a.Es), perm, flags, mypos);
        ARVAE_LAUNCH_CHECK("shard_merge_kernel");
        // Only now may the accumulators be cleared: every peer has published this step's run, hence finished pulling
        // the previous step's row sums from them.
        if (want_grad) ARVAE_CUDA_TRY(cudaMemsetAsync(a.acc_g, 0, L.acc_bytes, st));
        int *combo_cost = reinterpret_cast<int *>(ws + L.off_combo);
        plan_classes_kernel<<<(unsigned)L.n_rr, 256, 0, st>>>(a, combo_cost);
        ARVAE_LAUNCH_CHECK("plan_classes_kernel");
        plan_scan_kernel<<<1, 1024, 0, st>>>(combo_cost, L.n_rr, a.prefix);
        ARVAE_LAUNCH_CHECK("plan_scan_kernel");
        profile_begin(st);
        launch_tiles(a, (int)Gc, want_grad, false, st);
        profile_end(st);
        ARVAE_LAUNCH_CHECK("reg_tiles_kernel");
    }
    if (phases & 4) {
        RegProblem P;
        P.B = B; P.gamma = S.gamma; P.factor = S.factor;
        double lscale, gscale, pad_per_row;
        reg_scales(P, L.Bpad, lscale, gscale, pad_per_row);
        const int64_t work = n_local * S.R;
        shard_finalize_kernel<<<(unsigned)(work > 0 ? ceil_div(work, 256) : 1), 256, 0, st>>>(
            a, mypos, S.R, gscale, lscale, pad_per_row, S.grad_cols_out, S.loss_out, S.loss_f32_out);
        ARVAE_LAUNCH_CHECK("shard_finalize_kernel");
    }
    return 0;
}
