// reg_shard.cuh -- sorted runs -> global sorted columns (single GPU and sharded), and the attribute-regularization
// step sharded over the GPUs of one NVSwitch box with the exchange done by the kernels themselves over NVLink peer
// memory (included by reg_sorted.cu).
//
// One process per GPU; rank g holds the samples its own encoder produced (z_local [n_g, Z], labels_local [n_g, A]).
// The pair matrix is cut by ROW BLOCKS OF THE SORTED ORDER: every rank runs the same plan as a single GPU would
// (same row tiles, same tile classes, same cost model), cut into G x Gc CTA ranges of which rank g executes the g-th
// Gc.  Because row sums are fixed-point integers and the tile geometry is the single-GPU one, the loss and every
// gradient element are BITWISE identical to the single-GPU result for any G.
//
//   (A) publish   each rank radix-sorts ITS OWN rows per dim as runs of <= 8192 samples (sort.cu: one CTA per run,
//                 1/G of the sort work) and stores every sorted element -- key, latent and the step's epoch in one
//                 16-byte store -- straight into the run slots of EVERY peer's communication buffer (NVLink stores).
//                 No fence, no flag: an element says itself whether it has arrived.                  [all-gather]
//   (B) merge     every rank places every element of every run at its global sorted position (binary searches in
//                 windows of the other runs staged in shared memory; keys are unique: attribute, then global
//                 index), which rebuilds the single-GPU sorted columns on every rank; plan; pair kernel on this
//                 rank's CTA range, row sums into this rank's accumulators (integer atomics); the last CTA
//                 publishes the rank's exact loss partial and raises flag B at every peer.
//   (C) finalize  waits for the G flags B, PULLS the row sums of its own samples from the accumulators of the 1-2
//                 ranks that swept those sorted positions (NVLink loads), and all G loss partials -> grad_cols, loss.
//                                                                                [gradient return + all-reduce]
// No NCCL call and no host synchronisation in the step; waits are bounded (a peer that never shows up turns the
// loss into NaN instead of hanging the GPU).  System-scope fences cost ~7 us each on this platform and are avoided:
// published elements validate themselves, and what peers pull lives in the producer's own L2, where a device-scope
// fence before the flag store is enough.
#pragma once

namespace {

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
__device__ __forceinline__ unsigned long long *shard_flags(const ShardView &v, int h) {
    return reinterpret_cast<unsigned long long *>(v.peer[h] + v.off_flagB);
}
__device__ __forceinline__ acc_t *shard_acc(const ShardView &v, int h) {
    return reinterpret_cast<acc_t *>(v.peer[h] + v.off_acc);
}

// Longest a kernel waits for a peer (element, position or flag) before it gives up, marks the communicator broken and
// lets the step end with NaN results instead of hanging the GPU.  30 s by default -- ranks of a training job drift (data
// loading, checkpoints) -- and ARVAE_SHARD_WAIT_MS at arvae_shard_create overrides it.
__constant__ unsigned long long kShardWaitNs = 30000000000ull;

// Threads 0..G-1 of the CTA each wait for one peer's flag B to reach `epoch`; everybody leaves together.
__device__ __forceinline__ void shard_wait(const ShardView &v, unsigned long long epoch) {
    if ((int)threadIdx.x < v.G) {
        const unsigned long long *f = shard_flags(v, v.g) + threadIdx.x;
        volatile int *status = &shard_header(v, v.g)->status;
        unsigned long long t0 = 0;
        while (ld_volatile_u64(f) < epoch) {
            if (*status != 0) break;  // an earlier wait already gave up: do not stall again
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kShardWaitNs) {
                atomicExch(&shard_header(v, v.g)->status, 1);
                break;
            }
        }
    }
    __syncthreads();
}

// "Last CTA" ticket with device-scope fences: true in the CTA that arrives last, after which it sees what every
// other CTA wrote before arriving.  The caller resets the ticket.
__device__ __forceinline__ bool last_cta(unsigned int *ticket, unsigned int n_cta) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == n_cta - 1;
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// merge T sorted runs into the global sorted columns
// ------------------------------------------------------------------------------------------------------------
// Where a sharded step's merge stores the global positions of a rank's own run elements: slot (run, dim) of every
// rank's position region, element i of the run at [i] as {position, epoch} in one 8-byte store.  n_dest == 0: the
// merge writes the sorted columns itself (single GPU: every run is its own).
struct PosDest {
    int n_dest;
    char *base[kMaxShardRanks];
};
__host__ __device__ static inline uint2 *pos_slot(char *pos_base, int R_cap, int run, int r) {
    return reinterpret_cast<uint2 *>(pos_base) + ((int64_t)run * R_cap + r) * kRunCap;
}

struct RunSet {
    int T, R_cap;                         // runs; dims the slots are sized for
    int64_t run_off[kMaxRuns + 1];        // global index of each run's first sample; [T] = B
    int blk_off[kMaxRuns + 1];            // prefix of ceil(n_t / 256): merge CTA -> run
    char *base;                           // this GPU's run-slot region
    const unsigned long long *epoch_ctr;  // non-null: elements are valid once they carry (uint32)(*epoch_ctr + 1)
    int *status;                          // where a timed-out wait is recorded (with epoch_ctr)
};

__device__ __forceinline__ float sortable_to_float_bits(unsigned int u) {  // inverse of sort.cu's float_to_sortable
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

// One element of a run.  With validation the load bypasses L1 and spins until the element carries the step's epoch
// (its 16 bytes were stored at once by the producer, possibly another GPU).
template <bool VALIDATE>
__device__ __forceinline__ uint4 load_run_elem_once(const RunElem *p) {
    if (!VALIDATE) return *reinterpret_cast<const uint4 *>(p);
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool VALIDATE>
__device__ __forceinline__ uint4 load_run_elem(const RunElem *p, unsigned int epoch, int *status) {
    uint4 v = load_run_elem_once<VALIDATE>(p);
    if (!VALIDATE) return v;
    unsigned long long t0 = 0;
    while (v.w != epoch) {
        if (*reinterpret_cast<volatile int *>(status) != 0) break;
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > kShardWaitNs) {
            atomicExch(status, 1);
            break;
        }
        __nanosleep(100);  // the producer may share this SM (overlapped launches): leave it the issue slots
        v = load_run_elem_once<VALIDATE>(p);
    }
    return v;
}
// Keys of up to NB elements (null pointer: skipped), all loads in flight together; only an element that has not
// arrived yet is waited for.
template <bool VALIDATE, int NB>
__device__ __forceinline__ void load_run_keys(const RunElem *(&ptr)[NB], unsigned long long (&key)[NB], unsigned int epoch,
                                              int *status) {
    uint4 v[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k)
        if (ptr[k]) v[k] = load_run_elem_once<VALIDATE>(ptr[k]);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        if (!ptr[k]) continue;
        if (VALIDATE && v[k].w != epoch) v[k] = load_run_elem<VALIDATE>(ptr[k], epoch, status);
        key[k] = ((unsigned long long)v[k].y << 32) | v[k].x;
    }
}
__device__ __forceinline__ unsigned long long elem_key(const uint4 &e) { return ((unsigned long long)e.y << 32) | e.x; }

template <bool VALIDATE>
__device__ __forceinline__ int count_below_run(const RunElem *run, int n, unsigned long long key, unsigned int epoch, int *status) {
    int lo = 0, hi = n;  // first position whose key is >= `key`
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (elem_key(load_run_elem<VALIDATE>(run + mid, epoch, status)) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int count_below_smem(const unsigned long long *win, int n, unsigned long long key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (win[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int kMergeThreads = 256;

// n_in[r] = inliers of the dim = keys below the outlier bit, over all runs (one CTA per dim calls this).
template <bool VALIDATE>
__device__ __forceinline__ void count_inliers(const RunSet &rs, int r, unsigned int epoch, int *__restrict__ flags) {
    __shared__ int s_nin[kMaxRuns];
    const int tid = threadIdx.x;
    if (tid < rs.T)
        s_nin[tid] = count_below_run<VALIDATE>(run_slot(rs.base, rs.R_cap, tid, r), (int)(rs.run_off[tid + 1] - rs.run_off[tid]),
                                               1ull << 63, epoch, rs.status);
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        for (int o = 0; o < rs.T; ++o) c += s_nin[o];
        flags[kFlagNIn + r] = c;
    }
}

// Writes one run element into the sorted columns of dim r at its global position.
__device__ __forceinline__ void place_run_elem(const uint4 &e, int64_t pos, int64_t base, float cabs, float *__restrict__ Xs,
                                               float *__restrict__ As, float *__restrict__ Es, int *__restrict__ perm,
                                               int *__restrict__ flags, int r) {
    const unsigned long long key = elem_key(e);
    const float xs = __uint_as_float(e.z);
    Xs[base + pos] = xs;
    As[base + pos] = sortable_to_float_bits(key_sortable_attr(key));  // NaNs come back as one quiet NaN: only compared
    Es[base + pos] = key_is_outlier(key) ? 1.0f : exp2f(cabs * xs);
    perm[base + pos] = (int)(key & kKeyIdxMask);
    note_nonfinite(xs, flags, r);
    if (key_is_outlier(key) || !(fabsf(cabs * xs) <= kSharedMaxAbsU)) atomicOr(flags + kFlagNeedComplete, 1);
}

// One CTA per 256 consecutive elements of one run (and dim).  Their keys ascend, so inside any other run only the
// window between the positions of the CTA's first and last key can interleave with them: two binary searches per
// other run bound the windows, the windows are staged in shared memory, and every element finds its rank in each
// window there.  Global position = own index in its run + the ranks in all other runs.
// On one GPU the CTA then writes its elements into the sorted columns.  In a sharded step every rank ranks only ITS
// OWN runs (1/G of the searches, and of the L2 traffic of the windows) and stores the positions into every peer's
// position slots; runs_apply_kernel then moves all elements of all runs to their positions on every rank.
template <bool VALIDATE>
__global__ void __launch_bounds__(kMergeThreads)
runs_merge_kernel(RunSet rs, int64_t Bpad, float cabs, float *__restrict__ Xs, float *__restrict__ As,
                  float *__restrict__ Es, int *__restrict__ perm, int *__restrict__ flags, int *__restrict__ mypos,
                  int64_t my_lo, int64_t my_hi, int64_t mypos_stride, int win_cap, PosDest pd, int blk_begin) {
    pdl_trigger();  // (sharded step) the apply kernel may start: it waits for every position by itself
    extern __shared__ __align__(16) unsigned long long dyn_smem[];  // [T][kRunPivots] pivot keys, then [win_cap] window keys
    unsigned long long *piv = dyn_smem;
    unsigned long long *win = dyn_smem + (size_t)rs.T * kRunPivots;
    __shared__ int s_lb[kMaxRuns], s_len[kMaxRuns], s_woff[kMaxRuns + 1];
    __shared__ unsigned long long s_edge[2];
    __shared__ int s_total;
    const int r = blockIdx.y, tid = threadIdx.x;
    const int64_t B = rs.run_off[rs.T];
    const int64_t base = (int64_t)r * Bpad;
    const int n_blocks = rs.blk_off[rs.T];
    const int blk = blk_begin + (int)blockIdx.x;
    if (blk >= n_blocks) {  // padding: |t - s| = 1 and zero gradient for every row, in either tanh form
        const int64_t t = B + (int64_t)(blk - n_blocks) * kMergeThreads + tid;
        if (t < Bpad) {
            Xs[base + t] = ARVAE_PAD_U;
            As[base + t] = ARVAE_PAD_A;
            Es[base + t] = 8.5070592e37f;
            perm[base + t] = -1;
        }
        return;
    }
    const unsigned int epoch = VALIDATE ? (unsigned int)(*rs.epoch_ctr + 1ull) : 1u;
    int t = 0;
    while (blk >= rs.blk_off[t + 1]) ++t;
    const int n_t = (int)(rs.run_off[t + 1] - rs.run_off[t]);
    const int p0 = (blk - rs.blk_off[t]) * kMergeThreads;
    const int cnt = min(kMergeThreads, n_t - p0);
    const RunElem *mine_run = run_slot(rs.base, rs.R_cap, t, r);
    uint4 e = make_uint4(0, 0, 0, 0);
    if (tid < cnt) e = load_run_elem_once<VALIDATE>(mine_run + p0 + tid);
    // pivots (every kPivotStep-th key) of all other runs: the loads of a thread are in flight together
    constexpr int kBatch = 4;
    for (int i0 = tid; i0 < rs.T * kRunPivots; i0 += kBatch * kMergeThreads) {
        const RunElem *ptr[kBatch];
        unsigned long long key[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int i = i0 + k * kMergeThreads;
            const int o = i / kRunPivots, j = i % kRunPivots;
            ptr[k] = nullptr;
            key[k] = ~0ull;
            if (i < rs.T * kRunPivots && o != t && j * kPivotStep < (int)(rs.run_off[o + 1] - rs.run_off[o]))
                ptr[k] = run_slot(rs.base, rs.R_cap, o, r) + kRunCap + j;
        }
        load_run_keys<VALIDATE, kBatch>(ptr, key, epoch, rs.status);
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
            if (i0 + k * kMergeThreads < rs.T * kRunPivots) piv[i0 + k * kMergeThreads] = key[k];
    }
    if (VALIDATE && tid < cnt && e.w != epoch) e = load_run_elem<VALIDATE>(mine_run + p0 + tid, epoch, rs.status);
    const unsigned long long key = elem_key(e);
    if (tid == 0) s_edge[0] = key;
    if (tid == cnt - 1) s_edge[1] = key;
    __syncthreads();
    // Window of every other run that can interleave with this CTA's keys, to pivot granularity: with j pivots below a
    // key K, K's rank in the run lies in [kPivotStep (j - 1), kPivotStep j].
    if (tid < 2 * rs.T) {
        const int o = tid >> 1, which = tid & 1;
        const int n_o = (int)(rs.run_off[o + 1] - rs.run_off[o]);
        const int np = (n_o + kPivotStep - 1) / kPivotStep;
        const int j = o == t ? 0 : count_below_smem(piv + o * kRunPivots, np, s_edge[which]);
        if (which == 0) s_lb[o] = j >= 1 ? kPivotStep * (j - 1) : 0;
        else s_len[o] = min(kPivotStep * j, n_o);  // exclusive end, turned into a length below
    }
    __syncthreads();
    if (tid < 32) {  // lengths and their exclusive prefix (T <= 64: two entries per lane)
        int l0 = 0, l1 = 0;
        const int o0 = 2 * tid, o1 = 2 * tid + 1;
        if (o0 < rs.T && o0 != t) l0 = s_len[o0] - s_lb[o0];
        if (o1 < rs.T && o1 != t) l1 = s_len[o1] - s_lb[o1];
        const int mine = l0 + l1;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        const int excl = incl - mine;
        if (o0 < rs.T) { s_len[o0] = l0; s_woff[o0] = excl; }
        if (o1 < rs.T) { s_len[o1] = l1; s_woff[o1] = excl + l0; }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tid == 0) { s_woff[rs.T] = tot; s_total = tot; }
    }
    __syncthreads();
    const int total = s_total;
    const bool staged = total <= win_cap;
    if (staged) {
        for (int i0 = tid; i0 < total; i0 += kBatch * kMergeThreads) {
            const RunElem *ptr[kBatch];
            unsigned long long wk[kBatch];
            int o = 0;
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int i = i0 + k * kMergeThreads;
                ptr[k] = nullptr;
                if (i < total) {
                    while (i >= s_woff[o + 1]) ++o;  // runs with an empty window (the own run among them) are skipped
                    ptr[k] = run_slot(rs.base, rs.R_cap, o, r) + s_lb[o] + (i - s_woff[o]);
                }
            }
            load_run_keys<VALIDATE, kBatch>(ptr, wk, epoch, rs.status);
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
                if (ptr[k]) win[i0 + k * kMergeThreads] = wk[k];
        }
    }
    __syncthreads();
    if (tid < cnt) {
        int64_t pos = p0 + tid;
        for (int o = 0; o < rs.T; ++o) {
            if (o == t) continue;
            pos += s_lb[o];
            if (staged) pos += count_below_smem(win + s_woff[o], s_len[o], key);
            else pos += count_below_run<VALIDATE>(run_slot(rs.base, rs.R_cap, o, r) + s_lb[o], s_len[o], key, epoch, rs.status);
        }
        const int64_t idx = (int64_t)(key & kKeyIdxMask);
        if (pd.n_dest > 0) {
            const uint2 pe = make_uint2((unsigned int)pos, epoch);
            for (int h = 0; h < pd.n_dest; ++h) pos_slot(pd.base[h], rs.R_cap, t, r)[p0 + tid] = pe;
        } else {
            place_run_elem(e, pos, base, cabs, Xs, As, Es, perm, flags, r);
        }
        if (mypos && idx >= my_lo && idx < my_hi) mypos[(int64_t)r * mypos_stride + (idx - my_lo)] = (int)pos;
    }
    if (blockIdx.x == 0 && pd.n_dest == 0) count_inliers<VALIDATE>(rs, r, epoch, flags);
    // launched overlapped with the sort (sharded step): do not complete before it has, so that "this kernel is
    // complete" keeps implying "everything before it in the stream is complete" for the kernels that follow
    pdl_wait();
}

// A sharded step's second half of the merge: every element of every run of dim r goes to the position its owner
// computed (runs_merge_kernel, stored into this rank's position slots over NVLink).  Same grid as a full merge.
__global__ void __launch_bounds__(kMergeThreads)
runs_apply_kernel(RunSet rs, char *__restrict__ pos_base, int64_t Bpad, float cabs, float *__restrict__ Xs,
                  float *__restrict__ As, float *__restrict__ Es, int *__restrict__ perm, int *__restrict__ flags) {
    const int r = blockIdx.y, tid = threadIdx.x;
    const int64_t B = rs.run_off[rs.T];
    const int64_t base = (int64_t)r * Bpad;
    const int n_blocks = rs.blk_off[rs.T];
    pdl_trigger();
    if ((int)blockIdx.x >= n_blocks) {  // padding, as in the merge
        const int64_t t = B + (int64_t)(blockIdx.x - n_blocks) * kMergeThreads + tid;
        if (t < Bpad) {
            Xs[base + t] = ARVAE_PAD_U;
            As[base + t] = ARVAE_PAD_A;
            Es[base + t] = 8.5070592e37f;
            perm[base + t] = -1;
        }
        pdl_wait();
        return;
    }
    const unsigned int epoch = (unsigned int)(*rs.epoch_ctr + 1ull);
    int t = 0;
    while ((int)blockIdx.x >= rs.blk_off[t + 1]) ++t;
    const int n_t = (int)(rs.run_off[t + 1] - rs.run_off[t]);
    const int i = ((int)blockIdx.x - rs.blk_off[t]) * kMergeThreads + tid;
    if (i < n_t) {
        const RunElem *ep = run_slot(rs.base, rs.R_cap, t, r) + i;
        const uint2 *pp = pos_slot(pos_base, rs.R_cap, t, r) + i;
        uint4 e = load_run_elem_once<true>(ep);  // both loads in flight together
        uint2 pe;
        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(pe.x), "=r"(pe.y) : "l"(pp) : "memory");
        if (e.w != epoch) e = load_run_elem<true>(ep, epoch, rs.status);
        unsigned long long t0 = 0;
        while (pe.y != epoch) {
            if (*reinterpret_cast<volatile int *>(rs.status) != 0) break;
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kShardWaitNs) {
                atomicExch(rs.status, 1);
                break;
            }
            __nanosleep(100);
            asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(pe.x), "=r"(pe.y) : "l"(pp) : "memory");
        }
        if (pe.y == epoch && (int64_t)pe.x < B) place_run_elem(e, (int64_t)pe.x, base, cabs, Xs, As, Es, perm, flags, r);
    }
    if (blockIdx.x == 0) count_inliers<true>(rs, r, epoch, flags);
    pdl_wait();  // launched overlapped with the ranking kernel: complete only after it (and, through it, the sort)
}

// Fills `rs` for runs made of `n_parts` consecutive parts (ranks) of the batch, each cut into runs of kRunCap.
static int fill_run_set(RunSet &rs, const int64_t *part_sizes, int n_parts, int *first_run_of_part /* [n_parts] or null */) {
    rs.T = 0;
    rs.run_off[0] = 0;
    rs.blk_off[0] = 0;
    for (int h = 0; h < n_parts; ++h) {
        if (first_run_of_part) first_run_of_part[h] = rs.T;
        for (int64_t done = 0; done < part_sizes[h]; done += kRunCap) {
            if (rs.T >= kMaxRuns) return -1;
            const int64_t n = part_sizes[h] - done < kRunCap ? part_sizes[h] - done : kRunCap;
            rs.run_off[rs.T + 1] = rs.run_off[rs.T] + n;
            rs.blk_off[rs.T + 1] = rs.blk_off[rs.T] + (int)ceil_div(n, kMergeThreads);
            rs.T++;
        }
    }
    return 0;
}

// Merge CTAs [blk_begin, blk_begin + n_blk) of the run set (n_blk < 0: all of them plus the padding CTAs).
static int launch_runs_merge(const RunSet &rs, int R, int64_t Bpad, float cabs, float *Xs, float *As, float *Es, int *perm,
                             int *flags, int *mypos, int64_t my_lo, int64_t my_hi, int64_t mypos_stride, const PosDest &pd,
                             int blk_begin, int n_blk, cudaStream_t st) {
    const int64_t B = rs.run_off[rs.T];
    // expected window total ~ 256 (T - 1): stage up to 2x that (>= 4096 keys), within the opt-in shared memory
    int win_cap = 2 * kMergeThreads * (rs.T > 1 ? rs.T - 1 : 1);
    if (win_cap < 4096) win_cap = 4096;
    if (win_cap > 24576) win_cap = 24576;
    win_cap += 2 * kPivotStep * rs.T;  // pivot-granular windows are up to two buckets wider per run
    constexpr size_t kMergeSmemMax = 200 * 1024;
    const size_t piv_bytes = sizeof(unsigned long long) * (size_t)rs.T * kRunPivots;
    if (piv_bytes + sizeof(unsigned long long) * (size_t)win_cap > kMergeSmemMax)
        win_cap = (int)((kMergeSmemMax - piv_bytes) / sizeof(unsigned long long));
    const size_t smem = piv_bytes + sizeof(unsigned long long) * (size_t)win_cap;
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
        const int max_smem = 200 * 1024;
        ARVAE_CUDA_TRY(cudaFuncSetAttribute(runs_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        ARVAE_CUDA_TRY(cudaFuncSetAttribute(runs_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set[dev] = true;
    }
    if (n_blk < 0) n_blk = (int)(rs.blk_off[rs.T] + ceil_div(Bpad - B, kMergeThreads));
    if (n_blk == 0) return 0;
    dim3 grid((unsigned)n_blk, (unsigned)R);
    if (rs.epoch_ctr)  // sharded step: overlapped with the sort kernel before it (every load validates itself)
        ARVAE_CUDA_TRY(launch_kernel(runs_merge_kernel<true>, grid, dim3(kMergeThreads), smem, st, true, rs, Bpad, cabs, Xs, As, Es, perm,
                                     flags, mypos, my_lo, my_hi, mypos_stride, win_cap, pd, blk_begin));
    else
        runs_merge_kernel<false><<<grid, kMergeThreads, smem, st>>>(rs, Bpad, cabs, Xs, As, Es, perm, flags, mypos, my_lo, my_hi,
                                                                 mypos_stride, win_cap, pd, blk_begin);
    ARVAE_LAUNCH_CHECK("runs_merge_kernel");
    return 0;
}

// Last CTA of the pair kernel: this rank's exact loss partial -> its header, then flag B at every peer.  What the
// peers pull afterwards (row accumulators, loss partial) lives in THIS GPU's L2; device-scope fences order it before
// the flag stores.
__device__ void shard_pair_kernel_tail(const TilesArgs &a, acc_t *sh /* shared memory, 2 * kDuoThreads, free by now */) {
    ShardHeader *hdr = shard_header(a.shard, a.shard.g);
    const unsigned long long epoch = hdr->epoch + 1;
    if (!last_cta(&hdr->done_pair, gridDim.x)) return;
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
    }
    __syncthreads();
    // Every CTA's row-sum atomics were ordered before its ticket at device scope, and this CTA saw the last ticket: one
    // system-scope fence in each signalling thread (fences are cumulative) orders all of it before the flag a peer sees.
    if ((int)threadIdx.x < a.shard.G) {
        __threadfence_system();
        st_volatile_u64(shard_flags(a.shard, (int)threadIdx.x) + a.shard.g, epoch);
    }
}

// ------------------------------------------------------------------------------------------------------------
// (C) finalize: pull this rank's row sums and every rank's loss partial
// ------------------------------------------------------------------------------------------------------------
// Where the row sum of sample i (local index) of dim r lives: accumulator index and the 1-2 ranks that swept its sorted
// position.  Depends on the plan only, so it is worked out BEFORE the wait for the peers.
struct ShardRowRef {
    int64_t slot;   // rr * kTileRows + position within the row tile
    int h0, h1;
    bool poisoned;  // the reference's float arithmetic gives NaN for this row
};
__device__ __forceinline__ ShardRowRef shard_locate_row(const TilesArgs &a, const ShardView &v, const int *__restrict__ mypos,
                                                        int r, int64_t i) {
    const int64_t pos = mypos[(int64_t)r * v.n_cap + i];
    const int64_t rr = (int64_t)r * a.n_row_tiles + pos / kTileRows;
    const long long T = a.prefix[a.n_rr];
    ShardRowRef ref;
    ref.slot = rr * kTileRows + pos % kTileRows;
    ref.h0 = (int)(owner_of_pos(a.prefix[rr], T, a.G) / v.Gc);
    ref.h1 = (int)(owner_of_pos(a.prefix[rr + 1] - (long long)a.cost8[rr * a.S + a.S - 1], T, a.G) / v.Gc);
    ref.poisoned = row_is_poisoned(a.flags, r, a.Xs[(int64_t)r * a.Bpad + pos]);
    return ref;
}
__device__ __forceinline__ float shard_pull_grad(const ShardView &v, const ShardRowRef &ref, double gscale, bool broken) {
    if (broken || ref.poisoned) return __int_as_float(0x7fc00000);
    acc_t g = 0;
    for (int h = ref.h0; h <= ref.h1; ++h) g += ld_relaxed_sys_s64(shard_acc(v, h) + ref.slot);
    return (float)((double)g * kFixScale * gscale);
}

// Outputs: grad_cols [n, R] (one thread per element), or -- the host-buffer entry -- grad_z [n, Z] with the scatter
// into the latent columns done here (one thread per element of grad_z; it may be host memory mapped into the device).
__global__ void __launch_bounds__(256)
shard_finalize_kernel(TilesArgs a, const int *__restrict__ mypos, int R, double gscale, double lscale,
                      double pad_per_row, float *__restrict__ grad_cols, double *__restrict__ loss_out,
                      float *__restrict__ loss_f32_out, int *__restrict__ flags_rw, float *__restrict__ grad_z, int64_t Z,
                      RegDims dims) {
    const ShardView &v = a.shard;
    ShardHeader *hdr = shard_header(v, v.g);
    pdl_wait();  // everything this rank produced (plan, positions, its pair kernel) is complete
    const unsigned long long epoch = hdr->epoch + 1;
    const int64_t n = v.row_off[v.g + 1] - v.row_off[v.g];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // locate this thread's row sum(s) first, then wait for the peers, then pull over NVLink
    ShardRowRef ref;
    ref.slot = -1;
    ref.h0 = 0; ref.h1 = -1; ref.poisoned = false;
    const bool broken_early = *reinterpret_cast<volatile int *>(&hdr->status) != 0;  // an earlier wait of this step gave up: the
                                                                                    // positions may be garbage, do not use them
    if (broken_early) {
    } else if (grad_z) {  // over n * Z, column fastest: the (at most one, in practice) regularised dim of this latent column
        if (idx < n * Z) {
            const int zc = (int)(idx % Z);
            for (int r = 0; r < R; ++r)
                if (dims.zcol[r] == zc && ref.slot < 0) ref = shard_locate_row(a, v, mypos, r, idx / Z);
        }
    } else if (grad_cols && idx < n * R) {  // over n * R, dim fastest (coalesced stores)
        ref = shard_locate_row(a, v, mypos, (int)(idx % R), idx / R);
    }
    shard_wait(v, epoch);
    const bool broken = *reinterpret_cast<volatile int *>(&hdr->status) != 0;  // a wait gave up: positions may be garbage
    if (grad_z) {
        if (idx < n * Z) {
            float g = ref.slot >= 0 ? shard_pull_grad(v, ref, gscale, broken) : (broken ? __int_as_float(0x7fc00000) : 0.0f);
            const int zc = (int)(idx % Z);
            bool first = true;
            for (int r = 0; r < R; ++r) {  // a latent column regularised by several attributes (not a reference configuration)
                if (dims.zcol[r] != zc) continue;
                if (!first && !broken) g += shard_pull_grad(v, shard_locate_row(a, v, mypos, r, idx / Z), gscale, broken);
                first = false;
            }
            grad_z[idx] = g;
        }
    } else if (grad_cols && idx < n * R) {
        grad_cols[idx] = shard_pull_grad(v, ref, gscale, broken);
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
    // the step is complete on this rank once every CTA is past its wait and its pulls: clear the per-step flags for
    // the next step's merge and advance the epoch
    if (last_cta(&hdr->done_fin, gridDim.x)) {
        if ((int)threadIdx.x < kFlagClearInts) flags_rw[threadIdx.x] = 0;
        if (threadIdx.x == 0) {
            hdr->done_fin = 0;
            hdr->epoch = epoch;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
int shard_set_wait_ms(long long ms) {
    const unsigned long long ns = (unsigned long long)(ms > 0 ? ms : 1) * 1000000ull;
    ARVAE_CUDA_TRY(cudaMemcpyToSymbol(kShardWaitNs, &ns, sizeof(ns)));
    return 0;
}

static int shard_runs_cap(int64_t n_cap, int G) {
    const int64_t t = (int64_t)G * ceil_div(n_cap, kRunCap);
    return (int)(t < kMaxRuns ? t : kMaxRuns);
}

size_t shard_comm_bytes(int64_t n_cap, int R_cap, int G, ShardCtx *fill) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    take(sizeof(ShardHeader));
    const size_t fb = take(sizeof(unsigned long long) * kMaxShardRanks);
    const int runs_cap = shard_runs_cap(n_cap, G);
    const size_t orn = take(sizeof(RunElem) * (size_t)runs_cap * R_cap * kRunSlotElems);
    const size_t opo = take(sizeof(uint2) * (size_t)runs_cap * R_cap * kRunCap);
    const int64_t tiles = ceil_div((int64_t)G * n_cap, kTileRows);
    const size_t oa = take(sizeof(acc_t) * (size_t)R_cap * tiles * kTileRows);
    if (fill) {
        fill->off_flagB = fb; fill->off_runs = orn; fill->off_pos = opo; fill->off_acc = oa; fill->runs_cap = runs_cap;
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
    v.off_flagB = C.off_flagB; v.off_runs = C.off_runs; v.off_acc = C.off_acc;
    const int64_t B = v.row_off[G], n_local = S.n_all[g];
    if (S.R < 1 || S.R > C.R_cap || B < 1 || B > (int64_t)kKeyIdxMask) {
        set_error("shard step: R=%d (capacity %d), B=%lld out of range", S.R, C.R_cap, (long long)B);
        return ARVAE_E_BADARG;
    }
    RunSet rs;
    memset(&rs, 0, sizeof(rs));
    int first_run[kMaxShardRanks];
    if (fill_run_set(rs, S.n_all, G, first_run) != 0 || rs.T > C.runs_cap) {
        set_error("shard step: more than %d sorted runs", C.runs_cap);
        return ARVAE_E_BADARG;
    }
    rs.R_cap = C.R_cap;
    rs.base = C.comm + C.off_runs;
    rs.epoch_ctr = &reinterpret_cast<ShardHeader *>(C.comm)->epoch;
    rs.status = &reinterpret_cast<ShardHeader *>(C.comm)->status;
    const int phases = S.phases ? S.phases : 15;
    const SortedLayout L = sorted_layout(B, B, S.R, sm_count(), false);
    if (L.bytes > C.off_mypos) {
        set_error("shard step: workspace too small");
        return ARVAE_E_WORKSPACE;
    }
    char *ws = C.ws;
    int *flags = reinterpret_cast<int *>(ws + L.off_flags);  // cleared at creation and by every finalize
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
    a.n_in = flags + kFlagNIn;
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
    a.dbg_times = getenv("ARVAE_DEBUG_TIMES") ? reinterpret_cast<unsigned long long *>(ws + L.off_dbg) : nullptr;
    a.shard = v;
    a.dual = 1;
    const bool want_grad = S.grad_cols_out != nullptr || S.grad_z_out != nullptr;

    if (phases & 1) {
        KeySpec spec;
        spec.lab = S.lab; spec.lrs = S.lrs; spec.lcs = S.lcs;
        spec.z = S.z; spec.zrs = S.zrs; spec.zcs = S.zcs;
        spec.fsign = fsign; spec.cabs = cabs; spec.segment = 1;
        spec.idx_offset = v.row_off[g];
        spec.dims = S.dims;
        RunDest dest;
        memset(&dest, 0, sizeof(dest));
        dest.n_dest = G; dest.R_cap = C.R_cap;
        for (int h = 0; h < G; ++h) dest.base[h] = C.peer[h] + C.off_runs;
        timeline_mark(st, "begin");
        int rc = run_chunk_sort(spec, S.R, n_local, first_run[g], dest, rs.epoch_ctr, st);
        if (rc) return rc;
        timeline_mark(st, "sort+publish");
    }
    if (phases & 2) {
        timeline_mark(st, "begin B");
        // this rank's runs only: positions of their elements -> every peer's position slots
        PosDest pd;
        memset(&pd, 0, sizeof(pd));
        pd.n_dest = G;
        for (int h = 0; h < G; ++h) pd.base[h] = C.peer[h] + C.off_pos;
        const int run_end = g + 1 < G ? first_run[g + 1] : rs.T;
        int rc = launch_runs_merge(rs, S.R, L.Bpad, cabs, const_cast<float *>(a.Xs), const_cast<float *>(a.As),
                                   const_cast<float *>(a.Es), perm, flags, mypos, v.row_off[g], v.row_off[g + 1], C.n_cap, pd,
                                   rs.blk_off[first_run[g]], rs.blk_off[run_end] - rs.blk_off[first_run[g]], st);
        if (rc) return rc;
        timeline_mark(st, "rank own runs");
    }
    if (phases & 4) {
        timeline_mark(st, "begin C");
        {   // every element of every run to its position (waits for the peers' elements and positions)
            dim3 grid((unsigned)(rs.blk_off[rs.T] + ceil_div(L.Bpad - B, kMergeThreads)), (unsigned)S.R);
            ARVAE_CUDA_TRY(launch_kernel(runs_apply_kernel, grid, dim3(kMergeThreads), 0, st, true, rs, C.comm + C.off_pos, L.Bpad, cabs,
                                         const_cast<float *>(a.Xs), const_cast<float *>(a.As), const_cast<float *>(a.Es), perm, flags));
            ARVAE_LAUNCH_CHECK("runs_apply_kernel");
        }
        timeline_mark(st, "wait+apply");
        // The plan kernel also clears the row accumulators.  That is safe only now: every peer has published this
        // step's runs (the merge saw them), hence finished pulling the previous step's row sums.
        int *combo_cost = reinterpret_cast<int *>(ws + L.off_combo);
        ARVAE_CUDA_TRY(launch_kernel(plan_classes_kernel, dim3((unsigned)L.n_rr), dim3(256), 0, st, true, a, combo_cost, want_grad ? 1 : 0,
                                     reinterpret_cast<unsigned int *>(flags + kFlagTicket)));
        ARVAE_LAUNCH_CHECK("plan_classes_kernel");
        timeline_mark(st, "plan");
        profile_begin(st);
        launch_tiles(a, (int)Gc, want_grad, false, st);
        profile_end(st);
        ARVAE_LAUNCH_CHECK("reg_tiles_kernel");
        timeline_mark(st, "pairs");
    }
    if (phases & 8) {
        timeline_mark(st, "begin D");
        RegProblem P;
        P.B = B; P.gamma = S.gamma; P.factor = S.factor;
        double lscale, gscale, pad_per_row;
        reg_scales(P, L.Bpad, lscale, gscale, pad_per_row);
        const int64_t work = S.grad_z_out ? n_local * S.grad_z_cols : n_local * S.R;
        ARVAE_CUDA_TRY(launch_kernel(shard_finalize_kernel, dim3((unsigned)(work > 0 ? ceil_div(work, 256) : 1)), dim3(256), 0, st, true,
                                     a, mypos, S.R, gscale, lscale, pad_per_row, S.grad_cols_out, S.loss_out, S.loss_f32_out, flags,
                                     S.grad_z_out, S.grad_z_cols, S.dims));
        ARVAE_LAUNCH_CHECK("shard_finalize_kernel");
        timeline_mark(st, "wait+finalize");
    }
    return 0;
}
