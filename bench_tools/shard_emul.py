#!/usr/bin/env python
"""One rank's share of a G-GPU sharded step, timed on ONE GPU.

G virtual ranks of the NVLink-sharded step (arvae_b200.distributed.LocalShardGroup) run in one process; the step is
driven phase by phase (A sort + publish, B rank own runs, C apply + plan + pair kernel, D finalize) for every rank, and the phases of ONE rank
are bracketed with CUDA events.  What a real rank executes is exactly A + B + C of its own (plus NVLink latency and
the wait for the slowest peer), so max(A) + max(B) + max(C) + max(D) is the device time a G-GPU step needs per rank -- the
fixed costs of the multi-GPU path can be tuned on a 1-GPU box.

    python bench_tools/shard_emul.py [--world 8] [--batch 65536] [--steps 20]
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--workload", default="c4_mnist_b65536")
    args = ap.parse_args()
    import arvae_b200
    from arvae_b200 import _lib, synth
    from arvae_b200 import distributed as adist

    lib = _lib.load()
    c = synth.make_case(args.workload, args.batch)
    B, G = c["B"], args.world
    dims = tuple(c["reg_dims"])
    n = B // G
    z, lab = c["z"].cuda(), c["labels"].cuda()
    zp = [z[g * n:(g + 1) * n] for g in range(G)]
    lp = [lab[g * n:(g + 1) * n] for g in range(G)]
    n_all = [n] * G
    grp = adist.LocalShardGroup(G, n, len(dims))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def ev():
        return torch.cuda.Event(enable_timing=True)

    times = {1: [], 2: [], 4: [], 8: []}
    pair_ms = []
    for it in range(args.steps + 3):
        for phase in (1, 2, 4, 8):
            rec = []
            for g, h in enumerate(grp.ranks):
                if g == 0:
                    flush.zero_()
                    if phase == 4:
                        lib.arvae_profile_enable(1)
                e0, e1 = ev(), ev()
                e0.record()
                h.step(zp[g], lp[g], dims, dims, n_all, c["gamma"], c["delta"], True, phase)
                e1.record()
                rec.append((e0, e1))
                if g == 0 and phase == 4:
                    import ctypes
                    ks, kn = ctypes.c_float(), ctypes.c_int()
                    lib.arvae_profile_pair_kernel_ms(ctypes.byref(ks), ctypes.byref(kn))
                    lib.arvae_profile_enable(0)
                    if it >= 3 and kn.value:
                        pair_ms.append(ks.value / kn.value)
            torch.cuda.synchronize()
            if it >= 3:
                times[phase].append(rec[0][0].elapsed_time(rec[0][1]))
    # single-GPU step for comparison
    single = []
    for it in range(args.steps + 3):
        flush.zero_()
        e0, e1 = ev(), ev()
        e0.record()
        arvae_b200.reg_loss_rows(z, lab, dims, c["gamma"], c["delta"], 0, B, algo=arvae_b200.ALGO_SORTED)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            single.append(e0.elapsed_time(e1))
    # warm timeline of rank 0 (events between its launches; no L2 flush inside the step)
    import ctypes
    buf = ctypes.create_string_buffer(4096)
    tl = {}
    for it in range(5):
        for phase in (1, 2, 4, 8):
            for g, h in enumerate(grp.ranks):
                if g == 0:
                    lib.arvae_timeline_enable(1)
                h.step(zp[g], lp[g], dims, dims, n_all, c["gamma"], c["delta"], True, phase)
                if g == 0:
                    lib.arvae_timeline_report(buf, 4096)
                    lib.arvae_timeline_enable(0)
                    for item in buf.value.decode().split(";"):
                        if item:
                            k, v = item.rsplit(":", 1)
                            tl.setdefault(k, []).append(float(v))
            torch.cuda.synchronize()
    timeline = {k: round(statistics.median(v) * 1e3, 1) for k, v in tl.items()}
    tl1 = {}
    for it in range(5):
        lib.arvae_timeline_enable(1)
        arvae_b200.reg_loss_rows(z, lab, dims, c["gamma"], c["delta"], 0, B, algo=arvae_b200.ALGO_SORTED)
        lib.arvae_timeline_report(buf, 4096)
        lib.arvae_timeline_enable(0)
        for item in buf.value.decode().split(";"):
            if item:
                k, v = item.rsplit(":", 1)
                tl1.setdefault(k, []).append(float(v))
    timeline1 = {k: round(statistics.median(v) * 1e3, 1) for k, v in tl1.items()}
    cta = None
    if os.environ.get("ARVAE_DEBUG_TIMES"):
        import numpy as np
        lib.arvae_shard_debug_times.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
        arr = np.zeros((148, 2), dtype=np.uint64)
        lib.arvae_shard_debug_times(grp.ranks[0].ctx, B, len(dims), arr.ctypes.data_as(ctypes.c_void_p), 148)
        t = arr.astype(np.int64)
        t = t[t[:, 1] > 0]
        t0 = t[:, 0].min()
        st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
        cta = {"n": int(len(t)), "start_spread_us": float(st.max()), "end_min_p10_med_p90_max_us": [float(en.min()), float(np.percentile(en, 10)), float(np.median(en)), float(np.percentile(en, 90)), float(en.max())],
               "mean_busy_us": float(np.mean(en - st))}
    med = {k: statistics.median(v) for k, v in times.items()}
    out = {"world": G, "B": B, "rank0_publish_ms": med[1], "rank0_rank_own_ms": med[2], "rank0_apply_plan_pair_ms": med[4], "rank0_finalize_ms": med[8],
           "rank0_pair_kernel_ms": statistics.median(pair_ms) if pair_ms else None,
           "rank0_step_ms": med[1] + med[2] + med[4] + med[8], "single_gpu_fwd_ms": statistics.median(single),
           "fixed_ms": med[1] + med[2] + med[4] + med[8] - (statistics.median(pair_ms) if pair_ms else 0.0),
           "rank0_timeline_us": timeline, "single_gpu_timeline_us": timeline1, "rank0_cta_times": cta,
           "note": "flushed L2 before every timed phase; events bracket the launches of rank 0 only (cold start of each phase)"}
    print(json.dumps(out))
    grp.close()


if __name__ == "__main__":
    main()
