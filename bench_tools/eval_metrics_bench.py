#!/usr/bin/env python
"""Pairwise-rank evaluation metrics (Corr_score + SAP_score, utils/evaluation.py:146-219) at the reference's
evaluation size (201 batches of 128 = 25 728 samples, 16 codes x 6 attributes): the CUDA path through the C ABI
vs the oracle port on the host (vectorised numpy, i.e. already far faster than the reference's Z x A scipy calls).

First checks every golden fixture (reference outputs) so that a timing is never reported for a wrong result.
Writes one JSON line per section to stdout and, when gpurun_out/ exists, to gpurun_out/eval_metrics_bench.json."""
import glob, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from arvae_b200 import _lib, evaluation
from test_eval_metrics import check_against_golden, oracle_all

lines = []


def emit(obj):
    lines.append(obj)
    print(json.dumps(obj), flush=True)


parity = {}
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "eval_*.npz"))):
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    try:
        check_against_golden(evaluation.rank_metrics(g["mus"], g["ys"]), g)
        parity[name] = "ok"
    except Exception as e:  # noqa: BLE001 -- report every fixture, then fail
        parity[name] = f"FAIL: {type(e).__name__}: {str(e)[:200]}"
emit({"section": "parity_vs_reference_goldens", "results": parity})

g = np.load(os.path.join(ROOT, "tests", "golden", "eval_mnist_n25728.npz"))
mus_h, ys_h = g["mus"], g["ys"]
mus, ys = torch.from_numpy(mus_h).cuda(), torch.from_numpy(ys_h).cuda()
for _ in range(3):
    evaluation.rank_metrics(mus, ys)
_lib.load().arvae_launch_count(1)
evaluation.rank_metrics(mus, ys)
launches = int(_lib.load().arvae_launch_count(1))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record()
for _ in range(steps):
    evaluation.rank_metrics(mus, ys)          # device-resident inputs; includes the 3 KB result read-back
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / steps
t = time.perf_counter()
for _ in range(steps):
    evaluation.rank_metrics(mus_h, ys_h)      # numpy in, numpy out: what the reference's trainers would call
e2e_ms = (time.perf_counter() - t) * 1e3 / steps
t = time.perf_counter()
with np.errstate(all="ignore"):
    oracle_all(mus_h, ys_h)
cpu_ms = (time.perf_counter() - t) * 1e3
emit({"section": "timing", "N": int(mus_h.shape[0]), "Z": int(mus_h.shape[1]), "A": int(ys_h.shape[1]),
      "gpu_ms_device_inputs": gpu_ms, "e2e_ms_numpy_inputs": e2e_ms, "launches_per_call": launches,
      "cpu_oracle_numpy_ms": cpu_ms, "algorithmic_bytes": int(mus_h.nbytes + ys_h.nbytes),
      "note": "reference itself: Z*A scipy.stats.spearmanr + np.cov calls, ~0.87 s here (see tests/golden/make_golden_eval.py)"})
out_dir = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(out_dir):
    with open(os.path.join(out_dir, "eval_metrics_bench.json"), "w") as f:
        for obj in lines:
            f.write(json.dumps(obj) + "\n")
if any(v != "ok" for v in parity.values()):
    sys.exit(1)
