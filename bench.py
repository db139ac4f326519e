#!/usr/bin/env python
"""bench.py -- headline benchmark of the attribute-regularization hot path on B200.

Metric (BASELINE.json): fused reg-loss forward+backward throughput in Gpairs/s, pairs = B^2 * R
ordered pairs per step, on the "MnistRESNET large-batch" config C4 (B=65536, Z=16, R=6, gamma=10,
delta=1), synthetic Morpho-MNIST-shaped labels.  A step is one complete op: pack + pair kernel +
epilogue + gradient scatter (loss and dL/dz both produced).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

N>1 is strong scaling of the same global batch: each rank owns B/N rows (its own samples); the exchange
is done by the kernels themselves over NVLink peer memory (arvae_b200.distributed.ShardComm, csrc/reg_shard.cuh):
per-rank sort, sorted runs stored into every peer, merge, 1/N of the single-GPU plan swept per rank, row sums and
loss partials pulled from peers.  No NCCL call inside a step (--transport nccl selects the all-gather / all-reduce
form instead).

`value` is computed from the MEDIAN of the K per-step device times (each step has its own CUDA event pair; max over
ranks of the per-rank medians); `ms_per_step_mean` is the mean of the same K intervals.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

WORKLOAD = "c4_mnist_b65536"
MUFU_LANES_PER_CLK_SM = 16.0     # sm_100 MUFU issue rate; confirmed by bench_tools/pipe_rates.cu (profiles/)
# MUFU per evaluated pair: 2 (EX2 + RCP on the latent difference, SURVEY App. B's cheapest admissible
# dense form) on the dense path; 1 (RCP only, E_j/(E_i+E_j) with E = 2^u precomputed per element) where
# the attribute-sorted path's range guard holds.  Measured per run via arvae_b200.mufu_per_pair().
# The one-MUFU constant-sign loop (reg_sorted.cu: loop_const) works on column pairs with packed FP32 instructions, in two
# builds of the pair kernel (cuobjdump -sass of the loops):
#  * common case, every |2 f log2(e) z| <= 31 (kSharedMaxAbsU): two column pairs share one reciprocal, 1 / (a b), and only
#    the sums q_a + q_b and q_a^2 + q_b^2 are formed, from column-pair sums and products staged in shared memory; one of
#    the eight quads of a 4 x 8 pair group takes its reciprocals from a packed Newton iteration on the FMA pipe -- 580
#    instructions per 256 pairs: 112 MUFU.RCP, 296 FFMA2, 128 FADD2 (two FP32 operations each), 17 IADD3, 16 LDS.128,
#    11 others;
#  * complete build (outliers or latents beyond that range present): q = 1 / (1 + E_j F_i) per pair, 6 of the 16 slots of a
#    4 x 8 pair group take their reciprocals from a packed Newton iteration on the FMA pipe -- 914 instructions per
#    256 pairs: 160 MUFU.RCP, 496 FFMA2, 128 FADD2, 97 IADD3, 16 LDS.128, 17 others.
# tests/test_bench_contract.py holds these counts against cuobjdump -sass of the built library.
# Used for the roofs of the ACTUAL mix.
SHARED_MAX_ABS_U = 31.0
SHARED_LOOP = {"mufu_per_inlier_pair": 112.0 / 256.0, "instr_per_pair": 580.0 / 256.0, "fp32_ops_per_pair": (296 + 128) * 2 / 256.0,
               "sass": "580 instructions per 256 pairs (112 MUFU.RCP, 296 FFMA2, 128 FADD2, 17 IADD3, 16 LDS.128)",
               "what": "shared-reciprocal build (every |u| <= 31): one reciprocal per two pairs from staged column-pair sums and "
                       "products, 1 of 8 quads on packed Newton reciprocals"}
PLAIN_LOOP = {"mufu_per_inlier_pair": 160.0 / 256.0, "instr_per_pair": 914.0 / 256.0, "fp32_ops_per_pair": (496 + 128) * 2 / 256.0,
              "sass": "914 instructions per 256 pairs (160 MUFU.RCP, 496 FFMA2, 128 FADD2, 97 IADD3, 16 LDS.128)",
              "what": "complete build: 6 of 16 slots per pair group take packed Newton reciprocals on the FMA pipe"}
ISSUE_LANES_PER_CLK_SM = 128.0   # 4 schedulers x 32 lanes
FP32_LANES_PER_CLK_SM = 128.0    # FMA pipe (FFMA2 / FADD2 / FMUL2 measured at 123 FP32 operations/clk/SM, profiles/r2_pipe_rates.json)
ALGO_BYTES_PER_ROWCOL = 4        # float32 per latent / label / gradient element
# (B, R, n_gpus, algo) -> dram__bytes_read.sum + dram__bytes_write.sum of one pair-kernel launch (ncu, profiles/)
NCU_DRAM_BYTES_PER_LAUNCH = {(65536, 6, 1, 0): 8412928 + 0}


def load_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            inside = (t0 - 0.05) <= t <= (t1 + 0.15)
            try:
                if inside:
                    sm.append(float(parts[1]))
                    power.append(float(parts[3]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than one sample: use everything we saw
            for t, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's torch-CPU op chain (oracle/torch_port.py) on a bounded row sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(case, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times forward+backward of the reference op chain on rows [0, n) x all B columns x R dims.
    Returns (Gpairs/s, dict describing the sample)."""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    z, labels = case["z"], case["labels"]
    B, R = case["B"], len(case["reg_dims"])

    def run(n_rows):
        t = time.perf_counter()
        torch_port.reg_loss_dims_rows_fwdbwd(z, labels, case["reg_dims"], case["gamma"], case["delta"], (0, n_rows))
        return time.perf_counter() - t

    n = min(64, B)
    run(n)                       # touch pages, spin up the thread pool
    dt = run(n)
    rate = n * B * R / max(dt, 1e-9)
    # grow the sample toward the time target; cap by memory (~10 live [n, B] float32 temporaries)
    mem_cap = max(64, int(24e9 / (B * 4 * 12)))
    n = int(min(B, mem_cap, max(n, rate * target_seconds / (B * R))))
    for _ in range(warmup):
        run(n)
    times = [run(n) for _ in range(max(1, steps))]
    dt = statistics.median(times)
    pairs = float(n) * B * R
    return pairs / dt / 1e9, {
        "cores": threads, "kind": "port",
        "sample": f"rows [0,{n}) x all {B} columns x {R} dims of {WORKLOAD} ({pairs/1e9:.3f} Gpairs per step, "
                  f"{len(times)} steps, median {dt:.2f} s); torch {torch.__version__} CPU op chain of "
                  f"oracle/torch_port.py (= reference utils/trainer.py:378-403 + autograd), {threads} threads",
        "ms_per_step": dt * 1e3,
    }


def _import_reference_trainer():
    """The reference's own Trainer (utils/trainer.py) when its tree is on this box (the dev container), else None."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return None
    import types
    for name in ("tensorboardX", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        from utils.trainer import Trainer
        return Trainer
    except Exception:
        return None


def cpu_full_dense_chain(B: int, steps: int, budget_s: float):
    """The COMPLETE dense op chain (all B x B temporaries, forward + autograd backward, the trainers' per-dim loop) at
    a batch size that fits in host memory -- the unmodified reference when /root/reference exists, else the port."""
    from arvae_b200 import synth
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    case = synth.make_case(WORKLOAD, B)
    Trainer = _import_reference_trainer()
    fn = Trainer.compute_reg_loss if Trainer is not None else torch_port.compute_reg_loss
    kind = "reference" if Trainer is not None else "port (the reference tree is not on this box)"
    z0, labels = case["z"], case["labels"]

    def run():
        z = z0.clone().requires_grad_(True)
        t = time.perf_counter()
        loss = 0.0
        for dim in case["reg_dims"]:  # imagevae/image_vae_trainer.py:171-180
            loss = loss + fn(z, labels[:, dim], dim, gamma=case["gamma"], factor=case["delta"])
        loss.backward()
        return time.perf_counter() - t

    t_first = run()
    n = int(max(1, min(steps, (budget_s - t_first) // max(t_first, 1e-3))))
    times = [run() for _ in range(n)]
    dt = statistics.median(times)
    R = len(case["reg_dims"])
    return {"B": B, "Z": case["Z"], "R": R, "value": float(B) * B * R / dt / 1e9, "unit": "Gpairs/s",
            "ms_per_step": dt * 1e3, "steps": n, "cores": threads, "kind": kind, "same_config_except_B": True,
            "note": f"complete dense chain at the largest batch that fits in host memory; {WORKLOAD} itself (B=65536) "
                    "needs 16 GiB per B x B temporary and cannot run"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from arvae_b200 import synth
    case = synth.make_case(WORKLOAD, args.batch or None)
    B, R = case["B"], len(case["reg_dims"])
    n_runs = max(1, args.steps + args.warmup)
    value, info = cpu_reference_sample(case, min(20.0, 110.0 / n_runs), steps=args.steps, warmup=args.warmup)
    try:
        dense = cpu_full_dense_chain(8192, steps=3, budget_s=60.0)
    except Exception as e:  # pragma: no cover
        dense = {"B": 8192, "value": None, "note": f"failed: {e}"}
    kind = info["kind"] if os.path.isdir("/root/reference") else "port (the reference tree is not on this box)"
    line = {
        "impl": "reference", "metric": "reg_loss_fwd_bwd_throughput", "value": value, "unit": "Gpairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B": B, "Z": case["Z"], "R": R, "gamma": case["gamma"],
                   "delta": case["delta"], "pairs_per_step": float(B) * B * R},
        "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": info["cores"], "kind": kind,
                         "sample": info["sample"]},
        "full_dense_chain": dense,
        "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def parity_block(case, grad_local, r0, n_local, loss_val, dev, world, dist):
    """Gradient rows of 16 samples of this rank vs float64; max over ranks.  Also: is the loss bit-identical on all ranks."""
    dims = list(case["reg_dims"])
    B = case["B"]
    z = case["z"].to(dev, torch.float64)
    a = case["labels"].to(dev)
    rows = torch.linspace(0, n_local - 1, 16).long().unique().to(dev)
    gi = rows + r0
    worst = 0.0
    for d in dims:
        x, ad = z[:, d], a[:, d]
        t = torch.tanh(case["delta"] * (x[gi, None] - x[None, :]))
        s = torch.sign(ad[gi, None] - ad[None, :]).double()
        g_ref = (2.0 * case["gamma"] * case["delta"] / (float(B) * B)) * (torch.sign(t - s) * (1.0 - t * t)).sum(1)
        # the gate of tests/util.py: error against the column's magnitude (taken over the sampled rows)
        err = (grad_local[rows, d].double() - g_ref).abs().max() / g_ref.abs().max().clamp_min(1e-300)
        worst = max(worst, float(err))
    stats = torch.tensor([worst, loss_val, -loss_val], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    return {"grad_rows_max_rel_err_vs_f64": float(stats[0]), "rows_checked_per_rank": int(rows.numel()),
            "loss_bit_identical_on_all_ranks": bool(float(stats[1]) == -float(stats[2])), "gate": 1e-5,
            "ok": bool(float(stats[0]) <= 1e-5 and float(stats[1]) == -float(stats[2]))}


def small_batch_extra(dev):
    """the whole latent-loss head (reparametrize + KLD + reg loss, forward AND backward: one
    kernel launch each) on the reference's real configurations.
    device_us: device time per forward+backward, from 20 back-to-back replays of ONE CUDA graph holding both launches
               (and autograd's gradient accumulation), so that no host latency sits between the kernels;
    eager_us:  the same step issued eagerly through the Python API (autograd dispatch included), one step per event
               pair -- host-bound at these sizes."""
    import arvae_b200
    from arvae_b200 import graphs, synth
    out = {}
    for name, beta, cap in (("c1_mnist_b64", 4.0, 0.0), ("c2_dsprites_b4096", 4.0, 0.0), ("c3_measure_b2048", 0.001, 0.0)):
        c = synth.make_case(name)
        B, Z, dims = c["B"], c["Z"], tuple(c["reg_dims"])
        loc0, log_std0, eps0 = synth.make_latent_head(B, Z, 77)
        loc = loc0.to(dev).requires_grad_(True)
        scale = torch.exp(log_std0).to(dev).requires_grad_(True)
        eps, lab = eps0.to(dev), c["labels"].to(dev)

        def step():
            z, kld, reg = arvae_b200.reparam_kld_reg(loc, scale, eps, lab, dims, beta, cap, c["gamma"], c["delta"])
            (kld + reg).backward()
            return kld, reg

        res = {"B": B, "Z": Z, "R": len(dims), "pairs": float(B) * B * len(dims), "launches_fwd": 1, "launches_bwd": 1}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                loc.grad = scale.grad = None
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        ts = []
        for _ in range(30):
            loc.grad = scale.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res["eager_us"] = statistics.median(ts)
        try:
            graph = torch.cuda.CUDAGraph()
            loc.grad = scale.grad = None
            with graphs.quiet_gc(), torch.cuda.graph(graph):
                kld, reg = step()
            reps = 20
            ts = []
            for _ in range(12):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    graph.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / reps)
            res["device_us"] = statistics.median(ts[2:])
            res["gpairs_per_s_device"] = res["pairs"] / (res["device_us"] * 1e-6) / 1e9
            res["graph_losses"] = [float(kld.detach()), float(reg.detach())]
        except Exception as e:  # pragma: no cover
            res["graph_error"] = str(e)
        out[name] = res
    return out


def _claim_stdout():
    """Keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner to stdout) write to
    fd 1, so fd 1 is pointed at stderr for the whole run and the JSON goes to a private copy of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 dense, 2 sorted")
    ap.add_argument("--batch", type=int, default=0, help="override B (parity/scaling sweeps; 0 = config C4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small-batch", action="store_true", help="skip extra.small_batch (the C1-C3 head timings)")
    ap.add_argument("--transport", default="nvlink", choices=["nvlink", "nccl"],
                    help="N>1: nvlink = peer-memory exchange inside the kernels (ShardComm); nccl = all-gather + all-reduce")
    ap.add_argument("--graph", type=int, default=0,
                    help="1: replay the single-GPU device-resident step as CUDA graphs (measured: no gain at C4, and the "
                         "library's per-kernel timing hooks and launch counter do not see replays); default 0 = eager")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import arvae_b200
    from arvae_b200 import _lib, synth
    from arvae_b200 import distributed as adist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    case = synth.make_case(WORKLOAD, args.batch or None)
    B, Z, R = case["B"], case["Z"], len(case["reg_dims"])
    A = case["labels"].shape[1]
    dims = case["reg_dims"]
    gamma, delta = case["gamma"], case["delta"]
    assert B % world == 0
    n_local = B // world
    r0 = rank * n_local
    pairs = float(B) * B * R

    z_host = case["z"][r0:r0 + n_local].contiguous().pin_memory()
    lab_host = case["labels"][r0:r0 + n_local].contiguous().pin_memory()
    z_dev = z_host.to(dev)
    lab_dev = lab_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    comm = None
    if world > 1 and args.transport == "nvlink":
        comm = adist.ShardComm(n_local, R)
    n_all = _lib.i64_array([n_local] * world)

    graphed = None
    if args.graph and world == 1:  # (capturing the NCCL collectives of the sharded step deadlocked here: not offered)
        try:
            from arvae_b200 import graphs
            graphed = graphs.graphed_reg_loss(n_local, Z, A, dims, gamma, delta, device=dev, algo=args.algo)
        except Exception as e:  # capture not possible here: run eagerly and say so
            print(f"bench.py: CUDA-graph capture failed ({e}); running eagerly", file=sys.stderr)
            graphed = None

    def step_device():
        """One step with inputs resident in HBM: loss + dL/dz for this rank's rows."""
        z = z_dev.detach().requires_grad_(True)
        if graphed is not None:
            loss = graphed(z, lab_dev)
        elif world == 1:
            loss = arvae_b200.reg_loss_fused(z, lab_dev, dims, gamma, delta, algo=args.algo)
        else:
            loss = adist.reg_loss_sharded(z, lab_dev, dims, gamma, delta, algo=args.algo, comm=comm)
        loss.backward()
        return loss, z.grad

    grad_host = torch.empty((n_local, Z), dtype=torch.float32).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e():
        """The same step through the public API with HOST buffers: H2D of this step's inputs from pinned
        memory, the op, D2H of the loss and the gradient."""
        if world == 1:
            loss_c = ctypes.c_float()
            rc = lib.arvae_reg_loss_host_f32(ctypes.c_void_p(z_host.data_ptr()), n_local, Z,
                                             ctypes.c_void_p(lab_host.data_ptr()), A, _lib.i32_array(dims),
                                             _lib.i32_array(dims), R, gamma, delta, args.algo, ctypes.byref(loss_c),
                                             ctypes.c_void_p(grad_host.data_ptr()),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "arvae_reg_loss_host_f32")
            return loss_c.value
        if comm is not None:
            loss_c = ctypes.c_float()
            rc = lib.arvae_shard_reg_loss_host_f32(comm.h.ctx, ctypes.c_void_p(z_host.data_ptr()), Z,
                                                   ctypes.c_void_p(lab_host.data_ptr()), A, _lib.i32_array(dims),
                                                   _lib.i32_array(dims), R, n_all, gamma, delta, ctypes.byref(loss_c),
                                                   ctypes.c_void_p(grad_host.data_ptr()),
                                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "arvae_shard_reg_loss_host_f32")
            return loss_c.value
        z = z_host.to(dev, non_blocking=True).requires_grad_(True)
        lab = lab_host.to(dev, non_blocking=True)
        loss = adist.reg_loss_sharded(z, lab, dims, gamma, delta, algo=args.algo)
        loss.backward()
        grad_host.copy_(z.grad, non_blocking=True)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, do_flush):
        """K steps bracketed by barrier + synchronize on both sides.  Every step is timed on the device with
        its own CUDA event pair; the L2 flush (a 256 MiB memset, ~0.07 ms) runs BETWEEN the steps, outside the
        per-step intervals.  Returns (sum of the K step times in ms, ..., bracket time incl. flushes, median step
        time), each the max over ranks."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        br0 = torch.cuda.Event(enable_timing=True)
        br1 = torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        br0.record()
        out = None
        for e0, e1 in evs:
            if do_flush:
                flush.zero_()
            e0.record()
            out = fn()
            e1.record()
        br1.record()
        barrier()
        t1 = time.time()
        per_step = [e0.elapsed_time(e1) for e0, e1 in evs]
        ms = sum(per_step)
        ms_med = statistics.median(per_step)
        ms_bracket = br0.elapsed_time(br1)
        if world > 1:
            t = torch.tensor([ms, ms_bracket, ms_med], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_bracket, ms_med = float(t[0].item()), float(t[1].item()), float(t[2].item())
        return ms, out, t0, t1, ms_bracket, ms_med

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    for _ in range(3):
        step_e2e()
    barrier()

    # ---- device-resident timing (value), with the pair kernel timed alone through the library hooks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lib.arvae_launch_count(1)
    ms_total, out, t0, t1, ms_bracket, ms_median = timed(step_device, args.steps, True)
    launches = int(lib.arvae_launch_count(1))
    # the pair kernel alone, through the library's event hooks, in a few extra steps of its own: the hooks put event
    # records between the launches of a step, which would keep the timed steps above from overlapping their launches
    lib.arvae_profile_enable(1)
    timed(step_device, min(args.steps, 8), True)
    ksum, kn = ctypes.c_float(), ctypes.c_int()
    lib.arvae_profile_pair_kernel_ms(ctypes.byref(ksum), ctypes.byref(kn))
    lib.arvae_profile_enable(0)
    loss_val = float(out[0].item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_step_mean = ms_total / args.steps
    ms_step = ms_median
    value = pairs / (ms_step * 1e-3) / 1e9
    grad_dev = out[1]

    # ---- end-to-end timing (host buffers in, host results out) -----------------------------------
    ms_e2e_total, loss_e2e, _, _, _, ms_e2e = timed(step_e2e, args.steps, True)
    e2e_value = pairs / (ms_e2e * 1e-3) / 1e9
    h2d = z_host.numel() * 4 + lab_host.numel() * 4
    d2h = grad_host.numel() * 4 + 4

    # ---- parity of THIS run's results: sampled rows of every rank against a float64 evaluation of the reference
    # formula (utils/trainer.py:390-401 and its autograd backward) done with torch on the device
    parity = parity_block(case, grad_dev, r0, n_local, loss_val, dev, world, dist if world > 1 else None)

    if comm is not None:
        comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (pair) kernel -----------------------------------------------------
    peaks, peaks_src = load_peaks()
    sm_count = lib.arvae_device_sm_count()
    f_ghz = float(peaks.get("sm_max_mhz", 1965.0)) / 1e3
    per_dim = arvae_b200.mufu_per_pair(case["z"].to(dev), case["labels"].to(dev), dims, gamma, delta, algo=args.algo)
    MUFU_PER_PAIR = sum(per_dim) / len(per_dim)
    mufu_peak = MUFU_LANES_PER_CLK_SM * sm_count * f_ghz / MUFU_PER_PAIR  # Gpairs/s per GPU, one MUFU per inlier pair
    # the kernel's ACTUAL mix: which build of the pair kernel ran (reg_internal.cuh: kSharedMaxAbsU) and its loop's SASS;
    # an inlier pair costs the loop's MUFU share, a pair with an outlier (2 - MUFU_PER_PAIR of them per pair) keeps both MUFU
    zreg = case["z"].to(dev)[:, list(dims)]
    u_max = float((2.0 * abs(delta) * 1.4426950408889634 * zreg.abs()).max().item()) if zreg.numel() else 0.0
    loop = SHARED_LOOP if (MUFU_PER_PAIR == 1.0 and u_max <= SHARED_MAX_ABS_U) else PLAIN_LOOP
    mufu_mix = MUFU_PER_PAIR - (1.0 - loop["mufu_per_inlier_pair"]) * (2.0 - MUFU_PER_PAIR)
    mufu_peak_mix = MUFU_LANES_PER_CLK_SM * sm_count * f_ghz / mufu_mix
    issue_peak_mix = ISSUE_LANES_PER_CLK_SM * sm_count * f_ghz / loop["instr_per_pair"]
    fma_peak_mix = FP32_LANES_PER_CLK_SM * sm_count * f_ghz / loop["fp32_ops_per_pair"]
    peak_mix = min(mufu_peak_mix, issue_peak_mix, fma_peak_mix)
    mufu_peak_2 = MUFU_LANES_PER_CLK_SM * sm_count * f_ghz / 2.0
    k_ms = (ksum.value / kn.value) if kn.value else ms_step
    pairs_per_launch = pairs / world  # this rank's rows x all columns x R
    achieved = pairs_per_launch / (k_ms * 1e-3) / 1e9
    algo_bytes = (2 * B * R + n_local * R) * ALGO_BYTES_PER_ROWCOL  # columns in (u, a) + gradient columns out
    # DRAM traffic of the pair kernel per launch, from the ncu --set full capture of this same command committed
    # under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum); only known for the configuration profiled
    traffic = NCU_DRAM_BYTES_PER_LAUNCH.get((B, R, world, int(args.algo)))
    roofline = {
        "bound": "mufu", "achieved": achieved, "peak": peak_mix, "unit": "Gpairs/s", "frac": achieved / peak_mix,
        "traffic": traffic,
        "peak_of_actual_mix": {"mufu_per_pair": mufu_mix, "mufu_roof": mufu_peak_mix, "instr_per_pair": loop["instr_per_pair"],
                               "issue_roof": issue_peak_mix, "fp32_ops_per_pair": loop["fp32_ops_per_pair"],
                               "fma_pipe_roof": fma_peak_mix, "max_abs_u": u_max,
                               "note": f"{loop['what']} (reg_sorted.cu: loop_const); SASS of the constant-sign loop: {loop['sass']}; "
                                       "`peak` = the lowest of the three roofs of that mix"},
        "frac_of_1mufu_roof": achieved / mufu_peak, "peak_1mufu": mufu_peak,
        "traffic_source": "profiles/r2c_pair_ncu_full_summary.csv (ncu --set full, reg_tiles_kernel<1, 0, 1>)" if traffic else None,
        "peak_source": f"{MUFU_LANES_PER_CLK_SM:.0f} MUFU lanes/clk/SM x {sm_count} SMs x {f_ghz:.3f} GHz (sm_max_mhz, "
                       f"MEASURED_PEAKS.json {peaks_src}) / {mufu_mix:.4f} MUFU per evaluated pair (this run's "
                       "algorithm and instruction mix; every one of the B^2 R ordered pairs is evaluated); "
                       "lane rate confirmed by bench_tools/pipe_rates.cu (profiles/)",
        "frac_of_2mufu_dense_roof": achieved / mufu_peak_2, "peak_2mufu_dense": mufu_peak_2,
        "mufu_per_pair_by_dim": list(per_dim),
        "kernel": "reg pair kernel", "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step_mean,
        "evaluated_pairs_per_launch": pairs_per_launch, "mufu_per_pair": MUFU_PER_PAIR,
        "algorithmic_bytes_per_launch": algo_bytes,
        "hbm": {"achieved_gbs": algo_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
                "note": "not the bound: ~5e3 pairs per algorithmic byte"},
    }

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_sample(case, 12.0, steps=1)
            cpu_baseline = {"value": v, "unit": "Gpairs/s", "cores": info["cores"], "kind": info["kind"],
                            "sample": info["sample"]}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "Gpairs/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"failed: {e}"}

    extra = {}
    if world == 1 and not args.no_small_batch:
        try:
            extra["small_batch"] = small_batch_extra(dev)
            extra["small_batch"]["clocks"] = clocks
        except Exception as e:  # pragma: no cover
            extra["small_batch"] = {"error": str(e)}

    line = {
        "metric": "reg_loss_fwd_bwd_throughput", "value": value, "unit": "Gpairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_step_mean": ms_step_mean,
        "fixed_ms": ms_step - k_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B": B, "Z": Z, "R": R, "gamma": gamma, "delta": delta,
                   "pairs_per_step": pairs, "parallelism": f"row-block x{world}" if world > 1 else "single GPU",
                   "transport": (args.transport if world > 1 else None),
                   "algo": args.algo, "cuda_graph": graphed is not None,
                   "l2_flush": "256 MiB memset between the timed steps (each step has its own CUDA event pair; the "
                               "flush is outside the per-step intervals; inputs are ~6 MB, far below L2)",
                   "ms_per_step_incl_flush": ms_bracket / args.steps},
        "clocks": clocks, "e2e": {"value": e2e_value, "unit": "Gpairs/s", "ms_per_step": ms_e2e,
                                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity": parity, "extra": extra, "loss": loss_val, "loss_e2e": loss_e2e,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
