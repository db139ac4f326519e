"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, the
product arm refuses to run without a CUDA device (no CPU fallback), and the instruction mixes bench.py derives its
roofs from are the ones in the SASS of the built library."""
import collections
import json
import os
import re
import shutil
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--batch", "2048"], capture_output=True, text=True, timeout=280)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gpairs/s" and d["higher_is_better"] is True
    assert d["metric"] == "reg_loss_fwd_bwd_throughput" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "c4_mnist_b65536" and d["vs_baseline"] is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_fails_loudly_without_cuda():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=120)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def _hot_loops(lib, function):
    """Instruction histograms of the loops (backward branches) of one kernel that contain MUFU.RCP, in address order."""
    out = subprocess.run(["cuobjdump", "-sass", "-fun", function, lib], capture_output=True, text=True, timeout=120).stdout
    ins = []
    for line in out.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    loops = []
    for addr, text in ins:
        m = re.search(r"0x([0-9a-f]+)", text) if "BRA" in text else None
        if m and int(m.group(1), 16) < addr:
            body = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
            hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
            if hist.get("MUFU.RCP", 0) >= 8 and len(body) < 1200:
                loops.append((len(body), hist))
    return loops


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")
def test_roofline_instruction_mixes_are_the_ones_in_the_built_sass():
    """roofline.peak is the lowest roof of the constant-sign loop's ACTUAL mix: the MUFU / FP32 / issue counts per pair
    that bench.py uses must be what cuobjdump shows for the loop of each build of the pair kernel."""
    sys.path.insert(0, REPO)
    import bench
    from arvae_b200 import build as b
    lib = b.build()
    for only1, loop in ((1, bench.SHARED_LOOP), (0, bench.PLAIN_LOOP)):
        fn = f"_ZN5arvae16reg_tiles_kernelILb1ELb0ELb{only1}EEEvNS_9TilesArgsE"  # <GRAD, no sign sums, ONLY1>
        packed = [(n, h) for n, h in _hot_loops(lib, fn) if h.get("FFMA2", 0) > 0]
        assert packed, fn
        n, h = packed[0]
        pairs = 32 * h["LDS.128"] // 2   # one LDS.128 = four columns x four rows per lane; 32 pairs per two of them
        assert h["MUFU.RCP"] / pairs == pytest.approx(loop["mufu_per_inlier_pair"]), (fn, h)
        fp32 = 2 * (h.get("FFMA2", 0) + h.get("FADD2", 0) + h.get("FMUL2", 0))
        assert fp32 / pairs == pytest.approx(loop["fp32_ops_per_pair"]), (fn, h)
        assert n / pairs == pytest.approx(loop["instr_per_pair"], rel=0.02), (fn, n, pairs)
