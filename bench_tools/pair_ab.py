#!/usr/bin/env python
"""A/B timing of the pair kernel across builds of the library: one subprocess per .so (ARVAE_LIB_PATH), C4 workload,
the library's own CUDA events around the pair-kernel launch.   python bench_tools/pair_ab.py lib1.so lib2.so ..."""
import ctypes, json, os, statistics, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, REPO)
    import torch
    import arvae_b200
    from arvae_b200 import _lib, synth
    lib = _lib.load()
    B = int(os.environ.get("AB_BATCH", "65536"))
    c = synth.make_case(os.environ.get("AB_WORKLOAD", "c4_mnist_b65536"), B)
    delta = float(os.environ.get("AB_DELTA", c["delta"]))
    z, lab = c["z"].cuda(), c["labels"].cuda()
    ms = []
    for it in range(13):
        lib.arvae_profile_enable(1)
        loss, gc, _ = arvae_b200.reg_loss_rows(z, lab, c["reg_dims"], c["gamma"], delta, 0, c["B"], algo=2)
        ks, kn = ctypes.c_float(), ctypes.c_int()
        lib.arvae_profile_pair_kernel_ms(ctypes.byref(ks), ctypes.byref(kn))
        if it >= 3:
            ms.append(ks.value / max(kn.value, 1))
    print(json.dumps({"lib": os.environ.get("ARVAE_LIB_PATH", "default"), "pair_kernel_ms_median": statistics.median(ms),
                      "min": min(ms), "max": max(ms), "loss": loss.item(), "B": c["B"], "delta": delta}))
    sys.exit(0)
for libpath in sys.argv[1:] or ["default"]:
    env = dict(os.environ)
    if libpath != "default":
        env["ARVAE_LIB_PATH"] = os.path.abspath(libpath)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env)
