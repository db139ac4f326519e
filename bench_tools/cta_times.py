#!/usr/bin/env python
"""Per-CTA start/end times of the pair kernel (ARVAE_DEBUG_TIMES=1): how balanced is the persistent grid?"""
import ctypes, os, sys
os.environ["ARVAE_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arvae_b200 import _lib, ops, synth
lib = _lib.load()
lib.arvae_debug_times_offset.restype = ctypes.c_int64
lib.arvae_debug_times_offset.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]
ALGO = int(os.environ.get("ALGO", "2"))
c = synth.make_case("c4_mnist_b65536")
z, lab = c["z"].cuda(), c["labels"].cuda()
dims = c["reg_dims"]; R = len(dims); B = c["B"]
packed = ops.pack_columns(z, lab, dims, dims)
for shards in (1,):
    n = B // shards
    g = ctypes.c_int32()
    off = lib.arvae_debug_times_offset(B, n, R, ALGO, ctypes.byref(g))
    lib.arvae_reg_loss_workspace_bytes_algo.restype = ctypes.c_size_t
    ws_bytes = int(lib.arvae_reg_loss_workspace_bytes_algo(B, n, R, ALGO))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    loss = torch.empty((), dtype=torch.float64, device="cuda")
    gc = torch.empty((n, R), dtype=torch.float32, device="cuda")
    for it in range(3):
        rc = lib.arvae_reg_loss_fwdbwd_f32(packed.data_ptr(), 2 * R, 1, packed.data_ptr() + 4 * R, 2 * R, 1,
                                           _lib.i32_array(range(R)), _lib.i32_array(range(R)), R, 0, n, B,
                                           c["gamma"], c["delta"], int(os.environ.get("ALGO", "2")), loss.data_ptr(), None, gc.data_ptr(), None,
                                           ws.data_ptr(), ws_bytes, None)
        assert rc == 0
    torch.cuda.synchronize()
    t = ws[off:off + 16 * g.value].cpu().numpy().view(np.uint64).reshape(-1, 2).astype(np.int64)
    t = t[t[:, 1] > 0]
    t0 = t[:, 0].min()
    st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    print(f"shards {shards}: {len(t)} CTAs; start spread {st.max():.1f} us; end min/p10/median/p90/max = "
          f"{en.min():.1f} / {np.percentile(en,10):.1f} / {np.median(en):.1f} / {np.percentile(en,90):.1f} / {en.max():.1f} us; "
          f"mean busy {np.mean(en-st):.1f} us")
    busy = en - st
    order = np.argsort(en)
    print("busy us by CTA index (every 16th):", np.round(busy[::16], 0).tolist())
