"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, rejects bad arguments without touching a GPU, and the Python layer refuses CPU tensors
(there is no fallback path)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from arvae_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "arvae_b200.h")).read()
    return sorted(set(re.findall(r"ARVAE_API[^;(]*?\b(arvae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    from arvae_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/arvae_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string(lib):
    assert lib.arvae_version() == 200
    assert isinstance(lib.arvae_last_error(), bytes)


def test_argument_errors_need_no_gpu(lib):
    dims = (ctypes.c_int32 * 2)(0, 1)
    rc = lib.arvae_reg_loss_fwdbwd_f32(None, 1, 1, None, 1, 1, dims, dims, 99, 0, 4, 4, 1.0, 1.0, 0,
                                       None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"out of range" in lib.arvae_last_error()
    rc = lib.arvae_reg_loss_fwdbwd_f32(None, 1, 1, None, 1, 1, dims, dims, 2, 3, 2, 4, 1.0, 1.0, 0,
                                       None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"row range" in lib.arvae_last_error()
    neg = (ctypes.c_int32 * 1)(-1)
    rc = lib.arvae_reg_loss_scatter_bwd_f32(None, None, neg, 1, 0, 4, None, 4, None)
    assert rc == -1 and b"negative" in lib.arvae_last_error()
    assert lib.arvae_reg_loss_workspace_bytes(-1, 0, 1) == 0
    # sharded step: argument checks come before any CUDA call
    ctx = ctypes.c_void_p()
    assert lib.arvae_shard_create(3, 2, 128, 4, ctypes.byref(ctx)) == -1   # rank >= world
    assert lib.arvae_shard_create(0, 17, 128, 4, ctypes.byref(ctx)) == -1  # more ranks than one NVSwitch box
    assert lib.arvae_shard_comm_bytes(0, 4, 2) == 0
    assert lib.arvae_shard_comm_bytes(8192, 6, 8) > 8 * 6 * 8192 * 12


def test_python_layer_refuses_cpu_tensors():
    import arvae_b200
    z = torch.zeros(8, 4)
    a = torch.zeros(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        arvae_b200.compute_reg_loss(z, a, 0, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        arvae_b200.reg_loss_sign(z[:, 0], a)
    with pytest.raises(RuntimeError, match="CUDA"):
        arvae_b200.reparam_kld_reg(z, z, z, torch.zeros(8, 4), (0,), 1.0, 0.0, 1.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "arvae_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "arvae_oracle" not in text, f


def test_ctypes_table_has_the_arity_of_every_prototype(lib):
    """Guards against ABI drift: the number of parameters in each header prototype equals the number of
    ctypes argtypes bound for it."""
    from arvae_b200 import _lib
    text = open(os.path.join(REPO, "include", "arvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"ARVAE_API[^;(]*?\b(arvae_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(_lib.SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
