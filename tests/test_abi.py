"""The C-ABI shared library: it loads without a GPU or a CUDA driver, exports every symbol the header declares,
validates its arguments, and fails LOUDLY (never falls back) when no sm_100 device is present."""
import ctypes as C
import os
import re

import pytest
import torch

from diffsim_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "diffsim_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = N.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in the header but not exported"
        assert s in N.PROTOTYPES, f"{s} has no ctypes prototype"
    assert sorted(N.PROTOTYPES) == syms


def test_abi_version_and_error_string():
    lib = N.load()
    assert lib.ds_abi_version() == 1
    assert isinstance(N.last_error(), str)


def test_struct_layout_matches_header():
    # ds_tensor4: ptr, int64[4], int64[4], int32 (+pad) ; ds_tensor5: ptr, int64[5], int64[5], int32 (+pad)
    assert C.sizeof(N.Tensor4) == 8 + 32 + 32 + 8
    assert C.sizeof(N.Tensor5) == 8 + 40 + 40 + 8


def test_workspace_queries_need_no_device():
    lib = N.load()
    assert lib.ds_pair_reduce_workspace_bytes(16, 655360) >= 16 * 64 * 12 * 4
    assert lib.ds_simmat_workspace_bytes(2032, 2032, 655360) >= 2032 * 2032 * 4
    t = N.Tensor5()
    t.size[:] = [6, 2, 8, 256, 160]
    t.stride[:] = [655360, 327680, 160, 1280, 1]
    assert lib.ds_aas_pairs_workspace_bytes(t, 3) > 6 * 32 * 16
    assert lib.ds_aas_triplets_workspace_bytes(t, 2) > 8 * 32 * 16


def test_argument_validation_comes_first():
    lib = N.load()
    out = (C.c_float * 4)()
    # null pointers, bad modes and bad dtypes are rejected before anything touches the device
    assert lib.ds_pair_reduce(None, None, 1, 16, 16, 16, N.DS_F16, N.DS_SIM_COSINE, out, None, 0, None) == N.DS_ERR_INVALID
    assert "null" in N.last_error()
    buf = (C.c_uint16 * 64)()
    assert lib.ds_pair_reduce(buf, buf, 1, 16, 16, 16, N.DS_F16, 7, out, None, 0, None) == N.DS_ERR_INVALID
    assert lib.ds_pair_reduce(buf, buf, 1, 16, 16, 16, 9, N.DS_SIM_COSINE, out, None, 0, None) == N.DS_ERR_INVALID
    assert lib.ds_pair_reduce(buf, buf, 2, 16, 8, 16, N.DS_F16, N.DS_SIM_COSINE, out, None, 0, None) == N.DS_ERR_INVALID
    assert lib.ds_simmat(buf, 4, 12, buf, 4, 16, 16, N.DS_F16, N.DS_SIM_COSINE, out, 4, None, 0, None) == N.DS_ERR_INVALID
    assert lib.ds_twoafc(None, None, 4, N.DS_SIM_COSINE, None, None, None) == N.DS_ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    lib = N.load()
    assert lib.ds_device_ok() == N.DS_ERR_CUDA
    x = torch.randn(2, 64).half()
    out = torch.zeros(2)
    ws = torch.zeros(1 << 16, dtype=torch.uint8)
    rc = lib.ds_pair_reduce(x.data_ptr(), x.data_ptr(), 2, 64, 64, 64, N.DS_F16, N.DS_SIM_COSINE, out.data_ptr(),
                            ws.data_ptr(), ws.numel(), None)
    assert rc == N.DS_ERR_CUDA
    assert out.abs().sum().item() == 0.0  # nothing was computed on the host
    with pytest.raises(N.DiffSimError):
        N.check(rc)
    # the torch-facing layer refuses CPU tensors outright
    from diffsim_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.pair_reduce(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.attn_fwd(x.view(1, 1, 2, 64), x.view(1, 1, 2, 64), x.view(1, 1, 2, 64))


def test_product_path_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under diffsim_b200/ may import it."""
    pkg = os.path.join(ROOT, "diffsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
