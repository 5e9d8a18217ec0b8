"""The C-ABI shared library: builds with nvcc for sm_100a, loads without a GPU, and exports
every symbol include/b200rl.h declares (no compute calls here)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "b200rl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported():
    from b200rl import build, native
    names = _declared()
    assert len(names) >= 27
    assert set(names) == set(native.SIGNATURES), "ctypes table and include/b200rl.h disagree"
    lib = native.Lib(build.build())
    assert lib.missing == []
    assert lib.b200rl_version() >= 100
    assert lib.last_error() == ""


def test_bad_arguments_return_an_error_code_not_an_abort():
    from b200rl import build, native
    lib = native.Lib(build.build())
    rc = lib.b200rl_lstm_fwd(*([0] * 12), 0, 0, 4, 0, 0)          # B = 0 is rejected before any launch
    assert rc != 0 and "lstm_fwd" in lib.last_error()
    rc = lib.b200rl_ppo_head(*([0] * 13), 4, 1, 2, 0, *([0.0] * 6), 0)   # T = 1
    assert rc != 0 and "ppo_head" in lib.last_error()


def test_missing_library_fails_loudly(tmp_path):
    from b200rl import native
    with pytest.raises(native.NativeError, match="no CPU fallback"):
        native.Lib(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_refused_by_the_product_ops():
    import torch
    from b200rl import native, ops
    with pytest.raises(native.NativeError, match="CUDA device"):
        ops.Ops().compute_gae(torch.zeros(2, 3, 1), 0.99, 0.95)
