"""CPU: the C-ABI library builds, loads, and exports exactly what include/zsv_b200.h declares; host-only
entry points (shape / size queries, argument validation) behave.  No compute call is made without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from zeroshotvideoclassification_b200 import build, _lib
    build.build()
    return _lib.load()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "zsv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zsv_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from zeroshotvideoclassification_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in zsv_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared           # the ctypes table covers the header, nothing more
    assert lib.zsv_abi_version() == _lib.ABI_VERSION


def test_shape_queries(lib):
    from zeroshotvideoclassification_b200._lib import ConvDesc
    out = (C.c_int32 * 3)()
    # layer2.0.conv1.0.0 of R(2+1)D-18 at bs=22
    d = ConvDesc(22, 16, 56, 56, 64, 230, 1, 3, 3, 1, 2, 2, 0, 1, 1, 0)
    assert lib.zsv_conv3d_out_shape(C.byref(d), out) == 0
    assert list(out) == [16, 28, 28]
    assert lib.zsv_cpad(230) == 232 and lib.zsv_cpad(45) == 48 and lib.zsv_cpad(64) == 64
    assert lib.zsv_conv3d_packed_weight_bytes(C.byref(d), 0) == 9 * 230 * 64 * 2
    assert lib.zsv_conv3d_packed_weight_bytes(C.byref(d), 1) == 9 * 64 * 232 * 2
    assert lib.zsv_conv3d_stat_rows(C.byref(d)) > 0
    assert lib.zsv_conv3d_wgrad_workspace(C.byref(d)) > 0
    # stem through the W-folded layout
    s = ConvDesc(22, 16, 112, 112, 3, 45, 1, 7, 7, 1, 2, 2, 0, 3, 3, 1)
    assert lib.zsv_conv3d_out_shape(C.byref(s), out) == 0 and list(out) == [16, 56, 56]
    assert lib.zsv_conv3d_packed_weight_bytes(C.byref(s), 0) == 7 * 45 * 64 * 2
    assert lib.zsv_conv3d_packed_weight_bytes(C.byref(s), 1) == 0


def test_errors_are_reported_not_swallowed(lib):
    from zeroshotvideoclassification_b200._lib import ConvDesc
    out = (C.c_int32 * 3)()
    bad = ConvDesc(1, 4, 8, 8, 16, 16, 3, 3, 3, 3, 1, 1, 1, 1, 1, 0)     # stride 3 is outside the contract
    rc = lib.zsv_conv3d_out_shape(C.byref(bad), out)
    assert rc != 0 and b"stride" in lib.zsv_last_error()
    assert lib.zsv_nearest_class(None, None, 1, 1, 1, 1, None, None, None) != 0
    assert lib.zsv_bn_apply(None, None, None, None, None, None, None, None, 1, 8, 0, None) != 0


def test_no_cpu_fallback():
    """Product entry points refuse non-CUDA tensors instead of silently computing on the host."""
    import torch
    from zeroshotvideoclassification_b200 import ops, video_models as vm
    with pytest.raises(RuntimeError):
        ops.nearest_class(torch.zeros(2, 300), torch.zeros(3, 300))
    with pytest.raises(RuntimeError):
        ops.repack_input(torch.zeros(1, 3, 2, 4, 4))
    model = vm.get_network(vm.default_opt())
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 1, 3, 8, 32, 32))
    # optimizer, graph runner and the accuracy helpers added later follow the same rule
    from zeroshotvideoclassification_b200.optim import FusedAdam
    from zeroshotvideoclassification_b200.graph import GraphedStep
    from zeroshotvideoclassification_b200 import accuracy
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            GraphedStep(lambda x: x, (torch.zeros(2),))
        with pytest.raises((RuntimeError, AssertionError)):
            accuracy.class_overlap_mask(torch.zeros(3, 300), torch.zeros(2, 300), 0.05)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may import it."""
    pkg = os.path.join(ROOT, "zeroshotvideoclassification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
