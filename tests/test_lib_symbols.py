"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads without a
GPU, and exports every symbol include/mq_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from mannequin2_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    from mannequin2_b200 import build
    build.build()
    return _lib.lib()


def test_header_and_binding_agree(lib):
    header = open(os.path.join(ROOT, "include", "mq_b200.h")).read()
    declared = set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", header))
    declared -= {"mq_net", "mq_head"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_flag_constants_match_header():
    """The MQ_TC_* flags the host layer passes are the header's (a drifted bit would silently select another operand format)."""
    header = open(os.path.join(ROOT, "include", "mq_b200.h")).read()
    flags = {k: int(v, 0) for k, v in re.findall(r"#define\s+(MQ_TC_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)", header)}
    assert flags["MQ_TC_FORCE"] == _lib.TC_FORCE and flags["MQ_TC_OFF"] == _lib.TC_OFF
    assert flags["MQ_TC_FP16"] == _lib.TC_FP16 and flags["MQ_TC_KEEP_X"] == _lib.TC_KEEP_X
    assert _lib.TC_MODE_BITS == flags["MQ_TC_PLANES_MASK"] | flags["MQ_TC_FP16"]
    assert _lib.TC_MODES["fp16x2"] == 2 | flags["MQ_TC_FP16"] and _lib.TC_MODES["exact"] == 3
    bits = [flags[k] for k in ("MQ_TC_FORCE", "MQ_TC_OFF", "MQ_TC_NOPIPE", "MQ_TC_FP16", "MQ_TC_KEEP_X")]
    assert len(set(bits)) == len(bits) and all(b & flags["MQ_TC_PLANES_MASK"] == 0 and b & (b - 1) == 0 for b in bits)


def test_layered_mlp_workspace_holds_the_input_planes(lib):
    """mq_mlp_ws_bytes includes the region MQ_TC_KEEP_X keeps the operand planes of the input batch in
    (3 planes x batch padded to 128 x (784 + 1 ones column) padded to 128 16-bit words)."""
    net = _lib.make_net(784, [128, 128, 10], [2, 2, 0])
    for batch in (128, 8200, 65536):
        rows = (batch + 127) // 128 * 128
        assert lib.mq_mlp_ws_bytes(ctypes.byref(net), batch) >= 3 * rows * 896 * 2


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.MqNet) == 4 * 5 + 4 * 8 * 6
    assert ctypes.sizeof(_lib.MqHead) == 16 + 4 * 8 + 8 + 4 + 4 + 8
    net = _lib.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    assert net.n_params == 5126 and net.logstd_off == 5123
    assert list(net.w_off)[:3] == [0, 768, 4928] and list(net.b_off)[:3] == [704, 4864, 5120]
    net = _lib.make_net(784, [128, 128, 10], [2, 2, 0])
    assert net.n_params == 118282


def test_sizes_and_error_codes_without_gpu(lib):
    assert lib.mq_version() >= 100
    net = _lib.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    assert lib.mq_mlp_acts_floats(ctypes.byref(net), 10) == 10 * (64 + 64 + 3)
    assert lib.mq_scan_ws_bytes(1 << 20) > 0 and lib.mq_norm_ws_bytes(2048, 1) > 0
    # packed records and operand images of the tensor-core gradient kernel: sizes are host arithmetic
    assert lib.mq_record_width(ctypes.byref(net)) == 16                      # 11 obs + 3 act + adv + logp_old = 64 bytes
    sizes = [lib.mq_tc_image_bytes(ctypes.byref(net), pl) for pl in (1, 2, 3, 0)]
    assert 0 < sizes[0] < sizes[1] < sizes[2] == sizes[3] and all(b % 16 == 0 for b in sizes)
    assert lib.mq_tc_image_bytes(ctypes.byref(_lib.make_net(784, [128, 10], [2, 0])), 0) == 0     # does not fit the kernel
    # argument validation happens before any CUDA call
    assert lib.mq_adam_step(None, None, None, None, None, 4, 1, 0, 0.1, 0.1, 0.1, 1e-8, 0, None) == _lib.MQ_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.MQ_ERR_INVALID, lib)
    assert b"null pointer" in lib.mq_last_error()
    assert lib.mq_discounted(None, 5, 0.5, None, None, 0, None) == _lib.MQ_ERR_SHAPE
    with pytest.raises(AssertionError):
        _lib.check(_lib.MQ_ERR_SHAPE, lib)


def test_product_has_no_cpu_path():
    """Without a CUDA device every numeric entry point raises instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import mannequin2_b200 as mq
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mq.Adam([1.0, 2.0], horizon=10, lr=0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mq.discounted(np.ones(4), horizon=3)
    # nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "mannequin2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("no oracle", ""), f
