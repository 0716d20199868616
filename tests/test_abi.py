"""CPU checks of the C-ABI boundary: the library builds/loads, exports every symbol the header
declares, and validates arguments before touching the GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from dpdist_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpdist_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert "dpd_fv_forward" in declared and "dpd_head_forward" in declared
    for name in declared:
        assert hasattr(lib, name), "libdpdist_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES and include/dpdist_b200.h disagree"
    assert lib.dpd_version() == _lib.ABI_VERSION == 4     # DPD_ABI_VERSION


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for anchor in ("utils/dpdist_util.py:22-141", "utils/dpdist_util.py:459-492", "utils/dpdist_util.py:412-544",
                   "utils/tf_util.py:161-228"):
        assert anchor in src


def test_invalid_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    c = np.zeros(8, np.float32)
    rc = lib.dpd_fv_forward(None, 1, 64, 8, _lib.fptr(c), 0.125, 1, 0, None, None)
    assert rc == -1 and b"null" in lib.dpd_last_error()
    rc = lib.dpd_fv_forward(ctypes.c_void_p(256), 1, 64, 99, _lib.fptr(c), 0.125, 1, 0, ctypes.c_void_p(256), None)
    assert rc == -2 and b"G=99" in lib.dpd_last_error()
    rc = lib.dpd_fv_forward(ctypes.c_void_p(256), 1, 64, 8, _lib.fptr(c), -1.0, 1, 0, ctypes.c_void_p(256), None)
    assert rc == -1
    rc = lib.dpd_voxel_assign(None, 1, 1, 8, _lib.fptr(c), _lib.fptr(c), _lib.fptr(c), None, None, None, None)
    assert rc == -1
    cfg = _lib.HeadConfig(2, 64, 8, 20, 5, 1000, 0)      # H not a multiple of 16
    assert lib.dpd_head_packed_bytes(ctypes.byref(cfg)) == 0
    assert b"H=1000" in lib.dpd_last_error()


def test_backward_and_data_entry_points_validate_flags_and_pointers():
    lib = _lib.load()
    p = ctypes.c_void_p(256)
    c = np.zeros(8, np.float32)
    # input gradients need DPD_HEAD_TRAIN | DPD_HEAD_INPUT_GRAD; the latter alone is rejected by every head entry point
    cfg = _lib.HeadConfig(2, 64, 8, 20, 5, 1024, _lib.HEAD_TRAIN)
    assert lib.dpd_head_backward_inputs(ctypes.byref(cfg), p, p, p, p, 1 << 20, None) == -1
    assert b"DPD_HEAD_INPUT_GRAD" in lib.dpd_last_error()
    bad = _lib.HeadConfig(2, 64, 8, 20, 5, 1024, _lib.HEAD_INPUT_GRAD)
    assert lib.dpd_head_workspace_bytes(ctypes.byref(bad)) == 0 and b"DPD_HEAD_TRAIN" in lib.dpd_last_error()
    # backward without the training flag, bad stage
    inf = _lib.HeadConfig(2, 64, 8, 20, 5, 1024, 0)
    assert lib.dpd_head_backward(ctypes.byref(inf), p, p, p, 0, *([p] * 8), p, 1 << 20, None) == -1
    assert lib.dpd_head_backward(ctypes.byref(cfg), p, p, p, 9, *([p] * 8), p, 1 << 20, None) == -1
    # training keeps every activation: more rows than one chunk is refused, not silently chunked
    big = _lib.HeadConfig(8192, 64, 8, 20, 5, 1024, _lib.HEAD_TRAIN)
    assert lib.dpd_head_backward(ctypes.byref(big), p, p, p, 0, *([p] * 8), p, 1 << 20, None) == -3
    # training workspace: the flag costs memory, the input-gradient flag a little more
    w0 = lib.dpd_head_workspace_bytes(ctypes.byref(inf))
    w1 = lib.dpd_head_workspace_bytes(ctypes.byref(cfg))
    w2 = lib.dpd_head_workspace_bytes(ctypes.byref(_lib.HeadConfig(2, 64, 8, 20, 5, 1024, _lib.HEAD_TRAIN | _lib.HEAD_INPUT_GRAD)))
    assert 0 < w0 < w1 < w2
    assert lib.dpd_fv_backward(p, 1, 4, 8, _lib.fptr(c), 0.0, 1, 0, p, p, None) == -1          # sigma
    assert lib.dpd_adam_step_dev(p, p, p, p, 16, None, 0.9, 0.999, 1e-8, None) == -1
    want = 1e-4 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    assert abs(lib.dpd_adam_lr_t(1e-4, 0.9, 0.999, 1) - want) < 1e-4 * want          # float32 arguments


def test_head_sizes_are_consistent():
    lib = _lib.load()
    for flags in (_lib.HEAD_AUTO, _lib.HEAD_SIMT):
        cfg = _lib.HeadConfig(2048, 64, 8, 20, 5, 1024, flags)
        pb = lib.dpd_head_packed_bytes(ctypes.byref(cfg))
        wb = lib.dpd_head_workspace_bytes(ctypes.byref(cfg))
        assert pb >= 4666371 * 4
        rows = 2048 * 64
        assert wb >= 2 * rows * 1024 * 4
        assert wb < 64 * rows * 1024 * 4


def test_no_cpu_fallback():
    from dpdist_b200 import dpdist_util
    with pytest.raises(_lib.DPDistNativeError):
        dpdist_util.get_3dmfv_tf(torch.zeros(1, 4, 3), n_gaussians=27)
    with pytest.raises(_lib.DPDistNativeError):
        dpdist_util.local_z(torch.zeros(1, 27, 20), None, NUM_DIMS=3, k=3)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dpdist_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_layer_and_batch_norm_entry_points_validate_their_arguments():
    """dpd_layer_* / dpd_bn_* / dpd_gather_rows (ABI 3: --BN 1 and conv_version 3) reject bad calls before any launch."""
    lib = _lib.load()
    p = ctypes.c_void_p(256)
    need = lib.dpd_layer_workspace_bytes(1024, 1024, 1024)
    assert need > 8 * 1024 * 1024 * 4                                     # split-K partials of a 1024 x 1024 weight gradient
    assert lib.dpd_layer_workspace_bytes(0, 16, 16) == 0
    assert lib.dpd_layer_forward(p, 128, 64, None, p, 64, 0, p, None, None, None, 0, 0, 0, 0, None) == -1        # null weights
    assert lib.dpd_layer_forward(p, 128, 64, p, p, 64, 7, p, None, None, None, 0, 0, 0, 0, None) == -1           # bad act
    assert lib.dpd_layer_forward(p, 128, 64, p, p, 3, 1, p, None, None, None, 0, 0, 0, 0, None) == -1            # narrow + relu
    # gathered layer: packed weights must cover k^3*C + 3 rows; a bad grid is refused
    assert lib.dpd_layer_forward(None, 128, 64, p, p, 64, 0, p, p, p, p, 64, 8, 20, 5, None) == -1
    assert b"k^3*C + 3" in lib.dpd_last_error()
    assert lib.dpd_layer_forward(None, 128, 2528, p, p, 64, 0, p, p, p, p, 64, 99, 20, 5, None) == -1
    # backward: workspace size is checked; wide layers need N % 128 == 0
    assert lib.dpd_layer_backward(p, 1024, 1024, p, 1024, p, p, p, None, None, None, None, 0, 0, 0, 0, p, 1024, None) == -3
    assert lib.dpd_layer_backward(p, 1024, 1024, p, 64, p, p, p, None, None, None, None, 0, 0, 0, 0, p, need, None) == -2
    assert b"N % 128" in lib.dpd_last_error()
    assert lib.dpd_bn_forward(p, 128, 64, p, p, 1e-3, 5, p, p, p, p, 1 << 20, None) == -1                         # bad act
    assert lib.dpd_bn_forward(p, 128, 64, p, p, 1e-3, 1, p, p, p, p, 16, None) == -3                              # workspace
    assert lib.dpd_bn_backward(p, None, 128, 64, p, p, p, p, 1e-3, 1, p, p, p, p, 1 << 20, None) == -1
    assert lib.dpd_gather_rows(p, p, p, 0, 64, 8, 20, 5, p, None) == -1
    assert lib.dpd_relu_backward(None, p, 16, None) == -1
    assert lib.dpd_add_inplace(p, p, 6, None) == -1                                                               # n % 4


def test_tensor_core_operand_order_is_a_permutation_of_the_reference_row():
    """dpd_debug_tc_operand_order (host only): the packed layer-1 row of the fp16 tensor-core path -- channel-split with
    permuted 64-element blocks -- must hold every element of the reference row [patch | offsets] exactly once, padding
    everywhere else, the offsets and the padding tail in the LAST block (the kernel may skip that block's unused K-steps),
    and 16-byte units of the 16-channel part must not straddle a tap (they are copied with one 16-byte cp.async)."""
    import ctypes
    lib = _lib.load()
    for k, C, Kp in ((5, 20, 2560), (3, 20, 576), (5, 16, 2048), (3, 8, 256), (5, 4, 512), (2, 20, 192), (5, 12, 1536), (7, 20, 6912)):
        taps, E = k ** 3, k ** 3 * C
        out = (ctypes.c_int * Kp)()
        assert lib.dpd_debug_tc_operand_order(taps, C, Kp, out) == 0
        o = np.array(out[:])
        data = o[o < E + 3]
        assert sorted(data.tolist()) == list(range(E + 3)), (k, C, Kp)            # every element once
        assert (o >= 0).all() and len(set(o.tolist())) == Kp                       # padding positions are distinct too
        last = o[Kp - 64:]
        assert set(range(E, E + 3)) <= set(last.tolist()), "offsets live in the last block"
        cx = C & ~7
        if cx:
            x_pos = np.nonzero((o < E) & ((o % C) < cx))[0]
            for p0 in x_pos[::8][:200]:                                             # units of 8 start at multiples of 8
                unit = o[p0 - p0 % 8: p0 - p0 % 8 + 8]
                if ((unit < E) & ((unit % C) < cx)).all():
                    assert len(set((unit // C).tolist())) == 1 and (np.diff(unit) == 1).all()
    assert lib.dpd_debug_tc_operand_order(125, 20, 2500, (ctypes.c_int * 2500)()) != 0      # Kp % 64 != 0
    assert lib.dpd_debug_tc_operand_order(125, 20, 2560, None) != 0


def test_tensor_core_operand_order_random_shapes():
    """The same permutation property over random (k, C, padding) combinations, including rows whose padding spans several
    64-element blocks and channel counts without a 16-channel part (C = 4) or without a remainder (C = 8, 16, 24)."""
    import ctypes
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for _ in range(60):
        k = int(rng.integers(1, 8)); C = 4 * int(rng.integers(1, 9))
        E = k ** 3 * C
        Kp = ((E + 3 + 63) // 64 + int(rng.integers(0, 3))) * 64
        out = (ctypes.c_int * Kp)()
        assert lib.dpd_debug_tc_operand_order(k ** 3, C, Kp, out) == 0
        o = np.array(out[:])
        assert sorted(o[o < E + 3].tolist()) == list(range(E + 3)), (k, C, Kp)
        assert len(set(o.tolist())) == Kp and int(o.max()) < Kp
