"""CPU: the C-ABI library loads without a GPU driver and exports every symbol include/monoflex_b200.h declares; the
ctypes table in monoflex_b200/_lib.py covers the same set; product modules keep the reference's state_dict keys."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "monoflex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(ROOT, "monoflex_b200", "libmonoflex_b200.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(path)


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_ctypes_table_matches_header():
    from monoflex_b200 import _lib
    assert sorted(list(_lib.SIGNATURES) + ["mf_last_error"]) == header_symbols()


def test_version_and_error_string(lib):
    lib.mf_last_error.restype = ctypes.c_char_p
    assert lib.mf_version() == 100
    assert lib.mf_set_conv_impl(7) != 0
    assert b"impl must be 0 or 1" in lib.mf_last_error()
    assert lib.mf_dcn_v2_psroi_pooling_forward() == -2
    assert b"not built" in lib.mf_last_error()


def test_conv_block_n(lib):
    assert [lib.mf_conv_block_n(c) for c in (3, 16, 27, 32, 64, 128, 256, 2304)] == [16, 16, 32, 32, 64, 128, 128, 128]


def test_state_dict_keys_match_reference_layout():
    from monoflex_b200 import synthetic as syn
    from monoflex_b200.config import default_cfg
    from monoflex_b200.model.detector import KeypointDetector
    m = KeypointDetector(default_cfg())
    sd = syn.make_state_dict(0)
    assert set(m.state_dict()) == set(sd)
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert sum(p.numel() for p in m.parameters()) == 20952538      # SURVEY §8 [probe]
    m.load_state_dict(sd)
    for key in ("backbone.base.level3.tree1.tree2.conv1.weight", "backbone.dla_up.ida_1.proj_2.conv.conv_offset_mask.bias",
                "backbone.ida_up.up_2.weight", "heads.predictor.reg_features.5.0.weight",
                "heads.predictor.reg_heads.5.1.bias", "heads.predictor.trunc_offset_conv.0.weight"):
        assert key in sd


def test_no_cpu_fallback():
    """The product path refuses CPU tensors / training mode instead of silently falling back."""
    from monoflex_b200.config import default_cfg
    from monoflex_b200.model.detector import KeypointDetector
    m = KeypointDetector(default_cfg(width=64, height=64)).eval()
    with pytest.raises(RuntimeError):
        m.backbone(torch.zeros(1, 3, 64, 64))
    m.train()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 64, 64))
