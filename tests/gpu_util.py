"""Helpers for the -m gpu parity tests: build tiny engine plans around single ops and compare with torch-CPU fp32 math on
the same fp16-representable operands (so the only differences are accumulation order and the fp16 output rounding)."""
import numpy as np
import torch
import torch.nn.functional as F

from monoflex_b200 import engine
from monoflex_b200._lib import load


class FakeBN(object):
    def __init__(self, c, gen, abs_w=False):
        self.weight = torch.from_numpy(gen.uniform(0.5, 1.5, c).astype(np.float32)).cuda()
        self.bias = torch.from_numpy((gen.standard_normal(c) * 0.1).astype(np.float32)).cuda()
        self.running_mean = torch.from_numpy((gen.standard_normal(c) * 0.1).astype(np.float32)).cuda()
        self.running_var = torch.from_numpy(gen.uniform(0.5, 1.5, c).astype(np.float32)).cuda()
        self.eps = 1e-5

    def cpu_apply(self, y):
        return F.batch_norm(y, self.running_mean.cpu(), self.running_var.cpu(), self.weight.cpu(), self.bias.cpu(), False,
                            0.0, self.eps)


def h16(t):
    """round to fp16-representable fp32"""
    return t.half().float()


def to_rows(x_nchw, c_pad=None):
    """[B,C,H,W] fp32 -> [B*H*W, C_pad] fp16 rows (zero padded channels)"""
    B, C, H, W = x_nchw.shape
    c_pad = C if c_pad is None else c_pad
    rows = torch.zeros(B * H * W, c_pad, dtype=torch.half)
    rows[:, :C] = x_nchw.permute(0, 2, 3, 1).reshape(-1, C).half()
    return rows.cuda()


def to_rows_split(x_nchw):
    """[B,C,H,W] fp32 -> [B*H*W, 2C] fp16 pair rows [hi | lo] (strict-precision activation layout)"""
    B, C, H, W = x_nchw.shape
    r = x_nchw.permute(0, 2, 3, 1).reshape(-1, C).float()
    hi = r.half()
    lo = (r - hi.float()).half()
    return torch.cat([hi, lo], 1).cuda()


def from_rows(act):
    """engine.Act -> [B,C,H,W] fp32 CPU"""
    return act.nchw_view().float().cpu().contiguous()


def set_impl(impl):
    import os
    import pytest
    forced = os.environ.get("MF_CONV_IMPL")
    if forced is not None and int(forced) != impl:
        if impl == 0:
            impl = int(forced)        # "default implementation" requests follow the forced diagnostic mode
        else:
            pass
    assert load().mf_set_conv_impl(impl) == 0


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
