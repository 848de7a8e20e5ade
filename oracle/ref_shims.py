"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Shims that let the UNMODIFIED reference (`/root/reference`, read-only, absent on the GPU
box) be imported in this container so its outputs can be recorded as golden fixtures
(`oracle/make_golden.py` -> `tests/golden/*.npz`) and used to pin `oracle/monoflex_oracle.py`.

What is shimmed and why:
  * `_ext`            reference native op module (src/vision.cpp:4-9). Its THC/TH host code cannot be
                      compiled against torch 2.11, but `src/cpu/dcn_v2_im2col_cpu.cpp` compiles as-is
                      (oracle/Makefile -> oracle/_ref/libdcn_im2col_ref.so). The shim follows the host
                      sequence of `src/cpu/dcn_v2_cpu.cpp:17-107` (fwd) / `:109-233` (bwd) with torch.mm
                      in place of THFloatBlas_gemm, calling the reference's own C loops through ctypes.
  * `yacs.config`     not installed: attribute-dict CfgNode (merge_from_file / merge_from_list / freeze).
  * `inplace_abn`     third-party, unpinned (requirements.txt:14): BN(|gamma|+eps) + leaky_relu(0.01)
                      -- "parity unpinned" at this op, see DESIGN.md.
  * shapely / matplotlib / skimage / pycocotools / iopath / fvcore : MagicMock (visualiser-only imports).
  * `model.layers.utils.select_topk`: torch-1.4 integer `/` == floor division, CUDA-type asserts dropped
                      (SURVEY H5 / Appendix B.1).
"""
import ast
import ctypes
import os
import sys
import types
from unittest import mock

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libdcn_im2col_ref.so")


# ----------------------------------------------------------------------------- yacs
class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        if init:
            for k, v in init.items():
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        if self.__dict__.get("_frozen"):
            raise AttributeError("frozen cfg")
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    @staticmethod
    def _coerce(v):
        if isinstance(v, str):
            try:
                return ast.literal_eval(v)
            except Exception:
                return v
        return v

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self:
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = self._coerce(v)

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._coerce(v)


# ----------------------------------------------------------------------------- inplace_abn
class InPlaceABN(nn.Module):
    """mapillary/inplace_abn semantics as recalled (unpinned): y = leaky_relu(bn(x; |w|+eps, b), slope)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu",
                 activation_param=0.01):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.activation, self.activation_param = activation, activation_param
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        y = F.batch_norm(x, self.running_mean, self.running_var, self.weight.abs() + self.eps, self.bias,
                         self.training, self.momentum, self.eps)
        return F.leaky_relu(y, self.activation_param)


# ----------------------------------------------------------------------------- _ext
def _load_ref_so():
    lib = ctypes.CDLL(_REF_SO)
    fp, i = ctypes.c_void_p, ctypes.c_int
    lib.modulated_deformable_im2col_cpu.argtypes = [fp, fp, fp] + [i] * 15 + [fp]
    lib.modulated_deformable_col2im_cpu.argtypes = [fp, fp, fp] + [i] * 15 + [fp]
    lib.modulated_deformable_col2im_coord_cpu.argtypes = [fp, fp, fp, fp] + [i] * 15 + [fp, fp]
    return lib


class _RefExt:
    """`_ext` stand-in: reference C loops (oracle/_ref) + torch.mm, sequence of src/cpu/dcn_v2_cpu.cpp."""

    def __init__(self):
        self.lib = _load_ref_so()

    def dcn_v2_forward(self, input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        input, offset, mask = input.contiguous().float(), offset.contiguous().float(), mask.contiguous().float()
        B, C, H, W = input.shape
        Co = weight.shape[0]
        Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
        Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
        out = torch.empty(B, Co, Ho, Wo)
        cols = torch.empty(C * kh * kw, Ho * Wo)
        w2 = weight.reshape(Co, -1)
        for b in range(B):
            self.lib.modulated_deformable_im2col_cpu(
                input[b].data_ptr(), offset[b].data_ptr(), mask[b].data_ptr(), 1, C, H, W, Ho, Wo, kh, kw,
                ph, pw, sh, sw, dh, dw, dg, cols.data_ptr())
            out[b] = (bias[:, None] + w2 @ cols).view(Co, Ho, Wo)
        return out

    def dcn_v2_backward(self, input, weight, bias, offset, mask, grad_output, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        input, offset, mask = input.contiguous().float(), offset.contiguous().float(), mask.contiguous().float()
        grad_output = grad_output.contiguous().float()
        B, C, H, W = input.shape
        Co = weight.shape[0]
        Ho, Wo = grad_output.shape[2:]
        gi, gw, gb = torch.zeros_like(input), torch.zeros_like(weight), torch.zeros_like(bias)
        go, gm = torch.zeros_like(offset), torch.zeros_like(mask)
        w2 = weight.reshape(Co, -1)
        for b in range(B):
            gout = grad_output[b].reshape(Co, -1)
            cols = (w2.t() @ gout).contiguous()
            self.lib.modulated_deformable_col2im_coord_cpu(
                cols.data_ptr(), input[b].data_ptr(), offset[b].data_ptr(), mask[b].data_ptr(), 1, C, H, W, Ho, Wo,
                kh, kw, ph, pw, sh, sw, dh, dw, dg, go[b].data_ptr(), gm[b].data_ptr())
            self.lib.modulated_deformable_col2im_cpu(
                cols.data_ptr(), offset[b].data_ptr(), mask[b].data_ptr(), 1, C, H, W, Ho, Wo,
                kh, kw, ph, pw, sh, sw, dh, dw, dg, gi[b].data_ptr())
            cols2 = torch.empty(C * kh * kw, Ho * Wo)
            self.lib.modulated_deformable_im2col_cpu(
                input[b].data_ptr(), offset[b].data_ptr(), mask[b].data_ptr(), 1, C, H, W, Ho, Wo, kh, kw,
                ph, pw, sh, sw, dh, dw, dg, cols2.data_ptr())
            gw += (gout @ cols2.t()).view_as(gw)
            gb += gout.sum(1)
        return gi, go, gm, gw, gb


def _patched_select_topk(heat_map, K=100):
    # model/layers/utils.py:61-100 with torch-1.4 integer-division semantics restored.
    batch, cls, height, width = heat_map.size()
    heat_map = heat_map.view(batch, cls, -1)
    topk_scores_all, topk_inds_all = torch.topk(heat_map, K)
    topk_ys = (topk_inds_all // width).float()
    topk_xs = (topk_inds_all % width).float()
    topk_scores_all = topk_scores_all.view(batch, -1)
    topk_scores, topk_inds = torch.topk(topk_scores_all, K)
    topk_clses = (topk_inds // K).float()
    g = lambda f: f.view(batch, -1, 1).gather(1, topk_inds.unsqueeze(-1)).view(batch, K)
    return topk_scores, g(topk_inds_all), topk_clses, g(topk_ys), g(topk_xs)


_installed = False


def install():
    """Make `import model...` resolve to the reference with every shim in place."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    np.int = int  # numpy 2 removed it (data/datasets/kitti.py:434)
    yacs = types.ModuleType("yacs")
    yacs_cfg = types.ModuleType("yacs.config")
    yacs_cfg.CfgNode = CfgNode
    yacs.config = yacs_cfg
    sys.modules["yacs"], sys.modules["yacs.config"] = yacs, yacs_cfg
    abn = types.ModuleType("inplace_abn")
    abn.InPlaceABN = InPlaceABN
    sys.modules["inplace_abn"] = abn
    for name in ["shapely", "shapely.geometry", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.figure", "matplotlib.backends", "matplotlib.backends.backend_agg", "matplotlib.patches",
                 "matplotlib.gridspec", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "pycocotools", "pycocotools.mask", "iopath", "iopath.common", "iopath.common.file_io",
                 "skimage", "skimage.transform", "fvcore", "fvcore.common", "fvcore.common.file_io", "fire",
                 "tensorboardX", "torch.utils.tensorboard"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()
    sys.modules["_ext"] = _RefExt()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import model.layers.utils as lu
    lu.select_topk = _patched_select_topk
    _installed = True


def reference_cfg(width=1280, height=384, device="cpu"):
    """cfg = config/defaults.py merged with runs/monoflex.yaml, PRETRAIN off, CPU."""
    install()
    from config import cfg as _cfg
    cfg = _cfg.clone()
    cfg.defrost()
    cfg.merge_from_file(os.path.join(REF_ROOT, "runs/monoflex.yaml"))
    cfg.MODEL.PRETRAIN = False
    cfg.MODEL.DEVICE = device
    cfg.INPUT.WIDTH_TRAIN, cfg.INPUT.HEIGHT_TRAIN = width, height
    cfg.INPUT.WIDTH_TEST, cfg.INPUT.HEIGHT_TEST = width, height
    # same mutations tools/plain_train_net.py:86-108 applies by default
    cfg.DATASETS.TEST_SPLIT = "test"  # eval forward without label fields (detector_infer.py:58)
    return cfg


def build_reference_model(cfg):
    install()
    from model.detector import KeypointDetector
    import model.head.detector_infer as di
    di.select_topk = _patched_select_topk
    return KeypointDetector(cfg)
