"""CPU emulation of the rounding points of the CUDA forward (DESIGN.md §4 "precision modes"): runs the oracle with the
operands / stored activations of selected layer groups rounded to fp16 exactly where the kernels round them, everything else
fp32 (= what the hi/lo-split "strict" kernels compute up to 2^-22). Used to pick which layers need split operands for the
1e-3 end-to-end contract. TEST/DESIGN TOOLING: imports the oracle, never imported by the product.

    python tools/precision_emulation.py [H W]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import monoflex_oracle as mo          # noqa: E402
from monoflex_b200 import synthetic as syn        # noqa: E402


def r16(t):
    return t.half().float()


GROUPS = ["stem", "level2", "level3", "level4", "level5", "dla_up", "ida_up", "head"]


def group_of(name):
    if ".base.base_layer" in name or ".base.level0" in name or ".base.level1" in name:
        return "stem"
    for lv in ("level2", "level3", "level4", "level5"):
        if ".base." + lv in name:
            return lv
    if ".dla_up." in name:
        return "dla_up"
    if ".ida_up." in name:
        return "ida_up"
    if name.startswith("heads."):
        return "head"
    raise KeyError(name)


class Emu(object):
    """policy: {group: 'f16' | 'strict'}"""

    def __init__(self, policy):
        self.policy = policy

    def mode(self, name):
        return self.policy[group_of(name)]

    def qi(self, x, name):
        return r16(x) if self.mode(name) == "f16" else x

    qw = qi

    def qo(self, y, name):
        return r16(y) if self.mode(name) == "f16" else y

    def install(self):
        emu = self
        self.saved = {k: getattr(mo, k) for k in ("conv_bn", "dcn", "deform_conv", "ida_up", "iabn", "F")}

        def conv_bn(sd, conv, bn, x, stride=1, pad=1, relu=True, residual=None):
            y = F.conv2d(emu.qi(x, conv), emu.qw(sd[conv + '.weight'], conv), None, stride, pad)
            y = mo.bn_eval(sd, bn, y)
            if residual is not None:
                y = y + residual
            return emu.qo(F.relu(y) if relu else y, conv)

        def dcn(sd, p, x):
            xq = emu.qi(x, p)
            om = F.conv2d(xq, emu.qw(sd[p + '.conv_offset_mask.weight'], p), sd[p + '.conv_offset_mask.bias'], 1, 1)
            o1, o2, m = torch.chunk(om, 3, dim=1)
            B, C, H, W = x.shape
            cols = mo.dcn_columns(xq, torch.cat((o1, o2), 1), torch.sigmoid(m)).reshape(B, C * 9, H * W)
            cols = emu.qi(cols, p)                                       # blended samples are stored as the fp16 MMA operand
            w = emu.qw(sd[p + '.weight'], p)
            out = torch.matmul(w.reshape(w.shape[0], -1), cols) + sd[p + '.bias'].view(1, -1, 1)
            return out.view(B, -1, H, W)

        def deform_conv(sd, p, x):
            return emu.qo(F.relu(mo.bn_eval(sd, p + '.actf.0', dcn(sd, p + '.conv', x))), p)

        def ida_up(sd, p, layers, startp, endp, up_f):
            for i in range(startp + 1, endp):
                j = i - startp
                u = mo.up(sd, '%s.up_%d' % (p, j), deform_conv(sd, '%s.proj_%d' % (p, j), layers[i]), up_f[j])
                s = emu.qo(u + layers[i - 1], p + '.x')                  # fused upsample_add stores once
                layers[i] = deform_conv(sd, '%s.node_%d' % (p, j), s)

        def iabn(sd, p, x):
            return emu.qo(emu.saved["iabn"](sd, p, x), p)

        class FProxy(object):
            def __getattr__(self, k):
                return getattr(F, k)

            def conv2d(self, x, w, b=None, stride=1, pad=0):
                name = emu.names.get(id(w))
                if name is not None and name.startswith("heads."):
                    return F.conv2d(emu.qi(x, name), emu.qw(w, name), b, stride, pad)
                return F.conv2d(x, w, b, stride, pad)

        mo.conv_bn, mo.dcn, mo.deform_conv, mo.ida_up, mo.iabn, mo.F = conv_bn, dcn, deform_conv, ida_up, iabn, FProxy()

    def uninstall(self):
        for k, v in self.saved.items():
            setattr(mo, k, v)

    def run(self, sd, x, tg):
        self.names = {id(v): k for k, v in sd.items()}
        self.install()
        try:
            with torch.no_grad():
                xin = r16(x) if self.policy["stem"] == "f16" else x
                feats = mo.backbone(sd, xin)
                taps = {}
                pred = mo.predictor(sd, feats, tg["edge_indices"], tg["edge_len"], taps=taps)
        finally:
            self.uninstall()
        return feats, taps["cls_logits"], pred["cls"], pred["reg"]


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
    B = 1
    sd = syn.make_state_dict(0)
    x = syn.make_images(B, H, W)
    tg = syn.make_targets(B, W // 4, H // 4)
    ref = Emu({g: "strict" for g in GROUPS}).run(sd, x, tg)
    cases = {"all f16": {g: "f16" for g in GROUPS}}
    for g in GROUPS:
        cases["only %s f16" % g] = {k: ("f16" if k == g else "strict") for k in GROUPS}
    cases["head f16, backbone strict"] = {k: ("f16" if k == "head" else "strict") for k in GROUPS}
    cases["head+ida_up f16"] = {k: ("f16" if k in ("head", "ida_up") else "strict") for k in GROUPS}
    cases["head+level5 f16"] = {k: ("f16" if k in ("head", "level5") else "strict") for k in GROUPS}
    cases["strict: stem,level2,dla_up,ida_up"] = {k: ("strict" if k in ("stem", "level2", "dla_up", "ida_up") else "f16") for k in GROUPS}
    print("%-40s %10s %10s %10s %10s" % ("policy", "features", "logits", "cls", "reg"))
    for name, pol in cases.items():
        out = Emu(pol).run(sd, x, tg)
        print("%-40s %10.2e %10.2e %10.2e %10.2e" % ((name,) + tuple(rel(a, b) for a, b in zip(out, ref))))


if __name__ == "__main__":
    main()
