"""TEST INFRASTRUCTURE ONLY (oracle): CPU emulation of the fp16 rounding points of the CUDA "fast" forward on top of the
oracle's fp32 graph - operands / stored activations of selected layer groups are rounded to fp16 exactly where the kernels round
them, everything else stays fp32 (= what the hi/lo pair "strict" kernels compute up to 2^-22). The rounding is a cast pair
(`.half().float()`), i.e. differentiable with derivative 1: torch autograd through an emulated forward is the EXACT fp32 derivative
of the fp16-rounded function. Used by tools/precision_emulation.py (which layers need pair operands for the 1e-3 forward
contract) and tools/grad_emulation_cpu.py (how far the gradient of this chaotic synthetic network moves when only the forward is
rounded - the yardstick for the backward tape's element-wise checks). Never imported by the product."""
import torch
import torch.nn.functional as F

from oracle import monoflex_oracle as mo


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
