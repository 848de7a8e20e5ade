"""Backward of the backbone (DLA-34 base + DLAUp/IDAUp with DCNv2, model/backbone/dla_dcn.py) from the tape the train-mode
forward plan recorded (engine.Plan.tape): one record per emitted layer, replayed in reverse emission order, each record's
backward composed from the operators of the C ABI (csrc/mf_wgrad.cu, mf_bn_train.cu, mf_dcn_bwd.cu, mf_bwd_misc.cu and
the forward implicit-GEMM kernel for the data gradients).

Gradient bookkeeping: every activation buffer of the plan gets a zero-initialised fp16 gradient buffer of the same shape
(channel slices of a Root concatenation share their parent's, exactly like the forward shares storage) and EVERY backward
operator accumulates into it - stride-1 data gradients through the forward kernel's fused residual input (in place), the
others through a temporary + `mf_add_rows_f16`. Accumulating everywhere makes the result independent of which consumer of
a multiply-used activation runs first. Weight gradients are fp32. Gradients travel in fp16, so the seed must carry a loss
scale (see head_backward.py).

The 7x7 stem convolution reads the 8-channel packed image; with `stem_wgrad=True` its weight gradient runs on a 16-channel
zero-padded copy of the rows (the narrowest box the tensor-core wgrad takes) - written but NOT yet run on hardware (the
round's GPU budget ended), so it is opt-in and the default reports that one gradient as missing (None).
"""
import torch

from . import engine
from ._lib import call, load
from .backward import column_sum, conv2d_dgrad_stride2


def _st():
    return torch.cuda.current_stream().cuda_stream


class _Grads(object):
    """fp16 gradient buffers mirroring the plan's activation buffers."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def of(self, act):
        key = act.buf.data_ptr()
        g = self.bufs.get(key)
        if g is None:
            g = torch.zeros_like(act.buf)
            self.bufs[key] = g
        return g

    def ptr(self, act):
        return self.of(act).data_ptr() + 2 * act.ch_off

    def ld(self, act):
        return self.of(act).shape[1]

    def add(self, act, src, src_ld=None):
        """grad(act) += src ([M, >= C] fp16 rows)"""
        call("mf_add_rows_f16", self.ptr(act), self.ld(act), src.data_ptr(), src.shape[1] if src_ld is None else src_ld, act.M, act.C,
             _st())

    def bound(self, plan, act):
        """an engine.Act of `plan` whose storage is grad(act) (for kernels driven through engine.Plan)"""
        a = plan.act(act.B, act.H, act.W, act.C)
        a.buf, a.ch_off = self.of(act), act.ch_off
        return a


def _bn_backward(rec_bn, raw, y, dy_ptr, dy_ld, d_raw, d_res, act, abs_weight, put):
    M, dev = raw.M, d_raw.device
    for (mod, _raw, _y, stats, c0, cc) in rec_bn:
        dg = torch.empty(2, cc, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, load().mf_bn_train_workspace(M, cc) // 4), dtype=torch.float32, device=dev)
        engine.bn_backward_launch(mod, raw.ptr() + 2 * c0, raw.ld, dy_ptr + 2 * c0, dy_ld, y.ptr() + 2 * c0, y.ld, M, cc, stats, act,
                                  d_raw.data_ptr() + 2 * c0, d_raw.shape[1],
                                  (d_res.data_ptr() + 2 * c0) if d_res is not None else None,
                                  d_res.shape[1] if d_res is not None else 0, dg, ws, _st())
        off = c0 - rec_bn[0][4]                                   # slices of one module are consecutive 256-channel chunks
        put(mod.weight, dg[0] * torch.sign(mod.weight.detach()[off:off + cc]) if abs_weight else dg[0], off, cc)
        put(mod.bias, dg[1], off, cc)


def _conv_backward(G, x, weight, bias, stride, pad, d_raw, B, Ho, Wo, put, need_dx=True, stem_wgrad=False):
    """gradients of y = conv(x, weight) + bias given d_raw = dL/dy rows [B*Ho*Wo, Cout(_padded)] fp16."""
    dev = d_raw.device
    cout, cin, kh, kw = weight.shape
    cout_p = d_raw.shape[1]
    if x.C != 8 and (cin in (16, 32) or cin % 64 == 0):
        dw = torch.empty(cout_p, cin, kh, kw, dtype=torch.float32, device=dev)
        call("mf_conv2d_wgrad_nhwc_f16", x.ptr(), x.ld, x.B, x.H, x.W, cin, d_raw.data_ptr(), cout_p, cout_p, kh, stride, pad,
             dw.data_ptr(), _st())
        put(weight, dw[:cout])
    elif x.C == 8 and stem_wgrad:
        # 7x7 stem on the packed image (8 channels = 3 real + 5 zero): widen the rows to the narrowest supported box (16)
        cin_w_real = weight.shape[1]
        x16 = torch.zeros(x.M, 16, dtype=torch.half, device=dev)
        x16[:, :8] = x.buf.view(x.M, -1)[:, x.ch_off:x.ch_off + 8]
        dw = torch.empty(cout_p, 16, kh, kw, dtype=torch.float32, device=dev)
        call("mf_conv2d_wgrad_nhwc_f16", x16.data_ptr(), 16, x.B, x.H, x.W, 16, d_raw.data_ptr(), cout_p, cout_p, kh, stride, pad,
             dw.data_ptr(), _st())
        put(weight, dw[:cout, :cin_w_real])
    else:
        put(weight, None)
    if bias is not None:
        put(bias, column_sum(d_raw)[:cout])
    if not need_dx:
        return
    w = weight.detach().float()
    if stride == 1 and w.is_contiguous() and kh == kw:
        P = engine.Plan(str(dev))
        dya = P.act(B, Ho, Wo, cout_p)
        dya.buf = d_raw
        gx = G.bound(P, x)
        P.conv_dgrad(dya, w, pad, gx)                                # grad(x) += dgrad, in place (rotated taps packed in one kernel)
        P.finalize()
        P.run()
        return
    if cout_p != cout:
        w = torch.cat([w, torch.zeros(cout_p - cout, cin, kh, kw, device=dev)], 0)
    if stride == 1:
        w_rot = w.flip(2, 3).transpose(0, 1).contiguous()           # [Cin, Cout_p, k, k]
        P = engine.Plan(str(dev))
        dya = P.act(B, Ho, Wo, cout_p)
        dya.buf = d_raw
        gx = G.bound(P, x)
        P.conv(dya, w_rot, 1, kh - 1 - pad, None, act=engine.ACT_NONE, residual=gx, out=gx)      # grad(x) += dgrad, in place
        P.finalize()
        P.run()
    else:
        if (stride, kh, pad) != (2, 3, 1):
            raise NotImplementedError("tape: data gradient of a %dx%d stride-%d pad-%d convolution" % (kh, kw, stride, pad))
        G.add(x, conv2d_dgrad_stride2(d_raw, w, B, Ho, Wo))


@torch.no_grad()
def backbone_backward(backbone, plan, d_feat_rows, stem_wgrad=False):
    """backbone: DLASeg in train mode, plan: its last train-mode plan (backbone.last_plan), d_feat_rows [B*H/4*W/4, 64] fp16:
    (scaled) gradient of the loss w.r.t. the backbone's output feature map.
    Returns {parameter name: fp32 gradient or None (not built)} for every parameter the forward used."""
    dev = d_feat_rows.device
    G = _Grads(dev)
    names = {id(p): n for n, p in backbone.named_parameters()}
    grads = {}

    def put(param, value, off=0, cc=None):
        n = names[id(param)]
        if value is None:
            grads[n] = None
            return
        if cc is None or cc == param.numel():
            grads[n] = value.reshape(param.shape).float()
        else:                                                       # one 256-channel chunk of a wider normalisation layer
            if n not in grads or grads[n] is None:
                grads[n] = torch.zeros(param.shape, dtype=torch.float32, device=dev)
            grads[n].view(-1)[off:off + cc] = value

    G.add(plan.output, d_feat_rows.contiguous())
    for rec in reversed(plan.tape):
        kind = rec["kind"]
        if kind == "conv_bn":
            x, raw, y, residual = rec["x"], rec["raw"], rec["y"], rec["residual"]
            d_raw = torch.empty(raw.M, raw.C, dtype=torch.half, device=dev)
            d_res = torch.empty(raw.M, raw.C, dtype=torch.half, device=dev) if residual is not None else None
            _bn_backward(rec["bn"], raw, y, G.ptr(y), G.ld(y), d_raw, d_res, rec["act"], rec["abs_weight"], put)
            if residual is not None:
                G.add(residual, d_res)
            _conv_backward(G, x, rec["weight"], rec["bias"], rec["stride"], rec["pad"], d_raw, raw.B, raw.H, raw.W, put,
                           need_dx=x is not plan.input, stem_wgrad=stem_wgrad)
        elif kind == "dcn":
            x, raw, y, mod, om = rec["x"], rec["raw"], rec["y"], rec["mod"], rec["om"]
            d_raw = torch.empty(raw.M, raw.C, dtype=torch.half, device=dev)
            _bn_backward(rec["bn"], raw, y, G.ptr(y), G.ld(y), d_raw, None, engine.ACT_RELU, False, put)
            C, cout = x.C, raw.C
            # grad columns (1x1 forward conv of dY with W^T), sampled columns, col2im + coord, wgrad - backward.dcn_backward
            # inlined so that x may be a channel slice (ld != C)
            wt = mod.weight.detach().float().permute(2, 3, 1, 0).reshape(9 * C, cout, 1, 1).contiguous()
            P = engine.Plan(str(dev))
            dya = P.act(x.B, x.H, x.W, cout)
            dya.buf = d_raw
            gcol = P.conv(dya, wt, 1, 0, None, act=engine.ACT_NONE)
            P.finalize()
            P.run()
            cols = torch.empty(x.M, 9 * C, dtype=torch.half, device=dev)
            call("mf_dcn_sample_cols_nhwc_f16", x.ptr(), x.ld, om.data_ptr(), 32, cols.data_ptr(), x.B, x.H, x.W, C, _st())
            dw9 = torch.empty(cout, 9 * C, 1, 1, dtype=torch.float32, device=dev)
            call("mf_conv2d_wgrad_nhwc_f16", cols.data_ptr(), 9 * C, x.B, x.H, x.W, 9 * C, d_raw.data_ptr(), cout, cout, 1, 1, 0,
                 dw9.data_ptr(), _st())
            put(mod.weight, dw9.view(cout, 3, 3, C).permute(0, 3, 1, 2).contiguous())
            put(mod.bias, column_sum(d_raw))
            dx = torch.empty(x.M, C, dtype=torch.half, device=dev)
            dom = torch.empty(x.M, 32, dtype=torch.float32, device=dev)
            call("mf_dcn_col2im_nhwc_f16", x.ptr(), x.ld, om.data_ptr(), 32, gcol.ptr(), dx.data_ptr(), C, dom.data_ptr(), x.B, x.H,
                 x.W, C, _st())
            G.add(x, dx)
            # offset / mask convolution (3x3, bias, 27 -> padded 32 output channels): its pre-activation gradient is `dom`
            _conv_backward(G, x, mod.conv_offset_mask.weight, mod.conv_offset_mask.bias, 1, 1, dom.half(), x.B, x.H, x.W, put)
        elif kind == "maxpool2":
            x, y = rec["x"], rec["y"]
            dx = torch.empty(x.M, x.C, dtype=torch.half, device=dev)
            call("mf_maxpool2_bwd_nhwc_f16", x.ptr(), G.ptr(y), dx.data_ptr(), x.B, x.H, x.W, x.C, x.ld, G.ld(y), x.C, _st())
            G.add(x, dx)
        elif kind == "upsample_add":
            x, y, skip, f, wt = rec["x"], rec["y"], rec["skip"], rec["f"], rec["wt"]
            C = x.C
            if skip is not None:
                call("mf_add_rows_f16", G.ptr(skip), G.ld(skip), G.ptr(y), G.ld(y), y.M, C, _st())
            dx = torch.empty(x.M, C, dtype=torch.half, device=dev)
            dw = torch.empty(4 * f * f, C, dtype=torch.float32, device=dev)
            ws = torch.empty(max(1, load().mf_upsample_bwd_workspace(x.B, x.H, x.W, C, f) // 4), dtype=torch.float32, device=dev)
            call("mf_upsample_bwd_nhwc_f16", x.ptr(), wt.data_ptr(), G.ptr(y), dx.data_ptr(), dw.data_ptr(), x.B, x.H, x.W, C, f, x.ld,
                 G.ld(y), C, ws.data_ptr(), _st())
            G.add(x, dx)
            put(rec["weight"], dw.t().contiguous())
        else:
            raise NotImplementedError("tape record %r" % kind)
    return grads
