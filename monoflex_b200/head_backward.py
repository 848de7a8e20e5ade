"""Backward of the predictor (detector_predictor.py:121-165) composed from the operators of backward.py / the C ABI - the first
complete slice of the training tape (DESIGN.md "Training tape"): loss gradients -> sigmoid_hm -> edge-fusion tail -> 1x1 heads
-> InPlaceABN -> the nine 3x3 convs -> gradient of the backbone's feature map, plus the gradient of every predictor parameter.

Works on the activations the train-mode forward plan kept (`plan.ctx`, model/head/detector_predictor.py). Gradients flow in
fp16 NHWC rows (weight gradients fp32), so the caller multiplies the loss gradients by a loss scale (>= 128) and divides
the results by it, like any fp16 training step."""
import torch

from . import engine
from ._lib import call, load
from .backward import column_sum, conv2d_dgrad


def _st():
    return torch.cuda.current_stream().cuda_stream


def _bn_backward(entry, dy_ptr, dy_ld, dx_ptr, dx_ld, act, M, dev, mod=None):
    """mf_bn_train_backward on one <= 256-channel slice kept by Plan.bn_train -> (dgamma, dbeta). `mod`: the normalisation module
    (a converted torch.nn.SyncBatchNorm exchanges its sums across ranks, engine.bn_backward_launch)."""
    raw, y, stats, c0, cc = entry
    dg = torch.empty(2, cc, dtype=torch.float32, device=dev)
    ws = torch.empty(max(1, load().mf_bn_train_workspace(M, cc) // 4), dtype=torch.float32, device=dev)
    engine.bn_backward_launch(mod, raw.ptr() + 2 * c0, raw.ld, dy_ptr, dy_ld, y.ptr() + 2 * c0, y.ld, M, cc, stats, act, dx_ptr, dx_ld,
                              None, 0, dg, ws, _st())
    return dg[0], dg[1]


@torch.no_grad()
def predictor_backward(pred, plan, grad_cls, grad_reg):
    """pred: the `_predictor` module in train mode, plan: its last train-mode plan (pred.last_plan), grad_cls [B,3,H,W] /
    grad_reg [B,50,H,W] fp32: (scaled) loss gradients w.r.t. the predictor outputs.
    Returns (grads: {parameter name: fp32 gradient}, d_features: [B*H*W, 64] fp16 rows)."""
    ctx = plan.ctx
    x, hid = ctx["x"], ctx["hid"]
    B, H, W, hc = x.B, x.H, x.W, pred.head_conv
    M, dev = B * H * W, grad_reg.device
    HW = H * W
    grads = {}
    names = {id(p): n for n, p in pred.named_parameters()}

    def put(param, value):
        grads[names[id(param)]] = value.reshape(param.shape).float()

    branches = [pred.class_head] + list(pred.reg_features)
    head_lists = [[pred.class_head[2]]] + [list(h) for h in pred.reg_heads]
    nb = len(branches)
    # ---- sigmoid_hm
    g_logits = torch.empty_like(grad_cls)
    call("mf_sigmoid_clamp_bwd", plan.cls.data_ptr(), grad_cls.contiguous().data_ptr(), g_logits.data_ptr(), g_logits.numel(), _st())
    d_hid = torch.zeros(M, nb * hc, dtype=torch.half, device=dev)          # gradient of the activated hidden map [M, 2304]
    # ---- 1x1 heads: weight / bias gradients and d_hid = dY W
    dys = []
    for i, heads in enumerate(head_lists):
        n_i = sum(h.weight.shape[0] for h in heads)
        src = g_logits if i == 0 else grad_reg[:, ctx["ch0s"][i - 1]:ctx["ch0s"][i - 1] + n_i].contiguous()
        dy = torch.zeros(M, 32, dtype=torch.half, device=dev)
        call("mf_nchw_f32_to_nhwc_f16", src.data_ptr(), dy.data_ptr(), B, n_i, HW, 32, _st())
        dys.append(dy)
        dw = torch.empty(32, hc, 1, 1, dtype=torch.float32, device=dev)
        call("mf_conv2d_wgrad_nhwc_f16", hid.ptr() + 2 * i * hc, hid.ld, B, H, W, hc, dy.data_ptr(), 32, 32, 1, 1, 0, dw.data_ptr(), _st())
        db = column_sum(dy)
        r = 0
        for h in heads:
            n = h.weight.shape[0]
            put(h.weight, dw[r:r + n])
            put(h.bias, db[r:r + n])
            r += n
    P = engine.Plan(str(dev))
    dhid_act = P.act(B, H, W, nb * hc)
    dhid_act.buf = d_hid
    for i, heads in enumerate(head_lists):
        wcat = torch.cat([h.weight.detach().float().reshape(h.weight.shape[0], hc) for h in heads], 0)      # [n_i, 256]
        wd = torch.zeros(hc, 32, 1, 1, dtype=torch.float32, device=dev)
        wd[:, :wcat.shape[0], 0, 0] = wcat.t()
        dya = P.act(B, H, W, 32)
        dya.buf = dys[i]
        sl = P.act(B, H, W, hc)
        sl.owner, sl.ch_off = dhid_act, i * hc
        P.conv(dya, wd, 1, 0, None, act=engine.ACT_NONE, out=sl)
    P.finalize()
    P.run()
    # ---- edge-fusion tail
    if pred.enable_edge_fusion:
        K = ctx["K_edge"]
        d_e = []
        for seq, g_map, ctot, ch0, src in ((pred.trunc_heatmap_conv, g_logits, pred.num_classes, 0, ctx["ea"]),
                                          (pred.trunc_offset_conv, grad_reg.contiguous(), pred.num_reg, ctx["off_ch0"], ctx["eb"])):
            entry = ctx["bn"][id(seq[1])]
            t_raw, t_y = entry[0], entry[1]
            n_out = seq[3].weight.shape[0]
            w2 = seq[3].weight.detach().float().reshape(n_out, hc).contiguous()
            d_t = torch.empty(B * K, hc, dtype=torch.half, device=dev)
            dw2 = torch.empty(n_out, hc, dtype=torch.float32, device=dev)
            db2 = torch.empty(n_out, dtype=torch.float32, device=dev)
            call("mf_edge_head_add_bwd", t_y.ptr(), w2.data_ptr(), n_out, plan.edge_idx.data_ptr(), plan.edge_len.data_ptr(),
                 g_map.data_ptr(), ctot, ch0, d_t.data_ptr(), dw2.data_ptr(), db2.data_ptr(), B, K, H, W, _st())
            put(seq[3].weight, dw2)
            put(seq[3].bias, db2)
            d_raw = torch.empty(B * K, hc, dtype=torch.half, device=dev)
            dgam, dbet = _bn_backward(entry, d_t.data_ptr(), hc, d_raw.data_ptr(), hc, 0, B * K, dev, mod=seq[1])
            put(seq[1].weight, dgam)
            put(seq[1].bias, dbet)
            dw1 = torch.empty(hc, hc, 1, 3, dtype=torch.float32, device=dev)
            call("mf_conv2d_wgrad_rect_nhwc_f16", src.ptr(), src.ld, B, 1, K + 2, hc, d_raw.data_ptr(), hc, hc, 1, 3, 1, 0, 0,
                 dw1.data_ptr(), _st())
            put(seq[0].weight, dw1)
            put(seq[0].bias, column_sum(d_raw))
            # Conv1d data gradient: correlation of d_raw with the flipped, transposed taps, 2 columns of zero padding
            w5 = torch.zeros(hc, hc, 5, 3, dtype=torch.float32, device=dev)
            w5[:, :, 2, :] = seq[0].weight.detach().float().flip(2).transpose(0, 1)
            Pe = engine.Plan(str(dev))
            da = Pe.act(B, 1, K, hc)
            da.buf = d_raw
            de = Pe.conv(da, w5, 1, 2, None, act=engine.ACT_NONE)
            Pe.finalize()
            Pe.run()
            d_e.append(de)
        cols = (0, (pred.offset_index[0] + 1) * hc)
        call("mf_edge_gather_bwd", d_e[0].ptr(), d_e[1].ptr(), cols[0], cols[1], plan.edge_idx.data_ptr(), d_hid.data_ptr(), nb * hc,
             B, H, W, K, pred.output_width, pred.output_height, _st())
    # ---- InPlaceABN (leaky 0.01) per branch, in place on the gradient buffer: d_hid -> d_hid_raw
    for i, b in enumerate(branches):
        entry = ctx["bn"][id(b[1])]
        dgam, dbet = _bn_backward(entry, d_hid.data_ptr() + 2 * i * hc, nb * hc, d_hid.data_ptr() + 2 * i * hc, nb * hc, 2, M, dev)
        put(b[1].weight, dgam * torch.sign(b[1].weight.detach()))
        put(b[1].bias, dbet)
    # ---- the nine 3x3 convolutions: one wgrad with Cout = 2304, one dgrad with Cin = 2304
    dw_all = torch.empty(nb * hc, x.C, 3, 3, dtype=torch.float32, device=dev)
    call("mf_conv2d_wgrad_nhwc_f16", x.ptr(), x.ld, B, H, W, x.C, d_hid.data_ptr(), nb * hc, nb * hc, 3, 1, 1, dw_all.data_ptr(), _st())
    for i, b in enumerate(branches):
        put(b[0].weight, dw_all[i * hc:(i + 1) * hc])
    w_all = torch.cat([b[0].weight.detach() for b in branches], 0)
    d_feat = conv2d_dgrad(d_hid, w_all, B, H, W, 1)
    return grads, d_feat
