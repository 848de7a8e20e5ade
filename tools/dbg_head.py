import os, sys, torch
sys.path.insert(0, '/root/repo')
from monoflex_b200 import synthetic as syn
from monoflex_b200.config import default_cfg
from monoflex_b200.model.detector import KeypointDetector
H, W, B = 128, 256, 1
sd = syn.make_state_dict(0)
m = KeypointDetector(default_cfg(width=W, height=H)); m.load_state_dict(sd); m = m.cuda().eval()
x = syn.make_images(B, H, W).cuda()
tg = syn.make_targets(B, W // 4, H // 4)
targets = [t.to("cuda") for t in syn.make_param_lists(tg)]
with torch.no_grad():
    feats = m.backbone(x)
    os.environ["MF_NO_FUSED_HEAD"] = "1"
    p0 = m.heads.predictor(feats, targets); c0, r0 = p0['cls'].clone(), p0['reg'].clone()
    hid0 = m.heads.predictor.last_plan.hidden.nchw_view().float().clone()
    os.environ["MF_NO_FUSED_HEAD"] = "0"
    m.heads.predictor._plans = {}
    p1 = m.heads.predictor(feats, targets); c1, r1 = p1['cls'], p1['reg']
    torch.cuda.synchronize()
    print('cls maxabs', (c0 - c1).abs().max().item(), c0.abs().max().item())
    offs = [0, 4, 6, 26, 29, 32, 48, 49, 50]
    for i in range(8):
        a, b = r0[:, offs[i]:offs[i+1]], r1[:, offs[i]:offs[i+1]]
        print('reg branch', i, (a - b).abs().max().item(), a.abs().max().item())
    hb = m.heads.predictor.last_plan.keep
    hid_buf = [t for t in hb if isinstance(t, torch.Tensor) and t.dtype == torch.half and t.dim() == 2 and t.shape[1] == 512][0]
    h1 = hid_buf.view(B, H // 4, W // 4, 512).permute(0, 3, 1, 2).float()
    print('hid cls branch', (h1[:, :256] - hid0[:, :256]).abs().max().item(), hid0[:, :256].abs().max().item())
    print('hid off branch', (h1[:, 256:] - hid0[:, 512:768]).abs().max().item())
    # where are errors in cls? per pixel row pattern
    d = (c0 - c1).abs()[0, 0]
    print('cls err rows', d.max(1).values[:8].tolist())
    print('cls err by tile (128 px)', d.reshape(-1, 128).max(1).values.tolist()[:16])
    print('sample c0', c0[0, 0, 0, :4].tolist(), 'c1', c1[0, 0, 0, :4].tolist())
    # per-channel mean difference vs the 1x1 biases
    pr = m.heads.predictor
    print('reg per-channel mean diff (fused - ref):')
    print([round(v, 3) for v in (r1 - r0).mean(dim=(0, 2, 3)).tolist()])
    bias_all = torch.cat([h.bias for heads in pr.reg_heads for h in heads]).tolist()
    print('reg biases:')
    print([round(v, 3) for v in bias_all])
    print('cls bias', pr.class_head[2].bias.tolist())
    # interior pixels only (no edge fusion): logit diff of cls
    lc0 = torch.log(c0 / (1 - c0)); lc1 = torch.log(c1 / (1 - c1))
    print('cls logit mean diff per ch (interior)', (lc1 - lc0)[:, :, 8:20, 20:40].mean(dim=(0, 2, 3)).tolist())
    print('reg std of diff per channel', [round(v, 3) for v in (r1 - r0)[:, :, 8:20, 20:40].std(dim=(0, 2, 3)).tolist()])
