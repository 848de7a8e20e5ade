"""Deterministic synthetic KITTI-shaped workload: weights, images, edge indices, calibration.

Everything is drawn from a numpy PCG64 stream so that the build container, the GPU box, the tests and
bench.py all see bit-identical tensors without any dataset or checkpoint (there is no network).
Shapes follow the reference: data/datasets/kitti.py:126-179 (edge indices), :218-228 (padding),
config/defaults.py (sizes), runs/monoflex.yaml (heads).
"""
import math

import numpy as np
import torch

# typical KITTI P2
KITTI_P2 = ((721.5377, 0.0, 609.5593, 44.85728),
            (0.0, 721.5377, 172.854, 0.2163791),
            (0.0, 0.0, 1.0, 0.002745884))
PAD_SIZE = (19, 4)                       # 1242x375 image centred in 1280x384 (kitti.py:222-223)

REG_BRANCH_CH = [[4], [2], [20], [3], [3], [8, 8], [1], [1]]   # runs/monoflex.yaml:28
NUM_CLASSES = 3
HEAD_CONV = 256


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def dla34_param_shapes():
    """(key, shape, kind) for every tensor of KeypointDetector.state_dict() in reference order-independent form.
    kind in {conv, bn_w, bn_b, bn_m, bn_v, bn_n, bias, up, offw, offb, dcnw, dcnb, conv1d}."""
    out = []

    def conv(p, co, ci, k):
        out.append((p + '.weight', (co, ci, k, k), 'conv'))

    def bn(p, c):
        out.extend([(p + '.weight', (c,), 'bn_w'), (p + '.bias', (c,), 'bn_b'), (p + '.running_mean', (c,), 'bn_m'),
                    (p + '.running_var', (c,), 'bn_v'), (p + '.num_batches_tracked', (), 'bn_n')])

    def block(p, ci, co):
        conv(p + '.conv1', co, ci, 3); bn(p + '.bn1', co); conv(p + '.conv2', co, co, 3); bn(p + '.bn2', co)

    def tree(p, levels, ci, co, level_root, root_dim=0):
        if root_dim == 0:
            root_dim = 2 * co
        if level_root:
            root_dim += ci
        if levels == 1:
            block(p + '.tree1', ci, co); block(p + '.tree2', co, co)
            conv(p + '.root.conv', co, root_dim, 1); bn(p + '.root.bn', co)
        else:
            tree(p + '.tree1', levels - 1, ci, co, False, 0)
            tree(p + '.tree2', levels - 1, co, co, False, root_dim + co)
        if ci != co:
            conv(p + '.project.0', co, ci, 1); bn(p + '.project.1', co)

    b = 'backbone.base'
    conv(b + '.base_layer.0', 16, 3, 7); bn(b + '.base_layer.1', 16)
    conv(b + '.level0.0', 16, 16, 3); bn(b + '.level0.1', 16)
    conv(b + '.level1.0', 32, 16, 3); bn(b + '.level1.1', 32)
    ch, lv = [16, 32, 64, 128, 256, 512], [1, 1, 1, 2, 2, 1]
    for i in range(2, 6):
        tree('%s.level%d' % (b, i), lv[i], ch[i - 1], ch[i], i > 2)

    def dcn(p, ci, co):
        out.append((p + '.conv.weight', (co, ci, 3, 3), 'dcnw'))
        out.append((p + '.conv.bias', (co,), 'dcnb'))
        out.append((p + '.conv.conv_offset_mask.weight', (27, ci, 3, 3), 'offw'))
        out.append((p + '.conv.conv_offset_mask.bias', (27,), 'offb'))
        bn(p + '.actf.0', co)

    def ida(p, o, chans, up_f):
        for i in range(1, len(chans)):
            dcn('%s.proj_%d' % (p, i), chans[i], o)
            out.append(('%s.up_%d.weight' % (p, i), (o, 1, 2 * up_f[i], 2 * up_f[i]), 'up'))
            dcn('%s.node_%d' % (p, i), o, o)

    ida('backbone.dla_up.ida_0', 256, [256, 512], [1, 2])
    ida('backbone.dla_up.ida_1', 128, [128, 256, 256], [1, 2, 2])
    ida('backbone.dla_up.ida_2', 64, [64, 128, 128, 128], [1, 2, 2, 2])
    ida('backbone.ida_up', 64, [64, 128, 256], [1, 2, 4])

    h = 'heads.predictor'
    conv(h + '.class_head.0', HEAD_CONV, 64, 3)
    for s, k in (('weight', 'bn_w'), ('bias', 'bn_b'), ('running_mean', 'bn_m'), ('running_var', 'bn_v')):
        out.append(('%s.class_head.1.%s' % (h, s), (HEAD_CONV,), k))
    out.append((h + '.class_head.2.weight', (NUM_CLASSES, HEAD_CONV, 1, 1), 'conv'))
    out.append((h + '.class_head.2.bias', (NUM_CLASSES,), 'clsbias'))
    for i, br in enumerate(REG_BRANCH_CH):
        conv('%s.reg_features.%d.0' % (h, i), HEAD_CONV, 64, 3)
        for s, k in (('weight', 'bn_w'), ('bias', 'bn_b'), ('running_mean', 'bn_m'), ('running_var', 'bn_v')):
            out.append(('%s.reg_features.%d.1.%s' % (h, i, s), (HEAD_CONV,), k))
        for j, c in enumerate(br):
            out.append(('%s.reg_heads.%d.%d.weight' % (h, i, j), (c, HEAD_CONV, 1, 1), 'conv'))
            out.append(('%s.reg_heads.%d.%d.bias' % (h, i, j), (c,), 'bias'))
    for name, co in (('trunc_heatmap_conv', NUM_CLASSES), ('trunc_offset_conv', 2)):
        q = '%s.%s' % (h, name)
        out.append((q + '.0.weight', (HEAD_CONV, HEAD_CONV, 3), 'conv1d'))
        out.append((q + '.0.bias', (HEAD_CONV,), 'bias'))
        bn(q + '.1', HEAD_CONV)
        out.append((q + '.3.weight', (co, HEAD_CONV, 1), 'conv1d'))
        out.append((q + '.3.bias', (co,), 'bias'))
    return out


_CALIB = None


def _calibration():
    global _CALIB
    if _CALIB is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'synthetic_calib.npz')
        with np.load(path) as z:
            _CALIB = {k: z[k] for k in z.files}
    return _CALIB


def make_state_dict(seed=0, calibrated=True):
    """Random-init weights of the reference architecture (same keys/shapes as the reference state_dict).
    He-normal convs, BN/IABN stats randomised (gamma>0 so |gamma| == gamma), non-zero
    conv_offset_mask, bilinear up-sampling kernels (fill_up_weights dla_dcn.py:372)."""
    g = _rng(seed)
    sd = {}
    for key, shape, kind in dla34_param_shapes():
        if kind in ('conv', 'dcnw', 'conv1d'):
            fan_in = int(np.prod(shape[1:]))
            # residual branches sum two unit-variance signals: keep the variance stationary (gain 1 not 2)
            gain = 1.0 if key.endswith('conv2.weight') else 2.0
            if '.reg_heads.' in key or key.endswith('class_head.2.weight') or key.endswith('.3.weight'):
                gain = 0.125   # keep logits / regression outputs O(1): no sigmoid saturation ties
            v = g.standard_normal(shape) * math.sqrt(gain / fan_in)
        elif kind == 'offw':
            # offsets of ~1 px for unit-variance inputs whatever Cin (real DCN offsets are a few pixels)
            v = g.standard_normal(shape) / math.sqrt(int(np.prod(shape[1:])))
        elif kind == 'offb':
            v = g.standard_normal(shape) * 0.1
        elif kind == 'bn_w':
            v = g.uniform(0.5, 1.5, shape)
        elif kind == 'bn_b':
            v = g.standard_normal(shape) * 0.1
        elif kind == 'bn_m':
            v = g.standard_normal(shape) * 0.1
        elif kind == 'bn_v':
            v = g.uniform(0.5, 1.5, shape)
        elif kind == 'bn_n':
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        elif kind in ('bias', 'dcnb'):
            v = g.standard_normal(shape) * 0.05
        elif kind == 'clsbias':
            v = np.full(shape, -math.log(1 / 0.01 - 1))
        elif kind == 'up':
            k = shape[2]
            f = math.ceil(k / 2)
            c = (2 * f - 1 - f % 2) / (2.0 * f)
            w1 = np.array([1 - abs(i / f - c) for i in range(k)])
            v = np.broadcast_to(np.outer(w1, w1), shape) * g.uniform(0.9, 1.1, (shape[0], 1, 1, 1))
        else:
            raise KeyError(kind)
        sd[key] = _t(v)
    if calibrated:
        # per-DCN factor bringing the rms offset to ~1.5 px (oracle/calibrate_bn.py), as in a trained net
        for k, v in _calibration().items():
            sd[k] = sd[k] * float(v)
    return sd


def make_images(batch, height=384, width=1280, seed=1):
    """randn images ~ normalised KITTI (config/defaults.py:34-36)."""
    return _t(_rng(seed).standard_normal((batch, 3, height, width)))


def edge_indices(out_w=320, out_h=96, pad=PAD_SIZE, down_ratio=4):
    """Border pixel list of the un-padded image on the stride-4 map, (x,y) int64 [K_max,2] zero padded, and
    edge_len = count-1 (data/datasets/kitti.py:126-179, 279-285)."""
    img_w, img_h = out_w * down_ratio - 2 * pad[0], out_h * down_ratio - 2 * pad[1] - 1
    x_min, y_min = math.ceil(pad[0] / down_ratio), math.ceil(pad[1] / down_ratio)
    x_max, y_max = (pad[0] + img_w - 1) // down_ratio, (pad[1] + img_h - 1) // down_ratio
    pts = [(x_min, y) for y in range(y_min, y_max)]
    pts += [(x, y_max) for x in range(x_min, x_max)]
    pts += [(x_max, y) for y in range(y_max, y_min, -1)]
    pts += [(x, y_min) for x in range(x_max, x_min - 1, -1)]
    k_max = (out_w + out_h) * 2
    idx = torch.zeros(k_max, 2, dtype=torch.long)
    idx[:len(pts)] = torch.tensor(pts, dtype=torch.long)
    return idx, len(pts) - 1, (img_w, img_h)


def make_targets(batch, out_w=320, out_h=96):
    """Per-image inference-time target fields the hot path reads (detector_infer.py:53-58,
    detector_predictor.py:138-139): calib P, pad_size, size, edge_indices, edge_len."""
    idx, n, img_size = edge_indices(out_w, out_h)
    return dict(edge_indices=idx.unsqueeze(0).repeat(batch, 1, 1), edge_len=torch.full((batch,), n, dtype=torch.long),
                calib_P=[KITTI_P2 for _ in range(batch)],
                pad_size=torch.tensor([PAD_SIZE] * batch, dtype=torch.float32).view(batch, 2),
                size=[(out_w * 4, out_h * 4)] * batch)   # ParamsList.size = PADDED image (kitti.py:271,289)


def make_head_logits(batch, out_w=320, out_h=96, seed=2):
    """Decode-only workload: heat-map logits ~N(-3,1.5) -> >>50 distinct local maxima per class;
    regression map ~N(0,0.5)."""
    g = _rng(seed)
    cls = _t(g.standard_normal((batch, NUM_CLASSES, out_h, out_w)) * 1.5 - 3.0)
    reg = _t(g.standard_normal((batch, 50, out_h, out_w)) * 0.5)
    return cls, reg


def make_param_lists(tg):
    """The per-image ParamsList objects engine/inference.py hands to model(images, targets)."""
    from .structures import Calibration, ParamsList
    out = []
    for b in range(len(tg['calib_P'])):
        t = ParamsList(tg['size'][b], is_train=False)
        t.add_field('calib', Calibration(tg['calib_P'][b]))
        t.add_field('pad_size', tg['pad_size'][b])
        t.add_field('edge_indices', tg['edge_indices'][b])
        t.add_field('edge_len', tg['edge_len'][b])
        out.append(t)
    return out
