"""TEST/WORKLOAD TOOLING — generates monoflex_b200/synthetic_calib.npz (16 scalars).

The synthetic weights follow SURVEY.md §8d (random BN statistics, non-zero conv_offset_mask). With random
weights the activation scale drifts from level to level, so a fixed std for `conv_offset_mask.weight` gives
sub-pixel offsets in shallow layers and tens of pixels in deep ones. A trained DCN predicts offsets of a few
pixels; this script runs the CPU oracle once and stores, per DCN layer, the factor that brings the rms offset
to ~1.5 px. `synthetic.make_state_dict` multiplies the offset-conv weights by it, so every machine sees
bit-identical weights. Run from the repo root:  python -m oracle.calibrate_offsets
(BN statistics are NOT calibrated: mean removal puts a random ReLU net in the chaotic phase where rounding
noise is amplified ~1.4x per layer — measured — which says nothing about kernel correctness.)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import monoflex_oracle as mo          # noqa: E402
from monoflex_b200 import synthetic as syn        # noqa: E402

TARGET_RMS_PX = 1.5


def main(h=192, w=640, batch=1):
    sd = syn.make_state_dict(0, calibrated=False)
    scales = {}
    orig_conv = F.conv2d

    def conv(x, weight, bias=None, *a, **k):
        if weight.shape[0] == 27:
            key = [n for n, v in sd.items() if v is weight][0]
            y = orig_conv(x, weight, None, *a, **k)
            s = TARGET_RMS_PX / float(y[:, :18].pow(2).mean().sqrt())
            scales[key] = np.float32(s)
            weight.mul_(s)
        return orig_conv(x, weight, bias, *a, **k)

    F.conv2d = conv
    try:
        with torch.no_grad():
            mo.backbone(sd, syn.make_images(batch, h, w, seed=77))
    finally:
        F.conv2d = orig_conv
    out = os.path.join(os.path.dirname(os.path.abspath(syn.__file__)), 'synthetic_calib.npz')
    np.savez(out, **scales)
    for k, v in scales.items():
        print('%-60s %.4g' % (k, v))
    print('wrote', out, len(scales), 'scalars')


if __name__ == '__main__':
    main()
