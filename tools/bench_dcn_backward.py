"""Timing of the boundary-B operator pair `_ext.dcn_v2_forward/backward` (exact fp32, row R5) at the detector's DCN shapes.
Prints one JSON line per shape: ms (CUDA events, median of 5 after 2 warm-ups) and effective TFLOP/s counting
fwd 2*M*Co*9C and bwd 2x that (dgrad + wgrad contractions)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoflex_b200.model.backbone.DCNv2 import _ext            # noqa: E402

SHAPES = [(8, 64, 64, 96, 320), (8, 128, 128, 48, 160), (8, 256, 256, 24, 80), (8, 512, 256, 12, 40)]


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    only = int(os.environ.get("ONLY", "-1"))
    for i, (B, C, Co, H, W) in enumerate(SHAPES):
        if only >= 0 and i != only:
            continue
        g = torch.Generator(device="cuda").manual_seed(i)
        x = torch.randn(B, C, H, W, device="cuda", generator=g)
        w = torch.randn(Co, C, 3, 3, device="cuda", generator=g) * 0.05
        b = torch.randn(Co, device="cuda", generator=g) * 0.1
        off = torch.randn(B, 18, H, W, device="cuda", generator=g) * 1.5
        m = torch.sigmoid(torch.randn(B, 9, H, W, device="cuda", generator=g))
        dy = torch.randn(B, Co, H, W, device="cuda", generator=g)
        f_ms = timed(lambda: _ext.dcn_v2_forward(x, w, b, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 1))
        b_ms = timed(lambda: _ext.dcn_v2_backward(x, w, b, off, m, dy, 3, 3, 1, 1, 1, 1, 1, 1, 1))
        fl = 2.0 * B * H * W * Co * 9 * C
        print(json.dumps({"shape": [B, C, Co, H, W], "forward_ms": f_ms, "backward_ms": b_ms,
                          "forward_TFLOPs": fl / f_ms / 1e9, "backward_TFLOPs": 2 * fl / b_ms / 1e9}), flush=True)


if __name__ == "__main__":
    main()
