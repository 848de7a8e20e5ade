"""CPU emulation of the rounding points of the CUDA forward (DESIGN.md §4 "Precision modes"): which layer groups need hi/lo pair
operands for the 1e-3 end-to-end contract. TEST/DESIGN TOOLING: imports the oracle, never imported by the product.

    python tools/precision_emulation.py [H W]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.f16_emulation import GROUPS, Emu      # noqa: E402
from monoflex_b200 import synthetic as syn        # noqa: E402


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
