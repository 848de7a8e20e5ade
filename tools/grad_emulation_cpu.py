"""How far does the gradient of the synthetic detector move when ONLY THE FORWARD is rounded to fp16?

The CUDA training step computes its forward with fp16 operands / activations (fp32 accumulation) and the reference in fp32. On
this synthetic network (random weights + batch-statistics BatchNorm = the chaotic regime of a deep ReLU net) the derivative is
ill-conditioned with respect to the point it is evaluated at: this script evaluates torch autograd - an EXACT fp32 backward -
through the oracle forward with the fp16 rounding points of the kernels emulated (oracle/f16_emulation.py) and compares the
gradients with the unmodified reference's (tests/golden/train_step_2x384x1280.npz), element-wise. The cosines it prints are the
yardstick for the backward tape's element-wise test (tests/test_gpu_train.py::test_reference_training_loop_unchanged): a tape
that matches this yardstick is as close to the reference as ANY backward evaluated on an fp16 forward can be.

    python tools/grad_emulation_cpu.py          # ~4 min on 16 threads; writes tests/golden/grad_cos_f16_forward_emulation.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import monoflex_oracle as mo                     # noqa: E402
from oracle.f16_emulation import GROUPS, Emu                 # noqa: E402
from monoflex_b200 import synthetic as syn                   # noqa: E402


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "train_step_2x384x1280.npz"))
    skip = ("grad_names", "grad_norms", "grad_features_norm", "grad_features_sample")
    full = [k[5:] for k in gold.files if k.startswith("grad_") and k not in skip]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v)
          for k, v in syn.make_state_dict(seed=0).items()}
    fields = syn.make_train_targets(2, empty_image=0)
    images = syn.make_images(2, 384, 1280, seed=1)
    idx, n, _ = syn.edge_indices()
    emu = Emu({g: "f16" for g in GROUPS})
    emu.names = {id(v): k for k, v in sd.items()}
    emu.install()
    try:
        loss, _ = mo.detector_train_losses(sd, images.half().float(), fields, idx.unsqueeze(0).repeat(2, 1, 1), torch.tensor([n, n]),
                                           [syn.KITTI_P2] * 2)
    finally:
        emu.uninstall()
    total = sum(loss.values())
    total.backward()
    out = {"total_loss": float(total.item()), "reference_total_loss": float(gold["total"]), "cos": {}, "rel_l2": {}}
    for name in full:
        g, w = sd[name].grad.double().flatten(), torch.from_numpy(gold["grad_" + name]).double().flatten()
        out["cos"][name] = float((g @ w) / (g.norm() * w.norm()))
        out["rel_l2"][name] = float((g - w).norm() / w.norm())
        print("%-58s cos %.5f rel-l2 %.4f" % (name, out["cos"][name], out["rel_l2"][name]))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "grad_cos_f16_forward_emulation.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
