"""Diagnostics: element-wise fidelity of the backward tape against the unmodified reference's train-mode gradients stored in full
in tests/golden/train_step_2x384x1280.npz, for several loss scales:  python tools/grad_fidelity.py [scale ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoflex_b200 import synthetic as syn                      # noqa: E402
from monoflex_b200.config import default_cfg                    # noqa: E402
from monoflex_b200.model.detector import KeypointDetector       # noqa: E402

gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "train_step_2x384x1280.npz"))
full = [k[5:] for k in gold.files if k.startswith("grad_") and k not in ("grad_names", "grad_norms", "grad_features_norm", "grad_features_sample")]
scales = [float(a) for a in sys.argv[1:]] or [128.0, 2048.0, 32768.0]
fields = syn.make_train_targets(2, empty_image=0)
images = syn.make_images(2, 384, 1280, seed=1).cuda()
targets = [t.to("cuda") for t in syn.make_train_param_lists(fields)]
for S in scales:
    model = KeypointDetector(default_cfg()).cuda()
    model.load_state_dict(syn.make_state_dict(seed=0), strict=False)
    model.train()
    model.loss_scale = S
    loss_dict, _ = model(images, targets)
    sum(loss_dict.values()).backward()
    params = dict(model.named_parameters())
    print("loss scale %g, total loss %.4f (ref %.4f)" % (S, sum(v.item() for v in loss_dict.values()), float(gold["total"])))
    for n in full:
        g, w = params[n].grad.detach().double().cpu().flatten(), torch.from_numpy(gold["grad_" + n]).double().flatten()
        cos = float((g @ w) / (g.norm() * w.norm()).clamp_min(1e-300))
        print("   %-58s n=%6d cos %.5f  rel-l2 %.4f  |g|/|ref| %.4f  finite %s" % (n, g.numel(), cos, float((g - w).norm() / w.norm()),
                                                                                 float(g.norm() / w.norm()), bool(torch.isfinite(g).all())))
