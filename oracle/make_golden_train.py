"""TEST TOOLING — tests/golden/train_step_2x384x1280.npz: the UNMODIFIED reference KeypointDetector in TRAIN mode
(model/detector.py:32-34: backbone -> predictor -> Loss_Computation) on the synthetic batch, followed by
`sum(losses).backward()` (engine/trainer.py:109-117). Stores the 11 losses, the L2 norm of every parameter gradient and three
small gradient tensors in full. Pins oracle.detector_train_losses (tests/test_oracle_golden.py) and, through it, the CUDA
training operators. Build container only:  python -m oracle.make_golden_train"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims as rs            # noqa: E402
from oracle.make_golden_loss import ref_train_targets   # noqa: E402
from monoflex_b200 import synthetic as syn    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FULL = ["backbone.base.base_layer.0.weight", "backbone.base.level2.tree1.bn1.weight", "heads.predictor.class_head.2.bias",
        # round 2: element-wise gradient checks along the whole depth of the network (small tensors, stored in full)
        "backbone.base.level0.0.weight", "backbone.base.level1.0.weight", "backbone.base.level1.1.weight",
        "backbone.base.level2.tree1.conv1.weight", "backbone.base.level3.tree1.tree1.bn1.bias",
        "backbone.base.level5.root.bn.weight", "backbone.dla_up.ida_2.up_3.weight",
        "backbone.ida_up.node_1.conv.weight", "backbone.ida_up.node_1.conv.bias",
        "backbone.ida_up.node_1.conv.conv_offset_mask.bias", "backbone.ida_up.node_1.actf.0.weight",
        "heads.predictor.class_head.1.weight", "heads.predictor.reg_heads.2.0.weight", "heads.predictor.trunc_heatmap_conv.1.bias"]


def main(batch=2):
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    cfg = rs.reference_cfg()
    import model.head.detector_loss as dl
    dl.get_iou_3d = lambda a, b: torch.zeros(a.shape[0])
    model = rs.build_reference_model(cfg)
    sd = syn.make_state_dict(seed=0)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.missing_keys, missing.missing_keys[:5]
    model.train()
    fields = syn.make_train_targets(batch, empty_image=0)
    images = syn.make_images(batch, 384, 1280, seed=1)
    targets = ref_train_targets(fields)
    idx, n, _ = syn.edge_indices()
    for t in targets:
        t.add_field("edge_indices", idx)
        t.add_field("edge_len", torch.tensor(n, dtype=torch.long))
    feats = {}

    def keep(_m, _inp, out):                       # gradient w.r.t. the backbone's output feature map (head-backward check)
        out.retain_grad()
        feats["f"] = out
    hook = model.backbone.register_forward_hook(keep)
    loss_dict, log = model(images, targets)
    hook.remove()
    total = sum(v for v in loss_dict.values())
    total.backward()
    out = {"loss_" + k: np.float32(v.item()) for k, v in loss_dict.items()}
    out["total"] = np.float32(total.item())
    names, norms = [], []
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    gf = feats["f"].grad
    out["grad_features_norm"] = np.float64(gf.double().norm())
    out["grad_features_sample"] = gf.reshape(-1)[::997].numpy().astype(np.float32)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    for k, p in model.named_parameters():
        if k in FULL:
            out["grad_" + k] = p.grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "train_step_%dx384x1280.npz" % batch), **out)
    for k in sorted(out):
        if np.ndim(out[k]) == 0:
            print(k, out[k])
    print("params with grad:", sum(1 for v in norms if v > 0), "of", len(norms))


if __name__ == "__main__":
    main()
