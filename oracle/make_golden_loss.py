"""TEST TOOLING — tests/golden/loss_4x96x320.npz: the UNMODIFIED reference Loss_Computation (model/head/detector_loss.py, via
oracle/ref_shims.py) on the synthetic training targets / predictions of monoflex_b200/synthetic.py, plus autograd gradients of
the summed loss (engine/trainer.py:109-110). Build container only:  python -m oracle.make_golden_loss
shapely is mocked, so the logging-only '3D_IoU' (detector_loss.py:333) is patched to zeros and not recorded."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims as rs            # noqa: E402
from monoflex_b200 import synthetic as syn    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_train_targets(fields):
    from structures.params_3d import ParamsList
    from data.datasets.kitti_utils import Calibration
    out = []
    for f in fields:
        t = ParamsList((1280, 384), is_train=True)
        for k, v in f.items():
            t.add_field("2d_bboxes" if k == "bboxes" else k, torch.from_numpy(np.ascontiguousarray(v)))
        c = Calibration.__new__(Calibration)
        P = np.array(syn.KITTI_P2)
        c.P = P
        c.c_u, c.c_v, c.f_u, c.f_v = P[0, 2], P[1, 2], P[0, 0], P[1, 1]
        c.b_x, c.b_y = P[0, 3] / (-c.f_u), P[1, 3] / (-c.f_v)
        t.add_field("calib", c)
        out.append(t)
    return out


def main(batch=4):
    cfg = rs.reference_cfg()
    import model.head.detector_loss as dl
    dl.get_iou_3d = lambda a, b: torch.zeros(a.shape[0])
    lc = dl.Loss_Computation(cfg)
    fields = syn.make_train_targets(batch)
    cls, reg = syn.make_train_predictions(batch, fields)
    cls.requires_grad_(True)
    reg.requires_grad_(True)
    loss_dict, log = lc({"cls": cls, "reg": reg}, ref_train_targets(fields))
    total = sum(v for v in loss_dict.values())
    total.backward()
    out = {"loss_" + k: np.float32(v.item()) for k, v in loss_dict.items()}
    out.update({"log_" + k: np.float32(v) for k, v in log.items() if k != "3D_IoU"})
    centers = np.stack([f["target_centers"] for f in fields])
    mask = np.stack([f["reg_mask"] for f in fields]).astype(bool)
    g = reg.grad.numpy()
    rows = [g[b, :, centers[b, i, 1], centers[b, i, 0]] for b in range(batch) for i in range(mask.shape[1]) if mask[b, i]]
    out["grad_reg_at_centers"] = np.stack(rows).astype(np.float32)
    out["grad_reg_abs_sum"] = np.float64(np.abs(g).sum())
    gc = cls.grad.numpy().reshape(-1)
    out["grad_cls_sample"] = gc[::97].astype(np.float32)
    out["grad_cls_abs_sum"] = np.float64(np.abs(gc).sum())
    out["total"] = np.float32(total.item())
    np.savez_compressed(os.path.join(OUT, "loss_%dx96x320.npz" % batch), **out)
    for k in sorted(out):
        if np.ndim(out[k]) == 0:
            print(k, out[k])


if __name__ == "__main__":
    main()
