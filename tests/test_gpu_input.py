"""-m gpu: GPU input pipeline (csrc/mf_input.cu, monoflex_b200/data.py) against the oracle restatement of the
reference's pad_image + ToTensor + Normalize and heat-map drawing (oracle/input_oracle.py, pinned to the unmodified reference
by tests/golden/input_pipeline.npz). Images: bit-exact (IEEE fp32 divisions). Heat maps: the integer structure (support,
peaks == 1 at every centre, untouched pixels == 0) exactly, values within 1 fp32 ulp (device exp vs glibc exp in double)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from monoflex_b200.config import default_cfg
from monoflex_b200.data import GpuInputPipeline, heatmap_radii
from oracle import input_oracle as io

pytestmark = pytest.mark.gpu
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def test_preprocess_images_bit_exact_vs_oracle_and_golden():
    cfg = default_cfg()
    pipe = GpuInputPipeline(cfg)
    imgs, obj = io.synthetic_case(seed=0, B=5)
    flips = [False, True, False, True, False]
    out, pads = pipe.images([torch.from_numpy(im).pin_memory() for im in imgs], flips=flips)
    torch.cuda.synchronize()
    with np.load(os.path.join(GOLDEN, "input_pipeline.npz")) as z:
        gold = {k: z[k] for k in z.files}
    for b, im in enumerate(imgs):
        src = im[:, ::-1] if flips[b] else im                        # the reference flips the PIL image before padding
        padded, pad = io.pad_image(np.ascontiguousarray(src), 384, 1280)
        ref = io.to_tensor_normalize(padded, MEAN, STD)
        assert torch.equal(out[b].cpu(), ref), b
        assert pads[b].tolist() == pad.tolist()
        if b < 3 and not flips[b]:
            assert np.array_equal(out[b].cpu().numpy()[:, ::7, ::11], gold["image_samples"][b])
    # feeds the detector unchanged: same tensor type / shape as the reference's collate output
    assert out.shape == (5, 3, 384, 1280) and out.dtype == torch.float32


def test_heatmaps_vs_oracle_and_golden():
    cfg = default_cfg()
    pipe = GpuInputPipeline(cfg)
    for seed, B in ((0, 3), (7, 8)):
        _, obj = io.synthetic_case(seed=seed, B=B)
        hm = pipe.heatmaps(torch.from_numpy(obj)).cpu().numpy()
        ref = io.draw_heatmaps(obj, 3, 96, 320)
        assert np.array_equal(hm == 0, ref == 0)                      # identical support
        assert np.array_equal(hm == 1, ref == 1)                      # peaks
        ulp = np.abs(hm.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, ulp.max()
        for b in range(B):
            for valid, cls, cx, cy, rx, ry in obj[b]:
                if valid:
                    assert hm[b, cls, cy, cx] == 1.0
    with np.load(os.path.join(GOLDEN, "input_pipeline.npz")) as z:
        gold_hm, gold_obj = z["hm"], z["obj"]
    hm = pipe.heatmaps(torch.from_numpy(gold_obj)).cpu().numpy()
    assert np.abs(hm.view(np.int32).astype(np.int64) - gold_hm.view(np.int32).astype(np.int64)).max() <= 1


def test_heatmap_radii_host_logic():
    # inside object: circular radius from the float64 gaussian_radius; edge object: one-sided
    assert heatmap_radii([10, 10, 30, 20], (20, 15), approx_center=False) == (3, 3)
    rx, ry = heatmap_radii([0, 10, 12, 40], (0, 25), approx_center=True)
    assert rx == 0 and ry == 7
