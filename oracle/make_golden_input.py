"""Generates tests/golden/input_pipeline.npz by executing the UNMODIFIED reference functions of the input path
(/root/reference/model/heatmap_coder.py draw_umich_gaussian / draw_umich_gaussian_2D / gaussian_radius,
/root/reference/data/transforms/transforms.py ToTensor + Normalize, the pad_image method body of data/datasets/kitti.py:218-228)
on oracle.input_oracle.synthetic_case. Run in the build container (needs /root/reference):  python oracle/make_golden_input.py"""
import os
import sys
import types
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, REF)
for name in ("skimage", "skimage.transform"):            # heatmap_coder imports skimage only for get_transfrom_matrix
    sys.modules.setdefault(name, mock.MagicMock())

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_heatmap_coder", os.path.join(REF, "model", "heatmap_coder.py"))
hc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hc)
spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(REF, "data", "transforms", "transforms.py"))
tr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tr)

from PIL import Image  # noqa: E402
from oracle import input_oracle as io  # noqa: E402

H, W, NCLS = 384, 1280, 3
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def ref_pad_image(image, input_height, input_width):
    """body of KITTIDataset.pad_image (kitti.py:218-228) with self.input_* passed in"""
    img = np.array(image)
    h, w, c = img.shape
    ret_img = np.zeros((input_height, input_width, c))
    pad_y = (input_height - h) // 2
    pad_x = (input_width - w) // 2
    ret_img[pad_y: pad_y + h, pad_x: pad_x + w] = img
    pad_size = np.array([pad_x, pad_y])
    return Image.fromarray(ret_img.astype(np.uint8)), pad_size


def main():
    imgs, obj = io.synthetic_case(seed=0)
    tf = tr.Compose([tr.ToTensor(), tr.Normalize(mean=MEAN, std=STD, to_bgr=False)])
    out, pads = [], []
    for im in imgs:
        pil, pad = ref_pad_image(Image.fromarray(im), H, W)
        t, _ = tf(pil, None)
        out.append(t.numpy())
        pads.append(pad)
    hm = np.zeros((obj.shape[0], NCLS, H // 4, W // 4), dtype=np.float32)
    for b in range(obj.shape[0]):
        for valid, cls, cx, cy, rx, ry in obj[b]:
            if not valid:
                continue
            if rx == ry:
                hm[b, cls] = hc.draw_umich_gaussian(hm[b, cls], np.array([cx, cy]), int(rx))
            else:
                hm[b, cls] = hc.draw_umich_gaussian_2D(hm[b, cls], np.array([cx, cy]), int(rx), int(ry))
    radii = np.array([hc.gaussian_radius(hh, ww) for hh, ww in ((10.0, 20.0), (3.5, 7.25), (40.0, 12.0), (1.0, 1.0), (96.0, 300.0))])
    # images are large: keep a strided sample + a checksum per image
    samples = np.stack([o[:, ::7, ::11] for o in out])
    sums = np.array([o.astype(np.float64).sum() for o in out])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "input_pipeline.npz"), image_samples=samples, image_sums=sums,
                        pads=np.stack(pads), hm=hm, radii=radii, obj=obj)
    print("wrote input_pipeline.npz", samples.shape, hm.shape, radii)


if __name__ == "__main__":
    main()
