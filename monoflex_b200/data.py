"""GPU input pipeline: the per-image host work of the reference's loader that a 1-2 k img/s/GPU detector
cannot be fed by - `KITTIDataset.pad_image` (data/datasets/kitti.py:218-228), `ToTensor` + `Normalize`
(data/transforms/transforms.py:14-30) and the heat-map drawing of the target encoding (model/heatmap_coder.py:37-64, 83-124,
call sites data/datasets/kitti.py:449-463) - as two CUDA kernels (csrc/mf_input.cu) behind a small host API.

    pipe = GpuInputPipeline(cfg)
    images, pad_sizes = pipe.images(list_of_uint8_HWC_tensors, flips=None)     # -> [B,3,H,W] fp32 on the GPU, [B,2] int
    hm = pipe.heatmaps(obj6)                                                   # int32 [B,max_objs,6] -> [B,3,H/4,W/4] fp32

What stays on the host is O(objects) scalar work: which objects are valid, their integer centres and radii
(`heatmap_radii`, the reference's float64 `gaussian_radius`). There is no CPU fallback: tensors must be CUDA (or pinned
host memory for the uint8 images, which are copied with one async H2D each)."""
import numpy as np
import torch

from ._lib import call, stream


def gaussian_radius(height, width, min_overlap=0.7):
    """model/heatmap_coder.py:37-54 (float64, like the reference's numpy scalars)."""
    a1 = 1
    b1 = (height + width)
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * a1 * c1)) / 2
    a2 = 4
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 4 * a2 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def heatmap_radii(box2d, target_center, approx_center, edge_heatmap_ratio=0.5, adjust_edge_heatmap=True):
    """(rx, ry) of one object exactly as data/datasets/kitti.py:449-463 computes them: a circular Gaussian of
    `gaussian_radius(box h, box w)` for inside objects, a one-sided Gaussian (one radius 0) for objects whose projected centre
    was moved to the image border (`approx_center`). box2d / target_center are in feature-map (stride 4) coordinates."""
    box2d = np.asarray(box2d, dtype=np.float64)
    if adjust_edge_heatmap and approx_center:
        bw = min(target_center[0] - box2d[0], box2d[2] - target_center[0])
        bh = min(target_center[1] - box2d[1], box2d[3] - target_center[1])
        rx, ry = max(0, int(bw * edge_heatmap_ratio)), max(0, int(bh * edge_heatmap_ratio))
        if min(rx, ry) != 0:
            raise ValueError("edge object with two non-zero radii (kitti.py:457 asserts this cannot happen)")
        return rx, ry
    dim = box2d[2:] - box2d[:2]
    r = max(0, int(gaussian_radius(dim[1], dim[0])))
    return r, r


class GpuInputPipeline(object):
    def __init__(self, cfg, device="cuda"):
        inp = cfg.INPUT
        self.H, self.W = int(inp.HEIGHT_TRAIN), int(inp.WIDTH_TRAIN)
        self.down = int(cfg.MODEL.BACKBONE.DOWN_RATIO)
        self.mean = (np.ctypeslib.ctypes.c_float * 3)(*[float(v) for v in getattr(inp, "PIXEL_MEAN", (0.485, 0.456, 0.406))])
        self.std = (np.ctypeslib.ctypes.c_float * 3)(*[float(v) for v in getattr(inp, "PIXEL_STD", (0.229, 0.224, 0.225))])
        self.to_bgr = 1 if getattr(inp, "TO_BGR", False) else 0
        self.ncls = len(cfg.DATASETS.DETECT_CLASSES)
        self.max_objs = int(cfg.DATASETS.MAX_OBJECTS)
        self.device = torch.device(device)

    def images(self, imgs, flips=None, out=None):
        """imgs: list of uint8 [h, w, 3] tensors (CUDA, or pinned host memory: copied asynchronously). Returns
        (images [B,3,H,W] fp32 CUDA, pad_sizes int32 [B,2] = (pad_x, pad_y) like KITTIDataset.pad_image)."""
        import ctypes
        B = len(imgs)
        dev_imgs, hw = [], []
        for im in imgs:
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise ValueError("GpuInputPipeline.images: uint8 [h, w, 3] tensors expected")
            h, w = int(im.shape[0]), int(im.shape[1])
            if h > self.H or w > self.W:
                raise ValueError("image %dx%d larger than the padded input %dx%d" % (h, w, self.H, self.W))
            if not im.is_cuda:
                im = im.to(self.device, non_blocking=True)
            dev_imgs.append(im.contiguous())
            hw.append((h, w, (self.W - w) // 2, (self.H - h) // 2))
        ptrs = torch.tensor([t.data_ptr() for t in dev_imgs], dtype=torch.int64).to(self.device, non_blocking=True)
        hw_t = torch.tensor(hw, dtype=torch.int32).to(self.device, non_blocking=True)
        fl = None
        if flips is not None:
            fl = torch.tensor([1 if f else 0 for f in flips], dtype=torch.int32).to(self.device, non_blocking=True)
        if out is None:
            out = torch.empty(B, 3, self.H, self.W, dtype=torch.float32, device=self.device)
        call("mf_preprocess_images_u8", ptrs.data_ptr(), hw_t.data_ptr(), fl.data_ptr() if fl is not None else None, B, self.H,
             self.W, ctypes.addressof(self.mean), ctypes.addressof(self.std), self.to_bgr, out.data_ptr(), stream())
        self._keep = (dev_imgs, ptrs, hw_t, fl)          # alive until the kernel has run (stream ordered)
        return out, hw_t[:, 2:4]

    def heatmaps(self, obj6, out=None):
        """obj6: int32 [B, max_objs, 6] = (valid, class id, cx, cy, rx, ry) in feature-map coordinates (CUDA, or host: copied).
        Returns the target heat maps [B, ncls, H/4, W/4] fp32 (`hm` field of the reference's targets, kitti.py:302, 449-463)."""
        if obj6.dtype != torch.int32 or obj6.dim() != 3 or obj6.shape[2] != 6:
            raise ValueError("GpuInputPipeline.heatmaps: int32 [B, max_objs, 6] expected")
        obj6 = obj6.to(self.device, non_blocking=True).contiguous()
        B, M = obj6.shape[0], obj6.shape[1]
        h, w = self.H // self.down, self.W // self.down
        if out is None:
            out = torch.empty(B, self.ncls, h, w, dtype=torch.float32, device=self.device)
        call("mf_draw_heatmaps", obj6.data_ptr(), B, M, self.ncls, h, w, out.data_ptr(), stream())
        self._keep_obj = obj6
        return out
