"""TEST INFRASTRUCTURE ONLY (oracle) - CPU restatement of the reference's per-image input path:
pad_image (data/datasets/kitti.py:218-228), ToTensor + Normalize (data/transforms/transforms.py:14-30; torchvision
to_tensor = uint8 HWC -> float32 CHW / 255, normalize = (x - mean) / std in fp32) and the heat-map drawing
(model/heatmap_coder.py:37-64 gaussian_radius / gaussian2D, :83-124 draw_umich_gaussian / draw_umich_gaussian_2D,
:126-134 ellip_gaussian2D). Pinned against the unmodified reference functions by oracle/make_golden_input.py ->
tests/golden/input_pipeline.npz (tests/test_oracle_golden.py). Only tests import this file."""
import numpy as np
import torch


def pad_image(img, H, W):
    """kitti.py:218-228: centred zero padding of a uint8 HWC image -> (padded uint8 [H,W,3], pad_size (x, y))."""
    h, w, c = img.shape
    ret = np.zeros((H, W, c))
    pad_y, pad_x = (H - h) // 2, (W - w) // 2
    ret[pad_y:pad_y + h, pad_x:pad_x + w] = img
    return ret.astype(np.uint8), np.array([pad_x, pad_y])


def to_tensor_normalize(img_u8, mean, std, to_bgr=False):
    """transforms.py:14-30 on a uint8 HWC array -> float32 [3,H,W] torch tensor."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)
    x = x.sub(m).div(s)
    return x[[2, 1, 0]] if to_bgr else x


def gaussian2D(shape, sigma=1):
    """heatmap_coder.py:56-64"""
    m, n = [(ss - 1.) / 2. for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def ellip_gaussian2D(shape, sigma_x, sigma_y):
    """heatmap_coder.py:126-134"""
    m, n = [(ss - 1.) / 2. for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x) / (2 * sigma_x * sigma_x) - (y * y) / (2 * sigma_y * sigma_y))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_umich_gaussian(heatmap, center, radius, k=1):
    """heatmap_coder.py:83-105 (ignore=False branch)"""
    diameter = 2 * radius + 1
    gaussian = gaussian2D((diameter, diameter), sigma=diameter / 6)
    x, y = int(center[0]), int(center[1])
    height, width = heatmap.shape[0:2]
    left, right = min(x, radius), min(width - x, radius + 1)
    top, bottom = min(y, radius), min(height - y, radius + 1)
    masked_heatmap = heatmap[y - top:y + bottom, x - left:x + right]
    masked_gaussian = gaussian[radius - top:radius + bottom, radius - left:radius + right]
    if min(masked_gaussian.shape) > 0 and min(masked_heatmap.shape) > 0:
        np.maximum(masked_heatmap, masked_gaussian * k, out=masked_heatmap)
    return heatmap


def draw_umich_gaussian_2D(heatmap, center, radius_x, radius_y, k=1):
    """heatmap_coder.py:107-124"""
    diameter_x, diameter_y = 2 * radius_x + 1, 2 * radius_y + 1
    gaussian = ellip_gaussian2D((diameter_y, diameter_x), sigma_x=diameter_x / 6, sigma_y=diameter_y / 6)
    x, y = int(center[0]), int(center[1])
    height, width = heatmap.shape[0:2]
    left, right = min(x, radius_x), min(width - x, radius_x + 1)
    top, bottom = min(y, radius_y), min(height - y, radius_y + 1)
    masked_heatmap = heatmap[y - top:y + bottom, x - left:x + right]
    masked_gaussian = gaussian[radius_y - top:radius_y + bottom, radius_x - left:radius_x + right]
    if min(masked_gaussian.shape) > 0 and min(masked_heatmap.shape) > 0:
        np.maximum(masked_heatmap, masked_gaussian * k, out=masked_heatmap)
    return heatmap


def draw_heatmaps(obj6, ncls, h, w):
    """obj6 int [B, max_objs, 6] = (valid, cls, cx, cy, rx, ry) -> float32 [B, ncls, h, w], objects drawn in slot order like
    the reference's per-object loop (kitti.py:449-463: rx == ry -> circular, else one-sided)."""
    B = obj6.shape[0]
    hm = np.zeros((B, ncls, h, w), dtype=np.float32)
    for b in range(B):
        for valid, cls, cx, cy, rx, ry in obj6[b]:
            if not valid:
                continue
            if rx == ry:
                draw_umich_gaussian(hm[b, cls], (cx, cy), int(rx))
            else:
                draw_umich_gaussian_2D(hm[b, cls], (cx, cy), int(rx), int(ry))
    return hm


def synthetic_case(seed=0, B=3, H=384, W=1280, max_objs=40, ncls=3):
    """deterministic test inputs: uint8 images of KITTI-like sizes and an object table with inside (circular) and edge
    (one-sided) objects, some touching the borders, duplicates on one pixel and an empty image."""
    g = np.random.Generator(np.random.PCG64(seed))
    sizes = [(375, 1242), (370, 1224), (384, 1280), (376, 1241), (374, 1238)]
    imgs = [g.integers(0, 256, size=(sizes[i % len(sizes)][0], sizes[i % len(sizes)][1], 3), dtype=np.uint8) for i in range(B)]
    h, w = H // 4, W // 4
    obj = np.zeros((B, max_objs, 6), dtype=np.int32)
    for b in range(B):
        n = 0 if b == 1 else int(g.integers(5, 20))
        for j in range(n):
            cx, cy = int(g.integers(0, w)), int(g.integers(0, h))
            if g.uniform() < 0.3:
                cx = int(g.choice([0, 1, w - 1, w - 2, cx]))
            if g.uniform() < 0.25:                          # edge object: one radius 0
                r = int(g.integers(0, 14))
                rx, ry = (0, r) if g.uniform() < 0.5 else (r, 0)
            else:
                rx = ry = int(g.integers(0, 24))
            obj[b, j] = (1, int(g.integers(0, ncls)), cx, cy, rx, ry)
        if n > 2:
            obj[b, 1, 2:4] = obj[b, 0, 2:4]                  # two objects on the same centre pixel
            obj[b, 1, 1] = obj[b, 0, 1]
    return imgs, obj
