"""Batch containers the hot path touches, API-compatible with the reference's (structures/params_3d.py:5-56,
structures/image_list.py:6-71, data/datasets/kitti_utils.py:211-218). The reference's own objects are accepted too:
only `get_field`, `size`, `to` and the calibration attributes are used."""
import torch


class Calibration(object):
    """Camera intrinsics of one image (kitti_utils.py:211-218): P is the 3x4 projection matrix."""

    def __init__(self, P):
        self.P = [[float(v) for v in row] for row in P]
        self.f_u, self.f_v = self.P[0][0], self.P[1][1]
        self.c_u, self.c_v = self.P[0][2], self.P[1][2]
        self.b_x, self.b_y = self.P[0][3] / (-self.f_u), self.P[1][3] / (-self.f_v)


class ParamsList(object):
    def __init__(self, image_size, is_train=True):
        self.size = image_size
        self.is_train = is_train
        self.extra_fields = {}

    def add_field(self, field, field_data):
        if not isinstance(field_data, torch.Tensor) and not hasattr(field_data, "f_u"):
            field_data = torch.as_tensor(field_data)
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def to(self, device):
        out = ParamsList(self.size, self.is_train)
        for k, v in self.extra_fields.items():
            if torch.is_tensor(v):     # fields in pinned host memory (a pin_memory data loader) go up without blocking the host
                out.extra_fields[k] = v.to(device, non_blocking=v.device.type == "cpu" and v.is_pinned())
            else:
                out.extra_fields[k] = v.to(device) if hasattr(v, "to") else v
        return out

    def pin_memory(self):
        """the same fields in page-locked host memory (what DataLoader(pin_memory=True) does to a batch)"""
        out = ParamsList(self.size, self.is_train)
        for k, v in self.extra_fields.items():
            out.extra_fields[k] = v.pin_memory() if torch.is_tensor(v) and v.device.type == "cpu" else v
        return out

    def __len__(self):
        return int(torch.count_nonzero(self.extra_fields["reg_mask"])) if self.is_train else 0


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)


def to_image_list(tensors, size_divisible=0):
    """4-D tensor -> ImageList view (the only case the detector uses, model/detector.py:30)."""
    if hasattr(tensors, "tensors") and hasattr(tensors, "image_sizes"):
        return tensors
    if isinstance(tensors, torch.Tensor):
        if tensors.dim() == 3:
            tensors = tensors[None]
        assert tensors.dim() == 4
        return ImageList(tensors, [t.shape[-2:] for t in tensors])
    if isinstance(tensors, (tuple, list)):
        shape = tuple(max(s) for s in zip(*[t.shape for t in tensors]))
        out = tensors[0].new_zeros((len(tensors),) + shape)
        for src, dst in zip(tensors, out):
            dst[:src.shape[0], :src.shape[1], :src.shape[2]].copy_(src)
        return ImageList(out, [t.shape[-2:] for t in tensors])
    raise TypeError("Unsupported type for to_image_list: {}".format(type(tensors)))
