"""KeypointDetector (reference: model/detector.py:11-37): backbone -> heads, same constructor and forward signature so
engine/trainer.py:109 and engine/inference.py:38 call it unchanged."""
from torch import nn

from ..structures import to_image_list
from .backbone import build_backbone
from .head.detector_head import bulid_head


class KeypointDetector(nn.Module):
    def __init__(self, cfg):
        super(KeypointDetector, self).__init__()
        self.backbone = build_backbone(cfg)
        self.heads = bulid_head(cfg, self.backbone.out_channels)
        self.test = cfg.DATASETS.TEST_SPLIT == 'test'

    def forward(self, images, targets=None):
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        images = to_image_list(images)
        features = self.backbone(images.tensors)
        if self.training:
            return self.heads(features, targets)
        return self.heads(features, targets, test=self.test)
