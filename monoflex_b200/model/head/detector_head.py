"""Detect_Head glue (reference: model/head/detector_head.py:9-29)."""
from torch import nn

from .detector_infer import make_post_processor
from .detector_predictor import make_predictor


class _TrainingNotBuilt(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError("Loss_Computation (model/head/detector_loss.py) is a later SURVEY §8 row; the "
                                  "round-1 library builds the inference path only and has no PyTorch fallback")


class Detect_Head(nn.Module):
    def __init__(self, cfg, in_channels):
        super(Detect_Head, self).__init__()
        self.predictor = make_predictor(cfg, in_channels)
        self.loss_evaluator = _TrainingNotBuilt()
        self.post_processor = make_post_processor(cfg)

    def forward(self, features, targets=None, test=False):
        x = self.predictor(features, targets)
        if self.training:
            return self.loss_evaluator(x, targets)
        return self.post_processor(x, targets, test=test, features=features)


def bulid_head(cfg, in_channels):   # (sic) reference spelling, model/head/detector_head.py:27
    return Detect_Head(cfg, in_channels)
