"""Detect_Head glue (reference: model/head/detector_head.py:9-29)."""
from torch import nn

from .detector_infer import make_post_processor
from .detector_loss import make_loss_evaluator
from .detector_predictor import make_predictor


class Detect_Head(nn.Module):
    def __init__(self, cfg, in_channels):
        super(Detect_Head, self).__init__()
        self.predictor = make_predictor(cfg, in_channels)
        self.loss_evaluator = make_loss_evaluator(cfg)
        self.post_processor = make_post_processor(cfg)

    def forward(self, features, targets=None, test=False):
        if self.training:
            # detector_head.py:17-21. The predictions leave the hand-written kernels outside autograd: gradients reach the
            # parameters only through KeypointDetector's tape bridge (model/detector.py::_TapeBridge), which is why the
            # training entry point is KeypointDetector.forward and not this sub-module on its own.
            raise RuntimeError("Detect_Head.forward in training mode: call KeypointDetector.forward(images, targets) - the "
                               "backward tape spans backbone and head and is attached there")
        x = self.predictor(features, targets)
        return self.post_processor(x, targets, test=test, features=features)


def bulid_head(cfg, in_channels):   # (sic) reference spelling, model/head/detector_head.py:27
    return Detect_Head(cfg, in_channels)
