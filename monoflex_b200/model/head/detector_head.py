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
            # Loss_Computation itself is built (detector_loss.py); what is missing for a training step is the train-mode
            # forward (batch-statistics BN) and the conv / DCN backward kernels - never silently run the eval kernels here
            raise NotImplementedError("training-mode forward/backward of the backbone and predictor is a later SURVEY §8 "
                                      "row (R5/R13); call heads.loss_evaluator(predictions, targets) directly")
        x = self.predictor(features, targets)
        return self.post_processor(x, targets, test=test, features=features)


def bulid_head(cfg, in_channels):   # (sic) reference spelling, model/head/detector_head.py:27
    return Detect_Head(cfg, in_channels)
