"""First end-to-end training step on the CUDA kernels (reference loop: engine/trainer.py:103-126):

    forward (train-mode plans) -> Loss_Computation -> head_backward + tape (all parameter gradients) -> gradient arena ->
    FusedAdamW (one launch; `step_exchange()` instead when the optimiser was built with a symmetric group).

STATUS: written at the end of round 1 and NOT yet executed on hardware (the GPU budget was exhausted; every piece it calls
is hardware-verified on its own: train-mode forward, loss, head + backbone backward tape, FusedAdamW) - its test
(tests/test_gpu_train.py::test_end_to_end_train_steps) therefore only runs with MF_RUN_UNVERIFIED=1.

This wiring is eager Python over freshly rebuilt plans (the optimiser step changes the weights, the cached plans key on the
parameter versions, so every step re-packs the weights and re-allocates the activation buffers): it establishes CORRECTNESS
of the whole step, not its speed - see DESIGN.md "Training tape" for the static-plan / CUDA-graph version it is a stepping
stone to. Parameters the forward never uses (the outer `project` of the two-level trees, dla_dcn.py:249) get no gradient in
the reference (grad None -> skipped by AdamW); here their learning rate is set to 0 after the first backward so that the
kernel skips them too.
"""
import torch

from . import solver
from .head_backward import predictor_backward
from .tape import backbone_backward


class Trainer(object):
    def __init__(self, model, cfg, loss_scale=128.0, symmetric_group=None):
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("monoflex_b200 trains on sm_100a GPUs only; no CPU fallback")
        self.model = model.train()
        self.loss_scale = float(loss_scale)
        groups = solver.get_model_params(model, cfg)
        self.optimizer = solver.FusedAdamW(groups, lr=cfg.SOLVER.BASE_LR, weight_decay=cfg.SOLVER.WEIGHT_DECAY, betas=(0.9, 0.99),
                                           symmetric_group=symmetric_group)
        self.params = dict(model.named_parameters())
        self._unused_frozen = False
        self.last_grad_names = None

    def step(self, images, targets):
        """one optimisation step on a batch -> (loss_dict, log_loss_dict) of the forward that produced the gradients"""
        model, S = self.model, self.loss_scale
        model.train()
        feats = model.backbone.train_forward(images)
        pred_mod = model.heads.predictor
        pred = pred_mod.train_forward(feats, targets)
        c = pred["cls"].detach().clone().requires_grad_(True)
        r = pred["reg"].detach().clone().requires_grad_(True)
        loss_dict, log = model.heads.loss_evaluator({"cls": c, "reg": r}, targets)
        (S * sum(loss_dict.values())).backward()                              # fused loss backward -> d cls, d reg
        hgrads, d_feat = predictor_backward(pred_mod, pred_mod.last_plan, c.grad, r.grad)
        bgrads = backbone_backward(model.backbone, model.backbone.last_plan, d_feat, stem_wgrad=True)
        self.optimizer.zero_grad()
        got = set()
        with torch.no_grad():
            for prefix, grads in (("heads.predictor.", hgrads), ("backbone.", bgrads)):
                for name, g in grads.items():
                    if g is None:
                        raise RuntimeError("no gradient was produced for %s%s" % (prefix, name))
                    self.params[prefix + name].grad.copy_(g)                  # .grad is a view of the optimiser's arena
                    got.add(prefix + name)
        self.last_grad_names = got
        if not self._unused_frozen:                                           # parameters outside the forward graph: lr 0
            for group in self.optimizer.param_groups:
                p = group["params"][0]
                if not any(p is self.params[n] for n in got):
                    group["lr"] = 0.0
                    group["initial_lr"] = 0.0
            self._unused_frozen = True
        if self.optimizer._symm is not None:
            # the fused exchange kernel scales the reduced gradient by 1/world only: remove the loss scale first (one launch)
            self.optimizer.arena.grads.mul_(1.0 / S)
            self.optimizer.step_exchange()
        else:
            self.optimizer.step(grad_scale=1.0 / S)
        return {k: v.detach() for k, v in loss_dict.items()}, log
