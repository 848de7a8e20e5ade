"""Training step on the CUDA kernels, driven exactly like the reference loop (engine/trainer.py:103-126):

    loss_dict, log = model(images, targets); losses = sum(loss_dict.values())
    optimizer.zero_grad(); losses.backward(); [clip_grad_norm_]; optimizer.step(); scheduler.step()

`model(images, targets)` in training mode runs the train-mode plans + the fused loss; `losses.backward()` reaches the
backward tape through `model/detector.py::_TapeBridge` and accumulates into `p.grad`. `Trainer` is that loop as an object:
it owns the `FusedAdamW` arena optimiser, performs the data-parallel gradient exchange the reference gets from
DistributedDataParallel (tools/plain_train_net.py:100-104: mean over ranks) - bucketed NCCL all-reduce of the gradient arena
with the 1/world folded into the AdamW kernel, or the fused peer-memory kernel when built with `symmetric_group` - and guards
the update with a device-side finite check of the gradient arena (fp16 gradient flow under a fixed loss scale: a non-finite
gradient skips the step instead of poisoning params and moments; the reference trains in fp32 and stops on NaN losses,
detector_loss.py:485-489).
"""
import torch
import torch.distributed as dist

from . import solver


class Trainer(object):
    def __init__(self, model, cfg, loss_scale=None, symmetric_group=None, process_group=None, bucket_bytes=32 << 20):
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("monoflex_b200 trains on sm_100a GPUs only; no CPU fallback")
        self.model = model.train()
        if loss_scale is not None:
            model.loss_scale = float(loss_scale)
        self.cfg = cfg
        self.group = process_group
        self.bucket_bytes = bucket_bytes
        groups = solver.get_model_params(model, cfg)
        self.optimizer = solver.FusedAdamW(groups, lr=cfg.SOLVER.BASE_LR, weight_decay=cfg.SOLVER.WEIGHT_DECAY, betas=(0.9, 0.99),
                                           symmetric_group=symmetric_group)
        self.scheduler, _ = solver.build_scheduler(self.optimizer, cfg.SOLVER)
        self.grad_norm_clip = getattr(cfg.SOLVER, "GRAD_NORM_CLIP", -1)
        self.iteration = 0
        self._unused_marked = False

    @property
    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def step(self, images, targets):
        """one optimisation step on a batch -> (loss_dict, log_loss_dict) of the forward that produced the gradients"""
        model, opt = self.model, self.optimizer
        model.train()
        loss_dict, log = model(images, targets)
        losses = sum(loss_dict.values())
        opt.zero_grad()
        losses.backward()                                            # fused loss backward -> tape bridge -> p.grad (arena views)
        if not self._unused_marked:                                  # parameters outside the forward graph: the reference's
            opt.mark_unused({n for n, _ in model.named_parameters()} - set(model.last_grad_names), model)   # AdamW skips grad None
            self._unused_marked = True
        if self.grad_norm_clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), self.grad_norm_clip)
        world = self.world
        if opt._symm is not None:
            opt.step_exchange()                                      # reduce + AdamW + broadcast in one peer-memory kernel
        else:
            if world > 1:
                solver.allreduce_grads(opt.arena, self.bucket_bytes, self.group)
            opt.step(grad_scale=1.0 / world)
        self.scheduler.step()
        self.iteration += 1
        return {k: v.detach() for k, v in loss_dict.items()}, log
