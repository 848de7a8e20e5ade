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

`use_cuda_graph=True`: shapes are static (384x1280 padded images, 40 object slots), train-mode plans read
every parameter from its live storage, and nothing between the image copy and the optimiser update needs the host, so after
`graph_warmup` eager steps the whole step - forward plans, fused loss, loss backward, the backward tape with all its
temporaries, zero_grad and (single GPU) the guarded AdamW update - is captured ONCE in a CUDA graph and replayed; per step
the host copies the batch and its labels into static buffers and launches one graph (multi-GPU: graph, NCCL all-reduce,
AdamW).
"""
import torch
import torch.distributed as dist

from . import solver


class Trainer(object):
    def __init__(self, model, cfg, loss_scale=None, symmetric_group=None, process_group=None, bucket_bytes=32 << 20,
                 use_cuda_graph=False, graph_warmup=2):
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
        self.use_cuda_graph = bool(use_cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self._graph = None
        if self.use_cuda_graph and self.grad_norm_clip > 0:
            raise NotImplementedError("GRAD_NORM_CLIP > 0 reads the norm on the host; not available in the captured step")

    @property
    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    # ------------------------------------------------------------------ eager step (the reference loop, line by line)
    def step(self, images, targets, sync_log=True):
        """one optimisation step on a batch -> (loss_dict, log_loss_dict) of the forward that produced the gradients"""
        if self.use_cuda_graph and self.iteration >= self.graph_warmup:
            return self._step_graph(images, targets, sync_log)
        model, opt = self.model, self.optimizer
        model.train()
        loss_dict, log = model._forward_train(images, targets, sync_log=sync_log)
        losses = sum(loss_dict.values())
        opt.zero_grad()
        losses.backward()                                            # fused loss backward -> tape bridge -> p.grad (arena views)
        if not self._unused_marked:                                  # parameters outside the forward graph: the reference's
            opt.mark_unused({n for n, _ in model.named_parameters()} - set(model.last_grad_names), model)   # AdamW skips grad None
            self._unused_marked = True
        if self.grad_norm_clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), self.grad_norm_clip)
        self._exchange_and_update()
        self.scheduler.step()
        self.iteration += 1
        return {k: v.detach() for k, v in loss_dict.items()}, log

    def _exchange_and_update(self):
        opt, world = self.optimizer, self.world
        if opt._symm is not None:
            opt.step_exchange()                                      # reduce + AdamW + broadcast in one peer-memory kernel
        else:
            if world > 1:
                solver.allreduce_grads(opt.arena, self.bucket_bytes, self.group)
            opt.step(grad_scale=1.0 / world)

    # ------------------------------------------------------------------ captured step
    def _capture(self, images, targets):
        model, opt = self.model, self.optimizer
        dev = images.device
        pred, lossev = model.heads.predictor, model.heads.loss_evaluator
        g = {"x": images.clone(), "prep": tuple(t.clone() for t in lossev.prepare_targets(targets, dev)),
             "update_inside": self.world == 1 and opt._symm is None}
        opt._lr_table()
        pred._targets_preloaded = True
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(graph):
                loss_dict, log = model._forward_train(g["x"], targets, prepared=g["prep"], sync_log=False)
                losses = sum(loss_dict.values())
                opt.zero_grad()
                losses.backward()
                if g["update_inside"]:
                    opt.step(grad_scale=1.0)
        finally:
            pred._targets_preloaded = False
        if g["update_inside"]:                 # capture only records: undo the host-side bookkeeping of the captured step()
            opt.step_count -= 1
            opt._step_t -= 1
        g.update(graph=graph, loss_dict={k: v.detach() for k, v in loss_dict.items()}, log=log, plan_h=pred.last_plan)
        self._graph = g
        return g

    def _step_graph(self, images, targets, sync_log=True):
        model, opt = self.model, self.optimizer
        g = self._graph
        if g is None:
            g = self._capture(images, targets)
        if tuple(images.shape) != tuple(g["x"].shape):
            raise RuntimeError("captured training step: batch shape changed from %s to %s" % (tuple(g["x"].shape), tuple(images.shape)))
        g["x"].copy_(images, non_blocking=True)
        for dst, src in zip(g["prep"], model.heads.loss_evaluator.prepare_targets(targets, images.device)):
            dst.copy_(src, non_blocking=True)
        model.heads.predictor.load_targets(g["plan_h"], targets)
        opt._lr_table()                                              # in-place refresh when the scheduler moved the lr
        g["graph"].replay()
        if g["update_inside"]:
            opt.step_count += 1
            opt._step_t += 1
            for p in opt.arena.tensors:
                torch.autograd.graph.increment_version(p)
        else:
            self._exchange_and_update()
        self.scheduler.step()
        self.iteration += 1
        return g["loss_dict"], (g["log"].resolve() if sync_log else g["log"].snapshot())
