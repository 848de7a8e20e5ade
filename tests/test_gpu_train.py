"""-m gpu: training-side kernels (SURVEY §8 rows R11 backward, R13) against torch-CPU fp32 references of the same op."""
import numpy as np
import pytest
import torch

from monoflex_b200 import solver
from monoflex_b200.config import default_cfg
from oracle import monoflex_oracle as mo

pytestmark = pytest.mark.gpu


def _tensors(gen, shapes):
    return [torch.from_numpy(gen.standard_normal(s).astype(np.float32) * 0.1) for s in shapes]


def test_fused_adamw_matches_torch_adamw():
    """One launch over the arena == torch.optim.AdamW with the reference's groups (solver/__init__.py:10-37), 5 steps,
    ragged tensor sizes (1, chunk-1, chunk, chunk+1, ...), lr decay in between. Tolerance: fp32 re-association only."""
    gen = np.random.Generator(np.random.PCG64(3))
    shapes = [(1,), (511,), (512,), (513,), (16, 3, 7, 7), (27, 64, 3, 3), (27,), (256, 64, 3, 3), (3, 256, 1, 1), (3,)]
    names = ["w%d" % i if i % 2 == 0 else "b%d.bias" % i for i in range(len(shapes))]
    init = _tensors(gen, shapes)
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    lrs = [3e-4 * (2.0 if "bias" in n else 1.0) for n in names]
    ref = torch.optim.AdamW([{"params": [p], "lr": lr} for p, lr in zip(ref_p, lrs)], lr=3e-4, weight_decay=1e-5,
                            betas=(0.9, 0.99))
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ours = solver.FusedAdamW([{"params": [p], "lr": lr} for p, lr in zip(our_p, lrs)], lr=3e-4, weight_decay=1e-5,
                             betas=(0.9, 0.99))
    for it in range(5):
        grads = _tensors(gen, shapes)
        for p, q, g in zip(ref_p, our_p, grads):
            p.grad = g.clone()
            q.grad.copy_(g.cuda())                                    # .grad is a view of the gradient arena
        if it == 3:                                                    # scheduler-style lr change (LambdaLR rewrites group lr)
            for opt in (ref, ours):
                for g in opt.param_groups:
                    g["lr"] *= 0.1
        ref.step()
        ours.step()
        for p, q in zip(ref_p, our_p):
            d = (q.detach().cpu() - p.detach()).abs().max().item()
            assert d <= 2e-7 + 1e-6 * p.detach().abs().max().item(), (it, tuple(p.shape), d)
    for p, q in zip(ref_p, our_p):
        st = ref.state[p]
        assert torch.allclose(ours.state[q]["exp_avg"].cpu(), st["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(ours.state[q]["exp_avg_sq"].cpu(), st["exp_avg_sq"], rtol=1e-5, atol=1e-10)
        assert float(ours.state[q]["step"]) == float(st["step"]) == 5.0
    # padding between tensors is never touched
    a = ours.arena
    used = torch.zeros(a.numel, dtype=torch.bool)
    for o, p in zip(a.offsets, a.tensors):
        used[o:o + p.numel()] = True
    assert a.params.cpu()[~used].abs().sum() == 0 and ours.exp_avg_sq.cpu()[~used].abs().sum() == 0


def test_fused_adamw_grad_scale_is_ddp_mean():
    """grad_scale = 1/world on a SUM-reduced arena == AdamW on the mean gradient (DDP semantics)."""
    gen = np.random.Generator(np.random.PCG64(4))
    shapes = [(64, 64, 3, 3), (64,)]
    init, g0, g1 = _tensors(gen, shapes), _tensors(gen, shapes), _tensors(gen, shapes)
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref = torch.optim.AdamW(ref_p, lr=3e-4, weight_decay=1e-5, betas=(0.9, 0.99))
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ours = solver.FusedAdamW(our_p, lr=3e-4, weight_decay=1e-5, betas=(0.9, 0.99))
    for p, q, a, b in zip(ref_p, our_p, g0, g1):
        p.grad = (a + b) / 2
        q.grad.copy_((a + b).cuda())
    ref.step()
    ours.step(grad_scale=0.5)
    for p, q in zip(ref_p, our_p):
        assert torch.allclose(q.detach().cpu(), p.detach(), rtol=1e-6, atol=2e-7)


def test_detector_optimizer_layout_and_version_bump():
    """build_optimizer on the real detector: the reference's group census (280 tensors, 112 'bias' names at
    2x lr, 20.95 M parameters), one launch updates all of them and invalidates the cached kernel plans."""
    from monoflex_b200 import engine
    from monoflex_b200.model.detector import KeypointDetector
    cfg = default_cfg()
    model = KeypointDetector(cfg).cuda()
    opt = solver.build_optimizer(model, cfg)
    n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_bias = sum(1 for g in opt.param_groups if g["lr"] == pytest.approx(6e-4))
    names = [k for k, v in model.named_parameters() if v.requires_grad]
    assert len(opt.param_groups) == len(names) and n_bias == sum("bias" in k for k in names)
    assert opt.arena.numel >= n_par and opt.arena.numel - n_par < len(names) * opt.arena.chunk
    before = opt.arena.params.clone()
    fp0 = engine.fingerprint(model)
    opt.arena.grads.normal_()
    opt.step()
    torch.cuda.synchronize()
    assert engine.fingerprint(model) != fp0
    delta = (opt.arena.params - before).abs()
    # first AdamW step moves every weight by ~lr (m/sqrt(v) = +-1): between 0.5 lr and 2.1 lr where the gradient is non-zero
    used = opt.arena.chunk_table([1.0] * len(names)).cuda().repeat_interleave(opt.arena.chunk) > 0
    assert delta[used].max().item() < 2.1 * 6e-4 and delta.max().item() > 0.5 * 3e-4


def test_focal_loss_backward_matches_autograd():
    from monoflex_b200._lib import call, stream
    gen = np.random.Generator(np.random.PCG64(12))
    pred = torch.from_numpy(gen.uniform(1e-4, 1 - 1e-4, (2, 3, 24, 80)).astype(np.float32)).requires_grad_(True)
    tgt = torch.from_numpy((gen.uniform(0, 1, (2, 3, 24, 80)) ** 8).astype(np.float32))
    tgt.view(-1)[::97] = 1.0
    tgt.view(-1)[5::193] = -1.0                                        # ignored pixels
    loss, npos = mo.focal_loss(pred, tgt)
    (loss / torch.clamp(npos, 1)).backward()                           # detector_loss.py:276 with hm_loss weight 1
    p, t = pred.detach().cuda(), tgt.cuda()
    scale = (1.0 / torch.clamp(npos, 1)).reshape(1).cuda()
    grad = torch.empty_like(p)
    call("mf_focal_loss_backward", p.data_ptr(), t.data_ptr(), p.numel(), scale.data_ptr(), grad.data_ptr(), stream())
    ref = pred.grad
    assert (grad.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert grad.cpu()[tgt == -1].abs().sum() == 0
