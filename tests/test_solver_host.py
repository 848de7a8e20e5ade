"""CPU: host logic of the optimiser row (R13): arena layout, the reference's one-group-per-tensor lr rule
(solver/__init__.py:10-24), torch.optim plumbing, and the bucketed gradient all-reduce on gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from monoflex_b200 import solver
from monoflex_b200.config import default_cfg


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Flatten(),
                               torch.nn.Linear(5 * 7 * 7, 3))


def test_param_groups_follow_reference_rule():
    m, cfg = _toy(), default_cfg()
    groups = solver.get_model_params(m, cfg)
    names = [k for k, _ in m.named_parameters()]
    assert len(groups) == len(names)
    for k, g in zip(names, groups):
        assert len(g["params"]) == 1
        assert g["lr"] == (cfg.SOLVER.BASE_LR * cfg.SOLVER.BIAS_LR_FACTOR if "bias" in k else cfg.SOLVER.BASE_LR)


def test_arena_aliases_params_and_grads():
    m, cfg = _toy(), default_cfg()
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    opt = solver.build_optimizer(m, cfg)
    a = opt.arena
    assert a.numel % a.chunk == 0 and all(o % a.chunk == 0 for o in a.offsets)
    for (k, p), o in zip(m.named_parameters(), a.offsets):
        assert torch.equal(p.detach(), before[k])                       # values preserved
        assert p.data_ptr() == a.params.data_ptr() + 4 * o              # p.data is a view of the arena
        assert p.grad.data_ptr() == a.grads.data_ptr() + 4 * o
    m(torch.randn(2, 3, 9, 9)).sum().backward()                          # autograd accumulates INTO the arena
    assert a.grads.abs().sum() > 0
    assert torch.equal(a.view(a.grads, 0), m[0].weight.grad)
    opt.zero_grad()
    assert a.grads.abs().sum() == 0 and m[0].weight.grad is not None
    table = a.chunk_table([g["lr"] for g in opt.param_groups])
    assert table.numel() == a.n_chunks
    assert all(min(abs(v - r) for r in (0.0, 3e-4, 6e-4)) < 1e-9 for v in table.tolist())
    spans = a.buckets(4096)
    assert sorted(spans)[0][0] == 0 and sorted(spans)[-1][1] == a.numel
    assert all(lo % a.chunk == 0 for lo, _ in spans) and sum(hi - lo for lo, hi in spans) == a.numel


def test_state_dict_roundtrip_and_scheduler():
    m, cfg = _toy(), default_cfg()
    opt = solver.build_optimizer(m, cfg)
    opt.exp_avg.uniform_(-1, 1)
    opt.exp_avg_sq.uniform_(0, 1)
    opt.step_count = 7
    opt._bind_state()
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == len(list(m.parameters())) and sd["state"][0]["step"] == 7
    m2 = _toy()
    opt2 = solver.build_optimizer(m2, cfg)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 7
    for i in range(len(opt.arena.tensors)):
        assert torch.equal(opt2.arena.view(opt2.exp_avg, i), opt.arena.view(opt.exp_avg, i))
        assert torch.equal(opt2.arena.view(opt2.exp_avg_sq, i), opt.arena.view(opt.exp_avg_sq, i))
    sched, warm = solver.build_scheduler(opt, cfg.SOLVER)
    assert warm is None
    lr0 = [g["lr"] for g in opt.param_groups]
    sched.step(20000)                                                    # first decay step of config/defaults.py:265
    assert [g["lr"] for g in opt.param_groups] == pytest.approx([l * 0.1 for l in lr0])


def test_step_without_cuda_raises():
    m, cfg = _toy(), default_cfg()
    opt = solver.build_optimizer(m, cfg)
    with pytest.raises(RuntimeError):
        opt.step()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from monoflex_b200 import parallel
    parallel.init("gloo")
    m, cfg = _toy(), default_cfg()
    opt = solver.build_optimizer(m, cfg)
    torch.manual_seed(100 + rank)
    m(torch.randn(2, 3, 9, 9)).sum().backward()
    local = opt.arena.grads.clone()
    w, handles = solver.allreduce_grads(opt.arena, bucket_bytes=4096)
    assert w == world and handles == []
    torch.save({"local": local, "reduced": opt.arena.grads.clone()}, os.path.join(out, "g%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_bucketed_allreduce_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "g%d.pt" % r)) for r in range(world)]
    total = res[0]["local"] + res[1]["local"]
    for r in res:
        assert torch.allclose(r["reduced"], total, rtol=0, atol=0) or torch.allclose(r["reduced"], total, rtol=1e-6)
    assert not torch.equal(res[0]["local"], res[1]["local"])
