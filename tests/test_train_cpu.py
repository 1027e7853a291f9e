"""CPU: host-side logic of the training path that needs no kernel -- the flat parameter arena (views, adjacency of
fused operands, bucket marks) and the world-size-2 gloo gradient exchange of SegOFATrainer."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401


def _model():
    from ifseg_b200.segofa import SegOFAModel

    return SegOFAModel.from_config("segofa_tiny", 15, 64)


def test_param_arena_views_and_adjacency():
    from ifseg_b200.train_engine import ParamArena

    m = _model()
    a = m.encoder.layers[0].self_attn
    groups = [[a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias],
              [a.c_attn], [m.encoder.layer_norm.weight]]
    before = [p.detach().clone() for g in groups for p in g]
    ar = ParamArena(groups, torch.device("cpu"))
    after = [p for g in groups for p in g]
    for b, p in zip(before, after):
        assert torch.equal(b, p.detach())  # values preserved
        assert p.grad is not None and p.grad.shape == p.shape
    D = a.q_proj.weight.shape[0]
    fused = ar.view(ar.flat32, groups[0], (3 * D, D))
    assert torch.equal(fused[D:2 * D], a.k_proj.weight.detach())
    fused.zero_()  # the parameters are views of the arena
    assert a.v_proj.weight.abs().sum() == 0
    a.q_proj.weight.grad.fill_(2.0)
    assert ar.grad32[: D * D].eq(2.0).all()
    assert all(o % 8 == 0 for o, _ in (ar.slots[id(g[0])] for g in groups))  # 32-byte aligned group starts
    with pytest.raises(ValueError):
        ar.view(ar.flat32, [a.q_proj.weight, a.v_proj.weight], (2 * D, D))  # not adjacent


def test_arena_layout_is_forward_ordered_and_complete():
    """_build_arena without a GPU: marks are increasing in forward order, every trainable non-bias-path parameter
    is covered exactly once, fused groups are adjacent."""
    from ifseg_b200.train_engine import SegOFATrainEngine

    m = _model()
    eng = SegOFATrainEngine.__new__(SegOFATrainEngine)
    eng.model, eng.cfg, eng.device = m, m.cfg, torch.device("cpu")
    eng._build_arena()
    nl = len(m.encoder.layers)
    keys = [("enc", i) for i in range(nl + 1)] + ["cross_kv"] + [("dec", i) for i in range(len(m.decoder.layers) + 1)]
    offs = [eng.marks[k] for k in keys]
    assert offs == sorted(offs) and offs[0] > 0 and offs[-1] <= eng.arena.numel
    covered = {id(p) for p in eng.arena.params}
    for name, p in m.named_parameters():
        if p.requires_grad and not eng._is_bias_path(name):
            assert id(p) in covered, name
        else:
            assert id(p) not in covered, name
    d0 = m.decoder.layers[0].encoder_attn
    d1 = m.decoder.layers[1].encoder_attn
    ws = [d0.k_proj.weight, d0.v_proj.weight, d1.k_proj.weight, d1.v_proj.weight]
    D = ws[0].shape[0]
    eng.arena.view(eng.arena.flat32, ws, (4 * D, D))  # adjacency of the all-layer cross k/v operand
    # bucket hand-off: gradients become final from the top of the arena downwards
    got = []
    eng.grad_sync = lambda lo, hi: got.append((lo, hi))
    eng._sync_hi = eng.arena.numel
    for k in reversed(keys):
        eng._sync_down(k)
    eng._sync_down(None)
    assert got[0][1] == eng.arena.numel and got[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(got, got[1:]))  # contiguous, non-overlapping, covers everything


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ifseg_b200.trainer import SegOFATrainer

    class FakeArena:
        def __init__(self):
            self.grad32 = torch.full((1000,), float(rank + 1))

    class FakeEngine:
        arena = FakeArena()

    tr = SegOFATrainer.__new__(SegOFATrainer)
    tr.engine, tr.pg, tr._works = FakeEngine(), None, []
    for lo, hi in ((600, 1000), (200, 600), (0, 200)):  # buckets handed over top-down, as the backward does
        tr._sync_bucket(lo, hi)
    for w in tr._works:
        w.wait()
    q.put((rank, tr.engine.arena.grad32.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert torch.equal(res[r], torch.full((1000,), 3.0))  # sum over ranks; the optimizer applies 1/world
