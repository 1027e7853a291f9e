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


def test_bias_block_csr_groups_positions_by_bucket():
    """The static CSR the relative-position-table adjoint gathers over: every (i, j) of the block appears exactly once,
    grouped by bucket, with flat offsets into the padded [Tq, row_stride] bias layout."""
    from ifseg_b200 import ops

    g = torch.Generator().manual_seed(0)
    bucket = torch.randint(0, 13, (40, 40), generator=g)
    ids = torch.randperm(40, generator=g)[:9]
    lo, row_stride, num_rel = 5, 64, 13
    order, flat = ops.bias_block_csr(bucket, ids, lo, row_stride)
    counts = torch.bincount(flat, minlength=num_rel)
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    assert order.numel() == 81 and order.unique().numel() == 81
    dbias = torch.randn(lo + 9 + 3, row_stride, generator=g)
    want = torch.zeros(num_rel)
    for a in range(9):
        for b in range(9):
            want[bucket[ids[a], ids[b]]] += dbias[lo + a, lo + b]
    got = torch.stack([dbias.reshape(-1)[order[offsets[r]:offsets[r + 1]].long()].sum() for r in range(num_rel)])
    assert torch.allclose(got, want, atol=1e-5)


def test_synthetic_train_sample_layout():
    """Same layout as SegmentationDataset.collate for artificial_image_type='rand_k-1-33' (segmentation_dataset.py:303-347)."""
    from ifseg_b200.config import preset
    from ifseg_b200.synthetic import synthetic_train_sample

    cfg = preset("segofa_base", num_seg=15, patch_image_size=64, orig_patch_image_size=64)
    s = synthetic_train_sample(cfg, 3, 64, seed=4)
    P = 16
    aux = s["aux_input"]
    assert set(aux) == {"src_tokens", "src_lengths", "patch_images", "patch_masks", "prev_output_tokens"}
    assert aux["patch_masks"].shape == (3 * P,) and s["text2seg_target"].shape == (3, 64 * 64 + 1)
    ends = aux["patch_masks"].view(3, P)
    assert (ends[:, 1:] > ends[:, :-1]).all() and (ends[:, 0] >= 1).all()  # every bag holds at least one token
    for b in range(3):
        n = int(ends[b, -1])
        assert (aux["patch_images"][b, :n] != cfg.padding_idx).all() and (aux["patch_images"][b, n:] == cfg.padding_idx).all()
    t = s["text2seg_target"]
    assert (t[:, -1] == 2).all() and (t[:, :-1] >= 59457).all() and (t[:, :-1] < 59457 + 15).all()
    assert s["net_input"]["patch_images"].shape == (3, 3, 64, 64) and s["target"].shape == (3, 64 * 64 + 1)


def test_training_engine_fails_loudly_without_cuda():
    from ifseg_b200.train_engine import SegOFATrainEngine
    from ifseg_b200.trainer import SegOFATrainer

    m = _model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SegOFATrainEngine(m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SegOFATrainer(m)
