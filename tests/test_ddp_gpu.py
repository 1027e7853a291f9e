"""-m gpu: the drop-in boundary under the SHIPPED launch line (run_scripts/IFSeg/coco_unseen.sh:73-137): default
`--ddp-backend pytorch_ddp` = torch.nn.parallel.DistributedDataParallel with --find-unused-parameters
(custom_fairseq/fairseq/models/distributed_fairseq_model.py:57-83) and `--fp16` = model.half() with fp32 masters in
FP16Optimizer (trainer.py:95-101).  World size 2 on ONE device (gloo moves CUDA tensors through the host; NCCL refuses two
ranks per GPU) -- the reducer logic, not the transport, is what is under test."""
import os
import socket
import types

import pytest
import torch

from helpers import build_cuda_model, load_prompts, make_task, rel_l2

pytestmark = pytest.mark.gpu


class _Proxy(torch.nn.Module):
    """what fairseq's ModuleProxyWrapper does (distributed/module_proxy_wrapper.py): attribute fall-through to the
    twice-wrapped module"""

    def __init__(self, ddp):
        super().__init__()
        self.module = ddp

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            try:
                return getattr(self.module, name)
            except AttributeError:
                return getattr(self.module.module, name)

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _sample(model, B, S, C, seed):
    from ifseg_b200 import ops
    from ifseg_b200.synthetic import synthetic_inputs

    prompt = torch.tensor(load_prompts()["15"], dtype=torch.long)
    names = torch.full((C, 4), 1, dtype=torch.long)
    names[:, 0] = torch.arange(C) + 100
    lens = torch.ones(C, dtype=torch.int32)
    bag, ends, t2s = ops.artificial_sample(names.cuda(), lens.cuda(), B, S // 16, S, seed=seed)
    aux = dict(src_tokens=prompt.unsqueeze(0).repeat(B, 1).cuda(), patch_images=bag, patch_masks=ends,
               prev_output_tokens=torch.zeros(B, 1, dtype=torch.long).cuda())
    inp = {k: v.cuda() for k, v in synthetic_inputs(model.cfg, B, S, seed=seed, src_tokens=load_prompts()["15"]).items()}
    gen = torch.Generator().manual_seed(seed)
    target = torch.cat([torch.randint(0, C + 1, (B, S * S), generator=gen) + 59457, torch.full((B, 1), 2)], 1)
    return {"net_input": inp, "aux_input": aux, "target": target, "text2seg_target": t2s, "ntokens": 1, "nsentences": B}


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel

    from ifseg_b200.seg_criterion import SegCriterion

    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        C, S, B = 15, 64, 2
        model, _ = build_cuda_model("segofa_base", C, S, seed=0)
        if rank == 1:  # DDP must broadcast rank 0's parameters at construction
            with torch.no_grad():
                model.encoder.layers[0].fc1.weight.add_(1.0)
        model.train()
        ddp = _Proxy(DistributedDataParallel(model, device_ids=[0], output_device=0, find_unused_parameters=True,
                                             broadcast_buffers=False))
        crit = SegCriterion(make_task(C), init_seg_with_text="false")
        sample = _sample(model, B, S, C, seed=100 + rank)  # every rank its own data
        te = model.train_engine()
        te.stochastic = False  # deterministic: local and DDP passes must see the same function
        # local gradients of this rank's sample, no DDP involved
        loss_l, _, _ = crit(model, sample)
        loss_l.backward()
        names = ["encoder.layers.0.fc1.weight", "decoder.layers.5.encoder_attn.q_proj.bias", "encoder.layer_norm.weight",
                 "decoder.seg_rel_pos_table_list.3.weight", "encoder.embed_positions.weight", "decoder.layers.0.self_attn.c_attn"]
        prm = dict(model.named_parameters())
        local = {n: prm[n].grad.detach().clone() for n in names}
        model.zero_grad(set_to_none=True)
        # through DDP (two forwards per backward: the image-free one with grad, the real-image one without)
        loss, sample_size, log = crit(ddp, sample)
        loss.backward()
        torch.cuda.synchronize()
        errs = {}
        for n in names:
            both = [torch.empty_like(local[n]) for _ in range(world)]
            dist.all_gather(both, local[n])
            mean = sum(both) / world
            errs[n] = rel_l2(prm[n].grad, mean)
            assert prm[n].grad.data_ptr() >= te.arena.grad32.data_ptr(), "reduced gradient left the flat arena"
            assert prm[n].grad.data_ptr() < te.arena.grad32.data_ptr() + te.arena.grad32.numel() * 4
        unused = prm["decoder.embed_positions.weight"].grad  # one of the 20 never-reached tensors
        # gradient accumulation (--update-freq 2): a second backward adds, and the hooks fire again
        g1 = prm[names[0]].grad.detach().clone()
        loss2, _, _ = crit(ddp, sample)
        loss2.backward()
        acc_err = rel_l2(prm[names[0]].grad, 2 * g1)
        if rank == 0:
            torch.save(dict(errs=errs, acc_err=acc_err, unused_is_none_or_zero=unused is None or float(unused.abs().max()) == 0.0,
                            loss=float(loss), sample_size=sample_size), out)
    finally:
        dist.destroy_process_group()


def test_pytorch_ddp_find_unused_reduces_arena_gradients(cuda_device, tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    print(r)
    assert all(e < 2e-3 for e in r["errs"].values()), r["errs"]  # DDP's .grad == mean over ranks of the local gradients
    assert r["acc_err"] < 2e-3 and r["unused_is_none_or_zero"]


def test_half_precision_model_as_trainer_fp16_makes_it(cuda_device):
    """`--fp16`: trainer.py:95-101 calls model.half(); FP16Optimizer scales the loss and expects fp16 .grad on the fp16
    parameters.  The engines keep fp32 masters / bf16 operands internally."""
    from ifseg_b200.seg_criterion import SegCriterion

    C, S, B = 15, 64, 2
    ref_model, _ = build_cuda_model("segofa_base", C, S, seed=0)
    ref_model.train()
    ref_model.train_engine().stochastic = False
    crit = SegCriterion(make_task(C), init_seg_with_text="false")
    sample = _sample(ref_model, B, S, C, seed=3)
    loss32, _, _ = crit(ref_model, sample)
    loss32.backward()
    g32 = {n: p.grad.detach().float().clone() for n, p in ref_model.named_parameters() if p.grad is not None}
    for dtype in (torch.float16, torch.bfloat16):
        model, _ = build_cuda_model("segofa_base", C, S, seed=0)
        model = model.to(dtype).train()  # what trainer.py does for --fp16 / --bf16
        model.train_engine().stochastic = False
        crit = SegCriterion(make_task(C), init_seg_with_text="false")
        loss, _, log = crit(model, sample)
        assert abs(loss.item() - loss32.item()) < 2e-2 * abs(loss32.item())
        (loss * 128.0).backward()  # FP16Optimizer.backward: loss scaling (fp16_optimizer.py:96-106)
        n_checked = 0
        for n, p in model.named_parameters():
            if n in g32:
                assert p.grad is not None and p.grad.dtype == dtype, n
                if g32[n].norm() > 1e-6 and p.numel() > 1000:
                    assert rel_l2(p.grad.float() / 128.0, g32[n]) < 6e-2, (n, rel_l2(p.grad.float() / 128.0, g32[n]))
                    n_checked += 1
        assert n_checked > 100
        # an external optimizer (FP16Optimizer copies its fp32 masters back into the half parameters) is followed
        with torch.no_grad():
            for p in model.parameters():
                if p.requires_grad:
                    p.mul_(0.5)
        model.zero_grad(set_to_none=True)
        loss_b, _, _ = crit(model, sample)
        assert abs(loss_b.item() - loss.item()) > 1e-3
        model.eval()
        with torch.no_grad():
            x, _ = model(**sample["net_input"])
        assert torch.isfinite(x).all()
