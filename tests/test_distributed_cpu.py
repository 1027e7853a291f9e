"""CPU, gloo, world_size 2: the N>1 path of the benchmark (independent replicas, per-rank seeded
shards, max-over-ranks timing, rank-0-only reference arm)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ifseg_b200.config import preset
    from ifseg_b200.distributed import aggregate_throughput, max_over_ranks, rank_world, shard_range, shard_seed
    from ifseg_b200.synthetic import synthetic_inputs

    assert rank_world() == (rank, world)
    cfg = preset("segofa_base", num_seg=15, patch_image_size=32, orig_patch_image_size=32)
    inp = synthetic_inputs(cfg, 2, 32, seed=shard_seed(1, rank))
    sums = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(sums, inp["patch_images"].sum().reshape(1))
    t = max_over_ranks([1.0 + rank, 5.0 - rank])
    q.put((rank, [s.item() for s in sums], t, shard_range(10, rank, world), aggregate_throughput(8, world, t[0] / 1e3)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, sums0, t0, sh0, thr0), (r1, sums1, t1, sh1, thr1) = res
    assert sums0 == sums1 and abs(sums0[0] - sums0[1]) > 1e-3  # ranks drew different shards
    assert t0 == t1 == [2.0, 5.0]  # max over ranks, identical everywhere
    assert sh0 == (0, 5) and sh1 == (5, 10)
    assert thr0 == 8 * 2 / 2e-3


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.strip().splitlines()) == 1  # stdout carries exactly one JSON line
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
