#!/usr/bin/env python
"""Headline benchmark: images/sec of the segofa forward->mask hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

A "step" is one pass of the hot path over one batch of synthetic images:
  ResNet-101 stem -> OFA encoder -> surrogate decoder -> seg_projection -> x16 bilinear upsample
  -> argmax mask   (SegOFAModel.forward + seg_criterion.upsample_logits/argmax; label-propagation
  and CRF post-processing excluded).
Workload at N=1: BASELINE.json configs[1] = OFA-Base, 480x480, 15 COCO-unseen classes, batch 8,
bf16 tensor-core operands / fp32 accumulate, random-init weights of that architecture from the
seeded generator (no checkpoint exists offline), synthetic randn images, the real 36-token prompt.
N>1: one process per GPU (torchrun), independent replicas on per-rank seeded batches (weak
scaling; inference has no data-path collective), barrier + device sync on both
sides, max over ranks.

value   : whole-job images/s with the inputs already resident in HBM (CUDA events on the
          launching stream; the 22 MB input batch + ~1 GB of per-step workspace exceed nothing
          special, so L2 is flushed between timed steps by a 256 MB memset, excluded from timing).
e2e     : the same metric through the public serving API (ifseg_b200.serving.SegmentationSession):
          pinned host images -> H2D -> forward -> mask -> D2H inside the timed region.
roofline: dominant kernel family (tcgen05 GEMM incl. implicit-GEMM conv) -- algorithmic FLOPs of
          its launches / their CUDA-event durations vs MEASURED_PEAKS.json bf16 sustained peak;
          `step` gives the same for the whole forward (286.8 GFLOP/img).
cpu_baseline / --impl reference: the reference algorithm on the host cores.  The reference's own
          model cannot be imported on the GPU box (/root/reference is absent, fairseq not
          installable: DESIGN.md), so this is the pinned oracle port (oracle/restated.py, fp32,
          torch CPU, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline family's largest kernel, from the
# committed `ncu --set full` captures (the roofline is tensor-bound; traffic is reported to show there are no wasted
# re-reads: it is below the algorithmic operand bytes because C tiles stay in the 126 MB L2)
NCU_TRAFFIC = {
    2: (17.14e6, "gemm_ts_kernel<256,0,2> (fc1 shape: M=7488 N=3072 K=768; algorithmic operand bytes 62.2 MB: A and B are "
                 "read once from HBM, the bf16 C tile is still in L2 when the kernel ends), profiles/r02_gemm_ts_full.txt"),
    3: (68.95e6, "gemm_mm_tcgen05_kernel<128,3,1,1,1152> (split-K weight gradient, fp32 accumulate; "
                 "algorithmic 77.9 MB for the fc weight gradient), profiles/r01_trainmisc_final_full.txt"),
}
# forward FLOPs of the attention launches of one image (QK^T + PV, causal decoder self-attention counted as executed)
FLOPS_PER_IMG = {2: 286.8e9, 4: 364.6e9, 3: 1061.2e9, 5: 6465.3e9}  # SURVEY.md s8d (forward / train step, algorithmic)
CONFIGS = {
    # id: (arch, image, classes, batch, description)
    1: ("segofa_base", 128, 15, 1, "OFA-Base segofa 128x128, 15 COCO-unseen classes, batch 1"),
    2: ("segofa_base", 480, 15, 8, "OFA-Base segofa 480x480 inference, 15 COCO-unseen classes, batch 8"),
    4: ("segofa_base", 512, 171, 8, "OFA-Base segofa 512x512 inference, 171 COCO-Stuff classes, batch 8"),
    3: ("segofa_base", 480, 150, 8, "OFA-Base segofa 480x480 image-free finetune step, 150 ADE classes, per-GPU batch 8 "
                                    "(aux fwd+bwd + no-grad real-image fwd + metrics + grad all-reduce + Adam)"),
    5: ("segofa_large", 640, 150, 4, "OFA-Large segofa 640x640 image-free finetune step, 150 ADE classes, per-GPU batch 4 "
                                     "(aux fwd+bwd + no-grad real-image fwd + metrics + grad all-reduce + Adam)"),
}
TRAIN_CONFIGS = {3, 5}


_OUT_FD = None  # set in __main__: the real stdout, which carries exactly ONE JSON line


def emit(obj):
    line = json.dumps(obj) + "\n"
    if _OUT_FD is None:
        sys.stdout.write(line)
        sys.stdout.flush()
    else:
        os.write(_OUT_FD, line.encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def prompt_tokens(num_seg):
    import torch

    p = os.path.join(ROOT, "tests", "golden", "prompts.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if str(num_seg) in d:
            return torch.tensor(d[str(num_seg)], dtype=torch.long)
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(cfg_id, steps, warmup, sample_batch=None, want_outputs=False):
    """Times the oracle port (reference algorithm, fp32 torch CPU) on all host threads."""
    import torch

    from ifseg_b200.config import preset
    from ifseg_b200.synthetic import generate_state_dict, synthetic_inputs
    from oracle import restated as R

    arch, size, nseg, batch, _ = CONFIGS[cfg_id]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = sample_batch or min(batch, 2)
    cfg = preset(arch, num_seg=nseg, patch_image_size=size, orig_patch_image_size=size)
    sd = generate_state_dict(cfg, 0)
    ocfg = R.SegOFAConfig(**{k: getattr(cfg, k) for k in R.SegOFAConfig.__dataclass_fields__ if hasattr(cfg, k)})
    inp = synthetic_inputs(cfg, b, size, seed=1, src_tokens=prompt_tokens(nseg))
    hp = size // 16

    last = {}

    def step():
        with torch.no_grad():
            lg, _ = R.segofa_forward(sd, ocfg, inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
            last["logits"], last["mask"] = lg, R.predict_mask(lg, hp, hp, size, size)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    last["inputs"] = inp
    cb = dict(value=b / dt, unit="images/s", cores=cores, kind="port",
              sample=f"{steps} timed + {warmup} warm-up forward->mask passes of batch {b} (of {batch}) at {size}x{size}, "
                     f"fp32 torch CPU, {cores} threads; oracle/restated.py pinned on the reference")
    return (cb, dt, last) if want_outputs else (cb, dt)


def gpu_eager_context(cfg_id):
    """Context, not a target: the oracle port (the reference's algorithm, op for op) run by PyTorch eager on THIS GPU under
    bf16 autocast (ATen elementwise kernels + cuBLAS GEMMs), forward -> mask, full batch.  It answers "what does the
    reference's own PyTorch path do on one B200" (BASELINE.md s4.4); part of the cpu_baseline leg because it executes
    oracle/."""
    import torch

    from ifseg_b200.config import preset
    from ifseg_b200.synthetic import generate_state_dict, synthetic_inputs
    from oracle import restated as R

    arch, size, nseg, batch, _ = CONFIGS[cfg_id]
    try:
        cfg = preset(arch, num_seg=nseg, patch_image_size=size, orig_patch_image_size=size)
        sd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in generate_state_dict(cfg, 0).items()}
        ocfg = R.SegOFAConfig(**{k: getattr(cfg, k) for k in R.SegOFAConfig.__dataclass_fields__ if hasattr(cfg, k)})
        inp = {k: v.cuda() for k, v in synthetic_inputs(cfg, batch, size, seed=1, src_tokens=prompt_tokens(nseg)).items()}
        hp = size // 16

        def step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                lg, _ = R.segofa_forward(sd, ocfg, inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
            return R.predict_mask(lg, hp, hp, size, size)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        return {"value": batch / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "batch": batch,
                "what": "oracle/restated.py under torch eager + bf16 autocast on this GPU (ATen + cuBLAS), 3 timed steps"}
    except Exception as ex:  # noqa: BLE001 -- context only
        return {"unavailable": f"{type(ex).__name__}: {str(ex)[:160]}"}
    finally:
        torch.cuda.empty_cache()


def parity_against_oracle(model, last, size):
    """The timed configuration checked against the oracle run of the cpu_baseline leg (same seeded weights and inputs):
    logits rel-L2 against the fp32 oracle, mask agreement, and how many disagreeing pixels have a top-2 margin the logit
    error cannot explain (must be 0).  tests/test_parity_gpu.py holds the gates; this is the record next to the number."""
    import torch

    inp = last["inputs"]
    with torch.no_grad():
        x, extra = model(**{k: v.cuda() for k, v in inp.items()})
        hp, wp = extra["encoder_returns"]["image_embed_shape"][0]
        mask = model.engine().predict_mask(x, (hp, wp), (size, size)).view(x.shape[0], -1).cpu()
    ref, ref_mask = last["logits"].float(), last["mask"].view(x.shape[0], -1)
    xc = x.float().cpu()
    rel = ((xc - ref).norm() / ref.norm()).item()
    max_err = (xc - ref).abs().max().item()
    from oracle import restated as R

    up = R.upsample_logits(ref, hp, wp, size, size)[:, :-1]
    top2 = up.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).view(x.shape[0], -1)
    dis = mask != ref_mask
    return {"against": "oracle/restated.py (fp32, pinned on the reference), same seeded weights and inputs",
            "batch": int(x.shape[0]), "logits_rel_l2": rel, "logits_max_abs_err": max_err,
            "mask_agree": 1.0 - dis.float().mean().item(), "mask_pixels": int(dis.numel()),
            "mask_disagree_beyond_4x_logit_err": int((dis & (margin > 4 * max_err)).sum()),
            "bf16_floor_note": "the reference's own bf16 autocast forward sits at rel-L2 1.4e-2 from its fp32 forward "
                               "(tests/test_model_gpu.py); gates and the rounding-matched oracle: tests/test_parity_gpu.py"}


def cpu_reference_train_run(cfg_id, steps, warmup):
    """One image-free train step of the oracle port on the host cores (fp32 torch CPU autograd):
    aux forward + imfree loss + backward, no-grad real-image forward + mask.  Batch 1."""
    import torch

    from ifseg_b200.config import preset
    from ifseg_b200.synthetic import generate_state_dict, synthetic_train_sample
    from oracle import restated as R

    arch, size, nseg, batch, _ = CONFIGS[cfg_id]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = preset(arch, num_seg=nseg, patch_image_size=size, orig_patch_image_size=size)
    sd = generate_state_dict(cfg, 0)
    train_keys = [k for k, v in sd.items() if v.is_floating_point() and (".layers." in k or "layer_norm" in k)
                  and "embed_images" not in k]
    sd_g = {k: (v.clone().requires_grad_() if k in train_keys else v) for k, v in sd.items()}
    ocfg = R.SegOFAConfig(**{k: getattr(cfg, k) for k in R.SegOFAConfig.__dataclass_fields__ if hasattr(cfg, k)})
    smp = synthetic_train_sample(cfg, 1, size, seed=1, src_tokens=prompt_tokens(nseg))
    hp = size // 16

    def step():
        for v in sd_g.values():
            if v.requires_grad:
                v.grad = None
        x, _ = R.segofa_forward_aux(sd_g, ocfg, smp["aux_input"])
        R.imfree_loss(x, smp["text2seg_target"], ocfg).backward()
        ni = smp["net_input"]
        with torch.no_grad():
            lg, _ = R.segofa_forward(sd_g, ocfg, ni["src_tokens"], ni["patch_images"], ni["patch_masks"])
            R.predict_mask(lg, hp, hp, size, size)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=1 / dt, unit="images/s", cores=cores, kind="port",
                sample=f"{steps} timed + {warmup} warm-up train steps of batch 1 (of {batch}) at {size}x{size}: aux fwd+bwd "
                       f"(autograd) + no-grad real-image fwd, fp32 torch CPU, {cores} threads, no optimizer; "
                       f"oracle/restated.py pinned on the reference"), dt


def bench_train(args, rank, world, local_rank, config):
    import torch.distributed as dist

    out = measure_train(args, rank, world, local_rank, args.config, config, args.steps, max(args.warmup, 3),
                        cpu_baseline=not args.no_cpu_baseline)
    if rank == 0:
        emit(out)
    teardown(world)


def teardown(world):
    """Multi-rank exit.  A CUDA graph that captured NCCL all-reduces keeps the communicator's stream dependencies alive and
    destroy_process_group() then hangs with this torch/NCCL build (r01); the process has nothing left to flush, so every
    rank synchronises, meets the others at a barrier and leaves without tearing the communicator down."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stderr.flush()
    os._exit(0)


def measure_train(args, rank, world, local_rank, cfg_id, config, steps, warmup, cpu_baseline=False):
    """One image-free finetune step (aux fwd+bwd + no-grad real-image fwd + metrics + gradient all-reduce + Adam) timed
    on the device as the max over ranks; returns the record on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    from ifseg_b200 import ops
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.synthetic import generate_state_dict, synthetic_train_sample
    from ifseg_b200.trainer import SegOFATrainer, TrainSession

    arch, size, nseg, batch, desc = CONFIGS[cfg_id]
    peaks = load_peaks()
    model = SegOFAModel.from_config(arch, nseg, size)
    model.load_state_dict(generate_state_dict(model.cfg, 0), strict=True)
    model = model.cuda()
    trainer = SegOFATrainer(model)  # shipped recipe: Adam lr 5e-5, wd 0.1, clip 1.0 (Appendix B)
    if world > 1:
        os.environ.setdefault("SGF_GRAPH_NCCL", "1")  # capture the bucketed all-reduces too (see teardown())
    smp = synthetic_train_sample(model.cfg, batch, size, seed=1 + rank, src_tokens=prompt_tokens(nseg))
    pinned = {k: ({kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else
                  (v.pin_memory() if torch.is_tensor(v) else v)) for k, v in smp.items()}
    sess = TrainSession(trainer, smp, use_cuda_graph=not args.no_graph)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        sess.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    with torch.cuda.stream(sess.stream):
        for _ in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(sess.stream)
            sess.step()
            e.record(sess.stream)
            evs.append((s, e))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s.elapsed_time(e) for s, e in evs) / steps
    # end to end: pinned host sample -> device buffers -> step -> loss back on the host, every step
    def e2e_step():
        sess.load(pinned)
        out = sess.step()
        with torch.cuda.stream(sess.stream):
            return float(out["loss"])  # D2H of the loss on the session stream (host sync)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = sum(v.numel() * v.element_size() for d in (pinned["net_input"], pinned["aux_input"]) for v in d.values()) + \
        pinned["target"].numel() * 8 + pinned["text2seg_target"].numel() * 8
    # per-kernel attribution (eager)
    timer = ops.KernelTimer(fine=args.kernel_breakdown)
    ops.set_timer(timer)
    with torch.cuda.stream(sess.stream):
        trainer.train_step(sess.static, check_pads=False)
    fam_fine = timer.summary()
    ops.set_timer(None)
    fam = {}
    for k, v in fam_fine.items():
        d = fam.setdefault(k.split(":")[0], dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for kk in d:
            d[kk] += v[kk]
    total_k_ms = sum(v["ms"] for v in fam.values())
    if args.kernel_breakdown and rank == 0:
        for k, v in sorted(fam_fine.items(), key=lambda kv: -kv[1]["ms"]):
            sys.stderr.write(f"{k:34s} launches {v['launches']:4d}  {v['ms']:8.3f} ms  {100 * v['ms'] / total_k_ms:5.1f}%  "
                             f"{v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f} TFLOP/s  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s\n")
    t = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = t[0].item(), t[1].item()
    n = max(world, 1)
    if rank == 0:
        d = fam["gemm_tcgen05"]
        # family time inside the timed (graph) step = its share of the per-launch CUDA-event times x the step time
        achieved = d["flops"] / (d["ms"] / total_k_ms * dev_ms / 1e3) / 1e12
        step_tflops = FLOPS_PER_IMG[cfg_id] * batch / (dev_ms / 1e3) / 1e12
        out = {
            "metric": "images/sec (train step)", "value": n * batch / (dev_ms / 1e3), "unit": "images/s", "n_gpus": n,
            "steps": steps, "warmup": warmup, "ms_per_step": dev_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": n * batch * steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": sess.launches_per_step * steps, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05", "achieved": achieved,
                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                         "traffic": NCU_TRAFFIC.get(cfg_id, (None, None))[0],
                         "traffic_source": NCU_TRAFFIC.get(cfg_id, (None, None))[1],
                         "peak_source": peaks["source"] + " (bf16 sustained)",
                         "launches_per_step": d["launches"], "share_of_step_kernel_time": d["ms"] / total_k_ms,
                         "step": {"achieved": step_tflops, "frac": step_tflops / peaks["bf16_sustained"],
                                  "flops_per_image": FLOPS_PER_IMG[cfg_id]}},
            "kernel_families": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in fam.items()},
            "cuda_graph": sess.graph is not None,
            "train": {"dropout": trainer.engine.drop_p if trainer.engine.stochastic else 0.0,
                      "drop_path_max": max(trainer.engine.enc_dpr) if trainer.engine.stochastic else 0.0,
                      "note": "shipped recipe noise (dropout 0.1, DropPath 0..0.1) on; all 380 gradient tensors; Adam + clip"},
        }
        out["grad_allreduce"] = {"world": n, "bytes_per_step": int(trainer.engine.arena.grad32.numel() * 4) if n > 1 else 0,
                                 "dtype": "f32", "how": "bucketed NCCL all-reduce of the flat fp32 gradient arena, launched "
                                 "per layer bucket as the backward reaches it" if n > 1 else "single rank: none"}
        if cpu_baseline:
            cb, _ = cpu_reference_train_run(cfg_id, 1, 0)
            out["cpu_baseline"] = cb
        return out
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-record", action="store_true",
                    help="skip the short cfg-3 train-step measurement that is attached to the inference line as `train_step`")
    ap.add_argument("--kernel-breakdown", action="store_true", help="print the per-kernel-family table to stderr")
    ap.add_argument("--profile-step", action="store_true",
                    help="wrap ONE eager forward->mask step in cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    arch, size, nseg, batch, desc = CONFIGS[args.config]
    config = {"workload": desc, "arch": arch, "image_size": size, "num_classes": nseg, "per_gpu_batch": batch,
              "global_batch": batch * max(world, 1), "parallelism": f"dp{max(world, 1)} (independent replicas)",
              "l2_policy": "256 MB memset between timed steps (not timed)",
              "bias_cached": True}  # the additive position bias is parameter-only: built once per shape, not per step

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        if args.config in TRAIN_CONFIGS:
            cb, dt = cpu_reference_train_run(args.config, min(steps, 2), min(args.warmup, 1))
        else:
            cb, dt = cpu_reference_run(args.config, steps, min(args.warmup, 1))
        emit({
            "impl": "reference", "metric": "images/sec (train step)" if args.config in TRAIN_CONFIGS else "images/sec",
            "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0},
        })
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the segofa_b200 hot path has no CPU fallback "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    if args.config in TRAIN_CONFIGS:
        config["l2_policy"] = "256 MB memset between timed steps (not timed); a step streams > 4 GB of activations"
        return bench_train(args, rank, world, local_rank, config)
    from ifseg_b200 import ops
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.serving import SegmentationSession
    from ifseg_b200.synthetic import generate_state_dict, synthetic_inputs

    peaks = load_peaks()
    model = SegOFAModel.from_config(arch, nseg, size)
    model.load_state_dict(generate_state_dict(model.cfg, 0), strict=True)
    model = model.cuda().eval()
    tokens = prompt_tokens(nseg)
    inp = synthetic_inputs(model.cfg, batch, size, seed=1 + rank, src_tokens=tokens)
    sess = SegmentationSession(model, batch, size, inp["src_tokens"][0], use_cuda_graph=not args.no_graph)
    sess.images.copy_(inp["patch_images"].cuda())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        sess.step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    ops.reset_launch_count()
    with torch.cuda.stream(sess.compute):
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(sess.compute)
            sess.step_device()
            e.record(sess.compute)
            evs.append((s, e))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s.elapsed_time(e) for s, e in evs) / args.steps
    launches = sess.launches_per_step * args.steps + args.steps  # + the flush memset is torch's, not counted: ours only
    launches = sess.launches_per_step * args.steps

    # ---------------- end-to-end (host buffers) ----------------
    pinned = [inp["patch_images"].clone().pin_memory() for _ in range(2)]
    for _ in sess.infer_stream((pinned[i & 1] for i in range(max(args.warmup, 3))), pinned_inputs=True):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for out in sess.infer_stream((pinned[i & 1] for i in range(args.steps)), pinned_inputs=True):
        n_out += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert n_out == args.steps

    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        with torch.no_grad(), torch.cuda.stream(sess.compute):
            sess._forward()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---------------- per-kernel attribution (eager, CUDA events per launch) ----------------
    timer = ops.KernelTimer(fine=args.kernel_breakdown)
    ops.set_timer(timer)
    with torch.no_grad(), torch.cuda.stream(sess.compute):
        sess._forward()
    fam_fine = timer.summary()
    ops.set_timer(None)
    fam = {}
    for k, v in fam_fine.items():  # fold the call-site tags back into kernel families
        d = fam.setdefault(k.split(":")[0], dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for kk in d:
            d[kk] += v[kk]
    total_k_ms = sum(v["ms"] for v in fam.values())
    dom = max(fam.items(), key=lambda kv: kv[1]["ms"])
    if args.kernel_breakdown and rank == 0:
        for k, v in sorted(fam_fine.items(), key=lambda kv: -kv[1]["ms"]):
            sys.stderr.write(f"{k:34s} launches {v['launches']:4d}  {v['ms']:8.3f} ms  {100 * v['ms'] / total_k_ms:5.1f}%  "
                             f"{v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f} TFLOP/s  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s\n")

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = t[0].item(), t[1].item()
    n = max(world, 1)
    value = n * batch / (dev_ms / 1e3)
    e2e_value = n * batch * args.steps / e2e_s

    if rank == 0:
        dom_name = "gemm_tcgen05" if "gemm_tcgen05" in fam and fam["gemm_tcgen05"]["ms"] >= 0.3 * dom[1]["ms"] else dom[0]
        d = fam[dom_name]
        # time of a kernel family inside the timed (CUDA-graph) step = its share of the per-launch CUDA-event times of one
        # eager step x the graph step time (the eager per-launch events carry ~3 us of launch gap each; the SHARES agree
        # with the ncu launch list, profiles/r02_launches_*.txt)
        fam_ms = lambda v: v["ms"] / total_k_ms * dev_ms  # noqa: E731
        achieved = d["flops"] / (fam_ms(d) / 1e3) / 1e12
        step_tflops = FLOPS_PER_IMG.get(args.config, 0.0) * batch / (dev_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peaks["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                "traffic": NCU_TRAFFIC.get(args.config, (None, None))[0],
                "traffic_source": NCU_TRAFFIC.get(args.config, (None, None))[1],
                "peak_source": peaks["source"] + " (bf16 sustained: kernel timed inside a long step)",
                "launches_per_step": d["launches"], "share_of_step_kernel_time": d["ms"] / total_k_ms,
                "ms_in_step": fam_ms(d),
                "step": {"achieved": step_tflops, "frac": step_tflops / peaks["bf16_sustained"],
                         "flops_per_image": FLOPS_PER_IMG.get(args.config)}}
        if "attention_tcgen05" in fam:  # north_star: achieved fraction of the attention-GEMM roofline
            a = fam["attention_tcgen05"]
            a_tf = a["flops"] / (fam_ms(a) / 1e3) / 1e12
            roof["attention"] = {"kernel": "attention_tcgen05 (QK^T + PV tcgen05.mma, fused bias/softmax)", "achieved": a_tf,
                                 "frac": a_tf / peaks["bf16_sustained"], "unit": "TFLOP/s", "launches_per_step": a["launches"],
                                 "ms_in_step": fam_ms(a), "share_of_step_kernel_time": a["ms"] / total_k_ms}
        if "row_layernorm" in fam:  # the HBM-bound family: algorithmic bytes (rows*D*(in + residual + outs)) / time in the step
            ln = fam["row_layernorm"]
            gbps = ln["bytes"] / (fam_ms(ln) / 1e3) / 1e9
            roof["row_layernorm"] = {"bound": "hbm", "achieved": gbps, "peak": peaks["hbm"], "unit": "GB/s",
                                     "frac": gbps / peaks["hbm"], "launches_per_step": ln["launches"],
                                     "ms_in_step": fam_ms(ln), "algorithmic_bytes_per_step": ln["bytes"],
                                     "note": "inside the step the rows are L2-resident (written by the preceding GEMM), so the "
                                             "effective rate can exceed what a cold ncu capture of the same launch shows"}
        out = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": n, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": sess.h2d_bytes_per_step,
                    "d2h_bytes_per_step": sess.d2h_bytes_per_step},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "kernel_families": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in fam.items()},
            "cuda_graph": not args.no_graph,
        }
        if not args.no_cpu_baseline:
            cb, _, last = cpu_reference_run(args.config, 1, 1, sample_batch=1 if size >= 256 else None, want_outputs=True)
            out["cpu_baseline"] = cb
            out["parity"] = parity_against_oracle(model, last, size)
            out["gpu_eager_context"] = gpu_eager_context(args.config)
    else:
        out = None
    # ---------------- the training step of the same recipe on the same N ranks (cfg 3), as a sub-record ----------------
    if not args.no_train_record:
        del sess
        torch.cuda.empty_cache()
        t_arch, t_size, t_nseg, t_batch, t_desc = CONFIGS[3]
        t_cfg = {"workload": t_desc, "arch": t_arch, "image_size": t_size, "num_classes": t_nseg, "per_gpu_batch": t_batch,
                 "global_batch": t_batch * n, "parallelism": f"dp{n} (gradient all-reduce over NCCL)"}
        tr = measure_train(args, rank, world, local_rank, 3, t_cfg, min(args.steps, 8), 3, cpu_baseline=False)
        if rank == 0:
            keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "config", "e2e", "gpu_launches",
                    "roofline", "cuda_graph", "grad_allreduce", "train")
            out["train_step"] = {k: tr[k] for k in keep if k in tr}
    if rank == 0:
        emit(out)
    teardown(world)


if __name__ == "__main__":
    # anything a library prints to fd 1 while we run (NCCL's version banner under NCCL_DEBUG, for instance) goes to
    # stderr, so that stdout is the one JSON line the contract asks for
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)
    main()
