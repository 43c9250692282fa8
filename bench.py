#!/usr/bin/env python
"""Headline benchmark: edited images/sec @ SD-1.5 512^2 (64x64 latent), 50 DDIM steps, implicit h-Edit + P2P
(BASELINE.json configs[1]: batch 8 per GPU).  One JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 2 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port of its loop) on the host cores

A "step" = one full 50-timestep edit of one batch of 8 synthetic images per GPU (weights: seeded random init of the
SD-1.5 UNet geometry -- no pretrained weights exist offline; latents/noise: seeded Gaussians).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TFLOP_PER_SAMPLE_FORWARD = 0.8033            # SD-1.x UNet, 64x64 latent, 77 ctx tokens
GEMM_TFLOP_PER_SAMPLE_FORWARD = 2 * (200.16 + 138.42) * 1e-3      # conv3x3 + linear/1x1 GMAC of one sample-forward
PROMPT_PAIRS = [
    (["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"], ("lizard", "lizard")),
    (["a cat sitting next to a mirror", "a silver cat sculpture sitting next to a mirror"], ("cat", "cat")),
    (["a photo of a house on a hill", "a photo of a castle on a hill in winter"], ("house", "castle")),
    (["two birds on a wire", "two parrots on a wire"], ("birds", "parrots")),
]


def ncu_traffic(kernel_prefix):
    """dram read+write bytes per launch of the dominant kernel from the committed `ncu --set full` capture (tools/ncu_summary.py)."""
    p = os.path.join(ROOT, "profiles", "ncu_full_summary.json")
    if not os.path.exists(p):
        return None
    for name, d in json.load(open(p)).items():
        if kernel_prefix in name and "conv" in d.get("file", ""):
            return d["dram_read_bytes"] + d["dram_write_bytes"]
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1401.7), d.get("hbm_gbs", 6451.2), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampler running during the timed region."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            clocks = sorted(float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit())
            if clocks:
                out["sm_mhz"] = clocks[len(clocks) // 2]
                out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, n in enumerate(names):
                if any(len(r) > 2 + k and r[2 + k].strip().lower() == "active" for r in rows):
                    out["reasons"].append(n)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(T, n_timesteps, threads):
    """Times the oracle port of the reference loop (its own h_Edit_p2p_implicit restated, fp32, torch CPU) on a bounded
    sample: `n_timesteps` of the T-step schedule for ONE image (9 UNet sample-forwards each).  Returns sec/timestep."""
    import torch
    from oracle import h_edit as oh
    from oracle import p2p as op
    from oracle.pipeline import OraclePipeline
    from oracle.sd_unet import UNetConfig

    torch.set_num_threads(threads)
    model = OraclePipeline(UNetConfig.sd15(), seed=0)
    model.scheduler.set_timesteps(T)
    prompts, (bs, bt) = PROMPT_PAIRS[0]
    spec = op.make_edit_spec(prompts, False, 0.4, 0.35, ((bs,), (bt,)), {"words": (bt,), "values": (2.0,)}, T, model.tokenizer)
    enc = lambda p: model.text_encoder(model.tokenizer(p).input_ids)[0]
    g = torch.Generator().manual_seed(0)
    xT = torch.randn(1, 4, 64, 64, generator=g)
    zs = torch.randn(T, 4, 64, 64, generator=g)
    t0 = time.perf_counter()
    oh.h_edit_p2p_implicit(model.unet, model.scheduler, enc([""]), enc([prompts[0]]), enc([prompts[1]]), xT, zs, spec, [1.0, 5.0, 7.5], eta=1.0,
                           weight_reconstruction=0.1, optimization_steps=1, after_skip_steps=n_timesteps)
    return (time.perf_counter() - t0) / n_timesteps


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    T = args.timesteps
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(T, 1, threads)
    per_ts = []
    for _ in range(args.steps):
        per_ts.append(cpu_reference_sample(T, 1, threads))
    sec_ts = sum(per_ts) / len(per_ts)
    ips = 1.0 / (sec_ts * T)
    sample = f"1 of {T} timesteps of 1 image per step (9 UNet sample-forwards, fp32, torch CPU), extrapolated x{T}"
    line = {
        "impl": "reference", "metric": "edited images/sec @ SD-1.5 512^2, 50-step implicit h-Edit+P2P", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_ts * T * 8 * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "implicit h-Edit-R + P2P (Refine+Reweight+LocalBlend), SD-1.5 UNet geometry random-init, 64x64 latent, 50 DDIM steps, batch 8/GPU",
                   "global_batch": 8 * args.gpus, "timesteps": T},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    import hedit_b200
    from hedit_b200.dist import gather_results, max_over_ranks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B, T = args.batch, args.timesteps
    cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
               cross_attention_dim=768, norm_groups=32, ctx_len=77)
    eng = hedit_b200.UNetEngine(cfg, max_samples=5 * B, max_contexts=1 + 2 * B, device=local_rank)
    eng.load_random_weights(seed=0)
    tok = hedit_b200.WordTokenizer()
    sched = hedit_b200.DDIMTables(T, steps_offset=1)
    ts, coef = hedit_b200.step_tables(sched, T, 1.0, False)
    ctrls = []
    for b in range(B):
        prompts, (bs, bt) = PROMPT_PAIRS[(rank * B + b) % len(PROMPT_PAIRS)]
        ctrls.append(hedit_b200.make_controller(prompts, False, 0.4, 0.35, blend_word=((bs,), (bt,)),
                                                equilizer_params={"words": (bt,), "values": (2.0,)}, num_steps=T, tokenizer=tok))
    plan = hedit_b200.compile_edit_plan(ctrls, T)
    g = torch.Generator().manual_seed(1234 + rank)
    xT_h = torch.randn(B, 4, 64, 64, generator=g).pin_memory()
    zs_h = torch.randn(B, T, 4, 64, 64, generator=g).pin_memory()
    ctx_h = (torch.randn(1 + 2 * B, 77, 768, generator=g)).pin_memory()      # synthetic text-encoder outputs
    xT_d, zs_d, ctx_d = xT_h.to(dev), zs_h.to(dev), ctx_h.to(dev)
    cfgs = [1.0, 5.0, 7.5]

    def step(host, schedule=None):
        schedule = args.schedule if schedule is None else schedule
        if host:
            ed, rc = eng.edit(xT_h, zs_h, ctx_h, ts, coef, cfgs, plan, 0.1, 1, False, schedule)
            ed_dev = ed.to(dev, non_blocking=True) if world > 1 else ed
        else:
            ed, rc = eng.edit(xT_d, zs_d, ctx_d, ts, coef, cfgs, plan, 0.1, 1, False, schedule)
            ed_dev = ed
        if world > 1:
            gather_results(ed_dev, B * world)      # the only collective: final result gather over NCCL
        return eng.last_stats

    def timed(host, n, schedule=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fwd = launches = 0
        for _ in range(n):
            st = step(host, schedule)
            fwd += st["sample_forwards"]; launches += st["kernel_launches"]
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, fwd, launches

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, fwd, launches = timed(False, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    step(True)                                       # warm the host path once
    ms_e2e, _, _ = timed(True, args.steps)
    ips = B * world * args.steps / (ms / 1e3)
    ips_e2e = B * world * args.steps / (ms_e2e / 1e3)
    peak_tf, _, peak_src = measured_peaks()
    achieved_tf = fwd * TFLOP_PER_SAMPLE_FORWARD / (ms / 1e3)          # this rank's UNet work / max-over-ranks time
    h2d = (xT_h.numel() + zs_h.numel() + ctx_h.numel()) * 4 + plan.c_base.nbytes + plan.c_tar.nbytes + plan.mapper.nbytes + plan.blend_alpha.nbytes
    d2h = 2 * xT_h.numel() * 4
    # opt-in schedule 2 (cfg_src == 1: u + 1*(c-u) == c, 5 UNet sample-forwards per step), reported next to the headline
    skip = None
    if args.schedule == 1:
        ms2, fwd2, _ = timed(False, 1, schedule=2)
        skip = {"value": B * world / (ms2 / 1e3), "unit": "images/s", "sample_forwards_per_image": fwd2 / B,
                "note": "schedule 2: same edit up to one fp32 rounding per element of the source-guided noise; not the headline"}
    # live per-kernel timing (CUDA events around every launch of one UNet forward of 5B samples) -> the dominant kernel's roofline
    prof = gemm = None
    if rank == 0:
        pf = eng.profile_forward(5 * B, 2)
        prof = {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in sorted(pf.items(), key=lambda kv: -kv[1][0])}
        gemm_tags = [k for k in pf if k.startswith(("res.", "tf.", "upsample.conv", "downsample", "conv_out"))]
        gemm_ms = sum(pf[k][0] for k in gemm_tags)
        gemm_n = sum(pf[k][1] for k in gemm_tags)
        total_ms = sum(v[0] for v in pf.values())
        gemm_tf = 5 * B * GEMM_TFLOP_PER_SAMPLE_FORWARD / (gemm_ms / 1e3)
        gemm = {"kernel": "gemm_bf16_tcgen05_kernel (implicit-GEMM conv3x3 + linear launches)", "launches_per_forward": gemm_n,
                "avg_launch_us": 1e3 * gemm_ms / gemm_n, "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf,
                "share_of_forward": gemm_ms / total_ms,
                "algorithmic": "2*M*N*K per launch; 338.58 GMAC per UNet sample-forward over all GEMM launches (SURVEY 8d)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sec_ts = cpu_reference_sample(T, 1, threads)
        cpu = {"value": 1.0 / (sec_ts * T), "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"1 of {T} timesteps of 1 image (9 UNet sample-forwards, fp32 torch CPU oracle port of the reference loop), extrapolated x{T}"}
    if rank == 0:
        line = {
            "metric": "edited images/sec @ SD-1.5 512^2, 50-step implicit h-Edit+P2P", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": f"{hedit_b200._lib.load().hedit_operand_dtype().decode()} operands, f32 accumulate/residual/softmax/scheduler",
            "data": "synthetic",
            "config": {"workload": "implicit h-Edit-R + P2P (Refine+Reweight+LocalBlend), SD-1.5 UNet geometry random-init, 64x64 latent (512^2), 50 DDIM steps, batch 8/GPU",
                       "global_batch": B * world, "timesteps": T, "optimization_steps": 1,
                       "schedule": {0: "reference (9 UNet sample-forwards/step)", 1: "exact-reuse (7 UNet sample-forwards/step)",
                                    2: "exact-reuse + skip-uncond at cfg_src=1 (5 UNet sample-forwards/step)"}[args.schedule],
                       "sample_forwards_per_image": fwd / (B * args.steps),
                       "launch": "repeated UNet launches replayed from CUDA graphs (HEDIT_LOOP_GRAPH=0: direct); gpu_launches counts the kernels inside",
                       "l2": "working set per step (1.7 GB fp16 weights + GBs of activations) >> 126 MB L2; no explicit flush needed"},
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": ncu_traffic("gemm_bf16_tcgen05_kernel"), "peak_source": peak_src, "dominant_kernel": gemm,
                         "note": "achieved = executed UNet sample-forwards x 0.8033 TFLOP / device time of the timed region (per GPU), i.e. the whole "
                                 "step incl. attention, norms and the h-step kernels; dominant_kernel = the GEMM/conv kernel alone, timed live with CUDA "
                                 "events around each of its launches; traffic = dram read+write bytes of one conv launch (S=16, 64x64, 320->320: "
                                 "algorithmic 42 MB in + 84 MB out, L2-absorbed writes) from the committed ncu --set full capture (profiles/)"},
            "skip_uncond_schedule": skip,
            "cpu_baseline": cpu,
        }
        if prof and args.profile:
            line["kernel_breakdown_ms_per_forward"] = prof
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--timesteps", type=int, default=50)
    ap.add_argument("--schedule", type=int, default=1, choices=[0, 1, 2])
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
