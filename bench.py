#!/usr/bin/env python
"""Benchmark of the generator-inference hot path (BASELINE.json metric: clips/sec of 34-frame
gestures).  One "step" = one pass of the hot path over one batch of synthetic TED-shaped clips:
raw 16 kHz audio -> log-mel (F1-F4b) -> SE-ResNet audio encoder -> transformer generator ->
34-frame poses (+ the NCCL pose all_gather when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips-per-gpu B] [--impl reference]

Own arm prints ONE JSON line (rank 0) with value (inputs resident in HBM), e2e (pinned host
buffers, H2D/D2H inside the timed region), roofline of the dominant kernel family (trunk
convolutions; CUDA-event pairs recorded by libegx around every launch of the timed steps),
per-stage times, cpu_baseline (oracle port on the host cores, bounded sample) and clocks.
`--impl reference` times the CPU oracle port alone (the reference is pure PyTorch; its CPU
path restated in oracle/ and pinned against the real reference by oracle/make_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips_per_sec_34frame_generator_inference"
UNIT = "clips/s"
WORKLOAD = "Full_model generator inference, TED-Emotion shape (2.27 s 16 kHz audio -> 34 poses), random-init"

# algorithmic work per TED clip
S3_FLOP_PER_CLIP = 4.247e9           # trunk convolutions (layer1-3 + final conv), 2*MAC
STAGE_WORK = {                        # (bound, work per clip): bytes for hbm, flop for tensor
    "S1_frontend": ("hbm", 180_908.0),
    "S2_stem": ("hbm", 609_280.0),
    # trunk convolutions per layer (SURVEY.md §8(d) gives 991 / 1248 / 1963 + 45 MFLOP per clip).  Layers 2-3 are
    # tensor-bound.  Layer 1 (32 channels) is HBM-bound: 5 map transfers of 128*70*32*2 B per block (conv1: x in,
    # y1 out; conv2: y1 + residual in, block output out), 3 blocks, for 122 us of math at peak per conv
    "S3_conv_layer1": ("hbm", 3 * 5 * 128 * 70 * 32 * 2.0),
    "S3_conv_layer2": ("tensor", 1.248e9),
    "S3_conv_layer3": ("tensor", 1.963e9 + 0.045e9),
    # S4 (SE gate * y + residual + ReLU) is fused into conv2's epilogue; what is left under this tag are the three
    # small launches per block that compute the gate ahead of conv2, so SURVEY.md §8(d) has S3+S4 reported jointly
    # against the tensor roofline (see "S3+S4_trunk" below) and S4 alone carries no roofline of its own
    "S5_proj_gemm": ("tensor", 92e6),
    "S6_enc_dec": ("tensor", 442e6),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"],
                "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def conv_traffic_per_launch(B, layer=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per trunk-convolution launch (of one layer, or of all), from the
    committed ncu pass over one step (profiles/r1_conv_dram.json, captured at `clips` clips and scaled linearly:
    every conv streams its maps once)."""
    p = os.path.join(ROOT, "profiles", "r1_conv_dram.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if layer is not None:
        d = dict(d.get("layers", {}).get(layer, {}), clips=d["clips"])
        if "dram_bytes_per_launch" not in d:
            return None
    return d["dram_bytes_per_launch"] * B / d["clips"]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(sd, cfg, audio_np, prior):
    import torch
    from oracle import generator as og
    from oracle import logmel as ol
    spec = torch.from_numpy(ol.logmel(audio_np, cfg.spec_w, "log_in")).float()
    with torch.no_grad():
        return og.generator_forward(sd, cfg, spec, prior)[0]


def cpu_setup(n_clips):
    import torch
    from emotiongestures_b200 import TED, Transformer, randomize_norm_stats_
    from oracle import synth
    torch.manual_seed(0)
    gen = Transformer.from_config(TED).eval()
    randomize_norm_stats_(gen, 1)
    sd = {k: v.detach() for k, v in gen.state_dict().items()}
    audio = synth.synth_audio(n_clips, TED.n_audio, seed=1000)
    prior = torch.from_numpy(synth.synth_prior(n_clips, TED.prior_frames, TED.pose_dim, 1000))
    return gen, sd, audio, prior


def time_cpu(n_clips, steps, warmup):
    import torch
    from emotiongestures_b200 import TED
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, sd, audio, prior = cpu_setup(n_clips)
    for _ in range(warmup):
        cpu_reference_step(sd, TED, audio, prior)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, TED, audio, prior)
    dt = (time.perf_counter() - t0) / steps
    return n_clips / dt, dt * 1e3, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_clips = args.cpu_clips
    value, ms, cores = time_cpu(n_clips, args.steps, args.warmup)
    sample = f"{n_clips} TED clips per step (log-mel fp64 numpy + generator fp32 torch), {cores} torch threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": n_clips, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch
    import torch.distributed as dist

    from emotiongestures_b200 import LOGMEL_LOG_IN, TED, Transformer, randomize_norm_stats_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = TED
    B = args.clips_per_gpu
    torch.manual_seed(0)                      # identical random-init weights on every rank
    gen = Transformer.from_config(cfg).eval()
    randomize_norm_stats_(gen, 1)
    gen = gen.to(dev)
    gen.precision = args.precision
    eng = gen.engine(args.precision)

    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    audio = (0.1 * torch.randn(B, cfg.n_audio, generator=g, device=dev)).clamp_(-1, 1)
    prior = torch.randn(B, cfg.prior_frames, cfg.pose_dim, generator=g, device=dev)
    gathered = [torch.empty(B, cfg.frames, cfg.pose_dim, device=dev) for _ in range(world)] if world > 1 else None

    def step(a, p):
        spec = eng.logmel(a, LOGMEL_LOG_IN, True)
        poses = eng.generator_forward(spec, p, None)[0]
        if world > 1:
            dist.all_gather(gathered, poses)      # final pose gather over NVLink (north star)
        return poses

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step(audio, prior)
    sync_all()
    launches0 = eng.launch_count
    eng.profile_enable(200 * args.steps)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step(audio, prior)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    stages = eng.profile_read()
    eng.profile_enable(0)
    launches = eng.launch_count - launches0
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end: pinned host buffers in, poses out to the host, copies inside the timed region ----
    h_audio = audio.cpu().pin_memory()
    h_prior = prior.cpu().pin_memory()
    h_poses = torch.empty(B, cfg.frames, cfg.pose_dim).pin_memory()

    d_poses = torch.empty(B, cfg.frames, cfg.pose_dim, device=dev) if world > 1 else None

    def e2e_step():
        # public API: pinned host in -> pinned host out, chunked so PCIe copies overlap the kernels
        eng.infer_host(h_audio, h_prior, h_poses, chunk=args.e2e_chunk, mode=LOGMEL_LOG_IN, preemph=True, poses_dev=d_poses)
        if world > 1:
            dist.all_gather(gathered, d_poses)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    sync_all()
    e2e_steps = max(3, args.steps // 2)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (t.item() / e2e_steps * 1e-3)
    h2d = h_audio.numel() * 4 + h_prior.numel() * 4
    d2h = h_poses.numel() * 4

    # ---- small batches are launch-bound: eager launches vs one CUDA-graph replay (Engine.capture), 1 and 8 clips ----
    small = None
    if rank == 0 and world == 1:
        small = {}
        for nb in (1, 8):
            a, p_ = audio[:nb].contiguous(), prior[:nb].contiguous()
            path = eng.capture(nb, LOGMEL_LOG_IN, True)
            res = {}
            for name, fn in (("eager", lambda: step(a, p_)), ("graph", lambda: path(a, p_))):
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(100):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[name + "_ms"] = e0.elapsed_time(e1) / 100
            small["clips_%d" % nb] = res

    if rank == 0:
        peaks = load_peaks()
        per_stage = {}
        for nm, (ms, cnt) in stages.items():
            ms_per_step = ms / args.steps
            ent = {"ms_per_step": ms_per_step, "launches_per_step": cnt / args.steps}
            if nm in STAGE_WORK and ms_per_step > 0:
                bound, work = STAGE_WORK[nm]
                if bound == "hbm":
                    ach = work * B / (ms_per_step * 1e-3) / 1e9
                    ent.update(bound="hbm", achieved=ach, unit="GB/s", frac=ach / peaks["hbm_gbs"])
                else:
                    ach = work * B / (ms_per_step * 1e-3) / 1e12
                    ent.update(bound="tensor", achieved=ach, unit="TFLOP/s", frac=ach / peaks["tf_sustained"])
            per_stage[nm] = ent
        # all trunk convolutions together against the tensor roofline, alone and with what
        # is left of S4 (the SE gate ahead of conv2; its gate*y + residual pass is fused into conv2's epilogue)
        layers = {k: v for k, v in per_stage.items() if k.startswith("S3_conv_layer")}
        if layers:
            ms3 = sum(v["ms_per_step"] for v in layers.values())
            n3 = sum(v["launches_per_step"] for v in layers.values())
            ach = S3_FLOP_PER_CLIP * B / (ms3 * 1e-3) / 1e12
            per_stage["S3_trunk_conv"] = {"ms_per_step": ms3, "launches_per_step": n3, "bound": "tensor", "achieved": ach,
                                          "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"]}
            s4 = per_stage.get("S4_se")
            if s4:
                ms34 = ms3 + s4["ms_per_step"]
                ach = S3_FLOP_PER_CLIP * B / (ms34 * 1e-3) / 1e12
                per_stage["S3+S4_trunk"] = {"ms_per_step": ms34, "launches_per_step": n3 + s4["launches_per_step"],
                                            "bound": "tensor", "achieved": ach, "unit": "TFLOP/s",
                                            "frac": ach / peaks["tf_sustained"]}
        # roofline of the dominant kernel family: the trunk layer with the largest share of the step, against the
        # roofline that bounds it (layer 1: HBM; layers 2-3: tensor pipe)
        dom = max(layers, key=lambda k: layers[k]["ms_per_step"]) if layers else None
        roofline = None
        if dom:
            st = layers[dom]
            n_l = max(1.0, st["launches_per_step"])
            bound, work = STAGE_WORK[dom]
            lay = dom.replace("S3_conv_", "")
            roofline = {
                "bound": bound,
                "kernel": "trunk convolutions of %s (%d launches per step: %s)" % (
                    lay, n_l, {"layer1": "conv_tc_kernel<32,32> SE-sum and gated-residual flavours",
                               "layer2": "conv_tc_kernel<64,64> + the stride-2 / 1x1 convs of its first block",
                               "layer3": "conv128_tc_kernel + the stride-2 / 1x1 convs of its first block + final conv"}.get(lay, lay)),
                "achieved": st["achieved"], "peak": peaks["hbm_gbs"] if bound == "hbm" else peaks["tf_sustained"],
                "unit": st["unit"], "frac": st["frac"], "traffic": conv_traffic_per_launch(B, lay),
                "peak_source": peaks["src"] + (" (copy bandwidth)" if bound == "hbm" else
                                               " (sustained bf16/fp16 dense, kernel timed inside a long step)"),
                ("bytes_per_launch" if bound == "hbm" else "flop_per_launch"): work * B / n_l,
                "ms_per_launch": st["ms_per_step"] / n_l,
                "all_trunk_convs_vs_tensor_peak": per_stage.get("S3_trunk_conv", {}).get("frac"),
            }
        cpu_val, cpu_ms, cores = (None, None, os.cpu_count())
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_val, cpu_ms, cores = time_cpu(args.cpu_clips, 24, 2)
            cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_clips} TED clips x 24 steps (oracle port: log-mel fp64 numpy + "
                             f"generator fp32 torch, {cores} threads)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if args.precision == "tc" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": B, "global_batch": B * world,
                       "precision": args.precision, "logmel": "preemph+log+InstanceNorm (F4b)",
                       "parallelism": f"dp{world} (clip-sharded, pose all_gather)",
                       "l2": "inputs larger than L2 (%.0f MB audio per step)" % (audio.numel() * 4 / 1e6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "Engine.infer_host (pinned host in/out, %d-clip chunks, copies overlapped)" % args.e2e_chunk},
            "gpu_launches": launches, "roofline": roofline, "stages": per_stage, "cpu_baseline": cpu,
            "small_batch_latency": small,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=4096)
    ap.add_argument("--cpu-clips", type=int, default=32)   # the CPU path is fastest per clip around this batch
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=1024)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
