#!/usr/bin/env python
"""Benchmarks of the generator-inference hot path and of the networks / statistics either side of it
(BASELINE.json configs 2-5).  Every workload prints ONE JSON line (rank 0) carrying `value` (inputs resident in HBM),
`e2e` (pinned host buffers in / out, copies inside the timed region), `roofline`, `cpu_baseline`, `clocks`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ted|beat|emotion_net|cvae|fgd] [--impl reference]

  ted (default, BASELINE.json config 2 — the headline): one step = raw 16 kHz audio -> log-mel (PreEmphasis, STFT,
      mel, log, InstanceNorm) -> SE-ResNet audio encoder -> transformer generator -> 34-frame poses (+ the NCCL pose
      all-gather when N > 1, on a side stream under the next step).  `value` is weak scaling (4096 clips per GPU);
      the same line carries `strong_scaling` (BASELINE config 2 as written: 4096 clips TOTAL, 4096/N per GPU, the
      step replayed as one CUDA graph), the two collectives timed alone, and the FGD statistics leg.
  beat: the same path at the 60-frame BEAT geometry.
  emotion_net (config 3): log-mel + EmotionNet (four-stage SE-ResNet + FC chain) on synthetic speech.
  cvae (config 4): BEAT_CVAE MLP_Reconstruct.forward over 1 M rows (+ the CAVE v3 sampler).
  fgd (config 5): feature net over 100k generated clips -> [n | sum | gram] float64 statistics -> NCCL all-reduce ->
      mean / covariance -> Frechet distance.

`--impl reference` times the CPU oracle port of the same workload alone (the reference is pure PyTorch; its CPU path
is restated in oracle/ and pinned against the real reference by oracle/make_golden*.py).
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

UNIT = "clips/s"

# algorithmic work per clip
GEN = {
    "ted": dict(metric="clips_per_sec_34frame_generator_inference",
                workload="Full_model generator inference, TED-Emotion shape (2.27 s 16 kHz audio -> 34 poses), random-init",
                s3_flop=4.247e9, flop=4.79e9,
                stage_work={"S1_frontend": ("hbm", 180_908.0), "S2_stem": ("hbm", 609_280.0),
                            # layer 1 (32 channels) is HBM-bound: 5 map transfers of H*W*32*2 B per block (conv1: x in,
                            # y1 out; conv2: y1 + residual in, block output out), 3 blocks; layers 2-3 are tensor-bound
                            "S3_conv_layer1": ("hbm", 3 * 5 * 128 * 70 * 32 * 2.0),
                            "S3_conv_layer2": ("tensor", 1.248e9), "S3_conv_layer3": ("tensor", 1.963e9 + 0.045e9),
                            "S5_proj_gemm": ("tensor", 92e6), "S6_enc_dec": ("tensor", 442e6)}),
    "beat": dict(metric="clips_per_sec_60frame_generator_inference",
                 workload="Full_model generator inference, BEAT shape (4 s 16 kHz audio -> 60 poses), random-init",
                 s3_flop=7.483e9, flop=10.45e9,
                 stage_work={"S1_frontend": ("hbm", 319_488.0), "S2_stem": ("hbm", 1_079_296.0),
                             "S3_conv_layer1": ("hbm", 3 * 5 * 128 * 124 * 32 * 2.0),
                             "S3_conv_layer2": ("tensor", 2.210e9), "S3_conv_layer3": ("tensor", 3.381e9 + 0.137e9),
                             "S5_proj_gemm": ("tensor", 647e6), "S6_enc_dec": ("tensor", 2309e6)}),
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
    committed ncu pass over one TED step (profiles/r*_conv_dram.json, captured at `clips` clips and scaled linearly:
    every conv streams its maps once)."""
    for name in ("r2_conv_dram.json", "r1_conv_dram.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            break
    else:
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
        return self

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


class Dist:
    """torch.distributed plumbing of one rank (NCCL, one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa_node = None
        if self.world > 1:
            # one process per GPU: keep this rank's pinned buffers on the GPU's own NUMA node
            from emotiongestures_b200.sharding import bind_host_to_gpu_node
            self.numa_node = bind_host_to_gpu_node(self.local)
            dist.init_process_group("nccl", device_id=self.dev)

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def timed(self, fn, steps, warmup):
        """W untimed warm-up calls, then exactly K calls between a barrier + synchronize on both sides; CUDA events on
        the current stream, max over ranks.  Returns ms per call."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.sync()
        return self.max_ms(e0.elapsed_time(e1)) / steps

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# CPU arms: the oracle ports of the reference paths, on the host cores (bounded samples)
# ---------------------------------------------------------------------------------------------
def _cpu_threads():
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


def _time_host(fn, steps, warmup):
    for _ in range(warmup):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    return (time.perf_counter() - t0) / steps, out


def generator_weights(name):
    import torch
    from emotiongestures_b200 import BEAT, TED, Transformer, randomize_norm_stats_
    cfg = TED if name == "ted" else BEAT
    torch.manual_seed(0)                      # identical random-init weights on every rank and in the CPU arm
    gen = Transformer.from_config(cfg).eval()
    randomize_norm_stats_(gen, 1)
    return cfg, gen


def cpu_generator(name, n_clips, steps, warmup):
    """Oracle port of the path: log-mel in fp64 numpy (F1-F4b), generator forward in fp32 torch on all host cores.
    Returns clips/s, ms/step, cores, and (audio, prior, log-mel, poses) of the sample for the parity check."""
    import torch
    from oracle import generator as og
    from oracle import logmel as ol
    from oracle import synth
    cores = _cpu_threads()
    cfg, gen = generator_weights(name)
    sd = {k: v.detach() for k, v in gen.state_dict().items()}
    audio = synth.synth_audio(n_clips, cfg.n_audio, seed=1000)
    prior = torch.from_numpy(synth.synth_prior(n_clips, cfg.prior_frames, cfg.pose_dim, 1000))

    def step():
        spec = torch.from_numpy(ol.logmel(audio, cfg.spec_w, "log_in"))
        with torch.no_grad():
            return spec, og.generator_forward(sd, cfg, spec.float(), prior)[0]

    dt, (spec, poses) = _time_host(step, steps, warmup)
    return n_clips / dt, dt * 1e3, cores, (audio, prior, spec, poses)


def _aux_module(kind):
    """(mirror module with seeded random-init weights, its state_dict) for the small-network workloads."""
    import torch
    from emotiongestures_b200 import aux_models as mirrors
    from emotiongestures_b200.generator import randomize_norm_stats_
    torch.manual_seed(0)
    m = {"emotion_net": lambda: mirrors.EmotionNet(), "cvae": lambda: mirrors.MLP_Reconstruct(),
         "cvae3": lambda: mirrors.MLP_Reconstruct_v3(), "motion_ae": lambda: mirrors.MotionAE(126, 128),
         "fgd_mlp": lambda: mirrors.FGDNet()}[kind]().eval()
    randomize_norm_stats_(m, 1)
    return m, {k: v.detach() for k, v in m.state_dict().items()}


def line_base(args, metric, unit, value, ms_step, world, scaling, dtype, config):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config}


def reference_line(args, metric, unit, value, ms, cores, workload, sample, extra_cfg=None):
    cfg = {"workload": workload, "device": "host CPU"}
    cfg.update(extra_cfg or {})
    line = line_base(args, metric, unit, value, ms, args.gpus, "weak", "f32", cfg)
    line.update({"impl": "reference",
                 "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                 "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# generator workloads (ted / beat)
# ---------------------------------------------------------------------------------------------
def run_generator(args, name):
    import torch
    import torch.distributed as dist

    from emotiongestures_b200 import LOGMEL_LOG_IN, fgd

    spec_ = GEN[name]
    d = Dist()
    world, rank, dev = d.world, d.rank, d.dev
    cfg, gen = generator_weights(name)
    gen = gen.to(dev)
    gen.precision = args.precision
    eng = gen.engine(args.precision)
    B = args.clips_per_gpu or (4096 if name == "ted" else 2048)

    # ---- CPU baseline first (rank 0, N = 1): its sample doubles as the parity check that gates the timing ----
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_clips if name == "ted" else max(4, args.cpu_clips // 4)
        n_steps = 24 if name == "ted" else 8
        cpu_val, _, cores, (a_np, p_t, spec_ref, poses_ref) = cpu_generator(name, n_cpu, n_steps, 2)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} clips x {n_steps} steps (oracle port: log-mel fp64 numpy + generator fp32 torch, "
                         f"{cores} threads)"}
        spec_gpu = eng.logmel(torch.from_numpy(a_np).to(dev), LOGMEL_LOG_IN, True)
        poses_gpu = eng.generator_forward(spec_gpu, p_t.to(dev))[0]
        mel_err = float((spec_gpu.double().cpu() - spec_ref).abs().max())
        pose_err = float((poses_gpu.cpu().double() - poses_ref.double()).norm() / poses_ref.double().norm())
        tol = 2e-3 if args.precision == "tc" else 2e-5
        parity = {"logmel_max_abs": mel_err, "logmel_tol": 1e-4, "poses_rel_frobenius": pose_err, "poses_tol": tol,
                  "clips": n_cpu, "against": "oracle port on the same inputs and weights"}
        if not (mel_err <= 1e-4 and pose_err <= tol):
            raise SystemExit(f"parity check failed before timing: {parity}")

    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    audio = (0.1 * torch.randn(B, cfg.n_audio, generator=g, device=dev)).clamp_(-1, 1)
    prior = torch.randn(B, cfg.prior_frames, cfg.pose_dim, generator=g, device=dev)

    # pose all-gather (north star) as ONE all_gather_into_tensor on a side stream: the gather of step i runs under
    # the kernels of step i + 1; two output sets alternate so that a step never overwrites poses still in flight
    comm = torch.cuda.Stream(dev) if world > 1 else None
    outs = [tuple(torch.empty(s, device=dev) for s in ((B, cfg.frames, cfg.pose_dim), (B, cfg.frames, cfg.d_model),
                                                        (B, cfg.frames, cfg.d_model), (B, 8))) for _ in range(2)]
    gathered = [torch.empty(world * B, cfg.frames, cfg.pose_dim, device=dev) for _ in range(2)] if world > 1 else None
    done = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "used": [False, False]}

    # --gather peer (default): copy-engine pushes into every rank's buffer over NVLink (sharding.PeerGather) — no SMs
    # taken from the persistent compute kernels it overlaps; --gather nccl: one all_gather_into_tensor
    peer = None
    if world > 1 and args.gather == "peer":
        from emotiongestures_b200.sharding import PeerGather
        try:
            peer = PeerGather(B * cfg.frames * cfg.pose_dim, torch.float32, dev)
            gathered = [b.view(world * B, cfg.frames, cfg.pose_dim) for b in peer.bufs]
        except Exception as e:                                  # no CUDA IPC in this container: say so, use NCCL
            print(f"[bench] peer gather unavailable ({type(e).__name__}: {e}); using the NCCL all-gather", file=sys.stderr)
            peer = None
        if peer is None:
            gathered = [torch.empty(world * B, cfg.frames, cfg.pose_dim, device=dev) for _ in range(2)]

    def gather_async(poses, slot):
        main = torch.cuda.current_stream(dev)
        done[slot].record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(done[slot])
            if peer is not None:
                peer.gather(poses.reshape(-1), slot, comm)
            else:
                dist.all_gather_into_tensor(gathered[slot].view(-1), poses.reshape(-1))
            free[slot].record(comm)
        state["used"][slot] = True

    def step(a=audio, p=prior):
        slot = state["i"] & 1
        state["i"] += 1
        main = torch.cuda.current_stream(dev)
        if world > 1 and state["used"][slot]:
            main.wait_event(free[slot])
        spec = eng.logmel(a, LOGMEL_LOG_IN, True)
        poses = eng.generator_forward(spec, p, None, out=outs[slot] if a is audio else None)[0]
        if world > 1 and a is audio:
            gather_async(poses, slot)
        return poses

    def drain():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm)

    # ---- device-resident throughput (weak scaling: B clips per GPU) ----
    for _ in range(args.warmup):
        step()
    drain()
    d.sync()
    launches0 = eng.launch_count
    eng.profile_enable(200 * args.steps)
    sampler = ClockSampler(d.local).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.sync()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    d.sync()
    clocks = sampler.stop()
    stages = eng.profile_read()
    eng.profile_enable(0)
    launches = eng.launch_count - launches0
    ms_step = d.max_ms(e0.elapsed_time(e1)) / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end: pinned host buffers in, poses out to the host, copies inside the timed region ----
    h_audio = audio.cpu().pin_memory()
    h_prior = prior.cpu().pin_memory()
    h_poses = torch.empty(B, cfg.frames, cfg.pose_dim).pin_memory()
    d_poses = [torch.empty(B, cfg.frames, cfg.pose_dim, device=dev) for _ in range(2)] if world > 1 else [None, None]

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        if world > 1 and state["used"][slot]:
            torch.cuda.current_stream(dev).wait_event(free[slot])
        # public API: pinned host in -> pinned host out, chunked so PCIe copies overlap the kernels
        eng.infer_host(h_audio, h_prior, h_poses, chunk=args.e2e_chunk, mode=LOGMEL_LOG_IN, preemph=True,
                       poses_dev=d_poses[slot], join=False, graph=args.e2e_graph)
        if world > 1:
            gather_async(d_poses[slot], slot)

    e2e_steps = max(3, args.steps)
    for _ in range(3):
        e2e_step()
    eng.host_join()
    drain()
    d.sync()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    eng.host_join()          # every step's poses are in host memory before the closing event
    drain()
    e1.record()
    d.sync()
    e2e_value = world * B / (d.max_ms(e0.elapsed_time(e1)) / e2e_steps * 1e-3)
    h2d = h_audio.numel() * 4 + h_prior.numel() * 4
    d2h = h_poses.numel() * 4

    # the same loop fed 16-bit PCM (what a wav file holds; widened to float on the device): half the host->device bytes.
    # Reported next to `e2e`, not instead of it — the headline keeps the reference's float32 input format.
    h_audio_f32 = h_audio
    h_audio = (h_audio_f32 * 32767.0).round().to(torch.int16).pin_memory()
    for _ in range(3):
        e2e_step()
    eng.host_join()
    drain()
    d.sync()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    eng.host_join()
    drain()
    e1.record()
    d.sync()
    e2e_pcm16 = {"value": world * B / (d.max_ms(e0.elapsed_time(e1)) / e2e_steps * 1e-3), "unit": UNIT,
                 "h2d_bytes_per_step": h_audio.numel() * 2 + h_prior.numel() * 4, "d2h_bytes_per_step": d2h,
                 "input": "int16 PCM audio (Engine.infer_host widens it on the device, egx_audio_pcm16_to_f32)"}
    h_audio = h_audio_f32

    # ---- strong scaling (BASELINE.json config 2 as written): `total` clips over all GPUs, one CUDA-graph replay per step ----
    strong = None
    total = args.strong_total
    if total and total % world == 0 and total // world <= B and name == "ted":
        bs = total // world
        path = eng.capture(bs, LOGMEL_LOG_IN, True)
        path.audio.copy_(audio[:bs]); path.prior.copy_(prior[:bs])
        g_out = torch.empty(world * bs, cfg.frames, cfg.pose_dim, device=dev) if world > 1 else None

        def strong_step():
            path.graph.replay()
            if world > 1:
                dist.all_gather_into_tensor(g_out.view(-1), path.out[0].view(-1))

        ms = d.timed(strong_step, max(args.steps, 20), 3)
        strong = {"total_clips": total, "clips_per_gpu": bs, "ms_per_step": ms, "value": total / (ms * 1e-3), "unit": UNIT,
                  "step": "one CUDA-graph replay (log-mel + forward) + all_gather_into_tensor of the poses"}
        if world > 1:
            # N-GPU poses == 1-GPU poses bit for bit: every rank recomputes its neighbour's shard from the gathered inputs
            chk_a = torch.empty(world * bs, cfg.n_audio, device=dev)
            chk_p = torch.empty(world * bs, cfg.prior_frames, cfg.pose_dim, device=dev)
            dist.all_gather_into_tensor(chk_a.view(-1), audio[:bs].reshape(-1))
            dist.all_gather_into_tensor(chk_p.view(-1), prior[:bs].reshape(-1))
            other = (rank + 1) % world
            mine = eng.generator_forward(eng.logmel(chk_a[other * bs:(other + 1) * bs], LOGMEL_LOG_IN, True),
                                         chk_p[other * bs:(other + 1) * bs])[0]
            flag = torch.tensor([int(torch.equal(mine, g_out[other * bs:(other + 1) * bs]))], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            strong["gathered_poses_bit_identical_to_local_recompute"] = bool(flag.item())
            if not flag.item():
                raise SystemExit("sharded poses differ from a local recomputation of the same clips")
            del chk_a, chk_p
        del path

    # ---- the two collectives alone, and the FGD statistics leg (feature net -> [n|sum|gram] f64 -> all-reduce) ----
    collectives = fgd_leg = None
    if name == "ted":
        mae, _ = _aux_module("motion_ae")
        mae = mae.to(dev)
        poses = step()
        drain()
        acc = fgd.new_accumulator(128, dev)
        shift = torch.zeros(128, dtype=torch.float64, device=dev)

        def fgd_step():
            eng.fgd_accumulate(mae(poses)[1], acc, shift)

        ms_acc = d.timed(fgd_step, args.steps, 3)
        ms_ar = d.timed(lambda: fgd.all_reduce_stats(acc), 20, 3) if world > 1 else 0.0
        acc.zero_()
        fgd_step()
        fgd.all_reduce_stats(acc)
        mu, sigma = fgd.finalize_stats(acc, 128, shift)
        fgd_leg = {"feature_net": "MotionAE.encoder (34x126 -> 128)", "clips": world * B, "features_plus_accumulate_ms": ms_acc,
                   "all_reduce_ms": ms_ar, "all_reduce_bytes": acc.numel() * 8, "n": float(acc[0].item()),
                   "trace_sigma": float(sigma.trace())}
        if world > 1:
            t = torch.tensor([float(sigma.trace())], device=dev, dtype=torch.float64)
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            fgd_leg["identical_on_all_ranks"] = bool((lo == hi).item())
            buf = torch.empty(world * B, cfg.frames, cfg.pose_dim, device=dev)
            ms_ag = d.timed(lambda: dist.all_gather_into_tensor(buf.view(-1), poses.reshape(-1)), 20, 3)
            ms_pg = None
            if peer is not None:
                cur = torch.cuda.current_stream(dev)
                ms_pg = d.timed(lambda: peer.gather(poses.reshape(-1), 0, cur), 20, 3)
                same = torch.tensor([int(torch.equal(peer.bufs[0], buf.view(-1)))], device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                if not same.item():
                    raise SystemExit("peer gather and NCCL all-gather disagree")
            collectives = {"pose_gather": "peer copies on the copy engines + 4-byte all-reduce" if peer is not None else "nccl all_gather_into_tensor",
                           "pose_peer_gather_ms": ms_pg,
                           "pose_all_gather_ms": ms_ag, "pose_bytes_per_rank": poses.numel() * 4,
                           "pose_bytes_received_per_rank": poses.numel() * 4 * (world - 1),
                           "fgd_all_reduce_ms": ms_ar, "fgd_bytes": acc.numel() * 8,
                           "overlap": "the gather of step i runs on a side stream under the kernels of step i + 1"}
            del buf

    # ---- small batches are launch-bound: eager launches vs one CUDA-graph replay (Engine.capture), 1 and 8 clips ----
    small = None
    if rank == 0 and world == 1:
        small = {}
        for nb in (1, 8):
            a, p_ = audio[:nb].contiguous(), prior[:nb].contiguous()
            path = eng.capture(nb, LOGMEL_LOG_IN, True)
            res = {}
            for nm, fn in (("eager", lambda: step(a, p_)), ("graph", lambda: path(a, p_))):
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(100):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[nm + "_ms"] = e0.elapsed_time(e1) / 100
            small["clips_%d" % nb] = res

    if rank == 0:
        peaks = load_peaks()
        per_stage = {}
        for nm, (ms, cnt) in stages.items():
            ms_per_step = ms / args.steps
            ent = {"ms_per_step": ms_per_step, "launches_per_step": cnt / args.steps}
            if nm in spec_["stage_work"] and ms_per_step > 0:
                bound, work = spec_["stage_work"][nm]
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
            ach = spec_["s3_flop"] * B / (ms3 * 1e-3) / 1e12
            per_stage["S3_trunk_conv"] = {"ms_per_step": ms3, "launches_per_step": n3, "bound": "tensor", "achieved": ach,
                                          "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"]}
            s4 = per_stage.get("S4_se")
            if s4:
                ms34 = ms3 + s4["ms_per_step"]
                ach = spec_["s3_flop"] * B / (ms34 * 1e-3) / 1e12
                per_stage["S3+S4_trunk"] = {"ms_per_step": ms34, "launches_per_step": n3 + s4["launches_per_step"],
                                            "bound": "tensor", "achieved": ach, "unit": "TFLOP/s",
                                            "frac": ach / peaks["tf_sustained"]}
        whole = spec_["flop"] * B / (ms_step * 1e-3) / 1e12
        per_stage["whole_step"] = {"ms_per_step": ms_step, "bound": "tensor", "achieved": whole, "unit": "TFLOP/s",
                                   "frac": whole / peaks["tf_sustained"]}
        # roofline of the dominant kernel family: the trunk layer with the largest share of the step, against the
        # roofline that bounds it (layer 1: HBM; layers 2-3: tensor pipe)
        dom = max(layers, key=lambda k: layers[k]["ms_per_step"]) if layers else None
        roofline = None
        if dom:
            st = layers[dom]
            n_l = max(1.0, st["launches_per_step"])
            bound, work = spec_["stage_work"][dom]
            lay = dom.replace("S3_conv_", "")
            roofline = {
                "bound": bound,
                "kernel": "trunk convolutions of %s (%d launches per step: %s)" % (
                    lay, n_l, {"layer1": "conv_tc_kernel<32,32> SE-sum and gated-residual flavours",
                               "layer2": "conv_tc_kernel<64,64> + the stride-2 / 1x1 convs of its first block",
                               "layer3": "conv128_tc_kernel + the stride-2 / 1x1 convs of its first block + final conv"}.get(lay, lay)),
                "achieved": st["achieved"], "peak": peaks["hbm_gbs"] if bound == "hbm" else peaks["tf_sustained"],
                "unit": st["unit"], "frac": st["frac"], "traffic": conv_traffic_per_launch(B, lay) if name == "ted" else None,
                "peak_source": peaks["src"] + (" (copy bandwidth)" if bound == "hbm" else
                                               " (sustained bf16/fp16 dense, kernel timed inside a long step)"),
                ("bytes_per_launch" if bound == "hbm" else "flop_per_launch"): work * B / n_l,
                "ms_per_launch": st["ms_per_step"] / n_l,
                "all_trunk_convs_vs_tensor_peak": per_stage.get("S3_trunk_conv", {}).get("frac"),
            }
        line = line_base(args, spec_["metric"], UNIT, value, ms_step, world, "weak",
                         "f16" if args.precision == "tc" else "f32",
                         {"workload": spec_["workload"], "clips_per_gpu": B, "global_batch": B * world,
                          "precision": args.precision, "logmel": "preemph+log+InstanceNorm (F4b)",
                          "parallelism": f"dp{world} (clip-sharded; pose gather on a side stream: " + ("peer copies over NVLink" if peer is not None else "NCCL all_gather_into_tensor") + ")",
                          "l2": "inputs larger than L2 (%.0f MB audio per step)" % (audio.numel() * 4 / 1e6),
                          "host_numa_node_rank0": d.numa_node})
        line.update({
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "Engine.infer_host(join=False) per step + host_join() before the closing event: pinned host in/out, %d-clip "
                           "chunks, every copy inside the timed region, overlapped with the kernels of the neighbouring chunk / step"
                           % args.e2e_chunk},
            "gpu_launches": launches, "roofline": roofline, "stages": per_stage, "cpu_baseline": cpu, "parity": parity,
            "e2e_pcm16": e2e_pcm16, "strong_scaling": strong, "collectives": collectives, "fgd": fgd_leg, "small_batch_latency": small,
            "clocks": clocks})
        print(json.dumps(line), flush=True)
    d.close()


def reference_generator(args, name):
    n = args.cpu_clips if name == "ted" else max(4, args.cpu_clips // 4)
    value, ms, cores, _ = cpu_generator(name, n, args.steps, args.warmup)
    reference_line(args, GEN[name]["metric"], UNIT, value, ms, cores, GEN[name]["workload"],
                   f"{n} clips per step (log-mel fp64 numpy + generator fp32 torch), {cores} torch threads", {"clips_per_step": n})


# ---------------------------------------------------------------------------------------------
# config 3: EmotionNet (model/audio_emotion_classifer.py:38-49) on synthetic speech
# ---------------------------------------------------------------------------------------------
EMO_METRIC = "clips_per_sec_emotion_net_inference"
EMO_WORKLOAD = "EmotionNet audio emotion classifier: 4 s 16 kHz audio -> log-mel (128x124) -> four-stage SE-ResNet + FC chain -> 8 logits"
EMO_FLOP = 9.59e9            # 9.03 G trunk + 0.56 G FC per clip


def cpu_emotion_net(n_clips, steps, warmup):
    import torch
    from oracle import aux_models as oa
    from oracle import logmel as ol
    from oracle import synth
    cores = _cpu_threads()
    _, sd = _aux_module("emotion_net")
    audio = synth.synth_audio(n_clips, 64000, seed=1000)

    def step():
        spec = torch.from_numpy(ol.logmel(audio, 124, "log_in")).float()
        with torch.no_grad():
            return spec, oa.emotion_net(sd, spec)

    dt, (spec, logits) = _time_host(step, steps, warmup)
    return n_clips / dt, dt * 1e3, cores, (audio, spec, logits)


def run_emotion_net(args):
    import torch
    from emotiongestures_b200 import BEAT, LOGMEL_LOG_IN
    from emotiongestures_b200.engine import Engine
    d = Dist()
    dev, world = d.dev, d.world
    B = args.clips_per_gpu or 2048
    net, _ = _aux_module("emotion_net")
    net = net.to(dev)
    eng = net._engine()
    fe = Engine(BEAT, dev)                  # the front-end handle (log-mel tables only)
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        val, _, cores, (a_np, spec_ref, logits_ref) = cpu_emotion_net(8, 3, 1)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"8 clips x 3 steps (oracle port: log-mel fp64 numpy + EmotionNet fp32 torch, {cores} threads)"}
        got = net(fe.logmel(torch.from_numpy(a_np).to(dev), LOGMEL_LOG_IN, True, n_cols=124)).cpu()
        err = float((got - logits_ref).abs().max() / logits_ref.abs().max())
        parity = {"logits_rel_max": err, "tol": 2e-3, "clips": 8}
        if err > 2e-3:
            raise SystemExit(f"parity check failed before timing: {parity}")
    g = torch.Generator(device=dev).manual_seed(1000 + d.rank)
    audio = (0.1 * torch.randn(B, 64000, generator=g, device=dev)).clamp_(-1, 1)

    def step():
        return net(fe.logmel(audio, LOGMEL_LOG_IN, True, n_cols=124))

    l0 = eng.launch_count + fe.launch_count
    sampler = ClockSampler(d.local).start()
    ms = d.timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = (eng.launch_count + fe.launch_count - l0) * args.steps // (args.steps + args.warmup)
    h_audio = audio.cpu().pin_memory()
    h_out = torch.empty(B, 8).pin_memory()

    def e2e():
        c = args.e2e_chunk
        for lo in range(0, B, c):
            a = h_audio[lo:lo + c].to(dev, non_blocking=True)
            h_out[lo:lo + c].copy_(net(fe.logmel(a, LOGMEL_LOG_IN, True, n_cols=124)), non_blocking=True)

    ms_e2e = d.timed(e2e, max(3, args.steps // 2), 2)
    if d.rank == 0:
        peaks = load_peaks()
        ach = EMO_FLOP * B / (ms * 1e-3) / 1e12
        line = line_base(args, EMO_METRIC, UNIT, world * B / (ms * 1e-3), ms, world, "weak", "f16",
                         {"workload": EMO_WORKLOAD, "clips_per_gpu": B, "l2": "inputs larger than L2 (%.0f MB audio)" % (audio.numel() * 4 / 1e6)})
        line.update({"e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h_audio.numel() * 4,
                             "d2h_bytes_per_step": h_out.numel() * 4, "api": "EmotionNet.forward on chunks of pinned host audio"},
                     "gpu_launches": launches,
                     "roofline": {"bound": "tensor", "kernel": "whole step (trunk convolutions on conv_tc / conv128_tc dominate: 9.03 of 9.59 GFLOP per clip)",
                                  "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                                  "traffic": None, "flop_per_step": EMO_FLOP * B, "peak_source": peaks["src"]},
                     "cpu_baseline": cpu, "parity": parity, "clocks": clocks})
        print(json.dumps(line), flush=True)
    d.close()


# ---------------------------------------------------------------------------------------------
# config 4: BEAT_CVAE MLP_Reconstruct.forward (Full_model/BEAT_CVAE.py:98-114) + CAVE v3 sampler
# ---------------------------------------------------------------------------------------------
CVAE_METRIC = "rows_per_sec_beat_cvae_forward"
CVAE_WORKLOAD = "Full_model/BEAT_CVAE.MLP_Reconstruct.forward on (N,90) hand poses + (N,90) condition, noise passed in"
CVAE_BYTES = (90 + 90 + 32 + 90 + 32 + 32) * 4.0          # per row: x, y, eps in; out, mu, logvar out


def cpu_cvae(n, steps, warmup):
    import torch
    from oracle import aux_models as oa
    cores = _cpu_threads()
    _, sd = _aux_module("cvae")
    g = torch.Generator().manual_seed(1000)
    x, y, eps = (torch.randn(n, k, generator=g) for k in (90, 90, 32))

    def step():
        with torch.no_grad():
            return oa.cvae_forward(sd, x, y, eps)

    dt, out = _time_host(step, steps, warmup)
    return n / dt, dt * 1e3, cores, (x, y, eps, out)


def run_cvae(args):
    import torch
    d = Dist()
    dev, world = d.dev, d.world
    N = args.rows or 1_000_000
    net, _ = _aux_module("cvae")
    net = net.to(dev)
    eng = net._engine()
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        val, _, cores, (x_c, y_c, e_c, ref) = cpu_cvae(200_000, 10, 2)
        cpu = {"value": val, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"200000 rows x 10 steps (oracle port, fp32 torch, {cores} threads)"}
        got = net(x_c.to(dev), y_c.to(dev), eps=e_c.to(dev))
        err = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(got, ref))
        parity = {"rel_max": err, "tol": 2e-5, "rows": 200_000}
        if err > 2e-5:
            raise SystemExit(f"parity check failed before timing: {parity}")
    g = torch.Generator(device=dev).manual_seed(1000 + d.rank)
    x, y, eps = (torch.randn(N, k, generator=g, device=dev) for k in (90, 90, 32))
    l0 = eng.launch_count
    sampler = ClockSampler(d.local).start()
    ms = d.timed(lambda: net(x, y, eps=eps), args.steps, args.warmup)
    clocks = sampler.stop()
    launches = (eng.launch_count - l0) * args.steps // (args.steps + args.warmup)
    hx, hy, he = (t.cpu().pin_memory() for t in (x, y, eps))
    ho = [torch.empty(N, k).pin_memory() for k in (90, 32, 32)]

    def e2e():
        c = 1 << 18
        for lo in range(0, N, c):
            o = net(hx[lo:lo + c].to(dev, non_blocking=True), hy[lo:lo + c].to(dev, non_blocking=True),
                    eps=he[lo:lo + c].to(dev, non_blocking=True))
            for h_, t in zip(ho, o):
                h_[lo:lo + c].copy_(t, non_blocking=True)

    ms_e2e = d.timed(e2e, max(3, args.steps // 2), 2)
    # the CAVE v3 sampler (CAVE/BEAT_CVAE.py:427-447): (n,8) one-hot + (n,32) noise -> (n,60,512)
    v3, _ = _aux_module("cvae3")
    v3 = v3.to(dev)
    nb = 4096
    yy = torch.nn.functional.one_hot(torch.randint(0, 8, (nb,), generator=torch.Generator().manual_seed(3)), 8).float().to(dev)
    zz = torch.randn(nb, 32, generator=g, device=dev)
    ms_v3 = d.timed(lambda: v3.sample(yy, z=zz), args.steps, args.warmup)
    if d.rank == 0:
        peaks = load_peaks()
        ach = CVAE_BYTES * N / (ms * 1e-3) / 1e9
        line = line_base(args, CVAE_METRIC, "rows/s", world * N / (ms * 1e-3), ms, world, "weak", "f32",
                         {"workload": CVAE_WORKLOAD, "rows_per_gpu": N, "l2": "inputs larger than L2 (%.0f MB per step)" % (CVAE_BYTES * N / 1e6)})
        line.update({"e2e": {"value": world * N / (ms_e2e * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": N * 212 * 4,
                             "d2h_bytes_per_step": N * 154 * 4, "api": "MLP_Reconstruct.forward on 262144-row chunks of pinned host rows"},
                     "gpu_launches": launches,
                     "roofline": {"bound": "hbm", "kernel": "cvae_mlp_kernel", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": ach / peaks["hbm_gbs"], "traffic": None, "bytes_per_launch": CVAE_BYTES * N,
                                  "peak_source": peaks["src"] + " (copy bandwidth)"},
                     "cvae3_sampler": {"clips": nb, "ms": ms_v3, "clips_per_s": nb / (ms_v3 * 1e-3),
                                       "write_GBs": nb * 60 * 512 * 4 / (ms_v3 * 1e-3) / 1e9},
                     "cpu_baseline": cpu, "parity": parity, "clocks": clocks})
        print(json.dumps(line), flush=True)
    d.close()


# ---------------------------------------------------------------------------------------------
# config 5: FGD evaluation over 100k generated clips, statistics all-reduced over the GPUs
# ---------------------------------------------------------------------------------------------
FGD_METRIC = "clips_per_sec_fgd_evaluation"


def fgd_case(variant):
    """(feature net kind, frames, pose_dim, feature dim D, rows per clip)"""
    return {"ted": ("motion_ae", 34, 126, 128, 1), "beat": ("fgd_mlp", 60, 282, 512, 60)}[variant]


def cpu_fgd(variant, n_clips, steps):
    """The reference's arithmetic (test_emotion_gesture_diversity_iterative.py:226-232,251-254): feature rows to a
    float64 numpy array on the host, np.mean + np.cov(rowvar=False)."""
    import numpy as np
    import torch
    from oracle import aux_models as oa
    cores = _cpu_threads()
    kind, frames, pdim, D, rpc = fgd_case(variant)
    _, sd = _aux_module(kind)
    poses = torch.randn(n_clips, frames, pdim, generator=torch.Generator().manual_seed(1000)) * 0.3

    def step():
        with torch.no_grad():
            f = (oa.pose_encoder(sd, poses, "encoder.") if kind == "motion_ae" else oa.fgd_latent(sd, poses).reshape(-1, D))
        a = f.double().numpy()
        return np.mean(a, axis=0), np.cov(a, rowvar=False)

    dt, (mu, sigma) = _time_host(step, steps, 1)
    return n_clips / dt, dt * 1e3, cores, (poses, mu, sigma)


def run_fgd(args):
    import numpy as np
    import torch
    from emotiongestures_b200 import fgd
    d = Dist()
    dev, world, rank = d.dev, d.world, d.rank
    kind, frames, pdim, D, rpc = fgd_case(args.fgd_variant)
    total = args.fgd_clips
    n_local = total // world
    net, _ = _aux_module(kind)
    net = net.to(dev)
    eng = net._engine() if kind == "fgd_mlp" else net.encoder._engine()
    feats_of = (lambda p: net(p)[1].reshape(-1, D))
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 20000 if args.fgd_variant == "ted" else 1000
        val, _, cores, (p_c, mu_ref, sig_ref) = cpu_fgd(args.fgd_variant, n_cpu, 3)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} clips x 3 steps (oracle feature net fp32 torch + np.mean / np.cov float64, {cores} threads)"}
        f = feats_of(p_c.to(dev))
        acc = fgd.new_accumulator(D, dev)
        sh = f[:256].double().mean(0)
        eng.fgd_accumulate(f, acc, sh)
        mu, sig = fgd.finalize_stats(acc, D, sh)
        # statistics of the SAME feature rows: float64 accumulation vs numpy, rtol 1e-9
        f64 = f.double().cpu().numpy()
        np.testing.assert_allclose(mu, f64.mean(0), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(sig, np.cov(f64, rowvar=False), rtol=1e-9, atol=1e-12)
        parity = {"stats_vs_numpy_f64_on_the_same_rows": "rtol 1e-9 ok",
                  "mu_vs_cpu_feature_net_max_abs": float(np.abs(mu - mu_ref).max()),
                  "sigma_vs_cpu_feature_net_rel_fro": float(np.linalg.norm(sig - sig_ref) / np.linalg.norm(sig_ref)), "clips": n_cpu}
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    chunk = min(n_local, args.fgd_chunk)
    poses = [torch.randn(min(chunk, n_local - lo), frames, pdim, generator=g, device=dev) * 0.3 for lo in range(0, n_local, chunk)]
    real = torch.randn(4096, frames, pdim, generator=g, device=dev) * 0.35 + 0.02
    shift = feats_of(poses[0][:256]).double().mean(0)
    if world > 1:
        d.dist.broadcast(shift, src=0)                   # every rank centres on the same provisional mean
    acc = fgd.new_accumulator(D, dev)
    acc_real = fgd.new_accumulator(D, dev)
    eng.fgd_accumulate(feats_of(real), acc_real, shift)
    fgd.all_reduce_stats(acc_real)
    out = {}

    def step():
        acc.zero_()
        for p in poses:
            eng.fgd_accumulate(feats_of(p), acc, shift)
        fgd.all_reduce_stats(acc)
        out["fgd"] = fgd.frechet_distance_device(*fgd.finalize_stats_device(acc, D, shift),
                                                 *fgd.finalize_stats_device(acc_real, D, shift))

    l0 = eng.launch_count
    eng.profile_enable(64 * (args.steps + args.warmup) * len(poses))
    sampler = ClockSampler(d.local).start()
    ms = d.timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    st = eng.profile_read()
    eng.profile_enable(0)
    launches = (eng.launch_count - l0) * args.steps // (args.steps + args.warmup)
    ms_stats = st.get("S8_fgd", (0.0, 0))[0] / (args.steps + args.warmup)
    ms_ar = d.timed(lambda: fgd.all_reduce_stats(acc), 20, 3) if world > 1 else 0.0
    h_poses = [p.cpu().pin_memory() for p in poses]

    def e2e():
        acc.zero_()
        for hp in h_poses:
            eng.fgd_accumulate(feats_of(hp.to(dev, non_blocking=True)), acc, shift)
        fgd.all_reduce_stats(acc)
        out["fgd"] = fgd.frechet_distance_device(*fgd.finalize_stats_device(acc, D, shift),
                                                 *fgd.finalize_stats_device(acc_real, D, shift))

    ms_e2e = d.timed(e2e, max(2, args.steps // 2), 1)
    if rank == 0:
        peaks = load_peaks()
        rows = n_local * rpc
        fp64_peak = 37.0                           # B200 fp64 TFLOP/s (datasheet); MEASURED_PEAKS.json carries no fp64 line
        ach = 2.0 * D * D * rows / (max(ms_stats, 1e-6) * 1e-3) / 1e12
        hbm = 4.0 * D * rows / (max(ms_stats, 1e-6) * 1e-3) / 1e9
        tensor_bound = D >= 256
        line = line_base(args, FGD_METRIC, UNIT, world * n_local / (ms * 1e-3), ms, world, "strong", "f64",
                         {"workload": f"FGD evaluation over {total} generated clips ({args.fgd_variant}: {kind} features, D = {D}, "
                                      f"{rpc} row(s) per clip): feature net -> [n|sum|gram] f64 -> all-reduce -> mean/cov -> Frechet distance",
                          "clips_total": total, "clips_per_gpu": n_local, "rows_per_gpu": rows,
                          "l2": "inputs larger than L2 (%.0f MB poses per step)" % (n_local * frames * pdim * 4 / 1e6)})
        line.update({"e2e": {"value": world * n_local / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_local * frames * pdim * 4,
                             "d2h_bytes_per_step": 8, "api": "pinned host poses -> feature net -> fgd_accumulate -> all_reduce -> frechet_distance_device"},
                     "gpu_launches": launches,
                     "roofline": ({"bound": "tensor", "kernel": "fgd gram kernel (float64 syrk, 2*D*D flop per row)", "achieved": ach, "peak": fp64_peak,
                                   "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None, "flop_per_step": 2.0 * D * D * rows,
                                   "peak_source": "fallback: B200 datasheet fp64 (MEASURED_PEAKS.json carries no fp64 figure)"} if tensor_bound else
                                  {"bound": "hbm", "kernel": "fgd gram kernel (float64 syrk; D = %d rows are read once)" % D, "achieved": hbm,
                                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"], "traffic": None,
                                   "bytes_per_step": 4.0 * D * rows, "peak_source": peaks["src"] + " (copy bandwidth)"}),
                     "statistics_ms_per_step": ms_stats, "all_reduce_ms": ms_ar, "all_reduce_bytes": acc.numel() * 8,
                     "fgd_value": out.get("fgd"), "cpu_baseline": cpu, "parity": parity, "clocks": clocks})
        print(json.dumps(line), flush=True)
    d.close()


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = args.workload
    if w in GEN:
        return reference_generator(args, w)
    if w == "emotion_net":
        val, ms, cores, _ = cpu_emotion_net(8, args.steps, args.warmup)
        return reference_line(args, EMO_METRIC, UNIT, val, ms, cores, EMO_WORKLOAD,
                              f"8 clips per step (log-mel fp64 numpy + EmotionNet fp32 torch), {cores} torch threads")
    if w == "cvae":
        val, ms, cores, _ = cpu_cvae(200_000, args.steps, args.warmup)
        return reference_line(args, CVAE_METRIC, "rows/s", val, ms, cores, CVAE_WORKLOAD,
                              f"200000 rows per step (fp32 torch), {cores} torch threads")
    n = 20000 if args.fgd_variant == "ted" else 1000
    val, ms, cores, _ = cpu_fgd(args.fgd_variant, n, args.steps)
    reference_line(args, FGD_METRIC, UNIT, val, ms, cores, f"FGD evaluation ({args.fgd_variant})",
                   f"{n} clips per step (feature net fp32 torch + np.mean / np.cov float64), {cores} torch threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="ted", choices=["ted", "beat", "emotion_net", "cvae", "fgd"])
    ap.add_argument("--clips-per-gpu", type=int, default=0)   # 0: the workload's default (ted 4096, beat / emotion_net 2048)
    ap.add_argument("--strong-total", type=int, default=4096)  # BASELINE.json config 2: 4096 clips over all GPUs (0 = skip)
    ap.add_argument("--cpu-clips", type=int, default=32)   # the CPU path is fastest per clip around this batch
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # clips per pipeline chunk of Engine.infer_host.  Measured on one box against the device-resident value: 1024 -> 0.93,
    # 2048 -> 0.965, 4096 (= the whole step: step k+1's host->device copy runs under step k's kernels, step k's
    # device->host copy under step k+1's) -> 0.995
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--e2e-chunk", type=int, default=4096)
    # one CUDA-graph replay per staging slot inside infer_host instead of ~110 launches: measured equal at N = 1 and N = 8
    # (884k vs 876k end to end on eight GPUs: the gap to the device-resident value there is the PCIe fabric, not launches)
    ap.add_argument("--e2e-graph", action="store_true")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--fgd-variant", default="ted", choices=["ted", "beat"])
    ap.add_argument("--fgd-clips", type=int, default=100_000)
    ap.add_argument("--fgd-chunk", type=int, default=8192)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    {"ted": lambda: run_generator(args, "ted"), "beat": lambda: run_generator(args, "beat"),
     "emotion_net": lambda: run_emotion_net(args), "cvae": lambda: run_cvae(args), "fgd": lambda: run_fgd(args)}[args.workload]()


if __name__ == "__main__":
    main()
