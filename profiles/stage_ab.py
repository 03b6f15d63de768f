"""Same-box A/B of per-stage times under attribution switches (attribution build only: EGX_* variables are read by
`env_switch` at every launch).  Nothing here is a benchmark number; it ranks configurations on one box.
    python -m emotiongestures_b200.build --attribution
    python profiles/stage_ab.py 4096 "" "EGX_ATTN_CONTIG=0" "EGX_ATTN_CONTIG=0 EGX_ATTN_PROMO=128" ...
Every configuration is timed `ROUNDS` times, interleaved, 6 steps each (CUDA events around every launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emotiongestures_b200 import LOGMEL_LOG_IN, TED, Transformer, randomize_norm_stats_

B = int(sys.argv[1])
configs = sys.argv[2:] or [""]
ROUNDS, STEPS = 3, 6
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = Transformer.from_config(TED).eval()
randomize_norm_stats_(gen, 1)
gen = gen.to(dev)
eng = gen.engine("tc")
g = torch.Generator(device=dev).manual_seed(1000)
audio = (0.1 * torch.randn(B, TED.n_audio, generator=g, device=dev)).clamp_(-1, 1)
prior = torch.randn(B, TED.prior_frames, TED.pose_dim, generator=g, device=dev)


def step():
    spec = eng.logmel(audio, LOGMEL_LOG_IN, True)
    return eng.generator_forward(spec, prior, None)[0]


def set_env(cfg):
    for k in [k for k in os.environ if k.startswith("EGX_") and k != "EGX_ATTRIBUTION"]:
        del os.environ[k]
    for kv in cfg.split():
        k, v = kv.split("=")
        os.environ[k] = v


ref = None
for rnd in range(ROUNDS):
    for cfg in configs:
        set_env(cfg)
        for _ in range(2):
            out = step()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        eng.profile_enable(200 * STEPS)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(STEPS):
            step()
        e1.record()
        torch.cuda.synchronize()
        st = eng.profile_read()
        eng.profile_enable(0)
        line = " ".join(f"{k.split('_')[0]}={v[0] / STEPS:.3f}" for k, v in st.items())
        print(f"[{rnd}] {cfg or 'default':40s} step(with events)={e0.elapsed_time(e1) / STEPS:7.3f} ms  same={same}  {line}", flush=True)
