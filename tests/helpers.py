"""Shared test plumbing: synthetic modules/inputs built from oracle/synth.py."""
import os

import numpy as np
import torch

from emotiongestures_b200 import BEAT, TED, Transformer
from emotiongestures_b200.generator import MemoryTransformer
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# "tedmem": TED geometry with the Prior_MemoryEncoder of Full_model/Models_memory.py (args.chunk = 4)
CFGS = {"ted": TED, "beat": BEAT, "tedmem": TED}
MEM_CHUNK = 4
_cache = {}


def model_and_sd(name: str, seed: int):
    """Mirror module (CPU) + the synthetic state_dict for (cfg, seed)."""
    key = (name, seed)
    if key not in _cache:
        m = (MemoryTransformer.from_config(CFGS[name], MEM_CHUNK) if name == "tedmem"
             else Transformer.from_config(CFGS[name])).eval()
        sd = synth.synth_state_dict(m.state_dict(), seed)
        m.load_state_dict(sd)
        _cache[key] = (m, sd)
    return _cache[key]


def inputs(cfg, n_clips, seed, with_emotion=False):
    spec = torch.from_numpy(synth.synth_spec(n_clips, cfg.n_mels, cfg.spec_w, seed))
    prior = torch.from_numpy(synth.synth_prior(n_clips, cfg.prior_frames, cfg.pose_dim, seed))
    emo = torch.from_numpy(synth.synth_emotion(n_clips, cfg.frames, cfg.d_model, seed)) if with_emotion else None
    return spec, prior, emo


def load_golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def rel_fro(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rel_max(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
