"""`install(generator)`: put libegx behind a LIVE reference generator.

`generator` is the reference's own Full_model.Models.Transformer (or
Models_memory.Transformer, possibly unwrapped from nn.DataParallel), already
`.eval()`-ed and on a CUDA device, exactly as
test_emotion_gesture_diversity_iterative.py:135-145 leaves it.  Its `forward` is
swapped for one that keeps the signature and the 5-tuple but routes the pose path
through the C ABI; the text encoder (dead w.r.t. poses) keeps running as the
module's own PyTorch sub-module.  In training mode the module's original forward is
called untouched (the library is inference-only).
"""
from __future__ import annotations

import types

import torch

from .config import GeneratorConfig
from .engine import Engine


def config_from_module(gen) -> GeneratorConfig:
    sd = gen.state_dict()
    d_model = sd["emotion_proj.0.weight"].shape[0]
    if "prior_seq_encoder.pred_conv.0.weight" in sd:        # Models_memory.Transformer: conv maps p -> F - p frames
        n_pred, prior_frames, _ = sd["prior_seq_encoder.pred_conv.0.weight"].shape
        frames = n_pred + prior_frames
    else:
        frames, prior_frames, _ = sd["prior_seq_encoder.conv1.weight"].shape
    pose_dim = sd["post_projector.6.weight"].shape[0]
    fc1_in = sd["audio_encoder.fc1.weight"].shape[1]
    n_layers = len(gen.encoder.layer_stack)
    mha = gen.encoder.layer_stack[0].slf_attn
    spec_w = None
    for w in range(8, 512):                      # invert W -> ceil(ceil(W/2)/2) * 32
        if 32 * (((w + 1) // 2 + 1) // 2) == fc1_in:
            spec_w = w if spec_w is None else spec_w
            if w in (70, 124):
                spec_w = w
    if spec_w is None:
        raise RuntimeError(f"cannot infer spectrogram width from fc1 fan-in {fc1_in}")
    return GeneratorConfig(
        frames=frames, prior_frames=prior_frames, pose_dim=pose_dim, d_model=d_model,
        d_inner=sd["encoder.layer_stack.0.pos_ffn.w_1.weight"].shape[0], n_layers=n_layers,
        n_head=mha.n_head, d_k=mha.d_k, d_v=mha.d_v, spec_w=spec_w,
        n_audio=int(round(frames / 15 * 16000)),
        n_position=sd["encoder.position_enc.pos_table"].shape[1])


def install(gen: torch.nn.Module, precision: str = "tc", spec_w: int | None = None):
    """Swap `gen.forward` for the libegx path; returns the Engine (call `.load_state_dict`
    on it again, or `gen.egx_sync()`, after loading a new checkpoint)."""
    cfg = config_from_module(gen)
    if spec_w is not None:
        cfg = GeneratorConfig(**{**cfg.__dict__, "spec_w": spec_w})
    dev = next(gen.parameters()).device
    eng = Engine(cfg, dev, precision=precision)
    eng.load_state_dict(gen.state_dict())
    original = gen.forward

    def forward(self, input_spectrum, text, prior_seq, sampled_emotion_feature=None):
        if self.training:
            args = (input_spectrum, text, prior_seq)
            return original(*args) if sampled_emotion_feature is None else original(*args, sampled_emotion_feature)
        text_embedding = self.text_encoder(text)
        poses, emo, sem, logits = eng.generator_forward(input_spectrum, prior_seq, sampled_emotion_feature)
        return poses, emo, sem, logits, text_embedding

    gen.forward = types.MethodType(forward, gen)
    gen.egx_engine = eng
    gen.egx_sync = lambda: eng.load_state_dict(gen.state_dict())
    return eng
