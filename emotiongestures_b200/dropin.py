"""`install(generator)`: put libegx behind a LIVE reference generator.

`generator` is the reference's own Full_model.Models.Transformer (or
Models_memory.Transformer), bare or wrapped in nn.DataParallel, already
`.eval()`-ed and on a CUDA device, exactly as
test_emotion_gesture_diversity_iterative.py:135-145 leaves it.  Its `forward` is
overridden (at class level, so DataParallel replicas get it too) by one that keeps the
signature and the 5-tuple but routes the pose path through the C ABI on the device the
call runs on; the text encoder (dead w.r.t. poses) keeps running as the module's own
PyTorch sub-module.  In training mode the module's original forward is called untouched
(the library is inference-only).
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
    """Route the pose path of a live reference generator through libegx; returns the device-0 Engine.

    `gen` may be the bare module or an `nn.DataParallel` wrapper around it (the evaluation script wraps the generator
    whenever several GPUs are visible, test_emotion_gesture_diversity_iterative.py:137-138).  The module's class is
    swapped for a subclass of itself that overrides `forward` — a class-level method, because DataParallel rebuilds
    its replicas on every call as shallow copies of the instance (`type(self).__new__` + `__dict__.copy()`), so an
    instance-bound forward would make every replica call the ORIGINAL module on device 0.  The engines live in an
    `EngineSet` inside the instance `__dict__`, shared by reference with all replicas and keyed by device: each
    replica thread uses the handle of the GPU its inputs were scattered to.  After loading a new checkpoint call
    `module.egx_sync()`.  In training mode the reference's own forward runs untouched (the library is inference-only).
    """
    from .generator import EngineSet, _call_device
    module = gen.module if isinstance(gen, torch.nn.DataParallel) else gen
    if getattr(type(module), "_egx_installed", False):
        raise RuntimeError("install() was already applied to this module")
    cfg = config_from_module(module)
    if spec_w is not None:
        cfg = GeneratorConfig(**{**cfg.__dict__, "spec_w": spec_w})
    dev = next(module.parameters()).device
    base = type(module)

    def forward(self, input_spectrum, text, prior_seq, sampled_emotion_feature=None):
        if self.training:
            args = (input_spectrum, text, prior_seq)
            return (base.forward(self, *args) if sampled_emotion_feature is None
                    else base.forward(self, *args, sampled_emotion_feature))
        dev_ = _call_device(self, input_spectrum) or next(self.parameters()).device
        eng = self._egx_set.get(dev_, self._egx_precision)
        text_embedding = self.text_encoder(text)
        poses, emo, sem, logits = eng.generator_forward(input_spectrum, prior_seq, sampled_emotion_feature)
        return poses, emo, sem, logits, text_embedding

    patched = type(base.__name__, (base,), {"forward": forward, "_egx_installed": True, "__module__": base.__module__,
                                            "egx_sync": lambda self: self._egx_set.sync()})
    engines = EngineSet(module, cfg)
    eng = engines.get(dev, precision)        # fails loudly here (no library / not sm_100 / CPU module), module untouched
    module.__dict__["_egx_set"] = engines
    module.__dict__["_egx_precision"] = precision
    module.__dict__["egx_engine"] = eng
    module.__class__ = patched
    return eng
