"""ORACLE tooling: pin oracle/generator.py against the REAL reference and write tests/golden/.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
It (1) imports Full_model.Models.Transformer with a stub for the unused
`torch_dct` import (Full_model/Models.py:8), (2) loads the synthetic
state_dict of oracle/synth.py into it (TED shape needs audio_encoder.fc1
rebuilt as Linear(32*18, d) — SURVEY.md fact 4), (3) asserts the restatement
matches the reference forward, and (4) stores inputs' seeds + the reference
outputs as small .npz fixtures.  The GPU box has no /root/reference; tests
there use only the fixtures.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emotiongestures_b200 import BEAT, TED, Transformer  # noqa: E402
from emotiongestures_b200.generator import MemoryTransformer  # noqa: E402
from oracle import generator as og  # noqa: E402
from oracle import synth  # noqa: E402

REF = "/root/reference"


def load_reference(cfg, chunk=0):
    """Full_model.Models.Transformer, or (chunk > 0) Full_model.Models_memory.Transformer with args.chunk = chunk."""
    sys.path.insert(0, REF)
    sys.modules.setdefault("torch_dct", types.ModuleType("torch_dct"))
    if chunk:
        from Full_model.Models_memory import Transformer as RefTransformer
    else:
        from Full_model.Models import Transformer as RefTransformer

    class Args:
        freeze_wordembed = False
        hidden_size = cfg.tcn_hidden
        n_layers = cfg.tcn_layers
        wordembed_dim = cfg.wordembed_dim
        dropout_prob = 0.1

    Args.chunk = chunk

    class Lang:
        n_words = cfg.n_words
        word_embedding_weights = None

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = RefTransformer(Args(), Lang(), frames=cfg.frames, pose_dim=cfg.pose_dim,
                           prior_frames=cfg.prior_frames, d_word_vec=cfg.d_model,
                           d_model=cfg.d_model, d_inner=cfg.d_inner, n_layers=cfg.n_layers,
                           n_head=cfg.n_head, d_k=cfg.d_k, d_v=cfg.d_v, dropout=0.1,
                           n_position=cfg.n_position)
    if cfg.fc1_in != 32 * 31:
        g.audio_encoder.fc1 = torch.nn.Linear(cfg.fc1_in, cfg.d_model)
    return g.eval()


def run(cfg, name, n_clips, seed, with_emotion, chunk=0):
    ref = load_reference(cfg, chunk)
    mine = (MemoryTransformer.from_config(cfg, chunk) if chunk else Transformer.from_config(cfg)).eval()
    tmpl = mine.state_dict()
    ref_sd = ref.state_dict()
    assert list(tmpl.keys()) == list(ref_sd.keys()), "state_dict key layout differs from reference"
    for k in tmpl:
        assert tmpl[k].shape == ref_sd[k].shape, k
    assert torch.equal(tmpl["encoder.position_enc.pos_table"], ref_sd["encoder.position_enc.pos_table"])
    # the geometry install() derives from the LIVE reference module is the one the goldens are made at
    from emotiongestures_b200.dropin import config_from_module
    got = config_from_module(ref)
    for f in ("frames", "prior_frames", "pose_dim", "d_model", "d_inner", "n_layers", "n_head", "d_k", "d_v", "spec_w", "n_position"):
        assert getattr(got, f) == getattr(cfg, f), ("config_from_module", f)
    sd = synth.synth_state_dict(tmpl, seed)
    ref.load_state_dict(sd)
    spec = torch.from_numpy(synth.synth_spec(n_clips, cfg.n_mels, cfg.spec_w, seed))
    prior = torch.from_numpy(synth.synth_prior(n_clips, cfg.prior_frames, cfg.pose_dim, seed))
    text = torch.zeros(n_clips, 60, dtype=torch.int64)
    emo_in = torch.from_numpy(synth.synth_emotion(n_clips, cfg.frames, cfg.d_model, seed)) if with_emotion else None

    hooks, taps_ref = [], {}
    fe = ref.audio_encoder.feat_extractor
    for nm, m in (("stem", fe.bn1), ("layer1", fe.layer1), ("layer2", fe.layer2), ("layer3", fe.layer3),
                  ("spectrum_feature", ref.audio_encoder), ("prior_feature", ref.prior_seq_encoder),
                  ("enc_output", ref.encoder), ("dec_output", ref.decoder)):
        hooks.append(m.register_forward_hook(
            lambda _m, _i, o, nm=nm: taps_ref.__setitem__(nm, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    with torch.no_grad():
        if with_emotion and chunk:
            ref_out = ref(spec, text, prior, emo_in)          # Models_memory.py:521 takes the 4th argument itself
        elif with_emotion:
            # Models.py has no 4th argument; apply Models_memory.py:551-555 by patching the sum
            # exactly as that file does: fusion = sampled + semantic.
            sys.modules.setdefault("torch_dct", types.ModuleType("torch_dct"))
            ref_out = forward_with_emotion(ref, spec, text, prior, emo_in)
        else:
            ref_out = ref(spec, text, prior)
    for h in hooks:
        h.remove()
    taps = og.Taps()
    with torch.no_grad():
        out = og.generator_forward(sd, cfg, spec, prior, emo_in, taps)
        out64 = og.generator_forward(og.cast_state_dict(sd, torch.float64), cfg, spec.double(),
                                     prior.double(), None if emo_in is None else emo_in.double())
    names = ["poses", "emotion_feature", "semantic_feature", "emotion_logits"]
    for nm, a, b in zip(names, out, ref_out[:4]):
        err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
        print(f"  {name}: oracle vs reference {nm}: rel max-abs {err:.2e}")
        # fp32 re-association noise depends on the BLAS thread count; the 8-wide logits (sums of 8704 / 30720
        # products) scatter up to ~3e-6 between runs, everything else stays below 2e-6
        assert err <= (5e-6 if nm == "emotion_logits" else 2e-6), (nm, err)
    for nm in taps_ref:
        err = (taps[nm] - taps_ref[nm]).abs().max().item() / taps_ref[nm].abs().max().item()
        print(f"  {name}: oracle vs reference tap {nm}: rel max-abs {err:.2e}")
        assert err <= 2e-6, (nm, err)
    e64 = (out64[0].float() - ref_out[0]).norm().item() / ref_out[0].norm().item()
    print(f"  {name}: fp32 reference vs fp64 oracle poses rel-Frobenius {e64:.2e}")
    # text embedding of the mirror module (PyTorch) vs reference
    mine.load_state_dict(sd)
    with torch.no_grad():
        te = mine.text_encoder(text)
    assert torch.allclose(te, ref_out[4], atol=1e-6), "text encoder mirror differs"
    arrs = {"seed": np.int64(seed), "n_clips": np.int64(n_clips), "with_emotion": np.int64(with_emotion)}
    for nm, a in zip(names, ref_out[:4]):
        arrs[nm] = a.numpy()
    arrs["text_embedding_l2"] = np.float64(ref_out[4].double().norm().item())
    # keep intermediates small: clip 0 only, float16 storage is enough for stage localisation
    for nm in ("stem", "layer1", "layer2", "layer3"):
        arrs["tap_" + nm + "_mean"] = taps_ref[nm].double().mean(dim=(0, 2, 3)).numpy()
        arrs["tap_" + nm + "_absmax"] = np.float64(taps_ref[nm].abs().max().item())
    arrs["tap_layer3"] = taps_ref["layer3"][0].numpy().astype(np.float32)
    for nm in ("spectrum_feature", "prior_feature", "enc_output", "dec_output"):
        arrs["tap_" + nm] = taps_ref[nm][0].numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def forward_with_emotion(ref, spec, text, prior, sampled):
    """Reference Models.py forward with the one-line emotion injection of
    Models_memory.py:551-555, driven through the reference's own sub-modules."""
    text_embedding = ref.text_encoder(text)
    x = ref.audio_encoder(spec.unsqueeze(1))
    p = ref.prior_seq_encoder(prior)
    emo = ref.emotion_proj(x)
    sem = ref.semantic_proj(x)
    logits = ref.emotion_classifer_header(emo.reshape(emo.shape[0], -1))
    fusion = ref.fusion_proj(sampled + sem)
    enc, *_ = ref.encoder(fusion, None)
    dec, *_ = ref.decoder(p, None, enc, None)
    return ref.post_projector(dec), emo, sem, logits, text_embedding


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    print("TED"); run(TED, "ted_b2", 2, 0, False)
    print("TED + emotion injection"); run(TED, "ted_b2_emotion", 2, 3, True)
    print("BEAT"); run(BEAT, "beat_b1", 1, 1, False)
    # Models_memory.Transformer (Prior_MemoryEncoder); 3 clips so that the temporal memory's batch sum is exercised
    print("TED, memory prior encoder, chunk 4"); run(TED, "tedmem_b3", 3, 5, True, chunk=4)
