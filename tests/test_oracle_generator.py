"""CPU: the generator oracle against the golden vectors made from the REAL reference
(oracle/make_golden.py, run where /root/reference is mounted), plus structural pins."""
import os

import numpy as np
import pytest
import torch

from emotiongestures_b200 import BEAT, TED
from oracle import generator as og
from tests.helpers import CFGS, inputs, load_golden, model_and_sd, rel_max


@pytest.mark.parametrize("gold", ["ted_b2", "ted_b2_emotion", "beat_b1", "tedmem_b3"])
def test_oracle_reproduces_reference_golden(gold):
    g = load_golden(gold)
    name = gold.split("_")[0]
    cfg = CFGS[name]
    seed, n, with_emo = int(g["seed"]), int(g["n_clips"]), bool(g["with_emotion"])
    _, sd = model_and_sd(name, seed)
    spec, prior, emo = inputs(cfg, n, seed, with_emo)
    taps = og.Taps()
    with torch.no_grad():
        out = og.generator_forward(sd, cfg, spec, prior, emo, taps)
    for nm, got in zip(("poses", "emotion_feature", "semantic_feature", "emotion_logits"), out):
        assert rel_max(got, g[nm]) <= 5e-6, nm
    for nm in ("spectrum_feature", "prior_feature", "enc_output", "dec_output"):
        assert rel_max(taps[nm][0], g["tap_" + nm]) <= 5e-6, nm
    assert rel_max(taps["layer3"][0], g["tap_layer3"]) <= 5e-6
    for nm in ("stem", "layer1", "layer2", "layer3"):
        assert np.allclose(taps[nm].double().mean(dim=(0, 2, 3)).numpy(), g[f"tap_{nm}_mean"], atol=1e-5)


def test_state_dict_layout_is_the_reference_one():
    """Key count / names of SURVEY.md §8(b); order and shapes are asserted against the live
    reference inside oracle/make_golden.py."""
    m, sd = model_and_sd("ted", 0)
    keys = list(sd.keys())
    assert len(keys) == 424
    for k in ("audio_encoder.feat_extractor.layer3.5.se.fc.2.weight", "audio_encoder.final_conv1.bias",
              "prior_seq_encoder.conv2.weight", "emotion_classifer_header.6.bias",
              "encoder.position_enc.pos_table", "decoder.layer_stack.2.enc_attn.w_vs.weight",
              "decoder.layer_stack.0.slf_attn.fc.weight", "encoder.position_embeddings.weight",
              "text_encoder.tcn.network.2.net.4.weight_v", "post_projector.6.weight"):
        assert k in sd, k
    assert sd["audio_encoder.fc1.weight"].shape == (256, 32 * 18)
    mb, sdb = model_and_sd("beat", 1)
    assert sdb["audio_encoder.fc1.weight"].shape == (512, 32 * 31)
    assert sum(v.numel() for v in sdb.values() if v.is_floating_point()) > 46_000_000


def test_memory_prior_encoder_structure():
    """Full_model/Models_memory.py:335-345: the first prior_frames rows of the encoder input are the prior poses
    themselves, so their features do not depend on the memory nets (nor, through TM_Memory_Net's
    memory_encoding.t() @ pred_encoding batch sum, :287-288, on the clip's batch mates); only rows
    p .. p+chunk-1 see the spatial / temporal memory."""
    _, sd = model_and_sd("tedmem", 5)
    _, prior, _ = inputs(TED, 3, 5)
    p = TED.prior_frames
    with torch.no_grad():
        full = og.prior_memory_encoder(sd, prior)
        alone = og.prior_memory_encoder(sd, prior[:1])
        head = og._linear(sd, "prior_seq_encoder.post_header.2", og._linear(sd, "prior_seq_encoder.post_header.0", prior))
        sd2 = dict(sd)
        k = "prior_seq_encoder.temporal_memory.temporal_chunk_encoder.2.bias"
        sd2[k] = sd[k] + 1.0
        moved = og.prior_memory_encoder(sd2, prior)
    assert full.shape == (3, TED.frames, TED.d_model)
    assert torch.allclose(full[:, :p], head, atol=1e-6) and torch.allclose(alone[0, :p], full[0, :p], atol=1e-6)
    from tests.helpers import MEM_CHUNK
    assert torch.equal(moved[:, p + MEM_CHUNK:], full[:, p + MEM_CHUNK:])      # frames past the chunk are untouched
    assert not torch.equal(moved[:, p:p + MEM_CHUNK], full[:, p:p + MEM_CHUNK])


def test_structural_invariants():
    """emotion/semantic features are returned pre-fusion; the emotion injection
    replaces only the emotion addend; attention scales q before the product; sequence positions
    use the first F rows of the table."""
    _, sd = model_and_sd("ted", 0)
    spec, prior, emo = inputs(TED, 2, 7, True)
    with torch.no_grad():
        base = og.generator_forward(sd, TED, spec, prior)
        inj = og.generator_forward(sd, TED, spec, prior, emo)
        same = og.generator_forward(sd, TED, spec, prior, base[1])
    assert torch.equal(base[1], inj[1]) and torch.equal(base[2], inj[2]) and torch.equal(base[3], inj[3])
    assert not torch.allclose(base[0], inj[0])
    assert torch.allclose(base[0], same[0], atol=1e-6)       # injecting emotion_feature itself is a no-op
    q = torch.randn(1, 5, 256)
    out, attn = og.mha(sd, "encoder.layer_stack.0.slf_attn", q, q, 8, 64, 64)
    assert attn.shape == (1, 8, 5, 5) and torch.allclose(attn.sum(-1), torch.ones(1, 8, 5))
    # per-clip independence: batch composition never changes a clip's result
    with torch.no_grad():
        one = og.generator_forward(sd, TED, spec[1:], prior[1:])
    assert rel_max(one[0], base[0][1:]) <= 1e-5


def test_golden_files_are_small():
    d = os.path.join(os.path.dirname(__file__), "golden")
    assert sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) < 4 << 20
    assert BEAT.fc1_in == 992
