"""ORACLE tooling: pin oracle/aux_models.py against the REAL reference modules and write
tests/golden/aux_*.npz.  Build container only (needs /root/reference):

    python -m oracle.make_golden_aux

For every small network of SURVEY.md §8 rows C4 / E1 / D1 it instantiates the reference class
(stubbing the absent `fasttext` import of model/vocab.py), checks that the host mirror in
emotiongestures_b200/aux_models.py has the same state_dict keys and shapes, loads the synthetic
weights of oracle/synth.py into the reference, runs it with explicit noise (the reference's
torch.randn calls are replayed by seeding and, where the draw happens on the CPU inside the
module, by patching torch.randn for the duration of the call), asserts the restatement agrees
to <= 2e-6 and stores seeds + reference outputs as small fixtures.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emotiongestures_b200 import aux_models as mirrors  # noqa: E402
from oracle import aux_models as oa  # noqa: E402
from oracle import synth  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def rnd(shape, seed, tag):
    return torch.from_numpy(np.random.default_rng([seed, tag]).standard_normal(shape).astype(np.float32))


def same_layout(ref, mine):
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()), (set(a) ^ set(b))
    for k in a:
        assert a[k].shape == b[k].shape, k


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


class patched_randn:
    """Replay the module-internal torch.randn / randn_like draw with a tensor we control."""

    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self._randn, self._like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: self.value.clone()
        torch.randn_like = lambda *a, **k: self.value.clone()

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._like


def main():
    sys.path.insert(0, REF)
    sys.modules.setdefault("fasttext", types.ModuleType("fasttext"))
    sys.modules.setdefault("torch_dct", types.ModuleType("torch_dct"))
    import importlib
    out = {}

    # ---- C4: Full_model/BEAT_CVAE.py MLP_Reconstruct --------------------------------------------------
    ref = importlib.import_module("Full_model.BEAT_CVAE").MLP_Reconstruct().eval()
    mine = mirrors.MLP_Reconstruct()
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 11)
    ref.load_state_dict(sd)
    n = 37
    x, y, eps = rnd((n, 90), 11, 1), rnd((n, 90), 11, 2), rnd((n, 32), 11, 3)
    with torch.no_grad(), patched_randn(eps):
        r_out, r_mu, r_lv = ref(x, y)
        r_smp = ref.sample(y)
    with torch.no_grad():
        o_out, o_mu, o_lv = oa.cvae_forward(sd, x, y, eps)
        o_smp = oa.cvae_decode(sd, y, eps)
    for nm, a, b in (("out", o_out, r_out), ("mu", o_mu, r_mu), ("logvar", o_lv, r_lv), ("sample", o_smp, r_smp)):
        print(f"  cvae {nm}: oracle vs reference {rel(a, b):.2e}")
        assert rel(a, b) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_cvae.npz"), seed=11, n=n, out=r_out.numpy(), mu=r_mu.numpy(),
                        logvar=r_lv.numpy(), sample=r_smp.numpy())

    # ---- E1: CAVE/BEAT_CVAE.py MLP_Reconstruct_v3.sample ----------------------------------------------
    ref = importlib.import_module("CAVE.BEAT_CVAE").MLP_Reconstruct_v3().eval()
    mine = mirrors.MLP_Reconstruct_v3()
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 12)
    ref.load_state_dict(sd)
    n = 3
    lab = torch.tensor([1, 7, 4])
    y = torch.nn.functional.one_hot(lab, 8).float()
    z = rnd((n, 32), 12, 1)
    with torch.no_grad(), patched_randn(z):
        r = ref.sample(y)
    with torch.no_grad():
        o = oa.cvae3_sample(sd, y, z)
    print(f"  cvae3 sample: oracle vs reference {rel(o, r):.2e}", tuple(r.shape))
    assert rel(o, r) <= 2e-6 and tuple(r.shape) == (n, 60, 512)
    np.savez_compressed(os.path.join(GOLD, "aux_cvae3.npz"), seed=12, labels=lab.numpy(), sample=r.numpy().astype(np.float32))

    # ---- D1: MotionAE.encoder (TED) ------------------------------------------------------------------
    ref = importlib.import_module("model.motion_ae").MotionAE(126, 128).eval()
    mine = mirrors.MotionAE(126, 128)
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 13)
    ref.load_state_dict(sd)
    n = 5
    poses = rnd((n, 34, 126), 13, 1)
    with torch.no_grad():
        _, r = ref(poses)
        o = oa.pose_encoder(sd, poses, "encoder.")
    print(f"  motion_ae z: oracle vs reference {rel(o, r):.2e}", tuple(r.shape))
    assert rel(o, r) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_motion_ae.npz"), seed=13, n=n, z=r.numpy())

    # ---- D1: embedding_net.PoseEncoderConv (BEAT) ------------------------------------------------------
    ref = importlib.import_module("model.embedding_net").PoseEncoderConv(60, 282).eval()
    mine = mirrors.PoseEncoderConv(60, 282)
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 14)
    ref.load_state_dict(sd)
    n = 4
    poses = rnd((n, 60, 282), 14, 1)
    with torch.no_grad():
        _, r, _ = ref(poses, False)
        o = oa.pose_encoder(sd, poses, "", fc_mu=True)
    print(f"  pose_enc mu: oracle vs reference {rel(o, r):.2e}", tuple(r.shape))
    assert rel(o, r) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_pose_enc.npz"), seed=14, n=n, mu=r.numpy())

    # ---- D1: model/FGD.py MLP_Reconstruct latent -------------------------------------------------------
    ref = importlib.import_module("model.FGD").MLP_Reconstruct().eval()
    mine = mirrors.FGDNet()
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 15)
    ref.load_state_dict(sd)
    n = 2
    poses = rnd((n, 60, 282), 15, 1)
    with torch.no_grad():
        _, r = ref(poses)
        o = oa.fgd_latent(sd, poses)
    print(f"  fgd latent: oracle vs reference {rel(o, r):.2e}", tuple(r.shape))
    assert rel(o, r) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_fgd_mlp.npz"), seed=15, n=n, latent=r.numpy())
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    if not {"--emotion-net", "--skeleton", "--diversity"} & set(sys.argv):
        main()
    for f in sorted(os.listdir(GOLD)):
        if f.startswith("aux_"):
            print(f"  {f}: {os.path.getsize(os.path.join(GOLD, f)) / 1024:.0f} KiB")


def emotion_net_golden():
    """C3: model/audio_emotion_classifer.py EmotionNet (1 GB of fp32 weights: kept out of main())."""
    sys.path.insert(0, REF)
    sys.modules.setdefault("fasttext", types.ModuleType("fasttext"))
    import importlib
    ref = importlib.import_module("model.audio_emotion_classifer").EmotionNet().eval()
    mine = mirrors.EmotionNet()
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 16)
    del mine
    ref.load_state_dict(sd)
    n = 2
    spec = torch.from_numpy(synth.synth_spec(n, 128, 124, 16))
    taps_ref = {}
    hook = ref.emotion_encoder.register_forward_hook(lambda _m, _i, o: taps_ref.__setitem__("layer4", o.detach().clone()))
    with torch.no_grad():
        r = ref(spec)
    hook.remove()
    taps = {}
    with torch.no_grad():
        o = oa.emotion_net(sd, spec, taps)
    print(f"  emotion_net logits: oracle vs reference {rel(o, r):.2e}; layer4 {rel(taps['layer4'], taps_ref['layer4']):.2e}")
    assert rel(o, r) <= 2e-6 and rel(taps["layer4"], taps_ref["layer4"]) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_emotion_net.npz"), seed=16, n=n, logits=r.numpy(),
                        layer4_mean=taps_ref["layer4"].double().mean(dim=(2, 3)).numpy().astype(np.float32))


def skeleton_golden():
    """(f)2: skeleton_classifer/Models.py Transformer, built as test_emotion_gesture_diversity_iterative.py:158 does
    (d_model 512, d_k = d_v = 64, n_position 60; pose_dim 282, d_inner 2048 = BEAT geometry)."""
    sys.path.insert(0, REF)
    sys.modules.setdefault("torch_dct", types.ModuleType("torch_dct"))      # imported, never used (Models.py:8)
    import importlib
    kw = dict(class_dim=8, pose_dim=282, d_word_vec=512, d_model=512, d_inner=2048, n_layers=3, n_head=8, d_k=64, d_v=64,
              n_position=60)
    ref = importlib.import_module("skeleton_classifer.Models").Transformer(**kw).eval()
    mine = mirrors.SkeletonClassifier(**kw)
    same_layout(ref, mine)
    sd = synth.synth_state_dict(mine.state_dict(), 17)
    ref.load_state_dict(sd)
    n = 3
    poses = rnd((n, 60, 282), 17, 1)
    with torch.no_grad():
        r_logits, r_mid = ref(poses)
        o_logits, o_mid = oa.skeleton_classifier(sd, poses)
    print(f"  skeleton logits: oracle vs reference {rel(o_logits, r_logits):.2e}; mid_feature {rel(o_mid, r_mid):.2e}")
    assert rel(o_logits, r_logits) <= 2e-6 and rel(o_mid, r_mid) <= 2e-6
    np.savez_compressed(os.path.join(GOLD, "aux_skeleton.npz"), seed=17, n=n, logits=r_logits.numpy(),
                        mid=r_mid.numpy().astype(np.float32))


def diversity_golden():
    """model/FHD_score.py:244-280 diversity_score on a seeded array with np.random.seed(123): the reference draws its
    clip pairs from the global numpy RNG, so the same seed through np.random.RandomState reproduces it."""
    sys.path.insert(0, REF)
    sys.modules.setdefault("fasttext", types.ModuleType("fasttext"))
    import importlib
    from emotiongestures_b200.evaluate import diversity_score
    ref = importlib.import_module("model.FHD_score")
    n = 37
    x = np.random.default_rng(5).standard_normal((n * 60, 512)).astype(np.float32)
    np.random.seed(123)
    r_score, r_int = ref.diversity_score(x.copy(), "cpu")
    m_score, m_int = diversity_score(x, np.random.RandomState(123))
    assert np.allclose(r_score, m_score) and np.allclose(r_int[0], m_int[0]) and np.allclose(r_int[1], m_int[1])
    print(f"  diversity: reference {float(r_score[0]):.6f} mirror {float(m_score[0]):.6f}")
    np.savez_compressed(os.path.join(GOLD, "diversity.npz"), seed=123, data_seed=5, n=n, score=np.asarray(r_score),
                        lo=np.asarray(r_int[0]), hi=np.asarray(r_int[1]))


if __name__ == "__main__" and "--emotion-net" in sys.argv:
    emotion_net_golden()
if __name__ == "__main__" and "--diversity" in sys.argv:
    diversity_golden()
if __name__ == "__main__" and "--skeleton" in sys.argv:
    skeleton_golden()
