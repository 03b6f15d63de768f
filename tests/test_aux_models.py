"""Small networks either side of the generator.

CPU part: the oracle restatement (oracle/aux_models.py) reproduces the outputs the REAL reference
modules produced for the same synthetic weights and inputs (tests/golden/aux_*.npz, written by
oracle/make_golden_aux.py), and the host mirrors keep the reference's state_dict layout.
GPU part (-m gpu): the CUDA path through the C ABI against those reference outputs.

Tolerances.  The CUDA kernels compute in fp32 but fold eval-mode algebra on the host in fp64
(Linear chains without activations collapse, BatchNorm becomes scale/shift), so they differ from
the reference's fp32 chain only by rounding order: gate 2e-5 relative to the output scale.  The
per-frame FGD MLP runs on the fp16 tensor-core GEMM: gate 2e-3 (the bf16-path tolerance of the
north star)."""
import numpy as np
import pytest
import torch

from emotiongestures_b200 import aux_models as mirrors
from oracle import aux_models as oa
from oracle import synth
from tests.helpers import load_golden, rel_max


def rnd(shape, seed, tag):
    return torch.from_numpy(np.random.default_rng([seed, tag]).standard_normal(shape).astype(np.float32))


def build(cls, seed, *args):
    m = cls(*args).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m, sd


CASES = {
    "cvae": (mirrors.MLP_Reconstruct, 11, ()),
    "cvae3": (mirrors.MLP_Reconstruct_v3, 12, ()),
    "motion_ae": (mirrors.MotionAE, 13, (126, 128)),
    "pose_enc": (mirrors.PoseEncoderConv, 14, (60, 282)),
    "fgd_mlp": (mirrors.FGDNet, 15, ()),
    "emotion_net": (mirrors.EmotionNet, 16, ()),     # 1 GB of fp32 weights: built once per session
    # skeleton_classifer Transformer as test_emotion_gesture_diversity_iterative.py:158 builds it (BEAT geometry):
    # class_dim, pose_dim, src_pad, trg_pad, d_word_vec, d_model, d_inner, n_layers, n_head, d_k, d_v, dropout, n_position
    "skeleton": (mirrors.SkeletonClassifier, 17, (8, 282, 1, 1, 512, 512, 2048, 3, 8, 64, 64, 0.2, 60)),
}
_big = {}


def reference_and_inputs(name):
    g = load_golden("aux_" + name)
    cls, seed, args = CASES[name]
    if name == "emotion_net":
        if name not in _big:
            _big[name] = build(cls, seed, *args)
        m, sd = _big[name]
        ins = dict(mfcc=torch.from_numpy(synth.synth_spec(int(g["n"]), 128, 124, seed)))
        return m, sd, ins, {"logits": torch.from_numpy(g["logits"])}
    m, sd = build(cls, seed, *args)
    if name == "cvae":
        n = int(g["n"])
        ins = dict(x=rnd((n, 90), seed, 1), y=rnd((n, 90), seed, 2), eps=rnd((n, 32), seed, 3))
        ref = {k: torch.from_numpy(g[k]) for k in ("out", "mu", "logvar", "sample")}
    elif name == "cvae3":
        lab = torch.from_numpy(g["labels"])
        ins = dict(y=torch.nn.functional.one_hot(lab, 8).float(), z=rnd((len(lab), 32), seed, 1))
        ref = {"sample": torch.from_numpy(g["sample"])}
    elif name == "motion_ae":
        ins = dict(poses=rnd((int(g["n"]), 34, 126), seed, 1))
        ref = {"z": torch.from_numpy(g["z"])}
    elif name == "pose_enc":
        ins = dict(poses=rnd((int(g["n"]), 60, 282), seed, 1))
        ref = {"mu": torch.from_numpy(g["mu"])}
    elif name == "skeleton":
        ins = dict(poses=rnd((int(g["n"]), 60, 282), seed, 1))
        ref = {"logits": torch.from_numpy(g["logits"]), "mid": torch.from_numpy(g["mid"])}
    else:
        ins = dict(poses=rnd((int(g["n"]), 60, 282), seed, 1))
        ref = {"latent": torch.from_numpy(g["latent"])}
    return m, sd, ins, ref


def oracle_outputs(name, sd, ins):
    with torch.no_grad():
        if name == "cvae":
            out, mu, lv = oa.cvae_forward(sd, ins["x"], ins["y"], ins["eps"])
            return {"out": out, "mu": mu, "logvar": lv, "sample": oa.cvae_decode(sd, ins["y"], ins["eps"])}
        if name == "cvae3":
            return {"sample": oa.cvae3_sample(sd, ins["y"], ins["z"])}
        if name == "motion_ae":
            return {"z": oa.pose_encoder(sd, ins["poses"], "encoder.")}
        if name == "pose_enc":
            return {"mu": oa.pose_encoder(sd, ins["poses"], "", fc_mu=True)}
        if name == "emotion_net":
            return {"logits": oa.emotion_net(sd, ins["mfcc"])}
        if name == "skeleton":
            logits, mid = oa.skeleton_classifier(sd, ins["poses"])
            return {"logits": logits, "mid": mid}
        return {"latent": oa.fgd_latent(sd, ins["poses"])}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    _, sd, ins, ref = reference_and_inputs(name)
    got = oracle_outputs(name, sd, ins)
    for k, r in ref.items():
        assert got[k].shape == r.shape
        assert rel_max(got[k], r) <= 2e-6, (name, k)


def test_mirrors_refuse_cpu_and_training_mode():
    m, _ = build(mirrors.MLP_Reconstruct, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(torch.zeros(2, 90))
    m3, _ = build(mirrors.MLP_Reconstruct_v3, 1)
    with pytest.raises(RuntimeError, match="training-side"):
        m3(torch.zeros(1, 60, 512), torch.zeros(1, 8))


def test_pose_encoder_lengths_follow_the_reference():
    with pytest.raises(ValueError):
        mirrors.PoseEncoderConv(34, 126)         # embedding_net's out_net.0 is 800 wide: 60 frames only
    with pytest.raises(ValueError):
        mirrors.MotionAEEncoder(60, 126, 128)    # motion_ae's is 384 wide: 34 frames only


# ------------------------------------------------------------------------------------------------
# GPU parity through the C ABI
# ------------------------------------------------------------------------------------------------
def device_outputs(name, m, ins):
    m = m.cuda()
    d = {k: v.cuda() for k, v in ins.items()}
    with torch.no_grad():
        if name == "cvae":
            out, mu, lv = m(d["x"], d["y"], eps=d["eps"])
            return {"out": out, "mu": mu, "logvar": lv, "sample": m.sample(d["y"], z=d["eps"])}
        if name == "cvae3":
            return {"sample": m.sample(d["y"], z=d["z"])}
        if name == "motion_ae":
            return {"z": m(d["poses"])[1]}
        if name == "pose_enc":
            return {"mu": m(d["poses"], False)[1]}
        if name == "emotion_net":
            return {"logits": m(d["mfcc"])}
        if name == "skeleton":
            logits, mid = m(d["poses"])
            return {"logits": logits, "mid": mid}
        return {"latent": m(d["poses"])[1]}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_golden(name):
    m, _, ins, ref = reference_and_inputs(name)
    got = device_outputs(name, m, ins)
    tol = {"fgd_mlp": 2e-3, "emotion_net": 2e-3, "skeleton": 2e-3}.get(name, 2e-5)
    for k, r in ref.items():
        g = got[k].cpu()
        assert g.shape == r.shape and not torch.isnan(g).any()
        assert rel_max(g, r) <= tol, (name, k, rel_max(g, r))


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("cvae", 1), ("cvae", 1000), ("motion_ae", 300), ("pose_enc", 150), ("cvae3", 150),
                                    ("fgd_mlp", 77), ("skeleton", 1), ("skeleton", 77)])
def test_cuda_matches_oracle_at_ragged_sizes(name, n):
    """Sizes that do not divide the kernels' tiles / exceed one wave, against the oracle on the same inputs."""
    cls, seed, args = CASES[name]
    m, sd = build(cls, seed + 100, *args)
    if name == "cvae":
        ins = dict(x=rnd((n, 90), n, 1), y=rnd((n, 90), n, 2), eps=rnd((n, 32), n, 3))
    elif name == "cvae3":
        ins = dict(y=torch.nn.functional.one_hot(torch.arange(n) % 8, 8).float(), z=rnd((n, 32), n, 1))
    elif name == "motion_ae":
        ins = dict(poses=rnd((n, 34, 126), n, 1))
    else:
        ins = dict(poses=rnd((n, 60, 282), n, 1))
    ref = oracle_outputs(name, oa.cast(sd, torch.float64), {k: v.double() for k, v in ins.items()})
    got = device_outputs(name, m, ins)
    tol = 2e-3 if name in ("fgd_mlp", "skeleton") else 2e-5
    for k, r in ref.items():
        assert rel_max(got[k].cpu(), r) <= tol, (name, k, rel_max(got[k].cpu(), r))


@pytest.mark.gpu
def test_cuda_empty_batches_and_shape_errors():
    m, _ = build(mirrors.MLP_Reconstruct, 3)
    m = m.cuda()
    assert m.sample(torch.zeros(0, 90, device="cuda")).shape == (0, 90)
    enc, _ = build(mirrors.MotionAE, 3, 126, 128)
    enc = enc.cuda()
    with pytest.raises(RuntimeError, match="does not match"):
        enc(torch.zeros(2, 60, 126, device="cuda"))
    sk, _ = build(*[CASES["skeleton"][0], 3, *CASES["skeleton"][2]])
    sk = sk.cuda()
    logits, mid = sk(torch.zeros(0, 60, 282, device="cuda"))
    assert logits.shape == (0, 8) and mid.shape == (0, 60, 512)
    with pytest.raises(RuntimeError, match="do not match"):
        sk(torch.zeros(2, 34, 282, device="cuda"))            # post_projector flattens exactly n_position frames
    with pytest.raises(RuntimeError, match="d_k = d_v = 64"):
        mirrors.SkeletonClassifier().eval().cuda()(torch.zeros(1, 60, 242, device="cuda"))   # reference defaults: d_k = 32


@pytest.mark.gpu
def test_evaluation_loop_matches_the_oracle_composition():
    """emotiongestures_b200.evaluate.GestureEvaluator = test_emotion_gesture_diversity_iterative.py:191-255 (BEAT
    geometry) without the dataset and the beat metric: sampler -> generator -> skeleton classifier -> FGD features ->
    mean / covariance, against the same chain built from the oracle restatements on the CPU."""
    from emotiongestures_b200 import BEAT, Transformer
    from emotiongestures_b200 import fgd as fgd_mod
    from emotiongestures_b200.evaluate import GestureEvaluator
    from oracle import generator as og
    gen = Transformer.from_config(BEAT).eval()
    gsd = synth.synth_state_dict(gen.state_dict(), 21)
    gen.load_state_dict(gsd)
    vae, vsd = build(mirrors.MLP_Reconstruct_v3, 22)
    skel, ssd = build(CASES["skeleton"][0], 23, *CASES["skeleton"][2])
    fgdn, fsd = build(mirrors.FGDNet, 24)
    ev = GestureEvaluator(gen.cuda(), vae.cuda(), skel.cuda(), fgdn.cuda(), n_pre_poses=BEAT.prior_frames)
    ref_feats_p, ref_feats_t, ref = [], [], dict(acc=0.0, rot=0.0, l2=0.0, ambiguous=0)
    for it in range(2):
        n = 3
        spec = torch.from_numpy(synth.synth_spec(n, BEAT.n_mels, BEAT.spec_w, 40 + it))
        poses = rnd((n, BEAT.frames, BEAT.pose_dim), 41 + it, 1) * 0.3
        lab = torch.tensor([(it + i) % 8 for i in range(n)])
        y = torch.nn.functional.one_hot(lab, 8).float()
        z = rnd((n, 32), 42 + it, 2)
        text = torch.zeros(n, 60, dtype=torch.int64)
        rng = np.random.default_rng(50 + it)
        ons = [[np.sort(rng.choice(120, size=int(rng.integers(1, 8)), replace=False)) for _ in range(3)] for _ in range(n)]
        pred = ev.step(spec, text, poses, y, z=z, onsets=ons).cpu()
        # beat alignment (model/Beat_score_v2.py) of the clips the GPU path produced, against the oracle on those clips
        from oracle import beat as ob
        ref["beat"] = ref.get("beat", 0.0) + sum(
            ob.calculate_align(ons[i], ob.load_pose(pred[i].numpy(), 0, BEAT.frames // 15, 15, 2), 0.3, 15) for i in range(n))
        with torch.no_grad():
            sampled = oa.cvae3_sample(vsd, y, z)
            o_pred = og.generator_forward(gsd, BEAT, spec, poses[:, :BEAT.prior_frames], sampled)[0]
            o_logits, _ = oa.skeleton_classifier(ssd, o_pred)
            ref_feats_p.append(oa.fgd_latent(fsd, o_pred).reshape(-1, 512))
            ref_feats_t.append(oa.fgd_latent(fsd, poses).reshape(-1, 512))
        assert rel_max(pred, o_pred) <= 3e-3
        ref["acc"] += 100.0 * (o_logits.argmax(1) == lab).double().mean().item()
        # a clip's predicted class may legitimately differ only where the oracle's own top-2 margin is inside the
        # classifier's 2e-3 parity tolerance
        top2 = o_logits.topk(2, dim=1).values
        ref["ambiguous"] += int(((top2[:, 0] - top2[:, 1]) <= 4e-3 * o_logits.abs().max()).sum())
        ref["rot"] += (poses.reshape(n, -1, 6) - o_pred.reshape(n, -1, 6)).abs().mean().item()
        ref["l2"] += (poses - o_pred).norm(dim=-1).mean().item()
    out = ev.finalize()
    fp, ft = torch.cat(ref_feats_p).double().numpy(), torch.cat(ref_feats_t).double().numpy()
    mu_p, sig_p = out["pred_stats"]
    assert np.allclose(mu_p, fp.mean(0), rtol=0, atol=3e-3 * np.abs(fp).max())
    assert np.allclose(sig_p, np.cov(fp, rowvar=False), rtol=0, atol=6e-3 * np.abs(np.cov(fp, rowvar=False)).max())
    mu_t, sig_t = out["target_stats"]
    assert np.allclose(mu_t, ft.mean(0), rtol=0, atol=3e-3 * np.abs(ft).max())
    assert abs(out["rotation_error_deg"] - ref["rot"] / 2 * 57.2958) <= 3e-3 * ref["rot"] / 2 * 57.2958
    assert abs(out["l2_pose"] - ref["l2"] / 2) <= 3e-3 * ref["l2"] / 2
    assert abs(out["emotion_acc_percent"] - ref["acc"] / 2) <= 100.0 * ref["ambiguous"] / 6 + 1e-9
    assert abs(out["beat_score"] - ref["beat"] / 6) <= 1e-12
    want = fgd_mod.frechet_distance(fp.mean(0), np.cov(fp, rowvar=False), ft.mean(0), np.cov(ft, rowvar=False))
    assert np.isfinite(out["fgd"]) and abs(out["fgd"] - want) <= 2e-2 * max(abs(want), 1.0)
    assert abs(out["fgd"] - out["fgd_host"]) <= 1e-6 * max(abs(want), 1.0)      # device eigh tail == host numpy tail


@pytest.mark.gpu
def test_emotion_net_and_cvae_at_baseline_batch_sizes():
    """SURVEY.md §8(d) config 3 (EmotionNet, B = 256) and config 4 (CVAE, 1 M rows) at their full sizes, through
    size-independent properties: every clip's / row's result equals, bit for bit, what the same clip gives in a small
    batch (oracle-checked sizes) wherever it sits in the large one, and a sample agrees with the oracle."""
    m, sd = _big["emotion_net"] if "emotion_net" in _big else build(*CASES["emotion_net"][:2])
    _big["emotion_net"] = (m, sd)
    net = m.cuda()
    g = torch.Generator().manual_seed(5)
    spec = torch.randn(256, 128, 124, generator=g)
    with torch.no_grad():
        big = net(spec.cuda()).cpu()
        assert torch.isfinite(big).all()
        for lo in (0, 101, 250):
            small = net(spec[lo:lo + 6].cuda()).cpu()
            assert torch.equal(small, big[lo:lo + 6]), f"EmotionNet clips {lo}..: logits depend on the batch"
        ref = oa.emotion_net(sd, spec[[0, 255]])
    assert rel_max(big[[0, 255]], ref) <= 2e-3
    net.cpu()
    torch.cuda.empty_cache()
    cv, csd = build(*CASES["cvae"][:2])
    cv = cv.cuda()
    n = 1_000_000
    x, y, eps = (torch.randn(n, k, generator=g) for k in (90, 90, 32))
    with torch.no_grad():
        out, mu, lv = cv(x.cuda(), y.cuda(), eps=eps.cuda())
        for lo in (0, 499_999, n - 300):
            o2, m2, l2 = cv(x[lo:lo + 300].cuda(), y[lo:lo + 300].cuda(), eps=eps[lo:lo + 300].cuda())
            assert torch.equal(o2, out[lo:lo + 300]) and torch.equal(m2, mu[lo:lo + 300]) and torch.equal(l2, lv[lo:lo + 300])
        idx = torch.tensor([0, 123_456, n - 1])
        r_out, r_mu, r_lv = oa.cvae_forward(csd, x[idx], y[idx], eps[idx])
    assert rel_max(out[idx].cpu(), r_out) <= 2e-5 and rel_max(mu[idx].cpu(), r_mu) <= 2e-5 and rel_max(lv[idx].cpu(), r_lv) <= 2e-5
