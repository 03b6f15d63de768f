"""CPU: host-side mirror of the reference interface (constructor, state_dict, text encoder,
error behaviour), geometry helpers and sharding arithmetic."""
import pytest
import torch

from emotiongestures_b200 import (BEAT, TED, GeneratorConfig, Transformer, audio_length,
                                  randomize_norm_stats_, spectrogram_length)
from emotiongestures_b200.sharding import shard_bounds
from tests.helpers import model_and_sd


def test_geometry_matches_reference_formulas():
    # utils/data_utils.py:42-44 and lmdb_data_loader_expressive.py:95
    assert spectrogram_length(34, 15) == 70 and spectrogram_length(60, 15) == 124
    assert audio_length(34, 15) == 36267 and audio_length(60, 15) == 64000
    assert TED.trunk_hw == (32, 18) and BEAT.trunk_hw == (32, 31)
    assert TED.n_stft_frames == 71 and BEAT.n_stft_frames == 126
    with pytest.raises(ValueError):
        GeneratorConfig(spec_w=80).validate()
    with pytest.raises(ValueError):
        GeneratorConfig(frames=61).validate()


def test_constructor_signature_is_the_reference_one():
    class Args:
        freeze_wordembed = False; hidden_size = 300; n_layers = 3; wordembed_dim = 300; dropout_prob = 0.1

    class Lang:
        n_words = 50; word_embedding_weights = None

    # positional order of Full_model/Models.py:298-301
    g = Transformer(Args(), Lang(), 34, 126, 4, 1, 1, 256, 256, 1024, 3, 8, 64, 64, 0.1, 60, spec_w=70)
    assert g.cfg.frames == 34 and g.cfg.fc1_in == 576 and g.text_encoder.embedding.num_embeddings == 50
    with pytest.raises(AssertionError):
        Transformer(Args(), Lang(), d_word_vec=64, d_model=128)


def test_state_dict_roundtrip_and_init():
    m, sd = model_and_sd("ted", 0)
    m2 = Transformer.from_config(TED)
    missing = m2.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # Full_model/Models.py:381-383: xavier_uniform over every parameter with dim > 1
    torch.manual_seed(0)
    fresh = Transformer.from_config(TED)
    w = fresh.audio_encoder.feat_extractor.layer1[0].conv1.weight
    bound = (6.0 / (32 * 9 + 32 * 9)) ** 0.5
    assert w.abs().max() <= bound + 1e-6 and w.abs().max() > 0.8 * bound
    randomize_norm_stats_(fresh, 1)
    assert (fresh.audio_encoder.bn1.running_var != 1).all()


def test_text_encoder_runs_on_cpu_and_is_causal():
    m, _ = model_and_sd("ted", 0)
    t = torch.randint(0, 100, (2, 60))
    with torch.no_grad():
        out = m.text_encoder(t)
    assert out.shape == (2, 60, 512)


def test_training_mode_is_rejected():
    m = Transformer.from_config(TED)
    with pytest.raises(RuntimeError, match="inference path"):
        m.train()(torch.zeros(1, 128, 70), torch.zeros(1, 60, dtype=torch.int64), torch.zeros(1, 4, 126))


def test_shard_bounds_partition_the_batch():
    for n, world in ((4096, 8), (10, 4), (3, 8), (0, 2), (7, 1)):
        spans = [shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_dropin_reads_the_geometry_of_both_generator_variants():
    """install() derives egx_cfg from a live module's state_dict: Models.Transformer and Models_memory.Transformer
    (whose prior conv maps p -> F - p frames, Full_model/Models_memory.py:321-329)."""
    from emotiongestures_b200 import BEAT, TED, MemoryTransformer, Transformer
    from emotiongestures_b200.dropin import config_from_module
    for cfg, m in ((TED, Transformer.from_config(TED)), (BEAT, Transformer.from_config(BEAT)),
                   (TED, MemoryTransformer.from_config(TED, 4)), (BEAT, MemoryTransformer.from_config(BEAT, 10))):
        got = config_from_module(m)
        for f in ("frames", "prior_frames", "pose_dim", "d_model", "d_inner", "n_layers", "n_head", "d_k", "d_v", "spec_w",
                  "n_position"):
            assert getattr(got, f) == getattr(cfg, f), (type(m).__name__, f)


def test_diversity_score_matches_the_reference_golden():
    """evaluate.diversity_score vs model/FHD_score.py:244-280 run on the same array with np.random.seed(123)
    (tests/golden/diversity.npz, made in the build container against the real reference)."""
    import numpy as np
    from emotiongestures_b200.evaluate import diversity_score
    from tests.helpers import load_golden
    g = load_golden("diversity")
    x = np.random.default_rng(int(g["data_seed"])).standard_normal((int(g["n"]) * 60, 512)).astype(np.float32)
    score, (lo, hi) = diversity_score(x, np.random.RandomState(int(g["seed"])))
    assert np.allclose(score, g["score"], rtol=1e-6) and np.allclose(lo, g["lo"], rtol=1e-6) and np.allclose(hi, g["hi"], rtol=1e-6)


def test_frechet_distance_device_variant_matches_the_host_tail():
    """fgd.frechet_distance_device (torch.linalg.eigh, float64; runs on the CPU here, on the GPU in the evaluator)
    == fgd.frechet_distance (numpy) == the closed form for commuting covariances, incl. the rank-deficient case."""
    import numpy as np
    import torch
    from emotiongestures_b200 import fgd
    rng = np.random.default_rng(3)
    for n, d in ((400, 16), (10, 16)):                      # full rank, and fewer samples than dimensions
        a, b = rng.standard_normal((n, d)), rng.standard_normal((n, d)) * 1.5 + 0.3
        m1, s1, m2, s2 = a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False)
        host = fgd.frechet_distance(m1, s1, m2, s2)
        dev = fgd.frechet_distance_device(*(torch.from_numpy(x) for x in (m1, s1, m2, s2)))
        assert abs(host - dev) <= 1e-8 * max(1.0, abs(host))
    s = np.diag([1.0, 4.0, 9.0])
    want = 3 * 0.25 + (1 + 4 + 9) * (1 + 4 - 2 * 2)             # mu diff 0.5 each; Tr(S + 4S - 2 sqrt(4 S^2)) = Tr S
    got = fgd.frechet_distance_device(torch.zeros(3), torch.from_numpy(s), torch.full((3,), 0.5), torch.from_numpy(4 * s))
    assert abs(got - want) <= 1e-9
    acc = torch.zeros(1 + 3 + 9, dtype=torch.float64)
    x = torch.from_numpy(rng.standard_normal((50, 3)))
    acc[0], acc[1:4], acc[4:] = 50, x.sum(0), (x.T @ x).reshape(-1)
    mu, sig = fgd.finalize_stats_device(acc, 3)
    assert torch.allclose(mu, x.mean(0)) and torch.allclose(sig, torch.from_numpy(np.cov(x.numpy(), rowvar=False)))


def test_install_fails_loudly_on_cpu_and_leaves_the_module_untouched():
    """No CPU fallback: install() on a CPU module raises and does not swap anything."""
    from emotiongestures_b200.dropin import install
    m = Transformer.from_config(TED).eval()
    cls = type(m)
    with pytest.raises(RuntimeError, match="sm_100a CUDA devices only"):
        install(m)
    assert type(m) is cls and "egx_engine" not in m.__dict__


def test_engine_set_is_shared_with_dataparallel_replicas():
    """nn.DataParallel replicas are shallow __dict__ copies (torch Module._replicate_for_data_parallel): they must see
    the SAME per-device engine table as the original and pick their engine by the device of their inputs."""
    from emotiongestures_b200.generator import EngineSet, _call_device
    m = Transformer.from_config(TED).eval()
    rep = m._replicate_for_data_parallel()
    assert rep._egx_set is m._egx_set and isinstance(rep._egx_set, EngineSet)
    assert rep._is_replica and _call_device(m, torch.zeros(1)) is None
    assert type(rep).forward is type(m).forward          # class-level forward: `self` is the replica, not the original
    import pickle
    m2 = pickle.loads(pickle.dumps(m))                   # handles are per process: a copy starts with an empty table
    assert m2._egx_set.engines == {} and m2._egx_set is not m._egx_set
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m._egx_set.get(torch.device("cpu"))


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/Full_model"), reason="needs the reference tree (build container)")
def test_dropin_reads_the_geometry_of_the_real_reference_modules():
    """config_from_module / install() on the REAL Full_model.Models.Transformer and Models_memory.Transformer
    (CPU, build container only; the GPU box has no reference tree)."""
    from emotiongestures_b200.dropin import config_from_module, install
    from oracle.make_golden import load_reference
    for cfg, chunk in ((TED, 0), (BEAT, 0), (TED, 4)):
        ref = load_reference(cfg, chunk)
        got = config_from_module(ref)
        for f in ("frames", "prior_frames", "pose_dim", "d_model", "d_inner", "n_layers", "n_head", "d_k", "d_v", "spec_w",
                  "n_position"):
            assert getattr(got, f) == getattr(cfg, f), (chunk, f)
        cls = type(ref)
        with pytest.raises(RuntimeError, match="sm_100a CUDA devices only"):
            install(torch.nn.DataParallel(ref) if chunk else ref)
        assert type(ref) is cls


def test_bind_host_to_gpu_node_is_best_effort():
    """No GPU / no PCI topology: the helper reports None and leaves the affinity alone."""
    import os
    from emotiongestures_b200.sharding import bind_host_to_gpu_node
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert bind_host_to_gpu_node(0) is None
        assert os.sched_getaffinity(0) == before
