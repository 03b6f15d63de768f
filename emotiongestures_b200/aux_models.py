"""Host mirrors of the small networks either side of the generator.

Each class keeps the reference module's constructor, attribute names and therefore its exact
``state_dict`` layout, so reference checkpoints load unchanged; the arithmetic runs in libegx
(`csrc/k_aux.cu`).  As for the generator there is no CPU / PyTorch fallback: calling a forward on
a module that is not on an sm_100a device raises.

The reference draws Gaussian noise inside the modules (``torch.randn`` / ``randn_like``).  The
mirrors accept the draw as an optional argument (``eps=`` / ``z=``); when it is omitted they draw
it with ``torch.randn`` on the module's device exactly where the reference does, so seeding with
``torch.manual_seed`` keeps working, while parity tests and multi-GPU runs pass the noise in
(indexed by global clip id — SURVEY.md §8(e)).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import TED
from .engine import Engine, _ptr


def _chain(dims, drop=0.2):
    """Linear, Dropout, Linear, ... — the Sequential index layout the reference checkpoints use."""
    mods = []
    for i in range(len(dims) - 1):
        if i:
            mods.append(nn.Dropout(drop))
        mods.append(nn.Linear(dims[i], dims[i + 1]))
    return nn.Sequential(*mods)


class _DeviceModule(nn.Module):
    """Shared plumbing: one Engine per device, weights re-sent when the state_dict changes."""

    _family = ""

    def _engine(self) -> Engine:
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} runs on sm_100a CUDA devices only (no CPU fallback); "
                               f"module is on {p.device}")
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: inference only — call .eval() first")
        key = (p.device, tuple(t._version for t in self.state_dict().values()))
        cache = self.__dict__.setdefault("_egx_cache", {})
        if cache.get("key") != key:
            eng = cache.get("eng")
            if eng is None or eng.device != p.device:
                eng = Engine(TED, p.device)          # geometry is irrelevant for the small networks
            eng.load_state_dict({self._family + k: v for k, v in self.state_dict().items()})
            cache.update(key=key, eng=eng)
        return cache["eng"]


class MLP_Reconstruct(_DeviceModule):
    """Full_model/BEAT_CVAE.py:32-136 — MLP CVAE over 90-D hand poses (row C4)."""

    _family = "cvae."

    def __init__(self, bath=True):
        super().__init__()
        self.Encoder = _chain([90, 128, 128, 256, 256, 512])
        self.Posterior_Y_embedding = _chain([90, 64, 32])
        self.fc_mu = nn.Linear(512, 32)
        self.fc_var = nn.Linear(512, 32)
        self.Decoder = _chain([512, 256, 256, 128, 128, 90])
        self.fusion_z_posterior = _chain([64, 256, 512])

    def forward(self, Input, y, eps=None):
        eng = self._engine()
        x, y = eng._f32(Input, "Input"), eng._f32(y, "y")
        n = x.shape[0]
        if tuple(x.shape) != (n, 90) or tuple(y.shape) != (n, 90):
            raise RuntimeError("Input and y must be (N, 90)")
        eps = torch.randn(n, 32, device=eng.device) if eps is None else eng._f32(eps, "eps", (n, 32))
        out = torch.empty((n, 90), device=eng.device)
        mu, log_var = torch.empty((n, 32), device=eng.device), torch.empty((n, 32), device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_cvae_forward(eng._h, _ptr(x), _ptr(y), _ptr(eps), n, _ptr(out), _ptr(mu),
                                                _ptr(log_var), eng._stream()), "egx_cvae_forward")
        return out, mu, log_var

    def sample(self, y, z=None):
        eng = self._engine()
        y = eng._f32(y, "y")
        n = y.shape[0]
        z = torch.randn(n, 32, device=eng.device) if z is None else eng._f32(z, "z", (n, 32))
        out = torch.empty((n, 90), device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_cvae_sample(eng._h, _ptr(y), _ptr(z), n, _ptr(out), eng._stream()),
                       "egx_cvae_sample")
        return out


class MLP_Reconstruct_v3(_DeviceModule):
    """CAVE/BEAT_CVAE.py:313-447 — convolutional CVAE whose ``sample`` produces the BEAT generator's
    ``sampled_emotion_feature`` (row E1).  Only the sampler half runs on the device path; the
    encoder half exists so that checkpoints load (its forward is training-side)."""

    _family = "cvae3."

    def __init__(self, bath=True):
        super().__init__()
        act = lambda: nn.LeakyReLU(0.2, True)                              # noqa: E731
        self.Encoder = nn.Sequential(
            nn.Conv1d(60, 32, 3, padding=1), act(), nn.BatchNorm1d(32),
            nn.Conv1d(32, 16, 3, padding=1), act(), nn.BatchNorm1d(16),
            nn.Conv1d(16, 8, 5, stride=2, padding=2), act(), nn.BatchNorm1d(8),
            nn.Conv1d(8, 4, 5, stride=2, padding=2), act(), nn.BatchNorm1d(4))
        self.Posterior_Y_embedding = _chain([8, 16, 32])
        self.fc_mu = _chain([4 * 128, 128, 32])
        self.fc_var = _chain([4 * 128, 128, 32])
        self.Decoder = nn.Sequential(
            nn.ConvTranspose1d(4, 8, kernel_size=3, stride=2, padding=1, output_padding=1), act(), nn.BatchNorm1d(8),
            nn.ConvTranspose1d(8, 16, kernel_size=3, stride=2, padding=1, output_padding=1), act(), nn.BatchNorm1d(16),
            nn.Conv1d(16, 32, 3, padding=1), act(), nn.BatchNorm1d(32),
            nn.Conv1d(32, 60, 3, padding=1), act(), nn.BatchNorm1d(60),
            nn.Conv1d(60, 60, 3, padding=1))
        self.fusion_z_posterior = _chain([64, 128, 4 * 128])

    def sample(self, y, z=None):
        eng = self._engine()
        y = eng._f32(y, "y")
        n = y.shape[0]
        if tuple(y.shape) != (n, 8):
            raise RuntimeError("y must be (N, 8) one-hot emotion labels")
        z = torch.randn(n, 32, device=eng.device) if z is None else eng._f32(z, "z", (n, 32))
        out = torch.empty((n, 60, 512), device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_cvae3_sample(eng._h, _ptr(y), _ptr(z), n, _ptr(out), eng._stream()),
                       "egx_cvae3_sample")
        return out

    def forward(self, Input, y):
        raise RuntimeError("MLP_Reconstruct_v3.forward is the training-side path; the B200 path implements "
                           ".sample (CAVE/BEAT_CVAE.py:427-447)")


class FGDNet(_DeviceModule):
    """model/FGD.py:26-82 (``MLP_Reconstruct``): per-frame auto-encoder whose 512-D latent feeds the
    FGD statistics (test_emotion_gesture_diversity_iterative.py:226-229).  ``forward`` returns
    ``(None, latent)`` — the evaluation loop discards the reconstruction."""

    _family = "fgd_mlp."

    def __init__(self, bath=True, pose_dim=282, hidden=512):
        super().__init__()
        self.Encoder = _chain([pose_dim, hidden, hidden, hidden])
        self.Decoder = _chain([hidden, hidden, hidden, pose_dim])

    def forward(self, Input):
        eng = self._engine()
        x = eng._f32(Input, "Input")
        lead, d = x.shape[:-1], x.shape[-1]
        rows = x.reshape(-1, d)
        n = rows.shape[0]
        hidden = self.Encoder[0].out_features
        out = torch.empty((n, hidden), device=eng.device)
        ws = torch.empty(int(eng.lib.egx_row_features_workspace(eng._h, n)), dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_row_features(eng._h, _ptr(rows), n, d, _ptr(out), _ptr(ws), ws.numel(),
                                                eng._stream()), "egx_row_features")
        return None, out.view(*lead, hidden)


class EmotionNet(_DeviceModule):
    """model/audio_emotion_classifer.py:17-49 — the audio emotion classifier (row C3): a four-stage
    SE-ResNet ([3,4,6,3] blocks, 32..256 filters; model/emotion_ResNetSE34V2.py) over a (B,128,124)
    log-mel, then Linear+ReLU x5 and ``last_fc``.  Returns logits (the reference leaves ``acn`` off)."""

    _family = "emotion_net."

    def __init__(self):
        super().__init__()
        from .generator import _Trunk
        self.emotion_encoder = _Trunk(layers=(3, 4, 6, 3), filters=(32, 64, 128, 256))
        dims = [256 * 16 * 16, 4096, 2048, 512, 128, 64]
        mods = []
        for i in range(5):
            mods += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU(True)]
        self.emotion_eocder_fc = nn.Sequential(*mods)
        self.last_fc = nn.Linear(64, 8)
        self.acn = nn.Softmax(dim=1)

    def forward(self, mfcc):
        eng = self._engine()
        x = eng._f32(mfcc, "mfcc")
        if x.dim() != 3:
            raise RuntimeError("mfcc must be (B, n_mels, W)")
        b, n_mels, w = x.shape
        logits = torch.empty((b, 8), device=eng.device)
        ws = torch.empty(int(eng.lib.egx_emotion_net_workspace(eng._h, b, n_mels, w)), dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_emotion_net_forward(eng._h, _ptr(x), b, n_mels, w, _ptr(logits), _ptr(ws), ws.numel(),
                                                       eng._stream()), "egx_emotion_net_forward")
        return logits


class _SkeletonPriorEncoder(nn.Module):
    """skeleton_classifer/Models.py:87-121 (only fc1 -> Dropout -> fc2 is live)."""

    def __init__(self, pose_dim, d_model):
        super().__init__()
        self.fc1 = nn.Linear(pose_dim, d_model)
        self.fc2 = nn.Linear(d_model, d_model)


class SkeletonClassifier(_DeviceModule):
    """skeleton_classifer/Models.py:199-283 ``Transformer`` — the Emotion-ACC classifier the evaluation loop runs
    on every generated batch (test_emotion_gesture_diversity_iterative.py:158,217).
    Same constructor and state_dict layout; ``forward(prior_seq (B, n_position, pose_dim))`` returns
    ``(logits (B, class_dim), mid_feature (B, n_position, d_model))`` like the reference.  The tcgen05 attention
    kernel is specialised for d_k = d_v = 64, which is how the evaluation script builds the classifier."""

    _family = "skel."

    def __init__(self, class_dim=8, pose_dim=242, src_pad_idx=1, trg_pad_idx=1, d_word_vec=64, d_model=64,
                 d_inner=512, n_layers=3, n_head=8, d_k=32, d_v=32, dropout=0.2, n_position=60):
        super().__init__()
        from .generator import _Encoder
        assert d_model == d_word_vec
        self.d_model = d_model
        self.src_pad_idx, self.trg_pad_idx = src_pad_idx, trg_pad_idx
        self.prior_seq_encoder = _SkeletonPriorEncoder(pose_dim, d_model)
        self.post_projector = nn.Sequential(
            nn.Linear(n_position * d_model, d_model * 4), nn.ReLU(True), nn.Linear(d_model * 4, d_model), nn.ReLU(True),
            nn.Linear(d_model, 128), nn.ReLU(True), nn.Linear(128, 64), nn.ReLU(True), nn.Linear(64, class_dim))
        self.dropout = nn.Dropout(p=dropout)
        self.encoder = _Encoder(d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner, n_position)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._dk = (d_k, d_v)

    def forward(self, prior_seq):
        if self._dk != (64, 64):
            raise RuntimeError("SkeletonClassifier: the sm_100a attention kernel needs d_k = d_v = 64 "
                               "(test_emotion_gesture_diversity_iterative.py:158 builds it that way)")
        eng = self._engine()
        x = eng._f32(prior_seq, "prior_seq")
        if x.dim() != 3:
            raise RuntimeError("prior_seq must be (B, n_frames, pose_dim)")
        b, t, p = x.shape
        n_class = self.post_projector[8].out_features
        logits = torch.empty((b, n_class), device=eng.device)
        mid = torch.empty((b, t, self.d_model), device=eng.device)
        if b == 0:
            return logits, mid
        ws = torch.empty(int(eng.lib.egx_skeleton_workspace(eng._h, b)), dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_skeleton_forward(eng._h, _ptr(x), b, t, p, _ptr(logits), _ptr(mid), _ptr(ws), ws.numel(),
                                                    eng._stream()), "egx_skeleton_forward")
        return logits, mid


def _conv_norm_relu(cin, cout, downsample=False):
    k, s = (4, 2) if downsample else (3, 1)
    return nn.Sequential(nn.Conv1d(cin, cout, kernel_size=k, stride=s), nn.BatchNorm1d(cout), nn.LeakyReLU(0.2, True))


class _PoseEncoderBase(_DeviceModule):
    _kind = 0

    def _build(self, dim, flat, latent):
        self.net = nn.Sequential(_conv_norm_relu(dim, 32), _conv_norm_relu(32, 64), _conv_norm_relu(64, 64, True),
                                 nn.Conv1d(64, 32, 3))
        # nn.LeakyReLU(True): the reference passes True as negative_slope, i.e. slope 1.0 (identity)
        self.out_net = nn.Sequential(nn.Linear(flat, 256), nn.BatchNorm1d(256), nn.LeakyReLU(True),
                                     nn.Linear(256, 128), nn.BatchNorm1d(128), nn.LeakyReLU(True),
                                     nn.Linear(128, latent))

    def _features(self, poses):
        eng = self._engine()
        x = eng._f32(poses, "poses")
        if x.dim() != 3:
            raise RuntimeError("poses must be (B, frames, pose_dim)")
        b, length, dim = x.shape
        n_out = int(eng.lib.egx_pose_feature_dim(eng._h, self._kind))
        out = torch.empty((b, n_out), device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_pose_features(eng._h, self._kind, _ptr(x), b, length, dim, _ptr(out), eng._stream()),
                       "egx_pose_features")
        return out


class MotionAEEncoder(_PoseEncoderBase):
    """``MotionAE.encoder`` = ``PoseEncoderConv(34, pose_dim, latent_dim)`` (model/motion_ae.py:55-83)."""

    _family = "motion_ae.encoder."
    _kind = 0

    def __init__(self, length, pose_dim, latent_dim):
        super().__init__()
        if length != 34:
            raise ValueError("model/motion_ae.py fixes out_net.0 to 384 inputs, i.e. 34-frame clips")
        self._build(pose_dim, 384, latent_dim)

    def forward(self, poses):
        return self._features(poses)


class MotionAE(nn.Module):
    """model/motion_ae.py:117-130.  Only ``encoder`` carries device arithmetic (FGD features); the decoder's
    parameters are declared so that reference checkpoints load with ``strict=True``."""

    def __init__(self, pose_dim, latent_dim):
        super().__init__()
        self.encoder = MotionAEEncoder(34, pose_dim, latent_dim)
        self.decoder = _MotionAEDecoderParams(pose_dim, latent_dim)

    def forward(self, pose):
        pose = pose.view(pose.size(0), pose.size(1), -1)
        return None, self.encoder(pose)


class _MotionAEDecoderParams(nn.Module):
    """Parameter container for ``PoseDecoderConv(34, pose_dim, latent_dim)`` (model/motion_ae.py:85-136)."""

    def __init__(self, pose_dim, latent_dim):
        super().__init__()
        self.pre_net = nn.Sequential(nn.Linear(latent_dim, 64), nn.BatchNorm1d(64), nn.LeakyReLU(True), nn.Linear(64, 136))
        self.net = nn.Sequential(nn.ConvTranspose1d(4, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.ConvTranspose1d(32, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.Conv1d(32, 32, 3), nn.Conv1d(32, pose_dim, 3))


class PoseEncoderConv(_PoseEncoderBase):
    """model/embedding_net.py:37-83: ``forward(poses, variational_encoding)`` -> ``(z, mu, logvar)``.
    The FGD evaluator uses ``mu`` (variational_encoding=False); ``logvar`` is returned as None."""

    _family = "pose_enc."
    _kind = 1

    def __init__(self, length, dim):
        super().__init__()
        flat = 32 * (((length - 4 - 4) // 2 + 1) - 2)
        if flat != 800:
            raise ValueError("model/embedding_net.py fixes out_net.0 to 800 inputs, i.e. 60-frame clips")
        self._build(dim, flat, 32)
        self.fc_mu = nn.Linear(32, 32)
        self.fc_logvar = nn.Linear(32, 32)

    def forward(self, poses, variational_encoding=False):
        if variational_encoding:
            raise RuntimeError("variational_encoding=True is the training-side path")
        mu = self._features(poses)
        return mu, mu, None


__all__ = ["MLP_Reconstruct", "MLP_Reconstruct_v3", "FGDNet", "EmotionNet", "MotionAE", "MotionAEEncoder", "PoseEncoderConv"]
