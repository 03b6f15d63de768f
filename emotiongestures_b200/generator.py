"""Host-side mirror of the reference generator's nn.Module interface.

`Transformer` keeps the constructor, the `forward` signature, the 5-tuple it
returns and the `state_dict()` key layout of Full_model/Models.py:295-427
(4th forward argument from Full_model/Models_memory.py:521), so checkpoints and
calling code move over unchanged.  The sub-modules here are parameter holders:
all arithmetic of the pose path runs in libegx (hand-written sm_100a CUDA behind
the C ABI of include/egx.h).  There is no PyTorch fallback for that path; without
the library `forward` raises.

The text encoder is the one exception: its output never reaches the poses
(Full_model/Models.py:400,427) and is returned as output[4] only, so it stays a
small PyTorch module that keeps the tuple intact.
"""
from __future__ import annotations

import threading
import warnings
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import GeneratorConfig


class EngineSet:
    """The libegx engines of ONE generator, one per (CUDA device, precision).

    nn.DataParallel (test_emotion_gesture_diversity_iterative.py:137-138) rebuilds its replicas on every call as
    shallow `__dict__` copies of the wrapped module (torch.nn.Module._replicate_for_data_parallel) and runs them in
    one Python thread per device.  This object sits in that `__dict__`, so the original module and every replica
    share it by reference: a replica looks its engine up by the device of the inputs it was handed, the first call
    on a device packs the weights there (from the ORIGINAL module — replicas own no parameters), later calls reuse
    the handle.  libegx keeps no global state, so the per-device handles run concurrently."""

    def __init__(self, source, cfg):
        self._source = weakref.ref(source) if source is not None else None
        self.cfg = cfg
        self.engines = {}
        self.lock = threading.Lock()

    def __reduce__(self):            # pickling / deepcopy of the module: handles are per process, start empty
        return (EngineSet, (None, self.cfg))

    def bind(self, source):
        if self._source is None or self._source() is None:
            self._source = weakref.ref(source)

    def source(self):
        src = self._source() if self._source is not None else None
        if src is None:
            raise RuntimeError("the module these engines were packed from is gone")
        return src

    def get(self, device, precision="tc"):
        from .engine import Engine
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("emotiongestures_b200 runs on sm_100a CUDA devices only (no CPU fallback); the generator "
                               f"and its inputs are on {device}")
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, precision)
        eng = self.engines.get(key)
        if eng is None:
            with self.lock:
                eng = self.engines.get(key)
                if eng is None:
                    eng = Engine(self.cfg, torch.device("cuda", index), precision=precision)
                    eng.load_state_dict(self.source().state_dict())
                    self.engines[key] = eng
        return eng

    def sync(self):
        """Re-pack every live engine from the source module's current weights."""
        with self.lock:
            sd = self.source().state_dict() if self.engines else None
            for eng in self.engines.values():
                eng.load_state_dict(sd)


def _sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """Full_model/Models.py:34-44 — built in float64, stored as float32."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2.0 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.tensor(ang, dtype=torch.float32).unsqueeze(0)


class _PosTable(nn.Module):
    def __init__(self, d_hid, n_position):
        super().__init__()
        self.register_buffer("pos_table", _sinusoid_table(n_position, d_hid))
        self.register_buffer("pos_table2", _sinusoid_table(n_position, d_hid))


class _SE(nn.Module):
    """Full_model/ResNetBlocks.py:81-90 (reduction 8)."""

    def __init__(self, c, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c, c // reduction), nn.ReLU(inplace=True),
                                nn.Linear(c // reduction, c), nn.Sigmoid())


class _SEBlock(nn.Module):
    """Full_model/ResNetBlocks.py:7-19."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.se = _SE(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(cout))
        self.stride = stride


class _Trunk(nn.Module):
    """Full_model/ResNetSE34V2.py:13-55 with layers [3,4,6], filters [32,64,128]."""

    LAYERS = (3, 4, 6)
    FILTERS = (32, 64, 128)

    def __init__(self, layers=None, filters=None):
        super().__init__()
        f = filters or self.FILTERS
        self.conv1 = nn.Conv2d(1, f[0], 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(f[0])
        cin = f[0]
        for li, (n, c) in enumerate(zip(layers or self.LAYERS, f), start=1):
            blocks = []
            for b in range(n):
                blocks.append(_SEBlock(cin, c, 2 if (b == 0 and li > 1) else 1))
                cin = c
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _AudioEncoder(nn.Module):
    """Full_model/Models.py:92-107; fc1 fan-in follows the spectrogram width
    instead of the hard-coded 32*31."""

    def __init__(self, frames, d_model, fc1_in):
        super().__init__()
        self.feat_extractor = _Trunk()
        self.final_conv1 = nn.Conv2d(_Trunk.FILTERS[2], frames, 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(frames)
        self.fc1 = nn.Linear(fc1_in, d_model)
        self.fc2 = nn.Linear(d_model, d_model)


class _TemporalBlock(nn.Module):
    """Full_model/tcn.py:16-46 (weight-normed dilated causal conv pair)."""

    def __init__(self, cin, cout, k, dilation):
        super().__init__()
        pad = (k - 1) * dilation
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            wn = torch.nn.utils.weight_norm
            self.conv1 = wn(nn.Conv1d(cin, cout, k, padding=pad, dilation=dilation))
            self.conv2 = wn(nn.Conv1d(cout, cout, k, padding=pad, dilation=dilation))
        # the reference registers the same convs a second time inside `net`
        # (indices 0 and 4), which doubles their state_dict keys
        self.net = nn.Sequential(self.conv1, nn.Identity(), nn.ReLU(), nn.Identity(),
                                 self.conv2, nn.Identity(), nn.ReLU(), nn.Identity())
        self.downsample = nn.Conv1d(cin, cout, 1) if cin != cout else None
        self.pad = pad
        self.conv1.weight_v.data.normal_(0, 0.01)
        self.conv2.weight_v.data.normal_(0, 0.01)

    def forward(self, x):
        y = F.relu(self.conv1(x)[:, :, :-self.pad])
        y = F.relu(self.conv2(y)[:, :, :-self.pad])
        res = x if self.downsample is None else self.downsample(x)
        return F.relu(y + res)


class _TCN(nn.Module):
    def __init__(self, cin, channels, k):
        super().__init__()
        blocks = []
        for i, c in enumerate(channels):
            blocks.append(_TemporalBlock(cin if i == 0 else channels[i - 1], c, k, 2 ** i))
        self.network = nn.Sequential(*blocks)

    def forward(self, x):
        return self.network(x)


class _TextEncoder(nn.Module):
    """Full_model/Models.py:140-178.  Dead w.r.t. poses; runs in PyTorch."""

    def __init__(self, args, n_words, embed_size, pre_trained_embedding=None, text_len=60):
        super().__init__()
        if pre_trained_embedding is not None:
            self.embedding = nn.Embedding.from_pretrained(
                torch.as_tensor(np.asarray(pre_trained_embedding), dtype=torch.float32),
                freeze=args.freeze_wordembed)
        else:
            self.embedding = nn.Embedding(n_words, embed_size)
        self.tcn = _TCN(embed_size, [args.hidden_size] * args.n_layers, 2)
        self.decoder = nn.Linear(args.hidden_size, 512)
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.normal_(0, 0.01)
        self.fc1 = nn.Sequential(nn.Linear(text_len, text_len))

    def forward(self, tokens):
        y = self.tcn(self.embedding(tokens).transpose(1, 2))
        return self.decoder(self.fc1(y).transpose(1, 2)).contiguous()


class _PriorEncoder(nn.Module):
    """Full_model/Models.py:184-197."""

    def __init__(self, prior_frames, frames, pose_dim, d_model):
        super().__init__()
        self.conv1 = nn.Conv1d(prior_frames, frames, 3, padding=1)
        self.bn1 = nn.BatchNorm1d(frames)
        self.conv2 = nn.Conv1d(frames, frames, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(frames)
        self.fc1 = nn.Linear(pose_dim, d_model)
        self.fc2 = nn.Linear(d_model, d_model)


class _ChunkEncoder(nn.Sequential):
    """Linear -> Dropout(0.2) -> Linear, the Sequential index layout of the memory nets."""

    def __init__(self, d_in, d_out):
        super().__init__(nn.Linear(d_in, d_out), nn.Dropout(0.2), nn.Linear(d_out, d_out))


class _SpatialMemory(nn.Module):
    """Full_model/Models_memory.py:215-231 (SP_Memory_Net_v1 parameters)."""

    def __init__(self, chunk, pose_dim):
        super().__init__()
        self.spatial_chunk_encoder = _ChunkEncoder(chunk * pose_dim, pose_dim)


class _TemporalMemory(nn.Module):
    """Full_model/Models_memory.py:263-281 (TM_Memory_Net parameters)."""

    def __init__(self, chunk, pose_dim):
        super().__init__()
        self.temporal_chunk_encoder = _ChunkEncoder(chunk * pose_dim, pose_dim)
        self.temporal_memory_encoder = _ChunkEncoder(chunk * pose_dim, chunk)


class _PriorMemoryEncoder(nn.Module):
    """Full_model/Models_memory.py:296-345 (parameters only; the arithmetic runs in libegx, k_memory.cu)."""

    def __init__(self, args, prior_frames, frames, pose_dim, d_model):
        super().__init__()
        self.post_header = nn.Sequential(nn.Linear(pose_dim, d_model), nn.Dropout(0.2), nn.Linear(d_model, d_model))
        self.pred_length = frames - prior_frames
        n = self.pred_length
        self.pred_conv = nn.Sequential(
            nn.Conv1d(prior_frames, n, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True), nn.BatchNorm1d(n),
            nn.Conv1d(n, n, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True), nn.BatchNorm1d(n))
        self.spatial_memory = _SpatialMemory(args.chunk, pose_dim)
        self.temporal_memory = _TemporalMemory(args.chunk, pose_dim)


class _MHA(nn.Module):
    """Full_model/SubLayers.py:12-27 (no biases, LayerNorm eps 1e-6)."""

    def __init__(self, n_head, d_model, d_k, d_v):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class _FFN(nn.Module):
    """Full_model/SubLayers.py:67-72."""

    def __init__(self, d_in, d_hid):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.layer_norm = nn.LayerNorm(d_in, eps=1e-6)


class _EncoderLayer(nn.Module):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v):
        super().__init__()
        self.slf_attn = _MHA(n_head, d_model, d_k, d_v)
        self.pos_ffn = _FFN(d_model, d_inner)


class _DecoderLayer(nn.Module):
    """Full_model/Layers.py:41-58: slf_attn owns parameters but is never run."""

    def __init__(self, d_model, d_inner, n_head, d_k, d_v):
        super().__init__()
        self.slf_attn = _MHA(n_head, d_model, d_k, d_v)
        self.enc_attn = _MHA(n_head, d_model, d_k, d_v)
        self.pos_ffn = _FFN(d_model, d_inner)


class _Encoder(nn.Module):
    def __init__(self, d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner, n_position):
        super().__init__()
        self.position_embeddings = nn.Embedding(n_position, d_model)   # unused by forward
        self.position_enc = _PosTable(d_word_vec, n_position)
        self.layer_stack = nn.ModuleList(
            [_EncoderLayer(d_model, d_inner, n_head, d_k, d_v) for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)              # unused by forward


class _Decoder(nn.Module):
    def __init__(self, d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner, n_position):
        super().__init__()
        self.position_enc = _PosTable(d_word_vec, n_position)          # unused by forward
        self.layer_stack = nn.ModuleList(
            [_DecoderLayer(d_model, d_inner, n_head, d_k, d_v) for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)              # unused by forward


class Transformer(nn.Module):
    """Drop-in for Full_model.Models.Transformer (inference path).

    Extra keyword `spec_w` (spectrogram columns, 124 for BEAT / 70 for TED) sizes
    `audio_encoder.fc1`; the reference hard-codes the BEAT value.
    """

    def __init__(self, args, lang_model, frames=60, pose_dim=282, prior_frames=10,
                 src_pad_idx=1, trg_pad_idx=1, d_word_vec=64, d_model=64, d_inner=512,
                 n_layers=3, n_head=8, d_k=32, d_v=32, dropout=0.2, n_position=60, *,
                 spec_w=124, n_audio=None):
        super().__init__()
        assert d_model == d_word_vec, "d_model must equal d_word_vec (residual connections)"
        fps = 15
        if n_audio is None:
            n_audio = int(round(frames / fps * 16000))
        self.cfg = GeneratorConfig(
            frames=frames, prior_frames=prior_frames, pose_dim=pose_dim, d_model=d_model,
            d_inner=d_inner, n_layers=n_layers, n_head=n_head, d_k=d_k, d_v=d_v,
            spec_w=spec_w, n_audio=n_audio, n_position=n_position,
            n_words=lang_model.n_words, wordembed_dim=args.wordembed_dim,
            tcn_hidden=args.hidden_size, tcn_layers=args.n_layers)
        self.cfg.validate()
        self.d_model = d_model
        self.src_pad_idx, self.trg_pad_idx = src_pad_idx, trg_pad_idx
        self.audio_encoder = _AudioEncoder(frames, d_model, self.cfg.fc1_in)
        self.text_encoder = _TextEncoder(args, lang_model.n_words, args.wordembed_dim,
                                         lang_model.word_embedding_weights)
        self.emotion_proj = nn.Sequential(nn.Linear(d_model, d_model), nn.Dropout(0.2),
                                          nn.Linear(d_model, d_model))
        self.emotion_classifer_header = nn.Sequential(
            nn.Linear(frames * d_model, d_model), nn.ReLU(True), nn.Linear(d_model, 256),
            nn.ReLU(True), nn.Linear(256, 64), nn.ReLU(True), nn.Linear(64, 8))
        self.semantic_proj = nn.Sequential(nn.Linear(d_model, d_model), nn.Dropout(0.2),
                                           nn.Linear(d_model, d_model))
        self.fusion_proj = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(True),
                                         nn.Linear(d_model, d_model))
        self.prior_seq_encoder = _PriorEncoder(prior_frames, frames, pose_dim, d_model)
        self.post_projector = nn.Sequential(
            nn.Linear(d_model, d_model * 4), nn.Dropout(0.2), nn.Linear(d_model * 4, d_model),
            nn.Dropout(0.2), nn.Linear(d_model, pose_dim), nn.Dropout(0.2),
            nn.Linear(pose_dim, pose_dim))
        self.dropout = nn.Dropout(p=dropout)
        self.encoder = _Encoder(d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner,
                                n_position)
        self.decoder = _Decoder(d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner,
                                n_position)
        # Full_model/Models.py:381-383: xavier over every parameter with dim > 1
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._egx_set = EngineSet(self, self.cfg)

    @classmethod
    def from_config(cls, cfg: GeneratorConfig, dropout=0.1):
        class _Args:
            freeze_wordembed = False
            hidden_size = cfg.tcn_hidden
            n_layers = cfg.tcn_layers
            wordembed_dim = cfg.wordembed_dim
            dropout_prob = 0.1

        class _Lang:
            n_words = cfg.n_words
            word_embedding_weights = None

        return cls(_Args(), _Lang(), frames=cfg.frames, pose_dim=cfg.pose_dim,
                   prior_frames=cfg.prior_frames, d_word_vec=cfg.d_model, d_model=cfg.d_model,
                   d_inner=cfg.d_inner, n_layers=cfg.n_layers, n_head=cfg.n_head, d_k=cfg.d_k,
                   d_v=cfg.d_v, dropout=dropout, n_position=cfg.n_position, spec_w=cfg.spec_w,
                   n_audio=cfg.n_audio)

    # -- engine plumbing -------------------------------------------------
    def engine(self, precision="tc", device=None):
        """The libegx engine holding this module's weights on `device` (default: where the parameters live).
        nn.DataParallel replicas own no parameters and pass the device of their inputs."""
        if not getattr(self, "_is_replica", False):
            self._egx_set.bind(self)
        if device is None:
            device = next(self.parameters()).device
        return self._egx_set.get(device, precision)

    def sync_weights(self):
        """Re-pack weights after `load_state_dict` / in-place edits (SURVEY.md §5 checkpoint row)."""
        self._egx_set.bind(self)
        self._egx_set.sync()

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.sync_weights()
        return out

    def forward(self, input_spectrum, text, prior_seq, sampled_emotion_feature=None):
        """(B,128,W) f32, (B,60) i64, (B,p,P) f32[, (B,F,d) f32] -> the reference 5-tuple
        (Full_model/Models.py:389-427; emotion branch Models_memory.py:551-555)."""
        if self.training:
            raise RuntimeError(
                "emotiongestures_b200.Transformer is the inference path "
                "(the reference released no generator training code); call .eval() first")
        eng = self.engine(getattr(self, "precision", "tc"), _call_device(self, input_spectrum))
        text_embedding = self.text_encoder(text)
        poses, emo, sem, logits = eng.generator_forward(
            input_spectrum, prior_seq, sampled_emotion_feature)
        return poses, emo, sem, logits, text_embedding

    def forward_audio(self, audio, text, prior_seq, sampled_emotion_feature=None, *,
                      mode=None, preemph=False):
        """Raw 16 kHz audio (B,N) -> log-mel on the GPU -> forward (F1–F4 + generator).  By default the features are
        the ones the reference's checkpoints were trained on (config.LOGMEL_REFERENCE: no pre-emphasis, dB with
        ref=max, fp16 storage rounding); pass `mode=LOGMEL_LOG_IN, preemph=True` for the ResNetSE34V2 recipe."""
        from .config import LOGMEL_REFERENCE
        eng = self.engine(getattr(self, "precision", "tc"), _call_device(self, audio))
        spec = eng.logmel(audio, LOGMEL_REFERENCE if mode is None else mode, preemph)
        return self.forward(spec, text, prior_seq, sampled_emotion_feature)


def _call_device(module, x):
    """Device a forward call runs on: a DataParallel replica owns no parameters (they are plain attributes there), but
    its inputs were scattered to its device; otherwise the module's own device."""
    if getattr(module, "_is_replica", False) and isinstance(x, torch.Tensor) and x.is_cuda:
        return x.device
    return None


class MemoryTransformer(Transformer):
    """Drop-in for Full_model.Models_memory.Transformer (:428-560), the generator the evaluation script really loads
    (test_emotion_gesture_diversity_iterative.py:25,135): identical to ``Transformer`` except that the prior-pose
    encoder is ``Prior_MemoryEncoder`` (``args.chunk`` frames of spatial / temporal memory).  Its temporal memory
    sums over the batch of the call (Models_memory.py:287-288), so — exactly as in the reference — a clip's poses
    depend on which other clips share its batch / DataParallel replica."""

    def __init__(self, args, lang_model, *a, **k):
        super().__init__(args, lang_model, *a, **k)
        c = self.cfg
        self.prior_seq_encoder = _PriorMemoryEncoder(args, c.prior_frames, c.frames, c.pose_dim, c.d_model)
        for p in self.prior_seq_encoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    @classmethod
    def from_config(cls, cfg: GeneratorConfig, chunk: int = 4, dropout=0.1):
        class _Args:
            freeze_wordembed = False
            hidden_size = cfg.tcn_hidden
            n_layers = cfg.tcn_layers
            wordembed_dim = cfg.wordembed_dim
            dropout_prob = 0.1

        class _Lang:
            n_words = cfg.n_words
            word_embedding_weights = None

        _Args.chunk = chunk
        return cls(_Args(), _Lang(), frames=cfg.frames, pose_dim=cfg.pose_dim,
                   prior_frames=cfg.prior_frames, d_word_vec=cfg.d_model, d_model=cfg.d_model,
                   d_inner=cfg.d_inner, n_layers=cfg.n_layers, n_head=cfg.n_head, d_k=cfg.d_k,
                   d_v=cfg.d_v, dropout=dropout, n_position=cfg.n_position, spec_w=cfg.spec_w,
                   n_audio=cfg.n_audio)


def randomize_norm_stats_(module: nn.Module, seed: int = 1) -> None:
    """Give BatchNorm non-trivial running stats/affine so eval-mode BN is exercised
    (at default init BN is an almost-identity and would hide bugs)."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)

