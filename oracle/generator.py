"""ORACLE (test infrastructure, not product code): CPU restatement of the reference
generator forward, written against a plain `state_dict`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  Pinned: oracle/make_golden.py imports the real
reference from /root/reference in the build container, loads the same synthetic
state_dict into it, asserts this restatement agrees (fp32, <= 2e-6 max-abs
relative to the output scale) and commits its outputs under tests/golden/.

Every function cites the reference lines it follows (paths under the reference
repo root).  The arithmetic is torch fp32 by default; `dtype=torch.float64`
gives the noise-floor reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5     # torch.nn.BatchNorm default (the reference never overrides it)
LN_EPS = 1e-6     # Full_model/SubLayers.py:27,71


class Taps(dict):
    """Optional capture of intermediates (name -> tensor) for golden vectors."""


def _bn(sd, pre, x):
    """Eval-mode BatchNorm over channel dim 1."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    scale = sd[pre + ".weight"] / torch.sqrt(sd[pre + ".running_var"] + BN_EPS)
    shift = sd[pre + ".bias"] - sd[pre + ".running_mean"] * scale
    return x * scale.view(shape) + shift.view(shape)


def _linear(sd, pre, x):
    return F.linear(x, sd[pre + ".weight"], sd.get(pre + ".bias"))


def se_layer(sd, pre, x):
    """Full_model/ResNetBlocks.py:92-95."""
    s = x.mean(dim=(2, 3))
    g = torch.sigmoid(_linear(sd, pre + ".fc.2", F.relu(_linear(sd, pre + ".fc.0", s))))
    return x * g[:, :, None, None]


def se_basic_block(sd, pre, x, stride):
    """Full_model/ResNetBlocks.py:21-37: conv-ReLU-BN-conv-BN-SE-(+res)-ReLU."""
    out = F.conv2d(x, sd[pre + ".conv1.weight"], None, stride=stride, padding=1)
    out = _bn(sd, pre + ".bn1", F.relu(out))
    out = F.conv2d(out, sd[pre + ".conv2.weight"], None, stride=1, padding=1)
    out = _bn(sd, pre + ".bn2", out)
    out = se_layer(sd, pre + ".se", out)
    res = x
    if pre + ".downsample.0.weight" in sd:
        res = _bn(sd, pre + ".downsample.1",
                  F.conv2d(x, sd[pre + ".downsample.0.weight"], None, stride=stride))
    return F.relu(out + res)


def trunk(sd, pre, x, layers=(3, 4, 6), taps=None):
    """Full_model/ResNetSE34V2.py:62-74."""
    x = F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], padding=1)
    x = _bn(sd, pre + ".bn1", F.relu(x))
    if taps is not None:
        taps["stem"] = x
    for li, n in enumerate(layers, start=1):
        for b in range(n):
            x = se_basic_block(sd, f"{pre}.layer{li}.{b}", x, 2 if (b == 0 and li > 1) else 1)
        if taps is not None:
            taps[f"layer{li}"] = x
    return x


def audio_encoder(sd, spec, taps=None):
    """Full_model/Models.py:118-133 (input already has the channel dim)."""
    pre = "audio_encoder"
    x = trunk(sd, pre + ".feat_extractor", spec, taps=taps)
    x = F.conv2d(x, sd[pre + ".final_conv1.weight"], sd[pre + ".final_conv1.bias"], padding=1)
    x = _bn(sd, pre + ".bn1", x)
    b, f = x.shape[:2]
    x = x.reshape(b, f, -1)
    return _linear(sd, pre + ".fc2", _linear(sd, pre + ".fc1", x))


def prior_encoder(sd, prior):
    """Full_model/Models.py:199-212: Conv1d runs along the pose axis, channels = frames."""
    pre = "prior_seq_encoder"
    x = F.conv1d(prior, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], padding=1)
    x = _bn(sd, pre + ".bn1", F.relu(x))
    x = F.conv1d(x, sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], padding=1)
    x = _bn(sd, pre + ".bn2", F.relu(x))
    return _linear(sd, pre + ".fc2", _linear(sd, pre + ".fc1", x))


def prior_memory_encoder(sd, prior):
    """Full_model/Models_memory.py:335-345 with SP_Memory_Net_v1.forward (:233-249) and TM_Memory_Net.forward
    (:282-293).  The reference's B x chunk Python loop of torch.mm is one batched dot product here."""
    pre = "prior_seq_encoder."
    b, p, P = prior.shape
    chunk = sd[pre + "spatial_memory.spatial_chunk_encoder.0.weight"].shape[1] // P

    def chain(name, x):
        return _linear(sd, pre + name + ".2", _linear(sd, pre + name + ".0", x))

    x = F.conv1d(prior, sd[pre + "pred_conv.0.weight"], sd[pre + "pred_conv.0.bias"], padding=1)
    x = _bn(sd, pre + "pred_conv.2", F.relu(x))
    x = F.conv1d(x, sd[pre + "pred_conv.3.weight"], sd[pre + "pred_conv.3.bias"], padding=1)
    pred = _bn(sd, pre + "pred_conv.5", F.relu(x))
    tail = prior[:, p - chunk:, :].reshape(b, -1)
    # spatial memory
    mem = chain("spatial_memory.spatial_chunk_encoder", tail)                       # (B, P)
    head = pred[:, :chunk, :]
    s = torch.sigmoid(torch.einsum("bp,bcp->bc", mem, head)).unsqueeze(2)
    head = s * head + (1 - s) * mem.unsqueeze(1)
    # temporal memory (note the sum over the batch in mem^T e)
    mem_t = chain("temporal_memory.temporal_chunk_encoder", tail)                   # (B, P)
    e = chain("temporal_memory.temporal_memory_encoder", head.reshape(b, -1))       # (B, chunk)
    soft = torch.softmax(mem_t @ (mem_t.t() @ e), dim=1)
    head = head + head * soft.unsqueeze(2)
    out = torch.cat((prior, head, pred[:, chunk:, :]), 1)
    return chain("post_header", out)


def mha(sd, pre, q_in, kv_in, n_head, d_k, d_v):
    """Full_model/SubLayers.py:30-59 + Full_model/Modules.py:13-23 (mask None, eval)."""
    b, lq, _ = q_in.shape
    lk = kv_in.shape[1]
    q = _linear(sd, pre + ".w_qs", q_in).view(b, lq, n_head, d_k).transpose(1, 2)
    k = _linear(sd, pre + ".w_ks", kv_in).view(b, lk, n_head, d_k).transpose(1, 2)
    v = _linear(sd, pre + ".w_vs", kv_in).view(b, lk, n_head, d_v).transpose(1, 2)
    attn = torch.softmax(torch.matmul(q / (d_k ** 0.5), k.transpose(2, 3)), dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(b, lq, -1)
    o = _linear(sd, pre + ".fc", o) + q_in
    return F.layer_norm(o, (o.shape[-1],), sd[pre + ".layer_norm.weight"],
                        sd[pre + ".layer_norm.bias"], LN_EPS), attn


def ffn(sd, pre, x):
    """Full_model/SubLayers.py:74-84."""
    y = _linear(sd, pre + ".w_2", F.relu(_linear(sd, pre + ".w_1", x))) + x
    return F.layer_norm(y, (y.shape[-1],), sd[pre + ".layer_norm.weight"],
                        sd[pre + ".layer_norm.bias"], LN_EPS)


def generator_forward(sd, cfg, spec, prior, sampled_emotion=None, taps=None):
    """Full_model/Models.py:389-427 minus the text encoder (its output never feeds the poses);
    emotion injection per Full_model/Models_memory.py:551-555.

    Returns (poses, emotion_feature, semantic_feature, emotion_logits).
    """
    x = audio_encoder(sd, spec.unsqueeze(1), taps)
    memory = "prior_seq_encoder.pred_conv.0.weight" in sd           # Models_memory.Transformer checkpoints
    p = prior_memory_encoder(sd, prior) if memory else prior_encoder(sd, prior)
    if taps is not None:
        taps["spectrum_feature"] = x
        taps["prior_feature"] = p
    emo = _linear(sd, "emotion_proj.2", _linear(sd, "emotion_proj.0", x))
    sem = _linear(sd, "semantic_proj.2", _linear(sd, "semantic_proj.0", x))
    b = emo.shape[0]
    h = emo.reshape(b, -1)
    for i in (0, 2, 4):
        h = F.relu(_linear(sd, f"emotion_classifer_header.{i}", h))
    logits = _linear(sd, "emotion_classifer_header.6", h)
    fusion = (emo if sampled_emotion is None else sampled_emotion) + sem
    fusion = _linear(sd, "fusion_proj.2", F.relu(_linear(sd, "fusion_proj.0", fusion)))
    enc = fusion + sd["encoder.position_enc.pos_table"][:, :fusion.shape[1]].to(fusion.dtype)
    for i in range(cfg.n_layers):
        pre = f"encoder.layer_stack.{i}"
        enc, _ = mha(sd, pre + ".slf_attn", enc, enc, cfg.n_head, cfg.d_k, cfg.d_v)
        enc = ffn(sd, pre + ".pos_ffn", enc)
    if taps is not None:
        taps["enc_output"] = enc
    dec = p
    for i in range(cfg.n_layers):
        pre = f"decoder.layer_stack.{i}"
        dec, _ = mha(sd, pre + ".enc_attn", dec, enc, cfg.n_head, cfg.d_k, cfg.d_v)
        dec = ffn(sd, pre + ".pos_ffn", dec)
    if taps is not None:
        taps["dec_output"] = dec
    y = dec
    for i in (0, 2, 4, 6):
        y = _linear(sd, f"post_projector.{i}", y)
    return y, emo, sem, logits


def cast_state_dict(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
