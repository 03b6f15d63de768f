"""ORACLE (test infrastructure only — never imported by the product path): CPU restatements of the small
networks of SURVEY.md §8 rows C4, E1, D1, written functionally over a plain ``state_dict`` in the
reference's layout.  Pinned against the real reference modules by ``oracle/make_golden_aux.py``
(fixtures under ``tests/golden/aux_*.npz``).  Noise is always an explicit argument."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _lin(sd, pre, x):
    return F.linear(x, sd[pre + ".weight"], sd[pre + ".bias"])


def _chain(sd, pre, idx, x):
    # nn.Sequential(Linear, Dropout, Linear, ...) in eval mode: Dropout is the identity
    for i in idx:
        x = _lin(sd, f"{pre}.{i}", x)
    return x


def _bn(sd, pre, x):
    # eval-mode BatchNorm1d over (N, C) or (N, C, L); eps is the PyTorch default 1e-5
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"],
                        sd[pre + ".bias"], False, 0.0, 1e-5)


def cvae_forward(sd, x, y, eps):
    """Full_model/BEAT_CVAE.py:98-114 (reparameterize :84-94)."""
    latent = _chain(sd, "Encoder", (0, 2, 4, 6, 8), x)
    mu, log_var = _lin(sd, "fc_mu", latent), _lin(sd, "fc_var", latent)
    z = eps * torch.exp(0.5 * log_var) + mu
    return cvae_decode(sd, y, z), mu, log_var


def cvae_decode(sd, y, z):
    """Full_model/BEAT_CVAE.py:117-136 (``sample`` with the randn draw passed in as z)."""
    post_y = _chain(sd, "Posterior_Y_embedding", (0, 2), y)
    h = _chain(sd, "fusion_z_posterior", (0, 2), torch.cat([z, post_y], dim=1))
    return _chain(sd, "Decoder", (0, 2, 4, 6, 8), h)


def cvae3_sample(sd, y, z):
    """CAVE/BEAT_CVAE.py:427-447: (n,8) one-hot, (n,32) noise -> (n,60,512)."""
    n = y.shape[0]
    post_y = _chain(sd, "Posterior_Y_embedding", (0, 2), y)
    h = _chain(sd, "fusion_z_posterior", (0, 2), torch.cat([z, post_y], dim=1)).reshape(n, 4, 128)
    d = "Decoder."
    h = F.conv_transpose1d(h, sd[d + "0.weight"], sd[d + "0.bias"], stride=2, padding=1, output_padding=1)
    h = _bn(sd, d + "2", F.leaky_relu(h, 0.2))
    h = F.conv_transpose1d(h, sd[d + "3.weight"], sd[d + "3.bias"], stride=2, padding=1, output_padding=1)
    h = _bn(sd, d + "5", F.leaky_relu(h, 0.2))
    h = _bn(sd, d + "8", F.leaky_relu(F.conv1d(h, sd[d + "6.weight"], sd[d + "6.bias"], padding=1), 0.2))
    h = _bn(sd, d + "11", F.leaky_relu(F.conv1d(h, sd[d + "9.weight"], sd[d + "9.bias"], padding=1), 0.2))
    return F.conv1d(h, sd[d + "12.weight"], sd[d + "12.bias"], padding=1)


def pose_encoder(sd, poses, pre="", fc_mu=False):
    """PoseEncoderConv.forward: model/motion_ae.py:78-83 (``pre='encoder.'``) and
    model/embedding_net.py:67-83 (``fc_mu=True`` returns mu)."""
    h = poses.transpose(1, 2)
    for i, stride in ((0, 1), (1, 1), (2, 2)):
        h = F.conv1d(h, sd[f"{pre}net.{i}.0.weight"], sd[f"{pre}net.{i}.0.bias"], stride=stride)
        h = F.leaky_relu(_bn(sd, f"{pre}net.{i}.1", h), 0.2)
    h = F.conv1d(h, sd[pre + "net.3.weight"], sd[pre + "net.3.bias"]).flatten(1)
    # out_net's nn.LeakyReLU(True) has negative_slope == True == 1.0: the identity
    h = _bn(sd, pre + "out_net.1", _lin(sd, pre + "out_net.0", h))
    h = _bn(sd, pre + "out_net.4", _lin(sd, pre + "out_net.3", h))
    h = _lin(sd, pre + "out_net.6", h)
    return _lin(sd, pre + "fc_mu", h) if fc_mu else h


def fgd_latent(sd, rows):
    """model/FGD.py:66-82: latent_vector = Encoder(x) (three Linears, Dropouts between)."""
    return _chain(sd, "Encoder", (0, 2, 4), rows)


def emotion_net(sd, mfcc, taps=None):
    """model/audio_emotion_classifer.py:38-49 over the trunk of model/emotion_ResNetSE34V2.py:57-72 (the same
    SEBasicBlock as the generator's, one more stage)."""
    from oracle import generator as og
    feat = og.trunk(sd, "emotion_encoder", mfcc.unsqueeze(1), layers=(3, 4, 6, 3), taps=taps)
    x = feat.reshape(feat.shape[0], -1)
    for i in (0, 2, 4, 6, 8):
        x = F.relu(_lin(sd, f"emotion_eocder_fc.{i}", x))
    return _lin(sd, "last_fc", x)


def skeleton_classifier(sd, poses, n_head=8, d_k=64):
    """skeleton_classifer/Models.py:264-283: Prior_Encoder (:109-113) -> + pos_table[:, :T] (:49-50) ->
    EncoderLayers (Layers.py / SubLayers.py, byte-identical copies of the generator's) -> flatten -> post_projector."""
    from oracle import generator as og
    x = _lin(sd, "prior_seq_encoder.fc2", _lin(sd, "prior_seq_encoder.fc1", poses))
    x = x + sd["encoder.position_enc.pos_table"][:, :x.shape[1]].to(x.dtype)
    n_layers = 0
    while f"encoder.layer_stack.{n_layers}.slf_attn.w_qs.weight" in sd:
        n_layers += 1
    for l in range(n_layers):
        pre = f"encoder.layer_stack.{l}"
        x, _ = og.mha(sd, pre + ".slf_attn", x, x, n_head, d_k, d_k)
        x = og.ffn(sd, pre + ".pos_ffn", x)
    mid = x
    h = x.reshape(x.shape[0], -1)
    for i in range(5):
        h = _lin(sd, f"post_projector.{2 * i}", h)
        if i < 4:
            h = torch.relu(h)
    return h, mid


def cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
