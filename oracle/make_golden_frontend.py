"""ORACLE tooling: pin the parts of the front-end oracle that CAN be pinned against the real reference.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden_frontend

  F1   model.utils.PreEmphasis (model/utils.py:22-38) is torch-only and imports as it stands: the real module's
       float64 output pins oracle.logmel.preemphasis.
  F4b  the normalisation the commented-out recipe names (model/ResNetSE34V2.py:96-98; never constructed in the
       reference, `instancenorm` is torch.nn.InstanceNorm1d(n_mels) in the trainer that file was taken from) is pinned
       by running the REAL torch.nn.InstanceNorm1d(128) on log(M + 1e-6) of the real PreEmphasis output.
  F5   utils.data_utils.make_audio_fixed_length (utils/data_utils.py:69-75) is numpy-only; the file's top-level
       `import librosa` is stubbed (the function never touches it).
  F2-F4a stay unpinned: they are librosa calls (requirements.txt:5,16, no version) and librosa is absent here.

Writes tests/golden/frontend_pins.npz; tests/test_oracle_logmel.py checks the oracle against it on CPU and
tests/test_gpu_parity.py checks the CUDA path against it.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import logmel as ol  # noqa: E402
from oracle import synth  # noqa: E402

REF = "/root/reference"
RAGGED = (36267, 30000, 5, 50000, 12089, 1)      # clip lengths of the F5 golden, target 36267


def main():
    sys.path.insert(0, REF)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    from model.utils import PreEmphasis
    from utils.data_utils import calc_spectrogram_length_from_motion_length, make_audio_fixed_length

    audio = synth.synth_audio(3, 36267, seed=11)
    x64 = torch.from_numpy(audio).double()
    pre = PreEmphasis().double()
    with torch.no_grad():
        y_ref = pre(x64)                                   # the real module, float64
        y_ref32 = PreEmphasis()(torch.from_numpy(audio))   # and as shipped (float32)
    y_mine = ol.preemphasis(audio)
    e1 = np.abs(y_mine - y_ref.numpy()).max()
    print(f"F1  oracle preemphasis vs real PreEmphasis (float64): max-abs {e1:.2e}")
    assert e1 <= 1e-15
    assert np.abs(y_mine - y_ref32.double().numpy()).max() <= 1e-7

    # F4b on the real pre-emphasis output: oracle STFT/mel (unpinned), then the real InstanceNorm1d
    power = ol.stft_power(y_ref.numpy(), 70)
    mel = np.einsum("mk,bkt->bmt", ol.mel_filterbank(), power)
    inorm = torch.nn.InstanceNorm1d(128).double().eval()
    assert inorm.eps == 1e-5 and not inorm.affine and not inorm.track_running_stats
    with torch.no_grad():
        f4b_ref = inorm(torch.log(torch.from_numpy(mel) + 1e-6)).numpy()
    f4b_mine = ol.logmel(audio, 70, "log_in", preemph=True)
    e2 = np.abs(f4b_mine - f4b_ref).max()
    print(f"F4b oracle log+InstanceNorm vs real nn.InstanceNorm1d(128) (float64): max-abs {e2:.2e}")
    assert e2 <= 1e-11

    # F5 on ragged clips (shorter, longer, much shorter than the pad, one sample)
    rng = np.random.default_rng(5)
    flat = rng.standard_normal(sum(RAGGED)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(RAGGED)])
    fixed_ref = np.stack([make_audio_fixed_length(flat[off[i]:off[i + 1]], 36267) for i in range(len(RAGGED))])
    fixed_mine = np.stack([ol.make_audio_fixed_length(flat[off[i]:off[i + 1]], 36267) for i in range(len(RAGGED))])
    assert np.array_equal(fixed_ref, fixed_mine)
    print("F5  oracle make_audio_fixed_length == reference on", RAGGED)
    from emotiongestures_b200.config import spectrogram_length
    for n_frames in (34, 60, 30, 64):
        assert spectrogram_length(n_frames, 15) == calc_spectrogram_length_from_motion_length(n_frames, 15)

    path = os.path.join(ROOT, "tests", "golden", "frontend_pins.npz")
    np.savez_compressed(path, audio_seed=np.int64(11), preemph_head=y_ref.numpy()[:, :2048],
                        preemph_sum=y_ref.numpy().sum(axis=1), preemph_abs_sum=np.abs(y_ref.numpy()).sum(axis=1),
                        log_in=f4b_ref, ragged_seed=np.int64(5), ragged_lens=np.array(RAGGED, np.int64),
                        fixed_checksum=fixed_ref.astype(np.float64).sum(axis=1), fixed_tail=fixed_ref[:, -64:])
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
