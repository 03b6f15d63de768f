"""Object wrapper over the libegx C ABI: weight hand-off, workspace, stream plumbing.

PyTorch is used here for what the reference's host code uses it for — device memory
(the caching allocator owns every buffer the library touches) and the current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import LOGMEL_REFERENCE, GeneratorConfig

_PREC = {"fp32": _lib.EGX_PREC_FP32, "tc": _lib.EGX_PREC_TC}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One libegx handle bound to one CUDA device (no global state, so DataParallel-style
    per-replica threads each own an Engine)."""

    def __init__(self, cfg: GeneratorConfig, device, precision: str = "tc"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(
                "emotiongestures_b200 runs on sm_100a CUDA devices only (no CPU fallback); "
                f"got device {self.device}")
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        cfg.validate()
        self.cfg = cfg
        self.precision = precision
        self.lib = _lib.load_library()
        c = _lib.EgxCfg(cfg.frames, cfg.prior_frames, cfg.pose_dim, cfg.d_model, cfg.d_inner,
                        cfg.n_layers, cfg.n_head, cfg.d_k, cfg.d_v, cfg.n_mels, cfg.spec_w,
                        cfg.n_position, _PREC[precision])
        h = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        rc = self.lib.egx_create(C.byref(c), index, C.byref(h))
        if rc != 0:
            reasons = {2: "bad CUDA device", 3: "device is not compute capability 10.x (sm_100a)",
                       4: "unsupported geometry", 5: "device allocation failed"}
            raise RuntimeError(f"egx_create failed: {reasons.get(rc, rc)}")
        self._h = h
        self._ws = None
        self._ws_clips = 0
        self.batch_coupled = False

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.egx_destroy(h)

    # -- helpers ---------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.egx_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, name, shape=None):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        # borrowed caller tensors may be CPU, non-contiguous views or other float types
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    @property
    def launch_count(self) -> int:
        return int(self.lib.egx_launch_count(self._h))

    N_STAGES = 12
    STAGE_NAMES = ("other", "S1_frontend", "S2_stem", "S3_conv_layer1", "S4_se", "S5_proj_gemm",
                   "S6_enc_dec", "S7", "S8_fgd", "S3_conv_layer2", "S3_conv_layer3", "S3_conv_layer4")

    def profile_enable(self, max_launches: int):
        self._check(self.lib.egx_profile_enable(self._h, int(max_launches)), "egx_profile_enable")

    def profile_read(self):
        """{stage: (milliseconds, launches)} of everything recorded since profile_enable."""
        ms = (C.c_double * self.N_STAGES)()
        cnt = (C.c_int64 * self.N_STAGES)()
        self._check(self.lib.egx_profile_read(self._h, ms, cnt, self.N_STAGES), "egx_profile_read")
        return {self.STAGE_NAMES[i]: (ms[i], int(cnt[i])) for i in range(self.N_STAGES) if cnt[i]}

    # -- weights ---------------------------------------------------------------
    def load_state_dict(self, sd):
        """Hand every float tensor of a reference-layout state_dict to the library."""
        with torch.cuda.device(self.device):
            # Models_memory.Transformer (Prior_MemoryEncoder): its temporal memory sums over the clips of one call
            self.batch_coupled = "prior_seq_encoder.pred_conv.0.weight" in sd
            staged = {k: v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                      for k, v in sd.items() if v.is_floating_point()}
            # egx_set_weight copies synchronously on the legacy default stream, which does not order against a
            # non-blocking torch stream that may still be producing the fp32 copies above
            torch.cuda.current_stream(self.device).synchronize()
            for k, t in staged.items():
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                self._check(self.lib.egx_set_weight(self._h, k.encode(), _ptr(t), shape, t.dim(),
                                                    _lib.EGX_DTYPE_F32), f"egx_set_weight({k})")
            self._check(self.lib.egx_finalize_weights(self._h), "egx_finalize_weights")

    # -- workspace ---------------------------------------------------------------
    def workspace(self, n_clips: int):
        if self._ws is None or n_clips > self._ws_clips:
            nbytes = int(self.lib.egx_workspace_bytes(self._h, n_clips))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_clips = n_clips
        return self._ws

    # -- hot path ----------------------------------------------------------------
    def fixed_length_audio(self, clips, n_samples=None):
        """F5 (utils/data_utils.py:69-75): a list of ragged 1-D clips -> (B, n_samples) f32 on the device, each clip
        cropped or symmetric-padded at its end; the `audio` argument of `logmel`."""
        n_out = self.cfg.n_audio if n_samples is None else int(n_samples)
        lens = [int(c.numel()) for c in clips]
        if any(n == 0 for n in lens):
            raise RuntimeError("fixed_length_audio: empty clip (np.pad cannot mirror an empty signal either)")
        out = torch.empty((len(clips), n_out), dtype=torch.float32, device=self.device)
        if not clips:
            return out
        flat = torch.cat([self._f32(c.reshape(-1), "clip") for c in clips])
        offs = torch.tensor([0] + lens, dtype=torch.int64).cumsum(0).to(self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_audio_fixed_length(self._h, _ptr(flat), _ptr(offs), len(clips), n_out, _ptr(out),
                                                        self._stream()), "egx_audio_fixed_length")
        return out

    def pcm16_to_float(self, pcm, out=None):
        """int16 PCM device tensor -> float32 in [-1, 1) (x / 32768, exact), on the device (`egx_audio_pcm16_to_f32`)."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda or not pcm.is_contiguous():
            raise RuntimeError("pcm must be a contiguous int16 CUDA tensor")
        if out is None:
            out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_audio_pcm16_to_f32(self._h, _ptr(pcm), pcm.numel(), _ptr(out), self._stream()),
                        "egx_audio_pcm16_to_f32")
        return out

    def logmel(self, audio, mode: int = LOGMEL_REFERENCE, preemph: bool = False, n_cols=None, _global_tile=False):
        """(B,N) 16 kHz audio -> (B,128,n_cols) log-mel (F1–F4).  Defaults = the reference's live features
        (config.LOGMEL_REFERENCE); the north star's PreEmphasis + log + InstanceNorm recipe is
        `mode=LOGMEL_LOG_IN, preemph=True`."""
        a = self._f32(audio, "audio")
        if a.dim() != 2:
            raise RuntimeError("audio must be (B, N)")
        b, n = a.shape
        n_cols = self.cfg.spec_w if n_cols is None else int(n_cols)
        out = torch.empty((b, self.cfg.n_mels, n_cols), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            fn = self.lib.egx_debug_logmel_global_tile if _global_tile else self.lib.egx_logmel   # test hook
            self._check(fn(self._h, _ptr(a), b, n, n_cols, int(mode), int(bool(preemph)), _ptr(out), self._stream()),
                        "egx_logmel")
        return out

    def generator_forward(self, spec, prior, sampled_emotion=None, out=None):
        """`out`: optional preallocated (poses, emo, sem, logits) device tensors (chunked pipelines)."""
        cfg = self.cfg
        if spec.dim() != 3:
            raise RuntimeError("input_spectrum must be (B, n_mels, W)")
        b = spec.shape[0]
        spec = self._f32(spec, "input_spectrum", (b, cfg.n_mels, cfg.spec_w))
        prior = self._f32(prior, "prior_seq", (b, cfg.prior_frames, cfg.pose_dim))
        if sampled_emotion is not None:
            sampled_emotion = self._f32(sampled_emotion, "sampled_emotion_feature",
                                        (b, cfg.frames, cfg.d_model))
        dev = self.device
        if out is not None:
            poses, emo, sem, logits = out
            for t, shp in ((poses, (b, cfg.frames, cfg.pose_dim)), (emo, (b, cfg.frames, cfg.d_model)),
                           (sem, (b, cfg.frames, cfg.d_model)), (logits, (b, 8))):
                if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                    raise RuntimeError("`out` tensors must be contiguous float32 on the engine device")
        else:
            poses = torch.empty((b, cfg.frames, cfg.pose_dim), dtype=torch.float32, device=dev)
            emo = torch.empty((b, cfg.frames, cfg.d_model), dtype=torch.float32, device=dev)
            sem = torch.empty_like(emo)
            logits = torch.empty((b, 8), dtype=torch.float32, device=dev)
        if b == 0:
            return poses, emo, sem, logits
        ws = self.workspace(b)
        with torch.cuda.device(dev):
            self._check(self.lib.egx_generator_forward(
                self._h, _ptr(spec), _ptr(prior), _ptr(sampled_emotion), b, _ptr(poses), _ptr(emo),
                _ptr(sem), _ptr(logits), _ptr(ws), ws.numel(), self._stream()),
                "egx_generator_forward")
        self._last_b = b
        return poses, emo, sem, logits

    def infer_host(self, audio_h, prior_h, poses_h, chunk: int = 512, mode: int = LOGMEL_REFERENCE,
                   preemph: bool = False, poses_dev=None, join: bool = True, graph: bool = False):
        """End-to-end batch from PINNED host buffers: audio_h (B,N), prior_h (B,p,P) -> poses_h (B,F,P).

        The batch is cut into chunks; the host->device copy of chunk i+1, the kernels of chunk i and
        the device->host copy of chunk i-1 run on three streams, so PCIe time hides behind compute.
        `poses_dev` (optional, (B,F,P) device tensor) also keeps the poses on the GPU (pose gather).
        Returns after enqueueing; the caller synchronises (torch.cuda.synchronize / an event).
        `join=True` makes the CURRENT stream wait for the last device->host copy, so synchronising that stream is
        enough; a caller that streams batch after batch passes `join=False` (the next call's kernels then do not queue
        behind this call's last copy) and calls `host_join()` once before it reads the host buffers.
        `graph=True` (single-chunk calls, i.e. chunk >= B): the log-mel + forward of each staging slot is a captured
        CUDA graph replayed with one launch instead of ~110 — under a saturated PCIe link (8 ranks streaming audio)
        every kernel launch otherwise queues behind the copy traffic.
        """
        cfg, dev = self.cfg, self.device
        b = audio_h.shape[0]
        if not (audio_h.is_pinned() and prior_h.is_pinned() and poses_h.is_pinned()):
            raise RuntimeError("infer_host needs pinned host tensors (torch.Tensor.pin_memory())")
        pcm = audio_h.dtype == torch.int16
        if not pcm and audio_h.dtype != torch.float32:
            raise RuntimeError("audio_h must be float32 or int16 PCM")
        if self.batch_coupled and b > 0:
            # Prior_MemoryEncoder multiplies by memory_encoding.t() @ pred_encoding, a sum over the clips of the call
            # (Full_model/Models_memory.py:287-288): cutting the batch would change every clip's poses.  The whole
            # batch goes through as ONE forward — same result as forward() on it — and only the copies are chunked.
            return self._infer_host_whole(audio_h, prior_h, poses_h, mode, preemph, poses_dev)
        st = getattr(self, "_pipe", None)
        if st is None or st["chunk"] < min(chunk, b) or st["pcm"] != pcm:
            c = min(chunk, b)
            st = self._pipe = {
                "chunk": c, "pcm": pcm, "h2d": torch.cuda.Stream(dev), "d2h": torch.cuda.Stream(dev),
                "audio": [torch.empty((c, audio_h.shape[1]), dtype=audio_h.dtype, device=dev) for _ in range(2)],
                "audio_f32": torch.empty((c, audio_h.shape[1]), device=dev) if pcm else None,
                "prior": [torch.empty((c, cfg.prior_frames, cfg.pose_dim), device=dev) for _ in range(2)],
                "out": [(torch.empty((c, cfg.frames, cfg.pose_dim), device=dev),
                         torch.empty((c, cfg.frames, cfg.d_model), device=dev),
                         torch.empty((c, cfg.frames, cfg.d_model), device=dev),
                         torch.empty((c, 8), device=dev)) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)], "done": [torch.cuda.Event() for _ in range(2)],
                "in_free": [torch.cuda.Event() for _ in range(2)], "out_free": [torch.cuda.Event() for _ in range(2)],
            }
        c = st["chunk"]
        main = torch.cuda.current_stream(dev)
        # chunk bounds: a short first chunk (its host->device copy is the only one nothing can hide), then full ones
        bounds, lo = [], 0
        while lo < b:
            hi = min(b, lo + (max(1, c // 4) if lo == 0 and b > c else c))
            bounds.append((lo, hi))
            lo = hi
        used = st.setdefault("used", [False, False])
        for i, (lo, hi) in enumerate(bounds):
            slot = (st.get("next_slot", 0) + i) & 1
            n = hi - lo
            path = None
            if graph and len(bounds) == 1:
                key = (slot, n, int(mode), bool(preemph))
                path = st.setdefault("paths", {}).get(key)
                if path is None:                 # first use of this slot at this size: capture (synchronises once)
                    path = st["paths"][key] = GraphedPath(self, n, mode, preemph, False)
            with torch.cuda.stream(st["h2d"]):
                # the inputs are pinned HOST memory, already final when this call is made: the copy only has to wait
                # for the staging slot, so the first copy of a call overlaps the tail of the previous call's kernels
                if used[slot]:
                    st["h2d"].wait_event(st["in_free"][slot])
                # graph mode: float audio and the prior land directly in the graph's static inputs of this slot
                (path.audio if path is not None and not pcm else st["audio"][slot][:n]).copy_(audio_h[lo:hi], non_blocking=True)
                (path.prior if path is not None else st["prior"][slot][:n]).copy_(prior_h[lo:hi], non_blocking=True)
                st["ready"][slot].record(st["h2d"])
            main.wait_event(st["ready"][slot])
            if used[slot]:
                main.wait_event(st["out_free"][slot])
            a_dev = st["audio"][slot][:n]
            if path is not None:
                if pcm:        # widen the staged PCM into the graph's static input, then ONE launch
                    self.pcm16_to_float(a_dev, path.audio)
                path.graph.replay()
                out = path.out
            else:
                if pcm:        # the staging slot is free again as soon as the samples are widened
                    a_dev = self.pcm16_to_float(a_dev, st["audio_f32"][:n])
                spec = self.logmel(a_dev, mode, preemph)
                out = tuple(t[:n] for t in st["out"][slot])
                self.generator_forward(spec, st["prior"][slot][:n], None, out=out)
            if poses_dev is not None:
                poses_dev[lo:hi].copy_(out[0], non_blocking=True)
            st["in_free"][slot].record(main)
            st["done"][slot].record(main)
            with torch.cuda.stream(st["d2h"]):
                st["d2h"].wait_event(st["done"][slot])
                poses_h[lo:hi].copy_(out[0], non_blocking=True)
                st["out_free"][slot].record(st["d2h"])
            used[slot] = True
        st["next_slot"] = (st.get("next_slot", 0) + len(bounds)) & 1
        if join:
            main.wait_stream(st["d2h"])
        return poses_h

    def host_join(self):
        """Make the current stream wait for every device->host copy `infer_host(join=False)` has enqueued."""
        st = getattr(self, "_pipe", None)
        if st is not None:
            torch.cuda.current_stream(self.device).wait_stream(st["d2h"])

    def _infer_host_whole(self, audio_h, prior_h, poses_h, mode, preemph, poses_dev):
        dev = self.device
        audio = audio_h.to(dev, non_blocking=True)
        if audio.dtype == torch.int16:
            audio = self.pcm16_to_float(audio)
        prior = prior_h.to(dev, non_blocking=True)
        poses = self.generator_forward(self.logmel(audio, mode, preemph), prior)[0]
        if poses_dev is not None:
            poses_dev.copy_(poses, non_blocking=True)
        poses_h.copy_(poses, non_blocking=True)
        return poses_h

    def capture(self, n_clips: int, mode: int = LOGMEL_REFERENCE, preemph: bool = False, with_emotion: bool = False):
        """CUDA-graph the whole path (log-mel + generator forward, ~130 launches) for a fixed batch size.

        Small batches are launch-bound: a 1-clip forward is ~130 kernels of a few microseconds each.  The returned
        `GraphedPath` owns static input / output / workspace buffers and replays everything with one
        cudaGraphLaunch; call it with (audio, prior[, sampled_emotion]) and it returns its static
        (poses, emotion_feature, semantic_feature, emotion_logits) tensors (clone them to keep them)."""
        return GraphedPath(self, n_clips, mode, preemph, with_emotion)

    # -- parity probes (tests) ----------------------------------------------------
    def tap(self, name: str):
        b = self._last_b
        cfg = self.cfg
        out = torch.empty(b * cfg.frames * max(cfg.d_model, cfg.fc1_in), dtype=torch.float32,
                          device=self.device)
        n = C.c_size_t()
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_get_tap(self._h, name.encode(), _ptr(self._ws), b, _ptr(out),
                                             out.numel(), C.byref(n), self._stream()), "egx_get_tap")
        return out[: n.value].view(b, cfg.frames, -1)

    def trunk_stage(self, spec, stage: int):
        cfg = self.cfg
        b = spec.shape[0]
        spec = self._f32(spec, "input_spectrum", (b, cfg.n_mels, cfg.spec_w))
        ws = self.workspace(b)
        out = torch.empty(b * cfg.n_mels * cfg.spec_w * 32, dtype=torch.float32, device=self.device)
        n = C.c_size_t()
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_debug_trunk(self._h, _ptr(spec), b, stage, _ptr(out),
                                                 out.numel(), C.byref(n), _ptr(ws), ws.numel(),
                                                 self._stream()), "egx_debug_trunk")
        li = max(stage - 1, 0)
        c = (32, 64, 128)[li]
        h, w = cfg.n_mels, cfg.spec_w
        for _ in range(li):
            h, w = (h + 1) // 2, (w + 1) // 2
        return out[: n.value].view(b, c, h, w)

    def debug_linear_tc(self, a, w, bias=None, addend=None, addend_rows=0, relu=False):
        a, w = self._f32(a, "A"), self._f32(w, "W")
        m, k = a.shape
        n = w.shape[0]
        bias = None if bias is None else self._f32(bias, "bias")
        addend = None if addend is None else self._f32(addend, "addend")
        out = torch.empty((m, n), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_debug_linear_tc(self._h, _ptr(a), _ptr(w), _ptr(bias), _ptr(addend),
                                                     int(addend_rows), m, n, k, int(relu), _ptr(out),
                                                     self._stream()), "egx_debug_linear_tc")
        return out

    def debug_linear_ln_tc(self, a, w, bias, residual, ln_g, ln_b):
        """LayerNorm(a @ w.T + bias + residual) through the LN-epilogue GEMM (w: (256, K)); returns (f32, f16) outputs."""
        a, w = self._f32(a, "A"), self._f32(w, "W")
        m, k = a.shape
        bias = None if bias is None else self._f32(bias, "bias")
        residual = None if residual is None else self._f32(residual, "residual")
        ln_g, ln_b = self._f32(ln_g, "ln_g"), self._f32(ln_b, "ln_b")
        out = torch.empty((m, 256), dtype=torch.float32, device=self.device)
        out16 = torch.empty((m, 256), dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_debug_linear_ln_tc(self._h, _ptr(a), _ptr(w), _ptr(bias), _ptr(residual), _ptr(ln_g),
                                                        _ptr(ln_b), m, k, _ptr(out), _ptr(out16), self._stream()),
                        "egx_debug_linear_ln_tc")
        return out, out16

    def debug_ffn_tc(self, x, w1, b1, w2, b2, ln_g, ln_b):
        """LayerNorm(x + relu(x @ w1.T + b1) @ w2.T + b2) through the fused feed-forward kernel; (f32, f16) outputs."""
        x, w1, b1, w2, b2, ln_g, ln_b = (self._f32(t, "arg") for t in (x, w1, b1, w2, b2, ln_g, ln_b))
        m, d_inner = x.shape[0], w1.shape[0]
        out = torch.empty((m, 256), dtype=torch.float32, device=self.device)
        out16 = torch.empty((m, 256), dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_debug_ffn_tc(self._h, _ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(ln_g),
                                                  _ptr(ln_b), m, d_inner, _ptr(out), _ptr(out16), self._stream()),
                        "egx_debug_ffn_tc")
        return out, out16

    def debug_conv_tc(self, x, w, scale, shift, bias=None, stride=1, relu_first=False, nchw=False,
                      se_sums=False):
        """x (B,Cin,H,W) f32, w (Cout,Cin,ks,ks) f32 -> (B,Cout,Ho,Wo) f32 via the tcgen05 conv kernel."""
        dev = self.device
        b, cin, hh, ww = x.shape
        cout, _, ks, _ = w.shape
        x16 = x.to(dev).permute(0, 2, 3, 1).contiguous().half()
        w16 = w.to(dev).permute(0, 2, 3, 1).contiguous().half()          # [cout][kh][kw][cin]
        pad = ks // 2
        ho, wo = (hh + 2 * pad - ks) // stride + 1, (ww + 2 * pad - ks) // stride + 1
        shape = (b, cout, ho * wo) if nchw else (b, ho, wo, cout)
        out = torch.full(shape, float("nan"), dtype=torch.float16, device=dev)
        bias = None if bias is None else self._f32(bias, "bias")
        scale, shift = self._f32(scale, "scale"), self._f32(shift, "shift")
        part = torch.full((b * 256 * cout,), float("nan"), device=dev) if se_sums else None
        with torch.cuda.device(dev):
            self._check(self.lib.egx_debug_conv_tc(
                self._h, _ptr(x16), b, hh, ww, cin, _ptr(w16), cout, ks, stride, int(relu_first), _ptr(bias),
                _ptr(scale), _ptr(shift), _ptr(out), int(nchw), _ptr(part), self._stream()), "egx_debug_conv_tc")
        out = out.float()
        out = out.view(b, cout, ho, wo) if nchw else out.permute(0, 3, 1, 2).contiguous()
        if se_sums:
            torch.cuda.synchronize()
            n_used = int((~torch.isnan(part)).sum().item())       # slots are [B][tiles_per_clip][cout]
            return out, part[:n_used].view(b, -1, cout).sum(dim=1)
        return out

    def debug_attention_tc(self, q, k, v):
        """q (B,H,Lq,64), k/v (B,H,L,64) f32 -> (B,H,L,64) f32 through the tcgen05 attention kernel,
        using the packed [rows][3*H*64] layout the QKV GEMM produces."""
        b, nh, ll, dk = q.shape
        qkv = torch.cat([t.permute(0, 2, 1, 3).reshape(b * ll, nh * dk) for t in (q, k, v)], dim=1)
        qkv = qkv.to(self.device).half().contiguous()
        out = torch.full((b * ll, nh * dk), float("nan"), dtype=torch.float16, device=self.device)
        hk = nh * dk
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_debug_attention_tc(self._h, _ptr(qkv), 3 * hk, 0, _ptr(qkv), 3 * hk, hk, 2 * hk,
                                                        b, ll, nh, _ptr(out), hk, self._stream()),
                        "egx_debug_attention_tc")
        return out.float().view(b, ll, nh, dk).permute(0, 2, 1, 3).contiguous()

    def fgd_accumulate(self, feats, acc, shift=None):
        """Add the sufficient statistics of feats (n,D) f32 into acc [1+D+D*D] f64."""
        f = self._f32(feats, "feats")
        n, d = f.shape
        assert acc.dtype == torch.float64 and acc.numel() == 1 + d + d * d and acc.is_cuda
        with torch.cuda.device(self.device):
            self._check(self.lib.egx_fgd_accumulate(self._h, _ptr(f), n, d, _ptr(shift), _ptr(acc),
                                                    self._stream()), "egx_fgd_accumulate")
        return acc


class GraphedPath:
    """Fixed-batch CUDA graph over egx_logmel + egx_generator_forward (see Engine.capture)."""

    def __init__(self, eng: Engine, n_clips: int, mode: int, preemph: bool, with_emotion: bool):
        cfg, dev = eng.cfg, eng.device
        self.eng, self.n = eng, int(n_clips)
        n = self.n
        self.audio = torch.zeros((n, cfg.n_audio), device=dev)
        self.prior = torch.zeros((n, cfg.prior_frames, cfg.pose_dim), device=dev)
        self.sampled = torch.zeros((n, cfg.frames, cfg.d_model), device=dev) if with_emotion else None
        self.spec = torch.empty((n, cfg.n_mels, cfg.spec_w), device=dev)
        self.out = (torch.empty((n, cfg.frames, cfg.pose_dim), device=dev), torch.empty((n, cfg.frames, cfg.d_model), device=dev),
                    torch.empty((n, cfg.frames, cfg.d_model), device=dev), torch.empty((n, 8), device=dev))
        # a private workspace: the engine's shared one may be re-allocated when a larger batch comes along
        self.ws = torch.empty(int(eng.lib.egx_workspace_bytes(eng._h, n)), dtype=torch.uint8, device=dev)
        self._mode, self._preemph = int(mode), int(bool(preemph))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._enqueue()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()

    def _enqueue(self):
        eng, cfg, n = self.eng, self.eng.cfg, self.n
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_logmel(eng._h, _ptr(self.audio), n, cfg.n_audio, cfg.spec_w, self._mode, self._preemph,
                                          _ptr(self.spec), eng._stream()), "egx_logmel")
            eng._check(eng.lib.egx_generator_forward(
                eng._h, _ptr(self.spec), _ptr(self.prior), _ptr(self.sampled), n, *(_ptr(t) for t in self.out),
                _ptr(self.ws), self.ws.numel(), eng._stream()), "egx_generator_forward")

    def __call__(self, audio, prior, sampled_emotion=None):
        if (self.sampled is None) != (sampled_emotion is None):
            raise RuntimeError("this graph was captured %s sampled_emotion_feature" % ("without" if self.sampled is None else "with"))
        self.audio.copy_(audio, non_blocking=True)
        self.prior.copy_(prior, non_blocking=True)
        if sampled_emotion is not None:
            self.sampled.copy_(sampled_emotion, non_blocking=True)
        self.graph.replay()
        return self.out
