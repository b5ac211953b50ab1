"""``AudioEncoder``: the reference module's interface (model.py:120-169) over libasrb200.

Constructor signature, ``forward`` contract and ``state_dict`` keys are the reference's, so
``ours.load_state_dict(ref.state_dict())`` works and ``Model`` can hold this module in place
of its own (model.py:646, 665, 685).  The torch sub-modules built here are parameter
CONTAINERS only (they give identical key names and default init); ``forward`` never calls
them -- every FLOP runs in the CUDA library, and a missing library or CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

from . import _lib

THETA = 30000.0      # model.py:26


class _ChannelLayerNorm(nn.Module):          # parameter names of essentials.py:102-108
    def __init__(self, dims):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dims))
        self.beta = nn.Parameter(torch.zeros(dims))


class _ConvLite(nn.Module):                  # parameter names of model.py:93-107
    def __init__(self, dims, kernel_size=15):
        super().__init__()
        self.point1 = nn.Conv1d(dims, dims * 2, kernel_size=1)
        self.depth = nn.Conv1d(dims, dims, kernel_size=kernel_size, padding=(kernel_size - 1) // 2, groups=dims)
        self.bn = nn.BatchNorm1d(dims)
        self.point2 = nn.Conv1d(dims, dims, kernel_size=1)


class AudioEncoder(nn.Module):
    """Drop-in for ``model.AudioEncoder(mels, dims, head, layer, act, n_type, norm=False, enc=False)``.

    Extra keyword ``compute``: ``"bf16"`` (tcgen05 tensor-core path, returns bf16) or ``"fp32"``
    (FFMA path within 1e-4 of the reference, returns fp32).  Inference (eval) semantics only:
    Dropout off, BatchNorm running statistics.
    """

    def __init__(self, mels, dims, head, layer, act="gelu", n_type=None, norm=False, enc=False,
                 compute: str = "bf16", out_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if norm:
            raise NotImplementedError("norm=True is broken in the reference itself (shape error); unsupported")
        if act != "gelu":
            raise NotImplementedError("only act='gelu' (the reference configuration, model.py:746)")
        if compute not in ("bf16", "fp32"):
            raise ValueError("compute must be 'bf16' or 'fp32'")
        self.mels, self.dims, self.head, self.layer, self.enc = mels, dims, head, layer, bool(enc)
        self.compute = compute
        self.out_dtype = out_dtype or (torch.bfloat16 if compute == "bf16" else torch.float32)
        self.conv1 = nn.Sequential(nn.Conv1d(mels, dims, kernel_size=3, stride=1, padding=1), nn.Identity())
        self.conv2 = nn.Sequential(nn.Conv1d(1, dims, kernel_size=3, stride=1, padding=1), nn.Identity())
        self.EncoderLayer = (nn.TransformerEncoderLayer(d_model=dims, nhead=head, batch_first=True)
                             if enc else nn.Identity())
        self.encoder = nn.ModuleList()
        for _ in range(layer):
            self.encoder.append(nn.Sequential(
                nn.Identity(), weight_norm(nn.Conv1d(dims, dims, kernel_size=3, padding=1)),
                _ChannelLayerNorm(dims), _ConvLite(dims, 15), nn.Identity(),
                nn.Conv1d(dims, dims, kernel_size=3, stride=1, padding=1, groups=dims), nn.Identity(), nn.Identity()))
        self._handles = {}            # device index -> (ctypes handle, weight version)
        self._ws = {}
        self._lib = None

    # ------------------------------------------------------------------ weights -> handle
    def _weights_version(self):
        """Sum of the in-place version counters of every parameter and buffer (cached list: no state_dict() per forward)."""
        ts = self.__dict__.get("_vt")
        if ts is None:
            ts = [t for t in list(self.parameters()) + list(self.buffers())]
            self.__dict__["_vt"] = ts
        return sum(int(t._version) for t in ts)

    def prepare(self, device=None):
        """Fold weight-norm / BatchNorm, pack and upload the weights (done lazily by forward;
        call again after changing parameters in place)."""
        self._lib = _lib.load()
        dev = torch.device(device if device is not None else "cuda")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self._release(idx)
        sd = {k: v for k, v in self.state_dict().items() if torch.is_floating_point(v)}
        half = self.dims // 2
        # sinusoid scales with the reference's own ops (essentials.py:355) -> bit-equal table
        sd["__pos_scales"] = torch.exp(-torch.log(torch.tensor(float(THETA))) / (half - 1)
                                       * torch.arange(half, dtype=torch.float32))
        n, names, ptrs, nums, keep = _lib.state_dict_arrays(sd)
        cfg = _lib.EncoderConfig(self.mels, self.dims, self.head, self.layer, int(self.enc), 2048,
                                 _lib.BF16 if self.compute == "bf16" else _lib.F32, 0)
        h = C.c_void_p()
        with torch.cuda.device(idx):
            _lib.check(self._lib.asrb_encoder_create(C.byref(cfg), n, names, ptrs, nums, C.byref(h)),
                       "asrb_encoder_create")
        del keep
        self._handles[idx] = (h, self._weights_version())
        return self

    def _release(self, idx=None):
        for i in ([idx] if idx is not None else list(self._handles)):
            if i in self._handles:
                self._lib.asrb_encoder_destroy(self._handles.pop(i)[0])

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _apply(self, fn, *a, **kw):                     # .to() / .cuda() may replace parameter tensors: drop the cached list
        self.__dict__.pop("_vt", None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.__dict__.pop("_vt", None)
        if self._handles:
            self._release()
        return r

    def _handle(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        cur = self._handles.get(idx)
        if cur is None or cur[1] != self._weights_version():
            self.prepare(torch.device("cuda", idx))
        return self._handles[idx][0]

    def _workspace(self, device, nbytes):
        ws = self._ws.get(device)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._ws[device] = ws
        return ws

    # ------------------------------------------------------------------ forward
    def _process_feature(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:149-163: ``[B, mels, T]`` (or ``[mels, T]``; one channel selects conv2) -> ``[B, T, dims]``."""
        if self.training:
            raise _lib.AsrbError("AudioEncoder implements inference semantics: call .eval() first")
        if not x.is_cuda:
            raise _lib.AsrbError("AudioEncoder needs a CUDA tensor: there is no CPU path")
        if x.dim() == 2:
            x = x.unsqueeze(0)
        x = x.float().contiguous()
        B, Cin, T = x.shape
        h = self._handle(x.device)
        out = torch.empty(B, T, self.dims, device=x.device, dtype=self.out_dtype)
        need = self._lib.asrb_encoder_workspace_bytes(h, B, T)
        ws = self._workspace(x.device, need)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.asrb_encoder_forward(
                h, x.data_ptr(), B, Cin, T, out.data_ptr(),
                _lib.BF16 if self.out_dtype == torch.bfloat16 else _lib.F32,
                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "asrb_encoder_forward")
        return out

    def _process_streams(self, feats):
        """Several same-shaped feature streams (the reference's TensorDict {a, b, c}) through one pass of the layer
        stack: ``{name: [B, C_s, T]}`` -> ``{name: [B, T, dims]}`` (views of one tensor)."""
        names = list(feats)
        xs = [feats[k].unsqueeze(0) if feats[k].dim() == 2 else feats[k] for k in names]
        xs = [x.float().contiguous() for x in xs]
        B, _, T = xs[0].shape
        same = all(x.is_cuda and x.shape[0] == B and x.shape[2] == T and x.device == xs[0].device for x in xs)
        if self.training or not same or len(xs) < 2 or len(xs) > 8:
            return {k: self._process_feature(feats[k]) for k in names}
        dev = xs[0].device
        h = self._handle(dev)
        n = len(xs)
        out = torch.empty(n * B, T, self.dims, device=dev, dtype=self.out_dtype)
        ws = self._workspace(dev, self._lib.asrb_encoder_workspace_bytes(h, n * B, T))
        ptrs = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
        chans = (C.c_int32 * n)(*[x.shape[1] for x in xs])
        with torch.cuda.device(dev):
            _lib.check(self._lib.asrb_encoder_forward_streams(
                h, n, ptrs, chans, B, T, out.data_ptr(), _lib.BF16 if self.out_dtype == torch.bfloat16 else _lib.F32,
                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "asrb_encoder_forward_streams")
        return {k: out[i * B:(i + 1) * B] for i, k in enumerate(names)}

    def forward(self, x):
        if hasattr(x, "apply") and not torch.is_tensor(x):          # TensorDict (model.py:166-167)
            return x.apply(self._process_feature)
        if isinstance(x, dict):                                     # same streams as a plain dict: batched
            return self._process_streams({k: v for k, v in x.items() if v is not None})
        return self._process_feature(x)

    def forward_ragged(self, x: torch.Tensor, frames: torch.Tensor) -> torch.Tensor:
        """``[B, mels, T]`` features of a padded batch + ``frames [B]`` (frames that hold audio) -> ``[B, T, dims]`` whose rows
        ``t < frames[b]`` equal ``forward(x)`` bit for bit and whose padding rows are 0; on the tensor-core path without
        the TransformerEncoderLayer the frame tiles of the padding are never computed."""
        if self.training:
            raise _lib.AsrbError("AudioEncoder implements inference semantics: call .eval() first")
        if not x.is_cuda:
            raise _lib.AsrbError("AudioEncoder needs a CUDA tensor: there is no CPU path")
        x = x.float().contiguous()
        B, Cin, T = x.shape
        frames = torch.as_tensor(frames)
        if frames.numel() != B:
            raise ValueError(f"frames has {frames.numel()} entries for a batch of {B}")
        frames = frames.to(x.device, torch.int32).contiguous()
        h = self._handle(x.device)
        out = torch.empty(B, T, self.dims, device=x.device, dtype=self.out_dtype)
        ws = self._workspace(x.device, self._lib.asrb_encoder_workspace_bytes(h, B, T))
        with torch.cuda.device(x.device):
            _lib.check(self._lib.asrb_encoder_forward_ragged(
                h, x.data_ptr(), B, Cin, T, frames.data_ptr(), out.data_ptr(),
                _lib.BF16 if self.out_dtype == torch.bfloat16 else _lib.F32,
                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "asrb_encoder_forward_ragged")
        return out

    def forward_pcm(self, wave: torch.Tensor, frontend, lengths: Optional[torch.Tensor] = None,
                    return_logmel: bool = False, out: Optional[torch.Tensor] = None, skip_padding: bool = False):
        """Fused hot path: PCM ``[B, N]`` -> hidden states ``[B, T, dims]`` in one library call
        (``asrb_pcm_to_hidden``).  ``frontend`` is a ``LogMel`` plan with ``n_mels == mels``.
        ``skip_padding=True`` (needs ``lengths``): ``asrb_pcm_to_hidden_ragged`` -- the rows of each utterance's valid frames
        are unchanged, the rows of its padding are 0, and padded frame tiles are skipped from the FFT to the last block."""
        if self.training:
            raise _lib.AsrbError("AudioEncoder implements inference semantics: call .eval() first")
        if not wave.is_cuda:
            raise _lib.AsrbError("forward_pcm needs a CUDA tensor: there is no CPU path")
        if wave.dim() == 1:
            wave = wave.unsqueeze(0)
        wave = wave.float()
        if wave.stride(-1) != 1:
            wave = wave.contiguous()
        B, N = wave.shape
        T = frontend.num_frames(N)
        h = self._handle(wave.device)
        if out is None:
            out = torch.empty(B, T, self.dims, device=wave.device, dtype=self.out_dtype)
        elif tuple(out.shape) != (B, T, self.dims) or out.dtype != self.out_dtype or not out.is_contiguous():
            raise ValueError("out must be a contiguous [B, T, dims] tensor of the module's out_dtype")
        mel = torch.empty(B, self.mels, T, device=wave.device, dtype=torch.float32) if return_logmel else None
        from .frontend import check_lengths
        lengths = check_lengths(lengths, B, N, wave.device)
        need = self._lib.asrb_pcm_to_hidden_workspace_bytes(frontend.handle, h, B, N)
        ws = self._workspace(wave.device, need)
        if skip_padding and lengths is None:
            raise ValueError("skip_padding=True needs lengths")
        fn = self._lib.asrb_pcm_to_hidden_ragged if skip_padding else self._lib.asrb_pcm_to_hidden
        with torch.cuda.device(wave.device):
            _lib.check(fn(
                frontend.handle, h, wave.data_ptr(), B, N, wave.stride(0) if B > 1 else max(N, 1),
                lengths.data_ptr() if lengths is not None else None,
                mel.data_ptr() if mel is not None else None, out.data_ptr(),
                _lib.BF16 if self.out_dtype == torch.bfloat16 else _lib.F32,
                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "asrb_pcm_to_hidden")
        return (out, mel) if return_logmel else out
