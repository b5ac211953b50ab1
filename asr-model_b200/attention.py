"""``AudioAttention``: the ``attention`` block's live, deterministic branch applied to encoded
audio (model.py:234-317 with ``n_type="rmsnorm"``, ``xa=None, mask=None, pt=None``) including
``rotary`` (model.py:171-214).  Parameter names match the reference module so its
``state_dict`` loads unchanged.  Batched = the reference's B=1 semantics per utterance
(its own broadcast is only defined for B == 1)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class _Named(nn.Sequential):
    pass


class AudioAttention(nn.Module):
    def __init__(self, dims: int, head: int, layer: int = 1, n_type: str = "rmsnorm"):
        super().__init__()
        if n_type != "rmsnorm":
            raise NotImplementedError("only n_type='rmsnorm' is deterministic in the reference")
        self.dims, self.head = dims, head
        hd = dims // head
        self.q = _Named(nn.RMSNorm(dims), nn.Linear(dims, dims))
        self.kv = _Named(nn.RMSNorm(dims), nn.Linear(dims, dims * 2))
        self.c = _Named(nn.RMSNorm(dims), nn.Linear(dims, dims))          # unused by the live branch
        self.out = _Named(nn.Identity(), nn.Linear(dims, dims))
        self.ln = nn.RMSNorm(hd)
        self.rot = nn.Module()
        self.rot.lin = nn.Linear(dims, hd // 2, bias=True)                # unused parameter (model.py:178)
        self._handle = None
        self._ws = None
        self._lib = None

    def prepare(self):
        self._lib = _lib.load()
        self._release()
        hd = self.dims // self.head
        sd = dict(self.state_dict())
        # compute_f(mask=None) with the reference's own ops (model.py:191-194, essentials.py:237-240)
        g = torch.pow(torch.tensor(8000.0 / 200.0), torch.linspace(0, 1, hd // 2)) * 200.0 / 1000
        sd["__rot_freqs"] = (200 * g / 1000).float()
        n, names, ptrs, nums, keep = _lib.state_dict_arrays(sd)
        h = C.c_void_p()
        _lib.check(self._lib.asrb_attention_create(self.dims, self.head, _lib.F32, n, names, ptrs, nums, C.byref(h)),
                   "asrb_attention_create")
        self._handle = h
        return self

    def _release(self):
        if self._handle is not None:
            self._lib.asrb_attention_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._release() if self._lib is not None else None
        return r

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.AsrbError("AudioAttention needs a CUDA tensor: there is no CPU path")
        x = x.float().contiguous()
        B, T, D = x.shape
        with torch.cuda.device(x.device):
            if self._handle is None:
                self.prepare()
            out = torch.empty_like(x)
            need = self._lib.asrb_attention_workspace_bytes(self._handle, B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=x.device)
            _lib.check(self._lib.asrb_attention_forward(self._handle, x.data_ptr(), B, T, out.data_ptr(),
                                                        self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                       "asrb_attention_forward")
        return out
