"""``AudioAttention``: the ``attention`` block's live, deterministic branch applied to encoded
audio (model.py:234-317 with ``n_type="rmsnorm"``, ``xa=None, mask=None, pt=None``) including
``rotary`` (model.py:171-214).  Parameter names match the reference module so its
``state_dict`` loads unchanged.  Batched = the reference's B=1 semantics per utterance
(its own broadcast is only defined for B == 1).

``compute="fp32"`` is the CUDA-core variant (<= 1e-4 of the reference); ``compute="bf16"`` the tensor-core variant, which
also exposes the K|V reuse of SURVEY.md section 8f rank 3: ``kv = att.encode_kv(audio)`` once, then ``att(x, kv=kv)`` for every
query sequence that attends the same encoded audio (the reference recomputes the K|V branch on every call)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class _Named(nn.Sequential):
    pass


class AudioAttention(nn.Module):
    def __init__(self, dims: int, head: int, layer: int = 1, n_type: str = "rmsnorm", compute: str = "fp32"):
        super().__init__()
        if n_type != "rmsnorm":
            raise NotImplementedError("only n_type='rmsnorm' is deterministic in the reference")
        if compute not in ("fp32", "bf16"):
            raise ValueError("compute must be 'fp32' or 'bf16'")
        self.dims, self.head, self.compute = dims, head, compute
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
        _lib.check(self._lib.asrb_attention_create(self.dims, self.head, _lib.BF16 if self.compute == "bf16" else _lib.F32,
                                                   n, names, ptrs, nums, C.byref(h)),
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

    def _workspace(self, device, B, T):
        need = self._lib.asrb_attention_workspace_bytes(self._handle, B, T)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=device)
        return self._ws

    def encode_kv(self, xa: torch.Tensor) -> torch.Tensor:
        """K (rotary + per-head norm applied) | V of ``xa [B, Tk, D]`` as an opaque 16-bit ``[B, Tk, 2 D]`` cache tensor."""
        if self.compute != "bf16":
            raise _lib.AsrbError("the K|V cache belongs to the tensor-core variant (compute='bf16')")
        if not xa.is_cuda:
            raise _lib.AsrbError("AudioAttention needs a CUDA tensor: there is no CPU path")
        xa = xa.float().contiguous()
        B, Tk, D = xa.shape
        with torch.cuda.device(xa.device):
            if self._handle is None:
                self.prepare()
            kv = torch.empty(B, Tk, 2 * D, dtype=_lib.operand_dtype(), device=xa.device)
            ws = self._workspace(xa.device, B, Tk)
            _lib.check(self._lib.asrb_attention_encode_kv(self._handle, xa.data_ptr(), B, Tk, kv.data_ptr(), ws.data_ptr(),
                                                          ws.numel(), _lib.stream_ptr()), "asrb_attention_encode_kv")
        return kv

    def forward(self, x: torch.Tensor, kv: torch.Tensor = None) -> torch.Tensor:
        """``x [B, T, D]`` -> ``[B, T, D]`` fp32.  ``kv``: a cache from ``encode_kv`` -- the queries of ``x`` then attend the
        cached sequence instead of ``x`` itself (model.py:258-262 with ``xa`` given)."""
        if not x.is_cuda:
            raise _lib.AsrbError("AudioAttention needs a CUDA tensor: there is no CPU path")
        x = x.float().contiguous()
        B, T, D = x.shape
        with torch.cuda.device(x.device):
            if self._handle is None:
                self.prepare()
            out = torch.empty_like(x)
            ws = self._workspace(x.device, B, T)
            if kv is None:
                _lib.check(self._lib.asrb_attention_forward(self._handle, x.data_ptr(), B, T, out.data_ptr(),
                                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                           "asrb_attention_forward")
            else:
                if kv.shape[0] != B or kv.shape[2] != 2 * D or not kv.is_contiguous():
                    raise ValueError("kv must be the [B, Tk, 2 D] tensor encode_kv returned for the same batch")
                _lib.check(self._lib.asrb_attention_forward_cached(self._handle, x.data_ptr(), B, T, kv.data_ptr(), kv.shape[1],
                                                                   out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                           "asrb_attention_forward_cached")
        return out


class _TGate(nn.Module):                                # parameter names of model.py:525-530
    def __init__(self, dims, num_types):
        super().__init__()
        self.ga = nn.ModuleList([nn.Sequential(nn.Linear(dims, dims), nn.Identity()) for _ in range(num_types)])
        self.cs = nn.Sequential(nn.Linear(dims, num_types), nn.Identity())


class ResidualMLP(nn.Module):
    """``residual.mlp`` of the reference applied to encoded audio (model.py:573-574, 583): shared RMSNorm -> ``tgate`` ->
    ``Linear(D, n D)`` -> GELU -> ``Linear(n D, D)`` -> the same RMSNorm, on the tensor cores.  Parameter names are those of a
    reference ``residual`` module (``ln.weight``, ``mlp.1.ga.i.0.weight``, ``mlp.1.cs.0.weight``, ``mlp.2.weight``,
    ``mlp.4.weight``, ...), so ``load_state_dict(residual.state_dict(), strict=False)`` picks them up.
    ``forward(x, add_residual=True)`` returns ``x + mlp(x)`` like ``residual.forward``'s last line."""

    def __init__(self, dims: int, num_types: int = 3, act: str = "gelu"):
        super().__init__()
        if act != "gelu":
            raise NotImplementedError("only act='gelu' (the reference configuration)")
        self.dims, self.num_types = dims, num_types
        self.ln = nn.RMSNorm(dims)
        self.mlp = nn.Sequential(nn.Identity(), _TGate(dims, num_types), nn.Linear(dims, dims * num_types), nn.Identity(),
                                 nn.Linear(dims * num_types, dims), nn.Identity())
        self._handle, self._ws, self._lib = None, None, None

    def _release(self):
        if self._handle is not None:
            self._lib.asrb_mlp_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = set(self.state_dict())
        r = super().load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=strict, assign=assign)
        self._release() if self._lib is not None else None
        return r

    def prepare(self):
        self._lib = _lib.load()
        self._release()
        n, names, ptrs, nums, keep = _lib.state_dict_arrays(dict(self.state_dict()))
        h = C.c_void_p()
        _lib.check(self._lib.asrb_mlp_create(self.dims, self.num_types, n, names, ptrs, nums, C.byref(h)), "asrb_mlp_create")
        self._handle = h
        return self

    def forward(self, x: torch.Tensor, add_residual: bool = False) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.AsrbError("ResidualMLP needs a CUDA tensor: there is no CPU path")
        x = x.float().contiguous()
        B, T, D = x.shape
        with torch.cuda.device(x.device):
            if self._handle is None:
                self.prepare()
            out = torch.empty_like(x)
            need = self._lib.asrb_mlp_workspace_bytes(self._handle, B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=x.device)
            _lib.check(self._lib.asrb_mlp_forward(self._handle, x.data_ptr(), B, T, int(add_residual), out.data_ptr(),
                                                  self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()), "asrb_mlp_forward")
        return out
