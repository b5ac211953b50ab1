"""Oracle: the ``attention`` + ``rotary`` block applied to encoded audio (secondary
rows a11/a12 of SURVEY.md section 8; TEST INFRASTRUCTURE, see oracle/__init__.py).

Only the branch that is live and deterministic in the reference is restated:
``attention.forward(x, xa=None, mask=None, pt=None, pitch_bias=None)`` with
``n_type="rmsnorm"`` (model.py:258-262, 302-307, 316-317) and
``rotary.forward(x, xa, mask=None)`` (model.py:191-214).  The reference's rotary
broadcast ``[B,H,T,hd/2] * [B,T,hd/2]`` is only well defined for B == 1
(SURVEY.md section 8a row a12), so the batched semantics here are "B=1 applied to each
utterance" -- the loop below is that definition.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _rms_norm(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    # nn.RMSNorm(normalized_shape, eps=None): eps = finfo(dtype).eps  (essentials.py:207)
    return F.rms_norm(x, (x.shape[-1],), w, None)


def rotary_freqs(head_dim: int) -> torch.Tensor:
    """``compute_f(x=None, mask=None)`` model.py:191-194 with ``gammatone``
    essentials.py:237-240: ``200 * (40**linspace(0,1,hd/2) * 200 / 1000) / 1000``."""
    g = torch.pow(torch.tensor(8000.0 / 200.0), torch.linspace(0, 1, head_dim // 2)) * 200.0 / 1000
    return 200 * g / 1000


def rotary_apply(x: torch.Tensor, xa: torch.Tensor) -> torch.Tensor:
    """``rotary.forward(x, xa, mask=None)`` model.py:198-214 for ONE utterance:
    ``x [1, H, T, hd]``, ``xa [1, T, D]``.  Adjacent pairs ``(x_2j, x_2j+1)`` are a
    complex number multiplied by ``polar(||xa_t||_2, t f_j)``."""
    assert x.shape[0] == 1 and xa.shape[0] == 1
    T, hd = x.shape[2], x.shape[3]
    t = torch.arange(T, dtype=torch.float32)
    ang = torch.einsum("i,j->ij", t, rotary_freqs(hd))            # [T, hd/2]
    m = torch.norm(xa, dim=-1, keepdim=True)                      # [1, T, 1]
    f = torch.polar(m.expand(1, T, hd // 2).contiguous(), ang.unsqueeze(0).contiguous())
    xc = torch.view_as_complex(x.float().reshape(1, x.shape[1], T, hd // 2, 2).contiguous()) * f
    return torch.view_as_real(xc).flatten(-2).type_as(x)


def attention_forward(sd: SD, x: torch.Tensor, head: int, xa: torch.Tensor = None) -> torch.Tensor:
    """``attention.forward`` live branch, ``x [B, T, D]`` -> ``[B, T, D]``.  With ``xa [B, Tk, D]`` the keys and values
    come from ``xa`` (model.py:259: ``k, v = n.kv(aorb(xa, x))``) and the rotary magnitudes of ``k`` from ``xa``
    (model.py:306); ``q`` keeps ``x`` for both."""
    B, T, D = x.shape
    hd = D // head
    scale = hd ** -0.25                                           # model.py:239
    outs = []
    for b in range(B):
        xb = x[b:b + 1]
        ab = xb if xa is None else xa[b:b + 1]
        Tk = ab.shape[1]
        kv = F.linear(_rms_norm(ab, sd["kv.0.weight"]), sd["kv.1.weight"], sd["kv.1.bias"])
        k, v = kv.view(1, Tk, 2, head, hd).permute(2, 0, 3, 1, 4)  # 'b c (kv h d) -> kv b h c d'
        q = F.linear(_rms_norm(xb, sd["q.0.weight"]), sd["q.1.weight"], sd["q.1.bias"])
        q = q.view(1, T, head, hd).transpose(1, 2)
        q = rotary_apply(q * scale, xb)                           # model.py:303-306
        k = rotary_apply(k * scale, ab)
        qn, kn = _rms_norm(q, sd["ln.weight"]), _rms_norm(k, sd["ln.weight"])
        s = torch.matmul(qn, kn.transpose(-1, -2)) / math.sqrt(hd)   # SDPA default scale, :307
        a = torch.matmul(torch.softmax(s, dim=-1), v)
        a = a.transpose(1, 2).reshape(1, T, D)                    # 'b h c d -> b c (h d)'
        outs.append(F.linear(a, sd["out.1.weight"], sd["out.1.bias"]))
    return torch.cat(outs, dim=0)


def random_attention_state_dict(dims: int, head: int, seed: int = 0, perturb: bool = True) -> SD:
    """Weights with the reference's key names (``attention(dims, head, layer,
    n_type="rmsnorm")`` model.py:234-252) and default-init scales."""
    gen = torch.Generator().manual_seed(seed)
    hd = dims // head

    def U(shape, bound):
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound

    b = 1.0 / math.sqrt(dims)
    ones = (lambda n: 1.0 + U((n,), 0.3)) if perturb else (lambda n: torch.ones(n))
    return {
        "q.0.weight": ones(dims), "q.1.weight": U((dims, dims), b), "q.1.bias": U((dims,), b),
        "kv.0.weight": ones(dims), "kv.1.weight": U((2 * dims, dims), b), "kv.1.bias": U((2 * dims,), b),
        "c.0.weight": ones(dims), "c.1.weight": U((dims, dims), b), "c.1.bias": U((dims,), b),
        "out.1.weight": U((dims, dims), b), "out.1.bias": U((dims,), b),
        "ln.weight": ones(hd),
        "rot.lin.weight": U((hd // 2, dims), b), "rot.lin.bias": U((hd // 2,), b),
    }


def residual_mlp_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """``residual.mlp`` (model.py:573-574): ``ln -> tgate -> Linear(D, 3 D) -> GELU -> Linear(3 D, D) -> ln`` with ONE shared
    ``nn.RMSNorm`` (``n.ln``) at both ends and ``tgate`` (model.py:525-535) = ``sum_i softmax(cs(x))_i * sigmoid(ga_i(x))``
    -- a gate that does not multiply its input.  ``x [B, T, D]`` -> ``[B, T, D]`` (the caller adds the residual,
    model.py:583)."""
    n_types = sd["mlp.1.cs.0.weight"].shape[0]
    h = _rms_norm(x, sd["ln.weight"])
    types = torch.softmax(F.linear(h, sd["mlp.1.cs.0.weight"], sd["mlp.1.cs.0.bias"]), dim=-1)
    ga = torch.stack([torch.sigmoid(F.linear(h, sd[f"mlp.1.ga.{i}.0.weight"], sd[f"mlp.1.ga.{i}.0.bias"]))
                      for i in range(n_types)], dim=-1)
    t = torch.sum(ga * types.unsqueeze(2), dim=-1)
    y = F.linear(F.gelu(F.linear(t, sd["mlp.2.weight"], sd["mlp.2.bias"])), sd["mlp.4.weight"], sd["mlp.4.bias"])
    return _rms_norm(y, sd["ln.weight"])


def random_mlp_state_dict(dims: int, num_types: int = 3, seed: int = 0) -> SD:
    """The ``ln`` / ``mlp`` entries of a reference ``residual`` state_dict (default-init scales, perturbed norm weight)."""
    gen = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound

    b = 1.0 / math.sqrt(dims)
    sd = {"ln.weight": 1.0 + U((dims,), 0.3)}
    for i in range(num_types):
        sd[f"mlp.1.ga.{i}.0.weight"], sd[f"mlp.1.ga.{i}.0.bias"] = U((dims, dims), b), U((dims,), b)
    sd["mlp.1.cs.0.weight"], sd["mlp.1.cs.0.bias"] = U((num_types, dims), b), U((num_types,), b)
    sd["mlp.2.weight"], sd["mlp.2.bias"] = U((dims * num_types, dims), b), U((dims * num_types,), b)
    b2 = 1.0 / math.sqrt(dims * num_types)
    sd["mlp.4.weight"], sd["mlp.4.bias"] = U((dims, dims * num_types), b2), U((dims,), b2)
    return sd
