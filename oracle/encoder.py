"""Oracle: ``AudioEncoder`` forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

A functional, eval-mode restatement of model.py:93-169 driven by a reference
``state_dict`` (key names and shapes are the reference's: SURVEY.md section 8b).  It uses
the same ATen primitives the reference modules dispatch to (``F.conv1d``,
``F.layer_norm``, ``F.batch_norm``, ``F.gelu`` exact-erf, ``F.glu``, ``F.silu``,
``F.multi_head_attention_forward`` math) so fp32 results agree with the reference
module to rounding; ``oracle/pin_against_reference.py`` checks that.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

THETA = 30000.0          # model.py:26
FFN_WIDTH = 2048         # nn.TransformerEncoderLayer default dim_feedforward (model.py:138)
LN_EPS = 1e-5            # essentials.py:103 ; nn.TransformerEncoderLayer default
BN_EPS = 1e-5            # nn.BatchNorm1d default (model.py:103)

SD = Dict[str, torch.Tensor]


def sinusoids(ctx: int, dims: int, theta: float = THETA, dtype=torch.float32) -> torch.Tensor:
    """essentials.py:354-358: ``[sin(t s_j) || cos(t s_j)]``, ``s_j = exp(-ln(theta)/(dims/2-1) j)``."""
    tscales = torch.exp(-torch.log(torch.tensor(float(theta))) / (dims // 2 - 1)
                        * torch.arange(dims // 2, dtype=torch.float32))
    scaled = torch.arange(ctx, dtype=torch.float32).unsqueeze(1) * tscales.unsqueeze(0)
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=1).to(dtype)


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """``weight_norm`` parametrisation (model.py:143): ``W = g * v / ||v||`` with the
    norm over every dim but 0 (torch ``_weight_norm``, dim=0)."""
    norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
    return v * (g / norm)


def conv_lite(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """``ConvLite.forward`` model.py:109-118 on ``x [B, D, T]`` (eval: Dropout off,
    BatchNorm uses running statistics)."""
    D = x.shape[1]
    r = x
    y = F.conv1d(x, sd[p + "point1.weight"], sd[p + "point1.bias"])              # :111
    y = F.glu(y, dim=1)                                                           # :112
    y = F.conv1d(y, sd[p + "depth.weight"], sd[p + "depth.bias"], padding=7, groups=D)  # :113
    y = F.batch_norm(y, sd[p + "bn.running_mean"], sd[p + "bn.running_var"],
                     sd[p + "bn.weight"], sd[p + "bn.bias"], training=False, eps=BN_EPS)  # :114
    y = F.silu(y)                                                                 # :115
    y = F.conv1d(y, sd[p + "point2.weight"], sd[p + "point2.bias"])              # :116
    return r + y                                                                  # :118


def encoder_layer(sd: SD, i: int, x: torch.Tensor) -> torch.Tensor:
    """One ``nn.Sequential`` of model.py:142-147 on ``x [B, D, T]``."""
    D = x.shape[1]
    p = f"encoder.{i}."
    y = F.gelu(x)                                                                 # act_fn
    w = fold_weight_norm(sd[p + "1.parametrizations.weight.original0"],
                         sd[p + "1.parametrizations.weight.original1"])
    y = F.conv1d(y, w, sd[p + "1.bias"], padding=1)                              # weight_norm(Conv1d k3)
    # channel LayerNorm, essentials.py:102-113: transpose, layer_norm over D, transpose back
    y = F.layer_norm(y.transpose(1, -1), (D,), sd[p + "2.gamma"], sd[p + "2.beta"], LN_EPS).transpose(1, -1)
    y = conv_lite(sd, p + "3.", y)
    y = F.gelu(y)
    y = F.conv1d(y, sd[p + "5.weight"], sd[p + "5.bias"], padding=1, groups=D)   # depthwise k3
    return F.gelu(y)


def transformer_encoder_layer(sd: SD, x: torch.Tensor, head: int, p: str = "EncoderLayer.") -> torch.Tensor:
    """``nn.TransformerEncoderLayer(d_model, nhead, batch_first=True)`` in eval mode
    (model.py:138,163): post-norm, ReLU FFN of width 2048, no mask, dropout off.
    ``x [B, T, D]``."""
    B, T, D = x.shape
    hd = D // head
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, T, head, hd).transpose(1, 2)
    k = k.view(B, T, head, hd).transpose(1, 2)
    v = v.view(B, T, head, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    a = torch.matmul(torch.softmax(s, dim=-1), v)
    a = a.transpose(1, 2).reshape(B, T, D)
    a = F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + a, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    f = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(x + f, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)


def audio_encoder_forward(sd: SD, x: torch.Tensor, head: int, enc: Optional[bool] = None) -> torch.Tensor:
    """``AudioEncoder._process_feature`` model.py:149-163 with ``norm=False``:
    ``x [B, mels, T]`` (or ``[mels, T]``) -> ``[B, T, D]``.  A single input channel
    selects ``conv2`` (model.py:152-155)."""
    if x.dim() == 2:
        x = x.unsqueeze(0)
    stem = "conv1.0." if x.shape[1] > 1 else "conv2.0."
    x = F.conv1d(x, sd[stem + "weight"], sd[stem + "bias"], padding=1)
    n_layer = 1 + max([int(k.split(".")[1]) for k in sd if k.startswith("encoder.")], default=-1)
    for i in range(n_layer):
        x = encoder_layer(sd, i, x)
    x = x.permute(0, 2, 1).contiguous()
    x = x + sinusoids(x.shape[1], x.shape[-1], dtype=x.dtype)
    if enc is None:
        enc = "EncoderLayer.linear1.weight" in sd
    return transformer_encoder_layer(sd, x, head) if enc else x


def encoder_state_dict_spec(mels: int, dims: int, layer: int, enc: bool) -> Dict[str, tuple]:
    """Key -> shape of ``AudioEncoder(mels, dims, head, layer, ..., norm=False, enc=enc)``
    (verified against the instantiated reference by the pin script)."""
    D = dims
    spec = {"conv1.0.weight": (D, mels, 3), "conv1.0.bias": (D,),
            "conv2.0.weight": (D, 1, 3), "conv2.0.bias": (D,)}
    if enc:
        e = "EncoderLayer."
        spec.update({e + "self_attn.in_proj_weight": (3 * D, D), e + "self_attn.in_proj_bias": (3 * D,),
                     e + "self_attn.out_proj.weight": (D, D), e + "self_attn.out_proj.bias": (D,),
                     e + "linear1.weight": (FFN_WIDTH, D), e + "linear1.bias": (FFN_WIDTH,),
                     e + "linear2.weight": (D, FFN_WIDTH), e + "linear2.bias": (D,),
                     e + "norm1.weight": (D,), e + "norm1.bias": (D,),
                     e + "norm2.weight": (D,), e + "norm2.bias": (D,)})
    for i in range(layer):
        p = f"encoder.{i}."
        spec.update({p + "1.bias": (D,),
                     p + "1.parametrizations.weight.original0": (D, 1, 1),
                     p + "1.parametrizations.weight.original1": (D, D, 3),
                     p + "2.gamma": (D,), p + "2.beta": (D,),
                     p + "3.point1.weight": (2 * D, D, 1), p + "3.point1.bias": (2 * D,),
                     p + "3.depth.weight": (D, 1, 15), p + "3.depth.bias": (D,),
                     p + "3.bn.weight": (D,), p + "3.bn.bias": (D,),
                     p + "3.bn.running_mean": (D,), p + "3.bn.running_var": (D,),
                     p + "3.bn.num_batches_tracked": (),
                     p + "3.point2.weight": (D, D, 1), p + "3.point2.bias": (D,),
                     p + "5.weight": (D, 1, 3), p + "5.bias": (D,)})
    return spec


def random_encoder_state_dict(mels: int, dims: int, layer: int, enc: bool, seed: int = 0,
                              perturb: bool = False) -> SD:
    """Deterministic random-init weights with the reference's shapes and PyTorch's
    default init SCALES (uniform(+-1/sqrt(fan_in)) for conv/linear weights and biases,
    xavier-uniform ``in_proj_weight``, weight-norm ``g = ||v||``).  ``perturb=True``
    also randomises what default init leaves trivial -- BatchNorm running stats and
    affine, LayerNorm gamma/beta, weight-norm g -- so folding bugs are visible
    (SURVEY.md section 8d)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def U(shape, bound):
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound

    for k, shape in encoder_state_dict_spec(mels, dims, layer, enc).items():
        leaf = k.rsplit(".", 1)[-1]
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("in_proj_weight"):
            sd[k] = U(shape, math.sqrt(6.0 / (shape[0] + shape[1])))
        elif k.endswith("in_proj_bias") or k.endswith("out_proj.bias"):
            sd[k] = U(shape, 0.02) if perturb else torch.zeros(shape)
        elif k.endswith("original0"):
            sd[k] = torch.zeros(shape)            # filled below from v
        elif leaf in ("weight", "original1") and len(shape) >= 2:
            fan_in = int(torch.tensor(shape[1:]).prod())
            sd[k] = U(shape, 1.0 / math.sqrt(fan_in))
        elif leaf == "bias" and ("bn." not in k and "norm" not in k):
            # conv / linear bias: fan_in of the matching weight
            wk = k[:-4] + "weight"
            if wk not in sd and k.endswith("1.bias") and k.startswith("encoder."):
                fan_in = dims * 3
            else:
                fan_in = int(torch.tensor(sd[wk].shape[1:]).prod())
            sd[k] = U(shape, 1.0 / math.sqrt(fan_in))
        elif leaf in ("gamma",) or (leaf == "weight" and len(shape) == 1):
            sd[k] = 1.0 + U(shape, 0.5) if perturb else torch.ones(shape)
        elif leaf in ("beta",) or (leaf == "bias"):
            sd[k] = U(shape, 0.3) if perturb else torch.zeros(shape)
        elif leaf == "running_mean":
            sd[k] = torch.randn(shape, generator=gen) * 0.1 if perturb else torch.zeros(shape)
        elif leaf == "running_var":
            sd[k] = 0.5 + torch.rand(shape, generator=gen) if perturb else torch.ones(shape)
        else:
            raise KeyError(k)
    for i in range(layer):
        p = f"encoder.{i}.1.parametrizations.weight."
        g = sd[p + "original1"].flatten(1).norm(dim=1).view(-1, 1, 1)
        if perturb:
            g = g * (0.5 + torch.rand(g.shape, generator=gen))
        sd[p + "original0"] = g
    return sd
