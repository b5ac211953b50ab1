"""ctypes binding of libasrb200.so (include/asrb200.h).  Loading is lazy; a missing
library or a failing call raises -- there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasrb200.so")

OK, F32, BF16, F16 = 0, 0, 1, 2
_lib = None


class AsrbError(RuntimeError):
    pass


class EncoderConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("mels", "dims", "head", "layer", "enc", "ffn", "compute", "reserved")]


# every symbol include/asrb200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32, _sz, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t, C.c_int
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "asrb_version": (_int, []),
    "asrb_operand_format": (_int, []),
    "asrb_last_error": (C.c_char_p, []),
    "asrb_device_check": (_int, [_int]),
    "asrb_logmel_plan_create": (_int, [_int, _int, _int, _vp, _vp, _pp]),
    "asrb_logmel_plan_destroy": (None, [_vp]),
    "asrb_logmel_num_frames": (_i64, [_vp, _i64]),
    "asrb_logmel_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "asrb_logmel_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "asrb_waveform_pool_f32": (_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "asrb_logmel_waveform_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "asrb_encoder_create": (_int, [C.POINTER(EncoderConfig), _int, C.POINTER(C.c_char_p), _pp,
                                   C.POINTER(_i64), _pp]),
    "asrb_encoder_destroy": (None, [_vp]),
    "asrb_encoder_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "asrb_encoder_forward": (_int, [_vp, _vp, _i64, _i32, _i64, _vp, _int, _vp, _sz, _vp]),
    "asrb_encoder_forward_ragged": (_int, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _int, _vp, _sz, _vp]),
    "asrb_encoder_forward_streams": (_int, [_vp, _i32, _pp, C.POINTER(_i32), _i64, _i64, _vp, _int, _vp, _sz, _vp]),
    "asrb_pcm_to_hidden_workspace_bytes": (_sz, [_vp, _vp, _i64, _i64]),
    "asrb_pcm_to_hidden": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "asrb_pcm_to_hidden_ragged": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "asrb_attention_create": (_int, [_i32, _i32, _int, _int, C.POINTER(C.c_char_p), _pp,
                                     C.POINTER(_i64), _pp]),
    "asrb_attention_destroy": (None, [_vp]),
    "asrb_attention_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "asrb_attention_forward": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "asrb_attention_kv_bytes": (_sz, [_vp, _i64, _i64]),
    "asrb_attention_encode_kv": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "asrb_attention_forward_cached": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "asrb_mlp_create": (_int, [_i32, _i32, _int, C.POINTER(C.c_char_p), _pp, C.POINTER(_i64), _pp]),
    "asrb_mlp_destroy": (None, [_vp]),
    "asrb_mlp_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "asrb_mlp_forward": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _sz, _vp]),
    "asrb_profile_begin": (_int, []),
    "asrb_profile_end": (_int, []),
    "asrb_profile_get": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_double),
                                C.POINTER(C.c_double)]),
    "asrb_test_attention_tc": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp]),
    "asrb_test_gemm_tc": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _int, _int, _int,
                                 _vp, _vp, _int, _int, _vp, _vp]),
}


def load():
    """dlopen the in-tree library and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AsrbError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a). "
                        "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError = ABI drift: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.asrb_version() != 100:
        raise AsrbError(f"libasrb200.so version {lib.asrb_version()} != header 100")
    _lib = lib
    return lib


def operand_dtype():
    """torch dtype of the tensor-core variant's MMA operands (include/asrb200.h: asrb_operand_format)."""
    import torch
    return torch.float16 if load().asrb_operand_format() == F16 else torch.bfloat16


def check(code: int, what: str = ""):
    if code != OK:
        msg = load().asrb_last_error()
        raise AsrbError(f"{what} failed ({code}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def state_dict_arrays(sd):
    """Reference state_dict -> (n, names[], host fp32 pointers[], numels[], keepalive)."""
    import torch
    keep, names, ptrs, nums = [], [], [], []
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue                      # num_batches_tracked
        t = v.detach().to("cpu", torch.float32).contiguous()
        keep.append(t)
        names.append(k.encode())
        ptrs.append(t.data_ptr())
        nums.append(t.numel())
    n = len(names)
    return (n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs), (_i64 * n)(*nums), keep)


def profile_records():
    """Stop profiling and return [(tag, ms, flops, bytes), ...] in launch order."""
    lib = load()
    n = lib.asrb_profile_end()
    if n < 0:
        check(n, "asrb_profile_end")
    out = []
    for i in range(n):
        tag, ms, fl, by = C.c_char_p(), C.c_float(), C.c_double(), C.c_double()
        check(lib.asrb_profile_get(i, C.byref(tag), C.byref(ms), C.byref(fl), C.byref(by)), "asrb_profile_get")
        out.append((tag.value.decode(), ms.value, fl.value, by.value))
    return out
