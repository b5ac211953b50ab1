"""asr_model_b200 -- B200-native (sm_100a) log-mel front end + AudioEncoder forward
for sine2pi/ASR-model, behind the reference's own Python API.

Public surface (mirrors the reference):
  log_mel(wave, n_mels, n_fft, ...)      batched essentials.py:469-491
  extract_features(batch, tokenizer, spectrogram=True, ...)   per-utterance drop-in
  AudioEncoder(mels, dims, head, layer, act, n_type, norm=False, enc=False)
                                          model.py:120-169, same state_dict keys
  AudioAttention(dims, head)             model.py:234-317 live branch (secondary), K|V cache of the encoded audio
  ResidualMLP(dims, num_types)           model.py:573-574 residual.mlp (tgate + Linear-GELU-Linear between one shared RMSNorm)
  ShardedEncoder / gather_outputs        utterance-sharded multi-GPU driver

Everything numeric runs in hand-written CUDA behind the C ABI of
``libasrb200.so`` (include/asrb200.h).  There is no CPU or PyTorch fallback: a
missing library or a non-sm_100 device raises.
"""
import importlib as _importlib

__version__ = "0.1.0"

_LAZY = {
    "log_mel": "frontend", "extract_features": "frontend", "LogMel": "frontend", "waveform_feature": "frontend",
    "AudioEncoder": "encoder", "AudioAttention": "attention", "ResidualMLP": "attention",
    "ShardedEncoder": "sharded", "gather_outputs": "sharded", "shard_range": "sharded",
    "lib": "_lib", "synth": "synth",
}


def __getattr__(name):
    if name in _LAZY:
        mod = _importlib.import_module(f"asr_model_b200.{_LAZY[name]}")
        return mod if name in ("synth",) else (getattr(mod, name) if hasattr(mod, name) else mod)
    raise AttributeError(name)
