"""CPU oracle for the log-mel + AudioEncoder hot path of sine2pi/ASR-model.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and by nothing else.  The product
path (``asr_model_b200``) never imports it and fails loudly when its CUDA
library is missing.

What it is: a plain-PyTorch (CPU, fp32 or fp64) restatement of the arithmetic
the reference performs on this path, function by function, each citing the
reference ``file:line`` it follows (paths are relative to the reference checkout;
``ta:`` = the installed torchaudio 2.11.0, ``torch:`` = torch 2.11.0).

Pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run in the build container by ``oracle/pin_against_reference.py`` (which
imports the unmodified ``model.py`` / ``essentials.py`` from ``/root/reference``
with four non-arithmetic packages stubbed).  That script writes the small
fixtures under ``tests/golden/`` that the CPU test-suite replays on every run;
the fixtures travel to the GPU box, the reference checkout does not.
"""

from .logmel import (  # noqa: F401
    hann_periodic,
    melscale_fbanks_htk,
    power_spectrogram,
    log_mel_utterance,
    log_mel_batch,
    collate_spectrograms,
    waveform_feature,
)
from .encoder import (  # noqa: F401
    sinusoids,
    fold_weight_norm,
    conv_lite,
    encoder_layer,
    transformer_encoder_layer,
    audio_encoder_forward,
    encoder_state_dict_spec,
    random_encoder_state_dict,
)
from .attention import (rotary_apply, attention_forward, random_attention_state_dict,  # noqa: F401
                        residual_mlp_forward, random_mlp_state_dict)
