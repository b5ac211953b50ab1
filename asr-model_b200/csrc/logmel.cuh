// Internal view of the front-end plan, shared by logmel.cu and the fused pcm->hidden path.
#pragma once
#include "common.cuh"

struct asrb_logmel_plan {
    int n_fft, hop, n_mels, kmax, radix;
    float* d_window; float2* d_twiddle; int* d_lo; int* d_cnt; float* d_w;
};

namespace asrb {
int logmel_tile_frames(const asrb_logmel_plan* pl);                                    // frames per tile of pass 1
size_t logmel_keys_words(const asrb_logmel_plan* pl, int64_t batch, int64_t n_samples); // uint32 words behind `keys`
// Pass 1 of the front end: out = (log10(max(mel,1e-10)) + 4) / 4 without the dynamic-range
// floor, keys[b] = order-preserving image of max_t,m log10(mel).  The floor is applied by
// logmel_floor_kernel (asrb_logmel_f32) or on the fly by launch_to_channels_last.
// With out_cl the values go out as op16 (16-bit operand format) channels-last [B][T][CP] (channels >= n_mels zero) instead of `out`.
// With pool_out [batch][pool_target] the same pass also emits the average-pooled `waveform` feature (essentials.py:493-503).
int logmel_pass1(const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                 int64_t stride, const int32_t* lengths, float* out, uint32_t* keys, cudaStream_t st,
                 op16* out_cl = nullptr, int CP = 0, float* pool_out = nullptr, int64_t pool_target = 0);
int logmel_floor_cl(const asrb_logmel_plan* pl, op16* a, int CP, const uint32_t* keys, const int32_t* lengths,
                    int64_t batch, int64_t n_samples, cudaStream_t st);
}
