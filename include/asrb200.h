/* asrb200.h -- C ABI of libasrb200.so: the B200 (sm_100a) implementation of the
 * log-mel front end and AudioEncoder forward of sine2pi/ASR-model.
 *
 * The reference has no FFI of its own (it is pure Python); each entry point below
 * names the reference interface it replaces.  A binding (ctypes, cffi, pybind, cgo...)
 * passes plain device/host pointers and sizes -- no torch types cross this boundary.
 *
 * Conventions
 *   - every call returns int: 0 = OK, negative = error (ASRB_E_*); the message is
 *     available from asrb_last_error() (thread-local, valid until the next call);
 *   - nothing here throws, and the compute calls never allocate or free device memory:
 *     tensors are caller-owned, scratch comes from an explicit workspace whose size the
 *     *_workspace_bytes() functions report (create/destroy calls own the constants
 *     they upload);
 *   - compute calls are asynchronous on the given stream (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream) and are CUDA-graph capturable;
 *   - a handle may be used from one stream at a time; there is no mutable global state;
 *   - there is NO CPU fallback and no other GPU architecture: a device whose compute
 *     capability is not 10.x yields ASRB_E_DEVICE.
 */
#ifndef ASRB200_H_
#define ASRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASRB_VERSION 100          /* 0.1.0 */

#define ASRB_OK            0
#define ASRB_E_ARG        -1      /* bad argument (NULL, shape, unsupported size)      */
#define ASRB_E_DEVICE     -2      /* not an sm_100 device / no CUDA device              */
#define ASRB_E_WORKSPACE  -3      /* workspace NULL, misaligned or too small            */
#define ASRB_E_CUDA       -4      /* a CUDA runtime / driver call failed                */
#define ASRB_E_WEIGHTS    -5      /* a required state_dict tensor is missing / misshaped*/

#define ASRB_F32   0
#define ASRB_BF16  1
#define ASRB_F16   2

int         asrb_version(void);
/* Storage format of the tensor-core variant's MMA operands (activations between kernels and packed weights):
 * ASRB_F16 (IEEE fp16, the shipped build: same width and tensor throughput as bf16 with three more mantissa bits, which
 * is what it takes to meet allclose(atol 2e-2, rtol 1e-2) against the fp32 reference -- DESIGN.md section 5) or
 * ASRB_BF16 (a build with -DASRB_OPERAND_BF16).  Hidden states are returned as bf16 either way. */
int         asrb_operand_format(void);
const char* asrb_last_error(void);
/* ASRB_OK iff `device` (ordinal) is a compute-capability 10.x GPU. */
int         asrb_device_check(int device);

/* ------------------------------------------------------------------------------------
 * Front end.  Replaces the spectrogram branch of extract_features
 * (essentials.py:469-491: torchaudio MelSpectrogram -> clamp(1e-10).log10() ->
 * maximum(x, x.max()-8) -> (x+4)/4) and the zero right-padding of DataCollator
 * (essentials.py:555-572), batched.
 *
 * The plan holds the constants the reference rebuilds on every call: the window
 * (essentials.py:480, `window_fn=torch.hann_window`) and the HTK filterbank
 * (ta:functional/functional.py:518-587).  They are passed in as HOST arrays so the
 * binding can build them with the very torch ops the reference uses (bit-equal
 * constants); the plan converts the dense [n_fft/2+1, n_mels] filterbank to its banded
 * form and uploads it.  Supported n_fft: 400, 1024 (hop and n_mels are free).
 * ---------------------------------------------------------------------------------- */
typedef struct asrb_logmel_plan asrb_logmel_plan;

int  asrb_logmel_plan_create(int n_fft, int hop, int n_mels,
                             const float* window_host,   /* [n_fft]                    */
                             const float* fbank_host,    /* [n_fft/2+1][n_mels]        */
                             asrb_logmel_plan** plan);
void asrb_logmel_plan_destroy(asrb_logmel_plan* plan);

/* Frames produced for n samples: 1 + n / hop (center=True, essentials.py:478). */
int64_t asrb_logmel_num_frames(const asrb_logmel_plan* plan, int64_t n_samples);
size_t  asrb_logmel_workspace_bytes(const asrb_logmel_plan* plan, int64_t batch, int64_t n_samples);

/* pcm   [batch][pcm_stride] fp32 device, n_samples valid columns (|x| <= 1 like load_wave)
 * lengths  NULL, or [batch] int32 device: valid samples per utterance (<= n_samples);
 *          frames past 1 + len/hop are written as 0.0 (DataCollator's pad value)
 * out   [batch][n_mels][T] fp32 device, T = asrb_logmel_num_frames(n_samples)
 * The dynamic-range floor uses the maximum over each utterance, never over the batch. */
int asrb_logmel_f32(const asrb_logmel_plan* plan,
                    const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                    const int32_t* lengths,
                    float* out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* The `waveform` feature of extract_features (essentials.py:493-510): PCM resampled to the frame
 * rate.  n_samples > target: adaptive average pooling, out[b][i] = mean(pcm[b][floor(i n/target) : ceil((i+1) n/target)))
 * (essentials.py:503); otherwise F.interpolate(mode="linear", align_corners=False) (essentials.py:505-506).
 * `target` is computed by the binding exactly as the reference does
 * (int((n / sample_rate) * (sample_rate // hop))).  pcm [batch][pcm_stride], out [batch][target] fp32 device. */
int asrb_waveform_pool_f32(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                           int64_t target, float* out, void* stream);

/* extract_features(spectrogram=True, waveform=True) in ONE pass over the PCM: asrb_logmel_f32 plus the average-pooled
 * waveform feature taken from the samples the front-end kernel has staged in shared memory anyway.
 * pool_out [batch][pool_target] fp32 device; 0 < pool_target < n_samples and pool_target <= frames (the pooling branch:
 * essentials.py:502-503); results equal asrb_logmel_f32 + asrb_waveform_pool_f32 bit for bit.  pool_target = 0: no pooling. */
int asrb_logmel_waveform_f32(const asrb_logmel_plan* plan,
                             const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                             const int32_t* lengths, float* out, float* pool_out, int64_t pool_target,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Encoder.  Replaces AudioEncoder.__init__/forward (model.py:120-169) with norm=False:
 * conv stem, `layer` x [GELU, weight-normed Conv1d k3, channel LayerNorm, ConvLite, GELU,
 * depthwise k3, GELU], + sinusoids, optional nn.TransformerEncoderLayer (enc=1).
 *
 * Weights are handed over as the reference's own state_dict: parallel arrays of key
 * names, HOST fp32 pointers and element counts (SURVEY.md 8b lists the keys).  create()
 * folds weight-norm and eval-mode BatchNorm, repacks for the tensor-core kernels and
 * uploads; integer entries (num_batches_tracked) are ignored.
 * ---------------------------------------------------------------------------------- */
typedef struct asrb_encoder asrb_encoder;

typedef struct asrb_encoder_config {
    int32_t mels;        /* input channels of conv1 (80 | 128)                         */
    int32_t dims;        /* D; multiple of 64 (ASRB_F32) or 128 (ASRB_BF16), <= 1024   */
    int32_t head;        /* heads of the optional TransformerEncoderLayer              */
    int32_t layer;       /* number of conv blocks                                      */
    int32_t enc;         /* 1 = TransformerEncoderLayer present (model.py:138)         */
    int32_t ffn;         /* its feed-forward width (2048 in the reference)             */
    int32_t compute;     /* ASRB_BF16: tcgen05 tensor-core variant -- 16-bit MMA       */
                         /*   operands (asrb_operand_format()), fp32 accumulate, bf16  */
                         /*   hidden states;                                           */
                         /* ASRB_F32 : fp32 FFMA everywhere (the 1e-4 variant)         */
    int32_t reserved;
} asrb_encoder_config;

int  asrb_encoder_create(const asrb_encoder_config* cfg, int n_tensors,
                         const char* const* names, const float* const* host_data,
                         const int64_t* numels, asrb_encoder** enc);
void asrb_encoder_destroy(asrb_encoder* enc);
size_t asrb_encoder_workspace_bytes(const asrb_encoder* enc, int64_t batch, int64_t frames);

/* x    [batch][in_ch][frames] fp32 device; in_ch = mels selects conv1, in_ch = 1 selects
 *      conv2 (model.py:152-155)
 * out  [batch][frames][dims] device, out_dtype = ASRB_F32 | ASRB_BF16 */
int asrb_encoder_forward(asrb_encoder* enc, const float* x, int64_t batch, int32_t in_ch,
                         int64_t frames, void* out, int out_dtype,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Ragged batches (DataCollator pads every clip to the longest with 0.0, essentials.py:555-572):
 * frames [batch] int32 device = frames of each utterance that hold audio.  Output rows t < frames[b] equal
 * asrb_encoder_forward's bit for bit; rows t >= frames[b] are set to 0 instead of the encoding of the padding.  On the
 * tensor-core path without the TransformerEncoderLayer every frame tile past frames[b] + 9 * layer + 1 (the widest cone a
 * valid output frame depends on) is skipped by every kernel of the stack; with enc = 1 (attention sees every frame,
 * model.py:163) and on the fp32 path everything is computed and only the zeroing differs. */
int asrb_encoder_forward_ragged(asrb_encoder* enc, const float* x, int64_t batch, int32_t in_ch, int64_t frames_total,
                                const int32_t* frames, void* out, int out_dtype,
                                void* workspace, size_t workspace_bytes, void* stream);

/* Several feature streams of the same shape through ONE pass of the layer stack (Model.forward / generate encode
 * the TensorDict {a, b, c} = spectrogram / waveform / pitch with the same encoder, model.py:165-167, 657-665): stream s
 * is x[s] [batch][in_ch[s]][frames] (in_ch = mels -> conv1, 1 -> conv2); out is [n_streams * batch][frames][dims], stream s
 * at rows [s * batch, (s + 1) * batch).  Results equal n_streams calls of asrb_encoder_forward; the workspace is that of
 * asrb_encoder_workspace_bytes(enc, n_streams * batch, frames).  n_streams <= 8. */
int asrb_encoder_forward_streams(asrb_encoder* enc, int32_t n_streams, const float* const* x, const int32_t* in_ch,
                                 int64_t batch, int64_t frames, void* out, int out_dtype,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Fused hot path: PCM -> log-mel -> encoder without materialising the fp32 [B,M,T]
 * feature tensor unless `logmel_out` is non-NULL.  Equivalent to
 * asrb_logmel_f32 followed by asrb_encoder_forward. */
size_t asrb_pcm_to_hidden_workspace_bytes(const asrb_logmel_plan* plan, const asrb_encoder* enc,
                                          int64_t batch, int64_t n_samples);
int asrb_pcm_to_hidden(const asrb_logmel_plan* plan, asrb_encoder* enc,
                       const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                       const int32_t* lengths, float* logmel_out /* may be NULL */,
                       void* out, int out_dtype,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The same for ragged batches: lengths [batch] int32 device is required; valid rows (t < 1 + lengths[b] / hop) are bit-equal
 * to asrb_pcm_to_hidden with the same lengths, the rows of the padding are 0, and padded frame tiles are skipped from the
 * FFT to the last conv block (see asrb_encoder_forward_ragged). */
int asrb_pcm_to_hidden_ragged(const asrb_logmel_plan* plan, asrb_encoder* enc,
                              const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                              const int32_t* lengths, float* logmel_out /* may be NULL */,
                              void* out, int out_dtype,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Secondary: the `attention` block's live branch applied to encoded audio
 * (model.py:234-317 with n_type="rmsnorm", xa=None, mask=None) including `rotary`
 * (model.py:171-214).  Batched = the reference's B=1 semantics per utterance (per-sample rotary magnitudes).
 * compute = ASRB_F32 (CUDA cores, <= 1e-4) or ASRB_BF16 (tensor cores: tcgen05 projections + flash attention; dims % 128 == 0,
 * head_dim 64 or 128).
 * Keys: q.0.weight q.1.weight q.1.bias kv.0.weight kv.1.weight kv.1.bias out.1.weight
 * out.1.bias ln.weight (c.* and rot.lin.* are unused by the live branch).
 * ---------------------------------------------------------------------------------- */
typedef struct asrb_attention asrb_attention;

int  asrb_attention_create(int32_t dims, int32_t head, int compute, int n_tensors,
                           const char* const* names, const float* const* host_data,
                           const int64_t* numels, asrb_attention** att);
void asrb_attention_destroy(asrb_attention* att);
size_t asrb_attention_workspace_bytes(const asrb_attention* att, int64_t batch, int64_t frames);
/* x, out: [batch][frames][dims] fp32 device */
int asrb_attention_forward(asrb_attention* att, const float* x, int64_t batch, int64_t frames,
                           float* out, void* workspace, size_t workspace_bytes, void* stream);

/* K|V reuse (compute = ASRB_BF16 only).  The reference recomputes norm_kv -> Linear(D, 2D) -> rotary ->
 * per-head RMSNorm of the encoded audio in every `residual` call (8 per decoder block, model.py:617-626) and for every
 * generated token (model.py:691-699).  Here asrb_attention_encode_kv() does that once for xa [batch][kv_frames][dims] fp32 into a
 * caller-owned cache of asrb_attention_kv_bytes() bytes (256-B aligned; K with rotary and norm applied | V, 16-bit), and
 * asrb_attention_forward_cached() attends queries x [batch][q_frames][dims] (model.py:258-262 with xa given: rotary
 * magnitudes of q from x, of k from xa; no mask) against it.  encode_kv(x) + forward_cached(x) == asrb_attention_forward(x).
 * Workspace: asrb_attention_workspace_bytes(att, batch, frames of the call). */
size_t asrb_attention_kv_bytes(const asrb_attention* att, int64_t batch, int64_t kv_frames);
int asrb_attention_encode_kv(asrb_attention* att, const float* xa, int64_t batch, int64_t kv_frames, void* kv_cache,
                             void* workspace, size_t workspace_bytes, void* stream);
int asrb_attention_forward_cached(asrb_attention* att, const float* x, int64_t batch, int64_t q_frames,
                                  const void* kv_cache, int64_t kv_frames, float* out,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* residual.mlp applied to encoded audio (model.py:573-574, 583): shared RMSNorm -> tgate (model.py:525-535) ->
 * Linear(D, n D) -> GELU -> Linear(n D, D) -> the same RMSNorm, on the tensor cores (dims % 128 == 0).  Keys as in a reference
 * `residual` state_dict: ln.weight, mlp.1.ga.{i}.0.{weight,bias}, mlp.1.cs.0.{weight,bias}, mlp.2.{weight,bias},
 * mlp.4.{weight,bias}.  x, out: [batch][frames][dims] fp32 device; add_residual != 0: out = x + mlp(x) (model.py:583). */
typedef struct asrb_mlp asrb_mlp;
int  asrb_mlp_create(int32_t dims, int32_t n_types, int n_tensors, const char* const* names,
                     const float* const* host_data, const int64_t* numels, asrb_mlp** mlp);
void asrb_mlp_destroy(asrb_mlp* mlp);
size_t asrb_mlp_workspace_bytes(const asrb_mlp* mlp, int64_t batch, int64_t frames);
int asrb_mlp_forward(asrb_mlp* mlp, const float* x, int64_t batch, int64_t frames, int add_residual,
                     float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement aid (process-wide; off by default, not for production): between
 * asrb_profile_begin() and asrb_profile_end() every kernel launch of this library is
 * bracketed by CUDA events on its stream.  asrb_profile_end() returns the record count;
 * asrb_profile_get(i) yields the kernel tag, its device time and the ALGORITHMIC flops and
 * HBM bytes of that launch (the roofline numerators of DESIGN.md).
 * ---------------------------------------------------------------------------------- */
int asrb_profile_begin(void);
int asrb_profile_end(void);
int asrb_profile_get(int index, const char** tag, float* ms, double* flops, double* bytes);

/* ------------------------------------------------------------------------------------
 * Test hooks (exported so the GPU unit tests can check building blocks in isolation
 * through the same ABI; not needed by an integration).
 * ---------------------------------------------------------------------------------- */
/* The tcgen05/TMEM/TMA implicit-GEMM in isolation:
 *   out[b,t,:] = epilogue( sum_{tap,k} a[b, t+tap-taps/2, k] * w[n][tap*K + k] + bias[n] )
 * a [B][T][K], w [N][taps*K], out [B][T][N or N/2], res (or NULL) like out: all in the 16-bit operand format
 * (asrb_operand_format()).
 * epilogue: 0 bias+act | (1 reserved) |
 *           2 bias+res+act | 3 bias(+res)+LayerNorm(gamma, beta, eps=1e-5) |
 *           4 GLU (w rows interleaved [128 value | 128 gate] per 256) -> depthwise(dw_w [kw][N/2], dw_b)
 *             down the frames -> dw_act |
 *           5 bias+res+act -> depthwise(dw_w [kw][N], dw_b) -> dw_act (+ pos [T][N]).
 * act / dw_act: 0 none | 1 GELU | 2 ReLU | 3 SiLU | 4 GELU(GELU).  K % 64 == 0, N % 128 == 0
 * (N % 256 == 0 for 4).  The fused depthwise epilogues are built for the combinations the encoder uses:
 * 4 with kw = 15, dw_act = SiLU; 5 with kw = 3, act = GELU, dw_act in {GELU, GELU(GELU)}. */
int asrb_test_gemm_tc(const void* a, const void* w, const float* bias, const void* res,
                      const float* gamma, const float* beta, void* out,
                      int64_t B, int64_t T, int K, int N, int taps, int epilogue, int act,
                      const float* dw_w, const float* dw_b, int dw_kw, int dw_act, const float* pos,
                      void* stream);

/* The tcgen05 flash-style attention in isolation: qkv [B][T][3D] (q | k | v, heads contiguous inside each) ->
 * out [B][T][D] = softmax(q k^T / sqrt(D/H)) v per head, no mask (model.py:163); both in the 16-bit operand format.
 * D/H must be 64 or 128. */
int asrb_test_attention_tc(const void* qkv, void* out, int64_t B, int64_t T, int D, int H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASRB200_H_ */
