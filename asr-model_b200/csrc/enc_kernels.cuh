// Launch wrappers of the encoder's CUDA-core kernels (enc_kernels.cu) and of the tcgen05
// GEMM (gemm_tc.cu).  Activations are channels-last: [B][T][C], C contiguous.
#pragma once
#include "common.cuh"

namespace asrb {

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SILU = 3, ACT_GELU_GELU = 4 };
// storage types: fp32, bf16 (hidden states as they leave the encoder), op16 (the tensor-core operand format, common.cuh)
enum DType { DT_F32 = ASRB_F32, DT_BF16 = ASRB_BF16, DT_OP16 = 2, DT_SAME = -1 };

// ---- CUDA-core kernels (fp32 math; storage type per argument) ----------------------------

// [B][C][T] fp32 (reference layout, model.py:150-155) -> [B][T][CP] channels-last, channels
// >= C zero-filled.  floor_keys != NULL applies the log-mel dynamic-range floor on the fly
// (fused pcm->hidden path; essentials.py:489) and, if fix_src, also fixes src in place.
int launch_to_channels_last(const float* src, void* dst, DType dt, int64_t B, int C, int CP, int64_t T,
                            const uint32_t* floor_keys, const int32_t* lengths, int64_t n_samples, int hop,
                            bool fix_src, cudaStream_t st);

// out[b,t,n] = act( sum_{tap,k} A[b, t+tap-taps/2, k] W[n][tap][k] + bias[n] (+ res[b,t,n]) )
// A zero outside [0,T) (Conv1d padding).  FFMA, fp32 accumulate.
int launch_gemm_simt(const void* A, DType a_dt, const float* W, const float* bias, const void* res,
                     void* out, DType o_dt, int64_t B, int64_t T, int K, int N, int taps, Act act,
                     cudaStream_t st);

// out[r,:] = LayerNorm(x[r,:] (+ res[r,:])) * gamma + beta   (biased variance, eps)
int launch_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* out,
                     DType dt, int64_t rows, int D, float eps, cudaStream_t st, DType o_dt = DT_SAME);

// GLU over the channel dim: x [rows][2D] -> out [rows][D] = x[:, :D] * sigmoid(x[:, D:])
int launch_glu(const void* x, void* out, DType dt, int64_t rows, int D, cudaStream_t st);

// Depthwise conv along T, channels-last: out[b,t,c] = act(sum_j w[j][c] x[b,t+j-KW/2,c] + bias[c]);
// pos != NULL ([T][D] table from launch_pos_table) adds sinusoids(t, c) after the activation;
// fast = approximate erf/exp (tensor-core path); out may be a different storage type than x.  D % 64 == 0.
int launch_dwconv(const void* x, DType x_dt, const float* w, const float* bias, void* out, DType o_dt,
                  int64_t B, int64_t T, int D, int KW, Act act, const float* pos, bool fast, cudaStream_t st,
                  float* out32 = nullptr);     // optional fp32 copy of the output
int launch_pos_table(float* pos, const float* scales, int64_t T, int D, cudaStream_t st);

// Softmax attention, no mask: qkv [B][T][3D] (q | k | v, heads contiguous inside each) ->
// out [B][T][D].  fp32 math on CUDA cores (the fp32 variant and small shapes).
int launch_attention_simt(const void* qkv, void* out, DType dt, int64_t B, int64_t T, int D, int H,
                          float scale, cudaStream_t st);
// General form: separate q/k/v row strides (elements) so rotary / rms-normed copies can be used.
int launch_attention_simt_ex(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                             void* out, DType dt, int64_t B, int64_t T, int D, int H, float scale,
                             cudaStream_t st);

// Flash-style attention on tcgen05/TMEM (attn_tc.cu): op16 qkv [B][T][3D] -> op16 out [B][T][D]; head_dim 64 | 128
bool attention_tc_supported(int D, int H);
int launch_attention_tc(const void* qkv, void* out, int64_t B, int64_t T, int D, int H, float scale, cudaStream_t st);
// general form: separate query and key/value tensors (cross-attention against a cached K|V), Tq != Tk allowed
int launch_attention_tc_ex(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, void* out,
                           int64_t B, int64_t Tq, int64_t Tk, int D, int H, float scale, cudaStream_t st);

// RMSNorm rows (nn.RMSNorm, eps = fp32 machine eps): out = x / sqrt(mean(x^2)+eps) * w; out fp32 or op16 (an MMA operand)
int launch_rmsnorm(const float* x, const float* w, void* out, DType o_dt, int64_t rows, int D, cudaStream_t st,
                   const float* res = nullptr);            // optional: out = res + norm(x)
// tgate (model.py:525-535) on the gate / selector pre-activations of one GEMM: see enc_kernels.cu
int launch_tgate_combine(const void* g, int64_t ld, void* out, int64_t rows, int D, int n_types, cudaStream_t st);
// rotary (model.py:198-214) + per-head RMSNorm (model.py:307) in place on x [B*T][ld] (fp32 or op16), heads at h*hd
int launch_rotary_headnorm(void* x, DType dt, int64_t ld, const float* xa, const float* ln_w, const float* freqs,
                           int64_t B, int64_t T, int D, int H, float pre_scale, cudaStream_t st);

// ---- tcgen05 / TMEM / TMA GEMM (gemm_tc.cu), 16-bit (op16) operands, fp32 accumulate -----------------
enum TcEpilogue { TC_BIAS_ACT = 0, TC_GLU = 1, TC_RES_ACT = 2, TC_LN = 3,
                  TC_GLU_DW = 4,        // GLU -> depthwise conv down the frames (+folded BN) -> act2
                  TC_RES_ACT_DW = 5 };  // bias + residual + act -> depthwise conv -> act2 (+ sinusoids)

// The fp32 residual streams (res32 / out32) are workspace tensors in a blocked layout, [row / 32][column / 4][row % 32][4]
// (res32_index, gemm_tc_epi.cuh): allocate res32_rows(B * T) rows.
inline int64_t res32_rows(int64_t rows) { return (rows + 31) & ~(int64_t)31; }

struct TcGemmArgs {
    const op16* A;               // [B][T][K] channels-last
    const op16* W;               // [N][taps*K] tap-major (GLU: value/gate interleaved per tile)
    const float* bias;           // [N]
    const op16* res;             // [B][T][Nout] or NULL
    const float* res32;          // TC_LN: fp32 residual instead of res (post-norm residual streams stay fp32); BLOCKED layout, see res32_rows
    float* out32;                // TC_LN / TC_RES_ACT_DW: optional fp32 copy of the output, BLOCKED layout
    const float* gamma;          // TC_LN
    const float* beta;           // TC_LN
    void* out;                   // [B][T][Nout] op16 (bf16 when out_bf16: the encoder's result; fp32 when out_f32)
    int64_t B, T;
    int K, N, taps;
    int epilogue; int act; float eps;
    const float* dw_w;           // fused depthwise epilogues: [kw][Nout] taps, [Nout] bias, kw in {3, 15}
    const float* dw_b; int dw_kw; int dw_act;
    const float* pos;            // optional [T][Nout] sinusoid table added after dw_act
    int out_f32;                 // store fp32 (consumer is a depthwise conv, not an MMA); not with TC_LN
    int out_bf16;                // store bf16: this launch writes the hidden states the caller receives
    const int32_t* frames_eff;   // ragged batches: [B] device, frames per utterance still needed downstream (tiles past it are skipped), or NULL
};
bool tc_gemm_supported(int K, int N, int epilogue);
int  tc_glu_tile_n(int N);                    // BN the GLU weight interleave must use
int  launch_gemm_tc(const TcGemmArgs& a, cudaStream_t st);
// channel-major variants of TC_GLU_DW / TC_RES_ACT_DW (gemm_tct.cu): lanes = channels, columns = frames
bool tct_supported(const TcGemmArgs& a);
int  launch_gemm_tct(const TcGemmArgs& a, cudaStream_t st);
const op16* tct_identity();          // per-device 128 x 128 identity operand (created on first use)

}  // namespace asrb
