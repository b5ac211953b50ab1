// CUDA-core kernels of the encoder: the fp32 (1e-4) variant's GEMM, and the streaming ops
// both variants share (layout change, LayerNorm, GLU, depthwise convs, attention, rotary).
// All math is fp32; storage is fp32 or the 16-bit operand format (op16, common.cuh) per tensor.  Channels-last everywhere.
#include "enc_kernels.cuh"

namespace asrb {

// ------------------------------------------------------------------------------------------
// [B][C][T] fp32 -> [B][T][CP] channels-last (+ optional log-mel floor, essentials.py:489)
// ------------------------------------------------------------------------------------------
template <class TO>
__global__ void to_channels_last_kernel(float* src, TO* dst, int C, int CP, int64_t T,
                                        const uint32_t* keys, const int32_t* lengths, int64_t n_samples,
                                        int hop, int fix_src) {
    extern __shared__ float tile[];                       // [C][33]
    const int b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float floor_s = -INFINITY;
    int64_t Tb = T;
    if (keys) {
        floor_s = ((key2f(keys[b]) - 8.0f) + 4.0f) / 4.0f;
        Tb = 1 + clamp_len(lengths, b, n_samples) / hop;
    }
    float* s = src + (int64_t)b * C * T;
    for (int c = warp; c < C; c += nwarp) {
        const int64_t t = t0 + lane;
        float v = 0.f;
        if (t < T) {
            v = s[(int64_t)c * T + t];
            if (t < Tb && v < floor_s) { v = floor_s; if (fix_src) s[(int64_t)c * T + t] = v; }
        }
        tile[c * 33 + lane] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * CP; i += blockDim.x) {
        const int tt = i / CP, c = i - tt * CP;
        const int64_t t = t0 + tt;
        if (t < T) io<TO>::st(dst + ((int64_t)b * T + t) * CP + c, c < C ? tile[c * 33 + tt] : 0.f);
    }
}

int launch_to_channels_last(const float* src, void* dst, DType dt, int64_t B, int C, int CP, int64_t T,
                            const uint32_t* keys, const int32_t* lengths, int64_t n_samples, int hop,
                            bool fix_src, cudaStream_t st) {
    ProfScope ps("to_channels_last", st, 0.0, (double)B * T * (4.0 * C + (dt == DT_F32 ? 4.0 : 2.0) * CP));
    dim3 grid((unsigned)((T + 31) / 32), (unsigned)B);
    size_t smem = sizeof(float) * C * 33;
    if (dt == DT_F32)
        to_channels_last_kernel<float><<<grid, 256, smem, st>>>((float*)src, (float*)dst, C, CP, T, keys, lengths, n_samples, hop, fix_src);
    else
        to_channels_last_kernel<op16><<<grid, 256, smem, st>>>((float*)src, (op16*)dst, C, CP, T, keys, lengths, n_samples, hop, fix_src);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 implicit-GEMM conv (taps along T) on CUDA cores: 64x64x16 tiles, 4x4 per thread
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_GELU: return gelu_erf(v);
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_SILU: return siluf_(v);
        case ACT_GELU_GELU: return gelu_erf(gelu_erf(v));
        default: return v;
    }
}

template <class TA, class TO>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                 const TO* __restrict__ res, TO* __restrict__ out, int64_t T, int K, int N, int taps, int act) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int b = blockIdx.z;
    const int64_t t0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int pad = taps / 2;
    float acc[4][4] = {};
    const TA* Ab = A + (int64_t)b * T * K;
    for (int tap = 0; tap < taps; ++tap) {
        for (int k0 = 0; k0 < K; k0 += BK) {
            for (int i = threadIdx.x; i < BM * BK; i += 256) {
                const int r = i / BK, kk = i - r * BK;
                const int64_t t = t0 + r + tap - pad;
                const int k = k0 + kk;
                As[kk][r] = (t >= 0 && t < T && k < K) ? io<TA>::ld(Ab + t * K + k) : 0.f;
            }
            for (int i = threadIdx.x; i < BN * BK; i += 256) {
                const int r = i / BK, kk = i - r * BK;
                const int n = n0 + r, k = k0 + kk;
                Ws[kk][r] = (n < N && k < K) ? W[((int64_t)n * taps + tap) * K + k] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t t = t0 + ty * 4 + i;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            const int64_t o = ((int64_t)b * T + t) * N + n;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (res) v += io<TO>::ld(res + o);
            io<TO>::st(out + o, apply_act(v, act));
        }
    }
}

int launch_gemm_simt(const void* A, DType a_dt, const float* W, const float* bias, const void* res,
                     void* out, DType o_dt, int64_t B, int64_t T, int K, int N, int taps, Act act,
                     cudaStream_t st) {

    dim3 grid((unsigned)((T + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)B);
    if (a_dt == DT_F32 && o_dt == DT_F32)
        gemm_simt_kernel<float, float><<<grid, 256, 0, st>>>((const float*)A, W, bias, (const float*)res, (float*)out, T, K, N, taps, act);
    else if (a_dt == DT_F32 && o_dt == DT_OP16)
        gemm_simt_kernel<float, op16><<<grid, 256, 0, st>>>((const float*)A, W, bias, (const op16*)res, (op16*)out, T, K, N, taps, act);
    else if (a_dt == DT_OP16 && o_dt == DT_OP16)
        gemm_simt_kernel<op16, op16><<<grid, 256, 0, st>>>((const op16*)A, W, bias, (const op16*)res, (op16*)out, T, K, N, taps, act);
    else if (a_dt == DT_OP16 && o_dt == DT_F32)
        gemm_simt_kernel<op16, float><<<grid, 256, 0, st>>>((const op16*)A, W, bias, (const float*)res, (float*)out, T, K, N, taps, act);
    else return fail(ASRB_E_ARG, "gemm_simt: storage types %d -> %d unsupported", (int)a_dt, (int)o_dt);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the channel dim (essentials.py:102-113 / nn.LayerNorm), one warp per row
// ------------------------------------------------------------------------------------------
template <class T, class TO>
__global__ void layernorm_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 TO* __restrict__ out, int64_t rows, int D, float eps) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const T* xr = x + row * D;
    const T* rr = res ? res + row * D : nullptr;
    float v[32];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = i * 32 + lane;
        v[i] = 0.f;
        if (c < D) { v[i] = io<T>::ld(xr + c) + (rr ? io<T>::ld(rr + c) : 0.f); s += v[i]; }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const int c = i * 32 + lane; if (c < D) { const float d = v[i] - mean; q = fmaf(d, d, q); } }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = i * 32 + lane;
        if (c < D) io<TO>::st(out + row * D + c, (v[i] - mean) * rstd * gamma[c] + beta[c]);
    }
}

int launch_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* out,
                     DType dt, int64_t rows, int D, float eps, cudaStream_t st, DType o_dt) {
    if (o_dt == DT_SAME) o_dt = dt;
    if (D > 1024) return fail(ASRB_E_ARG, "layernorm: D=%d > 1024", D);
    ProfScope ps("layernorm", st, 0.0, (double)rows * D * (dt == DT_F32 ? 4.0 : 2.0) * (res ? 3.0 : 2.0));
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (dt == DT_F32 && o_dt == DT_F32)
        layernorm_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (const float*)res, gamma, beta, (float*)out, rows, D, eps);
    else if (dt == DT_OP16 && o_dt == DT_OP16)
        layernorm_kernel<op16, op16><<<grid, 256, 0, st>>>((const op16*)x, (const op16*)res, gamma, beta, (op16*)out, rows, D, eps);
    else if (dt == DT_OP16 && o_dt == DT_BF16)
        layernorm_kernel<op16, __nv_bfloat16><<<grid, 256, 0, st>>>((const op16*)x, (const op16*)res, gamma, beta, (__nv_bfloat16*)out, rows, D, eps);
    else if (dt == DT_OP16 && o_dt == DT_F32)
        layernorm_kernel<op16, float><<<grid, 256, 0, st>>>((const op16*)x, (const op16*)res, gamma, beta, (float*)out, rows, D, eps);
    else return fail(ASRB_E_ARG, "layernorm: storage types %d -> %d unsupported", (int)dt, (int)o_dt);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// ------------------------------------------------------------------------------------------
// GLU (model.py:112)
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void glu_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows, int D) {
    const int64_t total = rows * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / D; const int c = (int)(i - r * D);
        const float a = io<T>::ld(x + r * 2 * D + c), g = io<T>::ld(x + r * 2 * D + D + c);
        io<T>::st(out + i, a * (1.0f / (1.0f + expf(-g))));
    }
}

int launch_glu(const void* x, void* out, DType dt, int64_t rows, int D, cudaStream_t st) {
    const int64_t total = rows * D;
    ProfScope ps("glu", st, 0.0, (double)rows * D * (dt == DT_F32 ? 4.0 : 2.0) * 3.0);
    unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (grid == 0) grid = 1;
    if (dt == DT_F32) glu_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, rows, D);
    else glu_kernel<op16><<<grid, 256, 0, st>>>((const op16*)x, (op16*)out, rows, D);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// ------------------------------------------------------------------------------------------
// Depthwise conv along T (model.py:99-101 k=15 with eval BatchNorm folded in; model.py:146
// k=3), channels-last, two channels per thread, TT outputs per thread from a register window.
// ------------------------------------------------------------------------------------------
template <class T> struct pair_io;
template <> struct pair_io<float> {
    __device__ static float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
    __device__ static void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct pair_io<op16> {
    __device__ static float2 ld(const op16* p) { return unpack_op16x2(*reinterpret_cast<const uint32_t*>(p)); }
    __device__ static void st(op16* p, float2 v) { *reinterpret_cast<uint32_t*>(p) = pack_op16x2(v.x, v.y); }
};

template <bool FAST> __device__ __forceinline__ float dw_act(float v, int act) {
    if (FAST) {
        switch (act) {
            case ACT_GELU: return gelu_fast(v);
            case ACT_SILU: return v * sigmoid_fast(v);
            case ACT_GELU_GELU: return gelu_fast(gelu_fast(v));
            case ACT_RELU: return fmaxf(v, 0.f);
            default: return v;
        }
    }
    return apply_act(v, act);
}

// Block = 128 time steps x 64 channels of one utterance: the input tile (+ halo) is staged in
// shared memory with 16-byte loads, then every thread slides a KW-wide register window down
// 16 time steps of one channel pair (1 LDS.64 per output pair), storing channel-contiguous rows.
template <class TI, class TO, int KW, bool FAST>
__global__ void __launch_bounds__(256)
dwconv_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              TO* __restrict__ out, int64_t T, int D, int act, const float* __restrict__ pos, float* __restrict__ out32) {
    constexpr int TB = 128, CB = 64, TL = 16, HALO = KW / 2, ROWS = TB + KW - 1;
    __shared__ __align__(16) float tile[ROWS][CB];
    const int b = blockIdx.z;
    const int64_t t0 = (int64_t)blockIdx.y * TB;
    const int c0 = blockIdx.x * CB;
    const TI* xb = x + (int64_t)b * T * D + c0;
    constexpr int VEC = 16 / sizeof(TI);                        // elements per 16-byte load
    for (int i = threadIdx.x; i < ROWS * (CB / VEC); i += 256) {
        const int rr = i / (CB / VEC), cv = (i - rr * (CB / VEC)) * VEC;
        const int64_t t = t0 - HALO + rr;
        float vals[VEC];
        if (t >= 0 && t < T) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(xb + t * D + cv));
            if (sizeof(TI) == 4) {
                const float* f = reinterpret_cast<const float*>(&q);
#pragma unroll
                for (int j = 0; j < VEC; ++j) vals[j] = f[j];
            } else {
                const uint32_t* h = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
                for (int j = 0; j < VEC / 2; ++j) { const float2 f = unpack_op16x2(h[j]); vals[2 * j] = f.x; vals[2 * j + 1] = f.y; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) vals[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < VEC; j += 4) *reinterpret_cast<float4*>(&tile[rr][cv + j]) = make_float4(vals[j], vals[j + 1], vals[j + 2], vals[j + 3]);
    }
    __syncthreads();
    const int cp = (threadIdx.x & 31) * 2;                      // channel pair inside the block
    const int ts = (threadIdx.x >> 5) * TL;                     // first time step of this warp's slice
    const int c = c0 + cp;
    float2 wv[KW];
#pragma unroll
    for (int j = 0; j < KW; ++j) wv[j] = __ldg(reinterpret_cast<const float2*>(w + (int64_t)j * D + c));
    const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + c));
    float2 win[KW];
#pragma unroll
    for (int j = 0; j < KW - 1; ++j) win[j + 1] = *reinterpret_cast<const float2*>(&tile[ts + j][cp]);
#pragma unroll
    for (int o = 0; o < TL; ++o) {
#pragma unroll
        for (int j = 0; j < KW - 1; ++j) win[j] = win[j + 1];
        win[KW - 1] = *reinterpret_cast<const float2*>(&tile[ts + o + KW - 1][cp]);
        const int64_t t = t0 + ts + o;
        if (t < T) {
            float2 a = bv;
#pragma unroll
            for (int j = 0; j < KW; ++j) { a.x = fmaf(wv[j].x, win[j].x, a.x); a.y = fmaf(wv[j].y, win[j].y, a.y); }
            a.x = dw_act<FAST>(a.x, act); a.y = dw_act<FAST>(a.y, act);
            if (pos) { const float2 pv = __ldg(reinterpret_cast<const float2*>(pos + t * D + c)); a.x += pv.x; a.y += pv.y; }
            pair_io<TO>::st(out + ((int64_t)b * T + t) * D + c, a);
            if (out32) *reinterpret_cast<float2*>(out32 + ((int64_t)b * T + t) * D + c) = a;
        }
    }
}

// sinusoids(T, D) (essentials.py:354-358) from the host-built scale table: [T][D] fp32
__global__ void pos_table_kernel(float* __restrict__ pos, const float* __restrict__ scales, int64_t T, int D) {
    const int half = D / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T * D; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / D; const int c = (int)(i - t * D);
        const float tf = (float)t;
        pos[i] = c < half ? sinf(tf * scales[c]) : cosf(tf * scales[c - half]);
    }
}
int launch_pos_table(float* pos, const float* scales, int64_t T, int D, cudaStream_t st) {
    ProfScope ps("pos_table", st, 0.0, 4.0 * T * D);
    pos_table_kernel<<<(unsigned)((T * D + 255) / 256 < 2048 ? (T * D + 255) / 256 : 2048), 256, 0, st>>>(pos, scales, T, D);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

template <class TI, class TO>
static int dwconv_dispatch(const void* x, const float* w, const float* bias, void* out, int64_t B, int64_t T,
                           int D, int KW, int act, const float* pos, bool fast, cudaStream_t st, float* out32) {
    dim3 grid((unsigned)(D / 64), (unsigned)((T + 127) / 128), (unsigned)B);
#define ASRB_DW(KW_, F_) dwconv_kernel<TI, TO, KW_, F_><<<grid, 256, 0, st>>>((const TI*)x, w, bias, (TO*)out, T, D, act, pos, out32)
    if (KW == 15) { if (fast) ASRB_DW(15, true); else ASRB_DW(15, false); }
    else if (KW == 3) { if (fast) ASRB_DW(3, true); else ASRB_DW(3, false); }
    else return fail(ASRB_E_ARG, "dwconv: kernel width %d unsupported", KW);
#undef ASRB_DW
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

int launch_dwconv(const void* x, DType x_dt, const float* w, const float* bias, void* out, DType o_dt,
                  int64_t B, int64_t T, int D, int KW, Act act, const float* pos, bool fast, cudaStream_t st, float* out32) {
    if (D % 64) return fail(ASRB_E_ARG, "dwconv: D=%d must be a multiple of 64", D);
    ProfScope ps(KW == 15 ? "dwconv15_bn_silu" : (pos ? "dwconv3_gelu_pos" : "dwconv3_gelu"), st, 2.0 * B * T * (double)D * KW,
                 (double)B * T * D * ((x_dt == DT_F32 ? 4.0 : 2.0) + (o_dt == DT_F32 ? 4.0 : 2.0)));
    if (x_dt == DT_F32 && o_dt == DT_F32) return dwconv_dispatch<float, float>(x, w, bias, out, B, T, D, KW, act, pos, fast, st, out32);
    if (x_dt == DT_OP16 && o_dt == DT_OP16) return dwconv_dispatch<op16, op16>(x, w, bias, out, B, T, D, KW, act, pos, fast, st, out32);
    if (x_dt == DT_OP16 && o_dt == DT_F32) return dwconv_dispatch<op16, float>(x, w, bias, out, B, T, D, KW, act, pos, fast, st, out32);
    if (x_dt == DT_F32 && o_dt == DT_OP16) return dwconv_dispatch<float, op16>(x, w, bias, out, B, T, D, KW, act, pos, fast, st, out32);
    return fail(ASRB_E_ARG, "dwconv: storage types %d -> %d unsupported", (int)x_dt, (int)o_dt);
}

// ------------------------------------------------------------------------------------------
// Attention on CUDA cores: 32 queries x 4 dim-slices per block, 32-key tiles in shared memory,
// online softmax in sub-batches of 8 keys.  No mask (model.py:163 passes none).
// ------------------------------------------------------------------------------------------
template <class T, int HD>
__global__ void __launch_bounds__(128)
attention_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                      int64_t ldq, int64_t ldk, int64_t ldv, T* __restrict__ out, int64_t Tn, int D, float scale) {
    constexpr int DP = HD / 4, KT = 32;
    __shared__ float Ks[KT][HD + 1];
    __shared__ float Vs[KT][HD + 1];
    const int b = blockIdx.z, h = blockIdx.y;
    const int ql = threadIdx.x >> 2, part = threadIdx.x & 3;
    const int64_t tq = (int64_t)blockIdx.x * 32 + ql;
    const bool valid = tq < Tn;
    float qv[DP], o[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        qv[d] = valid ? io<T>::ld(q + ((int64_t)b * Tn + tq) * ldq + h * HD + part * DP + d) * scale : 0.f;
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int64_t k0 = 0; k0 < Tn; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HD; i += 128) {
            const int r = i / HD, d = i - r * HD;
            const int64_t t = k0 + r;
            Ks[r][d] = t < Tn ? io<T>::ld(k + ((int64_t)b * Tn + t) * ldk + h * HD + d) : 0.f;
            Vs[r][d] = t < Tn ? io<T>::ld(v + ((int64_t)b * Tn + t) * ldv + h * HD + d) : 0.f;
        }
        __syncthreads();
        const int nk = (int)((Tn - k0) < KT ? (Tn - k0) : KT);
        for (int j0 = 0; j0 < nk; j0 += 8) {
            float s[8];
            float mx = m;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d0 = 0.f;
#pragma unroll
                for (int d = 0; d < DP; ++d) d0 = fmaf(qv[d], Ks[j0 + j][part * DP + d], d0);
                d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
                d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
                s[j] = (j0 + j < nk) ? d0 : -INFINITY;
                mx = fmaxf(mx, s[j]);
            }
            const float alpha = (m == -INFINITY) ? 0.f : expf(m - mx);
            l *= alpha;
#pragma unroll
            for (int d = 0; d < DP; ++d) o[d] *= alpha;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = (s[j] == -INFINITY) ? 0.f : expf(s[j] - mx);
                l += p;
#pragma unroll
                for (int d = 0; d < DP; ++d) o[d] = fmaf(p, Vs[j0 + j][part * DP + d], o[d]);
            }
            m = mx;
        }
    }
    if (valid) {
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < DP; ++d) io<T>::st(out + ((int64_t)b * Tn + tq) * D + h * HD + part * DP + d, o[d] * inv);
    }
}

template <class T>
static int attention_dispatch(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                              void* out, int64_t B, int64_t Tn, int D, int H, float scale, cudaStream_t st) {
    const int hd = D / H;
    dim3 grid((unsigned)((Tn + 31) / 32), (unsigned)H, (unsigned)B);
#define ASRB_ATT(HD_) attention_simt_kernel<T, HD_><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, (T*)out, Tn, D, scale)
    switch (hd) {
        case 16: ASRB_ATT(16); break;
        case 32: ASRB_ATT(32); break;
        case 64: ASRB_ATT(64); break;
        case 128: ASRB_ATT(128); break;
        default: return fail(ASRB_E_ARG, "attention: head_dim %d unsupported (16/32/64/128)", hd);
    }
#undef ASRB_ATT
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

int launch_attention_simt_ex(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                             void* out, DType dt, int64_t B, int64_t T, int D, int H, float scale, cudaStream_t st) {
    ProfScope ps("attention_simt", st, 4.0 * B * (double)T * T * D, (double)B * T * D * 4.0 * (dt == DT_F32 ? 4.0 : 2.0));
    if (dt == DT_F32) return attention_dispatch<float>(q, k, v, ldq, ldk, ldv, out, B, T, D, H, scale, st);
    return attention_dispatch<op16>(q, k, v, ldq, ldk, ldv, out, B, T, D, H, scale, st);
}

int launch_attention_simt(const void* qkv, void* out, DType dt, int64_t B, int64_t T, int D, int H, float scale,
                          cudaStream_t st) {
    const size_t es = dt == DT_F32 ? 4 : 2;
    const char* p = (const char*)qkv;
    return launch_attention_simt_ex(p, p + es * D, p + es * 2 * D, 3 * D, 3 * D, 3 * D, out, dt, B, T, D, H, scale, st);
}

// ------------------------------------------------------------------------------------------
// Secondary block: RMSNorm rows; rotary (model.py:198-214) + per-head RMSNorm (model.py:307)
// ------------------------------------------------------------------------------------------
template <class TO>
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, TO* __restrict__ out,
                               int64_t rows, int D, float eps, const float* __restrict__ res) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = x[row * D + c]; s = fmaf(v, v, s); }
    const float r = rsqrtf(warp_sum(s) / (float)D + eps);
    for (int c = lane; c < D; c += 32)
        io<TO>::st(out + row * D + c, x[row * D + c] * r * w[c] + (res ? res[row * D + c] : 0.f));
}

int launch_rmsnorm(const float* x, const float* w, void* out, DType o_dt, int64_t rows, int D, cudaStream_t st, const float* res) {
    ProfScope ps("rmsnorm", st, 0.0, (double)rows * D * (o_dt == DT_F32 ? 8.0 : 6.0));
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (o_dt == DT_F32) rmsnorm_kernel<float><<<grid, 256, 0, st>>>(x, w, (float*)out, rows, D, 1.1920928955078125e-07f, res);
    else rmsnorm_kernel<op16><<<grid, 256, 0, st>>>(x, w, (op16*)out, rows, D, 1.1920928955078125e-07f, res);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// tgate (model.py:525-535): g [rows][ld] holds, per row, n_types blocks of D gate pre-activations followed by n_types
// selector logits; out[r][c] = sum_i softmax(logits)_i * sigmoid(g[r][i D + c]).  One warp per row.
__global__ void tgate_combine_kernel(const op16* __restrict__ g, int64_t ld, op16* __restrict__ out, int64_t rows, int D, int n_types) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const op16* gr = g + row * ld;
    float t[8], mx = -INFINITY, sum = 0.f;
    for (int i = 0; i < n_types; ++i) { t[i] = io<op16>::ld(gr + (int64_t)n_types * D + i); mx = fmaxf(mx, t[i]); }
    for (int i = 0; i < n_types; ++i) { t[i] = __expf(t[i] - mx); sum += t[i]; }
    const float inv = 1.0f / sum;
    for (int c = lane; c < D; c += 32) {
        float acc = 0.f;
        for (int i = 0; i < n_types; ++i) acc = fmaf(t[i] * inv, sigmoid_fast(io<op16>::ld(gr + (int64_t)i * D + c)), acc);
        io<op16>::st(out + row * D + c, acc);
    }
}
int launch_tgate_combine(const void* g, int64_t ld, void* out, int64_t rows, int D, int n_types, cudaStream_t st) {
    if (n_types < 1 || n_types > 8) return fail(ASRB_E_ARG, "tgate: %d types unsupported (1..8)", n_types);
    ProfScope ps("tgate_combine", st, 0.0, 2.0 * rows * ((double)n_types * D + D));
    tgate_combine_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>((const op16*)g, ld, (op16*)out, rows, D, n_types);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// One warp per (row = (b,t), head).  x <- rmsnorm_hd( (x * pre_scale) (*) polar(||xa_t||, t f_j) ) * ln_w
template <class TX>
__global__ void rotary_headnorm_kernel(TX* __restrict__ x_all, int64_t ld, const float* __restrict__ xa,
                                       const float* __restrict__ ln_w, const float* __restrict__ freqs,
                                       int64_t rows, int64_t T, int D, int H, float pre_scale, float eps) {
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int hd = D / H;
    const int64_t row = wid / H;
    if (row >= rows) return;
    const int h = (int)(wid - row * H);
    const int64_t t = row % T;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = xa[row * D + c]; s = fmaf(v, v, s); }
    const float mag = sqrtf(warp_sum(s));                       // torch.norm(xa, dim=-1) (model.py:201)
    TX* x = x_all + row * ld + h * hd;
    float yv[4];                                                // hd <= 128: up to two pairs per lane, kept in fp32
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int j = lane + 32 * i;
        yv[2 * i] = yv[2 * i + 1] = 0.f;
        if (j < hd / 2) {
            const float ang = (float)t * freqs[j];
            float sn, cs;
            sincosf(ang, &sn, &cs);
            const float fr = mag * cs, fi = mag * sn;           // torch.polar(m, f)
            const float xr = io<TX>::ld(x + 2 * j) * pre_scale, xi = io<TX>::ld(x + 2 * j + 1) * pre_scale;
            yv[2 * i] = xr * fr - xi * fi; yv[2 * i + 1] = xr * fi + xi * fr;
            ss = fmaf(yv[2 * i], yv[2 * i], fmaf(yv[2 * i + 1], yv[2 * i + 1], ss));
        }
    }
    const float r = rsqrtf(warp_sum(ss) / (float)hd + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int j = lane + 32 * i;
        if (j < hd / 2) {
            io<TX>::st(x + 2 * j, yv[2 * i] * r * ln_w[2 * j]);
            io<TX>::st(x + 2 * j + 1, yv[2 * i + 1] * r * ln_w[2 * j + 1]);
        }
    }
}

int launch_rotary_headnorm(void* x, DType dt, int64_t ld, const float* xa, const float* ln_w, const float* freqs,
                           int64_t B, int64_t T, int D, int H, float pre_scale, cudaStream_t st) {
    const int64_t warps = B * T * H;
    if (D / H > 128) return fail(ASRB_E_ARG, "rotary: head_dim %d > 128", D / H);
    ProfScope ps("rotary_headnorm", st, 0.0, (double)B * T * D * (dt == DT_F32 ? 12.0 : 8.0));
    const unsigned grid = (unsigned)((warps + 7) / 8);
    if (dt == DT_F32)
        rotary_headnorm_kernel<float><<<grid, 256, 0, st>>>((float*)x, ld, xa, ln_w, freqs, B * T, T, D, H, pre_scale, 1.1920928955078125e-07f);
    else
        rotary_headnorm_kernel<op16><<<grid, 256, 0, st>>>((op16*)x, ld, xa, ln_w, freqs, B * T, T, D, H, pre_scale, 1.1920928955078125e-07f);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

}  // namespace asrb
