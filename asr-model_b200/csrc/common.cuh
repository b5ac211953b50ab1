// Shared host/device helpers for libasrb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include "../../include/asrb200.h"

namespace asrb {

// ---- thread-local error string behind asrb_last_error() ----
inline std::string& err_slot() { static thread_local std::string s; return s; }
inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err_slot() = buf;
    return code;
}
#define ASRB_CUDA(expr)                                                                   \
    do { cudaError_t e__ = (expr);                                                        \
         if (e__ != cudaSuccess)                                                          \
             return ::asrb::fail(ASRB_E_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                                 cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)
#define ASRB_LAUNCH_CHECK()  ASRB_CUDA(cudaGetLastError())
#define ASRB_TRY(expr) do { int r__ = (expr); if (r__ != ASRB_OK) return r__; } while (0)

int require_sm100();                       // api.cu: current device must be cc 10.x
int sm_count();                            // api.cu: SM count of the current device (cached)

// ---- optional per-launch profiler (asrb_profile_*): CUDA events around every kernel launch
// on the launching stream, with the launch's algorithmic flops / bytes.  Off by default. ----
struct ProfRec { const char* tag; cudaEvent_t e0, e1; double flops, bytes; };
bool prof_on();
void prof_push(const char* tag, cudaStream_t st, double flops, double bytes);   // records e0
void prof_pop(cudaStream_t st);                                                 // records e1
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(const char* tag, cudaStream_t s, double flops, double bytes) : st(s), on(prof_on()) {
        if (on) prof_push(tag, st, flops, bytes);
    }
    ~ProfScope() { if (on) prof_pop(st); }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
    char* base; size_t cap; size_t off;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <class T> T* take(size_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ---- device math ----
__device__ __forceinline__ float gelu_erf(float x) {          // nn.GELU() exact (essentials.py:224)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }

// fast variants for the tensor-core path (error far below one rounding of the 16-bit result)
__device__ __forceinline__ float rcp_approx(float x) {           // MUFU.RCP, 1 ulp, no slow path
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_approx(float x) {          // MUFU.TANH, max relative error 2^-11
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// erf-GELU through one MUFU: 0.5 x (1 + tanh(x q(x^2))) with q fitted so that tanh(x q) = erf(x / sqrt 2)
// (|formula error| < 3e-5 on the whole line; q is evaluated at min(x^2, 49): tanh has long saturated there).
// Total error incl. the MUFU bound: < 2.5e-4 |x| -- far below the 2e-2 tolerance and of the order of an fp16 rounding.
__device__ __forceinline__ float gelu_fast(float x) {
    const float s = fminf(x * x, 49.0f);               // beyond |x| = 7 tanh has long saturated; keeps x q(s) monotone
    const float q = fmaf(fmaf(-3.58867440e-04f, s, 3.70510348e-02f), s, 7.97457818e-01f);
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(x * q), hx);
}
// sigmoid(x) = 0.5 + 0.5 tanh(x / 2)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// valid samples of utterance b: lengths[b] clamped to [0, n_samples] (a bad length must not read or write out of bounds)
__device__ __forceinline__ int64_t clamp_len(const int32_t* lengths, int b, int64_t n_samples) {
    if (!lengths) return n_samples;
    const int64_t l = (int64_t)lengths[b];
    return l < 0 ? 0 : (l > n_samples ? n_samples : l);
}

// order-preserving float <-> uint32 key (for atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- the 16-bit tensor-core operand format ------------------------------------------------
// Every MMA operand that lives in HBM or shared memory (activations X / Y / U / QKV / P / FFN hidden and the packed
// weights) is IEEE fp16, accumulated in fp32 by tcgen05.mma kind::f16; hidden states leave the encoder as bf16.
// Why not bf16 operands: with 8 mantissa bits per operand the encoder cannot meet allclose(atol 2e-2, rtol 1e-2)
// against the fp32 reference -- not even with the whole last layer computed exactly (tools/numerics_study.py:
// worst error / tolerance 1.5-2.7 for bf16 operands, 0.23-0.43 for fp16 at the same MMA rate and the same bytes).
// Range: every operand is either a LayerNorm / GELU / SiLU / softmax output or a weight; conversions saturate
// (cvt.rn.satfinite) instead of producing inf.  -DASRB_OPERAND_BF16 rebuilds the library with bf16 operands for A/B runs.
#ifdef ASRB_OPERAND_BF16
typedef __nv_bfloat16 op16;
#define ASRB_OP16_IS_F16 0
#else
typedef __half op16;
#define ASRB_OP16_IS_F16 1
#endif
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {           // a -> low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t pack_op16x2(float a, float b) {           // two fp32 -> packed operand pair, a -> low half
#if ASRB_OP16_IS_F16
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
#else
    return pack_bf16x2(a, b);
#endif
}
__device__ __forceinline__ float2 unpack_op16x2(uint32_t u) {
#if ASRB_OP16_IS_F16
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
#else
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
#endif
}
__device__ __forceinline__ op16 to_op16(float v) {
#if ASRB_OP16_IS_F16
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
#else
    return __float2bfloat16_rn(v);
#endif
}
inline op16 host_to_op16(float v) {
#if ASRB_OP16_IS_F16
    const float lim = 65504.0f;
    return __float2half_rn(v > lim ? lim : (v < -lim ? -lim : v));
#else
    return __float2bfloat16_rn(v);
#endif
}

template <class T> struct io;
template <> struct io<float> {
    __device__ static float ld(const float* p) { return *p; }
    __device__ static void st(float* p, float v) { *p = v; }
};
template <> struct io<__nv_bfloat16> {
    __device__ static float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct io<__half> {
    __device__ static float ld(const __half* p) { return __half2float(*p); }
    __device__ static void st(__half* p, float v) {
        unsigned short r;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
        *p = __ushort_as_half(r);
    }
};

}  // namespace asrb
