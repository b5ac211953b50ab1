// Shared host/device helpers for libasrb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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

// fast variants for the bf16 path (error far below one bf16 ulp)
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
// Total error incl. the MUFU bound: < 2.5e-4 |x| -- a few percent of a bf16 rounding of the result.
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

// order-preserving float <-> uint32 key (for atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
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

}  // namespace asrb
