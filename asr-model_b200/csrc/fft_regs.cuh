// In-register small DFTs (forward, exp(-2 pi i nk/N)) for the log-mel kernel, on TWO independent
// transforms at once: every value is a float2 whose lanes belong to two different FFTs ("structure of
// arrays": a complex number is a pair of float2, real parts | imaginary parts), so each arithmetic
// instruction is one packed FADD2 / FMUL2 / FFMA2 of sm_100 -- half the issue slots of the scalar code, and
// multiplying by -i costs nothing (it only renames re / im in the butterfly that follows).
// Every loop is fully unrolled: array indices are compile-time, data stays in registers.
// 20 = 4 x 5 by the prime-factor (Good-Thomas) map, which needs no internal twiddles; 32 = 4 x 8
// Cooley-Tukey with immediate twiddles.
#pragma once
#include <cuda_runtime.h>

namespace asrb {

typedef float2 V2;
__device__ __forceinline__ V2 vadd(V2 a, V2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }          // negation folds into the operand
__device__ __forceinline__ V2 vmul(float c, V2 a) { return __fmul2_rn(make_float2(c, c), a); }              // scalar operand is broadcast by the instruction
__device__ __forceinline__ V2 vfma(float c, V2 a, V2 b) { return __ffma2_rn(make_float2(c, c), a, b); }     // c a + b
__device__ __forceinline__ V2 vmul2(V2 a, V2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ V2 vfma2(V2 a, V2 b, V2 c) { return __ffma2_rn(a, b, c); }                       // a b + c

struct cx2 { V2 re, im; };
__device__ __forceinline__ cx2 cadd(cx2 a, cx2 b) { return {vadd(a.re, b.re), vadd(a.im, b.im)}; }
__device__ __forceinline__ cx2 csub(cx2 a, cx2 b) { return {vsub(a.re, b.re), vsub(a.im, b.im)}; }
__device__ __forceinline__ cx2 cadd_mi(cx2 a, cx2 b) { return {vadd(a.re, b.im), vsub(a.im, b.re)}; }       // a + (-i) b
__device__ __forceinline__ cx2 csub_mi(cx2 a, cx2 b) { return {vsub(a.re, b.im), vadd(a.im, b.re)}; }       // a - (-i) b
// a * (c + i s), c and s scalars shared by both transforms
__device__ __forceinline__ cx2 cmul_cs(cx2 a, float c, float s) {
    return {vfma(c, a.re, vmul(-s, a.im)), vfma(s, a.re, vmul(c, a.im))};
}

__device__ __forceinline__ void dft4(cx2& x0, cx2& x1, cx2& x2, cx2& x3) {
    const cx2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    x0 = cadd(a, c); x2 = csub(a, c);
    x1 = cadd_mi(b, d); x3 = csub_mi(b, d);
}

__device__ __forceinline__ void dft5(cx2& x0, cx2& x1, cx2& x2, cx2& x3, cx2& x4) {
    const float c1 = 0.30901699437494742f;    // cos(2pi/5)
    const float c2 = -0.80901699437494742f;   // cos(4pi/5)
    const float s1 = 0.95105651629515357f;    // sin(2pi/5)
    const float s2 = 0.58778525229247313f;    // sin(4pi/5)
    const cx2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    const cx2 m1 = {vfma(c2, t2.re, vfma(c1, t1.re, x0.re)), vfma(c2, t2.im, vfma(c1, t1.im, x0.im))};
    const cx2 m2 = {vfma(c1, t2.re, vfma(c2, t1.re, x0.re)), vfma(c1, t2.im, vfma(c2, t1.im, x0.im))};
    const cx2 u1 = {vfma(s2, t4.re, vmul(s1, t3.re)), vfma(s2, t4.im, vmul(s1, t3.im))};       // s1 t3 + s2 t4
    const cx2 u2 = {vfma(-s1, t4.re, vmul(s2, t3.re)), vfma(-s1, t4.im, vmul(s2, t3.im))};     // s2 t3 - s1 t4
    x0 = cadd(x0, cadd(t1, t2));
    x1 = cadd_mi(m1, u1); x4 = csub_mi(m1, u1);
    x2 = cadd_mi(m2, u2); x3 = csub_mi(m2, u2);
}

__device__ __forceinline__ void dft8(cx2 (&v)[8]) {
    // even / odd radix-2 split over two 4-point DFTs
    cx2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    cx2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752f;
    const cx2 w1 = {vmul(h, vadd(o1.re, o1.im)), vmul(h, vsub(o1.im, o1.re))};      // o1 * (h, -h)
    const cx2 w3 = {vmul(h, vsub(o3.im, o3.re)), vmul(-h, vadd(o3.re, o3.im))};     // o3 * (-h, -h)
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
    v[2] = cadd_mi(e2, o2); v[6] = csub_mi(e2, o2);                                  // o2 * (-i)
    v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}

// W_32^m = (cos, -sin)(2 pi m / 32), m = n2*k1 <= 21; after unrolling each one is a pair of immediates.
__device__ __forceinline__ cx2 mul_w32(cx2 a, int m) {
    switch (m) {
        case 0:  return a;
        case 1:  return cmul_cs(a, 9.807852804e-01f, -1.950903220e-01f);
        case 2:  return cmul_cs(a, 9.238795325e-01f, -3.826834324e-01f);
        case 3:  return cmul_cs(a, 8.314696123e-01f, -5.555702330e-01f);
        case 4:  return cmul_cs(a, 7.071067812e-01f, -7.071067812e-01f);
        case 5:  return cmul_cs(a, 5.555702330e-01f, -8.314696123e-01f);
        case 6:  return cmul_cs(a, 3.826834324e-01f, -9.238795325e-01f);
        case 7:  return cmul_cs(a, 1.950903220e-01f, -9.807852804e-01f);
        case 8:  return {a.im, make_float2(-a.re.x, -a.re.y)};
        case 9:  return cmul_cs(a, -1.950903220e-01f, -9.807852804e-01f);
        case 10: return cmul_cs(a, -3.826834324e-01f, -9.238795325e-01f);
        case 11: return cmul_cs(a, -5.555702330e-01f, -8.314696123e-01f);
        case 12: return cmul_cs(a, -7.071067812e-01f, -7.071067812e-01f);
        case 13: return cmul_cs(a, -8.314696123e-01f, -5.555702330e-01f);
        case 14: return cmul_cs(a, -9.238795325e-01f, -3.826834324e-01f);
        case 15: return cmul_cs(a, -9.807852804e-01f, -1.950903220e-01f);
        case 16: return {make_float2(-a.re.x, -a.re.y), make_float2(-a.im.x, -a.im.y)};
        case 17: return cmul_cs(a, -9.807852804e-01f, 1.950903220e-01f);
        case 18: return cmul_cs(a, -9.238795325e-01f, 3.826834324e-01f);
        case 19: return cmul_cs(a, -8.314696123e-01f, 5.555702330e-01f);
        case 20: return cmul_cs(a, -7.071067812e-01f, 7.071067812e-01f);
        case 21: return cmul_cs(a, -5.555702330e-01f, 8.314696123e-01f);
    }
    return a;
}

template <int R> struct SmallDFT;

template <> struct SmallDFT<20> {
    // n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20  (CRT maps; no twiddles)
    __device__ __forceinline__ static void run(cx2 (&x)[20]) {
        cx2 a[4][5];
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) {
            cx2 v0 = x[(4 * n2) % 20], v1 = x[(5 + 4 * n2) % 20];
            cx2 v2 = x[(10 + 4 * n2) % 20], v3 = x[(15 + 4 * n2) % 20];
            dft4(v0, v1, v2, v3);
            a[0][n2] = v0; a[1][n2] = v1; a[2][n2] = v2; a[3][n2] = v3;
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            dft5(a[k1][0], a[k1][1], a[k1][2], a[k1][3], a[k1][4]);
#pragma unroll
            for (int k2 = 0; k2 < 5; ++k2) x[(5 * k1 + 16 * k2) % 20] = a[k1][k2];
        }
    }
};

template <> struct SmallDFT<32> {
    // n = 8 n1 + n2, k = k1 + 4 k2
    __device__ __forceinline__ static void run(cx2 (&x)[32]) {
        cx2 a[4][8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            cx2 v0 = x[n2], v1 = x[8 + n2], v2 = x[16 + n2], v3 = x[24 + n2];
            dft4(v0, v1, v2, v3);
            a[0][n2] = v0;
            a[1][n2] = mul_w32(v1, n2);
            a[2][n2] = mul_w32(v2, 2 * n2);
            a[3][n2] = mul_w32(v3, 3 * n2);
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            dft8(a[k1]);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) x[k1 + 4 * k2] = a[k1][k2];
        }
    }
};

}  // namespace asrb
