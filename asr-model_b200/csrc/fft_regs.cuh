// In-register small DFTs (forward, exp(-2 pi i nk/N)) used by the log-mel kernel.
// Every loop is fully unrolled so array indices are compile-time and the data stays in
// registers.  20 = 4 x 5 by the prime-factor (Good-Thomas) map, which needs no internal
// twiddles; 32 = 4 x 8 Cooley-Tukey with immediate twiddles.
#pragma once
#include <cuda_runtime.h>

namespace asrb {

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__host__ __device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // -i a
__host__ __device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }   // +i a

__host__ __device__ __forceinline__ void dft2(float2& a, float2& b) {
    float2 t = csub(a, b); a = cadd(a, b); b = t;
}

__host__ __device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    x0 = cadd(a, c); x2 = csub(a, c);
    float2 md = mul_mi(d);
    x1 = cadd(b, md); x3 = csub(b, md);
}

__host__ __device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
    const float c1 = 0.30901699437494742f;    // cos(2pi/5)
    const float c2 = -0.80901699437494742f;   // cos(4pi/5)
    const float s1 = 0.95105651629515357f;    // sin(2pi/5)
    const float s2 = 0.58778525229247313f;    // sin(4pi/5)
    float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    float2 m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, x0.x)), fmaf(c2, t2.y, fmaf(c1, t1.y, x0.y)));
    float2 m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, x0.x)), fmaf(c1, t2.y, fmaf(c2, t1.y, x0.y)));
    float2 u1 = make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y));     // s1 t3 + s2 t4
    float2 u2 = make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y));   // s2 t3 - s1 t4
    x0 = cadd(x0, cadd(t1, t2));
    float2 iu1 = mul_mi(u1), iu2 = mul_mi(u2);
    x1 = cadd(m1, iu1); x4 = csub(m1, iu1);
    x2 = cadd(m2, iu2); x3 = csub(m2, iu2);
}

__host__ __device__ __forceinline__ void dft8(float2 (&v)[8]) {
    // even / odd radix-2 split over two 4-point DFTs
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752f;
    float2 w1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // o1 * (h, -h)
    float2 w2 = mul_mi(o2);                                              // o2 * (-i)
    float2 w3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // o3 * (-h, -h)
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
    v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
    v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}

// W_32^m, m = n2*k1 <= 21; a switch so that after unrolling each twiddle is an immediate.
__host__ __device__ __forceinline__ float2 tw32(int m) {
    switch (m) {
        case 0:  return make_float2(1.f, 0.f);
        case 1:  return make_float2(9.807852804e-01f, -1.950903220e-01f);
        case 2:  return make_float2(9.238795325e-01f, -3.826834324e-01f);
        case 3:  return make_float2(8.314696123e-01f, -5.555702330e-01f);
        case 4:  return make_float2(7.071067812e-01f, -7.071067812e-01f);
        case 5:  return make_float2(5.555702330e-01f, -8.314696123e-01f);
        case 6:  return make_float2(3.826834324e-01f, -9.238795325e-01f);
        case 7:  return make_float2(1.950903220e-01f, -9.807852804e-01f);
        case 8:  return make_float2(0.f, -1.f);
        case 9:  return make_float2(-1.950903220e-01f, -9.807852804e-01f);
        case 10: return make_float2(-3.826834324e-01f, -9.238795325e-01f);
        case 11: return make_float2(-5.555702330e-01f, -8.314696123e-01f);
        case 12: return make_float2(-7.071067812e-01f, -7.071067812e-01f);
        case 13: return make_float2(-8.314696123e-01f, -5.555702330e-01f);
        case 14: return make_float2(-9.238795325e-01f, -3.826834324e-01f);
        case 15: return make_float2(-9.807852804e-01f, -1.950903220e-01f);
        case 16: return make_float2(-1.f, 0.f);
        case 17: return make_float2(-9.807852804e-01f, 1.950903220e-01f);
        case 18: return make_float2(-9.238795325e-01f, 3.826834324e-01f);
        case 19: return make_float2(-8.314696123e-01f, 5.555702330e-01f);
        case 20: return make_float2(-7.071067812e-01f, 7.071067812e-01f);
        case 21: return make_float2(-5.555702330e-01f, 8.314696123e-01f);
    }
    return make_float2(1.f, 0.f);
}

template <int R> struct SmallDFT;

template <> struct SmallDFT<20> {
    // n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20  (CRT maps; no twiddles)
    __host__ __device__ __forceinline__ static void run(float2 (&x)[20]) {
        float2 a[4][5];
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) {
            float2 v0 = x[(4 * n2) % 20], v1 = x[(5 + 4 * n2) % 20];
            float2 v2 = x[(10 + 4 * n2) % 20], v3 = x[(15 + 4 * n2) % 20];
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
    __host__ __device__ __forceinline__ static void run(float2 (&x)[32]) {
        float2 a[4][8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            float2 v0 = x[n2], v1 = x[8 + n2], v2 = x[16 + n2], v3 = x[24 + n2];
            dft4(v0, v1, v2, v3);
            a[0][n2] = v0;
            a[1][n2] = n2 ? cmul(v1, tw32(n2)) : v1;
            a[2][n2] = n2 ? cmul(v2, tw32(2 * n2)) : v2;
            a[3][n2] = n2 ? cmul(v3, tw32(3 * n2)) : v3;
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
