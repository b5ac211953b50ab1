// Device-side pieces shared by the tcgen05 GEMM kernels (gemm_tc.cu: lanes = frames; gemm_tct.cu: lanes =
// channels, with the depthwise convolutions fused).
#pragma once
#include "tc_common.cuh"

namespace asrb {

// Per epilogue group (4 warps = 128 threads = one row each): the 16 KB staging tile of the TMA store
// ([128][64] 16-bit swizzled / [128][32] fp32 swizzled).
static constexpr int GROUP_SCRATCH = 16 * 1024;

template <int BN> struct TcCfg {
    static constexpr int NG = BN == 256 ? 4 : 2;                       // epilogue groups (column slices)
    static constexpr int THREADS = 64 + NG * 128;
    static constexpr int STAGES = BN == 256 ? 3 : 5;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int XCH_BYTES = NG * BM * 8 + 2 * BM * 8;         // LN partial sums per row and group + 2 peer slots (cluster LN)
    static constexpr int SMEM = STAGES * STAGE_BYTES + NG * GROUP_SCRATCH + XCH_BYTES + 256 /*barriers + tmem slot*/;
    static_assert((2 * STAGES + 10) * 8 + 4 <= 256, "barrier region");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// --------------------------------- fast epilogue math -----------------------------------
__device__ __forceinline__ float act_fast(float v, int act) {
    switch (act) {
        case ACT_GELU: return gelu_fast(v);
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_SILU: return v * sigmoid_fast(v);
        case ACT_GELU_GELU: return gelu_fast(gelu_fast(v));
        default: return v;
    }
}
// gelu_fast on a pair of values with packed f32x2 arithmetic (same roundings; half the issue slots for the multiplies and FMAs)
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
    float2 s = __fmul2_rn(x, x);
    s.x = fminf(s.x, 49.0f); s.y = fminf(s.y, 49.0f);
    float2 q = __ffma2_rn(make_float2(-3.58867440e-04f, -3.58867440e-04f), s, make_float2(3.70510348e-02f, 3.70510348e-02f));
    q = __ffma2_rn(q, s, make_float2(7.97457818e-01f, 7.97457818e-01f));
    const float2 hx = __fmul2_rn(make_float2(0.5f, 0.5f), x);
    float2 t = __fmul2_rn(x, q);
    t.x = tanh_approx(t.x); t.y = tanh_approx(t.y);
    return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ void act_fast32(float (&v)[32], int act) {      // switch hoisted out of the element loop
    switch (act) {
        case ACT_GELU:
#pragma unroll
            for (int i = 0; i < 32; i += 2) { const float2 y = gelu_fast2(make_float2(v[i], v[i + 1])); v[i] = y.x; v[i + 1] = y.y; }
            break;
        case ACT_GELU_GELU:
#pragma unroll
            for (int i = 0; i < 32; i += 2) { const float2 y = gelu_fast2(gelu_fast2(make_float2(v[i], v[i + 1]))); v[i] = y.x; v[i + 1] = y.y; }
            break;
        case ACT_SILU:
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] * sigmoid_fast(v[i]);
            break;
        case ACT_RELU:
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            break;
        default: break;
    }
}

struct TcParams {
    const float* bias; const op16* res; const float* res32; float* out32; const float* gamma; const float* beta;
    int T, K, N, taps, act, tiles_per_utt, m_tiles, n_chunks, n_out, out_f32, out_bf16;
    int cluster;                 // TC_LN with N = 512: 2 CTAs own one 256-column half each and swap row sums over DSMEM
    float eps;
    int halo, rows_out;          // rows_out = 128 frames per tile, halo = 0 (kept for the tile arithmetic)
    const int32_t* frames_eff;   // ragged batches: [B] frames per utterance that anything downstream still needs, or NULL
    int k3one;                   // 3 taps: the frame tile of a k-block is fetched ONCE (130 rows) and the taps read it at row offsets
};

// Ragged batches: a tile whose first produced frame lies at or past the utterance's effective extent
// is skipped by every warp role alike (the predicate only depends on the unit), so no pipeline state moves for it.
__device__ __forceinline__ bool tile_is_padding(const int32_t* frames_eff, int b, int first_frame) {
    return frames_eff != nullptr && first_frame >= __ldg(frames_eff + b);
}

// 32 fp32 values of one row -> 32 16-bit values into the swizzled staging tile (row r, columns cb..cb+31 of 64);
// as_bf16 (warp-uniform): this tile is the encoder's result, otherwise the next MMA's operand
__device__ __forceinline__ void stage_store32(unsigned char* staging, int r, int cb, const float (&v)[32], bool as_bf16 = false) {
    if (as_bf16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = (cb >> 3) + i;                          // 16-byte chunk index 0..7 in the 128-B row
            uint4 q;
            q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(staging + r * 128 + ((chunk ^ (r & 7)) << 4)) = q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = (cb >> 3) + i;
            uint4 q;
            q.x = pack_op16x2(v[8 * i + 0], v[8 * i + 1]); q.y = pack_op16x2(v[8 * i + 2], v[8 * i + 3]);
            q.z = pack_op16x2(v[8 * i + 4], v[8 * i + 5]); q.w = pack_op16x2(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(staging + r * 128 + ((chunk ^ (r & 7)) << 4)) = q;
        }
    }
}
// 32 fp32 values of one row -> one 128-byte swizzled staging row (fp32 output tiles are 32 columns wide)
__device__ __forceinline__ void stage_store32_f32(unsigned char* staging, int r, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(staging + r * 128 + ((i ^ (r & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void add_res32(float (&v)[32], const op16* rp) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 q = *reinterpret_cast<const uint4*>(rp + 8 * i);
        const uint32_t* h = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_op16x2(h[j]); v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y; }
    }
}
// ---- fp32 residual streams between kernels (workspace tensors, never user-visible) are stored in blocks of 32 rows,
// [row / 32][column / 4][row % 32][4]: the LayerNorm epilogues own one ROW per thread, and with rows N floats apart a
// warp's 16-byte access would touch 32 cache lines (the out-projection + LayerNorm launch of the TransformerEncoderLayer
// spent 0.5 ms on L1 tag cycles for 0.1 TFLOP of MMAs); in this layout the 32 rows of a column group are 512 contiguous
// bytes.  The channel-major producer (gemm_tct.cu, one CHANNEL per thread) writes 16-byte pieces, 8 sectors per warp store.
// `rows` of such a tensor round up to a multiple of 32 (res32_rows, enc_kernels.cuh). ----
__device__ __forceinline__ int64_t res32_index(int64_t row, int col, int N) {
    return ((row >> 5) * (N >> 2) + (col >> 2)) * 128 + ((row & 31) << 2) + (col & 3);
}
__device__ __forceinline__ void add_f32x32(float (&v)[32], const float* base, int64_t row, int col, int N) {     // plain (not read-only) loads; col % 4 == 0
    const float4* q = reinterpret_cast<const float4*>(base + res32_index(row, col, N));
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 f = q[32 * i]; v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w; }
}
__device__ __forceinline__ void store_f32x32(float* base, int64_t row, int col, int N, const float (&v)[32]) {
    float4* q = reinterpret_cast<float4*>(base + res32_index(row, col, N));
#pragma unroll
    for (int i = 0; i < 8; ++i) q[32 * i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void add_vec32(float (&v)[32], const float* p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 f = __ldg(reinterpret_cast<const float4*>(p) + i); v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w; }
}

}  // namespace asrb
