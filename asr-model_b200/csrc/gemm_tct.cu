// Channel-major tcgen05 GEMMs with the depthwise convolutions of ConvLite fused into the epilogue
// (sm_100a).  Same contraction as gemm_tc.cu with the operand roles swapped:
//
//   D[c, t] = sum_k W[c][k] * X[b, t, k]          (A operand = weight tile, B operand = frame tile)
//
// so a TMEM lane is an output CHANNEL and a TMEM column is a FRAME.  An epilogue thread therefore owns
// one channel and reads consecutive frames straight into registers: the depthwise convolution along
// the frames (model.py:113 k=15, model.py:146 k=3) is a sliding window over registers with per-thread
// scalar taps -- no shared-memory transpose, no barriers, and results leave as 64-byte row segments
// (32 consecutive channels of one frame per warp store).  Two kernels:
//
//   TC_RES_ACT_DW  point2 (1x1) + bias + residual + GELU -> depthwise-3 -> GELU[∘GELU] (+ sinusoids)
//                  model.py:116-118,145-147,160-161.  256 frames per tile (1 halo frame each side).  The
//                  residual is added by the tensor core as two extra k-blocks  I[128x128] * Y^T  (exact:
//                  operand x 1.0 into the fp32 accumulator), so the epilogue issues no residual loads.
//   TC_GLU_DW      point1 (1x1, D -> 2D) + GLU -> depthwise-15 (eval BatchNorm folded) -> SiLU
//                  model.py:111-115.  128 frames per tile (112 produced), value and gate halves of the
//                  same 128 channels in two accumulators (two MMAs per k-step share the frame tile).
//
// Roles per CTA (1 CTA / SM, persistent over (utterance, frame tile, channel tile) units):
//   warp 0       TMA producer  (4-stage ring of 48 KB: weight box + frame box, SWIZZLE_128B)
//   warp 1       MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, M = 128)
//   warps 2..17  epilogue      (4 frame groups x 4 TMEM lane quadrants; thread = channel)
// Accumulators: 2 x 256 TMEM columns, double-buffered across units.
#include "gemm_tc_epi.cuh"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace asrb {

namespace {

constexpr int TCT_STAGES = 4;
constexpr int TCT_STAGE_BYTES = 48 * 1024;
constexpr int TCT_THREADS = 64 + 16 * 32;
constexpr int TCT_SMEM = TCT_STAGES * TCT_STAGE_BYTES + 256;

template <int EPI> struct TctCfg;
template <> struct TctCfg<TC_RES_ACT_DW> {
    static constexpr int NF = 256, A_ROWS = 128, HALO = 1, ROWS_OUT = NF - 2;          // frames per tile / produced
};
template <> struct TctCfg<TC_GLU_DW> {
    static constexpr int NF = 128, A_ROWS = 256, HALO = 8, ROWS_OUT = NF - 16;         // 7 needed each side; 8 keeps groups aligned
};

struct TctParams {
    const float* bias; const float* dw_w; const float* dw_b; const float* pos; float* out32; uint16_t* out;
    int T, K, n_out, tiles_per_utt, m_tiles, n_ct;
    const int32_t* frames_eff;   // ragged batches: tiles whose first produced frame is past frames_eff[b] are skipped
};

// erf-GELU of x given h = x / 2 (common.cuh gelu_fast with the halving folded into the producer of h):
// gelu(x) = h + h tanh(h q'(h^2)),  q'(s) = 2 q(4 s)
__device__ __forceinline__ float gelu_h(float h) {
    const float s = fminf(h * h, 12.25f);
    const float q = fmaf(fmaf(-1.148375808e-02f, s, 2.964082784e-01f), s, 1.594915636f);
    return fmaf(h, tanh_approx(h * q), h);
}
template <int ACT2> __device__ __forceinline__ float act2_h(float ah) {          // act2(2 ah)
    if (ACT2 == ACT_GELU) return gelu_h(ah);
    if (ACT2 == ACT_GELU_GELU) return gelu_h(0.5f * gelu_h(ah));
    return fmaf(ah, tanh_approx(ah), ah);                                        // SiLU: x sigmoid(x) = h + h tanh(h)
}

// The same on a PAIR of values (two neighbouring frames of one channel) with packed f32x2 arithmetic: identical roundings,
// 5 packed + 2 min + 2 MUFU instructions for two GELUs instead of 14.  FFMA2 keeps the FMA pipe busy for two cycles, so the
// pipe does the same work either way; what is saved are issue slots, which the min / MUFU / convert / store instructions
// of these epilogues then find free.
__device__ __forceinline__ float2 gelu_h2(float2 h) {
    float2 s = __fmul2_rn(h, h);
    s.x = fminf(s.x, 12.25f); s.y = fminf(s.y, 12.25f);
    float2 q = __ffma2_rn(make_float2(-1.148375808e-02f, -1.148375808e-02f), s, make_float2(2.964082784e-01f, 2.964082784e-01f));
    q = __ffma2_rn(q, s, make_float2(1.594915636f, 1.594915636f));
    float2 t = __fmul2_rn(h, q);
    t.x = tanh_approx(t.x); t.y = tanh_approx(t.y);
    return __ffma2_rn(h, t, h);
}
template <int ACT2> __device__ __forceinline__ float2 act2_h2(float2 ah) {       // act2(2 ah), pairwise
    if (ACT2 == ACT_GELU) return gelu_h2(ah);
    if (ACT2 == ACT_GELU_GELU) return gelu_h2(__fmul2_rn(make_float2(0.5f, 0.5f), gelu_h2(ah)));
    float2 t = make_float2(tanh_approx(ah.x), tanh_approx(ah.y));                // SiLU
    return __ffma2_rn(ah, t, ah);
}

// TMEM loads without the trailing wait (several are batched before one tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4_nw(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2_nw(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one 16-bit store: OBF = the encoder's result (bf16), otherwise the next MMA's operand (op16)
template <bool OBF> __device__ __forceinline__ void st16(uint16_t* p, float v) {
    if (OBF) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
    else *reinterpret_cast<op16*>(p) = to_op16(v);
}

// Taps and bias of the depthwise-3, halved (the activation that follows takes x / 2)
struct Dw3 { float w0, w1, w2, b; };

// ---- depthwise-3 over one 32-frame chunk (thread = channel).  v[] = accumulator columns S..S+31 of this
// channel (frames tS..tS+31); (pa, pb) = the activated values of frames tS-2, tS-1.  Emits the 32 outputs of
// frames tS-1 .. tS+30 and leaves (pa, pb) = activated frames tS+30, tS+31 for the next chunk.
//   MASK = false: every frame of the chunk is inside [0, T) except possibly frame -1 in v[0] (zero_first), and
//                 every output is this group's except possibly the first two (skip2);
//   MASK = true : general warp-uniform predicates (tile touching the end of the utterance).
// NO = n_out as a compile-time constant (store offsets become immediates) or 0. ----
template <bool MASK, int NO, int ACT2, bool OBF>
__device__ __forceinline__ void dw3_chunk(float (&v)[32], float& pa, float& pb, float hbias, const Dw3& k, const TctParams& p,
                                          int tS, int t_lo, int t_hi, bool skip2, bool zero_first, int64_t row0 /* b*T */, int c) {
    const int ld = NO ? NO : p.n_out;
    const float2 half2 = make_float2(0.5f, 0.5f), hb2 = make_float2(hbias, hbias);
    float2 g[17];                                                             // g[m + 1] = activated frames (tS + 2m, tS + 2m + 1)
    g[0] = make_float2(pa, pb);                                               // frames tS - 2, tS - 1
#pragma unroll
    for (int m = 0; m < 16; ++m) g[m + 1] = gelu_h2(__ffma2_rn(make_float2(v[2 * m], v[2 * m + 1]), half2, hb2));   // GELU(acc + bias)   (model.py:145)
    if (MASK) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (tS + 2 * m < 0 || tS + 2 * m >= p.T) g[m + 1].x = 0.f;
            if (tS + 2 * m + 1 < 0 || tS + 2 * m + 1 >= p.T) g[m + 1].y = 0.f;
        }
    } else if (zero_first) g[1].x = 0.f;                                      // frame -1: conv zero padding
    pa = g[16].x; pb = g[16].y;
    const int64_t e0 = (row0 + tS - 1) * ld + c;                              // element of output j = 0
    const float2 w0 = make_float2(k.w0, k.w0), w2 = make_float2(k.w2, k.w2), kb = make_float2(k.b, k.b);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {                                          // two halves of 16 outputs: bounds the live registers
        float o[16];
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {                                      // outputs j = 2m, 2m + 1 (frames tS - 1 + j): taps on frames j - 2, j - 1, j of the chunk
            const int m = hf * 8 + mm;
            // output j uses activated chunk frames j - 2, j - 1, j; in pairs: (2m-2, 2m-1) = g[m], (2m, 2m+1) = g[m+1], the middle tap straddles both
            float2 t = __ffma2_rn(w0, g[m], kb);
            t.x = fmaf(k.w1, g[m].y, t.x);
            t.y = fmaf(k.w1, g[m + 1].x, t.y);
            t = act2_h2<ACT2>(__ffma2_rn(w2, g[m + 1], t));
            o[2 * mm] = t.x; o[2 * mm + 1] = t.y;
        }
        const int tj = tS - 1 + hf * 16;                                      // frame of o[0]
        const int64_t eh = e0 + (int64_t)(hf * 16) * ld;
        auto live = [&](int jj) { return MASK ? (tj + jj >= t_lo && tj + jj < t_hi) : !(skip2 && hf == 0 && jj < 2); };
        if (p.pos) {                                                          // last block: + sinusoids (model.py:160-161)
            const float* pp = p.pos + (int64_t)tj * ld + c;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) if (live(jj)) o[jj] += __ldg(pp + (int64_t)jj * ld);
        }
        if (p.out32) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) if (live(jj)) p.out32[res32_index(row0 + tj + jj, c, ld)] = o[jj];   // blocked fp32 stream (gemm_tc_epi.cuh)
        }
        uint16_t* op = p.out + eh;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) if (live(jj)) st16<OBF>(op + (int64_t)jj * ld, o[jj]);
    }
}

// ---- depthwise-15 over the 44-frame window h[] (window index i <-> frame tw + i) of one channel: emits the 28
// outputs of frames tw + 8 .. tw + 35: output o = bias + sum_t k[t] h[o + 1 + t].  k[0..14] taps and k[15] bias, halved.
// MASK: the window touches a frame outside [0, T) (conv zero padding; outputs past T are dropped).
// Packed f32x2 arithmetic on PAIRS of neighbouring outputs.  A packed operand must be an aligned register pair
// P_m = (h[2m], h[2m+1]); for outputs (o, o+1) tap t needs (h[o+1+t], h[o+2+t]), which is such a pair only when o + 1 + t is
// even.  So the odd taps are summed for the output pairs (2a, 2a+1) and the even taps (and the bias) for the pairs (2b-1, 2b),
// each with aligned operands only, and one scalar add per output joins the two: 218 FFMA2 + 28 FADD per 28 outputs instead
// of 420 FFMA + 28 FMUL + 28 FADD -- the same work for the FMA pipe, half the issue slots. ----
template <bool MASK, int NO, int ACT2>
__device__ __forceinline__ void dw15_emit(float (&h)[44], const float (&k)[16], uint16_t* op, int ldr, int tw, int T) {
    const int ld = NO ? NO : ldr;
    if (MASK) {
#pragma unroll
        for (int i = 0; i < 44; ++i) if (tw + i < 0 || tw + i >= T) h[i] = 0.f;
    }
    auto P = [&](int m) { return make_float2(h[2 * m], h[2 * m + 1]); };
    float2 kk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) kk[t] = make_float2(k[t], k[t]);
    // F_b: bias + even taps for outputs (2b - 1, 2b); computed one step ahead of the E pair that needs it
    auto even_taps = [&](int b) {
        float2 f = kk[15];
#pragma unroll
        for (int u = 0; u < 8; ++u) f = __ffma2_rn(kk[2 * u], P(b + u), f);
        return f;
    };
    float2 f_lo = even_taps(0);
#pragma unroll
    for (int a = 0; a < 14; ++a) {
        const float2 f_hi = even_taps(a + 1);
        float2 e = __fmul2_rn(kk[1], P(a + 1));                                // odd taps for outputs (2a, 2a + 1)
#pragma unroll
        for (int u = 1; u < 7; ++u) e = __ffma2_rn(kk[2 * u + 1], P(a + 1 + u), e);
        const float2 o = act2_h2<ACT2>(make_float2(e.x + f_lo.y, e.y + f_hi.x));
        if (!MASK || tw + 8 + 2 * a < T) st16<false>(op + (int64_t)(2 * a) * ld, o.x);
        if (!MASK || tw + 8 + 2 * a + 1 < T) st16<false>(op + (int64_t)(2 * a + 1) * ld, o.y);
        f_lo = f_hi;
    }
}

}  // namespace

template <int EPI, int NO, int ACT2, bool OBF>
__global__ void __launch_bounds__(TCT_THREADS, 1)
gemm_tct_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_i, const __grid_constant__ CUtensorMap map_r, const TctParams p) {
    using C = TctCfg<EPI>;
    constexpr bool RES = EPI == TC_RES_ACT_DW;
    constexpr int NF = C::NF, A_BYTES = C::A_ROWS * BK * 2;
    constexpr int STAGES = TCT_STAGES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * TCT_STAGE_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
    const int units = p.m_tiles * p.n_ct;
    const int kb_main = p.K / BK;
    const int num_kb = kb_main + (RES ? 2 : 0);
    const int ld = NO ? NO : p.n_out;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        if (RES) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_i) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // the residual kernel's 16 epilogue warps share a unit; the GLU kernel's two teams of 8 own one accumulator stage each
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), RES ? 16 * 32 : 8 * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        // convergent warp, one elected lane issues (keeps stage / coordinates in uniform registers)
        const bool leader = elect_one();
        int s = 0; uint32_t ph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int m = u / p.n_ct, ct = u - m * p.n_ct;
            const int b = m / p.tiles_per_utt, tb = (m - b * p.tiles_per_utt) * C::ROWS_OUT - C::HALO;
            if (tile_is_padding(p.frames_eff, b, tb + C::HALO)) continue;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_sleep(empty_bar(s), ph ^ 1, 64);
                if (leader) {
                    mbar_expect_tx(full_bar(s), TCT_STAGE_BYTES);
                    const uint32_t sa = smem_u32(smem + s * TCT_STAGE_BYTES);
                    if (!RES || kb < kb_main) {
                        tma_load_2d(sa, &map_w, kb * BK, ct * C::A_ROWS, full_bar(s));
                        tma_load_3d(sa + A_BYTES, &map_x, kb * BK, tb, b, full_bar(s));
                    } else {                                                  // residual: I[:, 64j..] * Y[b, frames, ct*128 + 64j..]^T
                        const int j = kb - kb_main;
                        tma_load_2d(sa, &map_i, j * BK, 0, full_bar(s));
                        tma_load_3d(sa + A_BYTES, &map_r, ct * 128 + j * BK, tb, b, full_bar(s));
                    }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // The whole warp runs the loop convergently (stage / phase / descriptors stay warp-uniform); one elected
        // lane issues the MMAs and commits.
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(NF);
        int s = 0; uint32_t ph = 0; int it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            if (p.frames_eff) {
                const int m = u / p.n_ct, b = m / p.tiles_per_utt;
                if (tile_is_padding(p.frames_eff, b, (m - b * p.tiles_per_utt) * C::ROWS_OUT)) continue;
            }
            const int a = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            ++it;
            mbar_wait_sleep(tempty_bar(a), aph ^ 1, 64);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * 256);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * TCT_STAGE_BYTES);
                const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
                if (leader) {
                    if (RES) {
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            tc_mma(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (uint32_t)((kb | kk) != 0));
                    } else {                                                  // value rows | gate rows of the same 128 channels
                        const uint64_t gdesc = make_smem_desc(sa + 128 * BK * 2);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            tc_mma(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (uint32_t)((kb | kk) != 0));
                            tc_mma(d_tmem + 128, gdesc + 2 * kk, bdesc + 2 * kk, idesc, (uint32_t)((kb | kk) != 0));
                        }
                    }
                    tc_commit(empty_bar(s));
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (leader) tc_commit(tfull_bar(a));
            __syncwarp();
        }
    } else {
        // ================================ epilogue ====================================
        const int ew = warp - 2;
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
        const int g = ew >> 2;                             // frame group
        const int cl = quad * 32 + lane;                   // channel inside the tile
        // per-channel constants, halved (every activation here takes x / 2); reloaded only when the channel tile
        // changes -- with gridDim % n_ct == 0 that is once per kernel
        int ct_cur = -1;
        float hb0 = 0.f, hb1 = 0.f;                        // RES: bias / 2, -            GLU: value bias / 2, gate bias / 2
        float k[RES ? 4 : 16];                             // RES: 3 taps + bias (halved); GLU: 15 taps + bias (halved)
        int it_next = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int m = u / p.n_ct, ct = u - m * p.n_ct;
            const int b = m / p.tiles_per_utt, ft = m - b * p.tiles_per_utt;
            if (tile_is_padding(p.frames_eff, b, ft * C::ROWS_OUT)) continue;
            const int it = it_next++;
            const int tb = ft * C::ROWS_OUT - C::HALO;      // frame of accumulator column 0
            const int a = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            const int c = ct * 128 + cl;                   // output channel of this thread
            const int64_t row0 = (int64_t)b * p.T;
            const uint32_t acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * 256);
            if (ct != ct_cur) {
                ct_cur = ct;
                constexpr int KW = RES ? 3 : 15;
#pragma unroll
                for (int j = 0; j < KW; ++j) k[j] = 0.5f * __ldg(p.dw_w + j * ld + c);
                k[KW] = 0.5f * __ldg(p.dw_b + c);
                if (RES) hb0 = 0.5f * __ldg(p.bias + c);
                else { hb0 = 0.5f * __ldg(p.bias + ct * 256 + cl); hb1 = 0.5f * __ldg(p.bias + ct * 256 + 128 + cl); }
            }

            if constexpr (RES) {
                Dw3 kk; kk.w0 = k[0]; kk.w1 = k[1]; kk.w2 = k[2]; kk.b = k[3];
                // group g emits tile columns [64g - 1, 64g + 63) ∩ [1, 255): frames [t_lo, t_hi)
                const int S0 = 64 * g;
                const int t_lo = tb + max(S0 - 1, 1), t_hi = min(tb + S0 + 63, p.T);
                mbar_wait_backoff(tfull_bar(a), aph, 32, 256);
                tc_fence_after();
                float pa = 0.f, pb = 0.f;
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    const int S = S0 + 32 * ch, tS = tb + S;
                    float v[32];
                    if (ch == 0 && g > 0) {                 // activated frames tS-2, tS-1 belong to the previous group: recompute
                        float e[2];
                        tmem_ld2_nw(acc + S - 2, e);
                        tmem_ld32_nw(acc + S, v);
                        tmem_ld_wait();
                        pa = (tS - 2 < p.T) ? gelu_h(fmaf(e[0], 0.5f, hb0)) : 0.f;
                        pb = (tS - 1 < p.T) ? gelu_h(fmaf(e[1], 0.5f, hb0)) : 0.f;
                    } else {
                        tmem_ld32_nw(acc + S, v);
                        tmem_ld_wait();
                    }
                    if (ch == 1) { tc_fence_before(); mbar_arrive(tempty_bar(a)); }   // last TMEM read of this unit
                    if (tS - 1 >= p.T) continue;            // nothing left to emit (warp-uniform)
                    if (tS + 31 < p.T)
                        dw3_chunk<false, NO, ACT2, OBF>(v, pa, pb, hb0, kk, p, tS, t_lo, t_hi, g == 0 && ch == 0, tS < 0, row0, c);
                    else
                        dw3_chunk<true, NO, ACT2, OBF>(v, pa, pb, hb0, kk, p, tS, t_lo, t_hi, false, false, row0, c);
                }
            } else {
                // ---- GLU -> depthwise-15 -> act2.  Two TEAMS of 8 warps alternate over the units (team = accumulator
                // stage), so while one team waits on TMEM loads at the head of its unit the other is in the FFMA-heavy
                // part of the previous one.  A warp emits two 28-frame groups of its unit one after the other: group gg
                // covers tile columns [8 + 28 gg, 36 + 28 gg) from the 44-column window starting at column 28 gg
                // (window index i <-> frame tw + i) ----
                if ((it & 1) != (g >> 1)) continue;         // the other team's unit
                mbar_wait_backoff(tfull_bar(a), aph, 32, 256);
                tc_fence_after();
                float h[44];
                // (v + bv) * sigmoid(gate + bg) = hv + hv tanh((gate + bg) / 2), hv = (v + bv) / 2, on pairs of frames
                const float2 half2 = make_float2(0.5f, 0.5f), hb0_2 = make_float2(hb0, hb0), hb1_2 = make_float2(hb1, hb1);
                auto glu2 = [&](float& v0, float& v1, float g0, float g1) {
                    const float2 hv = __ffma2_rn(make_float2(v0, v1), half2, hb0_2);
                    float2 ga = __ffma2_rn(make_float2(g0, g1), half2, hb1_2);
                    ga.x = tanh_approx(ga.x); ga.y = tanh_approx(ga.y);
                    const float2 r = __ffma2_rn(hv, ga, hv);
                    v0 = r.x; v1 = r.y;
                };
                {   // pass 0: the 44-column window of the warp's first group
                    const int ws = 28 * ((g & 1) * 2), tw = tb + ws;
                    {
                        float gt[32];
                        tmem_ld32_nw(acc + ws, h);
                        tmem_ld32_nw(acc + 128 + ws, gt);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 2) glu2(h[i], h[i + 1], gt[i], gt[i + 1]);
                    }
                    {
                        float gt[12];
                        tmem_ld8_nw(acc + ws + 32, h + 32);
                        tmem_ld4_nw(acc + ws + 40, h + 40);
                        tmem_ld8_nw(acc + 128 + ws + 32, gt);
                        tmem_ld4_nw(acc + 128 + ws + 40, gt + 8);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 12; i += 2) glu2(h[32 + i], h[33 + i], gt[i], gt[i + 1]);
                    }
                    uint16_t* op = p.out + (row0 + tw + 8) * ld + c;
                    if (tw >= 0 && tw + 43 < p.T) dw15_emit<false, NO, ACT2>(h, k, op, ld, tw, p.T);
                    else dw15_emit<true, NO, ACT2>(h, k, op, ld, tw, p.T);
                }
                {   // pass 1: the next group's window starts 28 columns on -- its first 16 GLU values are pass 0's last 16
                    const int ws = 28 * ((g & 1) * 2 + 1), tw = tb + ws;
                    // (frames outside [0, T) were zeroed by pass 0's masked path under the very rule pass 1 applies)
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = h[i + 28];
                    float gt[28];
                    tmem_ld8_nw(acc + ws + 16, h + 16); tmem_ld8_nw(acc + ws + 24, h + 24); tmem_ld8_nw(acc + ws + 32, h + 32);
                    tmem_ld4_nw(acc + ws + 40, h + 40);
                    tmem_ld8_nw(acc + 128 + ws + 16, gt); tmem_ld8_nw(acc + 128 + ws + 24, gt + 8); tmem_ld8_nw(acc + 128 + ws + 32, gt + 16);
                    tmem_ld4_nw(acc + 128 + ws + 40, gt + 24);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(tempty_bar(a));             // this warp's last TMEM read of the unit
#pragma unroll
                    for (int i = 0; i < 28; i += 2) glu2(h[16 + i], h[17 + i], gt[i], gt[i + 1]);
                    uint16_t* op = p.out + (row0 + tw + 8) * ld + c;
                    if (tw >= 0 && tw + 43 < p.T) dw15_emit<false, NO, ACT2>(h, k, op, ld, tw, p.T);
                    else dw15_emit<true, NO, ACT2>(h, k, op, ld, tw, p.T);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------ host side -------------------------------------------
namespace {

// 2-D K-major 16-bit matrix [rows][cols]: box 64 x box_rows, 128-byte swizzle
int make_mat_map(CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled(matrix %d x %d) -> %d", rows, cols, (int)r);
    return ASRB_OK;
}
// [B][T][C] 16-bit activations as the B operand: dims (C, T, B), box 64 x nf x 1, OOB -> 0
int make_frames_map(CUtensorMap* m, const void* base, int64_t B, int64_t T, int C, int nf) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)nf, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled(frames C=%d T=%lld B=%lld) -> %d", C, (long long)T, (long long)B, (int)r);
    return ASRB_OK;
}

}  // namespace

// 128 x 128 identity (operand format): the A operand of the residual k-blocks.  One per device, created on first use
// (asrb_encoder_create touches it, so a captured forward never allocates).
const op16* tct_identity() {
    static std::mutex mu;
    static op16* table[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!table[dev]) {
        std::vector<op16> h(128 * 128, host_to_op16(0.f));
        for (int i = 0; i < 128; ++i) h[i * 128 + i] = host_to_op16(1.f);
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(op16)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(op16), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        table[dev] = (op16*)d;
    }
    return table[dev];
}

bool tct_supported(const TcGemmArgs& a) {
    if (a.taps != 1 || a.K % 64 != 0 || a.out_f32 || !a.dw_w || !a.dw_b) return false;
    if (a.epilogue == TC_RES_ACT_DW)
        return a.N % 128 == 0 && a.dw_kw == 3 && a.res != nullptr && a.act == ACT_GELU &&
               (a.dw_act == ACT_GELU || (a.dw_act == ACT_GELU_GELU && !a.out_bf16));
    if (a.epilogue == TC_GLU_DW) return a.N % 256 == 0 && a.dw_kw == 15 && !a.pos && !a.out32 && !a.out_bf16 && a.dw_act == ACT_SILU;
    return false;
}

template <int EPI, int NO, int ACT2, bool OBF>
static int launch_tct(const TcGemmArgs& a, cudaStream_t st) {
    using C = TctCfg<EPI>;
    constexpr bool RES = EPI == TC_RES_ACT_DW;
    const int n_out = RES ? a.N : a.N / 2;
    CUtensorMap mw, mx, mi, mr;
    ASRB_TRY(make_mat_map(&mw, a.W, a.N, a.K, C::A_ROWS));
    ASRB_TRY(make_frames_map(&mx, a.A, a.B, a.T, a.K, C::NF));
    if (RES) {
        const op16* ident = tct_identity();
        if (!ident) return fail(ASRB_E_CUDA, "tcgen05 GEMM: identity operand unavailable");
        ASRB_TRY(make_mat_map(&mi, ident, 128, 128, 128));
        ASRB_TRY(make_frames_map(&mr, a.res, a.B, a.T, n_out, C::NF));
    } else { mi = mw; mr = mx; }
    TctParams p{};
    p.bias = a.bias; p.dw_w = a.dw_w; p.dw_b = a.dw_b; p.pos = a.pos; p.out32 = a.out32; p.out = (uint16_t*)a.out;
    p.T = (int)a.T; p.K = a.K; p.n_out = n_out; p.frames_eff = a.frames_eff;
    p.tiles_per_utt = (int)((a.T + C::ROWS_OUT - 1) / C::ROWS_OUT);
    p.m_tiles = (int)(a.B * p.tiles_per_utt);
    p.n_ct = n_out / 128;
    const int units = p.m_tiles * p.n_ct;
    auto kern = gemm_tct_kernel<EPI, NO, ACT2, OBF>;
    ASRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TCT_SMEM));
    const int grid = units < sm_count() ? units : sm_count();
    kern<<<grid, TCT_THREADS, TCT_SMEM, st>>>(mw, mx, mi, mr, p);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

int launch_gemm_tct(const TcGemmArgs& a, cudaStream_t st) {
    // n_out as a compile-time constant for the encoder widths of BASELINE.json (512, 1024); generic otherwise
#define ASRB_TCT(EPI_, ACT_, OBF_, no) \
    ((no) == 512 ? launch_tct<EPI_, 512, ACT_, OBF_>(a, st) : (no) == 1024 ? launch_tct<EPI_, 1024, ACT_, OBF_>(a, st) : launch_tct<EPI_, 0, ACT_, OBF_>(a, st))
    if (a.epilogue == TC_RES_ACT_DW) {
        if (a.dw_act == ACT_GELU_GELU) return ASRB_TCT(TC_RES_ACT_DW, ACT_GELU_GELU, false, a.N);     // feeds the next block's k3 conv
        return a.out_bf16 ? ASRB_TCT(TC_RES_ACT_DW, ACT_GELU, true, a.N) : ASRB_TCT(TC_RES_ACT_DW, ACT_GELU, false, a.N);
    }
    return ASRB_TCT(TC_GLU_DW, ACT_SILU, false, a.N / 2);
#undef ASRB_TCT
}

}  // namespace asrb
