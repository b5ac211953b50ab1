// Flash-style softmax attention on tcgen05 / TMEM for the optional TransformerEncoderLayer
// (model.py:138,163: nn.MultiheadAttention, no mask, non-causal, softmax(q k^T / sqrt(hd)) v).
//
// One CTA = one (utterance, head, 128-query tile); two neighbouring query tiles form a CLUSTER that shares every K and V
// tile: each CTA fetches half of it (64 keys) and TMA-multicasts it into both, so the L2 -> SM traffic of the kernel --
// its bound at T = 3001: 9.4 GB per call when every CTA streams all of K and V on its own -- is halved.  qkv is the packed op16 (16-bit operand format, common.cuh) [B][T][3D] output of
// the in_proj GEMM; Q, K and V tiles arrive by TMA (3-D map, OOB rows zero filled).
//   warp 0      TMA producer: Q once, then a 2-deep ring of K tiles and a 2-deep ring of V tiles
//   warp 1      MMA issuer:   S_j = Q K_j^T (SS, both K-major)  ->  TMEM S[j&1] (128 fp32 columns)
//                             O  += P_j V_j (A = P in TENSOR MEMORY, B = V in smem MN-major)  -> TMEM O
//   warps 2..9  softmax:      two warps per TMEM lane quadrant; a thread is a query row (TMEM lane) and owns ONE 64-key half
//                             of every key tile (kept in registers between the max and the exp pass); the two halves
//                             swap their row maxima through shared memory (one named barrier per tile) and keep
//                             separate row sums until the end.  Online softmax in the exp2 domain with a stale running
//                             max (O is rescaled in TMEM only when the max grew by more than 2^8, each half rescaling its
//                             half of the columns), P_j -> op16 -> TMEM (two keys per 32-bit column: the PV MMA takes its A operand straight from
//                             tensor memory, so P never crosses shared memory, whose bandwidth -- 128-wide MMAs read their
//                             operands at the full 128 B/clk while TMA refills the K / V rings -- is this kernel's bound),
//                             final O / l -> out op16.
//                             (Round 1 had one warp per quadrant: 1.5 warps per scheduler, 37 % tensor-pipe active.)
// S is never written to HBM; scores and probabilities live in TMEM / shared memory only.
#include "tc_common.cuh"

namespace asrb {

template <int HD> struct AttnCfg {
    static constexpr int SUB = HD / 64;                          // 64-column sub-tiles per head_dim
    static constexpr int Q_BYTES = BM * HD * 2;
    static constexpr int KV_BYTES = BM * HD * 2;                 // one K (or V) tile: 128 keys x HD
    static constexpr int SMEM = Q_BYTES + 4 * KV_BYTES + 256;
    static constexpr int THREADS = 64 + 256;                     // TMA + MMA warps, 8 softmax warps
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// Tq query frames against Tk key/value frames; q / k / v of head h start at columns q_col0 / k_col0 / v_col0 + h * HD of their
// tensors (packed qkv [B][T][3D]: 0 / D / 2D of one tensor; a cached K|V [B][Tk][2D] next to separate queries: 0 / 0 / D)
struct AttnParams { int Tq, Tk, D, H, n_kv, q_col0, k_col0, v_col0; float scale_log2; op16* out; };

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// MN-major, SWIZZLE_128B B operand (V tile: rows = keys (K), 64 head-dim columns (N) per 128-byte row):
// 8-key groups 1024 B apart (SBO), the next 64 columns of N one sub-tile (16 KB) further (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) |
           (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_ex(int n, bool b_mn_major) {
    return (1u << 4) | (ASRB_OP16_IS_F16 ? 0u : ((1u << 7) | (1u << 10))) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
// A operand from tensor memory (lane = row, two 16-bit K elements per column), B from a shared-memory descriptor
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(AttnCfg<HD>::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {   // map_kv: 64-key boxes
    using C = AttnCfg<HD>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* s_q = smem;                                   // [SUB][128][128 B]
    unsigned char* s_k = s_q + C::Q_BYTES;                       // [2][SUB][128][128 B]
    unsigned char* s_v = s_k + 2 * C::KV_BYTES;                  // [2][SUB][128][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + 2 * C::KV_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    // q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_ready[2], pv_done[2]
    const uint32_t q_full = bar0;
    auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bar0 + 8u * (3 + s); };
    auto v_full = [&](int s) { return bar0 + 8u * (5 + s); };
    auto v_empty = [&](int s) { return bar0 + 8u * (7 + s); };
    auto s_full = [&](int s) { return bar0 + 8u * (9 + s); };
    auto p_ready = [&](int s) { return bar0 + 8u * (11 + s); };
    auto pv_done = [&](int s) { return bar0 + 8u * (13 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    __shared__ float s_xch[2][2][BM];                            // [tile parity][key half][query row]: row maxima, at the end row sums

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
    const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = p.n_kv;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            // a K / V slot is half written by the peer CTA: it is free once BOTH CTAs' MMAs have consumed it
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2);
            mbar_init(s_full(s), 1); mbar_init(p_ready(s), 256); mbar_init(pv_done(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                              // the peer's barriers are initialised before anyone multicasts into it
    tc_fence_after();
    const uint32_t crank = cluster_ctarank();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_s = tmem_base;                 // S[2]: columns [0, 256)
    const uint32_t tm_o = tmem_base + 256;           // O: columns [256, 256 + HD)
    const uint32_t tm_p = tmem_base + 384;           // P[2]: 2 x 64 columns (128 keys x 16 bit per row)

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(q_full, C::Q_BYTES);
            for (int s = 0; s < C::SUB; ++s)
                tma_load_3d(smem_u32(s_q + s * (BM * 128)), &map_q, p.q_col0 + h * HD + s * 64, q0, b, q_full);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(k_empty(st), ph ^ 1);
                mbar_expect_tx(k_full(st), C::KV_BYTES);
                for (int s = 0; s < C::SUB; ++s)     // my 64 keys of the tile, delivered to both CTAs of the pair
                    tma_load_3d_mc(smem_u32(s_k + st * C::KV_BYTES + s * (BM * 128)) + crank * (64 * 128), &map_kv,
                                   p.k_col0 + h * HD + s * 64, j * BM + (int)crank * 64, b, k_full(st), (uint16_t)3);
                mbar_wait(v_empty(st), ph ^ 1);
                mbar_expect_tx(v_full(st), C::KV_BYTES);
                for (int s = 0; s < C::SUB; ++s)
                    tma_load_3d_mc(smem_u32(s_v + st * C::KV_BYTES + s * (BM * 128)) + crank * (64 * 128), &map_kv,
                                   p.v_col0 + h * HD + s * 64, j * BM + (int)crank * 64, b, v_full(st), (uint16_t)3);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_ex(BM, false);        // N = 128 keys
            constexpr uint32_t idesc_pv = make_idesc_ex(HD, true);         // N = head_dim, B = V is MN-major
            auto issue_qk = [&](int j) {
                const int st = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(k_full(st), ph);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * (BM * 128) + (kk & 3) * 32);
                    tc_mma(tm_s + (uint32_t)(st * BM), make_smem_desc(smem_u32(s_q) + off),
                           make_smem_desc(smem_u32(s_k + st * C::KV_BYTES) + off), idesc_qk, (uint32_t)(kk != 0));
                }
                tc_commit_mc(k_empty(st), (uint16_t)3);
                tc_commit(s_full(st));
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            issue_qk(0);
            if (n_kv > 1) issue_qk(1);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(p_ready(st), ph);                                 // P_j in smem, S[st] free, O rescaled if needed
                mbar_wait(v_full(st), ph);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < BM / 16; ++kk) {                      // 128 keys = 8 MMAs of K = 16
                    const uint64_t bdesc = make_smem_desc_mn(smem_u32(s_v + st * C::KV_BYTES) + (uint32_t)(kk * 16 * 128), BM * 128);
                    tc_mma_ts(tm_o, tm_p + (uint32_t)(st * 64 + kk * 8), bdesc, idesc_pv, (uint32_t)((j | kk) != 0));
                }
                tc_commit_mc(v_empty(st), (uint16_t)3);
                tc_commit(pv_done(st));
                if (j + 2 < n_kv) issue_qk(j + 2);
            }
        }
    } else {
        // ================================ softmax warps ===============================
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                            // which 64 keys of every 128-key tile
        const int r = quad * 32 + lane;                              // query row of this thread
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        float m_used = -INFINITY;                                    // max the exponents are taken against (same in both halves)
        float l = 0.f;                                               // this half's share of the row sum
        for (int j = 0; j < n_kv; ++j) {
            const int st = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            mbar_wait(s_full(st), ph);
            tc_fence_after();
            const uint32_t ts = tm_s + lane_off + (uint32_t)(st * BM + half * 64);
            const int kbase = j * BM + half * 64;
            // ---- pass A: this half's scores into registers, row max of the raw scores (the scale is positive) ----
            float sc[64];
            tmem_ld32(ts, *reinterpret_cast<float(*)[32]>(sc));
            tmem_ld32(ts + 32, *reinterpret_cast<float(*)[32]>(sc + 32));
            if (kbase + 64 > p.Tk) {                                 // keys >= Tk are padding: mask them
#pragma unroll
                for (int i = 0; i < 64; ++i) if (kbase + i >= p.Tk) sc[i] = -INFINITY;
            }
            float mx = sc[0];
#pragma unroll
            for (int i = 1; i < 64; ++i) mx = fmaxf(mx, sc[i]);
            s_xch[st][half][r] = mx;
            epi_bar(1, 256);                                         // both halves of every row have published their max
            mx = fmaxf(mx, s_xch[st][half ^ 1][r]) * p.scale_log2;
            // ---- rescale O only when the max grew by more than 2^8 (warp-uniform decision, identical in both halves) ----
            const bool grow = mx > m_used + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? mx : m_used;
                const float factor = (m_used == -INFINITY) ? 0.f : ex2f(m_used - m_new);
                if (j > 0) {
                    mbar_wait(pv_done((j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);     // O is quiescent
                    tc_fence_after();
#pragma unroll
                    for (int c = half * (HD / 2); c < (half + 1) * (HD / 2); c += 32) {  // this half's columns of O
                        float o[32];
                        tmem_ld32(tm_o + lane_off + c, o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] *= factor;
                        tmem_st32(tm_o + lane_off + c, o);
                    }
                    tc_fence_before();
                }
                l *= factor;
                m_used = m_new;
            }
            // ---- pass B: p = 2^(s * scale - m_used), row sum, P -> op16 -> TMEM (A operand of P V: row = lane, two keys per column) ----
            float rs = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) { sc[i] = ex2f(fmaf(sc[i], p.scale_log2, -m_used)); rs += sc[i]; }
            l += rs;
            float pw[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pw[i] = __uint_as_float(pack_op16x2(sc[2 * i], sc[2 * i + 1]));
            if (j >= 2) mbar_wait(pv_done(st), (uint32_t)((j - 2) >> 1) & 1u);          // P buffer st is free again
            tc_fence_after();
            tmem_st32(tm_p + lane_off + (uint32_t)(st * 64 + half * 32), pw);
            tc_fence_before();
            mbar_arrive(p_ready(st));
        }
        // ---- epilogue: O / l -> op16 -> out[b, q0 + r, h*HD ...]; the halves add their row sums and split the columns ----
        s_xch[n_kv & 1][half][r] = l;
        epi_bar(1, 256);
        l += s_xch[n_kv & 1][half ^ 1][r];
        mbar_wait(pv_done((n_kv - 1) & 1), (uint32_t)((n_kv - 1) >> 1) & 1u);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int t = q0 + r;
#pragma unroll
        for (int c = half * (HD / 2); c < (half + 1) * (HD / 2); c += 32) {
            float o[32];
            tmem_ld32(tm_o + lane_off + c, o);
            if (t < p.Tq) {
                op16* dst = p.out + ((int64_t)b * p.Tq + t) * p.D + h * HD + c;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 q;
                    q.x = pack_op16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv); q.y = pack_op16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
                    q.z = pack_op16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv); q.w = pack_op16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + 8 * i) = q;
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();                              // no CTA exits while its peer may still write to it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

bool attention_tc_supported(int D, int H) {
    const int hd = H > 0 ? D / H : 0;
    return H > 0 && D % H == 0 && (hd == 64 || hd == 128) && (3 * D) % 8 == 0;
}

template <int HD>
static int launch_attn(const CUtensorMap& mq, const CUtensorMap& mkv, const AttnParams& p, int64_t B, cudaStream_t st) {
    auto kern = attn_tc_kernel<HD>;
    ASRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<HD>::SMEM));
    const unsigned q_tiles = (unsigned)((p.Tq + BM - 1) / BM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((q_tiles + 1) & ~1u, (unsigned)p.H, (unsigned)B);   // pairs of query tiles (an odd last tile gets an idle partner that still loads its half)
    cfg.blockDim = dim3(AttnCfg<HD>::THREADS);
    cfg.dynamicSmemBytes = AttnCfg<HD>::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ASRB_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mkv, p));
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// q [B][Tq][ldq] and kv [B][Tk][ldkv] op16 (head h of q / k / v at columns *_col0 + h * hd) -> out [B][Tq][D] op16,
// softmax(q k^T * scale) v per head, no mask.
int launch_attention_tc_ex(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, void* out,
                           int64_t B, int64_t Tq, int64_t Tk, int D, int H, float scale, cudaStream_t st) {
    if (!attention_tc_supported(D, H) || (ldq & 7) || (ldkv & 7))
        return fail(ASRB_E_ARG, "tcgen05 attention: head_dim %d unsupported (64 or 128)", H ? D / H : 0);
    if (B <= 0 || Tq <= 0 || Tk <= 0) return ASRB_OK;
    if (B > 65535) return fail(ASRB_E_ARG, "tcgen05 attention: batch %lld > 65535", (long long)B);
    CUtensorMap mq, mkv;
    ASRB_TRY(make_act_map(&mq, q, B, Tq, ldq));
    ASRB_TRY(make_act_map(&mkv, kv, B, Tk, ldkv, BM / 2));              // 64-key boxes: each CTA of a pair loads half a tile
    AttnParams p;
    p.Tq = (int)Tq; p.Tk = (int)Tk; p.D = D; p.H = H; p.n_kv = (int)((Tk + BM - 1) / BM);
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.scale_log2 = scale * 1.4426950408889634f; p.out = (op16*)out;
    ProfScope ps("attention_tc", st, 4.0 * B * (double)Tq * Tk * D, 2.0 * B * D * (2.0 * Tq + 2.0 * Tk));
    if (D / H == 128) return launch_attn<128>(mq, mkv, p, B, st);
    return launch_attn<64>(mq, mkv, p, B, st);
}

// qkv [B][T][3D] op16 (q | k | v) -> out [B][T][D] op16, softmax(q k^T * scale) v per head.
int launch_attention_tc(const void* qkv, void* out, int64_t B, int64_t T, int D, int H, float scale, cudaStream_t st) {
    return launch_attention_tc_ex(qkv, 3 * D, 0, qkv, 3 * D, D, 2 * D, out, B, T, T, D, H, scale, st);
}

}  // namespace asrb

// Test hook (include/asrb200.h)
extern "C" int asrb_test_attention_tc(const void* qkv, void* out, int64_t B, int64_t T, int D, int H, void* stream) {
    using namespace asrb;
    if (!qkv || !out || H <= 0) return fail(ASRB_E_ARG, "asrb_test_attention_tc: bad argument");
    ASRB_TRY(require_sm100());
    return launch_attention_tc(qkv, out, B, T, D, H, 1.0f / sqrtf((float)(D / H)), (cudaStream_t)stream);
}
