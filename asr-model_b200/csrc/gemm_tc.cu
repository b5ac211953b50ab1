// Persistent, warp-specialised tcgen05 GEMM / implicit k-tap conv for the encoder (sm_100a).
//
//   out[b,t,:] = epilogue( sum_{tap,k} A[b, t+tap-taps/2, k] * W[n][tap*K + k] )
//
// A is channels-last op16 (the 16-bit operand format, common.cuh) [B][T][K]; a 3-D TMA map (K, T, B) with OOB zero fill supplies the
// conv halo (t = -1, t = T) and the ragged last tile of every utterance for free.  W is
// op16 [N][taps*K], K-major.  Accumulators live in TMEM (fp32).  Roles per CTA (1 CTA / SM):
//   warp 0      TMA producer   (A box 64x128, W box 64xBN, SWIZZLE_128B, 3-5 stage ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN)
//               both run their loops convergently with one elected lane issuing, so stage, phase and
//               descriptors stay in uniform registers and the UTCHMMAs of a k-block go out back to back
//   warps 2..   epilogue       (BN/64 column groups x four TMEM lane quadrants; thread = row)
// Epilogues (fp32 math, 16-bit store through swizzled smem + TMA store, which also clips the
// rows past T):  bias+act | bias+residual+act | bias(+residual)+LayerNorm over the full row (N <= 512:
// a CTA pair owns one 256-column half each and swaps row sums over DSMEM).  The two epilogues with a
// fused depthwise convolution (TC_GLU_DW, TC_RES_ACT_DW) live in gemm_tct.cu (lanes = channels).
#include "gemm_tc_epi.cuh"
#include <cstdlib>

namespace asrb {

template <int BN, int EPI>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_ah,
               const __grid_constant__ CUtensorMap map_ak, const TcParams p) {
    using C = TcCfg<BN>;
    constexpr int STAGES = C::STAGES, NG = C::NG;
    extern __shared__ __align__(1024) unsigned char smem[];                  // SWIZZLE_128B tiles need 1024-B alignment
    unsigned char* scratch = smem + STAGES * C::STAGE_BYTES;                 // NG x GROUP_SCRATCH
    float2* s_xch = reinterpret_cast<float2*>(scratch + NG * GROUP_SCRATCH); // [NG][BM] LN partial (sum, sumsq)
    float2* s_peer = s_xch + NG * BM;                                        // [2][BM] row sums written by the peer CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + NG * GROUP_SCRATCH + C::XCH_BYTES);
    // bars: full[STAGES], empty[STAGES], tfull[2], tempty[2], xbar[2] (peer row sums have landed)
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    auto xbar = [&](int a) { return bar0 + 8u * (2 * STAGES + 4 + a); };
    auto a_full = [&](int i) { return bar0 + 8u * (2 * STAGES + 6 + i); };   // k3one: the 130-row frame tile of a k-block
    auto a_empty = [&](int i) { return bar0 + 8u * (2 * STAGES + 8 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 10);
    // k3one layout of the operand area: [2 frame tiles of 130 rows x 128 B, 17 KB apart][STAGES weight tiles of B_BYTES]
    constexpr int A1_STRIDE = 17 * 1024, A1_BYTES = 130 * 128;
    static_assert(2 * A1_STRIDE + STAGES * C::B_BYTES <= STAGES * C::STAGE_BYTES, "k3one operand area");

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
    constexpr bool is_ln = EPI == TC_LN;
    const int cl = is_ln ? p.cluster : 1;                                    // CTAs per cluster (cluster LayerNorm)
    const uint32_t crank = cl > 1 ? cluster_ctarank() : 0;
    const int u_first = blockIdx.x / cl, u_step = gridDim.x / cl;
    const int nbu = is_ln ? (cl > 1 ? 1 : p.n_chunks) : 1;                   // accumulator chunks per unit (this CTA)
    const int units = is_ln ? p.m_tiles : p.m_tiles * p.n_chunks;
    const int acc_stages = (512 / (nbu * BN)) >= 2 ? 2 : 1;
    const int kb_per_tap = p.K / BK;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.taps / 2;
    const bool k3one = p.k3one != 0 && p.taps == 3 && nbu == 1;              // one accumulator chunk per unit: frame tile fetched once per k-block

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        if (is_ln) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ak) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // pair (cl = 2): both CTAs read the same frame tile, so each loads half of it and multicasts it to both; a
        // slot is free once BOTH CTAs' MMAs have consumed it (two arrivals on empty)
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), k3one ? 1u : (uint32_t)cl); }
        for (int i = 0; i < 2; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), (uint32_t)cl); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NG * 128); mbar_init(xbar(a), BM); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {                                                         // TMEM: all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    if (cl > 1) cluster_sync_all(); else __syncthreads();                    // peers' barriers are initialised before anyone arrives
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        // The warp runs the loop convergently (stage, phase and coordinates stay in uniform registers);
        // one elected lane issues.
        const bool leader = elect_one();
        int s = 0; uint32_t ph = 0;
        int ab = 0; uint32_t aph = 0;                                         // k3one: frame-tile ring
        for (int u = u_first; u < units; u += u_step) {
            const int m = is_ln ? u : u / p.n_chunks;
            const int nb0 = is_ln ? (int)crank : u % p.n_chunks;
            const int b = m / p.tiles_per_utt, t0 = (m % p.tiles_per_utt) * p.rows_out - p.halo;
            if (tile_is_padding(p.frames_eff, b, t0)) continue;               // ragged batch: nothing of this tile is needed
            if (k3one) {
                // one 130-row frame tile per k-block (rows t0 - 1 .. t0 + 128; in a pair CTA 0 fetches rows 0..63, CTA 1 rows
                // 64..129, both multicast), then the three taps' weight tiles
                const int n0 = nb0 * BN;
                for (int kc = 0; kc < kb_per_tap; ++kc) {
                    mbar_wait_sleep(a_empty(ab), aph ^ 1, 32);
                    if (leader) {
                        mbar_expect_tx(a_full(ab), A1_BYTES);
                        const uint32_t aa = smem_u32(smem + ab * A1_STRIDE);
                        if (cl == 1) tma_load_3d(aa, &map_ak, kc * BK, t0 - 1, b, a_full(ab));                     // 130-row box
                        else if (crank == 0) tma_load_3d_mc(aa, &map_ah, kc * BK, t0 - 1, b, a_full(ab), (uint16_t)3);
                        else tma_load_3d_mc(aa + 64 * 128, &map_ak, kc * BK, t0 - 1 + 64, b, a_full(ab), (uint16_t)3);   // 66-row box
                    }
                    __syncwarp();
                    if (++ab == 2) { ab = 0; aph ^= 1; }
                    for (int tap = 0; tap < 3; ++tap) {
                        mbar_wait_sleep(empty_bar(s), ph ^ 1, 32);
                        if (leader) {
                            mbar_expect_tx(full_bar(s), C::B_BYTES);
                            tma_load_2d(smem_u32(smem + 2 * A1_STRIDE + s * C::B_BYTES), &map_w, (tap * kb_per_tap + kc) * BK, n0, full_bar(s));
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                }
                continue;
            }
            for (int j = 0; j < nbu; ++j) {
                const int n0 = (nb0 + j) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_sleep(empty_bar(s), ph ^ 1, 32);
                    if (leader) {
                        mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
                        const int tap = kb / kb_per_tap, kc = kb - tap * kb_per_tap;
                        const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                        if (is_ln && cl > 1)       // my half of the shared frame tile (64 rows = 8 KB), delivered to both CTAs
                            tma_load_3d_mc(sa + crank * (C::A_BYTES / 2), &map_ah, kc * BK, t0 + tap - pad + (int)crank * (BM / 2), b, full_bar(s), (uint16_t)3);
                        else
                            tma_load_3d(sa, &map_a, kc * BK, t0 + tap - pad, b, full_bar(s));
                        tma_load_2d(sa + C::A_BYTES, &map_w, kb * BK, n0, full_bar(s));
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // convergent warp, one elected lane issues: the UTCHMMAs of a k-block go out back to back
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(BN);
        int s = 0; uint32_t ph = 0; int it = 0;
        int ab = 0; uint32_t abph = 0;                                        // k3one: frame-tile ring
        for (int u = u_first; u < units; u += u_step) {
            if (p.frames_eff) {
                const int m = is_ln ? u : u / p.n_chunks;
                if (tile_is_padding(p.frames_eff, m / p.tiles_per_utt, (m % p.tiles_per_utt) * p.rows_out - p.halo)) continue;
            }
            const int a = it % acc_stages;
            const uint32_t aph = (uint32_t)(it / acc_stages) & 1u;
            ++it;
            mbar_wait_sleep(tempty_bar(a), aph ^ 1, 32);
            tc_fence_after();
            if (k3one) {
                // tap `tap` of a k-block reads the frame tile from row `tap` on: the descriptor simply starts 128 B x tap into the
                // tile.  The 128-byte swizzle is a function of the shared-memory ADDRESS (bits 4-6 ^= bits 7-9), for the TMA
                // write and for the MMA read alike, so a start that is not 1024-byte aligned needs nothing else (measured: with
                // the descriptor's base-offset field set to the row phase the results are wrong, with 0 they are exact)
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
                for (int kc = 0; kc < kb_per_tap; ++kc) {
                    mbar_wait(a_full(ab), abph);
                    tc_fence_after();
                    for (int tap = 0; tap < 3; ++tap) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t aa = smem_u32(smem + ab * A1_STRIDE) + (uint32_t)(tap * 128);
                            const uint64_t adesc = make_smem_desc(aa);
                            const uint64_t bdesc = make_smem_desc(smem_u32(smem + 2 * A1_STRIDE + s * C::B_BYTES));
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk)
                                tc_mma(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (uint32_t)((kc | tap | kk) != 0));
                            tc_commit(empty_bar(s));                          // weight slots are this CTA's own
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                    if (leader) { if (cl > 1) tc_commit_mc(a_empty(ab), (uint16_t)3); else tc_commit(a_empty(ab)); }   // pair: the peer writes part of my frame tile
                    __syncwarp();
                    if (++ab == 2) { ab = 0; abph ^= 1; }
                }
                if (leader) tc_commit(tfull_bar(a));
                __syncwarp();
                continue;
            }
            for (int j = 0; j < nbu; ++j) {
                const uint32_t d_tmem = tmem_base + (uint32_t)((a * nbu + j) * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + C::A_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            tc_mma(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (uint32_t)((kb | kk) != 0));
                        if (is_ln && cl > 1) tc_commit_mc(empty_bar(s), (uint16_t)3);   // the peer writes half of my slot: tell both
                        else tc_commit(empty_bar(s));                         // frees the smem slot when the MMAs retire
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
            if (leader) tc_commit(tfull_bar(a));                              // accumulator complete
            __syncwarp();
        }
    } else {
        // ================================ epilogue ====================================
        // NG groups of 4 warps; a group owns a slice of the tile's columns, its warps own the four
        // TMEM lane quadrants (warp % 4), so in phase 1 a thread is one row of the tile.
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int group = ew >> 2;
        const int r = quad * 32 + lane;                   // row inside the tile
        const bool leader = (ew & 3) == 0 && lane == 0;   // issues this group's TMA stores
        unsigned char* stg = scratch + group * GROUP_SCRATCH;
        const uint32_t stg_u32 = smem_u32(stg);
        const int bar_id = 1 + group;
        int it_next = 0;
        for (int u = u_first; u < units; u += u_step) {
            const int m = is_ln ? u : u / p.n_chunks;
            const int nb0 = is_ln ? (int)crank : u % p.n_chunks;
            const int b = m / p.tiles_per_utt, t0 = (m % p.tiles_per_utt) * p.rows_out - p.halo;   // frame of tile row 0
            if (tile_is_padding(p.frames_eff, b, t0)) continue;
            const int it = it_next++;
            const int a = it % acc_stages;
            const uint32_t aph = (uint32_t)(it / acc_stages) & 1u;
            const bool row_ok = t0 + r >= 0 && t0 + r < p.T;
            const int64_t grow = (int64_t)b * p.T + t0 + r;
            mbar_wait_backoff(tfull_bar(a), aph, 32, 512);  // 16 warps wait here for most of a tile time: back off, leave issue slots and power to the MMA side
            tc_fence_after();
            const uint32_t acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * nbu * BN);

            if (EPI == TC_LN) {
                // ---- bias (+ residual) + LayerNorm over the whole row.  The row is nbu chunks side by side
                // in this CTA's TMEM; with cl = 2 the other half lives in the peer CTA and the two swap
                // their (sum, sum of squares) over distributed shared memory ----
                const int ncols = nbu * BN / NG, c_begin = group * ncols;     // this thread's slice of the accumulator
                const int gbase = nb0 * BN;                                   // global column of accumulator column 0
                float s1 = 0.f, s2 = 0.f;
                for (int c = c_begin; c < c_begin + ncols; c += 32) {
                    float v[32];
                    tmem_ld32(acc + c, v);
                    add_vec32(v, p.bias + gbase + c);
                    if (row_ok) {
                        if (p.res32) add_f32x32(v, p.res32, grow, gbase + c, p.N);
                        else if (p.res) add_res32(v, p.res + grow * p.N + gbase + c);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
                }
                s_xch[group * BM + r] = make_float2(s1, s2);
                epi_bar(NG + 1, NG * 128);
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) { const float2 o = s_xch[gq * BM + r]; t1 += o.x; t2 += o.y; }
                if (cl > 1) {
                    const int pb = it & 1;
                    if (group == 0) {                                         // one thread per row ships the half-row sums
                        st_cluster_f2(map_to_cta(smem_u32(s_peer + pb * BM + r), crank ^ 1), make_float2(t1, t2));
                        mbar_arrive_cluster(map_to_cta(xbar(pb), crank ^ 1));
                    }
                    mbar_wait_cluster(xbar(pb), (uint32_t)(it >> 1) & 1u);
                    const float2 o = s_peer[pb * BM + r];
                    t1 += o.x; t2 += o.y;
                }
                epi_bar(NG + 1, NG * 128);                                    // all partners have read before the next tile writes
                const float mean = t1 / (float)p.N;
                const float rstd = rsqrtf(fmaxf(t2 / (float)p.N - mean * mean, 0.f) + p.eps);
                for (int c = c_begin; c < c_begin + ncols; c += 32) {
                    float v[32];
                    tmem_ld32(acc + c, v);
                    add_vec32(v, p.bias + gbase + c);
                    if (row_ok) {
                        if (p.res32) add_f32x32(v, p.res32, grow, gbase + c, p.N);
                        else if (p.res) add_res32(v, p.res + grow * p.N + gbase + c);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + gbase + c) + i);
                        const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + gbase + c) + i);
                        v[4 * i + 0] = fmaf((v[4 * i + 0] - mean) * rstd, gm.x, bt.x);
                        v[4 * i + 1] = fmaf((v[4 * i + 1] - mean) * rstd, gm.y, bt.y);
                        v[4 * i + 2] = fmaf((v[4 * i + 2] - mean) * rstd, gm.z, bt.z);
                        v[4 * i + 3] = fmaf((v[4 * i + 3] - mean) * rstd, gm.w, bt.w);
                    }
                    if (p.out32 && row_ok) store_f32x32(p.out32, grow, gbase + c, p.N, v);   // fp32 copy: the next residual stream
                    const int cb = (c - c_begin) & 32;                        // which half of the 64-column staging tile
                    if (cb == 0) { if (leader) tma_wait_read0(); epi_bar(bar_id, 128); }
                    stage_store32(stg, r, cb, v, p.out_bf16 != 0);
                    if (cb == 32) {
                        fence_async_smem();
                        epi_bar(bar_id, 128);
                        if (leader) { tma_store_3d(&map_out, stg_u32, gbase + c - 32, t0, b); tma_commit(); }
                    }
                }
            } else {
                // ---- bias (+ residual) + activation; 64 output columns per group ----
                constexpr int COLS = BN / NG;
                static_assert(COLS == 64, "plain epilogue: one 64-column staging tile per group");
                const int tc0 = group * COLS;                                 // column inside the accumulator
                const int gc = nb0 * BN + tc0;                                // global output column
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
                    tmem_ld32(acc + tc0 + 32 * hh, v);
                    add_vec32(v, p.bias + gc + 32 * hh);
                    if (EPI == TC_RES_ACT && row_ok) add_res32(v, p.res + grow * p.n_out + gc + 32 * hh);
                    act_fast32(v, p.act);
                    if (p.out_f32) {                                          // consumer is not an MMA: keep fp32 (32-column boxes)
                        if (leader) tma_wait_read0();
                        epi_bar(bar_id, 128);
                        stage_store32_f32(stg, r, v);
                        fence_async_smem();
                        epi_bar(bar_id, 128);
                        if (leader) { tma_store_3d(&map_out, stg_u32, gc + 32 * hh, t0, b); tma_commit(); }
                    } else {
                        if (hh == 0) { if (leader) tma_wait_read0(); epi_bar(bar_id, 128); }
                        stage_store32(stg, r, 32 * hh, v, p.out_bf16 != 0);
                        if (hh == 1) {
                            fence_async_smem();
                            epi_bar(bar_id, 128);
                            if (leader) { tma_store_3d(&map_out, stg_u32, gc, t0, b); tma_commit(); }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(a));                                       // accumulator drained
        }
        if (leader) tma_wait_all();
    }

    tc_fence_before();
    if (cl > 1) cluster_sync_all(); else __syncthreads();                    // no CTA exits while its peer may still write to it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------ host side -------------------------------------------
// [B][T][C] fp32 output: box 32 x 128 x 1 (128-byte rows), 128-byte swizzle
static int make_out_f32_map(CUtensorMap* m, const void* base, int64_t B, int64_t T, int C) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)T * C * 4};
    cuuint32_t box[3] = {32, BM, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled(fp32 out C=%d) -> %d", C, (int)r);
    return ASRB_OK;
}
static int make_w_map(CUtensorMap* m, const void* base, int N, int Ktot, int bn) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASRB_E_CUDA, "cuTensorMapEncodeTiled(weights N=%d K=%d) -> %d", N, Ktot, (int)r);
    return ASRB_OK;
}

static int pick_bn(int N, int epi) {
    if (epi == TC_LN) return (N % 256 == 0) ? 256 : 128;
    return (N % 256 == 0) ? 256 : 128;
}
int tc_glu_tile_n(int) { return 256; }

bool tc_gemm_supported(int K, int N, int epi) {
    if (K % 64 != 0 || N % 128 != 0) return false;
    if (epi == TC_GLU) return false;                 // superseded by TC_GLU_DW
    if (epi == TC_GLU_DW) return N % 256 == 0;       // channel-major kernels (gemm_tct.cu)
    if (epi == TC_RES_ACT_DW) return true;
    if (epi == TC_LN && N > 512) return false;
    return true;
}

template <int BN, int EPI>
static int launch_one(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const CUtensorMap& mah,
                      const CUtensorMap& mak, const TcParams& p, int units, cudaStream_t st) {
    auto kern = gemm_tc_kernel<BN, EPI>;
    ASRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
    const int cl = p.cluster;
    int grid = sm_count() / cl * cl;
    if (grid > units * cl) grid = units * cl;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TcCfg<BN>::THREADS);
    cfg.dynamicSmemBytes = TcCfg<BN>::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ASRB_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, mo, mah, mak, p));
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// Validate one problem and build its tensor maps + kernel parameters (shared with gemm_tc_dual.cu).
int tc_prepare(const TcGemmArgs& a, CUtensorMap* ma, CUtensorMap* mw, CUtensorMap* mo, TcParams* pp, int* bn_out) {
    if (!tc_gemm_supported(a.K, a.N, a.epilogue))
        return fail(ASRB_E_ARG, "tcgen05 GEMM: unsupported shape K=%d N=%d epilogue=%d", a.K, a.N, a.epilogue);
    const int bn = pick_bn(a.N, a.epilogue);
    if (a.epilogue == TC_GLU_DW || a.epilogue == TC_RES_ACT_DW)
        return fail(ASRB_E_ARG, "tcgen05 GEMM: fused-depthwise epilogue %d with act=%d dw_act=%d kw=%d is not built", a.epilogue, a.act, a.dw_act, a.dw_kw);
    const int n_out = a.N;
    const int halo = 0, rows_out = BM;
    ASRB_TRY(make_act_map(ma, a.A, a.B, a.T, a.K));
    ASRB_TRY(make_w_map(mw, a.W, a.N, a.taps * a.K, bn));
    if ((a.res32 || a.out32) && a.epilogue != TC_LN) return fail(ASRB_E_ARG, "tcgen05 GEMM: fp32 residual streams are a LayerNorm-epilogue feature");
    if (a.out_f32 && a.epilogue == TC_LN) return fail(ASRB_E_ARG, "tcgen05 GEMM: LayerNorm epilogue stores 16-bit tiles only");
    if (a.out_f32) ASRB_TRY(make_out_f32_map(mo, a.out, a.B, a.T, n_out));
    else ASRB_TRY(make_act_map(mo, a.out, a.B, a.T, n_out));
    TcParams& p = *pp;
    p.bias = a.bias; p.res = a.res; p.res32 = a.res32; p.out32 = a.out32; p.gamma = a.gamma; p.beta = a.beta;
    p.T = (int)a.T; p.K = a.K; p.N = a.N; p.taps = a.taps; p.act = a.act;
    p.halo = halo; p.rows_out = rows_out;
    p.tiles_per_utt = (int)((a.T + rows_out - 1) / rows_out);
    p.m_tiles = (int)(a.B * p.tiles_per_utt);
    p.n_chunks = a.N / bn; p.n_out = n_out; p.eps = a.eps; p.out_f32 = a.out_f32; p.out_bf16 = a.out_bf16;
    p.frames_eff = a.frames_eff; p.k3one = 0;
    p.cluster = (a.epilogue == TC_LN && bn == 256 && a.N == 512) ? 2 : 1;     // the 128 x 512 row block fills TMEM: split it over a CTA pair
    *bn_out = bn;
    return ASRB_OK;
}

int launch_gemm_tc(const TcGemmArgs& a, cudaStream_t st) {
    if (a.B <= 0 || a.T <= 0) return ASRB_OK;
    static const char* const tags[6] = {"gemm_tc_bias_act", "gemm_tc_glu", "gemm_tc_res_act", "gemm_tc_layernorm",
                                        "gemm_tc_glu_dw15_silu", "gemm_tc_res_gelu_dw3_gelu"};
    if ((a.epilogue == TC_RES_ACT_DW || a.epilogue == TC_GLU_DW) && tct_supported(a)) {   // channel-major kernels (gemm_tct.cu)
        const int no = a.epilogue == TC_GLU_DW ? a.N / 2 : a.N;
        ProfScope ps(tags[a.epilogue], st, 2.0 * a.B * a.T * (double)a.N * a.K,
                     2.0 * a.B * a.T * ((double)a.K + no + (a.res ? no : 0)) + 2.0 * a.N * a.K);
        return launch_gemm_tct(a, st);
    }
    CUtensorMap ma, mw, mo;
    TcParams p;
    int bn = 0;
    ASRB_TRY(tc_prepare(a, &ma, &mw, &mo, &p, &bn));
    const int n_out = p.n_out;
    const int units = a.epilogue == TC_LN ? p.m_tiles : p.m_tiles * p.n_chunks;
    ProfScope ps(tags[a.epilogue], st, 2.0 * a.B * a.T * (double)a.N * a.K * a.taps,
                 2.0 * a.B * a.T * ((double)a.K + n_out * (a.out_f32 ? 2 : 1) + (a.res ? n_out : 0)) + 2.0 * a.N * a.K * a.taps);
    CUtensorMap mah = ma;                                                    // 64-row boxes of the same tensor (pair multicast)
    CUtensorMap mak = ma;                                                    // k3one: the 130-row frame tile (pair: its rows 64..129, 66-row boxes)
    if (p.cluster > 1) ASRB_TRY(make_act_map(&mah, a.A, a.B, a.T, a.K, BM / 2));
    if (a.taps == 3) ASRB_TRY(make_act_map(&mak, a.A, a.B, a.T, a.K, p.cluster > 1 ? 66 : 130));
    { static int k3 = -1; if (k3 < 0) { const char* e = getenv("ASRB_K3ONE"); k3 = (e && e[0] == '0') ? 0 : 1; } p.k3one = k3; }   // ASRB_K3ONE=0: one load per tap (A/B)
#define ASRB_TC(BN_, EPI_) return launch_one<BN_, EPI_>(ma, mw, mo, mah, mak, p, units, st)
    if (bn == 256) {
        switch (a.epilogue) {
            case TC_BIAS_ACT: ASRB_TC(256, TC_BIAS_ACT);
            case TC_RES_ACT: ASRB_TC(256, TC_RES_ACT);
            case TC_LN: ASRB_TC(256, TC_LN);
        }
    } else {
        switch (a.epilogue) {
            case TC_BIAS_ACT: ASRB_TC(128, TC_BIAS_ACT);
            case TC_RES_ACT: ASRB_TC(128, TC_RES_ACT);
            case TC_LN: ASRB_TC(128, TC_LN);
        }
    }
#undef ASRB_TC
    return fail(ASRB_E_ARG, "tcgen05 GEMM: no kernel for BN=%d epilogue=%d", bn, a.epilogue);
}

}  // namespace asrb

// Test hook: the kernel in isolation (include/asrb200.h).
extern "C" int asrb_test_gemm_tc(const void* a, const void* w, const float* bias, const void* res, const float* gamma,
                                 const float* beta, void* out, int64_t B, int64_t T, int K, int N, int taps,
                                 int epilogue, int act, const float* dw_w, const float* dw_b, int dw_kw, int dw_act,
                                 const float* pos, void* stream) {
    using namespace asrb;
    if (!a || !w || !bias || !out) return fail(ASRB_E_ARG, "asrb_test_gemm_tc: NULL tensor");
    if (epilogue < 0 || epilogue > 5 || (taps != 1 && taps != 3)) return fail(ASRB_E_ARG, "asrb_test_gemm_tc: bad epilogue/taps");
    ASRB_TRY(require_sm100());
    TcGemmArgs g{};
    g.A = (const op16*)a; g.W = (const op16*)w; g.bias = bias; g.res = (const op16*)res;
    g.gamma = gamma; g.beta = beta; g.out = out;
    g.B = B; g.T = T; g.K = K; g.N = N; g.taps = taps; g.epilogue = epilogue; g.act = act; g.eps = 1e-5f;
    g.dw_w = dw_w; g.dw_b = dw_b; g.dw_kw = dw_kw; g.dw_act = dw_act; g.pos = pos;
    return launch_gemm_tc(g, (cudaStream_t)stream);
}
