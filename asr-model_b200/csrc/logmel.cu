// Fused log-mel front end (essentials.py:469-491 batched): framing + Hann window +
// real FFT + |X|^2 + HTK mel projection + log10/clamp/normalise in ONE pass over the PCM.
//
// Data layout in HBM:  pcm [B][stride] fp32 (read once, staged in shared memory where the
// 2.5x / 6.4x frame overlap lives);  out [B][M][T] fp32 (written once) or, on the fused path, the stem
// GEMM's operand [B][T][CP] in the 16-bit operand format;  keys [B] uint32 -- order-preserving image of
// each utterance's running max of log10(mel);  optionally pool_out [B][target]: the average-pooled
// `waveform` feature (essentials.py:493-503), taken from the PCM span this pass has staged anyway.
//
// One block = FB consecutive frames of one utterance; a thread group of R threads owns a QUAD of four
// consecutive frames.  Two real frames ride one complex FFT (z = a + i b) and a thread carries TWO such FFTs
// side by side in the two lanes of packed f32x2 registers (fft_regs.cuh), so the whole transform issues as
// FADD2 / FMUL2 / FFMA2.  An N = R*R point FFT is two rounds of R-point DFTs held in registers (R = 20 for
// n_fft 400, 32 for n_fft 1024) with one transpose through shared memory in between (16-byte accesses, odd
// pitches: conflict free).  The conjugate-symmetric partner Z[N-k] that the split of the packed pair needs
// lives in exactly one other thread (R - j), so only the upper half of the spectrum is exchanged.  Power
// spectra go to shared memory as [bin][frame] rows; the banded mel projection reads them back four frames at
// a time (one LDS.128 + two FFMA2 per tap), then MUFU lg2.
// The per-utterance dynamic-range floor (max - 8) needs the max of the WHOLE utterance, so this pass writes
// (log10 + 4) / 4 and every warp publishes its max (and the tile's min) with one reduction each; logmel_floor_kernel then raises the
// values below the floor (monotone, so max((x+4)/4, (floor+4)/4) == (max(x, floor)+4)/4 bit for bit) and
// only writes where something changes.
#include "logmel.cuh"
#include "tc_common.cuh"
#include "fft_regs.cuh"
#include <vector>
#include <cmath>

namespace asrb {

struct LogmelParams {
    const float* pcm; int64_t stride; int64_t n_samples; const int32_t* lengths;
    int hop; int T; int M; int kmax;
    const float* window;       // [NFFT]
    const float2* twiddle;     // [R][R]: W_N^(k1*j) at [k1*R + j]
    const int* mel_lo;         // [M] first bin of filter m
    const int* mel_cnt;        // [M] taps (0 for an all-zero filter)
    const float* mel_w;        // [M][kmax], scaled by 1/4 (the split of the packed pair leaves 2 Re, 2 Im)
    float* out;                // [B][M][T] fp32 (reference layout), or NULL with out_cl set
    op16* out_cl;              // fused path: [B][T][CP] op16 channels-last, channels >= M zero (the stem GEMM's operand)
    int CP;
    uint32_t* keys;            // [B]
    uint32_t* tile_min;        // [B * tiles_per_utt]: order-preserving image of each tile's smallest log10(mel) (the floor pass skips tiles above the floor)
    float* pool_out;           // optional [B][pool_target]: adaptive average pooling of the PCM (essentials.py:493-503)
    int64_t pool_target;
    int a_bytes;               // shared-memory area A (see the kernel): max(partner exchange, staged PCM span + output staging tile)
};

template <int NFFT, int R, int FB, int HOP>
struct LogmelCfg {
    static constexpr int QUADS = FB / 4;               // frame quads = thread groups per block
    static constexpr int THREADS = QUADS * R;
    static constexpr int NB = NFFT / 2 + 1;            // one-sided bins
    static constexpr int YP = R + 1, EP = R / 2 + 1;   // row pitches in float4 (odd: 16-byte accesses are conflict free)
    static constexpr int Y_BYTES = QUADS * R * YP * 16;
    static constexpr int E_BYTES = QUADS * R * EP * 16;
    static constexpr int PP = FB + 4;                  // floats per power row (pitch 9 / 5 chunks of 16 B)
    static constexpr int P_ROWS = NB + 4;              // zero-weight padding taps read up to 3 rows past the last bin
    static constexpr int P_BYTES = P_ROWS * PP * 4;
    // Shared-memory timeline of one tile (area A = the first a_bytes, run-time: >= E_BYTES and >= PCM span + staging tile):
    //   PCM span (in A)  -> registers -> Y (transpose, A + beyond)  -> E (partner exchange, in A) | P (power rows, after A)
    //   -> during the mel stage A is dead again: the NEXT tile's PCM span streams into it, next to the 16-bit output staging
    __host__ __device__ static constexpr int region(int a_bytes) { return Y_BYTES > a_bytes + P_BYTES ? Y_BYTES : a_bytes + P_BYTES; }
    // hop = 160, R = 20: the four frames of a quad start 640 samples = 0 banks apart from the next quad, whose
    // first 12 threads share a warp with this quad's 20: a gap of 20 words per quad in the staged span puts them on
    // the 12 banks the first 20 leave free.  640 is a multiple of R, so which side of a gap a sample falls on is a
    // compile-time property of (frame in quad, n1).
    static constexpr bool SKEW = HOP == 160 && R == 20;
    static constexpr int GAP = SKEW ? 20 : 0;
    static_assert(R * R == NFFT, "two-round FFT needs n_fft = R^2");
    static_assert(E_BYTES + NB * PP * 4 >= Y_BYTES, "the transpose buffer must not reach the power rows' zero padding");
    static_assert((QUADS & (QUADS - 1)) == 0, "frame quads per block: a power of two");
};

// cp.async with zero fill: copies `bytes` (0..size) from src and zero-fills the rest of `size`
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 pack4(const cx2& z) { return make_float4(z.re.x, z.re.y, z.im.x, z.im.y); }
__device__ __forceinline__ cx2 unpack4(const float4& v) { return {make_float2(v.x, v.y), make_float2(v.z, v.w)}; }

// N taps-of-four of one filter against four frames' power rows: every load is issued before the first FFMA2
template <int N, int PP>
__device__ __forceinline__ void mel_taps(const float* px, const float4* w4, V2& a01, V2& a23) {
    float4 w[N], pv[N][4];
#pragma unroll
    for (int t = 0; t < N; ++t) {
        w[t] = w4[t];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[t][i] = *reinterpret_cast<const float4*>(px + (4 * t + i) * PP);
    }
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const float ws[4] = {w[t].x, w[t].y, w[t].z, w[t].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a01 = vfma(ws[i], make_float2(pv[t][i].x, pv[t][i].y), a01);
            a23 = vfma(ws[i], make_float2(pv[t][i].z, pv[t][i].w), a23);
        }
    }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {     // TMA 1-D bulk copy
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Persistent: gridDim.x blocks walk the (utterance, frame-tile) list; constants are staged once per block and the
// PCM span of the NEXT tile streams in under the mel stage -- TMA bulk copies for interior tiles, cp.async with zero fill
// outside [0, len) (= the center=True padding) for the first / last tiles of an utterance.  HOP = 0: hop is a run-time value (no bank skew).
template <int NFFT, int R, int FB, int HOP>
__global__ void __launch_bounds__(LogmelCfg<NFFT, R, FB, HOP>::THREADS)
logmel_kernel(const LogmelParams p, int tiles_per_utt, int total_tiles) {
    using C = LogmelCfg<NFFT, R, FB, HOP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int hop = HOP ? HOP : p.hop;
    const int span = (FB - 1) * hop + NFFT;
    const int span4 = (span + 3) & ~3;
    const int pcm_words = span4 + C::GAP * C::QUADS + 4;
    float4* s_y   = reinterpret_cast<float4*>(smem_raw);
    float4* s_e   = reinterpret_cast<float4*>(smem_raw);
    float*  s_p   = reinterpret_cast<float*>(smem_raw + p.a_bytes);
    float*  s_pcm = reinterpret_cast<float*>(smem_raw);                         // [pcm_words], area A
    op16*   s_cl  = reinterpret_cast<op16*>(smem_raw + ((pcm_words * 4 + 15) & ~15));   // area A, behind the PCM span
    float*  s_win = reinterpret_cast<float*>(smem_raw + C::region(p.a_bytes));  // [NFFT]
    float2* s_tw  = reinterpret_cast<float2*>(s_win + NFFT);                    // [R*R]
    float*  s_melw = reinterpret_cast<float*>(s_tw + R * R);                    // [M][kmax]
    int2*   s_meta = reinterpret_cast<int2*>(s_melw + p.M * p.kmax);            // [M]: (first bin * row pitch, taps / 4 rounded up)
    __shared__ __align__(8) uint64_t s_bar;                                     // completion of the TMA-fetched PCM span

    const int tid = threadIdx.x;
    const bool vec_ok = ((p.stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0) && ((hop & 3) == 0);
    const int clp = p.CP + 4;                                // staging pitch in 16-bit words: rows stay 8-byte aligned, transposed writes 2-way conflicted at worst

    auto pos = [&](int s) { return C::SKEW ? s + C::GAP * (s / (4 * 160)) : s; };       // span index -> shared-memory word
    // Async copy of one tile's PCM span; everything past the utterance (and before sample 0) is zero filled.  All index
    // arithmetic is 32-bit and relative to the span; with the bank skew a thread's chunk in round `it` lies in quad `it`
    // (THREADS * 4 samples = one quad of hops), so the gap offset is a compile-time constant per round.
    // A tile whose whole span lies inside the utterance (all but the first and the last two of an utterance) is fetched by
    // ONE thread with TMA bulk copies (one per frame quad when the bank skew is on) that complete on an mbarrier; the edge
    // tiles, which need zero fill, go through cp.async.
    // Tile geometry without divisions: a block walks the tile list in steps of gridDim.x, so (utterance, tile in utterance)
    // advance by a constant pair with one carry, and every "frame t is a frame of the utterance" test is t * hop <= len
    // (t < 1 + len / hop).  The length of an utterance is read once per tile, when its PCM span is requested.
    const int step_b = (int)gridDim.x / tiles_per_utt, step_t = (int)gridDim.x - step_b * tiles_per_utt;
    auto span_geom = [&](int t0, int64_t len, int64_t& s0, int& lo, int& hi) {
        s0 = (int64_t)t0 * hop - NFFT / 2;
        lo = s0 < 0 ? (int)(-s0) : 0;                                              // first span index that is a real sample
        const int64_t rem = len - s0;
        hi = rem < 0 ? 0 : (rem > span4 ? span4 : (int)rem);                       // one past the last real sample
    };
    const uint32_t pcm_bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(s_pcm);
    auto prefetch = [&](int b, int ti, int64_t len) -> bool {                      // returns: fetched by TMA (wait on the mbarrier)
        const int t0 = ti * FB;
        if (len < (int64_t)t0 * hop) { cp_async_commit(); return false; }          // a padding tile (t0 >= 1 + len / hop) reads no PCM
        int lo, hi; int64_t s0;
        span_geom(t0, len, s0, lo, hi);
        const float* src = p.pcm + (int64_t)b * p.stride + s0;                     // span index 0 (never dereferenced outside [lo, hi))
        if (vec_ok && lo == 0 && hi == span4) {
            if (tid == 0) {
                mbar_expect_tx(pcm_bar, (uint32_t)span4 * 4u);
                if (C::SKEW) {
                    for (int c = 0; c * 640 < span4; ++c) {
                        const int n = span4 - c * 640 < 640 ? span4 - c * 640 : 640;
                        bulk_g2s(d0 + 4u * (uint32_t)(c * (640 + C::GAP)), src + c * 640, (uint32_t)n * 4u, pcm_bar);
                    }
                } else bulk_g2s(d0, src, (uint32_t)span4 * 4u, pcm_bar);
            }
            return true;
        }
        if (vec_ok) {
            for (int i = tid * 4; i < span4; i += C::THREADS * 4) {
                int valid = hi - i;                                               // samples available from i
                valid = (i < lo || valid < 0) ? 0 : (valid > 4 ? 4 : valid);
                // (a source size of 0 reads nothing: the address may lie outside the utterance)
                cp_async16(d0 + 4u * (uint32_t)pos(i), (const void*)(src + i), valid * 4);
            }
        } else {
            for (int i = tid; i < span4; i += C::THREADS) {
                const bool in = i >= lo && i < hi;
                cp_async4(d0 + 4u * pos(i), in ? (const void*)(src + i) : (const void*)p.pcm, in ? 4 : 0);
            }
        }
        cp_async_commit();
        return false;
    };

    int tile = blockIdx.x;
    if (tid == 0) { mbar_init(pcm_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    bool by_tma = false; uint32_t tma_phase = 0;
    int nb = tile / tiles_per_utt, nti = tile - nb * tiles_per_utt;                // the one division of the block
    int64_t nlen = 0;                                                              // (nb, nti, nlen): the tile whose PCM is in flight
    if (tile < total_tiles) { nlen = clamp_len(p.lengths, nb, p.n_samples); by_tma = prefetch(nb, nti, nlen); }
    // ---- constants, once per block ----
    for (int i = tid; i < NFFT; i += C::THREADS) s_win[i] = p.window[i];
    for (int i = tid; i < R * R; i += C::THREADS) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < p.M * p.kmax; i += C::THREADS) s_melw[i] = p.mel_w[i];
    for (int i = tid; i < p.M; i += C::THREADS) s_meta[i] = make_int2(p.mel_lo[i] * C::PP, (p.mel_cnt[i] + 3) >> 2);
    for (int i = tid; i < 4 * C::PP; i += C::THREADS) s_p[C::NB * C::PP + i] = 0.f;    // rows the zero-weight padding taps touch: finite

    const int q = tid / R, j = tid - q * R;                  // frame quad, position inside the FFT
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
    constexpr int nwarp = C::THREADS / 32;
    float4* yq = s_y + q * (R * C::YP);
    float4* eq = s_e + q * (R * C::EP);

    for (; tile < total_tiles; tile += gridDim.x) {
        if (by_tma) { mbar_wait(pcm_bar, tma_phase); tma_phase ^= 1u; } else cp_async_wait<0>();
        __syncthreads();                                     // this tile's PCM (and the constants) are visible; the previous tile's readers are done

        const int b = nb, t0 = nti * FB;                     // this tile
        const int64_t len = nlen;
        auto request_next = [&]() {                          // (nb, nti, nlen) move on to the tile the block takes next; its PCM is requested
            nti += step_t; nb += step_b;
            if (nti >= tiles_per_utt) { nti -= tiles_per_utt; ++nb; }
            const bool has_next = tile + (int)gridDim.x < total_tiles;
            nlen = has_next ? clamp_len(p.lengths, nb, p.n_samples) : 0;
            if (has_next) by_tma = prefetch(nb, nti, nlen);
        };

        if (len < (int64_t)t0 * hop) {
            // ---- a tile wholly past the utterance (t0 >= 1 + len / hop, ragged batch): DataCollator's 0.0 padding, no transform ----
            __syncthreads();
            request_next();
            const int nf = min(FB, p.T - t0);
            if (p.out_cl) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(p.out_cl + ((int64_t)b * p.T + t0) * p.CP);
                for (int i = tid; i < nf * (p.CP >> 1); i += C::THREADS) dst[i] = 0u;
            } else {
                for (int i = tid; i < p.M * nf; i += C::THREADS) {
                    const int m = i / nf, f = i - m * nf;
                    p.out[((int64_t)b * p.M + m) * p.T + t0 + f] = 0.f;
                }
            }
            if (tid == 0) p.tile_min[tile] = f2key(INFINITY);
            continue;
        }

        // ---- optional: the average-pooled `waveform` feature of the bins this tile's frames name.  Bin i covers
        // samples [floor(i n / target), ceil((i + 1) n / target)): inside the staged span whenever n is a multiple
        // of hop (then the bin IS one hop); whatever falls outside comes from global memory. ----
        if (p.pool_out) {
            const int64_t n = p.n_samples, tg = p.pool_target;
            const int64_t s0 = (int64_t)t0 * hop - NFFT / 2;
            const float* src = p.pcm + (int64_t)b * p.stride;
            for (int i = warp; i < FB; i += nwarp) {
                const int64_t bin = (int64_t)t0 + i;
                if (bin >= tg) break;
                const int64_t s = (bin * n) / tg, e = ((bin + 1) * n + tg - 1) / tg;
                float acc = 0.f;
                for (int64_t k = s + lane; k < e; k += 32) {
                    const int64_t rel = k - s0;
                    acc += (rel >= 0 && rel < span) ? s_pcm[pos((int)rel)] : src[k];
                }
                acc = warp_sum(acc);
                if (lane == 0) p.pool_out[(int64_t)b * tg + bin] = acc / (float)(e - s);
            }
        }

        // ---- round 1: window, R-point DFT over n1 of z[R n1 + j], twiddle W_N^(j k1), transpose ----
        {
            cx2 x[R];
            // frames 4q .. 4q+3: (f0, f1) are the real parts of the two FFTs, (f2, f3) their imaginary parts
            const float* fq = s_pcm + q * (4 * hop + C::GAP) + j;
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1) {
                const float w = s_win[R * n1 + j];
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool wrap = C::SKEW && (i * 160 + R * n1 >= 4 * 160);       // past the quad's gap
                    v[i] = fq[i * hop + R * n1 + (wrap ? C::GAP : 0)];
                }
                x[n1].re = vmul(w, make_float2(v[0], v[1]));
                x[n1].im = vmul(w, make_float2(v[2], v[3]));
            }
            __syncthreads();                                 // the PCM span is in registers everywhere: the transpose may overwrite it
            SmallDFT<R>::run(x);
            yq[j] = pack4(x[0]);
            // twiddles are fetched five at a time AHEAD of the stores they feed: the compiler cannot move a shared-memory load
            // above a store that might alias it, and one load per store leaves every load's latency exposed
#pragma unroll
            for (int k0 = 1; k0 < R; k0 += 5) {
                float2 tw[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) if (k0 + i < R) tw[i] = s_tw[(k0 + i) * R + j];
#pragma unroll
                for (int i = 0; i < 5; ++i) if (k0 + i < R) yq[(k0 + i) * C::YP + j] = pack4(cmul_cs(x[k0 + i], tw[i].x, tw[i].y));
            }
        }
        __syncthreads();

        // ---- round 2: thread k1 = j does the R-point DFT over n2 -> Z[j + R k2] ----
        cx2 z[R];
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) z[n2] = unpack4(yq[j * C::YP + n2]);
        SmallDFT<R>::run(z);
        __syncthreads();                                     // everyone has read the transpose: E and P may overwrite it

        // ---- conjugate partner exchange: Z[N - k] of k = j + R k2 (k2 < R/2) is thread (R - j)'s Z at
        // k2' = R - 1 - k2 (thread 0: its own, at R - k2), i.e. always in an upper half ----
#pragma unroll
        for (int k2 = R / 2; k2 < R; ++k2) eq[j * C::EP + (k2 - R / 2)] = pack4(z[k2]);
        eq[j * C::EP + R / 2] = pack4(z[0]);                 // thread 0's partner of k = 0 is Z[0] itself
        __syncthreads();

        // ---- split the packed pair: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i; |.|^2 of
        // both for both FFTs = four frames of one bin -> one 16-byte store (the 1/4 lives in the mel weights) ----
        {
            const float4* ep = s_e + q * (R * C::EP) + (j ? R - j : 0) * C::EP + (j == 0 ? 1 : 0);
            float* pk = s_p + j * C::PP + 4 * q;
            cx2 zp[R / 2];                                   // all partners first (the upper half of z is dead: registers are free),
#pragma unroll                                               // then arithmetic and stores -- loads cannot pass the power-row stores
            for (int k2 = 0; k2 < R / 2; ++k2) zp[k2] = unpack4(ep[R / 2 - 1 - k2]);
#pragma unroll
            for (int k2 = 0; k2 < R / 2; ++k2) {
                const cx2 z2 = zp[k2];
                const V2 ar = vadd(z[k2].re, z2.re), ai = vsub(z[k2].im, z2.im);       // 2 Re A, 2 Im A
                const V2 br = vadd(z[k2].im, z2.im), bi = vsub(z2.re, z[k2].re);       // 2 Re B, 2 Im B
                const V2 pa = vfma2(ar, ar, vmul2(ai, ai)), pb = vfma2(br, br, vmul2(bi, bi));
                *reinterpret_cast<float4*>(pk + (R * k2) * C::PP) = make_float4(pa.x, pa.y, pb.x, pb.y);
            }
            if (j == 0) {                                    // bin N/2 (self-conjugate): A = Re Z, B = Im Z
                const V2 ar = vadd(z[R / 2].re, z[R / 2].re), br = vadd(z[R / 2].im, z[R / 2].im);
                const V2 pa = vmul2(ar, ar), pb = vmul2(br, br);
                *reinterpret_cast<float4*>(pk + (R * (R / 2)) * C::PP) = make_float4(pa.x, pa.y, pb.x, pb.y);
            }
        }
        __syncthreads();

        // area A is dead until the next tile's transpose: the next PCM span streams into it under the mel stage
        request_next();

        // ---- banded mel projection + log10 + (x+4)/4.  Work item = (filter m, frame quad): a thread keeps its quad and walks
        // the filters R apart; the 8 (4) lanes that share m read one power row per tap, conflict free, and the weights arrive
        // as broadcast 16-byte loads.  Taps go in groups of at most 3 x 4 with every load issued before the first FFMA2.
        // Max / min are tracked on the raw lg2 values (scaling by log10(2) is monotone); sv = (log10 + 4) / 4 is one FMUL (to
        // log10: keeps log10(1e-10) = -10 and hence silence = -1.5 exact) and one FFMA (essentials.py:488-490). ----
        float vmax = -INFINITY, vmin = INFINITY;            // lg2 of the tile's largest / smallest mel value (valid frames)
        {
            const int qq = tid & (C::QUADS - 1);
            const int tq = t0 + 4 * qq;                      // first frame of this thread's quad
            const bool full = (int64_t)(t0 + FB - 1) * hop <= len;   // every frame of the tile is a frame of the utterance (t0 + FB <= Tb <= T)
            const int M = p.M, kmax4 = p.kmax >> 2, T = p.T;
            const float* pq = s_p + 4 * qq;
            const float4* wbase = reinterpret_cast<const float4*>(s_melw);
            float* obase = p.out ? p.out + (int64_t)b * M * T + tq : nullptr;
            op16* cbase = s_cl + (4 * qq) * clp;
            // a thread's filters are R apart, walking its group index up and down in turns (g, 2R-1-g, 2R+g, ...): filter
            // bands widen with m, and this way every warp gets the same share of taps before the barrier
            // (the sequence g, 2R-1-g, 2R+g, 4R-1-g, ... is increasing: two running indices, no select, ends at the first m >= M)
            for (int m = tid / C::QUADS, m_next = 2 * R - 1 - m; m < M; ) {
                const int2 meta = s_meta[m];
                int n4 = meta.y;
                const float4* w4 = wbase + m * kmax4;
                const float* px = pq + meta.x;
                V2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
                while (n4 > 3) { mel_taps<3, C::PP>(px, w4, a01, a23); n4 -= 3; w4 += 3; px += 12 * C::PP; }
                if (n4 == 3) mel_taps<3, C::PP>(px, w4, a01, a23);
                else if (n4 == 2) mel_taps<2, C::PP>(px, w4, a01, a23);
                else if (n4 == 1) mel_taps<1, C::PP>(px, w4, a01, a23);
                float lg[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                for (int i = 0; i < 4; ++i)                  // MUFU lg2 (abs error ~2^-22); the argument is >= 1e-10: never denormal
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg[i]) : "f"(fmaxf(lg[i], 1e-10f)));
                auto norm = [](float l2) { return fmaf(l2 * 0.30102999566398120f, 0.25f, 1.0f); };
                if (full) {
                    vmax = fmaxf(vmax, fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3])));
                    vmin = fminf(vmin, fminf(fminf(lg[0], lg[1]), fminf(lg[2], lg[3])));
                    if (!obase) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) cbase[i * clp + m] = to_op16(norm(lg[i]));
                    } else {
                        float* o = obase + (int64_t)m * T;
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i] = norm(lg[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int t = tq + i;
                        if (t < T) {
                            float sv = 0.f;                                     // DataCollator pad value; t * hop <= len: t < 1 + len / hop, a frame of the utterance
                            if ((int64_t)t * hop <= len) { vmax = fmaxf(vmax, lg[i]); vmin = fminf(vmin, lg[i]); sv = norm(lg[i]); }
                            if (!obase) cbase[i * clp + m] = to_op16(sv);
                            else obase[(int64_t)m * T + i] = sv;
                        }
                    }
                }
                { const int t = m + 2 * R; m = m_next; m_next = t; }
            }
        }
        // each warp publishes its own max / min with fire-and-forget reductions (nothing waits for them): no block-wide
        // reduction, and without the staging tile no barrier either -- the next tile's first barrier orders the reuse of
        // the power rows
        vmax = warp_max(vmax);
        vmin = -warp_max(-vmin);
        if (lane == 0) {                                    // back to log10: monotone, so the max of the products is the product of the max
            if (vmax > -INFINITY) atomicMax(p.keys + b, f2key(vmax * 0.30102999566398120f));
            atomicMin(p.tile_min + tile, f2key(vmin));      // lg2 domain; +inf from a warp without valid frames: never below a floor
        }
        if (p.out_cl) {                                     // rows of CP 16-bit values leave as 8-byte words: at CP = 128 a warp moves one frame per load / store pair
            __syncthreads();                                // staging tile complete
            const int dpr = p.CP >> 2, M = p.M;             // 8-byte words per row
            const int nrows = min(FB, p.T - t0);
            uint2* dst = reinterpret_cast<uint2*>(p.out_cl + ((int64_t)b * p.T + t0) * p.CP);
            const uint2* src = reinterpret_cast<const uint2*>(s_cl);
            const int spitch = clp >> 2;
            for (int w = lane; w < dpr; w += 32) {
                // channels 4w .. 4w+3: keep the real ones, the rest (>= M) are zero
                const int c0 = 4 * w;
                const uint32_t k0 = c0 + 1 < M ? 0xffffffffu : (c0 < M ? 0xffffu : 0u);
                const uint32_t k1 = c0 + 3 < M ? 0xffffffffu : (c0 + 2 < M ? 0xffffu : 0u);
                const int ws = c0 < M ? w : 0;              // always a staged word
#pragma unroll 2
                for (int fr = warp; fr < nrows; fr += nwarp) {
                    const uint2 v = src[fr * spitch + ws];
                    dst[(int64_t)fr * dpr + w] = make_uint2(v.x & k0, v.y & k1);
                }
            }
        }
    }
}

// essentials.py:489: log_mel = maximum(log_mel, log_mel.max() - 8.0), applied in the normalised domain.  One block per
// frame tile of pass 1; a tile whose smallest value already clears the floor (pass 1 left its minimum in tile_min) is
// skipped without touching memory, and the others only write where something changes.
__device__ __forceinline__ bool floor_tile_is_clear(const uint32_t* keys, int64_t batch, int tiles_per_utt, int b, int tile, float& floor_s) {
    floor_s = ((key2f(keys[b]) - 8.0f) + 4.0f) / 4.0f;
    const float lg2_min = key2f(keys[batch + (int64_t)b * tiles_per_utt + tile]);
    return fmaf(lg2_min * 0.30102999566398120f, 0.25f, 1.0f) >= floor_s;           // the very expression pass 1 stored
}

__global__ void logmel_floor_kernel(float* out, const uint32_t* keys, int64_t batch, const int32_t* lengths,
                                    int64_t n_samples, int hop, int M, int T, int FB, int tiles_per_utt) {
    const int b = blockIdx.y, tile = blockIdx.x;
    float floor_s;
    if (floor_tile_is_clear(keys, batch, tiles_per_utt, b, tile, floor_s)) return;
    const int Tb = 1 + (int)(clamp_len(lengths, b, n_samples) / hop);
    const int t0 = tile * FB, nf = min(FB, Tb - t0);
    float* o = out + (int64_t)b * M * T + t0;
    for (int i = threadIdx.x; i < M * FB; i += blockDim.x) {
        const int m = i / FB, f = i - m * FB;
        if (f < nf) {
            float* q = o + (int64_t)m * T + f;
            if (*q < floor_s) *q = floor_s;
        }
    }
}

// The same floor on the fused path's 16-bit channels-last tensor, in place: rounding is monotone, so
// max(rn(x), rn(floor)) == rn(max(x, floor)) bit for bit.  One thread per (frame, 8 channels).
__global__ void logmel_floor_cl_kernel(op16* a, const uint32_t* keys, int64_t batch, const int32_t* lengths,
                                       int64_t n_samples, int hop, int M, int CP, int T, int FB, int tiles_per_utt) {
    const int b = blockIdx.y, tile = blockIdx.x;
    float fls;
    if (floor_tile_is_clear(keys, batch, tiles_per_utt, b, tile, fls)) return;
    const int Tb = min(1 + (int)(clamp_len(lengths, b, n_samples) / hop), T);
    const float fl = unpack_op16x2(pack_op16x2(fls, fls)).x;           // the floor, rounded like the values were
    const int cpr = (M + 7) >> 3;                          // 16-byte chunks per row that hold real channels
    const int t0 = tile * FB, nf = min(FB, Tb - t0);
    for (int i = threadIdx.x; i < nf * cpr; i += blockDim.x) {
        const int f = i / cpr, c8 = i - f * cpr;
        uint4* ptr = reinterpret_cast<uint4*>(a + ((int64_t)b * T + t0 + f) * CP + c8 * 8);
        uint4 q = *ptr;
        uint32_t* h = reinterpret_cast<uint32_t*>(&q);
        bool changed = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 v = unpack_op16x2(h[j]);
            if (c8 * 8 + 2 * j < M) v.x = fmaxf(v.x, fl);                  // padded channels stay 0
            if (c8 * 8 + 2 * j + 1 < M) v.y = fmaxf(v.y, fl);
            const uint32_t u = pack_op16x2(v.x, v.y);
            changed |= u != h[j];
            h[j] = u;
        }
        if (changed) *ptr = q;
    }
}

}  // namespace asrb

using namespace asrb;


extern "C" int asrb_logmel_plan_create(int n_fft, int hop, int n_mels, const float* window_host,
                                       const float* fbank_host, asrb_logmel_plan** plan) {
    if (!plan || !window_host || !fbank_host) return fail(ASRB_E_ARG, "asrb_logmel_plan_create: NULL argument");
    if (n_fft != 400 && n_fft != 1024)
        return fail(ASRB_E_ARG, "asrb_logmel_plan_create: n_fft=%d unsupported (400 or 1024)", n_fft);
    if (hop <= 0 || hop > n_fft || n_mels <= 0 || n_mels > 1024)
        return fail(ASRB_E_ARG, "asrb_logmel_plan_create: bad hop=%d / n_mels=%d", hop, n_mels);
    ASRB_TRY(require_sm100());
    const int R = n_fft == 400 ? 20 : 32;
    const int F = n_fft / 2 + 1;
    // dense [F][M] -> band per filter (triangles have contiguous support)
    std::vector<int> lo(n_mels, 0), cnt(n_mels, 0);
    int kmax = 1;
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int f = 0; f < F; ++f)
            if (fbank_host[(size_t)f * n_mels + m] != 0.0f) { if (first < 0) first = f; last = f; }
        if (first >= 0) { lo[m] = first; cnt[m] = last - first + 1; if (cnt[m] > kmax) kmax = cnt[m]; }
    }
    kmax = (kmax + 3) & ~3;                                // taps padded with zeros to a multiple of 4 (16-byte weight loads)
    std::vector<float> w((size_t)n_mels * kmax, 0.f);
    for (int m = 0; m < n_mels; ++m)
        for (int i = 0; i < cnt[m]; ++i)        // x 1/4 (exact): the kernel's power spectra are |2 X|^2
            w[(size_t)m * kmax + i] = 0.25f * fbank_host[(size_t)(lo[m] + i) * n_mels + m];
    std::vector<float2> tw((size_t)R * R);
    for (int k1 = 0; k1 < R; ++k1)
        for (int j = 0; j < R; ++j) {
            const double a = -2.0 * M_PI * (double)(k1 * j) / (double)n_fft;
            tw[(size_t)k1 * R + j] = make_float2((float)cos(a), (float)sin(a));
        }
    asrb_logmel_plan* pl = new asrb_logmel_plan{n_fft, hop, n_mels, kmax, R, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = up((void**)&pl->d_window, window_host, sizeof(float) * n_fft);
    if (e == cudaSuccess) e = up((void**)&pl->d_twiddle, tw.data(), sizeof(float2) * tw.size());
    if (e == cudaSuccess) e = up((void**)&pl->d_lo, lo.data(), sizeof(int) * n_mels);
    if (e == cudaSuccess) e = up((void**)&pl->d_cnt, cnt.data(), sizeof(int) * n_mels);
    if (e == cudaSuccess) e = up((void**)&pl->d_w, w.data(), sizeof(float) * w.size());
    if (e != cudaSuccess) {
        asrb_logmel_plan_destroy(pl);
        return fail(ASRB_E_CUDA, "asrb_logmel_plan_create: upload failed: %s", cudaGetErrorString(e));
    }
    *plan = pl;
    return ASRB_OK;
}

extern "C" void asrb_logmel_plan_destroy(asrb_logmel_plan* pl) {
    if (!pl) return;
    cudaFree(pl->d_window); cudaFree(pl->d_twiddle); cudaFree(pl->d_lo); cudaFree(pl->d_cnt); cudaFree(pl->d_w);
    delete pl;
}

extern "C" int64_t asrb_logmel_num_frames(const asrb_logmel_plan* pl, int64_t n) {
    return pl && n >= 0 ? 1 + n / pl->hop : -1;
}

namespace asrb {
int logmel_tile_frames(const asrb_logmel_plan* pl) { return pl->n_fft == 400 ? 32 : 16; }
// keys: [batch] per-utterance maxima, then [batch][tiles] per-tile minima
size_t logmel_keys_words(const asrb_logmel_plan* pl, int64_t batch, int64_t n_samples) {
    const int64_t T = 1 + n_samples / pl->hop, FB = logmel_tile_frames(pl);
    return (size_t)(batch > 0 ? batch : 1) * (size_t)(1 + (T + FB - 1) / FB);
}
}  // namespace asrb

extern "C" size_t asrb_logmel_workspace_bytes(const asrb_logmel_plan* pl, int64_t batch, int64_t n_samples) {
    if (!pl || batch < 0 || n_samples < 0) return 0;
    return align_up(sizeof(uint32_t) * logmel_keys_words(pl, batch, n_samples), 256);
}

namespace asrb {

template <int NFFT, int R, int FB, int HOP>
static int launch_logmel(const asrb_logmel_plan* pl, LogmelParams p, int64_t batch, cudaStream_t st) {
    using C = LogmelCfg<NFFT, R, FB, HOP>;
    const int span = (FB - 1) * pl->hop + NFFT;
    const size_t pcm_words = ((span + 3) & ~3) + C::GAP * C::QUADS + 4;
    size_t a_bytes = ((pcm_words * 4 + 15) & ~(size_t)15) + (p.out_cl ? ((sizeof(op16) * FB * (p.CP + 4) + 15) & ~(size_t)15) : 0);
    if (a_bytes < (size_t)C::E_BYTES) a_bytes = C::E_BYTES;
    p.a_bytes = (int)a_bytes;
    size_t smem = C::region((int)a_bytes) + sizeof(float) * NFFT + sizeof(float2) * (R * R) +
                  sizeof(float) * (size_t)pl->n_mels * pl->kmax + sizeof(int2) * pl->n_mels;
    if (smem > 227 * 1024) return fail(ASRB_E_ARG, "asrb_logmel_f32: hop/n_mels/channel pitch need %zu B of shared memory", smem);
    auto kern = logmel_kernel<NFFT, R, FB, HOP>;
    ASRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    ASRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int tiles_per_utt = (p.T + FB - 1) / FB;
    const int64_t total = (int64_t)tiles_per_utt * batch;
    if (total > 0x7fffffff) return fail(ASRB_E_ARG, "asrb_logmel_f32: too many frame tiles");
    int grid = sm_count() * per_sm;
    if (grid > total) grid = (int)total;
    kern<<<grid, C::THREADS, smem, st>>>(p, tiles_per_utt, (int)total);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// Shared by asrb_logmel_f32 and the fused pcm->hidden path: pass 1 only (values + keys).
int logmel_pass1(const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                 int64_t stride, const int32_t* lengths, float* out, uint32_t* keys, cudaStream_t st,
                 op16* out_cl, int CP, float* pool_out, int64_t pool_target) {
    LogmelParams p;
    p.out_cl = out_cl; p.CP = CP;
    if (out_cl && (CP < pl->n_mels || (CP & 7))) return fail(ASRB_E_ARG, "log-mel: channels-last pitch %d for %d mels", CP, pl->n_mels);
    p.pcm = pcm; p.stride = stride; p.n_samples = n_samples; p.lengths = lengths;
    p.hop = pl->hop; p.T = (int)(1 + n_samples / pl->hop); p.M = pl->n_mels; p.kmax = pl->kmax;
    p.window = pl->d_window; p.twiddle = pl->d_twiddle; p.mel_lo = pl->d_lo; p.mel_cnt = pl->d_cnt;
    p.mel_w = pl->d_w; p.out = out; p.keys = keys; p.tile_min = keys + batch;
    p.pool_out = pool_target > 0 ? pool_out : nullptr; p.pool_target = pool_target;
    ASRB_CUDA(cudaMemsetAsync(keys, 0, sizeof(uint32_t) * batch, st));                       // maxima: below every key
    {                                                                                        // tile minima: above every key (atomicMin)
        const int FB = logmel_tile_frames(pl);
        ASRB_CUDA(cudaMemsetAsync(keys + batch, 0xff, sizeof(uint32_t) * (size_t)batch * (size_t)((p.T + FB - 1) / FB), st));
    }
    const double frames = (double)batch * p.T;
    ProfScope ps("logmel_stft_mel", st, frames * 2.5 * pl->n_fft * log2((double)pl->n_fft),
                 4.0 * batch * ((double)n_samples + (double)pl->n_mels * p.T));
    if (pl->n_fft == 400) return pl->hop == 160 ? launch_logmel<400, 20, 32, 160>(pl, p, batch, st) : launch_logmel<400, 20, 32, 0>(pl, p, batch, st);
    return pl->hop == 160 ? launch_logmel<1024, 32, 16, 160>(pl, p, batch, st) : launch_logmel<1024, 32, 16, 0>(pl, p, batch, st);
}

// Floor of the fused path (after logmel_pass1 with out_cl).
int logmel_floor_cl(const asrb_logmel_plan* pl, op16* a, int CP, const uint32_t* keys, const int32_t* lengths,
                    int64_t batch, int64_t n_samples, cudaStream_t st) {
    const int T = (int)(1 + n_samples / pl->hop), FB = logmel_tile_frames(pl), tiles = (T + FB - 1) / FB;
    ProfScope ps("logmel_floor", st, 0.0, 4.0 * batch * tiles);
    logmel_floor_cl_kernel<<<dim3((unsigned)tiles, (unsigned)batch), 128, 0, st>>>(a, keys, batch, lengths, n_samples, pl->hop, pl->n_mels, CP, T, FB, tiles);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

}  // namespace asrb

static int logmel_f32_impl(const char* who, const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                           int64_t pcm_stride, const int32_t* lengths, float* out, float* pool_out, int64_t pool_target,
                           void* ws, size_t ws_bytes, void* stream) {
    if (!pl) return fail(ASRB_E_ARG, "%s: NULL plan", who);
    if (batch < 0 || n_samples < 0 || pcm_stride < n_samples)
        return fail(ASRB_E_ARG, "%s: bad shape batch=%lld n=%lld stride=%lld", who,
                    (long long)batch, (long long)n_samples, (long long)pcm_stride);
    if (batch == 0) return ASRB_OK;
    if (batch > 65535) return fail(ASRB_E_ARG, "%s: batch %lld > 65535", who, (long long)batch);
    if (!out || (!pcm && n_samples > 0)) return fail(ASRB_E_ARG, "%s: NULL tensor", who);
    if (1 + n_samples / pl->hop > 0x7fffffff / 2) return fail(ASRB_E_ARG, "%s: too many frames", who);
    if (pool_target > 0 && (!pool_out || pool_target >= n_samples || pool_target > 1 + n_samples / pl->hop))
        return fail(ASRB_E_ARG, "%s: pooled target %lld needs an output, target < samples and target <= frames", who, (long long)pool_target);
    if (!ws || ws_bytes < asrb_logmel_workspace_bytes(pl, batch, n_samples) || ((uintptr_t)ws & 3))
        return fail(ASRB_E_WORKSPACE, "%s: workspace too small or NULL (%zu B given)", who, ws_bytes);
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys = (uint32_t*)ws;
    ASRB_TRY(logmel_pass1(pl, pcm, batch, n_samples, pcm_stride, lengths, out, keys, st, nullptr, 0, pool_out, pool_target));
    const int T = (int)(1 + n_samples / pl->hop), FB = logmel_tile_frames(pl), tiles = (T + FB - 1) / FB;
    ProfScope ps("logmel_floor", st, 0.0, 4.0 * batch * tiles);
    logmel_floor_kernel<<<dim3((unsigned)tiles, (unsigned)batch), 256, 0, st>>>(out, keys, batch, lengths, n_samples, pl->hop, pl->n_mels, T, FB, tiles);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

extern "C" int asrb_logmel_f32(const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                               int64_t pcm_stride, const int32_t* lengths, float* out,
                               void* ws, size_t ws_bytes, void* stream) {
    return logmel_f32_impl("asrb_logmel_f32", pl, pcm, batch, n_samples, pcm_stride, lengths, out, nullptr, 0, ws, ws_bytes, stream);
}

extern "C" int asrb_logmel_waveform_f32(const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                                        int64_t pcm_stride, const int32_t* lengths, float* out, float* pool_out,
                                        int64_t pool_target, void* ws, size_t ws_bytes, void* stream) {
    return logmel_f32_impl("asrb_logmel_waveform_f32", pl, pcm, batch, n_samples, pcm_stride, lengths, out, pool_out, pool_target,
                           ws, ws_bytes, stream);
}

// ------------------------------------------------------------------------------------------
// "waveform" feature: the reference resamples the PCM to the frame rate
// with F.adaptive_avg_pool1d(audio, target) (essentials.py:493-503); output i is the mean of
// x[floor(i N / target) : ceil((i + 1) N / target)).  One warp per output, coalesced reads.
// ------------------------------------------------------------------------------------------
namespace asrb {
__global__ void waveform_pool_kernel(const float* __restrict__ pcm, int64_t stride, int64_t n, int64_t target,
                                     float* __restrict__ out, int64_t total) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= total) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = w / target, i = w - b * target;
    const int64_t s = (i * n) / target;
    const int64_t e = ((i + 1) * n + target - 1) / target;
    const float* x = pcm + b * stride;
    float acc = 0.f;
    for (int64_t k = s + lane; k < e; k += 32) acc += x[k];
    acc = warp_sum(acc);
    if (lane == 0) out[w] = acc / (float)(e - s);
}
// target >= n: F.interpolate(mode="linear", align_corners=False) (essentials.py:505-506): source position
// (i + 0.5) n / target - 0.5 clamped at 0, two taps, weights (1 - l, l) -- the arithmetic of ATen's upsample_linear1d.
__global__ void waveform_interp_kernel(const float* __restrict__ pcm, int64_t stride, int64_t n, int64_t target,
                                       float* __restrict__ out, int64_t total) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total) return;
    const int64_t b = w / target, i = w - b * target;
    const float scale = (float)n / (float)target;
    float src = scale * ((float)i + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    const int64_t i0 = (int64_t)src;
    const int64_t i1 = i0 + (i0 < n - 1 ? 1 : 0);
    const float l1 = src - (float)i0, l0 = 1.0f - l1;
    const float* x = pcm + b * stride;
    out[w] = l0 * x[i0] + l1 * x[i1];
}
}  // namespace asrb

extern "C" int asrb_waveform_pool_f32(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                                      int64_t target, float* out, void* stream) {
    if (batch < 0 || n_samples < 0 || pcm_stride < n_samples || target < 0)
        return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: bad shape");
    if (batch == 0 || target == 0) return ASRB_OK;
    if (!pcm || !out) return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: NULL tensor");
    if (n_samples == 0) return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: empty input");
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = batch * target;
    if (target >= n_samples) {                              // the reference's `else` branch: linear interpolation
        ProfScope ps("waveform_interp", st, 0.0, 4.0 * batch * ((double)n_samples + target));
        waveform_interp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pcm, pcm_stride, n_samples, target, out, total);
        ASRB_LAUNCH_CHECK();
        return ASRB_OK;
    }
    ProfScope ps("waveform_pool", st, 0.0, 4.0 * batch * ((double)n_samples + target));
    const int64_t blocks = (total * 32 + 255) / 256;
    if (blocks > 0x7fffffff) return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: too large");
    waveform_pool_kernel<<<(unsigned)blocks, 256, 0, st>>>(pcm, pcm_stride, n_samples, target, out, total);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}
