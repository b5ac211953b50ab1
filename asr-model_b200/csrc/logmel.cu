// Fused log-mel front end (essentials.py:469-491 batched): framing + Hann window +
// real FFT + |X|^2 + HTK mel projection + log10/clamp/normalise in ONE pass over the PCM.
//
// Data layout in HBM:  pcm [B][stride] fp32 (read once, staged in shared memory where the
// 2.5x / 6.4x frame overlap lives);  out [B][M][T] fp32 (written once);  keys [B] uint32 --
// order-preserving image of each utterance's running max of log10(mel).
//
// One block = FB consecutive frames of one utterance.  Two real frames ride one complex
// FFT (z = a + i b); an N = R*R point FFT is two rounds of R-point DFTs held in registers
// (R threads per FFT, R = 20 for n_fft 400, 32 for n_fft 1024) with one transpose through
// shared memory in between.  The per-utterance dynamic-range floor (max - 8) needs the max
// of the WHOLE utterance, so this pass writes (log10 + 4) / 4 and publishes the max with
// one atomicMax per block; logmel_floor_kernel then raises the values below the floor
// (monotone, so max((x+4)/4, (floor+4)/4) == (max(x, floor)+4)/4 bit for bit) and only
// writes where something changes.
#include "logmel.cuh"
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
    const float* mel_w;        // [M][kmax]
    float* out;                // [B][M][T] fp32 (reference layout), or NULL with out_cl set
    op16* out_cl;              // fused path: [B][T][CP] op16 channels-last, channels >= M zero (the stem GEMM's operand)
    int CP;
    uint32_t* keys;            // [B]
};

template <int NFFT, int R, int FB>
struct LogmelCfg {
    static constexpr int N = NFFT;
    static constexpr int GROUPS = FB / 2;              // complex FFTs per block
    static constexpr int THREADS = GROUPS * R;
    static constexpr int NB = NFFT / 2 + 1;            // one-sided bins
    static constexpr int YSTRIDE = R * (R + 1);        // float2 per group (padded transpose)
    static constexpr int BINS_PER_THREAD = (NB + R - 1) / R;
    static_assert(R * R == NFFT, "two-round FFT needs n_fft = R^2");
    static_assert(FB * NB <= GROUPS * YSTRIDE * 2, "power spectra must fit over the FFT buffer");
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

// Persistent: gridDim.x blocks walk the (utterance, frame-tile) list; constants are staged once
// per block and the PCM span of the NEXT tile streams in with cp.async (zero fill outside
// [0, len) = the center=True padding) while the current tile is transformed.
template <int NFFT, int R, int FB>
__global__ void __launch_bounds__(LogmelCfg<NFFT, R, FB>::THREADS)
logmel_kernel(const LogmelParams p, int tiles_per_utt, int total_tiles) {
    using C = LogmelCfg<NFFT, R, FB>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int span = (FB - 1) * p.hop + NFFT;
    const int span4 = (span + 3) & ~3;
    float*  s_pcm0 = reinterpret_cast<float*>(smem_raw);                // [2][span4]
    float*  s_win = s_pcm0 + 2 * span4;                                 // [NFFT]
    float2* s_tw  = reinterpret_cast<float2*>(s_win + NFFT);            // [R*R]
    float2* s_y   = s_tw + R * R;                                       // [GROUPS][YSTRIDE]
    float*  s_pow = reinterpret_cast<float*>(s_y);                      // aliases s_y: [FB][NB]
    float*  s_melw = reinterpret_cast<float*>(s_y + C::GROUPS * C::YSTRIDE);   // [M][kmax]
    int*    s_lo  = reinterpret_cast<int*>(s_melw + p.M * p.kmax);      // [M]
    int*    s_cnt = s_lo + p.M;                                         // [M]
    // fused path: the tile's [FB][CP] 16-bit output is transposed through shared memory (pitch CP + 2 halves:
    // frame-per-lane writes and row reads are both conflict free)
    op16* s_cl = reinterpret_cast<op16*>(s_cnt + p.M + (p.M & 1));
    const int clp = p.CP + 2;
    __shared__ float s_red[32];

    const int tid = threadIdx.x;
    const bool vec_ok = ((p.stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0) && ((p.hop & 3) == 0);

    auto prefetch = [&](int tile, float* dstbuf) {       // async copy of one tile's PCM span
        const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * FB;
        const int64_t len = clamp_len(p.lengths, b, p.n_samples);
        const float* src = p.pcm + (int64_t)b * p.stride;
        const int64_t s0 = (int64_t)t0 * p.hop - NFFT / 2;
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dstbuf);
        if (vec_ok) {                                     // s0 is a multiple of 4 here
            for (int i = tid * 4; i < span4; i += C::THREADS * 4) {
                const int64_t sidx = s0 + i;
                int64_t valid = len - sidx;               // samples available from sidx
                valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
                const bool in = sidx >= 0 && valid > 0;   // sidx < 0: whole chunk is padding (sidx multiple of 4)
                cp_async16(d0 + 4u * i, in ? (const void*)(src + sidx) : (const void*)src, in ? (int)valid * 4 : 0);
            }
        } else {
            for (int i = tid; i < span4; i += C::THREADS) {
                const int64_t sidx = s0 + i;
                const bool in = sidx >= 0 && sidx < len;
                cp_async4(d0 + 4u * i, in ? (const void*)(src + sidx) : (const void*)p.pcm, in ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    int tile = blockIdx.x;
    if (tile < total_tiles) prefetch(tile, s_pcm0);
    // ---- constants, once per block ----
    for (int i = tid; i < NFFT; i += C::THREADS) s_win[i] = p.window[i];
    for (int i = tid; i < R * R; i += C::THREADS) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < p.M * p.kmax; i += C::THREADS) s_melw[i] = p.mel_w[i];
    for (int i = tid; i < p.M; i += C::THREADS) { s_lo[i] = p.mel_lo[i]; s_cnt[i] = p.mel_cnt[i]; }
    if (p.out_cl) for (int i = tid; i < FB * clp; i += C::THREADS) s_cl[i] = to_op16(0.f);   // channels >= M stay 0

    const int g = tid / R, j = tid - g * R;
    float2* yg = s_y + g * C::YSTRIDE;
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: filter loops run on the uniform datapath
    constexpr int nwarp = C::THREADS / 32;
    constexpr int FW = FB < 32 ? FB : 32;                   // frames handled by one warp pass
    constexpr int MSUB = 32 / FW;                           // filters handled side by side

    for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
        float* s_pcm = s_pcm0 + (it & 1) * span4;
        const int next = tile + gridDim.x;
        if (next < total_tiles) { prefetch(next, s_pcm0 + ((it + 1) & 1) * span4); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                                    // this tile's PCM (and the constants) are visible

        const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * FB;
        const int64_t len = clamp_len(p.lengths, b, p.n_samples);
        const int Tb = 1 + (int)(len / p.hop);              // valid frames of this utterance (<= T: len <= n_samples)

        // ---- round 1: R-point DFT over n1 of z[R n1 + j], twiddle W_N^(j k1), transpose ----
        {
            float2 x[R];
            const float* fa = s_pcm + (2 * g) * p.hop;
            const float* fb = fa + p.hop;
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1) {
                const int n = R * n1 + j;
                const float w = s_win[n];
                x[n1] = make_float2(w * fa[n], w * fb[n]);
            }
            SmallDFT<R>::run(x);
            yg[j] = x[0];
#pragma unroll
            for (int k1 = 1; k1 < R; ++k1) yg[k1 * (R + 1) + j] = cmul(x[k1], s_tw[k1 * R + j]);
        }
        __syncthreads();

        // ---- round 2: thread k1 = j does the R-point DFT over n2 -> Z[j + R k2] ----
        {
            float2 y[R];
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2) y[n2] = yg[j * (R + 1) + n2];
            SmallDFT<R>::run(y);
            __syncthreads();                                // everyone has read the transpose
#pragma unroll
            for (int k2 = 0; k2 < R; ++k2) yg[j + R * k2] = y[k2];
        }
        __syncthreads();

        // ---- split the packed pair: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i ----
        float pa[C::BINS_PER_THREAD], pb[C::BINS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < C::BINS_PER_THREAD; ++i) {
            const int k = j + R * i;
            pa[i] = pb[i] = 0.f;
            if (k < C::NB) {
                const float2 z1 = yg[k];
                const float2 z2 = yg[k == 0 ? 0 : NFFT - k];
                const float ar = z1.x + z2.x, ai = z1.y - z2.y;     // 2 Re A, 2 Im A
                const float br = z1.y + z2.y, bi = z2.x - z1.x;     // 2 Re B, 2 Im B
                pa[i] = 0.25f * fmaf(ar, ar, ai * ai);
                pb[i] = 0.25f * fmaf(br, br, bi * bi);
            }
        }
        __syncthreads();                                    // Z is dead: reuse it for |X|^2
#pragma unroll
        for (int i = 0; i < C::BINS_PER_THREAD; ++i) {
            const int k = j + R * i;
            if (k < C::NB) {
                s_pow[(2 * g) * C::NB + k] = pa[i];
                s_pow[(2 * g + 1) * C::NB + k] = pb[i];
            }
        }
        __syncthreads();

        // ---- banded mel projection + log10 + (x+4)/4; lane -> frame so stores are coalesced ----
        const int f = lane % FW;
        float vmax = -INFINITY;
        for (int fbase = 0; fbase < FB; fbase += FW) {
            const int fr = fbase + f;
            const int t = t0 + fr;
            const float* pw = s_pow + fr * C::NB;
            for (int m = warp * MSUB + lane / FW; m < p.M; m += nwarp * MSUB) {
                // banded filter, taps zero-padded to a multiple of 4: one broadcast 16-byte weight load per 4 taps
                const int lo = s_lo[m], n4 = (s_cnt[m] + 3) >> 2;
                const float4* w4 = reinterpret_cast<const float4*>(s_melw + m * p.kmax);
                const float* px = pw + lo;
                float acc = 0.f;
#pragma unroll 1                                               // 1-3 trips: remainder code of an unrolled loop costs more than it saves
                for (int q = 0; q < n4; ++q) {
                    const float4 w = w4[q];
                    acc = fmaf(w.x, px[4 * q], acc);
                    acc = fmaf(w.y, px[4 * q + 1], acc);
                    acc = fmaf(w.z, px[4 * q + 2], acc);
                    acc = fmaf(w.w, px[4 * q + 3], acc);
                }
                // log10 through MUFU lg2 (abs error ~2^-22 in log2: 7e-8 in log10)   essentials.py:488
                float lg;                                                   // argument >= 1e-10: never denormal, plain MUFU.LG2
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(acc, 1e-10f)));
                lg *= 0.30102999566398120f;
                if (t < p.T) {
                    float sv = 0.f;                                         // DataCollator pad value
                    if (t < Tb) { vmax = fmaxf(vmax, lg); sv = (lg + 4.0f) / 4.0f; }   // essentials.py:490
                    if (p.out_cl) s_cl[fr * clp + m] = to_op16(sv);
                    else p.out[((int64_t)b * p.M + m) * p.T + t] = sv;
                }
            }
        }
        vmax = warp_max(vmax);
        if (lane == 0) s_red[warp] = vmax;
        __syncthreads();                                    // also: s_pow / s_pcm are free for the next tile
        if (p.out_cl) {                                     // rows of CP 16-bit values, two per 32-bit word: coalesced
            const int wpr = p.CP >> 1;
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.out_cl + ((int64_t)b * p.T + t0) * p.CP);
            for (int i = tid; i < FB * wpr; i += C::THREADS) {
                const int fr = i / wpr, w = i - fr * wpr;
                if (t0 + fr < p.T) dst[i] = *reinterpret_cast<const uint32_t*>(s_cl + fr * clp + 2 * w);
            }
        }
        if (warp == 0) {
            float v = lane < nwarp ? s_red[lane] : -INFINITY;
            v = warp_max(v);
            if (lane == 0 && v > -INFINITY) atomicMax(p.keys + b, f2key(v));
        }
    }
}

// essentials.py:489: log_mel = maximum(log_mel, log_mel.max() - 8.0), applied in the
// normalised domain.  Touches memory only where the floor is active.
__global__ void logmel_floor_kernel(float* out, const uint32_t* keys, const int32_t* lengths,
                                    int64_t n_samples, int hop, int M, int T) {
    const int b = blockIdx.y;
    const int64_t len = clamp_len(lengths, b, n_samples);
    const int Tb = 1 + (int)(len / hop);
    const float floor_s = ((key2f(keys[b]) - 8.0f) + 4.0f) / 4.0f;
    float* o = out + (int64_t)b * M * T;
    const int64_t total = (int64_t)M * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        if (t < Tb) {
            const float v = o[i];
            if (v < floor_s) o[i] = floor_s;
        }
    }
}

// The same floor on the fused path's 16-bit channels-last tensor, in place: rounding is monotone, so
// max(rn(x), rn(floor)) == rn(max(x, floor)) bit for bit.  One thread per (frame, 8 channels).
__global__ void logmel_floor_cl_kernel(op16* a, const uint32_t* keys, const int32_t* lengths,
                                       int64_t n_samples, int hop, int M, int CP, int T) {
    const int b = blockIdx.y;
    const int64_t len = clamp_len(lengths, b, n_samples);
    const int Tb = min(1 + (int)(len / hop), T);
    const float fls = ((key2f(keys[b]) - 8.0f) + 4.0f) / 4.0f;
    const float fl = unpack_op16x2(pack_op16x2(fls, fls)).x;           // the floor, rounded like the values were
    const int cpr = (M + 7) >> 3;                          // 16-byte chunks per row that hold real channels
    const int64_t total = (int64_t)Tb * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / cpr), c8 = (int)(i - (int64_t)t * cpr);
        uint4* ptr = reinterpret_cast<uint4*>(a + ((int64_t)b * T + t) * CP + c8 * 8);
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
        for (int i = 0; i < cnt[m]; ++i) w[(size_t)m * kmax + i] = fbank_host[(size_t)(lo[m] + i) * n_mels + m];
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

extern "C" size_t asrb_logmel_workspace_bytes(const asrb_logmel_plan* pl, int64_t batch, int64_t) {
    if (!pl || batch < 0) return 0;
    return align_up(sizeof(uint32_t) * (size_t)(batch > 0 ? batch : 1), 256);
}

namespace asrb {

template <int NFFT, int R, int FB>
static int launch_logmel(const asrb_logmel_plan* pl, LogmelParams p, int64_t batch, cudaStream_t st) {
    using C = LogmelCfg<NFFT, R, FB>;
    const int span = (FB - 1) * pl->hop + NFFT;
    size_t smem = sizeof(float) * (2 * ((span + 3) & ~3) + NFFT) + sizeof(float2) * (R * R + C::GROUPS * C::YSTRIDE) +
                  sizeof(float) * (size_t)pl->n_mels * pl->kmax + sizeof(int) * (2 * pl->n_mels + 1) +
                  (p.out_cl ? sizeof(op16) * FB * (p.CP + 2) : 0);
    if (smem > 227 * 1024) return fail(ASRB_E_ARG, "asrb_logmel_f32: hop/n_mels need %zu B of shared memory", smem);
    auto kern = logmel_kernel<NFFT, R, FB>;
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
                 op16* out_cl, int CP) {
    LogmelParams p;
    p.out_cl = out_cl; p.CP = CP;
    if (out_cl && (CP < pl->n_mels || (CP & 7))) return fail(ASRB_E_ARG, "log-mel: channels-last pitch %d for %d mels", CP, pl->n_mels);
    p.pcm = pcm; p.stride = stride; p.n_samples = n_samples; p.lengths = lengths;
    p.hop = pl->hop; p.T = (int)(1 + n_samples / pl->hop); p.M = pl->n_mels; p.kmax = pl->kmax;
    p.window = pl->d_window; p.twiddle = pl->d_twiddle; p.mel_lo = pl->d_lo; p.mel_cnt = pl->d_cnt;
    p.mel_w = pl->d_w; p.out = out; p.keys = keys;
    ASRB_CUDA(cudaMemsetAsync(keys, 0, sizeof(uint32_t) * batch, st));
    const double frames = (double)batch * p.T;
    ProfScope ps("logmel_stft_mel", st, frames * 2.5 * pl->n_fft * log2((double)pl->n_fft),
                 4.0 * batch * ((double)n_samples + (double)pl->n_mels * p.T));
    if (pl->n_fft == 400) return launch_logmel<400, 20, 32>(pl, p, batch, st);
    return launch_logmel<1024, 32, 16>(pl, p, batch, st);
}

// Floor of the fused path (after logmel_pass1 with out_cl).
int logmel_floor_cl(const asrb_logmel_plan* pl, op16* a, int CP, const uint32_t* keys, const int32_t* lengths,
                    int64_t batch, int64_t n_samples, cudaStream_t st) {
    const int T = (int)(1 + n_samples / pl->hop);
    const int64_t per = (int64_t)T * ((pl->n_mels + 7) / 8);
    int gx = (int)((per + 256 * 4 - 1) / (256 * 4));
    if (gx < 1) gx = 1;
    ProfScope ps("logmel_floor", st, 0.0, 4.0 * batch * T * pl->n_mels);
    logmel_floor_cl_kernel<<<dim3(gx, (unsigned)batch), 256, 0, st>>>(a, keys, lengths, n_samples, pl->hop, pl->n_mels, CP, T);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

}  // namespace asrb

extern "C" int asrb_logmel_f32(const asrb_logmel_plan* pl, const float* pcm, int64_t batch, int64_t n_samples,
                               int64_t pcm_stride, const int32_t* lengths, float* out,
                               void* ws, size_t ws_bytes, void* stream) {
    if (!pl) return fail(ASRB_E_ARG, "asrb_logmel_f32: NULL plan");
    if (batch < 0 || n_samples < 0 || pcm_stride < n_samples)
        return fail(ASRB_E_ARG, "asrb_logmel_f32: bad shape batch=%lld n=%lld stride=%lld",
                    (long long)batch, (long long)n_samples, (long long)pcm_stride);
    if (batch == 0) return ASRB_OK;
    if (batch > 65535) return fail(ASRB_E_ARG, "asrb_logmel_f32: batch %lld > 65535", (long long)batch);
    if (!out || (!pcm && n_samples > 0)) return fail(ASRB_E_ARG, "asrb_logmel_f32: NULL tensor");
    if (1 + n_samples / pl->hop > 0x7fffffff / 2) return fail(ASRB_E_ARG, "asrb_logmel_f32: too many frames");
    if (!ws || ws_bytes < asrb_logmel_workspace_bytes(pl, batch, n_samples) || ((uintptr_t)ws & 3))
        return fail(ASRB_E_WORKSPACE, "asrb_logmel_f32: workspace too small or NULL (%zu B given)", ws_bytes);
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys = (uint32_t*)ws;
    ASRB_TRY(logmel_pass1(pl, pcm, batch, n_samples, pcm_stride, lengths, out, keys, st));
    const int T = (int)(1 + n_samples / pl->hop);
    const int64_t per = (int64_t)pl->n_mels * T;
    int gx = (int)((per + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    ProfScope ps("logmel_floor", st, 0.0, 0.0);
    logmel_floor_kernel<<<dim3(gx, (unsigned)batch), 256, 0, st>>>(out, keys, lengths, n_samples, pl->hop, pl->n_mels, T);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
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
}  // namespace asrb

extern "C" int asrb_waveform_pool_f32(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_stride,
                                      int64_t target, float* out, void* stream) {
    if (batch < 0 || n_samples < 0 || pcm_stride < n_samples || target < 0)
        return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: bad shape");
    if (batch == 0 || target == 0) return ASRB_OK;
    if (!pcm || !out) return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: NULL tensor");
    if (target >= n_samples)
        return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: target %lld >= %lld samples (the reference interpolates there; unsupported)",
                    (long long)target, (long long)n_samples);
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = batch * target;
    ProfScope ps("waveform_pool", st, 0.0, 4.0 * batch * ((double)n_samples + target));
    const int64_t blocks = (total * 32 + 255) / 256;
    if (blocks > 0x7fffffff) return fail(ASRB_E_ARG, "asrb_waveform_pool_f32: too large");
    waveform_pool_kernel<<<(unsigned)blocks, 256, 0, st>>>(pcm, pcm_stride, n_samples, target, out, total);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}
