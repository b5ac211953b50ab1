// AudioEncoder handle: weight folding/packing at create(), forward orchestration.
// Mirrors model.py:120-169 (norm=False).  Two compute variants share the op sequence:
//   ASRB_BF16  tcgen05 GEMMs (gemm_tc.cu) with fused epilogues; 16-bit activations in HBM (op16 = IEEE fp16 MMA
//              operands, common.cuh), hidden states returned as bf16
//   ASRB_F32   FFMA GEMMs and fp32 activations (the <= 1e-4 variant)
#include "enc_kernels.cuh"
#include "logmel.cuh"
#include <map>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

using namespace asrb;

namespace {

struct HostTensors {
    std::map<std::string, std::pair<const float*, int64_t>> t;
    const float* get(const std::string& k, int64_t numel, int* err) const {
        auto it = t.find(k);
        if (it == t.end()) { *err = fail(ASRB_E_WEIGHTS, "state_dict tensor '%s' is missing", k.c_str()); return nullptr; }
        if (it->second.second != numel) {
            *err = fail(ASRB_E_WEIGHTS, "state_dict tensor '%s' has %lld elements, expected %lld", k.c_str(),
                        (long long)it->second.second, (long long)numel);
            return nullptr;
        }
        return it->second.first;
    }
    bool has(const std::string& k) const { return t.count(k) != 0; }
};

struct DevPool {                      // device constants owned by a handle
    std::vector<void*> ptrs;
    template <class T> int upload(const std::vector<T>& h, T** d) {
        void* p = nullptr;
        ASRB_CUDA(cudaMalloc(&p, h.size() * sizeof(T) + 16));
        ptrs.push_back(p);
        ASRB_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        *d = (T*)p;
        return ASRB_OK;
    }
    void release() { for (void* p : ptrs) cudaFree(p); ptrs.clear(); }
};

std::vector<op16> to_op16_vec(const std::vector<float>& v) {          // fp32 weights -> the 16-bit operand format
    std::vector<op16> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = host_to_op16(v[i]);
    return o;
}

struct LayerW {
    // k3 weight-normed conv: [D][3][D] (tap-major K); LN; ConvLite; depthwise k3
    float* wc_f = nullptr; op16* wc_h = nullptr; float* bc = nullptr;
    float* gamma = nullptr; float* beta = nullptr;
    float* w1_f = nullptr; op16* w1_h = nullptr; float* b1 = nullptr; float* b1_glu = nullptr;
    float* dw15 = nullptr; float* dw15_b = nullptr;          // [15][D], BatchNorm folded
    float* w2_f = nullptr; op16* w2_h = nullptr; float* b2 = nullptr;
    float* dw3 = nullptr; float* dw3_b = nullptr;            // [3][D]
};

}  // namespace

struct asrb_encoder {
    asrb_encoder_config cfg;
    int CP;                                                   // conv1 input channels padded to 64
    DevPool pool;
    float* stem1_f = nullptr; op16* stem1_h = nullptr; float* stem1_b = nullptr;
    float* stem2_f = nullptr; float* stem2_b = nullptr;
    std::vector<LayerW> layers;
    float* pos_scales = nullptr;
    // TransformerEncoderLayer
    float* win_f = nullptr; op16* win_h = nullptr; float* bin = nullptr;
    float* wo_f = nullptr; op16* wo_h = nullptr; float* bo = nullptr;
    float* wf1_f = nullptr; op16* wf1_h = nullptr; float* bf1 = nullptr;
    float* wf2_f = nullptr; op16* wf2_h = nullptr; float* bf2 = nullptr;
    float *n1g = nullptr, *n1b = nullptr, *n2g = nullptr, *n2b = nullptr;
};

// [N][C][taps] (PyTorch Conv1d) -> [N][taps][CPAD], zero-padded channels
static std::vector<float> pack_conv(const float* w, int N, int C, int taps, int CPAD) {
    std::vector<float> o((size_t)N * taps * CPAD, 0.f);
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < taps; ++j) o[((size_t)n * taps + j) * CPAD + c] = w[((size_t)n * C + c) * taps + j];
    return o;
}

extern "C" int asrb_encoder_create(const asrb_encoder_config* cfg, int n_tensors, const char* const* names,
                                   const float* const* host_data, const int64_t* numels, asrb_encoder** out) {
    if (!cfg || !out || (n_tensors > 0 && (!names || !host_data || !numels)))
        return fail(ASRB_E_ARG, "asrb_encoder_create: NULL argument");
    const int D = cfg->dims, M = cfg->mels, L = cfg->layer, F = cfg->ffn;
    if (D <= 0 || M <= 0 || L < 1 || (D % 64)) return fail(ASRB_E_ARG, "asrb_encoder_create: bad dims=%d (multiple of 64) mels=%d layer=%d (>= 1)", D, M, L);
    if (cfg->compute != ASRB_F32 && cfg->compute != ASRB_BF16) return fail(ASRB_E_ARG, "asrb_encoder_create: compute=%d", cfg->compute);
    if (D > 1024) return fail(ASRB_E_ARG, "asrb_encoder_create: dims=%d > 1024 unsupported", D);
    if (cfg->enc && (cfg->head <= 0 || D % cfg->head != 0 || F <= 0))
        return fail(ASRB_E_ARG, "asrb_encoder_create: bad head=%d / ffn=%d", cfg->head, F);
    if (cfg->enc) { const int hd = D / cfg->head; if (hd != 16 && hd != 32 && hd != 64 && hd != 128) return fail(ASRB_E_ARG, "asrb_encoder_create: head_dim %d unsupported", hd); }
    const bool bf = cfg->compute == ASRB_BF16;
    if (bf && (D % 128 != 0 || (cfg->enc && F % 128 != 0)))
        return fail(ASRB_E_ARG, "asrb_encoder_create: the tensor-core variant needs dims %% 128 == 0 (got %d)", D);
    ASRB_TRY(require_sm100());

    HostTensors ht;
    for (int i = 0; i < n_tensors; ++i) ht.t[names[i]] = {host_data[i], numels[i]};
    asrb_encoder* e = new asrb_encoder();
    e->cfg = *cfg;
    e->CP = (M + 63) / 64 * 64;
    int err = ASRB_OK;
#define GET(var, key, n) const float* var = ht.get(key, (int64_t)(n), &err); if (!var) { asrb_encoder_destroy(e); return err; }
#define UP(vec, dst) do { int r_ = e->pool.upload(vec, &(dst)); if (r_ != ASRB_OK) { asrb_encoder_destroy(e); return r_; } } while (0)
    auto vecf = [](const float* p, size_t n) { return std::vector<float>(p, p + n); };

    {   // stems (model.py:129-135)
        GET(w1, "conv1.0.weight", (int64_t)D * M * 3); GET(b1, "conv1.0.bias", D);
        if (bf) { auto p = to_op16_vec(pack_conv(w1, D, M, 3, e->CP)); UP(p, e->stem1_h); }
        else { auto p = pack_conv(w1, D, M, 3, M); UP(p, e->stem1_f); }
        { auto v = vecf(b1, D); UP(v, e->stem1_b); }
        if (ht.has("conv2.0.weight")) {
            GET(w2, "conv2.0.weight", (int64_t)D * 3); GET(b2, "conv2.0.bias", D);
            auto p = pack_conv(w2, D, 1, 3, 1); UP(p, e->stem2_f);
            auto v = vecf(b2, D); UP(v, e->stem2_b);
        }
    }
    {   // sinusoid scales s_j (essentials.py:355).  The binding passes the table it computed
        // with the reference's own torch ops ("__pos_scales"); expf is the fallback.
        std::vector<float> s(D / 2);
        if (ht.has("__pos_scales")) { GET(ps, "__pos_scales", D / 2); s.assign(ps, ps + D / 2); }
        else for (int j = 0; j < D / 2; ++j) s[j] = expf(-logf(30000.0f) / (float)(D / 2 - 1) * (float)j);
        UP(s, e->pos_scales);
    }
    e->layers.resize(L);
    for (int i = 0; i < L; ++i) {
        LayerW& lw = e->layers[i];
        const std::string p = "encoder." + std::to_string(i) + ".";
        GET(g, p + "1.parametrizations.weight.original0", D);
        GET(v, p + "1.parametrizations.weight.original1", (int64_t)D * D * 3);
        GET(bc, p + "1.bias", D);
        // weight_norm fold (model.py:143): W[n] = g[n] * v[n] / ||v[n]||_2 over (in, k)
        std::vector<float> w((size_t)D * D * 3);
        for (int n = 0; n < D; ++n) {
            double ss = 0; const float* vn = v + (size_t)n * D * 3;
            for (int k = 0; k < D * 3; ++k) ss += (double)vn[k] * vn[k];
            const float scale = (float)((double)g[n] / sqrt(ss));
            for (int k = 0; k < D * 3; ++k) w[(size_t)n * D * 3 + k] = vn[k] * scale;
        }
        { auto pk = pack_conv(w.data(), D, D, 3, D); if (bf) { auto h = to_op16_vec(pk); UP(h, lw.wc_h); } else UP(pk, lw.wc_f); }
        { auto t = vecf(bc, D); UP(t, lw.bc); }
        GET(gm, p + "2.gamma", D); GET(bt, p + "2.beta", D);
        { auto t = vecf(gm, D); UP(t, lw.gamma); } { auto t = vecf(bt, D); UP(t, lw.beta); }
        GET(w1, p + "3.point1.weight", (int64_t)2 * D * D); GET(b1, p + "3.point1.bias", 2 * D);
        if (bf) {
            // GLU interleave per 256-row tile: [128 value rows | 128 gate rows] for the same channels
            const int BN = tc_glu_tile_n(2 * D), H2 = BN / 2;
            std::vector<float> wi((size_t)2 * D * D), bi(2 * D);
            for (int tile = 0; tile < 2 * D / BN; ++tile)
                for (int r = 0; r < BN; ++r) {
                    const int src = r < H2 ? tile * H2 + r : D + tile * H2 + (r - H2);
                    memcpy(&wi[(size_t)(tile * BN + r) * D], w1 + (size_t)src * D, sizeof(float) * D);
                    bi[tile * BN + r] = b1[src];
                }
            auto h = to_op16_vec(wi); UP(h, lw.w1_h); UP(bi, lw.b1_glu);
        } else { auto t = vecf(w1, (size_t)2 * D * D); UP(t, lw.w1_f); auto tb = vecf(b1, 2 * D); UP(tb, lw.b1); }
        GET(dw, p + "3.depth.weight", (int64_t)D * 15); GET(db, p + "3.depth.bias", D);
        GET(bw, p + "3.bn.weight", D); GET(bb, p + "3.bn.bias", D);
        GET(rm, p + "3.bn.running_mean", D); GET(rv, p + "3.bn.running_var", D);
        {   // eval BatchNorm (model.py:103,114) folded into the depthwise conv
            std::vector<float> wf((size_t)15 * D), bfold(D);
            for (int c = 0; c < D; ++c) {
                const float s = bw[c] / sqrtf(rv[c] + 1e-5f);
                for (int j = 0; j < 15; ++j) wf[(size_t)j * D + c] = dw[(size_t)c * 15 + j] * s;
                bfold[c] = (db[c] - rm[c]) * s + bb[c];
            }
            UP(wf, lw.dw15); UP(bfold, lw.dw15_b);
        }
        GET(w2, p + "3.point2.weight", (int64_t)D * D); GET(b2, p + "3.point2.bias", D);
        { auto t = vecf(w2, (size_t)D * D); if (bf) { auto h = to_op16_vec(t); UP(h, lw.w2_h); } else UP(t, lw.w2_f); }
        { auto t = vecf(b2, D); UP(t, lw.b2); }
        GET(w5, p + "5.weight", (int64_t)D * 3); GET(b5, p + "5.bias", D);
        { std::vector<float> t((size_t)3 * D); for (int c = 0; c < D; ++c) for (int j = 0; j < 3; ++j) t[(size_t)j * D + c] = w5[(size_t)c * 3 + j]; UP(t, lw.dw3); }
        { auto t = vecf(b5, D); UP(t, lw.dw3_b); }
    }
    if (cfg->enc) {
        const std::string p = "EncoderLayer.";
        GET(wi, p + "self_attn.in_proj_weight", (int64_t)3 * D * D); GET(bi, p + "self_attn.in_proj_bias", 3 * D);
        GET(wo, p + "self_attn.out_proj.weight", (int64_t)D * D); GET(bo, p + "self_attn.out_proj.bias", D);
        GET(l1, p + "linear1.weight", (int64_t)F * D); GET(lb1, p + "linear1.bias", F);
        GET(l2, p + "linear2.weight", (int64_t)D * F); GET(lb2, p + "linear2.bias", D);
        GET(g1, p + "norm1.weight", D); GET(h1, p + "norm1.bias", D);
        GET(g2, p + "norm2.weight", D); GET(h2, p + "norm2.bias", D);
        auto put = [&](const float* w, size_t n, float** f, op16** h) -> int {
            auto t = vecf(w, n);
            if (bf) { auto hh = to_op16_vec(t); return e->pool.upload(hh, h); }
            return e->pool.upload(t, f);
        };
        int r = put(wi, (size_t)3 * D * D, &e->win_f, &e->win_h);
        if (r == ASRB_OK) r = put(wo, (size_t)D * D, &e->wo_f, &e->wo_h);
        if (r == ASRB_OK) r = put(l1, (size_t)F * D, &e->wf1_f, &e->wf1_h);
        if (r == ASRB_OK) r = put(l2, (size_t)D * F, &e->wf2_f, &e->wf2_h);
        if (r != ASRB_OK) { asrb_encoder_destroy(e); return r; }
        { auto t = vecf(bi, 3 * D); UP(t, e->bin); } { auto t = vecf(bo, D); UP(t, e->bo); }
        { auto t = vecf(lb1, F); UP(t, e->bf1); } { auto t = vecf(lb2, D); UP(t, e->bf2); }
        { auto t = vecf(g1, D); UP(t, e->n1g); } { auto t = vecf(h1, D); UP(t, e->n1b); }
        { auto t = vecf(g2, D); UP(t, e->n2g); } { auto t = vecf(h2, D); UP(t, e->n2b); }
    }
#undef GET
#undef UP
    *out = e;
    return ASRB_OK;
}

extern "C" void asrb_encoder_destroy(asrb_encoder* e) {
    if (!e) return;
    e->pool.release();
    delete e;
}

namespace {

struct EncBuffers {                    // carved from the caller's workspace
    void* a0; void* X; void* Y; void* G; void* U; void* H; void* wide; void* ffn; float* pos;
    int32_t* frames_valid; int32_t* frames_eff;
    bool ok;
};

size_t enc_ws_bytes(const asrb_encoder* e, int64_t B, int64_t T) {
    const size_t es = e->cfg.compute == ASRB_BF16 ? 2 : 4;
    const size_t D = e->cfg.dims;
    const size_t rows = (size_t)B * T;
    size_t n = 0;
    auto add = [&](size_t bytes) { n = align_up(n, 256) + bytes; };
    add(rows * (e->cfg.compute == ASRB_BF16 ? e->CP : e->cfg.mels) * es);   // a0
    add(rows * D * es); add(rows * D * es);                                 // X Y
    add(res32_rows(rows) * D * 4); add(rows * D * es); add(res32_rows(rows) * D * 4);   // G (fp32) U H (fp32): G, H feed depthwise convs and carry the blocked fp32 residual streams (rows rounded up to 32)
    const size_t wide = e->cfg.enc ? 3 * D : (e->cfg.compute == ASRB_F32 ? 2 * D : 0);
    add(rows * wide * es);                                                  // qkv | fp32 GLU input
    add(e->cfg.enc ? rows * e->cfg.ffn * es : 0);                           // FFN hidden
    add((size_t)T * D * 4);                                                 // sinusoid table
    add((size_t)B * 4); add((size_t)B * 4);                                 // ragged batches: valid / still-needed frames per utterance
    return align_up(n, 256) + 256;
}

EncBuffers carve(const asrb_encoder* e, int64_t B, int64_t T, void* ws, size_t ws_bytes) {
    const size_t es = e->cfg.compute == ASRB_BF16 ? 2 : 4;
    const size_t D = e->cfg.dims, rows = (size_t)B * T;
    Arena a(ws, ws_bytes);
    EncBuffers b;
    b.a0 = a.take<char>(rows * (e->cfg.compute == ASRB_BF16 ? e->CP : e->cfg.mels) * es);
    b.X = a.take<char>(rows * D * es); b.Y = a.take<char>(rows * D * es); b.G = a.take<char>(res32_rows(rows) * D * 4);
    b.U = a.take<char>(rows * D * es); b.H = a.take<char>(res32_rows(rows) * D * 4);
    const size_t wide = e->cfg.enc ? 3 * D : (e->cfg.compute == ASRB_F32 ? 2 * D : 0);
    b.wide = a.take<char>(rows * wide * es);
    b.ffn = a.take<char>(e->cfg.enc ? rows * e->cfg.ffn * es : 0);
    b.pos = a.take<float>((size_t)T * D);
    b.frames_valid = a.take<int32_t>((size_t)B); b.frames_eff = a.take<int32_t>((size_t)B);
    b.ok = a.ok();
    return b;
}

template <class TI, class TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        io<TO>::st(out + i, io<TI>::ld(in + i));
}

// Ragged batches (DataCollator pads to the longest clip, essentials.py:555-572).  valid[b] = frames of
// utterance b that hold audio; eff[b] = frames some valid output frame still depends on: every conv block widens the cone
// by 1 (k3) + 7 (depthwise-15) + 1 (depthwise-3) frames, the stem by 1.
__global__ void ragged_extents_kernel(const int32_t* __restrict__ lengths, int from_samples, int64_t n_samples, int hop, int T,
                                      int halo, int32_t* __restrict__ valid, int32_t* __restrict__ eff, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int64_t v = lengths[b];
    if (from_samples) { v = v < 0 ? 0 : (v > n_samples ? n_samples : v); v = 1 + v / hop; }
    v = v < 0 ? 0 : (v > T ? T : v);
    valid[b] = (int)v;
    eff[b] = (int)(v + halo > T ? T : v + halo);
}
// rows t >= valid[b] of out [B][T][D] <- 0 (16-byte stores; D * element size is a multiple of 16)
__global__ void zero_tail_kernel(uint4* __restrict__ out, const int32_t* __restrict__ valid, int T, int row_vec) {
    const int b = blockIdx.y;
    const int v = valid[b];
    const int64_t n = (int64_t)(T - v) * row_vec;
    uint4* o = out + ((int64_t)b * T + v) * row_vec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = make_uint4(0u, 0u, 0u, 0u);
}

// GEMM + LayerNorm on the tensor cores: fused epilogue when the row fits TMEM (N <= 512), else
// GEMM(+residual) to the 16-bit operand format followed by the row kernel.
int tc_gemm_ln(const op16* A, const op16* W, const float* bias, const op16* res,
               const float* gamma, const float* beta, void* out, void* tmp, int64_t B, int64_t T, int K, int N,
               int taps, cudaStream_t st, const float* res32 = nullptr, float* out32 = nullptr, int out_bf16 = 0,
               const int32_t* frames_eff = nullptr) {
    TcGemmArgs g{};
    g.frames_eff = frames_eff;
    g.A = A; g.W = W; g.bias = bias; g.res = res; g.gamma = gamma; g.beta = beta;
    g.B = B; g.T = T; g.K = K; g.N = N; g.taps = taps; g.act = ACT_NONE; g.eps = 1e-5f;
    if (tc_gemm_supported(K, N, TC_LN)) {
        g.epilogue = TC_LN; g.out = out; g.res32 = res32; g.out32 = out32; g.out_bf16 = out_bf16;
        return launch_gemm_tc(g, st);
    }
    g.epilogue = res ? TC_RES_ACT : TC_BIAS_ACT; g.out = tmp;
    ASRB_TRY(launch_gemm_tc(g, st));
    return launch_layernorm(tmp, nullptr, gamma, beta, out, DT_OP16, B * T, N, 1e-5f, st, out_bf16 ? DT_BF16 : DT_OP16);
}

// Stem of one feature stream: conv1 (mels -> D, k3; its input is already in w.a0) or conv2 (1 -> D, k3, straight
// from x), model.py:152-155, + layer 0's leading activation, written to X (B x T rows).
int encoder_stem(asrb_encoder* e, const EncBuffers& w, const float* x_c1, int in_ch, int64_t B, int64_t T, void* X, cudaStream_t st,
                 const int32_t* frames_eff = nullptr) {
    const int D = e->cfg.dims;
    const bool bf = e->cfg.compute == ASRB_BF16;
    const DType dt = bf ? DT_OP16 : DT_F32;
    const Act stem_act = ACT_GELU;                        // layer 0's leading act_fn (model.py:143)
    if (in_ch == 1) {
        if (!e->stem2_f) return fail(ASRB_E_WEIGHTS, "conv2.0.weight was not supplied: single-channel input unsupported");
        return launch_gemm_simt(x_c1, DT_F32, e->stem2_f, e->stem2_b, nullptr, X, dt, B, T, 1, D, 3, stem_act, st);
    }
    if (bf) {
        TcGemmArgs g{};
        g.A = (const op16*)w.a0; g.W = e->stem1_h; g.bias = e->stem1_b; g.out = X;
        g.B = B; g.T = T; g.K = e->CP; g.N = D; g.taps = 3; g.epilogue = TC_BIAS_ACT; g.act = stem_act;
        g.frames_eff = frames_eff;
        return launch_gemm_tc(g, st);
    }
    return launch_gemm_simt(w.a0, DT_F32, e->stem1_f, e->stem1_b, nullptr, X, DT_F32, B, T, e->cfg.mels, D, 3, stem_act, st);
}

int encoder_layers(asrb_encoder* e, const EncBuffers& w, int64_t B, int64_t T, void* out, int out_dtype, cudaStream_t st,
                   const int32_t* frames_eff = nullptr);

// Everything after the stem input is in place (a0 for in_ch == mels, x itself for in_ch == 1).
int encoder_body(asrb_encoder* e, const EncBuffers& w, const float* x_c1, int in_ch, int64_t B, int64_t T,
                 void* out, int out_dtype, cudaStream_t st, const int32_t* frames_eff = nullptr) {
    ASRB_TRY(encoder_stem(e, w, x_c1, in_ch, B, T, w.X, st, in_ch == 1 ? nullptr : frames_eff));
    return encoder_layers(e, w, B, T, out, out_dtype, st, frames_eff);
}

// Can padded tiles be skipped without changing any valid output frame?  Only in the conv stack on the tensor-core path: the
// TransformerEncoderLayer attends over every frame without a mask (model.py:163), so there the padding is part of the result.
bool can_skip_padding(const asrb_encoder* e) { return e->cfg.compute == ASRB_BF16 && !e->cfg.enc; }

// lengths -> (valid, eff) on the device; returns the array to hand to the kernels (NULL: compute everything)
int ragged_prepare(const asrb_encoder* e, const EncBuffers& w, const int32_t* lengths, bool from_samples, int64_t n_samples, int hop,
                   int64_t B, int64_t T, cudaStream_t st, const int32_t** frames_eff) {
    const int halo = 9 * e->cfg.layer + 1;
    ragged_extents_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(lengths, from_samples ? 1 : 0, n_samples, hop, (int)T, halo,
                                                                       w.frames_valid, w.frames_eff, (int)B);
    ASRB_LAUNCH_CHECK();
    *frames_eff = can_skip_padding(e) ? w.frames_eff : nullptr;
    return ASRB_OK;
}
int ragged_finish(const asrb_encoder* e, const EncBuffers& w, void* out, int out_dtype, int64_t B, int64_t T, cudaStream_t st) {
    const int row_vec = e->cfg.dims * (out_dtype == ASRB_BF16 ? 2 : 4) / 16;
    ProfScope ps("zero_padded_rows", st, 0.0, 0.0);
    zero_tail_kernel<<<dim3(64, (unsigned)B), 256, 0, st>>>((uint4*)out, w.frames_valid, (int)T, row_vec);
    ASRB_LAUNCH_CHECK();
    return ASRB_OK;
}

// The layer stack (and the optional TransformerEncoderLayer) over B x T rows whose stem output sits in w.X.
int encoder_layers(asrb_encoder* e, const EncBuffers& w, int64_t B, int64_t T, void* out, int out_dtype, cudaStream_t st,
                   const int32_t* frames_eff) {
    const int D = e->cfg.dims, L = e->cfg.layer;
    const bool bf = e->cfg.compute == ASRB_BF16;
    const DType dt = bf ? DT_OP16 : DT_F32;
    const int64_t rows = B * T;

    ASRB_TRY(launch_pos_table(w.pos, e->pos_scales, T, D, st));

    // ---- conv blocks (model.py:142-147) ----
    const bool direct_out = !e->cfg.enc;              // the last depthwise kernel stores the result itself
    for (int i = 0; i < L; ++i) {
        const LayerW& lw = e->layers[i];
        const bool last = i == L - 1;
        if (bf) {
            ASRB_TRY(tc_gemm_ln((const op16*)w.X, lw.wc_h, lw.bc, nullptr, lw.gamma, lw.beta, w.Y, w.H, B, T, D, D, 3, st,
                                nullptr, nullptr, 0, frames_eff));
            // point1 + GLU + depthwise-15 (BatchNorm folded) + SiLU in one kernel: Y -> U
            TcGemmArgs g{};
            g.A = (const op16*)w.Y; g.W = lw.w1_h; g.bias = lw.b1_glu; g.out = w.U;
            g.B = B; g.T = T; g.K = D; g.N = 2 * D; g.taps = 1; g.epilogue = TC_GLU_DW; g.act = ACT_NONE;
            g.dw_w = lw.dw15; g.dw_b = lw.dw15_b; g.dw_kw = 15; g.dw_act = ACT_SILU; g.frames_eff = frames_eff;
            ASRB_TRY(launch_gemm_tc(g, st));
            TcGemmArgs h{};
            h.A = (const op16*)w.U; h.W = lw.w2_h; h.bias = lw.b2; h.res = (const op16*)w.Y;
            h.B = B; h.T = T; h.K = D; h.N = D; h.taps = 1; h.act = ACT_GELU;
            // point2 + residual + GELU + depthwise-3 + GELU (+ next block's GELU | + sinusoids): U, Y -> X
            const bool to_out = last && !e->cfg.enc && out_dtype == ASRB_BF16;
            h.epilogue = TC_RES_ACT_DW; h.out = to_out ? out : w.X; h.out_bf16 = to_out;
            h.dw_w = lw.dw3; h.dw_b = lw.dw3_b; h.dw_kw = 3; h.dw_act = last ? ACT_GELU : ACT_GELU_GELU;
            h.pos = last ? w.pos : nullptr; h.frames_eff = frames_eff;
            h.out32 = (last && e->cfg.enc && tc_gemm_supported(D, D, TC_LN)) ? (float*)w.G : nullptr;
            ASRB_TRY(launch_gemm_tc(h, st));
            if (last && !e->cfg.enc && out_dtype != ASRB_BF16) {
                ProfScope ps("convert", st, 0.0, 6.0 * rows * D);
                convert_kernel<op16, float><<<148 * 8, 256, 0, st>>>((const op16*)w.X, (float*)out, rows * D);
                ASRB_LAUNCH_CHECK();
            }
            continue;
        } else {
            ASRB_TRY(launch_gemm_simt(w.X, DT_F32, lw.wc_f, lw.bc, nullptr, w.H, DT_F32, B, T, D, D, 3, ACT_NONE, st));
            ASRB_TRY(launch_layernorm(w.H, nullptr, lw.gamma, lw.beta, w.Y, DT_F32, rows, D, 1e-5f, st));
            ASRB_TRY(launch_gemm_simt(w.Y, DT_F32, lw.w1_f, lw.b1, nullptr, w.wide, DT_F32, B, T, D, 2 * D, 1, ACT_NONE, st));
            ASRB_TRY(launch_glu(w.wide, w.G, DT_F32, rows, D, st));
            ASRB_TRY(launch_dwconv(w.G, DT_F32, lw.dw15, lw.dw15_b, w.U, DT_F32, B, T, D, 15, ACT_SILU, nullptr, false, st));
            ASRB_TRY(launch_gemm_simt(w.U, DT_F32, lw.w2_f, lw.b2, w.Y, w.H, DT_F32, B, T, D, D, 1, ACT_GELU, st));
        }
        // fp32 variant: depthwise k3 + GELU (+ the next block's leading GELU; + sinusoids after the last block)
        const bool to_out = last && direct_out && out_dtype == ASRB_F32;
        ASRB_TRY(launch_dwconv(w.H, DT_F32, lw.dw3, lw.dw3_b, to_out ? out : w.X, DT_F32, B, T, D, 3, last ? ACT_GELU : ACT_GELU_GELU,
                               last ? w.pos : nullptr, false, st, nullptr));
        if (last && direct_out && !to_out) {               // fp32 variant asked for bf16 hidden states
            convert_kernel<float, __nv_bfloat16><<<148 * 8, 256, 0, st>>>((const float*)w.X, (__nv_bfloat16*)out, rows * D);
            ASRB_LAUNCH_CHECK();
        }
    }
    if (!e->cfg.enc) return ASRB_OK;

    // ---- nn.TransformerEncoderLayer, post-norm, ReLU, no mask (model.py:138,163) ----
    const int H = e->cfg.head, F = e->cfg.ffn;
    const float scale = 1.0f / sqrtf((float)(D / H));
    const bool same = (out_dtype == ASRB_BF16) == bf;
    void* fin = same ? out : w.U;
    if (bf) {
        TcGemmArgs q{};
        q.A = (const op16*)w.X; q.W = e->win_h; q.bias = e->bin; q.out = w.wide;
        q.B = B; q.T = T; q.K = D; q.N = 3 * D; q.taps = 1; q.epilogue = TC_BIAS_ACT; q.act = ACT_NONE;
        ASRB_TRY(launch_gemm_tc(q, st));
        if (attention_tc_supported(D, H)) ASRB_TRY(launch_attention_tc(w.wide, w.U, B, T, D, H, scale, st));
        else ASRB_TRY(launch_attention_simt(w.wide, w.U, DT_OP16, B, T, D, H, scale, st));
        const bool fused_ln = tc_gemm_supported(D, D, TC_LN);
        float* x32 = fused_ln ? (float*)w.G : nullptr;       // written by the last depthwise kernel
        float* y32 = fused_ln ? (float*)w.H : nullptr;
        ASRB_TRY(tc_gemm_ln((const op16*)w.U, e->wo_h, e->bo, (const op16*)w.X, e->n1g, e->n1b, w.Y, w.wide, B, T, D, D, 1, st, x32, y32));
        TcGemmArgs f1{};
        f1.A = (const op16*)w.Y; f1.W = e->wf1_h; f1.bias = e->bf1; f1.out = w.ffn;
        f1.B = B; f1.T = T; f1.K = D; f1.N = F; f1.taps = 1; f1.epilogue = TC_BIAS_ACT; f1.act = ACT_RELU;
        ASRB_TRY(launch_gemm_tc(f1, st));
        ASRB_TRY(tc_gemm_ln((const op16*)w.ffn, e->wf2_h, e->bf2, (const op16*)w.Y, e->n2g, e->n2b, fin, w.wide, B, T, F, D, 1, st, y32, nullptr, same ? 1 : 0));
        if (!same) {
            ProfScope ps("convert", st, 0.0, 6.0 * rows * D);
            convert_kernel<op16, float><<<148 * 8, 256, 0, st>>>((const op16*)fin, (float*)out, rows * D);
            ASRB_LAUNCH_CHECK();
        }
    } else {
        ASRB_TRY(launch_gemm_simt(w.X, DT_F32, e->win_f, e->bin, nullptr, w.wide, DT_F32, B, T, D, 3 * D, 1, ACT_NONE, st));
        ASRB_TRY(launch_attention_simt(w.wide, w.U, DT_F32, B, T, D, H, scale, st));
        ASRB_TRY(launch_gemm_simt(w.U, DT_F32, e->wo_f, e->bo, w.X, w.H, DT_F32, B, T, D, D, 1, ACT_NONE, st));
        ASRB_TRY(launch_layernorm(w.H, nullptr, e->n1g, e->n1b, w.Y, DT_F32, rows, D, 1e-5f, st));
        ASRB_TRY(launch_gemm_simt(w.Y, DT_F32, e->wf1_f, e->bf1, nullptr, w.ffn, DT_F32, B, T, D, F, 1, ACT_RELU, st));
        ASRB_TRY(launch_gemm_simt(w.ffn, DT_F32, e->wf2_f, e->bf2, w.Y, w.H, DT_F32, B, T, F, D, 1, ACT_NONE, st));
        ASRB_TRY(launch_layernorm(w.H, nullptr, e->n2g, e->n2b, fin, DT_F32, rows, D, 1e-5f, st));
        if (!same) { convert_kernel<float, __nv_bfloat16><<<148 * 8, 256, 0, st>>>((const float*)fin, (__nv_bfloat16*)out, rows * D); ASRB_LAUNCH_CHECK(); }
    }
    return ASRB_OK;
}

int check_forward_args(const asrb_encoder* e, int64_t B, int32_t in_ch, int64_t T, const void* out, int out_dtype) {
    if (!e) return fail(ASRB_E_ARG, "encoder forward: NULL handle");
    if (B < 0 || T < 0) return fail(ASRB_E_ARG, "encoder forward: bad shape B=%lld T=%lld", (long long)B, (long long)T);
    if (in_ch != e->cfg.mels && in_ch != 1)
        return fail(ASRB_E_ARG, "encoder forward: %d input channels, expected %d (conv1) or 1 (conv2)", in_ch, e->cfg.mels);
    if (out_dtype != ASRB_F32 && out_dtype != ASRB_BF16) return fail(ASRB_E_ARG, "encoder forward: out_dtype=%d", out_dtype);
    if (B > 65535) return fail(ASRB_E_ARG, "encoder forward: batch %lld > 65535", (long long)B);
    if (B * T > 0 && !out) return fail(ASRB_E_ARG, "encoder forward: NULL output");
    return ASRB_OK;
}

}  // namespace

extern "C" size_t asrb_encoder_workspace_bytes(const asrb_encoder* e, int64_t B, int64_t T) {
    if (!e || B < 0 || T < 0) return 0;
    return enc_ws_bytes(e, B, T);
}

static int encoder_forward_impl(const char* who, asrb_encoder* e, const float* x, int64_t B, int32_t in_ch, int64_t T,
                                const int32_t* frames, bool ragged, void* out, int out_dtype, void* ws, size_t ws_bytes, void* stream) {
    ASRB_TRY(check_forward_args(e, B, in_ch, T, out, out_dtype));
    if (B == 0 || T == 0) return ASRB_OK;
    if (!x) return fail(ASRB_E_ARG, "%s: NULL input", who);
    if (ragged && !frames) return fail(ASRB_E_ARG, "%s: NULL frame counts", who);
    if (!ws || ws_bytes < enc_ws_bytes(e, B, T) || ((uintptr_t)ws & 255))
        return fail(ASRB_E_WORKSPACE, "%s: workspace NULL, not 256-B aligned or smaller than %zu B", who, enc_ws_bytes(e, B, T));
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    EncBuffers w = carve(e, B, T, ws, ws_bytes);
    if (!w.ok) return fail(ASRB_E_WORKSPACE, "%s: workspace carve failed", who);
    const bool bf = e->cfg.compute == ASRB_BF16;
    if (in_ch != 1)
        ASRB_TRY(launch_to_channels_last(x, w.a0, bf ? DT_OP16 : DT_F32, B, in_ch, bf ? e->CP : in_ch, T,
                                         nullptr, nullptr, 0, 1, false, st));
    const int32_t* eff = nullptr;
    if (ragged) ASRB_TRY(ragged_prepare(e, w, frames, false, 0, 1, B, T, st, &eff));
    ASRB_TRY(encoder_body(e, w, x, in_ch, B, T, out, out_dtype, st, eff));
    return ragged ? ragged_finish(e, w, out, out_dtype, B, T, st) : ASRB_OK;
}

extern "C" int asrb_encoder_forward(asrb_encoder* e, const float* x, int64_t B, int32_t in_ch, int64_t T, void* out,
                                    int out_dtype, void* ws, size_t ws_bytes, void* stream) {
    return encoder_forward_impl("asrb_encoder_forward", e, x, B, in_ch, T, nullptr, false, out, out_dtype, ws, ws_bytes, stream);
}

extern "C" int asrb_encoder_forward_ragged(asrb_encoder* e, const float* x, int64_t B, int32_t in_ch, int64_t T,
                                           const int32_t* frames, void* out, int out_dtype, void* ws, size_t ws_bytes,
                                           void* stream) {
    return encoder_forward_impl("asrb_encoder_forward_ragged", e, x, B, in_ch, T, frames, true, out, out_dtype, ws, ws_bytes, stream);
}

extern "C" int asrb_encoder_forward_streams(asrb_encoder* e, int32_t n_streams, const float* const* x, const int32_t* in_ch,
                                            int64_t B, int64_t T, void* out, int out_dtype, void* ws, size_t ws_bytes,
                                            void* stream) {
    if (n_streams < 1 || n_streams > 8 || !x || !in_ch) return fail(ASRB_E_ARG, "asrb_encoder_forward_streams: bad stream list");
    for (int s = 0; s < n_streams; ++s) {
        ASRB_TRY(check_forward_args(e, B * n_streams, in_ch[s], T, out, out_dtype));
        if (B * T > 0 && !x[s]) return fail(ASRB_E_ARG, "asrb_encoder_forward_streams: NULL input %d", s);
    }
    if (B == 0 || T == 0) return ASRB_OK;
    const int64_t Bt = B * n_streams;
    if (!ws || ws_bytes < enc_ws_bytes(e, Bt, T) || ((uintptr_t)ws & 255))
        return fail(ASRB_E_WORKSPACE, "asrb_encoder_forward_streams: workspace NULL, not 256-B aligned or smaller than %zu B", enc_ws_bytes(e, Bt, T));
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    EncBuffers w = carve(e, Bt, T, ws, ws_bytes);
    if (!w.ok) return fail(ASRB_E_WORKSPACE, "asrb_encoder_forward_streams: workspace carve failed");
    const bool bf = e->cfg.compute == ASRB_BF16;
    const size_t es = bf ? 2 : 4;
    for (int s = 0; s < n_streams; ++s) {                  // stems one by one (conv1 | conv2), into stream s's rows of X
        if (in_ch[s] != 1)
            ASRB_TRY(launch_to_channels_last(x[s], w.a0, bf ? DT_OP16 : DT_F32, B, in_ch[s], bf ? e->CP : in_ch[s], T,
                                             nullptr, nullptr, 0, 1, false, st));
        ASRB_TRY(encoder_stem(e, w, x[s], in_ch[s], B, T, (char*)w.X + (size_t)s * B * T * e->cfg.dims * es, st));
    }
    return encoder_layers(e, w, Bt, T, out, out_dtype, st);   // one pass of the layer stack over all streams
}

extern "C" size_t asrb_pcm_to_hidden_workspace_bytes(const asrb_logmel_plan* pl, const asrb_encoder* e, int64_t B,
                                                     int64_t n_samples) {
    if (!pl || !e || B < 0 || n_samples < 0) return 0;
    const int64_t T = 1 + n_samples / pl->hop;
    return enc_ws_bytes(e, B, T) + align_up(sizeof(float) * (size_t)B * pl->n_mels * T, 256) +
           align_up(sizeof(uint32_t) * logmel_keys_words(pl, B, n_samples), 256) + 256;
}

static int pcm_to_hidden_impl(const char* who, const asrb_logmel_plan* pl, asrb_encoder* e, const float* pcm, int64_t B,
                              int64_t n_samples, int64_t pcm_stride, const int32_t* lengths, bool ragged, float* logmel_out,
                              void* out, int out_dtype, void* ws, size_t ws_bytes, void* stream) {
    if (!pl) return fail(ASRB_E_ARG, "%s: NULL plan", who);
    if (n_samples < 0 || pcm_stride < n_samples) return fail(ASRB_E_ARG, "%s: bad n_samples / stride", who);
    const int64_t T = 1 + n_samples / pl->hop;
    ASRB_TRY(check_forward_args(e, B, pl->n_mels, T, out, out_dtype));
    if (pl->n_mels != e->cfg.mels) return fail(ASRB_E_ARG, "%s: plan has %d mels, encoder %d", who, pl->n_mels, e->cfg.mels);
    if (B == 0) return ASRB_OK;
    if (!pcm && n_samples > 0) return fail(ASRB_E_ARG, "%s: NULL pcm", who);
    if (ragged && !lengths) return fail(ASRB_E_ARG, "%s: NULL lengths", who);
    const size_t need = asrb_pcm_to_hidden_workspace_bytes(pl, e, B, n_samples);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255))
        return fail(ASRB_E_WORKSPACE, "%s: workspace NULL, not 256-B aligned or smaller than %zu B", who, need);
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const size_t enc_bytes = enc_ws_bytes(e, B, T);
    EncBuffers w = carve(e, B, T, ws, enc_bytes);
    Arena tail((char*)ws + enc_bytes, ws_bytes - enc_bytes);
    float* mel = logmel_out ? logmel_out : tail.take<float>((size_t)B * pl->n_mels * T);
    uint32_t* keys = tail.take<uint32_t>(logmel_keys_words(pl, B, n_samples));
    if (!w.ok || !tail.ok()) return fail(ASRB_E_WORKSPACE, "%s: workspace carve failed", who);
    const int32_t* eff = nullptr;
    if (ragged) ASRB_TRY(ragged_prepare(e, w, lengths, true, n_samples, pl->hop, B, T, st, &eff));
    const bool bf = e->cfg.compute == ASRB_BF16;
    if (bf && !logmel_out) {
        // the front end writes the stem GEMM's operand itself (16-bit channels-last) and the floor runs in place
        ASRB_TRY(logmel_pass1(pl, pcm, B, n_samples, pcm_stride, lengths, nullptr, keys, st, (op16*)w.a0, e->CP));
        ASRB_TRY(logmel_floor_cl(pl, (op16*)w.a0, e->CP, keys, lengths, B, n_samples, st));
    } else {
        ASRB_TRY(logmel_pass1(pl, pcm, B, n_samples, pcm_stride, lengths, mel, keys, st));
        // the dynamic-range floor (essentials.py:489) is applied while changing layout for conv1
        ASRB_TRY(launch_to_channels_last(mel, w.a0, bf ? DT_OP16 : DT_F32, B, pl->n_mels, bf ? e->CP : pl->n_mels, T,
                                         keys, lengths, n_samples, pl->hop, logmel_out != nullptr, st));
    }
    ASRB_TRY(encoder_body(e, w, nullptr, pl->n_mels, B, T, out, out_dtype, st, eff));
    return ragged ? ragged_finish(e, w, out, out_dtype, B, T, st) : ASRB_OK;
}

extern "C" int asrb_pcm_to_hidden(const asrb_logmel_plan* pl, asrb_encoder* e, const float* pcm, int64_t B,
                                  int64_t n_samples, int64_t pcm_stride, const int32_t* lengths, float* logmel_out,
                                  void* out, int out_dtype, void* ws, size_t ws_bytes, void* stream) {
    return pcm_to_hidden_impl("asrb_pcm_to_hidden", pl, e, pcm, B, n_samples, pcm_stride, lengths, false, logmel_out, out, out_dtype,
                              ws, ws_bytes, stream);
}

extern "C" int asrb_pcm_to_hidden_ragged(const asrb_logmel_plan* pl, asrb_encoder* e, const float* pcm, int64_t B,
                                         int64_t n_samples, int64_t pcm_stride, const int32_t* lengths, float* logmel_out,
                                         void* out, int out_dtype, void* ws, size_t ws_bytes, void* stream) {
    return pcm_to_hidden_impl("asrb_pcm_to_hidden_ragged", pl, e, pcm, B, n_samples, pcm_stride, lengths, true, logmel_out, out,
                              out_dtype, ws, ws_bytes, stream);
}
