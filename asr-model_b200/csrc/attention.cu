// Secondary block: the live branch of `attention`
// (model.py:258-262, 302-307, 316-317, n_type="rmsnorm") with `rotary` (model.py:191-214)
// applied to encoded audio, batched as the reference's B=1 semantics per utterance (per-sample rotary magnitudes).
// Two variants: ASRB_F32 -- fp32 on CUDA cores (the <= 1e-4 check); ASRB_BF16 -- the tensor-core variant: RMSNorm -> 16-bit
// operand, q / kv / out projections on the tcgen05 GEMM, rotary + per-head RMSNorm in place on the 16-bit q and k, the
// tcgen05 flash attention of attn_tc.cu.  The tensor-core variant also splits the block at the K|V of the attended
// sequence: asrb_attention_encode_kv() computes them ONCE for an encoded-audio tensor, and
// asrb_attention_forward_cached() attends any number of query sequences against that cache -- the reference recomputes
// them for each of the 8 `residual` calls per decoder block and for every generated token (model.py:573-583, 617-626,
// 691-699).
#include "enc_kernels.cuh"
#include <map>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

using namespace asrb;

struct asrb_attention {
    int dims, head, compute;
    std::vector<void*> owned;
    float *q_norm, *q_w, *q_b, *kv_norm, *kv_w, *kv_b, *out_w, *out_b, *ln_w, *freqs;
    op16 *q_wh, *kv_wh, *out_wh;                        // tensor-core variant: weights in the 16-bit operand format
};

extern "C" int asrb_attention_create(int32_t dims, int32_t head, int compute, int n_tensors, const char* const* names,
                                     const float* const* host_data, const int64_t* numels, asrb_attention** out) {
    if (!out || dims <= 0 || head <= 0 || dims % head) return fail(ASRB_E_ARG, "asrb_attention_create: bad dims/head");
    if (compute != ASRB_F32 && compute != ASRB_BF16) return fail(ASRB_E_ARG, "asrb_attention_create: compute=%d", compute);
    const int hd = dims / head;
    if (hd != 16 && hd != 32 && hd != 64 && hd != 128) return fail(ASRB_E_ARG, "asrb_attention_create: head_dim %d unsupported", hd);
    if (compute == ASRB_BF16 && (dims % 128 || !attention_tc_supported(dims, head)))
        return fail(ASRB_E_ARG, "asrb_attention_create: the tensor-core variant needs dims %% 128 == 0 and head_dim 64 or 128");
    ASRB_TRY(require_sm100());
    std::map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n_tensors; ++i) t[names[i]] = {host_data[i], numels[i]};
    asrb_attention* a = new asrb_attention();
    a->dims = dims; a->head = head; a->compute = compute;
    a->q_wh = a->kv_wh = a->out_wh = nullptr;
    auto up16 = [&](const char* key, int64_t n, op16** dst) -> int {      // fp32 host weights -> 16-bit operand on the device
        auto it = t.find(key);
        if (it == t.end() || it->second.second != n) return fail(ASRB_E_WEIGHTS, "attention tensor '%s' missing or misshaped", key);
        std::vector<op16> h((size_t)n);
        for (int64_t i = 0; i < n; ++i) h[i] = host_to_op16(it->second.first[i]);
        void* p = nullptr;
        ASRB_CUDA(cudaMalloc(&p, sizeof(op16) * n));
        a->owned.push_back(p);
        ASRB_CUDA(cudaMemcpy(p, h.data(), sizeof(op16) * n, cudaMemcpyHostToDevice));
        *dst = (op16*)p;
        return ASRB_OK;
    };
    auto up = [&](const char* key, int64_t n, float** dst) -> int {
        auto it = t.find(key);
        if (it == t.end() || it->second.second != n) return fail(ASRB_E_WEIGHTS, "attention tensor '%s' missing or misshaped", key);
        void* p = nullptr;
        ASRB_CUDA(cudaMalloc(&p, sizeof(float) * n));
        a->owned.push_back(p);
        ASRB_CUDA(cudaMemcpy(p, it->second.first, sizeof(float) * n, cudaMemcpyHostToDevice));
        *dst = (float*)p;
        return ASRB_OK;
    };
    const int64_t D = dims;
    int r = up("q.0.weight", D, &a->q_norm);
    if (!r) r = up("q.1.weight", D * D, &a->q_w);
    if (!r) r = up("q.1.bias", D, &a->q_b);
    if (!r) r = up("kv.0.weight", D, &a->kv_norm);
    if (!r) r = up("kv.1.weight", 2 * D * D, &a->kv_w);
    if (!r) r = up("kv.1.bias", 2 * D, &a->kv_b);
    if (!r) r = up("out.1.weight", D * D, &a->out_w);
    if (!r) r = up("out.1.bias", D, &a->out_b);
    if (!r) r = up("ln.weight", hd, &a->ln_w);
    if (!r && compute == ASRB_BF16) {
        r = up16("q.1.weight", D * D, &a->q_wh);
        if (!r) r = up16("kv.1.weight", 2 * D * D, &a->kv_wh);
        if (!r) r = up16("out.1.weight", D * D, &a->out_wh);
    }
    if (!r) {
        if (t.count("__rot_freqs")) r = up("__rot_freqs", hd / 2, &a->freqs);
        else {  // compute_f(mask=None): 200 * (40**linspace(0,1,hd/2) * 200/1000) / 1000  (model.py:191-194)
            std::vector<float> f(hd / 2);
            for (int j = 0; j < hd / 2; ++j) {
                const float lin = hd / 2 > 1 ? (float)j / (float)(hd / 2 - 1) : 0.f;
                f[j] = 200.0f * (powf(40.0f, lin) * 200.0f / 1000.0f) / 1000.0f;
            }
            t["__rot_freqs"] = {f.data(), hd / 2};
            r = up("__rot_freqs", hd / 2, &a->freqs);
        }
    }
    if (r) { asrb_attention_destroy(a); return r; }
    *out = a;
    return ASRB_OK;
}

extern "C" void asrb_attention_destroy(asrb_attention* a) {
    if (!a) return;
    for (void* p : a->owned) cudaFree(p);
    delete a;
}

extern "C" size_t asrb_attention_workspace_bytes(const asrb_attention* a, int64_t B, int64_t T) {
    if (!a || B < 0 || T < 0) return 0;
    const size_t rows = (size_t)B * T, D = a->dims;
    if (a->compute == ASRB_BF16)          // xn, q, att (16-bit) + a K|V cache for the self-attention call
        return 3 * align_up(rows * D * 2, 256) + align_up(rows * 2 * D * 2, 256) + 256;
    return 5 * align_up(rows * D * 4, 256) + align_up(rows * 2 * D * 4, 256) + 256;
}

extern "C" size_t asrb_attention_kv_bytes(const asrb_attention* a, int64_t B, int64_t Tk) {
    if (!a || B < 0 || Tk < 0) return 0;
    return align_up((size_t)B * Tk * 2 * a->dims * sizeof(op16), 256);
}

// K (rotary + per-head RMSNorm applied) | V of the attended sequence xa [B][Tk][D] fp32 -> kv [B][Tk][2D] op16
static int attention_encode_kv(asrb_attention* a, const float* xa, int64_t B, int64_t Tk, op16* kv, op16* xn, cudaStream_t st) {
    const int D = a->dims, H = a->head, hd = D / H;
    const float pre = powf((float)hd, -0.25f);                       // n.scale, model.py:239
    ASRB_TRY(launch_rmsnorm(xa, a->kv_norm, xn, DT_OP16, B * Tk, D, st));
    TcGemmArgs g{};
    g.A = xn; g.W = a->kv_wh; g.bias = a->kv_b; g.out = kv;
    g.B = B; g.T = Tk; g.K = D; g.N = 2 * D; g.taps = 1; g.epilogue = TC_BIAS_ACT; g.act = ACT_NONE;
    ASRB_TRY(launch_gemm_tc(g, st));
    return launch_rotary_headnorm(kv, DT_OP16, 2 * D, xa, a->ln_w, a->freqs, B, Tk, D, H, pre, st);   // k = first D columns
}

// queries x [B][Tq][D] fp32 against a cache kv [B][Tk][2D] -> out [B][Tq][D] fp32
static int attention_queries(asrb_attention* a, const float* x, int64_t B, int64_t Tq, const op16* kv, int64_t Tk, float* out,
                             op16* xn, op16* q, op16* att, cudaStream_t st) {
    const int D = a->dims, H = a->head, hd = D / H;
    const float pre = powf((float)hd, -0.25f);
    ASRB_TRY(launch_rmsnorm(x, a->q_norm, xn, DT_OP16, B * Tq, D, st));
    TcGemmArgs g{};
    g.A = xn; g.W = a->q_wh; g.bias = a->q_b; g.out = q;
    g.B = B; g.T = Tq; g.K = D; g.N = D; g.taps = 1; g.epilogue = TC_BIAS_ACT; g.act = ACT_NONE;
    ASRB_TRY(launch_gemm_tc(g, st));
    ASRB_TRY(launch_rotary_headnorm(q, DT_OP16, D, x, a->ln_w, a->freqs, B, Tq, D, H, pre, st));      // model.py:303-307
    ASRB_TRY(launch_attention_tc_ex(q, D, 0, kv, 2 * D, 0, D, att, B, Tq, Tk, D, H, 1.0f / sqrtf((float)hd), st));
    TcGemmArgs o{};
    o.A = att; o.W = a->out_wh; o.bias = a->out_b; o.out = out; o.out_f32 = 1;
    o.B = B; o.T = Tq; o.K = D; o.N = D; o.taps = 1; o.epilogue = TC_BIAS_ACT; o.act = ACT_NONE;
    return launch_gemm_tc(o, st);
}

static int attention_check(const char* who, const asrb_attention* a, int64_t B, int64_t T, const void* p0, const void* p1,
                           const void* ws, size_t ws_bytes, size_t need) {
    if (!a) return fail(ASRB_E_ARG, "%s: NULL handle", who);
    if (B < 0 || T < 0 || B > 65535) return fail(ASRB_E_ARG, "%s: bad shape", who);
    if (B * T > 0 && (!p0 || !p1)) return fail(ASRB_E_ARG, "%s: NULL tensor", who);
    if (B * T > 0 && (!ws || ws_bytes < need || ((uintptr_t)ws & 255)))
        return fail(ASRB_E_WORKSPACE, "%s: workspace NULL, misaligned or smaller than %zu B", who, need);
    return ASRB_OK;
}

extern "C" int asrb_attention_encode_kv(asrb_attention* a, const float* xa, int64_t B, int64_t Tk, void* kv_cache,
                                        void* ws, size_t ws_bytes, void* stream) {
    ASRB_TRY(attention_check("asrb_attention_encode_kv", a, B, Tk, xa, kv_cache, ws, ws_bytes, a ? asrb_attention_workspace_bytes(a, B, Tk) : 0));
    if (a->compute != ASRB_BF16) return fail(ASRB_E_ARG, "asrb_attention_encode_kv: the K|V cache belongs to the tensor-core variant");
    if (B == 0 || Tk == 0) return ASRB_OK;
    if ((uintptr_t)kv_cache & 255) return fail(ASRB_E_ARG, "asrb_attention_encode_kv: cache not 256-B aligned");
    ASRB_TRY(require_sm100());
    Arena ar(ws, ws_bytes);
    op16* xn = ar.take<op16>((size_t)B * Tk * a->dims);
    return attention_encode_kv(a, xa, B, Tk, (op16*)kv_cache, xn, (cudaStream_t)stream);
}

extern "C" int asrb_attention_forward_cached(asrb_attention* a, const float* x, int64_t B, int64_t Tq, const void* kv_cache,
                                             int64_t Tk, float* out, void* ws, size_t ws_bytes, void* stream) {
    ASRB_TRY(attention_check("asrb_attention_forward_cached", a, B, Tq, x, out, ws, ws_bytes, a ? asrb_attention_workspace_bytes(a, B, Tq) : 0));
    if (a->compute != ASRB_BF16) return fail(ASRB_E_ARG, "asrb_attention_forward_cached: the K|V cache belongs to the tensor-core variant");
    if (B == 0 || Tq == 0) return ASRB_OK;
    if (!kv_cache || Tk <= 0 || ((uintptr_t)kv_cache & 255)) return fail(ASRB_E_ARG, "asrb_attention_forward_cached: bad cache");
    ASRB_TRY(require_sm100());
    const size_t rows = (size_t)B * Tq, D = a->dims;
    Arena ar(ws, ws_bytes);
    op16* xn = ar.take<op16>(rows * D); op16* q = ar.take<op16>(rows * D); op16* att = ar.take<op16>(rows * D);
    return attention_queries(a, x, B, Tq, (const op16*)kv_cache, Tk, out, xn, q, att, (cudaStream_t)stream);
}

extern "C" int asrb_attention_forward(asrb_attention* a, const float* x, int64_t B, int64_t T, float* out, void* ws,
                                      size_t ws_bytes, void* stream) {
    ASRB_TRY(attention_check("asrb_attention_forward", a, B, T, x, out, ws, ws_bytes, a ? asrb_attention_workspace_bytes(a, B, T) : 0));
    if (B == 0 || T == 0) return ASRB_OK;
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const int D = a->dims, H = a->head, hd = D / H;
    const int64_t rows = B * T;
    Arena ar(ws, ws_bytes);
    if (a->compute == ASRB_BF16) {                                   // self-attention: the cache lives in the workspace
        op16* xn = ar.take<op16>(rows * D); op16* q = ar.take<op16>(rows * D); op16* att = ar.take<op16>(rows * D);
        op16* kv = ar.take<op16>(rows * 2 * D);
        ASRB_TRY(attention_encode_kv(a, x, B, T, kv, xn, st));
        return attention_queries(a, x, B, T, kv, T, out, xn, q, att, st);
    }
    float* xn = ar.take<float>(rows * D);
    float* q = ar.take<float>(rows * D);
    float* kv = ar.take<float>(rows * 2 * D);
    float* att = ar.take<float>(rows * D);
    const float pre = powf((float)hd, -0.25f);                       // n.scale, model.py:239
    ASRB_TRY(launch_rmsnorm(x, a->q_norm, xn, DT_F32, rows, D, st));
    ASRB_TRY(launch_gemm_simt(xn, DT_F32, a->q_w, a->q_b, nullptr, q, DT_F32, B, T, D, D, 1, ACT_NONE, st));
    ASRB_TRY(launch_rmsnorm(x, a->kv_norm, xn, DT_F32, rows, D, st));
    ASRB_TRY(launch_gemm_simt(xn, DT_F32, a->kv_w, a->kv_b, nullptr, kv, DT_F32, B, T, D, 2 * D, 1, ACT_NONE, st));
    ASRB_TRY(launch_rotary_headnorm(q, DT_F32, D, x, a->ln_w, a->freqs, B, T, D, H, pre, st));       // model.py:303-307
    ASRB_TRY(launch_rotary_headnorm(kv, DT_F32, 2 * D, x, a->ln_w, a->freqs, B, T, D, H, pre, st));  // k = first D columns
    ASRB_TRY(launch_attention_simt_ex(q, kv, kv + D, D, 2 * D, 2 * D, att, DT_F32, B, T, D, H, 1.0f / sqrtf((float)hd), st));
    return launch_gemm_simt(att, DT_F32, a->out_w, a->out_b, nullptr, out, DT_F32, B, T, D, D, 1, ACT_NONE, st);
}

// ------------------------------------------------------------------------------------------------------------------
// residual.mlp (model.py:573-574, 583): ln -> tgate -> Linear(D, n D) -> GELU -> Linear(n D, D) -> ln with ONE shared RMSNorm,
// on the tensor cores: RMSNorm -> 16-bit operand; ONE GEMM for the n gate projections and the selector logits
// ([n D + n] rows, zero-padded to a multiple of 128); softmax-weighted sigmoid gates (CUDA cores, one warp per row);
// Linear + GELU epilogue; Linear -> fp32; RMSNorm (+ residual).
// ------------------------------------------------------------------------------------------------------------------
struct asrb_mlp {
    int dims, n_types, n_gate;                          // n_gate = rows of the fused gate / selector weight (multiple of 128)
    std::vector<void*> owned;
    float *ln_w, *gate_b, *b1, *b2;
    op16 *gate_w, *w1, *w2;
};

extern "C" int asrb_mlp_create(int32_t dims, int32_t n_types, int n_tensors, const char* const* names,
                               const float* const* host_data, const int64_t* numels, asrb_mlp** out) {
    if (!out || dims <= 0 || dims % 128 || n_types < 1 || n_types > 8) return fail(ASRB_E_ARG, "asrb_mlp_create: dims %% 128 == 0, 1 <= n_types <= 8");
    ASRB_TRY(require_sm100());
    std::map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n_tensors; ++i) t[names[i]] = {host_data[i], numels[i]};
    auto get = [&](const std::string& key, int64_t n) -> const float* {
        auto it = t.find(key);
        if (it == t.end() || it->second.second != n) { fail(ASRB_E_WEIGHTS, "mlp tensor '%s' missing or misshaped", key.c_str()); return nullptr; }
        return it->second.first;
    };
    asrb_mlp* m = new asrb_mlp();
    m->dims = dims; m->n_types = n_types;
    const int64_t D = dims, nD = (int64_t)n_types * D;
    m->n_gate = (int)((nD + n_types + 127) / 128 * 128);
    auto up = [&](const void* src, size_t bytes, void** dst) -> int {
        void* p = nullptr;
        ASRB_CUDA(cudaMalloc(&p, bytes));
        m->owned.push_back(p);
        ASRB_CUDA(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        *dst = p;
        return ASRB_OK;
    };
    auto half = [](const float* w, size_t n) { std::vector<op16> h(n); for (size_t i = 0; i < n; ++i) h[i] = host_to_op16(w[i]); return h; };
    int r = ASRB_OK;
    std::vector<float> gw((size_t)m->n_gate * D, 0.f), gb((size_t)m->n_gate, 0.f);
    for (int i = 0; i < n_types && r == ASRB_OK; ++i) {
        const std::string k = "mlp.1.ga." + std::to_string(i) + ".0.";
        const float* w = get(k + "weight", D * D); const float* b = get(k + "bias", D);
        if (!w || !b) { r = ASRB_E_WEIGHTS; break; }
        memcpy(&gw[(size_t)i * D * D], w, sizeof(float) * D * D);
        memcpy(&gb[(size_t)i * D], b, sizeof(float) * D);
    }
    const float *ln = nullptr, *csw = nullptr, *csb = nullptr, *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    if (r == ASRB_OK) {
        ln = get("ln.weight", D); csw = get("mlp.1.cs.0.weight", (int64_t)n_types * D); csb = get("mlp.1.cs.0.bias", n_types);
        w1 = get("mlp.2.weight", nD * D); b1 = get("mlp.2.bias", nD); w2 = get("mlp.4.weight", D * nD); b2 = get("mlp.4.bias", D);
        if (!ln || !csw || !csb || !w1 || !b1 || !w2 || !b2) r = ASRB_E_WEIGHTS;
    }
    if (r == ASRB_OK) {
        memcpy(&gw[(size_t)nD * D], csw, sizeof(float) * n_types * D);
        memcpy(&gb[(size_t)nD], csb, sizeof(float) * n_types);
        auto gh = half(gw.data(), gw.size()); auto h1 = half(w1, (size_t)nD * D); auto h2 = half(w2, (size_t)D * nD);
        r = up(gh.data(), gh.size() * sizeof(op16), (void**)&m->gate_w);
        if (!r) r = up(h1.data(), h1.size() * sizeof(op16), (void**)&m->w1);
        if (!r) r = up(h2.data(), h2.size() * sizeof(op16), (void**)&m->w2);
        if (!r) r = up(gb.data(), gb.size() * sizeof(float), (void**)&m->gate_b);
        if (!r) r = up(b1, sizeof(float) * nD, (void**)&m->b1);
        if (!r) r = up(b2, sizeof(float) * D, (void**)&m->b2);
        if (!r) r = up(ln, sizeof(float) * D, (void**)&m->ln_w);
    }
    if (r != ASRB_OK) { asrb_mlp_destroy(m); return r; }
    *out = m;
    return ASRB_OK;
}

extern "C" void asrb_mlp_destroy(asrb_mlp* m) {
    if (!m) return;
    for (void* p : m->owned) cudaFree(p);
    delete m;
}

extern "C" size_t asrb_mlp_workspace_bytes(const asrb_mlp* m, int64_t B, int64_t T) {
    if (!m || B < 0 || T < 0) return 0;
    const size_t rows = (size_t)B * T, D = m->dims;
    return align_up(rows * D * 2, 256) * 2 + align_up(rows * (size_t)m->n_gate * 2, 256) + align_up(rows * D * m->n_types * 2, 256) +
           align_up(rows * D * 4, 256) + 256;
}

extern "C" int asrb_mlp_forward(asrb_mlp* m, const float* x, int64_t B, int64_t T, int add_residual, float* out, void* ws,
                                size_t ws_bytes, void* stream) {
    if (!m) return fail(ASRB_E_ARG, "asrb_mlp_forward: NULL handle");
    if (B < 0 || T < 0 || B > 65535) return fail(ASRB_E_ARG, "asrb_mlp_forward: bad shape");
    if (B == 0 || T == 0) return ASRB_OK;
    if (!x || !out) return fail(ASRB_E_ARG, "asrb_mlp_forward: NULL tensor");
    if (!ws || ws_bytes < asrb_mlp_workspace_bytes(m, B, T) || ((uintptr_t)ws & 255))
        return fail(ASRB_E_WORKSPACE, "asrb_mlp_forward: workspace NULL, misaligned or too small");
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const int D = m->dims, nt = m->n_types;
    const int64_t rows = B * T;
    Arena ar(ws, ws_bytes);
    op16* xn = ar.take<op16>(rows * D); op16* tg = ar.take<op16>(rows * D);
    op16* g = ar.take<op16>(rows * (size_t)m->n_gate); op16* h = ar.take<op16>(rows * (size_t)D * nt);
    float* y = ar.take<float>(rows * D);
    ASRB_TRY(launch_rmsnorm(x, m->ln_w, xn, DT_OP16, rows, D, st));
    TcGemmArgs a{};
    a.A = xn; a.W = m->gate_w; a.bias = m->gate_b; a.out = g;
    a.B = B; a.T = T; a.K = D; a.N = m->n_gate; a.taps = 1; a.epilogue = TC_BIAS_ACT; a.act = ACT_NONE;
    ASRB_TRY(launch_gemm_tc(a, st));
    ASRB_TRY(launch_tgate_combine(g, m->n_gate, tg, rows, D, nt, st));
    TcGemmArgs l1{};
    l1.A = tg; l1.W = m->w1; l1.bias = m->b1; l1.out = h;
    l1.B = B; l1.T = T; l1.K = D; l1.N = D * nt; l1.taps = 1; l1.epilogue = TC_BIAS_ACT; l1.act = ACT_GELU;
    ASRB_TRY(launch_gemm_tc(l1, st));
    TcGemmArgs l2{};
    l2.A = h; l2.W = m->w2; l2.bias = m->b2; l2.out = y; l2.out_f32 = 1;
    l2.B = B; l2.T = T; l2.K = D * nt; l2.N = D; l2.taps = 1; l2.epilogue = TC_BIAS_ACT; l2.act = ACT_NONE;
    ASRB_TRY(launch_gemm_tc(l2, st));
    return launch_rmsnorm(y, m->ln_w, out, DT_F32, rows, D, st, add_residual ? x : nullptr);
}
