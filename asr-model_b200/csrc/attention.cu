// Secondary block: the live branch of `attention`
// (model.py:258-262, 302-307, 316-317, n_type="rmsnorm") with `rotary` (model.py:191-214)
// applied to encoded audio, batched as the reference's B=1 semantics per utterance.
// fp32 on CUDA cores.
#include "enc_kernels.cuh"
#include <map>
#include <string>
#include <vector>
#include <cmath>

using namespace asrb;

struct asrb_attention {
    int dims, head;
    std::vector<void*> owned;
    float *q_norm, *q_w, *q_b, *kv_norm, *kv_w, *kv_b, *out_w, *out_b, *ln_w, *freqs;
};

extern "C" int asrb_attention_create(int32_t dims, int32_t head, int compute, int n_tensors, const char* const* names,
                                     const float* const* host_data, const int64_t* numels, asrb_attention** out) {
    if (!out || dims <= 0 || head <= 0 || dims % head) return fail(ASRB_E_ARG, "asrb_attention_create: bad dims/head");
    if (compute != ASRB_F32) return fail(ASRB_E_ARG, "asrb_attention_create: only ASRB_F32 is implemented for this block");
    const int hd = dims / head;
    if (hd != 16 && hd != 32 && hd != 64 && hd != 128) return fail(ASRB_E_ARG, "asrb_attention_create: head_dim %d unsupported", hd);
    ASRB_TRY(require_sm100());
    std::map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n_tensors; ++i) t[names[i]] = {host_data[i], numels[i]};
    asrb_attention* a = new asrb_attention();
    a->dims = dims; a->head = head;
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
    return 5 * align_up(rows * D * 4, 256) + align_up(rows * 2 * D * 4, 256) + 256;
}

extern "C" int asrb_attention_forward(asrb_attention* a, const float* x, int64_t B, int64_t T, float* out, void* ws,
                                      size_t ws_bytes, void* stream) {
    if (!a) return fail(ASRB_E_ARG, "asrb_attention_forward: NULL handle");
    if (B < 0 || T < 0 || B > 65535) return fail(ASRB_E_ARG, "asrb_attention_forward: bad shape");
    if (B == 0 || T == 0) return ASRB_OK;
    if (!x || !out) return fail(ASRB_E_ARG, "asrb_attention_forward: NULL tensor");
    if (!ws || ws_bytes < asrb_attention_workspace_bytes(a, B, T) || ((uintptr_t)ws & 255))
        return fail(ASRB_E_WORKSPACE, "asrb_attention_forward: workspace NULL, misaligned or too small");
    ASRB_TRY(require_sm100());
    cudaStream_t st = (cudaStream_t)stream;
    const int D = a->dims, H = a->head, hd = D / H;
    const int64_t rows = B * T;
    Arena ar(ws, ws_bytes);
    float* xn = ar.take<float>(rows * D);
    float* q = ar.take<float>(rows * D);
    float* kv = ar.take<float>(rows * 2 * D);
    float* att = ar.take<float>(rows * D);
    const float pre = powf((float)hd, -0.25f);                       // n.scale, model.py:239
    ASRB_TRY(launch_rmsnorm(x, a->q_norm, xn, rows, D, st));
    ASRB_TRY(launch_gemm_simt(xn, DT_F32, a->q_w, a->q_b, nullptr, q, DT_F32, B, T, D, D, 1, ACT_NONE, st));
    ASRB_TRY(launch_rmsnorm(x, a->kv_norm, xn, rows, D, st));
    ASRB_TRY(launch_gemm_simt(xn, DT_F32, a->kv_w, a->kv_b, nullptr, kv, DT_F32, B, T, D, 2 * D, 1, ACT_NONE, st));
    ASRB_TRY(launch_rotary_headnorm(q, D, x, a->ln_w, a->freqs, B, T, D, H, pre, st));       // model.py:303-307
    ASRB_TRY(launch_rotary_headnorm(kv, 2 * D, x, a->ln_w, a->freqs, B, T, D, H, pre, st));  // k = first D columns
    ASRB_TRY(launch_attention_simt_ex(q, kv, kv + D, D, 2 * D, 2 * D, att, DT_F32, B, T, D, H, 1.0f / sqrtf((float)hd), st));
    return launch_gemm_simt(att, DT_F32, a->out_w, a->out_b, nullptr, out, DT_F32, B, T, D, D, 1, ACT_NONE, st);
}
