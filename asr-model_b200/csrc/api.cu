// Process-wide pieces of the C ABI: version, error string, device gate.
#include "common.cuh"
#include <vector>

namespace asrb {

int require_sm100() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(ASRB_E_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail(ASRB_E_DEVICE, "cannot query device %d: %s", dev, cudaGetErrorString(e));
    if (major != 10)
        return fail(ASRB_E_DEVICE, "device %d has compute capability %d.x; libasrb200 is sm_100a only "
                                   "(no fallback path exists)", dev, major);
    return ASRB_OK;
}

int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// ---- profiler state (process-wide, debugging instrumentation only) ----
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
bool prof_on() { return g_prof_on; }
void prof_push(const char* tag, cudaStream_t st, double flops, double bytes) {
    ProfRec r{tag, nullptr, nullptr, flops, bytes};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { g_prof_on = false; return; }
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
}
void prof_pop(cudaStream_t st) { if (!g_recs.empty()) cudaEventRecord(g_recs.back().e1, st); }

}  // namespace asrb

extern "C" int asrb_profile_begin(void) {
    for (auto& r : asrb::g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    asrb::g_recs.clear();
    asrb::g_prof_on = true;
    return ASRB_OK;
}
extern "C" int asrb_profile_end(void) {
    asrb::g_prof_on = false;
    if (!asrb::g_recs.empty()) {
        cudaError_t e = cudaEventSynchronize(asrb::g_recs.back().e1);
        if (e != cudaSuccess) return asrb::fail(ASRB_E_CUDA, "asrb_profile_end: %s", cudaGetErrorString(e));
    }
    return (int)asrb::g_recs.size();
}
extern "C" int asrb_profile_get(int i, const char** tag, float* ms, double* flops, double* bytes) {
    if (i < 0 || i >= (int)asrb::g_recs.size()) return asrb::fail(ASRB_E_ARG, "asrb_profile_get: index %d out of range", i);
    const asrb::ProfRec& r = asrb::g_recs[i];
    float t = 0.f;
    cudaError_t e = cudaEventElapsedTime(&t, r.e0, r.e1);
    if (e != cudaSuccess) return asrb::fail(ASRB_E_CUDA, "asrb_profile_get: %s", cudaGetErrorString(e));
    if (tag) *tag = r.tag; if (ms) *ms = t; if (flops) *flops = r.flops; if (bytes) *bytes = r.bytes;
    return ASRB_OK;
}

extern "C" int asrb_version(void) { return ASRB_VERSION; }
extern "C" int asrb_operand_format(void) { return ASRB_OP16_IS_F16 ? ASRB_F16 : ASRB_BF16; }
extern "C" const char* asrb_last_error(void) { return asrb::err_slot().c_str(); }
extern "C" int asrb_device_check(int device) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) return asrb::fail(ASRB_E_DEVICE, "no CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return asrb::fail(ASRB_E_DEVICE, "cannot select device %d", device);
    int r = asrb::require_sm100();
    cudaSetDevice(prev);
    return r;
}
