// Process-wide pieces of the C ABI: version, error string, device gate.
#include "common.cuh"

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

}  // namespace asrb

extern "C" int asrb_version(void) { return ASRB_VERSION; }
extern "C" const char* asrb_last_error(void) { return asrb::err_slot().c_str(); }
extern "C" int asrb_device_check(int device) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) return asrb::fail(ASRB_E_DEVICE, "no CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return asrb::fail(ASRB_E_DEVICE, "cannot select device %d", device);
    int r = asrb::require_sm100();
    cudaSetDevice(prev);
    return r;
}
