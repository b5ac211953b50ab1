// How many thread-block clusters of size 1/2/4/8 (576 threads, ~210 KB dynamic smem per CTA: the footprint of the
// persistent tcgen05 GEMM kernels) can be co-resident on this GPU?  Answers whether 148 SMs tile into 4-clusters.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) dummy(int* p) { extern __shared__ unsigned char s[]; if (p) p[0] = s[0]; }
int main() {
    const int smem = 210 * 1024;
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
