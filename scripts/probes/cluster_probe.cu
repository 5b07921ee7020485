// How many thread-block clusters of a given size can a B200 hold at once?  (scheduling granularity = GPC)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    for (int smem : {64 * 1024, 110 * 1024, 225 * 1024}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cs : {1, 2, 4, 6, 8, 10, 12, 14, 16}) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("smem %3d KB cluster %2d: max active clusters %d (CTAs %d) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return 0;
}
