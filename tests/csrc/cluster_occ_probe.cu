// How many non-portable 16-CTA (and 8-CTA) clusters can be co-resident on this device for a given block size /
// dynamic shared memory?  (cudaOccupancyMaxActiveClusters; decides particles per cluster of the wide
// cluster-resident sweeps.)  nvcc -arch=sm_100a -o cluster_occ_probe cluster_occ_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k512(float *p) { extern __shared__ float s[]; if (p) p[0] = s[threadIdx.x]; }
__global__ void __launch_bounds__(256, 1) k256(float *p) { extern __shared__ float s[]; if (p) p[0] = s[threadIdx.x]; }
template <typename K>
static void probe(K kern, int nt, const char *name) {
    const int smems[] = {64 * 1024, 160 * 1024, 200 * 1024, 227 * 1024};
    for (int C : {8, 16}) {
        for (int sm : smems) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(C * 64); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = sm; cfg.attrs = attr; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            printf("%s C=%2d smem=%3d KB -> max active clusters %d (%s)\n", name, C, sm / 1024, n, cudaGetErrorString(e));
        }
    }
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    probe(k512, 512, "512 threads");
    probe(k256, 256, "256 threads");
    return 0;
}
