// Probe: how many thread-block clusters of size c (1 CTA per SM: 221 KB dynamic shared memory, 512 threads)
// can be co-resident on this GPU?  nvcc -arch=sm_100a cluster_occ.cu -o cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  const size_t smem = 221 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c = 1; c <= 8; ++c) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c * 64, 1, 1);
    cfg.blockDim = dim3(512, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %d: max active clusters %d (%d CTAs)  %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
