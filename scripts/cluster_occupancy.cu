// How many thread-block clusters of a tensor-core kernel's shape (576 threads, ~197 KB dynamic shared memory, one CTA per SM)
// can be resident at once on this GPU, per cluster size.  nvcc -arch=sm_100a -o /tmp/cluster_occ scripts/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k_dummy(float* p) {
  extern __shared__ float sm[];
  if (p) p[0] = sm[threadIdx.x];
}
int main() {
  const int smem = 197 * 1024 + 1024;
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 32, 1, 1);
    cfg.blockDim = dim3(576, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (= %d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
