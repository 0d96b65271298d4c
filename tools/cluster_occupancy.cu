// How many clusters of the fused kernel can be co-resident on this GPU for cluster sizes 2/4/8
// (tuning aid: a 4-CTA cluster must sit inside one GPC, so some SMs may stay idle).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr \
//        -I zero-shot-aac_b200/csrc -o tools/cluster_occupancy tools/cluster_occupancy.cu
#include <cstdio>
#include "simtopk_kernel.cuh"

int main() {
  auto kern = zs::zs_simtopk_kernel<16, 2, zs::MODE_TOPK>;
  const int smem = zs::smem_bytes<2>();
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device %s, %d SMs, kernel smem %d\n", prop.name, prop.multiProcessorCount, smem);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
    cfg.blockDim = dim3(zs::NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
