// tma_probe.cu — which cp.async.bulk.tensor configurations does this driver / GPU accept? (diagnostic for sobel5_tma_kernel)
// usage: tma_probe <rank 2|3> <boxw> <x0> <y0> <l2promo 0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int bytes, unsigned* out) {
  extern __shared__ __align__(128) uint8_t t[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), t_a = (uint32_t)__cvta_generic_to_shared(t);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(bytes));
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                   ::"r"(t_a), "l"(&tmap), "r"(x0), "r"(y0), "r"(0), "r"(bar_a) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                   ::"r"(t_a), "l"(&tmap), "r"(x0), "r"(y0), "r"(bar_a) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar_a) : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += t[i];
  atomicAdd(out, s);
}
int main(int argc, char** argv) {
  int rank = atoi(argv[1]), boxw = atoi(argv[2]), x0 = atoi(argv[3]), y0 = atoi(argv[4]), l2 = atoi(argv[5]);
  const int W = 320, H = 240, BH = 36;
  uint8_t* d; cudaMalloc(&d, W * H * 2); cudaMemset(d, 1, W * H * 2);
  unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[3] = {W, H, 2}, strides[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)boxw, BH, 1}, es[3] = {1, 1, 1};
  CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d; ", (int)r);
  if (rank == 3) k<3><<<1, 128, boxw * BH>>>(tm, x0, y0, boxw * BH, out); else k<2><<<1, 128, boxw * BH>>>(tm, x0, y0, boxw * BH, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("rank %d boxw %d x0 %d y0 %d l2 %d -> %s sum %u\n", rank, boxw, x0, y0, l2, cudaGetErrorString(e), h);
  return 0;
}
