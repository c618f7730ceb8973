// dev test: 3-D byte-tensor TMA box load (the FAST patch fetch) in isolation
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int x0, int y0, int z, int box_h, uint8_t *out)
{
  __shared__ __align__(128) uint8_t s[70 * 80];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(80 * box_h) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(s)),
                 "l"(&map), "r"(x0), "r"(y0), "r"(z), "r"(smem_u32(&bar))
                 : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(&bar)), "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < 80 * box_h; i += blockDim.x) out[i] = s[i];
}
int main()
{
  const int pitch = 1248, h = 376, nimg = 4, box_h = 45;
  const size_t img_stride = 1464832 / 256 * 256; // any multiple of 256 >= pitch*h
  std::vector<uint8_t> hbuf(img_stride * nimg);
  for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (uint8_t)(i * 2654435761u >> 24);
  uint8_t *d, *o;
  cudaMalloc(&d, hbuf.size());
  cudaMalloc(&o, 80 * 70);
  cudaMemcpy(d, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)nimg};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
  const cuuint32_t box[3] = {80u, (cuuint32_t)box_h, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  for (int x0 : {16, 17, 1190})
  {
    int y0 = 47, z = 2;
    k<<<1, 128>>>(map, x0, y0, z, box_h, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("x0=%d: %s\n", x0, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> res(80 * box_h);
    cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < box_h; ++y)
      for (int x = 0; x < 80; ++x)
      {
        uint8_t e8 = (x0 + x < pitch && y0 + y < h) ? hbuf[(size_t)z * img_stride + (size_t)(y0 + y) * pitch + x0 + x] : 0;
        bad += res[y * 80 + x] != e8;
      }
    printf("  mismatches %d\n", bad);
  }
}
