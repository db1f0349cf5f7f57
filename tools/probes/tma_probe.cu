// Probe: which tensor-map layouts does the TMA unit accept for fp32 NCHW viewed as (w, n, c%8, h, c/8)?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda ; ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int nfloat, int c0, int c1, int c2, int c3, int c4) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nfloat * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(smem)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 4, C = 16, H = (variant >= 8 ? 16 : 32), W = (variant >= 8 ? (variant >= 10 ? 8 : 16) : 32), HW = H * W;
  std::vector<float> hx((size_t)B * C * HW);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)i;      // value = linear index (exact below 2^24)
  float* x; cudaMalloc(&x, hx.size() * 4);
  cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)B, 8, (cuuint64_t)H, (cuuint64_t)C / 8};
  cuuint64_t strides[4] = {(cuuint64_t)C * HW * 4, (cuuint64_t)HW * 4, (cuuint64_t)W * 4, (cuuint64_t)8 * HW * 4};
  cuuint32_t box[5] = {32, 1, 8, 6, 2};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  int c[5] = {-1, 1, 0, -1, 0};
  if (variant == 1) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (variant == 2) sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (variant == 3) { c[0] = 0; c[3] = 0; }                       // no negative coordinates
  if (variant == 4) {                                              // monotone strides: (w, h, c%8, c/8, n)
    dims[1] = H; dims[2] = 8; dims[3] = C / 8; dims[4] = B;
    strides[0] = W * 4; strides[1] = HW * 4; strides[2] = 8 * HW * 4; strides[3] = (cuuint64_t)C * HW * 4;
    box[1] = 6; box[2] = 8; box[3] = 2; box[4] = 1;
    c[1] = -1; c[2] = 0; c[3] = 0; c[4] = 1;
  }
  if (variant == 5) { sw = CU_TENSOR_MAP_SWIZZLE_128B; c[0] = 0; c[3] = 0; }
  if (variant == 6) { c[0] = 0; }                                  // only h negative
  if (variant == 7) { c[3] = 0; }                                  // only w negative
  if (variant == 8 || variant == 9) { box[0] = 16; box[1] = 2; c[0] = 0; c[1] = 0; c[3] = 0; if (variant == 9) sw = CU_TENSOR_MAP_SWIZZLE_NONE; }
  if (variant == 10) { box[0] = 8; box[1] = 4; c[0] = 0; c[1] = 0; c[3] = 0; }
  CUtensorMap tm;
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode=%d\n", variant, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int nfloat = box[0] * box[1] * box[2] * box[3] * box[4];
  float* out; cudaMalloc(&out, nfloat * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  probe<<<1, 128, 65536 + 64>>>(tm, out, nfloat, c[0], c[1], c[2], c[3], c[4]);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d run=%s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<float> ho(nfloat);
  cudaMemcpy(ho.data(), out, nfloat * 4, cudaMemcpyDeviceToHost);
  // print the first channel-group atom of row 1 (second row of the box): 8 channel rows x 32 floats, as (c, h, w) decoded
  for (int row = 0; row < 2; ++row) {
    printf("smem atom %d:\n", row);
    for (int cr = 0; cr < 8; ++cr) {
      printf(" c-row %d:", cr);
      for (int q = 0; q < 32; q += (variant >= 8 ? 2 : 4)) {
        const float v = ho[(row * 8 + cr) * 32 + q];
        const long li = (long)v; const int w = li % W, h = (li / W) % H, cc = (li / HW) % C, n = li / ((long)C * HW);
        if (v == 0.f) printf(" [   zero  ]"); else printf(" [n%d c%2d h%2d w%2d]", n, cc, h, w);
      }
      printf("\n");
    }
  }
  return 0;
}
