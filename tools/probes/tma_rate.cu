// Probe: sustained TMA load rate per SM for the box shapes of the conv / wgrad kernels (148 CTAs, 4 boxes in flight).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu -lcuda ; ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// each CTA loads `iters` rounds of 4 boxes; box b of round i uses coordinates derived from (cta, i, b)
__global__ void rate(const __grid_constant__ CUtensorMap tm, long long* cycles, int bytes, int iters, int mode, int B) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 32768);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int b = 0; b < 4; ++b) {
        const int item = (blockIdx.x + 148 * (i * 4 + b));
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
        if (mode == 0) { c1 = item % B; c3 = ((item / B) % 8) * 4; }                 // (w, n, c%8, h, c/8) box {32,1,8,6,2}
        else if (mode == 1) { c1 = (item * 96) % (B * 16 * 32 - 96); }                 // 2D (32, rows) box {32, 96}
        else if (mode == 2) { c1 = ((item / B) % 8) * 4; c2 = (item % B) * 16; }       // 3D (w, h, plane) box {32, 6, 16}
        else if (mode == 3) { c2 = ((item / B) % 8) * 16; c4 = item % B; }             // (p%8, c%8, p/8, c/8, n) box {8,8,16,2,1}
        else if (mode == 4) { c2 = ((item / B) % 8) * 4; c3 = item % B; }              // (p%32, c%8, p/32, n, c/8) box {32,8,4,1,2}
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + b)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(smem_u32(smem + b * 32768)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar + b)) : "memory");
      }
      for (int b = 0; b < 4; ++b) wait(bar + b, i & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int B = 128, C = 16, H = 32, W = 32, HW = H * W;
  float* x; cudaMalloc(&x, (size_t)B * C * HW * 4);
  cudaMemset(x, 0, (size_t)B * C * HW * 4);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 64);
  const char* names[5] = {"5D (w,n,c%8,h,c/8) box{32,1,8,6,2} SW128_ATOM32B 12KB", "5D-as-2D (32,rows) box{32,96} SW128 12KB",
                          "5D-as-3D (w,h,plane) box{32,6,16} SW128 12KB", "5D (p%8,c%8,p/8,c/8,n) box{8,8,16,2,1} SW32 8KB",
                          "5D (p%32,c%8,p/32,n,c/8) box{32,8,4,1,2} SW128_ATOM32B 8KB"};
  for (int mode = 0; mode < 5; ++mode) {
    cuuint64_t dims[5] = {1, 1, 1, 1, 1}, strides[4] = {16, 16, 16, 16};
    cuuint32_t box[5] = {1, 1, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    int bytes = 0;
    if (mode == 0) {
      dims[0] = W; dims[1] = B; dims[2] = 8; dims[3] = H; dims[4] = C / 8;
      strides[0] = (cuuint64_t)C * HW * 4; strides[1] = HW * 4; strides[2] = W * 4; strides[3] = 8 * HW * 4;
      box[0] = 32; box[1] = 1; box[2] = 8; box[3] = 6; box[4] = 2; sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    } else if (mode == 1) {
      dims[0] = 32; dims[1] = (cuuint64_t)B * C * H; strides[0] = 128; strides[1] = (cuuint64_t)B * C * HW * 4;
      strides[2] = strides[1]; strides[3] = strides[1]; box[0] = 32; box[1] = 96;
    } else if (mode == 2) {
      dims[0] = 32; dims[1] = H; dims[2] = (cuuint64_t)B * C; strides[0] = 128; strides[1] = HW * 4;
      strides[2] = (cuuint64_t)B * C * HW * 4; strides[3] = strides[2]; box[0] = 32; box[1] = 6; box[2] = 16;
    } else if (mode == 3) {
      dims[0] = 8; dims[1] = 8; dims[2] = HW / 8; dims[3] = C / 8; dims[4] = B;
      strides[0] = HW * 4; strides[1] = 32; strides[2] = 8 * HW * 4; strides[3] = (cuuint64_t)C * HW * 4;
      box[0] = 8; box[1] = 8; box[2] = 16; box[3] = 2; box[4] = 1; sw = CU_TENSOR_MAP_SWIZZLE_32B;
    } else {
      dims[0] = 32; dims[1] = 8; dims[2] = HW / 32; dims[3] = B; dims[4] = C / 8;
      strides[0] = HW * 4; strides[1] = 128; strides[2] = (cuuint64_t)C * HW * 4; strides[3] = 8 * HW * 4;
      box[0] = 32; box[1] = 8; box[2] = 4; box[3] = 1; box[4] = 2; sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    }
    bytes = box[0] * box[1] * box[2] * box[3] * box[4] * 4;
    CUtensorMap tm;
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("mode %d encode=%d\n", mode, (int)r); continue; }
    const int iters = 32;
    for (int rep = 0; rep < 2; ++rep) rate<<<148, 32, 4 * 32768 + 64>>>(tm, cyc, bytes, iters, mode, B);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d run=%s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    printf("%-62s: %7.0f cycles per 4 boxes -> %5.1f B/cycle/SM (%d-byte boxes, L2-warm)\n", names[mode], avg / iters,
           4.0 * bytes * iters / avg, bytes);
  }
  return 0;
}
