// Backward of the quantized conv for the layer kinds the tensor-core kernels do not take directly
// (autograd of models/quantized_conv.py:36 for models/resnet.py's stride-2 layers and
// models/mobilenet.py:64-74,106-116's depthwise layers):
//
//   * stride 2, dense: both gradients of conv(x, W, stride 2) equal the gradients of the STRIDE-1 conv whose
//     output gradient is the zero-inserted one, g_up[2p][2q] = g[p][q]:
//         gx = conv_transpose(g, W, 2) = dgrad_stride1(g_up, W),   gw = wgrad_stride1(g_up, x)
//     so one small kernel (dilate2_kernel) puts these layers on the TMA-fed tcgen05 kernels (K3T / K5T).
//   * depthwise 3x3 pad 1 (groups == C == K): data gradient = the depthwise forward kernel's arithmetic with
//     the filter rotated by 180 degrees (on g, or on g_up for stride 2); weight gradient = nine per-channel
//     inner products over the batch, a two-stage fixed-order reduction (deterministic).
#include "po2_common.cuh"

namespace po2 {

// g (planes, P, Q) -> g_up (planes, 2P, 2Q): g_up[2p][2q] = g[p][q], zero elsewhere.  One thread = 4 output
// columns of an even row pair: reads two inputs, writes two 128-bit vectors.
__global__ void __launch_bounds__(256) dilate2_kernel(const float* __restrict__ g, float* __restrict__ up, int Q2,
                                                      int rows, int Q) {
  // rows = planes * P input rows; Q2 = Q / 2 (pairs of input columns per row)
  const int total = rows * Q2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / Q2, c2 = i - row * Q2;
    const float2 v = __ldg(reinterpret_cast<const float2*>(g + (size_t)row * Q) + c2);
    float4* o = reinterpret_cast<float4*>(up + (size_t)row * 4 * Q) + c2;        // output row 2*row has 2Q floats
    o[0] = make_float4(v.x, 0.f, v.y, 0.f);
    o[Q2] = make_float4(0.f, 0.f, 0.f, 0.f);                                      // output row 2*row + 1
  }
}

// depthwise 3x3 pad 1 stride 1 data gradient on g (already zero-inserted for stride 2):
//   gx[h][w] = sum_{r,s} g[h + 1 - r][w + 1 - s] * w[c][r][s]
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                       float* __restrict__ gx, int C, int H, int W, int total) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int x = i % W;
    const int t = i / W;
    const int y = t % H;
    const int plane = t / H;
    const int c = plane % C;
    const float* pg = g + (size_t)plane * H * W;
    const float* pw = w + c * 9;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int gy = y + 1 - r;
      if (gy < 0 || gy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int gxx = x + 1 - s;
        if (gxx < 0 || gxx >= W) continue;
        acc = fmaf(__ldg(pg + gy * W + gxx), __ldg(pw + r * 3 + s), acc);
      }
    }
    gx[i] = acc;
  }
}

// depthwise 3x3 pad 1 weight gradient (stride-1 form; g zero-inserted for stride 2):
//   gw[c][r][s] = sum_{n,h,w} g[n][c][h][w] * x[n][c][h + r - 1][w + s - 1]
// CTA (split, c) sums the images n = split, split + S, ...; partials [c][split][9]; the last CTA of a channel
// (ticket) adds them in split order.
constexpr int DWG_THREADS = 256;
__global__ void __launch_bounds__(DWG_THREADS) dw_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                               float* __restrict__ gw, float* __restrict__ partial,
                                                               unsigned int* __restrict__ ticket, int B, int C, int H,
                                                               int W, int S) {
  __shared__ float sm[DWG_THREADS / 32][9];
  __shared__ int last;
  const int split = blockIdx.x, c = blockIdx.y, HW = H * W;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int n = split; n < B; n += S) {
    const float* pg = g + ((size_t)n * C + c) * HW;
    const float* px = x + ((size_t)n * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += DWG_THREADS) {
      const float gv = __ldg(pg + i);
      if (gv == 0.f) continue;                                  // zero-inserted gradients: 3 of 4 positions
      const int y = i / W, xx = i - y * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = y + r - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = xx + s - 1;
          if (ix < 0 || ix >= W) continue;
          acc[r * 3 + s] = fmaf(gv, __ldg(px + iy * W + ix), acc[r * 3 + s]);
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < DWG_THREADS / 32; ++q) v += sm[q][threadIdx.x];
    partial[((size_t)c * S + split) * 9 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket + c, 1u);
    last = (t == (unsigned int)(S - 1));
    if (last) ticket[c] = 0u;                                   // self-resetting
  }
  __syncthreads();
  if (last && threadIdx.x < 9) {
    __threadfence();
    float v = 0.f;
    for (int q = 0; q < S; ++q) v += __ldcg(partial + ((size_t)c * S + q) * 9 + threadIdx.x);
    gw[c * 9 + threadIdx.x] = v;
  }
}

constexpr int DWG_MAX_C = 4096;
constexpr int DWG_MAX_SPLIT = 32;

}  // namespace po2

using namespace po2;

extern "C" {

int po2_dilate2(const void* g, void* g_up, int planes, int P, int Q, void* stream) {
  if (!g || !g_up) return PO2_E_NULL;
  if (planes <= 0 || P <= 0 || Q <= 0) return PO2_E_SHAPE;
  if (Q % 2 || (reinterpret_cast<uintptr_t>(g) & 7) || (reinterpret_cast<uintptr_t>(g_up) & 15)) return PO2_E_UNSUPPORTED;
  if ((int64_t)planes * P * Q * 4 >= (1ll << 31)) return PO2_E_SIZE;
  const int rows = planes * P, total = rows * (Q / 2);
  const int blocks = (total + 255) / 256 < device_sm_count() * 16 ? (total + 255) / 256 : device_sm_count() * 16;
  dilate2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)g, (float*)g_up, Q / 2, rows, Q);
  return (int)cudaGetLastError();
}

// g: (B, C, H, W) -- the output gradient at the INPUT resolution (stride 2: zero-inserted with po2_dilate2)
int po2_conv2d_depthwise_dgrad(const void* g, const void* w, void* gx, int B, int C, int H, int W, void* stream) {
  if (!g || !w || !gx) return PO2_E_NULL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return PO2_E_SHAPE;
  if ((int64_t)B * C * H * W >= (1ll << 31)) return PO2_E_SIZE;
  const int total = B * C * H * W;
  const int blocks = (total + 255) / 256 < device_sm_count() * 16 ? (total + 255) / 256 : device_sm_count() * 16;
  dw_dgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)g, (const float*)w, (float*)gx, C, H, W, total);
  return (int)cudaGetLastError();
}

size_t po2_conv2d_depthwise_wgrad_workspace(int C) {
  return (size_t)DWG_MAX_C * sizeof(unsigned int) + (size_t)C * DWG_MAX_SPLIT * 9 * sizeof(float);
}

// workspace: po2_conv2d_depthwise_wgrad_workspace(C) bytes whose first DWG_MAX_C * 4 bytes (the tickets) are zero
// before the first call (the kernel leaves them zero)
int po2_conv2d_depthwise_wgrad(const void* g, const void* x, void* gw, int B, int C, int H, int W, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!g || !x || !gw || !workspace) return PO2_E_NULL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return PO2_E_SHAPE;
  if (C > DWG_MAX_C) return PO2_E_UNSUPPORTED;
  if (workspace_bytes < po2_conv2d_depthwise_wgrad_workspace(C)) return PO2_E_WORKSPACE;
  if ((int64_t)B * C * H * W >= (1ll << 31)) return PO2_E_SIZE;
  int S = (4 * device_sm_count() + C - 1) / C;
  if (S > B) S = B;
  if (S > DWG_MAX_SPLIT) S = DWG_MAX_SPLIT;
  if (S < 1) S = 1;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)DWG_MAX_C * sizeof(unsigned int));
  dw_wgrad_kernel<<<dim3(S, C), DWG_THREADS, 0, (cudaStream_t)stream>>>((const float*)g, (const float*)x, (float*)gw, partial,
                                                                         ticket, B, C, H, W, S);
  return (int)cudaGetLastError();
}

}  // extern "C"
