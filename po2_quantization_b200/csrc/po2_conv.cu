// Quantized-conv forward for B200 (sm_100a): out = conv2d(x, W_q), x/out fp32 NCHW, bias=None, dilation=1
// (reference models/quantized_conv.py:36,38 -> nn.Conv2d._conv_forward -> F.conv2d).
//
// Three kernels behind po2_conv2d_fwd:
//   K3  conv_umma_kernel     dense convs on the 5th-gen tensor cores: an implicit GEMM whose A operand is
//                            the activation strip converted fp32 NCHW -> bf16 *in shared memory* in a
//                            "flat padded" K-major layout, so every filter tap is the SAME smem tile
//                            read through a UMMA descriptor whose start address is shifted by
//                            (r*pitch + s) pixels (no im2col copy); B is the PO2 weight tensor decoded
//                            to exact bf16 (+-2^q) and pre-packed per tap; accumulators live in TMEM
//                            (tcgen05.mma, one issuing thread) and are read back with tcgen05.ld for
//                            the epilogue, which applies the per-tensor scale and stores NCHW fp32.
//   K4  conv_depthwise_kernel groups == C == K, 3x3: CUDA-core, HBM/L2-bound.
//   --  conv_direct_kernel    any dense/grouped shape in fp32 FMA (the "fp32-accumulate" parity path and
//                            the fallback for shapes K3 does not take).
#include "po2_common.cuh"

namespace po2 {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers (tcgen05 / mbarrier / bulk copy)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// contiguous global -> shared bulk copy (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 8-element K halves) |
//   [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0 (none)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// geometry shared by host and device
// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int B, C, H, W, K, R, S, stride, pad, groups;
  int P, Q;            // output height / width
  // K3 flat layout (stride 1): one zero column per row and one zero row per image are shared pads
  int pitch;           // W + (S == 3)
  int rows_img;        // H + (R == 3)
  int halo;            // (R==3)*pitch + (S==3)   flat positions of context before / after a tile
  int Ltot;            // flat positions that can hold an output
  int Cpad, CC, nchunk;  // channels padded to 16, channels per K chunk, chunks
  int NT, ntiles_n;    // padded out-channel tile (mult of 16, <= 256), number of N tiles
  int MT;              // 128-row M tiles per CTA
  int strip;           // smem positions per CTA = MT*128 + 2*halo
};

// ------------------------------------------------------------------------------------------------
// weight pack: fp32 PO2-grid weights (or codes) -> exact bf16 +-2^q in the B-operand layout
//   Bp[nt][chunk][tap][cg][n][8]   (cg: group of 8 channels inside the chunk, n: channel in N tile)
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, const uint8_t* __restrict__ codes,
                                    const float* __restrict__ scale, __nv_bfloat16* __restrict__ Bp,
                                    ConvGeom g, int bits, int fsr) {
  const int taps = g.R * g.S;
  const int64_t total = (int64_t)g.ntiles_n * g.nchunk * taps * (g.CC / 8) * g.NT * 8;
  const float s = scale ? *scale : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int j = (int)(t % 8); t /= 8;
    const int n = (int)(t % g.NT); t /= g.NT;
    const int cg = (int)(t % (g.CC / 8)); t /= (g.CC / 8);
    const int tap = (int)(t % taps); t /= taps;
    const int chunk = (int)(t % g.nchunk); t /= g.nchunk;
    const int nt = (int)t;
    const int c = chunk * g.CC + cg * 8 + j;
    const int k = nt * g.NT + n;
    float v = 0.0f;
    if (c < g.C && k < g.K) {
      const int64_t wi = ((int64_t)k * g.C + c) * taps + tap;
      if (codes) {
        const uint32_t code = (bits <= 4) ? ((codes[wi >> 1] >> ((wi & 1) * 4)) & 0xFu) : codes[wi];
        const int mag = code & ((1u << (bits - 1)) - 1u);
        v = exp2_int((fsr - 1) - mag);
        if ((code >> (bits - 1)) & 1u) v = -v;
      } else {
        // w = +-s*2^q exactly, so w/s is exactly +-2^q and the bf16 conversion is lossless; a weight
        // that is not on the grid (unquantized layer) is rounded to bf16 like any bf16 GEMM would
        v = (s == 1.0f) ? w[wi] : __fdiv_rn(w[wi], s);
      }
    }
    Bp[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// K3: tcgen05 implicit-GEMM conv (stride 1; 3x3 pad 1 or 1x1 pad 0; groups == 1)
// ------------------------------------------------------------------------------------------------
constexpr int K3_THREADS = 256;

__device__ __forceinline__ bool decode_pos(const ConvGeom& g, int L, int& img, int& h, int& w) {
  // flat position -> (image, row, col); false for the shared zero pads and out-of-range positions
  if (L < 0 || L >= g.Ltot) return false;
  const int row = L / g.pitch;
  w = L - row * g.pitch;
  const int r0 = row - (g.R == 3 ? 1 : 0);        // first flat row is the top pad of image 0
  if (r0 < 0) return false;
  img = r0 / g.rows_img;
  h = r0 - img * g.rows_img;
  return (w < g.W) && (h < g.H) && (img < g.B);
}

__global__ void __launch_bounds__(K3_THREADS) conv_umma_kernel(const float* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ Bp,
                                                               const float* __restrict__ scale,
                                                               float* __restrict__ out, ConvGeom g) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int taps = g.R * g.S;
  const uint32_t a_bytes = (uint32_t)(g.CC / 8) * g.strip * 16;
  const uint32_t b_bytes = (uint32_t)taps * (g.CC / 8) * g.NT * 16;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + a_bytes + b_bytes);
  uint64_t* bar_mma = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.y;
  const int L0 = blockIdx.x * g.MT * 128;                 // first output position of this CTA
  uint32_t ncols = 32;
  while ((int)ncols < g.MT * g.NT) ncols <<= 1;

  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  if (tid == 32) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc((uint32_t)g.NT);
  const int HW = g.H * g.W;

  for (int chunk = 0; chunk < g.nchunk; ++chunk) {
    if (chunk > 0) mbar_wait(bar_mma, (chunk - 1) & 1);   // previous chunk's MMAs have consumed smem
    if (tid == 0) {
      mbar_expect_tx(bar_b, b_bytes);
      bulk_g2s(sB, Bp + ((int64_t)(nt * g.nchunk + chunk) * b_bytes) / 2, b_bytes, bar_b);
    }
    // ---- A: fp32 NCHW -> bf16 [channel group][flat position][8 channels], zero pads materialised
    const int ngrp = g.CC / 8;
    const int items = ngrp * g.strip;
    for (int it = tid; it < items; it += K3_THREADS) {
      const int grp = it / g.strip;
      const int lloc = it - grp * g.strip;
      int img, h, w;
      uint4 packed = make_uint4(0u, 0u, 0u, 0u);
      if (decode_pos(g, L0 - g.halo + lloc, img, h, w)) {
        const int c0 = chunk * g.CC + grp * 8;
        const float* px = x + ((int64_t)img * g.C + c0) * HW + h * g.W + w;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c0 + j < g.C) ? __ldg(px + (int64_t)j * HW) : 0.0f;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        packed = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                            *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
      }
      *reinterpret_cast<uint4*>(sA + ((size_t)grp * g.strip + lloc) * 16) = packed;
    }
    fence_proxy_async();                                    // generic-proxy smem writes -> tensor core
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_b, chunk & 1);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const uint32_t a_plane = (uint32_t)g.strip * 16, b_plane = (uint32_t)g.NT * 16;
      for (int mt = 0; mt < g.MT; ++mt) {
        const uint32_t d = tmem_base + (uint32_t)(mt * g.NT);
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / g.S, s = tap - r * g.S;
          const uint32_t a_tap = a0 + (uint32_t)(mt * 128 + r * g.pitch + s) * 16;
          const uint32_t b_tap = b0 + (uint32_t)tap * ngrp * b_plane;
          for (int ks = 0; ks < g.CC / 16; ++ks) {
            const uint64_t ad = make_desc(a_tap + (uint32_t)ks * 2 * a_plane, a_plane, 128);
            const uint64_t bd = make_desc(b_tap + (uint32_t)ks * 2 * b_plane, b_plane, 128);
            umma_bf16(d, ad, bd, idesc, (chunk | tap | ks) != 0);
          }
        }
      }
      umma_commit(bar_mma);                                 // implies tcgen05.fence::before_thread_sync
    }
  }
  // ---- epilogue: TMEM -> registers -> scale -> NCHW fp32 (coalesced along w per out channel)
  mbar_wait(bar_mma, (g.nchunk - 1) & 1);
  tc_fence_after();
  const float sc = scale ? *scale : 1.0f;
  const int q4 = warp & 3, half = warp >> 2;                // lane quarter, column half
  const int nchunks16 = g.NT / 16;
  for (int mt = 0; mt < g.MT; ++mt) {
    const int L = L0 + mt * 128 + q4 * 32 + lane;
    int img = 0, h = 0, w = 0;
    const bool valid = decode_pos(g, L, img, h, w);
    for (int cb = half; cb < nchunks16; cb += 2) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(mt * g.NT + cb * 16), r);
      if (valid) {
        const int k0 = nt * g.NT + cb * 16;
        float* po = out + ((int64_t)img * g.K + k0) * HW + h * g.W + w;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (k0 + j < g.K) po[(int64_t)j * HW] = __uint_as_float(r[j]) * sc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// K4: depthwise 3x3 (groups == C == K), any stride/pad; one thread per output element
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_depthwise_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             float* __restrict__ out, ConvGeom g) {
  const int64_t total = (int64_t)g.B * g.C * g.P * g.Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % g.Q);
    int64_t t = i / g.Q;
    const int p = (int)(t % g.P); t /= g.P;
    const int c = (int)(t % g.C);
    const int n = (int)(t / g.C);
    const float* px = x + ((int64_t)n * g.C + c) * g.H * g.W;
    const float* pw = w + (int64_t)c * g.R * g.S;
    float acc = 0.0f;
    for (int r = 0; r < g.R; ++r) {
      const int ih = p * g.stride - g.pad + r;
      if (ih < 0 || ih >= g.H) continue;
      for (int s = 0; s < g.S; ++s) {
        const int iw = q * g.stride - g.pad + s;
        if (iw < 0 || iw >= g.W) continue;
        acc = fmaf(__ldg(px + ih * g.W + iw), __ldg(pw + r * g.S + s), acc);
      }
    }
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// direct fp32 conv: any groups/stride/pad/kernel.  One thread: one output pixel x DK out channels.
// ------------------------------------------------------------------------------------------------
constexpr int DK = 8;
__global__ void __launch_bounds__(128) conv_direct_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ w,
                                                          float* __restrict__ out, ConvGeom g) {
  extern __shared__ float sw[];                             // [DK][Cg*R*S] weights of this k block
  const int Cg = g.C / g.groups, Kg = g.K / g.groups;
  const int kblocks_g = (Kg + DK - 1) / DK;
  const int kb = blockIdx.y;                                // (group, k block)
  const int grp = kb / kblocks_g;
  const int k0 = grp * Kg + (kb - grp * kblocks_g) * DK;
  const int kend = min(k0 + DK, (grp + 1) * Kg);
  const int wsz = Cg * g.R * g.S;
  for (int i = threadIdx.x; i < DK * wsz; i += blockDim.x) {
    const int kk = i / wsz;
    sw[i] = (k0 + kk < kend) ? w[(int64_t)(k0 + kk) * wsz + (i - kk * wsz)] : 0.0f;
  }
  __syncthreads();
  const int64_t npix = (int64_t)g.B * g.P * g.Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % g.Q);
    const int64_t t = i / g.Q;
    const int p = (int)(t % g.P);
    const int n = (int)(t / g.P);
    float acc[DK];
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) acc[kk] = 0.0f;
    for (int c = 0; c < Cg; ++c) {
      const float* px = x + ((int64_t)n * g.C + grp * Cg + c) * g.H * g.W;
      for (int r = 0; r < g.R; ++r) {
        const int ih = p * g.stride - g.pad + r;
        if (ih < 0 || ih >= g.H) continue;
        for (int s = 0; s < g.S; ++s) {
          const int iw = q * g.stride - g.pad + s;
          if (iw < 0 || iw >= g.W) continue;
          const float v = __ldg(px + ih * g.W + iw);
          const float* pw = sw + (c * g.R + r) * g.S + s;
#pragma unroll
          for (int kk = 0; kk < DK; ++kk) acc[kk] = fmaf(v, pw[kk * wsz], acc[kk]);
        }
      }
    }
    for (int kk = 0; kk < DK && k0 + kk < kend; ++kk)
      out[(((int64_t)n * g.K + k0 + kk) * g.P + p) * g.Q + q] = acc[kk];
  }
}

// codes -> fp32 weights for the CUDA-core kernels
__global__ void decode_weights_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ scale,
                                      float* __restrict__ w, int64_t n, int bits, int fsr) {
  const float s = *scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t code = (bits <= 4) ? ((codes[i >> 1] >> ((i & 1) * 4)) & 0xFu) : codes[i];
    const int mag = code & ((1u << (bits - 1)) - 1u);
    float v = __fmul_rn(exp2_int((fsr - 1) - mag), s);
    w[i] = ((code >> (bits - 1)) & 1u) ? -v : v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_sms = 0;
static int sm_count() {
  if (!g_sms) {
    int d = 0, n = 0;
    if (cudaGetDevice(&d) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) == cudaSuccess) g_sms = n;
  }
  return g_sms > 0 ? g_sms : 148;
}

static bool fill_geom(ConvGeom& g, int B, int C, int H, int W, int K, int R, int S, int stride, int pad, int groups) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || R <= 0 || S <= 0 || stride <= 0 || pad < 0 || groups <= 0) return false;
  if (C % groups || K % groups) return false;
  g = ConvGeom{};
  g.B = B; g.C = C; g.H = H; g.W = W; g.K = K; g.R = R; g.S = S; g.stride = stride; g.pad = pad; g.groups = groups;
  g.P = (H + 2 * pad - R) / stride + 1;
  g.Q = (W + 2 * pad - S) / stride + 1;
  return g.P > 0 && g.Q > 0;
}

// K3 takes: dense, stride 1, square 3x3 pad 1 or 1x1 pad 0
static bool umma_eligible(const ConvGeom& g) {
  if (g.groups != 1 || g.stride != 1) return false;
  if (!((g.R == 3 && g.S == 3 && g.pad == 1) || (g.R == 1 && g.S == 1 && g.pad == 0))) return false;
  return true;
}

static size_t umma_smem_bytes(const ConvGeom& g) {
  return (size_t)(g.CC / 8) * g.strip * 16 + (size_t)g.R * g.S * (g.CC / 8) * g.NT * 16 + 64;
}

static void plan_umma(ConvGeom& g) {
  const bool k3 = (g.R == 3);
  g.pitch = g.W + (k3 ? 1 : 0);
  g.rows_img = g.H + (k3 ? 1 : 0);
  g.halo = k3 ? g.pitch + 1 : 0;
  g.Ltot = (g.B * g.rows_img + (k3 ? 1 : 0)) * g.pitch;
  g.Cpad = (g.C + 15) / 16 * 16;
  g.CC = g.Cpad < 64 ? g.Cpad : 64;
  g.nchunk = (g.Cpad + g.CC - 1) / g.CC;
  const int Kp = (g.K + 15) / 16 * 16;
  g.ntiles_n = (Kp + 255) / 256;
  g.NT = ((Kp + g.ntiles_n - 1) / g.ntiles_n + 15) / 16 * 16;
  // M tiles per CTA: as many as TMEM (512 cols) and smem allow while still giving every SM >= 2 CTAs
  const int mtiles = (g.Ltot + 127) / 128;
  int MT = 4;
  while (MT > 1 && (MT * g.NT > 512 || (mtiles + MT - 1) / MT * g.ntiles_n < 2 * sm_count())) MT >>= 1;
  g.MT = MT;
  g.strip = g.MT * 128 + 2 * g.halo;
  while (g.MT > 1 && umma_smem_bytes(g) > 200 * 1024) { g.MT >>= 1; g.strip = g.MT * 128 + 2 * g.halo; }
}

static size_t umma_pack_bytes(const ConvGeom& g) {
  return (size_t)g.ntiles_n * g.nchunk * g.R * g.S * (g.CC / 8) * g.NT * 16;
}

}  // namespace po2

using namespace po2;

extern "C" {

size_t po2_conv2d_workspace(int B, int C, int H, int W, int K, int R, int S, int stride, int pad,
                            int groups, int compute) {
  ConvGeom g;
  if (!fill_geom(g, B, C, H, W, K, R, S, stride, pad, groups)) return 0;
  size_t bytes = (size_t)K * (C / groups) * R * S * sizeof(float);       // decoded fp32 weights (codes input)
  if (compute == 0 && umma_eligible(g)) {
    plan_umma(g);
    if (umma_smem_bytes(g) <= 200 * 1024) bytes += umma_pack_bytes(g) + 256;
  }
  return (bytes + 255) / 256 * 256;
}

int po2_conv2d_fwd(const void* x, const void* w, const float* scale, void* out, int B, int C, int H,
                   int W, int K, int R, int S, int stride, int pad, int groups, int w_format,
                   int bits, int fsr, int compute, void* workspace, size_t workspace_bytes,
                   void* stream) {
  if (!x || !w || !out) return PO2_E_NULL;
  ConvGeom g;
  if (!fill_geom(g, B, C, H, W, K, R, S, stride, pad, groups)) return PO2_E_SHAPE;
  if (w_format != PO2_W_F32_PO2 && w_format != PO2_W_CODES) return PO2_E_UNSUPPORTED;
  if (w_format == PO2_W_CODES && (!scale || bits < 2 || bits > 8)) return PO2_E_BITS;
  if ((int64_t)B * C * H * W >= (1ll << 31) || (int64_t)B * K * g.P * g.Q >= (1ll << 31)) return PO2_E_SIZE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = po2_conv2d_workspace(B, C, H, W, K, R, S, stride, pad, groups, compute);
  const int64_t wn = (int64_t)K * (C / groups) * R * S;
  const size_t wbytes = ((size_t)wn * sizeof(float) + 255) / 256 * 256;

  if (compute == 0 && umma_eligible(g)) {
    plan_umma(g);
    const size_t smem = umma_smem_bytes(g);
    if (smem <= 200 * 1024) {
      if (!workspace || workspace_bytes < need) return PO2_E_WORKSPACE;
      __nv_bfloat16* Bp = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + wbytes);
      const int64_t total = (int64_t)umma_pack_bytes(g) / 2;
      const int pblocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      pack_weights_kernel<<<pblocks, 256, 0, st>>>(
          w_format == PO2_W_CODES ? nullptr : (const float*)w, w_format == PO2_W_CODES ? (const uint8_t*)w : nullptr,
          w_format == PO2_W_CODES ? nullptr : scale, Bp, g, bits, fsr);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
      static bool attr_set = false;
      if (!attr_set) {
        e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 64);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
      }
      const int mtiles = (g.Ltot + 127) / 128;
      dim3 grid((mtiles + g.MT - 1) / g.MT, g.ntiles_n);
      conv_umma_kernel<<<grid, K3_THREADS, smem, st>>>((const float*)x, Bp, scale, (float*)out, g);
      return (int)cudaGetLastError();
    }
  }
  // CUDA-core paths work on fp32 weights
  const float* wf = (const float*)w;
  if (w_format == PO2_W_CODES) {
    if (!workspace || workspace_bytes < wbytes) return PO2_E_WORKSPACE;
    decode_weights_kernel<<<(int)((wn + 255) / 256 < 1184 ? (wn + 255) / 256 : 1184), 256, 0, st>>>(
        (const uint8_t*)w, scale, (float*)workspace, wn, bits, fsr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    wf = (const float*)workspace;
  }
  if (groups == C && groups == K) {
    const int64_t total = (int64_t)B * C * g.P * g.Q;
    const int blocks = (int)((total + 255) / 256 < (int64_t)sm_count() * 16 ? (total + 255) / 256 : (int64_t)sm_count() * 16);
    conv_depthwise_kernel<<<blocks, 256, 0, st>>>((const float*)x, wf, (float*)out, g);
    return (int)cudaGetLastError();
  }
  const int Cg = C / groups, Kg = K / groups;
  const size_t smem = (size_t)DK * Cg * R * S * sizeof(float);
  if (smem > 200 * 1024) return PO2_E_SHAPE;
  static bool attr2 = false;
  if (!attr2) {
    cudaError_t e = cudaFuncSetAttribute(conv_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr2 = true;
  }
  const int64_t npix = (int64_t)B * g.P * g.Q;
  const int kblocks = groups * ((Kg + DK - 1) / DK);
  int bx = (int)((npix + 127) / 128);
  const int cap = sm_count() * 16 / (kblocks < 16 ? kblocks : 16) + 1;
  if (bx > cap) bx = cap;
  conv_direct_kernel<<<dim3(bx, kblocks), 128, smem, st>>>((const float*)x, wf, (float*)out, g);
  return (int)cudaGetLastError();
}

}  // extern "C"
