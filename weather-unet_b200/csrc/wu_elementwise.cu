// wu_elementwise.cu — the HBM-bound kernels of the cUNet generator hot path: first conv (K = 27),
// last 1x1 conv + tanh, MaxPool2d(2), AdaIN (statistics, style, apply) fused with the bilinear x2
// upsample and dropout, and their backward passes.  All activations are NHWC bf16; every thread
// moves 16-byte vectors (8 channels) and all reductions are fp32 (fp64 for the final variance).
//
// Reference arithmetic (file:line in the reference tree):
//   first conv   nets.py:20-21 via cunet.py:21,45        last conv+tanh  cunet.py:39-40,80-82
//   maxpool      cunet.py:27,46,49,52                    AdaIN           utils.py:26-51
//   upsample     cunet.py:26,60,67,74                    dropout         cunet.py:28,61,68,75
#include <cstdio>

#include "wu_host.h"
#include "wu_ptx.cuh"

namespace wu {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x);
  f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z);
  f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// streaming (read-once / write-once) variants: keep L1 for the reused operands
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Philox4x32 (Salmon et al. 2011), 7 rounds (the smallest round count that passes BigCrush):
// counter = 128-bit, key = 64-bit.  The seven round keys (key + r * Weyl constant) are computed on
// the host and arrive as a kernel parameter, so a round is two wide multiplies and two three-input
// XORs whose key operand comes straight from the constant bank.
struct PhiloxKeys {
  uint32_t x[7], y[7];
};
static inline PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
  for (int r = 0; r < 7; ++r) {
    k.x[r] = kx;
    k.y[r] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
  return k;
}
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, const PhiloxKeys& key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint64_t p0 = (uint64_t)M0 * ctr.x, p1 = (uint64_t)M1 * ctr.z;
    ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ key.x[r], (uint32_t)p1,
                     (uint32_t)(p0 >> 32) ^ ctr.w ^ key.y[r], (uint32_t)p0);
  }
  return ctr;
}
// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
// One block transposes a [32 channels][32 pixels] tile through shared memory.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                    int C, long long HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((long long)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[((long long)b * HW + p) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                    int C, long long HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(src[((long long)b * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((long long)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
// first layer: Conv2d(3, 64, 3, padding=1) + ReLU, NCHW fp32 image -> NHWC bf16
// ------------------------------------------------------------------------------------------------
// (the forward pass lives in wu_conv_first_tc.cu: TF32 im2col rows on tcgen05)

// dw[co][k] = sum_px dy[px][co] * patch[px][k]  (k = ci*9 + r*3 + s, slot 27 == 1 -> db).
// Block: 256 threads = 4 pixel slices x (16 channel quads x 4 tap octets); 32 accumulators/thread.
// Tiles of 64 pixels are double-buffered: the next tile's global loads are in flight (registers)
// while the current tile is multiplied out of shared memory.
// STRIDE 2: the discriminator stem's Conv2d(3, 64, 3, padding=1, stride=2); H, W = OUTPUT sizes.
constexpr int kFirstWgradTile = 64;  // pixels staged per iteration
template <int STRIDE>
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                        float* __restrict__ partial, int B, int H, int W) {
  const int Hin = H * STRIDE, Win = W * STRIDE;
  __shared__ __align__(16) __nv_bfloat16 sdy[2][kFirstWgradTile][64];
  __shared__ __align__(16) float sp[2][kFirstWgradTile][32];
  __shared__ __align__(16) float red[4096];
  const long long npix = (long long)B * H * W;
  const long long ntiles = (npix + kFirstWgradTile - 1) / kFirstWgradTile;
  const int q = threadIdx.x & 15;         // output channels 4q .. 4q+3
  const int kg = (threadIdx.x >> 4) & 3;  // patch slots 8kg .. 8kg+7
  const int ps = threadIdx.x >> 6;        // pixel slice
  const int sp_p = threadIdx.x >> 2, sp_part = threadIdx.x & 3;  // staging role: (pixel, channel)
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  uint4 rdy[2];
  float rp[9];
  auto fetch = [&](long long tile) {
    const long long p0 = tile * kFirstWgradTile;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + i * 256;  // 512 x 16 B
      const long long px = p0 + (idx >> 3);
      rdy[i] = px < npix ? ld_stream16(dy + px * 64 + (idx & 7) * 8) : make_uint4(0, 0, 0, 0);
    }
    const long long px = p0 + sp_p;
#pragma unroll
    for (int j = 0; j < 9; ++j) rp[j] = 0.f;
    if (px < npix) {
      if (sp_part < 3) {
        const int wq = (int)(px % W);
        const long long t = px / W;
        const int hq = (int)(t % H);
        const int b = (int)(t / H);
        const float* plane = x + ((long long)b * 3 + sp_part) * Hin * Win;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int hh = hq * STRIDE + r - 1, ww = wq * STRIDE + c - 1;
            if (hh >= 0 && hh < Hin && ww >= 0 && ww < Win)
              rp[r * 3 + c] = __ldg(plane + (long long)hh * Win + ww);
          }
      } else {
        rp[0] = 1.f;  // slot 27: bias gradient
      }
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + i * 256;
      *reinterpret_cast<uint4*>(&sdy[buf][idx >> 3][(idx & 7) * 8]) = rdy[i];
    }
    if (sp_part < 3) {
#pragma unroll
      for (int j = 0; j < 9; ++j) sp[buf][sp_p][sp_part * 9 + j] = rp[j];
    } else {
      sp[buf][sp_p][27] = rp[0];
#pragma unroll
      for (int j = 28; j < 32; ++j) sp[buf][sp_p][j] = 0.f;
    }
  };

  int buf = 0;
  if ((long long)blockIdx.x < ntiles) {
    fetch(blockIdx.x);
    commit(0);
  }
  __syncthreads();
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < ntiles) fetch(next);
#pragma unroll 4
    for (int p = ps; p < kFirstWgradTile; p += 4) {
      const uint2 dv = *reinterpret_cast<const uint2*>(&sdy[buf][p][q * 4]);
      const float4 a = *reinterpret_cast<const float4*>(&sp[buf][p][kg * 8]);
      const float4 c = *reinterpret_cast<const float4*>(&sp[buf][p][kg * 8 + 4]);
      const float dd[4] = {bf16lo(dv.x), bf16hi(dv.x), bf16lo(dv.y), bf16hi(dv.y)};
      const float pp[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(dd[i], pp[j], acc[i][j]);
    }
    if (next < ntiles) commit(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // fold the 4 pixel slices through shared memory, two slices per pass
  float* outp = partial + (size_t)blockIdx.x * 64 * 32;
  for (int pass = 0; pass < 2; ++pass) {
    if ((ps >> 1) == pass) {
      float* r = red + (ps & 1) * 2048;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[(q * 4 + i) * 32 + kg * 8 + j] = acc[i][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
      const float sum = red[i] + red[2048 + i];
      if (pass == 0) outp[i] = sum; else outp[i] += sum;
    }
    __syncthreads();
  }
}
__global__ void conv_first_wgrad_final_kernel(const float* __restrict__ partial,
                                              float* __restrict__ dw, float* __restrict__ db,
                                              int nblocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // co*32 + slot
  if (i >= 64 * 32) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 2048 + i];
  const int co = i >> 5, k = i & 31;
  if (k < 27) dw[co * 27 + k] = s;
  else if (k == 27 && db != nullptr) db[co] = s;
}

// ------------------------------------------------------------------------------------------------
// last layer: Conv2d(64, 3, 1) + Tanh, NHWC bf16 -> NCHW fp32
// ------------------------------------------------------------------------------------------------
// 8 lanes share a pixel (8 channels each); a warp covers 32 consecutive pixels in 8 steps.
__global__ void __launch_bounds__(256)
conv_last_tanh_fprop_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ y, long long npix,
                            long long HW) {
  __shared__ float so[8][3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7;
  float wr[3][8];
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[o][e] = __ldg(w + o * 64 + sub * 8 + e);
  const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f,
              b2 = bias ? __ldg(bias + 2) : 0.f;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long base = ((long long)blockIdx.x * 8 + warp) * 32; base < npix; base += nwarps * 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long px = base + i * 4 + (lane >> 3);
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (px < npix) unpack8(ld_stream16(x + px * 64 + sub * 8), f);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a0 = fmaf(f[e], wr[0][e], a0);
        a1 = fmaf(f[e], wr[1][e], a1);
        a2 = fmaf(f[e], wr[2][e], a2);
      }
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, m);
        a1 += __shfl_xor_sync(0xffffffffu, a1, m);
        a2 += __shfl_xor_sync(0xffffffffu, a2, m);
      }
      if (sub == 0) {
        const int p = i * 4 + (lane >> 3);
        so[warp][0][p] = tanhf(a0 + b0);
        so[warp][1][p] = tanhf(a1 + b1);
        so[warp][2][p] = tanhf(a2 + b2);
      }
    }
    __syncwarp();
    const long long px = base + lane;
    if (px < npix) {
      const long long b = px / HW, r = px - b * HW;
#pragma unroll
      for (int o = 0; o < 3; ++o) y[(b * 3 + o) * HW + r] = so[warp][o][lane];
    }
    __syncwarp();
  }
}

// gx[px][ci] = (x > 0) * sum_o w[o][ci] t[o],  t[o] = gy[o] * (1 - y[o]^2)
// partial[block][3*64 + 3]: dw[o][ci] = sum_px t[o] x[px][ci], db[o] = sum_px t[o]
__global__ void __launch_bounds__(256)
conv_last_tanh_bprop_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                            const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                            __nv_bfloat16* __restrict__ gx, float* __restrict__ partial,
                            long long npix, long long HW) {
  __shared__ float st[8][3][32];
  __shared__ float red[8][4][200];  // [warp][pixel-in-quad lane group][195]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  float wr[3][8];
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[o][e] = __ldg(w + o * 64 + sub * 8 + e);
  float dw[3][8];
  float dbl[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) dw[o][e] = 0.f;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long base = ((long long)blockIdx.x * 8 + warp) * 32; base < npix; base += nwarps * 32) {
    {
      const long long px = base + lane;
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      if (px < npix) {
        const long long b = px / HW, r = px - b * HW;
        const float y0 = y[(b * 3 + 0) * HW + r], y1 = y[(b * 3 + 1) * HW + r],
                    y2 = y[(b * 3 + 2) * HW + r];
        t0 = gy[(b * 3 + 0) * HW + r] * (1.f - y0 * y0);
        t1 = gy[(b * 3 + 1) * HW + r] * (1.f - y1 * y1);
        t2 = gy[(b * 3 + 2) * HW + r] * (1.f - y2 * y2);
      }
      st[warp][0][lane] = t0;
      st[warp][1][lane] = t1;
      st[warp][2][lane] = t2;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = i * 4 + grp;
      const long long px = base + p;
      const float t0 = st[warp][0][p], t1 = st[warp][1][p], t2 = st[warp][2][p];
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (px < npix) unpack8(ld_stream16(x + px * 64 + sub * 8), f);
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dw[0][e] = fmaf(t0, f[e], dw[0][e]);
        dw[1][e] = fmaf(t1, f[e], dw[1][e]);
        dw[2][e] = fmaf(t2, f[e], dw[2][e]);
        const float v = fmaf(t0, wr[0][e], fmaf(t1, wr[1][e], t2 * wr[2][e]));
        g[e] = f[e] > 0.f ? v : 0.f;
      }
      if (sub == 0) {
        dbl[0] += t0;
        dbl[1] += t1;
        dbl[2] += t2;
      }
      if (px < npix) st_stream16(gx + px * 64 + sub * 8, pack8(g));
    }
    __syncwarp();
  }
  // block reduction: each (warp, grp) writes its 8-lane row of 3x64 (+3) values, then fold 32 rows
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][grp][o * 64 + sub * 8 + e] = dw[o][e];
  if (sub == 0) {
    red[warp][grp][192] = dbl[0];
    red[warp][grp][193] = dbl[1];
    red[warp][grp][194] = dbl[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 195; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv)
#pragma unroll
      for (int g = 0; g < 4; ++g) s += red[wv][g][i];
    partial[(size_t)blockIdx.x * 195 + i] = s;
  }
}
__global__ void conv_last_bprop_final_kernel(const float* __restrict__ partial,
                                             float* __restrict__ dw, float* __restrict__ db,
                                             int nblocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 195) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 195 + i];
  if (i < 192) dw[i] = s;
  else if (db != nullptr) db[i - 192] = s;
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B,
                    int H, int W, int C) {
  const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const __nv_bfloat16* p = src + (((long long)b * H + 2 * ho) * W + 2 * wo) * C + v * 8;
    float a[8], c[8], d[8], e[8], m[8];
    unpack8(ld_stream16(p), a);
    unpack8(ld_stream16(p + C), c);
    unpack8(ld_stream16(p + (long long)W * C), d);
    unpack8(ld_stream16(p + (long long)W * C + C), e);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(a[j], c[j]), fmaxf(d[j], e[j]));
    *reinterpret_cast<uint4*>(dst + i * 8) = pack8(m);
  }
}
// thread = (pooled pixel, 8 channels): writes the four full-resolution gradients of its window.
// Everything stays in packed bf16x2: comparisons are 16-bit SIMD integer compares on an order-
// preserving key (bf16 is sign-magnitude), the one addition is a bf16x2 add (exact sum, one rounding:
// the same value an fp32 add + rounding gives).  A first version unpacked to fp32 (101 registers, 24 %
// of the warps resident, latency bound at 48 % of HBM peak).
__device__ __forceinline__ uint32_t bf16x2_order_key(uint32_t v) {
  return v ^ (__vcmplts2(v, 0u) & 0x7FFF7FFFu);  // negative lanes: flip the magnitude bits
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                   *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ g_pool,
                    const __nv_bfloat16* __restrict__ g_skip, __nv_bfloat16* __restrict__ g, int B,
                    int H, int W, int C) {
  const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const long long o00 = (((long long)b * H + 2 * ho) * W + 2 * wo) * C + v * 8;
    const long long offs[4] = {o00, o00 + C, o00 + (long long)W * C, o00 + (long long)W * C + C};
    uint4 yv[4], gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      yv[k] = ld_stream16(y + offs[k]);
      gs[k] = g_skip != nullptr ? ld_stream16(g_skip + offs[k]) : make_uint4(0, 0, 0, 0);
    }
    const uint4 gp = g_pool != nullptr ? ld_stream16(g_pool + i * 8) : make_uint4(0, 0, 0, 0);
    uint4 out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // two channels per 32-bit word
      const uint32_t y0 = (&yv[0].x)[e], y1 = (&yv[1].x)[e], y2 = (&yv[2].x)[e], y3 = (&yv[3].x)[e];
      const uint32_t k0 = bf16x2_order_key(y0), k1 = bf16x2_order_key(y1);
      const uint32_t k2 = bf16x2_order_key(y2), k3 = bf16x2_order_key(y3);
      // first maximum in window scan order (h-major), PyTorch's tie rule: a later element wins
      // only when strictly greater
      const uint32_t s01 = __vcmpgts2(k1, k0), s23 = __vcmpgts2(k3, k2);
      const uint32_t m01 = __vmaxs2(k0, k1), m23 = __vmaxs2(k2, k3);
      const uint32_t top = __vcmpgts2(m23, m01);  // the maximum is in the second row
      const uint32_t is[4] = {~top & ~s01, ~top & s01, top & ~s23, top & s23};
      const uint32_t gpe = (&gp.x)[e];
      const uint32_t ys[4] = {y0, y1, y2, y3};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t sum = bf16x2_add((&gs[k].x)[e], gpe & is[k]);
        (&out[k].x)[e] = sum & __vcmpgts2(ys[k], 0u);  // ReLU mask of y
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) st_stream16(g + offs[k], out[k]);
  }
}

// ------------------------------------------------------------------------------------------------
// AdaIN statistics: per-(b, chunk, c) sums of x and x^2
// ------------------------------------------------------------------------------------------------
constexpr int kStatChunk = 256;  // pixels per block

// Shared block reduction over pixel groups for two 8-wide accumulators.
// red must hold 2 * groups * C floats.
__device__ __forceinline__ void block_reduce_pairs(float* red, const float (&s1)[8],
                                                   const float (&s2)[8], int C, int lanes, int g,
                                                   int l, float* __restrict__ out /* [C][2] */) {
  const int groups = blockDim.x / lanes;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[(g * C + l * 8 + e) * 2 + 0] = s1[e];
    red[(g * C + l * 8 + e) * 2 + 1] = s2[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[gg * 2 * C + i];
    out[i] = s;
  }
}

__global__ void __launch_bounds__(256)
adain_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int HW, int C,
                   int nchunk) {
  extern __shared__ float red[];
  const int lanes = C >> 3, groups = blockDim.x / lanes;
  const int g = threadIdx.x / lanes, l = threadIdx.x % lanes;
  const int b = blockIdx.x / nchunk, chunk = blockIdx.x % nchunk;
  const int p0 = chunk * kStatChunk;
  const int p1 = min(HW, p0 + kStatChunk);
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = p0 + g; p < p1; p += groups) {
    float f[8];
    unpack8(ldg16(x + ((long long)b * HW + p) * C + l * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1[e] += f[e];
      s2[e] = fmaf(f[e], f[e], s2[e]);
    }
  }
  block_reduce_pairs(red, s1, s2, C, lanes, g, l, partial + ((size_t)b * nchunk + chunk) * C * 2);
}

// Sum of the per-chunk (s1, s2) pairs of one (b, c) in fp64, in chunk order.  The loads of eight
// chunks are issued together: with one dependent load per iteration the style kernels (a few
// thousand threads) spent 64 L2 round trips, 35 us, on this loop.
__device__ __forceinline__ void fold_partials(const float* __restrict__ p0, int nchunk, int C,
                                              double& s1, double& s2) {
  const size_t pitch = (size_t)C * 2;
  int k = 0;
  for (; k + 8 <= nchunk; k += 8) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float2*>(p0 + (k + u) * pitch));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s1 += (double)v[u].x;
      s2 += (double)v[u].y;
    }
  }
  for (; k < nchunk; ++k) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p0 + k * pitch));
    s1 += (double)v.x;
    s2 += (double)v.y;
  }
}

// style = l1(cond) -> (y_mean, y_std) over the 4 numbers of a channel; combine with x statistics.
// Block = 32 channels x 8 chunk slices: slice j folds chunks j, j + 8, ... (fp64), the slices are then
// added in a fixed order through shared memory.  (One thread per channel walking up to 256 chunks — the
// 64-pixel chunks the convolution epilogues emit — was a 30 us chain of dependent L2 round trips.)
constexpr int kStyleSlices = 8;
__global__ void __launch_bounds__(32 * kStyleSlices)
adain_style_fwd_kernel(const float* __restrict__ cond, const float* __restrict__ lw,
                       const float* __restrict__ lb, const float* __restrict__ partial,
                       float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ ystd,
                       float* __restrict__ scale, float* __restrict__ shift, int B, int C, int nc, int HW,
                       int nchunk, float eps, int xb_mul) {
  __shared__ double sh1[kStyleSlices][32], sh2[kStyleSlices][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;  // (b, c) linear; C is a multiple of 32 or the tail is masked
  const bool live = i < B * C;
  const int b = live ? i / C : 0, c = live ? i - b * C : 0;
  const int bx = b * xb_mul;  // batch index of the statistics (0 when one x serves every condition)
  double s1 = 0.0, s2 = 0.0;
  if (live) {
    const float* p0 = partial + ((size_t)bx * nchunk * C + c) * 2;
    const size_t pitch = (size_t)C * 2;
    int k = ty;
    for (; k + 3 * kStyleSlices < nchunk; k += 4 * kStyleSlices) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(reinterpret_cast<const float2*>(p0 + (size_t)(k + u * kStyleSlices) * pitch));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1 += (double)v[u].x;
        s2 += (double)v[u].y;
      }
    }
    for (; k < nchunk; k += kStyleSlices) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p0 + (size_t)k * pitch));
      s1 += (double)v.x;
      s2 += (double)v.y;
    }
  }
  sh1[ty][tx] = s1;
  sh2[ty][tx] = s2;
  __syncthreads();
  if (ty != 0 || !live) return;
#pragma unroll
  for (int j = 1; j < kStyleSlices; ++j) {
    s1 += sh1[j][tx];
    s2 += sh2[j][tx];
  }
  float h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = lb[4 * c + j];
    for (int k = 0; k < nc; ++k) a = fmaf(cond[b * nc + k], lw[(4 * c + j) * nc + k], a);
    h[j] = a;
  }
  const float ym = 0.25f * (h[0] + h[1] + h[2] + h[3]);
  float yv = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) yv += (h[j] - ym) * (h[j] - ym);
  const float ys = sqrtf(yv * (1.f / 3.f) + eps);
  const double m = s1 / HW;
  double var = (s2 - s1 * m) / (HW > 1 ? HW - 1 : 1);
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[i] = (float)m;
  rstd[i] = r;
  ystd[i] = ys;
  const float sc = ys * r;
  scale[i] = sc;
  shift[i] = ym - (float)m * sc;
}

// ------------------------------------------------------------------------------------------------
// AdaIN apply + bilinear x2 (align_corners=True) + dropout
// ------------------------------------------------------------------------------------------------
// PyTorch's source index for align_corners=True: src = dst * (in - 1) / (out - 1), in fp32.
__device__ __forceinline__ void bilinear_src(int dst, float ratio, int in, int& i0, int& i1,
                                             float& lam) {
  const float s = ratio * (float)dst;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = s - (float)i0;
}

// Dropout keep decisions of one 8-channel vector as a byte (bit j = channel j) and as packed
// bf16x2 lane masks.  The forward pass decides (Philox stream or injected uint8 mask) and stores the
// byte; the backward pass only reads bytes back (1/16 of the activation bytes, no RNG replay).
enum { kDropNone = 0, kDropPhilox = 1, kDropInjected = 2 };

__device__ __forceinline__ uint32_t lanes_to_byte(const uint4& km) {
  return (km.x & 1u) | ((km.x >> 15) & 2u) | ((km.y & 1u) << 2) | ((km.y >> 13) & 8u) |
         ((km.z & 1u) << 4) | ((km.z >> 11) & 32u) | ((km.w & 1u) << 6) | ((km.w >> 9) & 128u);
}
// prmt.b32, generic mode: selector nibble n < 8 copies byte n of {a (0-3), b (4-7)}; n >= 8 fills the
// byte with the sign bit of byte n - 8.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// (a multiply + sign-replicating PRMT version of this, eight instructions, made the adjoint kernel 9 %
// slower than the shift pairs the compiler builds from the selects below)
__device__ __forceinline__ uint32_t byte_pair_to_lanes(uint32_t two_bits) {
  return ((two_bits & 1u) ? 0x0000FFFFu : 0u) | ((two_bits & 2u) ? 0xFFFF0000u : 0u);
}
__device__ __forceinline__ uint4 byte_to_lanes(uint32_t b) {
  return make_uint4(byte_pair_to_lanes(b), byte_pair_to_lanes(b >> 2), byte_pair_to_lanes(b >> 4),
                    byte_pair_to_lanes(b >> 6));
}
// injected uint8 keep mask (tests, golden fixtures): 8 bytes per vector -> bf16x2 lane masks
__device__ __forceinline__ uint4 injected_keep8(const uint8_t* __restrict__ mask, uint32_t vi) {
  const uint2 mv = __ldg(reinterpret_cast<const uint2*>(mask + (size_t)vi * 8));
  auto lanes = [](uint32_t two_bytes) -> uint32_t {
    return ((two_bytes & 0xFFu) ? 0x0000FFFFu : 0u) | ((two_bytes & 0xFF00u) ? 0xFFFF0000u : 0u);
  };
  return make_uint4(lanes(mv.x), lanes(mv.x >> 16), lanes(mv.y), lanes(mv.y >> 16));
}

// Philox keep decisions of one 8-channel vector for the forward kernel: 15 random bits per element
// (rate resolution 2^-15), decided without a compare instruction: (u15 + 32768 - thr) carries into
// bit 15 of its 16-bit lane iff u15 >= thr (no carry crosses the lane: the sum is < 65536), and one
// PRMT replicates that bit over the lane (selector nibble 8 + k = "sign of byte k").
__host__ __device__ inline uint32_t dropout_threshold15(float p) {
  float t = p * 32768.f + 0.5f;
  if (t < 0.f) t = 0.f;
  if (t > 32767.f) t = 32767.f;
  return (uint32_t)t;
}
__device__ __forceinline__ void philox_keep8(const PhiloxKeys& keys, uint32_t vimg, uint32_t b,
                                             uint32_t k2, uint4& lanes, uint32_t& byte) {
  // counter = (8-channel vector index inside the image, image index | epoch << 16, tag, 0); the
  // epoch (0 unless the caller keeps a device-side draw counter) is merged into b by the kernel
  const uint4 r = philox4x32(make_uint4(vimg, b, 0x77755555u, 0u), keys);
  const uint32_t s0 = (r.x & 0x7FFF7FFFu) + k2, s1 = (r.y & 0x7FFF7FFFu) + k2;
  const uint32_t s2 = (r.z & 0x7FFF7FFFu) + k2, s3 = (r.w & 0x7FFF7FFFu) + k2;
  lanes = make_uint4(prmt(s0, 0u, 0xBB99u), prmt(s1, 0u, 0xBB99u), prmt(s2, 0u, 0xBB99u),
                     prmt(s3, 0u, 0xBB99u));
  // bytes 1 and 3 of every word carry the decision in their top bit: gather the four decisions of two
  // words as 0x00 / 0xFF bytes, keep one bit of each, and move them to bits 24..27 with one multiply
  // (the partial products do not collide)
  const uint32_t t0 = prmt(s0, s1, 0xFDB9u) & 0x01010101u;
  const uint32_t t1 = prmt(s2, s3, 0xFDB9u) & 0x01010101u;
  byte = ((t0 * 0x01020408u) >> 24) | (((t1 * 0x01020408u) >> 20) & 0xF0u);
}

// The blend is evaluated separably, on values that already carry the AdaIN affine map (the bilinear
// weights sum to 1, so the map commutes with the blend; 1/(1-p) is folded into sc / sh):
//   p' = sc * p + sh per source vector, H = p'_left + lx * (p'_right - p'_left) per source row and
//   output column, o = H_top + ly * (H_bottom - H_top) per output vector.
// Shared between the eight vectors of a thread's block this is 27 fp32 operations per output vector
// instead of 49 for the four-weight form (the fma pipe, which also carries Philox's wide multiplies,
// is the busiest pipe of this kernel).  Every step is an explicit fmaf / __fsub_rn so that the block
// path and the general path round identically.
__device__ __forceinline__ void affine8(const uint4& raw, const float (&sc)[8], const float (&sh)[8],
                                        float (&o)[8]) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], sc[j], sh[j]);
}
__device__ __forceinline__ void lerp8(const float (&a)[8], const float (&diff)[8], float t,
                                      float (&o)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(t, diff[j], a[j]);
}
__device__ __forceinline__ void sub8(const float (&b)[8], const float (&a)[8], float (&o)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __fsub_rn(b[j], a[j]);
}
// One output vector from its two horizontally blended rows: vertical blend, dropout, store.  u, keep_bits
// and mask point at image b; vimg is the vector index inside the image.
template <int MODE>
__device__ __forceinline__ void adain_emit(const float (&ht)[8], const float (&dv)[8], float ly,
                                           uint32_t vimg, uint32_t b, __nv_bfloat16* __restrict__ u,
                                           uint8_t* __restrict__ keep_bits, uint32_t thr,
                                           const PhiloxKeys& keys, const uint8_t* __restrict__ mask) {
  float o[8];
  lerp8(ht, dv, ly, o);
  uint4 ov = pack8(o);
  if (MODE == kDropPhilox) {
    uint4 km;
    uint32_t kb;
    philox_keep8(keys, vimg, b, thr, km, kb);
    ov.x &= km.x; ov.y &= km.y; ov.z &= km.z; ov.w &= km.w;
    keep_bits[vimg] = (uint8_t)kb;
  } else if (MODE == kDropInjected) {
    const uint4 km = injected_keep8(mask, vimg);
    ov.x &= km.x; ov.y &= km.y; ov.z &= km.z; ov.w &= km.w;
    keep_bits[vimg] = (uint8_t)lanes_to_byte(km);
  }
  st_stream16(u + (size_t)vimg * 8, ov);
}

// grid = (ceil((w + 1) * C/8 / 256), h/2 + 1, B).  With align_corners=True and an exact x2 scale the
// output columns 2k-1 and 2k blend the SAME source columns k-1 and k (src = dst (w-1)/(2w-1) lies in
// [k-1, k) for both), and likewise for rows.  A thread therefore owns one 8-channel vector of a 4-row x
// 2-column output block (rows 4j-1 .. 4j+2, columns 2k-1, 2k; k = 0 .. w) and produces its eight
// output vectors from SIX source vectors (rows 2j-1, 2j, 2j+1 x columns k-1, k) that are loaded and
// converted once: 0.75 loads and 6 conversions per output vector instead of 4 and 32 (the kernel is
// instruction bound, not HBM bound).  Source indices and weights are still PyTorch's fp32 arithmetic
// (bilinear_src); a thread whose computed indices differ from the block pattern (possible only where
// src is an exact integer, i.e. the last row / column, through rounding) takes the general
// four-loads-per-vector path, so the result never depends on the pattern being right.  An index may
// differ where its weight is exactly zero (first row / column: PyTorch reads row 1 with weight 0).
constexpr int kRowsPerThread = 4;
// (four blocks per SM, 64 registers: 3 % faster than the natural 70 registers / three blocks)
template <int MODE>
__global__ void __launch_bounds__(256, 4)
adain_up_drop_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                         const float* __restrict__ shift, __nv_bfloat16* __restrict__ u,
                         uint8_t* __restrict__ keep_bits, int h, int w, int C, float inv_keep,
                         uint32_t thr, const __grid_constant__ PhiloxKeys keys,
                         const uint8_t* __restrict__ mask, int xb_mul, float rh, float rw, int cv_shift,
                         const uint32_t* __restrict__ epoch) {
  // rh = (h-1)/(2h-1), rw = (w-1)/(2w-1) in fp32 (PyTorch's ratio) and cv_shift = log2(C/8) or -1 come
  // from the host: two fp32 divisions and an integer division per thread were a tenth of the kernel
  const int cv = C >> 3, Ho = 2 * h, Wo = 2 * w;
  const int xi = blockIdx.x * blockDim.x + threadIdx.x;
  if (xi >= (w + 1) * cv) return;
  const int kx = cv_shift >= 0 ? (xi >> cv_shift) : xi / cv;
  const int v = xi - kx * cv;
  const int b = blockIdx.z;
  // Philox counter word 1: image index (< 65536) | low 16 bits of the device-side draw counter.  The
  // counter lives in device memory so that a CUDA-graph replay of the same launch draws a NEW mask.
  const uint32_t ctr_b = (uint32_t)b | ((MODE == kDropPhilox && epoch != nullptr) ? (__ldg(epoch) << 16) : 0u);
  const int j4 = blockIdx.y;  // row group: rows 4j-1 .. 4j+2, j = 0 .. h/2
  const int Y0 = 4 * j4 - 1;
  const size_t img_vecs = (size_t)Ho * Wo * cv;
  u += (size_t)b * img_vecs * 8;
  if (MODE != kDropNone) keep_bits += (size_t)b * img_vecs;
  if (MODE == kDropInjected) mask += (size_t)b * img_vecs * 8;
  const __nv_bfloat16* xb = x + (long long)(b * xb_mul) * h * w * C + v * 8;
  const int row_pitch = w * C;  // 32-bit offsets inside one image
  const float4* scp = reinterpret_cast<const float4*>(scale + (long long)b * C + v * 8);
  const float4* shp = reinterpret_cast<const float4*>(shift + (long long)b * C + v * 8);
  const float4 sc0 = __ldg(scp), sc1 = __ldg(scp + 1), sh0 = __ldg(shp), sh1 = __ldg(shp + 1);
  const float sc[8] = {sc0.x * inv_keep, sc0.y * inv_keep, sc0.z * inv_keep, sc0.w * inv_keep,
                       sc1.x * inv_keep, sc1.y * inv_keep, sc1.z * inv_keep, sc1.w * inv_keep};
  const float sh[8] = {sh0.x * inv_keep, sh0.y * inv_keep, sh0.z * inv_keep, sh0.w * inv_keep,
                       sh1.x * inv_keep, sh1.y * inv_keep, sh1.z * inv_keep, sh1.w * inv_keep};

  // the block pattern: source columns cA, cB and source rows rr[0..2] (clamped to the image)
  const int cA = max(kx - 1, 0), cB = min(kx, w - 1);
  const int rr[3] = {max(2 * j4 - 1, 0), min(2 * j4, h - 1), min(2 * j4 + 1, h - 1)};
  // PyTorch's indices and weights of the two columns and four rows, and whether they fit the pattern
  const int X[2] = {2 * kx - 1, 2 * kx};
  const bool okx[2] = {kx >= 1, kx < w};
  int x0[2], x1[2];
  float lx[2];
  bool fast = true;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    bilinear_src(okx[q] ? X[q] : 0, rw, w, x0[q], x1[q], lx[q]);
    if (okx[q]) fast = fast && x0[q] == cA && (x1[q] == cB || lx[q] == 0.f);
  }
  int y0[4], y1[4];
  float ly[4];
  bool oky[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    oky[i] = (Y0 + i >= 0) && (Y0 + i < Ho);
    bilinear_src(oky[i] ? Y0 + i : 0, rh, h, y0[i], y1[i], ly[i]);
    if (oky[i]) fast = fast && y0[i] == rr[i >> 1] && (y1[i] == rr[(i >> 1) + 1] || ly[i] == 0.f);
  }
  const int vi_row = Wo * cv;
  const int vi00 = (Y0 * Wo + X[0]) * cv + v;  // (row Y0, column 2k-1); only valid (row, column) are used

  if (fast) {
    uint4 raw[3][2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      raw[r][0] = ldg16(xb + rr[r] * row_pitch + cA * C);
      raw[r][1] = ldg16(xb + rr[r] * row_pitch + cB * C);
    }
    // H[r][q]: source row r blended horizontally for output column q (affine map applied first)
    float H[3][2][8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float pl[8], pr[8], dx[8];
      affine8(raw[r][0], sc, sh, pl);
      affine8(raw[r][1], sc, sh, pr);
      sub8(pr, pl, dx);
      lerp8(pl, dx, lx[0], H[r][0]);
      lerp8(pl, dx, lx[1], H[r][1]);
    }
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {  // rows 4j-1, 4j blend H[0], H[1]; rows 4j+1, 4j+2 H[1], H[2]
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!okx[q]) continue;
        float dv[8];
        sub8(H[pair + 1][q], H[pair][q], dv);
#pragma unroll
        for (int i = 2 * pair; i < 2 * pair + 2; ++i)
          if (oky[i])
            adain_emit<MODE>(H[pair][q], dv, ly[i], (uint32_t)(vi00 + i * vi_row + q * cv), ctr_b, u,
                             keep_bits, thr, keys, mask);
      }
    }
    return;
  }
  // general path: four loads per output vector at PyTorch's own indices (recomputed here so that the
  // index arrays above stay in registers)
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int Y = Y0 + i;
    if (Y < 0 || Y >= Ho) continue;
    int ya, yb;
    float lyy;
    bilinear_src(Y, rh, h, ya, yb, lyy);
    const __nv_bfloat16* r0 = xb + ya * row_pitch;
    const __nv_bfloat16* r1 = xb + yb * row_pitch;
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      const int XX = 2 * kx - 1 + q;
      if (XX < 0 || XX >= Wo) continue;
      int xa, xc;
      float lxx;
      bilinear_src(XX, rw, w, xa, xc, lxx);
      float pl[8], pr[8], dx[8], ht[8], hb[8], dv[8];
      affine8(ldg16(r0 + xa * C), sc, sh, pl);
      affine8(ldg16(r0 + xc * C), sc, sh, pr);
      sub8(pr, pl, dx);
      lerp8(pl, dx, lxx, ht);
      affine8(ldg16(r1 + xa * C), sc, sh, pl);
      affine8(ldg16(r1 + xc * C), sc, sh, pr);
      sub8(pr, pl, dx);
      lerp8(pl, dx, lxx, hb);
      sub8(hb, ht, dv);
      adain_emit<MODE>(ht, dv, lyy, (uint32_t)(vi00 + i * vi_row + q * cv), ctr_b, u, keep_bits, thr,
                       keys, mask);
    }
  }
}

// Weight with which destination index D contributes to source index s (0 if it does not).
__device__ __forceinline__ float bilinear_adjoint_w(int D, int s, float ratio, int in, int out) {
  if (D < 0 || D >= out) return 0.f;
  int i0, i1;
  float lam;
  bilinear_src(D, ratio, in, i0, i1, lam);
  float wgt = 0.f;
  if (i0 == s) wgt += 1.f - lam;
  if (i1 == s) wgt += lam;
  return wgt;
}

// Adjoint of dropout o upsample, plus the per-(b, chunk, c) sums S1 = sum gz, S2 = sum gz * xhat, in
// ONE pass over gu (a first version ran a horizontal pass into a scratch tensor and a vertical pass
// out of it: 6.25 E + 4 E bytes moved for E = bytes of the low-resolution tensor, here 4.25 E + 2 E):
//   gz[b,y,x,c] = sum_Y wy(Y->y) sum_X wx(X->x) keep(b,Y,X,c)/(1-p) gu[b,Y,X,c]
// A block owns `groups` = 256 / (C/8) low-resolution columns x kAdjRows rows of one image; a thread
// owns one (column, 8-channel vector) and walks DOWN the upsampled rows Y: the horizontally reduced row
// t(Y) (six candidate X, four of them with non-zero weight) feeds the two output rows i0(Y), i0(Y)+1,
// and because i0 never decreases two accumulators suffice: when i0 advances, the finished row is
// written, added to the sums, and the accumulators shift.  i0(Y) is the same for the whole block.
// keep comes from the byte tensor the forward pass stored (keep_bits == nullptr: no dropout).
constexpr int kAdjRows = 16;
// (three blocks per SM: 80 registers; at the natural 112 the kernel was latency bound at 25 % occupancy)
__global__ void __launch_bounds__(256, 3)
adain_up_drop_adjoint_kernel(const __nv_bfloat16* __restrict__ gu, const uint8_t* __restrict__ keep_bits,
                             const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                             const float* __restrict__ rstd, __nv_bfloat16* __restrict__ gz,
                             float* __restrict__ partial, int h, int w, int C, float inv_keep,
                             int col_blocks, int row_blocks) {
  extern __shared__ float red[];
  const int lanes = C >> 3, groups = blockDim.x / lanes;
  const int g = threadIdx.x / lanes, l = threadIdx.x % lanes;
  const int chunks = col_blocks * row_blocks;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int rb = chunk / col_blocks, cb = chunk - rb * col_blocks;
  const int Ho = 2 * h, Wo = 2 * w;
  const int y0 = rb * kAdjRows, y1 = min(h, y0 + kAdjRows);  // output rows [y0, y1)
  const int xx = cb * groups + g;
  const bool live = xx < w;
  const float rh = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float rw = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  float wx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j)
    wx[j] = live ? bilinear_adjoint_w(2 * xx - 2 + j, xx, rw, w, Wo) * inv_keep : 0.f;
  const float4* mp = reinterpret_cast<const float4*>(mean + (long long)b * C + l * 8);
  const float4* rp = reinterpret_cast<const float4*>(rstd + (long long)b * C + l * 8);
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float accA[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accB[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cur = y0 - 1;  // accA collects output row `cur`, accB row `cur + 1`
  auto finish_row = [&](int r, const float (&acc)[8]) {
    if (!live || r < y0 || r >= y1) return;
    const long long off = (((long long)b * h + r) * w + xx) * C + l * 8;
    float xv[8];
    unpack8(ldg16(x + off), xv);
    // mean / rstd are re-read (L1 hits) once per finished row instead of living in 16 registers
    const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), r0 = __ldg(rp), r1 = __ldg(rp + 1);
    const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float rs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1[e] += acc[e];
      s2[e] = fmaf(acc[e], (xv[e] - mu[e]) * rs[e], s2[e]);
    }
    *reinterpret_cast<uint4*>(gz + off) = pack8(acc);
  };
  for (int Y = max(0, 2 * y0 - 3); Y < Ho; ++Y) {
    int i0, i1;
    float lam;
    bilinear_src(Y, rh, h, i0, i1, lam);
    if (i0 < y0 - 1) continue;  // feeds rows above this block only
    if (i0 >= y1) break;        // rows [y0, y1) are complete
    while (cur < i0) {          // block-uniform: row `cur` has all its contributions
      finish_row(cur, accA);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        accA[e] = accB[e];
        accB[e] = 0.f;
      }
      ++cur;
    }
    float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
      const long long vbase = (((long long)b * Ho + Y) * Wo + 2 * xx - 2) * lanes + l;
      uint4 gv[6];
      uint32_t kb[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        gv[j] = make_uint4(0, 0, 0, 0);
        kb[j] = 0xFFu;
        if (wx[j] != 0.f) {
          gv[j] = ld_stream16(gu + (vbase + (long long)j * lanes) * 8);
          if (keep_bits != nullptr) kb[j] = __ldg(keep_bits + vbase + (long long)j * lanes);
        }
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (wx[j] != 0.f) {
          const uint4 km = byte_to_lanes(kb[j]);
          uint4 gq = gv[j];
          gq.x &= km.x; gq.y &= km.y; gq.z &= km.z; gq.w &= km.w;
          float f[8];
          unpack8(gq, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = fmaf(wx[j], f[e], t[e]);
        }
      }
    }
    const float wa = 1.f - lam, wb = (i1 != i0) ? lam : 0.f;  // i1 == i0 only on the last row (lam = 0)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      accA[e] = fmaf(wa, t[e], accA[e]);
      accB[e] = fmaf(wb, t[e], accB[e]);
    }
  }
  finish_row(cur, accA);
  finish_row(cur + 1, accB);  // complete only when the loop ran out of rows (cur + 1 == h - 1 < y1)
  block_reduce_pairs(red, s1, s2, C, lanes, g, l, partial + ((size_t)b * chunks + chunk) * C * 2);
}

// thread per (b, c): fold the partials, emit k1, k2 and the gradient of the 4 style numbers.
__global__ void adain_style_bwd_kernel(const float* __restrict__ cond, const float* __restrict__ lw,
                                       const float* __restrict__ lb,
                                       const float* __restrict__ partial,
                                       const float* __restrict__ ystd,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ rstd, float* __restrict__ k1,
                                       float* __restrict__ k2, float* __restrict__ coef,
                                       float* __restrict__ gh, int B, int C, int nc, int HW,
                                       int nchunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  double s1 = 0.0, s2 = 0.0;
  fold_partials(partial + ((size_t)b * nchunk * C + c) * 2, nchunk, C, s1, s2);
  const float kk1 = (float)(s1 / HW), kk2 = (float)(s2 / (HW > 1 ? HW - 1 : 1));
  k1[i] = kk1;
  k2[i] = kk2;
  {
    // gx = relu'(x) * rstd*ystd * (gz - k1 - xhat*k2) = relu'(x) * (A*gz + Bc*x + Cc)
    const float r = rstd[i], A = r * ystd[i];
    coef[i] = A;
    coef[(size_t)B * C + i] = -A * kk2 * r;
    coef[2 * (size_t)B * C + i] = A * (kk2 * r * mean[i] - kk1);
  }
  float h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = lb[4 * c + j];
    for (int k = 0; k < nc; ++k) a = fmaf(cond[b * nc + k], lw[(4 * c + j) * nc + k], a);
    h[j] = a;
  }
  const float ym = 0.25f * (h[0] + h[1] + h[2] + h[3]);
  const float ys = ystd[i];
  // d y_mean / d h_j = 1/4 ;  d y_std / d h_j = (h_j - y_mean) / (3 y_std)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    gh[(size_t)b * 4 * C + 4 * c + j] = (float)s1 * 0.25f + (float)s2 * (h[j] - ym) / (3.f * ys);
}
// thread per row r of l1.weight: dlb[r] = sum_b gh[b][r]; dlw[r][k] = sum_b gh[b][r] cond[b][k]
__global__ void adain_style_bwd_params_kernel(const float* __restrict__ gh,
                                              const float* __restrict__ cond,
                                              float* __restrict__ dlw, float* __restrict__ dlb,
                                              int B, int R, int nc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float sb = 0.f;
  for (int b = 0; b < B; ++b) sb += gh[(size_t)b * R + r];
  dlb[r] = sb;
  for (int k = 0; k < nc; ++k) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(gh[(size_t)b * R + r], cond[b * nc + k], s);
    dlw[(size_t)r * nc + k] = s;
  }
}

// AdaIN on its own (no upsample / dropout): out = x * scale[b,c] + shift[b,c]
__global__ void __launch_bounds__(256)
adain_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                   const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int B, int HW,
                   int C) {
  const int cv = C >> 3;
  const long long total = (long long)B * HW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const int b = (int)(i / ((long long)HW * cv));
    const long long pc = (long long)b * C + v * 8;
    float f[8];
    unpack8(ld_stream16(x + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], __ldg(scale + pc + j), __ldg(shift + pc + j));
    st_stream16(out + i * 8, pack8(f));
  }
}

// gx = (x > 0) * (A*gz + Bc*x + Cc), coefficients per (b, c) from adain_style_bwd_kernel
__global__ void __launch_bounds__(256)
adain_bwd_apply_kernel(const __nv_bfloat16* __restrict__ gz, const __nv_bfloat16* __restrict__ x,
                       const float* __restrict__ coef, __nv_bfloat16* __restrict__ gx, int B, int HW,
                       int C) {
  // blockIdx.y = image: a thread keeps the 24 coefficients of its 8 channels in registers and walks
  // the image's pixels four at a time (eight independent 16-byte loads in flight)
  const int cv = C >> 3;
  const int b = blockIdx.y;
  const int l = threadIdx.x % cv, grp = threadIdx.x / cv, groups = blockDim.x / cv;
  const size_t plane = (size_t)B * C;
  const float4* cp = reinterpret_cast<const float4*>(coef + (size_t)b * C + l * 8);
  const float4 a0 = __ldg(cp), a1 = __ldg(cp + 1);
  const float4 b0 = __ldg(cp + plane / 4), b1 = __ldg(cp + plane / 4 + 1);
  const float4 c0 = __ldg(cp + plane / 2), c1 = __ldg(cp + plane / 2 + 1);
  const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float Bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float Cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const size_t img = (size_t)b * HW * C + l * 8;
  const int stride = gridDim.x * groups;
  auto one = [&](const uint4& gq, const uint4& xq, int p) {
    float g[8], xv[8], o[8];
    unpack8(gq, g);
    unpack8(xq, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = xv[j] > 0.f ? fmaf(A[j], g[j], fmaf(Bc[j], xv[j], Cc[j])) : 0.f;
    st_stream16(gx + img + (size_t)p * C, pack8(o));
  };
  int p = blockIdx.x * groups + grp;
  for (; p + 3 * stride < HW; p += 4 * stride) {
    uint4 gq[4], xq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      gq[u] = ld_stream16(gz + img + (size_t)(p + u * stride) * C);
      xq[u] = ld_stream16(x + img + (size_t)(p + u * stride) * C);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(gq[u], xq[u], p + u * stride);
  }
  for (; p < HW; p += stride)
    one(ld_stream16(gz + img + (size_t)p * C), ld_stream16(x + img + (size_t)p * C), p);
}

static inline int grid_for(long long work_items, int block, int max_waves = 16) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace wu

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace wu;
typedef __nv_bfloat16 bf16;

extern "C" int wu_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W,
                                        wu_stream_t stream) {
  WU_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "wu_nchw_f32_to_nhwc_bf16: bad args");
  WU_REQUIRE(B <= 65535, "wu_nchw_f32_to_nhwc_bf16: B=%d too large", B);
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, (bf16*)dst, C, HW);
  WU_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return WU_OK;
}
extern "C" int wu_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W,
                                        wu_stream_t stream) {
  WU_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "wu_nhwc_bf16_to_nchw_f32: bad args");
  WU_REQUIRE(B <= 65535, "wu_nhwc_bf16_to_nchw_f32: B=%d too large", B);
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)src, dst, C, HW);
  WU_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return WU_OK;
}

extern "C" int wu_conv_first_fprop(const float* x, const float* w, const float* bias, void* dst,
                                   int B, int H, int W, wu_stream_t stream) {
  WU_REQUIRE(x && w && dst && B > 0 && H > 0 && W > 0, "wu_conv_first_fprop: bad args");
  return conv_k27_fprop_tc(x, w, bias, 0.f, dst, B, H, W, 1, (cudaStream_t)stream);
}
static int first_wgrad_blocks(long long npix) {
  long long tiles = (npix + kFirstWgradTile - 1) / kFirstWgradTile;
  long long g = 2LL * num_sms();
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}
static size_t first_wgrad_ws(long long npix) {
  const size_t fma = (size_t)first_wgrad_blocks(npix) * 2048 * sizeof(float);
  const size_t tc = conv_k27_wgrad_tc_workspace_bytes();
  return fma > tc ? fma : tc;
}
extern "C" size_t wu_conv_first_wgrad_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return first_wgrad_ws((long long)B * H * W);
}
extern "C" int wu_conv_first_wgrad(const float* x, const void* dy, float* dw, float* db, int B,
                                   int H, int W, void* workspace, size_t workspace_bytes,
                                   wu_stream_t stream) {
  WU_REQUIRE(x && dy && dw && workspace && B > 0 && H > 0 && W > 0, "wu_conv_first_wgrad: bad args");
  WU_REQUIRE(workspace_bytes >= first_wgrad_ws((long long)B * H * W),
             "wu_conv_first_wgrad: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (conv_k27_wgrad_tc_supported(x, W))  // tcgen05 path (wu_conv_first_tc.cu)
    return conv_k27_wgrad_tc(x, dy, dw, db, B, H, W, 1, workspace, st);
  // image rows not 16-byte aligned (TMA cannot fetch them): FMA kernel
  const int blocks = first_wgrad_blocks((long long)B * H * W);
  conv_first_wgrad_kernel<1><<<blocks, 256, 0, st>>>(x, (const bf16*)dy, (float*)workspace, B, H, W);
  WU_CHECK_LAUNCH("conv_first_wgrad_kernel");
  conv_first_wgrad_final_kernel<<<8, 256, 0, st>>>((const float*)workspace, dw, db, blocks);
  WU_CHECK_LAUNCH("conv_first_wgrad_final_kernel");
  return WU_OK;
}
extern "C" int wu_conv_last_tanh_fprop(const void* x, const float* w, const float* bias, float* y,
                                       int B, int H, int W, wu_stream_t stream) {
  WU_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "wu_conv_last_tanh_fprop: bad args");
  const long long npix = (long long)B * H * W;
  conv_last_tanh_fprop_kernel<<<grid_for(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, w, bias, y, npix, (long long)H * W);
  WU_CHECK_LAUNCH("conv_last_tanh_fprop_kernel");
  return WU_OK;
}
static int last_bprop_blocks(long long npix) { return grid_for(npix, 256, 4); }
extern "C" size_t wu_conv_last_tanh_bprop_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)last_bprop_blocks((long long)B * H * W) * 195 * sizeof(float);
}
extern "C" int wu_conv_last_tanh_bprop(const float* gy, const float* y, const void* x,
                                       const float* w, void* gx, float* dw, float* db, int B, int H,
                                       int W, void* workspace, size_t workspace_bytes,
                                       wu_stream_t stream) {
  WU_REQUIRE(gy && y && x && w && gx && dw && workspace && B > 0 && H > 0 && W > 0,
             "wu_conv_last_tanh_bprop: bad args");
  const long long npix = (long long)B * H * W;
  const int blocks = last_bprop_blocks(npix);
  WU_REQUIRE(workspace_bytes >= (size_t)blocks * 195 * sizeof(float),
             "wu_conv_last_tanh_bprop: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  conv_last_tanh_bprop_kernel<<<blocks, 256, 0, st>>>(gy, y, (const bf16*)x, w, (bf16*)gx,
                                                      (float*)workspace, npix, (long long)H * W);
  WU_CHECK_LAUNCH("conv_last_tanh_bprop_kernel");
  conv_last_bprop_final_kernel<<<1, 256, 0, st>>>((const float*)workspace, dw, db, blocks);
  WU_CHECK_LAUNCH("conv_last_bprop_final_kernel");
  return WU_OK;
}

extern "C" int wu_maxpool2_fwd(const void* src, void* dst, int B, int H, int W, int C,
                               wu_stream_t stream) {
  WU_REQUIRE(src && dst && B > 0 && H > 0 && W > 0, "wu_maxpool2_fwd: bad args");
  WU_REQUIRE(C > 0 && C % 8 == 0 && H % 2 == 0 && W % 2 == 0,
             "wu_maxpool2_fwd: need C %% 8 == 0 and even H, W (C=%d H=%d W=%d)", C, H, W);
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)src, (bf16*)dst, B, H, W, C);
  WU_CHECK_LAUNCH("maxpool2_fwd_kernel");
  return WU_OK;
}
extern "C" int wu_maxpool2_bwd(const void* y, const void* g_pool, const void* g_skip, void* g,
                               int B, int H, int W, int C, wu_stream_t stream) {
  WU_REQUIRE(y && g && B > 0 && H > 0 && W > 0, "wu_maxpool2_bwd: bad args");
  WU_REQUIRE(C > 0 && C % 8 == 0 && H % 2 == 0 && W % 2 == 0,
             "wu_maxpool2_bwd: need C %% 8 == 0 and even H, W (C=%d H=%d W=%d)", C, H, W);
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)y, (const bf16*)g_pool, (const bf16*)g_skip, (bf16*)g, B, H, W, C);
  WU_CHECK_LAUNCH("maxpool2_bwd_kernel");
  return WU_OK;
}

extern "C" int wu_adain_stats_chunks(int HW) { return HW <= 0 ? 0 : (HW + kStatChunk - 1) / kStatChunk; }

#define WU_REQUIRE_ADAIN_C(fn, C)                                                           \
  WU_REQUIRE((C) >= 8 && (C) <= 2048 && pow2(C), fn ": C=%d must be a power of two in [8, 2048]", C)

extern "C" int wu_adain_stats(const void* x, float* partial, int B, int HW, int C,
                              wu_stream_t stream) {
  WU_REQUIRE(x && partial && B > 0 && HW > 0, "wu_adain_stats: bad args");
  WU_REQUIRE_ADAIN_C("wu_adain_stats", C);
  const int nchunk = wu_adain_stats_chunks(HW);
  const int lanes = C / 8, groups = 256 / lanes;
  adain_stats_kernel<<<B * nchunk, 256, 2 * groups * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)x, partial, HW, C, nchunk);
  WU_CHECK_LAUNCH("adain_stats_kernel");
  return WU_OK;
}
extern "C" int wu_adain_style_fwd_n(const float* cond, const float* lw, const float* lb,
                                  const float* partial, float* mean, float* rstd, float* ystd,
                                  float* scale, float* shift, int B, int C, int nc, int HW,
                                  int nchunk, float eps, int x_bcast, wu_stream_t stream) {
  WU_REQUIRE(cond && lw && lb && partial && mean && rstd && ystd && scale && shift,
             "wu_adain_style_fwd: null pointer");
  WU_REQUIRE(B > 0 && C > 0 && nc > 0 && HW > 0 && nchunk > 0, "wu_adain_style_fwd: bad shape");
  adain_style_fwd_kernel<<<(B * C + 31) / 32, 32 * kStyleSlices, 0, (cudaStream_t)stream>>>(
      cond, lw, lb, partial, mean, rstd, ystd, scale, shift, B, C, nc, HW,
      nchunk, eps, x_bcast ? 0 : 1);
  WU_CHECK_LAUNCH("adain_style_fwd_kernel");
  return WU_OK;
}
extern "C" int wu_adain_style_fwd(const float* cond, const float* lw, const float* lb,
                                  const float* partial, float* mean, float* rstd, float* ystd,
                                  float* scale, float* shift, int B, int C, int nc, int HW, float eps,
                                  int x_bcast, wu_stream_t stream) {
  return wu_adain_style_fwd_n(cond, lw, lb, partial, mean, rstd, ystd, scale, shift, B, C, nc, HW,
                              wu_adain_stats_chunks(HW), eps, x_bcast, stream);
}
extern "C" int wu_adain_up_drop_fwd(const void* x, const float* scale, const float* shift, void* u,
                                    uint8_t* keep_bits, int B, int h, int w, int C, float p_drop,
                                    uint64_t seed, const uint8_t* mask, int x_bcast,
                                    wu_stream_t stream) {
  return wu_adain_up_drop_fwd_epoch(x, scale, shift, u, keep_bits, B, h, w, C, p_drop, seed, nullptr,
                                    mask, x_bcast, stream);
}
extern "C" int wu_adain_up_drop_fwd_epoch(const void* x, const float* scale, const float* shift, void* u,
                                          uint8_t* keep_bits, int B, int h, int w, int C, float p_drop,
                                          uint64_t seed, const uint32_t* epoch, const uint8_t* mask,
                                          int x_bcast, wu_stream_t stream) {
  WU_REQUIRE(x && scale && shift && u && B > 0 && h > 0 && w > 0, "wu_adain_up_drop_fwd: bad args");
  WU_REQUIRE(C > 0 && C % 8 == 0, "wu_adain_up_drop_fwd: C=%d must be a multiple of 8", C);
  WU_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "wu_adain_up_drop_fwd: p_drop=%f out of [0,1)", p_drop);
  WU_REQUIRE(p_drop == 0.f || keep_bits != nullptr,
             "wu_adain_up_drop_fwd: keep_bits is required when p_drop > 0");
  WU_REQUIRE(B <= 65535 && h / 2 + 1 <= 65535, "wu_adain_up_drop_fwd: B=%d or h=%d exceeds the grid", B, h);
  WU_REQUIRE((long long)4 * h * w * (C / 8) < (1ll << 31),
             "wu_adain_up_drop_fwd: more than 2^31 vectors per image");
  dim3 grid((unsigned)(((w + 1) * (C / 8) + 255) / 256), (unsigned)(h / 2 + 1), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_keep = 1.f / (1.f - p_drop);
  const int xm = x_bcast ? 0 : 1;
  const PhiloxKeys keys = philox_keys(seed);
  // PyTorch's fp32 ratio for align_corners=True (area_pixel_compute_scale): (in - 1) / (out - 1)
  const float rh = 2 * h > 1 ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  const float rw = 2 * w > 1 ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  const int cv = C / 8;
  int cv_shift = -1;
  if ((cv & (cv - 1)) == 0)
    for (cv_shift = 0; (1 << cv_shift) < cv; ++cv_shift) {}
  if (p_drop == 0.f)
    adain_up_drop_fwd_kernel<kDropNone><<<grid, 256, 0, st>>>(
        (const bf16*)x, scale, shift, (bf16*)u, nullptr, h, w, C, 1.f, 0u, keys, nullptr, xm, rh, rw,
        cv_shift, nullptr);
  else if (mask != nullptr)
    adain_up_drop_fwd_kernel<kDropInjected><<<grid, 256, 0, st>>>(
        (const bf16*)x, scale, shift, (bf16*)u, keep_bits, h, w, C, inv_keep, 1u, keys, mask, xm, rh, rw,
        cv_shift, nullptr);
  else  // thr argument = (32768 - thr15) in both 16-bit lanes (philox_keep8)
    adain_up_drop_fwd_kernel<kDropPhilox><<<grid, 256, 0, st>>>(
        (const bf16*)x, scale, shift, (bf16*)u, keep_bits, h, w, C, inv_keep,
        (32768u - dropout_threshold15(p_drop)) * 0x00010001u, keys, nullptr, xm, rh, rw, cv_shift, epoch);
  WU_CHECK_LAUNCH("adain_up_drop_fwd_kernel");
  return WU_OK;
}
extern "C" int wu_adain_bwd_chunks(int h, int w, int C) {
  if (h <= 0 || w <= 0 || C < 8 || C > 2048 || !pow2(C)) return 0;
  const int groups = 256 / (C / 8);
  return ((w + groups - 1) / groups) * ((h + kAdjRows - 1) / kAdjRows);
}
extern "C" int wu_adain_up_drop_bwd(const void* gu, const void* x, const float* mean,
                                    const float* rstd, void* gz, float* partial, int B, int h, int w,
                                    int C, float p_drop, const uint8_t* keep_bits,
                                    wu_stream_t stream) {
  WU_REQUIRE(gu && x && mean && rstd && gz && partial && B > 0 && h > 0 && w > 0,
             "wu_adain_up_drop_bwd: bad args");
  WU_REQUIRE_ADAIN_C("wu_adain_up_drop_bwd", C);
  WU_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "wu_adain_up_drop_bwd: p_drop=%f out of [0,1)", p_drop);
  WU_REQUIRE(p_drop == 0.f || keep_bits != nullptr,
             "wu_adain_up_drop_bwd: keep_bits (from the forward pass) is required when p_drop > 0");
  const int lanes = C / 8, groups = 256 / lanes;
  const int col_blocks = (w + groups - 1) / groups, row_blocks = (h + kAdjRows - 1) / kAdjRows;
  const long long blocks = (long long)B * col_blocks * row_blocks;
  WU_REQUIRE(blocks < (1LL << 31), "wu_adain_up_drop_bwd: too many blocks");
  adain_up_drop_adjoint_kernel<<<(unsigned)blocks, 256, 2 * groups * C * sizeof(float),
                                 (cudaStream_t)stream>>>(
      (const bf16*)gu, p_drop > 0.f ? keep_bits : nullptr, (const bf16*)x, mean, rstd, (bf16*)gz, partial,
      h, w, C, 1.f / (1.f - p_drop), col_blocks, row_blocks);
  WU_CHECK_LAUNCH("adain_up_drop_adjoint_kernel");
  return WU_OK;
}
extern "C" int wu_adain_style_bwd(const float* cond, const float* lw, const float* lb,
                                  const float* partial, int nchunk, const float* ystd,
                                  const float* mean, const float* rstd, float* k1, float* k2,
                                  float* coef, float* gh, float* dlw, float* dlb, int B, int C, int nc,
                                  int HW, wu_stream_t stream) {
  WU_REQUIRE(cond && lw && lb && partial && ystd && mean && rstd && k1 && k2 && coef && gh && dlw &&
                 dlb,
             "wu_adain_style_bwd: null pointer");
  WU_REQUIRE(B > 0 && C > 0 && nc > 0 && HW > 0 && nchunk > 0, "wu_adain_style_bwd: bad shape");
  WU_REQUIRE((B * C) % 4 == 0, "wu_adain_style_bwd: B*C=%d must be a multiple of 4", B * C);
  cudaStream_t st = (cudaStream_t)stream;
  adain_style_bwd_kernel<<<(B * C + 127) / 128, 128, 0, st>>>(cond, lw, lb, partial, ystd, mean, rstd,
                                                              k1, k2, coef, gh, B, C, nc, HW, nchunk);
  WU_CHECK_LAUNCH("adain_style_bwd_kernel");
  adain_style_bwd_params_kernel<<<(4 * C + 127) / 128, 128, 0, st>>>(gh, cond, dlw, dlb, B, 4 * C, nc);
  WU_CHECK_LAUNCH("adain_style_bwd_params_kernel");
  return WU_OK;
}
extern "C" int wu_adain_apply(const void* x, const float* scale, const float* shift, void* out,
                              int B, int HW, int C, wu_stream_t stream) {
  WU_REQUIRE(x && scale && shift && out && B > 0 && HW > 0, "wu_adain_apply: bad args");
  WU_REQUIRE(C > 0 && C % 8 == 0, "wu_adain_apply: C=%d must be a multiple of 8", C);
  const long long total = (long long)B * HW * (C / 8);
  adain_apply_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, scale, shift, (bf16*)out, B, HW, C);
  WU_CHECK_LAUNCH("adain_apply_kernel");
  return WU_OK;
}
extern "C" int wu_adain_bwd_apply(const void* gz, const void* x, const float* coef, void* gx, int B,
                                  int HW, int C, wu_stream_t stream) {
  WU_REQUIRE(gz && x && coef && gx && B > 0 && HW > 0, "wu_adain_bwd_apply: bad args");
  WU_REQUIRE(C > 0 && C % 8 == 0, "wu_adain_bwd_apply: C=%d must be a multiple of 8", C);
  WU_REQUIRE(B <= 65535 && C <= 2048 && 256 % (C / 8) == 0,
             "wu_adain_bwd_apply: B=%d / C=%d outside the kernel's decomposition", B, C);
  const int groups = 256 / (C / 8);
  int bx = (HW + groups * 4 - 1) / (groups * 4);  // four pixels per thread and iteration
  const int cap = (148 * 16 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  adain_bwd_apply_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)gz, (const bf16*)x, coef, (bf16*)gx, B, HW, C);
  WU_CHECK_LAUNCH("adain_bwd_apply_kernel");
  return WU_OK;
}

// ---- discriminator stem, second convolution: Conv2d(3, 64, 3, padding=1, stride=2) + LeakyReLU
extern "C" int wu_conv3to64_s2_fprop(const float* h1, const float* w, const float* bias, float slope,
                                     void* dst, int B, int Hin, int Win, wu_stream_t stream) {
  WU_REQUIRE(h1 && w && dst && B > 0 && Hin > 0 && Win > 0, "wu_conv3to64_s2_fprop: bad args");
  WU_REQUIRE(Hin % 2 == 0 && Win % 2 == 0,
             "wu_conv3to64_s2_fprop: need even Hin and Win (got %d x %d)", Hin, Win);
  return conv_k27_fprop_tc(h1, w, bias, slope, dst, B, Hin, Win, 2, (cudaStream_t)stream);
}
extern "C" size_t wu_conv3to64_s2_wgrad_workspace_bytes(int B, int Hin, int Win) {
  if (B <= 0 || Hin <= 0 || Win <= 0) return 0;
  return first_wgrad_ws((long long)B * (Hin / 2) * (Win / 2));
}
extern "C" int wu_conv3to64_s2_wgrad(const float* h1, const void* g, float* dw, float* db, int B,
                                     int Hin, int Win, void* workspace, size_t workspace_bytes,
                                     wu_stream_t stream) {
  WU_REQUIRE(h1 && g && dw && workspace && B > 0 && Hin > 0 && Win > 0,
             "wu_conv3to64_s2_wgrad: bad args");
  WU_REQUIRE(Hin % 2 == 0 && Win % 2 == 0, "wu_conv3to64_s2_wgrad: need even Hin, Win");
  const int H = Hin / 2, W = Win / 2;
  WU_REQUIRE(workspace_bytes >= first_wgrad_ws((long long)B * H * W),
             "wu_conv3to64_s2_wgrad: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (conv_k27_wgrad_tc_supported(h1, Win))
    return conv_k27_wgrad_tc(h1, g, dw, db, B, Hin, Win, 2, workspace, st);
  const int blocks = first_wgrad_blocks((long long)B * H * W);
  conv_first_wgrad_kernel<2><<<blocks, 256, 0, st>>>(h1, (const bf16*)g, (float*)workspace, B, H, W);
  WU_CHECK_LAUNCH("conv_first_wgrad_kernel<2>");
  conv_first_wgrad_final_kernel<<<8, 256, 0, st>>>((const float*)workspace, dw, db, blocks);
  WU_CHECK_LAUNCH("conv_first_wgrad_final_kernel");
  return WU_OK;
}
