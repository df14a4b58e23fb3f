// wu_pointwise.cu — bias + LeakyReLU epilogue kernels for NHWC bf16 tensors, forward (in place)
// and backward (masked gradient + deterministic bias gradient in one pass).  They replace the
// separate bias-add, LeakyReLU, LeakyReLU-backward and bias-gradient-reduction kernels that follow
// the spectral-norm convolutions of the reference discriminator (nets.py:26-33, disc.py:28-31)
// when those convolutions run through a library that does not fuse them.
#include "wu_host.h"
#include "wu_ptx.cuh"

namespace wu {

__device__ __forceinline__ void unpack8p(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x);
  f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z);
  f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8p(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

__global__ void __launch_bounds__(256)
bias_act_fwd_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ bias, float slope,
                    long long nvec, int cv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + v * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + v * 8) + 1);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float f[8];
    unpack8p(*reinterpret_cast<const uint4*>(x + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = f[j] + bb[j];
      f[j] = t > 0.f ? t : t * slope;
    }
    *reinterpret_cast<uint4*>(x + i * 8) = pack8p(f);
  }
}

// g = gy * (y > 0 ? 1 : slope); partial[block][c] = sum over the block's pixels of g
__global__ void __launch_bounds__(256)
bias_act_bwd_kernel(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                    __nv_bfloat16* __restrict__ g, float* __restrict__ partial, float slope,
                    long long npix, int C) {
  extern __shared__ float red[];  // [groups][C]
  const int lanes = C / 8;
  const int groups = blockDim.x / lanes;
  const int gi = threadIdx.x / lanes, l = threadIdx.x % lanes;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool masked = slope != 1.f;
  const long long stride = (long long)gridDim.x * groups;
  auto one = [&](const uint4& gq, const uint4& yq, long long off) {
    float f[8];
    unpack8p(gq, f);
    if (masked) {
      float yv[8];
      unpack8p(yq, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = yv[j] > 0.f ? f[j] : f[j] * slope;
    }
    if (masked || g != gy) *reinterpret_cast<uint4*>(g + off) = pack8p(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  };
  long long px = (long long)blockIdx.x * groups + gi;
  for (; px + 3 * stride < npix; px += 4 * stride) {  // eight independent 16-byte loads in flight
    uint4 gq[4], yq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long off = (px + u * stride) * C + l * 8;
      gq[u] = __ldg(reinterpret_cast<const uint4*>(gy + off));
      yq[u] = masked ? __ldg(reinterpret_cast<const uint4*>(y + off)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(gq[u], yq[u], (px + u * stride) * C + l * 8);
  }
  for (; px < npix; px += stride) {
    const long long off = px * C + l * 8;
    one(__ldg(reinterpret_cast<const uint4*>(gy + off)),
        masked ? __ldg(reinterpret_cast<const uint4*>(y + off)) : make_uint4(0, 0, 0, 0), off);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[gi * C + l * 8 + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[gg * C + c];
    partial[(size_t)blockIdx.x * C + c] = s;
  }
}
// db[c] = sum_b partial[b][c]: one block per 32 channels, 8 threads share a channel's block range
__global__ void __launch_bounds__(256)
bias_act_bwd_final_kernel(const float* __restrict__ partial, float* __restrict__ db, int nblocks,
                          int C) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (int b = part; b < nblocks; b += 8) s += partial[(size_t)b * C + c];
  red[part][cl] = s;
  __syncthreads();
  if (part == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cl];
    db[c] = t;
  }
}


// ------------------------------------------------------------------------------------------------
// discriminator stem (disc.py:28 / nets.py:26-33 with in_channels = 3):
//   x (B,3,H,W) fp32 NCHW -> h1 = Conv2d(3,3,3,padding=1)(x)  (fp32 NCHW, no activation)
//   -> c1 = LeakyReLU(Conv2d(3,64,3,padding=1,stride=2)(h1))   (NHWC bf16, wu_conv3to64_s2_*)
// Three-channel tensors are FMA / HBM work, not tensor-core work (K = 27).
// ------------------------------------------------------------------------------------------------
// One thread = a strip of 4 consecutive output pixels of a row: the 3 x 3 x 6 input values of the
// strip are loaded once and reused by the 4 pixels, and each of the 81 weights is read from shared
// memory once per strip (a first version, one pixel per thread, re-read all 81 per pixel and was
// bound by shared-memory loads: 100 us for 100 MB of traffic).
// FLIP: use w'[ci][co][r][s] = w[co][ci][2-r][2-s], i.e. the transposed convolution = data gradient.
template <bool FLIP>
__global__ void __launch_bounds__(256)
conv3to3_strip_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W) {
  __shared__ float ws[81];  // [co][ci][r][s] of the convolution actually applied
  __shared__ float bs[3];
  if (threadIdx.x < 81) {
    const int co = threadIdx.x / 27, rem = threadIdx.x % 27, ci = rem / 9, t = rem % 9;
    ws[threadIdx.x] = FLIP ? w[ci * 27 + co * 9 + (8 - t)] : w[threadIdx.x];
  }
  if (threadIdx.x < 3) bs[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int W4 = (W + 3) >> 2;
  const long long HW = (long long)H * W, nstrips = (long long)B * H * W4;
  const bool vec = (W & 3) == 0;
  for (long long st = blockIdx.x * (long long)blockDim.x + threadIdx.x; st < nstrips;
       st += (long long)gridDim.x * blockDim.x) {
    const int w0 = (int)(st % W4) * 4;
    const long long t = st / W4;
    const int hq = (int)(t % H);
    const long long b = t / H;
    float acc[3][4];
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[co][p] = bs[co];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = hq + r - 1;
        if (hh < 0 || hh >= H) continue;
        const float* row = x + (b * 3 + ci) * HW + (long long)hh * W;
        float v[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const int ww = w0 + c - 1;
          v[c] = (ww >= 0 && ww < W) ? __ldg(row + ww) : 0.f;
        }
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
          for (int co = 0; co < 3; ++co) {
            const float wv = ws[co * 27 + ci * 9 + r * 3 + s3];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[co][p] = fmaf(v[p + s3], wv, acc[co][p]);
          }
      }
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      float* o = y + (b * 3 + co) * HW + (long long)hq * W + w0;
      if (vec) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[co][0], acc[co][1], acc[co][2], acc[co][3]);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (w0 + p < W) o[p] = acc[co][p];
      }
    }
  }
}

// Weight / bias gradient of Conv2d(3,3,3,padding=1): per-block partial sums of
// dw0[co][ci][r][s] = sum g_h1[co] * x[ci][shifted], db0[co] = sum g_h1[co]; same 4-pixel strips,
// 84 accumulators per thread.
constexpr int kStemBwdBlocks = 148 * 4;
__global__ void __launch_bounds__(256, 2)
conv3to3_wgrad_kernel(const float* __restrict__ gh, const float* __restrict__ x,
                      float* __restrict__ partial, int B, int H, int W) {
  __shared__ float red[8][84];
  const int W4 = (W + 3) >> 2;
  const long long HW = (long long)H * W, nstrips = (long long)B * H * W4;
  float dw[81];
  float db[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 81; ++i) dw[i] = 0.f;
  for (long long st = blockIdx.x * (long long)blockDim.x + threadIdx.x; st < nstrips;
       st += (long long)gridDim.x * blockDim.x) {
    const int w0 = (int)(st % W4) * 4;
    const long long t = st / W4;
    const int hq = (int)(t % H);
    const long long b = t / H;
    float gv[3][4];
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      const float* grow = gh + (b * 3 + co) * HW + (long long)hq * W + w0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        gv[co][p] = (w0 + p < W) ? __ldg(grow + p) : 0.f;
        db[co] += gv[co][p];
      }
    }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = hq + r - 1;
        if (hh < 0 || hh >= H) continue;
        const float* row = x + (b * 3 + ci) * HW + (long long)hh * W;
        float v[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const int ww = w0 + c - 1;
          v[c] = (ww >= 0 && ww < W) ? __ldg(row + ww) : 0.f;
        }
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
          for (int co = 0; co < 3; ++co) {
            float a = dw[co * 27 + ci * 9 + r * 3 + s3];
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fmaf(gv[co][p], v[p + s3], a);
            dw[co * 27 + ci * 9 + r * 3 + s3] = a;
          }
      }
  }
  // block reduction of the 84 partial sums: warp shuffles, then 8 warps through shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 84; ++i) {
    float v = i < 81 ? dw[i] : db[i - 81];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 84) {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sum += red[wv][threadIdx.x];
    partial[(size_t)blockIdx.x * 84 + threadIdx.x] = sum;
  }
}
__global__ void conv3to3_bprop_final_kernel(const float* __restrict__ partial,
                                            float* __restrict__ dw, float* __restrict__ db,
                                            int nblocks) {
  const int i = threadIdx.x;
  if (i >= 84) return;
  float sum = 0.f;
  for (int b = 0; b < nblocks; ++b) sum += partial[(size_t)b * 84 + i];
  if (i < 81) dw[i] = sum;
  else if (db != nullptr) db[i - 81] = sum;
}

constexpr int kBiasActBlocks = 148 * 4;

}  // namespace wu

using namespace wu;

static bool bias_act_c_ok(int C) { return C >= 8 && C <= 2048 && (C & (C - 1)) == 0; }

extern "C" int wu_bias_act_fwd(void* x, const float* bias, float slope, long long npix, int C,
                               wu_stream_t stream) {
  WU_REQUIRE(x && bias && npix > 0, "wu_bias_act_fwd: bad args");
  WU_REQUIRE(bias_act_c_ok(C), "wu_bias_act_fwd: C=%d must be a power of two in [8, 2048]", C);
  const long long nvec = npix * (C / 8);
  long long g = (nvec + 255) / 256;
  if (g > 148LL * 32) g = 148LL * 32;
  bias_act_fwd_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, bias, slope, nvec,
                                                               C / 8);
  WU_CHECK_LAUNCH("bias_act_fwd_kernel");
  return WU_OK;
}
extern "C" size_t wu_bias_act_bwd_workspace_bytes(int C) {
  return C > 0 ? (size_t)kBiasActBlocks * C * sizeof(float) : 0;
}
extern "C" int wu_bias_act_bwd(const void* gy, const void* y, void* g, float* db, float slope,
                               long long npix, int C, void* workspace, size_t workspace_bytes,
                               wu_stream_t stream) {
  WU_REQUIRE(gy && y && g && db && workspace && npix > 0, "wu_bias_act_bwd: bad args");
  WU_REQUIRE(bias_act_c_ok(C), "wu_bias_act_bwd: C=%d must be a power of two in [8, 2048]", C);
  WU_REQUIRE(workspace_bytes >= wu_bias_act_bwd_workspace_bytes(C),
             "wu_bias_act_bwd: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int lanes = C / 8, groups = 256 / lanes;
  WU_REQUIRE(groups >= 1, "wu_bias_act_bwd: C=%d too wide", C);
  bias_act_bwd_kernel<<<kBiasActBlocks, 256, groups * C * sizeof(float), st>>>(
      (const __nv_bfloat16*)gy, (const __nv_bfloat16*)y, (__nv_bfloat16*)g, (float*)workspace, slope,
      npix, C);
  WU_CHECK_LAUNCH("bias_act_bwd_kernel");
  bias_act_bwd_final_kernel<<<(C + 31) / 32, 256, 0, st>>>((const float*)workspace, db,
                                                           kBiasActBlocks, C);
  WU_CHECK_LAUNCH("bias_act_bwd_final_kernel");
  return WU_OK;
}

// ---- discriminator stem entry points
static int strip_grid(int B, int H, int W) {
  const long long nstrips = (long long)B * H * ((W + 3) / 4);
  long long g = (nstrips + 255) / 256;
  if (g > 148LL * 32) g = 148LL * 32;
  return (int)g;
}
extern "C" int wu_conv3to3_fprop(const float* x, const float* w, const float* bias, float* y, int B,
                                 int H, int W, wu_stream_t stream) {
  WU_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "wu_conv3to3_fprop: bad args");
  conv3to3_strip_kernel<false><<<strip_grid(B, H, W), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B,
                                                                                    H, W);
  WU_CHECK_LAUNCH("conv3to3_strip_kernel");
  return WU_OK;
}
extern "C" size_t wu_conv3to3_bprop_workspace_bytes(void) {
  return (size_t)kStemBwdBlocks * 84 * sizeof(float);
}
extern "C" int wu_conv3to3_bprop(const float* g_h1, const float* x, const float* w, float* g_x,
                                 float* dw, float* db, int B, int H, int W, void* workspace,
                                 size_t workspace_bytes, wu_stream_t stream) {
  WU_REQUIRE(g_h1 && x && w && B > 0 && H > 0 && W > 0, "wu_conv3to3_bprop: bad args");
  WU_REQUIRE(g_x || dw, "wu_conv3to3_bprop: nothing to compute");
  cudaStream_t st = (cudaStream_t)stream;
  if (dw != nullptr) {
    WU_REQUIRE(workspace && workspace_bytes >= wu_conv3to3_bprop_workspace_bytes(),
               "wu_conv3to3_bprop: workspace %zu too small", workspace_bytes);
    int blocks = strip_grid(B, H, W);
    if (blocks > kStemBwdBlocks) blocks = kStemBwdBlocks;
    conv3to3_wgrad_kernel<<<blocks, 256, 0, st>>>(g_h1, x, (float*)workspace, B, H, W);
    WU_CHECK_LAUNCH("conv3to3_wgrad_kernel");
    conv3to3_bprop_final_kernel<<<1, 96, 0, st>>>((const float*)workspace, dw, db, blocks);
    WU_CHECK_LAUNCH("conv3to3_bprop_final_kernel");
  }
  if (g_x != nullptr) {  // data gradient = the same convolution with transposed, flipped weights
    conv3to3_strip_kernel<true><<<strip_grid(B, H, W), 256, 0, st>>>(g_h1, w, nullptr, g_x, B, H, W);
    WU_CHECK_LAUNCH("conv3to3_strip_kernel<flip>");
  }
  return WU_OK;
}
