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
  for (long long px = (long long)blockIdx.x * groups + gi; px < npix;
       px += (long long)gridDim.x * groups) {
    const long long off = px * C + l * 8;
    float f[8];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(gy + off)), f);
    if (masked) {
      float yv[8];
      unpack8p(__ldg(reinterpret_cast<const uint4*>(y + off)), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = yv[j] > 0.f ? f[j] : f[j] * slope;
    }
    if (masked || g != gy) *reinterpret_cast<uint4*>(g + off) = pack8p(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
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
