// wu_optim.cu — multi-tensor Adam: every parameter tensor of a model updated by ONE launch.
// Semantics of torch.optim.Adam as the reference configures it (t_cls_train.py:184-185,
// t_est_train.py:172-173): betas (0, 0.999), eps 1e-8, L2 weight decay lr/20 added to the gradient
// (not AdamW), bias correction, no amsgrad.
#include "wu_host.h"

#include <cuda_bf16.h>

namespace wu {

// One record of the device-side tensor table (64 bytes).  wf / wd are NULL for an ordinary tensor;
// for a 3x3 convolution weight [cout][cin][3][3] they are the bf16 operand layouts of
// wu_pack_conv3x3_weights, rewritten by the same launch that updates the fp32 master.
struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  __nv_bfloat16* wf;  // [cout][9*cin], k = tap*cin + ci
  __nv_bfloat16* wd;  // [cin][9*cout], k = (8-tap)*cout + co
  int cout, cin;
};
static_assert(sizeof(AdamTensor) == 64, "AdamTensor layout is part of the C ABI");
struct AdamChunk {  // work item: `count` elements of tensor `t` starting at `start`;
  int t;            // packed tensors: count == 0 and start = tile index (16 co x 64 ci x 9 taps)
  int count;
  long long start;
};
// Device-resident step state (so that the update can sit inside a CUDA graph): advanced by
// adam_tick_kernel before every update.
struct AdamState {
  int step;
  float bc1;        // 1 - beta1^step
  float rsqrt_bc2;  // 1 / sqrt(1 - beta2^step)
  int pad;
};

__global__ void adam_tick_kernel(AdamState* st, float beta1, float beta2) {
  const int s = st->step + 1;
  st->step = s;
  st->bc1 = (float)(1.0 - pow((double)beta1, (double)s));
  st->rsqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)s)));
}

constexpr int kPackCo = 16, kPackCi = 64;  // tile of a packed weight handled by one CTA

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float beta1,
                                          float beta2, float eps, float wd, float step_size,
                                          float rsqrt_bc2) {
  g = fmaf(wd, p, g);
  m = fmaf(beta1, m, (1.f - beta1) * g);
  v = fmaf(beta2, v, (1.f - beta2) * g * g);
  const float denom = sqrtf(v) * rsqrt_bc2 + eps;
  p = p - step_size * (m / denom);
  return p;
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks,
                  float lr, float beta1, float beta2, float eps, float wd, float bc1_host,
                  float rsqrt_bc2_host, const AdamState* __restrict__ state) {
  __shared__ __align__(16) __nv_bfloat16 tile[kPackCo * 9 * kPackCi];  // [co][tap][ci], 18 KB
  const AdamChunk ck = chunks[blockIdx.x];
  const AdamTensor T = tensors[ck.t];
  const float bc1 = state ? state->bc1 : bc1_host;
  const float rsqrt_bc2 = state ? state->rsqrt_bc2 : rsqrt_bc2_host;
  const float step_size = lr / bc1;
  if (T.wf == nullptr) {
    for (int i = threadIdx.x; i < ck.count; i += blockDim.x) {
      const long long j = ck.start + i;
      float p = T.p[j], m = T.m[j], v = T.v[j];
      adam_one(p, T.g[j], m, v, beta1, beta2, eps, wd, step_size, rsqrt_bc2);
      T.m[j] = m;
      T.v[j] = v;
      T.p[j] = p;
    }
    return;
  }
  // 3x3 convolution weight: tile (co0 .. co0+15) x (ci0 .. ci0+63) x 9 taps.  The fp32 master is
  // contiguous over (ci, tap) for one co: 576 floats per row of the tile.
  const int ci_tiles = T.cin / kPackCi;
  const int co0 = (int)(ck.start / ci_tiles) * kPackCo;
  const int ci0 = (int)(ck.start % ci_tiles) * kPackCi;
  constexpr int kRow = kPackCi * 9;
  for (int e = threadIdx.x; e < kPackCo * kRow; e += blockDim.x) {
    const int co_l = e / kRow, r = e - co_l * kRow;
    const int ci_l = r / 9, tap = r - ci_l * 9;
    const long long j = ((long long)(co0 + co_l) * T.cin + ci0) * 9 + r;
    float p = T.p[j], m = T.m[j], v = T.v[j];
    adam_one(p, T.g[j], m, v, beta1, beta2, eps, wd, step_size, rsqrt_bc2);
    T.m[j] = m;
    T.v[j] = v;
    T.p[j] = p;
    tile[(co_l * 9 + tap) * kPackCi + ci_l] = __float2bfloat16_rn(p);
  }
  __syncthreads();
  // w_fprop: per (co, tap) 64 consecutive ci = 128 bytes = 8 x 16 bytes
  for (int e = threadIdx.x; e < kPackCo * 9 * 8; e += blockDim.x) {
    const int q = e & 7, row = e >> 3;  // row = co_l*9 + tap
    const int co_l = row / 9, tap = row - co_l * 9;
    const uint4 val = *reinterpret_cast<const uint4*>(&tile[row * kPackCi + q * 8]);
    *reinterpret_cast<uint4*>(T.wf + ((long long)(co0 + co_l) * 9 + tap) * T.cin + ci0 + q * 8) = val;
  }
  // w_dgrad: per (ci, tap) 16 consecutive co = 32 bytes = 2 x 16 bytes
  if (T.wd != nullptr) {
    for (int e = threadIdx.x; e < kPackCi * 9 * 2; e += blockDim.x) {
      const int ci_l = e & (kPackCi - 1);
      const int rest = e >> 6;  // tap*2 + half
      const int half = rest & 1, tap = rest >> 1;
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = tile[((half * 8 + k) * 9 + tap) * kPackCi + ci_l];
      *reinterpret_cast<uint4*>(T.wd + ((long long)(ci0 + ci_l) * 9 + (8 - tap)) * T.cout + co0 +
                                half * 8) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

}  // namespace wu

extern "C" int wu_adam_pack_tile(int* co, int* ci) {
  if (co) *co = wu::kPackCo;
  if (ci) *ci = wu::kPackCi;
  return WU_OK;
}

extern "C" int wu_adam_multi(const void* tensors, const void* chunks, int n_chunks, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step,
                             void* step_state, wu_stream_t stream) {
  WU_REQUIRE(tensors && chunks && n_chunks > 0, "wu_adam_multi: bad args");
  WU_REQUIRE(step_state != nullptr || step >= 1, "wu_adam_multi: step must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  double bc1 = 1.0, bc2 = 1.0;
  if (step_state != nullptr) {
    wu::adam_tick_kernel<<<1, 1, 0, st>>>((wu::AdamState*)step_state, beta1, beta2);
    WU_CHECK_LAUNCH("adam_tick_kernel");
  } else {
    bc1 = 1.0 - pow((double)beta1, (double)step);
    bc2 = 1.0 - pow((double)beta2, (double)step);
  }
  wu::adam_multi_kernel<<<n_chunks, 256, 0, st>>>(
      (const wu::AdamTensor*)tensors, (const wu::AdamChunk*)chunks, lr, beta1, beta2, eps,
      weight_decay, (float)bc1, (float)(1.0 / sqrt(bc2)), (const wu::AdamState*)step_state);
  WU_CHECK_LAUNCH("adam_multi_kernel");
  return WU_OK;
}

// ------------------------------------------------------------------------------------------------
// per-sample L1 distance (the reconstruction terms of the generator loss)
// ------------------------------------------------------------------------------------------------
// t_cls_train.py:255,259-266: g_loss_l1 = F.l1_loss(fake, images) (logged) and
// loss_con = mean_b( mean_chw |fake_b - images_b| / (lambda_b + eps) ).  Through PyTorch that is
// sub, abs, mean (three passes over two 50 MB tensors plus temporaries) and sign, mul, div, expand in
// the backward pass; here one pass forward (d[b] = mean |a_b - b_b|) and one pass backward
// (ga = sign(a - b) * gd[b] / n).  fp32, n elements per sample.
namespace wu {
constexpr int kL1Blocks = 64;  // blocks per sample

__global__ void __launch_bounds__(256)
l1_per_sample_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                             float* __restrict__ partial, long long n) {
  __shared__ float red[8];
  const int s = blockIdx.y;
  const float4* pa = reinterpret_cast<const float4*>(a + (size_t)s * n);
  const float4* pb = reinterpret_cast<const float4*>(b + (size_t)s * n);
  const long long n4 = n >> 2;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(pa + i), y = __ldg(pb + i);
    acc += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
      acc += fabsf(a[(size_t)s * n + i] - b[(size_t)s * n + i]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[s * kL1Blocks + blockIdx.x] = t;
  }
}
__global__ void l1_per_sample_final_kernel(const float* __restrict__ partial, float* __restrict__ d,
                                           int B, long long n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  float t = 0.f;
  for (int i = 0; i < kL1Blocks; ++i) t += partial[s * kL1Blocks + i];
  d[s] = t / (float)n;
}
__global__ void __launch_bounds__(256)
l1_per_sample_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                         const float* __restrict__ gd, float* __restrict__ ga, long long n) {
  const int s = blockIdx.y;
  const float coef = gd[s] / (float)n;
  const size_t base = (size_t)s * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float df = a[base + i] - b[base + i];
    ga[base + i] = df > 0.f ? coef : (df < 0.f ? -coef : 0.f);  // torch.sign: 0 at 0
  }
}
}  // namespace wu

extern "C" size_t wu_l1_per_sample_workspace_bytes(int B) {
  return B > 0 ? (size_t)B * wu::kL1Blocks * sizeof(float) : 0;
}
extern "C" int wu_l1_per_sample_fwd(const float* a, const float* b, float* d, int B, long long n,
                                    void* workspace, size_t workspace_bytes, wu_stream_t stream) {
  WU_REQUIRE(a && b && d && workspace && B > 0 && B <= 65535 && n > 0, "wu_l1_per_sample_fwd: bad args");
  WU_REQUIRE(workspace_bytes >= wu_l1_per_sample_workspace_bytes(B), "wu_l1_per_sample_fwd: workspace too small");
  WU_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 && n % 4 == 0,
             "wu_l1_per_sample_fwd: 16-byte aligned tensors with n %% 4 == 0 required");
  cudaStream_t st = (cudaStream_t)stream;
  wu::l1_per_sample_partial_kernel<<<dim3(wu::kL1Blocks, B), 256, 0, st>>>(a, b, (float*)workspace, n);
  WU_CHECK_LAUNCH("l1_per_sample_partial_kernel");
  wu::l1_per_sample_final_kernel<<<(B + 127) / 128, 128, 0, st>>>((const float*)workspace, d, B, n);
  WU_CHECK_LAUNCH("l1_per_sample_final_kernel");
  return WU_OK;
}
extern "C" int wu_l1_per_sample_bwd(const float* a, const float* b, const float* gd, float* ga, int B,
                                    long long n, wu_stream_t stream) {
  WU_REQUIRE(a && b && gd && ga && B > 0 && B <= 65535 && n > 0, "wu_l1_per_sample_bwd: bad args");
  wu::l1_per_sample_bwd_kernel<<<dim3(wu::kL1Blocks, B), 256, 0, (cudaStream_t)stream>>>(a, b, gd, ga, n);
  WU_CHECK_LAUNCH("l1_per_sample_bwd_kernel");
  return WU_OK;
}
