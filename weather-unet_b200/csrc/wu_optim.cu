// wu_optim.cu — multi-tensor Adam: every parameter tensor of a model updated by ONE launch.
// Semantics of torch.optim.Adam as the reference configures it (t_cls_train.py:184-185,
// t_est_train.py:172-173): betas (0, 0.999), eps 1e-8, L2 weight decay lr/20 added to the gradient
// (not AdamW), bias correction, no amsgrad.
#include "wu_host.h"

namespace wu {

struct AdamTensor {  // one record of the device-side table (5 x 8 bytes)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};
struct AdamChunk {  // work item: `count` elements of tensor `t` starting at `start`
  int t;
  int count;
  long long start;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks,
                  float lr, float beta1, float beta2, float eps, float wd, float bc1, float rsqrt_bc2) {
  const AdamChunk ck = chunks[blockIdx.x];
  const AdamTensor T = tensors[ck.t];
  const float step_size = lr / bc1;
  for (int i = threadIdx.x; i < ck.count; i += blockDim.x) {
    const long long j = ck.start + i;
    const float p = T.p[j];
    const float g = fmaf(wd, p, T.g[j]);
    const float m = fmaf(beta1, T.m[j], (1.f - beta1) * g);
    const float v = fmaf(beta2, T.v[j], (1.f - beta2) * g * g);
    T.m[j] = m;
    T.v[j] = v;
    const float denom = sqrtf(v) * rsqrt_bc2 + eps;
    T.p[j] = p - step_size * (m / denom);
  }
}

}  // namespace wu

extern "C" int wu_adam_multi(const void* tensors, const void* chunks, int n_chunks, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step,
                             wu_stream_t stream) {
  WU_REQUIRE(tensors && chunks && n_chunks > 0 && step >= 1, "wu_adam_multi: bad args");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  wu::adam_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      (const wu::AdamTensor*)tensors, (const wu::AdamChunk*)chunks, lr, beta1, beta2, eps,
      weight_decay, (float)bc1, (float)(1.0 / sqrt(bc2)));
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
