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
