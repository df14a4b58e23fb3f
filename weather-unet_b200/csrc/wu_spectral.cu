// wu_spectral.cu — spectral normalisation of every weight of the discriminator in a handful of
// multi-tensor launches.
//
// The reference wraps each discriminator layer in torch.nn.utils.spectral_norm (nets.py:26-33,
// disc.py:21,24): in training mode every forward runs one power iteration on the weight viewed as a
// matrix W [rows = out][cols = in*kh*kw],
//     v <- normalize(W^T u),  u <- normalize(W v),  sigma = u . (W v),  weight = W / sigma
// (normalize(x) = x / max(||x||_2, eps), eps = 1e-12; u, v are buffers updated in place; in eval mode
// the stored u, v are used as they are).  Through PyTorch that is ~25 tiny kernels per layer and
// forward (750 launches, ~2.5 ms per training iteration).  Here all layers go through
//     sn_wtu (t = W^T u, partial ||t||^2)  ->  sn_wv (v = t/||t||, s = W v, partial ||s||^2)
//     ->  sn_sigma (u = s/||s||, sigma)  ->  sn_scale (W_sn = W / sigma)
// and the backward of  W_sn = W / sigma(W)  (sigma differentiated through u v^T with u, v constant,
// exactly what autograd does for the reference) is
//     dW = G / sigma - (<G, W> / sigma^2) u v^T :  sn_dot (partial <G, W>)  ->  sn_bwd.
// Everything is fp32, like the reference.
#include <cuda_bf16.h>

#include "wu_host.h"

namespace wu {

struct SnTensor {   // one record of the device-side table (14 x 8 + 8 bytes)
  const float* w;   // [rows][cols] weight_orig
  float* u;         // [rows] weight_u (updated in place in training mode)
  float* v;         // [cols] weight_v
  float* u_snap;    // [rows] u used by this forward (kept for backward)
  float* v_snap;    // [cols]
  float* sigma;     // [1]
  float* t;         // [cols] scratch: W^T u
  float* s;         // [rows] scratch: W v
  float* part;      // [2 * kSnParts] scratch: partial sums of squares / dots
  float* w_sn;      // [rows][cols] W / sigma
  const float* g;   // backward: gradient w.r.t. w_sn
  float* dw;        // backward: gradient w.r.t. w
  // 3x3 convolution weights that feed the tcgen05 kernels: W / sigma goes straight into the packed bf16
  // operand layouts of wu_pack_conv3x3_weights (w_sn is then not written); null otherwise
  __nv_bfloat16* wf;  // [rows][9 * cin], k = tap * cin + ci
  __nv_bfloat16* wd;  // [cin][9 * rows], k = (8 - tap) * rows + co
  int rows, cols;
};
struct SnChunk {  // work item
  int t;          // tensor
  int begin;      // first column (sn_wtu) / row (sn_wv) / element index in units of 1024 (others)
  int count;
  int index;      // index of this chunk within its tensor (sn_bwd records: number of sn_dot chunks)
};
constexpr int kSnParts = 128;  // partial slots per tensor per reduction
constexpr int kSnCols = 32;    // columns per sn_wtu chunk
constexpr int kSnRows = 8;    // rows per sn_wv chunk

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// Sum over the block (256 threads); result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red /* [8] */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

// t[c] = sum_r W[r][c] u[r] for 32 columns (8 row groups per block, one warp = 128 contiguous bytes
// of a row); part[index] = sum over these columns of t^2
__global__ void __launch_bounds__(256)
sn_wtu_kernel(const SnTensor* __restrict__ tensors, const SnChunk* __restrict__ chunks) {
  __shared__ float grp[8][kSnCols];
  const SnChunk ck = chunks[blockIdx.x];
  const SnTensor T = tensors[ck.t];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = ck.begin + lane;
  float acc = 0.f;
  if (c < T.cols) {
    int r = g;
    for (; r + 24 < T.rows; r += 32) {  // four independent loads in flight
      const float a0 = T.w[(size_t)r * T.cols + c], a1 = T.w[(size_t)(r + 8) * T.cols + c];
      const float a2 = T.w[(size_t)(r + 16) * T.cols + c], a3 = T.w[(size_t)(r + 24) * T.cols + c];
      acc = fmaf(a0, T.u[r], acc);
      acc = fmaf(a1, T.u[r + 8], acc);
      acc = fmaf(a2, T.u[r + 16], acc);
      acc = fmaf(a3, T.u[r + 24], acc);
    }
    for (; r < T.rows; r += 8) acc = fmaf(T.w[(size_t)r * T.cols + c], T.u[r], acc);
  }
  grp[g][lane] = acc;
  __syncthreads();
  if (g == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += grp[i][lane];
    if (c < T.cols) T.t[c] = tot;
    const float sq = warp_sum(c < T.cols ? tot * tot : 0.f);
    if (lane == 0) T.part[ck.index] = sq;
  }
}

// training: v = t / max(||t||, eps) (this chunk writes its share of v), s[r] = W[r] . v for 8 rows;
// eval: v as stored.  part[index] = sum over these rows of s^2.
__global__ void __launch_bounds__(256)
sn_wv_kernel(const SnTensor* __restrict__ tensors, const SnChunk* __restrict__ chunks, int training,
             float eps) {
  __shared__ float red[8];
  const SnChunk ck = chunks[blockIdx.x];
  const SnTensor T = tensors[ck.t];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float inv = 1.f;
  if (training) {
    const int nparts = (T.cols + kSnCols - 1) / kSnCols;
    float n2 = 0.f;
    for (int i = 0; i < nparts; ++i) n2 += T.part[i];  // same order in every thread: identical value
    inv = 1.f / fmaxf(sqrtf(n2), eps);
    // this chunk's share of the normalised v (the chunks of a tensor tile its columns)
    const int nchunks = (T.rows + kSnRows - 1) / kSnRows;
    const int per = (T.cols + nchunks - 1) / nchunks;
    const int c0 = ck.index * per, c1 = min(T.cols, c0 + per);
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
      const float vn = T.t[c] * inv;
      T.v[c] = vn;
      T.v_snap[c] = vn;
    }
  } else {
    const int nchunks = (T.rows + kSnRows - 1) / kSnRows;
    const int per = (T.cols + nchunks - 1) / nchunks;
    const int c0 = ck.index * per, c1 = min(T.cols, c0 + per);
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) T.v_snap[c] = T.v[c];
  }
  const float* vec = training ? T.t : T.v;
  const int r = ck.begin + warp;
  float acc = 0.f;
  if (r < T.rows) {
    const float* wr = T.w + (size_t)r * T.cols;
    for (int c = lane; c < T.cols; c += 32) acc = fmaf(wr[c], vec[c], acc);
    acc = warp_sum(acc) * inv;
    if (lane == 0) T.s[r] = acc;
  }
  // part slots kSnParts.. hold the partial ||s||^2 (the first kSnParts hold ||t||^2 / dot partials)
  const float sq = (lane == 0 && r < T.rows) ? acc * acc : 0.f;
  const float tot = block_sum(sq, red);
  if (threadIdx.x == 0) T.part[kSnParts + ck.index] = tot;
}

// one block per tensor: training: u = s / max(||s||, eps), sigma = u . s; eval: sigma = u_stored . s
__global__ void __launch_bounds__(256)
sn_sigma_kernel(const SnTensor* __restrict__ tensors, int training, float eps) {
  __shared__ float red[8];
  const SnTensor T = tensors[blockIdx.x];
  float sig;
  if (training) {
    const int nchunks = (T.rows + kSnRows - 1) / kSnRows;
    float n2 = 0.f;
    for (int i = 0; i < nchunks; ++i) n2 += T.part[kSnParts + i];
    const float inv = 1.f / fmaxf(sqrtf(n2), eps);
    for (int r = threadIdx.x; r < T.rows; r += blockDim.x) {
      const float un = T.s[r] * inv;
      T.u[r] = un;
      T.u_snap[r] = un;
    }
    sig = n2 * inv;  // sum_r (s_r * inv) * s_r
  } else {
    float acc = 0.f;
    for (int r = threadIdx.x; r < T.rows; r += blockDim.x) {
      const float uu = T.u[r];
      T.u_snap[r] = uu;
      acc = fmaf(uu, T.s[r], acc);
    }
    sig = block_sum(acc, red);
  }
  if (threadIdx.x == 0) T.sigma[0] = sig;
}

// w_sn = w / sigma (1024 elements per chunk unit)
__global__ void __launch_bounds__(256)
sn_scale_kernel(const SnTensor* __restrict__ tensors, const SnChunk* __restrict__ chunks) {
  const SnChunk ck = chunks[blockIdx.x];
  const SnTensor T = tensors[ck.t];
  const float sig = T.sigma[0];
  const size_t n = (size_t)T.rows * T.cols;
  const size_t e0 = (size_t)ck.begin * 1024, e1 = min(n, e0 + (size_t)ck.count * 1024);
  if (T.wf == nullptr) {
    for (size_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) T.w_sn[i] = T.w[i] / sig;
    return;
  }
  const int cout = T.rows, cin = T.cols / 9;
  for (size_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) {  // i = (co * cin + ci) * 9 + tap
    const int tap = (int)(i % 9);
    const size_t cc = i / 9;
    const int ci = (int)(cc % cin), co = (int)(cc / cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(T.w[i] / sig);
    T.wf[(size_t)co * T.cols + (size_t)tap * cin + ci] = v;
    T.wd[(size_t)ci * 9 * cout + (size_t)(8 - tap) * cout + co] = v;
  }
}

// backward, step 1: part[index] = sum over the chunk of g * w
__global__ void __launch_bounds__(256)
sn_dot_kernel(const SnTensor* __restrict__ tensors, const SnChunk* __restrict__ chunks) {
  __shared__ float red[8];
  const SnChunk ck = chunks[blockIdx.x];
  const SnTensor T = tensors[ck.t];
  const size_t n = (size_t)T.rows * T.cols;
  const size_t e0 = (size_t)ck.begin * 1024, e1 = min(n, e0 + (size_t)ck.count * 1024);
  float acc = 0.f;
  for (size_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) acc = fmaf(T.g[i], T.w[i], acc);
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) T.part[ck.index] = tot;
}
// backward, step 2: dw = g / sigma - (<g, w> / sigma^2) u v^T
__global__ void __launch_bounds__(256)
sn_bwd_kernel(const SnTensor* __restrict__ tensors, const SnChunk* __restrict__ chunks) {
  const SnChunk ck = chunks[blockIdx.x];
  const SnTensor T = tensors[ck.t];
  const size_t n = (size_t)T.rows * T.cols;
  float dot = 0.f;
  for (int i = 0; i < ck.index; ++i) dot += T.part[i];  // bwd records carry the tensor's dot-chunk count
  const float sig = T.sigma[0];
  const float coef = dot / (sig * sig);
  const float isig = 1.f / sig;
  const size_t e0 = (size_t)ck.begin * 1024, e1 = min(n, e0 + (size_t)ck.count * 1024);
  for (size_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
    const int r = (int)(i / T.cols), c = (int)(i - (size_t)r * T.cols);
    T.dw[i] = T.g[i] * isig - coef * T.u_snap[r] * T.v_snap[c];
  }
}

}  // namespace wu

using namespace wu;

extern "C" int wu_sn_forward(const void* tensors, int n_tensors, const void* wtu_chunks,
                             int n_wtu_chunks, const void* wv_chunks, int n_wv_chunks,
                             const void* elem_chunks, int n_elem_chunks, int training, float eps,
                             wu_stream_t stream) {
  WU_REQUIRE(tensors && wv_chunks && elem_chunks && n_tensors > 0 && n_wv_chunks > 0 && n_elem_chunks > 0,
             "wu_sn_forward: bad args");
  WU_REQUIRE(!training || (wtu_chunks && n_wtu_chunks > 0), "wu_sn_forward: training needs the W^T u chunks");
  cudaStream_t st = (cudaStream_t)stream;
  const SnTensor* T = (const SnTensor*)tensors;
  if (training) {
    sn_wtu_kernel<<<n_wtu_chunks, 256, 0, st>>>(T, (const SnChunk*)wtu_chunks);
    WU_CHECK_LAUNCH("sn_wtu_kernel");
  }
  sn_wv_kernel<<<n_wv_chunks, 256, 0, st>>>(T, (const SnChunk*)wv_chunks, training, eps);
  WU_CHECK_LAUNCH("sn_wv_kernel");
  sn_sigma_kernel<<<n_tensors, 256, 0, st>>>(T, training, eps);
  WU_CHECK_LAUNCH("sn_sigma_kernel");
  sn_scale_kernel<<<n_elem_chunks, 256, 0, st>>>(T, (const SnChunk*)elem_chunks);
  WU_CHECK_LAUNCH("sn_scale_kernel");
  return WU_OK;
}

extern "C" int wu_sn_backward(const void* tensors, const void* dot_chunks, int n_dot_chunks,
                              const void* bwd_chunks, int n_bwd_chunks, wu_stream_t stream) {
  WU_REQUIRE(tensors && dot_chunks && bwd_chunks && n_dot_chunks > 0 && n_bwd_chunks > 0,
             "wu_sn_backward: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  sn_dot_kernel<<<n_dot_chunks, 256, 0, st>>>((const SnTensor*)tensors, (const SnChunk*)dot_chunks);
  WU_CHECK_LAUNCH("sn_dot_kernel");
  sn_bwd_kernel<<<n_bwd_chunks, 256, 0, st>>>((const SnTensor*)tensors, (const SnChunk*)bwd_chunks);
  WU_CHECK_LAUNCH("sn_bwd_kernel");
  return WU_OK;
}

extern "C" int wu_sn_parts(void) { return 2 * kSnParts; }
extern "C" int wu_sn_wtu_cols(void) { return kSnCols; }
extern "C" int wu_sn_wv_rows(void) { return kSnRows; }
