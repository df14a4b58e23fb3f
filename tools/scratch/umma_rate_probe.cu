// Probe: intrinsic issue rate of tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128, K = 16, SS mode)
// as a function of N, with operands that walk through shared memory the way the convolution kernels'
// operands do.  One CTA per SM (148), one thread issues `iters` groups of G MMAs into alternating TMEM
// accumulators, one commit + wait at the end; cycles per MMA = elapsed / (iters * G).
// Question it answers: is conv3x3_igemm_v2<128,2> (74 % tensor pipe) / <64,4> (51 %) losing time in its
// own pipeline (barrier waits, TMA), or is that the tensor core's rate for these shapes?
//   variants: N in {64, 128, 256};  A walk: `a_tiles` distinct 16 KB tiles;  B walk: `b_tiles` tiles;
//   mn = 1: both operands MN-major (the weight-gradient kernels);  tma = 1: a second warp keeps
//   bulk-copying global -> shared (no dependency) to add the fill traffic of the real kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -o tools/scratch/umma_rate_probe \
//        tools/scratch/umma_rate_probe.cu weather-unet_b200/csrc/wu_host.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../weather-unet_b200/csrc/wu_host.h"
#include "../../weather-unet_b200/csrc/wu_ptx.cuh"
using namespace wu;

constexpr int kATile = 16384;       // 128 rows x 64 k x 2 B
constexpr int kMaxSmem = 200 * 1024;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int N, int MN>
__global__ void __launch_bounds__(128, 1)
rate_kernel(long long* cycles, int iters, int a_tiles, int b_tiles, int tma, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t raw = smem_u32(sm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* s = sm + (base - raw);
  constexpr int kBTile = N * 128;  // N rows x 64 k x 2 B
  const uint32_t a_base = base, b_base = base + a_tiles * kATile;
  const uint32_t fill_base = b_base + b_tiles * kBTile;  // 16 KB scratch the bulk copies land in
  const uint32_t bar = fill_base + 16384, done = bar + 8, slot = bar + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (a_tiles * kATile + b_tiles * kBTile) / 16; i += 128)
    reinterpret_cast<uint4*>(s)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(s + (slot - base));
  __shared__ volatile int stop;
  if (tid == 0) stop = 0;
  __syncthreads();
  if (warp == 1 && tma) {
    if ((tid & 31) == 0) {  // unrelated fill traffic: 16 KB per round trip, as fast as it completes
      uint32_t ph = 0;
      const uint8_t* src = gsrc + (size_t)blockIdx.x * (1 << 20);
      int j = 0;
      while (!stop) {
        mbar_arrive_expect_tx(bar, 16384);
        bulk_g2s(fill_base, src + (size_t)(j & 63) * 16384, 16384, bar);
        mbar_wait(bar, ph);
        ph ^= 1u;
        ++j;
      }
    }
  } else if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MN, MN);
    // K-major: k step = 32 B inside the 128-byte row; MN-major: k step of 16 = two 8-row groups = 2 KB
    const uint64_t ad0 = MN ? umma_smem_desc_sw128(a_base, 8192, 1024) : umma_smem_desc_sw128(a_base, 16, 1024);
    const uint64_t bd0 = MN ? umma_smem_desc_sw128(b_base, 8192, 1024) : umma_smem_desc_sw128(b_base, 16, 1024);
    const uint32_t kstep = MN ? 2048u : 32u;
    int ai = 0, bi = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t ad = ad0 + (uint64_t)((ai * kATile) >> 4);
      const uint64_t bd = bd0 + (uint64_t)((bi * kBTile) >> 4);
#pragma unroll
      for (int t = 0; t < 512 / N / 2; ++t) {  // T accumulators share B (as the conv kernels' stacked tiles)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + t * N, ad + (uint64_t)((k * kstep) >> 4), bd + (uint64_t)((k * kstep) >> 4), idesc,
                    (it | k) ? 1u : 0u);
      }
      if (++ai == a_tiles) ai = 0;
      if (++bi == b_tiles) bi = 0;
    }
    umma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// CTA-pair variant: M = 256 across two CTAs, each supplying its own 128 rows of A and HALF of B.
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
rate_pair_kernel(long long* cycles, int iters, int a_tiles, int b_tiles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t raw = smem_u32(sm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* s = sm + (base - raw);
  constexpr int kBTile = (N / 2) * 128;  // this CTA's half: N/2 rows x 64 k x 2 B
  const uint32_t a_base = base, b_base = base + a_tiles * kATile;
  const uint32_t done = b_base + b_tiles * kBTile, slot = done + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = cluster_ctarank() == 0;
  for (int i = tid; i < (a_tiles * kATile + b_tiles * kBTile) / 16; i += 128)
    reinterpret_cast<uint4*>(s)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(s + (slot - base));
  if (leader && tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const uint64_t ad0 = umma_smem_desc_sw128(a_base, 16, 1024);
    const uint64_t bd0 = umma_smem_desc_sw128(b_base, 16, 1024);
    int ai = 0, bi = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t ad = ad0 + (uint64_t)((ai * kATile) >> 4);
      const uint64_t bd = bd0 + (uint64_t)((bi * kBTile) >> 4);
#pragma unroll
      for (int t = 0; t < 512 / N / 2; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_pair(tmem + t * N, ad + (uint64_t)((k * 32) >> 4), bd + (uint64_t)((k * 32) >> 4), idesc,
                         (it | k) ? 1u : 0u);
      }
      if (++ai == a_tiles) ai = 0;
      if (++bi == b_tiles) bi = 0;
    }
    umma_commit_pair(done);
    mbar_wait(done, 0);
    cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

template <int N>
static void run_pair(int a_tiles, int b_tiles, long long* dcyc) {
  const int iters = 4000;
  const int smem = a_tiles * kATile + b_tiles * (N / 2) * 128 + 64 + 1024;
  cudaFuncSetAttribute(rate_pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148;
  for (int rep = 0; rep < 2; ++rep) {
    rate_pair_kernel<N><<<grid, 128, smem>>>(dcyc, iters, a_tiles, b_tiles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(grid / 2);
  cudaMemcpy(h.data(), dcyc, (grid / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (auto v : h) sum += (double)v;
  const double per_group = 4.0 * (512 / N / 2);
  const double cyc = sum / (grid / 2) / iters / per_group;
  printf("CTA pair M=256 N=%3d K-major a_tiles=%d b_tiles=%d: %.1f cycles per MMA (ideal %d) -> %.1f %% of the "
         "tensor pipe\n", N, a_tiles, b_tiles, cyc, N / 2, 100.0 * (N / 2) / cyc);
}

template <int N, int MN>
static void run(int a_tiles, int b_tiles, int tma, const uint8_t* gsrc, long long* dcyc) {
  const int iters = 4000;
  const int smem = a_tiles * kATile + b_tiles * N * 128 + 16384 + 64 + 1024;
  if (smem > kMaxSmem) { printf("skip N=%d a=%d b=%d (smem)\n", N, a_tiles, b_tiles); return; }
  cudaFuncSetAttribute(rate_kernel<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148;
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<N, MN><<<grid, 128, smem>>>(dcyc, iters, a_tiles, b_tiles, tma, gsrc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0, mx = 0;
  for (auto v : h) { sum += (double)v; if ((double)v > mx) mx = (double)v; }
  const double per_group = 4.0 * (512 / N / 2);
  const double cyc = sum / grid / iters / per_group;
  printf("N=%3d %s a_tiles=%d b_tiles=%d tma=%d: %.1f cycles per MMA (ideal %d) -> %.1f %% of the tensor pipe; "
         "slowest SM %.1f\n", N, MN ? "MN-major" : "K-major ", a_tiles, b_tiles, tma, cyc, N / 2,
         100.0 * (N / 2) / cyc, mx / iters / per_group);
}

int main() {
  long long* dcyc;
  uint8_t* gsrc;
  cudaMalloc(&dcyc, 148 * sizeof(long long));
  cudaMalloc(&gsrc, (size_t)148 << 20);
  cudaMemset(gsrc, 0, (size_t)148 << 20);
  for (int tma = 0; tma < 2; ++tma) {
    run<64, 0>(1, 1, tma, gsrc, dcyc);
    run<64, 0>(4, 4, tma, gsrc, dcyc);
    run<128, 0>(1, 1, tma, gsrc, dcyc);
    run<128, 0>(4, 4, tma, gsrc, dcyc);
    run<256, 0>(1, 1, tma, gsrc, dcyc);
    run<256, 0>(4, 3, tma, gsrc, dcyc);
    run<64, 1>(4, 4, tma, gsrc, dcyc);
    run<128, 1>(4, 4, tma, gsrc, dcyc);
    run<256, 1>(4, 3, tma, gsrc, dcyc);
  }
  run_pair<64>(4, 4, dcyc);
  run_pair<128>(4, 4, dcyc);
  run_pair<256>(4, 3, dcyc);
  printf("RESULT done\n");
  return 0;
}
