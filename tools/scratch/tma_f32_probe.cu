// Probe: fp32 NCHW image box loads through TMA (no swizzle), negative start coordinates.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../weather-unet_b200/csrc/wu_host.h"
#include "../../weather-unet_b200/csrc/wu_ptx.cuh"
using namespace wu;
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
  uint8_t* s = sm + (base - smem_u32(sm));
  const uint32_t bar = base + 16384;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, n * 4);
    tma_load_4d(base, &tm, bar, c0, c1, c2, c3);
  }
  // probe: do not rely on the transaction count; give the copy time, then look at shared memory
  for (int k = 0; k < 2000; ++k) __nanosleep(1000);
  if (threadIdx.x == 0) printf("barrier complete: %d\n", (int)mbar_try_wait(bar, 0));
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(s)[i];
}
int main(int argc, char** argv) {
  const int bx = atoi(argv[1]), by = atoi(argv[2]), x0 = atoi(argv[3]), y0 = atoi(argv[4]);
  const int bc = 3;
  const int B = 2, C = 3, H = 32, W = 32;
  std::vector<float> h(B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  int rc = make_image_tmap(&tm, d, B, C, H, W, bx, by);
  printf("box %dx%d at (%d,%d): encode rc=%d (%s)\n", bx, by, x0, y0, rc, rc ? wu_last_error() : "ok");
  if (rc) return 1;
  const int n = bx * by * bc;
  cudaMalloc(&o, n * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(tm, o, n, x0, y0, 0, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  run: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> r(n);
  cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bc; ++c)
    for (int y = 0; y < by; ++y)
      for (int x = 0; x < bx; ++x) {
        const int gx = x0 + x, gy = y0 + y;
        float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((1 * C + c) * H + gy) * W + gx] : 0.f;
        if (r[(c * by + y) * bx + x] != want) ++bad;
      }
  printf("  mismatches: %d of %d\n", bad, n);
  return 0;
}
