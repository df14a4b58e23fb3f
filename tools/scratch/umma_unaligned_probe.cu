// Probe: does a SWIZZLE_128B K-major UMMA operand work when its start address is only 128-byte
// aligned and its 8-row groups are 1280 bytes apart?
//
// Why: the 3x3 convolution kernels load THREE column-shifted copies of every activation tile so that
// each operand start stays 1024-byte aligned.  If the tensor core applies the 128-byte
// swizzle to the absolute shared-memory address (as TMA does when it writes), ONE box that is 10
// pixels wide serves all nine taps: tap (r, s) starts at r * 1280 + s * 128 bytes, row pitch 1280.
//
// Test: TMA loads X[18 rows][10 px][64 ch] (bf16, every element unique up to its 16-byte chunk) as one
// swizzled box; B is a 64 x 64 identity; for the nine (r, s) one M128 N64 K64 MMA group is issued with
// A start = base + r * 1280 + s * 128, SBO = 1280, and D (= the A operand as the tensor core saw it)
// is compared with X[r + m / 8][s + m % 8][:].  mode 0: base_offset field 0; mode 1: base_offset =
// (start >> 7) & 7.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o tools/scratch/umma_unaligned_probe \
//        tools/scratch/umma_unaligned_probe.cu weather-unet_b200/csrc/wu_host.cu -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../weather-unet_b200/csrc/wu_host.h"
#include "../../weather-unet_b200/csrc/wu_ptx.cuh"
using namespace wu;

constexpr int kRows = 18, kPx = 10, kCh = 64;
constexpr int kABytes = kRows * kPx * kCh * 2;  // 23040
constexpr int kAAlloc = 24576;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmX, float* out, int mode, int row_pitch_bytes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t raw = smem_u32(sm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* s = sm + (base - raw);
  const uint32_t a_base = base, b_base = base + kAAlloc;
  const uint32_t full = b_base + 8192, done = full + 8, slot = full + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(slot);
  // B = 64 x 64 identity, K-major, 128-byte swizzle: row n at n * 128, chunk (k / 8) ^ (n & 7)
  if (tid < 64) {
    for (int ck = 0; ck < 8; ++ck) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ck == (tid >> 3)) {
        uint16_t e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e[tid & 7] = 0x3F80;  // bf16 1.0
        v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      }
      *reinterpret_cast<uint4*>(s + kAAlloc + tid * 128 + ((ck ^ (tid & 7)) << 4)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(s + kAAlloc + 8192 + 16);
  if (tid == 0) {
    mbar_arrive_expect_tx(full, kABytes);
    tma_load_4d(a_base, &tmX, full, 0, 0, 0, 0);
    mbar_wait(full, 0);
  }
  __syncthreads();
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  for (int cs = 0; cs < 9; ++cs) {
    const int r = cs / 3, sft = cs % 3;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t start = a_base + r * row_pitch_bytes + sft * 128;
      uint64_t adesc = umma_smem_desc_sw128(start, 16, row_pitch_bytes);
      if (mode == 1) adesc |= (uint64_t)((start >> 7) & 7u) << 49;
      const uint64_t bdesc = umma_smem_desc_sw128(b_base, 16, 1024);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, adesc + (uint64_t)((k * 32) >> 4), bdesc + (uint64_t)((k * 32) >> 4), idesc,
                  k > 0 ? 1u : 0u);
      umma_commit(done);
      mbar_wait(done, cs & 1);
    }
    __syncthreads();
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_ld_32x32(taddr, v0);
    tmem_ld_32x32(taddr + 32, v1);
    tmem_ld_wait();
    float* o = out + ((size_t)cs * 128 + tid) * 64;
    for (int j = 0; j < 32; ++j) {
      o[j] = __uint_as_float(v0[j]);
      o[32 + j] = __uint_as_float(v1[j]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<64>(tmem);
}

// Second question (weight-gradient kernels, both operands MN-major: the contiguous 128 bytes are the
// 64 channels of one pixel, GEMM-K is the pixel): D[m][n] = sum_px A[px][m] * B[px][n] with
// A = [tap (r, s) | tap (r, s + 1)] (two 64-channel atoms ONE pixel = 128 bytes apart: LBO = 128),
// 8-pixel K groups one image row = 1280 bytes apart (SBO = 1280), start = base + r * 1280 + s * 128,
// and B = the 64 x 64 identity over (pixel, column).  Then D[m][n] = X[r + n / 8][s + (m >= 64) + n % 8][m % 64].
__global__ void __launch_bounds__(128, 1)
probe_mn(const __grid_constant__ CUtensorMap tmX, float* out, int row_pitch_bytes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t raw = smem_u32(sm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* s = sm + (base - raw);
  const uint32_t a_base = base, b_base = base + kAAlloc;
  const uint32_t full = b_base + 8192, done = full + 8, slot = full + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(slot);
  // B[k = pixel][n]: row k at k * 128 (8-pixel groups 1 KiB apart), chunk (n / 8) ^ (k & 7); identity
  if (tid < 64) {
    for (int ck = 0; ck < 8; ++ck) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ck == (tid >> 3)) {
        uint16_t e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e[tid & 7] = 0x3F80;
        v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      }
      *reinterpret_cast<uint4*>(s + kAAlloc + tid * 128 + ((ck ^ (tid & 7)) << 4)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(s + kAAlloc + 8192 + 16);
  if (tid == 0) {
    mbar_arrive_expect_tx(full, kABytes);
    tma_load_4d(a_base, &tmX, full, 0, 0, 0, 0);
    mbar_wait(full, 0);
  }
  __syncthreads();
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);  // both MN-major
  for (int cs = 0; cs < 6; ++cs) {  // r = 0..2, s = 0..1 (the second atom is column s + 1 <= 2)
    const int r = cs / 2, sft = cs % 2;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t start = a_base + r * row_pitch_bytes + sft * 128;
      const uint64_t adesc = umma_smem_desc_sw128(start, 128, row_pitch_bytes);
      const uint64_t bdesc = umma_smem_desc_sw128(b_base, 8192, 1024);
      for (int k = 0; k < 4; ++k)  // K = 16 pixels = two image rows
        umma_bf16(tmem, adesc + (uint64_t)((k * 2 * row_pitch_bytes) >> 4),
                  bdesc + (uint64_t)((k * 2048) >> 4), idesc, k > 0 ? 1u : 0u);
      umma_commit(done);
      mbar_wait(done, cs & 1);
    }
    __syncthreads();
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_ld_32x32(taddr, v0);
    tmem_ld_32x32(taddr + 32, v1);
    tmem_ld_wait();
    float* o = out + ((size_t)cs * 128 + tid) * 64;
    for (int j = 0; j < 32; ++j) {
      o[j] = __uint_as_float(v0[j]);
      o[32 + j] = __uint_as_float(v1[j]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<64>(tmem);
}

static float xval(int row, int px, int ch) {
  // (row, px) in the mantissa (1..180 < 256: exact in bf16), the 16-byte chunk in the exponent, the
  // parity of the channel in the sign
  const float m = (float)(row * kPx + px + 1);
  return ((ch & 1) ? -1.f : 1.f) * std::ldexp(m, (ch >> 3) - 4);
}

int main() {
  std::vector<__nv_bfloat16> h((size_t)kRows * kPx * kCh);
  for (int r = 0; r < kRows; ++r)
    for (int p = 0; p < kPx; ++p)
      for (int c = 0; c < kCh; ++c) h[((size_t)r * kPx + p) * kCh + c] = __float2bfloat16(xval(r, p, c));
  __nv_bfloat16* dx;
  float* dout;
  cudaMalloc(&dx, h.size() * 2);
  cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 9 * 128 * 64 * 4);
  CUtensorMap tm;
  int rc = make_act_tmap(&tm, dx, 1, kRows, kPx, kCh, kCh, kPx, kRows);
  printf("tensor map (64 ch, %d px, %d rows): rc=%d (%s)\n", kPx, kRows, rc, rc ? wu_last_error() : "ok");
  if (rc) return 1;
  const int smem = kAAlloc + 8192 + 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int all_ok = 1;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dout, 0xFF, 9 * 128 * 64 * 4);
    probe<<<1, 128, smem>>>(tm, dout, mode, kPx * 128);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d (base_offset %s): run %s\n", mode, mode ? "= (start >> 7) & 7" : "= 0",
           cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(9 * 128 * 64);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int cs = 0; cs < 9; ++cs) {
      const int r = cs / 3, s = cs % 3;
      int bad = 0, first_bad = -1;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 64; ++c) {
          const float want = xval(r + m / 8, s + m % 8, c);
          if (o[((size_t)cs * 128 + m) * 64 + c] != want) {
            if (first_bad < 0) first_bad = m;
            ++bad;
          }
        }
      printf("  tap r=%d s=%d: %s (%d of 8192 wrong)", r, s, bad ? "MISMATCH" : "ok", bad);
      if (bad) {
        all_ok = 0;
        // decode where rows first_bad .. +2 came from: chunk 4 (channels 32..39) has exponent 0
        printf("  first wrong row m=%d; D[m][32] decodes to (row*10+px+1) =", first_bad);
        for (int m = first_bad; m < first_bad + 3 && m < 128; ++m)
          printf(" %g (want %d)", o[((size_t)cs * 128 + m) * 64 + 32], (r + m / 8) * kPx + s + m % 8 + 1);
        printf("; chunks of row m:");
        for (int ck = 0; ck < 8; ++ck)
          printf(" %g", std::ldexp(o[((size_t)cs * 128 + first_bad) * 64 + ck * 8], 4 - ck));
      }
      printf("\n");
    }
  }
  printf(all_ok ? "RESULT (K-major): unaligned starts with a 1280-byte row pitch read correctly in both modes\n"
                : "RESULT (K-major): see mismatches above\n");
  // ---- MN-major operands (weight gradient)
  cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(dout, 0xFF, 9 * 128 * 64 * 4);
  probe_mn<<<1, 128, smem>>>(tm, dout, kPx * 128);
  cudaError_t e = cudaDeviceSynchronize();
  printf("MN-major, A = [tap (r,s) | tap (r,s+1)] (LBO 128, SBO 1280): run %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> o(6 * 128 * 64);
  cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
  int mn_ok = 1;
  for (int cs = 0; cs < 6; ++cs) {
    const int r = cs / 2, s = cs % 2;
    int bad = 0, fm = -1, fn = -1;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        const float want = xval(r + n / 8, s + (m >= 64 ? 1 : 0) + n % 8, m % 64);
        if (o[((size_t)cs * 128 + m) * 64 + n] != want) {
          if (fm < 0) { fm = m; fn = n; }
          ++bad;
        }
      }
    printf("  tap r=%d s=%d|%d: %s (%d of 8192 wrong)", r, s, s + 1, bad ? "MISMATCH" : "ok", bad);
    if (bad) {
      mn_ok = 0;
      printf("  first wrong D[%d][%d] = %g, want %g", fm, fn, o[((size_t)cs * 128 + fm) * 64 + fn],
             xval(r + fn / 8, s + (fm >= 64 ? 1 : 0) + fn % 8, fm % 64));
    }
    printf("\n");
  }
  printf(mn_ok ? "RESULT (MN-major): one 10-pixel box serves neighbouring column shifts of the weight gradient\n"
               : "RESULT (MN-major): see mismatches above\n");
  return 0;
}
