// wu_conv3x3.cu — 3x3 / stride 1 / pad 1 convolution of NHWC bf16 activations as tcgen05
// implicit GEMMs (fprop == dgrad kernel, wgrad kernel), plus the weight pack / reduce helpers.
//
// Replaces nn.Conv2d(cin, cout, 3, padding=1) [+ nn.ReLU] of the reference's r_double_conv
// (nets.py:18-24) and its autograd, including the torch.cat skip concatenation in front of the
// decoder blocks (cunet.py:62,69,76) by walking two sources in the K loop.
//
// fprop/dgrad tile (one CTA, persistent over tiles):
//   M = 128 output pixels (a bw x bh patch of one image), N = BN output channels, K = 9 taps x Cin.
//   A (activations): one 4-D TMA box (64 ch, bw, bh, 1) per (tap, channel block), fetched at the
//     tap-shifted coordinate; TMA zero-fills the halo, so padding costs nothing.  Lands as a
//     K-major [128 px][64 ch] 128B-swizzled tile — exactly the UMMA canonical layout.
//   B (weights): 2-D TMA box (64 k, BN rows) of the packed [Cout][9*Cin] matrix, K-major.
//   D: fp32 in TMEM, two BN-column buffers so the epilogue of tile i overlaps the MMAs of i+1.
//   Epilogue: tcgen05.ld -> bias/ReLU/mask -> bf16 -> swizzled smem staging -> TMA store (clips).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include <cstdio>
#include <cstdlib>

#include "wu_host.h"
#include "wu_ptx.cuh"

namespace wu {

// Debug build only (-DWU_PIPE_STATS, tools/pipe_stats.sh): cycles the single MMA-issuing thread and
// the TMA producer thread spend waiting on each barrier class, summed over all CTAs.  Not compiled
// into the product library.
#ifdef WU_PIPE_STATS
__device__ unsigned long long g_pipe_stats[16];
#define WU_STAT_DECL(n) long long _st[n] = {}
#define WU_STAT_WAIT(slot, ...)        \
  do {                                  \
    const long long _t0 = clock64();    \
    __VA_ARGS__;                        \
    _st[slot] += clock64() - _t0;       \
  } while (0)
#define WU_STAT_FLUSH(base, n) \
  for (int _i = 0; _i < (n); ++_i) atomicAdd(&g_pipe_stats[(base) + _i], (unsigned long long)_st[_i])
#else
#define WU_STAT_DECL(n)
#define WU_STAT_WAIT(slot, ...) __VA_ARGS__
#define WU_STAT_FLUSH(base, n)
#endif


// ------------------------------------------------------------------------------------------------
// fprop / dgrad
// ------------------------------------------------------------------------------------------------
struct ConvParams {
  int c0_blocks;    // 64-channel blocks of source 0
  int ctot_blocks;  // 64-channel blocks of source 0 + source 1
  int tiles_w, tiles_h, batch;
  int bw, bh, log2_bw;
  int n_tiles;    // cout / BN
  int num_tiles;  // batch * tiles_h * tiles_w * n_tiles
  int H, W, cout;
  int relu;
  int b1_mul;                 // 1, or 0 when source 1 has batch 1 and is shared by every image
  const float* bias;          // [cout] or null
  const __nv_bfloat16* mask;  // NHWC [B][H][W][cout] or null: dst = mask > 0 ? dst : 0
  // STATS instantiation: per-(image, 64-pixel half tile, channel) sums of the stored bf16 output
  // and of its square, [B][stats_chunks][cout][2] fp32 — AdaIN's statistics (utils.py:34-39) without
  // another pass over the tensor
  float* stats;
  int stats_chunks;
};

// Sum and sum of squares of one channel over 64 pixel rows of a staged [128 px][64 ch] bf16 tile
// (128-byte swizzle: 16-byte chunk k of row r sits at chunk position k ^ (r & 7)).  Thread et owns
// channel et & 63 and pixel rows (et >> 6) * 64 ... + 63; `valid(r)` excludes rows outside the image.
template <typename Valid>
__device__ __forceinline__ void staged_tile_stats(const uint8_t* stile, int et, Valid valid, float& s1,
                                                  float& s2) {
  const int c = et & 63, r0 = (et >> 6) * 64;
  const int kc = c >> 3, off = (c & 7) * 2;
  s1 = 0.f;
  s2 = 0.f;
#pragma unroll 8
  for (int i = 0; i < 64; ++i) {
    const int r = r0 + i;
    const uint16_t raw = *reinterpret_cast<const uint16_t*>(stile + r * 128 + ((kc ^ (r & 7)) << 4) + off);
    const float v = valid(r) ? __uint_as_float((uint32_t)raw << 16) : 0.f;
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
}

template <int BN>
struct ConvCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = 128 * 128;  // 128 pixels x 64 ch x 2 B
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 2 * 16384;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
  static constexpr uint32_t kTmemCols = 2 * BN;  // 128 / 256 / 512: powers of two
};

struct TileCoord {
  int b, h0, w0, n0;
};
template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  t.b = mt / p.tiles_h;
  t.h0 = th * p.bh;
  t.w0 = tw * p.bw;
  t.n0 = nt * BN;
  return t;
}

// One 32-column half of an epilogue chunk: fp32 accumulators -> (+bias, ReLU) -> packed bf16 pairs
// -> optional ReLU-gradient mask taken from `maskp` (4 x 16 B of bf16, keep where value > 0).
__device__ __forceinline__ void epilogue_half(uint32_t (&v)[32], const float* bias, int relu,
                                              const uint4* maskp, uint32_t (&pk)[16]) {
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j));
      v[j + 0] = __float_as_uint(__uint_as_float(v[j + 0]) + bv.x);
      v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + bv.y);
      v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + bv.z);
      v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + bv.w);
    }
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  if (maskp != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 m = __ldg(maskp + j);
      const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t lo = mm[e] & 0xFFFFu, hi = mm[e] >> 16;
        uint32_t keep = 0;
        if (lo != 0 && (lo & 0x8000u) == 0) keep |= 0x0000FFFFu;  // bf16 value > 0
        if (hi != 0 && (hi & 0x8000u) == 0) keep |= 0xFFFF0000u;
        pk[4 * j + e] &= keep;
      }
    }
  }
}

// Same, with the ReLU-gradient mask already in registers (prefetched while the MMAs were running).
__device__ __forceinline__ void epilogue_half_r(uint32_t (&v)[32], const float* bias, int relu,
                                                bool use_mask, const uint4* mreg, uint32_t (&pk)[16]) {
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j));
      v[j + 0] = __float_as_uint(__uint_as_float(v[j + 0]) + bv.x);
      v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + bv.y);
      v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + bv.z);
      v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + bv.w);
    }
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  if (use_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // bf16 value > 0  <=>  its bit pattern > 0 as a signed 16-bit integer: one SIMD compare
      // per element pair (keeps the epilogue warps' issue slots free for the MMA thread)
      pk[4 * j + 0] &= __vcmpgts2(mreg[j].x, 0u);
      pk[4 * j + 1] &= __vcmpgts2(mreg[j].y, 0u);
      pk[4 * j + 2] &= __vcmpgts2(mreg[j].z, 0u);
      pk[4 * j + 3] &= __vcmpgts2(mreg[j].w, 0u);
    }
  }
}

template <int BN, bool STATS>
__global__ void __launch_bounds__(192, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmA0,
                     const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmD, const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);

  const uint32_t staging_base = base + S * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kblocks = 9 * p.ctot_blocks;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      WU_STAT_DECL(2);
#ifdef WU_PIPE_STATS
      const long long _p0 = clock64();
#endif
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<BN>(p, tile);
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          for (int cb = 0; cb < p.ctot_blocks; ++cb) {
            WU_STAT_WAIT(1, mbar_wait(empty_bar(stage), phase ^ 1u));
            const uint32_t fb = full_bar(stage);
            mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
            const uint32_t a_dst = base + stage * Cfg::kStageBytes;
            const uint32_t b_dst = a_dst + Cfg::kABytes;
            if (cb < p.c0_blocks)
              tma_load_4d(a_dst, &tmA0, fb, cb * 64, t.w0 + s - 1, t.h0 + r - 1, t.b);
            else
              tma_load_4d(a_dst, &tmA1, fb, (cb - p.c0_blocks) * 64, t.w0 + s - 1, t.h0 + r - 1,
                          t.b * p.b1_mul);
            tma_load_2d(b_dst, &tmB, fb, (tap * p.ctot_blocks + cb) * 64, t.n0);
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _p0;
#endif
      WU_STAT_FLUSH(4, 2);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      WU_STAT_DECL(3);
#ifdef WU_PIPE_STATS
      const long long _m0 = clock64();
#endif
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        WU_STAT_WAIT(1, mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1u));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          WU_STAT_WAIT(2, mbar_wait(full_bar(stage), phase));
          tc_fence_after();
          const uint32_t a_addr = base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16) per 64-channel block
            const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(buf));  // accumulator complete -> epilogue
      }
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _m0;
#endif
      WU_STAT_FLUSH(0, 3);
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // pixel within the tile == TMEM lane
    const bool issuer = (threadIdx.x == 64);
    const int ph = row >> p.log2_bw;
    const int pw = row & (p.bw - 1);
    uint32_t store_count = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const TileCoord t = decode_tile<BN>(p, tile);
      const int h = t.h0 + ph, w = t.w0 + pw;
      const bool inb = (h < p.H) && (w < p.W);
      mbar_wait(tfull_bar(buf), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + chunk * 64;
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {  // this thread has drained its part of the accumulator
          tc_fence_before();
          mbar_arrive(tempty_bar(buf));
        }
        const int cbase = t.n0 + chunk * 64;
        const float* bptr = p.bias != nullptr ? p.bias + cbase : nullptr;
        const uint4* mptr = nullptr;
        if (p.mask != nullptr && inb)
          mptr = reinterpret_cast<const uint4*>(
              p.mask + ((size_t)(t.b * p.H + h) * p.W + w) * p.cout + cbase);
        uint32_t pk0[16], pk1[16];
        epilogue_half(v0, bptr, p.relu, mptr, pk0);
        epilogue_half(v1, bptr ? bptr + 32 : nullptr, p.relu, mptr ? mptr + 4 : nullptr, pk1);
        // stage the [128 px][64 ch] bf16 tile in the 128B-swizzled layout the store map expects
        const uint32_t sb = staging_base + (store_count & 1u) * 16384u;
        ++store_count;
        if (issuer) tma_store_wait_read<1>();  // the store that last used this buffer has read it
        named_bar_sync(1, 128);
        uint8_t* srow = smem + (sb - base) + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(srow + ((j ^ (row & 7)) << 4)) =
              make_uint4(pk0[4 * j], pk0[4 * j + 1], pk0[4 * j + 2], pk0[4 * j + 3]);
          *reinterpret_cast<uint4*>(srow + (((j + 4) ^ (row & 7)) << 4)) =
              make_uint4(pk1[4 * j], pk1[4 * j + 1], pk1[4 * j + 2], pk1[4 * j + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          tma_store_4d(&tmD, sb, cbase, t.w0, t.h0, t.b);
          tma_store_commit();
        }
        if (STATS) {
          // the staged tile is rewritten only after the next-but-one chunk's barrier, which every
          // thread reaches after these reads
          const int et = threadIdx.x - 64;
          float s1, s2;
          staged_tile_stats(smem + (sb - base), et, [&](int r) {
            return (t.h0 + (r >> p.log2_bw) < p.H) && (t.w0 + (r & (p.bw - 1)) < p.W); }, s1, s2);
          const int mt = tile / p.n_tiles;  // (b, th, tw) linear
          const int in_img = mt - t.b * p.tiles_h * p.tiles_w;
          float* o = p.stats + (((size_t)t.b * p.stats_chunks + in_img * 2 + (et >> 6)) * p.cout + cbase +
                                (et & 63)) * 2;
          *reinterpret_cast<float2*>(o) = make_float2(s1, s2);
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN, bool STATS = false>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& bm,
                       const CUtensorMap& dm, const ConvParams& p, cudaStream_t st) {
  using Cfg = ConvCfg<BN>;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_igemm_kernel<BN, STATS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  conv3x3_igemm_kernel<BN, STATS><<<grid, 192, Cfg::kSmemBytes, st>>>(a0, a1, bm, dm, p);
  WU_CHECK_LAUNCH("conv3x3_igemm_kernel");
  return WU_OK;
}

// ------------------------------------------------------------------------------------------------
// N = 256 on CTA PAIRS (tcgen05 cta_group::2, M = 256): conv3x3_igemm_pair_kernel
// ------------------------------------------------------------------------------------------------
// Why (profiles/r02_pipeline_experiments.txt): the single-CTA N = 256 kernel above moves 48 KiB per
// 512 MMA cycles from L2 into every SM (a 16 KiB pixel tile + the whole 32 KiB weight tile), and its
// MMA-issuing thread waits 15-26 % of the time for those bytes.  In a pair each CTA keeps its own 128
// pixels but only HALF of the weight tile (128 of the 256 rows of B): 32 KiB per 512 cycles and SM,
// stages of 32 KiB instead of 48 (six instead of four in flight), 64 instead of 96 B/clk of operand
// reads.  Same tiling, same accumulation order per output element as the single-CTA kernel, hence
// bit-identical results.
// Pair protocol (validated in round 1 on the N = 64 kernel, where it did not pay because that shape is
// bound by the operand read rate, not by delivery):
//   * cluster (2,1,1); pair-tile = two M-tiles (128 pixels each) x one N-tile; CTA r owns M-tile 2m + r;
//   * "full" barriers live in the leader (CTA 0): it posts expect_tx for the bytes of BOTH CTAs and both
//     producers' cp.async.bulk.tensor ... .cta_group::2 loads complete on it;
//   * "empty" and "accumulator full" barriers exist in both CTAs, signalled by the leader's
//     tcgen05.commit.cta_group::2 ... multicast;
//   * "accumulator drained" lives in the leader (256 arrivals: both CTAs' epilogue threads).
struct ConvPairCfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = 128 * 128;  // 128 pixels x 64 ch x 2 B
  static constexpr int kBBytes = 128 * 128;  // this CTA's 128 of the 256 weight rows x 64 k x 2 B
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 2 * 16384;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 1024;
  static constexpr uint32_t kTmemCols = 512;  // two accumulator buffers of 256 columns
};

template <bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv3x3_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA0,
                          const __grid_constant__ CUtensorMap tmA1,
                          const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmD, const ConvParams p) {
  using Cfg = ConvPairCfg;
  constexpr int S = Cfg::kStages, BN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t staging_base = base + S * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = p.batch * p.tiles_h * p.tiles_w;
  const int num_pairs_total = ((m_tiles + 1) >> 1) * p.n_tiles;  // the last M pair may have one live half

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 256);  // both CTAs' epilogue threads (used in the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int kblocks = 9 * p.ctot_blocks;

  // M-tile of this CTA inside pair-tile `pt`; the dead half of an odd tail re-does the last tile
  auto decode = [&](int pt, int& mtile, int& b, int& h0, int& w0, int& n0, bool& live) {
    const int nt = pt % p.n_tiles;
    const int mp = pt / p.n_tiles;
    mtile = 2 * mp + (int)rank;
    live = mtile < m_tiles;
    if (!live) mtile = m_tiles - 1;
    const int tw = mtile % p.tiles_w;
    const int mt = mtile / p.tiles_w;
    const int th = mt % p.tiles_h;
    b = mt / p.tiles_h;
    h0 = th * p.bh;
    w0 = tw * p.bw;
    n0 = nt * BN;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      // One thread feeds the CTA: keep its per-stage instruction count minimal (see the weight-gradient
      // kernel): the leader's barrier window is mapped once, the weight-tile K coordinate and the
      // channel-block coordinate advance by increments, no division inside the stage loop.
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader = cluster_map(full_bar(0), 0);  // + 8 * stage
      const int brow = 128 * (int)rank;
      for (int pt = pair; pt < num_pairs_total; pt += npairs) {
        int mtile, b, h0, w0, n0;
        bool live;
        decode(pt, mtile, b, h0, w0, n0, live);
        const int b1 = b * p.b1_mul;
        const int nrow = n0 + brow;
        int kcoord = 0;  // (tap * ctot_blocks + cb) * 64
        for (int r = 0; r < 3; ++r) {
          for (int s = 0; s < 3; ++s) {
            const int wc = w0 + s - 1, hc = h0 + r - 1;
            for (int cb = 0; cb < p.ctot_blocks; ++cb, kcoord += 64) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
              const uint32_t fb = full_leader + 8u * stage;
              const uint32_t a_dst = base + stage * Cfg::kStageBytes;
              if (cb < p.c0_blocks) tma_load_4d_pair(a_dst, &tmA0, fb, cb * 64, wc, hc, b);
              else tma_load_4d_pair(a_dst, &tmA1, fb, (cb - p.c0_blocks) * 64, wc, hc, b1);
              tma_load_2d_pair(a_dst + Cfg::kABytes, &tmB, fb, kcoord, nrow);
              if (++stage == S) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ------------------------------------------------------------ MMA issuer (leader, one thread)
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = pair; pt < num_pairs_total; pt += npairs, ++it) {
        const int buf = it & 1;
        mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16) per 64-channel block
            const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(empty_bar(stage));  // frees the slot in BOTH CTAs once these MMAs retire
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(tfull_bar(buf));  // accumulator complete -> both epilogues
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5, both CTAs)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // pixel within the tile == TMEM lane
    const bool issuer = (threadIdx.x == 64);
    const int ph = row >> p.log2_bw;
    const int pw = row & (p.bw - 1);
    const uint32_t tempty_leader[2] = {cluster_map(tempty_bar(0), 0), cluster_map(tempty_bar(1), 0)};
    uint32_t store_count = 0;
    int it = 0;
    for (int pt = pair; pt < num_pairs_total; pt += npairs, ++it) {
      const int buf = it & 1;
      int mtile, b, h0, w0, n0;
      bool live;
      decode(pt, mtile, b, h0, w0, n0, live);
      const int h = h0 + ph, w = w0 + pw;
      const bool inb = live && (h < p.H) && (w < p.W);
      mbar_wait(tfull_bar(buf), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + chunk * 64;
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {  // this thread has drained its part of the accumulator
          tc_fence_before();
          mbar_arrive_cluster(tempty_leader[buf]);
        }
        const int cbase = n0 + chunk * 64;
        const float* bptr = p.bias != nullptr ? p.bias + cbase : nullptr;
        const uint4* mptr = nullptr;
        if (p.mask != nullptr && inb)
          mptr = reinterpret_cast<const uint4*>(
              p.mask + ((size_t)(b * p.H + h) * p.W + w) * p.cout + cbase);
        uint32_t pk0[16], pk1[16];
        epilogue_half(v0, bptr, p.relu, mptr, pk0);
        epilogue_half(v1, bptr ? bptr + 32 : nullptr, p.relu, mptr ? mptr + 4 : nullptr, pk1);
        const uint32_t sb = staging_base + (store_count & 1u) * 16384u;
        ++store_count;
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        uint8_t* srow = smem + (sb - base) + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(srow + ((j ^ (row & 7)) << 4)) =
              make_uint4(pk0[4 * j], pk0[4 * j + 1], pk0[4 * j + 2], pk0[4 * j + 3]);
          *reinterpret_cast<uint4*>(srow + (((j + 4) ^ (row & 7)) << 4)) =
              make_uint4(pk1[4 * j], pk1[4 * j + 1], pk1[4 * j + 2], pk1[4 * j + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          if (live) tma_store_4d(&tmD, sb, cbase, w0, h0, b);
          tma_store_commit();
        }
        if (STATS) {
          if (live) {
            const int et = threadIdx.x - 64;
            float s1, s2;
            staged_tile_stats(smem + (sb - base), et, [&](int r) {
              return (h0 + (r >> p.log2_bw) < p.H) && (w0 + (r & (p.bw - 1)) < p.W); }, s1, s2);
            const int in_img = mtile - b * p.tiles_h * p.tiles_w;
            float* o = p.stats + (((size_t)b * p.stats_chunks + in_img * 2 + (et >> 6)) * p.cout + cbase +
                                  (et & 63)) * 2;
            *reinterpret_cast<float2*>(o) = make_float2(s1, s2);
          }
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody frees tensor memory (or exits) while the pair still uses it
  if (warp == 2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
}

template <bool STATS>
static int launch_conv_pair(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& bm,
                            const CUtensorMap& dm, const ConvParams& p, cudaStream_t st) {
  using Cfg = ConvPairCfg;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_igemm_pair_kernel<STATS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int m_tiles = p.batch * p.tiles_h * p.tiles_w;
  const int pairs = ((m_tiles + 1) / 2) * p.n_tiles;
  const int max_pairs = num_sms() / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  conv3x3_igemm_pair_kernel<STATS><<<grid, 192, Cfg::kSmemBytes, st>>>(a0, a1, bm, dm, p);
  WU_CHECK_LAUNCH("conv3x3_igemm_pair_kernel");
  return WU_OK;
}

// ------------------------------------------------------------------------------------------------
// fprop / dgrad, v2: operand reuse in shared memory
// ------------------------------------------------------------------------------------------------
// v1 above fetches one (tap, channel block) A tile per MMA group, i.e. every input pixel crosses
// L2 -> SMEM nine times, and every weight tile once per 128 output pixels; the N = 64 / 128 layers
// are L2-bandwidth bound that way (ncu: tensor pipe 31 % / 57 %).  v2 changes the tiling:
//   * output super-tile = T vertically stacked M-tiles of 16 rows x 8 columns (T*128 pixels);
//   * A: per 64-channel block only THREE TMA boxes (64 ch, 8 px, 16T+2 rows), one per column shift
//     s-1 in {-1, 0, +1}.  A row shift r is a start-address offset of r KiB (8 pixels x 128 B), which
//     keeps every UMMA descriptor 1024-byte aligned, so the three taps (r, s) of a column shift and
//     all T M-tiles read the same copy: A traffic drops from 9 x 16 KiB to 3 x (16T+2) KiB per T tiles;
//   * B: one (tap, channel block) weight tile feeds T M-tiles (T accumulators live in TMEM);
//   * TMEM: 2 x T x BN = 512 columns, double buffered across super-tiles as before.
struct ConvParams2 {
  int c0_blocks, ctot_blocks;
  int tiles_w, tiles_h, batch;
  int n_tiles, num_tiles;
  int H, W, cout, relu;
  int b1_mul;  // 1, or 0 when source 1 has batch 1 and is shared by every image
  const float* bias;
  const __nv_bfloat16* mask;
  // LAST instantiation (dconv_up1.2): Conv2d(64, 3, 1) + Tanh (cunet.py:39-40,80-82) on the tile while
  // it is still in registers: last_w fp32 [3][64], last_b [3], last_y fp32 NCHW [B][3][H][W]
  const float* last_w;
  const float* last_b;
  float* last_y;
  // POOL instantiation (dconv_down1.2 / dconv_down2.2): nn.MaxPool2d(2) (cunet.py:46,49) of the tile,
  // taken from the staged bf16 tile before it leaves shared memory: pool_dst NHWC [B][H/2][W/2][cout]
  __nv_bfloat16* pool_dst;
  // STATS instantiation (dconv_up2.2): AdaIN sums of the stored output, see ConvParams::stats
  float* stats;
  int stats_chunks;
};

// (A variant that loads ONE 10-pixel-wide TMA box per channel block and reaches the three column
// shifts through 128-byte start offsets — UMMA operands may start at any 128-byte boundary under
// SWIZZLE_128B, profiles/r01_umma_unaligned_start_probe.txt — was validated and timed in round 2:
// bit-identical results, 2.4x fewer bytes across L2 -> SM, and no faster (fprop 73.0 vs 73.2 % of peak
// over the 13 layers, dgrad 69.1 vs 70.9 %): these kernels are paced by the tensor core's own
// shared-memory operand reads, not by the fabric.  Removed; numbers in profiles/r02_conv_impl_*.txt.)
template <int BN, int T>
struct ConvCfg2 {
  static constexpr int kARows = 16 * T + 2;
  static constexpr int kRowPitch = 1024;  // bytes per image row of an A stage (8 px x 128 B)
  static constexpr int kATx = kARows * kRowPitch;       // bytes one TMA box delivers
  static constexpr int kABytes = (kATx + 1023) / 1024 * 1024;  // stage stride (1024-byte aligned)
  static constexpr int kSA = (T == 4) ? 2 : 3;
  static constexpr int kSB = BN == 64 ? 5 : (BN == 128 ? 4 : 3);
  static constexpr int kBBytes = BN * 128;
  static constexpr int kNStg = 2;  // epilogue staging buffers
  static constexpr int kStagingBytes = kNStg * 16384;
  static constexpr int kMaskBytes = 16384;  // dgrad: ReLU-mask tile of the next chunk (cp.async)
  static constexpr int kSmemBytes =
      kSA * kABytes + kSB * kBBytes + kStagingBytes + kMaskBytes + 1024 + 1024;
  static constexpr uint32_t kTmemCols = 2 * T * BN;
  static_assert(kTmemCols == 512, "TMEM budget: 2 x T x BN must be 512 columns");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int BN, int T, bool LAST, bool POOL, bool STATS>
__global__ void __launch_bounds__(192, 1)
conv3x3_igemm_v2_kernel(const __grid_constant__ CUtensorMap tmA0,
                        const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmD, const ConvParams2 p) {
  using Cfg = ConvCfg2<BN, T>;
  constexpr int SA = Cfg::kSA, SB = Cfg::kSB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const uint32_t a_base = base;
  const uint32_t b_base = a_base + SA * Cfg::kABytes;
  const uint32_t staging_base = b_base + SB * Cfg::kBBytes;
  const uint32_t mask_base = staging_base + Cfg::kStagingBytes;
  const uint32_t bar_base = mask_base + Cfg::kMaskBytes;
  auto fullA = [&](int i) { return bar_base + 8u * i; };
  auto emptyA = [&](int i) { return bar_base + 8u * (SA + i); };
  auto fullB = [&](int i) { return bar_base + 8u * (2 * SA + i); };
  auto emptyB = [&](int i) { return bar_base + 8u * (2 * SA + SB + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * SA + 2 * SB + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + 2 * SB + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SA; ++i) {
      mbar_init(fullA(i), 1);
      mbar_init(emptyA(i), 1);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(fullB(i), 1);
      mbar_init(emptyB(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 128);
    }
    fence_mbar_init();
  }
  if (LAST) {  // the 1x1 weights live in the ReLU-mask staging area (unused in a forward pass)
    float* lw = reinterpret_cast<float*>(smem + (mask_base - base));
    for (int i = threadIdx.x; i < 3 * 64 + 3; i += blockDim.x)
      lw[i] = i < 192 ? p.last_w[i] : (p.last_b != nullptr ? p.last_b[i - 192] : 0.f);
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto decode = [&](int tile, int& b, int& h0, int& w0, int& n0) {
    const int nt = tile % p.n_tiles;
    int mt = tile / p.n_tiles;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    b = mt / p.tiles_h;
    h0 = th * 16 * T;
    w0 = tw * 8;
    n0 = nt * BN;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      WU_STAT_DECL(3);
#ifdef WU_PIPE_STATS
      const long long _p0 = clock64();
#endif
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b, h0, w0, n0;
        decode(tile, b, h0, w0, n0);
        for (int cb = 0; cb < p.ctot_blocks; ++cb) {
          for (int s = 0; s < 3; ++s) {
            {
              const int wl = w0 + s - 1;
              WU_STAT_WAIT(1, mbar_wait(emptyA(sa), pa ^ 1u));
              mbar_arrive_expect_tx(fullA(sa), Cfg::kATx);
              if (cb < p.c0_blocks)
                tma_load_4d(a_base + sa * Cfg::kABytes, &tmA0, fullA(sa), cb * 64, wl, h0 - 1, b);
              else
                tma_load_4d(a_base + sa * Cfg::kABytes, &tmA1, fullA(sa), (cb - p.c0_blocks) * 64, wl,
                            h0 - 1, b * p.b1_mul);
              if (++sa == SA) { sa = 0; pa ^= 1u; }
            }
            for (int r = 0; r < 3; ++r) {
              WU_STAT_WAIT(2, mbar_wait(emptyB(sb), pb ^ 1u));
              mbar_arrive_expect_tx(fullB(sb), Cfg::kBBytes);
              tma_load_2d(b_base + sb * Cfg::kBBytes, &tmB, fullB(sb),
                          ((r * 3 + s) * p.ctot_blocks + cb) * 64, n0);
              if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _p0;
#endif
      WU_STAT_FLUSH(4, 3);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      // descriptors are built once; per MMA only (byte offset >> 4) is added to the start-address
      // field (the single issuing thread's instruction count paces the tensor pipe at small N)
      constexpr int RP = Cfg::kRowPitch;  // bytes between the 8-pixel row groups of an A stage
      const uint64_t adesc0 = umma_smem_desc_sw128(a_base, 16, RP);
      const uint64_t bdesc0 = umma_smem_desc_sw128(b_base, 16, 1024);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      WU_STAT_DECL(4);
#ifdef WU_PIPE_STATS
      const long long _m0 = clock64();
#endif
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        WU_STAT_WAIT(1, mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1u));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (T * BN);
        for (int cb = 0; cb < p.ctot_blocks; ++cb) {
          for (int s = 0; s < 3; ++s) {
            WU_STAT_WAIT(2, mbar_wait(fullA(sa), pa));
            tc_fence_after();
            const uint64_t adesc_s = adesc0 + (uint64_t)((sa * Cfg::kABytes) >> 4);
            for (int r = 0; r < 3; ++r) {
              WU_STAT_WAIT(3, mbar_wait(fullB(sb), pb));
              tc_fence_after();
              const uint64_t bdesc_s = bdesc0 + (uint64_t)((sb * Cfg::kBBytes) >> 4);
              const uint64_t adesc_r = adesc_s + (uint64_t)((r * RP) >> 4);
              const uint32_t first = (cb | s | r) == 0 ? 1u : 0u;
#pragma unroll
              for (int t = 0; t < T; ++t) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t adesc = adesc_r + (uint64_t)((16 * t * RP + k * 32) >> 4);
                  const uint64_t bdesc = bdesc_s + (uint64_t)((k * 32) >> 4);
                  umma_bf16(d_tmem + t * BN, adesc, bdesc, idesc, (first && k == 0) ? 0u : 1u);
                }
              }
              umma_commit(emptyB(sb));
              if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
            umma_commit(emptyA(sa));
            if (++sa == SA) { sa = 0; pa ^= 1u; }
          }
        }
        umma_commit(tfull_bar(buf));
      }
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _m0;
#endif
      WU_STAT_FLUSH(0, 4);
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // pixel within an M-tile == TMEM lane
    const bool issuer = (threadIdx.x == 64);
    const int ph = row >> 3, pw = row & 7;
    constexpr int NCH = BN / 64;
    uint32_t store_count = 0;
    // dgrad: the ReLU-gradient mask tile (128 px x 64 ch of the layer below) of the NEXT chunk is
    // copied global -> shared with coalesced 16-byte cp.async while the current chunk is processed;
    // each thread then reads its own pixel row from shared memory (XOR-swizzled, conflict free).
    // (Per-thread row loads from global kept the LSU pipe ~90 % busy: 32 lines per request.)
    const bool has_mask = p.mask != nullptr;
    const int et = threadIdx.x - 64;  // 0..127 within the epilogue warps
    auto issue_mask = [&](int tile, int t, int chunk) {
      if (!has_mask || tile >= p.num_tiles) return;
      int b, h0, w0, n0;
      decode(tile, b, h0, w0, n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int u = et + 128 * j;      // 16-byte unit of the [128 px][64 ch] tile
        const int r = u >> 3, c = u & 7;  // pixel row of the tile, 8-channel chunk
        const int h = h0 + 16 * t + (r >> 3), w = w0 + (r & 7);
        const bool ok = (h < p.H) && (w < p.W);
        const __nv_bfloat16* src =
            p.mask + ((size_t)(b * p.H + (ok ? h : 0)) * p.W + (ok ? w : 0)) * p.cout + n0 +
            chunk * 64 + c * 8;
        cp_async16(mask_base + r * 128 + ((c ^ (r & 7)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    uint4 mk[8];
    issue_mask(blockIdx.x, 0, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int b, h0, w0, n0;
      decode(tile, b, h0, w0, n0);
      mbar_wait(tfull_bar(buf), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
#pragma unroll 1
        for (int chunk = 0; chunk < NCH; ++chunk) {
          if (has_mask) {
            cp_async_wait_all();
            named_bar_sync(3, 128);  // every thread's copies have landed
            const uint8_t* mrow = smem + (mask_base - base) + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              mk[j] = *reinterpret_cast<const uint4*>(mrow + ((j ^ (row & 7)) << 4));
            named_bar_sync(4, 128);  // everyone has read: the buffer may be refilled
            if (chunk + 1 < NCH) issue_mask(tile, t, chunk + 1);
            else if (t + 1 < T) issue_mask(tile, t + 1, 0);
            else issue_mask(tile + gridDim.x, 0, 0);
          }
          uint32_t v0[32], v1[32];
          const uint32_t taddr =
              tmem_base + ((uint32_t)(q * 32) << 16) + buf * (T * BN) + t * BN + chunk * 64;
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32, v1);
          tmem_ld_wait();
          if (t == T - 1 && chunk == NCH - 1) {
            tc_fence_before();
            mbar_arrive(tempty_bar(buf));
          }
          const int cbase = n0 + chunk * 64;
          const float* bptr = p.bias != nullptr ? p.bias + cbase : nullptr;
          uint32_t pk0[16], pk1[16];
          epilogue_half_r(v0, bptr, p.relu, has_mask, mk, pk0);
          epilogue_half_r(v1, bptr ? bptr + 32 : nullptr, p.relu, has_mask, mk + 4, pk1);
          if (LAST) {
            // y[b, o, h, w] = tanh(last_b[o] + sum_c last_w[o][c] * relu(conv)[c]): this thread holds
            // all 64 channels of its pixel (BN == 64), fp32, after bias + ReLU
            const float4* lw4 = reinterpret_cast<const float4*>(smem + (mask_base - base));
            const float* lb = reinterpret_cast<const float*>(smem + (mask_base - base)) + 192;
            float a0 = lb[0], a1 = lb[1], a2 = lb[2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w0v = lw4[j], w1v = lw4[16 + j], w2v = lw4[32 + j];
              const float x0 = __uint_as_float(v0[4 * j]), x1 = __uint_as_float(v0[4 * j + 1]);
              const float x2 = __uint_as_float(v0[4 * j + 2]), x3 = __uint_as_float(v0[4 * j + 3]);
              a0 = fmaf(x0, w0v.x, a0); a0 = fmaf(x1, w0v.y, a0); a0 = fmaf(x2, w0v.z, a0); a0 = fmaf(x3, w0v.w, a0);
              a1 = fmaf(x0, w1v.x, a1); a1 = fmaf(x1, w1v.y, a1); a1 = fmaf(x2, w1v.z, a1); a1 = fmaf(x3, w1v.w, a1);
              a2 = fmaf(x0, w2v.x, a2); a2 = fmaf(x1, w2v.y, a2); a2 = fmaf(x2, w2v.z, a2); a2 = fmaf(x3, w2v.w, a2);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w0v = lw4[8 + j], w1v = lw4[24 + j], w2v = lw4[40 + j];
              const float x0 = __uint_as_float(v1[4 * j]), x1 = __uint_as_float(v1[4 * j + 1]);
              const float x2 = __uint_as_float(v1[4 * j + 2]), x3 = __uint_as_float(v1[4 * j + 3]);
              a0 = fmaf(x0, w0v.x, a0); a0 = fmaf(x1, w0v.y, a0); a0 = fmaf(x2, w0v.z, a0); a0 = fmaf(x3, w0v.w, a0);
              a1 = fmaf(x0, w1v.x, a1); a1 = fmaf(x1, w1v.y, a1); a1 = fmaf(x2, w1v.z, a1); a1 = fmaf(x3, w1v.w, a1);
              a2 = fmaf(x0, w2v.x, a2); a2 = fmaf(x1, w2v.y, a2); a2 = fmaf(x2, w2v.z, a2); a2 = fmaf(x3, w2v.w, a2);
            }
            const int hh = h0 + 16 * t + ph, ww = w0 + pw;
            if (hh < p.H && ww < p.W) {
              const size_t hw = (size_t)p.H * p.W;
              float* yo = p.last_y + (size_t)b * 3 * hw + (size_t)hh * p.W + ww;
              yo[0] = tanhf(a0);
              yo[hw] = tanhf(a1);
              yo[2 * hw] = tanhf(a2);
            }
          }
          const uint32_t sbuf = staging_base + (store_count % Cfg::kNStg) * 16384u;
          ++store_count;
          if (issuer) tma_store_wait_read<Cfg::kNStg - 1>();
          named_bar_sync(1, 128);
          uint8_t* srow = smem + (sbuf - base) + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(srow + ((j ^ (row & 7)) << 4)) =
                make_uint4(pk0[4 * j], pk0[4 * j + 1], pk0[4 * j + 2], pk0[4 * j + 3]);
            *reinterpret_cast<uint4*>(srow + (((j + 4) ^ (row & 7)) << 4)) =
                make_uint4(pk1[4 * j], pk1[4 * j + 1], pk1[4 * j + 2], pk1[4 * j + 3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (issuer) {
            if (h0 + 16 * t < p.H) tma_store_4d(&tmD, sbuf, cbase, w0, h0 + 16 * t, b);
            tma_store_commit();
          }
          if (STATS) {
            float s1, s2;
            staged_tile_stats(smem + (sbuf - base), et, [&](int r) {
              return (h0 + 16 * t + (r >> 3) < p.H) && (w0 + (r & 7) < p.W); }, s1, s2);
            const int mt = tile / p.n_tiles;  // (b, th, tw) linear
            const int in_img = mt - b * p.tiles_h * p.tiles_w;
            float* o = p.stats + (((size_t)b * p.stats_chunks + (in_img * T + t) * 2 + (et >> 6)) * p.cout +
                                  cbase + (et & 63)) * 2;
            *reinterpret_cast<float2*>(o) = make_float2(s1, s2);
          }
          if (POOL) {
            // 2x2 max over the staged [16 x 8 px][64 ch] tile -> 8 x 4 pooled pixels; a task is one
            // 16-byte chunk of one pooled pixel (256 tasks, two per thread).  Values are post-ReLU
            // (non-negative bf16), so the unsigned 16-bit SIMD max is the float max.  The staging
            // buffer is only rewritten after the next chunk's barrier, which every thread reaches
            // after these reads.
            const uint8_t* stile = smem + (sbuf - base);
            const int Hp = p.H >> 1, Wp = p.W >> 1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int task = et + 128 * half;
              const int pp = task >> 3, c = task & 7;
              const int pr = pp >> 2, pc = pp & 3;
              const int r00 = (2 * pr) * 8 + 2 * pc;
              uint4 m = *reinterpret_cast<const uint4*>(stile + r00 * 128 + ((c ^ (r00 & 7)) << 4));
#pragma unroll
              for (int k = 1; k < 4; ++k) {
                const int r = r00 + (k & 1) + (k >> 1) * 8;
                const uint4 q = *reinterpret_cast<const uint4*>(stile + r * 128 + ((c ^ (r & 7)) << 4));
                m.x = __vmaxu2(m.x, q.x); m.y = __vmaxu2(m.y, q.y);
                m.z = __vmaxu2(m.z, q.z); m.w = __vmaxu2(m.w, q.w);
              }
              const int hp = ((h0 + 16 * t) >> 1) + pr, wp = (w0 >> 1) + pc;
              if (hp < Hp && wp < Wp)
                *reinterpret_cast<uint4*>(p.pool_dst + (((size_t)b * Hp + hp) * Wp + wp) * p.cout + cbase +
                                          c * 8) = m;
            }
          }
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN, int T, bool LAST = false, bool POOL = false, bool STATS = false>
static int launch_conv2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& bm,
                        const CUtensorMap& dm, const ConvParams2& p, cudaStream_t st) {
  using Cfg = ConvCfg2<BN, T>;
  static bool attr_done = false;
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_igemm_v2_kernel<BN, T, LAST, POOL, STATS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  conv3x3_igemm_v2_kernel<BN, T, LAST, POOL, STATS>
      <<<grid, 192, Cfg::kSmemBytes, st>>>(a0, a1, bm, dm, p);
  WU_CHECK_LAUNCH("conv3x3_igemm_v2_kernel");
  return WU_OK;
}

// WU_CONV_IMPL: 0 / unset = per-shape choice, 1 = v1 everywhere, 2 = v2 everywhere (read once; A/B
// measurements only)
static int conv_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WU_CONV_IMPL");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
  }
  return v;
}

// WU_CONV_PAIR=0 keeps the single-CTA N = 256 kernel (A/B measurements); default: CTA pairs
#ifndef WU_PAIR_DEFAULT
#define WU_PAIR_DEFAULT 1
#endif
static bool conv_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WU_CONV_PAIR");
    v = e ? (atoi(e) != 0) : WU_PAIR_DEFAULT;
  }
  return v != 0;
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[(tap,ci)][co] = sum_px X[px + shift(tap)][ci] * dY[px][co]
// ------------------------------------------------------------------------------------------------
// Both operands are "MN-major" for the tensor core (the contiguous smem dimension is the channel,
// the GEMM K dimension is the pixel).  One CTA owns M = 128 = two 64-wide (tap, channel-block)
// atoms and N = BN output channels, and walks a contiguous range of 64-pixel boxes (split-K).
struct WgradParams {
  int c0_blocks, ctot_blocks;
  int atoms;  // 9 * ctot_blocks
  int n_tiles, splits;
  int tiles_w, tiles_h, batch;
  int bw, bh;
  int pix_tiles;
  int cout, cin_total;
  float* partial;  // [splits][9*cin_total][cout]
};

template <int BN>
struct WgradCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kAtomBytes = 64 * 128;  // 64 pixels x 64 ch x 2 B
  static constexpr int kABytes = 2 * kAtomBytes;
  static constexpr int kBBytes = (BN / 64) * kAtomBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
  static constexpr uint32_t kTmemCols = BN;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0,
                     const __grid_constant__ CUtensorMap tmX1,
                     const __grid_constant__ CUtensorMap tmY, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // blockIdx.x -> (pair, n tile, split)
  int id = blockIdx.x;
  const int z = id % p.splits;
  id /= p.splits;
  const int nt = id % p.n_tiles;
  const int pair = id / p.n_tiles;
  const int n0 = nt * BN;
  const int pt_begin = (int)(((long long)p.pix_tiles * z) / p.splits);
  const int pt_end = (int)(((long long)p.pix_tiles * (z + 1)) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int a_idx[2], a_tap_r[2], a_tap_s[2], a_cb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int a = 2 * pair + j;
        if (a >= p.atoms) a = p.atoms - 1;  // odd atom count: duplicate, result discarded
        a_idx[j] = a;
        const int tap = a / p.ctot_blocks;
        a_cb[j] = a - tap * p.ctot_blocks;
        a_tap_r[j] = tap / 3;
        a_tap_s[j] = tap - 3 * a_tap_r[j];
      }
      (void)a_idx;
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        int m = pt;
        const int tw = m % p.tiles_w;
        m /= p.tiles_w;
        const int th = m % p.tiles_h;
        const int b = m / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t fb = full_bar(stage);
        mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
        const uint32_t a_dst = base + stage * Cfg::kStageBytes;
        const uint32_t b_dst = a_dst + Cfg::kABytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cw = w0 + a_tap_s[j] - 1, ch = h0 + a_tap_r[j] - 1;
          if (a_cb[j] < p.c0_blocks)
            tma_load_4d(a_dst + j * Cfg::kAtomBytes, &tmX0, fb, a_cb[j] * 64, cw, ch, b);
          else
            tma_load_4d(a_dst + j * Cfg::kAtomBytes, &tmX1, fb, (a_cb[j] - p.c0_blocks) * 64, cw,
                        ch, b);
        }
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(b_dst + j * Cfg::kAtomBytes, &tmY, fb, n0 + j * 64, w0, h0, b);
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::kStageBytes;
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 pixels); 16 pixel rows = 2048 bytes
          const uint64_t adesc =
              umma_smem_desc_sw128(a_addr + k * 2048, Cfg::kAtomBytes, 1024);
          const uint64_t bdesc =
              umma_smem_desc_sw128(b_addr + k * 2048, Cfg::kAtomBytes, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (pt != pt_begin || k != 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int a = 2 * pair + (row >> 6);
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* out = p.partial + ((size_t)z * 9 * p.cin_total + (size_t)a * 64 + (row & 63)) * p.cout + n0;
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + chunk * 32, v);
      tmem_ld_wait();
      if (a < p.atoms) {
        float4* o = reinterpret_cast<float4*>(out + chunk * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// wgrad, v2: operand reuse in shared memory
// ------------------------------------------------------------------------------------------------
// v1 gives every CTA two (tap, channel block) atoms, so the dY tile is re-fetched for each pair
// and every X tile once per tap (ncu: L2 -> SMEM bound, tensor pipe ~22 %).  v2 gives a CTA up to
// NC "copies" = (channel block cb, column shift s) pairs; one TMA box (64 ch, 8 px, 8+2 rows) per
// copy and pixel tile serves the three taps r = 0..2 of that column shift through start-address
// offsets of r KiB (8 pixels x 128 B: 1024-byte aligned), and one dY tile feeds all 3*NC atoms:
//   M-blocks (128 accumulator lanes = two 64-row atoms) per CTA:
//     [copy c: r=0 | r=1]            atoms 1 KiB apart   -> LBO = 1024      (one per copy)
//     [r=2 of copy c | r=2 of c+1]   atoms one copy apart -> LBO = 10 KiB   (one per copy pair)
//   TMEM columns = (NC + ceil(NC/2)) * BN  (NC = 5, BN = 64: 512;  NC = 2, BN = 128: 384).
// A lone r=2 atom is issued as an M=128 block whose upper half is never read back.
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

struct WgradParams2 {
  int c0_blocks, ctot_blocks;
  int copies;  // 3 * ctot_blocks
  int groups;  // ceil(copies / NC)
  int n_tiles, splits;
  int tiles_w, tiles_h, batch;
  int pix_tiles;
  int cout, cin_total;
  float* partial;       // [splits][9*cin_total][cout]
  float* bias_partial;  // [4*splits][cout] column sums of dY (bias gradient), or null
  int merge;            // every CTA's NC copies come from one source tensor: one 5-D TMA box per stage
};

template <int BN, int NC>
struct WgradCfg2 {
  static constexpr int kCopyBytes = 10 * 1024;  // (8+2) rows x 8 px x 128 B
  static constexpr int kAtomBytes = 64 * 128;
  static constexpr int kABytes = NC * kCopyBytes;
  static constexpr int kBBytes = (BN / 64) * kAtomBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (220 * 1024) / kStageBytes > 6 ? 6 : (220 * 1024) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
  static constexpr int kBlocks = NC + (NC + 1) / 2;
  static constexpr uint32_t kTmemColsUsed = kBlocks * BN;
  static constexpr uint32_t kTmemCols = kTmemColsUsed <= 256 ? 256 : 512;
  static_assert(kTmemColsUsed <= 512, "TMEM budget");
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kStageBytes % 1024 == 0, "stage alignment");
};

template <int BN, int NC>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_v2_kernel(const __grid_constant__ CUtensorMap tmX0,
                        const __grid_constant__ CUtensorMap tmX1,
                        const __grid_constant__ CUtensorMap tmX0m,
                        const __grid_constant__ CUtensorMap tmX1m,
                        const __grid_constant__ CUtensorMap tmY, const WgradParams2 p) {
  using Cfg = WgradCfg2<BN, NC>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + S * Cfg::kStageBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (S + i); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // blockIdx.x -> (split, copy group, n tile) with the split slowest: the CTAs that walk the same
  // pixel range (all copy groups / n tiles of one split) are neighbours in launch order, run at
  // the same time and share their X / dY tiles through L2 instead of re-reading them from HBM
  int id = blockIdx.x;
  const int nt = id % p.n_tiles;
  id /= p.n_tiles;
  const int grp = id % p.groups;
  const int z = id / p.groups;
  const int n0 = nt * BN;
  const int g0 = grp * NC;                                       // first global copy id
  const int nc = (p.copies - g0) < NC ? (p.copies - g0) : NC;    // copies of this CTA
  const int pt_begin = (int)(((long long)p.pix_tiles * z) / p.splits);
  const int pt_end = (int)(((long long)p.pix_tiles * (z + 1)) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmX0m);
    tma_prefetch_desc(&tmX1m);
    tma_prefetch_desc(&tmY);
  }
  // Bias gradient db[co] = sum_px dY[px][co]: the dY tiles pass through this CTA's shared memory
  // anyway, so the four warps that otherwise only drain the accumulator at the end add them up
  // column-wise while the MMAs run (copy group 0 of every (split, n tile) only: all groups of a split
  // see the same tiles).  A separate pass re-read every dY from HBM (0.86 ms per training iteration).
  const bool do_bias = p.bias_partial != nullptr && grp == 0;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), do_bias ? 5 : 1);  // MMA commit (+ one arrival per summing warp)
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(nc * Cfg::kCopyBytes + Cfg::kBBytes);
      const bool merged = p.merge != 0 && nc == NC;
      WU_STAT_DECL(2);
#ifdef WU_PIPE_STATS
      const long long _p0 = clock64();
#endif
      // This ONE thread feeds the whole CTA: every instruction it executes per stage delays the next TMA
      // issue, and the MMAs starve behind it (measured: adding two integer divisions per copy and stage
      // cost the N = 64 layers 18-27 %; merging two dY boxes into one gained 4-6 %).  So everything that
      // does not change from stage to stage is computed here, once: which tensor map, channel
      // coordinate and column shift each copy uses, and the tile coordinates advance by increments.
      // Global copy id g = s * ctot_blocks + cb: a CTA's copies are consecutive channel blocks of ONE
      // column shift (the other shifts of the same blocks belong to neighbouring CTAs of the split and
      // are fetched at about the same time: L2 hits).  When they all come from one source tensor, ONE
      // 5-D TMA box fetches them (the copies land back to back in shared memory).
      const CUtensorMap* cmap[NC];
      int cch[NC], csh[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int g = g0 + (c < nc ? c : 0);
        const int s = g / p.ctot_blocks, cb = g - s * p.ctot_blocks;
        const bool first = cb < p.c0_blocks;
        csh[c] = s - 1;
        if (merged) {
          cmap[c] = first ? &tmX0m : &tmX1m;
          cch[c] = first ? cb : cb - p.c0_blocks;  // block coordinate of the 5-D map
        } else {
          cmap[c] = first ? &tmX0 : &tmX1;
          cch[c] = (first ? cb : cb - p.c0_blocks) * 64;
        }
      }
      const int ny = n0 >> 6;
      int tw = pt_begin % p.tiles_w;
      int th = (pt_begin / p.tiles_w) % p.tiles_h;
      int b = pt_begin / (p.tiles_w * p.tiles_h);
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int w0 = tw * 8, h0 = th * 8;
        WU_STAT_WAIT(1, mbar_wait(empty_bar(stage), phase ^ 1u));
        const uint32_t fb = full_bar(stage);
        mbar_arrive_expect_tx(fb, tx);
        const uint32_t a_dst = base + stage * Cfg::kStageBytes;
        if (merged) {
          tma_load_5d(a_dst, cmap[0], fb, 0, w0 + csh[0], h0 - 1, cch[0], b);
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            if (c < nc) tma_load_4d(a_dst + c * Cfg::kCopyBytes, cmap[c], fb, cch[c], w0 + csh[c], h0 - 1, b);
        }
        // one 5-D box fetches the BN / 64 channel-block tiles of dY (they land back to back)
        tma_load_5d(a_dst + Cfg::kABytes, &tmY, fb, 0, w0, h0, ny, b);
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++b;
          }
        }
      }
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _p0;
#endif
      WU_STAT_FLUSH(12, 2);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);  // both operands MN-major
      WU_STAT_DECL(2);
#ifdef WU_PIPE_STATS
      const long long _m0 = clock64();
#endif
      const int npair = (nc + 1) >> 1;
      const int nblocks = nc + npair;
      // One thread issues every MMA, so its instruction count per MMA is what paces the tensor
      // pipe at small N: build each accumulator block's A descriptor ONCE (stage 0, k = 0) and
      // only add (byte offset >> 4) to the start-address field per stage / k step.
      uint64_t adesc0[Cfg::kBlocks];
#pragma unroll
      for (int blk = 0; blk < Cfg::kBlocks; ++blk) {
        uint32_t addr, lbo;
        if (blk < NC) {  // [r=0 | r=1] of copy blk
          addr = base + blk * Cfg::kCopyBytes;
          lbo = 1024u;
        } else {         // [r=2 of copy 2j | r=2 of copy 2j+1]
          const int j = blk - NC;
          addr = base + 2 * j * Cfg::kCopyBytes + 2 * 1024;
          lbo = (2 * j + 1 < nc) ? (uint32_t)Cfg::kCopyBytes : 1024u;
        }
        adesc0[blk] = umma_smem_desc_sw128(addr, lbo, 1024);
      }
      const uint64_t bdesc0 = umma_smem_desc_sw128(base + Cfg::kABytes, Cfg::kAtomBytes, 1024);
      int stage = 0;
      uint32_t phase = 0;
      // FULL: every copy slot of the CTA is in use (nc == NC) -> no per-MMA liveness tests
      auto issue_stage = [&](auto full_tag, int pt) {
        constexpr bool FULL = decltype(full_tag)::value;
        WU_STAT_WAIT(1, mbar_wait(full_bar(stage), phase));
        tc_fence_after();
        const uint64_t soff = (uint64_t)((stage * Cfg::kStageBytes) >> 4);
        const uint32_t acc = pt != pt_begin ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 pixels = 2 rows of 8 = 2 KiB)
          const uint64_t koff = soff + (uint64_t)((k * 2048) >> 4);
          const uint64_t bdesc = bdesc0 + koff;
          const uint32_t accum = (acc | (uint32_t)k) != 0 ? 1u : 0u;
#pragma unroll
          for (int blk = 0; blk < Cfg::kBlocks; ++blk) {
            // TMEM column of block blk: copies first (blk < nc), then the r=2 pairs
            const bool is_copy = blk < NC;
            if (FULL) {
              umma_bf16(tmem_base + blk * BN, adesc0[blk] + koff, bdesc, idesc, accum);
            } else {
              const bool live = is_copy ? (blk < nc) : (blk - NC < npair);
              if (live) {
                const int col = is_copy ? blk : nc + (blk - NC);
                umma_bf16(tmem_base + col * BN, adesc0[blk] + koff, bdesc, idesc, accum);
              }
            }
          }
        }
        umma_commit(empty_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      };
      if (nc == NC) {
        for (int pt = pt_begin; pt < pt_end; ++pt) issue_stage(TrueTag{}, pt);
      } else {
        for (int pt = pt_begin; pt < pt_end; ++pt) issue_stage(FalseTag{}, pt);
      }
      (void)nblocks;
      umma_commit(tfull_bar);
#ifdef WU_PIPE_STATS
      _st[0] = clock64() - _m0;
#endif
      WU_STAT_FLUSH(8, 2);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = row >> 6, r64 = row & 63;
    if (do_bias) {
      // warp q sums pixel rows 16q .. 16q+15 of every dY tile; lane = a pair of adjacent channels
      // (4 bytes) of each 64-channel atom.  The tile is 128B-swizzled: 16-byte chunk c of row r
      // sits at chunk position c ^ (r & 7); a warp reads one 128-byte row per load (no conflicts).
      float acc[BN / 64][2];
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) acc[j][0] = acc[j][1] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(full_bar(stage), phase);
        const uint8_t* bt = smem + stage * Cfg::kStageBytes + Cfg::kABytes;
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = q * 16 + rr;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(
                bt + j * Cfg::kAtomBytes + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
            acc[j][0] += bf16lo(v);
            acc[j][1] += bf16hi(v);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      float* bp = p.bias_partial + ((size_t)(z * 4 + q)) * p.cout + n0;
#pragma unroll
      for (int j = 0; j < BN / 64; ++j)
        *reinterpret_cast<float2*>(bp + j * 64 + lane * 2) = make_float2(acc[j][0], acc[j][1]);
    }
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int nblocks = nc + ((nc + 1) >> 1);
#pragma unroll 1
    for (int blk = 0; blk < nblocks; ++blk) {
      // which (tap row r, copy c) does this lane's atom hold?
      int r, c;
      if (blk < nc) {
        c = blk;
        r = half;
      } else {
        c = 2 * (blk - nc) + half;
        r = 2;
      }
      const bool valid = c < nc;
      const int g = g0 + (valid ? c : 0);
      const int s = g / p.ctot_blocks, cb = g - s * p.ctot_blocks;
      const int tap = r * 3 + s;
      float* out = p.partial +
                   ((size_t)z * 9 * p.cin_total + (size_t)tap * p.cin_total + cb * 64 + r64) * p.cout + n0;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + blk * BN + chunk * 32, v);
        tmem_ld_wait();
        if (valid) {
          float4* o = reinterpret_cast<float4*>(out + chunk * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN, int NC>
static int launch_wgrad2(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& x0m,
                         const CUtensorMap& x1m, const CUtensorMap& ym, const WgradParams2& p, int grid,
                         cudaStream_t st) {
  using Cfg = WgradCfg2<BN, NC>;
  static bool attr_done = false;
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_v2_kernel<BN, NC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  conv3x3_wgrad_v2_kernel<BN, NC><<<grid, 192, Cfg::kSmemBytes, st>>>(x0, x1, x0m, x1m, ym, p);
  WU_CHECK_LAUNCH("conv3x3_wgrad_v2_kernel");
  return WU_OK;
}

// dw[co][ci][tap] = sum_z partial[z][tap*cin + ci][co]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                    int splits, int cin, int cout) {
  const long long total = 9LL * cin * cout;
  const long long stride_z = total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const long long rc = i / cout;  // tap*cin + ci
    const int ci = (int)(rc % cin);
    const int tap = (int)(rc / cin);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += partial[z * stride_z + i];
    dw[((long long)co * cin + ci) * 9 + tap] = acc;
  }
}

// Same, and the blocks past `main_blocks` fold the in-kernel bias partials (rows x cout) into db:
// one launch instead of two for every convolution's weight + bias gradient.
__global__ void __launch_bounds__(256)
wgrad_reduce_bias_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int cin,
                         int cout, int main_blocks, const float* __restrict__ bias_partial, int rows,
                         float* __restrict__ db) {
  if ((int)blockIdx.x < main_blocks) {
    const long long total = 9LL * cin * cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)main_blocks * blockDim.x) {
      const int co = (int)(i % cout);
      const long long rc = i / cout;
      const int ci = (int)(rc % cin);
      const int tap = (int)(rc / cin);
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += partial[z * total + i];
      dw[((long long)co * cin + ci) * 9 + tap] = acc;
    }
    return;
  }
  __shared__ float red[32][8];
  const int cl = threadIdx.x & 7, part = threadIdx.x >> 3;
  const int c = ((int)blockIdx.x - main_blocks) * 8 + cl;
  float sum = 0.f;
  if (c < cout)
    for (int b = part; b < rows; b += 32) sum += bias_partial[(size_t)b * cout + c];
  red[part][cl] = sum;
  __syncthreads();
  if (part == 0 && c < cout) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][cl];
    db[c] = t;
  }
}

// db[c] = sum_p dy[p][c]; stage 1: per-block partial sums (deterministic), stage 2: fold.
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                         long long npix, int C) {
  extern __shared__ float red[];  // [groups][C]
  const int lanes = C / 8;        // threads per pixel (16 B each)
  const int groups = blockDim.x / lanes;
  const int g = threadIdx.x / lanes, l = threadIdx.x % lanes;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g < groups) {
    const long long stride = (long long)gridDim.x * groups;
    long long px = (long long)blockIdx.x * groups + g;
    for (; px + 3 * stride < npix; px += 4 * stride) {  // four independent 16-byte loads in flight
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(reinterpret_cast<const uint4*>(dy + (px + u * stride) * C) + l);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16lo(v[u].x); acc[1] += bf16hi(v[u].x);
        acc[2] += bf16lo(v[u].y); acc[3] += bf16hi(v[u].y);
        acc[4] += bf16lo(v[u].z); acc[5] += bf16hi(v[u].z);
        acc[6] += bf16lo(v[u].w); acc[7] += bf16hi(v[u].w);
      }
    }
    for (; px < npix; px += stride) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + px * C) + l);
      acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x);
      acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
      acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z);
      acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[g * C + l * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[gg * C + c];
    partial[(size_t)blockIdx.x * C + c] = s;
  }
}
// db[c] = sum_b partial[b][c]: one block per 8 channels, 32 threads share a channel's row range
// (rows = 4 x splits <= 592: a longer chain per thread made this a 10 us kernel for 64 channels)
__global__ void __launch_bounds__(256)
bias_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ db, int nblocks,
                       int C) {
  __shared__ float red[32][8];
  const int cl = threadIdx.x & 7, part = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float s = 0.f;
  if (c < C)
    for (int b = part; b < nblocks; b += 32) s += partial[(size_t)b * C + c];
  red[part][cl] = s;
  __syncthreads();
  if (part == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][cl];
    db[c] = t;
  }
}

template <int BN>
static int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& ym,
                        const WgradParams& p, int grid, cudaStream_t st) {
  using Cfg = WgradCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  conv3x3_wgrad_kernel<BN><<<grid, 192, Cfg::kSmemBytes, st>>>(x0, x1, ym, p);
  WU_CHECK_LAUNCH("conv3x3_wgrad_kernel");
  return WU_OK;
}

struct WgradPlan {
  int bn, n_tiles, pairs, atoms, splits, bw, bh, tiles_w, tiles_h, pix_tiles;
  size_t partial_bytes, bias_bytes;
  int bias_blocks;
  int v2, nc, groups;
};

static WgradPlan plan_wgrad(int cin_total, int cout, int B, int H, int W) {
  WgradPlan pl;
  pl.v2 = conv_impl() != 1;
  if (pl.v2) {
    pl.bn = cout >= 128 ? 128 : 64;
    // cout = 64: five copies per CTA when the 3 * cin/64 copies divide evenly, else three (cin = 192 has
    // nine: 5 + 4 left half of the SMs 20 % short of work in the second wave; 3 + 3 + 3 is 3 % faster)
    const int copies64 = 3 * (cin_total / 64);
    pl.nc = pl.bn == 64 ? (copies64 % 5 == 0 ? 5 : 3) : 2;
    pl.n_tiles = cout / pl.bn;
    const int copies = 3 * (cin_total / 64);
    pl.groups = (copies + pl.nc - 1) / pl.nc;
    pl.atoms = 9 * (cin_total / 64);
    pl.pairs = 0;
    pl.bw = 8;
    pl.bh = 8;
    pl.tiles_w = (W + 7) / 8;
    pl.tiles_h = (H + 7) / 8;
    pl.pix_tiles = B * pl.tiles_w * pl.tiles_h;
    // one CTA per SM (shared memory): size the grid to fill whole waves — the largest split count
    // whose grid still fits in two waves (a 297-CTA grid on 148 SMs would run three)
    const int tiles = pl.groups * pl.n_tiles;
    const int slots = 2 * num_sms();
    int splits = slots / tiles;
    if (splits > pl.pix_tiles) splits = pl.pix_tiles;
    if (splits < 1) splits = 1;
    if (splits > num_sms()) splits = num_sms();
    if (splits > kBiasGradBlocks / 4) splits = kBiasGradBlocks / 4;  // in-kernel bias sums: 4 rows per split
    pl.splits = splits;
    pl.partial_bytes = (size_t)splits * 9 * cin_total * cout * sizeof(float);
    pl.bias_blocks = kBiasGradBlocks;
    pl.bias_bytes = (size_t)pl.bias_blocks * cout * sizeof(float);
    return pl;
  }
  pl.nc = pl.groups = 0;
  pl.bn = cout >= 256 ? 256 : cout;  // 64 / 128 / 256
  pl.n_tiles = cout / pl.bn;
  pl.atoms = 9 * (cin_total / 64);
  pl.pairs = (pl.atoms + 1) / 2;
  pick_box(H, W, 64, &pl.bw, &pl.bh);
  pl.tiles_w = (W + pl.bw - 1) / pl.bw;
  pl.tiles_h = (H + pl.bh - 1) / pl.bh;
  pl.pix_tiles = B * pl.tiles_w * pl.tiles_h;
  const int tiles = pl.pairs * pl.n_tiles;
  int splits = (2 * 148 + tiles - 1) / tiles;  // about two waves of CTAs
  if (splits > pl.pix_tiles) splits = pl.pix_tiles;
  if (splits < 1) splits = 1;
  if (splits > 128) splits = 128;
  pl.splits = splits;
  pl.partial_bytes = (size_t)splits * 9 * cin_total * cout * sizeof(float);
  pl.bias_blocks = kBiasGradBlocks;
  pl.bias_bytes = (size_t)pl.bias_blocks * cout * sizeof(float);
  return pl;
}

// ------------------------------------------------------------------------------------------------
// weight pack: fp32 [cout][cin][3][3] -> bf16 fprop [cout][9*cin] and dgrad [cin][9*cout]
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                    __nv_bfloat16* __restrict__ wd, int cout, int cin) {
  const long long total = (long long)cout * cin * 9;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const long long cc = i / 9;
    const int ci = (int)(cc % cin);
    const int co = (int)(cc / cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    wf[(long long)co * 9 * cin + (long long)tap * cin + ci] = v;
    if (wd != nullptr) wd[(long long)ci * 9 * cout + (long long)(8 - tap) * cout + co] = v;
  }
}

}  // namespace wu

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace wu;

#ifdef WU_PIPE_STATS
extern "C" int wu_debug_pipe_stats(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, g_pipe_stats, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_pipe_stats, z, sizeof(z));
  }
  return WU_OK;
}
#endif

extern "C" int wu_pack_conv3x3_weights(const float* w, int cout, int cin, void* w_fprop,
                                       void* w_dgrad, wu_stream_t stream) {
  WU_REQUIRE(w && w_fprop, "wu_pack_conv3x3_weights: null pointer");
  WU_REQUIRE(cout > 0 && cin > 0, "wu_pack_conv3x3_weights: bad shape cout=%d cin=%d", cout, cin);
  const long long total = (long long)cout * cin * 9;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      w, (__nv_bfloat16*)w_fprop, (__nv_bfloat16*)w_dgrad, cout, cin);
  WU_CHECK_LAUNCH("pack_weights_kernel");
  return WU_OK;
}

extern "C" int wu_conv3x3_fprop(const void* src0, int c0, const void* src1, int c1,
                                const void* w_packed, const float* bias, int relu,
                                const void* relu_mask_src, void* dst, int cout, int B, int H, int W,
                                wu_stream_t stream) {
  return wu_conv3x3_fprop_bcast(src0, c0, src1, c1, 0, w_packed, bias, relu, relu_mask_src, dst, cout,
                                B, H, W, stream);
}

// Number of 64-pixel statistics chunks per image the STATS instantiations write (0: unsupported)
static int fprop_stats_chunks(int cout, int H, int W) {
  if (cout != 128 && cout % 256 != 0) return 0;
  if (cout == 128) {  // v2 <128, 2>: super-tiles of 32 rows x 8 columns, two 16-row tiles each
    return ((W + 7) / 8) * ((H + 31) / 32) * 2 * 2;
  }
  int bw, bh;  // v1 <256>: one 128-pixel box per tile
  pick_box(H, W, 128, &bw, &bh);
  return ((W + bw - 1) / bw) * ((H + bh - 1) / bh) * 2;
}

extern "C" int wu_conv3x3_stats_chunks(int cout, int H, int W) {
  if (cout <= 0 || H <= 0 || W <= 0 || conv_impl() != 0) return 0;
  return fprop_stats_chunks(cout, H, W);
}

static int conv3x3_fprop_impl(const void* src0, int c0, const void* src1, int c1, int src1_bcast,
                              const void* w_packed, const float* bias, int relu,
                              const void* relu_mask_src, void* dst, int cout, int B, int H, int W,
                              float* stats, wu_stream_t stream);

extern "C" int wu_conv3x3_fprop_bcast(const void* src0, int c0, const void* src1, int c1,
                                      int src1_bcast, const void* w_packed, const float* bias,
                                      int relu, const void* relu_mask_src, void* dst, int cout,
                                      int B, int H, int W, wu_stream_t stream) {
  return conv3x3_fprop_impl(src0, c0, src1, c1, src1_bcast, w_packed, bias, relu, relu_mask_src, dst,
                            cout, B, H, W, nullptr, stream);
}

extern "C" int wu_conv3x3_fprop_stats(const void* src0, int c0, const void* src1, int c1,
                                      int src1_bcast, const void* w_packed, const float* bias,
                                      void* dst, float* stats, int cout, int B, int H, int W,
                                      wu_stream_t stream) {
  WU_REQUIRE(stats != nullptr, "wu_conv3x3_fprop_stats: null stats pointer");
  WU_REQUIRE(wu_conv3x3_stats_chunks(cout, H, W) > 0,
             "wu_conv3x3_fprop_stats: cout=%d must be 128 or a multiple of 256", cout);
  WU_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 7) == 0, "wu_conv3x3_fprop_stats: stats unaligned");
  return conv3x3_fprop_impl(src0, c0, src1, c1, src1_bcast, w_packed, bias, 1, nullptr, dst, cout, B, H,
                            W, stats, stream);
}

static int conv3x3_fprop_impl(const void* src0, int c0, const void* src1, int c1, int src1_bcast,
                              const void* w_packed, const float* bias, int relu,
                              const void* relu_mask_src, void* dst, int cout, int B, int H, int W,
                              float* stats, wu_stream_t stream) {
  WU_REQUIRE(src0 && w_packed && dst, "wu_conv3x3_fprop: null pointer");
  WU_REQUIRE(B > 0 && H > 0 && W > 0, "wu_conv3x3_fprop: bad shape B=%d H=%d W=%d", B, H, W);
  WU_REQUIRE(c0 > 0 && c0 % 64 == 0, "wu_conv3x3_fprop: c0=%d must be a positive multiple of 64", c0);
  WU_REQUIRE(c1 >= 0 && c1 % 64 == 0 && (c1 == 0) == (src1 == nullptr),
             "wu_conv3x3_fprop: c1=%d must be a multiple of 64 and match src1", c1);
  WU_REQUIRE(cout > 0 && cout % 64 == 0 && cout != 192 && (cout <= 256 || cout % 256 == 0),
             "wu_conv3x3_fprop: cout=%d must be 64, 128 or a multiple of 256", cout);
  const int bn = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  if (conv_impl() == 2 || (conv_impl() == 0 && bn < 256)) {
    const int T = bn == 64 ? 4 : (bn == 128 ? 2 : 1);
    const int bw = 8;  // pixels per row of an A box
    ConvParams2 q;
    q.c0_blocks = c0 / 64;
    q.ctot_blocks = (c0 + c1) / 64;
    q.tiles_w = (W + 7) / 8;
    q.tiles_h = (H + 16 * T - 1) / (16 * T);
    q.batch = B;
    q.n_tiles = cout / bn;
    const long long nt2 = (long long)B * q.tiles_w * q.tiles_h * q.n_tiles;
    WU_REQUIRE(nt2 < (1LL << 31), "wu_conv3x3_fprop: too many tiles");
    q.num_tiles = (int)nt2;
    q.H = H;
    q.W = W;
    q.cout = cout;
    q.relu = relu;
    q.b1_mul = src1_bcast ? 0 : 1;
    q.bias = bias;
    q.mask = (const __nv_bfloat16*)relu_mask_src;
    q.last_w = q.last_b = nullptr;
    q.last_y = nullptr;
    q.pool_dst = nullptr;
    q.stats = stats;
    q.stats_chunks = stats ? fprop_stats_chunks(cout, H, W) : 0;
    CUtensorMap a0, a1, bm, dm;
    int rc;
    if ((rc = make_act_tmap(&a0, src0, B, H, W, c0, c0, bw, 16 * T + 2)) != WU_OK) return rc;
    if (c1 > 0) {
      if ((rc = make_act_tmap(&a1, src1, src1_bcast ? 1 : B, H, W, c1, c1, bw, 16 * T + 2)) != WU_OK)
        return rc;
    } else {
      a1 = a0;
    }
    if ((rc = make_mat_tmap(&bm, w_packed, cout, 9 * (c0 + c1), bn)) != WU_OK) return rc;
    if ((rc = make_act_tmap(&dm, dst, B, H, W, cout, cout, 8, 16)) != WU_OK) return rc;
    cudaStream_t st2 = (cudaStream_t)stream;
    if (stats != nullptr) {
      WU_REQUIRE(bn == 128, "wu_conv3x3_fprop_stats: internal: v2 statistics need N = 128");
      return launch_conv2<128, 2, false, false, true>(a0, a1, bm, dm, q, st2);
    }
    switch (bn) {
      case 64: return launch_conv2<64, 4>(a0, a1, bm, dm, q, st2);
      case 128: return launch_conv2<128, 2>(a0, a1, bm, dm, q, st2);
      default: return launch_conv2<256, 1>(a0, a1, bm, dm, q, st2);
    }
  }
  ConvParams p;
  pick_box(H, W, 128, &p.bw, &p.bh);
  p.log2_bw = ilog2(p.bw);
  p.c0_blocks = c0 / 64;
  p.ctot_blocks = (c0 + c1) / 64;
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.batch = B;
  p.n_tiles = cout / bn;
  const long long nt = (long long)B * p.tiles_w * p.tiles_h * p.n_tiles;
  WU_REQUIRE(nt < (1LL << 31), "wu_conv3x3_fprop: too many tiles");
  p.num_tiles = (int)nt;
  p.H = H;
  p.W = W;
  p.cout = cout;
  p.relu = relu;
  p.b1_mul = src1_bcast ? 0 : 1;
  p.bias = bias;
  p.mask = (const __nv_bfloat16*)relu_mask_src;
  p.stats = stats;
  p.stats_chunks = stats ? fprop_stats_chunks(cout, H, W) : 0;
  CUtensorMap a0, a1, bm, dm;
  int rc;
  if ((rc = make_act_tmap(&a0, src0, B, H, W, c0, c0, p.bw, p.bh)) != WU_OK) return rc;
  if (c1 > 0) {
    if ((rc = make_act_tmap(&a1, src1, src1_bcast ? 1 : B, H, W, c1, c1, p.bw, p.bh)) != WU_OK)
      return rc;
  } else {
    a1 = a0;
  }
  const bool pair = bn == 256 && conv_pair();  // CTA pairs (cta_group::2): half a weight tile per CTA
  if ((rc = make_mat_tmap(&bm, w_packed, cout, 9 * (c0 + c1), pair ? 128 : bn)) != WU_OK) return rc;
  if ((rc = make_act_tmap(&dm, dst, B, H, W, cout, cout, p.bw, p.bh)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (pair)
    return stats != nullptr ? launch_conv_pair<true>(a0, a1, bm, dm, p, st)
                            : launch_conv_pair<false>(a0, a1, bm, dm, p, st);
  if (stats != nullptr) {
    WU_REQUIRE(bn == 256, "wu_conv3x3_fprop_stats: internal: v1 statistics need N = 256");
    return launch_conv<256, true>(a0, a1, bm, dm, p, st);
  }
  switch (bn) {
    case 64: return launch_conv<64>(a0, a1, bm, dm, p, st);
    case 128: return launch_conv<128>(a0, a1, bm, dm, p, st);
    default: return launch_conv<256>(a0, a1, bm, dm, p, st);
  }
}

extern "C" int wu_conv3x3_fprop_last(const void* src, int cin, const void* w_packed, const float* bias,
                                     void* dst, const float* last_w, const float* last_b, float* y,
                                     int B, int H, int W, wu_stream_t stream) {
  WU_REQUIRE(src && w_packed && dst && last_w && y, "wu_conv3x3_fprop_last: null pointer");
  WU_REQUIRE(B > 0 && H > 0 && W > 0, "wu_conv3x3_fprop_last: bad shape B=%d H=%d W=%d", B, H, W);
  WU_REQUIRE(cin > 0 && cin % 64 == 0, "wu_conv3x3_fprop_last: cin=%d must be a positive multiple of 64", cin);
  constexpr int T = 4;
  ConvParams2 q;
  q.c0_blocks = q.ctot_blocks = cin / 64;
  q.tiles_w = (W + 7) / 8;
  q.tiles_h = (H + 16 * T - 1) / (16 * T);
  q.batch = B;
  q.n_tiles = 1;
  const long long nt = (long long)B * q.tiles_w * q.tiles_h;
  WU_REQUIRE(nt < (1LL << 31), "wu_conv3x3_fprop_last: too many tiles");
  q.num_tiles = (int)nt;
  q.H = H;
  q.W = W;
  q.cout = 64;
  q.relu = 1;
  q.b1_mul = 1;
  q.bias = bias;
  q.mask = nullptr;
  q.last_w = last_w;
  q.last_b = last_b;
  q.last_y = y;
  q.pool_dst = nullptr;
  q.stats = nullptr;
  q.stats_chunks = 0;
  CUtensorMap a0, bm, dm;
  int rc;
  if ((rc = make_act_tmap(&a0, src, B, H, W, cin, cin, 8, 16 * T + 2)) != WU_OK) return rc;
  if ((rc = make_mat_tmap(&bm, w_packed, 64, 9 * cin, 64)) != WU_OK) return rc;
  if ((rc = make_act_tmap(&dm, dst, B, H, W, 64, 64, 8, 16)) != WU_OK) return rc;
  return launch_conv2<64, T, true>(a0, a0, bm, dm, q, (cudaStream_t)stream);
}

extern "C" int wu_conv3x3_fprop_pool(const void* src, int cin, const void* w_packed, const float* bias,
                                     void* dst, void* pool_dst, int cout, int B, int H, int W,
                                     wu_stream_t stream) {
  WU_REQUIRE(src && w_packed && dst && pool_dst, "wu_conv3x3_fprop_pool: null pointer");
  WU_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
             "wu_conv3x3_fprop_pool: bad shape B=%d H=%d W=%d (H, W must be even)", B, H, W);
  WU_REQUIRE(cin > 0 && cin % 64 == 0, "wu_conv3x3_fprop_pool: cin=%d must be a positive multiple of 64", cin);
  WU_REQUIRE(cout == 64 || cout == 128, "wu_conv3x3_fprop_pool: cout=%d must be 64 or 128", cout);
  const int T = cout == 64 ? 4 : 2;
  ConvParams2 q;
  q.c0_blocks = q.ctot_blocks = cin / 64;
  q.tiles_w = (W + 7) / 8;
  q.tiles_h = (H + 16 * T - 1) / (16 * T);
  q.batch = B;
  q.n_tiles = 1;
  const long long nt = (long long)B * q.tiles_w * q.tiles_h;
  WU_REQUIRE(nt < (1LL << 31), "wu_conv3x3_fprop_pool: too many tiles");
  q.num_tiles = (int)nt;
  q.H = H;
  q.W = W;
  q.cout = cout;
  q.relu = 1;
  q.b1_mul = 1;
  q.bias = bias;
  q.mask = nullptr;
  q.last_w = q.last_b = nullptr;
  q.last_y = nullptr;
  q.pool_dst = (__nv_bfloat16*)pool_dst;
  q.stats = nullptr;
  q.stats_chunks = 0;
  CUtensorMap a0, bm, dm;
  int rc;
  if ((rc = make_act_tmap(&a0, src, B, H, W, cin, cin, 8, 16 * T + 2)) != WU_OK) return rc;
  if ((rc = make_mat_tmap(&bm, w_packed, cout, 9 * cin, cout)) != WU_OK) return rc;
  if ((rc = make_act_tmap(&dm, dst, B, H, W, cout, cout, 8, 16)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return cout == 64 ? launch_conv2<64, 4, false, true>(a0, a0, bm, dm, q, st)
                    : launch_conv2<128, 2, false, true>(a0, a0, bm, dm, q, st);
}

// split-K fold into the reference layout + bias gradient (shared with wu_conv_s2.cu)
namespace wu {
int wgrad_fold(const float* partial, int splits, int cin, int cout, float* dw, const void* dy,
               long long npix, float* db, float* bias_scratch, cudaStream_t st) {
  {
    const long long total = 9LL * cin * cout;
    int g = (int)((total + 255) / 256);
    if (g > 148 * 16) g = 148 * 16;
    wgrad_reduce_kernel<<<g, 256, 0, st>>>(partial, dw, splits, cin, cout);
    WU_CHECK_LAUNCH("wgrad_reduce_kernel");
  }
  if (db != nullptr) {
    const int lanes = cout / 8;
    const int groups = 256 / lanes;
    WU_REQUIRE(groups >= 1, "wgrad: cout=%d too wide for the bias-grad kernel", cout);
    bias_grad_partial_kernel<<<kBiasGradBlocks, 256, groups * cout * sizeof(float), st>>>(
        (const __nv_bfloat16*)dy, bias_scratch, npix, cout);
    WU_CHECK_LAUNCH("bias_grad_partial_kernel");
    bias_grad_final_kernel<<<(cout + 7) / 8, 256, 0, st>>>(bias_scratch, db, kBiasGradBlocks, cout);
    WU_CHECK_LAUNCH("bias_grad_final_kernel");
  }
  return WU_OK;
}
}  // namespace wu

static int wgrad_finish(const float* partial, float* dw, float* db, const void* dy, void* workspace,
                        const WgradPlan& pl, int cin, int cout, int B, int H, int W,
                        cudaStream_t st) {
  float* bpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + pl.partial_bytes);
  return wgrad_fold(partial, pl.splits, cin, cout, dw, dy, (long long)B * H * W, db, bpart, st);
}

extern "C" size_t wu_conv3x3_wgrad_workspace_bytes(int cin_total, int cout, int B, int H, int W) {
  if (cin_total <= 0 || cout <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  const WgradPlan pl = plan_wgrad(cin_total, cout, B, H, W);
  return pl.partial_bytes + pl.bias_bytes + 256;
}

extern "C" int wu_conv3x3_wgrad(const void* src0, int c0, const void* src1, int c1, const void* dy,
                                int cout, int B, int H, int W, float* dw, float* db,
                                void* workspace, size_t workspace_bytes, wu_stream_t stream) {
  WU_REQUIRE(src0 && dy && dw && workspace, "wu_conv3x3_wgrad: null pointer");
  WU_REQUIRE(B > 0 && H > 0 && W > 0, "wu_conv3x3_wgrad: bad shape B=%d H=%d W=%d", B, H, W);
  WU_REQUIRE(c0 > 0 && c0 % 64 == 0, "wu_conv3x3_wgrad: c0=%d must be a positive multiple of 64", c0);
  WU_REQUIRE(c1 >= 0 && c1 % 64 == 0 && (c1 == 0) == (src1 == nullptr),
             "wu_conv3x3_wgrad: c1=%d must be a multiple of 64 and match src1", c1);
  WU_REQUIRE(cout > 0 && cout % 64 == 0 && (cout <= 256 || cout % 256 == 0) && cout != 192,
             "wu_conv3x3_wgrad: unsupported cout=%d", cout);
  const int cin = c0 + c1;
  const WgradPlan pl = plan_wgrad(cin, cout, B, H, W);
  WU_REQUIRE(workspace_bytes >= pl.partial_bytes + pl.bias_bytes,
             "wu_conv3x3_wgrad: workspace %zu < required %zu", workspace_bytes,
             pl.partial_bytes + pl.bias_bytes);
  WU_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "wu_conv3x3_wgrad: workspace unaligned");
  if (pl.v2) {
    WgradParams2 q;
    q.c0_blocks = c0 / 64;
    q.ctot_blocks = cin / 64;
    q.copies = 3 * q.ctot_blocks;
    q.groups = pl.groups;
    q.n_tiles = pl.n_tiles;
    q.splits = pl.splits;
    q.tiles_w = pl.tiles_w;
    q.tiles_h = pl.tiles_h;
    q.batch = B;
    q.pix_tiles = pl.pix_tiles;
    q.cout = cout;
    q.cin_total = cin;
    q.partial = reinterpret_cast<float*>(workspace);
    float* bscratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + pl.partial_bytes);
    q.bias_partial = db != nullptr ? bscratch : nullptr;  // 4 * splits <= kBiasGradBlocks rows
    CUtensorMap x0, x1, x0m, x1m, ym;
    int rc;
    if ((rc = make_act_tmap(&x0, src0, B, H, W, c0, c0, 8, 10)) != WU_OK) return rc;
    if (c1 > 0) {
      if ((rc = make_act_tmap(&x1, src1, B, H, W, c1, c1, 8, 10)) != WU_OK) return rc;
    } else {
      x1 = x0;
    }
    // groups of pl.nc copies never straddle a column shift or the two sources -> merged 5-D loads
    q.merge = (q.ctot_blocks % pl.nc == 0 && q.c0_blocks % pl.nc == 0) ? 1 : 0;
    x0m = x0;
    x1m = x1;
    if (q.merge) {
      if ((rc = make_act_tmap_blocks(&x0m, src0, B, H, W, c0, c0, 8, 10, pl.nc)) != WU_OK) return rc;
      if (c1 > 0) {
        if ((rc = make_act_tmap_blocks(&x1m, src1, B, H, W, c1, c1, 8, 10, pl.nc)) != WU_OK) return rc;
      } else {
        x1m = x0m;
      }
    }
    if ((rc = make_act_tmap_blocks(&ym, dy, B, H, W, cout, cout, 8, 8, pl.bn / 64)) != WU_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = pl.groups * pl.n_tiles * pl.splits;
    if (pl.bn == 64 && pl.nc == 3) rc = launch_wgrad2<64, 3>(x0, x1, x0m, x1m, ym, q, grid, st);
    else if (pl.bn == 64) rc = launch_wgrad2<64, 5>(x0, x1, x0m, x1m, ym, q, grid, st);
    else rc = launch_wgrad2<128, 2>(x0, x1, x0m, x1m, ym, q, grid, st);
    if (rc != WU_OK) return rc;
    if (db == nullptr) return wgrad_fold(q.partial, pl.splits, cin, cout, dw, dy, 0, nullptr, nullptr, st);
    {  // split-K fold + fold of the in-kernel column sums ([4 * splits][cout] -> db), one launch
      const long long total = 9LL * cin * cout;
      int gmain = (int)((total + 255) / 256);
      if (gmain > 148 * 16) gmain = 148 * 16;
      wgrad_reduce_bias_kernel<<<gmain + (cout + 7) / 8, 256, 0, st>>>(q.partial, dw, pl.splits, cin, cout,
                                                                      gmain, bscratch, 4 * pl.splits, db);
      WU_CHECK_LAUNCH("wgrad_reduce_bias_kernel");
    }
    return WU_OK;
  }
  WgradParams p;
  p.c0_blocks = c0 / 64;
  p.ctot_blocks = cin / 64;
  p.atoms = pl.atoms;
  p.n_tiles = pl.n_tiles;
  p.splits = pl.splits;
  p.tiles_w = pl.tiles_w;
  p.tiles_h = pl.tiles_h;
  p.batch = B;
  p.bw = pl.bw;
  p.bh = pl.bh;
  p.pix_tiles = pl.pix_tiles;
  p.cout = cout;
  p.cin_total = cin;
  p.partial = reinterpret_cast<float*>(workspace);
  CUtensorMap x0, x1, ym;
  int rc;
  if ((rc = make_act_tmap(&x0, src0, B, H, W, c0, c0, pl.bw, pl.bh)) != WU_OK) return rc;
  if (c1 > 0) {
    if ((rc = make_act_tmap(&x1, src1, B, H, W, c1, c1, pl.bw, pl.bh)) != WU_OK) return rc;
  } else {
    x1 = x0;
  }
  if ((rc = make_act_tmap(&ym, dy, B, H, W, cout, cout, pl.bw, pl.bh)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = pl.pairs * pl.n_tiles * pl.splits;
  switch (pl.bn) {
    case 64: rc = launch_wgrad<64>(x0, x1, ym, p, grid, st); break;
    case 128: rc = launch_wgrad<128>(x0, x1, ym, p, grid, st); break;
    default: rc = launch_wgrad<256>(x0, x1, ym, p, grid, st); break;
  }
  if (rc != WU_OK) return rc;
  return wgrad_finish(p.partial, dw, db, dy, workspace, pl, cin, cout, B, H, W, st);
}
