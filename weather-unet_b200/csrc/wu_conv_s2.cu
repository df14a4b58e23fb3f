// wu_conv_s2.cu — 3x3 / stride 2 / pad 1 convolution of NHWC bf16 activations on tcgen05:
// forward (+ bias + LeakyReLU), data gradient (a transposed convolution) and weight gradient.
//
// Replaces the second convolution of the reference's discriminator block
//   spectral_norm(nn.Conv2d(cin, cout, 3, padding=1, stride=2)) -> nn.LeakyReLU(0.2)   (nets.py:26-33,
// used by disc.py:12-15,28-31) and its autograd.  SURVEY §8 f1.
//
// A stride-2 window never touches two neighbouring input pixels of the same parity, so the input is
// addressed through its four PARITY VIEWS  x[b][2u+py][2v+px][c]  (py, px in {0,1}): each is an
// ordinary strided 4-D tensor for TMA (pitches 2*C, 2*W*C, H*W*C elements), and every filter tap
// (r, s) becomes a plain box load from ONE view at an offset of -1 or 0 pixels:
//     input row 2*yo + r - 1  ==  view py = (r+1)&1, row yo + (r == 0 ? -1 : 0)        (same for columns)
// TMA zero-fills rows / columns outside a view, which is exactly the convolution's padding.
//   * fprop:  out[b][yo][xo][:] = sum over 9 taps of  view(tap)[yo+dy][xo+dx][:] . W[tap]
//   * dgrad:  the four parity views of dX are four independent small convolutions of dY with
//             1 / 2 / 2 / 4 taps (9 tap-GEMMs in total: no multiplication by inserted zeros);
//             each result tile is TMA-stored straight into its parity view of dX.
//   * wgrad:  D[(tap, ci)][co] = sum_px view(tap)[px + d][ci] . dY[px][co]   (both operands MN-major).
// Tile / pipeline structure as in wu_conv3x3.cu (v1): 128 output pixels x BN channels per tile,
// one TMA producer thread, one MMA issuer thread, four epilogue warps, double-buffered TMEM.
#include <cstdio>

#include "wu_host.h"
#include "wu_ptx.cuh"

namespace wu {

struct TapMaps {
  CUtensorMap a[4];  // sources  (fprop: parity views of x;  dgrad: a[0] = dY)
  CUtensorMap d[4];  // outputs  (fprop: d[0] = out;         dgrad: parity views of dX)
};

struct TapParams {
  int c_blocks;  // 64-channel blocks of the contraction dimension
  int n_tiles;   // N / BN
  int bw, bh, log2_bw;
  int num_tiles;
  int n_classes;
  int cls_end[4];  // cumulative tile count per class (classes ordered heaviest first)
  int cls_tw[4], cls_th[4];
  int cls_ntaps[4];
  int cls_out[4];  // index into TapMaps::d
  signed char tap_map[4][9], tap_dx[4][9], tap_dy[4][9], tap_w[4][9];
  float slope;  // epilogue: v > 0 ? v : v * slope   (1 = identity, 0 = ReLU)
  const float* bias;
};

template <int BN>
struct TapCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 2 * 16384;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 1024;
  static constexpr uint32_t kTmemCols = 2 * BN;
};

struct TapTile {
  int cls, b, h0, w0, n0;
};
template <int BN>
__device__ __forceinline__ TapTile decode_tap_tile(const TapParams& p, int tile) {
  TapTile t;
  int cls = 0, begin = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    if (c + 1 < p.n_classes && tile >= p.cls_end[c]) {
      cls = c + 1;
      begin = p.cls_end[c];
    }
  t.cls = cls;
  int local = tile - begin;
  const int nt = local % p.n_tiles;
  int mt = local / p.n_tiles;
  const int tw = mt % p.cls_tw[cls];
  mt /= p.cls_tw[cls];
  const int th = mt % p.cls_th[cls];
  t.b = mt / p.cls_th[cls];
  t.h0 = th * p.bh;
  t.w0 = tw * p.bw;
  t.n0 = nt * BN;
  return t;
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_taps_kernel(const __grid_constant__ TapMaps maps, const __grid_constant__ CUtensorMap tmB,
                 const TapParams p) {
  using Cfg = TapCfg<BN>;
  constexpr int S = Cfg::kStages;
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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.d[0]);
    tma_prefetch_desc(&tmB);
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

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TapTile t = decode_tap_tile<BN>(p, tile);
        const int ntaps = p.cls_ntaps[t.cls];
        for (int k = 0; k < ntaps; ++k) {
          const CUtensorMap* am = &maps.a[p.tap_map[t.cls][k]];
          const int cw = t.w0 + p.tap_dx[t.cls][k], ch = t.h0 + p.tap_dy[t.cls][k];
          const int wk = p.tap_w[t.cls][k] * p.c_blocks;
          for (int cb = 0; cb < p.c_blocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t fb = full_bar(stage);
            mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
            const uint32_t a_dst = base + stage * Cfg::kStageBytes;
            tma_load_4d(a_dst, am, fb, cb * 64, cw, ch, t.b);
            tma_load_2d(a_dst + Cfg::kABytes, &tmB, fb, (wk + cb) * 64, t.n0);
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const uint64_t adesc0 = umma_smem_desc_sw128(base, 16, 1024);
      const uint64_t bdesc0 = umma_smem_desc_sw128(base + Cfg::kABytes, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const TapTile t = decode_tap_tile<BN>(p, tile);
        const int kblocks = p.cls_ntaps[t.cls] * p.c_blocks;
        const int buf = it & 1;
        mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = (uint64_t)((stage * Cfg::kStageBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16) per 64-channel block
            const uint64_t koff = soff + (uint64_t)((k * 32) >> 4);
            umma_bf16(d_tmem, adesc0 + koff, bdesc0 + koff, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // pixel within the tile == TMEM lane
    const bool issuer = (threadIdx.x == 64);
    const float slope = p.slope;
    const bool act = slope != 1.f;
    uint32_t store_count = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const TapTile t = decode_tap_tile<BN>(p, tile);
      const CUtensorMap* dm = &maps.d[p.cls_out[t.cls]];
      mbar_wait(tfull_bar(buf), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + chunk * 64;
        tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {  // this thread has drained its part of the accumulator
          tc_fence_before();
          mbar_arrive(tempty_bar(buf));
        }
        const int cbase = t.n0 + chunk * 64;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
            v[j + 0] = __float_as_uint(__uint_as_float(v[j + 0]) + bv.x);
            v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + bv.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + bv.z);
            v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + bv.w);
          }
        }
        if (act) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float f = __uint_as_float(v[j]);
            v[j] = __float_as_uint(f > 0.f ? f : f * slope);
          }
        }
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        const uint32_t sb = staging_base + (store_count & 1u) * 16384u;
        ++store_count;
        if (issuer) tma_store_wait_read<1>();  // the store that last used this buffer has read it
        named_bar_sync(1, 128);
        uint8_t* srow = smem + (sb - base) + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(srow + ((j ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          tma_store_4d(dm, sb, cbase, t.w0, t.h0, t.b);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN>
static int launch_taps(const TapMaps& maps, const CUtensorMap& bm, const TapParams& p,
                       cudaStream_t st) {
  using Cfg = TapCfg<BN>;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv_taps_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  conv_taps_kernel<BN><<<grid, 192, Cfg::kSmemBytes, st>>>(maps, bm, p);
  WU_CHECK_LAUNCH("conv_taps_kernel");
  return WU_OK;
}

static int ilog2i(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Parity view (py, px) of an NHWC bf16 tensor [B][H][W][C]: rows py, py+2, ..; columns px, px+2, ..
static int make_parity_tmap(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int py,
                            int px, int bw, int bh) {
  const int Hv = (H - py + 1) / 2, Wv = (W - px + 1) / 2;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(ptr) + ((size_t)py * W + px) * C * 2;
  return make_act_tmap_strided(out, p, B, Hv, Wv, C, 2LL * C * 2, 2LL * W * C * 2,
                               (long long)H * W * C * 2, bw, bh);
}

// ------------------------------------------------------------------------------------------------
// wgrad, stride 2:  D[(tap, ci)][co] = sum_px view(tap)[px + d(tap)][ci] * dY[px][co]
// ------------------------------------------------------------------------------------------------
struct WgradS2Maps {
  CUtensorMap x[4];  // parity views of the input
  CUtensorMap y;     // dY
};
struct WgradS2Params {
  int c_blocks, atoms, n_tiles, splits;
  int tiles_w, tiles_h, bw, bh, pix_tiles;
  int cout, cin;
  float* partial;  // [splits][9*cin][cout]
};

template <int BN>
struct WgradS2Cfg {
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
conv3x3_s2_wgrad_kernel(const __grid_constant__ WgradS2Maps maps, const WgradS2Params p) {
  using Cfg = WgradS2Cfg<BN>;
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

  // blockIdx.x -> (split, atom pair, n tile), split slowest: CTAs walking the same pixel range are
  // neighbours in launch order and share their X / dY tiles through L2
  int id = blockIdx.x;
  const int nt = id % p.n_tiles;
  id /= p.n_tiles;
  const int npairs = (p.atoms + 1) / 2;
  const int pair = id % npairs;
  const int z = id / npairs;
  const int n0 = nt * BN;
  const int pt_begin = (int)(((long long)p.pix_tiles * z) / p.splits);
  const int pt_end = (int)(((long long)p.pix_tiles * (z + 1)) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x[0]);
    tma_prefetch_desc(&maps.y);
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
      int a_map[2], a_dx[2], a_dy[2], a_cb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int a = 2 * pair + j;
        if (a >= p.atoms) a = p.atoms - 1;  // odd atom count: duplicate, result discarded
        const int tap = a / p.c_blocks;
        a_cb[j] = a - tap * p.c_blocks;
        const int r = tap / 3, s = tap - 3 * r;
        a_map[j] = ((r + 1) & 1) * 2 + ((s + 1) & 1);
        a_dy[j] = r == 0 ? -1 : 0;
        a_dx[j] = s == 0 ? -1 : 0;
      }
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
        for (int j = 0; j < 2; ++j)
          tma_load_4d(a_dst + j * Cfg::kAtomBytes, &maps.x[a_map[j]], fb, a_cb[j] * 64, w0 + a_dx[j],
                      h0 + a_dy[j], b);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(b_dst + j * Cfg::kAtomBytes, &maps.y, fb, n0 + j * 64, w0, h0, b);
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);  // both operands MN-major
      const uint64_t adesc0 = umma_smem_desc_sw128(base, Cfg::kAtomBytes, 1024);
      const uint64_t bdesc0 = umma_smem_desc_sw128(base + Cfg::kABytes, Cfg::kAtomBytes, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t soff = (uint64_t)((stage * Cfg::kStageBytes) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 pixels); 16 pixel rows = 2048 bytes
          const uint64_t koff = soff + (uint64_t)((k * 2048) >> 4);
          umma_bf16(tmem_base, adesc0 + koff, bdesc0 + koff, idesc, (pt != pt_begin || k != 0) ? 1u : 0u);
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
    float* out = p.partial + ((size_t)z * 9 * p.cin + (size_t)a * 64 + (row & 63)) * p.cout + n0;
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + chunk * 32, v);
      tmem_ld_wait();
      if (a < p.atoms && pt_end > pt_begin) {
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

template <int BN>
static int launch_wgrad_s2(const WgradS2Maps& maps, const WgradS2Params& p, int grid,
                           cudaStream_t st) {
  using Cfg = WgradS2Cfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    WU_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_s2_wgrad_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  conv3x3_s2_wgrad_kernel<BN><<<grid, 192, Cfg::kSmemBytes, st>>>(maps, p);
  WU_CHECK_LAUNCH("conv3x3_s2_wgrad_kernel");
  return WU_OK;
}

struct WgradS2Plan {
  int bn, n_tiles, atoms, pairs, splits, bw, bh, tiles_w, tiles_h, pix_tiles;
  size_t partial_bytes, bias_bytes;
};
static WgradS2Plan plan_wgrad_s2(int cin, int cout, int B, int Hin, int Win) {
  WgradS2Plan pl;
  const int Ho = (Hin + 1) / 2, Wo = (Win + 1) / 2;
  pl.bn = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  pl.n_tiles = cout / pl.bn;
  pl.atoms = 9 * (cin / 64);
  pl.pairs = (pl.atoms + 1) / 2;
  pick_box(Ho, Wo, 64, &pl.bw, &pl.bh);
  pl.tiles_w = (Wo + pl.bw - 1) / pl.bw;
  pl.tiles_h = (Ho + pl.bh - 1) / pl.bh;
  pl.pix_tiles = B * pl.tiles_w * pl.tiles_h;
  const int tiles = pl.pairs * pl.n_tiles;
  int splits = (2 * num_sms()) / tiles;  // whole waves: at most two
  if (splits > pl.pix_tiles) splits = pl.pix_tiles;
  if (splits < 1) splits = 1;
  pl.splits = splits;
  pl.partial_bytes = (size_t)splits * 9 * cin * cout * sizeof(float);
  pl.bias_bytes = (size_t)kBiasGradBlocks * cout * sizeof(float);
  return pl;
}

// ---- data gradient of the 3 -> 64 stride-2 stem convolution: GEMM first, shift afterwards ----------
// With only three input channels the usual "shift the input, then contract" order wastes the tensor
// core (N = 3) and re-reads g once per tap.  Here g is contracted ONCE with all 27 (ci, r, s) weight
// columns, T[b,u,v,(ci,r,s)] = sum_co g[b,u,v,co] * w[co][ci][r][s] (a 1x1 convolution, N = 64 with 27
// live columns, one pass over g), and a small second kernel scatters T to the output positions
// (2u + r - 1, 2v + s - 1): every 2x2 block of output pixels reads the four T pixels around it and uses
// each of their 27 values exactly once.
__global__ void pack_w3to64_t_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // n * 64 + co
  if (i >= 64 * 64) return;
  const int n = i >> 6, co = i & 63;
  wt[i] = __float2bfloat16_rn(n < 27 ? w[co * 27 + n] : 0.f);
}

__global__ void __launch_bounds__(256)
col2im_s2_kernel(const __nv_bfloat16* __restrict__ T, float* __restrict__ out, int B, int Hin, int Win,
                 int Ho, int Wo) {
  const long long nblk = (long long)B * Ho * Wo;
  const long long HW = (long long)Hin * Win;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nblk;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % Wo);
    const long long t = i / Wo;
    const int Y = (int)(t % Ho);
    const long long b = t / Ho;
    float acc[2][2][3];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) acc[dy][dx][ci] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (Y + a >= Ho || X + c >= Wo) continue;
        const uint4* tp = reinterpret_cast<const uint4*>(T + (((b * Ho + Y + a) * Wo) + X + c) * 64);
        uint32_t wds[16];  // 32 bf16: the 27 live values of this T pixel
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 v = __ldg(tp + q);
          wds[4 * q] = v.x; wds[4 * q + 1] = v.y; wds[4 * q + 2] = v.z; wds[4 * q + 3] = v.w;
        }
        // tap r lands on output row parity dy = (r + 1) & 1 from T row a = (dy + 1 - r) / 2
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int dy = (r + 1) & 1;
          if (((dy + 1 - r) >> 1) != a) continue;
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3) {
            const int dx = (s3 + 1) & 1;
            if (((dx + 1 - s3) >> 1) != c) continue;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              const int n = ci * 9 + r * 3 + s3;
              const uint32_t wd = wds[n >> 1];
              acc[dy][dx][ci] += (n & 1) ? bf16hi(wd) : bf16lo(wd);
            }
          }
        }
      }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float* o = out + (b * 3 + ci) * HW + (long long)(2 * Y) * Win + 2 * X;
      *reinterpret_cast<float2*>(o) = make_float2(acc[0][0][ci], acc[0][1][ci]);
      if (2 * Y + 1 < Hin) *reinterpret_cast<float2*>(o + Win) = make_float2(acc[1][0][ci], acc[1][1][ci]);
    }
  }
}

}  // namespace wu

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace wu;

static bool s2_channels_ok(int c) { return c > 0 && c % 64 == 0 && (c <= 256 || c % 256 == 0) && c != 192; }

extern "C" int wu_conv3x3_s2_fprop(const void* src, int cin, const void* w_packed, const float* bias,
                                   float slope, void* dst, int cout, int B, int Hin, int Win,
                                   wu_stream_t stream) {
  WU_REQUIRE(src && w_packed && dst, "wu_conv3x3_s2_fprop: null pointer");
  WU_REQUIRE(B > 0 && Hin >= 2 && Win >= 2, "wu_conv3x3_s2_fprop: bad shape B=%d H=%d W=%d", B, Hin, Win);
  WU_REQUIRE(cin > 0 && cin % 64 == 0, "wu_conv3x3_s2_fprop: cin=%d must be a positive multiple of 64", cin);
  WU_REQUIRE(s2_channels_ok(cout), "wu_conv3x3_s2_fprop: cout=%d must be 64, 128 or a multiple of 256", cout);
  const int Ho = (Hin + 1) / 2, Wo = (Win + 1) / 2;
  const int bn = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  TapParams p{};
  pick_box(Ho, Wo, 128, &p.bw, &p.bh);
  p.log2_bw = ilog2i(p.bw);
  p.c_blocks = cin / 64;
  p.n_tiles = cout / bn;
  p.n_classes = 1;
  p.cls_tw[0] = (Wo + p.bw - 1) / p.bw;
  p.cls_th[0] = (Ho + p.bh - 1) / p.bh;
  const long long nt = (long long)B * p.cls_tw[0] * p.cls_th[0] * p.n_tiles;
  WU_REQUIRE(nt < (1LL << 31), "wu_conv3x3_s2_fprop: too many tiles");
  p.num_tiles = (int)nt;
  p.cls_end[0] = p.num_tiles;
  p.cls_ntaps[0] = 9;
  p.cls_out[0] = 0;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int k = r * 3 + s;
      p.tap_map[0][k] = (signed char)(((r + 1) & 1) * 2 + ((s + 1) & 1));
      p.tap_dy[0][k] = (signed char)(r == 0 ? -1 : 0);
      p.tap_dx[0][k] = (signed char)(s == 0 ? -1 : 0);
      p.tap_w[0][k] = (signed char)k;
    }
  p.slope = slope;
  p.bias = bias;
  TapMaps maps;
  CUtensorMap bm;
  int rc;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      if ((rc = make_parity_tmap(&maps.a[py * 2 + px], src, B, Hin, Win, cin, py, px, p.bw, p.bh)) != WU_OK)
        return rc;
  if ((rc = make_act_tmap(&maps.d[0], dst, B, Ho, Wo, cout, cout, p.bw, p.bh)) != WU_OK) return rc;
  for (int i = 1; i < 4; ++i) maps.d[i] = maps.d[0];
  if ((rc = make_mat_tmap(&bm, w_packed, cout, 9 * cin, bn)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 64: return launch_taps<64>(maps, bm, p, st);
    case 128: return launch_taps<128>(maps, bm, p, st);
    default: return launch_taps<256>(maps, bm, p, st);
  }
}

extern "C" int wu_conv3x3_s2_dgrad(const void* dy, int cout, const void* w_dgrad, void* dx, int cin,
                                   int B, int Hin, int Win, wu_stream_t stream) {
  WU_REQUIRE(dy && w_dgrad && dx, "wu_conv3x3_s2_dgrad: null pointer");
  WU_REQUIRE(B > 0 && Hin >= 2 && Win >= 2, "wu_conv3x3_s2_dgrad: bad shape B=%d H=%d W=%d", B, Hin, Win);
  WU_REQUIRE(cout > 0 && cout % 64 == 0, "wu_conv3x3_s2_dgrad: cout=%d must be a positive multiple of 64", cout);
  WU_REQUIRE(s2_channels_ok(cin), "wu_conv3x3_s2_dgrad: cin=%d must be 64, 128 or a multiple of 256", cin);
  const int Ho = (Hin + 1) / 2, Wo = (Win + 1) / 2;
  const int bn = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
  TapParams p{};
  pick_box(Ho, Wo, 128, &p.bw, &p.bh);
  p.log2_bw = ilog2i(p.bw);
  p.c_blocks = cout / 64;
  p.n_tiles = cin / bn;
  p.n_classes = 4;
  // classes heaviest first: (py, px) = (1,1): 4 taps, (1,0) and (0,1): 2 taps, (0,0): 1 tap
  const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};
  long long total = 0;
  for (int c = 0; c < 4; ++c) {
    const int py = order[c][0], px = order[c][1];
    const int Hv = (Hin - py + 1) / 2, Wv = (Win - px + 1) / 2;
    p.cls_tw[c] = (Wv + p.bw - 1) / p.bw;
    p.cls_th[c] = (Hv + p.bh - 1) / p.bh;
    total += (long long)B * p.cls_tw[c] * p.cls_th[c] * p.n_tiles;
    WU_REQUIRE(total < (1LL << 31), "wu_conv3x3_s2_dgrad: too many tiles");
    p.cls_end[c] = (int)total;
    p.cls_out[c] = py * 2 + px;
    // dX[2u+py][2v+px] += dY[(2u+py+1-r)/2][(2v+px+1-s)/2] . W[r][s]   for r = py+1 (mod 2), s = px+1 (mod 2)
    int n = 0;
    for (int r = 0; r < 3; ++r) {
      if (((py + 1 - r) & 1) != 0) continue;
      for (int s = 0; s < 3; ++s) {
        if (((px + 1 - s) & 1) != 0) continue;
        p.tap_map[c][n] = 0;
        p.tap_dy[c][n] = (signed char)((py + 1 - r) / 2);  // 0 or +1 (r = 0 with py = 1)
        p.tap_dx[c][n] = (signed char)((px + 1 - s) / 2);
        p.tap_w[c][n] = (signed char)(8 - (r * 3 + s));    // column block of w_dgrad: k = (8-tap)*cout + co
        ++n;
      }
    }
    p.cls_ntaps[c] = n;
  }
  p.num_tiles = (int)total;
  p.slope = 1.f;
  p.bias = nullptr;
  TapMaps maps;
  CUtensorMap bm;
  int rc;
  if ((rc = make_act_tmap(&maps.a[0], dy, B, Ho, Wo, cout, cout, p.bw, p.bh)) != WU_OK) return rc;
  for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      if ((rc = make_parity_tmap(&maps.d[py * 2 + px], dx, B, Hin, Win, cin, py, px, p.bw, p.bh)) != WU_OK)
        return rc;
  if ((rc = make_mat_tmap(&bm, w_dgrad, cin, 9 * cout, bn)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 64: return launch_taps<64>(maps, bm, p, st);
    case 128: return launch_taps<128>(maps, bm, p, st);
    default: return launch_taps<256>(maps, bm, p, st);
  }
}

extern "C" size_t wu_conv3x3_s2_wgrad_workspace_bytes(int cin, int cout, int B, int Hin, int Win) {
  if (cin <= 0 || cout <= 0 || B <= 0 || Hin <= 0 || Win <= 0) return 0;
  const WgradS2Plan pl = plan_wgrad_s2(cin, cout, B, Hin, Win);
  return pl.partial_bytes + pl.bias_bytes + 256;
}

extern "C" int wu_conv3x3_s2_wgrad(const void* src, int cin, const void* dy, int cout, int B, int Hin,
                                   int Win, float* dw, float* db, void* workspace,
                                   size_t workspace_bytes, wu_stream_t stream) {
  WU_REQUIRE(src && dy && dw && workspace, "wu_conv3x3_s2_wgrad: null pointer");
  WU_REQUIRE(B > 0 && Hin >= 2 && Win >= 2, "wu_conv3x3_s2_wgrad: bad shape B=%d H=%d W=%d", B, Hin, Win);
  WU_REQUIRE(cin > 0 && cin % 64 == 0, "wu_conv3x3_s2_wgrad: cin=%d must be a positive multiple of 64", cin);
  WU_REQUIRE(s2_channels_ok(cout), "wu_conv3x3_s2_wgrad: unsupported cout=%d", cout);
  const WgradS2Plan pl = plan_wgrad_s2(cin, cout, B, Hin, Win);
  WU_REQUIRE(workspace_bytes >= pl.partial_bytes + pl.bias_bytes,
             "wu_conv3x3_s2_wgrad: workspace %zu < required %zu", workspace_bytes,
             pl.partial_bytes + pl.bias_bytes);
  WU_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "wu_conv3x3_s2_wgrad: workspace unaligned");
  const int Ho = (Hin + 1) / 2, Wo = (Win + 1) / 2;
  WgradS2Params p;
  p.c_blocks = cin / 64;
  p.atoms = pl.atoms;
  p.n_tiles = pl.n_tiles;
  p.splits = pl.splits;
  p.tiles_w = pl.tiles_w;
  p.tiles_h = pl.tiles_h;
  p.bw = pl.bw;
  p.bh = pl.bh;
  p.pix_tiles = pl.pix_tiles;
  p.cout = cout;
  p.cin = cin;
  p.partial = reinterpret_cast<float*>(workspace);
  WgradS2Maps maps;
  int rc;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      if ((rc = make_parity_tmap(&maps.x[py * 2 + px], src, B, Hin, Win, cin, py, px, pl.bw, pl.bh)) != WU_OK)
        return rc;
  if ((rc = make_act_tmap(&maps.y, dy, B, Ho, Wo, cout, cout, pl.bw, pl.bh)) != WU_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = pl.pairs * pl.n_tiles * pl.splits;
  switch (pl.bn) {
    case 64: rc = launch_wgrad_s2<64>(maps, p, grid, st); break;
    case 128: rc = launch_wgrad_s2<128>(maps, p, grid, st); break;
    default: rc = launch_wgrad_s2<256>(maps, p, grid, st); break;
  }
  if (rc != WU_OK) return rc;
  float* bpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + pl.partial_bytes);
  return wgrad_fold(p.partial, pl.splits, cin, cout, dw, dy, (long long)B * Ho * Wo, db, bpart, st);
}

// Data gradient of the discriminator stem's Conv2d(3, 64, 3, padding=1, stride=2) (nets.py:30-31 with
// in_channels = 3): g NHWC bf16 [B][Ho][Wo][64] -> g_h1 fp32 NCHW [B][3][Hin][Win] (Hin, Win even).
// workspace: packed weights (8 KiB) + T bf16 [B][Ho][Wo][64].
extern "C" size_t wu_conv3to64_s2_dgrad_workspace_bytes(int B, int Hin, int Win) {
  if (B <= 0 || Hin <= 0 || Win <= 0) return 0;
  return 8192 + (size_t)B * ((Hin + 1) / 2) * ((Win + 1) / 2) * 128;
}

extern "C" int wu_conv3to64_s2_dgrad(const void* g, const float* w, float* g_h1, int B, int Hin,
                                     int Win, void* workspace, size_t workspace_bytes,
                                     wu_stream_t stream) {
  WU_REQUIRE(g && w && g_h1 && workspace, "wu_conv3to64_s2_dgrad: null pointer");
  WU_REQUIRE(B > 0 && Hin >= 2 && Win >= 2 && Hin % 2 == 0 && Win % 2 == 0,
             "wu_conv3to64_s2_dgrad: bad shape B=%d H=%d W=%d (H, W must be even)", B, Hin, Win);
  WU_REQUIRE(workspace_bytes >= wu_conv3to64_s2_dgrad_workspace_bytes(B, Hin, Win) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 127) == 0,
             "wu_conv3to64_s2_dgrad: workspace too small or not 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + 8192);
  pack_w3to64_t_kernel<<<16, 256, 0, st>>>(w, wt);
  WU_CHECK_LAUNCH("pack_w3to64_t_kernel");
  const int Ho = Hin / 2, Wo = Win / 2;
  TapParams p{};
  pick_box(Ho, Wo, 128, &p.bw, &p.bh);
  p.log2_bw = ilog2i(p.bw);
  p.c_blocks = 1;
  p.n_tiles = 1;
  p.n_classes = 1;
  p.cls_tw[0] = (Wo + p.bw - 1) / p.bw;
  p.cls_th[0] = (Ho + p.bh - 1) / p.bh;
  const long long nt = (long long)B * p.cls_tw[0] * p.cls_th[0];
  WU_REQUIRE(nt < (1LL << 31), "wu_conv3to64_s2_dgrad: too many tiles");
  p.num_tiles = (int)nt;
  p.cls_end[0] = p.num_tiles;
  p.cls_ntaps[0] = 1;  // a 1x1 convolution: one tap, no shift
  p.cls_out[0] = 0;
  p.slope = 1.f;
  p.bias = nullptr;
  TapMaps maps;
  CUtensorMap bm;
  int rc;
  if ((rc = make_act_tmap(&maps.a[0], g, B, Ho, Wo, 64, 64, p.bw, p.bh)) != WU_OK) return rc;
  if ((rc = make_act_tmap(&maps.d[0], T, B, Ho, Wo, 64, 64, p.bw, p.bh)) != WU_OK) return rc;
  for (int i = 1; i < 4; ++i) {
    maps.a[i] = maps.a[0];
    maps.d[i] = maps.d[0];
  }
  if ((rc = make_mat_tmap(&bm, wt, 64, 64, 64)) != WU_OK) return rc;
  if ((rc = launch_taps<64>(maps, bm, p, st)) != WU_OK) return rc;
  const long long nblk = (long long)B * Ho * Wo;
  long long grid = (nblk + 255) / 256;
  if (grid > 148LL * 32) grid = 148LL * 32;
  col2im_s2_kernel<<<(int)grid, 256, 0, st>>>(T, g_h1, B, Hin, Win, Ho, Wo);
  WU_CHECK_LAUNCH("col2im_s2_kernel");
  return WU_OK;
}
