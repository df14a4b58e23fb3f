// wu_host.h — host-side helpers shared by the C-ABI translation units (error slot, TMA maps).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/wu_b200.h"

namespace wu {

// Thread-local last-error slot behind wu_last_error(); returns `code` for `return fail(...)`.
int fail(int code, const char* fmt, ...);

#define WU_CHECK_CUDA(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return ::wu::fail(WU_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
                        cudaGetErrorString(_e));                                       \
  } while (0)

#define WU_REQUIRE(cond, ...)                                         \
  do {                                                                \
    if (!(cond)) return ::wu::fail(WU_ERR_INVALID, __VA_ARGS__);      \
  } while (0)

// NHWC bf16 activation view: `ptr` addresses channel 0 of the view, `ctot` is the pixel pitch in
// channels of the underlying tensor, `c` the number of channels the view exposes.
// Builds a 4-D (C, W, H, B) tiled map with box (64, bw, bh, 1) and 128-byte swizzle.
int make_act_tmap(CUtensorMap* out, const void* ptr, int B, int H, int W, int c, int ctot, int bw,
                  int bh);
// The same view with the 64-channel block index as a dimension of its own: 5-D (64, W, H, c/64, B), box
// (64, bw, bh, nblk, 1).  ONE TMA instruction then fetches the tiles of nblk consecutive channel blocks,
// which land back to back in shared memory (bw * bh * 128 bytes each) — the layout several 4-D loads
// would have produced.  Coordinates: (0, w, h, first block, b).
int make_act_tmap_blocks(CUtensorMap* out, const void* ptr, int B, int H, int W, int c, int ctot, int bw,
                         int bh, int nblk);
// Row-major bf16 matrix [rows][cols] (cols contiguous), box (64, box_rows), 128-byte swizzle.
int make_mat_tmap(CUtensorMap* out, const void* ptr, int rows, int cols, int box_rows);

// Pick a (bw, bh) pixel box with bw*bh == npix minimising padded work for an H x W image.
void pick_box(int H, int W, int npix, int* bw, int* bh);

int num_sms();

// Strided NHWC bf16 view (e.g. the even/odd row/column "parity" views a stride-2 convolution walks):
// dims (c, Wv, Hv, B) with byte pitches between consecutive view columns / rows / images; `ptr`
// addresses element (0, 0, 0, 0) of the view.  Box (64, bw, bh, 1), 128-byte swizzle.
int make_act_tmap_strided(CUtensorMap* out, const void* ptr, int B, int Hv, int Wv, int c,
                          long long pitch_w_bytes, long long pitch_h_bytes, long long pitch_b_bytes,
                          int bw, int bh);

// fp32 NCHW image [B][C][H][W] as a 4-D map (W, H, C, B), box (bx, by, C, 1), no swizzle, zero fill
// outside the image.  Needs W % 4 == 0 (16-byte row pitch) and a 16-byte aligned pointer.
int make_image_tmap(CUtensorMap* out, const float* ptr, int B, int C, int H, int W, int bx, int by);

// Split-K fold of weight-gradient partials [splits][9*cin][cout] into dw [cout][cin][3][3] and,
// when db != NULL, the bias gradient db[c] = sum_px dy[px][c] (dy bf16 [npix][cout]) through
// `bias_scratch` (kBiasGradBlocks * cout floats).  Defined in wu_conv3x3.cu.
constexpr int kBiasGradBlocks = 148 * 4;
int wgrad_fold(const float* partial, int splits, int cin, int cout, float* dw, const void* dy,
               long long npix, float* db, float* bias_scratch, cudaStream_t st);

// Conv2d(3, 64, 3, padding=1, stride) + bias + LeakyReLU(slope) of an fp32 NCHW image into NHWC bf16
// on tcgen05 (TF32 im2col rows written by producer threads).  Defined in wu_conv_first_tc.cu.
int conv_k27_fprop_tc(const float* x, const float* w, const float* bias, float slope, void* dst, int B,
                      int Hin, int Win, int stride, cudaStream_t st);

// Weight / bias gradient of the same layers on tcgen05 (image row pitch must be a multiple of 16 bytes:
// conv_k27_wgrad_tc_supported); workspace = conv_k27_wgrad_tc_workspace_bytes().
bool conv_k27_wgrad_tc_supported(const float* x, int Win);
size_t conv_k27_wgrad_tc_workspace_bytes();
int conv_k27_wgrad_tc(const float* x, const void* dy, float* dw, float* db, int B, int Hin, int Win,
                      int stride, void* workspace, cudaStream_t st);

// Launch accounting behind wu_launch_count().
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// Error check after a kernel launch (launch-configuration errors only; no sync).
#define WU_CHECK_LAUNCH(what)                                                             \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::wu::fail(WU_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e)); \
    ::wu::count_launch();                                                                 \
  } while (0)

}  // namespace wu
