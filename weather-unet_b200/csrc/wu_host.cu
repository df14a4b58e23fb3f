// wu_host.cu — error slot, device query, TMA tensor-map construction (driver entry point fetched
// at run time so the library links against cudart only).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "wu_host.h"

namespace wu {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_act_tmap(CUtensorMap* out, const void* ptr, int B, int H, int W, int c, int ctot, int bw,
                  int bh) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(WU_ERR_INVALID, "activation pointer %p not 16-byte aligned", ptr);
  if (c % 64 != 0 || ctot % 8 != 0)
    return fail(WU_ERR_INVALID, "activation view needs c %% 64 == 0 (got %d) and pitch %% 8 == 0 (%d)",
                c, ctot);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ctot * 2, (cuuint64_t)W * ctot * 2,
                           (cuuint64_t)H * W * ctot * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled(act B=%d H=%d W=%d c=%d pitch=%d box=%dx%d) -> %d",
                B, H, W, c, ctot, bw, bh, (int)r);
  return WU_OK;
}

int make_act_tmap_blocks(CUtensorMap* out, const void* ptr, int B, int H, int W, int c, int ctot, int bw,
                         int bh, int nblk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(WU_ERR_INVALID, "activation pointer %p not 16-byte aligned", ptr);
  if (c % 64 != 0 || ctot % 8 != 0 || nblk < 1 || nblk > c / 64)
    return fail(WU_ERR_INVALID, "blocked activation view: c=%d pitch=%d nblk=%d", c, ctot, nblk);
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(c / 64), (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)ctot * 2, (cuuint64_t)W * ctot * 2, 128,
                           (cuuint64_t)H * W * ctot * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)nblk, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled(act5 B=%d H=%d W=%d c=%d pitch=%d box=%dx%dx%d) -> %d",
                B, H, W, c, ctot, bw, bh, nblk, (int)r);
  return WU_OK;
}

int make_act_tmap_strided(CUtensorMap* out, const void* ptr, int B, int Hv, int Wv, int c,
                          long long pitch_w_bytes, long long pitch_h_bytes, long long pitch_b_bytes,
                          int bw, int bh) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(WU_ERR_INVALID, "activation pointer %p not 16-byte aligned", ptr);
  if (c % 64 != 0 || pitch_w_bytes % 16 != 0 || pitch_h_bytes % 16 != 0 || pitch_b_bytes % 16 != 0)
    return fail(WU_ERR_INVALID, "strided activation view needs c %% 64 == 0 (got %d) and 16-byte pitches", c);
  if (B <= 0 || Hv <= 0 || Wv <= 0)
    return fail(WU_ERR_INVALID, "strided activation view is empty (B=%d Hv=%d Wv=%d)", B, Hv, Wv);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)pitch_w_bytes, (cuuint64_t)pitch_h_bytes,
                           (cuuint64_t)pitch_b_bytes};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled(strided act B=%d Hv=%d Wv=%d c=%d box=%dx%d) -> %d",
                B, Hv, Wv, c, bw, bh, (int)r);
  return WU_OK;
}

int make_image_tmap(CUtensorMap* out, const float* ptr, int B, int C, int H, int W, int bx, int by) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || W % 4 != 0 || bx % 4 != 0 || bx > 256 || by > 256)
    return fail(WU_ERR_INVALID, "image map needs a 16-byte aligned pointer, W %% 4 == 0 and a box <= 256 (W=%d box=%dx%d)",
                W, bx, by);
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)C, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled(image B=%d C=%d H=%d W=%d box=%dx%d) -> %d", B, C, H,
                W, bx, by, (int)r);
  return WU_OK;
}

int make_mat_tmap(CUtensorMap* out, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(WU_ERR_INVALID, "matrix pointer %p not 16-byte aligned", ptr);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WU_ERR_CUDA, "cuTensorMapEncodeTiled(mat %dx%d box_rows=%d) -> %d", rows, cols,
                box_rows, (int)r);
  return WU_OK;
}

void pick_box(int H, int W, int npix, int* bw_out, int* bh_out) {
  long best = -1;
  int best_bw = npix, best_bh = 1;
  for (int bw = npix; bw >= 8; bw >>= 1) {
    int bh = npix / bw;
    if (bw > 256 || bh > 256) continue;
    long padded = (long)((W + bw - 1) / bw) * bw * (long)((H + bh - 1) / bh) * bh;
    if (best < 0 || padded < best) {
      best = padded;
      best_bw = bw;
      best_bh = bh;
    }
  }
  *bw_out = best_bw;
  *bh_out = best_bh;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace wu

extern "C" {

const char* wu_last_error(void) { return wu::g_err; }
int wu_version(void) { return 100; }
unsigned long long wu_launch_count(void) { return wu::g_launches.load(); }

int wu_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return wu::fail(WU_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess)
    return wu::fail(WU_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10)
    return wu::fail(WU_ERR_UNSUPPORTED, "device %d has compute capability %d.x; sm_100a required",
                    dev, major);
  return WU_OK;
}

}  // extern "C"
