// Library-wide state of libicaf_b200: version, thread-local error string, device properties.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "icaf_internal.cuh"

namespace icaf {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return ICAF_ERR_CUDA;
}
int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, where);
  return ICAF_OK;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ICAF_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}
int sm_count_cached() {
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  if (sms[dev] <= 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 256-byte L2 promotion pulls the neighbouring 128-byte line along with every miss: right for dense rows (the neighbour
// is the next K block or the next pixel of the same box), wrong for a channel slice of a wider buffer whose last line
// is followed by channels this tensor does not own (ncu: 2x DRAM reads on the 64-of-128-channel C3 inputs).
static CUtensorMapL2promotion l2_promotion(uint64_t row_bytes, uint64_t pitch_bytes) {
  return (pitch_bytes > row_bytes && (row_bytes % 256) != 0) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                   uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(ICAF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  // staged rows are box_inner*2 bytes wide; the swizzle span equals the row (32 / 64 / 128 B)
  const CUtensorMapSwizzle swz = box_inner >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2_promotion(inner * 2, row_pitch_bytes),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(2d inner=%llu rows=%llu pitch=%llu box=%ux%u) failed: %d",
             (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_inner, box_rows, int(r));
    return set_error(ICAF_ERR_CUDA, msg);
  }
  return ICAF_OK;
}

int encode_tmap_nhwc(CUtensorMap* out, const void* base, int C, int W, int H, int B, int64_t ld, uint32_t box_c,
                     uint32_t box_w, uint32_t box_h, uint32_t sw, uint32_t sh) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(ICAF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(B)};
  cuuint64_t strides[3] = {cuuint64_t(ld) * 2, cuuint64_t(ld) * 2 * W, cuuint64_t(ld) * 2 * W * H};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, sw, sh, 1};
  // rows of the staged tile are box_c*2 bytes wide; the swizzle span equals the row (32 / 64 / 128 B)
  const CUtensorMapSwizzle swz = box_c >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2_promotion(uint64_t(C) * 2, uint64_t(ld) * 2),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(nhwc C=%d W=%d H=%d B=%d ld=%lld box=%u,%u,%u stride=%u,%u) failed: %d", C, W, H, B,
             (long long)ld, box_c, box_w, box_h, sw, sh, int(r));
    return set_error(ICAF_ERR_CUDA, msg);
  }
  return ICAF_OK;
}

}  // namespace icaf

extern "C" int icaf_version(void) { return 100; }   // 0.1.0
extern "C" const char* icaf_last_error(void) { return icaf::g_err; }
extern "C" long long icaf_kernel_launches(void) { return icaf::launches_so_far(); }
extern "C" int icaf_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}

// Device-side offset added to every dropout seed (read by the kernels at run time): a captured CUDA graph of the training step
// keeps its host-side seeds, so the caller bumps this counter on the device between replays to draw fresh masks.
static const uint32_t* g_seed_offset = nullptr;
namespace icaf { const uint32_t* seed_offset_ptr() { return g_seed_offset; } }
extern "C" int icaf_set_seed_offset(const void* device_u32) {
  g_seed_offset = (const uint32_t*)device_u32;
  return ICAF_OK;
}

