// Library bookkeeping: error strings, launch counter, device-architecture gate.
#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstring>

#include "scp_tc.cuh"

namespace scp {

static thread_local char g_detail[512] = "";
static std::atomic<int> g_launches{0};

void set_error_detail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_device_arch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(SCP_ERR_CUDA, "cudaGetDevice failed");
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = SCP_OK;
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return fail(SCP_ERR_CUDA, "cudaDeviceGetAttribute failed");
  cached_dev = dev;
  cached_rc = major == 10 ? SCP_OK : fail(SCP_ERR_ARCH, "device %d has compute capability %d.x; sm_100a required", dev, major);
  return cached_rc;
}

// ---- helper stream (fork / join inside one entry point) ----------------------------------------------------------
namespace {
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false, tried = false;
};
constexpr int kMaxDevices = 64;
SideStream g_side[kMaxDevices];
std::mutex g_side_mutex;

SideStream* side_for_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  SideStream& s = g_side[dev];
  if (!s.tried) {
    s.tried = true;
    s.ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    if (!s.ok) cudaGetLastError();
  }
  return s.ok ? &s : nullptr;
}
}  // namespace

cudaStream_t fork_to_side(cudaStream_t main) {
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream* s = side_for_current_device();
  if (!s) return nullptr;
  if (cudaEventRecord(s->fork, main) != cudaSuccess || cudaStreamWaitEvent(s->stream, s->fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return s->stream;
}

int join_from_side(cudaStream_t main) {
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream* s = side_for_current_device();
  if (!s) return fail(SCP_ERR_CUDA, "helper stream missing at join");
  if (cudaEventRecord(s->join, s->stream) != cudaSuccess || cudaStreamWaitEvent(main, s->join, 0) != cudaSuccess)
    return fail(SCP_ERR_CUDA, "helper stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
  return SCP_OK;
}

// ---- TMA tensor maps ------------------------------------------------------------------------------------------
namespace tc {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(SCP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // cuTensorMapEncodeTiled is a DRIVER call and needs a context bound to the calling thread.  A thread that has only
  // issued cudaSetDevice / cudaGetDevice so far (e.g. PyTorch's autograd worker when one of our backward functions is
  // the first node it runs) has none yet: the call then fails with CUDA_ERROR_INVALID_CONTEXT (201).  cudaFree(0) binds
  // the device's primary context to this thread; done once per thread.
  static thread_local bool context_bound = false;
  if (!context_bound) {
    if (cudaFree(nullptr) != cudaSuccess) return fail(SCP_ERR_CUDA, "could not initialise the CUDA context on this thread");
    context_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0 || box_rows < 1 || box_rows > 256)
    return fail(SCP_ERR_INVALID, "tensor map: base/pitch must be 16-byte aligned, box rows in [1,256]");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SCP_OK;
}
// 3-D view of a row-major fp16 matrix (rows, 64*n_blocks) with row pitch `ld` elements, for MN-major operands:
//   dim0 = 64 contiguous elements of one row, dim1 = row, dim2 = 64-element column block.
// One box {64, box_rows, box_blocks} lands in shared memory as [block][row][64] -- 8 KB (box_rows = 64) SWIZZLE_128B
// tiles one after the other, which is the layout the MN-major UMMA descriptor walks with LBO = box_rows * 128 B.
// Coordinates of a load: {0, first row, first block}.
int make_tmap_f16_blocked(CUtensorMap* out, const void* base, int64_t rows, int64_t n_blocks, int64_t ld, int box_rows,
                          int box_blocks) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(SCP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  static thread_local bool context_bound = false;
  if (!context_bound) {
    if (cudaFree(nullptr) != cudaSuccess) return fail(SCP_ERR_CUDA, "could not initialise the CUDA context on this thread");
    context_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0 || box_rows < 1 || box_rows > 256 || box_blocks < 1 ||
      box_blocks > n_blocks)
    return fail(SCP_ERR_INVALID, "blocked tensor map: bad base / pitch / box");
  const cuuint64_t gdim[3] = {(cuuint64_t)kChunkK, (cuuint64_t)rows, (cuuint64_t)n_blocks};
  const cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)kChunkK * 2};
  const cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows, (cuuint32_t)box_blocks};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCP_ERR_CUDA, "cuTensorMapEncodeTiled (blocked) failed with CUresult %d", (int)r);
  return SCP_OK;
}
}  // namespace tc

}  // namespace scp

extern "C" int scp_version(void) { return 0 * 10000 + 2 * 100 + 0; }

extern "C" int scp_num_launches(void) { return scp::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* scp_last_error_string(int code) {
  static thread_local char buf[640];
  const char* base = "unknown error";
  switch (code) {
    case SCP_OK: base = "ok"; break;
    case SCP_ERR_INVALID: base = "invalid argument"; break;
    case SCP_ERR_UNSUPPORTED: base = "unsupported shape or dtype"; break;
    case SCP_ERR_WORKSPACE: base = "workspace too small"; break;
    case SCP_ERR_CUDA: base = "CUDA error"; break;
    case SCP_ERR_ARCH: base = "wrong GPU architecture (needs sm_100a)"; break;
  }
  if (code != SCP_OK && scp::g_detail[0])
    snprintf(buf, sizeof(buf), "%s: %s", base, scp::g_detail);
  else
    snprintf(buf, sizeof(buf), "%s", base);
  return buf;
}
