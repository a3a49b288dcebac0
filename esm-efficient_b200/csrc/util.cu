// Host utilities: thread-local error string, TMA tensor-map encoding through the
// driver entry point (no link-time libcuda dependency), device properties.
#include <stdlib.h>

#include "common.cuh"
#include "esmk_internal.h"

#include <atomic>
#include <mutex>

namespace esmk {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

int fail(const char* where, const std::string& msg) {
  g_last_error = std::string(where) + ": " + msg;
  return 1;
}

int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  g_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return 2;
}

// ---------------------------------------------------------------------------
// Asynchronous (device-detected) errors: one sticky word in mapped pinned host memory (same address on the host and on
// every device under UVA).  Kernels that meet invalid DATA (a token id outside the embedding table, a position
// beyond the learned positional table) OR a code into it -- a zero-copy store that only happens on the error path --
// and every C-ABI entry point checks it before doing anything else: the failure surfaces at the next API call
// after the offending kernel ran (the reference would hit a device-side assert at its next synchronisation).
// ---------------------------------------------------------------------------
uint32_t* async_error_word() {
  static uint32_t* word = [] {
    void* p = nullptr;
    if (cudaHostAlloc(&p, sizeof(uint32_t), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return static_cast<uint32_t*>(nullptr);
    }
    *static_cast<volatile uint32_t*>(p) = 0;
    return static_cast<uint32_t*>(p);
  }();
  return word;
}

int consume_async_error() {
  static uint32_t* word = nullptr;
  if (word == nullptr) {
    word = async_error_word();
    if (word == nullptr) return 0;
  }
  const uint32_t code = *reinterpret_cast<volatile uint32_t*>(word);
  if (code == 0) return 0;
  *reinterpret_cast<volatile uint32_t*>(word) = 0;
  std::string msg = "a previous kernel reported:";
  if (code & ESMK_ASYNC_BAD_TOKEN) msg += " token id outside [0, embedding rows) in esmk_embed / esmk_forward;";
  if (code & ESMK_ASYNC_BAD_POSITION) msg += " sequence longer than the learned positional table in esmk_add_positions;";
  if (code & ESMK_ASYNC_PEER_TIMEOUT) msg += " a rank did not arrive within 30 s in esmk_peer_allgather_logits;";
  msg += " the outputs of that call are invalid";
  return fail("esmk (asynchronous)", msg);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  ESMK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  ESMK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  ESMK_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  ESMK_REQUIRE(box_rows <= 256 && box_cols <= 256, "TMA box dims must be <= 256");
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else ESMK_REQUIRE(swizzle_bytes == 0, "bad swizzle");
  if (swizzle_bytes) ESMK_REQUIRE(box_cols * 2 == (uint32_t)swizzle_bytes, "box inner bytes must equal swizzle span");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled", "CUresult " + std::to_string((int)r));
  return 0;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("ESMK_PDL"); return e == nullptr || e[0] != '0'; }();
  return on;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev;
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  if (dev >= kMaxDevices) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncSetAttribute is per device (context): `mask` remembers on which devices a kernel was configured
bool needs_config(std::atomic<uint64_t>& mask) {
  const int dev = current_device();
  if (dev >= kMaxDevices) return true;                       // beyond the bitmask: configure every time
  return (mask.load(std::memory_order_acquire) & (1ull << dev)) == 0;
}
void mark_configured(std::atomic<uint64_t>& mask) {
  const int dev = current_device();
  if (dev < kMaxDevices) mask.fetch_or(1ull << dev, std::memory_order_release);
}

}  // namespace esmk
