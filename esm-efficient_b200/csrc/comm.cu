// The one collective of the path behind the C ABI, two implementations:
//   esmk_allgather_logits      an NCCL all-gather of the per-rank, padded logits followed by a row gather that restores
//                              the original packed order on every rank;
//   esmk_peer_allgather_logits the NVLink / NVSwitch peer path: every rank owns a window (cudaMalloc + CUDA IPC) that
//                              all ranks map; ONE kernel stores the rank's logits rows straight into their final packed
//                              position in every rank's window (order restore and transfer in the same pass, no padded
//                              intermediate, no NCCL on the data path) and raises a per-rank flag with a system-scope
//                              release; a one-block kernel waits for all flags.  NCCL is used once, at set-up, to
//                              exchange the 64-byte IPC handles.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2: the copy PyTorch has already loaded in a torch process, the
// system library for a plain C caller), so libesmk.so has no link-time dependency on it and single-GPU users never
// touch it.  The handful of NCCL declarations used are restated here (nccl.h: ncclUniqueId = 128 opaque bytes,
// ncclBfloat16 = 9, ncclSuccess = 0).
#include <dlfcn.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "esmk_internal.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclBfloat16 = 9;

struct Nccl {
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllGatherFn all_gather = nullptr;
  GetErrorStringFn error_string = nullptr;
  std::string why;
  bool ok = false;
};

const Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    const char* override_path = getenv("ESMK_NCCL_LIB");
    for (const char* name : {override_path, "libnccl.so.2", "libnccl.so"}) {
      if (name == nullptr) continue;
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h != nullptr) break;
    }
    if (h == nullptr) {
      n.why = std::string("libnccl.so.2 could not be loaded (") + (dlerror() ? dlerror() : "?") + ")";
      return;
    }
    n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
    n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
    n.all_gather = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
    n.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_gather;
    if (!n.ok) n.why = "libnccl.so.2 lacks an expected symbol";
  });
  return n;
}

int nccl_fail(const char* what, int rc) {
  const Nccl& n = nccl();
  return esmk::fail(what, std::string("NCCL error ") + std::to_string(rc) + (n.error_string ? std::string(": ") + n.error_string(rc) : ""));
}

// out[t, :] = gathered[perm[t], :]: one warp per output row, 16-byte chunks when the row allows, else 2-byte elements
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int64_t* __restrict__ perm,
                                   __nv_bfloat16* __restrict__ dst, int T, int V) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const __nv_bfloat16* s = src + (size_t)perm[t] * V;
  __nv_bfloat16* d = dst + (size_t)t * V;
  if ((V & 7) == 0) {
    for (int c = lane; c < (V >> 3); c += 32) reinterpret_cast<uint4*>(d)[c] = reinterpret_cast<const uint4*>(s)[c];
  } else {
    for (int c = lane; c < V; c += 32) d[c] = s[c];
  }
}

constexpr size_t kWindowHeader = 1024;      // uint32 arrival flags[world] at the start of every window
constexpr int kMaxPeerWorld = 64;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Every rank stores its rows at their packed position into the window of EVERY rank (its own included), then the
// last block to finish publishes `epoch` in slot `rank` of every window's flag array.
//   window layout: [flags | buffer 0 | buffer 1], row t of the packed batch at buf_off + t * V * 2
template <bool VEC>
__global__ void __launch_bounds__(256)
peer_store_rows_kernel(const __nv_bfloat16* __restrict__ local, const int32_t* __restrict__ dest_rows,
                       uint8_t* const* __restrict__ windows, int world, int rank, int rows, int V, size_t buf_off,
                       unsigned* __restrict__ counter, uint32_t epoch) {
  const int per_row = VEC ? (V >> 3) : V;                       // 16-byte chunks or 2-byte elements per row
  const long n = (long)rows * per_row;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per_row), c = (int)(i - (long)t * per_row);
    const size_t off = buf_off + (size_t)dest_rows[t] * V * 2;
    if (VEC) {
      const uint4 v = reinterpret_cast<const uint4*>(local + (size_t)t * V)[c];
      for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(windows[p] + off)[c] = v;
    } else {
      const __nv_bfloat16 v = local[(size_t)t * V + c];
      for (int p = 0; p < world; ++p) reinterpret_cast<__nv_bfloat16*>(windows[p] + off)[c] = v;
    }
  }
  __threadfence_system();                                       // this thread's peer stores are ordered before ...
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {              // ... the last block's flag stores (fence cumulativity)
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys(reinterpret_cast<uint32_t*>(windows[p]) + rank, epoch);
      *counter = 0;                                             // (the next call's kernel is stream-ordered after this one)
    }
  }
}

// One thread per rank waits until that rank's rows of this epoch have landed in OUR window.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ flags, int world, uint32_t epoch, uint32_t* __restrict__ err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const uint64_t t0 = esmk::global_timer_ns();
  while ((int32_t)(ld_acquire_sys(flags + r) - epoch) < 0) {
    if (esmk::global_timer_ns() - t0 > 30000000000ull) {        // 30 s: a peer died or never called
      if (err != nullptr) atomicOr(err, (uint32_t)ESMK_ASYNC_PEER_TIMEOUT);
      return;
    }
    __nanosleep(200);
  }
}

}  // namespace

struct esmk_comm {
  NcclComm comm = nullptr;
  int world = 0, rank = 0, device = 0;
  // ---- NVLink peer path (esmk_comm_enable_peer) ----
  uint8_t* window = nullptr;            // this rank's window: [flags | buffer 0 | buffer 1]
  size_t half_bytes = 0;                // size of one buffer
  std::vector<uint8_t*> mapped;         // every rank's window as mapped into this process (own entry = window)
  uint8_t** windows_dev = nullptr;      // the same table on the device
  unsigned* counter = nullptr;          // block counter of the store kernel
  uint32_t epoch = 0;                   // one per collective call; buffer = epoch & 1
};

namespace esmk {

int comm_unique_id(void* id128) {
  ESMK_REQUIRE(id128 != nullptr, "null argument");
  const Nccl& n = nccl();
  if (!n.ok) return fail("esmk_comm_unique_id", n.why);
  NcclUniqueId id;
  const int rc = n.get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, id.internal, sizeof(id.internal));
  return 0;
}

int comm_create(esmk_comm** out, int world, int rank, const void* id128) {
  ESMK_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  const Nccl& n = nccl();
  if (!n.ok) return fail("esmk_comm_create", n.why);
  NcclUniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  esmk_comm* c = new esmk_comm();
  c->world = world;
  c->rank = rank;
  c->device = current_device();
  const int rc = n.comm_init_rank(&c->comm, world, id, rank);
  if (rc != 0) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  *out = c;
  return 0;
}

// Collective: every rank unmaps its peers' windows, all ranks meet (one-byte ncclAllGather), and only then does each
// rank free its own window -- CUDA leaves a cudaFree of exported memory that another process still has mapped undefined.
int comm_disable_peer(esmk_comm* c) {
  ESMK_REQUIRE(c != nullptr, "null communicator");
  if (c->window == nullptr) return 0;
  ESMK_REQUIRE(c->device == current_device(), "the communicator belongs to another device");
  ESMK_CUDA(cudaDeviceSynchronize());
  for (int p = 0; p < (int)c->mapped.size(); ++p)
    if (p != c->rank && c->mapped[p] != nullptr) cudaIpcCloseMemHandle(c->mapped[p]);
  c->mapped.clear();
  uint8_t* vote = nullptr;
  ESMK_CUDA(cudaMalloc(&vote, (size_t)c->world + 1));
  const int rc = nccl().all_gather(vote + c->world, vote, 1, 0, c->comm, nullptr);
  if (rc == 0) cudaStreamSynchronize(nullptr);
  cudaFree(vote);
  if (rc != 0) return nccl_fail("ncclAllGather (peer tear-down barrier)", rc);   // (the window stays allocated)
  cudaFree(c->windows_dev);
  cudaFree(c->counter);
  cudaFree(c->window);
  c->windows_dev = nullptr;
  c->counter = nullptr;
  c->window = nullptr;
  c->half_bytes = 0;
  return 0;
}

void comm_destroy(esmk_comm* c) {
  if (c == nullptr) return;
  // Not a collective: if the peer windows are still up (esmk_comm_disable_peer was not called) the imports are
  // closed, but this rank's own window is deliberately NOT freed -- a peer may still have it mapped -- and goes
  // with the process.
  for (int p = 0; p < (int)c->mapped.size(); ++p)
    if (p != c->rank && c->mapped[p] != nullptr) cudaIpcCloseMemHandle(c->mapped[p]);
  if (c->windows_dev != nullptr) cudaFree(c->windows_dev);
  if (c->counter != nullptr) cudaFree(c->counter);
  if (c->comm != nullptr && nccl().ok) nccl().comm_destroy(c->comm);
  delete c;
}

// Collective over all ranks: allocate this rank's window (two buffers of `buffer_bytes`), exchange the CUDA IPC handles
// (one small ncclAllGather), map every peer's window.  Fails (and leaves the NCCL path usable) where CUDA IPC or peer
// access is unavailable; the outcome is the same on every rank of one node.
int comm_enable_peer(esmk_comm* c, size_t buffer_bytes) {
  ESMK_REQUIRE(c != nullptr && buffer_bytes > 0, "bad peer-window arguments");
  ESMK_REQUIRE(c->device == current_device(), "the communicator belongs to another device");
  ESMK_REQUIRE(c->window == nullptr, "the peer window of this communicator exists already");
  ESMK_REQUIRE(c->world <= kMaxPeerWorld, "peer path supports up to 64 ranks");
  const size_t half = (buffer_bytes + 255) & ~size_t(255);
  uint8_t* win = nullptr;
  uint8_t* stage = nullptr;               // [world][64] handles
  ESMK_CUDA(cudaMalloc(&win, kWindowHeader + 2 * half));
  ESMK_CUDA(cudaMemset(win, 0, kWindowHeader));
  cudaIpcMemHandle_t mine;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle(&mine, win);
  // every rank must reach the all-gather below whatever happened locally: a failed rank sends an all-zero handle
  std::vector<cudaIpcMemHandle_t> all(c->world);
  bool ok = (e == cudaSuccess);
  if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); }
  ESMK_CUDA(cudaMalloc(&stage, (size_t)(c->world + 1) * 64));
  ESMK_CUDA(cudaMemcpy(stage + (size_t)c->world * 64, &mine, 64, cudaMemcpyHostToDevice));
  const int rc = nccl().all_gather(stage + (size_t)c->world * 64, stage, 64, /*ncclInt8*/ 0, c->comm, nullptr);
  if (rc != 0) { cudaFree(stage); cudaFree(win); return nccl_fail("ncclAllGather (IPC handles)", rc); }
  ESMK_CUDA(cudaStreamSynchronize(nullptr));
  ESMK_CUDA(cudaMemcpy(all.data(), stage, (size_t)c->world * 64, cudaMemcpyDeviceToHost));
  cudaFree(stage);
  const cudaIpcMemHandle_t zero = {};
  for (int p = 0; p < c->world; ++p) ok = ok && memcmp(&all[p], &zero, 64) != 0;
  std::vector<uint8_t*> mapped(c->world, nullptr);
  std::string why = ok ? "" : "cudaIpcGetMemHandle failed on a rank";
  for (int p = 0; ok && p < c->world; ++p) {
    if (p == c->rank) { mapped[p] = win; continue; }
    void* ptr = nullptr;
    e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      why = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(p) + "): " + cudaGetErrorString(e);
      cudaGetLastError();
      ok = false;
    } else {
      mapped[p] = static_cast<uint8_t*>(ptr);
    }
  }
  // second handshake: the windows are used only if EVERY rank mapped every peer (and nobody unmaps a window a peer
  // may still be opening)
  uint8_t* vote = nullptr;
  ESMK_CUDA(cudaMalloc(&vote, (size_t)c->world + 1));
  const uint8_t my_vote = ok ? 1 : 0;
  ESMK_CUDA(cudaMemcpy(vote + c->world, &my_vote, 1, cudaMemcpyHostToDevice));
  const int rc2 = nccl().all_gather(vote + c->world, vote, 1, 0, c->comm, nullptr);
  std::vector<uint8_t> votes(c->world, 0);
  if (rc2 == 0) {
    cudaStreamSynchronize(nullptr);
    cudaMemcpy(votes.data(), vote, c->world, cudaMemcpyDeviceToHost);
  }
  cudaFree(vote);
  bool all_ok = rc2 == 0;
  for (int p = 0; p < c->world; ++p) all_ok = all_ok && votes[p] == 1;
  if (!all_ok) {
    for (int p = 0; p < c->world; ++p)
      if (p != c->rank && mapped[p] != nullptr) cudaIpcCloseMemHandle(mapped[p]);
    cudaFree(win);
    return fail("esmk_comm_enable_peer", why.empty() ? "a peer rank could not map the windows" : why);
  }
  uint8_t** table = nullptr;
  unsigned* counter = nullptr;
  ESMK_CUDA(cudaMalloc(&table, sizeof(uint8_t*) * c->world));
  ESMK_CUDA(cudaMemcpy(table, mapped.data(), sizeof(uint8_t*) * c->world, cudaMemcpyHostToDevice));
  ESMK_CUDA(cudaMalloc(&counter, sizeof(unsigned)));
  ESMK_CUDA(cudaMemset(counter, 0, sizeof(unsigned)));
  c->window = win;
  c->half_bytes = half;
  c->mapped = mapped;
  c->windows_dev = table;
  c->counter = counter;
  c->epoch = 0;
  return 0;
}

int peer_allgather_logits(esmk_comm* c, const void* local, int rows, int V, const int32_t* dest_rows, int T, void* out,
                          cudaStream_t st) {
  ESMK_REQUIRE(c && out && rows >= 0 && V >= 1 && T >= 1, "bad peer all-gather arguments");
  ESMK_REQUIRE(rows == 0 || (local && dest_rows), "bad peer all-gather arguments");
  ESMK_REQUIRE(c->device == current_device(), "the communicator belongs to another device");
  ESMK_REQUIRE(c->window != nullptr, "esmk_comm_enable_peer has not been called (or failed)");
  ESMK_REQUIRE((size_t)T * V * 2 <= c->half_bytes, "the gathered result does not fit the peer window");
  const uint32_t epoch = ++c->epoch;
  const size_t buf_off = kWindowHeader + (size_t)(epoch & 1) * c->half_bytes;
  const bool vec = (V & 7) == 0 && (reinterpret_cast<uintptr_t>(local) & 15) == 0;
  const long n = (long)rows * (vec ? V / 8 : V);
  const int blocks = (int)std::max(1L, std::min((n + 255) / 256, 4L * sm_count()));
  if (vec)
    peer_store_rows_kernel<true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)local, dest_rows, c->windows_dev, c->world,
                                                         c->rank, rows, V, buf_off, c->counter, epoch);
  else
    peer_store_rows_kernel<false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)local, dest_rows, c->windows_dev, c->world,
                                                          c->rank, rows, V, buf_off, c->counter, epoch);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  peer_wait_kernel<<<1, kMaxPeerWorld, 0, st>>>(reinterpret_cast<const uint32_t*>(c->window), c->world, epoch,
                                                async_error_word());
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  ESMK_CUDA(cudaMemcpyAsync(out, c->window + buf_off, (size_t)T * V * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int allgather_logits(esmk_comm* c, const void* local, int t_max, int V, const int64_t* perm, int T, void* gathered,
                     void* out, cudaStream_t st) {
  ESMK_REQUIRE(c && local && gathered && t_max >= 1 && V >= 1, "bad all-gather arguments");
  ESMK_REQUIRE(c->device == current_device(), "the communicator belongs to another device");
  const int rc = nccl().all_gather(local, gathered, (size_t)t_max * V, kNcclBfloat16, c->comm, st);
  if (rc != 0) return nccl_fail("ncclAllGather", rc);
  if (out != nullptr && T > 0) {
    ESMK_REQUIRE(perm != nullptr, "perm required to restore the packed order");
    gather_rows_kernel<<<(T + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)gathered, perm, (__nv_bfloat16*)out, T, V);
    count_launch();
    ESMK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace esmk
