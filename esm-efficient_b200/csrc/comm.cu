// The one collective of the path behind the C ABI: an NCCL all-gather of the per-rank, padded
// logits followed by a row gather that restores the original packed order on every rank.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2: the copy PyTorch has already loaded in a torch process, the
// system library for a plain C caller), so libesmk.so has no link-time dependency on it and single-GPU users never
// touch it.  The handful of NCCL declarations used are restated here (nccl.h: ncclUniqueId = 128 opaque bytes,
// ncclBfloat16 = 9, ncclSuccess = 0).
#include <dlfcn.h>

#include <mutex>

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

}  // namespace

struct esmk_comm {
  NcclComm comm = nullptr;
  int world = 0, rank = 0, device = 0;
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

void comm_destroy(esmk_comm* c) {
  if (c == nullptr) return;
  if (c->comm != nullptr && nccl().ok) nccl().comm_destroy(c->comm);
  delete c;
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
