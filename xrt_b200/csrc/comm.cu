// Collectives of the sharded ray / wave path behind the C ABI (r02), for hosts that are not Python.
//
// The reference has no communication backend: its multi-device model is XRT_CL.run_parallel_max
// (xrt/backends/raycing/myopencl.py:455-583) -- sliced arguments are cut into one chunk per device, non-sliced ones
// are replicated, results are copied back into the host slices.  One process per GPU replaces that here, and the
// exchange steps the path then has are:
//   * xrt_allreduce_slots  : min / max / sum of the 64 beam-wide reduction slots between the pre-passes and the main
//                            pass of reflect (oes/base.py:861-871, :1257-1263: numpy's whole-array min / max);
//   * xrt_allgather_beam   : the Beam SoA of every rank, in rank order, at a Screen -- ONE grouped NCCL operation for
//                            the 14 arrays, ragged shard lengths allowed (all-gather-v as grouped broadcasts);
//   * xrt_allgather_bytes  : the packed 128-byte input records of the Kirchhoff sum (waves.diffract with ray-sharded
//                            samples), ragged as well.
// NCCL is resolved with dlopen at first use (libnccl.so.2: the copy torch has already loaded when the host is
// Python), so libxrt_b200.so itself has no link-time dependency on it; the communicator is created through
// xrt_nccl_unique_id / xrt_nccl_comm_init (the 128-byte id travels by whatever channel the host has) or handed in as
// an existing ncclComm_t.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "../../include/xrt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace xrtb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    set_last_error("NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
    return nullptr;
  }
#define XRTB_NCCL_SYM(field, sym)                                               \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));    \
  if (!api.field) {                                                             \
    set_last_error("NCCL: symbol %s not found", sym);                           \
    api.handle = nullptr;                                                       \
    return nullptr;                                                             \
  }
  XRTB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  XRTB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  XRTB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  XRTB_NCCL_SYM(AllGather, "ncclAllGather")
  XRTB_NCCL_SYM(Broadcast, "ncclBroadcast")
  XRTB_NCCL_SYM(GroupStart, "ncclGroupStart")
  XRTB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  XRTB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef XRTB_NCCL_SYM
  return &api;
}

#define XRTB_NCCL_CHECK(api, expr)                                                            \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, (api)->GetErrorString(_r)); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

// ---------------------------------------------------------------- reduction slots
// 0 = max, 1 = min (of the order-preserving encodings), 2 = sum; layout: reflect.cu (the numpy mirror of this table is
// xrt_b200.parallel.slot_kinds, a test compares the two)
__host__ __device__ inline int slot_kind(int k) {
  if (k == 3) return 1;                                  // first good ray key
  if (k >= 4 && k < 28) return ((k - 4) & 3) == 0 ? 1 : 0;  // min t1 | max t2, max |dz1|, max |dz2| per variant
  if (k >= 28 && k < 31) return 1;                       // footprint min
  if (k >= 31 && k < 34) return 0;                       // footprint max
  if (k == 40) return 1;                                 // multiple_reflect, derivative stage: min t1
  if (k >= 41 && k < 44) return 0;                       // ... max t2, max |dz1|, max |dz2|
  if (k >= 34) return 2;                                 // counters (and the scratch slots, zero between passes)
  return 0;                                              // 0..2: max |a|, |b|, |c|
}

__global__ void slots_combine_kernel(int world, const unsigned long long* __restrict__ gathered, int sums,
                                     unsigned long long* __restrict__ red) {
  const int k = threadIdx.x;
  if (k >= XRT_REDUCTION_SLOTS) return;
  const int kind = slot_kind(k);
  unsigned long long acc = gathered[k];
  for (int r = 1; r < world; ++r) {
    const unsigned long long v = gathered[(size_t)r * XRT_REDUCTION_SLOTS + k];
    acc = kind == 0 ? (v > acc ? v : acc) : (kind == 1 ? (v < acc ? v : acc) : acc + v);
  }
  if (kind == 2 && !sums) return;  // counters were summed by an earlier all-reduce: keep this rank's
  red[k] = acc;
}

}  // namespace xrtb

using namespace xrtb;

extern "C" {

int xrt_slot_kind(int slot) { return slot_kind(slot); }

int xrt_nccl_unique_id(void* id128) {
  XRTB_REQUIRE(id128 != nullptr, "xrt_nccl_unique_id: NULL argument");
  NcclApi* api = nccl_api();
  if (!api) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  XRTB_NCCL_CHECK(api, api->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return 0;
}

int xrt_nccl_comm_init(int world, int rank, const void* id128, void** comm) {
  XRTB_REQUIRE(id128 && comm && world >= 1 && rank >= 0 && rank < world, "xrt_nccl_comm_init: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return 1;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c;
  XRTB_NCCL_CHECK(api, api->CommInitRank(&c, world, id, rank));
  *comm = c;
  return 0;
}

int xrt_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  NcclApi* api = nccl_api();
  if (!api) return 1;
  XRTB_NCCL_CHECK(api, api->CommDestroy(reinterpret_cast<ncclComm_t>(comm)));
  return 0;
}

int xrt_allreduce_slots(void* comm, int world, uint64_t* reductions, int sums, uint64_t* scratch, void* stream) {
  XRTB_REQUIRE(comm && reductions && scratch && world >= 1, "xrt_allreduce_slots: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  XRTB_NCCL_CHECK(api, api->AllGather(reductions, scratch, XRT_REDUCTION_SLOTS, ncclUint64,
                                      reinterpret_cast<ncclComm_t>(comm), st));
  slots_combine_kernel<<<1, XRT_REDUCTION_SLOTS, 0, st>>>(world, reinterpret_cast<unsigned long long*>(scratch), sums,
                                                          reinterpret_cast<unsigned long long*>(reductions));
  XRTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// all-gather-v of raw bytes: rank r contributes counts[r] bytes; out receives them in rank order
int xrt_allgather_bytes(void* comm, int world, int rank, const uint64_t* counts, const void* in, void* out,
                        void* stream) {
  XRTB_REQUIRE(comm && counts && out && world >= 1 && rank >= 0 && rank < world, "xrt_allgather_bytes: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  ncclComm_t c = reinterpret_cast<ncclComm_t>(comm);
  bool equal = true;
  for (int r = 1; r < world; ++r) equal = equal && counts[r] == counts[0];
  if (equal) {
    if (counts[0]) XRTB_NCCL_CHECK(api, api->AllGather(in, out, counts[0], ncclUint8, c, st));
    return 0;
  }
  XRTB_NCCL_CHECK(api, api->GroupStart());
  size_t off = 0;
  for (int r = 0; r < world; ++r) {
    if (counts[r])
      XRTB_NCCL_CHECK(api, api->Broadcast(r == rank ? in : static_cast<unsigned char*>(out) + off,
                                          static_cast<unsigned char*>(out) + off, counts[r], ncclUint8, r, c, st));
    off += counts[r];
  }
  XRTB_NCCL_CHECK(api, api->GroupEnd());
  return 0;
}

// The single all-gather of the Beam at a Screen (north star): every array of `in` (counts[rank] rays) lands in `out`
// (sum(counts) rays) in rank order; one NCCL group = one fused operation on the wire.
int xrt_allgather_beam(void* comm, int world, int rank, const uint64_t* counts, const xrt_beam* in, const xrt_beam* out,
                       void* stream) {
  XRTB_REQUIRE(comm && counts && in && out && world >= 1 && rank >= 0 && rank < world,
               "xrt_allgather_beam: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  ncclComm_t c = reinterpret_cast<ncclComm_t>(comm);
  struct Arr {
    const void* src;
    void* dst;
    size_t elem;
  };
  std::vector<Arr> arrs;
  auto add = [&](const void* s, void* d, size_t elem) {
    if (s && d) arrs.push_back(Arr{s, d, elem});
  };
  add(in->x, out->x, 8);
  add(in->y, out->y, 8);
  add(in->z, out->z, 8);
  add(in->a, out->a, 8);
  add(in->b, out->b, 8);
  add(in->c, out->c, 8);
  add(in->path, out->path, 8);
  add(in->E, out->E, 8);
  add(in->Jss, out->Jss, 8);
  add(in->Jpp, out->Jpp, 8);
  add(in->Jsp, out->Jsp, 16);
  add(in->Es, out->Es, 16);
  add(in->Ep, out->Ep, 16);
  add(in->state, out->state, 4);
  add(in->theta, out->theta, 8);
  add(in->s, out->s, 8);
  add(in->phi, out->phi, 8);
  add(in->r, out->r, 8);
  bool equal = true;
  for (int r = 1; r < world; ++r) equal = equal && counts[r] == counts[0];
  XRTB_NCCL_CHECK(api, api->GroupStart());
  for (const Arr& a : arrs) {
    if (equal) {
      if (counts[0]) XRTB_NCCL_CHECK(api, api->AllGather(a.src, a.dst, counts[0] * a.elem, ncclUint8, c, st));
    } else {
      size_t off = 0;
      for (int r = 0; r < world; ++r) {
        if (counts[r])
          XRTB_NCCL_CHECK(api, api->Broadcast(r == rank ? a.src : static_cast<unsigned char*>(a.dst) + off,
                                              static_cast<unsigned char*>(a.dst) + off, counts[r] * a.elem, ncclUint8, r,
                                              c, st));
        off += counts[r] * a.elem;
      }
    }
  }
  XRTB_NCCL_CHECK(api, api->GroupEnd());
  return 0;
}

}  // extern "C"
