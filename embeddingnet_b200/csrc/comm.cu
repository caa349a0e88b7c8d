// NCCL plumbing of the sharded bank paths behind the C ABI: the exchange step of
//   bank kNN      : all-gather of the per-shard (Q, k) top-k records, then en_knn_merge_packed on every rank;
//   bank mining   : all-gather of the per-shard candidate counts, all-reduce(max) of the selected ids.
// The Python host uses torch.distributed for the same collectives; these entry points let a host WITHOUT PyTorch run
// the sharded path through this library alone (one process per GPU, as everywhere in this repo).
//
// NCCL is bound at first use with dlopen("libnccl.so.2"): if a copy is already loaded in the process (PyTorch's) that
// one is reused, and the library has no link-time dependency on NCCL.  Only stable entry points of the NCCL 2.x ABI
// are used (ncclGetUniqueId, ncclCommInitRank, ncclAllGather, ncclAllReduce, ncclCommDestroy, ncclGetErrorString).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "common.cuh"

namespace en {
namespace {

struct NcclUniqueId {
  char internal[EN_COMM_ID_BYTES];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char* (*fn_error_string)(int);
constexpr int kNcclInt8 = 0, kNcclInt64 = 4, kNcclMax = 2;

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_error_string error_string = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return;
    a.get_unique_id = reinterpret_cast<fn_get_unique_id>(dlsym(a.handle, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<fn_comm_init_rank>(dlsym(a.handle, "ncclCommInitRank"));
    a.all_gather = reinterpret_cast<fn_all_gather>(dlsym(a.handle, "ncclAllGather"));
    a.all_reduce = reinterpret_cast<fn_all_reduce>(dlsym(a.handle, "ncclAllReduce"));
    a.comm_destroy = reinterpret_cast<fn_comm_destroy>(dlsym(a.handle, "ncclCommDestroy"));
    a.error_string = reinterpret_cast<fn_error_string>(dlsym(a.handle, "ncclGetErrorString"));
    a.ok = a.get_unique_id && a.comm_init_rank && a.all_gather && a.all_reduce && a.comm_destroy;
  });
  return a;
}

int need_nccl(const char* who) {
  if (api().ok) return EN_OK;
  return fail(EN_ERR_DRIVER, "%s: libnccl.so.2 could not be loaded (%s)", who, dlerror() ? dlerror() : "missing symbols");
}

int nccl_fail(int rc, const char* who) {
  const NcclApi& a = api();
  return fail(EN_ERR_COMM, "%s: NCCL error %d: %s", who, rc, a.error_string ? a.error_string(rc) : "?");
}

struct Comm {
  NcclComm comm;
  int nranks, rank;
};

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

int en_comm_unique_id(void* id_bytes_host) {
  EN_REQUIRE(id_bytes_host != nullptr, "en_comm_unique_id: null output");
  if (int rc = need_nccl("en_comm_unique_id")) return rc;
  NcclUniqueId id;
  if (int rc = api().get_unique_id(&id)) return nccl_fail(rc, "en_comm_unique_id");
  std::memcpy(id_bytes_host, id.internal, EN_COMM_ID_BYTES);
  return EN_OK;
}

int en_comm_init(int nranks, int rank, const void* id_bytes_host, void** comm_out) {
  EN_REQUIRE(id_bytes_host && comm_out && nranks >= 1 && rank >= 0 && rank < nranks,
             "en_comm_init: bad arguments (nranks=%d rank=%d)", nranks, rank);
  if (int rc = need_nccl("en_comm_init")) return rc;
  NcclUniqueId id;
  std::memcpy(id.internal, id_bytes_host, EN_COMM_ID_BYTES);
  NcclComm c = nullptr;
  if (int rc = api().comm_init_rank(&c, nranks, id, rank)) return nccl_fail(rc, "en_comm_init");
  *comm_out = new Comm{c, nranks, rank};
  return EN_OK;
}

int en_comm_allgather(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  EN_REQUIRE(comm && send && recv, "en_comm_allgather: bad arguments");
  if (bytes_per_rank == 0) return EN_OK;
  Comm* c = static_cast<Comm*>(comm);
  if (int rc = api().all_gather(send, recv, bytes_per_rank, kNcclInt8, c->comm, as_stream(stream)))
    return nccl_fail(rc, "en_comm_allgather");
  return EN_OK;
}

int en_comm_allreduce_max_i64(void* comm, const int64_t* send, int64_t* recv, size_t count, void* stream) {
  EN_REQUIRE(comm && send && recv, "en_comm_allreduce_max_i64: bad arguments");
  if (count == 0) return EN_OK;
  Comm* c = static_cast<Comm*>(comm);
  if (int rc = api().all_reduce(send, recv, count, kNcclInt64, kNcclMax, c->comm, as_stream(stream)))
    return nccl_fail(rc, "en_comm_allreduce_max_i64");
  return EN_OK;
}

int en_comm_destroy(void* comm) {
  if (comm == nullptr) return EN_OK;
  Comm* c = static_cast<Comm*>(comm);
  const int rc = api().ok ? api().comm_destroy(c->comm) : 0;
  delete c;
  return rc ? nccl_fail(rc, "en_comm_destroy") : EN_OK;
}

}  // extern "C"
