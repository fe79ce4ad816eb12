// nccl_comm.cpp - see nccl_comm.h.  Minimal run-time binding of the five NCCL entry points this path needs
// (prototypes as published in nccl.h; ncclDataType_t ncclFloat64 = 8, ncclRedOp_t ncclSum = 0 / ncclMax = 2).
#include "nccl_comm.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace g2o_b200 {
namespace {

struct UniqueId { char internal[kNcclUniqueIdBytes]; };
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct Api {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  GetVersionFn get_version = nullptr;
  std::string load_error;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    // 1. the copy already mapped into this process (torch ships its own libnccl.so.2: two NCCLs in one process must
    //    be avoided)  2. an explicit path  3. the system library
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) if (const char* p = getenv("G2O_B200_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { a.load_error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
    a.handle = h;
    a.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
    a.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
    a.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
    a.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
    a.get_error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
    a.get_version = (GetVersionFn)dlsym(h, "ncclGetVersion");
    if (!a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.comm_destroy) {
      a.load_error = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy";
      a.handle = nullptr;
    }
  });
  return a;
}

int check(int rc, const char* what, std::string* err) {
  if (rc == 0) return 0;
  if (err) {
    Api& a = api();
    *err = std::string(what) + " failed: " + (a.get_error_string ? a.get_error_string(rc) : "NCCL error");
  }
  return rc;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

}  // namespace

int NcclComm::version() {
  Api& a = api();
  int v = 0;
  if (a.handle && a.get_version) a.get_version(&v);
  return v;
}

int NcclComm::unique_id(void* out128, std::string* err) {
  Api& a = api();
  if (!a.handle) { if (err) *err = a.load_error; return -1; }
  UniqueId id;
  memset(&id, 0, sizeof(id));
  if (int rc = check(a.get_unique_id(&id), "ncclGetUniqueId", err)) return rc;
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int NcclComm::init(const void* id128, int rank, int world, std::string* err) {
  Api& a = api();
  if (!a.handle) { if (err) *err = a.load_error; return -1; }
  destroy();
  UniqueId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  if (int rc = check(a.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank", err)) return rc;
  comm_ = comm;
  rank_ = rank;
  world_ = world;
  return 0;
}

int NcclComm::allreduce_sum(double* p, long long count, cudaStream_t s, std::string* err) {
  if (!comm_) { if (err) *err = "no NCCL communicator"; return -1; }
  return check(api().all_reduce(p, p, (size_t)count, kNcclFloat64, kNcclSum, comm_, s), "ncclAllReduce", err);
}
int NcclComm::allreduce_max(double* p, long long count, cudaStream_t s, std::string* err) {
  if (!comm_) { if (err) *err = "no NCCL communicator"; return -1; }
  return check(api().all_reduce(p, p, (size_t)count, kNcclFloat64, kNcclMax, comm_, s), "ncclAllReduce", err);
}

void NcclComm::destroy() {
  if (comm_) {
    api().comm_destroy(comm_);
    comm_ = nullptr;
  }
  rank_ = 0;
  world_ = 1;
}

}  // namespace g2o_b200
