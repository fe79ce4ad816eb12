// nccl_comm.h - the one collective of the path: ncclAllReduce(double, sum) over NVLink / NVSwitch for landmark-sharded
// bundle adjustment, issued natively from the C++ host on the solver's CUDA stream (capturable into the
// trial CUDA graph).  NCCL is bound at run time (dlopen of libnccl.so.2 - the copy already mapped into the process when
// the host also uses torch.distributed, else the system one), so the library has no link-time dependency on it and a
// single-GPU user never loads it.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace g2o_b200 {

constexpr int kNcclUniqueIdBytes = 128;  // sizeof(ncclUniqueId)

class NcclComm {
 public:
  ~NcclComm() { destroy(); }
  // rank 0 creates the id (ncclGetUniqueId) and hands it to the other ranks over any out-of-band channel
  static int unique_id(void* out128, std::string* err);
  // ncclCommInitRank on the CURRENT device; blocks until every rank of the job has called it
  int init(const void* id128, int rank, int world, std::string* err);
  bool active() const { return comm_ != nullptr; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  // in-place sum over all ranks of `count` doubles at device pointer p, ordered on stream s.  0 = ok
  int allreduce_sum(double* p, long long count, cudaStream_t s, std::string* err);
  int allreduce_max(double* p, long long count, cudaStream_t s, std::string* err);
  void destroy();
  static int version();  // NCCL_VERSION_CODE of the bound library, 0 if unavailable

 private:
  void* comm_ = nullptr;
  int rank_ = 0, world_ = 1;
};

}  // namespace g2o_b200
