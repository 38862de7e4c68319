// nccl_dyn.h - NCCL bound at run time (dlopen), so that libpbrgpu.so has no link-time NCCL dependency:
//   * inside a PyTorch process the already loaded libnccl.so.2 (torch's bundled copy) is reused - two NCCL copies in one
//     process are never mixed;
//   * a plain C++ application (the reference's Main/main.cpp with the drop-in integrator) picks up the system libnccl.so.2;
//   * single-GPU use never touches NCCL at all.
// Only the handful of entry points the film reduce needs (ONE ncclReduce(sum, fp32, root 0) per render).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

namespace pbr {

struct NcclApi {
  // mirrors of the NCCL ABI (nccl.h): ncclResult_t = int (0 = ncclSuccess), ncclComm_t = opaque pointer,
  // ncclUniqueId = 128 bytes, ncclFloat32 = 7, ncclSum = 0
  typedef void* Comm;
  struct UniqueId { char internal[128]; };
  static constexpr int kFloat32 = 7, kSum = 0;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommInitAll)(Comm*, int, const int*) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string path;
};

// nullptr + *err when no usable NCCL is found (PBRGPU_NCCL_LIB overrides the search)
const NcclApi* nccl_api(std::string* err);

}  // namespace pbr
