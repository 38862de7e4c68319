// nccl_dyn.cpp - see nccl_dyn.h
#include "nccl_dyn.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace pbr {

const NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static std::string failure;
  static bool ok = false;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[4] = {getenv("PBRGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    void* h = nullptr;
    for (const char* n : names) {
      if (!n || !*n) continue;
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) { api.path = n; break; }
    }
    if (!h) { failure = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
    bool all = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { all = false; failure = std::string("NCCL symbol missing: ") + name; } return p; };
    api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    api.GetUniqueId = (int (*)(NcclApi::UniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclApi::Comm*, int, NcclApi::UniqueId, int))sym("ncclCommInitRank");
    api.CommInitAll = (int (*)(NcclApi::Comm*, int, const int*))sym("ncclCommInitAll");
    api.CommDestroy = (int (*)(NcclApi::Comm))sym("ncclCommDestroy");
    api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, NcclApi::Comm, cudaStream_t))sym("ncclReduce");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    ok = all;
  });
  if (!ok) { if (err) *err = failure; return nullptr; }
  return &api;
}

}  // namespace pbr
