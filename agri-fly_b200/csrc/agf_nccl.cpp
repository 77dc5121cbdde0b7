// agf_nccl.cpp -- run-time binding of NCCL (agf_nccl.h) and the communicator helpers of include/agrifly_b200.h for
// hosts that do not link NCCL themselves (the C++ fleet example, bench.py through ctypes).
#include "agf_nccl.h"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>

#include "agrifly_b200.h"

namespace agf {

int fail_from(int code, const char* what, int cuda_error);
void release_stats_graphs_for(void* comm);  // agf_batch.cu: CUDA graphs that captured a collective of this communicator

const NcclApi* nccl_api(const char** why) {
  static NcclApi api;
  static bool ok = false;
  static std::string err;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return;
    }
    struct { const char* name; void** slot; } syms[] = {
        {"ncclGetUniqueId", (void**)&api.GetUniqueId},   {"ncclCommInitRank", (void**)&api.CommInitRank},
        {"ncclCommInitAll", (void**)&api.CommInitAll},   {"ncclCommDestroy", (void**)&api.CommDestroy},
        {"ncclCommCount", (void**)&api.CommCount},       {"ncclCommUserRank", (void**)&api.CommUserRank},
        {"ncclAllGather", (void**)&api.AllGather},       {"ncclGroupStart", (void**)&api.GroupStart},
        {"ncclGroupEnd", (void**)&api.GroupEnd},         {"ncclGetErrorString", (void**)&api.GetErrorString},
        {"ncclGetVersion", (void**)&api.GetVersion}};
    for (auto& s : syms) {
      *s.slot = dlsym(h, s.name);
      if (!*s.slot) {
        err = std::string("libnccl.so.2 lacks ") + s.name;
        return;
      }
    }
    ok = true;
  });
  if (!ok && why) *why = err.c_str();
  return ok ? &api : nullptr;
}

static int nccl_fail(const NcclApi* a, const char* what, int rc) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, a ? a->GetErrorString(rc) : "NCCL unavailable");
  return fail_from(AGF_ENCCL, buf, 0);
}

}  // namespace agf

using agf::NcclApi;
using agf::NcclComm;
using agf::NcclUniqueId;

extern "C" {

int agf_nccl_version(int* version) {
  const char* why = "";
  const NcclApi* a = agf::nccl_api(&why);
  if (!a) return agf::fail_from(AGF_ENCCL, why, 0);
  if (!version) return agf::fail_from(AGF_EINVAL, "null argument", 0);
  const int rc = a->GetVersion(version);
  return rc == agf::kNcclSuccess ? AGF_OK : agf::nccl_fail(a, "ncclGetVersion", rc);
}

int agf_nccl_get_unique_id(uint8_t id[AGF_NCCL_UNIQUE_ID_BYTES]) {
  const char* why = "";
  const NcclApi* a = agf::nccl_api(&why);
  if (!a) return agf::fail_from(AGF_ENCCL, why, 0);
  if (!id) return agf::fail_from(AGF_EINVAL, "null argument", 0);
  static_assert(sizeof(NcclUniqueId) == AGF_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclUniqueId u;
  const int rc = a->GetUniqueId(&u);
  if (rc != agf::kNcclSuccess) return agf::nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(id, &u, sizeof u);
  return AGF_OK;
}

int agf_nccl_comm_init_rank(const uint8_t id[AGF_NCCL_UNIQUE_ID_BYTES], int nranks, int rank, int device, void** comm_out) {
  const char* why = "";
  const NcclApi* a = agf::nccl_api(&why);
  if (!a) return agf::fail_from(AGF_ENCCL, why, 0);
  if (!id || !comm_out || nranks < 1 || rank < 0 || rank >= nranks) return agf::fail_from(AGF_EINVAL, "bad communicator arguments", 0);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return agf::fail_from(AGF_ECUDA, "cudaSetDevice", (int)e);
  NcclUniqueId u;
  memcpy(&u, id, sizeof u);
  NcclComm c = nullptr;
  const int rc = a->CommInitRank(&c, nranks, u, rank);
  if (rc != agf::kNcclSuccess) return agf::nccl_fail(a, "ncclCommInitRank", rc);
  *comm_out = c;
  return AGF_OK;
}

int agf_nccl_comm_init_all(int ndev, const int* devices, void** comms_out) {
  const char* why = "";
  const NcclApi* a = agf::nccl_api(&why);
  if (!a) return agf::fail_from(AGF_ENCCL, why, 0);
  if (ndev < 1 || !comms_out) return agf::fail_from(AGF_EINVAL, "bad communicator arguments", 0);
  const int rc = a->CommInitAll(reinterpret_cast<NcclComm*>(comms_out), ndev, devices);
  return rc == agf::kNcclSuccess ? AGF_OK : agf::nccl_fail(a, "ncclCommInitAll", rc);
}

int agf_nccl_comm_destroy(void* comm) {
  if (!comm) return AGF_OK;
  const char* why = "";
  const NcclApi* a = agf::nccl_api(&why);
  if (!a) return agf::fail_from(AGF_ENCCL, why, 0);
  agf::release_stats_graphs_for(comm);  // NCCL would otherwise wait for those graphs to be destroyed
  const int rc = a->CommDestroy(reinterpret_cast<NcclComm>(comm));
  return rc == agf::kNcclSuccess ? AGF_OK : agf::nccl_fail(a, "ncclCommDestroy", rc);
}

}  // extern "C"
