// agf_nccl.h -- internal: the handful of NCCL entry points the library uses, resolved at run time.
//
// The product's only collective is the Monte-Carlo statistics read-out.  libnccl.so.2 is dlopen'ed on
// first use instead of being a link-time dependency: a single-GPU host needs no NCCL at all, and inside a process that
// already carries an NCCL (torch bundles one) the loader hands back that very library, so communicators created by
// either side are interchangeable.  Declarations restate the public NCCL 2.x C API (nccl.h); the ABI of these entry
// points has been stable since 2.4.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace agf {

struct NcclUniqueId { char internal[128]; };  // ncclUniqueId
typedef struct ncclComm* NcclComm;            // ncclComm_t
enum { kNcclSuccess = 0, kNcclFloat64 = 8 };  // ncclResult_t::ncclSuccess, ncclDataType_t::ncclFloat64

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int nranks, NcclUniqueId id, int rank);
  int (*CommInitAll)(NcclComm*, int ndev, const int* devlist);
  int (*CommDestroy)(NcclComm);
  int (*CommCount)(const NcclComm, int*);
  int (*CommUserRank)(const NcclComm, int*);
  int (*AllGather)(const void* send, void* recv, size_t sendcount, int datatype, NcclComm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
};

// nullptr (and *why set) when libnccl.so.2 cannot be loaded
const NcclApi* nccl_api(const char** why);

}  // namespace agf
