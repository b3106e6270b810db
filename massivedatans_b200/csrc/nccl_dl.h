// nccl_dl.h -- NCCL bound at run time (dlopen), so that libmdns_b200.so carries no link-time
// dependency on it: single-GPU users never load NCCL, and a process that already holds a copy
// (e.g. the one bundled with PyTorch, same SONAME) shares it.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>      // types and enums only; no symbol is linked

namespace mdns {

struct NcclApi {
	ncclResult_t (*GetVersion)(int *) = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
	                          cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
	                          cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

// nullptr (and mdns_last_error set) when libnccl.so.2 cannot be loaded
const NcclApi *nccl_api();

}  // namespace mdns
