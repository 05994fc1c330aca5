// z-slab decomposition: shared declarations (see cs_dist.cu)
#pragma once
#include "cs_common.cuh"
#include <nccl.h>
#include <vector>

namespace cs {

struct NcclApi {
    bool ok = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi* nccl_api();      // nullptr (and cs_last_error) when libnccl cannot be loaded

// geometry of this rank's slab: owned global cell levels [z_lo, z_hi); the local mesh covers global cell levels [k0, k1)
struct DistGeom {
    int rank = 0, nranks = 1;
    int z_lo = 0, z_hi = 0, gnz = 0, k0 = 0, k1 = 0;
};
struct DistBlock { long long off, per; int nlev; bool node; };

void dist_blocks(const MeshD& m, int family, std::vector<DistBlock>& out);
void dist_mask_host(const MeshD& m, const DistGeom& g, int family, std::vector<double>& mask);
int dist_halo_exchange(const MeshD& m, const DistGeom& g, void* comm, int family, void* vec, long long n, int nvec, bool cplx_vec,
                       void* sendbuf, void* recvbuf, size_t half_bytes, cudaStream_t st);
int dist_allreduce_sum(const DistGeom& g, void* comm, double* buf_d, size_t count, cudaStream_t st);

}  // namespace cs
