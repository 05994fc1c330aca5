// z-slab domain decomposition of ONE large system across GPUs (BASELINE config 5): NCCL plumbing, the ownership masks
// and the halo pack / unpack kernels (K10).  The reference has no counterpart (single process, run.py:50-53).
//
// Partition.  Rank r owns the global cell levels [z_lo, z_hi) and the node levels [z_lo, z_hi) (the last rank also the
// top node level nz).  Its context holds a LOCAL mesh: the cells [z_lo - 1, z_hi) -- one ghost cell layer below -- whose
// node levels z_lo - 1 .. z_hi include one ghost node level on each side.  Every stencil of the path reaches one level,
// so the unmodified single-GPU kernels applied to the local mesh are exact on all owned entities as long as the ghost
// entries of their INPUT are valid; outputs at ghost entities are refreshed from the owning neighbour:
//   down (rank - 1) <- my node level z_lo (its upper ghost);           up (rank + 1) <- my node level z_hi - 1 and
//   down            -> my ghost node level z_lo - 1, cell level z_lo - 1     cell level z_hi - 1 (its lower ghosts)
// one packed message per neighbour and direction (ncclSend / ncclRecv in one group, NVLink through NVSwitch).
// Inner products run over owned entities (mask) and are summed with one ncclAllReduce of all scalars of the call.
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy torch already loaded when there is one): the library has
// no link-time dependency on it and single-GPU use never touches it.
#include "cs_internal.h"
#include "cs_dist.h"
#include <dlfcn.h>
#include <cstring>

namespace cs {

static NcclApi g_nccl;
static bool g_nccl_tried = false;

const NcclApi* nccl_api() {
    if (g_nccl_tried) return g_nccl.ok ? &g_nccl : nullptr;
    g_nccl_tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error(std::string("cannot load libnccl.so.2: ") + dlerror()); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) ok = false; return p; };
    g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))sym("ncclCommDestroy");
    g_nccl.AllReduce = (ncclResult_t(*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
    g_nccl.Send = (ncclResult_t(*)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (ncclResult_t(*)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))sym("ncclRecv");
    g_nccl.GroupStart = (ncclResult_t(*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (ncclResult_t(*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(ncclResult_t))sym("ncclGetErrorString");
    g_nccl.ok = ok;
    if (!ok) { set_error("libnccl.so.2 lacks a required symbol"); return nullptr; }
    return &g_nccl;
}

// ---------------------------------------------------------------- ownership ----
// blocks of a family's local vector: (offset, entries per level, number of levels, node-level?)
void dist_blocks(const MeshD& m, int family, std::vector<DistBlock>& out) {
    out.clear();
    const long long nxy = m.nxy;
    if (family == 0) { out.push_back({0, nxy, m.nz, false}); return; }
    if (family == 1) {
        if (m.sym) { out.push_back({0, nxy, m.nz + 1, true}); return; }
        out.push_back({0, nxy, m.nz + 1, true});                      // Ex
        out.push_back({m.nEx, nxy, m.nz + 1, true});                  // Ey
        out.push_back({m.nEx + m.nEy, m.ezp, m.nz, false});           // Ez (axis edge first in every level)
        return;
    }
    out.push_back({0, nxy, m.nz, false});                             // Fx
    if (!m.sym) out.push_back({m.nFx, nxy, m.nz, false});             // Fy
    out.push_back({m.nFx + m.nFy, nxy, m.nz + 1, true});              // Fz
}

// 1.0 on the entities this rank owns
void dist_mask_host(const MeshD& m, const DistGeom& g, int family, std::vector<double>& mask) {
    std::vector<DistBlock> blk;
    dist_blocks(m, family, blk);
    long long n = family == 0 ? m.nC : (family == 1 ? m.nE : m.nF);
    mask.assign((size_t)n, 0.0);
    for (auto& b : blk)
        for (int kl = 0; kl < b.nlev; ++kl) {
            int kg = kl + g.k0;
            bool own = (kg >= g.z_lo && kg < g.z_hi) || (b.node && kg == g.gnz && g.z_hi == g.gnz);
            if (own) for (long long e = 0; e < b.per; ++e) mask[(size_t)(b.off + (long long)kl * b.per + e)] = 1.0;
        }
}

// ---------------------------------------------------------------- K10: halo pack / unpack ----
// one plane = `per` consecutive entries of a block at local level kl; planes are laid out back to back in the buffer,
// system after system
struct HaloPlan { long long off[8]; long long per[8]; int n; long long total; };

template <typename T>
__global__ void k_halo_copy(T* __restrict__ vec, long long nvecstride, T* __restrict__ buf, HaloPlan hp, int pack) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= hp.total) return;
    int sys = blockIdx.y;
    long long o = t, src = 0;
    for (int s = 0; s < hp.n; ++s) {
        if (o < hp.per[s]) { src = hp.off[s] + o; break; }
        o -= hp.per[s];
    }
    T* v = vec + (size_t)sys * nvecstride + src;
    T* b = buf + (size_t)sys * hp.total + t;
    if (pack) *b = *v; else *v = *b;
}

static HaloPlan make_plan(const std::vector<DistBlock>& blk, const DistGeom& g, int which) {
    // which: 0 send down, 1 send up, 2 recv from down, 3 recv from up   (global levels -> local by - k0)
    HaloPlan hp{}; hp.n = 0; hp.total = 0;
    for (auto& b : blk) {
        int lev = -1;
        if (which == 0) lev = b.node ? g.z_lo : -1;                   // the neighbour below misses my lowest owned node level only
        if (which == 1) lev = g.z_hi - 1;                             // node level z_hi-1 and cell level z_hi-1
        if (which == 2) lev = g.z_lo - 1;                             // both kinds: my ghost levels below
        if (which == 3) lev = b.node ? g.z_hi : -1;                   // my upper ghost node level
        if (lev < 0) continue;
        int kl = lev - g.k0;
        if (kl < 0 || kl >= b.nlev) continue;
        hp.off[hp.n] = b.off + (long long)kl * b.per; hp.per[hp.n] = b.per; hp.total += b.per; hp.n++;
    }
    return hp;
}

// refresh the ghost entries of nvec local vectors of a family from the neighbouring slabs
int dist_halo_exchange(const MeshD& m, const DistGeom& g, void* comm, int family, void* vec, long long n, int nvec, bool cplx_vec,
                       void* sendbuf /* 2 halves */, void* recvbuf, size_t half_bytes, cudaStream_t st) {
    if (g.nranks == 1) return CS_OK;
    const NcclApi* nc = nccl_api();
    CS_CHECK(nc, "NCCL is not available");
    std::vector<DistBlock> blk;
    dist_blocks(m, family, blk);
    const bool has_dn = g.rank > 0, has_up = g.rank + 1 < g.nranks;
    HaloPlan sd = make_plan(blk, g, 0), su = make_plan(blk, g, 1), rd = make_plan(blk, g, 2), ru = make_plan(blk, g, 3);
    const size_t es = cplx_vec ? sizeof(cplx) : sizeof(double);
    CS_CHECK((size_t)std::max(std::max(sd.total, su.total), std::max(rd.total, ru.total)) * nvec * es <= half_bytes, "halo buffers too small");
    char* sb = (char*)sendbuf; char* rb = (char*)recvbuf;
    auto run = [&](const HaloPlan& hp, char* buf, int pack) {
        if (hp.total == 0) return;
        dim3 gr((unsigned)cdiv(hp.total, 256), nvec);
        if (cplx_vec) k_halo_copy<cplx><<<gr, 256, 0, st>>>((cplx*)vec, n, (cplx*)buf, hp, pack);
        else k_halo_copy<double><<<gr, 256, 0, st>>>((double*)vec, n, (double*)buf, hp, pack);
    };
    if (has_dn) run(sd, sb, 1);
    if (has_up) run(su, sb + half_bytes, 1);
    const size_t dper = cplx_vec ? 2 : 1;
    ncclResult_t r = nc->GroupStart();
    if (has_dn) { r = nc->Send(sb, sd.total * nvec * dper, ncclDouble, g.rank - 1, (ncclComm_t)comm, st);
                  r = nc->Recv(rb, rd.total * nvec * dper, ncclDouble, g.rank - 1, (ncclComm_t)comm, st); }
    if (has_up) { r = nc->Send(sb + half_bytes, su.total * nvec * dper, ncclDouble, g.rank + 1, (ncclComm_t)comm, st);
                  r = nc->Recv(rb + half_bytes, ru.total * nvec * dper, ncclDouble, g.rank + 1, (ncclComm_t)comm, st); }
    r = nc->GroupEnd();
    CS_CHECK(r == ncclSuccess, "NCCL halo exchange failed");
    if (has_dn) run(rd, rb, 0);
    if (has_up) run(ru, rb + half_bytes, 0);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// sum of `count` doubles over all ranks, in place (K11: all scalars of a Krylov phase in one message)
int dist_allreduce_sum(const DistGeom& g, void* comm, double* buf_d, size_t count, cudaStream_t st) {
    if (g.nranks == 1) return CS_OK;
    const NcclApi* nc = nccl_api();
    CS_CHECK(nc, "NCCL is not available");
    ncclResult_t r = nc->AllReduce(buf_d, buf_d, count, ncclDouble, ncclSum, (ncclComm_t)comm, st);
    CS_CHECK(r == ncclSuccess, "ncclAllReduce failed");
    return CS_OK;
}

}  // namespace cs
