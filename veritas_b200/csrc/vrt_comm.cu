// Multi-GPU x-slab plumbing (no counterpart in the reference): per-stage exchange of the 3-column
// f halo between x-neighbours (grouped ncclSend/ncclRecv, walls have one neighbour) and the in-place all-gather of
// the per-slab charge/current moments so that every rank runs the 1-D Poisson + Maxwell solve redundantly.
// NCCL is bound at run time (dlopen) so that single-GPU use needs no NCCL at all and a host process that already
// loaded a libnccl (e.g. torch's) shares it.
#include "vrt_internal.cuh"
#include <cstring>
#include <dlfcn.h>
#if defined(__has_include) && __has_include(<nccl.h>)
#include <nccl.h>
#else
// No NCCL development header on this machine: the handful of declarations the dlopen'ed entry points need (NCCL's public ABI:
// opaque communicator, 128-byte unique id, result / datatype enumerators), so that the library — single-GPU path included — still builds.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
#endif

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

bool load_nccl(std::string* err) {
    if (g_nccl.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
    if (!g_nccl.h) { *err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name); if (!g_nccl.field) { *err = std::string("dlsym ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}
}  // namespace

#define VRT_NCCL(ctx, call)                                                                   \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) { (ctx)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return VRT_ERR_NCCL; } \
    } while (0)

extern "C" int vrt_nccl_unique_id(void* out128) {
    std::string err;
    if (!out128 || !load_nccl(&err)) return VRT_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    return g_nccl.GetUniqueId((ncclUniqueId*)out128) == ncclSuccess ? 0 : VRT_ERR_NCCL;
}

extern "C" int vrt_comm_init(vrt_ctx* c, const void* unique_id128, int rank, int n_ranks) {
    if (!c || !unique_id128) return VRT_ERR_ARG;
    if (rank != c->rank || n_ranks != c->n_ranks) { c->err = "vrt_comm_init: rank/n_ranks differ from vrt_set_slab"; return VRT_ERR_ARG; }
    if (!load_nccl(&c->err)) return VRT_ERR_NCCL;
    cudaSetDevice(c->device);
    ncclUniqueId id; memcpy(&id, unique_id128, sizeof(id));
    ncclComm_t comm;
    VRT_NCCL(c, g_nccl.CommInitRank(&comm, n_ranks, id, rank));
    c->nccl_comm = comm;
    if (!c->comm_stream) {
        VRT_CUDA(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
        VRT_CUDA(c, cudaEventCreateWithFlags(&c->ev_m, cudaEventDisableTiming));
        VRT_CUDA(c, cudaEventCreateWithFlags(&c->ev_ag, cudaEventDisableTiming));
        for (auto& S : c->S) {
            VRT_CUDA(c, cudaEventCreateWithFlags(&S.ev_k, cudaEventDisableTiming));
            VRT_CUDA(c, cudaEventCreateWithFlags(&S.ev_h, cudaEventDisableTiming));
        }
    }
    return 0;
}

// halo exchange of species s behind its stage kernel: enqueued on the communication stream, signalled through S.ev_h
int vrt_comm_halo_exchange(vrt_ctx* c, int s) {
    if (!c->nccl_comm) { c->err = "multi-rank context without vrt_comm_init"; return VRT_ERR_STATE; }
    VrtSpeciesState& S = c->S[s];
    VrtSlabDev& L = S.slab;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    double* f = L.f[S.i_f1];
    const size_t n = (size_t)L.gx * L.pitch;
    cudaStream_t main_stream = c->stream;
    VRT_CUDA(c, cudaEventRecord(S.ev_k, main_stream));
    VRT_CUDA(c, cudaStreamWaitEvent(c->comm_stream, S.ev_k, 0));
    struct Restore { vrt_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, main_stream};
    c->stream = c->comm_stream;
    VRT_NCCL(c, g_nccl.GroupStart());
    if (c->rank > 0) {
        VRT_NCCL(c, g_nccl.Send(f + (long)L.gx * L.pitch, n, ncclDouble, c->rank - 1, comm, c->stream));        // own columns [0,3)
        VRT_NCCL(c, g_nccl.Recv(f, n, ncclDouble, c->rank - 1, comm, c->stream));                               // halo [-3,0)
    }
    if (c->rank < c->n_ranks - 1) {
        VRT_NCCL(c, g_nccl.Send(f + (long)L.n_x * L.pitch, n, ncclDouble, c->rank + 1, comm, c->stream));       // own [n_x-3,n_x)
        VRT_NCCL(c, g_nccl.Recv(f + (long)(L.n_x + L.gx) * L.pitch, n, ncclDouble, c->rank + 1, comm, c->stream));   // halo [n_x,n_x+3)
    }
    VRT_NCCL(c, g_nccl.GroupEnd());
    VRT_CUDA(c, cudaEventRecord(S.ev_h, c->comm_stream));
    S.halo_pending = true;
    return 0;
}

// the stream the context computes on waits for the outstanding halo exchange of species s (all species: s < 0)
int vrt_comm_wait_halo(vrt_ctx* c, int s) {
    for (int k = 0; k < c->n_species; k++) {
        VrtSpeciesState& S = c->S[k];
        if ((s < 0 || s == k) && S.halo_pending) { VRT_CUDA(c, cudaStreamWaitEvent(c->stream, S.ev_h, 0)); S.halo_pending = false; }
    }
    return 0;
}

// all-gather of the slab moments, on the communication stream between two events of the compute stream
int vrt_comm_gather_moments(vrt_ctx* c) {
    if (!c->nccl_comm) { c->err = "multi-rank context without vrt_comm_init"; return VRT_ERR_STATE; }
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const size_t n = (size_t)(c->x_end - c->x_begin);
    cudaStream_t main_stream = c->stream;
    VRT_CUDA(c, cudaEventRecord(c->ev_m, main_stream));
    VRT_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_m, 0));
    struct Restore { vrt_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, main_stream};
    c->stream = c->comm_stream;
    VRT_NCCL(c, g_nccl.GroupStart());
    for (int s = 0; s < c->n_species; s++)
        VRT_NCCL(c, g_nccl.AllGather(c->S[s].d_charges + c->x_begin, c->S[s].d_charges, n, ncclDouble, comm, c->stream));
    VRT_NCCL(c, g_nccl.AllGather(c->F.J + c->x_begin, c->F.J, n, ncclDouble, comm, c->stream));
    VRT_NCCL(c, g_nccl.GroupEnd());
    VRT_CUDA(c, cudaEventRecord(c->ev_ag, c->comm_stream));
    VRT_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_ag, 0));
    return 0;
}

void vrt_comm_destroy(vrt_ctx* c) {
    if (c->comm_stream) {
        cudaStreamSynchronize(c->comm_stream);
        for (auto& S : c->S) { if (S.ev_k) cudaEventDestroy(S.ev_k); if (S.ev_h) cudaEventDestroy(S.ev_h); S.ev_k = S.ev_h = nullptr; }
        if (c->ev_m) cudaEventDestroy(c->ev_m);
        if (c->ev_ag) cudaEventDestroy(c->ev_ag);
        cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr;
    }
    if (c->nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_t)c->nccl_comm); c->nccl_comm = nullptr; }
}
