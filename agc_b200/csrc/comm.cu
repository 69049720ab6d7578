// comm.cu -- the exchange step of the multi-GPU path over NCCL, inside the library: one process per GPU, one
// communicator per process, device-to-device all-gathers on the library's stream (NVLink / NVSwitch).  NCCL is resolved at run
// time (dlopen of libnccl.so.2 -- the copy torch already loaded when the host program is Python, the system one otherwise), so
// single-GPU programs do not depend on it.  The unique id travels over whatever side channel the launcher has
// (torch.distributed broadcast in agc_b200/dist.py, a file or MPI for a C++ host).
#include "internal.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#if __has_include(<nccl.h>)
#include <nccl.h>
#else
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
#define ncclSuccess 0
#define ncclUint8 1
#endif

namespace {
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
    int device = 0;
    uint64_t collectives = 0, bytes = 0;
    std::string err;
} g;
std::mutex g_mu;

bool load_nccl()
{
    if (g.lib) return true;
    const char* names[] = { getenv("AGCGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) { if (!n) continue; g.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g.lib) break; }
    if (!g.lib) { g.err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define SYM(field, name) g.field = (decltype(g.field))dlsym(g.lib, name); if (!g.field) { g.err = std::string("libnccl lacks ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(CommCount, "ncclCommCount") SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}
}  // namespace

bool agc_comm_active() { return g.comm != nullptr && g.world > 1; }
uint32_t agc_comm_rank() { return g.rank; }
uint32_t agc_comm_world() { return g.world; }

// all-gather of `bytes` bytes per rank between device buffers, on `st`
int agc_comm_allgather(agcgpu_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st)
{
    if (!g.comm) return agc_fail(ctx, AGCGPU_EINVAL, "comm: no communicator (agcgpu_comm_init)");
    ncclResult_t r = g.AllGather(d_send, d_recv, bytes, ncclUint8, g.comm, st);
    if (r != ncclSuccess) return agc_fail(ctx, AGCGPU_ECUDA, "ncclAllGather: %s", g.GetErrorString(r));
    g.collectives++; g.bytes += bytes * g.world;
    return 0;
}

// variable-size all-gather of device blocks: sizes first (8 bytes per rank), then the blocks padded to the largest one.
// d_mine: this rank's block (my_bytes).  On return the blocks of all ranks lie in ctx->scr_gather at r * stride (stride returned).
int agc_comm_allgatherv(agcgpu_ctx* ctx, const void* d_mine, uint64_t my_bytes, int local_status, std::vector<uint64_t>& sizes, uint64_t* stride_out)
{
    // Every rank ALWAYS enters the size exchange, carrying its status: a rank that failed locally (out of memory, an input outside
    // the coder's envelope ...) makes all ranks fail together instead of leaving the others waiting in the collective.
    const uint32_t W = g.world;
    if (int r = agc_reserve(ctx, ctx->scr_gsz, (size_t)(W + 1) * 16 + 64)) return r;
    uint64_t* d_sz = (uint64_t*)ctx->scr_gsz.p;                      // [0..1] mine (bytes, status), then everybody's pairs
    uint64_t mine[2] = { local_status ? 0 : my_bytes, (uint64_t)(uint32_t)local_status };
    CK(cudaMemcpyAsync(d_sz, mine, 16, cudaMemcpyHostToDevice, ctx->st));
    if (int r = agc_comm_allgather(ctx, d_sz, d_sz + 2, 16, ctx->st)) return r;
    std::vector<uint64_t> pairs((size_t)W * 2, 0);
    CK(cudaMemcpyAsync(pairs.data(), d_sz + 2, (size_t)W * 16, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    sizes.assign(W, 0);
    for (uint32_t r = 0; r < W; ++r) {
        sizes[r] = pairs[2 * r];
        if (pairs[2 * r + 1]) {
            if (r == g.rank) return local_status;                    // the caller's own error message stands
            return agc_fail(ctx, AGCGPU_ECUDA, "sharded step failed on rank %u (code %d)", r, (int)(int32_t)pairs[2 * r + 1]);
        }
    }
    if (sizes[g.rank] != my_bytes) return agc_fail(ctx, AGCGPU_ECUDA, "comm: size all-gather returned a wrong entry for this rank");
    uint64_t mx = 0; for (auto s : sizes) mx = std::max(mx, s);
    const uint64_t stride = (mx + 15) / 16 * 16;
    *stride_out = stride;
    if (!stride) return 0;
    if (int r = agc_reserve(ctx, ctx->scr_gather, (size_t)stride * (W + 1) + 64)) return r;
    uint8_t* d_all = (uint8_t*)ctx->scr_gather.p;
    uint8_t* d_send = d_all + (size_t)stride * W;                    // padded copy of this rank's block
    if (my_bytes) CK(cudaMemcpyAsync(d_send, d_mine, my_bytes, cudaMemcpyDeviceToDevice, ctx->st));
    return agc_comm_allgather(ctx, d_send, d_all, stride, ctx->st);
}

extern "C" {

int agcgpu_comm_unique_id(uint8_t* out_id)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!out_id) return AGCGPU_EINVAL;
    if (!load_nccl()) return AGCGPU_EUNSUPPORTED;
    ncclUniqueId id;
    if (g.GetUniqueId(&id) != ncclSuccess) { g.err = "ncclGetUniqueId failed"; return AGCGPU_ECUDA; }
    static_assert(sizeof(ncclUniqueId) == AGCGPU_UNIQUE_ID_BYTES, "unique id size");
    memcpy(out_id, &id, sizeof id);
    return 0;
}

int agcgpu_comm_init(uint32_t rank, uint32_t world, const uint8_t* id_bytes, int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!id_bytes || world == 0 || rank >= world) return AGCGPU_EINVAL;
    if (!load_nccl()) return AGCGPU_EUNSUPPORTED;
    if (g.comm) { g.CommDestroy(g.comm); g.comm = nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { g.err = "cudaSetDevice failed"; return AGCGPU_ENODEV; }
    ncclUniqueId id; memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = g.CommInitRank(&g.comm, (int)world, id, (int)rank);
    if (r != ncclSuccess) { g.err = std::string("ncclCommInitRank: ") + g.GetErrorString(r); g.comm = nullptr; return AGCGPU_ECUDA; }
    g.rank = rank; g.world = world; g.device = device; g.collectives = 0; g.bytes = 0;
    return 0;
}

int agcgpu_comm_destroy(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.comm) { g.CommDestroy(g.comm); g.comm = nullptr; }
    g.world = 1; g.rank = 0;
    return 0;
}

int agcgpu_comm_get_stats(agcgpu_comm_stats* out)
{
    if (!out) return AGCGPU_EINVAL;
    memset(out, 0, sizeof *out);
    out->rank = g.rank; out->nranks = 1;
    if (g.comm) { int n = 0; if (g.CommCount(g.comm, &n) == ncclSuccess) out->nranks = (uint32_t)n; }
    out->collectives = g.collectives; out->bytes_gathered = g.bytes;
    return 0;
}

const char* agcgpu_comm_last_error(void) { return g.err.c_str(); }

}  // extern "C"
