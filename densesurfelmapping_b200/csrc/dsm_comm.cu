// Multi-GPU side of the C ABI (include/dsm.h "multi-GPU"): one context per GPU (one process or thread each), frames of a
// batch sharded across the contexts, and ONE gather of the per-GPU surfel deltas onto a root rank at the end of a batch
// (BASELINE configs[3]).  The reference has no counterpart: fuse_initialize_map is single-process
// (fusion_functions.cpp:30-83); what travels is what SurfelMap::fuse_map consumes after the call (surfel_map.cpp:1077-1109):
// the new surfels and the updated local surfels of every frame.
//
// Wire format of one rank's payload (little endian; tests/ and densesurfelmapping_b200/gather.py restate it):
//   int32  magic 'DSMD', n_frames, n_new_total, n_pool_total
//   int32  n_new[n_frames]
//   int32  pool_ofs[n_frames + 1]
//   pad to a multiple of 16 bytes
//   dsm_surfel_t new[n_new_total]     valid new surfels only, frame by frame, seed-index order (fusion_functions.cpp:320-359)
//   dsm_surfel_t pool[n_pool_total]   the frames' updated local surfels, batch order
// Only valid records travel (the padded [n_frames][S] new-surfel buffer never does).  Counts must be known on the host
// to size the messages, so the call waits for the batch's kernels once (new-surfel counts to pinned memory), then
// packs on a side stream (the next batch's kernels may start as soon as the pack has read the buffers), exchanges the
// byte counts with one ncclAllGather and moves the payloads with one grouped ncclSend / ncclRecv.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the NCCL torch already loaded,
// in a plain C++ process the system one), so the library has no link-time dependency on it.
#include "dsm_ctx.hpp"
#include <cstring>
#include <new>
#include <dlfcn.h>
#include <nccl.h>

struct DsmNccl
{
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static DsmNccl *nccl_api()
{
    static DsmNccl api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
#define SYM(field, name)                                        \
    api.field = (decltype(api.field))dlsym(h, name);            \
    if (!api.field) return nullptr;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    api.lib = h;
    return &api;
}

struct DsmComm
{
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;       // pack + collectives
    cudaEvent_t ev_kernels = nullptr;    // the batch's kernels have finished (counts readable)
    cudaEvent_t ev_packed = nullptr;     // the pack has read newsurf / pool: the next batch may overwrite them
    unsigned char *pack = nullptr;       // device: this rank's payload
    size_t pack_cap = 0;
    unsigned char *recv = nullptr;       // device, root only: all payloads back to back
    size_t recv_cap = 0;
    int64_t *d_bytes = nullptr;          // device [nranks]
    int64_t *h_bytes = nullptr;          // pinned [nranks]
    int32_t *h_hdr = nullptr;            // pinned header staging + new-surfel counts
    size_t hdr_cap = 0;
    std::vector<size_t> ofs;             // root: byte offset of every rank's payload in recv (nranks + 1)
    int last_root = -1;
};

#define NK(call)                                                                                                  \
    do                                                                                                            \
    {                                                                                                             \
        ncclResult_t _r = (call);                                                                                 \
        if (_r != ncclSuccess)                                                                                    \
        {                                                                                                         \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, api->GetErrorString(_r)); \
            return DSM_E_NCCL;                                                                                    \
        }                                                                                                         \
    } while (0)

static size_t header_bytes(int n_frames) { return ((size_t)(4 + n_frames + n_frames + 1) * 4 + 15) / 16 * 16; }
static size_t payload_cap(const dsm_ctx *ctx)
{
    return header_bytes(ctx->p.max_batch) + ((size_t)ctx->p.max_batch * ctx->S + (size_t)(ctx->p.max_local_surfels > 0 ? ctx->p.max_local_surfels : 1)) * sizeof(dsm_surfel_t);
}

// valid new surfels of frame (frame0 + blockIdx.y) -> packed[offset of the frame .. ), one thread per 4-byte word
__global__ void k_pack_new(const dsm_surfel_t *newsurf, const int32_t *nnew, int S, int nb, dsm_surfel_t *packed)
{
    const int f = blockIdx.y;
    int before = 0;
    for (int i = 0; i < f; i++) before += nnew[i]; // nb <= max_batch: a few hundred adds at most
    const int n = nnew[f];
    const uint32_t *src = reinterpret_cast<const uint32_t *>(newsurf + (size_t)f * S);
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + before);
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n * 11; w += gridDim.x * blockDim.x) dst[w] = src[w];
    (void)nb;
}

extern "C" int dsm_comm_unique_id(void *id_out)
{
    if (!id_out) return DSM_E_INVALID;
    DsmNccl *api = nccl_api();
    if (!api) return DSM_E_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return DSM_E_NCCL;
    memcpy(id_out, &id, DSM_COMM_ID_BYTES);
    return DSM_OK;
}

extern "C" int dsm_comm_destroy(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    DsmComm *c = ctx->comm;
    if (!c) return DSM_OK;
    cudaSetDevice(ctx->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DsmNccl *api = nccl_api();
    if (c->comm && api) api->CommDestroy(c->comm);
    cudaFree(c->pack);
    cudaFree(c->recv);
    cudaFree(c->d_bytes);
    cudaFreeHost(c->h_bytes);
    cudaFreeHost(c->h_hdr);
    if (c->ev_kernels) cudaEventDestroy(c->ev_kernels);
    if (c->ev_packed) cudaEventDestroy(c->ev_packed);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    ctx->comm = nullptr;
    return DSM_OK;
}

extern "C" int dsm_comm_init(dsm_ctx *ctx, const void *id, int rank, int nranks)
{
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return DSM_E_INVALID;
    if (ctx->comm) return DSM_E_STATE;
    DsmNccl *api = nccl_api();
    if (!api)
    {
        snprintf(ctx->err, sizeof(ctx->err), "libnccl.so.2 could not be loaded: %s", dlerror());
        return DSM_E_NCCL;
    }
    CK(cudaSetDevice(ctx->device));
    DsmComm *c = new (std::nothrow) DsmComm();
    if (!c) return DSM_E_NOMEM;
    ctx->comm = c;
    c->rank = rank, c->nranks = nranks;
    ncclUniqueId uid;
    memcpy(&uid, id, DSM_COMM_ID_BYTES);
    NK(api->CommInitRank(&c->comm, nranks, uid, rank));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_kernels, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    c->pack_cap = payload_cap(ctx);
    CK(cudaMalloc((void **)&c->pack, c->pack_cap));
    CK(cudaMalloc((void **)&c->d_bytes, (size_t)nranks * sizeof(int64_t)));
    CK(cudaMallocHost((void **)&c->h_bytes, (size_t)nranks * sizeof(int64_t)));
    c->hdr_cap = header_bytes(ctx->p.max_batch) + (size_t)ctx->p.max_batch * 4;
    CK(cudaMallocHost((void **)&c->h_hdr, c->hdr_cap));
    c->ofs.assign((size_t)nranks + 1, 0);
    return DSM_OK;
}

extern "C" int dsm_gather_deltas(dsm_ctx *ctx, int root)
{
    if (!ctx) return DSM_E_INVALID;
    DsmComm *c = ctx->comm;
    DsmNccl *api = nccl_api();
    if (!c || !api) return DSM_E_STATE;
    if (root < 0 || root >= c->nranks) return DSM_E_INVALID;
    if (!ctx->ran || ctx->in_flight || ctx->res_active) return DSM_E_STATE; // the deltas of a dsm_batch_run batch
    CK(cudaSetDevice(ctx->device));
    const int nb = ctx->nb, S = ctx->S;
    // (1) new-surfel counts of the batch -> pinned memory; this waits for the batch's kernels
    int32_t *h_cnt = c->h_hdr + header_bytes(ctx->p.max_batch) / 4;
    CK(cudaStreamSynchronize(c->stream)); // the previous gather has consumed the staging buffers
    CK(cudaMemcpyAsync(h_cnt, ctx->d.nnew, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(c->ev_kernels, ctx->stream));
    CK(cudaEventSynchronize(c->ev_kernels));
    // (2) header on the host
    int n_new_total = 0;
    for (int f = 0; f < nb; f++) n_new_total += h_cnt[f];
    const int n_pool = ctx->n_pool;
    int32_t *h = c->h_hdr;
    h[0] = 0x444d5344; // 'DSMD'
    h[1] = nb, h[2] = n_new_total, h[3] = n_pool;
    for (int f = 0; f < nb; f++) h[4 + f] = h_cnt[f];
    for (int f = 0; f <= nb; f++) h[4 + nb + f] = ctx->h_ofs[f];
    const size_t hb = header_bytes(nb);
    const size_t my_bytes = hb + ((size_t)n_new_total + (size_t)n_pool) * sizeof(dsm_surfel_t);
    if (my_bytes > c->pack_cap) return DSM_E_CAPACITY;
    // (3) pack on the side stream
    cudaStream_t cs = c->stream;
    CK(cudaMemcpyAsync(c->pack, h, hb, cudaMemcpyHostToDevice, cs));
    dsm_surfel_t *body = reinterpret_cast<dsm_surfel_t *>(c->pack + hb);
    if (n_new_total > 0)
    {
        dim3 grid(16, nb);
        k_pack_new<<<grid, 256, 0, cs>>>(ctx->d.newsurf, ctx->d.nnew, S, nb, body);
    }
    if (n_pool > 0)
        CK(cudaMemcpyAsync(body + n_new_total, ctx->d.pool, (size_t)n_pool * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, cs));
    CK(cudaEventRecord(c->ev_packed, cs));
    CK(cudaStreamWaitEvent(ctx->stream, c->ev_packed, 0)); // later work on the compute stream may overwrite newsurf / pool
    // (4) byte counts of every rank
    c->h_bytes[c->rank] = (int64_t)my_bytes;
    CK(cudaMemcpyAsync(c->d_bytes + c->rank, c->h_bytes + c->rank, sizeof(int64_t), cudaMemcpyHostToDevice, cs));
    if (c->nranks > 1)
    {
        NK(api->AllGather(c->d_bytes + c->rank, c->d_bytes, 1, ncclInt64, c->comm, cs));
        CK(cudaMemcpyAsync(c->h_bytes, c->d_bytes, (size_t)c->nranks * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
        CK(cudaStreamSynchronize(cs));
    }
    // (5) payloads
    if (c->rank == root)
    {
        size_t total = 0;
        for (int r = 0; r < c->nranks; r++)
        {
            c->ofs[r] = total;
            total += ((size_t)c->h_bytes[r] + 255) / 256 * 256; // every payload starts 256-byte aligned
        }
        c->ofs[c->nranks] = total;
        if (total > c->recv_cap)
        { // grown on demand (first call: nranks payloads of this context's capacity)
            CK(cudaStreamSynchronize(cs));
            cudaFree(c->recv);
            c->recv = nullptr;
            size_t want = (size_t)c->nranks * ((c->pack_cap + 255) / 256 * 256);
            if (want < total) want = total;
            CK(cudaMalloc((void **)&c->recv, want));
            c->recv_cap = want;
        }
        CK(cudaMemcpyAsync(c->recv + c->ofs[root], c->pack, my_bytes, cudaMemcpyDeviceToDevice, cs));
    }
    if (c->nranks > 1)
    {
        NK(api->GroupStart());
        if (c->rank == root)
        {
            for (int r = 0; r < c->nranks; r++)
                if (r != root) NK(api->Recv(c->recv + c->ofs[r], (size_t)c->h_bytes[r], ncclInt8, r, c->comm, cs));
        }
        else
            NK(api->Send(c->pack, my_bytes, ncclInt8, root, c->comm, cs));
        NK(api->GroupEnd());
    }
    c->last_root = root;
    CK(cudaGetLastError());
    return DSM_OK;
}

extern "C" int dsm_gather_wait(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->comm) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->comm->stream));
    return DSM_OK;
}

extern "C" int dsm_gathered_device(dsm_ctx *ctx, void **dev_ptr, size_t *rank_offsets)
{
    if (!ctx || !dev_ptr || !rank_offsets) return DSM_E_INVALID;
    DsmComm *c = ctx->comm;
    if (!c || c->last_root != c->rank) return DSM_E_STATE;
    *dev_ptr = c->recv;
    for (int r = 0; r <= c->nranks; r++) rank_offsets[r] = c->ofs[r];
    return DSM_OK;
}

extern "C" int dsm_gathered_rank_bytes(dsm_ctx *ctx, int rank, size_t *bytes)
{
    if (!ctx || !bytes) return DSM_E_INVALID;
    DsmComm *c = ctx->comm;
    if (!c || c->last_root != c->rank) return DSM_E_STATE;
    if (rank < 0 || rank >= c->nranks) return DSM_E_INVALID;
    *bytes = (size_t)c->h_bytes[rank];
    return DSM_OK;
}

extern "C" int dsm_gathered_download(dsm_ctx *ctx, int rank, void *host_out, size_t cap)
{
    if (!ctx || !host_out) return DSM_E_INVALID;
    DsmComm *c = ctx->comm;
    if (!c || c->last_root != c->rank) return DSM_E_STATE;
    if (rank < 0 || rank >= c->nranks) return DSM_E_INVALID;
    const size_t nbytes = (size_t)c->h_bytes[rank];
    if (nbytes > cap) return DSM_E_CAPACITY;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(host_out, c->recv + c->ofs[rank], nbytes, cudaMemcpyDeviceToHost));
    return DSM_OK;
}
