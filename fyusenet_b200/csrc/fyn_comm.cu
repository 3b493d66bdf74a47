// fyn_comm.cu -- multi-GPU entry points of the C ABI: one process per GPU.
//
// The reference is a single-GPU, batch-1 engine (README.md:72); the two workloads that shard naturally are
//   * ResNet-50 batch shards: contiguous image ranges per rank, one all-gather of the [B/world, 1000] logits
//     (DeepGEMMLayer output, fyusenet/gpu/deep/deepgemmlayer.cpp:66-140 -> DeepDownloadLayer, deepdownloadlayer.cpp:136-160);
//   * StyleNet row bands: every layer needs a few rows of its band neighbours' output (tap geometry of
//     fyusenet/gpu/vanilla/convlayerbase_vanilla.cpp:347-371, fractionalconvlayerNxN_vanilla.cpp:43-51).
//
// Plumbing: NCCL (dlopen'ed -- inside a torch process this resolves to the NCCL torch has already loaded) carries the
// bootstrap (the caller distributes the 128-byte unique id any way it likes) and the logit all-gather.  The halo exchange does
// NOT go through NCCL: tensors are registered once, their CUDA IPC handles travel over the communicator, and every exchange
// is ONE kernel that stores the band-edge rows straight into the neighbours' tensor memory over NVLink (peer stores, 16 bytes
// per thread) and hand-shakes through flag words in peer memory:
//     post ready[seq] to both neighbours   ("my margin rows of this tensor may be overwritten": in stream order every
//                                            earlier reader of the tensor on this rank has finished)
//     wait ready[seq] from both            (write-after-read safety on the neighbours)
//     push rows, __threadfence_system()
//     last CTA: post arrived[seq] to both neighbours, then wait for arrived[seq] from both (the margins are complete when
//     the kernel ends, so plain stream order protects the next layer).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "fyn_internal.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi *nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

#define FYN_NCCL(expr)                                                                                       \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess) {                                                                             \
            fyn_set_error("%s failed: %s (%s:%d)", #expr, nccl()->GetErrorString(_r), __FILE__, __LINE__);   \
            return FYN_ERR_CUDA;                                                                             \
        }                                                                                                    \
    } while (0)

constexpr int kMaxHaloTensors = 64;

// flag words every rank exposes to its band neighbours (device memory, CUDA IPC); [0] = written by the rank above (rank - 1),
// [1] = written by the rank below (rank + 1)
struct HaloFlags {
    unsigned ready[2];
    unsigned arrived[2];
    unsigned done;         // CTAs of the running exchange kernel that have finished pushing
    unsigned pad[3];
};

struct Wire {              // what every rank publishes per registered tensor (and once for its flag block)
    cudaIpcMemHandle_t mem;
    fyn_tensor_desc desc;
    int valid;
    int pad[3];
};

struct HaloTensor {
    fyn_tensor *mine = nullptr;
    void *peer[2] = {nullptr, nullptr};          // mapped tensor memory of rank - 1 / rank + 1
    fyn_tensor_desc peerDesc[2]{};
    fyn_tensor_geom peerGeom[2]{};
};

}  // namespace

struct fyn_comm {
    fyn_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    ncclComm_t nccl = nullptr;
    float *logitStage = nullptr;
    size_t logitStageFloats = 0;
    Wire *d_wire = nullptr;                      // [world + 1] device scratch for the handle all-gathers
    HaloFlags *flags = nullptr;                  // mine
    HaloFlags *peerFlags[2] = {nullptr, nullptr};
    HaloTensor tensors[kMaxHaloTensors];
    unsigned seq = 0;                            // exchanges issued so far (identical on all ranks)
    uint64_t halo_bytes = 0;                     // bytes this rank has pushed to its neighbours
};

namespace {

// all-gather of one Wire per rank through the communicator (tiny, blocking; set-up path only)
int gather_wires(fyn_comm *c, const Wire &mine, std::vector<Wire> &all) {
    all.assign(c->world, Wire{});
    if (c->world == 1) {
        all[0] = mine;
        return FYN_OK;
    }
    FYN_CUDA(cudaMemcpy(c->d_wire + c->world, &mine, sizeof(Wire), cudaMemcpyHostToDevice));
    FYN_NCCL(nccl()->AllGather(c->d_wire + c->world, c->d_wire, sizeof(Wire), ncclChar, c->nccl, (cudaStream_t)0));
    FYN_CUDA(cudaStreamSynchronize((cudaStream_t)0));
    FYN_CUDA(cudaMemcpy(all.data(), c->d_wire, sizeof(Wire) * c->world, cudaMemcpyDeviceToHost));
    return FYN_OK;
}

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// sequence numbers wrap: "a has reached b"
__device__ __forceinline__ bool reached(unsigned a, unsigned b) { return (int)(a - b) >= 0; }

// deep 1x1-spatial logits tensor [batch][TH][TW][4] (fp16 or fp32) -> float32 [rows][C], rows >= batch zero-filled
__global__ void k_logits_to_f32(TView t, int C, int batch, int rows, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const int n = i / C, c = i - n * C;
    float v = 0.f;
    if (n < batch) {
        const float4 tx = fyn_fetch(t, n, c >> 2, t.P, t.P);
        v = (c & 3) == 0 ? tx.x : ((c & 3) == 1 ? tx.y : ((c & 3) == 2 ? tx.z : tx.w));
    }
    out[i] = v;
}

struct HaloArgs {
    const uint4 *src;            // my tensor
    uint4 *dst[2];               // neighbours' tensors (NULL = no neighbour on that side)
    long long srcPlane, dstPlane[2];      // plane strides in 16-byte units
    long long srcRow[2], dstRow[2];       // first source / destination texture row per side, in 16-byte units from the plane start
    int run16;                   // 16-byte units per side and plane (rows * row pitch: the rows are contiguous)
    int planes;                  // planes x batch images (image stride == planes * plane stride)
    HaloFlags *mine, *peer[2];
    unsigned seq;
};

__global__ void __launch_bounds__(256) k_halo_exchange(const HaloArgs a) {
    // 1. hand-shake: my margins may be overwritten / wait until the neighbours' may
    if (threadIdx.x < 2 && a.peer[threadIdx.x]) {
        const int side = threadIdx.x;
        if (blockIdx.x == 0) st_release_sys(&a.peer[side]->ready[side ^ 1], a.seq);   // I am the neighbour's other side
        while (!reached(ld_acquire_sys(&a.mine->ready[side]), a.seq)) __nanosleep(64);
    }
    __syncthreads();
    // 2. push: per side, `planes` runs of `run16` 16-byte units
    const long long total = (long long)a.run16 * a.planes;
    for (int side = 0; side < 2; side++) {
        if (!a.dst[side]) continue;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long p = i / a.run16, o = i - p * a.run16;
            a.dst[side][p * a.dstPlane[side] + a.dstRow[side] + o] = a.src[p * a.srcPlane + a.srcRow[side] + o];
        }
    }
    __threadfence_system();
    __syncthreads();
    // 3. the last CTA publishes "arrived" and waits for the neighbours' data
    __shared__ bool last;
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&a.mine->done, 1u);
        last = prev + 1 == gridDim.x;
        if (last) a.mine->done = 0;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < 2 && a.peer[threadIdx.x]) {
        const int side = threadIdx.x;
        __threadfence_system();
        st_release_sys(&a.peer[side]->arrived[side ^ 1], a.seq);
        while (!reached(ld_acquire_sys(&a.mine->arrived[side]), a.seq)) __nanosleep(64);
    }
}

}  // namespace

extern "C" {

int fyn_comm_unique_id(void *id128) {
    if (!id128) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    if (!nccl()) FYN_FAIL(FYN_ERR_UNSUPPORTED, "NCCL library (libnccl.so.2) not found");
    static_assert(sizeof(ncclUniqueId) == FYN_COMM_ID_BYTES, "unique id size");
    ncclUniqueId id;
    FYN_NCCL(nccl()->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return FYN_OK;
}

int fyn_comm_init(fyn_ctx *ctx, int rank, int world, const void *id128, fyn_comm **out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) FYN_FAIL(FYN_ERR_INVALID, "comm: bad argument (rank %d of %d)", rank, world);
    *out = nullptr;
    FYN_CUDA(cudaSetDevice(ctx->device));
    fyn_comm *c = new fyn_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    auto fail = [&](int rc) {
        fyn_comm_destroy(c);
        return rc;
    };
    if (cudaMalloc((void **)&c->flags, sizeof(HaloFlags)) != cudaSuccess || cudaMemset(c->flags, 0, sizeof(HaloFlags)) != cudaSuccess) {
        fyn_set_error("comm: cannot allocate the flag block");
        return fail(FYN_ERR_NOMEM);
    }
    if (world > 1) {
        if (!id128) {
            fyn_set_error("comm: unique id is NULL");
            return fail(FYN_ERR_INVALID);
        }
        if (!nccl()) {
            fyn_set_error("NCCL library (libnccl.so.2) not found");
            return fail(FYN_ERR_UNSUPPORTED);
        }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = nccl()->CommInitRank(&c->nccl, world, id, rank);
        if (r != ncclSuccess) {
            fyn_set_error("ncclCommInitRank failed: %s", nccl()->GetErrorString(r));
            c->nccl = nullptr;
            return fail(FYN_ERR_CUDA);
        }
        if (cudaMalloc((void **)&c->d_wire, sizeof(Wire) * (world + 1)) != cudaSuccess) {
            fyn_set_error("comm: cannot allocate the handle scratch");
            return fail(FYN_ERR_NOMEM);
        }
        // flag blocks of the band neighbours
        Wire mine{};
        mine.valid = 1;
        if (cudaIpcGetMemHandle(&mine.mem, c->flags) != cudaSuccess) {
            fyn_set_error("cudaIpcGetMemHandle(flags) failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(FYN_ERR_CUDA);
        }
        std::vector<Wire> all;
        if (int rc = gather_wires(c, mine, all)) return fail(rc);
        for (int side = 0; side < 2; side++) {
            const int peer = rank + (side == 0 ? -1 : 1);
            if (peer < 0 || peer >= world) continue;
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[peer].mem, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                fyn_set_error("cudaIpcOpenMemHandle(flags of rank %d) failed: %s (peer access over NVLink / PCIe is required)", peer, cudaGetErrorString(e));
                return fail(FYN_ERR_CUDA);
            }
            c->peerFlags[side] = static_cast<HaloFlags *>(p);
        }
    }
    *out = c;
    return FYN_OK;
}

int fyn_comm_destroy(fyn_comm *c) {
    if (!c) return FYN_OK;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (HaloTensor &t : c->tensors)
        for (void *p : t.peer)
            if (p) cudaIpcCloseMemHandle(p);
    for (HaloFlags *p : c->peerFlags)
        if (p) cudaIpcCloseMemHandle(p);
    if (c->flags) cudaFree(c->flags);
    if (c->d_wire) cudaFree(c->d_wire);
    if (c->logitStage) cudaFree(c->logitStage);
    if (c->nccl && nccl()) nccl()->CommDestroy(c->nccl);
    delete c;
    return FYN_OK;
}

int fyn_comm_info(const fyn_comm *c, int *rank, int *world, int *nccl_version, uint64_t *halo_bytes_pushed) {
    if (!c) FYN_FAIL(FYN_ERR_INVALID, "comm is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) {
        *nccl_version = 0;
        if (nccl() && nccl()->GetVersion) nccl()->GetVersion(nccl_version);
    }
    if (halo_bytes_pushed) *halo_bytes_pushed = c->halo_bytes;
    return FYN_OK;
}

int fyn_allgather_logits(fyn_comm *c, const fyn_tensor *logits, int images_per_rank, float *device_out, void *stream) {
    if (!c || !logits || !device_out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const fyn_tensor_desc &d = logits->desc;
    if (d.width != 1 || d.height != 1) FYN_FAIL(FYN_ERR_INVALID, "allgather_logits: tensor must be 1x1 spatial (got %dx%d)", d.width, d.height);
    if (images_per_rank < d.batch) FYN_FAIL(FYN_ERR_INVALID, "allgather_logits: %d images per rank < batch %d", images_per_rank, d.batch);
    FYN_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int C = d.channels, n = images_per_rank * C;
    float *stage = device_out;
    if (c->world > 1) {
        if (c->logitStageFloats < (size_t)n) {
            if (c->logitStage) cudaFree(c->logitStage);
            c->logitStage = nullptr;
            FYN_CUDA(cudaMalloc((void **)&c->logitStage, (size_t)n * sizeof(float)));
            c->logitStageFloats = (size_t)n;
        }
        stage = c->logitStage;
    }
    k_logits_to_f32<<<(n + 255) / 256, 256, 0, s>>>(fyn_make_view(logits), C, d.batch, images_per_rank, stage);
    FYN_CHECK_LAUNCH(c->ctx);
    if (c->world > 1) FYN_NCCL(nccl()->AllGather(stage, device_out, (size_t)n, ncclFloat, c->nccl, s));
    return FYN_OK;
}

int fyn_comm_register_tensor(fyn_comm *c, fyn_tensor *t, int *slot) {
    if (!c || !t || !slot) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *slot = -1;
    for (int i = 0; i < kMaxHaloTensors; i++)
        if (c->tensors[i].mine == t) {
            *slot = i;
            return FYN_OK;
        }
    int s = 0;
    while (s < kMaxHaloTensors && c->tensors[s].mine) s++;
    if (s == kMaxHaloTensors) FYN_FAIL(FYN_ERR_NOMEM, "comm: more than %d registered tensors", kMaxHaloTensors);
    if (t->desc.order != FYN_ORDER_SHALLOW || t->geom.packing != 4) FYN_FAIL(FYN_ERR_UNSUPPORTED, "halo exchange needs shallow RGBA tensors");
    FYN_CUDA(cudaSetDevice(c->ctx->device));
    HaloTensor &h = c->tensors[s];
    if (c->world > 1) {
        if (!t->owns) FYN_FAIL(FYN_ERR_UNSUPPORTED, "halo exchange needs tensors allocated by this library (CUDA IPC on the allocation base)");
        Wire mine{};
        mine.valid = 1;
        mine.desc = t->desc;
        FYN_CUDA(cudaIpcGetMemHandle(&mine.mem, t->dptr));
        std::vector<Wire> all;
        if (int rc = gather_wires(c, mine, all)) return rc;
        for (int side = 0; side < 2; side++) {
            const int peer = c->rank + (side == 0 ? -1 : 1);
            if (peer < 0 || peer >= c->world) continue;
            const fyn_tensor_desc &pd = all[peer].desc;
            if (pd.width != t->desc.width || pd.channels != t->desc.channels || pd.padding != t->desc.padding || pd.dtype != t->desc.dtype ||
                pd.batch != t->desc.batch || pd.order != t->desc.order)
                FYN_FAIL(FYN_ERR_INVALID, "halo exchange: rank %d registered a different tensor shape in this slot", peer);
            FYN_CUDA(cudaIpcOpenMemHandle(&h.peer[side], all[peer].mem, cudaIpcMemLazyEnablePeerAccess));
            h.peerDesc[side] = pd;
            if (int rc = fyn_tensor_geometry(&pd, &h.peerGeom[side])) return rc;
        }
    }
    h.mine = t;
    *slot = s;
    return FYN_OK;
}

int fyn_halo_exchange(fyn_comm *c, int slot, int rows, void *stream) {
    if (!c || slot < 0 || slot >= kMaxHaloTensors || !c->tensors[slot].mine) FYN_FAIL(FYN_ERR_INVALID, "halo exchange: bad slot %d", slot);
    c->seq++;                                          // every rank counts every exchange, also the ones without work
    if (c->world == 1 || rows <= 0) return FYN_OK;
    HaloTensor &h = c->tensors[slot];
    const fyn_tensor *t = h.mine;
    const int P = t->desc.padding, H = t->desc.height;
    const bool up = c->rank > 0, down = c->rank + 1 < c->world;
    const int mt = up ? rows : 0, mb = down ? rows : 0;
    if (H - mt - mb < rows) FYN_FAIL(FYN_ERR_INVALID, "halo exchange: band of %d rows is thinner than the margin %d", H - mt - mb, rows);
    const size_t esz = t->geom.elem_size;
    const long long pitch16 = (long long)t->geom.tex_width * 4 * esz;     // bytes per texture row
    if (pitch16 % 16 || (t->geom.plane_elems * esz) % 16 || ((uintptr_t)t->dptr & 15))
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "halo exchange: rows are not 16-byte aligned (texture width %d)", t->geom.tex_width);
    FYN_CUDA(cudaSetDevice(c->ctx->device));
    HaloArgs a{};
    a.src = static_cast<const uint4 *>(t->dptr);
    a.srcPlane = (long long)(t->geom.plane_elems * esz / 16);
    a.run16 = (int)(rows * pitch16 / 16);
    a.planes = t->geom.planes * t->desc.batch;
    a.mine = c->flags;
    a.seq = c->seq;
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? up : down)) continue;
        const fyn_tensor_geom &pg = h.peerGeom[side];
        const int pH = h.peerDesc[side].height;
        if ((pg.plane_elems * esz) % 16) FYN_FAIL(FYN_ERR_UNSUPPORTED, "halo exchange: neighbour planes are not 16-byte aligned");
        a.dst[side] = static_cast<uint4 *>(h.peer[side]);
        a.dstPlane[side] = (long long)(pg.plane_elems * esz / 16);
        a.peer[side] = c->peerFlags[side];
        // side 0 (rank above): my first band rows -> its bottom margin; side 1 (rank below): my last band rows -> its top margin
        const int srcRow = side == 0 ? P + mt : P + H - mb - rows;
        const int dstRow = side == 0 ? P + pH - rows : P;
        a.srcRow[side] = srcRow * pitch16 / 16;
        a.dstRow[side] = dstRow * pitch16 / 16;
        c->halo_bytes += (uint64_t)a.run16 * 16 * a.planes;
    }
    const long long total = (long long)a.run16 * a.planes;
    const int grid = (int)std::max<long long>(1, std::min<long long>(64, (total + 1023) / 1024));
    k_halo_exchange<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    FYN_CHECK_LAUNCH(c->ctx);
    return FYN_OK;
}

}  // extern "C"
