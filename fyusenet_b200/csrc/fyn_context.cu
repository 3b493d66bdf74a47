// fyn_context.cu -- device context, streams, events, pinned memory.
// Replaces the roles of GfxContextManager / GfxContextLink (fyusenet/gpu/gfxcontextmanager.h),
// GLsync fences (fyusenet/base/engine.cpp:779-780) and PBOPool (fyusenet/gl/pbopool.cpp).
#include <algorithm>
#include <cstring>

#include "fyn_internal.h"

static thread_local char g_err[1024] = "";

void fyn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int fyn_abi_version(void) { return FYN_ABI_VERSION; }

const char *fyn_last_error(void) { return g_err; }

int fyn_device_count(int *count) {
    if (!count) FYN_FAIL(FYN_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        FYN_FAIL(FYN_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return FYN_OK;
}

int fyn_cuda_init(int device, fyn_ctx **out) {
    if (!out) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        FYN_FAIL(FYN_ERR_NODEVICE, "no CUDA device available (%s); this backend has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) FYN_FAIL(FYN_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    FYN_CUDA(cudaSetDevice(device));
    fyn_ctx *c = new fyn_ctx();
    c->device = device;
    e = cudaGetDeviceProperties(&c->prop, device);
    if (e != cudaSuccess) {
        delete c;
        FYN_FAIL(FYN_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (c->prop.major < 10) {
        int maj = c->prop.major, min = c->prop.minor;
        delete c;
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", maj, min);
    }
    FYN_CUDA(cudaFree(0));
    *out = c;
    return FYN_OK;
}

int fyn_cuda_shutdown(fyn_ctx *ctx) {
    if (!ctx) return FYN_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    delete ctx;
    return FYN_OK;
}

int fyn_get_device_info(fyn_ctx *ctx, fyn_device_info *info) {
    if (!ctx || !info) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    memset(info, 0, sizeof(*info));
    info->device = ctx->device;
    info->sm_count = ctx->prop.multiProcessorCount;
    info->cc_major = ctx->prop.major;
    info->cc_minor = ctx->prop.minor;
    info->total_mem = ctx->prop.totalGlobalMem;
    info->smem_per_block_optin = ctx->prop.sharedMemPerBlockOptin;
    memcpy(info->name, ctx->prop.name, std::min(sizeof(info->name) - 1, strlen(ctx->prop.name)));   // (memset above keeps it terminated)
    return FYN_OK;
}

int fyn_launch_count(fyn_ctx *ctx, uint64_t *count, int reset) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (count) *count = ctx->launches;
    if (reset) ctx->launches = 0;
    return FYN_OK;
}

int fyn_stream_create(fyn_ctx *ctx, void **stream) {
    if (!ctx || !stream) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s;
    FYN_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return FYN_OK;
}

int fyn_stream_destroy(fyn_ctx *ctx, void *stream) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (stream) FYN_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return FYN_OK;
}

int fyn_stream_sync(fyn_ctx *ctx, void *stream) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    FYN_CUDA(cudaSetDevice(ctx->device));
    FYN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FYN_OK;
}

int fyn_event_create(fyn_ctx *ctx, void **event) {
    if (!ctx || !event) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaSetDevice(ctx->device));
    cudaEvent_t e;
    FYN_CUDA(cudaEventCreate(&e));
    *event = (void *)e;
    return FYN_OK;
}

int fyn_event_destroy(fyn_ctx *ctx, void *event) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (event) FYN_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return FYN_OK;
}

int fyn_event_record(fyn_ctx *ctx, void *event, void *stream) {
    if (!ctx || !event) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return FYN_OK;
}

int fyn_event_sync(fyn_ctx *ctx, void *event) {
    if (!ctx || !event) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return FYN_OK;
}

int fyn_event_elapsed_ms(fyn_ctx *ctx, void *start, void *stop, float *ms) {
    if (!ctx || !start || !stop || !ms) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return FYN_OK;
}

int fyn_stream_wait_event(fyn_ctx *ctx, void *stream, void *event) {
    if (!ctx || !event) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return FYN_OK;
}

int fyn_stream_add_callback(fyn_ctx *ctx, void *stream, fyn_host_fn fn, void *user) {
    if (!ctx || !fn) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaLaunchHostFunc((cudaStream_t)stream, (cudaHostFn_t)fn, user));
    return FYN_OK;
}

int fyn_graph_begin_capture(fyn_ctx *ctx, void *stream) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    FYN_CUDA(cudaSetDevice(ctx->device));
    FYN_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return FYN_OK;
}

int fyn_graph_end_capture(fyn_ctx *ctx, void *stream, void **graph_exec) {
    if (!ctx || !graph_exec) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    FYN_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) FYN_FAIL(FYN_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    *graph_exec = (void *)exec;
    return FYN_OK;
}

int fyn_graph_launch(fyn_ctx *ctx, void *graph_exec, void *stream) {
    if (!ctx || !graph_exec) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return FYN_OK;
}

int fyn_graph_destroy(fyn_ctx *ctx, void *graph_exec) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (graph_exec) FYN_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return FYN_OK;
}

int fyn_device_alloc(fyn_ctx *ctx, size_t bytes, void **ptr) {
    if (!ctx || !ptr) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaSetDevice(ctx->device));
    FYN_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return FYN_OK;
}

int fyn_device_free(fyn_ctx *ctx, void *ptr) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (ptr) FYN_CUDA(cudaFree(ptr));
    return FYN_OK;
}

int fyn_memcpy_async(fyn_ctx *ctx, void *dst, const void *src, size_t bytes, int device_to_host, void *stream) {
    if (!ctx || !dst || !src) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaMemcpyAsync(dst, src, bytes, device_to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return FYN_OK;
}

int fyn_host_alloc(fyn_ctx *ctx, size_t bytes, void **ptr) {
    if (!ctx || !ptr) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    FYN_CUDA(cudaSetDevice(ctx->device));
    FYN_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return FYN_OK;
}

int fyn_host_free(fyn_ctx *ctx, void *ptr) {
    if (!ctx) FYN_FAIL(FYN_ERR_INVALID, "ctx is NULL");
    if (ptr) FYN_CUDA(cudaFreeHost(ptr));
    return FYN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// host fp16 helpers
// ---------------------------------------------------------------------------------------------
float fyn_half_round_host(float x) { return __half2float(__float2half_rn(x)); }

// Truncating conversion with the semantics of the reference's table-driven converter
// (fyusenet/gpu/floatconversion.cpp:44-58,85-127): mantissa bits are dropped, never rounded;
// |x| < 2^-24 -> +-0, exponent > 15 -> +-inf.  Expressed arithmetically instead of via tables.
float fyn_half_trunc_host(float x) {
    uint32_t f;
    memcpy(&f, &x, 4);
    uint32_t sign = (f >> 16) & 0x8000u;
    int e = (int)((f >> 23) & 0xff) - 127;
    uint32_t man = f & 0x007fffffu;
    uint16_t h;
    if (e < -24) h = (uint16_t)sign;
    else if (e < -14) h = (uint16_t)(sign | ((0x0400u >> (-e - 14)) + (man >> (-e - 1))));
    else if (e <= 15) h = (uint16_t)(sign | (((uint32_t)(e + 15) << 10) + (man >> 13)));
    else if (e < 128) h = (uint16_t)(sign | 0x7c00u);
    else h = (uint16_t)(sign | (0x7c00u + (man >> 13)));
    __half_raw r;
    r.x = h;
    return __half2float(__half(r));
}

ActParams fyn_act_from_flags(unsigned flags, float leaky, float lo, float hi) {
    // flag -> activation mapping of GPULayerBase::handlePreprocFlags (fyusenet/gpu/gpulayerbase.cpp:704-802)
    ActParams a{0, 0.f, 0.f, 0.f};
    if (flags & FYN_FLAG_PRE_RELU) {
        a.type = (leaky != 0.f) ? 2 : 1;
        a.leak = leaky;
    } else if (flags & FYN_FLAG_PRE_CLIP) {
        a.type = 3;
        a.lo = lo;
        a.hi = hi;
    }
    return a;
}
