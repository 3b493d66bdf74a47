// fyn_tensor.cu -- CUDA device-tensor manager primitives and host<->device I/O.
// Replaces BufferManager::createTexture (fyusenet/base/buffermanager.cpp:650-708), UploadLayer
// (fyusenet/gpu/uploadlayer.cpp:360-380), DownloadLayer / DeepDownloadLayer
// (fyusenet/gpu/downloadlayer.cpp:257-283, fyusenet/gpu/deep/deepdownloadlayer.cpp:136-160) and
// copyResult (fyusenet/gpu/gpulayerbase.cpp:525-560).  Layout contract: see fyusenet_b200.h.
#include <cstring>

#include "fyn_internal.h"

// deep tiling rule (cpu/cpubuffershape.cpp:430-447): minimise |x-y| + (x*y - tiles), first minimum
static void deep_tiling(int channels, int *tx, int *ty) {
    int tiles = (channels + 3) / 4;
    long best = -1;
    *tx = *ty = 1;
    for (int y = 1; y <= tiles; y++) {
        for (int x = y; x <= tiles; x++) {
            if (x * y < tiles) continue;
            long cost = (long)(x - y) + (long)(x * y - tiles);
            if (best < 0 || cost < best) {
                best = cost;
                *tx = x;
                *ty = y;
            }
            break;  // larger x in this row only costs more
        }
    }
}

extern "C" int fyn_tensor_geometry(const fyn_tensor_desc *d, fyn_tensor_geom *g) {
    if (!d || !g) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    if (d->width <= 0 || d->height <= 0 || d->channels <= 0 || d->padding < 0 || d->batch < 1)
        FYN_FAIL(FYN_ERR_INVALID, "bad tensor desc w=%d h=%d c=%d pad=%d batch=%d", d->width, d->height,
                 d->channels, d->padding, d->batch);
    if (d->dtype != FYN_F16 && d->dtype != FYN_F32) FYN_FAIL(FYN_ERR_INVALID, "bad dtype %d", d->dtype);
    memset(g, 0, sizeof(*g));
    g->elem_size = d->dtype == FYN_F16 ? 2 : 4;
    int packing = d->packing == 0 ? 4 : d->packing;
    if (packing < 1 || packing > 4) FYN_FAIL(FYN_ERR_INVALID, "bad packing %d", d->packing);
    if (packing != 4 && (d->order != FYN_ORDER_SHALLOW || d->channels > packing))
        FYN_FAIL(FYN_ERR_INVALID, "packing %d needs a shallow tensor with <= %d channels", packing, packing);
    g->packing = packing;
    if (d->order == FYN_ORDER_SHALLOW) {
        g->tex_width = d->width + 2 * d->padding;
        g->tex_height = d->height + 2 * d->padding;
        g->planes = (d->channels + 3) / 4;
        g->tiles_x = g->tiles_y = 1;
    } else if (d->order == FYN_ORDER_DEEP) {
        deep_tiling(d->channels, &g->tiles_x, &g->tiles_y);
        g->tex_width = g->tiles_x * (d->width + d->padding) + d->padding;
        g->tex_height = g->tiles_y * (d->height + d->padding) + d->padding;
        g->planes = 1;
    } else {
        FYN_FAIL(FYN_ERR_INVALID, "bad order %d", d->order);
    }
    g->plane_elems = (size_t)g->tex_width * g->tex_height * packing;
    g->image_elems = g->plane_elems * g->planes;
    g->bytes = g->image_elems * d->batch * g->elem_size;
    return FYN_OK;
}

TView fyn_make_view(const fyn_tensor *t) {
    TView v{};
    if (!t) return v;
    v.ptr = t->dptr;
    v.dtype = t->desc.dtype;
    v.packing = t->geom.packing;
    v.deep = t->desc.order == FYN_ORDER_DEEP;
    v.texW = t->geom.tex_width;
    v.texH = t->geom.tex_height;
    v.W = t->desc.width;
    v.H = t->desc.height;
    v.P = t->desc.padding;
    v.tx = t->geom.tiles_x;
    v.tileRows = t->geom.tiles_y;
    v.tileW = t->desc.width + t->desc.padding;
    v.tileH = t->desc.height + t->desc.padding;
    v.planeElems = (long long)t->geom.plane_elems;
    v.imageElems = (long long)t->geom.image_elems;
    return v;
}

static int tensor_new(fyn_ctx *ctx, const fyn_tensor_desc *desc, void *wrap, fyn_tensor **out) {
    if (!ctx || !desc || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    fyn_tensor_geom g;
    int rc = fyn_tensor_geometry(desc, &g);
    if (rc) return rc;
    FYN_CUDA(cudaSetDevice(ctx->device));
    fyn_tensor *t = new fyn_tensor();
    t->ctx = ctx;
    t->desc = *desc;
    t->desc.packing = g.packing;
    t->geom = g;
    if (wrap) {
        t->dptr = wrap;
        t->owns = false;
    } else {
        // 16 bytes of slack: the tcgen05 conv family stages rows with 16-byte-granular bulk copies
        cudaError_t e = cudaMalloc(&t->dptr, g.bytes + 16);
        if (e != cudaSuccess) {
            delete t;
            FYN_FAIL(FYN_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", g.bytes, cudaGetErrorString(e));
        }
        t->owns = true;
        e = cudaMemset(t->dptr, 0, g.bytes);
        if (e != cudaSuccess) {
            cudaFree(t->dptr);
            delete t;
            FYN_FAIL(FYN_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
        }
    }
    *out = t;
    return FYN_OK;
}

// ---------------------------------------------------------------------------------------------
// conversion kernels (bandwidth-bound, one texel per thread, 8/16-byte accesses)
// ---------------------------------------------------------------------------------------------

// host-order float32 [batch][H][W][C] (staged on the device) -> single plane/tile tensor interior
__global__ void k_upload_convert(const float *__restrict__ src, TView dst, int C, int batch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int n = blockIdx.z;
    if (x >= dst.W || n >= batch) return;
    const float *p = src + (((long long)n * dst.H + y) * dst.W + x) * C;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; c++) r[c] = p[c];
    long long idx = fyn_texel_index(dst, n, 0, dst.P + x, dst.P + y);
    if (dst.dtype == FYN_F16) {
        __half *o = reinterpret_cast<__half *>(dst.ptr) + idx;
        for (int c = 0; c < dst.packing; c++) o[c] = __float2half_rn(r[c]);
    } else {
        float *o = reinterpret_cast<float *>(dst.ptr) + idx;
        for (int c = 0; c < dst.packing; c++) o[c] = r[c];
    }
}

// whole tensor (including padding) -> float32 RGBA texels in the same texel order
__global__ void k_download_widen(TView src, float4 *__restrict__ dst, long long texels) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= texels) return;
    dst[i] = fyn_load_texel(src, i * src.packing);
}

// host-order uint8 [batch][H][W][C] (staged on the device) -> tensor interior, value / 255: the normalised 8-bit texture of an
// UBYTE upload (gpu/uploadlayer.cpp:51-66,365-375; the samples divide by 255 on the host instead, samples/desktop/stylenet.cpp:45-47)
__global__ void k_upload_convert_u8(const unsigned char *__restrict__ src, TView dst, int C, int batch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int n = blockIdx.z;
    if (x >= dst.W || n >= batch) return;
    const unsigned char *p = src + (((long long)n * dst.H + y) * dst.W + x) * C;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; c++) r[c] = __fdiv_rn((float)p[c], 255.f);
    long long idx = fyn_texel_index(dst, n, 0, dst.P + x, dst.P + y);
    if (dst.dtype == FYN_F16) {
        __half *o = reinterpret_cast<__half *>(dst.ptr) + idx;
        for (int c = 0; c < dst.packing; c++) o[c] = __float2half_rn(r[c]);
    } else {
        float *o = reinterpret_cast<float *>(dst.ptr) + idx;
        for (int c = 0; c < dst.packing; c++) o[c] = r[c];
    }
}

// RGB bytes -> the unpadded RGB32F upload texture, four pixels (12 bytes in, 48 bytes out) per thread; the texture is one
// contiguous [batch * H * W][3] array, W % 4 == 0
__global__ void k_upload_rgb8_to_rgb32f(const uint32_t *__restrict__ src, float4 *__restrict__ dst, long long quads) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const uint32_t w0 = __ldg(src + 3 * i), w1 = __ldg(src + 3 * i + 1), w2 = __ldg(src + 3 * i + 2);
    auto f = [](uint32_t w, int b) { return __fdiv_rn((float)((w >> (8 * b)) & 0xffu), 255.f); };
    dst[3 * i] = make_float4(f(w0, 0), f(w0, 1), f(w0, 2), f(w0, 3));
    dst[3 * i + 1] = make_float4(f(w1, 0), f(w1, 1), f(w1, 2), f(w1, 3));
    dst[3 * i + 2] = make_float4(f(w2, 0), f(w2, 1), f(w2, 2), f(w2, 3));
}

// whole tensor (including padding) -> 8-bit RGBA texels in the same texel order: (uint8)(clamp(v, 0, 1) * 255), the
// conversion of the samples' writeImage (samples/desktop/stylenet.cpp:52-62)
__device__ __forceinline__ uint32_t fyn_pack_rgba8(float4 v) {
    auto q = [](float x) { return __float2uint_rz(fminf(fmaxf(x, 0.f), 1.f) * 255.f); };
    return q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
}
__global__ void k_download_rgba8(TView src, uint32_t *__restrict__ dst, long long texels) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= texels) return;
    dst[i] = fyn_pack_rgba8(fyn_load_texel(src, i * src.packing));
}
// fp16 RGBA tensors: four texels (32 bytes in, 16 bytes out) per thread
__global__ void k_download_rgba8_h4(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long quads) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const uint4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
    auto tex = [](uint32_t lo, uint32_t hi) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&lo)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
        return fyn_pack_rgba8(make_float4(f0.x, f0.y, f1.x, f1.y));
    };
    dst[i] = make_uint4(tex(a.x, a.y), tex(a.z, a.w), tex(b.x, b.y), tex(b.z, b.w));
}

static int ensure_staging(fyn_tensor *t, size_t bytes) {
    if (t->staging_bytes >= bytes) return FYN_OK;
    if (t->staging) cudaFree(t->staging);
    t->staging = nullptr;
    t->staging_bytes = 0;
    cudaError_t e = cudaMalloc(&t->staging, bytes);
    if (e != cudaSuccess) FYN_FAIL(FYN_ERR_NOMEM, "cudaMalloc(staging %zu) failed: %s", bytes, cudaGetErrorString(e));
    t->staging_bytes = bytes;
    return FYN_OK;
}

extern "C" {

int fyn_tensor_create(fyn_ctx *ctx, const fyn_tensor_desc *desc, fyn_tensor **tensor) {
    return tensor_new(ctx, desc, nullptr, tensor);
}

int fyn_tensor_wrap(fyn_ctx *ctx, const fyn_tensor_desc *desc, void *device_ptr, fyn_tensor **tensor) {
    if (!device_ptr) FYN_FAIL(FYN_ERR_INVALID, "device_ptr is NULL");
    if (((uintptr_t)device_ptr) & 15) FYN_FAIL(FYN_ERR_INVALID, "device_ptr must be 16-byte aligned");
    return tensor_new(ctx, desc, device_ptr, tensor);
}

int fyn_tensor_destroy(fyn_tensor *t) {
    if (!t) return FYN_OK;
    cudaSetDevice(t->ctx->device);
    if (t->owns && t->dptr) cudaFree(t->dptr);
    if (t->staging) cudaFree(t->staging);
    delete t;
    return FYN_OK;
}

int fyn_tensor_clear(fyn_tensor *t, void *stream) {
    if (!t) FYN_FAIL(FYN_ERR_INVALID, "tensor is NULL");
    FYN_CUDA(cudaMemsetAsync(t->dptr, 0, t->geom.bytes, (cudaStream_t)stream));
    return FYN_OK;
}

int fyn_tensor_get_desc(const fyn_tensor *t, fyn_tensor_desc *desc, fyn_tensor_geom *geom) {
    if (!t) FYN_FAIL(FYN_ERR_INVALID, "tensor is NULL");
    if (desc) *desc = t->desc;
    if (geom) *geom = t->geom;
    return FYN_OK;
}

void *fyn_tensor_device_ptr(const fyn_tensor *t) { return t ? t->dptr : nullptr; }

int fyn_upload_f32_async(fyn_tensor *t, const float *host, void *stream) {
    if (!t || !host) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const fyn_tensor_desc &d = t->desc;
    if (d.channels > 4) FYN_FAIL(FYN_ERR_UNSUPPORTED, "upload supports <= 4 channels (got %d)", d.channels);
    cudaStream_t s = (cudaStream_t)stream;
    size_t n = (size_t)d.batch * d.height * d.width * d.channels;
    if (d.dtype == FYN_F32 && d.padding == 0 && t->geom.packing == d.channels) {
        // the reference's RGB32F upload texture: verbatim copy (gpu/uploadlayer.cpp:371-375)
        FYN_CUDA(cudaMemcpyAsync(t->dptr, host, n * sizeof(float), cudaMemcpyHostToDevice, s));
        return FYN_OK;
    }
    int rc = ensure_staging(t, n * sizeof(float));
    if (rc) return rc;
    FYN_CUDA(cudaMemcpyAsync(t->staging, host, n * sizeof(float), cudaMemcpyHostToDevice, s));
    dim3 block(128), grid((d.width + 127) / 128, d.height, d.batch);
    k_upload_convert<<<grid, block, 0, s>>>((const float *)t->staging, fyn_make_view(t), d.channels, d.batch);
    FYN_CHECK_LAUNCH(t->ctx);
    return FYN_OK;
}

int fyn_upload_u8_async(fyn_tensor *t, const unsigned char *host, void *stream) {
    if (!t || !host) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const fyn_tensor_desc &d = t->desc;
    if (d.channels > 4) FYN_FAIL(FYN_ERR_UNSUPPORTED, "upload supports <= 4 channels (got %d)", d.channels);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)d.batch * d.height * d.width * d.channels;
    int rc = ensure_staging(t, (n + 15) & ~(size_t)15);
    if (rc) return rc;
    FYN_CUDA(cudaMemcpyAsync(t->staging, host, n, cudaMemcpyHostToDevice, s));
    if (d.dtype == FYN_F32 && d.padding == 0 && t->geom.packing == 3 && d.channels == 3 && d.width % 4 == 0) {
        const long long quads = (long long)(n / 12);
        k_upload_rgb8_to_rgb32f<<<(unsigned)((quads + 255) / 256), 256, 0, s>>>((const uint32_t *)t->staging, (float4 *)t->dptr, quads);
    } else {
        dim3 block(128), grid((d.width + 127) / 128, d.height, d.batch);
        k_upload_convert_u8<<<grid, block, 0, s>>>((const unsigned char *)t->staging, fyn_make_view(t), d.channels, d.batch);
    }
    FYN_CHECK_LAUNCH(t->ctx);
    return FYN_OK;
}

size_t fyn_download_u8_bytes(const fyn_tensor *t) { return fyn_download_f32_elems(t); }   // one byte per element of the RGBA texels

int fyn_download_u8_convert(fyn_tensor *t, unsigned char *device_staging, void *stream) {
    if (!t || !device_staging) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long long texels = (long long)(fyn_download_f32_elems(t) / 4);
    if (t->desc.dtype == FYN_F16 && t->geom.packing == 4 && texels % 4 == 0 && (((uintptr_t)device_staging) & 15) == 0) {
        const long long quads = texels / 4;
        k_download_rgba8_h4<<<(unsigned)((quads + 255) / 256), 256, 0, s>>>((const uint4 *)t->dptr, (uint4 *)device_staging, quads);
    } else {
        k_download_rgba8<<<(unsigned)((texels + 255) / 256), 256, 0, s>>>(fyn_make_view(t), (uint32_t *)device_staging, texels);
    }
    FYN_CHECK_LAUNCH(t->ctx);
    return FYN_OK;
}

int fyn_download_u8_async(fyn_tensor *t, unsigned char *host, void *stream) {
    if (!t || !host) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const size_t bytes = fyn_download_u8_bytes(t);
    int rc = ensure_staging(t, bytes);
    if (rc) return rc;
    rc = fyn_download_u8_convert(t, (unsigned char *)t->staging, stream);
    if (rc) return rc;
    FYN_CUDA(cudaMemcpyAsync(host, t->staging, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return FYN_OK;
}

size_t fyn_download_f32_elems(const fyn_tensor *t) {
    if (!t) return 0;
    return (size_t)t->geom.tex_width * t->geom.tex_height * t->geom.planes * t->desc.batch * 4;
}

int fyn_download_convert(fyn_tensor *t, float *device_staging, void *stream) {
    if (!t || !device_staging) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t elems = fyn_download_f32_elems(t);
    if (t->desc.dtype == FYN_F32 && t->geom.packing == 4) {
        FYN_CUDA(cudaMemcpyAsync(device_staging, t->dptr, elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return FYN_OK;
    }
    const long long texels = (long long)(elems / 4);
    const int block = 256;
    const long long grid = (texels + block - 1) / block;
    k_download_widen<<<(unsigned)grid, block, 0, s>>>(fyn_make_view(t), (float4 *)device_staging, texels);
    FYN_CHECK_LAUNCH(t->ctx);
    return FYN_OK;
}

int fyn_download_f32_async(fyn_tensor *t, float *host, void *stream) {
    if (!t || !host) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    size_t elems = fyn_download_f32_elems(t);
    if (t->desc.dtype == FYN_F32 && t->geom.packing == 4) {
        FYN_CUDA(cudaMemcpyAsync(host, t->dptr, elems * sizeof(float), cudaMemcpyDeviceToHost, s));
        return FYN_OK;
    }
    int rc = ensure_staging(t, elems * sizeof(float));
    if (rc) return rc;
    long long texels = (long long)(elems / 4);
    int block = 256;
    long long grid = (texels + block - 1) / block;
    k_download_widen<<<(unsigned)grid, block, 0, s>>>(fyn_make_view(t), (float4 *)t->staging, texels);
    FYN_CHECK_LAUNCH(t->ctx);
    FYN_CUDA(cudaMemcpyAsync(host, t->staging, elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    return FYN_OK;
}

// CHW <-> layout conversion on the host (debug / parity path, blocking)
static size_t host_index(const fyn_tensor *t, int n, int c, int y, int x) {
    const fyn_tensor_desc &d = t->desc;
    const fyn_tensor_geom &g = t->geom;
    size_t base = (size_t)n * g.image_elems;
    int P = d.padding;
    if (d.order == FYN_ORDER_DEEP) {
        int tile = c / 4;
        int ox = P + (tile % g.tiles_x) * (d.width + P), oy = P + (tile / g.tiles_x) * (d.height + P);
        return base + ((size_t)(oy + y) * g.tex_width + ox + x) * 4 + (c % 4);
    }
    return base + (size_t)(c / 4) * g.plane_elems + ((size_t)(y + P) * g.tex_width + x + P) * g.packing + (c % 4);
}

int fyn_tensor_write_chw_f32(fyn_tensor *t, const float *chw) {
    if (!t || !chw) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const fyn_tensor_desc &d = t->desc;
    size_t total = t->geom.image_elems * d.batch;
    std::vector<unsigned char> buf(total * t->geom.elem_size, 0);
    for (int n = 0; n < d.batch; n++)
        for (int c = 0; c < d.channels; c++)
            for (int y = 0; y < d.height; y++)
                for (int x = 0; x < d.width; x++) {
                    float v = chw[(((size_t)n * d.channels + c) * d.height + y) * d.width + x];
                    size_t i = host_index(t, n, c, y, x);
                    if (d.dtype == FYN_F16) reinterpret_cast<__half *>(buf.data())[i] = __float2half_rn(v);
                    else reinterpret_cast<float *>(buf.data())[i] = v;
                }
    FYN_CUDA(cudaSetDevice(t->ctx->device));
    FYN_CUDA(cudaDeviceSynchronize());
    FYN_CUDA(cudaMemcpy(t->dptr, buf.data(), buf.size(), cudaMemcpyHostToDevice));
    return FYN_OK;
}

int fyn_tensor_read_chw_f32(fyn_tensor *t, float *chw) {
    if (!t || !chw) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    const fyn_tensor_desc &d = t->desc;
    size_t total = t->geom.image_elems * d.batch;
    std::vector<unsigned char> buf(total * t->geom.elem_size);
    FYN_CUDA(cudaSetDevice(t->ctx->device));
    FYN_CUDA(cudaDeviceSynchronize());
    FYN_CUDA(cudaMemcpy(buf.data(), t->dptr, buf.size(), cudaMemcpyDeviceToHost));
    for (int n = 0; n < d.batch; n++)
        for (int c = 0; c < d.channels; c++)
            for (int y = 0; y < d.height; y++)
                for (int x = 0; x < d.width; x++) {
                    size_t i = host_index(t, n, c, y, x);
                    float v = d.dtype == FYN_F16 ? __half2float(reinterpret_cast<__half *>(buf.data())[i])
                                                 : reinterpret_cast<float *>(buf.data())[i];
                    chw[(((size_t)n * d.channels + c) * d.height + y) * d.width + x] = v;
                }
    return FYN_OK;
}

}  // extern "C"
