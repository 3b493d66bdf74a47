// fyn_internal.h -- shared declarations of the CUDA backend (not part of the public C ABI).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdlib>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fyusenet_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void fyn_set_error(const char *fmt, ...);

#define FYN_FAIL(code, ...)         \
    do {                            \
        fyn_set_error(__VA_ARGS__); \
        return (code);              \
    } while (0)

#define FYN_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fyn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FYN_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// Launch with programmatic dependent launch allowed: the kernel may become resident while its predecessor in the stream still runs;
// it must execute griddepcontrol.wait before its first global-memory access (FYN_PDL_PROLOGUE at the top of the kernel).
#ifdef __CUDACC__
#define FYN_PDL_PROLOGUE()                                              \
    do {                                                                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
        asm volatile("griddepcontrol.wait;" ::: "memory");              \
    } while (0)
template <typename... ExpTypes, typename... ActTypes>
inline cudaError_t fyn_launch_pdl(void (*kernel)(ExpTypes...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, ActTypes &&...args) {
    static const bool noPdl = getenv("FYN_TC_NO_PDL") != nullptr;   // debugging aid: plain stream-ordered launches
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = noPdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<ExpTypes>(args)...);
}
#endif

#define FYN_CHECK_LAUNCH(ctx)                                                                  \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            fyn_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FYN_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------
// host objects
// ---------------------------------------------------------------------------------------------
struct fyn_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    uint64_t launches = 0;
};

struct fyn_tensor {
    fyn_ctx *ctx = nullptr;
    fyn_tensor_desc desc{};
    fyn_tensor_geom geom{};
    void *dptr = nullptr;
    bool owns = false;
    void *staging = nullptr;  // float32 staging for up/download conversions (lazily allocated)
    size_t staging_bytes = 0;
};

enum fyn_op_kind { FYN_OP_CONV = 1, FYN_OP_POOL, FYN_OP_BN, FYN_OP_SIGMOID, FYN_OP_SCALE, FYN_OP_ARITH, FYN_OP_CONCAT, FYN_OP_RGB2BGR, FYN_OP_RELAYOUT, FYN_OP_DWCONV, FYN_OP_TRANSCONV };

// Device-side view of a tensor: everything a kernel needs to address texels.
struct TView {
    void *ptr;
    int dtype;       // fyn_dtype
    int packing;     // elements per texel
    int deep;        // 1 = tiled texture
    int texW, texH;  // clamp bounds
    int W, H, P;
    int tx;          // tiles per texture row (deep)
    int tileRows;    // rows of tiles (deep)
    int tileW, tileH;  // W+P, H+P (deep)
    long long planeElems, imageElems;
};

TView fyn_make_view(const fyn_tensor *t);

struct ActParams {
    int type;  // 0 none, 1 relu, 2 leaky, 3 clip
    float leak, lo, hi;
};

ActParams fyn_act_from_flags(unsigned flags, float leaky, float lo, float hi);

struct ConvTcPlan;  // tcgen05 plan (fyn_conv_tc.cu)
struct DeepTcPlan;  // tcgen05 plan of the deep-tiled family (fyn_conv_deep_tc.cu)

struct fyn_op {
    fyn_ctx *ctx = nullptr;
    int kind = 0;
    // conv
    fyn_conv_desc conv{};
    int Wo = 0, Ho = 0;
    float *d_w = nullptr;      // direct-kernel weights [nOut][nIn][K][K][4ci][4co] fp32
    float *d_bias = nullptr;   // [nOut*4] folded bias
    float *d_scale = nullptr;  // [nOut*4] BN scale (1 without post-BN)
    int backend = 0;           // 1 direct, 2 tcgen05
    int lastKernel = 0;        // fyn_conv2d_last_kernel
    int epilogue = 0;          // FYN_EPILOGUE_*: element-wise function fused behind the convolution
    float *d_innorm = nullptr; // fused input batch-norm: scale[Cin4], bias[Cin4] (fyn_conv2d_set_input_norm)
    int innorm = 0;
    ConvTcPlan *tc = nullptr;
    DeepTcPlan *dtc = nullptr;
    // pool
    fyn_pool_desc pool{};
    // bn
    fyn_bn_desc bn{};
    // unary
    fyn_unary_desc unary{};
    // scale / arithmetic / concat (fyn_gather.cu)
    fyn_scale_desc scale{};
    fyn_arith_desc arith{};
    fyn_concat_desc concat{};
    fyn_dwconv_desc dw{};
    fyn_transconv_desc tconv{};
};

// tcgen05 path (fyn_conv_tc.cu)
int fyn_conv_tc_supported(const fyn_conv_desc *d, int dtype_hint);
int fyn_conv_tc_create(fyn_op *op, const float *wb);
int fyn_conv_tc_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t s);
void fyn_conv_tc_destroy(fyn_op *op);

// tcgen05 path for deep-tiled tensors (fyn_conv_deep_tc.cu)
int fyn_conv_deep_tc_supported(const fyn_conv_desc *d);
int fyn_conv_deep_tc_create(fyn_op *op, const float *wb);
int fyn_conv_deep_tc_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t s);
void fyn_conv_deep_tc_destroy(fyn_op *op);

// direct path (fyn_conv_direct.cu)
int fyn_conv_direct_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t s);

// host fp16 helpers
float fyn_half_trunc_host(float x);   // gpu/floatconversion.cpp:44-58 semantics
float fyn_half_round_host(float x);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fyn_act(float v, const ActParams &a) {
    // gpu/shaders/activation.inc:3-18
    if (a.type == 1) return fmaxf(v, 0.f);
    if (a.type == 2) return v >= 0.f ? v : a.leak * v;
    if (a.type == 3) return fminf(a.hi, fmaxf(a.lo, v));
    return v;
}

// shaders/sigmoid.frag:10-13
// (fast reciprocal: 2 ulp in fp32, far below the fp16 / 1e-5 parity tolerances; shared by the sigmoid layer and the
// fused conv epilogue so that both give identical values)
__device__ __forceinline__ float fyn_sigmoid(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
__device__ __forceinline__ float fyn_round_half(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ float4 fyn_act4(float4 v, const ActParams &a) {
    return make_float4(fyn_act(v.x, a), fyn_act(v.y, a), fyn_act(v.z, a), fyn_act(v.w, a));
}

// texel address (element index) of plane/tile `pt`, texel (x,y) given in plane coordinates
// (shallow) or tile-local coordinates whose origin is the tile's top-left padding texel (deep).
// CLAMP_TO_EDGE over the whole texture (base/buffermanager.cpp:657-670).
__device__ __forceinline__ long long fyn_texel_index(const TView &v, int n, int pt, int x, int y) {
    long long base = (long long)n * v.imageElems;
    if (v.deep) {
        x += (pt % v.tx) * v.tileW;
        y += (pt / v.tx) * v.tileH;
    } else {
        base += (long long)pt * v.planeElems;
    }
    x = min(max(x, 0), v.texW - 1);
    y = min(max(y, 0), v.texH - 1);
    return base + ((long long)y * v.texW + x) * v.packing;
}

// texel load; packing < 4 (upload textures) reads the stored lanes only, missing lanes are 0.  Written without
// dynamically indexed temporaries so everything stays in registers.
__device__ __forceinline__ float4 fyn_load_texel(const TView &v, long long idx) {
    if (v.dtype == FYN_F16) {
        const __half *p = reinterpret_cast<const __half *>(v.ptr) + idx;
        if (v.packing == 4) {
            uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
            __half2 a = *reinterpret_cast<__half2 *>(&raw.x), b = *reinterpret_cast<__half2 *>(&raw.y);
            float2 fa = __half22float2(a), fb = __half22float2(b);
            return make_float4(fa.x, fa.y, fb.x, fb.y);
        }
        const float x = __half2float(p[0]);
        const float y = v.packing > 1 ? __half2float(p[1]) : 0.f;
        const float z = v.packing > 2 ? __half2float(p[2]) : 0.f;
        return make_float4(x, y, z, 0.f);
    }
    const float *p = reinterpret_cast<const float *>(v.ptr) + idx;
    if (v.packing == 4) return __ldg(reinterpret_cast<const float4 *>(p));
    const float x = __ldg(p);
    const float y = v.packing > 1 ? __ldg(p + 1) : 0.f;
    const float z = v.packing > 2 ? __ldg(p + 2) : 0.f;
    return make_float4(x, y, z, 0.f);
}

__device__ __forceinline__ float4 fyn_fetch(const TView &v, int n, int pt, int x, int y) {
    return fyn_load_texel(v, fyn_texel_index(v, n, pt, x, y));
}

// store a texel (outputs always have packing 4); no clamping: caller guarantees in-range
__device__ __forceinline__ void fyn_store_texel(const TView &v, int n, int pt, int x, int y, float4 val) {
    long long base = (long long)n * v.imageElems;
    if (v.deep) {
        x += (pt % v.tx) * v.tileW;
        y += (pt / v.tx) * v.tileH;
    } else {
        base += (long long)pt * v.planeElems;
    }
    long long idx = base + ((long long)y * v.texW + x) * 4;
    if (v.dtype == FYN_F16) {
        __half2 a = __floats2half2_rn(val.x, val.y), b = __floats2half2_rn(val.z, val.w);
        uint2 raw;
        raw.x = *reinterpret_cast<unsigned *>(&a);
        raw.y = *reinterpret_cast<unsigned *>(&b);
        *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(v.ptr) + idx) = raw;
    } else {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(v.ptr) + idx) = val;
    }
}
#endif
