// fyn_gather.cu -- the remaining bandwidth-bound layers of the reference's GPU layer set:
// scaling (also the PADDING2D / RELU / CLIP pseudo-layers), add / sub / singleton arithmetic, channel concatenation,
// RGB<->BGR swizzle and the shallow <-> deep layout conversion.
// One thread per output texel (4 channels); every source texel goes through fyn_fetch(), i.e. the clamp-to-edge
// sampler of the reference's textures (base/buffermanager.cpp:657-670) in the tensor's own (planar or tiled) layout.
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>

#include "fyn_internal.h"

namespace {

enum { G_SCALE = 0, G_ARITH, G_CONCAT, G_RGB2BGR, G_RELAYOUT };

struct GatherArgs {
    TView in[FYN_CONCAT_MAX_INPUTS];
    TView out;
    int mode;
    int Wo, Ho, tiles, batch, outP;
    // scale
    int W, H, linear;
    // arith
    int op, singleton;
    float operand;
    // concat: first output channel of every input, total channel count
    int nin, chOff[FYN_CONCAT_MAX_INPUTS + 1];
    ActParams act;
};

__device__ __forceinline__ float lane_of(const float4 &v, int l) { return l == 0 ? v.x : (l == 1 ? v.y : (l == 2 ? v.z : v.w)); }

// floor(num / den) for den > 0
__device__ __forceinline__ int floor_div(int num, int den) { return num >= 0 ? num / den : -((-num + den - 1) / den); }

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float f) {
    return make_float4(a.x + (b.x - a.x) * f, a.y + (b.y - a.y) * f, a.z + (b.z - a.z) * f, a.w + (b.w - a.w) * f);
}

__global__ void __launch_bounds__(128) k_gather(const __grid_constant__ GatherArgs a) {
    unsigned bid = blockIdx.x;
    const int xBlocks = (a.Wo + 31) / 32, yBlocks = (a.Ho + 3) / 4;
    const int xb = bid % xBlocks;
    bid /= xBlocks;
    const int yb = bid % yBlocks;
    bid /= yBlocks;
    const int t = bid % a.tiles;
    const int n = bid / a.tiles;
    const int xo = xb * 32 + threadIdx.x, yo = yb * 4 + threadIdx.y;
    if (xo >= a.Wo || yo >= a.Ho) return;
    const int P = a.in[0].P;
    float4 r;
    if (a.mode == G_SCALE) {
        if (!a.linear) {
            // texel containing P + (o + 0.5) * W / Wo (exact rational arithmetic)
            const int sx = ((2 * xo + 1) * a.W) / (2 * a.Wo), sy = ((2 * yo + 1) * a.H) / (2 * a.Ho);
            r = fyn_act4(fyn_fetch(a.in[0], n, t, P + sx, P + sy), a.act);
        } else {
            // GL_LINEAR: u = coordinate - 0.5, texels floor(u) and floor(u) + 1 weighted by frac(u); u * 2Wo is an integer
            const int nx = (2 * xo + 1) * a.W - a.Wo, ny = (2 * yo + 1) * a.H - a.Ho;
            const int ix = floor_div(nx, 2 * a.Wo), iy = floor_div(ny, 2 * a.Ho);
            const float fx = (float)(nx - ix * 2 * a.Wo) / (float)(2 * a.Wo), fy = (float)(ny - iy * 2 * a.Ho) / (float)(2 * a.Ho);
            const float4 v00 = fyn_fetch(a.in[0], n, t, P + ix, P + iy), v10 = fyn_fetch(a.in[0], n, t, P + ix + 1, P + iy);
            const float4 v01 = fyn_fetch(a.in[0], n, t, P + ix, P + iy + 1), v11 = fyn_fetch(a.in[0], n, t, P + ix + 1, P + iy + 1);
            r = fyn_act4(lerp4(lerp4(v00, v10, fx), lerp4(v01, v11, fx), fy), a.act);
        }
    } else if (a.mode == G_ARITH) {
        const float4 p = fyn_act4(fyn_fetch(a.in[0], n, t, P + xo, P + yo), a.act);
        const float4 q = a.singleton ? make_float4(a.operand, a.operand, a.operand, a.operand)
                                     : fyn_act4(fyn_fetch(a.in[1], n, t, a.in[1].P + xo, a.in[1].P + yo), a.act);
        if (a.op == FYN_ARITH_ADD) r = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
        else if (a.op == FYN_ARITH_SUB) r = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
        else if (a.op == FYN_ARITH_MUL) r = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
        else r = make_float4(p.x / q.x, p.y / q.y, p.z / q.z, p.w / q.w);
    } else if (a.mode == G_CONCAT) {
        float o[4];
        int lastK = -1, lastPlane = -1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const int c = 4 * t + l;
            o[l] = 0.f;
            if (c >= a.chOff[a.nin]) continue;
            int k = 0;
            while (c >= a.chOff[k + 1]) k++;
            const int ci = c - a.chOff[k];
            if (k != lastK || (ci >> 2) != lastPlane) {
                // (dynamic indexing of the views would put them in local memory: select with a uniform switch)
                switch (k) {
                    case 0: v = fyn_fetch(a.in[0], n, ci >> 2, a.in[0].P + xo, a.in[0].P + yo); break;
                    case 1: v = fyn_fetch(a.in[1], n, ci >> 2, a.in[1].P + xo, a.in[1].P + yo); break;
                    case 2: v = fyn_fetch(a.in[2], n, ci >> 2, a.in[2].P + xo, a.in[2].P + yo); break;
                    case 3: v = fyn_fetch(a.in[3], n, ci >> 2, a.in[3].P + xo, a.in[3].P + yo); break;
                    case 4: v = fyn_fetch(a.in[4], n, ci >> 2, a.in[4].P + xo, a.in[4].P + yo); break;
                    case 5: v = fyn_fetch(a.in[5], n, ci >> 2, a.in[5].P + xo, a.in[5].P + yo); break;
                    case 6: v = fyn_fetch(a.in[6], n, ci >> 2, a.in[6].P + xo, a.in[6].P + yo); break;
                    default: v = fyn_fetch(a.in[7], n, ci >> 2, a.in[7].P + xo, a.in[7].P + yo); break;
                }
                v = fyn_act4(v, a.act);
                lastK = k;
                lastPlane = ci >> 2;
            }
            o[l] = lane_of(v, ci & 3);
        }
        r = make_float4(o[0], o[1], o[2], o[3]);
    } else if (a.mode == G_RGB2BGR) {
        const float4 v = fyn_fetch(a.in[0], n, t, P + xo, P + yo);
        r = make_float4(v.z, v.y, v.x, v.w);   // shaders/rgb2bgr.frag: val.bgra
    } else {
        r = fyn_act4(fyn_fetch(a.in[0], n, t, P + xo, P + yo), a.act);
    }
    fyn_store_texel(a.out, n, t, a.outP + xo, a.outP + yo, r);
}

// ---------------------------------------------------------------------------------------------
// fp16 fast path of the 1:1 modes (add / sub / singleton arithmetic, concatenation of 4-aligned inputs, RGB<->BGR, layout
// conversion): a block owns a chunk of ONE 4-channel plane, so tile origins are block-uniform and a texel costs one
// multiplication instead of the five divisions of k_gather; U 8-byte accesses per tensor in flight per thread.
// Arithmetic as in k_gather (fp32, one rounding to fp16 at the store), so both kernels give the same bits.
// ---------------------------------------------------------------------------------------------
enum { PL_COPY = 0, PL_ADD, PL_SUB, PL_MUL, PL_DIV, PL_BGR };

struct PlaneArgs {
    TView in0, in1, out;
    int tiles, batch, outTile0, outP;   // block (n, t) writes output tile outTile0 + t
    int mode, two;
    float operand;
    ActParams act;
};

__device__ __forceinline__ long long plane_origin(const TView &v, unsigned n, unsigned t, int P) {
    long long b = (long long)n * v.imageElems;
    unsigned x0 = P, y0 = P;
    if (v.deep) {
        x0 += (t % v.tx) * v.tileW;
        y0 += (t / v.tx) * v.tileH;
    } else {
        b += (long long)t * v.planeElems;
    }
    return b + ((long long)y0 * v.texW + x0) * 4;
}

__device__ __forceinline__ float4 h4_to_f4(uint2 raw) {
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    return make_float4(f0.x, f0.y, f1.x, f1.y);
}

template <int U>
__global__ void __launch_bounds__(256) k_plane_h4(const __grid_constant__ PlaneArgs a, unsigned W, unsigned HW, unsigned chunks, unsigned magic) {
    unsigned bid = blockIdx.x;
    const unsigned chunk = bid % chunks;
    bid /= chunks;
    const unsigned t = bid % (unsigned)a.tiles, n = bid / (unsigned)a.tiles;
    const __half *s0 = reinterpret_cast<const __half *>(a.in0.ptr) + plane_origin(a.in0, n, t, a.in0.P);
    const __half *s1 = a.two ? reinterpret_cast<const __half *>(a.in1.ptr) + plane_origin(a.in1, n, t, a.in1.P) : nullptr;
    __half *dst = reinterpret_cast<__half *>(a.out.ptr) + plane_origin(a.out, n, a.outTile0 + t, a.outP);
    const unsigned p0 = chunk * (256u * U) + threadIdx.x;
    uint2 r0[U], r1[U];
    unsigned oo[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const unsigned p = p0 + u * 256u;
        oo[u] = ~0u;
        if (p < HW) {
            const unsigned y = magic ? __umulhi(p, magic) : p / W, x = p - y * W;
            oo[u] = (y * a.out.texW + x) * 4;
            r0[u] = __ldg(reinterpret_cast<const uint2 *>(s0 + (y * a.in0.texW + x) * 4));
            if (a.two) r1[u] = __ldg(reinterpret_cast<const uint2 *>(s1 + (y * a.in1.texW + x) * 4));
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (oo[u] == ~0u) continue;
        float4 r;
        if (a.mode == PL_BGR) {
            const float4 v = h4_to_f4(r0[u]);
            r = make_float4(v.z, v.y, v.x, v.w);
        } else {
            const float4 p = fyn_act4(h4_to_f4(r0[u]), a.act);
            if (a.mode == PL_COPY) {
                r = p;
            } else {
                const float4 q = a.two ? fyn_act4(h4_to_f4(r1[u]), a.act) : make_float4(a.operand, a.operand, a.operand, a.operand);
                if (a.mode == PL_ADD) r = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
                else if (a.mode == PL_SUB) r = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
                else if (a.mode == PL_MUL) r = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
                else r = make_float4(p.x / q.x, p.y / q.y, p.z / q.z, p.w / q.w);
            }
        }
        const __half2 h0 = __floats2half2_rn(r.x, r.y), h1 = __floats2half2_rn(r.z, r.w);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned *>(&h0);
        o.y = *reinterpret_cast<const unsigned *>(&h1);
        *reinterpret_cast<uint2 *>(dst + oo[u]) = o;
    }
}

// Two-tensor add / sub on fp16 tensors of IDENTICAL geometry (same padding, same tiling): the whole texture is one flat array,
// 16 bytes per access and four accesses per tensor in flight. Border / unused texels hold zeros on both sides and
// act(0) +- act(0) = 0 keeps them zero (the launcher refuses a clip range that excludes 0).
__global__ void __launch_bounds__(256) k_addsub_flat(const uint4 *__restrict__ s0, const uint4 *__restrict__ s1, uint4 *__restrict__ dst,
                                                      unsigned long long units, int sub, ActParams act) {
    for (unsigned long long base = (unsigned long long)blockIdx.x * 1024u + threadIdx.x; base < units; base += (unsigned long long)gridDim.x * 1024u) {
        uint4 r0[4], r1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long i = base + u * 256u;
            if (i < units) {
                r0[u] = __ldg(s0 + i);
                r1[u] = __ldg(s1 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long i = base + u * 256u;
            if (i >= units) break;
            const float4 p0 = fyn_act4(h4_to_f4(make_uint2(r0[u].x, r0[u].y)), act), p1 = fyn_act4(h4_to_f4(make_uint2(r0[u].z, r0[u].w)), act);
            float4 q0 = fyn_act4(h4_to_f4(make_uint2(r1[u].x, r1[u].y)), act), q1 = fyn_act4(h4_to_f4(make_uint2(r1[u].z, r1[u].w)), act);
            if (sub) {
                q0 = make_float4(-q0.x, -q0.y, -q0.z, -q0.w);
                q1 = make_float4(-q1.x, -q1.y, -q1.z, -q1.w);
            }
            const __half2 h0 = __floats2half2_rn(p0.x + q0.x, p0.y + q0.y), h1 = __floats2half2_rn(p0.z + q0.z, p0.w + q0.w);
            const __half2 h2 = __floats2half2_rn(p1.x + q1.x, p1.y + q1.y), h3 = __floats2half2_rn(p1.z + q1.z, p1.w + q1.w);
            uint4 o;
            o.x = *reinterpret_cast<const unsigned *>(&h0);
            o.y = *reinterpret_cast<const unsigned *>(&h1);
            o.z = *reinterpret_cast<const unsigned *>(&h2);
            o.w = *reinterpret_cast<const unsigned *>(&h3);
            dst[i] = o;
        }
    }
}

bool same_geometry(const TView &a, const TView &b) {
    return a.dtype == b.dtype && a.packing == b.packing && a.deep == b.deep && a.texW == b.texW && a.texH == b.texH && a.P == b.P && a.tx == b.tx &&
           a.tileW == b.tileW && a.tileH == b.tileH && a.planeElems == b.planeElems && a.imageElems == b.imageElems;
}

bool plane_view_ok(const TView &v) { return v.dtype == FYN_F16 && v.packing == 4 && (long long)v.texW * v.texH < (1 << 28); }

// true if the fast path took the launch (FYN_GATHER_GENERIC=1 forces the generic kernel: tests compare the two)
bool launch_plane(fyn_ctx *ctx, PlaneArgs &a, int W, int H, void *stream, int *rc) {
    const char *env = getenv("FYN_GATHER_GENERIC");   // per launch, so that a test can toggle it
    const bool generic = env && atoi(env) != 0;
    if (generic || !plane_view_ok(a.in0) || !plane_view_ok(a.out) || (a.two && !plane_view_ok(a.in1))) return false;
    const long long halves = (long long)a.batch * a.out.imageElems;
    const bool zeroStays = a.act.type != 3 || (a.act.lo <= 0.f && a.act.hi >= 0.f);
    if (a.two && (a.mode == PL_ADD || a.mode == PL_SUB) && a.outTile0 == 0 && zeroStays && same_geometry(a.in0, a.out) && same_geometry(a.in1, a.out) &&
        a.out.P == a.outP && (halves % 8) == 0 && (((uintptr_t)a.in0.ptr | (uintptr_t)a.in1.ptr | (uintptr_t)a.out.ptr) & 15) == 0) {
        const unsigned long long units = (unsigned long long)halves / 8;
        const unsigned long long want = (units + 1023) / 1024, cap = (unsigned long long)ctx->prop.multiProcessorCount * 8;
        k_addsub_flat<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4 *>(a.in0.ptr), reinterpret_cast<const uint4 *>(a.in1.ptr), reinterpret_cast<uint4 *>(a.out.ptr), units, a.mode == PL_SUB, a.act);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        *rc = FYN_OK;
        if (e != cudaSuccess) {
            fyn_set_error("gather: launch failed: %s", cudaGetErrorString(e));
            *rc = FYN_ERR_CUDA;
        }
        return true;
    }
    const unsigned HW = (unsigned)W * (unsigned)H;
    const int U = HW > 1024 ? 8 : (HW > 512 ? 4 : (HW > 256 ? 2 : 1));
    const unsigned chunks = (HW + 256u * U - 1) / (256u * U);
    const long long blocks = (long long)chunks * a.tiles * a.batch;
    if (blocks <= 0 || blocks > 0x7fffffffll) return false;
    const unsigned magic = ((unsigned long long)(HW + 256u * U) * (unsigned)W < (1ull << 32) && W > 1) ? (unsigned)((1ull << 32) / (unsigned)W) + 1u : 0u;
    cudaStream_t s = (cudaStream_t)stream;
    if (U == 8) k_plane_h4<8><<<(unsigned)blocks, 256, 0, s>>>(a, (unsigned)W, HW, chunks, magic);
    else if (U == 4) k_plane_h4<4><<<(unsigned)blocks, 256, 0, s>>>(a, (unsigned)W, HW, chunks, magic);
    else if (U == 2) k_plane_h4<2><<<(unsigned)blocks, 256, 0, s>>>(a, (unsigned)W, HW, chunks, magic);
    else k_plane_h4<1><<<(unsigned)blocks, 256, 0, s>>>(a, (unsigned)W, HW, chunks, magic);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    *rc = FYN_OK;
    if (e != cudaSuccess) {
        fyn_set_error("gather: launch failed: %s", cudaGetErrorString(e));
        *rc = FYN_ERR_CUDA;
    }
    return true;
}

int check_tensor(const char *who, const fyn_tensor *t, int w, int h, int c, int pad, int deep /* -1: any order */) {
    if (!t) FYN_FAIL(FYN_ERR_INVALID, "%s: tensor is NULL", who);
    const fyn_tensor_desc &d = t->desc;
    const bool order_ok = deep < 0 || ((d.order == FYN_ORDER_DEEP) == (deep != 0)) || c <= 4;
    if (d.width != w || d.height != h || d.channels != c || d.padding != pad || !order_ok)
        FYN_FAIL(FYN_ERR_INVALID, "%s: tensor mismatch: got %dx%dx%d pad %d order %d, need %dx%dx%d pad %d", who, d.width, d.height,
                 d.channels, d.padding, d.order, w, h, c, pad);
    return FYN_OK;
}

int launch(fyn_ctx *ctx, GatherArgs &a, void *stream) {
    const long long blocks = (long long)((a.Wo + 31) / 32) * ((a.Ho + 3) / 4) * a.tiles * a.batch;
    if (blocks <= 0 || blocks > 0x7fffffffll) FYN_FAIL(FYN_ERR_INVALID, "gather: grid of %lld blocks", blocks);
    k_gather<<<(unsigned)blocks, dim3(32, 4), 0, (cudaStream_t)stream>>>(a);
    FYN_CHECK_LAUNCH(ctx);
    return FYN_OK;
}

fyn_op *new_op(fyn_ctx *ctx, int kind) {
    fyn_op *op = new fyn_op();
    op->ctx = ctx;
    op->kind = kind;
    return op;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// scaling
// ---------------------------------------------------------------------------------------------
int fyn_scale_out_size(const fyn_scale_desc *d, int *width, int *height) {
    if (!d || !width || !height) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    if (d->upsample_x < 1 || d->upsample_y < 1 || d->downsample_x < 1 || d->downsample_y < 1) FYN_FAIL(FYN_ERR_INVALID, "scale: factors must be >= 1");
    // gpu/scalelayer.cpp:44-47
    *width = (int)(((float)d->upsample_x / (float)d->downsample_x) * (float)d->width);
    *height = (int)(((float)d->upsample_y / (float)d->downsample_y) * (float)d->height);
    return FYN_OK;
}

int fyn_scale_create(fyn_ctx *ctx, const fyn_scale_desc *d, fyn_op **out) {
    if (!ctx || !d || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (d->width <= 0 || d->height <= 0 || d->channels <= 0) FYN_FAIL(FYN_ERR_INVALID, "scale: bad shape");
    int w, h;
    int rc = fyn_scale_out_size(d, &w, &h);
    if (rc) return rc;
    if (w < 1 || h < 1) FYN_FAIL(FYN_ERR_INVALID, "scale: empty output");
    fyn_op *op = new_op(ctx, FYN_OP_SCALE);
    op->scale = *d;
    op->Wo = w;
    op->Ho = h;
    *out = op;
    return FYN_OK;
}

int fyn_scale_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_SCALE) FYN_FAIL(FYN_ERR_INVALID, "not a scale op");
    const fyn_scale_desc &d = op->scale;
    const int deep = (d.flags & FYN_FLAG_DEEP) ? 1 : 0;
    int rc = check_tensor("scale input", in, d.width, d.height, d.channels, d.in_padding, deep);
    if (rc) return rc;
    rc = check_tensor("scale output", out, op->Wo, op->Ho, d.channels, d.out_padding, deep);
    if (rc) return rc;
    if (in->desc.batch != out->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "scale: batch mismatch");
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    GatherArgs a{};
    a.mode = G_SCALE;
    a.in[0] = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.W = d.width;
    a.H = d.height;
    a.linear = d.linear && !(deep && (d.width == 1 || d.height == 1));   // gpu/deep/deepscalelayer.cpp:34
    a.tiles = (d.channels + 3) / 4;
    a.batch = in->desc.batch;
    a.outP = d.out_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    return launch(op->ctx, a, stream);
}

// ---------------------------------------------------------------------------------------------
// add / sub / singleton arithmetic
// ---------------------------------------------------------------------------------------------
int fyn_arith_create(fyn_ctx *ctx, const fyn_arith_desc *d, fyn_op **out) {
    if (!ctx || !d || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (d->width <= 0 || d->height <= 0 || d->channels <= 0) FYN_FAIL(FYN_ERR_INVALID, "arith: bad shape");
    if (d->op < FYN_ARITH_ADD || d->op > FYN_ARITH_DIV) FYN_FAIL(FYN_ERR_INVALID, "arith: unknown operation %d", d->op);
    if (!d->singleton && d->op > FYN_ARITH_SUB) FYN_FAIL(FYN_ERR_UNSUPPORTED, "arith: the two-tensor layer supports ADD and SUB only (gpu/addsublayer.cpp)");
    fyn_op *op = new_op(ctx, FYN_OP_ARITH);
    op->arith = *d;
    *out = op;
    return FYN_OK;
}

int fyn_arith_run(fyn_op *op, const fyn_tensor *in1, const fyn_tensor *in2, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_ARITH) FYN_FAIL(FYN_ERR_INVALID, "not an arithmetic op");
    const fyn_arith_desc &d = op->arith;
    const int deep = (d.flags & FYN_FLAG_DEEP) ? 1 : 0;
    int rc = check_tensor("arith input 1", in1, d.width, d.height, d.channels, d.in_padding, deep);
    if (rc) return rc;
    if (!d.singleton) {
        rc = check_tensor("arith input 2", in2, d.width, d.height, d.channels, d.in_padding, deep);
        if (rc) return rc;
        if (in2->desc.batch != in1->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "arith: batch mismatch");
    } else if (in2) {
        FYN_FAIL(FYN_ERR_INVALID, "arith: a singleton layer takes one input tensor");
    }
    rc = check_tensor("arith output", out, d.width, d.height, d.channels, d.out_padding, deep);
    if (rc) return rc;
    if (in1->desc.batch != out->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "arith: batch mismatch");
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    GatherArgs a{};
    a.mode = G_ARITH;
    a.in[0] = fyn_make_view(in1);
    if (in2) a.in[1] = fyn_make_view(in2);
    a.out = fyn_make_view(out);
    a.Wo = d.width;
    a.Ho = d.height;
    a.op = d.op;
    a.singleton = d.singleton;
    a.operand = d.operand;
    a.tiles = (d.channels + 3) / 4;
    a.batch = in1->desc.batch;
    a.outP = d.out_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    {
        PlaneArgs pa{};
        pa.in0 = a.in[0];
        pa.in1 = a.in[1];
        pa.out = a.out;
        pa.tiles = a.tiles;
        pa.batch = a.batch;
        pa.outP = a.outP;
        pa.mode = PL_ADD + (d.op - FYN_ARITH_ADD);
        pa.two = in2 ? 1 : 0;
        pa.operand = d.operand;
        pa.act = a.act;
        if (launch_plane(op->ctx, pa, d.width, d.height, stream, &rc)) return rc;
    }
    return launch(op->ctx, a, stream);
}

// ---------------------------------------------------------------------------------------------
// concatenation
// ---------------------------------------------------------------------------------------------
int fyn_concat_create(fyn_ctx *ctx, const fyn_concat_desc *d, fyn_op **out) {
    if (!ctx || !d || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (d->width <= 0 || d->height <= 0) FYN_FAIL(FYN_ERR_INVALID, "concat: bad shape");
    if (d->num_inputs < 1 || d->num_inputs > FYN_CONCAT_MAX_INPUTS) FYN_FAIL(FYN_ERR_INVALID, "concat: 1..%d inputs", FYN_CONCAT_MAX_INPUTS);
    for (int i = 0; i < d->num_inputs; i++)
        if (d->channels[i] < 1) FYN_FAIL(FYN_ERR_INVALID, "concat: input %d has %d channels", i, d->channels[i]);
    fyn_op *op = new_op(ctx, FYN_OP_CONCAT);
    op->concat = *d;
    *out = op;
    return FYN_OK;
}

int fyn_concat_run(fyn_op *op, const fyn_tensor *const *inputs, int num_inputs, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_CONCAT) FYN_FAIL(FYN_ERR_INVALID, "not a concat op");
    const fyn_concat_desc &d = op->concat;
    if (!inputs || num_inputs != d.num_inputs) FYN_FAIL(FYN_ERR_INVALID, "concat: %d inputs given, %d expected", num_inputs, d.num_inputs);
    const int deep = (d.flags & FYN_FLAG_DEEP) ? 1 : 0;
    GatherArgs a{};
    a.mode = G_CONCAT;
    a.nin = d.num_inputs;
    int total = 0;
    for (int i = 0; i < d.num_inputs; i++) {
        int rc = check_tensor("concat input", inputs[i], d.width, d.height, d.channels[i], d.in_padding, deep);
        if (rc) return rc;
        a.in[i] = fyn_make_view(inputs[i]);
        a.chOff[i] = total;
        total += d.channels[i];
    }
    a.chOff[d.num_inputs] = total;
    int rc = check_tensor("concat output", out, d.width, d.height, total, d.out_padding, deep);
    if (rc) return rc;
    for (int i = 0; i < d.num_inputs; i++)
        if (inputs[i]->desc.batch != out->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "concat: batch mismatch");
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    a.out = fyn_make_view(out);
    a.Wo = d.width;
    a.Ho = d.height;
    a.tiles = (total + 3) / 4;
    a.batch = out->desc.batch;
    a.outP = d.out_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    // inputs of whole texels: every output plane is one input plane -> one plane-copy launch per input
    bool aligned = plane_view_ok(a.out);
    for (int i = 0; i < d.num_inputs; i++) aligned = aligned && (d.channels[i] % 4) == 0 && plane_view_ok(a.in[i]);
    if (aligned) {
        for (int i = 0; i < d.num_inputs; i++) {
            PlaneArgs pa{};
            pa.in0 = a.in[i];
            pa.out = a.out;
            pa.tiles = d.channels[i] / 4;
            pa.batch = a.batch;
            pa.outTile0 = a.chOff[i] / 4;
            pa.outP = a.outP;
            pa.mode = PL_COPY;
            pa.act = a.act;
            if (!launch_plane(op->ctx, pa, d.width, d.height, stream, &rc)) {
                if (i == 0) break;   // fast path disabled: the generic kernel writes everything
                FYN_FAIL(FYN_ERR_CUDA, "concat: plane copy refused after the first input");
            }
            if (rc) return rc;
            if (i == d.num_inputs - 1) return FYN_OK;
        }
    }
    return launch(op->ctx, a, stream);
}

// ---------------------------------------------------------------------------------------------
// RGB <-> BGR, shallow <-> deep
// ---------------------------------------------------------------------------------------------
static int unary_create(fyn_ctx *ctx, const fyn_unary_desc *d, int kind, fyn_op **out) {
    if (!ctx || !d || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (d->width <= 0 || d->height <= 0 || d->channels <= 0) FYN_FAIL(FYN_ERR_INVALID, "bad shape");
    fyn_op *op = new_op(ctx, kind);
    op->unary = *d;
    *out = op;
    return FYN_OK;
}

static int unary_run(fyn_op *op, int mode, int inDeep, int outDeep, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    const fyn_unary_desc &d = op->unary;
    int rc = check_tensor("input", in, d.width, d.height, d.channels, d.in_padding, inDeep);
    if (rc) return rc;
    rc = check_tensor("output", out, d.width, d.height, d.channels, d.out_padding, outDeep);
    if (rc) return rc;
    if (in->desc.batch != out->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "batch mismatch");
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    GatherArgs a{};
    a.mode = mode;
    a.in[0] = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.Wo = d.width;
    a.Ho = d.height;
    a.tiles = (d.channels + 3) / 4;
    a.batch = in->desc.batch;
    a.outP = d.out_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    {
        PlaneArgs pa{};
        pa.in0 = a.in[0];
        pa.out = a.out;
        pa.tiles = a.tiles;
        pa.batch = a.batch;
        pa.outP = a.outP;
        pa.mode = mode == G_RGB2BGR ? PL_BGR : PL_COPY;
        pa.act = a.act;
        if (launch_plane(op->ctx, pa, d.width, d.height, stream, &rc)) return rc;
    }
    return launch(op->ctx, a, stream);
}

int fyn_rgb2bgr_create(fyn_ctx *ctx, const fyn_unary_desc *d, fyn_op **out) { return unary_create(ctx, d, FYN_OP_RGB2BGR, out); }

int fyn_rgb2bgr_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_RGB2BGR) FYN_FAIL(FYN_ERR_INVALID, "not an rgb2bgr op");
    const int deep = (op->unary.flags & FYN_FLAG_DEEP) ? 1 : 0;
    return unary_run(op, G_RGB2BGR, deep, deep, in, out, stream);
}

int fyn_relayout_create(fyn_ctx *ctx, const fyn_unary_desc *d, fyn_op **out) { return unary_create(ctx, d, FYN_OP_RELAYOUT, out); }

int fyn_relayout_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_RELAYOUT) FYN_FAIL(FYN_ERR_INVALID, "not a relayout op");
    return unary_run(op, G_RELAYOUT, -1, -1, in, out, stream);
}

}  // extern "C"
