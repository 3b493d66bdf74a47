// fyn_dwconv.cu -- depthwise 3x3 convolution, shallow and deep-tiled, channel multiplier and residual input.
// Bandwidth-bound: one thread per output texel (4 channels), nine clamped texel fetches, per-channel weights as float4.
#include <vector>

#include "fyn_internal.h"

namespace {

struct DwArgs {
    TView in, out, res;
    const float4 *w;        // [output tiles][9]
    const float4 *bias;     // [output tiles]
    const float4 *scale;    // [output tiles]
    int ds, dil, Wo, Ho, tiles, inTiles, batch, outP, resP;
    int hasRes, reluRes, bnRes;
    ActParams act;
};

__global__ void __launch_bounds__(128) k_dwconv3x3(const DwArgs a) {
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
    const int cx = a.in.P + a.ds * xo, cy = a.in.P + a.ds * yo;
    const int ti = t % a.inTiles;      // channel multiplier: output tile t = multiplier (t / inTiles) of input tile ti
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const float4 v = fyn_act4(fyn_fetch(a.in, n, ti, cx + (kx - 1) * a.dil, cy + (ky - 1) * a.dil), a.act);
            const float4 w = __ldg(a.w + t * 9 + ky * 3 + kx);
            acc.x += v.x * w.x;
            acc.y += v.y * w.y;
            acc.z += v.z * w.z;
            acc.w += v.w * w.w;
        }
    const float4 s = __ldg(a.scale + t), b = __ldg(a.bias + t);
    float4 r = make_float4(acc.x * s.x + b.x, acc.y * s.y + b.y, acc.z * s.z + b.z, acc.w * s.w + b.w);
    if (a.hasRes) {
        float4 q = fyn_fetch(a.res, n, t, a.resP + xo, a.resP + yo);
        if (a.reluRes) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
        if (a.bnRes) q = make_float4(q.x * s.x, q.y * s.y, q.z * s.z, q.w * s.w);
        r = make_float4(r.x + q.x, r.y + q.y, r.z + q.z, r.w + q.w);
    }
    fyn_store_texel(a.out, n, t, a.outP + xo, a.outP + yo, r);
}

int validate(const fyn_dwconv_desc *d) {
    if (d->width <= 0 || d->height <= 0 || d->channels <= 0) FYN_FAIL(FYN_ERR_INVALID, "dwconv: bad shape");
    if (d->downsample < 1 || d->dilation < 1) FYN_FAIL(FYN_ERR_INVALID, "dwconv: stride and dilation must be >= 1");
    if (d->in_padding < 0 || d->out_padding < 0) FYN_FAIL(FYN_ERR_INVALID, "dwconv: negative padding");
    if (!(d->flags & FYN_FLAG_DEEP) && d->dilation != 1) FYN_FAIL(FYN_ERR_UNSUPPORTED, "dwconv: the shallow depthwise layer has no dilation (conv_dw_3x3.frag)");
    if (d->multiplier < 0 || d->res_padding < 0) FYN_FAIL(FYN_ERR_INVALID, "dwconv: bad channel multiplier / residual padding");
    if (d->multiplier > 1 && (!(d->flags & FYN_FLAG_DEEP) || (d->channels & 3)))
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "Channel multipliers > 1 are only supported on deep layers with input channels being a multiple of 4");
    if ((d->flags & (FYN_FLAG_RELU_ON_RESIDUAL | FYN_FLAG_BATCHNORM_ON_RESIDUAL)) && !(d->flags & FYN_FLAG_RESIDUAL_INPUT))
        FYN_FAIL(FYN_ERR_INVALID, "dwconv: residual modifiers without RESIDUAL_INPUT");
    if ((d->flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) && !(d->flags & FYN_FLAG_DEEP))
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "dwconv: the shallow depthwise layer has no batch-norm on its residual (conv_dw_3x3.frag:133-140)");
    if (d->width / d->downsample < 1 || d->height / d->downsample < 1) FYN_FAIL(FYN_ERR_INVALID, "dwconv: empty output");
    return FYN_OK;
}

}  // namespace

extern "C" {

// device parameter block: [fp32 weights][fp32 bias][fp32 scale] and, for deep layers, a second copy with fp16-truncated
// weights (gpu/floatconversion.cpp:44-58) and fp16-rounded bias / scale (RGBA16F bias texture, deepdwconvlayerbase.cpp:47-53)
int fyn_dwconv3x3_load_weights(fyn_op *op, const float *wb) {
    if (!op || op->kind != FYN_OP_DWCONV || !wb) FYN_FAIL(FYN_ERR_INVALID, "bad dwconv op / weights");
    const fyn_dwconv_desc &d = op->dw;
    const int C = d.channels, M = d.multiplier > 1 ? d.multiplier : 1, Co = C * M, tiles = (Co + 3) / 4;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    const size_t setFloats = (size_t)tiles * 4 * 11;
    std::vector<float> h(setFloats * (deep ? 2 : 1), 0.f);
    // shallow quirk: batch-norm data read at the start of the block (convlayer_dw_3x3_vanilla.cpp:66)
    const float *bn = (!deep && (d.quirks & FYN_QUIRK_DW_BN_OFFSET)) ? wb : wb + Co + (size_t)C * 9 * M;
    for (int set = 0; set < (deep ? 2 : 1); set++) {
        float *w = h.data() + set * setFloats, *bias = w + (size_t)tiles * 36, *scale = bias + (size_t)tiles * 4;
        for (int o = 0; o < Co; o++) {
            const int m = o / C, c = o - m * C;        // output channel m * C + c: input channel c, multiplier m
            for (int k = 0; k < 9; k++) {
                const float v = wb[Co + ((size_t)c * 9 + k) * M + m];
                w[((size_t)(o / 4) * 9 + k) * 4 + (o & 3)] = set ? fyn_half_trunc_host(v) : v;
            }
            float b = wb[o], s = 1.f;
            if (d.flags & FYN_FLAG_POST_BATCHNORM) {
                s = bn[o];
                b = b * s + bn[Co + o];
            }
            bias[o] = set ? fyn_half_round_host(b) : b;
            scale[o] = set ? fyn_half_round_host(s) : s;
        }
    }
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    if (!op->d_w) FYN_CUDA(cudaMalloc((void **)&op->d_w, h.size() * sizeof(float)));
    FYN_CUDA(cudaMemcpy(op->d_w, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return FYN_OK;
}

int fyn_dwconv3x3_create(fyn_ctx *ctx, const fyn_dwconv_desc *desc, const float *wb, fyn_op **out) {
    if (!ctx || !desc || !wb || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = validate(desc);
    if (rc) return rc;
    fyn_op *op = new fyn_op();
    op->ctx = ctx;
    op->kind = FYN_OP_DWCONV;
    op->dw = *desc;
    op->Wo = desc->width / desc->downsample;    // gpu/convlayerbase.cpp:47-48
    op->Ho = desc->height / desc->downsample;
    rc = fyn_dwconv3x3_load_weights(op, wb);
    if (rc) {
        fyn_op_destroy(op);
        return rc;
    }
    *out = op;
    return FYN_OK;
}

int fyn_dwconv3x3_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream) { return fyn_dwconv3x3_run_residual(op, in, nullptr, out, stream); }

int fyn_dwconv3x3_run_residual(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_DWCONV) FYN_FAIL(FYN_ERR_INVALID, "not a dwconv op");
    const fyn_dwconv_desc &d = op->dw;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    if (!in || !out) FYN_FAIL(FYN_ERR_INVALID, "dwconv: tensor is NULL");
    const int M = d.multiplier > 1 ? d.multiplier : 1, Co = d.channels * M;
    const fyn_tensor_desc &i = in->desc, &o = out->desc;
    const bool orderOk = Co <= 4 || (((i.order == FYN_ORDER_DEEP) == deep) && ((o.order == FYN_ORDER_DEEP) == deep));
    if (i.width != d.width || i.height != d.height || i.channels != d.channels || i.padding != d.in_padding || o.width != op->Wo ||
        o.height != op->Ho || o.channels != Co || o.padding != d.out_padding || !orderOk || i.batch != o.batch)
        FYN_FAIL(FYN_ERR_INVALID, "dwconv: tensor mismatch: in %dx%dx%d pad %d, out %dx%dx%d pad %d; need %dx%dx%d pad %d -> %dx%dx%d pad %d", i.width,
                 i.height, i.channels, i.padding, o.width, o.height, o.channels, o.padding, d.width, d.height, d.channels, d.in_padding, op->Wo,
                 op->Ho, Co, d.out_padding);
    const bool hasRes = (d.flags & FYN_FLAG_RESIDUAL_INPUT) != 0;
    if (hasRes) {
        if (!res) FYN_FAIL(FYN_ERR_INVALID, "dwconv: residual tensor is NULL");
        const fyn_tensor_desc &r = res->desc;
        if (r.width != op->Wo || r.height != op->Ho || r.channels != Co || r.padding != d.res_padding || r.batch != o.batch ||
            (Co > 4 && (r.order == FYN_ORDER_DEEP) != deep))
            FYN_FAIL(FYN_ERR_INVALID, "dwconv: residual tensor mismatch: got %dx%dx%d pad %d, need %dx%dx%d pad %d", r.width, r.height, r.channels, r.padding, op->Wo,
                     op->Ho, Co, d.res_padding);
    }
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    DwArgs a{};
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    const int tiles = (Co + 3) / 4;
    a.inTiles = (d.channels + 3) / 4;
    if (hasRes) a.res = fyn_make_view(res);
    a.hasRes = hasRes ? 1 : 0;
    a.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
    a.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
    a.resP = d.res_padding;
    // fp16 storage of a deep layer selects the reduced-precision parameter set (the reference decides at build time)
    const size_t set = (deep && in->desc.dtype == FYN_F16) ? (size_t)tiles * 44 : 0;
    a.w = reinterpret_cast<const float4 *>(op->d_w + set);
    a.bias = a.w + (size_t)tiles * 9;
    a.scale = a.bias + tiles;
    a.ds = d.downsample;
    a.dil = d.dilation;
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.tiles = tiles;
    a.batch = i.batch;
    a.outP = d.out_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    const long long blocks = (long long)((a.Wo + 31) / 32) * ((a.Ho + 3) / 4) * tiles * a.batch;
    if (blocks > 0x7fffffffll) FYN_FAIL(FYN_ERR_INVALID, "dwconv: grid of %lld blocks", blocks);
    k_dwconv3x3<<<(unsigned)blocks, dim3(32, 4), 0, (cudaStream_t)stream>>>(a);
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}

}  // extern "C"
