// fyn_transconv.cu -- stride-2 transpose convolution (2x2 / 3x3, shallow and deep-tiled tensors).
// The reference renders four "strata" (output parity classes) through a stencil buffer; here one thread computes one
// output texel (4 output channels) and picks the taps of its parity class directly.  The deep-tiled layers
// (gpu/deep/deeptransconvlayerbase.cpp, shaders/deep/deeptransconv{2x2,3x3}_stride2.*) align their taps differently -- even
// output texels take tap 0 on input i and tap 2 on input i - 1, odd ones tap 1 on input i; 2x2: tap = parity on input i -- and
// read zero outside the image whatever the padding.
#include <vector>

#include "fyn_internal.h"

namespace {

struct TcvArgs {
    TView in, out;
    const float4 *w;       // [nOut][nIn][K*K][4 ci] float4 over co
    const float4 *bias, *scale;
    int K, nIn, nOut, Wo, Ho, batch, outP, next2, deepTaps;
    ActParams act;
};

__global__ void __launch_bounds__(128) k_transconv(const TcvArgs a) {
    unsigned bid = blockIdx.x;
    const int xBlocks = (a.Wo + 31) / 32, yBlocks = (a.Ho + 3) / 4;
    const int xb = bid % xBlocks;
    bid /= xBlocks;
    const int yb = bid % yBlocks;
    bid /= yBlocks;
    const int t = bid % a.nOut;
    const int n = bid / a.nOut;
    const int xo = xb * 32 + threadIdx.x, yo = yb * 4 + threadIdx.y;
    if (xo >= a.Wo || yo >= a.Ho) return;
    const int i = xo >> 1, j = yo >> 1, ox = xo & 1, oy = yo & 1;
    // taps of this parity class: (kernel index, input offset) per axis
    int kxs[2], dxs[2], nx, kys[2], dys[2], ny;
    if (a.K == 3 && a.deepTaps) {
        if (ox) { nx = 1; kxs[0] = 1; dxs[0] = 0; } else { nx = 2; kxs[0] = 0; dxs[0] = 0; kxs[1] = 2; dxs[1] = -1; }
        if (oy) { ny = 1; kys[0] = 1; dys[0] = 0; } else { ny = 2; kys[0] = 0; dys[0] = 0; kys[1] = 2; dys[1] = -1; }
    } else if (a.K == 3) {
        if (ox) { nx = 2; kxs[0] = 0; dxs[0] = 0; kxs[1] = 2; dxs[1] = 1; } else { nx = 1; kxs[0] = 1; dxs[0] = 0; }
        if (oy) { ny = 2; kys[0] = 0; dys[0] = 0; kys[1] = 2; dys[1] = 1; } else { ny = 1; kys[0] = 1; dys[0] = 0; }
    } else {
        nx = ny = 1;
        kxs[0] = ox;
        kys[0] = oy;
        // convtrans2x2_stride2.frag: STEP 2 samples tc - hstep (column i), STEP 3 tc + vstep (row j + 1), STEP 4 tc + step
        dxs[0] = (a.next2 && ox && oy) ? 1 : 0;
        dys[0] = (a.next2 && oy) ? 1 : 0;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int P = a.in.P, KK = a.K * a.K;
    for (int ty = 0; ty < ny; ty++)
        for (int tx = 0; tx < nx; tx++) {
            const int tap = kys[ty] * a.K + kxs[tx];
            // deep layers: zero outside the image (clampedTexture, deeptransconv3x3_stride2.frag:21-25), then the activation
            const bool inside = !a.deepTaps || (i + dxs[tx] >= 0 && j + dys[ty] >= 0);
            for (int p = 0; p < a.nIn; p++) {
                const float4 v = fyn_act4(inside ? fyn_fetch(a.in, n, p, P + i + dxs[tx], P + j + dys[ty]) : make_float4(0.f, 0.f, 0.f, 0.f), a.act);
                const float4 *w = a.w + (((size_t)t * a.nIn + p) * KK + tap) * 4;
                const float4 w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
                acc.x += v.x * w0.x + v.y * w1.x + v.z * w2.x + v.w * w3.x;
                acc.y += v.x * w0.y + v.y * w1.y + v.z * w2.y + v.w * w3.y;
                acc.z += v.x * w0.z + v.y * w1.z + v.z * w2.z + v.w * w3.z;
                acc.w += v.x * w0.w + v.y * w1.w + v.z * w2.w + v.w * w3.w;
            }
        }
    const float4 s = __ldg(a.scale + t), b = __ldg(a.bias + t);
    fyn_store_texel(a.out, n, t, a.outP + xo, a.outP + yo, make_float4(acc.x * s.x + b.x, acc.y * s.y + b.y, acc.z * s.z + b.z, acc.w * s.w + b.w));
}

int validate(const fyn_transconv_desc *d) {
    if (d->width <= 0 || d->height <= 0 || d->in_channels <= 0 || d->out_channels <= 0) FYN_FAIL(FYN_ERR_INVALID, "transconv: bad shape");
    if (d->kernel != 2 && d->kernel != 3) FYN_FAIL(FYN_ERR_UNSUPPORTED, "transconv: only 2x2 and 3x3 kernels (stride 2) are supported");
    if (d->in_padding < 0 || d->out_padding < 0) FYN_FAIL(FYN_ERR_INVALID, "transconv: negative padding");
    if (d->flags & FYN_FLAG_RESIDUAL_INPUT) FYN_FAIL(FYN_ERR_UNSUPPORTED, "Transpose convolutions do not support residuals as of now");   // (deeptransconvlayer3x3.cpp:44-46)
    return FYN_OK;
}

}  // namespace

extern "C" {

int fyn_transconv2d_load_weights(fyn_op *op, const float *wb) {
    if (!op || op->kind != FYN_OP_TRANSCONV || !wb) FYN_FAIL(FYN_ERR_INVALID, "bad transconv op / weights");
    const fyn_transconv_desc &d = op->tconv;
    const int Ci = d.in_channels, Co = d.out_channels, K = d.kernel, nIn = (Ci + 3) / 4, nOut = (Co + 3) / 4;
    const size_t wFloats = (size_t)nOut * nIn * K * K * 16, setFloats = wFloats + (size_t)nOut * 8;
    // deep layers keep a second parameter set for fp16 storage: weights fp16-truncated (toFP16UI, deeptransconvlayerbase.cpp:177-186),
    // bias / scale fp16-rounded (RGBA16F bias texture, :228-232); selected at run time by the tensors' data type
    const int sets = (d.flags & FYN_FLAG_DEEP) ? 2 : 1;
    std::vector<float> h(setFloats * sets, 0.f);
    const float *src = wb + Co;   // W[Co][K][K][Ci]
    const float *bn = src + (size_t)Co * K * K * Ci;
    for (int set = 0; set < sets; set++) {
        float *w = h.data() + set * setFloats;
        for (int o = 0; o < Co; o++)
            for (int tap = 0; tap < K * K; tap++)
                for (int c = 0; c < Ci; c++) {
                    const float v = src[((size_t)o * K * K + tap) * Ci + c];
                    w[((((size_t)(o / 4) * nIn + c / 4) * K * K + tap) * 4 + (c & 3)) * 4 + (o & 3)] = set ? fyn_half_trunc_host(v) : v;
                }
        float *bias = w + wFloats, *scale = bias + (size_t)nOut * 4;
        for (int o = 0; o < Co; o++) {
            float b = wb[o], s = 1.f;
            if (d.flags & FYN_FLAG_POST_BATCHNORM) {
                s = bn[o];
                b = b * s + bn[Co + o];   // transconvweightarray3x3xNxM.cpp:168-179
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

int fyn_transconv2d_create(fyn_ctx *ctx, const fyn_transconv_desc *desc, const float *wb, fyn_op **out) {
    if (!ctx || !desc || !wb || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = validate(desc);
    if (rc) return rc;
    fyn_op *op = new fyn_op();
    op->ctx = ctx;
    op->kind = FYN_OP_TRANSCONV;
    op->tconv = *desc;
    op->Wo = 2 * desc->width;    // transconvlayerbase_vanilla.cpp:60-62
    op->Ho = 2 * desc->height;
    rc = fyn_transconv2d_load_weights(op, wb);
    if (rc) {
        fyn_op_destroy(op);
        return rc;
    }
    *out = op;
    return FYN_OK;
}

int fyn_transconv2d_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_TRANSCONV) FYN_FAIL(FYN_ERR_INVALID, "not a transconv op");
    if (!in || !out) FYN_FAIL(FYN_ERR_INVALID, "transconv: tensor is NULL");
    const fyn_transconv_desc &d = op->tconv;
    const fyn_tensor_desc &i = in->desc, &o = out->desc;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    const bool orderOk = ((i.order == FYN_ORDER_DEEP) == deep || d.in_channels <= 4) && ((o.order == FYN_ORDER_DEEP) == deep || d.out_channels <= 4);
    if (i.width != d.width || i.height != d.height || i.channels != d.in_channels || i.padding != d.in_padding || o.width != op->Wo ||
        o.height != op->Ho || o.channels != d.out_channels || o.padding != d.out_padding || !orderOk || i.batch != o.batch || out->geom.packing != 4)
        FYN_FAIL(FYN_ERR_INVALID, "transconv: tensor mismatch: in %dx%dx%d pad %d, out %dx%dx%d pad %d; need %dx%dx%d pad %d -> %dx%dx%d pad %d", i.width,
                 i.height, i.channels, i.padding, o.width, o.height, o.channels, o.padding, d.width, d.height, d.in_channels, d.in_padding, op->Wo,
                 op->Ho, d.out_channels, d.out_padding);
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    TcvArgs a{};
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.K = d.kernel;
    a.nIn = (d.in_channels + 3) / 4;
    a.nOut = (d.out_channels + 3) / 4;
    const size_t setFloats = (size_t)a.nOut * a.nIn * d.kernel * d.kernel * 16 + (size_t)a.nOut * 8;
    a.w = reinterpret_cast<const float4 *>(op->d_w + ((deep && in->desc.dtype == FYN_F16) ? setFloats : 0));
    a.deepTaps = deep ? 1 : 0;
    a.bias = a.w + (size_t)a.nOut * a.nIn * d.kernel * d.kernel * 4;
    a.scale = a.bias + a.nOut;
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.batch = i.batch;
    a.outP = d.out_padding;
    a.next2 = (!deep && (d.quirks & FYN_QUIRK_TRANS2X2_NEXT)) ? 1 : 0;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    const long long blocks = (long long)((a.Wo + 31) / 32) * ((a.Ho + 3) / 4) * a.nOut * a.batch;
    if (blocks > 0x7fffffffll) FYN_FAIL(FYN_ERR_INVALID, "transconv: grid of %lld blocks", blocks);
    k_transconv<<<(unsigned)blocks, dim3(32, 4), 0, (cudaStream_t)stream>>>(a);
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}

}  // extern "C"
