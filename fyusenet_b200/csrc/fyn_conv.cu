// fyn_conv.cu -- convolution op: validation, weight repacking, kernel-family dispatch.
// Replaces ConvLayerBase::loadWeightsAndBiases + ConvWeightArrayKxKxNxM
// (fyusenet/gpu/vanilla/convlayerbase_vanilla.cpp:253-262, fyusenet/gpu/convweightarrayKxKxNxM.cpp:151-234)
// and DeepConvLayerBase::loadWeightsAndBiases (fyusenet/gpu/deep/deepconvlayerbase.cpp:293-395).
#include <cmath>
#include <cstring>

#include "fyn_internal.h"

extern "C" int fyn_conv2d_output_size(const fyn_conv_desc *d, int *ow, int *oh) {
    if (!d || !ow || !oh) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    if (d->downsample < 1) FYN_FAIL(FYN_ERR_INVALID, "downsample must be >= 1");
    if (d->fractional) {
        // gpu/vanilla/fractionalconvlayerNxN_vanilla.cpp:46-49
        if (!(d->source_step > 0.f)) FYN_FAIL(FYN_ERR_INVALID, "source_step must be > 0");
        *ow = (int)((float)d->width / (d->source_step * (float)d->downsample));
        *oh = (int)((float)d->height / (d->source_step * (float)d->downsample));
    } else {
        // gpu/convlayerbase.cpp:47-48
        *ow = d->width / d->downsample;
        *oh = d->height / d->downsample;
    }
    return FYN_OK;
}

static int validate(const fyn_conv_desc *d) {
    if (d->width <= 0 || d->height <= 0 || d->in_channels <= 0 || d->out_channels <= 0)
        FYN_FAIL(FYN_ERR_INVALID, "conv: bad shape %dx%d %d->%d", d->width, d->height, d->in_channels, d->out_channels);
    if (d->kernel < 1 || d->kernel > 9 || !(d->kernel & 1)) FYN_FAIL(FYN_ERR_INVALID, "conv: kernel %d not in {1,3,5,7,9}", d->kernel);
    if (d->dilation < 1) FYN_FAIL(FYN_ERR_INVALID, "conv: dilation must be >= 1");
    if (d->in_padding < 0 || d->out_padding < 0 || d->res_padding < 0) FYN_FAIL(FYN_ERR_INVALID, "conv: negative padding");
    if (d->fractional) {
        // gpu/gpulayerfactory.cpp:447-457 (shallow only), fractionalconvlayerNxN_vanilla.cpp:44 (no dilation)
        if (d->flags & FYN_FLAG_DEEP) FYN_FAIL(FYN_ERR_UNSUPPORTED, "fractional convolution has no deep variant");
        if (d->dilation > 1) FYN_FAIL(FYN_ERR_UNSUPPORTED, "Dilations not supported for fractional convolution");
    }
    if ((d->flags & (FYN_FLAG_RELU_ON_RESIDUAL | FYN_FLAG_BATCHNORM_ON_RESIDUAL)) && !(d->flags & FYN_FLAG_RESIDUAL_INPUT))
        FYN_FAIL(FYN_ERR_INVALID, "conv: residual modifiers without RESIDUAL_INPUT");
    return FYN_OK;
}

// Folded epilogue parameters.  shallow: b' = b*s + beta in fp32 (convweightarrayKxKxNxM.cpp:167-181);
// deep with fp16 storage: the bias / BN texture is RGBA16F (deepconvlayerbase.cpp:371-394) -> values
// pass through fp16 (round to nearest) -- selected by `half_params`.
static void fold(const fyn_conv_desc *d, const float *wb, bool half_params, std::vector<float> &bias,
                 std::vector<float> &scale) {
    const int Co = d->out_channels, nOut = (Co + 3) / 4;
    bias.assign((size_t)nOut * 4, 0.f);
    scale.assign((size_t)nOut * 4, 0.f);
    const float *bn = wb + Co + (size_t)d->kernel * d->kernel * d->in_channels * Co;
    for (int o = 0; o < Co; o++) {
        float b = wb[o], s = 1.f;
        if (d->flags & FYN_FLAG_POST_BATCHNORM) {
            s = bn[o];
            b = b * s + bn[Co + o];
        }
        if (half_params) {
            b = fyn_half_round_host(b);
            s = fyn_half_round_host(s);
        }
        bias[o] = b;
        scale[o] = s;
    }
}

static int upload_floats(float **dptr, const std::vector<float> &h) {
    if (!*dptr) FYN_CUDA(cudaMalloc((void **)dptr, h.size() * sizeof(float)));
    FYN_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return FYN_OK;
}

extern "C" {

int fyn_conv2d_load_weights(fyn_op *op, const float *wb) {
    if (!op || op->kind != FYN_OP_CONV || !wb) FYN_FAIL(FYN_ERR_INVALID, "bad conv op / weights");
    const fyn_conv_desc &d = op->conv;
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    // Hot swap (weights reloaded after set-up): the images below are overwritten in place with blocking copies on the legacy
    // stream, which is NOT ordered against the engine's non-blocking streams -- wait until every forward pass that may
    // still read the old images has left the device (the reference serialises the swap on the GL command stream).
    if (op->d_w) FYN_CUDA(cudaDeviceSynchronize());
    const int Ci = d.in_channels, Co = d.out_channels, K = d.kernel;
    const int nIn = (Ci + 3) / 4, nOut = (Co + 3) / 4;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    // The weight precision rule follows the storage mode of the layer's tensors, which the op only
    // learns at run time; the reference decides at build time (HIGH_PRECISION).  We prepare both:
    // d_w holds fp32 weights for FYN_F32 storage; for deep layers a second set truncated to fp16
    // (gpu/floatconversion.cpp:44-58) is appended and selected when the input tensor is FYN_F16.
    std::vector<float> w((size_t)nOut * nIn * K * K * 16 * (deep ? 2 : 1), 0.f);
    const size_t half_off = (size_t)nOut * nIn * K * K * 16;
    const float *src = wb + Co;
    for (int o = 0; o < Co; o++)
        for (int ky = 0; ky < K; ky++)
            for (int kx = 0; kx < K; kx++)
                for (int c = 0; c < Ci; c++) {
                    float v = src[(((size_t)o * K + ky) * K + kx) * Ci + c];
                    size_t idx = ((((size_t)(o / 4) * nIn + c / 4) * K + ky) * K + kx) * 16 + (c % 4) * 4 + (o % 4);
                    w[idx] = v;
                    if (deep) w[half_off + idx] = fyn_half_trunc_host(v);
                }
    int rc = upload_floats(&op->d_w, w);
    if (rc) return rc;
    std::vector<float> bias, scale, both;
    fold(&d, wb, false, bias, scale);
    both = bias;
    both.insert(both.end(), scale.begin(), scale.end());
    if (deep) {
        std::vector<float> hb, hs;
        fold(&d, wb, true, hb, hs);
        both.insert(both.end(), hb.begin(), hb.end());
        both.insert(both.end(), hs.begin(), hs.end());
    }
    rc = upload_floats(&op->d_bias, both);
    if (rc) return rc;
    op->d_scale = op->d_bias + (size_t)nOut * 4;
    if (op->backend == 2) {
        // (re)packs the fp16 operand images of the tcgen05 families
        rc = deep ? fyn_conv_deep_tc_create(op, wb) : fyn_conv_tc_create(op, wb);
        if (rc) return rc;
    }
    return FYN_OK;
}

int fyn_conv2d_create(fyn_ctx *ctx, const fyn_conv_desc *desc, const float *wb, fyn_op **out) {
    if (!ctx || !desc || !wb || !out) FYN_FAIL(FYN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = validate(desc);
    if (rc) return rc;
    fyn_op *op = new fyn_op();
    op->ctx = ctx;
    op->kind = FYN_OP_CONV;
    op->conv = *desc;
    if (op->conv.dilation < 1) op->conv.dilation = 1;
    fyn_conv2d_output_size(desc, &op->Wo, &op->Ho);
    if (op->Wo <= 0 || op->Ho <= 0) {
        int wo = op->Wo, ho = op->Ho;
        delete op;
        FYN_FAIL(FYN_ERR_INVALID, "conv: empty output %dx%d", wo, ho);
    }
    op->backend = 1;
    if (desc->backend != 1 && (fyn_conv_tc_supported(desc, FYN_F16) || fyn_conv_deep_tc_supported(desc))) {
        op->backend = 2;
    } else if (desc->backend == 2) {
        delete op;
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "conv: tcgen05 kernel family does not cover this configuration");
    }
    rc = fyn_conv2d_load_weights(op, wb);
    if (rc) {
        fyn_op_destroy(op);
        return rc;
    }
    *out = op;
    return FYN_OK;
}

int fyn_conv2d_backend(const fyn_op *op) { return (op && op->kind == FYN_OP_CONV) ? op->backend : 0; }
int fyn_conv2d_last_kernel(const fyn_op *op) { return (op && op->kind == FYN_OP_CONV) ? op->lastKernel : 0; }

int fyn_conv2d_set_input_norm(fyn_op *op, const float *sb) {
    if (!op || op->kind != FYN_OP_CONV) FYN_FAIL(FYN_ERR_INVALID, "not a convolution op");
    if (!sb) {
        op->innorm = 0;
        return FYN_OK;
    }
    const fyn_conv_desc &d = op->conv;
    // padding texels of the stand-alone layer's output are zero, not bn(0): only kernels that never read padding qualify
    if (op->epilogue != FYN_EPILOGUE_NONE) FYN_FAIL(FYN_ERR_UNSUPPORTED, "conv: input batch-norm fusion and a fused epilogue function exclude each other");
    if (!op->dtc || d.kernel != 1 || d.in_channels % 64 != 0 || (d.flags & FYN_FLAG_PRE_CLIP))
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "conv: input batch-norm fusion needs a 1x1 layer of the deep-tiled tcgen05 family");
    const int C = d.in_channels;
    std::vector<float> h((size_t)2 * C);
    for (int c = 0; c < C; c++) {
        h[c] = sb[c];
        h[(size_t)C + c] = sb[(size_t)C + c];
    }
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    if (op->d_innorm) FYN_CUDA(cudaDeviceSynchronize());   // parameters of a live op: see fyn_conv2d_load_weights
    if (!op->d_innorm) FYN_CUDA(cudaMalloc((void **)&op->d_innorm, h.size() * sizeof(float)));
    FYN_CUDA(cudaMemcpy(op->d_innorm, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    op->innorm = 1;
    return FYN_OK;
}

int fyn_conv2d_set_epilogue(fyn_op *op, int function) {
    if (!op || op->kind != FYN_OP_CONV) FYN_FAIL(FYN_ERR_INVALID, "not a convolution op");
    if (function != FYN_EPILOGUE_NONE && function != FYN_EPILOGUE_SIGMOID) FYN_FAIL(FYN_ERR_INVALID, "conv: unknown epilogue function %d", function);
    // the deep-tiled tcgen05 family has no fused function: fusing would silently drop the layer to the CUDA-core kernel
    // (and collide with a fused input batch-norm, which only that family implements)
    if (function != FYN_EPILOGUE_NONE && (op->dtc || op->innorm))
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "conv: no fused epilogue function in the deep-tiled tcgen05 family");
    op->epilogue = function;
    return FYN_OK;
}

static int check_tensor(const fyn_tensor *t, const char *what, int w, int h, int c, int pad, bool deep) {
    if (!t) FYN_FAIL(FYN_ERR_INVALID, "conv: %s tensor is NULL", what);
    const fyn_tensor_desc &d = t->desc;
    if (d.width != w || d.height != h || d.channels != c || d.padding != pad || (d.order == FYN_ORDER_DEEP) != deep) {
        // single-tile deep == single-plane shallow when the layouts coincide (<= 4 channels)
        bool same_layout = c <= 4 && d.width == w && d.height == h && d.channels == c && d.padding == pad;
        if (!same_layout)
            FYN_FAIL(FYN_ERR_INVALID, "conv: %s tensor mismatch: got %dx%dx%d pad %d %s, need %dx%dx%d pad %d %s", what,
                     d.width, d.height, d.channels, d.padding, d.order == FYN_ORDER_DEEP ? "deep" : "shallow", w, h, c, pad,
                     deep ? "deep" : "shallow");
    }
    return FYN_OK;
}

int fyn_conv2d_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, void *stream) {
    if (!op || op->kind != FYN_OP_CONV) FYN_FAIL(FYN_ERR_INVALID, "not a conv op");
    const fyn_conv_desc &d = op->conv;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    int rc = check_tensor(in, "input", d.width, d.height, d.in_channels, d.in_padding, deep);
    if (rc) return rc;
    rc = check_tensor(out, "output", op->Wo, op->Ho, d.out_channels, d.out_padding, deep);
    if (rc) return rc;
    if (out->geom.packing != 4) FYN_FAIL(FYN_ERR_INVALID, "conv: output packing must be 4");
    if (d.flags & FYN_FLAG_RESIDUAL_INPUT) {
        rc = check_tensor(res, "residual", op->Wo, op->Ho, d.out_channels, d.res_padding, deep);
        if (rc) return rc;
        if (res->desc.batch != in->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "conv: residual batch mismatch");
    }
    if (in->desc.batch != out->desc.batch) FYN_FAIL(FYN_ERR_INVALID, "conv: batch mismatch %d vs %d", in->desc.batch, out->desc.batch);
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (op->backend == 2 && (op->tc || op->dtc)) {
        // > 0 means "tensor formats not covered by the tcgen05 family": use the direct kernel
        rc = op->dtc ? fyn_conv_deep_tc_run(op, in, res, out, s) : fyn_conv_tc_run(op, in, res, out, s);
        if (rc == 0 && !op->dtc) op->lastKernel = 2;
        if (rc <= 0) return rc;
    }
    op->lastKernel = 1;
    if (op->innorm) FYN_FAIL(FYN_ERR_UNSUPPORTED, "conv: a fused input batch-norm needs the deep-tiled tcgen05 kernel (fp16 tensors)");
    // direct family; deep layers on fp16 tensors use the fp16-truncated weight / fp16 bias sets
    fyn_op view = *op;
    if (deep && in->desc.dtype == FYN_F16) {
        const int nIn = (d.in_channels + 3) / 4, nOut = (d.out_channels + 3) / 4;
        view.d_w = op->d_w + (size_t)nOut * nIn * d.kernel * d.kernel * 16;
        view.d_bias = op->d_bias + (size_t)nOut * 8;
        view.d_scale = view.d_bias + (size_t)nOut * 4;
    }
    return fyn_conv_direct_run(&view, in, res, out, s);
}

int fyn_op_destroy(fyn_op *op) {
    if (!op) return FYN_OK;
    cudaSetDevice(op->ctx->device);
    if (op->tc) fyn_conv_tc_destroy(op);
    if (op->dtc) fyn_conv_deep_tc_destroy(op);
    if (op->d_w) cudaFree(op->d_w);
    if (op->d_bias) cudaFree(op->d_bias);
    if (op->d_innorm) cudaFree(op->d_innorm);
    delete op;
    return FYN_OK;
}

}  // extern "C"
