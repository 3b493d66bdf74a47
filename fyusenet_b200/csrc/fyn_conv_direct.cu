// fyn_conv_direct.cu -- generic CUDA-core convolution (every shape / layout / dtype the ABI accepts).
//
// This is the exact-fp32 kernel family: it serves FYN_F32 storage (the reference's HIGH_PRECISION
// mode), odd shapes the tcgen05 family does not cover, and is the in-library cross-check for the
// tensor-core kernels.  Semantics restated from:
//   shallow   gpu/vanilla/convlayerbase_vanilla.cpp:347-371 (texel centre P + s*(ds*o+0.5); vertical
//             taps shift by s*(ky-m): no vertical dilation), shaders/vanilla/conv3x3.frag:14-21,
//             conv9x9.frag:40-48, conv.inc:1-34, residual.inc; CLAMP_TO_EDGE addressing
//   fraction  gpu/vanilla/fractionalconvlayerNxN_vanilla.cpp:43-51,88; shaders/vanilla/fraconv3x3.frag:13-19
//             (quirk: taps -2s,-s,0), fractional.inc:11-12 vs :69-70 (quirk: act on first tap only)
//   deep      gpu/deep/deeptiler.cpp:109-203 (base texel P + ds*o inside the tile),
//             shaders/deep/deepconv3x3_tiled.frag:32-34, deepconv1x1_tiled.frag:24-34, batchnorm.inc:2-8
// One thread = one output texel (4 output channels); fp32 accumulation over all taps / input planes,
// epilogue = *bnScale + foldedBias (+ residual [relu] [*bnScale]).
#include "fyn_internal.h"

struct DirectConvArgs {
    TView in, out, res;
    const float4 *w;      // [nOut][nIn][K][K][4ci] float4 over co
    const float4 *bias;   // [nOut]
    const float4 *scale;  // [nOut]
    int K, ds, dilx, dily, m;
    float step;
    int fractional;
    int tapx[9];          // horizontal tap offsets (units of step / dilation)
    int Wo, Ho, nIn, nOut, batch, outP, resP, xBlocks, yBlocks;
    ActParams act;
    int actFirstOnly, hasRes, reluRes, bnRes;
    int epilogue;         // FYN_EPILOGUE_*
    int ksplit;           // input planes are split over blockDim.z slices and reduced through shared memory (small grids)
};

// Small grids (ResNet-50's 7x7 and 14x14 layers at batch 1: ~128 blocks whose threads each walk 128 input planes x 9 taps)
// are bound by the serial depth of that loop, so the input planes are split over blockDim.z slices whose partial sums
// meet in shared memory (fp32, fixed order: slice 0 + 1 + ... -- deterministic).
template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 1024 : 128) k_conv_direct(const DirectConvArgs a) {
    __shared__ float4 partial[SPLIT ? 7 : 1][128];
    // linear block index -> (image, output plane, y block, x block); grid.x only (no 65535 limits)
    unsigned bid = blockIdx.x;
    const int xb = bid % a.xBlocks;
    bid /= a.xBlocks;
    const int yb = bid % a.yBlocks;
    bid /= a.yBlocks;
    const int op = bid % a.nOut;
    const int n = bid / a.nOut;
    const int xo = xb * 32 + threadIdx.x;
    const int yo = yb * 4 + threadIdx.y;
    const bool active = xo < a.Wo && yo < a.Ho;
    if (!active && !SPLIT) return;
    const int P = a.in.P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float cx = (float)P + a.step * ((float)(a.ds * xo) + 0.5f);
    const float cy = (float)P + a.step * ((float)(a.ds * yo) + 0.5f);
    const int per = SPLIT ? (a.nIn + a.ksplit - 1) / a.ksplit : a.nIn;
    const int ip0 = SPLIT ? threadIdx.z * per : 0, ip1 = min(a.nIn, ip0 + per);
    for (int ip = ip0; active && ip < ip1; ip++) {
        const float4 *wp = a.w + (size_t)(op * a.nIn + ip) * a.K * a.K * 4;
        for (int ky = 0; ky < a.K; ky++) {
            const int iy = a.fractional ? (int)floorf(cy + a.step * (float)(ky - a.m))
                                        : P + a.ds * yo + (ky - a.m) * a.dily;
            for (int kx = 0; kx < a.K; kx++) {
                const int ix = a.fractional ? (int)floorf(cx + a.step * (float)a.tapx[kx])
                                            : P + a.ds * xo + a.tapx[kx] * a.dilx;
                float4 v = fyn_fetch(a.in, n, ip, ix, iy);
                if (!(a.actFirstOnly && kx > 0)) v = fyn_act4(v, a.act);
                const float4 w0 = __ldg(wp + 0), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                wp += 4;
                acc.x = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, acc.x))));
                acc.y = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, acc.y))));
                acc.z = fmaf(v.x, w0.z, fmaf(v.y, w1.z, fmaf(v.z, w2.z, fmaf(v.w, w3.z, acc.z))));
                acc.w = fmaf(v.x, w0.w, fmaf(v.y, w1.w, fmaf(v.z, w2.w, fmaf(v.w, w3.w, acc.w))));
            }
        }
    }
    if (SPLIT) {
        const int t = threadIdx.y * 32 + threadIdx.x;
        if (threadIdx.z > 0) partial[threadIdx.z - 1][t] = acc;
        __syncthreads();
        if (threadIdx.z > 0 || !active) return;
        for (int z = 1; z < a.ksplit; z++) {
            const float4 q = partial[z - 1][t];
            acc.x += q.x;
            acc.y += q.y;
            acc.z += q.z;
            acc.w += q.w;
        }
    }
    const float4 s = __ldg(a.scale + op), b = __ldg(a.bias + op);
    float4 r = make_float4(fmaf(acc.x, s.x, b.x), fmaf(acc.y, s.y, b.y), fmaf(acc.z, s.z, b.z), fmaf(acc.w, s.w, b.w));
    if (a.hasRes) {
        float4 q = fyn_fetch(a.res, n, op, a.resP + xo, a.resP + yo);
        if (a.reluRes) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
        if (a.bnRes) q = make_float4(q.x * s.x, q.y * s.y, q.z * s.z, q.w * s.w);
        r.x += q.x;
        r.y += q.y;
        r.z += q.z;
        r.w += q.w;
    }
    if (a.epilogue == FYN_EPILOGUE_SIGMOID) {
        // fused FunctionLayer: same values as storing the convolution and running the sigmoid layer on the stored texel
        if (a.out.dtype == FYN_F16) r = make_float4(fyn_round_half(r.x), fyn_round_half(r.y), fyn_round_half(r.z), fyn_round_half(r.w));
        r = make_float4(fyn_sigmoid(r.x), fyn_sigmoid(r.y), fyn_sigmoid(r.z), fyn_sigmoid(r.w));
    }
    fyn_store_texel(a.out, n, op, a.outP + xo, a.outP + yo, r);
}

int fyn_conv_direct_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t s) {
    const fyn_conv_desc &d = op->conv;
    DirectConvArgs a{};
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.res = fyn_make_view(res);
    a.w = reinterpret_cast<const float4 *>(op->d_w);
    a.bias = reinterpret_cast<const float4 *>(op->d_bias);
    a.scale = reinterpret_cast<const float4 *>(op->d_scale);
    a.K = d.kernel;
    a.m = (d.kernel - 1) / 2;
    a.ds = d.downsample;
    const bool deep = (d.flags & FYN_FLAG_DEEP) != 0;
    a.dilx = d.dilation;
    a.dily = deep ? d.dilation : 1;  // shallow path has no vertical dilation (convlayerbase_vanilla.cpp:364)
    a.fractional = d.fractional;
    a.step = d.fractional ? d.source_step : 1.f;
    for (int k = 0; k < d.kernel; k++) a.tapx[k] = k - a.m;
    if (d.fractional && d.kernel == 3 && (d.quirks & FYN_QUIRK_FRAC3_ASYM)) {
        a.tapx[0] = -2;
        a.tapx[1] = -1;
        a.tapx[2] = 0;
    }
    a.actFirstOnly = d.fractional && (d.quirks & FYN_QUIRK_FRAC_ACT_FIRST);
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.nIn = (d.in_channels + 3) / 4;
    a.nOut = (d.out_channels + 3) / 4;
    a.batch = in->desc.batch;
    a.outP = d.out_padding;
    a.resP = d.res_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    a.hasRes = (d.flags & FYN_FLAG_RESIDUAL_INPUT) != 0;
    a.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
    a.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
    a.epilogue = op->epilogue;
    a.xBlocks = (a.Wo + 31) / 32;
    a.yBlocks = (a.Ho + 3) / 4;
    long long blocks = (long long)a.xBlocks * a.yBlocks * a.nOut * a.batch;
    if (blocks > 0x7fffffffLL) FYN_FAIL(FYN_ERR_UNSUPPORTED, "direct conv: %lld blocks exceed the grid limit", blocks);
    // split the input planes when the grid alone cannot fill the GPU and the reduction is deep
    a.ksplit = 1;
    const long long depth = (long long)a.nIn * a.K * a.K;
    const int sms = op->ctx->prop.multiProcessorCount;
    while (a.ksplit < 8 && blocks * a.ksplit < 4LL * sms && depth / (a.ksplit * 2) >= 36 && a.nIn >= a.ksplit * 2) a.ksplit *= 2;
    dim3 block(32, 4, a.ksplit), grid((unsigned)blocks);
    if (a.ksplit > 1) k_conv_direct<true><<<grid, block, 0, s>>>(a);
    else k_conv_direct<false><<<grid, block, 0, s>>>(a);
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}
