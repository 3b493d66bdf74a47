/* ---------------------------------------------------------------------------------------------
 * fyn_oracle.c -- CPU restatement of the FyuseNet GPU-layer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fyusenet_b200/) may link, import or
 * call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: the reference's shader path cannot be executed in this container (no GL/EGL,
 * see SURVEY.md section 8c) and its data/ weights are git-LFS stubs, so this oracle is pinned
 * against the reference's own *layer-level* known-answer tests (unit_tests/convlayertests.cpp,
 * pooltests.cpp, misctests.cpp, networktests.cpp -- replayed in tests/test_oracle_kat.py).
 * Whole-network outputs are "parity unpinned" (the reference stores no golden outputs).
 *
 * All formulas are restated from the reference sources cited next to each function (paths are
 * relative to /root/reference/fyusenet unless noted).  No reference code is copied: the
 * reference expresses these semantics as GLSL shader passes + GL state, this file expresses
 * them as loops over "textures" (RGBA float images with clamp-to-edge addressing).
 *
 * Data model
 *   texture      : w x h pixels of 4 floats, CLAMP_TO_EDGE + NEAREST (base/buffermanager.cpp:657-670)
 *   shallow      : ceil(C/4) textures of (W+2P) x (H+2P); channel c -> texture c/4, lane c%4
 *                  (unit_tests/layertestbase.cpp:281-317)
 *   deep         : one texture; tile i (channels 4i..4i+3) at pixel (P+(i%tx)(W+P), P+(i/tx)(H+P))
 *                  (gpu/deep/deeptiler.cpp:63-95, unit_tests/layertestbase.cpp:235-280)
 *   CHW          : [C][H][W] float32 without padding = the reference's dump format
 *                  (base/layerbase.h:160-172)
 *
 * Precision modes (gpu/gpulayerbase.h:100-110, README.md:61-67)
 *   FYO_FP32        : HIGH_PRECISION build, everything float32
 *   FYO_FP16_STORE  : results rounded to fp16 when written to a texture; deep conv weights
 *                     truncated to fp16 (gpu/floatconversion.cpp:44-58), deep bias/BN texture
 *                     fp16 (gpu/deep/deepconvlayerbase.cpp:371-394); accumulation in fp32.
 *                     This is what a one-kernel-per-layer fp32-accumulate backend computes.
 *   FYO_FP16_BLEND  : like STORE, plus the render target is rounded to fp16 after EVERY blend
 *                     pass (one pass per (input plane|tile, kernel row)) = the reference default.
 * ------------------------------------------------------------------------------------------- */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fyn_oracle.h"

/* ------------------------------------------------------------------------------------------- */
/* fp16 helpers                                                                                */
/* ------------------------------------------------------------------------------------------- */

/* Round-to-nearest-even float -> half -> float (what a GL driver does on an RGBA16F store). */
static inline float h_rn(float x) {
    _Float16 h = (_Float16)x;
    return (float)h;
}

/* Truncating float -> half -> float following the table scheme of
 * gpu/floatconversion.cpp:44-58 (base/shift tables built in :85-127): the mantissa is shifted
 * right without rounding; |x| < 2^-24 -> 0; exponents > 15 -> inf; denormals are truncated. */
static inline uint16_t h_trunc_bits(float x) {
    uint32_t f;
    memcpy(&f, &x, 4);
    uint32_t sign = (f >> 16) & 0x8000u;
    int e = (int)((f >> 23) & 0xff) - 127;
    uint32_t man = f & 0x007fffffu;
    if (e < -24) return (uint16_t)sign;
    if (e < -14) return (uint16_t)(sign | ((0x0400u >> (-e - 14)) + (man >> (-e - 1))));
    if (e <= 15) return (uint16_t)(sign | (((uint32_t)(e + 15) << 10) + (man >> 13)));
    if (e < 128) return (uint16_t)(sign | 0x7c00u);
    return (uint16_t)(sign | (0x7c00u + (man >> 13)));
}

static inline float h_bits_to_float(uint16_t h) {
    _Float16 v;
    memcpy(&v, &h, 2);
    return (float)v;
}

static inline float h_trunc(float x) { return h_bits_to_float(h_trunc_bits(x)); }

float fyo_half_round(float x) { return h_rn(x); }
float fyo_half_trunc(float x) { return h_trunc(x); }
uint16_t fyo_half_trunc_bits(float x) { return h_trunc_bits(x); }

static inline float store(float x, int prec) { return prec == FYO_FP32 ? x : h_rn(x); }

/* ------------------------------------------------------------------------------------------- */
/* textures                                                                                    */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
    int w, h;
    float *px; /* [h][w][4] */
} tex_t;

static int tex_alloc(tex_t *t, int w, int h) {
    t->w = w;
    t->h = h;
    t->px = (float *)calloc((size_t)w * h * 4, sizeof(float));
    return t->px ? 0 : -1;
}

static void tex_free(tex_t *t) {
    free(t->px);
    t->px = NULL;
}

/* CLAMP_TO_EDGE + NEAREST texel fetch (base/buffermanager.cpp:657-670) */
static inline const float *tex_fetch(const tex_t *t, int x, int y) {
    if (x < 0) x = 0;
    if (x >= t->w) x = t->w - 1;
    if (y < 0) y = 0;
    if (y >= t->h) y = t->h - 1;
    return t->px + ((size_t)y * t->w + x) * 4;
}

/* activation at fetch (gpu/shaders/activation.inc:3-18) */
static inline float act1(float v, const fyo_act *a) {
    switch (a->type) {
    case FYO_ACT_RELU: return v > 0.f ? v : 0.f;
    case FYO_ACT_LEAKY: {
        /* sg = step(0,x); (sg + leak*(1-sg))*x */
        float sg = (v >= 0.f) ? 1.f : 0.f;
        return (sg + a->leak * (1.f - sg)) * v;
    }
    case FYO_ACT_CLIP: {
        float m = v > a->lo ? v : a->lo;
        return m < a->hi ? m : a->hi;
    }
    default: return v;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* layouts                                                                                     */
/* ------------------------------------------------------------------------------------------- */

/* cpu/cpubuffershape.cpp:430-447: minimise |x-y| + (x*y - tiles) over 1<=y<=x, first minimum
 * in (y ascending, x ascending) enumeration order. */
void fyo_deep_tiling(int channels, int *tx, int *ty) {
    int tiles = (channels + 3) / 4;
    float best = 1e30f;
    int bx = 1, by = 1;
    for (int y = 1; y <= tiles; y++) {
        for (int x = y; x <= tiles; x++) {
            if (x * y >= tiles) {
                float cost = (float)(x - y) + (float)(x * y - tiles);
                if (cost < best) {
                    best = cost;
                    bx = x;
                    by = y;
                }
            }
        }
    }
    *tx = bx;
    *ty = by;
}

void fyo_deep_texture_size(int channels, int w, int h, int pad, int *tw, int *th) {
    int tx, ty;
    fyo_deep_tiling(channels, &tx, &ty);
    *tw = tx * (w + pad) + pad; /* gpu/deep/deeptiler.cpp:91-94 */
    *th = ty * (h + pad) + pad;
}

/* unit_tests/layertestbase.cpp:235-280 (deep) */
void fyo_pack_deep(const float *chw, int C, int H, int W, int pad, float *texels) {
    int tx, ty, tw, th;
    fyo_deep_tiling(C, &tx, &ty);
    fyo_deep_texture_size(C, W, H, pad, &tw, &th);
    memset(texels, 0, (size_t)tw * th * 4 * sizeof(float));
    for (int c = 0; c < C; c++) {
        int tile = c / 4, lane = c % 4;
        int ox = pad + (tile % tx) * (W + pad), oy = pad + (tile / tx) * (H + pad);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                texels[((size_t)(oy + y) * tw + ox + x) * 4 + lane] = chw[((size_t)c * H + y) * W + x];
    }
}

/* gpu/deep/deeplayerbase.cpp:137-176 (copyResult) */
void fyo_unpack_deep(const float *texels, int C, int H, int W, int pad, float *chw) {
    int tx, ty, tw, th;
    fyo_deep_tiling(C, &tx, &ty);
    fyo_deep_texture_size(C, W, H, pad, &tw, &th);
    for (int c = 0; c < C; c++) {
        int tile = c / 4, lane = c % 4;
        int ox = pad + (tile % tx) * (W + pad), oy = pad + (tile / tx) * (H + pad);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                chw[((size_t)c * H + y) * W + x] = texels[((size_t)(oy + y) * tw + ox + x) * 4 + lane];
    }
}

/* unit_tests/layertestbase.cpp:281-317 (shallow): planes[ceil(C/4)][H+2P][W+2P][4] */
void fyo_pack_shallow(const float *chw, int C, int H, int W, int pad, float *planes) {
    int np = (C + 3) / 4, pw = W + 2 * pad, ph = H + 2 * pad;
    memset(planes, 0, (size_t)np * pw * ph * 4 * sizeof(float));
    for (int c = 0; c < C; c++) {
        float *pl = planes + (size_t)(c / 4) * pw * ph * 4;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                pl[((size_t)(y + pad) * pw + x + pad) * 4 + (c % 4)] = chw[((size_t)c * H + y) * W + x];
    }
}

/* gpu/gpulayerbase.cpp:525-560 (copyResult) */
void fyo_unpack_shallow(const float *planes, int C, int H, int W, int pad, float *chw) {
    int pw = W + 2 * pad, ph = H + 2 * pad;
    for (int c = 0; c < C; c++) {
        const float *pl = planes + (size_t)(c / 4) * pw * ph * 4;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                chw[((size_t)c * H + y) * W + x] = pl[((size_t)(y + pad) * pw + x + pad) * 4 + (c % 4)];
    }
}

/* internal: shallow tensor as an array of textures */
static tex_t *shallow_from_chw(const float *chw, int C, int H, int W, int pad) {
    int np = (C + 3) / 4;
    tex_t *t = (tex_t *)calloc((size_t)np, sizeof(tex_t));
    if (!t) return NULL;
    for (int p = 0; p < np; p++)
        if (tex_alloc(&t[p], W + 2 * pad, H + 2 * pad)) return NULL;
    for (int c = 0; c < C; c++) {
        tex_t *pl = &t[c / 4];
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                pl->px[((size_t)(y + pad) * pl->w + x + pad) * 4 + (c % 4)] = chw[((size_t)c * H + y) * W + x];
    }
    return t;
}

static void shallow_free(tex_t *t, int C) {
    if (!t) return;
    for (int p = 0; p < (C + 3) / 4; p++) tex_free(&t[p]);
    free(t);
}

/* ------------------------------------------------------------------------------------------- */
/* output geometry                                                                             */
/* ------------------------------------------------------------------------------------------- */

/* regular: gpu/convlayerbase.cpp:47-48; fractional: gpu/vanilla/fractionalconvlayerNxN_vanilla.cpp:46-49 */
void fyo_conv2d_outdims(const fyo_conv *p, int *Wo, int *Ho) {
    if (p->fractional) {
        *Wo = (int)((float)p->width / (p->sourceStep * (float)p->downsample));
        *Ho = (int)((float)p->height / (p->sourceStep * (float)p->downsample));
    } else {
        *Wo = p->width / p->downsample;
        *Ho = p->height / p->downsample;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* shallow convolution (regular + fractional)                                                  */
/* ------------------------------------------------------------------------------------------- */

/* Folded bias / scale per output channel.
 * shallow: gpu/convweightarrayKxKxNxM.cpp:151-181 (b' = b*s + beta), fp32 uniforms.
 * deep   : gpu/deep/deepconvlayerbase.cpp:353-394, stored in an RGBA16F texture unless HIGH_PRECISION. */
static void fold_bias(const fyo_conv *p, const float *wb, float *bias, float *scale) {
    int Co = p->outChannels, Ci = p->inChannels, K = p->kernel;
    const float *bn = wb + Co + (size_t)K * K * Ci * Co;
    for (int o = 0; o < Co; o++) {
        float b = wb[o], s = 1.f;
        if (p->flags & FYO_POST_BATCHNORM) {
            s = bn[o];
            b = b * s + bn[Co + o];
        }
        if (p->deep && p->prec != FYO_FP32) {
            b = h_rn(b);
            s = h_rn(s);
        }
        bias[o] = b;
        scale[o] = s;
    }
}

/*
 * Shallow conv.  Sources restated:
 *   geometry   gpu/vanilla/convlayerbase_vanilla.cpp:347-371  (texel centre X = P + s*(ds*xo+0.5),
 *              vertical tap = quad shifted by s*(ky-m) texels -- NO vertical dilation)
 *   taps       gpu/shaders/vanilla/conv3x3.frag:14-21, conv9x9.frag:40-48 (textureOffset (kx-m)*dil)
 *              gpu/shaders/vanilla/fraconv3x3.frag:13-19 (Q1: taps -2s,-s,0), fraconv9x9.frag:13-31,
 *              fractional.inc:11-12 vs :69-70 (Q2: activation only on the first horizontal tap)
 *   math       gpu/shaders/vanilla/conv.inc:1-34 ((M*pix)*bnscale per tap, summed), residual.inc
 *   passes     gpu/vanilla/convlayerNxN_vanilla.cpp:104-127 (for input plane, for kernel row;
 *              bias-by-clear if outPad==0 else bias in pass (plane 0,row K-1); residual in the same pass)
 */
static int conv_shallow(const fyo_conv *p, const float *in_chw, const float *wb, const float *res_chw,
                        float *out_chw) {
    const int W = p->width, H = p->height, Ci = p->inChannels, Co = p->outChannels, K = p->kernel;
    const int P = p->inPadding, ds = p->downsample, dil = p->dilation, m = (K - 1) / 2;
    const float s = p->fractional ? p->sourceStep : 1.f;
    int Wo, Ho;
    fyo_conv2d_outdims(p, &Wo, &Ho);
    const int nip = (Ci + 3) / 4;
    tex_t *in = shallow_from_chw(in_chw, Ci, H, W, P);
    if (!in) return -1;
    float *bias = (float *)malloc(sizeof(float) * Co), *scale = (float *)malloc(sizeof(float) * Co);
    fold_bias(p, wb, bias, scale);
    /* weights [Co][K][K][Ci] (base/convlayerinterface.h:31-57) re-ordered to [ip][ky][kx][c][Co]
     * so the per-output-channel loop is contiguous; the arithmetic order per output channel
     * (sum over the 4 lanes of one texel, then *bnscale, then add to the pass) is unchanged. */
    const int nip4 = nip * 4;
    float *wT = (float *)calloc((size_t)nip4 * K * K * Co, sizeof(float));
    for (int o = 0; o < Co; o++)
        for (int ky = 0; ky < K; ky++)
            for (int kx = 0; kx < K; kx++)
                for (int c = 0; c < Ci; c++)
                    wT[((((size_t)(c / 4) * K + ky) * K + kx) * 4 + (c % 4)) * Co + o] =
                        wb[Co + (((size_t)o * K + ky) * K + kx) * Ci + c];
    const int blend = (p->prec == FYO_FP16_BLEND);
    const int postbn = (p->flags & FYO_POST_BATCHNORM) != 0;
    const int relu_res = (p->flags & FYO_RELU_ON_RESIDUAL) != 0;
    const int bn_res = (p->flags & FYO_BATCHNORM_ON_RESIDUAL) != 0;
    const int has_res = (p->flags & FYO_RESIDUAL_INPUT) != 0 && res_chw;
    /* horizontal tap offsets in units of s */
    int *tap = (int *)malloc(sizeof(int) * K);
    for (int k = 0; k < K; k++) tap[k] = k - m;
    if (p->fractional && K == 3 && (p->quirks & FYO_Q1_FRAC3_ASYM)) {
        tap[0] = -2;
        tap[1] = -1;
        tap[2] = 0;
    }
    const int act_first_only = p->fractional && (p->quirks & FYO_Q2_FRAC_ACT_FIRST);

#pragma omp parallel for schedule(dynamic, 4)
    for (int yo = 0; yo < Ho; yo++) {
        float *acc = (float *)malloc(sizeof(float) * Co * 3);
        float *pass = acc + Co, *tmp = acc + 2 * Co;
        for (int xo = 0; xo < Wo; xo++) {
            /* render-target initial value: cleared to the (folded) bias when the output is
             * unpadded (fp16 target in BLEND mode), else cleared to 0 and the bias is added in
             * the shader on pass (plane 0, row K-1) (convlayerbase_vanilla.cpp:280-292) */
            for (int o = 0; o < Co; o++) acc[o] = (p->outPadding == 0) ? (blend ? h_rn(bias[o]) : bias[o]) : 0.f;
            const float cx = (float)P + s * ((float)(ds * xo) + 0.5f);
            const float cy = (float)P + s * ((float)(ds * yo) + 0.5f);
            for (int ip = 0; ip < nip; ip++) {
                for (int ky = 0; ky < K; ky++) {
                    for (int o = 0; o < Co; o++) pass[o] = 0.f;
                    int iy = p->fractional ? (int)floorf(cy + s * (float)(ky - m)) : (P + ds * yo + (ky - m));
                    for (int kx = 0; kx < K; kx++) {
                        int ix = p->fractional ? (int)floorf(cx + s * (float)tap[kx]) : (P + ds * xo + (kx - m) * dil);
                        const float *px = tex_fetch(&in[ip], ix, iy);
                        float v[4];
                        int do_act = !(act_first_only && kx > 0);
                        for (int c = 0; c < 4; c++) v[c] = do_act ? act1(px[c], &p->act) : px[c];
                        const float *w = wT + ((((size_t)ip * K + ky) * K + kx) * 4) * Co;
                        for (int o = 0; o < Co; o++) tmp[o] = w[o] * v[0];
                        for (int c = 1; c < 4; c++)
                            for (int o = 0; o < Co; o++) tmp[o] += w[(size_t)c * Co + o] * v[c];
                        if (postbn)
                            for (int o = 0; o < Co; o++) pass[o] += tmp[o] * scale[o];
                        else
                            for (int o = 0; o < Co; o++) pass[o] += tmp[o];
                    }
                    if (ip == 0 && ky == K - 1) {
                        if (p->outPadding > 0)
                            for (int o = 0; o < Co; o++) pass[o] += bias[o];
                        if (has_res) {
                            for (int o = 0; o < Co; o++) {
                                float r = res_chw[((size_t)o * Ho + yo) * Wo + xo];
                                if (relu_res) r = r > 0.f ? r : 0.f;
                                if (bn_res) r *= scale[o];
                                pass[o] += r;
                            }
                        }
                    }
                    for (int o = 0; o < Co; o++) acc[o] = blend ? h_rn(acc[o] + pass[o]) : acc[o] + pass[o];
                }
            }
            for (int o = 0; o < Co; o++) out_chw[((size_t)o * Ho + yo) * Wo + xo] = store(acc[o], p->prec);
        }
        free(acc);
    }
    free(wT);
    free(tap);
    free(bias);
    free(scale);
    shallow_free(in, Ci);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* deep convolution / GEMM                                                                     */
/* ------------------------------------------------------------------------------------------- */

/*
 * Deep conv on the tiled texture.  Sources restated:
 *   tiling     gpu/deep/deeptiler.cpp:63-95,109-203 (base texel inside a tile = P + ds*o)
 *   taps       gpu/shaders/deep/deepconv3x3_tiled.frag:32-34 (textureOffset +-dil horizontally),
 *              vertical displacement (ky-m)*dil via gpu/deep/deepconvlayerbase.cpp:701-709
 *   weights    gpu/deep/deepconvlayerbase.cpp:293-349 (fp16 truncated unless HIGH_PRECISION),
 *              gpu/shaders/deep/computeconv.inc:2-12 (tex*weights)
 *   passes     gpu/shaders/deep/deepconv3x3_tiled.vert:54-56 (instance -> input tile = i/K, row = i%K),
 *              deepconv1x1_tiled.frag:24-34 + batchnorm.inc:2-8 (instance 0: *scale + foldedBias
 *              [+ residual(*scale if BN on residual)], others: *scale only)
 * Reads use whole-texture clamp-to-edge, so an under-padded kernel spills into the neighbouring
 * tile exactly as in the reference (gpu/gpulayerbase.h:84-90).
 */
static int conv_deep(const fyo_conv *p, const float *in_chw, const float *wb, const float *res_chw,
                     float *out_chw) {
    const int W = p->width, H = p->height, Ci = p->inChannels, Co = p->outChannels, K = p->kernel;
    const int P = p->inPadding, ds = p->downsample, dil = p->dilation, m = (K - 1) / 2;
    int Wo, Ho;
    fyo_conv2d_outdims(p, &Wo, &Ho);
    int itx, ity, tw, th;
    fyo_deep_tiling(Ci, &itx, &ity);
    fyo_deep_texture_size(Ci, W, H, P, &tw, &th);
    tex_t in;
    if (tex_alloc(&in, tw, th)) return -1;
    fyo_pack_deep(in_chw, Ci, H, W, P, in.px);
    const int nit = (Ci + 3) / 4;
    float *bias = (float *)malloc(sizeof(float) * Co), *scale = (float *)malloc(sizeof(float) * Co);
    fold_bias(p, wb, bias, scale);
    /* weights, optionally fp16-truncated */
    size_t nw = (size_t)Co * K * K * Ci;
    float *wt = (float *)malloc(sizeof(float) * nw);
    for (size_t i = 0; i < nw; i++) wt[i] = (p->prec == FYO_FP32) ? wb[Co + i] : h_trunc(wb[Co + i]);
    const int blend = (p->prec == FYO_FP16_BLEND);
    const int postbn = (p->flags & FYO_POST_BATCHNORM) != 0;
    const int relu_res = (p->flags & FYO_RELU_ON_RESIDUAL) != 0;
    const int bn_res = (p->flags & FYO_BATCHNORM_ON_RESIDUAL) != 0;
    const int has_res = (p->flags & FYO_RESIDUAL_INPUT) != 0 && res_chw;

#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < Co; o++) {
        for (int yo = 0; yo < Ho; yo++) {
            for (int xo = 0; xo < Wo; xo++) {
                float acc = 0.f;
                for (int it = 0; it < nit; it++) {
                    int bx = P + (it % itx) * (W + P) + ds * xo;
                    int by = P + (it / itx) * (H + P) + ds * yo;
                    int nc = Ci - it * 4 < 4 ? Ci - it * 4 : 4;
                    for (int ky = 0; ky < K; ky++) {
                        float pass = 0.f;
                        for (int kx = 0; kx < K; kx++) {
                            const float *px = tex_fetch(&in, bx + (kx - m) * dil, by + (ky - m) * dil);
                            const float *w = wt + (((size_t)o * K + ky) * K + kx) * Ci + it * 4;
                            float t = 0.f;
                            for (int c = 0; c < nc; c++) t += act1(px[c], &p->act) * w[c];
                            pass += t;
                        }
                        if (it == 0 && ky == 0) {
                            pass = postbn ? pass * scale[o] + bias[o] : pass + bias[o];
                            if (has_res) {
                                float r = res_chw[((size_t)o * Ho + yo) * Wo + xo];
                                if (relu_res) r = r > 0.f ? r : 0.f;
                                if (bn_res) r *= scale[o];
                                pass += r;
                            }
                        } else if (postbn) {
                            pass *= scale[o];
                        }
                        acc = blend ? h_rn(acc + pass) : acc + pass;
                    }
                }
                out_chw[((size_t)o * Ho + yo) * Wo + xo] = store(acc, p->prec);
            }
        }
    }
    free(wt);
    free(bias);
    free(scale);
    tex_free(&in);
    return 0;
}

int fyo_conv2d(const fyo_conv *p, const float *in_chw, const float *wb, const float *res_chw, float *out_chw) {
    if (p->kernel < 1 || !(p->kernel & 1)) return -2;
    if (p->fractional && p->deep) return -3; /* gpu/gpulayerfactory.cpp:447-457: shallow only */
    return p->deep ? conv_deep(p, in_chw, wb, res_chw, out_chw) : conv_shallow(p, in_chw, wb, res_chw, out_chw);
}

/* ------------------------------------------------------------------------------------------- */
/* pooling (deep layout)                                                                       */
/* ------------------------------------------------------------------------------------------- */

/*
 * gpu/deep/deeppoolinglayer.cpp:38-54 (global: pool = downsample = (W,H)),
 * gpu/shaders/deep/deepmaxpool.frag:12-61 (window offsets [-P, pool-1-P] around texel P+ds*o;
 *   Q7: for pool==3 the third column is fetched WITHOUT activate()),
 * gpu/shaders/deep/deepavgpool.frag:15-58 (offsets [0,pool-1] for 2/4 and the loop version,
 *   [-1,1] for 3; mean = sum * 1/n).
 * Channels are independent, so the tiled texture is emulated per channel with zero padding
 * and whole-texture clamping only matters at the outer border (kept via clamp on the tile
 * when the tensor has a single tile; multi-tile spill is not modelled: P >= needed in all
 * reference uses).
 */
int fyo_pool2d(const fyo_pool *p, const float *in_chw, float *out_chw) {
    const int W = p->width, H = p->height, C = p->channels, P = p->inPadding;
    int px = p->global ? W : p->poolX, py = p->global ? H : p->poolY;
    int dx = p->global ? W : p->downsample, dy = p->global ? H : p->downsample;
    int Wo = W / dx, Ho = H / dy;
    int pw = W + 2 * P, ph = H + 2 * P;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        float *pad = (float *)calloc((size_t)pw * ph, sizeof(float));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) pad[(size_t)(y + P) * pw + x + P] = in_chw[((size_t)c * H + y) * W + x];
        for (int yo = 0; yo < Ho; yo++) {
            for (int xo = 0; xo < Wo; xo++) {
                int bx = P + dx * xo, by = P + dy * yo;
                float r;
                if (p->isMax) {
                    int off = -P;
                    r = -INFINITY;
                    int first = 1;
                    for (int j = 0; j < py; j++) {
                        for (int i = 0; i < px; i++) {
                            int x = bx + off + i, y = by + off + j;
                            x = x < 0 ? 0 : (x >= pw ? pw - 1 : x);
                            y = y < 0 ? 0 : (y >= ph ? ph - 1 : y);
                            float v = pad[(size_t)y * pw + x];
                            int noact = (px == 3 && py == 3 && i == 2 && (p->quirks & FYO_Q7_MAXPOOL3_COL));
                            if (!noact) v = act1(v, &p->act);
                            if (first || v > r) r = v;
                            first = 0;
                        }
                    }
                } else {
                    int off = (px == 3 && py == 3 && !p->global) ? -1 : 0;
                    float sum = 0.f;
                    for (int j = 0; j < py; j++) {
                        for (int i = 0; i < px; i++) {
                            int x = bx + off + i, y = by + off + j;
                            x = x < 0 ? 0 : (x >= pw ? pw - 1 : x);
                            y = y < 0 ? 0 : (y >= ph ? ph - 1 : y);
                            sum += act1(pad[(size_t)y * pw + x], &p->act);
                        }
                    }
                    r = sum / (float)(px * py);
                }
                out_chw[((size_t)c * Ho + yo) * Wo + xo] = store(r, p->prec);
            }
        }
        free(pad);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* batchnorm, sigmoid                                                                          */
/* ------------------------------------------------------------------------------------------- */

/* shallow: gpu/batchnormlayer.cpp:71-92 + shaders/batchnorm.frag:60-68  (bias + x*scale, NO activation)
 * deep   : gpu/deep/deepbatchnormlayer.cpp:78-88 + shaders/deep/deepbatchnorm.frag:57-58 (act(x)*scale+bias)
 * data   : scale[C] then bias[C] (base/batchnorminterface.h:33-48) */
int fyo_batchnorm(const float *in_chw, int C, int H, int W, const float *scaleBias, int deep,
                  const fyo_act *act, int prec, float *out_chw) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = (deep && act) ? act : &none;
    for (int c = 0; c < C; c++) {
        float s = scaleBias[c], b = scaleBias[C + c];
        for (size_t i = 0; i < (size_t)H * W; i++) {
            float x = act1(in_chw[(size_t)c * H * W + i], a);
            float r = deep ? x * s + b : b + x * s;
            out_chw[(size_t)c * H * W + i] = store(r, prec);
        }
    }
    return 0;
}

/* gpu/sigmoidlayer.cpp:77-92 + shaders/sigmoid.frag:10-13: 1/(1+exp(-act(x))) */
int fyo_sigmoid(const float *in_chw, size_t n, const fyo_act *act, int prec, float *out_chw) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    for (size_t i = 0; i < n; i++) {
        float x = act1(in_chw[i], a);
        out_chw[i] = store(1.0f / (1.0f + expf(-x)), prec);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* scaling, arithmetic, concatenation, swizzle                               */
/* ------------------------------------------------------------------------------------------- */

/* One texel lane of channel c's texture at ABSOLUTE texture coordinates (X, Y), with the sampler's clamp to the
 * texture edge (base/buffermanager.cpp:657-670).  Shallow: plane c/4 of (W+2P)x(H+2P), interior at (P,P).
 * Deep: tile (c/4) of the tiled texture (gpu/deep/deeptiler.cpp:63-95); a coordinate that leaves the tile lands in the
 * padding gap (zero) or in the NEIGHBOURING tile, exactly as a GL sampler would see it. */
static float tex_lane(const float *chw, int C, int H, int W, int P, int deep, int c, int X, int Y) {
    if (!deep) {
        int tw = W + 2 * P, th = H + 2 * P;
        X = X < 0 ? 0 : (X >= tw ? tw - 1 : X);
        Y = Y < 0 ? 0 : (Y >= th ? th - 1 : Y);
        int x = X - P, y = Y - P;
        if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
        return chw[((size_t)c * H + y) * W + x];
    }
    int tx, ty, tw, th;
    fyo_deep_tiling(C, &tx, &ty);
    fyo_deep_texture_size(C, W, H, P, &tw, &th);
    X = X < 0 ? 0 : (X >= tw ? tw - 1 : X);
    Y = Y < 0 ? 0 : (Y >= th ? th - 1 : Y);
    if (X < P || Y < P) return 0.f;
    int col = (X - P) / (W + P), row = (Y - P) / (H + P);
    int x = (X - P) - col * (W + P), y = (Y - P) - row * (H + P);
    if (x >= W || y >= H || col >= tx || row >= ty) return 0.f;
    int ch = (row * tx + col) * 4 + (c & 3);
    if (ch >= C || (row * tx + col) >= (C + 3) / 4) return 0.f;
    return chw[((size_t)ch * H + y) * W + x];
}

static int floor_div(int num, int den) { return num >= 0 ? num / den : -((-num + den - 1) / den); }

void fyo_scale_outdims(int W, int H, int upx, int upy, int dnx, int dny, int *Wo, int *Ho) {
    /* gpu/scalelayer.cpp:44-47 */
    *Wo = (int)(((float)upx / (float)dnx) * (float)W);
    *Ho = (int)(((float)upy / (float)dny) * (float)H);
}

/* ScaleLayer / DeepScaleLayer: gpu/scalelayer.cpp:40-60, gpu/deep/deepscalelayer.cpp:30-75, shaders/scaling.frag
 * (activate(texture(...))), quad geometry gpu/functionlayer.cpp:194-204: output texel o samples texel-space
 * coordinate P + (o+0.5)*W/Wo of its channel's texture; NEAREST = the texel containing it, LINEAR = GL_LINEAR. */
int fyo_scale(const float *in_chw, int C, int H, int W, int in_pad, int deep, int upx, int upy, int dnx, int dny,
              int linear, const fyo_act *act, int prec, float *out_chw) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    int Wo, Ho;
    if (upx < 1 || upy < 1 || dnx < 1 || dny < 1) return -1;
    fyo_scale_outdims(W, H, upx, upy, dnx, dny, &Wo, &Ho);
    if (Wo < 1 || Ho < 1) return -1;
    if (deep && (W == 1 || H == 1)) linear = 0; /* deepscalelayer.cpp:34 */
    int tx = 1, ty = 1;
    if (deep) fyo_deep_tiling(C, &tx, &ty);
    for (int c = 0; c < C; c++) {
        int t = c / 4;
        int ox = deep ? in_pad + (t % tx) * (W + in_pad) : in_pad;
        int oy = deep ? in_pad + (t / tx) * (H + in_pad) : in_pad;
        for (int yo = 0; yo < Ho; yo++)
            for (int xo = 0; xo < Wo; xo++) {
                float v;
                if (!linear) {
                    int sx = ((2 * xo + 1) * W) / (2 * Wo), sy = ((2 * yo + 1) * H) / (2 * Ho);
                    v = tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + sx, oy + sy);
                } else {
                    int nx = (2 * xo + 1) * W - Wo, ny = (2 * yo + 1) * H - Ho;
                    int ix = floor_div(nx, 2 * Wo), iy = floor_div(ny, 2 * Ho);
                    float fx = (float)(nx - ix * 2 * Wo) / (float)(2 * Wo), fy = (float)(ny - iy * 2 * Ho) / (float)(2 * Ho);
                    float v00 = tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + ix, oy + iy);
                    float v10 = tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + ix + 1, oy + iy);
                    float v01 = tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + ix, oy + iy + 1);
                    float v11 = tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + ix + 1, oy + iy + 1);
                    float top = v00 + (v10 - v00) * fx, bot = v01 + (v11 - v01) * fx;
                    v = top + (bot - top) * fy;
                }
                out_chw[((size_t)c * Ho + yo) * Wo + xo] = store(act1(v, a), prec);
            }
    }
    return 0;
}

/* Depthwise 3x3 convolution, channel multiplier 1.
 * shallow: gpu/vanilla/convlayer_dw_3x3_vanilla.cpp:22-75 + shaders/vanilla/conv_dw_3x3.frag (taps +-1 texel via
 *          textureOffset, accu += act(pix)*coeff in (ky,kx) order, then *= bnscale, += bias), weights
 *          gpu/convweightarray_dw_KxKxNxM.cpp:120-150 (W[c][ky][kx]), bias fold :84-95; quirk bit 8: the layer passes the
 *          block start as batch-norm offset (convlayer_dw_3x3_vanilla.cpp:66), so scale = blob[0..C), beta = blob[C..2C).
 * deep   : gpu/deep/deepdwconvlayer3x3.cpp + shaders/deep/deepconv_dw3x3_tiled.frag (taps +-dilation, row sums added up,
 *          applyBN then += bias), weights fp16-truncated and bias / scale through an RGBA16F texture when prec != FP32
 *          (gpu/deep/deepdwconvlayerbase.cpp:40-75,255-262).
 * data: bias[C], W[C][3][3], (bnScale[C], bnBias[C]). */
int fyo_dwconv3x3_ex(const float *in_chw, int C, int H, int W, int in_pad, int deep, int ds, int dil, int post_bn, int quirks,
                     const float *wb, const fyo_act *act, int prec, int mult, const float *res_chw, int res_flags, float *out_chw) {
    /* Channel multiplier (deep layers only; the shallow layer throws for multipliers != 1, convlayer_dw_3x3_vanilla.cpp:49-50):
     * output tile t + m * tiles(C) is input tile t filtered with multiplier m (deepdwconvlayerbase.cpp:288-297), i.e. output
     * channel m * C + c reads input channel c with W[c][ky][kx][m] (weights [C][3][3][mult], :234-246); bias / batch-norm data
     * are indexed by output channel (:96-125).  Residual: added after bias / batch-norm, optionally through ReLU
     * (shaders/vanilla/conv_dw_3x3.frag:133-140; shaders/deep/residual.inc) and, for deep layers, scaled by the batch-norm
     * scale (BATCHNORM_ON_RESIDUAL, residual.inc:9-11).  res_flags: bit 0 = ReLU on the residual, bit 1 = batch-norm on it. */
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    if (ds < 1 || dil < 1 || (!deep && dil != 1) || mult < 1) return -1;
    if (mult > 1 && (!deep || (C & 3))) return -1;
    int Wo = W / ds, Ho = H / ds;
    if (Wo < 1 || Ho < 1) return -1;
    int tx = 1, ty = 1;
    if (deep) fyo_deep_tiling(C, &tx, &ty);
    const int Co = C * mult;
    const float *bn = (!deep && (quirks & 8)) ? wb : wb + Co + (size_t)C * 9 * mult;
    int reduced = deep && prec != FYO_FP32;
    for (int o = 0; o < Co; o++) {
        const int m = o / C, c = o - m * C;
        int t = c / 4;
        int ox = deep ? in_pad + (t % tx) * (W + in_pad) : in_pad;
        int oy = deep ? in_pad + (t / tx) * (H + in_pad) : in_pad;
        float b = wb[o], s = 1.f;
        if (post_bn) {
            s = bn[o];
            b = b * s + bn[Co + o];
        }
        if (reduced) {
            b = h_rn(b);
            s = h_rn(s);
        }
        for (int yo = 0; yo < Ho; yo++)
            for (int xo = 0; xo < Wo; xo++) {
                float acc = 0.f;
                for (int ky = 0; ky < 3; ky++) {
                    float row = 0.f;
                    for (int kx = 0; kx < 3; kx++) {
                        float w = wb[Co + ((size_t)c * 9 + ky * 3 + kx) * mult + m];
                        if (reduced) w = fyo_half_trunc(w);
                        float v = act1(tex_lane(in_chw, C, H, W, in_pad, deep, c, ox + ds * xo + (kx - 1) * dil, oy + ds * yo + (ky - 1) * dil), a);
                        if (deep) row += v * w;
                        else acc += v * w;
                    }
                    if (deep) acc += row;
                }
                float r = acc * s + b;
                if (res_chw) {
                    float q = res_chw[((size_t)o * Ho + yo) * Wo + xo];
                    if (res_flags & 1) q = q > 0.f ? q : 0.f;
                    if ((res_flags & 2) && deep) q *= s;
                    r += q;
                }
                out_chw[((size_t)o * Ho + yo) * Wo + xo] = store(r, prec);
            }
    }
    return 0;
}

int fyo_dwconv3x3(const float *in_chw, int C, int H, int W, int in_pad, int deep, int ds, int dil, int post_bn, int quirks,
                  const float *wb, const fyo_act *act, int prec, float *out_chw) {
    return fyo_dwconv3x3_ex(in_chw, C, H, W, in_pad, deep, ds, dil, post_bn, quirks, wb, act, prec, 1, NULL, 0, out_chw);
}

/* Transpose convolution, stride 2, shallow: gpu/vanilla/transconvlayerbase_vanilla.cpp (viewport 2W x 2H :60-62; quad maps
 * the output interior onto the input interior :365-372, so output texel o samples texel coordinate P + (o+0.5)/2; texStep =
 * half a texel :206-209; strata by output parity via the stencil :499-506, stratum s+1 = (x&1) + 2*(y&1) + 1 :233),
 * shaders/vanilla/convtrans3x3_stride2.frag / convtrans2x2_stride2.frag (taps per stratum), weights
 * gpu/transconvweightarray3x3xNxM.cpp:261-420 (stratum 1: tap 4; 2: taps 3,5; 3: taps 1,7; 4: taps 0,2,6,8 of
 * W[Co][ky][kx][Ci]) and transconvweightarray2x2xNxM.cpp:277-414 (stratum s: tap s-1), bias fold :168-179.
 * quirks & 16: the 2x2 shader samples tc - hstep in stratum 2 (column i) but tc + vstep / tc + step in strata 3 / 4 (row
 * j+1, column i+1 for odd/odd); without the bit every stratum reads input (i, j). */
int fyo_transconv(const float *in_chw, int Ci, int H, int W, int in_pad, int Co, int K, int post_bn, int quirks, const float *wb,
                  const fyo_act *act, int prec, float *out_chw) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    if (K != 2 && K != 3) return -1;
    const float *wsrc = wb + Co, *bn = wsrc + (size_t)Co * K * K * Ci;
    int Wo = 2 * W, Ho = 2 * H;
    for (int o = 0; o < Co; o++) {
        float b = wb[o], s = 1.f;
        if (post_bn) {
            s = bn[o];
            b = b * s + bn[Co + o];
        }
        for (int yo = 0; yo < Ho; yo++)
            for (int xo = 0; xo < Wo; xo++) {
                int i = xo / 2, j = yo / 2, ox = xo & 1, oy = yo & 1;
                int kxs[2], dxs[2], nx, kys[2], dys[2], ny;
                if (K == 3) {
                    if (ox) { nx = 2; kxs[0] = 0; dxs[0] = 0; kxs[1] = 2; dxs[1] = 1; } else { nx = 1; kxs[0] = 1; dxs[0] = 0; }
                    if (oy) { ny = 2; kys[0] = 0; dys[0] = 0; kys[1] = 2; dys[1] = 1; } else { ny = 1; kys[0] = 1; dys[0] = 0; }
                } else {
                    nx = ny = 1;
                    kxs[0] = ox;
                    kys[0] = oy;
                    dxs[0] = ((quirks & 16) && ox && oy) ? 1 : 0;
                    dys[0] = ((quirks & 16) && oy) ? 1 : 0;
                }
                float acc = 0.f;
                for (int ty = 0; ty < ny; ty++)
                    for (int tx = 0; tx < nx; tx++)
                        for (int c = 0; c < Ci; c++) {
                            float v = act1(tex_lane(in_chw, Ci, H, W, in_pad, 0, c, in_pad + i + dxs[tx], in_pad + j + dys[ty]), a);
                            acc += v * wsrc[(((size_t)o * K + kys[ty]) * K + kxs[tx]) * Ci + c];
                        }
                out_chw[((size_t)o * Ho + yo) * Wo + xo] = store(acc * s + b, prec);
            }
    }
    return 0;
}

/* Transpose convolution, stride 2, deep-tiled: gpu/deep/deeptransconvlayerbase.cpp (four passes selected by a stencil of the
 * output parity, pass = (x & 1) + 2 (y & 1), :376-383 and deeptransconvlayer3x3.cpp:68-71; weights W[Co][fy][fx][Ci] as 4x4
 * blocks per (output tile, kernel row, input tile, fx), :150-170; fp16-truncated with fp16 storage :177-186; bias RGBA16F
 * :228-232), shaders/deep/deeptransconv3x3_stride2.{vert,frag} and deeptransconv2x2_stride2.{vert,frag}.  With texStep = half
 * an input texel (deeptransconvlayer3x3.cpp:124-128) output texel o = 2 i + a samples
 *   3x3: a = 0: tap 0 on input i and tap 2 on input i - 1 (vert pass 0 / 2 columns intile+0,+4; frag tc, tc - 2 step);
 *        a = 1: tap 1 on input i (columns intile+2; frag tc - step)                       -- per axis;
 *   2x2: tap a on input i (vert pass & 1 -> column intile + TSTEP; frag tc - step * (pass & 1)),
 * i.e. out = the full convolution of the zero-stuffed input with the kernel, out[o] = sum_k W[k] u[o - k].  Reads outside the
 * tile's image are ZERO whatever the padding (clampedTexture, deeptransconv3x3_stride2.frag:21-25), the prefix activation is
 * applied to the (masked) texel (computeconv.inc).  POST_BATCHNORM: the reference writes the scales past the end of its bias
 * array and never uploads them (deeptransconvlayerbase.cpp:219-232: the buffer has 1 + tiles texels, the scale index starts at
 * bs / 2); this restatement uses the documented meaning, out = acc * s + (b * s + beta) -- unpinned for that flag. */
int fyo_transconv_deep(const float *in_chw, int Ci, int H, int W, int in_pad, int Co, int K, int post_bn, const float *wb,
                       const fyo_act *act, int prec, float *out_chw) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    if (K != 2 && K != 3) return -1;
    const float *wsrc = wb + Co, *bn = wsrc + (size_t)Co * K * K * Ci;
    const int Wo = 2 * W, Ho = 2 * H, reduced = prec != FYO_FP32;
    for (int o = 0; o < Co; o++) {
        float b = wb[o], s = 1.f;
        if (post_bn) {
            s = bn[o];
            b = b * s + bn[Co + o];
        }
        if (reduced) {
            b = h_rn(b);
            s = h_rn(s);
        }
        for (int yo = 0; yo < Ho; yo++)
            for (int xo = 0; xo < Wo; xo++) {
                const int i = xo / 2, j = yo / 2, ox = xo & 1, oy = yo & 1;
                int kxs[2], dxs[2], nx, kys[2], dys[2], ny;
                if (K == 3) {
                    if (ox) { nx = 1; kxs[0] = 1; dxs[0] = 0; } else { nx = 2; kxs[0] = 0; dxs[0] = 0; kxs[1] = 2; dxs[1] = -1; }
                    if (oy) { ny = 1; kys[0] = 1; dys[0] = 0; } else { ny = 2; kys[0] = 0; dys[0] = 0; kys[1] = 2; dys[1] = -1; }
                } else {
                    nx = ny = 1;
                    kxs[0] = ox;
                    kys[0] = oy;
                    dxs[0] = dys[0] = 0;
                }
                float acc = 0.f;
                for (int ty = 0; ty < ny; ty++)
                    for (int tx = 0; tx < nx; tx++) {
                        const int x = i + dxs[tx], y = j + dys[ty];
                        for (int c = 0; c < Ci; c++) {
                            float v = (x < 0 || y < 0) ? 0.f : in_chw[((size_t)c * H + y) * W + x];
                            float w = wsrc[(((size_t)o * K + kys[ty]) * K + kxs[tx]) * Ci + c];
                            if (reduced) w = fyo_half_trunc(w);
                            acc += act1(v, a) * w;
                        }
                    }
                out_chw[((size_t)o * Ho + yo) * Wo + xo] = store(acc * s + b, prec);
            }
    }
    (void)in_pad;
    return 0;
}

/* AddSubLayer: gpu/addsublayer.cpp + shaders/add.frag:84-135 (fetch = activate(texture)); SingletonArithmeticLayer:
 * gpu/singleton_arithlayer.cpp + shaders/singleton_arith.frag (activate(texture) op operand).
 * op: 0 add, 1 sub, 2 mul, 3 div; in2 == NULL: scalar operand. */
int fyo_arith(const float *in1, const float *in2, size_t n, int op, float operand, const fyo_act *act, int prec, float *out) {
    fyo_act none = {FYO_ACT_NONE, 0, 0, 0};
    const fyo_act *a = act ? act : &none;
    if (op < 0 || op > 3 || (in2 && op > 1)) return -1;
    for (size_t i = 0; i < n; i++) {
        float p = act1(in1[i], a), q = in2 ? act1(in2[i], a) : operand, r;
        switch (op) {
        case 0: r = p + q; break;
        case 1: r = p - q; break;
        case 2: r = p * q; break;
        default: r = p / q; break;
        }
        out[i] = store(r, prec);
    }
    return 0;
}

/* RGB2BGRLayer: shaders/rgb2bgr.frag (val.bgra per texel): lanes 0 and 2 of every 4-channel group trade places; a
 * lane that does not exist in the tensor reads as 0 and a value moved to a non-existing lane is dropped. */
int fyo_rgb2bgr(const float *in_chw, int C, int H, int W, int prec, float *out_chw) {
    size_t hw = (size_t)H * W;
    for (int c = 0; c < C; c++) {
        int l = c & 3, src = (l == 0) ? c + 2 : (l == 2 ? c - 2 : c);
        for (size_t i = 0; i < hw; i++) out_chw[(size_t)c * hw + i] = src < C ? store(in_chw[(size_t)src * hw + i], prec) : 0.f;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* host <-> texture conversions                                                                */
/* ------------------------------------------------------------------------------------------- */

/* gpu/uploadlayer.cpp:360-380: host [H][W][C] float32 (GPU_SHALLOW order, C<=4) becomes a C-channel
 * float32 texture verbatim (no rounding: the upload texture is RGB32F). -> CHW */
void fyo_upload_hwc_to_chw(const float *hwc, int C, int H, int W, float *chw) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < C; c++) chw[((size_t)c * H + y) * W + x] = hwc[((size_t)y * W + x) * C + c];
}

/* gpu/downloadlayer.cpp:257-283: each 4-channel texture is read back as an RGBA float block:
 * host = [planes][H][W][4]; lanes beyond C hold whatever the texture holds (0 for conv outputs,
 * sigmoid(0)=0.5 after a sigmoid layer -- SURVEY A.5). 'fill' is that lane value. */
void fyo_download_shallow(const float *chw, int C, int H, int W, float fill, float *host) {
    int np = (C + 3) / 4;
    for (int p = 0; p < np; p++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int l = 0; l < 4; l++) {
                    int c = p * 4 + l;
                    host[(((size_t)p * H + y) * W + x) * 4 + l] = c < C ? chw[((size_t)c * H + y) * W + x] : fill;
                }
}

/* Thread control for the timed CPU baseline (bench.py): launchers such as torchrun export OMP_NUM_THREADS=1. */
#ifdef _OPENMP
#include <omp.h>
int fyo_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int fyo_set_threads(int n) {
    (void)n;
    return 1;
}
#endif
