/* fyn_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY, see fyn_oracle.c). */
#ifndef FYN_ORACLE_H
#define FYN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layer flag bits, numerically identical to fyusenet/base/layerflags.h:33-53 */
enum {
    FYO_RESIDUAL_INPUT = 1,
    FYO_RELU_ON_RESIDUAL = 2,
    FYO_BATCHNORM_ON_RESIDUAL = 4,
    FYO_POST_BATCHNORM = 8,
    FYO_DEEP = 16,
    FYO_PRE_RELU = 64,
    FYO_PRE_CLIP = 128
};

enum { FYO_ACT_NONE = 0, FYO_ACT_RELU = 1, FYO_ACT_LEAKY = 2, FYO_ACT_CLIP = 3 };
enum { FYO_FP32 = 0, FYO_FP16_STORE = 1, FYO_FP16_BLEND = 2 };

/* reference quirks, all ON = bit-faithful to the shaders */
enum {
    FYO_Q1_FRAC3_ASYM = 1,      /* fraconv3x3.frag:14-19 horizontal taps -2s,-s,0 */
    FYO_Q2_FRAC_ACT_FIRST = 2,  /* fractional.inc:11-12 vs :69-70 activation on first tap only */
    FYO_Q7_MAXPOOL3_COL = 4,    /* deepmaxpool.frag: 3rd column of a 3x3 max-pool not activated */
    FYO_Q8_DW_BN_OFFSET = 8,    /* convlayer_dw_3x3_vanilla.cpp:66: shallow depthwise conv reads its BN data at the block start */
    FYO_Q16_TRANS2X2_NEXT = 16, /* convtrans2x2_stride2.frag: strata 3 / 4 sample the NEXT input row / texel */
    FYO_QUIRKS_REFERENCE = 31
};

typedef struct {
    int type; /* FYO_ACT_* */
    float leak, lo, hi;
} fyo_act;

typedef struct {
    int width, height;        /* input net size (without padding) */
    int inChannels, outChannels;
    int kernel, downsample, dilation;
    int inPadding, outPadding;
    unsigned flags;           /* FYO_* layer flags (activation comes from .act) */
    fyo_act act;              /* prefix activation */
    float sourceStep;         /* fractional convs only */
    int fractional;           /* LayerType::FRACCONVOLUTION2D */
    int deep;                 /* deep-tiled variant */
    int quirks;               /* FYO_Q* */
    int prec;                 /* FYO_FP32 / FYO_FP16_STORE / FYO_FP16_BLEND */
} fyo_conv;

typedef struct {
    int width, height, channels;
    int poolX, poolY, downsample, inPadding;
    int isMax, global;
    fyo_act act;
    int quirks, prec;
} fyo_pool;

float fyo_half_round(float x);
float fyo_half_trunc(float x);
uint16_t fyo_half_trunc_bits(float x);

void fyo_deep_tiling(int channels, int *tx, int *ty);
void fyo_deep_texture_size(int channels, int w, int h, int pad, int *tw, int *th);
void fyo_pack_deep(const float *chw, int C, int H, int W, int pad, float *texels);
void fyo_unpack_deep(const float *texels, int C, int H, int W, int pad, float *chw);
void fyo_pack_shallow(const float *chw, int C, int H, int W, int pad, float *planes);
void fyo_unpack_shallow(const float *planes, int C, int H, int W, int pad, float *chw);

void fyo_conv2d_outdims(const fyo_conv *p, int *Wo, int *Ho);
/* wb = bias[Co], W[Co][K][K][Ci], then (POST_BATCHNORM) bnScale[Co], bnBias[Co]; res_chw may be NULL */
int fyo_conv2d(const fyo_conv *p, const float *in_chw, const float *wb, const float *res_chw, float *out_chw);
int fyo_pool2d(const fyo_pool *p, const float *in_chw, float *out_chw);
int fyo_batchnorm(const float *in_chw, int C, int H, int W, const float *scaleBias, int deep,
                  const fyo_act *act, int prec, float *out_chw);
int fyo_sigmoid(const float *in_chw, size_t n, const fyo_act *act, int prec, float *out_chw);
void fyo_scale_outdims(int W, int H, int upx, int upy, int dnx, int dny, int *Wo, int *Ho);
int fyo_scale(const float *in_chw, int C, int H, int W, int in_pad, int deep, int upx, int upy, int dnx, int dny,
              int linear, const fyo_act *act, int prec, float *out_chw);
int fyo_arith(const float *in1, const float *in2, size_t n, int op, float operand, const fyo_act *act, int prec, float *out);
int fyo_dwconv3x3_ex(const float *in_chw, int C, int H, int W, int in_pad, int deep, int ds, int dil, int post_bn, int quirks,
                     const float *wb, const fyo_act *act, int prec, int mult, const float *res_chw, int res_flags, float *out_chw);
int fyo_dwconv3x3(const float *in_chw, int C, int H, int W, int in_pad, int deep, int ds, int dil, int post_bn, int quirks,
                  const float *wb, const fyo_act *act, int prec, float *out_chw);
int fyo_transconv_deep(const float *in_chw, int Ci, int H, int W, int in_pad, int Co, int K, int post_bn, const float *wb,
                       const fyo_act *act, int prec, float *out_chw);
int fyo_transconv(const float *in_chw, int Ci, int H, int W, int in_pad, int Co, int K, int post_bn, int quirks, const float *wb,
                  const fyo_act *act, int prec, float *out_chw);
int fyo_rgb2bgr(const float *in_chw, int C, int H, int W, int prec, float *out_chw);
void fyo_upload_hwc_to_chw(const float *hwc, int C, int H, int W, float *chw);
void fyo_download_shallow(const float *chw, int C, int H, int W, float fill, float *host);
/* sets (n > 0) and returns the number of OpenMP threads the oracle uses */
int fyo_set_threads(int n);

#ifdef __cplusplus
}
#endif

#endif
