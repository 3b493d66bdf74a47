/* ---------------------------------------------------------------------------------------------
 * fyusenet_b200.h -- C ABI of the B200-native GPU-layer backend for FyuseNet.
 *
 * This is the drop-in boundary.  The reference has no C ABI on this
 * path: its seam is the C++ virtual interface
 *     LayerFactoryBackend::createLayer(LayerType, LayerBuilder*, int)      fyusenet/base/layerfactory.h:149-176
 * whose products implement LayerBase (fyusenet/base/layerbase.h:109-203), the GPULayerBase
 * texture-slot API (fyusenet/gpu/gpulayerbase.h:127-142), ConvLayerInterface::loadWeightsAndBiases
 * (fyusenet/base/convlayerinterface.h:58) and BatchNormInterface::loadScaleAndBias
 * (fyusenet/base/batchnorminterface.h:48).  The C++ layer classes of the new backend
 * (fyusenet_b200/host/) implement those interfaces and call ONLY the functions below; every
 * entry point cites the reference mechanism it replaces.
 *
 * Conventions
 *   - extern "C", POD structs, opaque handles, plain pointers and sizes; no C++/torch types.
 *   - every function returns 0 (FYN_OK) or a negative fyn_status; fyn_last_error() returns a
 *     thread-local message for the last failing call (the C++ wrapper turns it into FynException,
 *     mirroring THROW_EXCEPTION_ARGS, fyusenet/common/fynexception.h:24-25,75-107).
 *   - a context is bound to one CUDA device; calls on one context must come from one thread at
 *     a time (the reference's "GL context must be current" rule, base/layerbase.h:106-108).
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = legacy default stream).  Ops
 *     only enqueue work; nothing synchronises except the functions documented as blocking.
 *   - there is NO CPU fallback: without a CUDA device fyn_cuda_init fails.
 * ------------------------------------------------------------------------------------------- */
#ifndef FYUSENET_B200_H
#define FYUSENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FYN_ABI_VERSION 1

typedef enum {
    FYN_OK = 0,
    FYN_ERR_INVALID = -1,   /* bad argument / shape mismatch */
    FYN_ERR_CUDA = -2,      /* CUDA runtime error (message in fyn_last_error) */
    FYN_ERR_NOMEM = -3,
    FYN_ERR_UNSUPPORTED = -4,
    FYN_ERR_NODEVICE = -5
} fyn_status;

/* layer flag bits, numerically identical to fyusenet/base/layerflags.h:33-53 */
enum {
    FYN_FLAG_RESIDUAL_INPUT = 1,
    FYN_FLAG_RELU_ON_RESIDUAL = 2,
    FYN_FLAG_BATCHNORM_ON_RESIDUAL = 4,
    FYN_FLAG_POST_BATCHNORM = 8,
    FYN_FLAG_DEEP = 16,
    FYN_FLAG_PRE_RELU = 64,
    FYN_FLAG_PRE_CLIP = 128
};

/* reference shader quirks that the kernels reproduce by default */
enum {
    FYN_QUIRK_FRAC3_ASYM = 1,      /* fraconv3x3.frag:14-19: horizontal taps at -2s,-s,0 */
    FYN_QUIRK_FRAC_ACT_FIRST = 2,  /* fractional.inc:11-12 vs :69-70: activation on the first tap only */
    FYN_QUIRK_MAXPOOL3_COL = 4,    /* deepmaxpool.frag: 3rd column of a 3x3 max-pool bypasses activate() */
    FYN_QUIRK_DW_BN_OFFSET = 8,    /* convlayer_dw_3x3_vanilla.cpp:66: the shallow depthwise layer reads its post-BN data at the block start */
    FYN_QUIRK_TRANS2X2_NEXT = 16,  /* convtrans2x2_stride2.frag: odd output rows read input row j+1, odd/odd reads (i+1, j+1) */
    FYN_QUIRKS_REFERENCE = 31
};

typedef enum { FYN_ORDER_SHALLOW = 0, FYN_ORDER_DEEP = 1 } fyn_order;
typedef enum { FYN_F16 = 0, FYN_F32 = 1 } fyn_dtype;

typedef struct fyn_ctx fyn_ctx;
typedef struct fyn_tensor fyn_tensor;
typedef struct fyn_op fyn_op;

/* ----------------------------------------------------------------------------------------- */
/* context  (replaces GfxContextManager / GfxContextLink, fyusenet/gpu/gfxcontextmanager.h)   */
/* ----------------------------------------------------------------------------------------- */

typedef struct {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t total_mem;
    size_t smem_per_block_optin;
    char name[64];
} fyn_device_info;

int fyn_abi_version(void);
const char *fyn_last_error(void);
int fyn_device_count(int *count);
int fyn_cuda_init(int device, fyn_ctx **ctx);
int fyn_cuda_shutdown(fyn_ctx *ctx);
int fyn_get_device_info(fyn_ctx *ctx, fyn_device_info *info);
/* number of kernels this library has launched on the context since the last reset (bench "gpu_launches") */
int fyn_launch_count(fyn_ctx *ctx, uint64_t *count, int reset);

/* streams / events (replace GLsync fences, fyusenet/base/engine.cpp:779-780; timings :208-233) */
int fyn_stream_create(fyn_ctx *ctx, void **stream);
int fyn_stream_destroy(fyn_ctx *ctx, void *stream);
int fyn_stream_sync(fyn_ctx *ctx, void *stream);              /* blocking */
int fyn_event_create(fyn_ctx *ctx, void **event);
int fyn_event_destroy(fyn_ctx *ctx, void *event);
int fyn_event_record(fyn_ctx *ctx, void *event, void *stream);
int fyn_event_sync(fyn_ctx *ctx, void *event);                /* blocking */
int fyn_event_elapsed_ms(fyn_ctx *ctx, void *start, void *stop, float *ms);
int fyn_stream_wait_event(fyn_ctx *ctx, void *stream, void *event);

/* host function executed in stream order (replaces the GLsync waiter threads of the asynchronous download,
 * fyusenet/gpu/downloadlayer.cpp:307-323): fn(user) runs on a driver thread once all prior work of `stream` is done */
typedef void (*fyn_host_fn)(void *user);
int fyn_stream_add_callback(fyn_ctx *ctx, void *stream, fyn_host_fn fn, void *user);

/* CUDA-graph capture / replay of a fixed launch sequence (static networks at small batch are launch-latency bound: ResNet-50
 * batch 1 is 59 kernels; the reference pays the same price per GL draw call, README.md:293-295).  Everything the library
 * enqueues on `stream` between begin and end becomes one graph; programmatic dependent launches are kept as such. */
int fyn_graph_begin_capture(fyn_ctx *ctx, void *stream);
int fyn_graph_end_capture(fyn_ctx *ctx, void *stream, void **graph_exec);
int fyn_graph_launch(fyn_ctx *ctx, void *graph_exec, void *stream);
int fyn_graph_destroy(fyn_ctx *ctx, void *graph_exec);

/* pinned host memory (replaces PBOPool / ManagedPBO, fyusenet/gl/pbopool.cpp) */
int fyn_host_alloc(fyn_ctx *ctx, size_t bytes, void **ptr);
int fyn_host_free(fyn_ctx *ctx, void *ptr);
/* raw device staging memory + asynchronous copies for the pipelined (asynchronous) upload / download path */
int fyn_device_alloc(fyn_ctx *ctx, size_t bytes, void **ptr);
int fyn_device_free(fyn_ctx *ctx, void *ptr);
int fyn_memcpy_async(fyn_ctx *ctx, void *dst, const void *src, size_t bytes, int device_to_host, void *stream);

/* ----------------------------------------------------------------------------------------- */
/* device tensors  (replace BufferManager::createTexture, fyusenet/base/buffermanager.cpp:650-708) */
/* ----------------------------------------------------------------------------------------- */

/*
 * Layout contract (FyuseNet's 4-channel-packed, spatially padded layouts; PIXEL_PACKING = 4,
 * fyusenet/base/layerflags.h:191):
 *   SHALLOW: [batch][ceil(C/4)][H+2P][W+2P][packing]   channel c -> plane c/4, lane c%4
 *            (unit_tests/layertestbase.cpp:281-317).  packing is 4, or 1..3 for a single-plane
 *            texture with fewer channels (upload textures, convlayerbase_vanilla.cpp:205-210).
 *   DEEP   : [batch][TH][TW][4] with T = ceil(C/4) tiles on a tx x ty grid
 *            (cpu/cpubuffershape.cpp:430-447), tile i at pixel (P+(i%tx)(W+P), P+(i/tx)(H+P)),
 *            TW = tx(W+P)+P, TH = ty(H+P)+P (gpu/deep/deeptiler.cpp:63-95).
 * Padding texels and unused lanes are zero: they are cleared at creation and no kernel writes them.
 * dtype F16 is the reference default (RGBA16F), F32 = HIGH_PRECISION (gpu/gpulayerbase.h:100-110).
 * batch is new (the reference is batch-1, README.md:72); batch=1 reproduces it exactly.
 */
typedef struct {
    int width, height;  /* net size, without padding */
    int channels;
    int padding;
    int order;          /* fyn_order */
    int dtype;          /* fyn_dtype */
    int batch;          /* >= 1 */
    int packing;        /* 0 or 4 = RGBA; 1..3 only for shallow single-plane tensors */
} fyn_tensor_desc;

typedef struct {
    int tex_width, tex_height; /* texels per plane (shallow) or of the tiled texture (deep) */
    int planes;                /* shallow: ceil(C/4); deep: 1 */
    int tiles_x, tiles_y;      /* deep tiling (1,1 for shallow) */
    int packing;               /* resolved packing */
    size_t elem_size;          /* 2 or 4 */
    size_t plane_elems;        /* elements per plane / tiled texture */
    size_t image_elems;        /* elements per batch image */
    size_t bytes;              /* whole tensor */
} fyn_tensor_geom;

int fyn_tensor_geometry(const fyn_tensor_desc *desc, fyn_tensor_geom *geom); /* host only, no device */
int fyn_tensor_create(fyn_ctx *ctx, const fyn_tensor_desc *desc, fyn_tensor **tensor);
/* wrap caller-owned device memory of at least geom.bytes (e.g. a torch allocation); never freed here */
int fyn_tensor_wrap(fyn_ctx *ctx, const fyn_tensor_desc *desc, void *device_ptr, fyn_tensor **tensor);
int fyn_tensor_destroy(fyn_tensor *tensor);
int fyn_tensor_clear(fyn_tensor *tensor, void *stream);
int fyn_tensor_get_desc(const fyn_tensor *tensor, fyn_tensor_desc *desc, fyn_tensor_geom *geom);
void *fyn_tensor_device_ptr(const fyn_tensor *tensor);

/* ----------------------------------------------------------------------------------------- */
/* host <-> device I/O                                                                         */
/* ----------------------------------------------------------------------------------------- */

/* UploadLayer::syncUpload (fyusenet/gpu/uploadlayer.cpp:360-380): host float32 [batch][H][W][C]
 * (CPUBuffer GPU_SHALLOW order, C = tensor channels <= 4) -> single-plane shallow tensor.  When the
 * tensor is F32 with packing == C and padding 0 (the reference's RGB32F upload texture) this is one
 * cudaMemcpyAsync; otherwise the data goes through a device staging buffer and a convert kernel.
 * Asynchronous w.r.t. the host when `host` is pinned. */
int fyn_upload_f32_async(fyn_tensor *tensor, const float *host, void *stream);

/* DownloadLayer::pboBlit + CPUBuffer::readFromPBO (fyusenet/gpu/downloadlayer.cpp:112-131,257-283)
 * and DeepDownloadLayer (fyusenet/gpu/deep/deepdownloadlayer.cpp:136-160): tensor -> host float32 in
 * the tensor's own texel order INCLUDING padding: shallow [batch][planes][H+2P][W+2P][4],
 * deep [batch][TH][TW][4] (CPUBuffer GPU_SHALLOW / GPU_DEEP orders).  F16 tensors are widened by a
 * convert kernel into a device staging buffer first.  Asynchronous when `host` is pinned. */
int fyn_download_f32_async(fyn_tensor *tensor, float *host, void *stream);
size_t fyn_download_f32_elems(const fyn_tensor *tensor);
/* the device half of a download: tensor -> float32 RGBA texels in DEVICE memory `device_staging`
 * (fyn_download_f32_elems floats); the host copy is then a plain fyn_memcpy_async on another stream */
int fyn_download_convert(fyn_tensor *tensor, float *device_staging, void *stream);

/* 8-bit I/O.  Upload: UploadLayer with UpDownLayerBuilder::dataType(UBYTE) (fyusenet/gpu/uploadlayer.cpp:51-66,365-375: the bytes
 * become a normalised 8-bit texture, i.e. the layers read value / 255); host order [batch][H][W][C] like the float upload.
 * Download: an extension -- the reference's DownloadLayer reads float32 texels only (gpu/downloadlayer.cpp:257-283) and the
 * samples quantise on the host, (uint8_t)(v * 255) (samples/desktop/stylenet.cpp:52-62); the same conversion (after a clamp to
 * [0, 1]) runs on the device here, so an RGBA frame leaves as 4 instead of 16 bytes per pixel.  Texel order and size in
 * elements are those of the float download (fyn_download_f32_elems). */
int fyn_upload_u8_async(fyn_tensor *tensor, const unsigned char *host_hwc, void *stream);
size_t fyn_download_u8_bytes(const fyn_tensor *tensor);
int fyn_download_u8_convert(fyn_tensor *tensor, unsigned char *device_staging, void *stream);
int fyn_download_u8_async(fyn_tensor *tensor, unsigned char *host, void *stream);

/* Debug / parity interchange (LayerBase::writeResult format, fyusenet/base/layerbase.h:160-172 and
 * GPULayerBase::copyResult, fyusenet/gpu/gpulayerbase.cpp:525-560): float32 [batch][C][H][W]
 * without padding.  Both calls are BLOCKING and go through pageable host memory. */
int fyn_tensor_write_chw_f32(fyn_tensor *tensor, const float *host_chw);
int fyn_tensor_read_chw_f32(fyn_tensor *tensor, float *host_chw);

/* ----------------------------------------------------------------------------------------- */
/* layer ops.  create = repack weights to the device; run = enqueue kernels on `stream`.        */
/* ----------------------------------------------------------------------------------------- */

/* ConvLayerNxN / ConvLayer1x1 / FractionalConvLayerNxN (fyusenet/gpu/vanilla/) and
 * DeepConvLayer1x1 / DeepConvLayerNxN / DeepGEMMLayer (fyusenet/gpu/deep/).
 * Weight blob = ConvLayerInterface::loadWeightsAndBiases format (base/convlayerinterface.h:31-57):
 * bias[Co], W[Co][Ky][Kx][Ci], then with FYN_FLAG_POST_BATCHNORM bnScale[Co], bnBias[Co]. */
typedef struct {
    int width, height;          /* input net size */
    int in_channels, out_channels;
    int kernel;                 /* odd, 1..9 */
    int downsample;             /* isotropic stride (ds) */
    int dilation;
    int in_padding, out_padding, res_padding;
    unsigned flags;             /* FYN_FLAG_* (DEEP selects the tiled layout) */
    float leaky;                /* with PRE_RELU: leak factor (0 = plain ReLU) */
    float clip_lo, clip_hi;     /* with PRE_CLIP */
    float source_step;          /* fractional convs */
    int fractional;             /* LayerType::FRACCONVOLUTION2D */
    int quirks;                 /* FYN_QUIRK_* */
    int backend;                /* 0 = auto, 1 = force direct (CUDA-core) kernel, 2 = force tcgen05 kernel */
} fyn_conv_desc;

int fyn_conv2d_output_size(const fyn_conv_desc *desc, int *out_width, int *out_height);
int fyn_conv2d_create(fyn_ctx *ctx, const fyn_conv_desc *desc, const float *bias_weights_bn, fyn_op **op);
/* hot-swap weights (StyleNet9x9::loadWeightsAndBiases after setup, stylenet9x9.cpp:87-95) */
int fyn_conv2d_load_weights(fyn_op *op, const float *bias_weights_bn);
int fyn_conv2d_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *residual, fyn_tensor *out, void *stream);
/* Engine-level fusion of a DeepBatchNormLayer (fyusenet/gpu/deep/deepbatchnormlayer.cpp:78-147) into the 1x1 convolution
 * that consumes it: the convolution reads the batch-norm layer's INPUT tensor and applies x*scale+bias, rounded to the
 * storage precision exactly like the stand-alone layer's store, before its own prefix activation.  scale_bias = scale[Cin],
 * bias[Cin] (base/batchnorminterface.h:33-48); NULL switches it off.  Only the deep-tiled tcgen05 family with fp16 storage
 * implements it: FYN_ERR_UNSUPPORTED otherwise (the caller keeps the separate layer). */
int fyn_conv2d_set_input_norm(fyn_op *op, const float *scale_bias);
/* which kernel family the op resolved to: 1 = direct, 2 = tcgen05 */
int fyn_conv2d_backend(const fyn_op *op);
/* which kernel the op's LAST run launched (diagnostics / tests; 0 = not run yet): 1 = direct (CUDA cores), 2 = tcgen05 shallow
 * row-ring kernel, 10 = deep one-tile-per-CTA, 11 = deep persistent, 12 = deep halo-tile 3x3,
 * 13 | cluster size << 8 | columns per CTA << 16 = deep split-K cluster kernel (small grids) */
int fyn_conv2d_last_kernel(const fyn_op *op);
/* Fuses the element-wise FunctionLayer that consumes this convolution into its epilogue (engine-level layer
 * fusion; the reference runs it as its own render pass, fyusenet/gpu/functionlayer.cpp:145-179).  `function`:
 * FYN_EPILOGUE_NONE or FYN_EPILOGUE_SIGMOID (fyusenet/gpu/sigmoidlayer.cpp:77-112, shaders/sigmoid.frag:10-17).
 * The convolution result is first rounded to the storage type of the output tensor, so the values equal those of
 * the unfused layer pair. */
#define FYN_EPILOGUE_NONE 0
#define FYN_EPILOGUE_SIGMOID 1
int fyn_conv2d_set_epilogue(fyn_op *op, int function);

/* Persistent chain of shallow convolutions (engine-level fusion; no counterpart in the reference, which renders every layer
 * with its own blend passes: fyusenet/gpu/vanilla/convlayerNxN_vanilla.cpp:72-145 and, for the residual input,
 * shaders/vanilla/residual.inc).  `ops` = n >= 2 convolution ops of identical geometry (stride 1, as many outputs as inputs,
 * tensor padding = kernel / 2) that run on the shallow tcgen05 family, layer i+1 reading layer i's output;
 * residual_from[i] names the layer whose output layer i adds as its residual: it must be i - 2, where -1 stands for the
 * chain input (ignored for layers without FYN_FLAG_RESIDUAL_INPUT).  One kernel launch runs all layers: a strip of layer
 * i+1 waits only for the neighbouring strips of layer i.  Results are bit-identical to running the ops one by one.  The ops
 * stay owned by the caller and must outlive the chain; their weights may be hot-swapped (fyn_conv2d_load_weights).
 * fyn_conv_chain_create returns FYN_ERR_UNSUPPORTED when the layers cannot be chained; fyn_conv_chain_run returns 1
 * (nothing enqueued) when the tensors' formats or batch size are not covered -- run the ops one by one then. */
typedef struct fyn_conv_chain fyn_conv_chain;
int fyn_conv_chain_create(fyn_ctx *ctx, fyn_op *const *ops, const int *residual_from, int n, fyn_conv_chain **chain);
int fyn_conv_chain_layers(const fyn_conv_chain *chain);
int fyn_conv_chain_run(fyn_conv_chain *chain, const fyn_tensor *in, fyn_tensor *out, void *stream);
int fyn_conv_chain_destroy(fyn_conv_chain *chain);

/* Diagnostics (no device needed): the shared-memory plan the tcgen05 family would use for `desc`.  `stack_rows` = 1 or 2
 * job rows per accumulator.  Returns FYN_ERR_UNSUPPORTED if the family does not cover the layer (the direct kernel runs
 * it).  Used by the device-free planner tests and by tools. */
typedef struct {
    int mode;            /* 0 = plane-pair chunks, 1 = pixel-pair chunks (<= 4 input channels) */
    int n;               /* accumulator columns (output channels x stacked phases, padded to 16) */
    int steps;           /* tcgen05.mma instructions per job (without the folded-bias step) */
    int window_rows, row_advance, phases_x, phases_y;
    int ring_slots, mirror_slots, slot_bytes;
    int staged_rows, stage_bytes, row_items;
    int loader_groups, epilogue_warps, bias_folded;
    size_t weight_image_bytes, shared_bytes;
} fyn_conv_plan_info;
int fyn_conv2d_plan_query(const fyn_conv_desc *desc, int stack_rows, fyn_conv_plan_info *info);

/* DeepMaxPoolLayer / DeepAvgPoolLayer / MaxPoolLayer / AvgPoolLayer
 * (fyusenet/gpu/deep/deeppoolinglayer.cpp:38-54,109-190; shaders deep/deepmaxpool.frag, deepavgpool.frag) */
typedef struct {
    int width, height, channels;
    int pool_x, pool_y, downsample;
    int in_padding, out_padding;
    int is_max, global;
    unsigned flags;             /* PRE_RELU / PRE_CLIP / DEEP */
    float leaky, clip_lo, clip_hi;
    int quirks;
} fyn_pool_desc;

int fyn_pool2d_create(fyn_ctx *ctx, const fyn_pool_desc *desc, fyn_op **op);
int fyn_pool2d_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* BatchNormLayer (fyusenet/gpu/batchnormlayer.cpp:71-128, shaders/batchnorm.frag:60-68: x*scale+bias,
 * no activation) and DeepBatchNormLayer (fyusenet/gpu/deep/deepbatchnormlayer.cpp:78-147,
 * deep/deepbatchnorm.frag:57-58: act(x)*scale+bias).  Data = scale[C] then bias[C]. */
typedef struct {
    int width, height, channels;
    int in_padding, out_padding;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
} fyn_bn_desc;

int fyn_batchnorm_create(fyn_ctx *ctx, const fyn_bn_desc *desc, const float *scale_bias, fyn_op **op);
int fyn_batchnorm_load(fyn_op *op, const float *scale_bias);
int fyn_batchnorm_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* SigmoidLayer (fyusenet/gpu/sigmoidlayer.cpp:77-112, shaders/sigmoid.frag:10-17): 1/(1+exp(-act(x)))
 * on every stored lane (unused lanes become 0.5 exactly as in the reference). */
typedef struct {
    int width, height, channels;
    int in_padding, out_padding;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
} fyn_unary_desc;

int fyn_sigmoid_create(fyn_ctx *ctx, const fyn_unary_desc *desc, fyn_op **op);
int fyn_sigmoid_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* Depthwise 3x3 convolution: vanilla::DepthwiseConvLayer3x3
 * (fyusenet/gpu/vanilla/convlayer_dw_3x3_vanilla.cpp:22-75, shaders/vanilla/conv_dw_3x3.frag, weights
 * gpu/convweightarray_dw_KxKxNxM.cpp:120-150) and deep::DeepDepthwiseConvLayer3x3
 * (gpu/deep/deepdwconvlayer3x3.cpp, deepdwconvlayerbase.cpp:40-75, shaders/deep/deepconv_dw3x3_tiled.frag):
 *   out[c,yo,xo] = s[c] * sum_{ky,kx} W[c][ky][kx] * act(in[c, ds*yo + (ky-1)*dil, ds*xo + (kx-1)*dil]) + b'[c]
 * with the clamp-to-edge / zero-padding sampling of the regular convolutions.  Data: bias[C], W[C][3][3], then with
 * POST_BATCHNORM bnScale[C], bnBias[C]; b' = b*s + beta.  The shallow shader has no dilation (taps are +-1 texel);
 * the deep variant keeps its weights fp16-truncated and its bias / scale fp16-rounded with fp16 storage, like the deep
 * convolutions.  FYN_QUIRK_DW_BN_OFFSET reproduces the shallow layer's batch-norm read position (block start instead
 * of behind the weights).
 * Channel multiplier M > 1 (deep layers with channels % 4 == 0 only: the shallow layer throws, convlayer_dw_3x3_vanilla.cpp:49-50;
 * deepdwconvlayerbase.cpp:40-44): the output has M * channels channels, output channel m * channels + c is input channel c
 * filtered with W[c][ky][kx][m] (output tile t + m * tiles reads input tile t, deepdwconvlayerbase.cpp:288-297); data =
 * bias[Co], W[C][3][3][M], then bnScale[Co], bnBias[Co].
 * Residual input (FYN_FLAG_RESIDUAL_INPUT): added behind bias / batch-norm, through ReLU with FYN_FLAG_RELU_ON_RESIDUAL
 * (shaders/vanilla/conv_dw_3x3.frag:133-140, shaders/deep/residual.inc) and, deep layers only, times the batch-norm scale with
 * FYN_FLAG_BATCHNORM_ON_RESIDUAL. */
typedef struct {
    int width, height, channels;
    int downsample, dilation;
    int in_padding, out_padding;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
    int quirks;
    int multiplier;             /* channel multiplier, 0 or 1 = one output per input channel */
    int res_padding;
} fyn_dwconv_desc;

int fyn_dwconv3x3_create(fyn_ctx *ctx, const fyn_dwconv_desc *desc, const float *bias_weights_bn, fyn_op **op);
int fyn_dwconv3x3_load_weights(fyn_op *op, const float *bias_weights_bn);
int fyn_dwconv3x3_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);
int fyn_dwconv3x3_run_residual(fyn_op *op, const fyn_tensor *in, const fyn_tensor *residual, fyn_tensor *out, void *stream);

/* Transpose convolution, stride 2, kernels 2x2 and 3x3, shallow tensors: vanilla::TransConvLayer2x2 / TransConvLayer3x3
 * (fyusenet/gpu/vanilla/transconvlayerbase_vanilla.cpp:44-62,195-215,360-420,499-506, transconvlayer{2x2,3x3}_vanilla.cpp,
 * shaders/vanilla/convtrans{2x2,3x3}_stride2.frag, weights gpu/transconvweightarray{2x2,3x3}xNxM.cpp).  The output is
 * 2W x 2H; output texel o samples the input at texel coordinate P + (o + 0.5) / 2 -+ half a texel, so with i = o / 2:
 *   3x3: even o uses the centre tap on input i; odd o uses tap 0 on input i and tap 2 on input i + 1 (per axis) -- the
 *        correlation of the kernel with the zero-stuffed input; input W (one past the edge) is the padding texel (zero
 *        with in_padding >= 1, the clamped edge texel without).
 *   2x2: out[2j+b][2i+a] = W[b][a] * in[j][i]; with FYN_QUIRK_TRANS2X2_NEXT (reference behaviour) odd rows read input
 *        row j + 1 and odd/odd output texels input (i + 1, j + 1), while odd columns of even rows read column i.
 * Data: bias[Co], W[Co][K][K][Ci], then with POST_BATCHNORM bnScale[Co], bnBias[Co] (b' = b*s + beta).
 * FYN_FLAG_DEEP selects deep::DeepTransConvLayer2x2 / 3x3 (gpu/deep/deeptransconvlayerbase.cpp:44-232, deeptransconvlayer{2x2,3x3}.cpp,
 * shaders/deep/deeptransconv{2x2,3x3}_stride2.{vert,frag}) on deep-tiled tensors, whose passes align the taps differently:
 *   3x3: even o uses tap 0 on input i and tap 2 on input i - 1, odd o tap 1 on input i (per axis); 2x2: tap (o & 1) on input i
 *   -- the full convolution of the zero-stuffed input with the kernel; reads outside the image are zero whatever the padding;
 *   with fp16 storage the weights are fp16-truncated and bias / scale fp16-rounded like the other deep layers.
 * A residual input is refused like in the reference ("Transpose convolutions do not support residuals as of now"). */
typedef struct {
    int width, height;
    int in_channels, out_channels;
    int kernel;                 /* 2 or 3 */
    int in_padding, out_padding;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
    int quirks;
} fyn_transconv_desc;

int fyn_transconv2d_create(fyn_ctx *ctx, const fyn_transconv_desc *desc, const float *bias_weights_bn, fyn_op **op);
int fyn_transconv2d_load_weights(fyn_op *op, const float *bias_weights_bn);
int fyn_transconv2d_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* ScaleLayer / DeepScaleLayer (fyusenet/gpu/scalelayer.cpp:40-60,120-140, gpu/deep/deepscalelayer.cpp:30-75,
 * shaders/scaling.frag, geometry of gpu/functionlayer.cpp:194-204): output texel o samples the input at the
 * texel-space coordinate P + (o + 0.5) * W / Wo, with Wo = (int)(W * up / down) (scalelayer.cpp:44-47).
 * NEAREST takes the texel that contains the coordinate, LINEAR (GL_LINEAR) blends the four texels around it --
 * including the padding texels / clamped texture edge exactly like the sampler -- and the activation is applied
 * to the sampled value.  Deep layers of 1-texel width or height always sample NEAREST (deepscalelayer.cpp:34).
 * With all factors 1 this is also the reference's PADDING2D / RELU / CLIP pseudo-layer
 * (gpu/gpulayerfactory.cpp:125-140,317-322).  Rotation is not supported. */
typedef struct {
    int width, height, channels;          /* input size */
    int in_padding, out_padding;
    int upsample_x, upsample_y;           /* integer factors (gpu/scalelayerbuilder.h:60-95) */
    int downsample_x, downsample_y;
    int linear;                           /* 0 = ScalingType::NEAREST, 1 = ScalingType::LINEAR */
    unsigned flags;                       /* FYN_FLAG_DEEP, FYN_FLAG_PRE_* */
    float leaky, clip_lo, clip_hi;
} fyn_scale_desc;

int fyn_scale_create(fyn_ctx *ctx, const fyn_scale_desc *desc, fyn_op **op);
int fyn_scale_out_size(const fyn_scale_desc *desc, int *width, int *height);
int fyn_scale_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* AddSubLayer (fyusenet/gpu/addsublayer.cpp:60-120, shaders/add.frag:84-135: act(a) +/- act(b)) and
 * SingletonArithmeticLayer (fyusenet/gpu/singleton_arithlayer.cpp, shaders/singleton_arith.frag:
 * act(x) (+|-|*|/) operand).  Known answers: unit_tests/arithtests.cpp:97-250. */
#define FYN_ARITH_ADD 0
#define FYN_ARITH_SUB 1
#define FYN_ARITH_MUL 2
#define FYN_ARITH_DIV 3
typedef struct {
    int width, height, channels;
    int in_padding, out_padding;
    int op;                               /* FYN_ARITH_*; two-tensor form: ADD / SUB only */
    int singleton;                        /* 1: second operand is the scalar `operand` */
    float operand;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
} fyn_arith_desc;

int fyn_arith_create(fyn_ctx *ctx, const fyn_arith_desc *desc, fyn_op **op);
int fyn_arith_run(fyn_op *op, const fyn_tensor *in1, const fyn_tensor *in2, fyn_tensor *out, void *stream);

/* ConcatLayer / DeepConcatLayer (fyusenet/gpu/concatlayer.cpp:60-75,112-180, gpu/deep/deepconcatlayer.cpp,
 * shaders/vanilla/concat.frag): the output holds the inputs' channels back to back (inputs whose channel count is
 * not a multiple of 4 are shifted across texels: the reference's "consolidation render"); the activation applies
 * to every input (the reference supports ReLU on all inputs or none, concatlayer.cpp:20-26). */
#define FYN_CONCAT_MAX_INPUTS 8
typedef struct {
    int width, height;
    int num_inputs;
    int channels[FYN_CONCAT_MAX_INPUTS];
    int in_padding, out_padding;
    unsigned flags;
    float leaky, clip_lo, clip_hi;
} fyn_concat_desc;

int fyn_concat_create(fyn_ctx *ctx, const fyn_concat_desc *desc, fyn_op **op);
int fyn_concat_run(fyn_op *op, const fyn_tensor *const *inputs, int num_inputs, fyn_tensor *out, void *stream);

/* RGB2BGRLayer (fyusenet/gpu/rgb2bgrlayer.cpp, shaders/rgb2bgr.frag: every texel becomes val.bgra, no
 * activation). */
int fyn_rgb2bgr_create(fyn_ctx *ctx, const fyn_unary_desc *desc, fyn_op **op);
int fyn_rgb2bgr_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

/* Shallow2DeepLayer / Deep2ShallowLayer (fyusenet/gpu/shallow2deep.cpp, gpu/deep2shallow.cpp,
 * shaders/shallow2deep.frag, deep2shallow.frag): layout conversion between the per-plane and the tiled texture
 * format with the activation applied at the fetch.  The direction follows from the two tensors' orders. */
int fyn_relayout_create(fyn_ctx *ctx, const fyn_unary_desc *desc, fyn_op **op);
int fyn_relayout_run(fyn_op *op, const fyn_tensor *in, fyn_tensor *out, void *stream);

int fyn_op_destroy(fyn_op *op);

/* ----------------------------------------------------------------------------------------- */
/* multi-GPU: one process per GPU.  The reference is a single-GPU, batch-1  */
/* engine (README.md:72); these entry points serve the two workloads that shard naturally.      */
/* ----------------------------------------------------------------------------------------- */
typedef struct fyn_comm fyn_comm;
#define FYN_COMM_ID_BYTES 128

/* Bootstrap: rank 0 creates the id (an ncclUniqueId), the caller hands the 128 bytes to every rank over any channel it has
 * (torch.distributed, MPI, a file), every rank then calls fyn_comm_init -- a collective.  world == 1 needs no id and no NCCL.
 * NCCL is loaded with dlopen("libnccl.so.2") (inside a torch process: the copy torch already loaded).  The communicator also
 * exchanges the CUDA IPC handles of the flag blocks the halo exchange hand-shakes through. */
int fyn_comm_unique_id(void *id128);
int fyn_comm_init(fyn_ctx *ctx, int rank, int world, const void *id128, fyn_comm **comm);
int fyn_comm_destroy(fyn_comm *comm);
int fyn_comm_info(const fyn_comm *comm, int *rank, int *world, int *nccl_version, uint64_t *halo_bytes_pushed);

/* Batch-sharded ResNet-50 (BASELINE configs[3]): every rank holds the logits of its image shard in the DeepGEMMLayer's output
 * tensor (deep layout, 1x1 spatial, C channels, batch = local images; fyusenet/gpu/deep/deepgemmlayer.cpp:66-140).  Converts
 * them to float32 [images_per_rank][C] (rows beyond the local batch are zero: shards may differ by one image) and all-gathers
 * over NCCL into `device_out` = float32 [world][images_per_rank][C] in DEVICE memory, on `stream`.  What DeepDownloadLayer +
 * CPUBuffer::toChannelWise deliver for one image (deepdownloadlayer.cpp:136-160, cpu/cpubuffer.cpp:131-142), for the job. */
int fyn_allgather_logits(fyn_comm *comm, const fyn_tensor *logits, int images_per_rank, float *device_out, void *stream);

/* Row-banded StyleNet (BASELINE configs[4]): rank r holds rows [b_r - m, e_r + m) of every layer's tensor -- its band plus a
 * margin of m rows towards each existing neighbour (none at the true image border, where the reference's clamp-to-edge
 * applies).  A layer computes its band rows exactly as long as its taps stay inside band + margin (tap geometry:
 * fyusenet/gpu/vanilla/convlayerbase_vanilla.cpp:347-371, fractionalconvlayerNxN_vanilla.cpp:43-51); its margin rows are then
 * REPLACED by the neighbours' band-edge rows.
 *   fyn_comm_register_tensor  collective, set-up time: publishes the tensor's CUDA IPC handle and maps the tensors the ranks
 *                             r-1 / r+1 registered in the same call (same order on all ranks).  Shallow RGBA tensors
 *                             allocated by this library only.  Returns the slot to pass to fyn_halo_exchange.
 *   fyn_halo_exchange         enqueues ONE kernel on `stream`: pushes the first / last `rows` band rows of the tensor into the
 *                             neighbours' margins with peer stores over NVLink (no NCCL, no host copy), hand-shaking through
 *                             flag words in peer memory; when the kernel ends this rank's margins hold the neighbours' rows.
 *                             Every rank must issue the same sequence of exchanges. */
int fyn_comm_register_tensor(fyn_comm *comm, fyn_tensor *tensor, int *slot);
int fyn_halo_exchange(fyn_comm *comm, int slot, int rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FYUSENET_B200_H */
