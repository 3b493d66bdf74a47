// Device layer classes: thin state holders that translate builder parameters into C-ABI descriptors and
// enqueue the ops on the network stream.
#include <cstdlib>
#include <cstring>

#include "fyusenet/gpu/cudalayers.h"

namespace fyusion {
namespace fyusenet {
namespace gpu {

static int envInt(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// GPULayerBase
// ------------------------------------------------------------------------------------------------
int GPULayerBase::outputBatch() const {
    fyn_tensor_desc d{};
    FYN_ABI_CALL(fyn_tensor_get_desc(out(), &d, nullptr));
    return d.batch;
}

void GPULayerBase::copyResult(float *memory, bool includePadding) {
    if (includePadding) THROW_EXCEPTION_ARGS(FynException, "Padded result copies are not supported");
    FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
    FYN_ABI_CALL(fyn_tensor_read_chw_f32(out(), memory));
}

void GPULayerBase::writeResult(const char *fileName, bool includePadding) {
    fyn_tensor_desc d{};
    FYN_ABI_CALL(fyn_tensor_get_desc(out(), &d, nullptr));
    std::vector<float> data((size_t)d.batch * d.channels * d.height * d.width);
    copyResult(data.data(), includePadding);
    FILE *f = fopen(fileName, "wb");
    if (!f) THROW_EXCEPTION_ARGS(FynException, "Cannot open file %s for writing", fileName);
    fwrite(data.data(), sizeof(float), data.size(), f);
    fclose(f);
}

// ------------------------------------------------------------------------------------------------
// ConvLayerBase
// ------------------------------------------------------------------------------------------------
void ConvLayerBase::init(int kernel, int dilation, float sourceStep, bool fractional) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.in_channels = inputChannels_;
    desc_.out_channels = outputChannels_;
    desc_.kernel = kernel;
    desc_.dilation = dilation;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.res_padding = residualPadding_;
    desc_.flags = flags_ & (LayerFlags::RESIDUAL_INPUT | LayerFlags::RELU_ON_RESIDUAL | LayerFlags::BATCHNORM_ON_RESIDUAL |
                            LayerFlags::POST_BATCHNORM | LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    desc_.source_step = sourceStep;
    desc_.fractional = fractional ? 1 : 0;
    desc_.quirks = envInt("FYN_QUIRKS", FYN_QUIRKS_REFERENCE);
    desc_.backend = envInt("FYN_CONV_BACKEND", 0);
    FYN_ABI_CALL(fyn_conv2d_output_size(&desc_, &outWidth_, &outHeight_));
    viewport_[0] = outWidth_ + 2 * outputPadding_;
    viewport_[1] = outHeight_ + 2 * outputPadding_;
}

ConvLayerBase::ConvLayerBase(const ConvLayerBuilder &builder, int layerNumber, bool fractional) : GPULayerBase(builder, layerNumber) {
    if (builder.downsample_[0] != builder.downsample_[1])
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: anisotropic downsampling not supported", name_.c_str());
    if (builder.dilation_[0] != builder.dilation_[1])
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: anisotropic dilation not supported", name_.c_str());
    if (builder.kernel_ < 1 || !(builder.kernel_ & 1))
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: kernel size %d not supported", name_.c_str(), builder.kernel_);
    if (fractional && builder.dilation_[0] > 1)
        THROW_EXCEPTION_ARGS(FynException, "Dilations not supported for fractional convolution");
    desc_.downsample = builder.downsample_[0];
    init(builder.kernel_, builder.dilation_[0], builder.sourceStep_, fractional);
}

ConvLayerBase::ConvLayerBase(const GPULayerBuilder &builder, int layerNumber) : GPULayerBase(builder, layerNumber) {
    // GEMM: generalized matrix/vector product run as a 1x1 convolution (reference: gpu/deep/deepgemmlayer.cpp:66-140)
    desc_.downsample = 1;
    init(1, 1, 1.f, false);
}

ConvLayerBase::~ConvLayerBase() {}

std::vector<BufferSpec> ConvLayerBase::getRequiredInputBuffers() const {
    std::vector<BufferSpec> r;
    BufferSpec in0(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_SOURCE);
    // inputs with fewer than 4 channels may come straight from an upload texture (RGB32F):
    // reference gpu/vanilla/convlayerbase_vanilla.cpp:205-210
    if (inputChannels_ < PIXEL_PACKING) in0.anyType();
    r.push_back(in0);
    if (flags_ & LayerFlags::RESIDUAL_INPUT)
        r.push_back(BufferSpec(1, outWidth_, outHeight_, outputChannels_, residualPadding_, order(), storagePrecision(), BufferSpec::RESIDUAL_SOURCE));
    return r;
}

std::vector<BufferSpec> ConvLayerBase::getRequiredOutputBuffers() const {
    return {BufferSpec(0, outWidth_, outHeight_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_DEST)};
}

void ConvLayerBase::loadWeightsAndBiases(const float *biasAndWeights, size_t offset) {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!biasAndWeights) THROW_EXCEPTION_ARGS(FynException, "Layer %s: null weight pointer", name_.c_str());
    size_t n = (size_t)outputChannels_ + (size_t)desc_.kernel * desc_.kernel * inputChannels_ * outputChannels_;
    if (flags_ & LayerFlags::POST_BATCHNORM) n += 2 * (size_t)outputChannels_;
    const float *src = biasAndWeights + offset;
    if (op_) {
        // hot swap (reference: stylenet9x9.cpp:87-95, serialised there by the GL command stream): the library waits for the
        // device before it overwrites live weight images (fyn_conv2d_load_weights)
        FYN_ABI_CALL(fyn_conv2d_load_weights(op_, src));
        graphEpoch()++;
    } else {
        pendingWeights_.assign(src, src + n);              // weights are only read during the call
    }
}

void ConvLayerBase::setup() {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (pendingWeights_.empty())
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: loadWeightsAndBiases() must be called before setup()", name_.c_str());
    FYN_ABI_CALL(fyn_conv2d_create(context_.handle(), &desc_, pendingWeights_.data(), &op_));
    pendingWeights_.clear();
    pendingWeights_.shrink_to_fit();
    valid_ = true;
}

void ConvLayerBase::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}

void ConvLayerBase::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    if (chainMember_) return;                          // the head of the chain has computed this layer's output
    if (chain_) {
        std::lock_guard<std::recursive_mutex> lck(processingLock_);
        ConvLayerBase *last = chainFollowers_.back();
        const int rc = fyn_conv_chain_run(chain_, in(0), last->out(), context_.stream());
        if (rc == 0) return;
        if (rc != 1) FYN_ABI_CALL(rc);
        // tensor formats the chain does not cover (e.g. fp32 storage): the layers run one by one
        forwardSingle();
        for (ConvLayerBase *f : chainFollowers_) f->forwardSingle();
        return;
    }
    forwardSingle();
}

void ConvLayerBase::setChainHead(fyn_conv_chain *chain, const std::vector<ConvLayerBase *> &followers) {
    chain_ = chain;
    chainFollowers_ = followers;
    graphEpoch()++;
}

void ConvLayerBase::unchain() {
    if (chain_ || chainMember_) graphEpoch()++;
    chain_ = nullptr;
    chainFollowers_.clear();
    chainMember_ = false;
}

void ConvLayerBase::forwardSingle() {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    TensorHandle res = nullptr;
    if (flags_ & LayerFlags::RESIDUAL_INPUT) {
        if (residuals_.empty() || !residuals_[0])
            THROW_EXCEPTION_ARGS(FynException, "Residual flag configured, but no such texture found.");
        res = residuals_[0];
    }
    FYN_ABI_CALL(fyn_conv2d_run(op_, fusedInput_ ? fusedInput_ : in(0), res, fusedTarget_ ? fusedTarget_ : out(), context_.stream()));
}

bool ConvLayerBase::fuseFunction(int function, TensorHandle target) {
    if (!op_ || !target || !hasOutputTexture(0)) return false;
    // the consumer's output must be laid out exactly like ours and must not alias anything this layer reads
    fyn_tensor_desc mine{}, theirs{};
    FYN_ABI_CALL(fyn_tensor_get_desc(out(), &mine, nullptr));
    FYN_ABI_CALL(fyn_tensor_get_desc(target, &theirs, nullptr));
    if (memcmp(&mine, &theirs, sizeof(mine)) != 0) return false;
    if (target == in(0) || (!residuals_.empty() && target == residuals_[0])) return false;
    if (fusedInput_) return false;                       // one fusion per convolution: the kernel families implement either
    if (fyn_conv2d_set_epilogue(op_, function) != 0) return false;   // (fails for the deep-tiled tcgen05 family: no fused function there)
    fusedFunction_ = function;
    fusedTarget_ = target;
    return true;
}

bool ConvLayerBase::fuseInputNorm(const float *scaleAndBias, TensorHandle source) {
    if (!op_ || !source || !scaleAndBias || !hasInputTexture(0) || hasOutputTexture(0) == false) return false;
    // the batch-norm layer's input must be laid out exactly like its output (our regular input), in fp16 storage
    fyn_tensor_desc mine{}, theirs{};
    FYN_ABI_CALL(fyn_tensor_get_desc(in(0), &mine, nullptr));
    FYN_ABI_CALL(fyn_tensor_get_desc(source, &theirs, nullptr));
    if (memcmp(&mine, &theirs, sizeof(mine)) != 0 || mine.dtype != FYN_F16) return false;
    if (source == out() || source == fusedTarget_ || fusedFunction_ != 0) return false;
    if (fyn_conv2d_set_input_norm(op_, scaleAndBias) != 0) return false;
    fusedInput_ = source;
    return true;
}

void ConvLayerBase::unfuseInput() {
    if (op_) fyn_conv2d_set_input_norm(op_, nullptr);
    fusedInput_ = nullptr;
}

void ConvLayerBase::unfuse() {
    if (op_) fyn_conv2d_set_epilogue(op_, FYN_EPILOGUE_NONE);
    fusedFunction_ = 0;
    fusedTarget_ = nullptr;
}

// ------------------------------------------------------------------------------------------------
// PoolingLayer
// ------------------------------------------------------------------------------------------------
PoolingLayer::PoolingLayer(const PoolLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.global = b.global_ ? 1 : 0;
    // global pooling: window = stride = spatial size (reference: gpu/deep/deeppoolinglayer.cpp:38-54)
    desc_.pool_x = b.global_ ? width_ : b.poolsize_[0];
    desc_.pool_y = b.global_ ? height_ : b.poolsize_[1];
    if (!b.global_ && b.downsample_[0] != b.downsample_[1])
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: anisotropic pooling stride not supported", name_.c_str());
    desc_.downsample = b.downsample_[0];
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.is_max = (b.operation_ == PoolLayerBuilder::POOL_MAX) ? 1 : 0;
    desc_.flags = flags_ & (LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    desc_.quirks = envInt("FYN_QUIRKS", FYN_QUIRKS_REFERENCE);
    outWidth_ = b.global_ ? 1 : width_ / desc_.downsample;
    outHeight_ = b.global_ ? 1 : height_ / desc_.downsample;
    viewport_[0] = outWidth_ + 2 * outputPadding_;
    viewport_[1] = outHeight_ + 2 * outputPadding_;
}

std::vector<BufferSpec> PoolingLayer::getRequiredInputBuffers() const {
    return {BufferSpec(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE)};
}
std::vector<BufferSpec> PoolingLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, outWidth_, outHeight_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void PoolingLayer::setup() {
    FYN_ABI_CALL(fyn_pool2d_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void PoolingLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void PoolingLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    FYN_ABI_CALL(fyn_pool2d_run(op_, in(0), out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// BatchNormLayer
// ------------------------------------------------------------------------------------------------
BatchNormLayer::BatchNormLayer(const GPULayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = flags_ & (LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
}
std::vector<BufferSpec> BatchNormLayer::getRequiredInputBuffers() const {
    BufferSpec s(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE);
    if (inputChannels_ < PIXEL_PACKING) s.anyType();
    return {s};
}
std::vector<BufferSpec> BatchNormLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void BatchNormLayer::loadScaleAndBias(const float *scaleAndBias, size_t sbOffset) {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!scaleAndBias) THROW_EXCEPTION_ARGS(FynException, "Layer %s: null parameter pointer", name_.c_str());
    params_.assign(scaleAndBias + sbOffset, scaleAndBias + sbOffset + 2 * (size_t)outputChannels_);
    if (op_) {
        FYN_ABI_CALL(fyn_batchnorm_load(op_, params_.data()));
        graphEpoch()++;
    }
    if (fusedConsumer_ && hasInputTexture(0) && !fusedConsumer_->fuseInputNorm(params_.data(), in(0))) {
        fusedConsumer_->unfuseInput();   // new parameters could not be handed over: run as a layer again
        fusedConsumer_ = nullptr;
    }
}
void BatchNormLayer::setup() {
    if (params_.empty()) THROW_EXCEPTION_ARGS(FynException, "Layer %s: loadScaleAndBias() must be called before setup()", name_.c_str());
    FYN_ABI_CALL(fyn_batchnorm_create(context_.handle(), &desc_, params_.data(), &op_));
    valid_ = true;
}
void BatchNormLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void BatchNormLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (fusedConsumer_) return;   // evaluated at the consumer's fetch
    FYN_ABI_CALL(fyn_batchnorm_run(op_, in(0), out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// SigmoidLayer
// ------------------------------------------------------------------------------------------------
SigmoidLayer::SigmoidLayer(const GPULayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = flags_ & (LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
}
std::vector<BufferSpec> SigmoidLayer::getRequiredInputBuffers() const {
    return {BufferSpec(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE)};
}
std::vector<BufferSpec> SigmoidLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void SigmoidLayer::setup() {
    FYN_ABI_CALL(fyn_sigmoid_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void SigmoidLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void SigmoidLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (bypass_) return;   // evaluated in the producer's epilogue
    FYN_ABI_CALL(fyn_sigmoid_run(op_, in(0), out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// DepthwiseConvLayer
// ------------------------------------------------------------------------------------------------
DepthwiseConvLayer::DepthwiseConvLayer(const ConvLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    if (b.kernel_ != 3) THROW_EXCEPTION_ARGS(FynException, "Layer %s: depthwise convolution supports 3x3 kernels only", name_.c_str());
    // channel multiplier = outputs / group size (convlayer_dw_3x3_vanilla.cpp:49-50: shallow layers throw for != 1;
    // deepdwconvlayerbase.cpp:40-44: deep layers need input channels % 4 == 0)
    if (b.groupSize_ != inputChannels_ || outputChannels_ % inputChannels_ != 0)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: depthwise convolution needs group size == input channels and outputs a multiple of it", name_.c_str());
    const int multiplier = outputChannels_ / b.groupSize_;
    if (multiplier != 1 && !(flags_ & LayerFlags::DEEP)) THROW_EXCEPTION_ARGS(FynException, "Channel multipliers are currently not supported");
    if (multiplier > 1 && (inputChannels_ & 3))
        THROW_EXCEPTION_ARGS(FynException, "Channel multipliers > 1 are only supported on input channels being a multiple of 4");
    if (b.downsample_[0] != b.downsample_[1] || b.dilation_[0] != b.dilation_[1])
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: anisotropic downsampling / dilation not supported", name_.c_str());
    desc_.multiplier = multiplier;
    desc_.res_padding = residualPadding_;
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.downsample = b.downsample_[0];
    desc_.dilation = b.dilation_[0];
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = flags_ & (LayerFlags::POST_BATCHNORM | LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP | LayerFlags::RESIDUAL_INPUT |
                            LayerFlags::RELU_ON_RESIDUAL | LayerFlags::BATCHNORM_ON_RESIDUAL);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    desc_.quirks = envInt("FYN_QUIRKS", FYN_QUIRKS_REFERENCE);
    outWidth_ = width_ / desc_.downsample;
    outHeight_ = height_ / desc_.downsample;
    viewport_[0] = outWidth_ + 2 * outputPadding_;
    viewport_[1] = outHeight_ + 2 * outputPadding_;
}
std::vector<BufferSpec> DepthwiseConvLayer::getRequiredInputBuffers() const {
    BufferSpec in0(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_SOURCE);
    if (inputChannels_ < PIXEL_PACKING) in0.anyType();
    std::vector<BufferSpec> r{in0};
    if (flags_ & LayerFlags::RESIDUAL_INPUT)
        r.push_back(BufferSpec(1, outWidth_, outHeight_, outputChannels_, residualPadding_, order(), storagePrecision(), BufferSpec::RESIDUAL_SOURCE));
    return r;
}
std::vector<BufferSpec> DepthwiseConvLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, outWidth_, outHeight_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_DEST)};
}
void DepthwiseConvLayer::loadWeightsAndBiases(const float *biasAndWeights, size_t offset) {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!biasAndWeights) THROW_EXCEPTION_ARGS(FynException, "Layer %s: null weight pointer", name_.c_str());
    size_t n = (size_t)outputChannels_ + (size_t)outputChannels_ * 9;          // bias[Co], W[Ci][3][3][multiplier]
    if (flags_ & LayerFlags::POST_BATCHNORM) n += 2 * (size_t)outputChannels_;
    const float *src = biasAndWeights + offset;
    if (op_) {
        FYN_ABI_CALL(fyn_dwconv3x3_load_weights(op_, src));
        graphEpoch()++;
    }
    else pendingWeights_.assign(src, src + n);
}
void DepthwiseConvLayer::setup() {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (pendingWeights_.empty())
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: loadWeightsAndBiases() must be called before setup()", name_.c_str());
    FYN_ABI_CALL(fyn_dwconv3x3_create(context_.handle(), &desc_, pendingWeights_.data(), &op_));
    pendingWeights_.clear();
    pendingWeights_.shrink_to_fit();
    valid_ = true;
}
void DepthwiseConvLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void DepthwiseConvLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    TensorHandle res = nullptr;
    if (flags_ & LayerFlags::RESIDUAL_INPUT) {
        if (residuals_.empty() || !residuals_[0]) THROW_EXCEPTION_ARGS(FynException, "Residual flag configured, but no such texture found.");
        res = residuals_[0];
    }
    FYN_ABI_CALL(fyn_dwconv3x3_run_residual(op_, in(0), res, out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// TransConvLayer
// ------------------------------------------------------------------------------------------------
TransConvLayer::TransConvLayer(const ConvLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    // transconvlayerbase_vanilla.cpp:44-62
    if (b.upsample_[0] != 2 || b.upsample_[1] != 2) THROW_EXCEPTION_ARGS(FynException, "Only stride 2 transpose conv layers are supported for now");
    if (b.kernel_ != 2 && b.kernel_ != 3) THROW_EXCEPTION_ARGS(FynException, "Layer %s: transpose convolution supports 2x2 and 3x3 kernels", name_.c_str());
    // (deep::DeepTransConvLayer2x2 / 3x3 are the same class with the DEEP flag; both variants refuse a residual input like
    // the reference: deeptransconvlayer3x3.cpp:44-46)
    if (flags_ & LayerFlags::RESIDUAL_INPUT) THROW_EXCEPTION_ARGS(FynException, "Transpose convolutions do not support residuals as of now");
    desc_.width = width_;
    desc_.height = height_;
    desc_.in_channels = inputChannels_;
    desc_.out_channels = outputChannels_;
    desc_.kernel = b.kernel_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = flags_ & (LayerFlags::POST_BATCHNORM | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP | LayerFlags::DEEP);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    desc_.quirks = envInt("FYN_QUIRKS", FYN_QUIRKS_REFERENCE);
    viewport_[0] = 2 * width_ + 2 * outputPadding_;
    viewport_[1] = 2 * height_ + 2 * outputPadding_;
}
std::vector<BufferSpec> TransConvLayer::getRequiredInputBuffers() const {
    BufferSpec in0(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_SOURCE);
    if (inputChannels_ < PIXEL_PACKING) in0.anyType();
    return {in0};
}
std::vector<BufferSpec> TransConvLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, 2 * width_, 2 * height_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::CONVOLUTION_DEST)};
}
void TransConvLayer::loadWeightsAndBiases(const float *biasAndWeights, size_t offset) {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!biasAndWeights) THROW_EXCEPTION_ARGS(FynException, "Layer %s: null weight pointer", name_.c_str());
    size_t n = (size_t)outputChannels_ + (size_t)desc_.kernel * desc_.kernel * inputChannels_ * outputChannels_;
    if (flags_ & LayerFlags::POST_BATCHNORM) n += 2 * (size_t)outputChannels_;
    const float *src = biasAndWeights + offset;
    if (op_) {
        FYN_ABI_CALL(fyn_transconv2d_load_weights(op_, src));
        graphEpoch()++;
    }
    else pendingWeights_.assign(src, src + n);
}
void TransConvLayer::setup() {
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (pendingWeights_.empty())
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: loadWeightsAndBiases() must be called before setup()", name_.c_str());
    FYN_ABI_CALL(fyn_transconv2d_create(context_.handle(), &desc_, pendingWeights_.data(), &op_));
    pendingWeights_.clear();
    pendingWeights_.shrink_to_fit();
    valid_ = true;
}
void TransConvLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void TransConvLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    FYN_ABI_CALL(fyn_transconv2d_run(op_, in(0), out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// ScaleLayer (+ PADDING2D / RELU / CLIP), ArithLayer, ConcatLayer, UnaryCopyLayer
// ------------------------------------------------------------------------------------------------
static unsigned gatherFlags(layerflags f) { return f & (LayerFlags::DEEP | LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP); }

void ScaleLayer::init(int upx, int upy, int dnx, int dny, ScalingType type) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.upsample_x = upx;
    desc_.upsample_y = upy;
    desc_.downsample_x = dnx;
    desc_.downsample_y = dny;
    desc_.linear = type == ScalingType::LINEAR;
    desc_.flags = gatherFlags(flags_);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    if (inputChannels_ != outputChannels_)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: scaling cannot change the channel count (%d -> %d)", name_.c_str(), inputChannels_, outputChannels_);
    FYN_ABI_CALL(fyn_scale_out_size(&desc_, &outWidth_, &outHeight_));
    viewport_[0] = outWidth_ + 2 * outputPadding_;   // gpu/scalelayer.cpp:44-47
    viewport_[1] = outHeight_ + 2 * outputPadding_;
}
ScaleLayer::ScaleLayer(const ScaleLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    if (b.rotation_ != 0) THROW_EXCEPTION_ARGS(FynException, "Layer %s: rotation is not supported by the CUDA backend", name_.c_str());
    init(b.upsample_[0], b.upsample_[1], b.downsample_[0], b.downsample_[1], b.scaleType_);
}
ScaleLayer::ScaleLayer(const GPULayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) { init(1, 1, 1, 1, ScalingType::NEAREST); }
std::vector<BufferSpec> ScaleLayer::getRequiredInputBuffers() const {
    return {BufferSpec(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE).anyType()};
}
std::vector<BufferSpec> ScaleLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, outWidth_, outHeight_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void ScaleLayer::setup() {
    FYN_ABI_CALL(fyn_scale_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void ScaleLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void ScaleLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    FYN_ABI_CALL(fyn_scale_run(op_, in(0), out(), context_.stream()));
}

void ArithLayer::init() {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = gatherFlags(flags_);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    if (flags_ & LayerFlags::POST_BATCHNORM)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: post-batchnorm on arithmetic layers is not supported by the CUDA backend", name_.c_str());
}
ArithLayer::ArithLayer(const GPULayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    if (b.type_ != LayerType::ADD && b.type_ != LayerType::SUB)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: unsupported operation (only ADD / SUB)", name_.c_str());   // gpu/addsublayer.cpp
    desc_.op = b.type_ == LayerType::ADD ? FYN_ARITH_ADD : FYN_ARITH_SUB;
    desc_.singleton = 0;
    init();
}
ArithLayer::ArithLayer(const SingletonArithLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    desc_.op = (int)b.opType_;   // ArithType and FYN_ARITH_* share their numbering
    desc_.singleton = 1;
    desc_.operand = b.operand_;
    init();
}
std::vector<BufferSpec> ArithLayer::getRequiredInputBuffers() const {
    std::vector<BufferSpec> specs;
    for (int port = 0; port < numInputPorts(); port++)
        specs.push_back(BufferSpec(port, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE).anyType());
    return specs;
}
std::vector<BufferSpec> ArithLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void ArithLayer::setup() {
    FYN_ABI_CALL(fyn_arith_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void ArithLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void ArithLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    FYN_ABI_CALL(fyn_arith_run(op_, in(0), desc_.singleton ? nullptr : in(1), out(), context_.stream()));
}

ConcatLayer::ConcatLayer(const ConcatLayerBuilder &b, int layerNumber) : GPULayerBase(b, layerNumber) {
    if (b.inputs_.empty()) THROW_EXCEPTION_ARGS(FynException, "No inputs allocated, please use input()");
    if ((int)b.inputs_.size() > FYN_CONCAT_MAX_INPUTS)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: at most %d inputs", name_.c_str(), FYN_CONCAT_MAX_INPUTS);
    desc_.width = width_;
    desc_.height = height_;
    desc_.num_inputs = (int)b.inputs_.size();
    int total = 0, relu = 0;
    for (size_t i = 0; i < b.inputs_.size(); i++) {
        // gpu/concatlayer.cpp:60-64
        if (b.inputs_[i].padding != inputPadding_)
            THROW_EXCEPTION_ARGS(FynException, "Mismatch on input padding (%d vs %d)", inputPadding_, b.inputs_[i].padding);
        desc_.channels[i] = b.inputs_[i].channels;
        total += b.inputs_[i].channels;
        if (b.inputs_[i].flags & LayerFlags::PRE_RELU) relu++;
    }
    // ReLU on all inputs or on none (gpu/concatlayer.cpp:20-26)
    if (relu == desc_.num_inputs) flags_ |= LayerFlags::PRE_RELU;
    else if (relu > 0) THROW_EXCEPTION_ARGS(FynException, "Layer %s: reLU/non-reLU concats are not supported", name_.c_str());
    if (outputChannels_ != total)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: %d output channels, inputs add up to %d", name_.c_str(), outputChannels_, total);
    inputChannels_ = total;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = gatherFlags(flags_);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
}
int ConcatLayer::numInputChannels(int port) const {
    if (port < 0 || port >= desc_.num_inputs) THROW_EXCEPTION_ARGS(FynException, "Illegal input port %d specified", port);
    return desc_.channels[port];
}
std::vector<BufferSpec> ConcatLayer::getRequiredInputBuffers() const {
    std::vector<BufferSpec> specs;
    for (int port = 0; port < desc_.num_inputs; port++)
        specs.push_back(BufferSpec(port, width_, height_, desc_.channels[port], inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE).anyType());
    return specs;
}
std::vector<BufferSpec> ConcatLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void ConcatLayer::setup() {
    FYN_ABI_CALL(fyn_concat_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void ConcatLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void ConcatLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    const fyn_tensor *ins[FYN_CONCAT_MAX_INPUTS];
    for (int port = 0; port < desc_.num_inputs; port++) ins[port] = in(port);
    FYN_ABI_CALL(fyn_concat_run(op_, ins, desc_.num_inputs, out(), context_.stream()));
}

UnaryCopyLayer::UnaryCopyLayer(const GPULayerBuilder &b, int layerNumber, Kind kind) : GPULayerBase(b, layerNumber), kind_(kind) {
    desc_.width = width_;
    desc_.height = height_;
    desc_.channels = inputChannels_;
    desc_.in_padding = inputPadding_;
    desc_.out_padding = outputPadding_;
    desc_.flags = gatherFlags(flags_);
    desc_.leaky = leakyReLU_;
    desc_.clip_lo = lowClip_;
    desc_.clip_hi = highClip_;
    if (inputChannels_ != outputChannels_)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: channel count must not change (%d -> %d)", name_.c_str(), inputChannels_, outputChannels_);
}
std::vector<BufferSpec> UnaryCopyLayer::getRequiredInputBuffers() const {
    const BufferSpec::order ord = kind_ == SHALLOW2DEEP ? BufferSpec::order::GPU_SHALLOW : (kind_ == DEEP2SHALLOW ? BufferSpec::order::GPU_DEEP : order());
    return {BufferSpec(0, width_, height_, inputChannels_, inputPadding_, ord, storagePrecision(), BufferSpec::FUNCTION_SOURCE).anyType()};
}
std::vector<BufferSpec> UnaryCopyLayer::getRequiredOutputBuffers() const {
    const BufferSpec::order ord = kind_ == SHALLOW2DEEP ? BufferSpec::order::GPU_DEEP : (kind_ == DEEP2SHALLOW ? BufferSpec::order::GPU_SHALLOW : order());
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, ord, storagePrecision(), BufferSpec::FUNCTION_DEST)};
}
void UnaryCopyLayer::setup() {
    if (kind_ == RGB2BGR) FYN_ABI_CALL(fyn_rgb2bgr_create(context_.handle(), &desc_, &op_));
    else FYN_ABI_CALL(fyn_relayout_create(context_.handle(), &desc_, &op_));
    valid_ = true;
}
void UnaryCopyLayer::cleanup() {
    if (op_) fyn_op_destroy(op_);
    op_ = nullptr;
    GPULayerBase::cleanup();
}
void UnaryCopyLayer::forward(uint64_t) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (kind_ == RGB2BGR) FYN_ABI_CALL(fyn_rgb2bgr_run(op_, in(0), out(), context_.stream()));
    else FYN_ABI_CALL(fyn_relayout_run(op_, in(0), out(), context_.stream()));
}

// ------------------------------------------------------------------------------------------------
// UploadLayer: host float32 [H][W][C] -> C-channel float32 texture, verbatim (gpu/uploadlayer.cpp:360-380)
// ------------------------------------------------------------------------------------------------
UploadLayer::UploadLayer(const UpDownLayerBuilder &b, int layerNumber)
    : GPULayerBase(b, layerNumber), async_(b.async_), dataType_(b.dataType_), callback_(b.callback_) {
    if (inputChannels_ > PIXEL_PACKING)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s: upload supports at most %d channels", name_.c_str(), PIXEL_PACKING);
}
std::vector<BufferSpec> UploadLayer::getRequiredInputBuffers() const {
    return {BufferSpec(0, width_, height_, inputChannels_, 0, BufferSpec::order::GPU_SHALLOW, dataType_ == BufferSpec::UBYTE ? BufferSpec::UBYTE : BufferSpec::FLOAT32,
                       BufferSpec::CPU_SOURCE)
                .device(BufferSpec::COMP_STOR_CPU)};
}
std::vector<BufferSpec> UploadLayer::getRequiredOutputBuffers() const {
    // RGB32F-style texture: packing == channel count, float32, no padding
    return {BufferSpec(0, width_, height_, outputChannels_, outputPadding_, order(), BufferSpec::FLOAT32, BufferSpec::GPU_DEST)
                .packing(outputPadding_ == 0 ? outputChannels_ : 4).async(async_).multi(async_ ? Engine::ASYNC_SLOTS : 1)};
}
void UploadLayer::forward(uint64_t sequence) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!input_) THROW_EXCEPTION_ARGS(FynException, "No input buffer set for upload layer %s", name_.c_str());
    if (callback_) callback_(sequence, input_, AsyncLayer::UPLOAD_COMMENCED);
    uploadFrom(input_, out(), context_.stream());
    if (callback_) callback_(sequence, input_, AsyncLayer::UPLOAD_DONE);
}

TensorHandle UploadLayer::asyncUpload(uint64_t sequence, int slot, void *stream) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!input_) THROW_EXCEPTION_ARGS(FynException, "No input buffer set for upload layer %s", name_.c_str());
    TensorHandle target = getOutputTexture(0, slot);
    if (!target) THROW_EXCEPTION_ARGS(FynException, "Upload layer %s has no output buffer %d", name_.c_str(), slot);
    // the copy reads the caller's pinned buffer asynchronously: UPLOAD_COMMENCED / UPLOAD_DONE are fired by
    // notifyUploaded() once it has completed, never before (a caller may refill its buffer on UPLOAD_COMMENCED)
    (void)sequence;
    pendingInput_ = input_;
    uploadFrom(input_, target, stream);
    return target;
}

// FLOAT32 buffers are copied verbatim into the RGB32F upload texture; UBYTE buffers (reference: gpu/uploadlayer.cpp:51-66,365-375,
// an 8-bit normalised texture) are converted to value / 255 on the device, so only a quarter of the bytes cross PCIe
void UploadLayer::uploadFrom(CPUBuffer *buffer, TensorHandle target, void *stream) {
    const bool bytes = buffer->shape().dataType() == cpu::CPUBufferShape::UINT8;
    if (bytes != (dataType_ == BufferSpec::UBYTE))
        THROW_EXCEPTION_ARGS(FynException, "Upload layer %s: input buffer data type does not match the layer's", name_.c_str());
    if (bytes) {
        const uint8_t *src = buffer->map<uint8_t>();
        buffer->unmap();
        FYN_ABI_CALL(fyn_upload_u8_async(target, src, stream));
    } else {
        const float *src = buffer->map<float>();
        buffer->unmap();
        FYN_ABI_CALL(fyn_upload_f32_async(target, src, stream));
    }
}

void UploadLayer::notifyUploaded(uint64_t sequence) {
    if (!callback_) return;
    callback_(sequence, pendingInput_, AsyncLayer::UPLOAD_COMMENCED);
    callback_(sequence, pendingInput_, AsyncLayer::UPLOAD_DONE);
}

// ------------------------------------------------------------------------------------------------
// DownloadLayer: tensor -> host float32 in texel order (gpu/downloadlayer.cpp:112-131,257-283)
// ------------------------------------------------------------------------------------------------
DownloadLayer::DownloadLayer(const UpDownLayerBuilder &b, int layerNumber)
    : GPULayerBase(b, layerNumber), async_(b.async_), dataType_(b.dataType_ == BufferSpec::UBYTE ? BufferSpec::UBYTE : BufferSpec::FLOAT32), callback_(b.callback_) {
    if (flags_ & LayerFlags::PRE_ACT_MASK) THROW_EXCEPTION_ARGS(FynException, "Activation on download not implemented yet");
    if (flags_ & LayerFlags::RESIDUAL_INPUT) THROW_EXCEPTION_ARGS(FynException, "Residual add on download not implemented yet");
}
std::vector<BufferSpec> DownloadLayer::getRequiredInputBuffers() const {
    return {BufferSpec(0, width_, height_, inputChannels_, inputPadding_, order(), storagePrecision(), BufferSpec::FUNCTION_SOURCE)};
}
std::vector<BufferSpec> DownloadLayer::getRequiredOutputBuffers() const {
    return {BufferSpec(0, width_, height_, outputChannels_, inputPadding_, order(), dataType_, BufferSpec::CPU_DEST)
                .device(BufferSpec::COMP_STOR_CPU)};
}
size_t DownloadLayer::hostBytes() const {
    return dataType_ == BufferSpec::UBYTE ? fyn_download_u8_bytes(in(0)) : fyn_download_f32_elems(in(0)) * sizeof(float);
}
void DownloadLayer::forward(uint64_t sequence) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    std::lock_guard<std::recursive_mutex> lck(processingLock_);
    if (!output_) THROW_EXCEPTION_ARGS(FynException, "No output buffer set for download layer %s", name_.c_str());
    size_t need = hostBytes();
    if (output_->bytes() < need)
        THROW_EXCEPTION_ARGS(FynException, "Download buffer too small (%zu < %zu bytes)", output_->bytes(), need);
    if (callback_) callback_(sequence, output_, AsyncLayer::DOWNLOAD_COMMENCED);
    if (dataType_ == BufferSpec::UBYTE) FYN_ABI_CALL(fyn_download_u8_async(in(0), static_cast<unsigned char *>(output_->raw()), context_.stream()));
    else FYN_ABI_CALL(fyn_download_f32_async(in(0), static_cast<float *>(output_->raw()), context_.stream()));
    if (!async_) {
        // synchronous path blocks like the reference's glReadPixels + readFromPBO
        FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
        output_->setSequence(sequence);
        if (callback_) callback_(sequence, output_, AsyncLayer::DOWNLOAD_DONE);
    }
}
CPUBuffer *DownloadLayer::asyncBuffer(int slot) {
    if (slot == 0) return output_;
    if (!asyncOutputs_[slot]) {
        if (!output_) THROW_EXCEPTION_ARGS(FynException, "No output buffer set for download layer %s", name_.c_str());
        asyncOutputs_[slot] = output_->shape().createBuffer(context_);   // second pinned buffer for the pipeline
    }
    return asyncOutputs_[slot];
}

void DownloadLayer::asyncConvert(int slot, void *stream) {
    if (!valid_) THROW_EXCEPTION_ARGS(FynException, "Trying to invoke forward() on invalid layer");
    if (!staging_[slot]) {
        void *p = nullptr;
        FYN_ABI_CALL(fyn_device_alloc(context_.handle(), hostBytes(), &p));
        staging_[slot] = static_cast<float *>(p);
    }
    if (dataType_ == BufferSpec::UBYTE) FYN_ABI_CALL(fyn_download_u8_convert(in(0), reinterpret_cast<unsigned char *>(staging_[slot]), stream));
    else FYN_ABI_CALL(fyn_download_convert(in(0), staging_[slot], stream));
}

CPUBuffer *DownloadLayer::asyncCopy(uint64_t sequence, int slot, void *stream) {
    CPUBuffer *buf = asyncBuffer(slot);
    const size_t bytes = hostBytes();
    if (buf->bytes() < bytes) THROW_EXCEPTION_ARGS(FynException, "Download buffer too small (%zu < %zu bytes)", buf->bytes(), bytes);
    if (callback_) callback_(sequence, buf, AsyncLayer::DOWNLOAD_COMMENCED);
    FYN_ABI_CALL(fyn_memcpy_async(context_.handle(), buf->raw(), staging_[slot], bytes, 1, stream));
    return buf;
}

void DownloadLayer::cleanup() {
    for (int i = 0; i < Engine::ASYNC_SLOTS; i++) {
        if (staging_[i]) fyn_device_free(context_.handle(), staging_[i]);
        staging_[i] = nullptr;
        delete asyncOutputs_[i];
        asyncOutputs_[i] = nullptr;
    }
    GPULayerBase::cleanup();
}

void DownloadLayer::writeResult(const char *fileName, bool) {
    if (!output_) return;
    CPUBuffer *cw = output_->toChannelWise();
    FILE *f = fopen(fileName, "wb");
    if (f) {
        fwrite(cw->raw(), 1, cw->bytes(), f);
        fclose(f);
    }
    delete cw;
}

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
