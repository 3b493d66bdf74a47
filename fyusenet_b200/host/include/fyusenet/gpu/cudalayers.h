// Concrete device layers of the CUDA backend and the factory backend that creates them.
// One class per reference layer family; all arithmetic happens in libfyusenet_b200.so behind the C ABI.
//   ConvLayer        <- vanilla::ConvLayer1x1 / ConvLayerNxN / FractionalConvLayerNxN (gpu/vanilla/*),
//                       deep::DeepConvLayer1x1 / DeepConvLayerNxN / DeepGEMMLayer (gpu/deep/*)
//   DepthwiseConvLayer <- vanilla::DepthwiseConvLayer3x3, deep::DeepDepthwiseConvLayer3x3
//   TransConvLayer   <- vanilla::TransConvLayer2x2 / TransConvLayer3x3 (stride 2, shallow)
//   PoolingLayer     <- deep::DeepMaxPoolLayer / DeepAvgPoolLayer, MaxPoolLayer / AvgPoolLayer
//   BatchNormLayer   <- BatchNormLayer, deep::DeepBatchNormLayer
//   SigmoidLayer     <- SigmoidLayer
//   ScaleLayer       <- ScaleLayer, deep::DeepScaleLayer (also the PADDING2D / RELU / CLIP pseudo-layers)
//   AddSubLayer, SingletonArithmeticLayer, ConcatLayer (deep::DeepConcatLayer), RGB2BGRLayer,
//   Shallow2DeepLayer / Deep2ShallowLayer
//   UploadLayer      <- UploadLayer (gpu/uploadlayer.cpp)
//   DownloadLayer    <- DownloadLayer, deep::DeepDownloadLayer (gpu/downloadlayer.cpp, deep/deepdownloadlayer.cpp)
#pragma once
#include "../base/batchnorminterface.h"
#include "../base/convlayerinterface.h"
#include "../base/engine.h"
#include "../base/layerfactory.h"
#include "../cpu/cpubuffer.h"
#include "gpulayerbase.h"

namespace fyusion {
namespace fyusenet {
namespace gpu {

class ConvLayerBase : public GPULayerBase, public ConvLayerInterface {
 public:
    ConvLayerBase(const ConvLayerBuilder &builder, int layerNumber, bool fractional);
    ConvLayerBase(const GPULayerBuilder &builder, int layerNumber);  // GEMM: 1x1 conv on 1x1 spatial
    ~ConvLayerBase() override;
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void loadWeightsAndBiases(const float *biasAndWeights, size_t offset = 0) override;
    int kernel() const { return desc_.kernel; }
    int backendFamily() const { return op_ ? fyn_conv2d_backend(op_) : 0; }  // 1 direct, 2 tcgen05
    const fyn_conv_desc &descriptor() const { return desc_; }
    // Engine-level layer fusion: evaluate the element-wise FunctionLayer that consumes this layer in the convolution's
    // epilogue and write straight into that layer's output tensor (the reference runs one render pass per layer,
    // gpu/functionlayer.cpp:145-179).  Returns false when the pair cannot be fused.
    bool fuseFunction(int function, TensorHandle target);
    void unfuse();
    bool fused() const { return fusedTarget_ != nullptr; }
    // Fusion of the stand-alone batch-norm layer that produces this layer's input (deep 1x1 convolutions): read `source`
    // (the batch-norm layer's input tensor) and apply scale / bias at the fetch.  Returns false when the kernel family
    // cannot do it (the batch-norm layer then simply runs).
    bool fuseInputNorm(const float *scaleAndBias, TensorHandle source);
    void unfuseInput();
    bool inputFused() const { return fusedInput_ != nullptr; }
    // Chain fusion (Engine::updateFusion, fyn_conv_chain): the first layer of a run of convolutions of identical geometry
    // launches one persistent kernel for the whole run and writes the last layer's output tensor; the other layers of
    // the run do nothing in forward().  If the chain declines the tensors at run time, the layers run one by one.
    fyn_op *op() const { return op_; }
    TensorHandle residualTexture() const { return residuals_.empty() ? nullptr : residuals_[0]; }
    void setChainHead(fyn_conv_chain *chain, const std::vector<ConvLayerBase *> &followers);
    void setChainMember(bool on) { chainMember_ = on; }
    void unchain();
    bool chained() const { return chain_ != nullptr || chainMember_; }
    int chainLength() const { return chain_ ? 1 + (int)chainFollowers_.size() : 0; }   // > 0: this layer launches the chain

 protected:
    void init(int kernel, int dilation, float sourceStep, bool fractional);
    void forwardSingle();
    fyn_conv_chain *chain_ = nullptr;                 // owned by the engine
    std::vector<ConvLayerBase *> chainFollowers_;
    bool chainMember_ = false;
    int fusedFunction_ = 0;
    TensorHandle fusedTarget_ = nullptr;
    TensorHandle fusedInput_ = nullptr;
    fyn_conv_desc desc_{};
    fyn_op *op_ = nullptr;
    std::vector<float> pendingWeights_;  // weights handed over before setup()
    int outWidth_ = 0, outHeight_ = 0;
};

class DepthwiseConvLayer;
class TransConvLayer;
namespace vanilla {
using ConvLayerNxN = gpu::ConvLayerBase;
using ConvLayer1x1 = gpu::ConvLayerBase;
using FractionalConvLayerNxN = gpu::ConvLayerBase;
using DepthwiseConvLayer3x3 = gpu::DepthwiseConvLayer;
using TransConvLayer2x2 = gpu::TransConvLayer;
using TransConvLayer3x3 = gpu::TransConvLayer;
}  // namespace vanilla

// vanilla::DepthwiseConvLayer3x3 (gpu/vanilla/convlayer_dw_3x3_vanilla.cpp) / deep::DeepDepthwiseConvLayer3x3
// (gpu/deep/deepdwconvlayer3x3.cpp): selected by the factory when groupSize == input channels (gpulayerfactory.cpp:358-396)
class DepthwiseConvLayer : public GPULayerBase, public ConvLayerInterface {
 public:
    DepthwiseConvLayer(const ConvLayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void loadWeightsAndBiases(const float *biasAndWeights, size_t offset = 0) override;

 protected:
    fyn_dwconv_desc desc_{};
    fyn_op *op_ = nullptr;
    std::vector<float> pendingWeights_;
    int outWidth_ = 0, outHeight_ = 0;
};

// vanilla::TransConvLayer2x2 / TransConvLayer3x3 (gpu/vanilla/transconvlayerbase_vanilla.cpp): stride-2 transpose convolution
class TransConvLayer : public GPULayerBase, public ConvLayerInterface {
 public:
    TransConvLayer(const ConvLayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void loadWeightsAndBiases(const float *biasAndWeights, size_t offset = 0) override;

 protected:
    fyn_transconv_desc desc_{};
    fyn_op *op_ = nullptr;
    std::vector<float> pendingWeights_;
};

class PoolingLayer : public GPULayerBase {
 public:
    PoolingLayer(const PoolLayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;

 protected:
    fyn_pool_desc desc_{};
    fyn_op *op_ = nullptr;
    int outWidth_ = 0, outHeight_ = 0;
};

class BatchNormLayer : public GPULayerBase, public BatchNormInterface {
 public:
    BatchNormLayer(const GPULayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void loadScaleAndBias(const float *scaleAndBias, size_t sbOffset = 0) override;
    // fusion support (Engine::updateFusion): a batch-norm layer without prefix activation whose only consumer is a
    // convolution can be evaluated at that convolution's fetch; a bypassed layer does nothing in forward()
    bool plainFunction() const { return (flags_ & (LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP)) == 0; }
    const std::vector<float> &parameters() const { return params_; }
    void setBypass(ConvLayerBase *consumer) { fusedConsumer_ = consumer; }
    bool bypassed() const { return fusedConsumer_ != nullptr; }

 protected:
    ConvLayerBase *fusedConsumer_ = nullptr;
    fyn_bn_desc desc_{};
    fyn_op *op_ = nullptr;
    std::vector<float> params_;
};

class SigmoidLayer : public GPULayerBase {
 public:
    SigmoidLayer(const GPULayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    // fusion support: a sigmoid without prefix activation can be evaluated by its producer (ConvLayerBase::fuseFunction);
    // a bypassed layer does nothing in forward(), its output tensor is written by the producer
    bool plainFunction() const { return (flags_ & (LayerFlags::PRE_RELU | LayerFlags::PRE_CLIP)) == 0; }
    void setBypass(bool on) { bypass_ = on; }
    bool bypassed() const { return bypass_; }

 protected:
    fyn_unary_desc desc_{};
    fyn_op *op_ = nullptr;
    bool bypass_ = false;
};

// ScaleLayer / deep::DeepScaleLayer (gpu/scalelayer.cpp:40-75, gpu/deep/deepscalelayer.cpp:30-45); constructed from a plain
// GPULayerBuilder it is the identity-size copy that implements PADDING2D / RELU / CLIP (gpu/gpulayerfactory.cpp:125-140)
class ScaleLayer : public GPULayerBase {
 public:
    ScaleLayer(const ScaleLayerBuilder &builder, int layerNumber);
    ScaleLayer(const GPULayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;

 protected:
    void init(int upx, int upy, int dnx, int dny, ScalingType type);
    fyn_scale_desc desc_{};
    fyn_op *op_ = nullptr;
    int outWidth_ = 0, outHeight_ = 0;
};

// AddSubLayer (gpu/addsublayer.cpp) and SingletonArithmeticLayer (gpu/singleton_arithlayer.cpp)
class ArithLayer : public GPULayerBase {
 public:
    ArithLayer(const GPULayerBuilder &builder, int layerNumber);                 // LayerType::ADD / SUB: ports 0 and 1
    ArithLayer(const SingletonArithLayerBuilder &builder, int layerNumber);     // port 0 (op) operand
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    int numInputPorts() const override { return desc_.singleton ? 1 : 2; }

 protected:
    void init();
    fyn_arith_desc desc_{};
    fyn_op *op_ = nullptr;
};
using AddSubLayer = ArithLayer;
using SingletonArithmeticLayer = ArithLayer;

// ConcatLayer / deep::DeepConcatLayer (gpu/concatlayer.cpp:60-110): one input port per concatenated tensor
class ConcatLayer : public GPULayerBase {
 public:
    ConcatLayer(const ConcatLayerBuilder &builder, int layerNumber);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    int numInputPorts() const override { return desc_.num_inputs; }
    int numInputChannels(int port = 0) const override;

 protected:
    fyn_concat_desc desc_{};
    fyn_op *op_ = nullptr;
};

// RGB2BGRLayer (gpu/rgb2bgrlayer.cpp), Shallow2DeepLayer (gpu/shallow2deep.cpp), Deep2ShallowLayer (gpu/deep2shallow.cpp)
class UnaryCopyLayer : public GPULayerBase {
 public:
    enum Kind { RGB2BGR, SHALLOW2DEEP, DEEP2SHALLOW };
    UnaryCopyLayer(const GPULayerBuilder &builder, int layerNumber, Kind kind);
    void setup() override;
    void cleanup() override;
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;

 protected:
    Kind kind_;
    fyn_unary_desc desc_{};
    fyn_op *op_ = nullptr;
};
using RGB2BGRLayer = UnaryCopyLayer;
using Shallow2DeepLayer = UnaryCopyLayer;
using Deep2ShallowLayer = UnaryCopyLayer;

class UploadLayer : public GPULayerBase, public cpu::CPULayerInterface {
 public:
    UploadLayer(const UpDownLayerBuilder &builder, int layerNumber);
    void setup() override { valid_ = true; }
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void setInputBuffer(CPUBuffer *buf, int port) override { (void)port; input_ = buf; }
    CPUBuffer *getInputBuffer(int = 0) const override { return input_; }
    void addOutputBuffer(CPUBuffer *, int = 0) override { THROW_EXCEPTION_ARGS(FynException, "Upload layers have no CPU output"); }
    CPUBuffer *getOutputBuffer(int = 0) const override { return nullptr; }
    bool hasOutputBuffer(int = 0) const override { return false; }
    void clearOutputBuffers(int = -1) override {}
    void clearInputBuffers(int = -1) override { input_ = nullptr; }
    bool isAsync() const { return async_; }
    // asynchronous path (reference: gpu/uploadlayer.cpp:395-541): copy the current input buffer into output buffer
    // `slot` on `stream`; returns the tensor that now carries the sequence
    TensorHandle asyncUpload(uint64_t sequence, int slot, void *stream);
    // Fires UPLOAD_COMMENCED ("input buffer may be changed") and UPLOAD_DONE once the host->device copy of `sequence` has
    // completed; the asynchronous engine calls it from a host function queued behind the copy (reference contract:
    // gpu/uploadlayer.cpp:395-541, AsyncLayer::state)
    void notifyUploaded(uint64_t sequence);
    bool hasCallback() const { return (bool)callback_; }
    BufferSpec::dtype dataType() const { return dataType_; }

 protected:
    void uploadFrom(CPUBuffer *buffer, TensorHandle target, void *stream);
    CPUBuffer *input_ = nullptr;
    CPUBuffer *pendingInput_ = nullptr;     // buffer of the most recent asynchronous upload (callback argument)
    bool async_ = false;
    BufferSpec::dtype dataType_ = BufferSpec::FLOAT32;
    UpDownLayerBuilder::callback_t callback_;
};

// (dataType(UBYTE) on a download layer is an extension: 8-bit RGBA texels, (uint8)(clamp(v, 0, 1) * 255) on the device)
class DownloadLayer : public GPULayerBase, public cpu::CPULayerInterface {
 public:
    DownloadLayer(const UpDownLayerBuilder &builder, int layerNumber);
    void setup() override { valid_ = true; }
    void forward(uint64_t sequence = 0) override;
    std::vector<BufferSpec> getRequiredInputBuffers() const override;
    std::vector<BufferSpec> getRequiredOutputBuffers() const override;
    void setInputBuffer(CPUBuffer *, int) override { THROW_EXCEPTION_ARGS(FynException, "Download layers have no CPU input"); }
    CPUBuffer *getInputBuffer(int = 0) const override { return nullptr; }
    void addOutputBuffer(CPUBuffer *buf, int = 0) override { output_ = buf; }
    void updateOutputBuffer(CPUBuffer *buf, int = 0) { output_ = buf; }
    // DOWNLOAD_DONE of the asynchronous path, fired by the engine's completion function once the copy has landed
    void notifyDownloaded(uint64_t sequence, CPUBuffer *buf) { if (callback_) callback_(sequence, buf, AsyncLayer::DOWNLOAD_DONE); }
    CPUBuffer *getOutputBuffer(int = 0) const override { return output_; }
    bool hasOutputBuffer(int = 0) const override { return output_ != nullptr; }
    void clearOutputBuffers(int = -1) override { output_ = nullptr; }
    void clearInputBuffers(int = -1) override {}
    void writeResult(const char *fileName, bool includePadding = false) override;
    bool isAsync() const { return async_; }
    // the engine synchronises the stream after the last layer unless the download is asynchronous
    bool needsSync() const { return !async_; }
    // asynchronous path (reference: gpu/downloadlayer.cpp:139-157,307-323), split in its device and host halves:
    // convert the input tensor into device staging buffer `slot` (compute stream), then copy staging -> host buffer
    // `slot` (download stream).  Both staging and the second host buffer are created on first use.
    void asyncConvert(int slot, void *stream);
    CPUBuffer *asyncCopy(uint64_t sequence, int slot, void *stream);
    CPUBuffer *asyncBuffer(int slot);
    void cleanup() override;

 protected:
    CPUBuffer *output_ = nullptr;
    CPUBuffer *asyncOutputs_[Engine::ASYNC_SLOTS] = {};
    float *staging_[Engine::ASYNC_SLOTS] = {};
    bool async_ = false;
    BufferSpec::dtype dataType_ = BufferSpec::FLOAT32;
    size_t hostBytes() const;          // size of one downloaded frame in the layer's host data type
    UpDownLayerBuilder::callback_t callback_;
};

namespace deep {
using DeepDownloadLayer = gpu::DownloadLayer;
using DeepConvLayer1x1 = gpu::ConvLayerBase;
using DeepConvLayerNxN = gpu::ConvLayerBase;
using DeepGEMMLayer = gpu::ConvLayerBase;
using DeepMaxPoolLayer = gpu::PoolingLayer;
using DeepAvgPoolLayer = gpu::PoolingLayer;
using DeepBatchNormLayer = gpu::BatchNormLayer;
using DeepDepthwiseConvLayer3x3 = gpu::DepthwiseConvLayer;
using DeepScaleLayer = gpu::ScaleLayer;
using DeepConcatLayer = gpu::ConcatLayer;
}  // namespace deep

// The plugin: creates CUDA layers behind LayerFactoryBackend::createLayer.
// Dispatch restated from GPULayerFactoryBackend::createLayer (gpu/gpulayerfactory.cpp:112-184,358-396,447-457).
class CUDALayerFactoryBackend : public LayerFactoryBackend {
 public:
    explicit CUDALayerFactoryBackend(GfxContextLink ctx = GfxContextLink()) : context_(ctx) {}
    std::string getName() const override { return "CUDA-sm100a"; }
    LayerBase *createLayer(LayerType type, LayerBuilder *builder, int layerNumber) override;

 private:
    GfxContextLink context_;
};

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
