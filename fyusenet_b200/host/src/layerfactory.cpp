// LayerFactory and the CUDA factory backend.
#include "fyusenet/base/layerfactory.h"

#include "fyusenet/gpu/cudalayers.h"

namespace fyusion {
namespace fyusenet {

LayerFactory::~LayerFactory() {
    // the factory owns pushed builders and its backend (reference: layerfactory.cpp:41-52)
    for (auto &kv : builders_) delete kv.second;
    builders_.clear();
    delete backend_;
}

void LayerFactory::pushBuilder(LayerBuilder *builder) {
    // every builder starts with a LayerBuilderData sub-object at offset 0 (single inheritance chain)
    LayerBuilderData *data = reinterpret_cast<LayerBuilderData *>(builder);
    if (!data) THROW_EXCEPTION_ARGS(FynException, "Null builder pushed");
    if (data->number_ < 0) THROW_EXCEPTION_ARGS(FynException, "Builder %s has no layer number", data->name_.c_str());
    if (builders_.count(data->number_))
        THROW_EXCEPTION_ARGS(FynException, "Layer number %d (%s) already in use", data->number_, data->name_.c_str());
    if (data->type_ == LayerType::ILLEGAL) THROW_EXCEPTION_ARGS(FynException, "Builder %s has no layer type", data->name_.c_str());
    builders_[data->number_] = data;
}

CompiledLayers LayerFactory::compileLayers() {
    CompiledLayers result;
    for (auto &kv : builders_) {
        LayerBuilderData *b = kv.second;
        if (b->device_ != compute_device::DEV_GPU)
            THROW_EXCEPTION_ARGS(FynException, "Layer %s: only GPU layers are supported by this backend (no CPU fallback)", b->name_.c_str());
        LayerBase *layer = backend_->createLayer(b->type_, reinterpret_cast<LayerBuilder *>(b), b->number_);
        if (!layer) THROW_EXCEPTION_ARGS(FynException, "Cannot create layer %s (type %d)", b->name_.c_str(), (int)b->type_);
        result.setLayer(layer);
    }
    layers_ = result;
    return result;
}

LayerFactoryBackend *LayerFactory::GPUFactoryType::createBackend() { return new gpu::CUDALayerFactoryBackend(gfxContext); }

namespace gpu {

template <typename T>
static const T &as(LayerBuilderData *data, const char *what) {
    T *b = dynamic_cast<T *>(data);
    if (!b) THROW_EXCEPTION_ARGS(FynException, "Layer %s needs a %s", data->name_.c_str(), what);
    return *b;
}

LayerBase *CUDALayerFactoryBackend::createLayer(LayerType type, LayerBuilder *builder, int layerNumber) {
    LayerBuilderData *data = reinterpret_cast<LayerBuilderData *>(builder);
    switch (type) {
        case LayerType::CONVOLUTION2D: {
            const ConvLayerBuilder &cb = as<ConvLayerBuilder>(data, "ConvLayerBuilder");
            // depthwise when the group size equals the input channel count (gpu/gpulayerfactory.cpp:358-396)
            if (cb.groupSize_ != 1) {
                if (cb.groupSize_ == (short)data->inputChannels_ && cb.kernel_ == 3) return new DepthwiseConvLayer(cb, layerNumber);
                THROW_EXCEPTION_ARGS(FynException, "Layer %s: grouped convolution (group size %d, kernel %d) is not supported", data->name_.c_str(),
                                     (int)cb.groupSize_, (int)cb.kernel_);
            }
            return new ConvLayerBase(cb, layerNumber, false);
        }
        case LayerType::TRANSCONVOLUTION2D:
            return new TransConvLayer(as<ConvLayerBuilder>(data, "ConvLayerBuilder"), layerNumber);
        case LayerType::FRACCONVOLUTION2D: {
            const ConvLayerBuilder &cb = as<ConvLayerBuilder>(data, "ConvLayerBuilder");
            if (cb.isDeep()) THROW_EXCEPTION_ARGS(FynException, "Layer %s: fractional convolution has no deep variant", data->name_.c_str());
            return new ConvLayerBase(cb, layerNumber, true);
        }
        case LayerType::MAXPOOL2D:
        case LayerType::AVGPOOL2D:
            return new PoolingLayer(as<PoolLayerBuilder>(data, "PoolLayerBuilder"), layerNumber);
        case LayerType::BATCHNORM:
            return new BatchNormLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber);
        case LayerType::SIGMOID:
            return new SigmoidLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber);
        case LayerType::GEMM:
            return new ConvLayerBase(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber);
        // the reference emulates padding / ReLU / clip with a scaling layer (gpu/gpulayerfactory.cpp:125-140,317-322)
        case LayerType::PADDING2D:
            return new ScaleLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber);
        case LayerType::RELU:
        case LayerType::CLIP:
        case LayerType::SCALE2D: {
            if (ScaleLayerBuilder *sb = dynamic_cast<ScaleLayerBuilder *>(data)) {
                if (type == LayerType::RELU) sb->prefixAct(ActType::RELU);
                if (type == LayerType::CLIP) sb->prefixAct(ActType::CLIP);
                return new ScaleLayer(*sb, layerNumber);
            }
            GPULayerBuilder &gb = const_cast<GPULayerBuilder &>(as<GPULayerBuilder>(data, "ScaleLayerBuilder or GPULayerBuilder"));
            if (type == LayerType::RELU) gb.prefixAct(ActType::RELU);
            if (type == LayerType::CLIP) gb.prefixAct(ActType::CLIP);
            return new ScaleLayer(gb, layerNumber);
        }
        case LayerType::ADD:
        case LayerType::SUB:
            return new ArithLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber);
        case LayerType::SINGLETON_ARITH:
            return new ArithLayer(as<SingletonArithLayerBuilder>(data, "SingletonArithLayerBuilder"), layerNumber);
        case LayerType::CONCAT:
            return new ConcatLayer(as<ConcatLayerBuilder>(data, "ConcatLayerBuilder"), layerNumber);
        case LayerType::RGB2BGR:
            return new UnaryCopyLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber, UnaryCopyLayer::RGB2BGR);
        case LayerType::SHALLOW2DEEP:
            return new UnaryCopyLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber, UnaryCopyLayer::SHALLOW2DEEP);
        case LayerType::DEEP2SHALLOW:
            return new UnaryCopyLayer(as<GPULayerBuilder>(data, "GPULayerBuilder"), layerNumber, UnaryCopyLayer::DEEP2SHALLOW);
        case LayerType::UPLOAD:
            return new UploadLayer(as<UpDownLayerBuilder>(data, "UpDownLayerBuilder"), layerNumber);
        case LayerType::DOWNLOAD:
            return new DownloadLayer(as<UpDownLayerBuilder>(data, "UpDownLayerBuilder"), layerNumber);
        default:
            THROW_EXCEPTION_ARGS(FynException, "Layer %s: layer type %d is outside the CUDA backend's hot path", data->name_.c_str(), (int)type);
    }
    return nullptr;
}

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
