// Fluent layer builders.  Same public surface as the reference's LayerBuilderTempl
// (fyusenet/base/layerbuilder.h:53-558): shape/size/downsample/upsample/*Padding/prefixAct/postfixNorm/
// residual/deep/leakyReLU/clip/number/type/push and getFlags().  The parameter bag lives in a plain
// struct so factories and layers can read it without knowing the concrete builder type.
#pragma once
#include <cassert>
#include <memory>
#include <string>

#include "../common/fynexception.h"
#include "layerfactoryinterface.h"
#include "layerflags.h"

namespace fyusion {
namespace fyusenet {

class LayerFactory;

// Everything a builder collects.  Field names follow the reference so that ported network
// definitions (and code that pokes at builder members) keep compiling.
struct LayerBuilderData {
    explicit LayerBuilderData(const std::string &name) : name_(name) {}
    virtual ~LayerBuilderData() = default;

    virtual short width() const { return (short)width_; }
    virtual short height() const { return (short)height_; }
    virtual short in() const { return (short)inputChannels_; }
    virtual short out() const { return (short)outputChannels_; }
    bool isDeep() const { return (flags_ & LayerFlags::DEEP) != 0; }

    // Builder state -> layer flag word (reference: layerbuilder.h:445-493).
    layerflags getFlags() const {
        layerflags f = flags_;
        if (preAct_ == ActType::RELU || preAct_ == ActType::LEAKY_RELU) f |= LayerFlags::PRE_RELU;
        else if (preAct_ == ActType::CLIP) f |= LayerFlags::PRE_CLIP;
        else if (preAct_ != ActType::NONE) THROW_EXCEPTION_ARGS(FynException, "Activation type not supported yet");
        if (postAct_ == ActType::RELU || postAct_ == ActType::LEAKY_RELU) f |= LayerFlags::POST_RELU;
        else if (postAct_ != ActType::NONE) THROW_EXCEPTION_ARGS(FynException, "Activation type not supported yet");
        if (postNorm_ == NormType::BATCHNORM) f |= LayerFlags::POST_BATCHNORM;
        if (resAct_ == ActType::RELU) f |= LayerFlags::RELU_ON_RESIDUAL;
        if (residualNorm_) f |= LayerFlags::BATCHNORM_ON_RESIDUAL;
        return f;
    }

    std::string name_;
    short inputPadding_ = 0, outputPadding_ = 0, residualPadding_ = 0;
    short downsample_[2] = {1, 1};
    short upsample_[2] = {1, 1};
    ActType preAct_ = ActType::NONE, postAct_ = ActType::NONE, resAct_ = ActType::NONE;
    NormType postNorm_ = NormType::NONE;
    float leakyReLU_ = 0.0f, clipLow_ = 0.0f, clipHigh_ = 0.0f;
    int number_ = -1;
    LayerType type_ = LayerType::ILLEGAL;
    bool residualNorm_ = false;
    compute_device device_ = compute_device::DEV_CPU;
    uint16_t width_ = 0, height_ = 0, inputChannels_ = 0, outputChannels_ = 0;
    layerflags flags_ = LayerFlags::NO_LAYER_FLAGS;
};

struct LayerBuilder;
struct BuilderLeaf {};

#define FYN_FLUENT(signature, body) \
    D &signature {                  \
        body;                       \
        return *static_cast<D *>(this); \
    }

template <typename D = BuilderLeaf>
struct LayerBuilderTempl : LayerBuilderData {
    explicit LayerBuilderTempl(const std::string &name) : LayerBuilderData(name) {}

    // hands the builder to the factory, which takes ownership (reference: layerbuilder.h:87-91)
    void push(std::shared_ptr<LayerFactory> &factory);

    FYN_FLUENT(type(LayerType t), type_ = t)
    FYN_FLUENT(number(int no), assert(no >= 0); number_ = no)
    FYN_FLUENT(size(short w, short h), width_ = w; height_ = h)
    FYN_FLUENT(downsample(int ds), downsample_[0] = downsample_[1] = (short)ds)
    FYN_FLUENT(downsample(int horizontal, int vertical), downsample_[0] = (short)horizontal; downsample_[1] = (short)vertical)
    FYN_FLUENT(upsample(short us), upsample_[0] = upsample_[1] = us)
    FYN_FLUENT(upsample(short horizontal, short vertical), upsample_[0] = horizontal; upsample_[1] = vertical)
    FYN_FLUENT(inputPadding(short p), inputPadding_ = p)
    FYN_FLUENT(outputPadding(short p), outputPadding_ = p)
    FYN_FLUENT(residualPadding(short p), residualPadding_ = p)
    FYN_FLUENT(prefixAct(ActType a), preAct_ = a)
    FYN_FLUENT(postfixAct(ActType a), postAct_ = a)
    FYN_FLUENT(postfixNorm(NormType n), postNorm_ = n)
    FYN_FLUENT(deep(), flags_ |= LayerFlags::DEEP)
    FYN_FLUENT(shape(int outChannels, int h, int w, int inChannels),
               width_ = (uint16_t)w; height_ = (uint16_t)h; inputChannels_ = (uint16_t)inChannels; outputChannels_ = (uint16_t)outChannels)
    FYN_FLUENT(shape(int h, int w, int chans),
               width_ = (uint16_t)w; height_ = (uint16_t)h; inputChannels_ = outputChannels_ = (uint16_t)chans)
    FYN_FLUENT(channels(short c), inputChannels_ = outputChannels_ = (uint16_t)c)
    FYN_FLUENT(inChannels(short c), inputChannels_ = (uint16_t)c)
    FYN_FLUENT(outChannels(short c), outputChannels_ = (uint16_t)c)
    FYN_FLUENT(leakyReLU(float leak), leakyReLU_ = leak)
    FYN_FLUENT(clip(float low, float high), clipLow_ = low; clipHigh_ = high)

    // residual input; only NONE / RELU may be applied to it (reference: layerbuilder.h:290-297)
    D &residual(ActType act = ActType::NONE, bool postfixNorm = false) {
        if (act != ActType::RELU && act != ActType::NONE)
            THROW_EXCEPTION_ARGS(FynException, "Activation type %d not supported on residual", (int)act);
        flags_ |= LayerFlags::RESIDUAL_INPUT;
        if (act == ActType::RELU) flags_ |= LayerFlags::RELU_ON_RESIDUAL;
        else flags_ &= ~LayerFlags::RELU_ON_RESIDUAL;
        residualNorm_ = postfixNorm;
        return *static_cast<D *>(this);
    }
};

struct LayerBuilder : LayerBuilderTempl<LayerBuilder> {
    explicit LayerBuilder(const std::string &name) : LayerBuilderTempl<LayerBuilder>(name) {}
};

}  // namespace fyusenet
}  // namespace fyusion
