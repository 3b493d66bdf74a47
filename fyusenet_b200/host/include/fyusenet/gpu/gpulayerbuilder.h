// GPU layer builders: GPULayerBuilder, ConvLayerBuilder, PoolLayerBuilder, UpDownLayerBuilder, ScaleLayerBuilder,
// ConcatLayerBuilder, SingletonArithLayerBuilder.
// Reference: fyusenet/gpu/gpulayerbuilder.h, convlayerbuilder.h, poollayerbuilder.h, updownlayerbuilder.h,
// scalelayerbuilder.h, concatlayerbuilder.h, singleton_arithlayerbuilder.h.
#pragma once
#include <functional>
#include <cmath>
#include <string>
#include <vector>

#include "../base/bufferspec.h"
#include "../base/layerbuilder.h"
#include "gfxcontextlink.h"

namespace fyusion {
namespace fyusenet {
namespace cpu { class CPUBuffer; }

// upload / download progress states handed to asynchronous callbacks (reference: base/asynclayerinterface.h)
struct AsyncLayer {
    enum state { UPLOAD_COMMENCED = 0, UPLOAD_DONE, DOWNLOAD_COMMENCED, DOWNLOAD_DONE, ERROR };
};

namespace gpu {

template <typename D = LayerBuilderTempl<>>
struct GPULayerBuilderTempl : LayerBuilderTempl<D> {
    explicit GPULayerBuilderTempl(const std::string &name) : LayerBuilderTempl<D>(name) {
        LayerBuilderData::device_ = compute_device::DEV_GPU;
    }
    FYN_FLUENT(context(GfxContextLink ctx), context_ = ctx)
    GfxContextLink context_;
};

struct GPULayerBuilder : GPULayerBuilderTempl<GPULayerBuilder> {
    explicit GPULayerBuilder(const std::string &name) : GPULayerBuilderTempl<GPULayerBuilder>(name) {}
};

// kernel size, dilation, group size, fractional source step (reference: gpu/convlayerbuilder.h:36-100)
template <typename D = LayerBuilderTempl<>>
struct ConvLayerBuilderTempl : GPULayerBuilderTempl<D> {
    ConvLayerBuilderTempl(short kernel, const std::string &name) : GPULayerBuilderTempl<D>(name), kernel_(kernel) {}
    FYN_FLUENT(dilation(short dilate), dilation_[0] = dilation_[1] = dilate)
    FYN_FLUENT(dilation(short horizontal, short vertical), dilation_[0] = horizontal; dilation_[1] = vertical)
    FYN_FLUENT(sourceStep(float step), sourceStep_ = step)
    FYN_FLUENT(groupSize(short gs), groupSize_ = gs)
    short kernel_ = 1;
    short dilation_[2] = {1, 1};
    short groupSize_ = 1;
    float sourceStep_ = 1.f;
};

struct ConvLayerBuilder : ConvLayerBuilderTempl<ConvLayerBuilder> {
    ConvLayerBuilder(short kernel, const std::string &name) : ConvLayerBuilderTempl<ConvLayerBuilder>(kernel, name) {}
};

// pooling window / global flag (reference: gpu/poollayerbuilder.h:36-125)
template <typename D = LayerBuilderTempl<>>
struct PoolLayerBuilderTempl : GPULayerBuilderTempl<D> {
    enum op { POOL_AVG = 0, POOL_MAX };
    PoolLayerBuilderTempl(op operation, const std::string &name) : GPULayerBuilderTempl<D>(name), operation_(operation) {}
    FYN_FLUENT(poolSize(short win), poolsize_[0] = poolsize_[1] = win)
    FYN_FLUENT(poolSize(short winx, short winy), poolsize_[0] = winx; poolsize_[1] = winy)
    D &global() {
        if (LayerBuilderData::width_ == 0 || LayerBuilderData::height_ == 0)
            THROW_EXCEPTION_ARGS(FynException, "Must set size before specifying global pooling");
        LayerBuilderData::downsample_[0] = (short)LayerBuilderData::width_;
        LayerBuilderData::downsample_[1] = (short)LayerBuilderData::height_;
        global_ = true;
        return *static_cast<D *>(this);
    }
    op operation_;
    short poolsize_[2] = {1, 1};
    bool global_ = false;
};

struct PoolLayerBuilder : PoolLayerBuilderTempl<PoolLayerBuilder> {
    PoolLayerBuilder(op operation, const std::string &name) : PoolLayerBuilderTempl<PoolLayerBuilder>(operation, name) {}
};

// upload / download (reference: gpu/updownlayerbuilder.h:36-110)
template <typename D = LayerBuilderTempl<>>
struct UpDownLayerBuilderTempl : GPULayerBuilderTempl<D> {
    enum dir { UPLOAD = 0, DOWNLOAD };
    using callback_t = std::function<void(uint64_t, cpu::CPUBuffer *, AsyncLayer::state)>;
    UpDownLayerBuilderTempl(dir direction, const std::string &name) : GPULayerBuilderTempl<D>(name), direction_(direction) {
        LayerBuilderData::type_ = (direction == UPLOAD) ? LayerType::UPLOAD : LayerType::DOWNLOAD;
    }
    FYN_FLUENT(async(), async_ = true)
    FYN_FLUENT(dataType(BufferSpec::dtype dt), dataType_ = dt)
    FYN_FLUENT(callback(callback_t cb), callback_ = cb)
    dir direction_;
    bool async_ = false;
    callback_t callback_;
    BufferSpec::dtype dataType_ = BufferSpec::FLOAT;
};

struct UpDownLayerBuilder : UpDownLayerBuilderTempl<UpDownLayerBuilder> {
    UpDownLayerBuilder(dir direction, const std::string &name) : UpDownLayerBuilderTempl<UpDownLayerBuilder>(direction, name) {}
};

// scaling type, integer up / down factors, rotation (reference: gpu/scalelayerbuilder.h:36-120)
template <typename D = LayerBuilderTempl<>>
struct ScaleLayerBuilderTempl : GPULayerBuilderTempl<D> {
    explicit ScaleLayerBuilderTempl(const std::string &name) : GPULayerBuilderTempl<D>(name) { LayerBuilderData::type_ = LayerType::SCALE2D; }
    FYN_FLUENT(scaleType(ScalingType typ), scaleType_ = typ)
    FYN_FLUENT(rotate(int angle), rotation_ = angle)
    D &scale(float sc) { return scale(sc, sc); }
    D &scale(float scaleX, float scaleY) {
        setFactor(scaleX, 0);
        setFactor(scaleY, 1);
        return *static_cast<D *>(this);
    }
    bool equal() const {
        return LayerBuilderData::upsample_[0] == LayerBuilderData::upsample_[1] && LayerBuilderData::downsample_[0] == LayerBuilderData::downsample_[1];
    }
    ScalingType scaleType_ = ScalingType::NEAREST;
    int rotation_ = 0;

 private:
    void setFactor(float sc, int axis) {
        if (sc > 1.0f) {
            LayerBuilderData::upsample_[axis] = (short)sc;
            if (std::fabs((float)LayerBuilderData::upsample_[axis] - sc) > 1e-4f) THROW_EXCEPTION_ARGS(FynException, "Only supporting integer upscales for now");
        } else if (sc < 1.0f) {
            const float dn = 1.0f / sc;
            LayerBuilderData::downsample_[axis] = (short)(dn + 1e-4f);
            if (std::fabs((float)LayerBuilderData::downsample_[axis] - dn) > 1e-3f) THROW_EXCEPTION_ARGS(FynException, "Only supporting integer downscales for now");
        }
    }
};

struct ScaleLayerBuilder : ScaleLayerBuilderTempl<ScaleLayerBuilder> {
    explicit ScaleLayerBuilder(const std::string &name) : ScaleLayerBuilderTempl<ScaleLayerBuilder>(name) {}
};

// one entry per concatenated input (reference: gpu/concatlayerbuilder.h:36-75)
template <typename D = LayerBuilderTempl<>>
struct ConcatLayerBuilderTempl : GPULayerBuilderTempl<D> {
    struct Input {
        Input(short chan, short pad, int fl) : channels(chan), padding(pad), flags((layerflags)fl) {}
        short channels;
        short padding;
        layerflags flags;
    };
    explicit ConcatLayerBuilderTempl(const std::string &name) : GPULayerBuilderTempl<D>(name) { LayerBuilderData::type_ = LayerType::CONCAT; }
    D &input(short channels, short padding, int flags = LayerFlags::NO_LAYER_FLAGS) {
        inputs_.push_back(Input(channels, padding, flags));
        LayerBuilderData::inputChannels_ = (uint16_t)(LayerBuilderData::inputChannels_ + channels);
        return *static_cast<D *>(this);
    }
    std::vector<Input> inputs_;
};

struct ConcatLayerBuilder : ConcatLayerBuilderTempl<ConcatLayerBuilder> {
    explicit ConcatLayerBuilder(const std::string &name) : ConcatLayerBuilderTempl<ConcatLayerBuilder>(name) {}
};

// tensor (op) scalar (reference: gpu/singleton_arithlayerbuilder.h:36-70)
template <typename D = LayerBuilderTempl<>>
struct SingletonArithLayerBuilderTempl : GPULayerBuilderTempl<D> {
    SingletonArithLayerBuilderTempl(const std::string &name, ArithType type) : GPULayerBuilderTempl<D>(name), opType_(type) {
        LayerBuilderData::type_ = LayerType::SINGLETON_ARITH;
    }
    FYN_FLUENT(operand(float opd), operand_ = opd)
    ArithType opType_;
    float operand_ = 0.0f;
};

struct SingletonArithLayerBuilder : SingletonArithLayerBuilderTempl<SingletonArithLayerBuilder> {
    SingletonArithLayerBuilder(const std::string &name, ArithType type) : SingletonArithLayerBuilderTempl<SingletonArithLayerBuilder>(name, type) {}
};

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
