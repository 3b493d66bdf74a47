// GPU layer builders: GPULayerBuilder, ConvLayerBuilder, PoolLayerBuilder, UpDownLayerBuilder.
// Reference: fyusenet/gpu/gpulayerbuilder.h, convlayerbuilder.h, poollayerbuilder.h, updownlayerbuilder.h.
#pragma once
#include <functional>
#include <string>

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

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
