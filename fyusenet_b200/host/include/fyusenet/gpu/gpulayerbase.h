// GPULayerBase: base of all device layers -- tensor slots, precision mode, result dumps.
// Reference: fyusenet/gpu/gpulayerbase.h:127-142 (addInputTexture / addResidualTexture / addOutputTexture /
// updateInputTexture / getOutputTexture / hasOutputTexture), :100-110 (fp16 default, fp32 with HIGH_PRECISION),
// gpulayerbase.cpp:443-564 (writeResult / copyResult).  A "texture" here is a device tensor of the C ABI that
// holds ALL channel planes of a port, so the per-plane channelIndex of the reference collapses to 0.
#pragma once
#include <atomic>
#include <cstdio>
#include <mutex>
#include <vector>

#include "../base/layerbase.h"
#include "gfxcontextlink.h"
#include "gpulayerbuilder.h"

namespace fyusion {
namespace fyusenet {
namespace gpu {

using TensorHandle = fyn_tensor *;

// storage precision of activations: FYN_F16 (reference default, RGBA16F) or FYN_F32 (HIGH_PRECISION build).
// Run-time switch FYN_STORAGE={fp16,fp32} / setStoragePrecision() replaces the reference's compile-time macro.
BufferSpec::dtype storagePrecision();
void setStoragePrecision(BufferSpec::dtype dt);

// Bumped whenever something a captured CUDA graph may have baked in changes (tensor bindings, weight images, fusions):
// the engine re-captures when its graph is older than this (Engine::enableGraph).
inline std::atomic<uint64_t> &graphEpoch() {
    static std::atomic<uint64_t> epoch{1};
    return epoch;
}

class GPULayerBase : public LayerBase, public GfxContextTracker {
 public:
    template <typename B>
    GPULayerBase(const B &builder, int layerNumber) : LayerBase(builder, layerNumber) {
        setContext(builder.context_);
        assertContext();
        viewport_[0] = width_ + 2 * outputPadding_;
        viewport_[1] = height_ + 2 * outputPadding_;
    }

    void cleanup() override {
        // layers never free tensors: the BufferManager owns them (reference: gpulayerbase.cpp:102-111)
        inputs_.clear();
        residuals_.clear();
        outputs_.clear();
        valid_ = false;
    }

    virtual void addInputTexture(TensorHandle t, int port) { put(inputs_, port, t); graphEpoch()++; }
    virtual void addResidualTexture(TensorHandle t, int index = 0) { put(residuals_, index, t); graphEpoch()++; }
    virtual void addOutputTexture(TensorHandle t, int index = 0, int shadowIndex = 0) {
        if (shadowIndex == 0) put(outputs_, index, t);
        else put(shadowOutputs_, shadowIndex - 1, t);
        graphEpoch()++;
    }
    virtual void updateInputTexture(TensorHandle t, int port) {
        if (port < (int)inputs_.size() && inputs_[port] == t) return;
        put(inputs_, port, t);
        graphEpoch()++;
    }
    virtual bool hasInputTexture(int port = 0) const { return port < (int)inputs_.size() && inputs_[port]; }
    virtual bool hasOutputTexture(int index = 0) const { return index < (int)outputs_.size() && outputs_[index]; }
    virtual TensorHandle getOutputTexture(int index = 0) const { return hasOutputTexture(index) ? outputs_[index] : nullptr; }
    virtual TensorHandle getInputTexture(int port = 0) const { return hasInputTexture(port) ? inputs_[port] : nullptr; }
    // output tensor of buffer `slot` (0 = primary, >0 = shadow buffers of asynchronous producers)
    TensorHandle getOutputTexture(int index, int slot) const {
        if (slot == 0) return getOutputTexture(index);
        return (slot - 1 < (int)shadowOutputs_.size()) ? shadowOutputs_[slot - 1] : nullptr;
    }
    // layers reading this layer's output (needed to re-point them at a shadow buffer)
    void addOutputConnection(int port, LayerBase *receiver, int receiverPort) override {
        LayerBase::addOutputConnection(port, receiver, receiverPort);
        if (receiver) receivers_.push_back({receiver, receiverPort});
    }
    const std::vector<std::pair<LayerBase *, int>> &receivers() const { return receivers_; }

    // float32 [C][H][W] dump without padding (reference: layerbase.h:160-172, gpulayerbase.cpp:443-523)
    void writeResult(const char *fileName, bool includePadding = false) override;
    // same data into caller memory (reference: copyResult, debug builds only there)
    virtual void copyResult(float *memory, bool includePadding = false);
    int outputBatch() const;

 protected:
    static void put(std::vector<TensorHandle> &v, int idx, TensorHandle t) {
        if (idx < 0) THROW_EXCEPTION_ARGS(FynException, "Illegal slot %d", idx);
        if ((int)v.size() <= idx) v.resize(idx + 1, nullptr);
        v[idx] = t;
    }
    TensorHandle in(int port = 0) const {
        if (!hasInputTexture(port)) THROW_EXCEPTION_ARGS(FynException, "Layer %s: input port %d not connected", name_.c_str(), port);
        return inputs_[port];
    }
    TensorHandle out() const {
        if (!hasOutputTexture(0)) THROW_EXCEPTION_ARGS(FynException, "Layer %s: output not connected", name_.c_str());
        return outputs_[0];
    }
    BufferSpec::order order() const { return (flags_ & LayerFlags::DEEP) ? BufferSpec::order::GPU_DEEP : BufferSpec::order::GPU_SHALLOW; }

    std::vector<TensorHandle> inputs_, residuals_, outputs_, shadowOutputs_;
    std::vector<std::pair<LayerBase *, int>> receivers_;
    int viewport_[2] = {0, 0};
    std::recursive_mutex processingLock_;  // reference: gpulayerbase.h:191
};

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
