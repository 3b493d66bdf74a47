// LayerBase: what the engine, the buffer manager and the factory see of a layer.
// Reference: fyusenet/base/layerbase.h:81-404.  Activation-at-fetch semantics (:50-59): a layer's
// prefix activation is applied when it READS its input; stored tensors hold pre-activation values.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../common/fynexception.h"
#include "bufferspec.h"
#include "layerbuilder.h"
#include "layerflags.h"

namespace fyusion {
namespace fyusenet {

class LayerBase {
 public:
    constexpr static int PIXEL_PACKING = gpu::PIXEL_PACKING;

    explicit LayerBase(const LayerBuilderData &builder, int layerNumber = -1)
        : name_(builder.name_), leakyReLU_(builder.leakyReLU_), lowClip_(builder.clipLow_), highClip_(builder.clipHigh_),
          flags_(builder.getFlags()), width_(builder.width()), height_(builder.height()), inputChannels_(builder.in()),
          outputChannels_(builder.out()), layerNumber_(layerNumber >= 0 ? layerNumber : builder.number_),
          inputPadding_((uint8_t)builder.inputPadding_), outputPadding_((uint8_t)builder.outputPadding_),
          residualPadding_((uint8_t)builder.residualPadding_), device_(builder.device_) {
        if (flags_ & LayerFlags::POST_RELU) THROW_EXCEPTION_ARGS(FynException, "Post-ReLU not supported by GPU layers");
    }
    virtual ~LayerBase() = default;

    // GPU resources are created in setup() and released in cleanup(), not in the destructor (:112-123)
    virtual void setup() = 0;
    virtual void cleanup() = 0;
    virtual void forward(uint64_t sequence = 0) = 0;
    virtual std::vector<BufferSpec> getRequiredInputBuffers() const = 0;
    virtual std::vector<BufferSpec> getRequiredOutputBuffers() const = 0;
    // dump of the result as float32 [C][H][W] (optionally with padding), the parity interchange format (:160-172)
    virtual void writeResult(const char *fileName, bool includePadding = false) = 0;

    virtual int numInputPorts() const { return (flags_ & LayerFlags::RESIDUAL_INPUT) ? 2 : 1; }
    virtual bool isConnected() const {
        for (int p = 0; p < numInputPorts(); p++)
            if (!isConnected(p)) return false;
        return outputConnected_;
    }
    virtual bool isConnected(int port) const {
        for (int p : connectedInputPorts_)
            if (p == port) return true;
        return false;
    }
    virtual void addInputConnection(int port, LayerBase *, int) { connectedInputPorts_.push_back(port); }
    virtual void addOutputConnection(int, LayerBase *, int) { outputConnected_ = true; }

    int getInputPadding() const { return inputPadding_; }
    int getOutputPadding() const { return outputPadding_; }
    int getResidualPadding() const { return residualPadding_; }
    int getWidth() const { return width_; }
    int getHeight() const { return height_; }
    layerflags getFlags() const { return flags_; }
    int getNumber() const { return layerNumber_; }
    virtual int numInputChannels(int = 0) const { return inputChannels_; }
    int numOutputChannels() const { return outputChannels_; }
    const std::string &getName() const { return name_; }
    bool isValid() const { return valid_; }
    compute_device getDevice() const { return device_; }

 protected:
    std::string name_;
    float leakyReLU_ = 0.f, lowClip_ = 0.f, highClip_ = 0.f;
    layerflags flags_ = LayerFlags::NO_LAYER_FLAGS;
    int width_ = 0, height_ = 0, inputChannels_ = 0, outputChannels_ = 0;
    int layerNumber_ = -1;
    uint8_t inputPadding_ = 0, outputPadding_ = 0, residualPadding_ = 0;
    bool outputConnected_ = false;
    std::vector<int> connectedInputPorts_;
    compute_device device_ = compute_device::DEV_ILLEGAL;
    bool valid_ = false;
};

}  // namespace fyusenet
}  // namespace fyusion
