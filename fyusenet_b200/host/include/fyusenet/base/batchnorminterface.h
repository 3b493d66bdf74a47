// Parameter-loading interface of stand-alone batchnorm layers (reference: base/batchnorminterface.h:33-48).
// Data: scale[C] followed by bias[C].
#pragma once
#include <cstddef>
namespace fyusion {
namespace fyusenet {
class BatchNormInterface {
 public:
    virtual ~BatchNormInterface() = default;
    virtual void loadScaleAndBias(const float *scaleAndBias, size_t sbOffset = 0) = 0;
};
}  // namespace fyusenet
}  // namespace fyusion
