// Weight-loading interface of all convolution-type layers (reference: base/convlayerinterface.h:31-58).
// Blob format: bias[Co], W[Co][Ky][Kx][Ci], then with post-BN: bnScale[Co], bnBias[Co]; raw float32.
#pragma once
#include <cstddef>
namespace fyusion {
namespace fyusenet {
class ConvLayerInterface {
 public:
    virtual ~ConvLayerInterface() = default;
    virtual void loadWeightsAndBiases(const float *biasAndWeights, size_t offset = 0) = 0;
};
}  // namespace fyusenet
}  // namespace fyusion
