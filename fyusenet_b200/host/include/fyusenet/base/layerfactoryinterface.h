// Minimal interface the builders push themselves through (reference: base/layerfactoryinterface.h).
#pragma once
namespace fyusion {
namespace fyusenet {
struct LayerBuilder;
class LayerFactoryInterface {
 public:
    virtual ~LayerFactoryInterface() = default;
    virtual void pushBuilder(LayerBuilder *builder) = 0;
};
}  // namespace fyusenet
}  // namespace fyusion
