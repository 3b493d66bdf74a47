// LayerFactory + LayerFactoryBackend: THE PLUGIN SEAM.
// Reference: fyusenet/base/layerfactory.h:94-176, layerfactory.cpp:41-101,146-166.  Builders are pushed
// (the factory owns them), compileLayers() asks the backend for one layer object per builder.  The only
// GPU backend here is the CUDA one (gpu/cudalayerfactory.h); there is no CPU backend and no fallback.
#pragma once
#include <memory>
#include <string>
#include <unordered_map>

#include "../gpu/gfxcontextlink.h"
#include "compiledlayers.h"
#include "layerbase.h"
#include "layerfactoryinterface.h"

namespace fyusion {
namespace fyusenet {

class LayerFactoryBackend {
    friend class LayerFactory;

 public:
    virtual ~LayerFactoryBackend() = default;
    virtual std::string getName() const = 0;
    virtual LayerBase *createLayer(LayerType type, LayerBuilder *builder, int layerNumber) = 0;
};

class LayerFactory : public LayerFactoryInterface {
 public:
    struct FactoryType {
        explicit FactoryType(compute_device t) : factoryType(t) {}
        virtual ~FactoryType() = default;
        virtual LayerFactoryBackend *createBackend() = 0;
        compute_device factoryType;
    };
    struct GPUFactoryType : FactoryType {
        enum gputype { VANILLA = 0, SPECIALIZED };
        explicit GPUFactoryType(gputype tp, GfxContextLink context = GfxContextLink())
            : FactoryType(compute_device::DEV_GPU), gpuType(tp), gfxContext(context) {}
        LayerFactoryBackend *createBackend() override;  // -> CUDALayerFactoryBackend
        gputype gpuType;
        GfxContextLink gfxContext;
    };

    ~LayerFactory() override;
    std::string getName() const { return backend_->getName(); }
    void pushBuilder(LayerBuilder *builder) override;
    virtual CompiledLayers compileLayers();

    template <class T>
    static std::shared_ptr<LayerFactory> instance(T typ) {
        return std::shared_ptr<LayerFactory>(new LayerFactory(typ.createBackend()));
    }
    // plug in a foreign backend (what a maintainer of the reference does with this library)
    static std::shared_ptr<LayerFactory> withBackend(LayerFactoryBackend *backend) {
        return std::shared_ptr<LayerFactory>(new LayerFactory(backend));
    }

 protected:
    explicit LayerFactory(LayerFactoryBackend *backend) : backend_(backend) {}
    LayerFactoryBackend *backend_ = nullptr;
    std::unordered_map<int, LayerBuilderData *> builders_;
    CompiledLayers layers_;
};

template <typename D>
void LayerBuilderTempl<D>::push(std::shared_ptr<LayerFactory> &factory) {
    if (!factory) THROW_EXCEPTION_ARGS(FynException, "No factory supplied");
    static_cast<LayerFactoryInterface *>(factory.get())->pushBuilder(reinterpret_cast<LayerBuilder *>(this));
}

}  // namespace fyusenet
}  // namespace fyusion
