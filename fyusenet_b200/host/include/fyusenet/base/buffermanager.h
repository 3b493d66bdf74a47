// BufferManager: CUDA device-tensor manager.
// Reference: fyusenet/base/buffermanager.cpp:86-184 (connectLayers / createCPUOutput / createGPUOutput),
// :370-442 (connect + reuse), :462-492 (checkIOMatch), :512-525 (findTexture liveness rule), :650-708.
// Tensors keep FyuseNet's 4-channel-packed padded layouts (see include/fyusenet_b200.h); the pool reuses a
// tensor as the output of layer o feeding layer i iff it is unlocked, has the identical descriptor,
// lastInputLayer < i-1 and o > lastInputLayer (same rule as the reference's texture pool).
#pragma once
#include <vector>

#include "../cpu/cpubuffer.h"
#include "../gpu/gpulayerbase.h"
#include "bufferspec.h"
#include "layerbase.h"

namespace fyusion {
namespace fyusenet {

class BufferManager : public GfxContextTracker {
 public:
    explicit BufferManager(const GfxContextLink &ctx = GfxContextLink(), int batch = 1);
    ~BufferManager();
    void cleanup();
    void connectLayers(LayerBase *outputLayer, LayerBase *inputLayer, int port, bool lock = false);
    void createCPUOutput(LayerBase *outputLayer, bool lock = false);
    void createGPUOutput(gpu::GPULayerBase *outputLayer);
    size_t estimateTextureMemory() const { return deviceBytes_; }
    int numTensors() const { return (int)pool_.size(); }
    int batch() const { return batch_; }

 private:
    struct Entry {
        fyn_tensor *tensor = nullptr;
        fyn_tensor_desc desc{};
        int lastInputLayer = -1;
        bool locked = false;
    };
    static bool sameDesc(const fyn_tensor_desc &a, const fyn_tensor_desc &b);
    fyn_tensor_desc toDesc(const BufferSpec &spec) const;
    int findTensor(int inputLayer, int outputLayer, const fyn_tensor_desc &d) const;
    Entry &createTensor(const fyn_tensor_desc &d);
    void touch(fyn_tensor *t, int inputLayer, bool lock);

    std::vector<Entry> pool_;
    std::vector<CPUBuffer *> cpuBuffers_;
    size_t deviceBytes_ = 0;
    int batch_ = 1;
};

}  // namespace fyusenet
}  // namespace fyusion
