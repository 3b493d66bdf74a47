// NeuralNetwork: base class of all networks (reference: fyusenet/base/neuralnetwork.h:171-183,
// neuralnetwork.cpp:57-250).  Subclasses implement buildLayers / connectLayers / initializeWeights; setup()
// builds the engine, forward() runs one inference (not re-entrant, single caller thread).
#pragma once
#include <cstdint>
#include <functional>
#include <memory>

#include "buffermanager.h"
#include "compiledlayers.h"
#include "engine.h"
#include "layerfactory.h"

namespace fyusion {
namespace fyusenet {

class NeuralNetwork : public GfxContextTracker {
    friend class Engine;

 public:
    using state = Engine::execstate;
    struct execstate {
        state status = Engine::EXEC_DONE;
        uint64_t sequenceNo = 0;
    };

    explicit NeuralNetwork(const GfxContextLink &ctx = GfxContextLink());
    virtual ~NeuralNetwork();
    virtual void cleanup();
    virtual void setup();
    virtual execstate forward();
    virtual execstate finish();
    // Asynchronous operation (reference: neuralnetwork.h:96-160, neuralnetwork.cpp:205-211).  Must be called before
    // setup().  forward() then only enqueues (EXEC_DEFERRED, at most two sequences in flight) and the adapter's
    // callbacks fire from a driver thread when an upload buffer may be refilled / a download has landed.
    class AsyncAdapter {
     public:
        AsyncAdapter &newSequence(const std::function<void(uint64_t)> &cb) { newSeq_ = cb; return *this; }
        AsyncAdapter &sequenceDone(const std::function<void(uint64_t)> &cb) { seqDone_ = cb; return *this; }
        AsyncAdapter &downloadReady(const std::function<void(const std::string &, uint64_t, cpu::CPUBuffer *)> &cb) { downReady_ = cb; return *this; }
        AsyncAdapter &uploadReady(const std::function<void(const std::string &, uint64_t)> &cb) { upReady_ = cb; return *this; }
        std::function<void(uint64_t)> newSeq_, seqDone_;
        std::function<void(const std::string &, uint64_t, cpu::CPUBuffer *)> downReady_;
        std::function<void(const std::string &, uint64_t)> upReady_;
    };
    virtual void asynchronous(const AsyncAdapter &adapter = AsyncAdapter());
    bool isAsynchronous() const { return async_; }
    uint64_t nextSequenceNo() const { return engine_ ? engine_->nextSequenceNo() : 0; }
    uint64_t lastSequenceNo() const { return engine_ ? engine_->lastSequenceNo() : 0; }
    // batch is new on this backend (the reference is batch-1, README.md:72); must be set before setup()
    void setBatch(int batch);
    int batch() const { return batch_; }
    Engine *engine() const { return engine_; }
    BufferManager *bufferManager() const { return bufferMgr_; }

 protected:
    virtual CompiledLayers glSetup();
    std::shared_ptr<LayerFactory> getLayerFactory(compute_device dev = compute_device::DEV_GPU);
    virtual void initializeWeights(CompiledLayers &layers) = 0;
    virtual CompiledLayers buildLayers() = 0;
    virtual void connectLayers(CompiledLayers &layers, BufferManager *buffers) = 0;

    bool async_ = false;
    AsyncAdapter asyncCallbacks_;
    Engine *engine_ = nullptr;
    BufferManager *bufferMgr_ = nullptr;
    bool setup_ = false;
    int batch_ = 1;
};

}  // namespace fyusenet
}  // namespace fyusion
