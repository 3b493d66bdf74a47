// Engine + NeuralNetwork.
#include "fyusenet/base/engine.h"

#include <cstdio>

#include "fyusenet/base/neuralnetwork.h"
#include "fyusenet/common/performance.h"
#include "fyusenet/gpu/cudalayers.h"

namespace fyusion {
namespace fyusenet {

Engine::Engine(const GfxContextLink &ctx, bool async) : async_(async) { setContext(ctx); }

Engine::~Engine() {}

void Engine::setup(NeuralNetwork *net) {
    if (!net) THROW_EXCEPTION_ARGS(FynException, "Null network");
    assertContext();
    layers_ = net->glSetup();
    setup_ = true;
}

void Engine::cleanup() {
    if (setup_) {
        FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
        collectTimings(false);
        for (void *e : freeEvents_) fyn_event_destroy(context_.handle(), e);
        freeEvents_.clear();
        layers_.cleanup();
    }
    layers_ = CompiledLayers();
    setup_ = false;
}

void Engine::resetTimings() {
    collectTimings(true);
    timingData_.clear();
    deviceTimingData_.clear();
    runs_ = 0;
}

Engine::execstate Engine::forwardLayers() {
    if (!setup_) return EXEC_ERROR;
    uint64_t seq = sequenceNo_++;
    return execute(seq);
}

// strict ascending-layer-number execution (reference: engine.cpp:386-683, hot loop 1).
// Timings: host microseconds around each forward() like the reference (:443-449,595-601) plus device time from
// CUDA event pairs recorded around every layer WITHOUT host synchronisation; the pairs are resolved lazily
// (collectTimings) once the stream has been synchronised, so enabling timings does not serialise the step.
Engine::execstate Engine::execute(uint64_t sequence) {
    size_t slot = 0;
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        LayerBase *layer = it.second;
        if (timings_) {
            if (pendingEvents_.size() >= 4096) collectTimings(true);
            void *evA = nullptr, *evB = nullptr;
            if (freeEvents_.size() >= 2) {
                evA = freeEvents_.back(); freeEvents_.pop_back();
                evB = freeEvents_.back(); freeEvents_.pop_back();
            } else {
                FYN_ABI_CALL(fyn_event_create(context_.handle(), &evA));
                FYN_ABI_CALL(fyn_event_create(context_.handle(), &evB));
            }
            fyn_event_record(context_.handle(), evA, context_.stream());
            tstamp t0 = fy_get_stamp();
            layer->forward(sequence);
            tstamp t1 = fy_get_stamp();
            fyn_event_record(context_.handle(), evB, context_.stream());
            timingData_[it.first] += (uint32_t)fy_elapsed_micros(t0, t1);
            pendingEvents_.push_back({it.first, evA, evB});
        } else {
            layer->forward(sequence);
        }
        if (writeResults_) {
            char fname[1024];
            snprintf(fname, sizeof(fname), "%s/%s_%llu.bin", outputDir_.c_str(), layer->getName().c_str(), (unsigned long long)sequence);
            layer->writeResult(fname, false);
        }
        slot++;
    }
    if (timings_) runs_++;
    return EXEC_DONE;
}

void Engine::collectTimings(bool sync) {
    if (pendingEvents_.empty()) return;
    if (sync) FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
    for (auto &p : pendingEvents_) {
        float ms = 0.f;
        if (fyn_event_elapsed_ms(context_.handle(), p.start, p.stop, &ms) == 0) deviceTimingData_[p.layer] += ms;
        freeEvents_.push_back(p.start);
        freeEvents_.push_back(p.stop);
    }
    pendingEvents_.clear();
}

Engine::execstate Engine::finish() {
    if (!setup_) return EXEC_ERROR;
    FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
    collectTimings(false);
    return EXEC_DONE;
}

// ------------------------------------------------------------------------------------------------
// NeuralNetwork
// ------------------------------------------------------------------------------------------------
NeuralNetwork::NeuralNetwork(const GfxContextLink &ctx) {
    // like the reference, the main context is used when none is passed (neuralnetwork.cpp:57-66); it is
    // resolved lazily in setup() so that network objects (layer tables, weight offsets) exist without a device
    setContext(ctx);
}

NeuralNetwork::~NeuralNetwork() {
    if (engine_ || bufferMgr_) cleanup();
}

void NeuralNetwork::setBatch(int batch) {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "Batch size must be set before setup()");
    if (batch < 1) THROW_EXCEPTION_ARGS(FynException, "Illegal batch size %d", batch);
    batch_ = batch;
}

void NeuralNetwork::setup() {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "Network already set up");
    if (!context_.isValid()) setContext(GfxContextManager::instance(0)->createMainContext());
    engine_ = new Engine(context_, async_);
    engine_->setup(this);
    setup_ = true;
}

void NeuralNetwork::cleanup() {
    if (engine_) {
        engine_->cleanup();
        delete engine_;
        engine_ = nullptr;
    }
    if (bufferMgr_) {
        bufferMgr_->cleanup();
        delete bufferMgr_;
        bufferMgr_ = nullptr;
    }
    setup_ = false;
}

NeuralNetwork::execstate NeuralNetwork::forward() {
    execstate st;
    if (!setup_ || !engine_) {
        st.status = Engine::EXEC_ERROR;
        return st;
    }
    st.sequenceNo = engine_->nextSequenceNo();
    st.status = engine_->forwardLayers();
    return st;
}

NeuralNetwork::execstate NeuralNetwork::finish() {
    execstate st;
    st.status = engine_ ? engine_->finish() : Engine::EXEC_ERROR;
    st.sequenceNo = engine_ ? engine_->lastSequenceNo() : 0;
    return st;
}

// build -> connect -> load weights -> per-layer setup (reference: neuralnetwork.cpp:238-250)
CompiledLayers NeuralNetwork::glSetup() {
    CompiledLayers layers = buildLayers();
    bufferMgr_ = new BufferManager(context_, batch_);
    connectLayers(layers, bufferMgr_);
    initializeWeights(layers);
    for (auto it = layers.begin(); it != layers.end(); ++it) it.second->setup();
    return layers;
}

std::shared_ptr<LayerFactory> NeuralNetwork::getLayerFactory(compute_device dev) {
    if (dev != compute_device::DEV_GPU) THROW_EXCEPTION_ARGS(FynException, "Device type not supported");
    return LayerFactory::instance(LayerFactory::GPUFactoryType(LayerFactory::GPUFactoryType::VANILLA, context_));
}

}  // namespace fyusenet
}  // namespace fyusion
