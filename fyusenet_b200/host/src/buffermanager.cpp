// BufferManager: device-tensor pool with liveness-based reuse + host output buffers.
#include "fyusenet/base/buffermanager.h"

#include "fyusenet/gpu/cudalayers.h"

namespace fyusion {
namespace fyusenet {

BufferManager::BufferManager(const GfxContextLink &ctx, int batch) : batch_(batch < 1 ? 1 : batch) { setContext(ctx); }

BufferManager::~BufferManager() { cleanup(); }

void BufferManager::cleanup() {
    for (auto &e : pool_)
        if (e.tensor) fyn_tensor_destroy(e.tensor);
    pool_.clear();
    for (CPUBuffer *b : cpuBuffers_) delete b;
    cpuBuffers_.clear();
    deviceBytes_ = 0;
}

bool BufferManager::sameDesc(const fyn_tensor_desc &a, const fyn_tensor_desc &b) {
    return a.width == b.width && a.height == b.height && a.channels == b.channels && a.padding == b.padding &&
           a.order == b.order && a.dtype == b.dtype && a.batch == b.batch && a.packing == b.packing;
}

fyn_tensor_desc BufferManager::toDesc(const BufferSpec &s) const {
    fyn_tensor_desc d{};
    d.width = s.width_;
    d.height = s.height_;
    d.channels = s.channels_;
    d.padding = s.padding_;
    d.order = (s.dataOrder_ == BufferSpec::order::GPU_DEEP) ? FYN_ORDER_DEEP : FYN_ORDER_SHALLOW;
    d.dtype = (s.type_ == BufferSpec::FLOAT16) ? FYN_F16 : FYN_F32;
    d.batch = batch_;
    d.packing = s.packing_;
    return d;
}

// liveness rule of the reference's texture pool (buffermanager.cpp:512-525): a pooled tensor can become
// the output of layer `outputLayer` consumed by `inputLayer` iff it is unlocked, was last read by a layer
// before inputLayer-1, and the new producer runs after that last reader.
int BufferManager::findTensor(int inputLayer, int outputLayer, const fyn_tensor_desc &d) const {
    for (int i = 0; i < (int)pool_.size(); i++) {
        const Entry &e = pool_[i];
        if (e.locked || !sameDesc(e.desc, d)) continue;
        if (e.lastInputLayer < inputLayer - 1 && outputLayer > e.lastInputLayer) return i;
    }
    return -1;
}

BufferManager::Entry &BufferManager::createTensor(const fyn_tensor_desc &d) {
    Entry e;
    e.desc = d;
    FYN_ABI_CALL(fyn_tensor_create(context_.handle(), &d, &e.tensor));
    fyn_tensor_geom g{};
    fyn_tensor_get_desc(e.tensor, &e.desc, &g);
    deviceBytes_ += g.bytes;
    pool_.push_back(e);
    return pool_.back();
}

void BufferManager::touch(fyn_tensor *t, int inputLayer, bool lock) {
    for (auto &e : pool_)
        if (e.tensor == t) {
            if (inputLayer > e.lastInputLayer) e.lastInputLayer = inputLayer;
            e.locked |= lock;
        }
}

static bool specsMatch(const BufferSpec &in, const BufferSpec &out) {
    // the reference matches per 4-channel texture on (width, height, channelIndex, format); with one tensor per
    // port this becomes: same net size + padding + plane count + layout, and same storage type unless the
    // consumer takes any texture type (upload textures)
    if (in.device_ != out.device_) return false;
    if (in.width_ != out.width_ || in.height_ != out.height_ || in.padding_ != out.padding_) return false;
    if ((in.channels_ + 3) / 4 != (out.channels_ + 3) / 4) return false;
    bool singleTile = in.channels_ <= 4 && out.channels_ <= 4;  // deep == shallow for one tile
    if (in.dataOrder_ != out.dataOrder_ && !singleTile) return false;
    if (!in.anyType_ && (in.type_ != out.type_ || in.packing_ != out.packing_)) return false;
    return true;
}

void BufferManager::connectLayers(LayerBase *outputLayer, LayerBase *inputLayer, int port, bool lock) {
    if (!outputLayer || !inputLayer)
        THROW_EXCEPTION_ARGS(FynException, "Illegal parameters out=%p in=%p", (void *)outputLayer, (void *)inputLayer);
    if (inputLayer->getNumber() <= outputLayer->getNumber())
        THROW_EXCEPTION_ARGS(FynException, "Layer %s (#%d) cannot feed layer %s (#%d): execution is in ascending layer number",
                             outputLayer->getName().c_str(), outputLayer->getNumber(), inputLayer->getName().c_str(), inputLayer->getNumber());
    const std::vector<BufferSpec> inputs = inputLayer->getRequiredInputBuffers();
    const std::vector<BufferSpec> outputs = outputLayer->getRequiredOutputBuffers();
    if (inputs.empty()) THROW_EXCEPTION_ARGS(FynException, "Input layer %s has no inputs", inputLayer->getName().c_str());
    if (outputs.empty()) THROW_EXCEPTION_ARGS(FynException, "Output layer %s has no outputs", outputLayer->getName().c_str());
    if (inputLayer->isConnected(port))
        THROW_EXCEPTION_ARGS(FynException, "Inputs/outputs do not match (I/O) for layers %s and %s", inputLayer->getName().c_str(),
                             outputLayer->getName().c_str());
    const BufferSpec *inSpec = nullptr;
    for (const BufferSpec &s : inputs)
        if (s.port_ == port && specsMatch(s, outputs[0])) inSpec = &s;
    if (!inSpec)
        THROW_EXCEPTION_ARGS(FynException, "Inputs/outputs do not match (I/O) for layers %s and %s", inputLayer->getName().c_str(),
                             outputLayer->getName().c_str());
    const BufferSpec &outSpec = outputs[0];
    gpu::GPULayerBase *outL = dynamic_cast<gpu::GPULayerBase *>(outputLayer);
    gpu::GPULayerBase *inL = dynamic_cast<gpu::GPULayerBase *>(inputLayer);
    if (!outL || !inL) THROW_EXCEPTION_ARGS(FynException, "Only GPU layers can be connected by this buffer manager");
    lock = lock || outSpec.lock_ || outSpec.async_;  // asynchronous producers always have locked outputs
    fyn_tensor *t = nullptr;
    if (outL->hasOutputTexture(0)) {
        // second consumer of an existing output
        t = outL->getOutputTexture(0);
        touch(t, inputLayer->getNumber(), lock);
    } else {
        fyn_tensor_desc d = toDesc(outSpec);
        int idx = lock ? -1 : findTensor(inputLayer->getNumber(), outputLayer->getNumber(), d);
        if (idx >= 0) {
            t = pool_[idx].tensor;
            touch(t, inputLayer->getNumber(), lock);
        } else {
            Entry &e = createTensor(d);
            e.lastInputLayer = inputLayer->getNumber();
            e.locked = lock;
            t = e.tensor;
            for (int m = 1; m < outSpec.multiplicity_; m++) {
                Entry &s = createTensor(d);  // shadow buffers of asynchronous producers
                s.lastInputLayer = inputLayer->getNumber();
                s.locked = true;
                outL->addOutputTexture(s.tensor, 0, m);
            }
        }
        outL->addOutputTexture(t, 0);
    }
    if (inSpec->usage_ == BufferSpec::RESIDUAL_SOURCE) inL->addResidualTexture(t, 0);
    else inL->addInputTexture(t, port);
    inputLayer->addInputConnection(port, outputLayer, 0);
    outputLayer->addOutputConnection(0, inputLayer, port);
}

void BufferManager::createCPUOutput(LayerBase *outputLayer, bool lock) {
    (void)lock;
    cpu::CPULayerInterface *cpuL = dynamic_cast<cpu::CPULayerInterface *>(outputLayer);
    if (!cpuL) THROW_EXCEPTION_ARGS(FynException, "Layer %s cannot write to a CPU buffer", outputLayer->getName().c_str());
    const std::vector<BufferSpec> outs = outputLayer->getRequiredOutputBuffers();
    if (outs.empty() || outs[0].device_ != BufferSpec::COMP_STOR_CPU)
        THROW_EXCEPTION_ARGS(FynException, "Layer %s has no CPU output", outputLayer->getName().c_str());
    const BufferSpec &s = outs[0];
    CPUBufferShape shape(s.height_, s.width_, s.channels_, s.padding_, s.type_ == BufferSpec::UBYTE ? CPUBufferShape::UINT8 : CPUBufferShape::FLOAT32,
                         s.dataOrder_, batch_);
    shape.uploadStyle(false);
    CPUBuffer *buf = shape.createBuffer(context_);  // pinned: the download is a true async copy
    cpuBuffers_.push_back(buf);
    cpuL->addOutputBuffer(buf, 0);
    outputLayer->addOutputConnection(0, nullptr, 0);
}

void BufferManager::createGPUOutput(gpu::GPULayerBase *outputLayer) {
    if (!outputLayer) THROW_EXCEPTION_ARGS(FynException, "Null layer");
    const std::vector<BufferSpec> outs = outputLayer->getRequiredOutputBuffers();
    if (outs.empty()) THROW_EXCEPTION_ARGS(FynException, "Layer %s has no outputs", outputLayer->getName().c_str());
    if (!outputLayer->hasOutputTexture(0)) {
        Entry &e = createTensor(toDesc(outs[0]));
        e.locked = true;
        e.lastInputLayer = 1 << 30;
        outputLayer->addOutputTexture(e.tensor, 0);
    }
    outputLayer->addOutputConnection(0, nullptr, 0);
}

}  // namespace fyusenet
}  // namespace fyusion
