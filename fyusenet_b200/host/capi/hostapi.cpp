// hostapi.cpp -- small extern "C" surface over the C++ host engine, in the style of the reference's own C
// surfaces (samples/web/stylenet.cpp:190-258, samples/android/app/src/main/cpp/styletransfer.cpp:49-140):
// opaque handle, everything caught, int status.  Used by tests/ and bench.py (ctypes) to drive the same
// NeuralNetwork::setup()/forward() path a C++ user of the library calls.
#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>

#include <fyusenet/fyusenet.h>

#include "../samplenetworks/layerzoo.h"
#include "../samplenetworks/resnet50.h"
#include "../samplenetworks/stylenet.h"

using namespace fyusion;
using namespace fyusion::fyusenet;

namespace {
thread_local std::string g_error;

struct NetHandle {
    enum Kind { STYLE, RESNET, ZOO } kind;
    std::unique_ptr<StyleNetBase> style;
    std::unique_ptr<ResNet50> resnet;
    std::unique_ptr<LayerZoo> zoo;
    GfxContextLink ctx;
    NeuralNetwork *net() {
        if (kind == STYLE) return style.get();
        if (kind == RESNET) return resnet.get();
        return zoo.get();
    }
    void setInput(const float *hwc) {
        if (kind == STYLE) style->setInputBuffer(hwc);
        else if (kind == RESNET) resnet->setInputBuffer(hwc);
        else zoo->setInputBuffer(hwc);
    }
    cpu::CPUBuffer *inputBuffer() { return kind == STYLE ? style->inputBuffer() : (kind == RESNET ? resnet->inputBuffer() : zoo->inputBuffer()); }
    cpu::CPUBuffer *outputBuffer() { return kind == STYLE ? style->getOutputBuffer() : (kind == RESNET ? resnet->getOutputBuffer() : zoo->getOutputBuffer()); }
};

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &ex) {
        g_error = ex.what();
    } catch (...) {
        g_error = "unknown exception";
    }
    return -1;
}
}  // namespace

// asynchronous (pipelined) operation; must be called before setup.  Completed sequences are counted and the last
// delivered download buffer is remembered (both updated from the engine's completion callback).
struct AsyncState {
    std::atomic<uint64_t> completed{0};
    std::atomic<uint64_t> lastSequence{0};
    std::atomic<const float *> lastData{nullptr};
};
static std::unordered_map<void *, std::shared_ptr<AsyncState>> g_async;


extern "C" {

const char *fynhost_last_error(void) { return g_error.c_str(); }

// device < 0 in the create calls builds the network object without a device context (layer tables and weight
// offsets only; setup() then needs a GPU).
// storage: 0 = fp16 (reference default), 1 = fp32 (HIGH_PRECISION)
int fynhost_set_storage_precision(int fp32) {
    return guarded([&] { gpu::setStoragePrecision(fp32 ? BufferSpec::FLOAT32 : BufferSpec::FLOAT16); });
}

void *fynhost_stylenet_create(int kernel, int width, int height, int upload, int download, int device) {
    NetHandle *h = nullptr;
    int rc = guarded([&] {
        if (kernel != 3 && kernel != 9) THROW_EXCEPTION_ARGS(FynException, "StyleNet kernel must be 3 or 9");
        std::unique_ptr<NetHandle> nh(new NetHandle());
        nh->kind = NetHandle::STYLE;
        if (device >= 0) nh->ctx = GfxContextManager::instance(device)->createMainContext();
        if (kernel == 3) nh->style.reset(new StyleNet3x3(width, height, upload != 0, download != 0, nh->ctx));
        else nh->style.reset(new StyleNet9x9(width, height, upload != 0, download != 0, nh->ctx));
        h = nh.release();
    });
    return rc == 0 ? h : nullptr;
}

// 8-bit frames in and out (StyleNetBase::setByteIO; before setup)
int fynhost_stylenet_set_byte_io(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        if (h->kind != NetHandle::STYLE) THROW_EXCEPTION_ARGS(FynException, "Not a StyleNet");
        h->style->setByteIO(on != 0);
    });
}

// 8-bit images in (ResNet50::setByteInput; before setup)
int fynhost_resnet50_set_byte_input(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        if (h->kind != NetHandle::RESNET) THROW_EXCEPTION_ARGS(FynException, "Not a ResNet-50");
        h->resnet->setByteInput(on != 0);
    });
}

// raw views of the pinned input (slot < 0: the synchronous buffer) / output buffers with their size in bytes and element type
// (0 float32, 1 float16, 2 uint8), for networks with 8-bit I/O
void *fynhost_net_input_raw(void *handle, int slot, size_t *numBytes, int *dataType) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    void *ptr = nullptr;
    guarded([&] {
        cpu::CPUBuffer *buf = (slot >= 0 && h->kind == NetHandle::STYLE) ? h->style->inputBuffer(slot) : h->inputBuffer();
        if (numBytes) *numBytes = buf->bytes();
        if (dataType) *dataType = (int)buf->shape().dataType();
        ptr = buf->raw();
    });
    return ptr;
}

const void *fynhost_net_output_raw(void *handle, size_t *numBytes, int *dataType) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    const void *ptr = nullptr;
    guarded([&] {
        cpu::CPUBuffer *buf = h->outputBuffer();
        if (!buf) THROW_EXCEPTION_ARGS(FynException, "Network has no output buffer");
        if (numBytes) *numBytes = buf->bytes();
        if (dataType) *dataType = (int)buf->shape().dataType();
        ptr = buf->raw();
    });
    return ptr;
}

void *fynhost_resnet50_create(int device, int batch) {
    NetHandle *h = nullptr;
    int rc = guarded([&] {
        std::unique_ptr<NetHandle> nh(new NetHandle());
        nh->kind = NetHandle::RESNET;
        if (device >= 0) nh->ctx = GfxContextManager::instance(device)->createMainContext();
        nh->resnet.reset(new ResNet50(nh->ctx));
        nh->resnet->setBatch(batch);
        h = nh.release();
    });
    return rc == 0 ? h : nullptr;
}

// LayerZoo (samplenetworks/layerzoo.h): every SURVEY 8f rank-2 layer in one weight-free network
void *fynhost_layerzoo_create(int width, int height, int device) {
    NetHandle *h = nullptr;
    int rc = guarded([&] {
        std::unique_ptr<NetHandle> nh(new NetHandle());
        nh->kind = NetHandle::ZOO;
        if (device >= 0) nh->ctx = GfxContextManager::instance(device)->createMainContext();
        nh->zoo.reset(new LayerZoo(width, height, nh->ctx));
        h = nh.release();
    });
    return rc == 0 ? h : nullptr;
}

void fynhost_net_destroy(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    if (!h) return;
    guarded([&] { h->net()->cleanup(); });
    g_async.erase(handle);
    delete h;
}

size_t fynhost_net_weight_floats(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return h->kind == NetHandle::STYLE ? h->style->weightSize() : h->resnet->weightSize();
}

// float offset of a layer's block inside the weight file, or -1
long long fynhost_net_weight_offset(void *handle, int layerNumber) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    if (h->kind == NetHandle::STYLE) {
        auto &m = h->style->weightOffsets();
        auto it = m.find(layerNumber);
        return it == m.end() ? -1 : (long long)it->second;
    }
    auto &m = h->resnet->weightOffsets();
    auto it = m.find(layerNumber);
    return it == m.end() ? -1 : (long long)it->second;
}

int fynhost_net_load_weights(void *handle, const float *weights, size_t numFloats) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        if (h->kind == NetHandle::STYLE) h->style->loadWeightsAndBiases(weights, numFloats);
        else h->resnet->loadWeightsAndBiases(weights, numFloats);
    });
}

int fynhost_net_asynchronous(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        auto st = std::make_shared<AsyncState>();
        g_async[handle] = st;
        NeuralNetwork::AsyncAdapter adapter;
        adapter.downloadReady([st](const std::string &, uint64_t seq, cpu::CPUBuffer *buf) {
            st->lastSequence = seq;
            st->lastData = buf ? static_cast<const float *>(buf->raw()) : nullptr;
            st->completed++;
        });
        h->net()->asynchronous(adapter);
    });
}

// number of sequences whose download has been delivered; *lastSequence / *data describe the most recent one
uint64_t fynhost_net_async_completed(void *handle, uint64_t *lastSequence, const float **data) {
    auto it = g_async.find(handle);
    if (it == g_async.end()) return 0;
    if (lastSequence) *lastSequence = it->second->lastSequence;
    if (data) *data = it->second->lastData;
    return it->second->completed;
}

// pinned input buffer `slot` of an asynchronous StyleNet: sequence s uploads from slot s % fynhost_async_slots()
float *fynhost_stylenet_input_buffer_slot(void *handle, int slot, size_t *numFloats) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    float *ptr = nullptr;
    guarded([&] {
        if (h->kind != NetHandle::STYLE) THROW_EXCEPTION_ARGS(FynException, "Not a StyleNet");
        cpu::CPUBuffer *buf = h->style->inputBuffer(slot);
        if (numFloats) *numFloats = buf->bytes() / sizeof(float);
        ptr = static_cast<float *>(buf->raw());
    });
    return ptr;
}

int fynhost_async_slots(void) { return Engine::ASYNC_SLOTS; }

int fynhost_net_set_batch(void *handle, int batch) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->setBatch(batch); });
}

int fynhost_net_setup(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->setup(); });
}

// host float32 [batch][H][W][3]; copied into the network's pinned upload buffer
int fynhost_net_set_input(void *handle, const float *hwc) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        h->setInput(hwc);
    });
}

// pinned upload buffer of the network (created on first use and attached to the upload layer): callers that
// write their frames straight into it avoid the extra host copy of setInputBuffer ("one deep-copy operation too
// many", stylenet_base.cpp:150)
float *fynhost_net_input_buffer(void *handle, size_t *numFloats) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    float *ptr = nullptr;
    guarded([&] {
        cpu::CPUBuffer *buf = h->inputBuffer();
        if (numFloats) *numFloats = buf->bytes() / sizeof(float);
        ptr = buf->map<float>();
        buf->unmap();
    });
    return ptr;
}

int fynhost_net_forward(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        NeuralNetwork::execstate st = h->net()->forward();
        if (st.status != Engine::EXEC_DONE && st.status != Engine::EXEC_DEFERRED)
            THROW_EXCEPTION_ARGS(FynException, "forward() returned state %d", (int)st.status);
    });
}

int fynhost_net_finish(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->finish(); });
}

// pointer into the download layer's (pinned) CPU buffer; valid until the next forward()
const float *fynhost_net_output(void *handle, size_t *numFloats) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    const float *ptr = nullptr;
    guarded([&] {
        cpu::CPUBuffer *buf = h->outputBuffer();
        if (!buf) THROW_EXCEPTION_ARGS(FynException, "Network has no output buffer");
        if (numFloats) *numFloats = buf->bytes() / sizeof(float);
        ptr = buf->map<float>();
        buf->unmap();
    });
    return ptr;
}

// device-resident I/O for networks built without upload / download layers
int fynhost_stylenet_set_input_tensor(void *handle, fyn_tensor *t) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        if (h->kind != NetHandle::STYLE) THROW_EXCEPTION_ARGS(FynException, "Not a StyleNet");
        h->style->setInputTexture(t);
    });
}

fyn_tensor *fynhost_stylenet_output_tensor(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return h->kind == NetHandle::STYLE ? h->style->getOutputTexture() : nullptr;
}

fyn_ctx *fynhost_net_context(void *handle) { return static_cast<NetHandle *>(handle)->ctx.handle(); }
void *fynhost_net_stream(void *handle) { return static_cast<NetHandle *>(handle)->ctx.stream(); }

int fynhost_net_use_stream(void *handle, void *stream) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->ctx.interface()->setStream(stream); });
}

int fynhost_net_num_layers(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    int n = 0;
    guarded([&] { n = (int)h->net()->engine()->getLayers().size(); });
    return n;
}

// layer listing: fills number / channels / width / height / conv backend family for the idx-th layer in
// execution order; name copied into `name` (cap bytes)
int fynhost_net_layer_info(void *handle, int idx, int *number, int *channels, int *width, int *height, int *family,
                           char *name, int cap) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        CompiledLayers &layers = h->net()->engine()->getLayers();
        int i = 0;
        for (auto it = layers.begin(); it != layers.end(); ++it, ++i) {
            if (i != idx) continue;
            LayerBase *l = it.second;
            if (number) *number = l->getNumber();
            if (channels) *channels = l->numOutputChannels();
            std::vector<BufferSpec> outs = l->getRequiredOutputBuffers();
            if (width) *width = outs.empty() ? 0 : outs[0].width_;
            if (height) *height = outs.empty() ? 0 : outs[0].height_;
            auto *conv = dynamic_cast<gpu::ConvLayerBase *>(l);
            if (family) *family = conv ? conv->backendFamily() : 0;
            if (name && cap > 0) snprintf(name, cap, "%s", l->getName().c_str());
            return;
        }
        THROW_EXCEPTION_ARGS(FynException, "Layer index %d out of range", idx);
    });
}

// CHW float32 result of a layer (LayerBase::writeResult / copyResult format), blocking
int fynhost_net_copy_layer_result(void *handle, int layerNumber, float *chw, size_t capFloats) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        LayerBase *l = h->net()->engine()->getLayers()[layerNumber];
        auto *g = dynamic_cast<gpu::GPULayerBase *>(l);
        if (!g || !g->hasOutputTexture(0)) THROW_EXCEPTION_ARGS(FynException, "Layer %d has no device output", layerNumber);
        fyn_tensor_desc d{};
        FYN_ABI_CALL(fyn_tensor_get_desc(g->getOutputTexture(0), &d, nullptr));
        size_t need = (size_t)d.batch * d.channels * d.height * d.width;
        if (capFloats < need) THROW_EXCEPTION_ARGS(FynException, "Buffer too small (%zu < %zu)", capFloats, need);
        g->copyResult(chw, false);
    });
}

int fynhost_net_enable_dumps(void *handle, const char *dir) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->enableIntermediateOutput(dir); });
}

// layer fusion (conv + sigmoid in one kernel): on by default, suspended while dumps are written
int fynhost_net_enable_fusion(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->enableFusion(on != 0); });
}

int fynhost_net_fused_layers(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    int n = -1;
    guarded([&] { n = h->net()->engine()->fusedLayers(); });
    return n;
}

// chains of same-geometry convolutions as one persistent kernel (Engine::enableChains); number of layers running in chains
int fynhost_net_enable_chains(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->enableChains(on != 0); });
}

int fynhost_net_chained_layers(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    int n = -1;
    guarded([&] { n = h->net()->engine()->chainedLayers(); });
    return n;
}

int fynhost_net_halo_exchanges(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    int n = -1;
    guarded([&] { n = h->net()->engine()->haloExchanges(); });
    return n;
}

// like fynhost_net_enable_timings(handle, 1) but with an event pair around one layer only
int fynhost_net_enable_layer_timing(void *handle, int layerNumber) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        h->net()->engine()->resetTimings();
        h->net()->engine()->enableTimings(layerNumber);
    });
}

int fynhost_net_enable_timings(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        if (on) {
            h->net()->engine()->resetTimings();
            h->net()->engine()->enableTimings();
        } else {
            h->net()->engine()->disableTimings();
        }
    });
}

// accumulated device milliseconds / host microseconds of a layer since timings were enabled
int fynhost_net_layer_timing(void *handle, int layerNumber, float *deviceMs, unsigned *hostUs) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] {
        const auto &dev = h->net()->engine()->getDeviceTimings();
        auto &host = h->net()->engine()->getTimings();
        auto d = dev.find(layerNumber);
        auto u = host.find(layerNumber);
        if (deviceMs) *deviceMs = d == dev.end() ? 0.f : d->second;
        if (hostUs) *hostUs = u == host.end() ? 0u : u->second;
    });
}

// CUDA-graph replay of the device layers of the synchronous path (Engine::enableGraph)
int fynhost_net_enable_graph(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->enableGraph(on != 0); });
}

int fynhost_net_graph_active(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    int v = 0;
    guarded([&] { v = h->net()->engine()->graphActive() ? 1 : 0; });
    return v;
}

// device-resident operation of a network with upload / download layers: both are skipped (Engine::skipIO)
int fynhost_net_skip_io(void *handle, int on) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->skipIO(on != 0); });
}

// output tensor of a layer (device-side consumers: fyn_allgather_logits on ResNet-50's GEMM72 output)
fyn_tensor *fynhost_net_layer_tensor(void *handle, int layerNumber) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    fyn_tensor *t = nullptr;
    guarded([&] {
        auto *g = dynamic_cast<gpu::GPULayerBase *>(h->net()->engine()->getLayers()[layerNumber]);
        if (!g || !g->hasOutputTexture(0)) THROW_EXCEPTION_ARGS(FynException, "Layer %d has no device output", layerNumber);
        t = g->getOutputTexture(0);
    });
    return t;
}

// row-banded operation over several GPUs: margin rows refreshed from the band neighbours after every layer (Engine::setHaloExchange)
int fynhost_net_set_halo_exchange(void *handle, fyn_comm *comm, int marginRows, int inputHeight) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return guarded([&] { h->net()->engine()->setHaloExchange(comm, marginRows, inputHeight); });
}

size_t fynhost_net_device_bytes(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return h->net()->bufferManager() ? h->net()->bufferManager()->estimateTextureMemory() : 0;
}

int fynhost_net_num_tensors(void *handle) {
    NetHandle *h = static_cast<NetHandle *>(handle);
    return h->net()->bufferManager() ? h->net()->bufferManager()->numTensors() : 0;
}

}  // extern "C"
