#include "stylenet.h"

#include <algorithm>
#include <cstring>

using namespace fyusion;
using namespace fyusion::fyusenet;

StyleNetBase::StyleNetBase(int kernel, int resBlocks, int width, int height, bool upload, bool download, const GfxContextLink &ctx)
    : NeuralNetwork(ctx), kernel_(kernel), resBlocks_(resBlocks), width_(width), height_(height), upload_(upload), download_(download) {
    // width / height must be multiples of 4 (two stride-2 stages; samples/desktop/stylenet.cpp:108-111)
    if ((width & 3) || (height & 3)) THROW_EXCEPTION_ARGS(FynException, "Width and height must be multiples of 4 (got %dx%d)", width, height);
    // layer table; file order: conv1..3, deconv1..3, then the residual blocks
    convs_.push_back({"conv1", kernel, 3, 12, 1, 1, false, 1.f, true, false, false, 0});
    convs_.push_back({"conv2", 3, 12, 20, 1, 2, false, 1.f, true, false, false, 1});
    convs_.push_back({"conv3", 3, 20, 40, 2, 2, false, 1.f, true, false, false, 2});
    static const char *resNames[5][2] = {{"res1_1", "res1_2"}, {"res2_1", "res2_2"}, {"res3_1", "res3_2"}, {"res4_1", "res4_2"}, {"res5_1", "res5_2"}};
    for (int r = 0; r < resBlocks; r++) {
        // res2_1 carries no prefix activation; res1_2 ReLUs its residual (stylenet9x9.cpp:149-163)
        convs_.push_back({resNames[r][0], 3, 40, 40, 4, 1, false, 1.f, r != 1, false, false, 6 + 2 * r});
        convs_.push_back({resNames[r][1], 3, 40, 40, 4, 1, false, 1.f, true, true, r == 0, 7 + 2 * r});
    }
    convs_.push_back({"deconv1", 3, 40, 20, 4, 2, true, 0.5f, false, false, false, 3});
    convs_.push_back({"deconv2", 3, 20, 12, 4, 2, true, 0.25f, true, false, false, 4});
    convs_.push_back({"deconv3", kernel, 12, 3, 2, 1, true, 0.5f, true, false, false, 5});
    // offsets by walking the table in file order
    std::vector<int> order(convs_.size());
    for (size_t i = 0; i < convs_.size(); i++) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return convs_[a].fileOrder < convs_[b].fileOrder; });
    size_t off = 0;
    for (int idx : order) {
        const ConvSpec &c = convs_[idx];
        weightOffsets_[CONV1 + idx] = off;
        off += (size_t)c.cout + (size_t)c.cout * c.kernel * c.kernel * c.cin;
    }
    totalWeights_ = off;
    wbData_.assign(totalWeights_, 0.f);
    sigmoidLayer_ = CONV1 + (int)convs_.size();
    downloadLayer_ = sigmoidLayer_ + 1;
    lastLayer_ = download ? downloadLayer_ : sigmoidLayer_;
}

StyleNetBase::~StyleNetBase() {
    cleanup();
    delete inBuffer_;
    for (CPUBuffer *b : inBuffers_) delete b;
}

// In asynchronous mode the upload layer reads input buffer (sequence % ASYNC_SLOTS): the caller fills inputBuffer(slot) for
// the next sequence while the previous one is still being processed (the reference cycles two upload buffers the
// same way, stylenet_base.cpp:110-155).
NeuralNetwork::execstate StyleNetBase::forward() {
    if (async_ && upload_ && setup_) {
        const int slot = (int)(engine_->nextSequenceNo() % Engine::ASYNC_SLOTS);
        static_cast<gpu::UploadLayer *>(engine_->getLayers()["upload"])->setInputBuffer(inputBuffer(slot), 0);
    }
    return NeuralNetwork::forward();
}

StyleNetBase::CPUBuffer *StyleNetBase::inputBuffer(int slot) {
    if (!setup_) THROW_EXCEPTION_ARGS(FynException, "Please run setup() before setting input buffers");
    if (!upload_) THROW_EXCEPTION_ARGS(FynException, "Network was created without an upload layer");
    if (slot < 0 || slot >= Engine::ASYNC_SLOTS) THROW_EXCEPTION_ARGS(FynException, "Illegal input buffer %d", slot);
    if (!inBuffers_[slot]) {
        cpu::CPUBufferShape shape(height_, width_, 3, 0, byteIO_ ? cpu::CPUBufferShape::UINT8 : cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_SHALLOW, batch_);
        inBuffers_[slot] = shape.createBuffer(context());
    }
    return inBuffers_[slot];
}

void StyleNetBase::loadWeightsAndBiases(const float *weightsAndBiases, size_t size) {
    if (size != totalWeights_) THROW_EXCEPTION_ARGS(FynException, "Weight blob has %zu floats, expected %zu", size, totalWeights_);
    memcpy(wbData_.data(), weightsAndBiases, size * sizeof(float));
    if (setup_) initializeWeights(engine_->getLayers());  // change the style of a live net (stylenet9x9.cpp:87-95)
}

void StyleNetBase::initializeWeights(CompiledLayers &layers) {
    for (auto it = layers.begin(); it != layers.end(); ++it) {
        ConvLayerInterface *conv = dynamic_cast<ConvLayerInterface *>(it.second);
        if (conv) conv->loadWeightsAndBiases(wbData_.data(), weightOffsets_.at(it.second->getNumber()));
    }
}

CompiledLayers StyleNetBase::buildLayers() {
    std::shared_ptr<LayerFactory> factory = getLayerFactory();
    if (upload_) {
        auto *up = new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::UPLOAD, "upload");
        up->shape(3, height_, width_, 3).context(context()).number(UPLOAD);
        if (byteIO_) up->dataType(BufferSpec::UBYTE);
        if (async_) up->async();
        up->push(factory);
    }
    for (size_t i = 0; i < convs_.size(); i++) {
        const ConvSpec &c = convs_[i];
        auto *b = new gpu::ConvLayerBuilder((short)c.kernel, c.name);
        b->shape(c.cout, height_ / c.scaleDiv, width_ / c.scaleDiv, c.cin)
            .type(c.fractional ? LayerType::FRACCONVOLUTION2D : LayerType::CONVOLUTION2D)
            .context(context())
            .number(CONV1 + (int)i);
        if (c.downsample > 1) b->downsample(c.downsample);
        if (c.fractional) b->sourceStep(c.sourceStep);
        if (c.preRelu) b->prefixAct(ActType::RELU);
        if (c.residual) b->residual(c.reluOnResidual ? ActType::RELU : ActType::NONE);
        b->push(factory);
    }
    auto *sig = new gpu::GPULayerBuilder("sigmoid");
    sig->shape(3, height_, width_, 3).type(LayerType::SIGMOID).context(context()).number(sigmoidLayer_);
    sig->push(factory);
    if (download_) {
        auto *down = new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::DOWNLOAD, "download");
        down->shape(4, height_, width_, 4).context(context()).number(downloadLayer_);
        if (byteIO_) down->dataType(BufferSpec::UBYTE);
        if (async_) down->async();
        down->push(factory);
    }
    return factory->compileLayers();
}

void StyleNetBase::connectLayers(CompiledLayers &layers, BufferManager *buffers) {
    if (upload_) buffers->connectLayers(layers[UPLOAD], layers[CONV1], 0);
    for (size_t i = 1; i < convs_.size(); i++) {
        int no = CONV1 + (int)i;
        // connection order of the reference: block input -> resN_1 (port 0), block input -> resN_2 (port 1),
        // resN_1 -> resN_2 (port 0)  (stylenet9x9.cpp:230-246); the order decides which pooled tensors get reused
        if (convs_[i].residual) buffers->connectLayers(layers[no - 2], layers[no], 1);
        buffers->connectLayers(layers[no - 1], layers[no], 0);
    }
    buffers->connectLayers(layers[sigmoidLayer_ - 1], layers[sigmoidLayer_], 0);
    if (download_) {
        buffers->connectLayers(layers[sigmoidLayer_], layers[downloadLayer_], 0);
        buffers->createCPUOutput(layers[downloadLayer_], true);
    } else {
        buffers->createGPUOutput(static_cast<gpu::GPULayerBase *>(layers[sigmoidLayer_]));
    }
    if (!upload_ && inputTexture_) static_cast<gpu::GPULayerBase *>(layers[CONV1])->addInputTexture(inputTexture_, 0);
}

StyleNetBase::CPUBuffer *StyleNetBase::inputBuffer() {
    if (!setup_) THROW_EXCEPTION_ARGS(FynException, "Please run setup() before setting input buffers");
    if (!upload_) THROW_EXCEPTION_ARGS(FynException, "Network was created without an upload layer");
    if (!inBuffer_) {
        cpu::CPUBufferShape shape(height_, width_, 3, 0, byteIO_ ? cpu::CPUBufferShape::UINT8 : cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_SHALLOW, batch_);
        inBuffer_ = shape.createBuffer(context());  // pinned staging, the role of the upload PBO
    }
    static_cast<gpu::UploadLayer *>(engine_->getLayers()["upload"])->setInputBuffer(inBuffer_, 0);
    return inBuffer_;
}

void StyleNetBase::setByteIO(bool on) {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "The I/O data type must be chosen before setup()");
    byteIO_ = on;
}

void StyleNetBase::setInputBuffer(const float *data) {
    if (byteIO_) THROW_EXCEPTION_ARGS(FynException, "Network takes 8-bit frames (setByteIO)");
    CPUBuffer *buf = async_ ? inputBuffer((int)(engine_->nextSequenceNo() % Engine::ASYNC_SLOTS)) : inputBuffer();
    float *tgt = buf->map<float>();
    memcpy(tgt, data, buf->bytes());
    buf->unmap();
}

void StyleNetBase::setInputBuffer(const uint8_t *data) {
    if (!byteIO_) THROW_EXCEPTION_ARGS(FynException, "Network takes float32 frames (see setByteIO)");
    CPUBuffer *buf = async_ ? inputBuffer((int)(engine_->nextSequenceNo() % Engine::ASYNC_SLOTS)) : inputBuffer();
    memcpy(buf->raw(), data, buf->bytes());
}

StyleNetBase::CPUBuffer *StyleNetBase::getOutputBuffer() {
    if (!download_ || !setup_) return nullptr;
    auto *dwn = static_cast<gpu::DownloadLayer *>(engine_->getLayers()["download"]);
    return dwn->getOutputBuffer(0);
}

void StyleNetBase::setInputTexture(fyn_tensor *texture) {
    inputTexture_ = texture;
    if (setup_ && texture) {
        auto *layer = static_cast<gpu::GPULayerBase *>(engine_->getLayers()[CONV1]);
        if (layer->hasInputTexture(0)) layer->updateInputTexture(texture, 0);
        else layer->addInputTexture(texture, 0);
    }
}

fyn_tensor *StyleNetBase::getOutputTexture() const {
    if (!engine_) return nullptr;
    auto *layer = static_cast<gpu::GPULayerBase *>(engine_->getLayers()["sigmoid"]);
    return layer ? layer->getOutputTexture(0) : nullptr;
}

void StyleNetBase::enableDebugOutput(const std::string &outDir) {
    if (!engine_) THROW_EXCEPTION_ARGS(FynException, "Please run setup() before setting debug output");
    engine_->enableIntermediateOutput(outDir);
}
