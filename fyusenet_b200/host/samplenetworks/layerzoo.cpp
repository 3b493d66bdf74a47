#include "layerzoo.h"

#include <cstring>

using namespace fyusion;
using namespace fyusion::fyusenet;

LayerZoo::LayerZoo(int width, int height, const GfxContextLink &ctx) : NeuralNetwork(ctx), width_(width), height_(height) {
    if (width < 2 || height < 2) THROW_EXCEPTION_ARGS(FynException, "LayerZoo needs at least 2x2 pixels (got %dx%d)", width, height);
}

LayerZoo::~LayerZoo() {
    cleanup();
    delete inBuffer_;
}

CompiledLayers LayerZoo::buildLayers() {
    std::shared_ptr<LayerFactory> factory = getLayerFactory();
    const int w = width_, h = height_;
    (new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::UPLOAD, "upload"))->shape(3, h, w, 3).context(context()).number(UPLOAD).push(factory);
    (new gpu::GPULayerBuilder("bgr"))->shape(3, h, w, 3).type(LayerType::RGB2BGR).context(context()).number(BGR).push(factory);
    (new gpu::ScaleLayerBuilder("upscale"))->scale(2.0f).scaleType(ScalingType::LINEAR).shape(3, h, w, 3).outputPadding(1).context(context()).number(UPSCALE).push(factory);
    (new gpu::SingletonArithLayerBuilder("twice", ArithType::MUL))->operand(2.0f).shape(3, 2 * h, 2 * w, 3).inputPadding(1).outputPadding(1).context(context()).number(TWICE).push(factory);
    (new gpu::GPULayerBuilder("diff"))->shape(3, 2 * h, 2 * w, 3).type(LayerType::SUB).inputPadding(1).outputPadding(1).context(context()).number(DIFF).push(factory);
    (new gpu::ConcatLayerBuilder("concat"))->input(3, 1).input(3, 1).input(3, 1).shape(9, 2 * h, 2 * w, 9).inputPadding(1).context(context()).number(CONCAT).push(factory);
    (new gpu::GPULayerBuilder("clip"))->shape(9, 2 * h, 2 * w, 9).type(LayerType::CLIP).clip(CLIP_LOW, CLIP_HIGH).outputPadding(1).context(context()).number(CLIP).push(factory);
    (new gpu::GPULayerBuilder("todeep"))->shape(9, 2 * h, 2 * w, 9).type(LayerType::SHALLOW2DEEP).inputPadding(1).outputPadding(1).context(context()).number(TODEEP).push(factory);
    (new gpu::ScaleLayerBuilder("downscale"))->scale(0.5f).shape(9, 2 * h, 2 * w, 9).deep().inputPadding(1).context(context()).number(DOWNSCALE).push(factory);
    (new gpu::GPULayerBuilder("toshallow"))->shape(9, h, w, 9).type(LayerType::DEEP2SHALLOW).context(context()).number(TOSHALLOW).push(factory);
    (new gpu::GPULayerBuilder("pad"))->shape(9, h, w, 9).type(LayerType::PADDING2D).outputPadding(2).context(context()).number(PAD).push(factory);
    (new gpu::GPULayerBuilder("sum"))->shape(9, h, w, 9).type(LayerType::ADD).inputPadding(2).context(context()).number(SUM).push(factory);
    (new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::DOWNLOAD, "download"))->shape(9, h, w, 9).context(context()).number(DOWNLOAD).push(factory);
    return factory->compileLayers();
}

void LayerZoo::connectLayers(CompiledLayers &layers, BufferManager *buffers) {
    buffers->connectLayers(layers[UPLOAD], layers[BGR], 0);
    buffers->connectLayers(layers[BGR], layers[UPSCALE], 0);
    buffers->connectLayers(layers[UPSCALE], layers[TWICE], 0);
    buffers->connectLayers(layers[TWICE], layers[DIFF], 0);
    buffers->connectLayers(layers[UPSCALE], layers[DIFF], 1);
    buffers->connectLayers(layers[UPSCALE], layers[CONCAT], 0);
    buffers->connectLayers(layers[DIFF], layers[CONCAT], 1);
    buffers->connectLayers(layers[TWICE], layers[CONCAT], 2);
    buffers->connectLayers(layers[CONCAT], layers[CLIP], 0);
    buffers->connectLayers(layers[CLIP], layers[TODEEP], 0);
    buffers->connectLayers(layers[TODEEP], layers[DOWNSCALE], 0);
    buffers->connectLayers(layers[DOWNSCALE], layers[TOSHALLOW], 0);
    buffers->connectLayers(layers[TOSHALLOW], layers[PAD], 0);
    buffers->connectLayers(layers[PAD], layers[SUM], 0);
    buffers->connectLayers(layers[PAD], layers[SUM], 1);
    buffers->connectLayers(layers[SUM], layers[DOWNLOAD], 0);
    buffers->createCPUOutput(layers[DOWNLOAD], true);
}

LayerZoo::CPUBuffer *LayerZoo::inputBuffer() {
    if (!setup_) THROW_EXCEPTION_ARGS(FynException, "Please run setup() before setting input buffers");
    if (!inBuffer_) {
        cpu::CPUBufferShape shape(height_, width_, 3, 0, cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_SHALLOW, batch_);
        inBuffer_ = shape.createBuffer(context());
    }
    static_cast<gpu::UploadLayer *>(engine_->getLayers()["upload"])->setInputBuffer(inBuffer_, 0);
    return inBuffer_;
}

void LayerZoo::setInputBuffer(const float *data) {
    CPUBuffer *buf = inputBuffer();
    memcpy(buf->map<float>(), data, buf->bytes());
    buf->unmap();
}

LayerZoo::CPUBuffer *LayerZoo::getOutputBuffer() {
    if (!setup_) return nullptr;
    return static_cast<gpu::DownloadLayer *>(engine_->getLayers()["download"])->getOutputBuffer(0);
}
