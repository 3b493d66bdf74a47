#include "resnet50.h"

#include <algorithm>
#include <cstring>
#include <string>

using namespace fyusion;
using namespace fyusion::fyusenet;

namespace {
struct Stage { int mid, out, blocks, size; };
const Stage kStages[4] = {{64, 256, 3, 56}, {128, 512, 4, 28}, {256, 1024, 6, 14}, {512, 2048, 3, 7}};
}  // namespace

ResNet50::ResNet50(const GfxContextLink &ctx) : NeuralNetwork(ctx) {
    buildGraph();
    wbData_.assign(totalWeightBytes_ / sizeof(float), 0.f);
}

ResNet50::~ResNet50() {
    cleanup();
    delete inBuffer_;
}

// Generates the node list.  Per stage: first block = {1x1 reduce (for stages > 0 it already exists as the tail
// of the previous stage and runs at the previous resolution), 1x1 projection shortcut (stride 2 for stages > 0),
// 3x3 (stride 2 for stages > 0), 1x1 expand + residual(shortcut)}; further blocks = {BN, 1x1, 3x3, 1x1 + residual(
// previous block)}; the last block's expand conv carries post-BN that is also applied to its residual.
void ResNet50::buildGraph() {
    auto conv = [&](int no, int k, int cin, int cout, int size, int ds, int inPad, int outPad, bool relu, bool postBN,
                    bool bnRes, int input, int residual) {
        nodes_.push_back({Node::CONV, no, "Conv", k, cin, cout, size, ds, inPad, outPad, true, relu, postBN, bnRes, input, residual});
    };
    auto bn = [&](int no, int c, int size, bool deep, int outPad, int input) {
        nodes_.push_back({Node::BN, no, "BN", 1, c, c, size, 1, 0, outPad, deep, false, false, false, input, -1});
    };
    // weight-file order = layer-number order, except that each stage's projection shortcut is stored after the
    // first block's expand conv (6,8,9,7 / 18,20,21,19 / 34,36,37,35 / 58,60,61,59: resnet50.cpp:539-677)
    std::vector<int> fileOrder;
    nodes_.push_back({Node::UPLOAD, 0, "upload", 1, 3, 3, IMAGE_SIZE, 1, 0, 0, false, false, false, false, -1, -1});
    bn(2, 3, IMAGE_SIZE, false, 1, 0);
    conv(3, 7, 3, 64, IMAGE_SIZE, 2, 1, 1, false, true, false, 2, -1);
    nodes_.push_back({Node::MAXPOOL, 4, "MaxPool", 3, 64, 64, 112, 2, 1, 0, true, true, false, false, 3, -1});
    bn(5, 64, 56, true, 0, 4);
    fileOrder = {2, 3, 5};
    int no = 6, feed = 5, blockOut = -1, reduceOfNextStage = -1;
    for (int s = 0; s < 4; s++) {
        const Stage &st = kStages[s];
        const int cin = s == 0 ? 64 : kStages[s - 1].out;
        const int inSize = s == 0 ? st.size : st.size * 2;
        const int ds = s == 0 ? 1 : 2;
        int reduce;
        if (s == 0) {
            reduce = no;
            fileOrder.push_back(no);
            conv(no++, 1, cin, st.mid, inSize, 1, 0, 1, true, true, false, feed, -1);
        } else {
            reduce = reduceOfNextStage;
        }
        const int shortcut = no;
        conv(no++, 1, cin, st.out, inSize, ds, 0, 0, true, false, false, feed, -1);
        const int mid3 = no;
        conv(no++, 3, st.mid, st.mid, inSize, ds, 1, 0, true, true, false, reduce, -1);
        blockOut = no;
        conv(no++, 1, st.mid, st.out, st.size, 1, 0, 0, true, false, false, mid3, shortcut);
        fileOrder.insert(fileOrder.end(), {mid3, blockOut, shortcut});
        for (int b = 1; b < st.blocks; b++) {
            const bool last = (b == st.blocks - 1);
            const int bnNo = no;
            bn(no++, st.out, st.size, true, 0, blockOut);
            const int a = no;
            conv(no++, 1, st.out, st.mid, st.size, 1, 0, 1, true, true, false, bnNo, -1);
            const int m = no;
            conv(no++, 3, st.mid, st.mid, st.size, 1, 1, 0, true, true, false, a, -1);
            const int c = no;
            conv(no++, 1, st.mid, st.out, st.size, 1, 0, 0, true, last, last, m, blockOut);
            fileOrder.insert(fileOrder.end(), {bnNo, a, m, c});
            blockOut = c;
        }
        feed = blockOut;
        if (s < 3) {
            reduceOfNextStage = no;
            fileOrder.push_back(no);
            conv(no++, 1, st.out, kStages[s + 1].mid, st.size, 1, 0, 1, true, true, false, feed, -1);
        }
    }
    nodes_.push_back({Node::GAP, 70, "GlobAvg", 7, 2048, 2048, 7, 7, 0, 0, true, true, false, false, 69, -1});
    nodes_.push_back({Node::GEMM, 72, "GEMM", 1, 2048, 1000, 1, 1, 0, 0, true, false, false, false, 70, -1});
    nodes_.push_back({Node::DOWNLOAD, 73, "download", 1, 1000, 1000, 1, 1, 0, 0, true, false, false, false, 72, -1});
    fileOrder.push_back(72);
    std::vector<const Node *> order;
    for (int n : fileOrder)
        for (const Node &node : nodes_)
            if (node.no == n) order.push_back(&node);
    uint32_t off = 0;
    for (const Node *n : order) {
        uint32_t sz;
        if (n->kind == Node::BN) sz = 2u * n->cin;
        else sz = (uint32_t)n->cout + (uint32_t)n->kernel * n->kernel * n->cin * n->cout + (n->postBN ? 2u * n->cout : 0u);
        weightOffsets_[n->no] = off;
        weightSizes_[n->no] = sz;
        off += sz;
    }
    totalWeightBytes_ = (size_t)off * sizeof(float);
}

void ResNet50::loadWeightsAndBiases(const float *data, size_t numFloats) {
    if (numFloats * sizeof(float) != totalWeightBytes_)
        THROW_EXCEPTION_ARGS(FynException, "Weight blob has %zu floats, expected %zu", numFloats, totalWeightBytes_ / sizeof(float));
    memcpy(wbData_.data(), data, totalWeightBytes_);
    if (setup_) initializeWeights(engine_->getLayers());
}

void ResNet50::initializeWeights(CompiledLayers &layers) {
    for (auto it = layers.begin(); it != layers.end(); ++it) {
        if (auto *conv = dynamic_cast<ConvLayerInterface *>(it.second))
            conv->loadWeightsAndBiases(wbData_.data(), weightOffsets_.at(it.second->getNumber()));
        else if (auto *bn = dynamic_cast<BatchNormInterface *>(it.second))
            bn->loadScaleAndBias(wbData_.data(), weightOffsets_.at(it.second->getNumber()));
    }
}

CompiledLayers ResNet50::buildLayers() {
    std::shared_ptr<LayerFactory> factory = getLayerFactory();
    for (const Node &n : nodes_) {
        const std::string name = (n.kind == Node::UPLOAD || n.kind == Node::DOWNLOAD) ? n.prefix : std::string(n.prefix) + std::to_string(n.no);
        switch (n.kind) {
            case Node::UPLOAD: {
                auto *b = new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::UPLOAD, name);
                b->shape(3, n.size, n.size, 3).context(context()).number(n.no);
                if (byteInput_) b->dataType(BufferSpec::UBYTE);
                b->push(factory);
                break;
            }
            case Node::DOWNLOAD: {
                auto *b = new gpu::UpDownLayerBuilder(gpu::UpDownLayerBuilder::DOWNLOAD, name);
                b->shape(n.cout, 1, 1, n.cin).context(context()).deep().number(n.no);
                b->push(factory);
                break;
            }
            case Node::BN: {
                auto *b = new gpu::GPULayerBuilder(name);
                b->type(LayerType::BATCHNORM).number(n.no).shape(n.cout, n.size, n.size, n.cin).outputPadding((short)n.outPad).context(context());
                if (n.deep) b->deep();
                b->push(factory);
                break;
            }
            case Node::CONV: {
                auto *b = new gpu::ConvLayerBuilder((short)n.kernel, name);
                b->type(LayerType::CONVOLUTION2D).number(n.no).shape(n.cout, n.size, n.size, n.cin).deep()
                    .inputPadding((short)n.inPad).outputPadding((short)n.outPad).context(context());
                if (n.ds > 1) b->downsample(n.ds);
                if (n.preRelu) b->prefixAct(ActType::RELU);
                if (n.postBN) b->postfixNorm(NormType::BATCHNORM);
                if (n.residual >= 0) b->residual(ActType::NONE, n.bnOnResidual);
                b->push(factory);
                break;
            }
            case Node::MAXPOOL: {
                auto *b = new gpu::PoolLayerBuilder(gpu::PoolLayerBuilder::POOL_MAX, name);
                b->type(LayerType::MAXPOOL2D).number(n.no).shape(n.cout, n.size, n.size, n.cin).poolSize(3, 3).downsample(2).deep()
                    .inputPadding(1).prefixAct(ActType::RELU).context(context());
                b->push(factory);
                break;
            }
            case Node::GAP: {
                auto *b = new gpu::PoolLayerBuilder(gpu::PoolLayerBuilder::POOL_AVG, name);
                b->type(LayerType::AVGPOOL2D).number(n.no).shape(n.cout, n.size, n.size, n.cin).global().deep().prefixAct(ActType::RELU).context(context());
                b->push(factory);
                break;
            }
            case Node::GEMM: {
                auto *b = new gpu::GPULayerBuilder(name);
                b->type(LayerType::GEMM).number(n.no).shape(n.cout, 1, 1, n.cin).deep().context(context());
                b->push(factory);
                break;
            }
        }
    }
    return factory->compileLayers();
}

void ResNet50::connectLayers(CompiledLayers &layers, BufferManager *bufMgr) {
    // Edges are issued sorted by (producer, consumer) like the hand-written list of resnet50.cpp:426-516;
    // the order matters for pooled-tensor reuse.
    struct Edge { int from, to, port; };
    std::vector<Edge> edges;
    for (const Node &n : nodes_) {
        if (n.input >= 0) edges.push_back({n.input, n.no, 0});
        if (n.residual >= 0) edges.push_back({n.residual, n.no, 1});
    }
    std::sort(edges.begin(), edges.end(), [](const Edge &a, const Edge &b) { return a.from != b.from ? a.from < b.from : a.to < b.to; });
    for (const Edge &e : edges) bufMgr->connectLayers(layers[e.from], layers[e.to], e.port);
    bufMgr->createCPUOutput(layers[73], true);
}

ResNet50::CPUBuffer *ResNet50::inputBuffer() {
    if (!setup_) THROW_EXCEPTION_ARGS(FynException, "Please run setup() before setting input buffers");
    if (!inBuffer_) {
        cpu::CPUBufferShape shape(IMAGE_SIZE, IMAGE_SIZE, 3, 0, byteInput_ ? cpu::CPUBufferShape::UINT8 : cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_SHALLOW, batch_);
        inBuffer_ = shape.createBuffer(context());
    }
    static_cast<gpu::UploadLayer *>(engine_->getLayers()["upload"])->setInputBuffer(inBuffer_, 0);
    return inBuffer_;
}

void ResNet50::setByteInput(bool on) {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "setByteInput() must be called before setup()");
    byteInput_ = on;
}

void ResNet50::setInputBuffer(const float *data) {
    if (byteInput_) THROW_EXCEPTION_ARGS(FynException, "Network takes 8-bit images (setByteInput)");
    CPUBuffer *buf = inputBuffer();
    float *tgt = buf->map<float>();
    memcpy(tgt, data, buf->bytes());
    buf->unmap();
}

ResNet50::CPUBuffer *ResNet50::getOutputBuffer() {
    if (!setup_) return nullptr;
    return static_cast<gpu::DownloadLayer *>(engine_->getLayers()["download"])->getOutputBuffer(0);
}
