// ResNet-50 sample network on the CUDA backend (reference: samples/samplenetworks/resnet50.h:23-33, resnet50.cpp).
// Pre-activation bottleneck network, 224x224 input, raw logits out.  The layer graph is generated from the
// four stage descriptions instead of being written out layer by layer; numbering, flags, paddings and the
// weight-file order reproduce resnet50.cpp:200-516 and :539-677 (asserted in tests against those tables).
// New on this backend: a batch dimension (setBatch before setup) -- images are stacked along the batch axis of
// every device tensor; batch 1 is the reference behaviour.
#pragma once
#include <unordered_map>
#include <vector>

#include <fyusenet/fyusenet.h>

class ResNet50 : public fyusion::fyusenet::NeuralNetwork {
 public:
    using CPUBuffer = fyusion::fyusenet::cpu::CPUBuffer;
    constexpr static int IMAGE_SIZE = 224;
    constexpr static size_t TOTAL_WEIGHT_BYTES = 102304184;

    explicit ResNet50(const fyusion::fyusenet::GfxContextLink &ctx = fyusion::fyusenet::GfxContextLink());
    ~ResNet50() override;
    void setInputBuffer(const float *data);
    // 8-bit images in (before setup()): the upload layer takes UBYTE data and the bytes become value / 255 on the device -- the
    // conversion samples/desktop/resnet.cpp:48-51 does on the host, bit for bit -- so a batch crosses PCIe as 3 instead of 12 bytes per pixel
    void setByteInput(bool on);
    bool byteInput() const { return byteInput_; }
    CPUBuffer *getOutputBuffer();
    // the pinned host buffer the upload layer reads from (created on demand and attached to the upload layer)
    CPUBuffer *inputBuffer();
    void loadWeightsAndBiases(const float *data, size_t numFloats);
    const std::unordered_map<int, uint32_t> &weightOffsets() const { return weightOffsets_; }
    size_t weightSize() const { return totalWeightBytes_ / sizeof(float); }

    struct Node {
        enum Kind { UPLOAD, BN, CONV, MAXPOOL, GAP, GEMM, DOWNLOAD } kind;
        int no;
        const char *prefix;
        int kernel, cin, cout, size, ds, inPad, outPad;
        bool deep, preRelu, postBN, bnOnResidual;
        int input, residual;  // producer layer numbers (-1 = none)
    };
    const std::vector<Node> &nodes() const { return nodes_; }

 protected:
    fyusion::fyusenet::CompiledLayers buildLayers() override;
    void initializeWeights(fyusion::fyusenet::CompiledLayers &layers) override;
    void connectLayers(fyusion::fyusenet::CompiledLayers &layers, fyusion::fyusenet::BufferManager *buffers) override;
    void buildGraph();

    std::vector<Node> nodes_;
    std::unordered_map<int, uint32_t> weightOffsets_, weightSizes_;
    std::vector<float> wbData_;
    size_t totalWeightBytes_ = 0;
    CPUBuffer *inBuffer_ = nullptr;
    bool byteInput_ = false;
};
