// StyleNet 3x3 / 9x9 sample networks on the CUDA backend.
// Same public API as the reference's samples/samplenetworks/stylenet_base.h:36-61 (+ stylenet3x3.h, stylenet9x9.h):
// construct -> loadWeightsAndBiases -> setup -> setInputBuffer / setInputTexture -> forward -> getOutputBuffer.
// Topology restated from stylenet9x9.cpp:120-273 / stylenet3x3.cpp:114-234 as a table instead of one builder
// statement per layer; weight-file offsets are derived from the layer table (they equal the hard-coded
// offsets of stylenet9x9.cpp:41-56 / stylenet3x3.cpp:41-50, see tests).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include <fyusenet/fyusenet.h>

class StyleNetBase : public fyusion::fyusenet::NeuralNetwork {
 public:
    using CPUBuffer = fyusion::fyusenet::cpu::CPUBuffer;
    enum { UNPACK = 0, UPLOAD = 0, CONV1 = 1 };

    StyleNetBase(int kernel, int resBlocks, int width, int height, bool upload, bool download,
                 const fyusion::fyusenet::GfxContextLink &ctx = fyusion::fyusenet::GfxContextLink());
    ~StyleNetBase() override;

    void loadWeightsAndBiases(const float *weightsAndBiases, size_t size);
    // 8-bit frames in and out (call before setup()): the upload layer takes UBYTE buffers -- UpDownLayerBuilder::dataType(UBYTE),
    // an 8-bit normalised texture in the reference (gpu/uploadlayer.cpp:51-66) -- and the download layer delivers RGBA8 texels,
    // quantised on the device the way samples/desktop/stylenet.cpp:52-62 does on the host.  A frame then moves 3 + 4 instead of
    // 12 + 16 bytes per pixel over PCIe.  Default: float32 both ways, like the reference's sample.
    void setByteIO(bool on);
    bool byteIO() const { return byteIO_; }
    void setInputBuffer(const float *data);
    void setInputBuffer(const uint8_t *data);
    CPUBuffer *getOutputBuffer();
    // the pinned host buffer the upload layer reads from (created on demand and attached to the upload layer)
    CPUBuffer *inputBuffer();
    // asynchronous operation keeps Engine::ASYNC_SLOTS input buffers; sequence s uploads from slot s % ASYNC_SLOTS
    CPUBuffer *inputBuffer(int slot);
    // device-tensor in / out (the reference's setInputTexture / getOutputTexture, stylenet_base.cpp:188-233)
    void setInputTexture(fyn_tensor *texture);
    fyn_tensor *getOutputTexture() const;
    void enableDebugOutput(const std::string &outDir);
    execstate forward() override;

    size_t weightSize() const { return totalWeights_; }
    int numLayers() const { return lastLayer_ + 1; }
    int width() const { return width_; }
    int height() const { return height_; }
    const std::unordered_map<int, size_t> &weightOffsets() const { return weightOffsets_; }

 protected:
    struct ConvSpec {
        const char *name;
        int kernel, cin, cout, scaleDiv, downsample;  // input size = (W,H)/scaleDiv
        bool fractional;
        float sourceStep;
        bool preRelu, residual, reluOnResidual;
        int fileOrder;  // position in the weight file
    };
    fyusion::fyusenet::CompiledLayers buildLayers() override;
    void connectLayers(fyusion::fyusenet::CompiledLayers &layers, fyusion::fyusenet::BufferManager *buffers) override;
    void initializeWeights(fyusion::fyusenet::CompiledLayers &layers) override;

    int kernel_, resBlocks_, width_, height_;
    bool upload_, download_;
    bool byteIO_ = false;
    std::vector<ConvSpec> convs_;                       // in layer-number order, layer number = index + CONV1
    std::unordered_map<int, size_t> weightOffsets_;     // layer number -> float offset
    size_t totalWeights_ = 0;
    int sigmoidLayer_ = 0, downloadLayer_ = 0, lastLayer_ = 0;
    std::vector<float> wbData_;
    CPUBuffer *inBuffer_ = nullptr;
    CPUBuffer *inBuffers_[fyusion::fyusenet::Engine::ASYNC_SLOTS] = {};
    fyn_tensor *inputTexture_ = nullptr;
};

class StyleNet3x3 : public StyleNetBase {
 public:
    constexpr static int STYLENET_SIZE = 77235;
    StyleNet3x3(int width, int height, bool upload, bool download,
                const fyusion::fyusenet::GfxContextLink &ctx = fyusion::fyusenet::GfxContextLink())
        : StyleNetBase(3, 2, width, height, upload, download, ctx) {}
};

class StyleNet9x9 : public StyleNetBase {
 public:
    constexpr static int STYLENET_SIZE = 169059;
    StyleNet9x9(int width, int height, bool upload, bool download,
                const fyusion::fyusenet::GfxContextLink &ctx = fyusion::fyusenet::GfxContextLink())
        : StyleNetBase(9, 5, width, height, upload, download, ctx) {}
};
