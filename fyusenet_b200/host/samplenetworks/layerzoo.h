// LayerZoo: a small network that chains every bandwidth-bound layer outside the two sample networks
// through the builders, the factory and the buffer manager exactly as an application would:
//   upload -> RGB2BGR -> SCALE2D (x2, linear) -> SINGLETON_ARITH (*2) -> SUB -> CONCAT (3+3+3) -> CLIP -> SHALLOW2DEEP
//   -> deep SCALE2D (/2, nearest) -> DEEP2SHALLOW -> PADDING2D -> ADD -> download
// It has no weights; tests compare every layer with the oracle chain (tests/test_gpu_networks.py).
#pragma once
#include <fyusenet/fyusenet.h>

class LayerZoo : public fyusion::fyusenet::NeuralNetwork {
 public:
    using CPUBuffer = fyusion::fyusenet::cpu::CPUBuffer;
    enum { UPLOAD = 0, BGR, UPSCALE, TWICE, DIFF, CONCAT, CLIP, TODEEP, DOWNSCALE, TOSHALLOW, PAD, SUM, DOWNLOAD };
    static constexpr float CLIP_LOW = 0.2f, CLIP_HIGH = 0.7f;

    LayerZoo(int width, int height, const fyusion::fyusenet::GfxContextLink &ctx = fyusion::fyusenet::GfxContextLink());
    ~LayerZoo() override;
    void setInputBuffer(const float *data);
    CPUBuffer *inputBuffer();
    CPUBuffer *getOutputBuffer();

 protected:
    fyusion::fyusenet::CompiledLayers buildLayers() override;
    void connectLayers(fyusion::fyusenet::CompiledLayers &layers, fyusion::fyusenet::BufferManager *buffers) override;
    void initializeWeights(fyusion::fyusenet::CompiledLayers &) override {}

    int width_, height_;
    CPUBuffer *inBuffer_ = nullptr;
};
