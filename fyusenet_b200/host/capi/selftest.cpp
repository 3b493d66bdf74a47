// selftest.cpp -- device-free checks of the host engine, callable from the CPU test-suite through
// fynhost_selftest().  They mirror what the reference asserts implicitly in its builders / factory
// (base/layerbuilder.h:445-493, base/layerfactory.cpp:67-101, base/compiledlayers.h, cpu/cpubuffer.cpp).
#include <cstring>
#include <sstream>
#include <string>

#include <fyusenet/fyusenet.h>

using namespace fyusion;
using namespace fyusion::fyusenet;

namespace {

struct Report {
    std::ostringstream out;
    int failures = 0;
    void check(bool ok, const char *what) {
        out << (ok ? "ok   " : "FAIL ") << what << "\n";
        if (!ok) failures++;
    }
};

// a layer that records what the factory handed it
class ProbeLayer : public LayerBase {
 public:
    ProbeLayer(const LayerBuilderData &b, int no) : LayerBase(b, no) {}
    void setup() override { valid_ = true; }
    void cleanup() override { valid_ = false; }
    void forward(uint64_t) override {}
    std::vector<BufferSpec> getRequiredInputBuffers() const override { return {}; }
    std::vector<BufferSpec> getRequiredOutputBuffers() const override { return {}; }
    void writeResult(const char *, bool) override {}
};

class ProbeBackend : public LayerFactoryBackend {
 public:
    std::string getName() const override { return "probe"; }
    LayerBase *createLayer(LayerType, LayerBuilder *builder, int layerNumber) override {
        return new ProbeLayer(*reinterpret_cast<LayerBuilderData *>(builder), layerNumber);
    }
};

template <typename F>
bool throws(F &&f) {
    try {
        f();
    } catch (const FynException &) {
        return true;
    }
    return false;
}

}  // namespace

extern "C" int fynhost_selftest(char *report, int cap) {
    Report r;
    // ---- builder -> flags (layerbuilder.h:445-493)
    {
        gpu::ConvLayerBuilder b(3, "c");
        b.shape(40, 16, 24, 20).type(LayerType::CONVOLUTION2D).prefixAct(ActType::RELU).residual(ActType::RELU).number(3);
        layerflags f = b.getFlags();
        r.check(f == (LayerFlags::PRE_RELU | LayerFlags::RESIDUAL_INPUT | LayerFlags::RELU_ON_RESIDUAL), "relu + relu-residual flags");
        r.check(b.width() == 24 && b.height() == 16 && b.in() == 20 && b.out() == 40, "shape(out,h,w,in) argument order");
        r.check(b.kernel_ == 3 && b.device_ == compute_device::DEV_GPU, "conv builder defaults");
    }
    {
        gpu::ConvLayerBuilder b(1, "c");
        b.shape(256, 56, 56, 64).deep().postfixNorm(NormType::BATCHNORM).residual(ActType::NONE, true).prefixAct(ActType::LEAKY_RELU).leakyReLU(0.1f);
        layerflags f = b.getFlags();
        r.check(f == (LayerFlags::DEEP | LayerFlags::POST_BATCHNORM | LayerFlags::RESIDUAL_INPUT | LayerFlags::BATCHNORM_ON_RESIDUAL | LayerFlags::PRE_RELU),
                "deep + postBN + BN-on-residual + leaky flags");
        b.residual(ActType::RELU);
        b.residual(ActType::NONE);
        r.check(!(b.getFlags() & LayerFlags::RELU_ON_RESIDUAL), "residual(NONE) clears RELU_ON_RESIDUAL");
        r.check(throws([&] { b.residual(ActType::CLIP); }), "residual(CLIP) throws");
        gpu::GPULayerBuilder g("g");
        g.prefixAct(ActType::SIGMOID);
        r.check(throws([&] { g.getFlags(); }), "prefix SIGMOID not supported yet");
        g.prefixAct(ActType::CLIP).clip(-1.f, 2.f);
        r.check(g.getFlags() == LayerFlags::PRE_CLIP && g.clipLow_ == -1.f && g.clipHigh_ == 2.f, "clip activation");
    }
    {
        gpu::PoolLayerBuilder p(gpu::PoolLayerBuilder::POOL_AVG, "p");
        r.check(throws([&] { p.global(); }), "global() before size throws");
        p.shape(2048, 7, 7, 2048).global();
        r.check(p.global_ && p.downsample_[0] == 7 && p.downsample_[1] == 7, "global pooling sets downsample to the spatial size");
        gpu::UpDownLayerBuilder u(gpu::UpDownLayerBuilder::UPLOAD, "u"), d(gpu::UpDownLayerBuilder::DOWNLOAD, "d");
        r.check(u.type_ == LayerType::UPLOAD && d.type_ == LayerType::DOWNLOAD, "up/download builders set their type");
    }
    // ---- factory / compiled layers (layerfactory.cpp:67-101, compiledlayers.h)
    {
        std::shared_ptr<LayerFactory> factory = LayerFactory::withBackend(new ProbeBackend());
        r.check(factory->getName() == "probe", "foreign backend plugs into the factory");
        const int numbers[] = {7, 2, 5};
        const char *names[] = {"seven", "two", "five"};
        for (int i = 0; i < 3; i++) {
            auto *b = new gpu::GPULayerBuilder(names[i]);
            b->shape(8, 4, 4, 8).type(LayerType::SIGMOID).number(numbers[i]);
            b->push(factory);
        }
        auto *dup = new gpu::GPULayerBuilder("dup");
        dup->shape(8, 4, 4, 8).type(LayerType::SIGMOID).number(5);
        r.check(throws([&] { dup->push(factory); }), "duplicate layer number is refused");
        delete dup;
        auto *untyped = new gpu::GPULayerBuilder("untyped");
        untyped->shape(8, 4, 4, 8).number(9);
        r.check(throws([&] { untyped->push(factory); }), "builder without type is refused");
        delete untyped;
        CompiledLayers layers = factory->compileLayers();
        std::string order;
        for (auto it = layers.begin(); it != layers.end(); ++it) order += std::to_string(it.first) + ":" + it.second->getName() + " ";
        r.check(order == "2:two 5:five 7:seven ", "layers iterate in ascending layer-number order");
        r.check(layers[5] && layers[5]->getName() == "five" && layers["seven"] && layers["seven"]->getNumber() == 7 && !layers[3] && !layers["x"],
                "lookup by number and by name");
        auto *cpuB = new LayerBuilder("cpu");
        cpuB->shape(8, 4, 4, 8).type(LayerType::SIGMOID).number(11);
        std::shared_ptr<LayerFactory> f2 = LayerFactory::withBackend(new ProbeBackend());
        cpuB->push(f2);
        r.check(throws([&] { f2->compileLayers(); }), "non-GPU layers are refused (no CPU fallback)");
    }
    {
        gpu::CUDALayerFactoryBackend backend;
        gpu::GPULayerBuilder t("tanh");
        t.shape(8, 4, 4, 8).type(LayerType::TANH).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::TANH, reinterpret_cast<LayerBuilder *>(&t), 1); }), "out-of-scope layer type throws");
        gpu::GPULayerBuilder wrong("conv");
        wrong.shape(8, 4, 4, 8).type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&wrong), 1); }),
                "conv needs a ConvLayerBuilder");
        gpu::ConvLayerBuilder noctx(3, "noctx");
        noctx.shape(8, 4, 4, 8).type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&noctx), 1); }),
                "layer without a context throws");
    }
    // ---- scale / concat / singleton builders (gpu/scalelayerbuilder.h:60-95, concatlayerbuilder.h:45-55)
    {
        gpu::ScaleLayerBuilder up("up");
        up.scale(2.0f, 3.0f);
        r.check(up.type_ == LayerType::SCALE2D && up.upsample_[0] == 2 && up.upsample_[1] == 3 && up.downsample_[0] == 1 && !up.equal(),
                "ScaleLayerBuilder::scale(2,3) -> integer upsample factors");
        gpu::ScaleLayerBuilder dn("dn");
        dn.scale(0.5f).scaleType(ScalingType::LINEAR);
        r.check(dn.downsample_[0] == 2 && dn.downsample_[1] == 2 && dn.equal() && dn.scaleType_ == ScalingType::LINEAR, "scale(0.5) -> downsample 2");
        gpu::ScaleLayerBuilder third("third");
        third.scale(1.0f / 3.0f);
        r.check(third.downsample_[0] == 3, "scale(1/3) -> downsample 3");
        r.check(throws([&] { gpu::ScaleLayerBuilder("x").scale(1.5f); }), "fractional upscale throws");
        r.check(throws([&] { gpu::ScaleLayerBuilder("x").scale(0.4f); }), "non-integer downscale throws");
        gpu::ConcatLayerBuilder cat("cat");
        cat.input(3, 1).input(8, 1, LayerFlags::PRE_RELU);
        r.check(cat.type_ == LayerType::CONCAT && cat.inputs_.size() == 2 && cat.inputChannels_ == 11 && cat.inputs_[1].flags == LayerFlags::PRE_RELU,
                "ConcatLayerBuilder collects its inputs");
        gpu::SingletonArithLayerBuilder mul("mul", ArithType::MUL);
        mul.operand(2.5f);
        r.check(mul.type_ == LayerType::SINGLETON_ARITH && mul.opType_ == ArithType::MUL && mul.operand_ == 2.5f, "SingletonArithLayerBuilder");
        gpu::CUDALayerFactoryBackend backend;
        gpu::GPULayerBuilder plain("cat2");
        plain.shape(8, 4, 4, 8).type(LayerType::CONCAT).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONCAT, reinterpret_cast<LayerBuilder *>(&plain), 1); }), "concat needs a ConcatLayerBuilder");
        // grouped convolutions: only depthwise 3x3 with channel multiplier 1 (gpu/gpulayerfactory.cpp:358-396)
        gpu::ConvLayerBuilder grouped(3, "grouped");
        grouped.groupSize(4).shape(8, 4, 4, 8).type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&grouped), 1); }), "grouped (non-depthwise) convolution throws");
        gpu::ConvLayerBuilder tc1(3, "tc1");
        tc1.shape(8, 4, 4, 8).type(LayerType::TRANSCONVOLUTION2D).number(1);   // upsample defaults to 1
        r.check(throws([&] { backend.createLayer(LayerType::TRANSCONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&tc1), 1); }), "transpose convolution needs stride 2");
        gpu::ConvLayerBuilder dw5(5, "dw5");
        dw5.groupSize(8).shape(8, 4, 4, 8).type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&dw5), 1); }), "5x5 depthwise convolution throws");
        // channel multipliers: deep layers with input channels % 4 == 0 only (convlayer_dw_3x3_vanilla.cpp:49-50, deepdwconvlayerbase.cpp:40-44)
        gpu::ConvLayerBuilder dwm(3, "dwm");
        dwm.groupSize(8).shape(16, 4, 4, 8).type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&dwm), 1); }), "shallow depthwise convolution with a channel multiplier throws");
        gpu::ConvLayerBuilder dwm6(3, "dwm6");
        dwm6.groupSize(6).shape(12, 4, 4, 6).deep().type(LayerType::CONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::CONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&dwm6), 1); }), "deep channel multiplier needs input channels % 4 == 0");
        gpu::ConvLayerBuilder tcr(3, "tcr");
        tcr.shape(8, 4, 4, 8).deep().upsample(2).residual(ActType::NONE).type(LayerType::TRANSCONVOLUTION2D).number(1);
        r.check(throws([&] { backend.createLayer(LayerType::TRANSCONVOLUTION2D, reinterpret_cast<LayerBuilder *>(&tcr), 1); }), "transpose convolution with a residual input throws");
    }
    // ---- host tensors (cpubuffershape.cpp:430-447, cpubuffer.cpp:121-158)
    {
        r.check(cpu::CPUBufferShape::computeDeepTiling(64) == std::make_pair(4, 4), "deep tiling 64 ch -> 4x4");
        r.check(cpu::CPUBufferShape::computeDeepTiling(1000) == std::make_pair(18, 14), "deep tiling 1000 ch -> 18x14");
        r.check(cpu::CPUBufferShape::computeDeepTiling(3) == std::make_pair(1, 1), "deep tiling 3 ch -> 1x1");
        cpu::CPUBufferShape up(6, 5, 3, 0, cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_SHALLOW);
        r.check(up.bytes() == 6 * 5 * 3 * 4, "upload buffer is [H][W][3] float32");
        cpu::CPUBufferShape deep(1, 1, 1000, 0, cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_DEEP);
        r.check(deep.bytes() == 18 * 14 * 4 * 4, "1x1x1000 deep buffer is 18x14 RGBA texels");
        // deep -> channel-wise with padding and several tiles
        const int C = 10, H = 3, W = 2, P = 1;
        cpu::CPUBufferShape ds(H, W, C, P, cpu::CPUBufferShape::FLOAT32, BufferSpec::order::GPU_DEEP);
        cpu::CPUBuffer buf(ds);
        auto tl = cpu::CPUBufferShape::computeDeepTiling(C);
        int tw = tl.first * (W + P) + P;
        float *m = buf.map<float>();
        for (int c = 0; c < C; c++)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int t = c / 4, ox = P + (t % tl.first) * (W + P), oy = P + (t / tl.first) * (H + P);
                    m[((size_t)(oy + y) * tw + ox + x) * 4 + (c % 4)] = (float)(c * 100 + y * 10 + x);
                }
        buf.unmap();
        cpu::CPUBuffer *cw = buf.toChannelWise();
        const float *q = cw->map<float>();
        bool ok = true;
        for (int c = 0; c < C; c++)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) ok &= q[(c * H + y) * W + x] == (float)(c * 100 + y * 10 + x);
        cw->unmap();
        delete cw;
        r.check(ok, "deep -> channel-wise conversion");
    }
    std::string text = r.out.str();
    if (report && cap > 0) snprintf(report, cap, "%s", text.c_str());
    return r.failures;
}
