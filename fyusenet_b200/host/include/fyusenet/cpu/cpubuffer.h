// Host tensors: CPUBufferShape + CPUBuffer (upload source / download target).
// Reference: fyusenet/cpu/cpubuffershape.cpp:75-87,430-447 and cpu/cpubuffer.cpp:121-158,319-.
// Orders: CHANNELWISE = [C][H][W]; GPU_SHALLOW = [H][W][C] for uploads and [planes][H+2P][W+2P][4] for
// downloads; GPU_DEEP = tiled texture [TH][TW][4].  Buffers are pinned (cudaHostAlloc through the C ABI)
// when created with a context so that uploads / downloads are truly asynchronous -- the role PBOs play
// in the reference.
#pragma once
#include <cstdint>
#include <cstring>
#include <mutex>
#include <utility>

#include "../base/bufferspec.h"
#include "../common/fynexception.h"
#include "../gpu/gfxcontextlink.h"

namespace fyusion {
namespace fyusenet {
namespace cpu {

class CPUBuffer;

class CPUBufferShape {
 public:
    enum type : uint8_t { FLOAT32 = 0, FLOAT16, UINT8 };
    using order = BufferSpec::order;

    CPUBufferShape() = default;
    CPUBufferShape(int height, int width, int channels, int padding, type dt, order ord = order::CHANNELWISE, int batch = 1)
        : width_(width), height_(height), channels_(channels), padding_(padding), type_(dt), order_(ord), batch_(batch) {
        if (ord == order::GPU_DEEP) {
            auto t = computeDeepTiling(channels);
            tileWidth_ = t.first;
            tileHeight_ = t.second;
        }
    }
    int width() const { return width_; }
    int height() const { return height_; }
    int channels() const { return channels_; }
    int padding() const { return padding_; }
    int batch() const { return batch_; }
    type dataType() const { return type_; }
    order dataOrder() const { return order_; }
    static size_t typeSize(type t) { return t == FLOAT32 ? 4 : (t == FLOAT16 ? 2 : 1); }

    // number of elements the buffer holds in the given order
    size_t elements(order ord) const {
        int pc = 4 * ((channels_ + 3) / 4);
        size_t per;
        switch (ord) {
            case order::GPU_DEEP: {
                auto t = computeDeepTiling(channels_);
                per = (size_t)(t.first * (width_ + padding_) + padding_) * (t.second * (height_ + padding_) + padding_) * 4;
                break;
            }
            case order::GPU_SHALLOW:
                // uploads (no padding, <= 4 channels) are stored [H][W][C]; everything else as RGBA planes
                per = (padding_ == 0 && channels_ <= 4 && uploadStyle_) ? (size_t)width_ * height_ * channels_
                                                                        : (size_t)(width_ + 2 * padding_) * (height_ + 2 * padding_) * pc;
                break;
            default:
                per = (size_t)(width_ + 2 * padding_) * (height_ + 2 * padding_) * channels_;
        }
        return per * batch_;
    }
    size_t bytes(order ord) const { return elements(ord) * typeSize(type_); }
    size_t bytes() const { return bytes(order_); }
    CPUBufferShape &uploadStyle(bool on) { uploadStyle_ = on; return *this; }

    // tile arrangement for `channels` channels: minimise |x-y| + unused tiles over y <= x, first minimum
    // (reference: cpu/cpubuffershape.cpp:430-447)
    static std::pair<int, int> computeDeepTiling(int channels) {
        int tiles = (channels + 3) / 4, bx = 1, by = 1;
        long best = -1;
        for (int y = 1; y <= tiles; y++) {
            int x = (tiles + y - 1) / y;
            if (x < y) x = y;
            long cost = (long)(x - y) + (long)(x * y - tiles);
            if (best < 0 || cost < best) { best = cost; bx = x; by = y; }
        }
        return {bx, by};
    }
    CPUBuffer *createBuffer(const GfxContextLink &ctx = GfxContextLink()) const;

 private:
    int width_ = 0, height_ = 0, channels_ = 0, padding_ = 0;
    type type_ = FLOAT32;
    order order_ = order::CHANNELWISE;
    int batch_ = 1;
    int tileWidth_ = 1, tileHeight_ = 1;
    bool uploadStyle_ = true;
};

class CPUBuffer {
 public:
    explicit CPUBuffer(const CPUBufferShape &shape, const GfxContextLink &ctx = GfxContextLink()) : shape_(shape), ctx_(ctx) {
        size_t n = shape.bytes();
        if (ctx_.isValid()) {
            FYN_ABI_CALL(fyn_host_alloc(ctx_.handle(), n, &memory_));
            pinned_ = true;
        } else {
            memory_ = ::operator new(n ? n : 1);
        }
        memset(memory_, 0, n);
    }
    ~CPUBuffer() {
        if (pinned_) fyn_host_free(ctx_.handle(), memory_);
        else ::operator delete(memory_);
    }
    CPUBuffer(const CPUBuffer &) = delete;
    CPUBuffer &operator=(const CPUBuffer &) = delete;

    const CPUBufferShape &shape() const { return shape_; }
    size_t bytes() const { return shape_.bytes(); }
    bool isPinned() const { return pinned_; }
    template <typename T> T *map() { lock_.lock(); return static_cast<T *>(memory_); }
    template <typename T> const T *map() const { lock_.lock(); return static_cast<const T *>(memory_); }
    void unmap() const { lock_.unlock(); }
    void *raw() { return memory_; }
    uint64_t sequence() const { return sequence_; }
    void setSequence(uint64_t s) { sequence_ = s; }

    // deep / shallow GPU order -> [C][H][W] (reference: cpu/cpubuffer.cpp:121-158; layout ground truth is
    // unit_tests/layertestbase.cpp:235-317, see SURVEY A.1 caveat).  Returns a new unpinned buffer.
    CPUBuffer *toChannelWise() const {
        const CPUBufferShape &s = shape_;
        if (s.dataType() != CPUBufferShape::FLOAT32) THROW_EXCEPTION_ARGS(FynException, "Only float32 buffers supported");
        CPUBufferShape cw(s.height(), s.width(), s.channels(), 0, CPUBufferShape::FLOAT32, BufferSpec::order::CHANNELWISE, s.batch());
        CPUBuffer *out = new CPUBuffer(cw);
        const float *src = static_cast<const float *>(memory_);
        float *dst = static_cast<float *>(out->memory_);
        const int W = s.width(), H = s.height(), C = s.channels(), P = s.padding();
        if (s.dataOrder() == BufferSpec::order::CHANNELWISE) {
            memcpy(dst, src, out->bytes());
            return out;
        }
        auto tiling = CPUBufferShape::computeDeepTiling(C);
        const int tw = tiling.first * (W + P) + P, th = tiling.second * (H + P) + P;
        const int pw = W + 2 * P, ph = H + 2 * P, planes = (C + 3) / 4;
        for (int n = 0; n < s.batch(); n++)
            for (int c = 0; c < C; c++)
                for (int y = 0; y < H; y++)
                    for (int x = 0; x < W; x++) {
                        size_t si;
                        if (s.dataOrder() == BufferSpec::order::GPU_DEEP) {
                            int t = c / 4, ox = P + (t % tiling.first) * (W + P), oy = P + (t / tiling.first) * (H + P);
                            si = (size_t)n * tw * th * 4 + ((size_t)(oy + y) * tw + ox + x) * 4 + (c % 4);
                        } else {
                            si = ((size_t)n * planes + c / 4) * pw * ph * 4 + ((size_t)(y + P) * pw + x + P) * 4 + (c % 4);
                        }
                        dst[(((size_t)n * C + c) * H + y) * W + x] = src[si];
                    }
        return out;
    }

 private:
    CPUBufferShape shape_;
    GfxContextLink ctx_;
    void *memory_ = nullptr;
    bool pinned_ = false;
    uint64_t sequence_ = 0;
    mutable std::recursive_mutex lock_;
};

inline CPUBuffer *CPUBufferShape::createBuffer(const GfxContextLink &ctx) const { return new CPUBuffer(*this, ctx); }

// layers that read from / write to host buffers (reference: cpu/cpulayerinterface.h)
class CPULayerInterface {
 public:
    virtual ~CPULayerInterface() = default;
    virtual void setInputBuffer(CPUBuffer *buf, int port) = 0;
    virtual CPUBuffer *getInputBuffer(int port = 0) const = 0;
    virtual void addOutputBuffer(CPUBuffer *buf, int port = 0) = 0;
    virtual CPUBuffer *getOutputBuffer(int port = 0) const = 0;
    virtual bool hasOutputBuffer(int port = 0) const = 0;
    virtual void clearOutputBuffers(int port = -1) = 0;
    virtual void clearInputBuffers(int port = -1) = 0;
};

}  // namespace cpu
using cpu::CPUBuffer;
using cpu::CPUBufferShape;
}  // namespace fyusenet
}  // namespace fyusion
