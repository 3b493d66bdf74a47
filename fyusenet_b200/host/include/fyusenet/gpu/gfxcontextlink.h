// GfxContextLink / GfxContextManager on CUDA.
// In the reference these wrap an OpenGL context (fyusenet/gpu/gfxcontextlink.h, gfxcontextmanager.h);
// here a "context" is one CUDA device context of the C ABI (fyn_ctx) plus the stream the network's
// layers are enqueued on.  API names are kept so network code (`context()`, `.context(ctx)`,
// GfxContextManager::instance()->createMainContext()) ports unchanged.
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "../../../../../include/fyusenet_b200.h"
#include "../common/fynexception.h"

namespace fyusion {
namespace fyusenet {

// converts a non-zero C-ABI status into a FynException carrying fyn_last_error()
#define FYN_ABI_CALL(expr)                                                                     \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc != 0) THROW_EXCEPTION_ARGS(fyusion::FynException, "%s failed (%d): %s", #expr, _rc, fyn_last_error()); \
    } while (0)

class CudaContext {
 public:
    explicit CudaContext(int device) : device_(device) {
        FYN_ABI_CALL(fyn_cuda_init(device, &ctx_));
        FYN_ABI_CALL(fyn_stream_create(ctx_, &stream_));
    }
    ~CudaContext() {
        if (ctx_) {
            if (uploadStream_) fyn_stream_destroy(ctx_, uploadStream_);
            if (downloadStream_) fyn_stream_destroy(ctx_, downloadStream_);
            if (notifyStream_) fyn_stream_destroy(ctx_, notifyStream_);
            fyn_stream_destroy(ctx_, stream_);
            fyn_cuda_shutdown(ctx_);
        }
    }
    // side streams of the asynchronous engine (the role of the AsyncPool's shared GL contexts, gl/asyncpool.h:28-50)
    void *uploadStream() {
        if (!uploadStream_) FYN_ABI_CALL(fyn_stream_create(ctx_, &uploadStream_));
        return uploadStream_;
    }
    void *downloadStream() {
        if (!downloadStream_) FYN_ABI_CALL(fyn_stream_create(ctx_, &downloadStream_));
        return downloadStream_;
    }
    // carries only the host notifications of completed downloads: a host function blocks its stream until the driver
    // thread has run it (~0.4 ms under load on the B200 boxes), which must not delay the next device-to-host copy
    void *notifyStream() {
        if (!notifyStream_) FYN_ABI_CALL(fyn_stream_create(ctx_, &notifyStream_));
        return notifyStream_;
    }
    CudaContext(const CudaContext &) = delete;
    CudaContext &operator=(const CudaContext &) = delete;
    fyn_ctx *handle() const { return ctx_; }
    void *stream() const { return stream_; }
    void setStream(void *s) { externalStream_ = s; useExternal_ = true; }
    void *activeStream() const { return useExternal_ ? externalStream_ : stream_; }
    int device() const { return device_; }

 private:
    int device_ = 0;
    fyn_ctx *ctx_ = nullptr;
    void *stream_ = nullptr;
    void *uploadStream_ = nullptr, *downloadStream_ = nullptr, *notifyStream_ = nullptr;
    void *externalStream_ = nullptr;
    bool useExternal_ = false;
};

class GfxContextLink {
 public:
    GfxContextLink() = default;
    explicit GfxContextLink(std::shared_ptr<CudaContext> c) : ctx_(std::move(c)) {}
    bool isValid() const { return (bool)ctx_; }
    CudaContext *interface() const { return ctx_.get(); }
    fyn_ctx *handle() const {
        if (!ctx_) THROW_EXCEPTION_ARGS(FynException, "No (CUDA) context linked");
        return ctx_->handle();
    }
    void *stream() const { return ctx_ ? ctx_->activeStream() : nullptr; }
    int device() const { return ctx_ ? ctx_->device() : -1; }

 private:
    std::shared_ptr<CudaContext> ctx_;
};

class GfxContextManager {
 public:
    // one manager per device (the reference keeps one per display/device as well)
    static std::shared_ptr<GfxContextManager> instance(int device = 0);
    GfxContextLink createMainContext() {
        std::lock_guard<std::mutex> lck(lock_);
        if (!main_) main_ = std::make_shared<CudaContext>(device_);
        return GfxContextLink(main_);
    }
    GfxContextLink getMain() const { return GfxContextLink(main_); }
    // PBO pools of the reference (setupPBOPools) have no equivalent: pinned buffers are owned by CPUBuffer
    void setupPBOPools(int, int) {}
    void tearDown() { main_.reset(); }

 private:
    explicit GfxContextManager(int device) : device_(device) {}
    int device_;
    std::mutex lock_;
    std::shared_ptr<CudaContext> main_;
};

// "the GL context must be current" bookkeeping of the reference (gpu/gfxcontexttracker.h) reduces to
// remembering the link; assertContext() checks that one exists.
class GfxContextTracker {
 public:
    void setContext(const GfxContextLink &ctx) { context_ = ctx; }
    const GfxContextLink &context() const { return context_; }
    void assertContext() const {
        if (!context_.isValid()) THROW_EXCEPTION_ARGS(FynException, "No valid (CUDA) context");
    }

 protected:
    GfxContextLink context_;
};

}  // namespace fyusenet
}  // namespace fyusion
