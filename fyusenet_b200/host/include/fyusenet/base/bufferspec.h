// BufferSpec: what a layer needs on an input port / produces on its output.
// Reference: fyusenet/base/bufferspec.h:59-368.  GL texture formats are replaced by the device-tensor
// properties of the C ABI (order, dtype, packing); one spec describes a WHOLE tensor (all channel planes)
// instead of one 4-channel texture.
#pragma once
#include <cstdint>

namespace fyusion {
namespace fyusenet {

struct BufferSpec {
    enum usage : uint8_t { CONVOLUTION_SOURCE = 0, CONVOLUTION_DEST, FUNCTION_SOURCE, FUNCTION_DEST, RESIDUAL_SOURCE,
                           GPU_DEST /* upload target */, CPU_SOURCE, CPU_DEST };
    // UBYTE: host-side data of upload / download layers only (reference: BufferSpec::UBYTE, gpu/uploadlayer.cpp:51-66)
    enum dtype : uint8_t { FLOAT16 = 0, FLOAT32 = 1, FLOAT = 1, UBYTE = 2, UINT8 = 2 };
    enum class order : uint8_t { CHANNELWISE = 0, GPU_SHALLOW, GPU_DEEP };
    enum csdevice : uint8_t { COMP_STOR_GPU = 0, COMP_STOR_CPU };

    BufferSpec() = default;
    BufferSpec(int port, int netWidth, int netHeight, int channels, int padding, order ord, dtype dt, usage us)
        : port_(port), width_(netWidth), height_(netHeight), channels_(channels), padding_(padding), dataOrder_(ord),
          type_(dt), usage_(us) {}

    BufferSpec &device(csdevice d) { device_ = d; return *this; }
    BufferSpec &packing(int p) { packing_ = p; return *this; }
    BufferSpec &async(bool a) { async_ = a; return *this; }
    BufferSpec &multi(int m) { multiplicity_ = m; return *this; }
    BufferSpec &lock(bool l) { lock_ = l; return *this; }
    BufferSpec &anyType() { anyType_ = true; return *this; }

    int port_ = 0;
    int width_ = 0, height_ = 0;   // net size (without padding)
    int channels_ = 0;
    int padding_ = 0;
    order dataOrder_ = order::GPU_SHALLOW;
    dtype type_ = FLOAT16;
    usage usage_ = FUNCTION_SOURCE;
    csdevice device_ = COMP_STOR_GPU;
    int packing_ = 4;              // channels per texel (3 for the RGB32F upload texture)
    bool anyType_ = false;         // consumer accepts any dtype / packing (e.g. reads an upload texture)
    bool async_ = false;
    int multiplicity_ = 1;         // shadow buffers for asynchronous producers
    bool lock_ = false;            // never reuse
};

}  // namespace fyusenet
}  // namespace fyusion
