// Layer flags, activation / normalisation identifiers and layer types.
// The numeric values are part of the drop-in contract and equal those of the reference
// (fyusenet/base/layerflags.h:33-192); include/fyusenet_b200.h uses the same flag bits.
#pragma once
#include <cstdint>

namespace fyusion {
namespace fyusenet {

using layerflags = uint32_t;

namespace LayerFlags {
constexpr layerflags NO_LAYER_FLAGS = 0;
constexpr layerflags RESIDUAL_INPUT = 1u << 0;         // a second tensor is added to the layer result
constexpr layerflags RELU_ON_RESIDUAL = 1u << 1;       // ReLU the residual when it is fetched
constexpr layerflags BATCHNORM_ON_RESIDUAL = 1u << 2;  // post-BN scale also multiplies the residual
constexpr layerflags POST_BATCHNORM = 1u << 3;         // scale/bias applied when the result is written
constexpr layerflags DEEP = 1u << 4;                   // deep (tiled) tensor layout
constexpr layerflags POST_RELU = 1u << 5;              // not supported by GPU layers
constexpr layerflags PRE_RELU = 1u << 6;               // (leaky) ReLU applied when the input is fetched
constexpr layerflags PRE_CLIP = 1u << 7;               // clip applied when the input is fetched
constexpr layerflags PRE_SIGMOID = 1u << 8;            // declared, not implemented (as in the reference)
constexpr layerflags PRE_TANH = 1u << 9;               // declared, not implemented (as in the reference)
constexpr layerflags PRE_ACT_MASK = PRE_RELU | PRE_CLIP | PRE_SIGMOID | PRE_TANH;
constexpr layerflags ACT_MASK = PRE_ACT_MASK | POST_RELU;
}  // namespace LayerFlags

enum class ActType : uint8_t { NONE = 0, RELU = 1, LEAKY_RELU, CLIP, SIGMOID, TANH };
enum class NormType : uint8_t { NONE = 0, BATCHNORM = 1 };
enum class ScalingType : uint8_t { NEAREST = 0, LINEAR };
enum class ArithType : uint8_t { ADD = 0, SUB, MUL, DIV };

enum class LayerType : uint16_t {
    ADD = 1, SUB, ARGMAX, CAST, CONCAT, CONVOLUTION2D, FRACCONVOLUTION2D, TRANSCONVOLUTION2D, AVGPOOL2D,
    MAXPOOL2D, PADDING2D, SCALE2D, SINGLETON_ARITH, RELU, CLIP, TANH, SIGMOID, REDUCE, TRANSPOSE, IMGEXTRACT,
    BLUR2D, NONMAX2D, RGB2BGR, DEEP2SHALLOW, SHALLOW2DEEP, DOWNLOAD, UPLOAD, RESIDUAL, OESCONV, BATCHNORM,
    GEMM, CUSTOM, LAST_SUPPORTED, ILLEGAL = 1000
};

enum class compute_device : uint8_t { DEV_GPU = 0, DEV_CPU, DEV_NPU, DEV_ILLEGAL };

namespace gpu {
static constexpr int PIXEL_PACKING = 4;  // channels per texel
}

}  // namespace fyusenet
}  // namespace fyusion
