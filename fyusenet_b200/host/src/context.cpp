// Context manager singleton + storage-precision switch.
#include <cstdlib>
#include <cstring>
#include <map>

#include "fyusenet/gpu/gfxcontextlink.h"
#include "fyusenet/gpu/gpulayerbase.h"

namespace fyusion {
namespace fyusenet {

std::shared_ptr<GfxContextManager> GfxContextManager::instance(int device) {
    static std::mutex lock;
    static std::map<int, std::shared_ptr<GfxContextManager>> managers;
    std::lock_guard<std::mutex> lck(lock);
    auto it = managers.find(device);
    if (it != managers.end()) return it->second;
    std::shared_ptr<GfxContextManager> mgr(new GfxContextManager(device));
    managers[device] = mgr;
    return mgr;
}

namespace gpu {

// FYN_STORAGE=fp32 mirrors the reference's HIGH_PRECISION build option (CMakeLists.txt:21-27,
// gpu/gpulayerbase.h:100-110); default is fp16 activations like the reference's RGBA16F textures.
static BufferSpec::dtype initialPrecision() {
    const char *env = getenv("FYN_STORAGE");
    if (env && (!strcmp(env, "fp32") || !strcmp(env, "f32") || !strcmp(env, "FP32"))) return BufferSpec::FLOAT32;
    return BufferSpec::FLOAT16;
}
static BufferSpec::dtype g_precision = initialPrecision();
BufferSpec::dtype storagePrecision() { return g_precision; }
void setStoragePrecision(BufferSpec::dtype dt) { g_precision = dt; }

}  // namespace gpu
}  // namespace fyusenet
}  // namespace fyusion
