// Engine: executes the compiled layers in ascending layer-number order on the network's CUDA stream.
// Reference: fyusenet/base/engine.cpp:78-158 (setup/cleanup), :247-335 (finish/forwardLayers), :386-683 (execute),
// :208-233 (timings), :174-181,434-437,602-604 (intermediate dumps).  The synchronous path ends with a stream
// synchronise after the download layer (the reference's blocking glReadPixels); optional CUDA-graph replay
// removes per-layer launch latency for static networks.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../gpu/gfxcontextlink.h"
#include "../cpu/cpubuffer.h"
#include "compiledlayers.h"

namespace fyusion {
namespace fyusenet {

class NeuralNetwork;
namespace gpu {
class UploadLayer;
class DownloadLayer;
}

class Engine : public GfxContextTracker {
 public:
    enum execstate { EXEC_DONE = 0, EXEC_DEFERRED, EXEC_STOPPED, EXEC_ERROR };
    constexpr static int ASYNC_SLOTS = 3;       // buffers per pipeline interface == max sequences in flight
    constexpr static int MAX_IN_FLIGHT = ASYNC_SLOTS;
    struct Completion { Engine *engine; uint64_t sequence; cpu::CPUBuffer *buffer; gpu::DownloadLayer *download; };
    struct UploadNote { gpu::UploadLayer *layer; uint64_t sequence; };

    explicit Engine(const GfxContextLink &ctx = GfxContextLink(), bool async = false);
    ~Engine();
    void setup(NeuralNetwork *net);
    void cleanup();
    execstate forwardLayers();
    execstate finish();
    CompiledLayers &getLayers() { return layers_; }
    uint64_t nextSequenceNo() const { return sequenceNo_; }
    uint64_t lastSequenceNo() const { return sequenceNo_ - 1; }

    // host-side microseconds per layer number around each forward() (issue time, not device time: :201-204),
    // plus device milliseconds from CUDA events when enabled
    void enableTimings() { timings_ = true; timingOnly_ = -1; }
    // time a single layer: its event pair is the only one in the stream, so the other layers keep their dependent-launch overlap
    void enableTimings(int layerNumber) { timings_ = true; timingOnly_ = layerNumber; }
    void disableTimings() { timings_ = false; }
    void resetTimings();
    const std::unordered_map<int, uint32_t> &getTimings() const { return timingData_; }
    const std::unordered_map<int, float> &getDeviceTimings() { collectTimings(true); return deviceTimingData_; }
    int timedRuns() const { return runs_; }
    // <dir>/<layername>_<seq>.bin dumps, CHW float32 without padding
    void enableIntermediateOutput(const std::string &outputDir) { outputDir_ = outputDir; writeResults_ = true; updateFusion(); }
    void disableIntermediateOutput() { writeResults_ = false; updateFusion(); }
    // Layer fusion (conv + element-wise function in one kernel) is on by default and suspended while intermediate
    // results are written, so that every layer's dump shows that layer's own output.
    void enableFusion(bool on) { fusion_ = on; updateFusion(); }
    int fusedLayers() const { return fusedLayers_; }
    // chains of same-geometry convolutions as one persistent kernel (part of the fusion switch; separately switchable)
    void enableChains(bool on) { chainFusion_ = on; updateFusion(); }
    int chainedLayers() const { return chainedLayers_; }
    int haloExchanges() const { return (int)haloSteps_.size(); }   // exchanges per forward in row-banded operation (setHaloExchange)
    // Synchronous path: capture the device layers (everything between the upload and the download layer) into a CUDA graph on
    // the next forward and replay it afterwards; re-captured when tensor bindings, weights or fusions change.  Suspended
    // while timings, dumps or a halo exchange are active.
    void enableGraph(bool on) { useGraph_ = on; dropGraph(); }
    bool graphActive() const { return graphExec_ != nullptr; }
    // Device-resident operation of a network WITH upload / download layers: both are skipped, the upload tensor keeps the
    // data of the last real upload and the result stays in the last layer's output tensor (bench.py's `value` arm).
    void skipIO(bool on) { skipIO_ = on; }
    // Row-banded operation over several GPUs (BASELINE configs[4]): this rank's network runs on its band plus
    // `marginRows` full-resolution rows towards each neighbour; after every layer whose output feeds a layer with spatial
    // taps the margin rows of the output tensor are replaced by the neighbours' band-edge rows (fyn_halo_exchange: peer
    // stores over NVLink).  Registers the output tensors with the communicator: a collective, same order on every rank.
    void setHaloExchange(fyn_comm *comm, int marginRows, int inputHeight);
    // asynchronous operation: callbacks fired from a driver thread when a sequence's download has landed
    using DownloadCallback = std::function<void(uint64_t sequence, cpu::CPUBuffer *buffer)>;
    void setDownloadCallback(const DownloadCallback &cb) { downloadCallback_ = cb; }
    bool isAsync() const { return async_; }
    int sequencesInFlight();
    // host-side completion hook (called through fyn_stream_add_callback)
    void sequenceCompleted(uint64_t sequence, cpu::CPUBuffer *buffer);

 private:
    execstate execute(uint64_t sequence);
    execstate executeAsync(uint64_t sequence);
    // pipeline state of the asynchronous path.  The reference allows two sequences in flight (engine.cpp:314-316);
    // with three pipeline stages (upload / layers / host copy) three slots are needed to keep all of them busy.
    void *uploadDone_[ASYNC_SLOTS] = {}, *computeDone_[ASYNC_SLOTS] = {}, *copyDone_[ASYNC_SLOTS] = {};
    // FYN_ASYNC_TRACE=1: per-sequence stage boundaries (events) printed by finish(); a tuning aid
    struct TraceEntry { uint64_t seq; void *ev[6]; };
    std::vector<TraceEntry> trace_;
    void *traceBase_ = nullptr;
    void dumpTrace();
    bool slotUsed_[ASYNC_SLOTS] = {};
    Completion completions_[ASYNC_SLOTS];
    UploadNote uploadNotes_[ASYNC_SLOTS];
    std::mutex flightLock_;
    std::condition_variable flightCv_;
    int inFlight_ = 0;
    DownloadCallback downloadCallback_;
    void collectTimings(bool sync);
    struct EventPair { int layer; void *start; void *stop; };
    std::vector<EventPair> pendingEvents_;
    std::vector<void *> freeEvents_;
    CompiledLayers layers_;
    uint64_t sequenceNo_ = 1;
    bool setup_ = false;
    bool async_ = false;
    bool timings_ = false;
    int timingOnly_ = -1;
    bool writeResults_ = false;
    bool fusion_ = true;
    int fusedLayers_ = 0, chainedLayers_ = 0;
    void updateFusion();
    void updateChains(bool want);
    std::vector<fyn_conv_chain *> chains_;   // persistent multi-layer convolution kernels (fyn_conv_chain), owned here
    bool chainFusion_ = true;
    bool useGraph_ = false;
    void *graphExec_ = nullptr;
    uint64_t graphEpoch_ = 0;
    bool graphWarm_ = false;                           // one eager forward has run at the current epoch
    void dropGraph();
    bool skipIO_ = false;
    fyn_comm *haloComm_ = nullptr;
    int haloMargin_ = 0, haloInputHeight_ = 0;
    struct HaloStep { int slot; int rows; };
    std::unordered_map<int, HaloStep> haloSteps_;      // layer number -> exchange issued after that layer (the active ones)
    std::unordered_map<int, HaloStep> haloSlots_;      // layer number -> registered tensor (every candidate)
    bool haloNoChains_ = false;                        // the margin does not cover a whole chain: layers run one by one
    bool planHalo();
    std::string outputDir_;
    std::unordered_map<int, uint32_t> timingData_;
    std::unordered_map<int, float> deviceTimingData_;
    int runs_ = 0;
};

}  // namespace fyusenet
}  // namespace fyusion
