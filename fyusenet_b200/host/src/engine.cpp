// Engine + NeuralNetwork.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include "fyusenet/base/engine.h"

#include <cstdio>

#include "fyusenet/base/neuralnetwork.h"
#include "fyusenet/common/performance.h"
#include "fyusenet/gpu/cudalayers.h"

// NVTX ranges per layer (the reference brackets every layer with GL debug groups / timers, base/engine.cpp:208-233,443-449):
// header-only NVTX 3, costs nothing unless a tool is attached; FYN_NVTX=0 switches the calls off entirely
#include <nvtx3/nvToolsExt.h>

namespace fyusion {
namespace fyusenet {

Engine::Engine(const GfxContextLink &ctx, bool async) : async_(async) { setContext(ctx); }

Engine::~Engine() {}

void Engine::setup(NeuralNetwork *net) {
    if (!net) THROW_EXCEPTION_ARGS(FynException, "Null network");
    assertContext();
    layers_ = net->glSetup();
    setup_ = true;
    updateFusion();
}

// Fuses conv -> sigmoid pairs (the only consumer of the convolution is a plain sigmoid layer): the convolution's
// epilogue evaluates the function and writes the sigmoid layer's output tensor, the sigmoid layer is bypassed.
void Engine::updateFusion() {
    if (!setup_) return;
    if (async_) finish();
    const bool want = fusion_ && !writeResults_;
    fusedLayers_ = 0;
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        auto *conv = dynamic_cast<gpu::ConvLayerBase *>(it.second);
        if (!conv) continue;
        const auto &recv = conv->receivers();
        auto *sig = recv.size() == 1 ? dynamic_cast<gpu::SigmoidLayer *>(recv[0].first) : nullptr;
        if (!sig) continue;
        conv->unfuse();
        sig->setBypass(false);
        // (consecutive layer numbers: no layer in between can still read the pooled tensor the sigmoid layer writes)
        if (want && sig->getNumber() == conv->getNumber() + 1 && sig->plainFunction() && sig->hasOutputTexture(0) && conv->fuseFunction(FYN_EPILOGUE_SIGMOID, sig->getOutputTexture(0))) {
            sig->setBypass(true);
            fusedLayers_++;
        }
    }
    // batch-norm -> convolution pairs (the batch-norm layer's only consumer reads it on port 0): the convolution fetches the
    // batch-norm layer's input and normalises at the fetch (ResNet-50: BN9 ... BN66 in front of the 1x1 reduce convolutions)
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        auto *bn = dynamic_cast<gpu::BatchNormLayer *>(it.second);
        if (!bn) continue;
        const auto &recv = bn->receivers();
        auto *conv = (recv.size() == 1 && recv[0].second == 0) ? dynamic_cast<gpu::ConvLayerBase *>(recv[0].first) : nullptr;
        if (!conv) continue;
        conv->unfuseInput();
        bn->setBypass(nullptr);
        // (consecutive layer numbers: nothing can run, and reuse the batch-norm layer's input tensor, between the two)
        if (want && conv->getNumber() == bn->getNumber() + 1 && bn->plainFunction() && bn->hasInputTexture(0) && !bn->parameters().empty() &&
            conv->fuseInputNorm(bn->parameters().data(), bn->getInputTexture(0))) {
            bn->setBypass(conv);
            fusedLayers_++;
        }
    }
    // (row-banded operation: chains stay as long as the margin covers a whole chain's taps, see planHalo)
    updateChains(want && chainFusion_ && !(haloComm_ && haloNoChains_) && getenv("FYN_NO_CHAIN") == nullptr);
    if (haloComm_ && !haloSlots_.empty() && !planHalo()) {
        haloNoChains_ = true;
        updateChains(false);
        if (!planHalo()) THROW_EXCEPTION_ARGS(FynException, "Halo margin of %d rows does not cover the taps of a single layer", haloMargin_);
    }
}

// Runs of consecutive shallow convolutions of identical geometry (StyleNet: res1_1 ... res5_2) -> one persistent kernel
// (fyn_conv_chain).  A run qualifies when layer i+1 reads layer i's output on port 0, a residual input is the tensor the
// previous layer of the run reads, layer numbers are consecutive and no tensor inside the run has a reader outside of it.
// fyn_conv_chain_create decides whether the kernels can do it; everything else keeps running layer by layer.
void Engine::updateChains(bool want) {
    for (auto it = layers_.begin(); it != layers_.end(); ++it)
        if (auto *conv = dynamic_cast<gpu::ConvLayerBase *>(it.second)) conv->unchain();
    for (fyn_conv_chain *c : chains_) fyn_conv_chain_destroy(c);
    chains_.clear();
    chainedLayers_ = 0;
    if (!want) return;
    std::vector<gpu::ConvLayerBase *> convs;
    for (auto it = layers_.begin(); it != layers_.end(); ++it) convs.push_back(dynamic_cast<gpu::ConvLayerBase *>(it.second));
    size_t i = 0;
    while (i < convs.size()) {
        gpu::ConvLayerBase *head = convs[i];
        if (!head || !head->op() || head->fused() || head->inputFused() || !head->hasInputTexture(0) || !head->hasOutputTexture(0) ||
            (head->getFlags() & LayerFlags::RESIDUAL_INPUT)) {
            i++;
            continue;
        }
        std::vector<gpu::ConvLayerBase *> run{head};
        std::vector<int> resFrom{-2};
        while (i + run.size() < convs.size()) {
            gpu::ConvLayerBase *prev = run.back(), *next = convs[i + run.size()];
            if (!next || !next->op() || next->fused() || next->inputFused() || next->getNumber() != prev->getNumber() + 1 || !next->hasInputTexture(0) ||
                !next->hasOutputTexture(0) || next->getInputTexture(0) != prev->getOutputTexture(0))
                break;
            {
                // same geometry as the head (fyn_conv_chain_create checks the rest: plans, flags, padding)
                const fyn_conv_desc &a = head->descriptor(), &b = next->descriptor();
                if (a.width != b.width || a.height != b.height || a.in_channels != b.in_channels || a.out_channels != b.out_channels || a.kernel != b.kernel ||
                    a.downsample != b.downsample || a.fractional != b.fractional || a.dilation != b.dilation)
                    break;
            }
            int rf = -2;
            if (next->getFlags() & LayerFlags::RESIDUAL_INPUT) {
                // the residual must be the tensor the previous layer of the run reads
                if (next->residualTexture() != prev->getInputTexture(0)) break;
                rf = (int)run.size() - 2;
            }
            run.push_back(next);
            resFrom.push_back(rf);
        }
        // Longest prefix of the run that (a) keeps every tensor inside it private -- no layer outside may read the output of
        // any layer but the last -- and (b) the chain kernel accepts.
        auto closed = [&](size_t len) {
            for (size_t k = 0; k + 1 < len; k++)
                for (auto &rcv : run[k]->receivers()) {
                    bool inside = false;
                    for (size_t m = 0; m < len; m++) inside = inside || run[m] == rcv.first;
                    if (!inside) return false;
                }
            return true;
        };
        bool done = false;
        for (size_t len = run.size(); len >= 2 && !done; len--) {
            if (!closed(len)) continue;
            std::vector<fyn_op *> ops;
            for (size_t k = 0; k < len; k++) ops.push_back(run[k]->op());
            fyn_conv_chain *chain = nullptr;
            if (fyn_conv_chain_create(context_.handle(), ops.data(), resFrom.data(), (int)len, &chain) == 0) {
                chains_.push_back(chain);
                head->setChainHead(chain, std::vector<gpu::ConvLayerBase *>(run.begin() + 1, run.begin() + (long)len));
                for (size_t k = 1; k < len; k++) run[k]->setChainMember(true);
                chainedLayers_ += (int)len;
                i += len;
                done = true;
            }
        }
        if (!done) i++;
    }
}

void Engine::dropGraph() {
    if (graphExec_) fyn_graph_destroy(context_.handle(), graphExec_);
    graphExec_ = nullptr;
}

void Engine::setHaloExchange(fyn_comm *comm, int marginRows, int inputHeight) {
    if (!setup_) THROW_EXCEPTION_ARGS(FynException, "setHaloExchange() needs a network that has been set up");
    dropGraph();
    haloSteps_.clear();
    haloSlots_.clear();
    haloNoChains_ = false;
    haloComm_ = comm;
    haloMargin_ = marginRows;
    haloInputHeight_ = inputHeight;
    if (!comm) {
        updateFusion();
        return;
    }
    if (async_) THROW_EXCEPTION_ARGS(FynException, "Row-banded operation is synchronous");
    if (marginRows <= 0 || marginRows % 4 || inputHeight <= 0)
        THROW_EXCEPTION_ARGS(FynException, "Halo margin must be a positive multiple of 4 full-resolution rows (got %d)", marginRows);
    // every tensor whose margins may have to be refreshed is registered once (a collective over the communicator: the same
    // sequence on every rank); planHalo() then picks the exchanges that are needed
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        auto *g = dynamic_cast<gpu::GPULayerBase *>(it.second);
        if (!g || dynamic_cast<gpu::UploadLayer *>(g) || dynamic_cast<gpu::DownloadLayer *>(g) || !g->hasOutputTexture(0)) continue;
        // only outputs that a layer with spatial taps reads need their margins refreshed
        bool spatial = false;
        for (auto &rcv : g->receivers())
            if (!dynamic_cast<gpu::SigmoidLayer *>(rcv.first) && !dynamic_cast<gpu::DownloadLayer *>(rcv.first)) spatial = true;
        if (!spatial) continue;
        fyn_tensor_desc d{};
        FYN_ABI_CALL(fyn_tensor_get_desc(g->getOutputTexture(0), &d, nullptr));
        if (d.order != FYN_ORDER_SHALLOW) THROW_EXCEPTION_ARGS(FynException, "Layer %s: row bands need shallow tensors", g->getName().c_str());
        // this tensor's resolution relative to the network input: margin rows scale with it
        const long long num = (long long)marginRows * d.height;
        if (num % inputHeight) THROW_EXCEPTION_ARGS(FynException, "Layer %s: margin %d does not map to whole rows at height %d / %d", g->getName().c_str(), marginRows, d.height, inputHeight);
        HaloStep st{};
        st.rows = (int)(num / inputHeight);
        FYN_ABI_CALL(fyn_comm_register_tensor(comm, g->getOutputTexture(0), &st.slot));
        haloSlots_[it.first] = st;
    }
    updateFusion();            // (re)builds the chains and plans the exchanges
}

// Which margins have to be refreshed, and when.  A band carries `haloMargin_` full-resolution rows of its neighbours on either
// side; they are exact after an exchange (or in the uploaded input) and every layer with spatial taps spoils the outermost
// ones -- the texture edge of a band is not the image edge.  `valid` follows, per tensor, how many full-resolution rows
// beyond the band edge are still exact; an exchange (fyn_halo_exchange, peer stores over NVLink) is issued on a layer's
// input only when the layer would otherwise reach into spoilt rows.  With the 8-row margin of round 2's first version that is
// after every layer (15 exchanges per StyleNet frame); 44 rows cover the ten 3x3 layers of the residual trunk at 1/4
// resolution, so the trunk runs as ONE chain kernel between two exchanges.  Returns false when an exchange would fall
// inside a chain (its inner tensors are private to the kernel) even with the chain's input freshly exchanged.
bool Engine::planHalo() {
    using gpu::TensorHandle;
    haloSteps_.clear();
    const long long M = haloMargin_;
    std::vector<int> forcedBefore;                 // chain heads whose input is exchanged whatever its state
    for (int attempt = 0; attempt < 16; attempt++) {
        haloSteps_.clear();
        std::unordered_map<TensorHandle, long long> valid;
        std::unordered_map<TensorHandle, int> producer;
        int conflictHead = -1;
        int chainHead = -1, chainLeft = 0;          // inside a chain: its head, layers still to come (this one included)
        auto scaleOf = [&](TensorHandle t) {
            fyn_tensor_desc d{};
            FYN_ABI_CALL(fyn_tensor_get_desc(t, &d, nullptr));
            if (d.height <= 0 || haloInputHeight_ % d.height) THROW_EXCEPTION_ARGS(FynException, "Row bands: tensor height %d does not divide the band height %d", d.height, haloInputHeight_);
            return (long long)(haloInputHeight_ / d.height);
        };
        auto validOf = [&](TensorHandle t) {
            auto f = valid.find(t);
            return f == valid.end() ? M : f->second;          // (a tensor handed in by the caller: margins as uploaded)
        };
        auto refresh = [&](TensorHandle t) {                   // exchange on t, issued behind its producer
            auto pr = producer.find(t);
            if (pr == producer.end()) return false;            // the network input: nothing to exchange with
            auto slot = haloSlots_.find(pr->second);
            if (slot == haloSlots_.end()) return false;
            haloSteps_[pr->second] = slot->second;
            const long long sc = scaleOf(t);
            valid[t] = (M / sc) * sc;
            // an exchange behind an inner layer of a chain cannot be issued
            auto *pc = dynamic_cast<gpu::ConvLayerBase *>(layers_[pr->second]);
            if (pc && pc->chained() && chainHead >= 0 && pr->second >= chainHead && chainLeft > 0) conflictHead = chainHead;
            return true;
        };
        for (auto it = layers_.begin(); it != layers_.end() && conflictHead < 0; ++it) {
            auto *g = dynamic_cast<gpu::GPULayerBase *>(it.second);
            if (!g || dynamic_cast<gpu::DownloadLayer *>(g) || !g->hasOutputTexture(0)) continue;
            TensorHandle out = g->getOutputTexture(0);
            producer[out] = it.first;
            if (dynamic_cast<gpu::UploadLayer *>(g)) {
                valid[out] = M;
                continue;
            }
            auto *conv = dynamic_cast<gpu::ConvLayerBase *>(g);
            if (conv && conv->chainLength() > 0) {
                chainHead = it.first;
                chainLeft = conv->chainLength();
                for (int f : forcedBefore)
                    if (f == it.first && conv->hasInputTexture(0)) refresh(conv->getInputTexture(0));
                conflictHead = -1;                               // (refreshing the head's own input is no conflict)
            }
            const long long so = scaleOf(out);
            auto evaluate = [&]() -> long long {
                long long vo;
                if (conv) {
                    const fyn_conv_desc &d = conv->descriptor();
                    TensorHandle in = conv->getInputTexture(0);
                    const long long si = scaleOf(in), viRows = validOf(in) / si;
                    const long long mh = (long long)((d.kernel - 1) / 2) * std::max(1, d.dilation);
                    long long voRows;
                    if (!d.fractional) {
                        const long long ds = std::max(1, d.downsample);
                        const long long num = viRows - mh - (ds - 1);
                        voRows = ds == 1 ? viRows - mh : (num >= 0 ? num / ds : -1);
                    } else {
                        // source rows floor(s * (ds * o + 0.5 + tap)), |tap| <= mh (gpu/vanilla/convlayerbase_vanilla.cpp:352-371): at most
                        // ceil(s * (mh + 1)) source rows around the source position of the output row, one more for the floor
                        const long long reach = (long long)std::ceil((double)d.source_step * (double)(mh + 1));
                        const long long left = viRows - reach - 1;
                        voRows = left >= 0 ? left * si / so : -1;
                    }
                    vo = voRows * so;
                    if (TensorHandle res = conv->residualTexture()) vo = std::min(vo, (validOf(res) / so) * so);
                } else if (dynamic_cast<gpu::SigmoidLayer *>(g) || dynamic_cast<gpu::BatchNormLayer *>(g)) {
                    vo = validOf(g->getInputTexture(0));
                } else {
                    // any other layer: assume it uses up the margin (an exchange behind every such layer, the conservative rule)
                    vo = validOf(g->getInputTexture(0)) >= M ? 0 : -1;
                }
                return vo;
            };
            long long vo = evaluate();
            if (vo < 0) {
                bool any = false;
                if (g->hasInputTexture(0) && validOf(g->getInputTexture(0)) < (M / scaleOf(g->getInputTexture(0))) * scaleOf(g->getInputTexture(0))) any = refresh(g->getInputTexture(0)) || any;
                if (conv && conv->residualTexture() && validOf(conv->residualTexture()) < (M / so) * so) any = refresh(conv->residualTexture()) || any;
                vo = any ? evaluate() : vo;
                if (vo < 0) {
                    if (chainHead >= 0 && chainLeft > 0 && it.first != chainHead) conflictHead = chainHead;   // only a shorter chain would do
                    else THROW_EXCEPTION_ARGS(FynException, "Layer %s: a halo margin of %d rows does not cover its taps", g->getName().c_str(), haloMargin_);
                }
            }
            valid[out] = vo;
            if (chainLeft > 0 && --chainLeft == 0) chainHead = -1;
        }
        if (conflictHead < 0) return true;
        for (int f : forcedBefore)
            if (f == conflictHead) return false;                 // already tried with a fresh input: the chain is too long for this margin
        forcedBefore.push_back(conflictHead);
    }
    return false;
}

void Engine::cleanup() {
    dropGraph();
    if (setup_) {
        if (async_) finish();
        FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
        updateChains(false);
        collectTimings(false);
        for (void *e : freeEvents_) fyn_event_destroy(context_.handle(), e);
        freeEvents_.clear();
        for (int i = 0; i < ASYNC_SLOTS; i++) {
            if (uploadDone_[i]) fyn_event_destroy(context_.handle(), uploadDone_[i]);
            if (computeDone_[i]) fyn_event_destroy(context_.handle(), computeDone_[i]);
            if (copyDone_[i]) fyn_event_destroy(context_.handle(), copyDone_[i]);
            uploadDone_[i] = computeDone_[i] = copyDone_[i] = nullptr;
            slotUsed_[i] = false;
        }
        layers_.cleanup();
    }
    layers_ = CompiledLayers();
    setup_ = false;
}

void Engine::resetTimings() {
    collectTimings(true);
    timingData_.clear();
    deviceTimingData_.clear();
    runs_ = 0;
}

Engine::execstate Engine::forwardLayers() {
    if (!setup_) return EXEC_ERROR;
    if (async_) {
        // back-pressure: wait until fewer than two sequences are in flight (reference: engine.cpp:310-329)
        std::unique_lock<std::mutex> lck(flightLock_);
        flightCv_.wait(lck, [this]() { return inFlight_ < MAX_IN_FLIGHT; });
        inFlight_++;
        lck.unlock();
        uint64_t seq = sequenceNo_++;
        return executeAsync(seq);
    }
    uint64_t seq = sequenceNo_++;
    return execute(seq);
}

int Engine::sequencesInFlight() {
    std::lock_guard<std::mutex> lck(flightLock_);
    return inFlight_;
}

void Engine::sequenceCompleted(uint64_t sequence, cpu::CPUBuffer *buffer) {
    if (buffer) buffer->setSequence(sequence);
    if (downloadCallback_) downloadCallback_(sequence, buffer);
    {
        std::lock_guard<std::mutex> lck(flightLock_);
        inFlight_--;
    }
    flightCv_.notify_all();
}

static void engineCompletionTrampoline(void *user) {
    auto *c = static_cast<Engine::Completion *>(user);
    if (c->download) c->download->notifyDownloaded(c->sequence, c->buffer);
    c->engine->sequenceCompleted(c->sequence, c->buffer);
}

static void engineUploadTrampoline(void *user) {
    auto *n = static_cast<Engine::UploadNote *>(user);
    n->layer->notifyUploaded(n->sequence);
}

// Pipelined execution on three streams (upload / compute / download), double-buffered at both ends:
//   upload(n+1)  ||  layers(n)  ||  host copy(n-1)
// The reference gets the same overlap from PBO uploads on AsyncPool threads with shadow textures
// (gpu/uploadlayer.cpp:395-541) and fenced PBO read-backs (gpu/downloadlayer.cpp:139-157,307-323).
Engine::execstate Engine::executeAsync(uint64_t sequence) {
    fyn_ctx *ctx = context_.handle();
    CudaContext *cc = context_.interface();
    void *sC = context_.stream(), *sU = cc->uploadStream(), *sD = cc->downloadStream();
    const int slot = (int)(sequence % ASYNC_SLOTS);
    for (int i = 0; i < ASYNC_SLOTS; i++) {
        if (!uploadDone_[i]) {
            FYN_ABI_CALL(fyn_event_create(ctx, &uploadDone_[i]));
            FYN_ABI_CALL(fyn_event_create(ctx, &computeDone_[i]));
            FYN_ABI_CALL(fyn_event_create(ctx, &copyDone_[i]));
        }
    }
    gpu::UploadLayer *upload = nullptr;
    gpu::DownloadLayer *download = nullptr;
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        if (!upload) upload = dynamic_cast<gpu::UploadLayer *>(it.second);
        if (auto *d = dynamic_cast<gpu::DownloadLayer *>(it.second)) download = d;
    }
    static const bool traceOn = getenv("FYN_ASYNC_TRACE") != nullptr;
    TraceEntry te{sequence, {}};
    auto mark = [&](int k, void *stream) {
        if (!traceOn) return;
        FYN_ABI_CALL(fyn_event_create(ctx, &te.ev[k]));
        FYN_ABI_CALL(fyn_event_record(ctx, te.ev[k], stream));
    };
    if (traceOn && !traceBase_) {
        FYN_ABI_CALL(fyn_event_create(ctx, &traceBase_));
        FYN_ABI_CALL(fyn_event_record(ctx, traceBase_, sC));
    }
    // ---- upload: buffer `slot` is free once the layers of sequence-ASYNC_SLOTS have consumed it
    mark(0, sU);
    if (upload) {
        if (slotUsed_[slot]) FYN_ABI_CALL(fyn_stream_wait_event(ctx, sU, computeDone_[slot]));
        gpu::TensorHandle t = upload->asyncUpload(sequence, slot, sU);
        FYN_ABI_CALL(fyn_event_record(ctx, uploadDone_[slot], sU));
        FYN_ABI_CALL(fyn_stream_wait_event(ctx, sC, uploadDone_[slot]));
        if (upload->hasCallback()) {
            // UPLOAD_COMMENCED / UPLOAD_DONE once the data has left the caller's buffer (host function behind the copy)
            uploadNotes_[slot] = UploadNote{upload, sequence};
            FYN_ABI_CALL(fyn_stream_wait_event(ctx, cc->notifyStream(), uploadDone_[slot]));
            FYN_ABI_CALL(fyn_stream_add_callback(ctx, cc->notifyStream(), engineUploadTrampoline, &uploadNotes_[slot]));
        }
        for (auto &rcv : upload->receivers())
            if (auto *g = dynamic_cast<gpu::GPULayerBase *>(rcv.first)) g->updateInputTexture(t, rcv.second);
    }
    mark(1, sU);
    // ---- layers on the compute stream
    mark(2, sC);
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        LayerBase *layer = it.second;
        if (layer == upload || layer == download) continue;
        layer->forward(sequence);
    }
    // ---- download: device half on the compute stream (staging buffer `slot` is free once copy(sequence-2) is done),
    //      host half on the download stream
    cpu::CPUBuffer *buffer = nullptr;
    if (download) {
        if (slotUsed_[slot]) FYN_ABI_CALL(fyn_stream_wait_event(ctx, sC, copyDone_[slot]));
        download->asyncConvert(slot, sC);
    }
    mark(3, sC);
    FYN_ABI_CALL(fyn_event_record(ctx, computeDone_[slot], sC));
    FYN_ABI_CALL(fyn_stream_wait_event(ctx, sD, computeDone_[slot]));
    mark(4, sD);
    if (download) buffer = download->asyncCopy(sequence, slot, sD);
    mark(5, sD);
    if (traceOn) trace_.push_back(te);
    FYN_ABI_CALL(fyn_event_record(ctx, copyDone_[slot], sD));
    completions_[slot] = Completion{this, sequence, buffer, download};
    void *sN = cc->notifyStream();
    FYN_ABI_CALL(fyn_stream_wait_event(ctx, sN, copyDone_[slot]));
    FYN_ABI_CALL(fyn_stream_add_callback(ctx, sN, engineCompletionTrampoline, &completions_[slot]));
    slotUsed_[slot] = true;
    return EXEC_DEFERRED;
}

// strict ascending-layer-number execution (reference: engine.cpp:386-683, hot loop 1).
// Timings: host microseconds around each forward() like the reference (:443-449,595-601) plus device time from
// CUDA event pairs recorded around every layer WITHOUT host synchronisation; the pairs are resolved lazily
// (collectTimings) once the stream has been synchronised, so enabling timings does not serialise the step.
Engine::execstate Engine::execute(uint64_t sequence) {
    size_t slot = 0;
    static const bool debugSyncOn = getenv("FYN_DEBUG_SYNC") != nullptr;
    static const bool nvtxOn = !(getenv("FYN_NVTX") && atoi(getenv("FYN_NVTX")) == 0);
    struct NvtxRange {
        bool on;
        NvtxRange(bool enabled, const char *name) : on(enabled) { if (on) nvtxRangePushA(name); }
        ~NvtxRange() { if (on) nvtxRangePop(); }
    };
    NvtxRange sequenceRange(nvtxOn, "fyusenet forward");
    // CUDA-graph replay of the device layers (launch-latency-bound networks: ResNet-50 at batch 1 is 59 launches)
    const bool graphOk = useGraph_ && !timings_ && !writeResults_ && !haloComm_ && !debugSyncOn;
    const uint64_t epochNow = gpu::graphEpoch().load();
    if (!graphOk || graphEpoch_ != epochNow) {
        dropGraph();
        graphWarm_ = false;
        graphEpoch_ = epochNow;
    }
    // the first forward after a change runs eagerly (lazy per-kernel attribute set-up must not happen inside a capture), the
    // second one is captured, later ones replay
    const bool capturing = graphOk && !graphExec_ && graphWarm_;
    bool inCapture = false;
    auto isIO = [](LayerBase *l) { return dynamic_cast<gpu::UploadLayer *>(l) || dynamic_cast<gpu::DownloadLayer *>(l); };
    bool replayed = false;
    for (auto it = layers_.begin(); it != layers_.end(); ++it) {
        LayerBase *layer = it.second;
        const bool io = isIO(layer);
        if (io && skipIO_) continue;
        if (graphOk && !io) {
            if (graphExec_) {
                if (!replayed) {
                    FYN_ABI_CALL(fyn_graph_launch(context_.handle(), graphExec_, context_.stream()));
                    replayed = true;
                }
                continue;
            }
            if (capturing && !inCapture && !graphExec_) {
                FYN_ABI_CALL(fyn_graph_begin_capture(context_.handle(), context_.stream()));
                inCapture = true;
            }
        } else if (inCapture) {
            // an I/O layer ends the captured run of device layers (the download synchronises the stream)
            FYN_ABI_CALL(fyn_graph_end_capture(context_.handle(), context_.stream(), &graphExec_));
            FYN_ABI_CALL(fyn_graph_launch(context_.handle(), graphExec_, context_.stream()));
            inCapture = false;
        }
        NvtxRange layerRange(nvtxOn, layer->getName().c_str());
        if (timings_ && (timingOnly_ < 0 || timingOnly_ == it.first)) {
            if (pendingEvents_.size() >= 4096) collectTimings(true);
            void *evA = nullptr, *evB = nullptr;
            if (freeEvents_.size() >= 2) {
                evA = freeEvents_.back(); freeEvents_.pop_back();
                evB = freeEvents_.back(); freeEvents_.pop_back();
            } else {
                FYN_ABI_CALL(fyn_event_create(context_.handle(), &evA));
                FYN_ABI_CALL(fyn_event_create(context_.handle(), &evB));
            }
            fyn_event_record(context_.handle(), evA, context_.stream());
            tstamp t0 = fy_get_stamp();
            layer->forward(sequence);
            tstamp t1 = fy_get_stamp();
            fyn_event_record(context_.handle(), evB, context_.stream());
            timingData_[it.first] += (uint32_t)fy_elapsed_micros(t0, t1);
            pendingEvents_.push_back({it.first, evA, evB});
        } else {
            layer->forward(sequence);
        }
        if (haloComm_) {
            auto hs = haloSteps_.find(it.first);
            if (hs != haloSteps_.end()) FYN_ABI_CALL(fyn_halo_exchange(haloComm_, hs->second.slot, hs->second.rows, context_.stream()));
        }
        // FYN_DEBUG_SYNC=1: synchronise after every layer and name it (finds the layer a device fault or hang belongs to)
        const bool debugSync = debugSyncOn;
        if (debugSync) {
            fprintf(stderr, "[fyn debug] seq %llu layer %s ...", (unsigned long long)sequence, layer->getName().c_str());
            fflush(stderr);
            FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
            fprintf(stderr, " done\n");
        }
        if (writeResults_) {
            char fname[1024];
            snprintf(fname, sizeof(fname), "%s/%s_%llu.bin", outputDir_.c_str(), layer->getName().c_str(), (unsigned long long)sequence);
            layer->writeResult(fname, false);
        }
        slot++;
    }
    if (inCapture) {
        FYN_ABI_CALL(fyn_graph_end_capture(context_.handle(), context_.stream(), &graphExec_));
        FYN_ABI_CALL(fyn_graph_launch(context_.handle(), graphExec_, context_.stream()));
    }
    if (graphOk && !graphExec_) graphWarm_ = true;
    if (timings_) runs_++;
    return EXEC_DONE;
}

void Engine::collectTimings(bool sync) {
    if (pendingEvents_.empty()) return;
    if (sync) FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
    for (auto &p : pendingEvents_) {
        float ms = 0.f;
        if (fyn_event_elapsed_ms(context_.handle(), p.start, p.stop, &ms) == 0) deviceTimingData_[p.layer] += ms;
        freeEvents_.push_back(p.start);
        freeEvents_.push_back(p.stop);
    }
    pendingEvents_.clear();
}

Engine::execstate Engine::finish() {
    if (!setup_) return EXEC_ERROR;
    if (async_) {
        // all deferred sequences must have delivered their download (reference: engine.cpp:264-274, 5 s timeout there)
        std::unique_lock<std::mutex> lck(flightLock_);
        flightCv_.wait(lck, [this]() { return inFlight_ == 0; });
    }
    FYN_ABI_CALL(fyn_stream_sync(context_.handle(), context_.stream()));
    collectTimings(false);
    if (!trace_.empty()) dumpTrace();
    return EXEC_DONE;
}

void Engine::dumpTrace() {
    fyn_ctx *ctx = context_.handle();
    CudaContext *cc = context_.interface();
    fyn_stream_sync(ctx, cc->uploadStream());
    fyn_stream_sync(ctx, cc->downloadStream());
    fyn_stream_sync(ctx, cc->notifyStream());
    fprintf(stderr, "[async trace] ms since first sequence: seq | upload begin-end | layers begin-end | host copy begin-end\n");
    for (TraceEntry &t : trace_) {
        float v[6] = {};
        for (int k = 0; k < 6; k++) {
            fyn_event_elapsed_ms(ctx, traceBase_, t.ev[k], &v[k]);
            fyn_event_destroy(ctx, t.ev[k]);
        }
        fprintf(stderr, "[async trace] %4llu | %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f\n", (unsigned long long)t.seq, v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    trace_.clear();
}

// ------------------------------------------------------------------------------------------------
// NeuralNetwork
// ------------------------------------------------------------------------------------------------
NeuralNetwork::NeuralNetwork(const GfxContextLink &ctx) {
    // like the reference, the main context is used when none is passed (neuralnetwork.cpp:57-66); it is
    // resolved lazily in setup() so that network objects (layer tables, weight offsets) exist without a device
    setContext(ctx);
}

NeuralNetwork::~NeuralNetwork() {
    if (engine_ || bufferMgr_) cleanup();
}

void NeuralNetwork::asynchronous(const AsyncAdapter &adapter) {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "Cannot switch to asynchronous operation after setup()");
    async_ = true;
    asyncCallbacks_ = adapter;
}

void NeuralNetwork::setBatch(int batch) {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "Batch size must be set before setup()");
    if (batch < 1) THROW_EXCEPTION_ARGS(FynException, "Illegal batch size %d", batch);
    batch_ = batch;
}

void NeuralNetwork::setup() {
    if (setup_) THROW_EXCEPTION_ARGS(FynException, "Network already set up");
    if (!context_.isValid()) setContext(GfxContextManager::instance(0)->createMainContext());
    engine_ = new Engine(context_, async_);
    engine_->setup(this);
    if (async_) {
        AsyncAdapter cbs = asyncCallbacks_;
        engine_->setDownloadCallback([cbs](uint64_t seq, cpu::CPUBuffer *buf) {
            if (cbs.downReady_) cbs.downReady_("download", seq, buf);
            if (cbs.seqDone_) cbs.seqDone_(seq);
        });
    }
    setup_ = true;
}

void NeuralNetwork::cleanup() {
    if (engine_) {
        engine_->cleanup();
        delete engine_;
        engine_ = nullptr;
    }
    if (bufferMgr_) {
        bufferMgr_->cleanup();
        delete bufferMgr_;
        bufferMgr_ = nullptr;
    }
    setup_ = false;
}

NeuralNetwork::execstate NeuralNetwork::forward() {
    execstate st;
    if (!setup_ || !engine_) {
        st.status = Engine::EXEC_ERROR;
        return st;
    }
    st.sequenceNo = engine_->nextSequenceNo();
    if (async_ && asyncCallbacks_.newSeq_) asyncCallbacks_.newSeq_(st.sequenceNo);
    st.status = engine_->forwardLayers();
    return st;
}

NeuralNetwork::execstate NeuralNetwork::finish() {
    execstate st;
    st.status = engine_ ? engine_->finish() : Engine::EXEC_ERROR;
    st.sequenceNo = engine_ ? engine_->lastSequenceNo() : 0;
    return st;
}

// build -> connect -> load weights -> per-layer setup (reference: neuralnetwork.cpp:238-250)
CompiledLayers NeuralNetwork::glSetup() {
    CompiledLayers layers = buildLayers();
    bufferMgr_ = new BufferManager(context_, batch_);
    connectLayers(layers, bufferMgr_);
    initializeWeights(layers);
    for (auto it = layers.begin(); it != layers.end(); ++it) it.second->setup();
    return layers;
}

std::shared_ptr<LayerFactory> NeuralNetwork::getLayerFactory(compute_device dev) {
    if (dev != compute_device::DEV_GPU) THROW_EXCEPTION_ARGS(FynException, "Device type not supported");
    return LayerFactory::instance(LayerFactory::GPUFactoryType(LayerFactory::GPUFactoryType::VANILLA, context_));
}

}  // namespace fyusenet
}  // namespace fyusion
