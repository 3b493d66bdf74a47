// fyn_conv_chain.cu -- one persistent tcgen05 kernel for a CHAIN of shallow NxN convolutions of identical geometry
// (StyleNet's residual trunk: res1_1 ... res5_2, ten 3x3 40->40 layers on the same 381x464 tensor).
//
// The reference renders every layer as its own sequence of blend passes (fyusenet/gpu/vanilla/convlayerNxN_vanilla.cpp:72-145,
// residual input: shaders/vanilla/residual.inc); fyn_conv_tc.cu turned a layer into one kernel.  For thin layers on
// L2-resident tensors that kernel spends a third of its time before its first epilogue (prologue, first rows' copy
// latency, first transform) and another ~2 us between launches, because a dependent launch needs the WHOLE previous
// grid.  Here one launch walks all layers; a strip of layer l+1 only waits for the strips of layer l it reads.
//
// What changes against fyn_conv_tc.cu (the arithmetic does not: same step table, same weight images, same epilogue, so the
// result is bit-identical to the unfused layers):
//  * Intermediate tensors are private to the chain and stored in the OPERAND layout of the next layer:
//      image[n][row + P][8-channel chunk][pixel + mh][8 x fp16]   (P = tensor padding, mh = kernel / 2),
//    whose border rows / columns are zero like the padding texels of a tensor; with P = 0 (StyleNet: CLAMP_TO_EDGE reads
//    at the borders, base/buffermanager.cpp:657-670) the border columns replicate the edge pixel and rows clamp,
//    with the consumer's prefix activation already applied (activation-at-fetch, base/layerbase.h:50-59, moved to the
//    producer's store: the value is computed from the fp16-rounded result exactly as the loader warps would).  A row of a
//    strip therefore arrives in the ring slot by `nchunks` bulk copies (TMA unit) and is a valid UMMA operand as it
//    lands: no loader warps, no staging buffers, no generic-proxy transform, no fence.proxy.async on shared memory.
//    The un-activated values a later layer adds as its residual live in one more image that every thread updates in
//    place (it is read and written by the same thread).
//  * Work: the image is cut into strips of SH rows x 128 columns; a CTA owns one strip in the upper and one in the lower
//    half of the image and processes (layer 0: upper, lower), (layer 1: upper, lower), ...  While a CTA works on its lower
//    strip every neighbour of its upper strip finishes the same layer, so the next layer's rows can be fetched without a
//    bubble.  Rows and jobs are numbered cumulatively over the whole program, so the ring of row slots, the two TMEM
//    accumulators and all mbarrier phases simply run on across strips and layers.
//  * Cross-CTA dependencies: progress[image of layer l][strip][image n][column block] = rows of that strip stored so far,
//    tagged with a per-launch epoch (no reset between launches).  Epilogue threads count themselves into a shared-memory
//    counter per accumulator buffer after their stores (red.release); a publisher warp polls the two counters, derives
//    the number of completed jobs and turns it into `fence; st.release` of the strips' counters -- it may skip values
//    when it falls behind and never holds the epilogue up.  The producer warp
//    polls (ld.acquire, cached) the counters of the <= 3 column blocks a row spans before it issues the row's copies,
//    followed by fence.proxy.async (generic-proxy stores of other CTAs -> async-proxy reads of the bulk copies).
//    All CTAs of the grid must be co-resident (grid <= SM count, one CTA per SM): checked on the host.
//  * The weight image of layer l+1 is fetched while layer l runs (two buffers).
//
// Warp roles ((E + 4) x 32 threads, E = 8 or 12): warps [0, E) epilogue, E producer, E+1 / E+2 MMA issuers (even / odd
// jobs, one TMEM accumulator each), E+3 publisher.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda.h>      // CUtensorMap and the cuTensorMapEncodeTiled prototype (resolved at run time through cudaGetDriverEntryPoint)

#include "fyn_internal.h"
#include "fyn_tc_common.cuh"

namespace {

constexpr int kMaxChain = 16;
constexpr int kChainBufs = 4;     // rotating input images: layer l writes image (l+1) % 4 while neighbours may still read images l and l-1

struct ChainLayer {
    const uint4 *wimg;   // weight image of the layer (ConvTcPlan::d_wimg)
    ActParams act;       // prefix activation of the layer: applied by whoever stores the layer's input image
    int resSrc;          // 0 no residual, 1 the chain input (plane layout), 2 the raw image
    int reluRes, bnRes;
    int writeRaw;        // a later layer adds this layer's un-activated output: keep it in the raw image
};

struct ChainArgs {
    // tensor maps of the rotating input images: [rows of all images][chunk][pixel][8 x fp16], box = one strip row
    // (all chunks x rowpx pixels): a row arrives by ONE cp.async.bulk.tensor (SASS UTMALDG)
    // (in global memory, written once by the host: with the maps inside this struct the kernel parameters exceeded 4 KB and
    // the whole kernel ran 13 % slower -- the MMA issuers read their step table from the parameter bank)
    const CUtensorMap *tmap;
    int useTma;
    TView in, out;                    // chain input / output tensors (shallow plane layout, fp16, packing 4)
    __half *A[kChainBufs];
    __half *R;
    unsigned long long *progress;
    unsigned long long epoch;
    int nlayers;
    ChainLayer layer[kMaxChain];
    uint32_t wbytes, idesc, b_lbo;
    int nsteps;
    TcStep steps[kMaxSteps];
    TcStep stepsRev[kMaxSteps];       // the same steps with the window rows in reverse order (layers that sweep bottom -> top)
    int N, nchunks, rowpx, K, mh, slotBytes, nslots, nmirror;
    int biasFolded;
    uint32_t biasB16, onesOff, epiOff;
    int Wj, Hj, P;
    int nxs, nG, SH, nsub, batch, NS;
    int pitchPx, imgRows;
    long long imgElems;               // halves per image n of an internal image
    int epiWarps;
    int nInPlanes, nOutPlanes;
    int fenceMode;                    // experiment switch (FYN_CHAIN_FENCE): bit 0 acquire fence, bit 1 proxy fence after a poll
    int rawInput;                     // layer 1 adds the chain input: the input relayout also fills the raw image
    int simpleAct;
};

// (gpu-scope acquire / release operations flush the SM's L1 and drain its store queue: measured, a polling loop of
// ld.acquire.gpu more than doubled the time of the epilogue's stores on the same SM.  Polls are relaxed; ONE fence follows.)
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// generic-proxy writes (of other CTAs, made visible by the acquire) before async-proxy reads of global memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_global_16(const __half *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_cta_shared_inc(uint32_t *p) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

#ifdef FYN_CHAIN_PROFILE
// event trace of three blocks (0, the middle one and its right neighbour): [block][segment][event] in ns (%globaltimer)
__device__ unsigned long long g_ctrace[3][64][8];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CTRACE(seg, ev) do { if (tb >= 0 && (seg) < 64) g_ctrace[tb][seg][ev] = gtimer(); } while (0)
__device__ unsigned long long g_jtrace[128][4];
#define JTRACE(job, ev) do { if (tb == 1 && (job) < 128) g_jtrace[job][ev] = gtimer(); } while (0)
#define CPROF_DECL(n) long long n = 0
#define CPROF_T() clock64()
#define CPROF_ADD(acc, t0) acc += clock64() - (t0)
#define CPROF_ON(x) x
#else
#define CPROF_DECL(n)
#define CPROF_T() 0
#define CPROF_ADD(acc, t0)
#define CPROF_ON(x)
#define CTRACE(seg, ev)
#define JTRACE(job, ev)
#endif

// one strip row through the tensor map: box [1 row][nchunks][rowpx pixels][8 halves] -> one ring slot, completion on `bar`
__device__ __forceinline__ void tma_load_row(void *dst, const CUtensorMap *map, uint64_t *bar, int px, int row) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(px), "r"(0), "r"(row)
                 : "memory");
}

struct Ring {
    int slot, fill;
    __device__ __forceinline__ void next(int nslots) {
        if (++slot == nslots) {
            slot = 0;
            fill++;
        }
    }
};

// NOCT: accumulator octets (= 8-channel chunks) per epilogue thread, (N / 8) / (epilogue warps / 4).  SIMPLE: every prefix
// activation of the chain is ReLU or none.
template <int NOCT, bool SIMPLE>
__global__ void __launch_bounds__(512, 1) k_conv_tc_chain(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t wpad = (a.wbytes + 127u) & ~127u;
    unsigned char *sW = smem;                                                                    // [2] weight images
    unsigned char *sRing = smem + 2 * (size_t)wpad;                                              // [nslots + nmirror] rows
    float4 *sEpi = reinterpret_cast<float4 *>(sRing + (size_t)(a.nslots + a.nmirror) * a.slotBytes);   // [2][32]: bias[16], scale[16]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sEpi + 64);
    uint64_t *full = bars;                    // [nslots] row landed             (bulk copies -> MMA)
    uint64_t *empty = bars + a.nslots;        // [nslots] row retired            (both MMA warps -> producer)
    uint64_t *tfull = bars + 2 * a.nslots;    // [2]
    uint64_t *tempty = tfull + 2;             // [2]
    uint64_t *wbar = tempty + 2;              // [2] weight image landed
    uint32_t *sCnt = reinterpret_cast<uint32_t *>(wbar + 2);   // [2] epilogue threads that have stored their part of a job, per buffer
    uint32_t *tmemBase = sCnt + 2;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int E = a.epiWarps, prodWarp = E, mmaWarp0 = E + 1, pubWarp = E + 3;

    // CTA -> (image n, row group g, column block xb); the CTA owns strips t = h * nG + g, h = 0 .. nsub-1
    int bid = blockIdx.x;
    const int xb = bid % a.nxs;
    bid /= a.nxs;
    const int g = bid % a.nG;
    const int n = bid / a.nG;
    const int j0 = xb * kTileM;
#ifdef FYN_CHAIN_PROFILE
    const int tb = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x / 2 + 1 ? 2 : -1));
#endif
    auto strip = [&](int h, int &t, int &ja, int &nj) {
        t = h * a.nG + g;
        ja = t * a.SH;
        nj = max(0, min(a.Hj, ja + a.SH) - ja);
    };
    auto prog = [&](int image, int t, int xq) -> unsigned long long * {
        return a.progress + (((size_t)image * a.NS + t) * a.batch + n) * a.nxs + xq;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nslots; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kMmaWarps);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], E * 32);
            mbar_init(&wbar[b], 1);
            sCnt[b] = 0;
        }
        fence_barrier_init();
        // weights are not produced by the previous kernel: both buffers are filled before the grid dependency resolves
        for (int l = 0; l < 2 && l < a.nlayers; l++) {
            mbar_expect_tx(&wbar[l], a.wbytes);
            bulk_g2s(sW + (size_t)l * wpad, a.layer[l].wimg, a.wbytes, &wbar[l]);
        }
    }
    if (warp == mmaWarp0) tmem_alloc(tmemBase, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    grid_dep_launch();
    grid_dep_wait();

    if (warp == prodWarp) {
        // ===================== producer: rows of the input images -> ring slots =====================
        // Layers alternate their sweep direction (even layers top -> bottom, odd layers bottom -> top).  A strip of layer l
        // therefore starts next to the rows its vertical neighbour finished FIRST in layer l-1, continues with its own rows
        // (finished a moment ago, but without any skew between CTAs) and ends next to the rows the other vertical neighbour
        // finished LAST -- a whole strip of work earlier.  The rows of a segment are issued in groups by source strip:
        // (1) lanes 0..8 poll the counters of that strip in the three column blocks a row spans until they cover the group's
        // rows (relaxed loads, values cached per segment); (2) one (row, chunk) bulk copy per lane.  A group is short enough
        // that no slot it waits for can depend on a row of the same group (<= nslots - K + 1 rows).
        const uint32_t chunkBytes = (uint32_t)a.rowpx * 16u;
        const unsigned long long epochTag = a.epoch << 32;
        const int batch = min(32, a.nslots - a.K + 1);
        int rowCum = 0;
        CPROF_DECL(pPoll); CPROF_DECL(pIssue); CPROF_DECL(nPolls);
        [[maybe_unused]] const long long pStart = CPROF_T();
        for (int l = 0; l < a.nlayers; l++) {
            const __half *img = a.A[l % kChainBufs] + (long long)n * a.imgElems;
            const bool rev = (l & 1) != 0;                   // this layer's sweep
            const bool prevRev = l > 0 && ((l - 1) & 1) != 0;   // sweep of the layer that wrote the image (the relayout of the chain input publishes whole strips)
            for (int h = 0; h < a.nsub; h++) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                if (nj == 0) continue;
                const int R = nj + a.K - 1;
                // r-th row of the segment in issue order: texture row with CLAMP_TO_EDGE, as the samplers of the reference
                // address it (yr), and the data row it holds (yr - P; padding rows are nobody's)
                auto tex_row = [&](int r) { return min(max((rev ? ja + nj - 1 + a.mh - r : ja - a.mh + r) + a.P, 0), a.imgRows - 1); };
                // lane i < 9 watches strip t - 1 + i / 3 in column block xb - 1 + i % 3
                const int t2 = t - 1 + lane / 3, xq = xb - 1 + lane % 3;
                const bool watcher = lane < 9 && t2 >= 0 && t2 < a.NS && xq >= 0 && xq < a.nxs;
                const int s2 = t2 * a.SH, n2 = min(a.Hj, s2 + a.SH) - s2;      // first row / rows of the watched strip
                const unsigned long long *p = watcher ? prog(l, t2, xq) : nullptr;
                if (lane == 0) CTRACE(l * a.nsub + h, 0);
                // one look at all nine counters up front (one L2 round trip): in the steady state it covers the whole segment
                unsigned long long seen = watcher && n2 > 0 ? ld_relaxed_gpu(p) : 0;
                CPROF_ON(nPolls++);
                if (a.fenceMode & 1) fence_acq_rel_gpu();
                if (a.fenceMode & 2) fence_proxy_async_all();
                // ring position of the segment's first row
                const int fill0 = rowCum / a.nslots, slot0 = rowCum - fill0 * a.nslots;
                for (int rb = 0; rb < R; rb += 32) {
                    // lane = row rb + lane: its source strip (-1 = padding row) and what it needs of it, in the strip's sweep order
                    const int r = rb + lane;
                    const bool rowOk = r < R;
                    const int yr = tex_row(rowOk ? r : R - 1), yd = yr - a.P;
                    // (rows of a segment come from its own strip and the two next to it: no division needed)
                    int src = -1;
                    if (yd >= 0 && yd < a.Hj) src = yd < t * a.SH ? t - 1 : (yd >= (t + 1) * a.SH ? t + 1 : t);
                    const int ss = src * a.SH, sn = min(a.Hj - ss, a.SH);
                    const unsigned needRow = src < 0 ? 0u : (unsigned)(prevRev ? ss + sn - yd : yd - ss + 1);
                    const int srcPrev = __shfl_up_sync(0xffffffffu, src, 1);
                    // groups: consecutive rows of one source, at most `batch` rows
                    unsigned starts = __ballot_sync(0xffffffffu, rowOk && (lane == 0 || src != srcPrev));
                    const int rowsHere = min(32, R - rb);
                    while (starts) {
                        const int g0 = __ffs(starts) - 1;
                        starts &= starts - 1;
                        const int gEnd = starts ? __ffs(starts) - 1 : rowsHere;
                        const int src0 = __shfl_sync(0xffffffffu, src, g0);
                        for (int b0 = g0; b0 < gEnd; b0 += batch) {
                            const int b1 = min(gEnd, b0 + batch);
                            if (src0 >= 0) {
                                const unsigned need = __reduce_max_sync(0xffffffffu, (lane >= b0 && lane < b1) ? needRow : 0u);
                                const bool mine = watcher && t2 == src0 && n2 > 0;
                                const unsigned long long want = epochTag | need;
                                bool ok = !mine || seen >= want;
                                if (!__all_sync(0xffffffffu, ok)) {
                                    const long long t0 = clock64();
                                    for (;;) {
                                        if (!ok) {
                                            seen = ld_relaxed_gpu(p);
                                            ok = seen >= want;
                                        }
                                        CPROF_ON(nPolls++);
                                        if (__all_sync(0xffffffffu, ok)) break;
                                        __nanosleep(64);
                                        if (clock64() - t0 > 6000000000ll) {     // ~3 s: a lost dependency must not hang the device
                                            if (!ok) printf("[fyn chain] block %d: layer %d strip %d never got %u rows of strip %d in column block %d\n", (int)blockIdx.x, l, t, need, t2, xq);
                                            __trap();
                                        }
                                    }
                                    if (a.fenceMode & 1) fence_acq_rel_gpu();
                                    if (a.fenceMode & 2) fence_proxy_async_all();
                                    CPROF_ADD(pPoll, t0);
                                }
                            }
                            if (lane == 0 && rb == 0 && b0 == 0) CTRACE(l * a.nsub + h, 1);
                            [[maybe_unused]] const long long pt = CPROF_T();
                            if (a.useTma) {
                                // one tensor copy per row (one lane each): box = all chunks x rowpx pixels of image row yr
                                const int rl = b0 + lane;
                                if (rl < b1) {
                                    int slot = slot0 + rb + rl, fill = fill0;
                                    while (slot >= a.nslots) {
                                        slot -= a.nslots;
                                        fill++;
                                    }
                                    mbar_wait(&empty[slot], (fill & 1) ^ 1);
                                    mbar_expect_tx(&full[slot], (uint32_t)a.nchunks * chunkBytes);
                                    tma_load_row(sRing + (size_t)slot * a.slotBytes, &a.tmap[l % kChainBufs], &full[slot], j0, n * a.imgRows + tex_row(rb + rl));
                                }
                                __syncwarp();
                                CPROF_ADD(pIssue, pt);
                                continue;
                            }
                            // one (row, chunk) copy per lane: ceil(rows * nchunks / 32) copy instructions per group
                            const int items = (b1 - b0) * a.nchunks;
                            for (int it = lane; it < items; it += 32) {
                                const int rr = it / a.nchunks, c = it - rr * a.nchunks, rl = b0 + rr;
                                int slot = slot0 + rb + rl, fill = fill0;
                                while (slot >= a.nslots) {
                                    slot -= a.nslots;
                                    fill++;
                                }
                                const int yrr = tex_row(rb + rl);
                                mbar_wait(&empty[slot], (fill & 1) ^ 1);
                                if (c == 0) mbar_expect_tx(&full[slot], (uint32_t)a.nchunks * chunkBytes);
                                const __half *srcp = img + (((long long)yrr * a.nchunks + c) * a.pitchPx + j0) * 8;
                                unsigned char *dst = sRing + (size_t)slot * a.slotBytes + (size_t)c * chunkBytes;
                                bulk_g2s(dst, srcp, chunkBytes, &full[slot]);
                            }
                            __syncwarp();
                            CPROF_ADD(pIssue, pt);
                        }
                    }
                }
                if (lane == 0) CTRACE(l * a.nsub + h, 2);
                rowCum += R;
            }
        }
#ifdef FYN_CHAIN_PROFILE
        if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && lane == 0)
            printf("[chain prof] block %d producer: total %lld poll %lld (%lld looks) issue incl. waitEmpty %lld\n", (int)blockIdx.x,
                   (long long)(clock64() - pStart), pPoll, nPolls, pIssue);
#endif
    } else if (warp == mmaWarp0 || warp == mmaWarp0 + 1) {
        // ===================== MMA issuers =====================
        // As in fyn_conv_tc.cu: warp w issues the jobs with (cumulative index & 1) == w into TMEM buffer w.  Row bookkeeping
        // is cumulative: a warp waits for the rows [waited, window end) and, before it issues a job, releases the rows
        // below that job's window -- all of which it has waited for, and which only its earlier MMAs (covered by the
        // commit) can still be reading.
        const int mw = warp - mmaWarp0;
        const uint32_t rbase16 = smem_u32(sRing) >> 4, slot16 = (uint32_t)a.slotBytes >> 4;
        const uint64_t hiA = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
        const uint32_t d = tmem + (uint32_t)mw * 64u;
        Ring wt{0, 0}, rl{0, 0}, st{0, 0};     // next row to wait for / to release; first row of the current segment
        int waited = 0, released = 0, rowBase = 0, Q = 0;
        CPROF_DECL(pWaitT); CPROF_DECL(pWaitF); CPROF_DECL(pIss); CPROF_DECL(pRel); CPROF_DECL(pCommit);
        [[maybe_unused]] const long long pStart = CPROF_T();
        for (int l = 0; l < a.nlayers; l++) {
            const uint32_t sWl = smem_u32(sW) + (uint32_t)(l & 1) * wpad;
            const uint32_t bconst = (sWl >> 4) | ((a.b_lbo >> 4) << 16);
            const uint32_t onesDesc = ((sWl + a.onesOff) >> 4) | ((16u >> 4) << 16);
            bool haveW = false;
            for (int h = 0; h < a.nsub; h++) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                if (nj == 0) continue;
                const int R = nj + a.K - 1;
                Ring win = st;
                for (int q = 0; q < nj; q++, win.next(a.nslots)) {
                    const int Qg = Q + q;
                    if ((Qg & 1) != mw) continue;
                    const int use = Qg >> 1;
                    const int start = rowBase + q, needTo = start + a.K;
                    if (elect_one()) {
                        // rows below this job's window that this warp has already seen land go back first (a short ring
                        // would otherwise deadlock at a strip boundary, where the window jumps by K rows) ...
                        [[maybe_unused]] long long pt = CPROF_T();
                        Ring rr = rl;
                        int k = released;
                        for (; k < start && k < waited; k++, rr.next(a.nslots)) umma_commit(&empty[rr.slot]);
                        CPROF_ADD(pRel, pt);
                        pt = CPROF_T();
                        mbar_wait(&tempty[mw], (use & 1) ^ 1);
                        CPROF_ADD(pWaitT, pt);
                        pt = CPROF_T();
                        Ring w = wt;
                        for (int kk = waited; kk < needTo; kk++, w.next(a.nslots)) mbar_wait(&full[w.slot], w.fill & 1);
                        if (!haveW) mbar_wait(&wbar[l & 1], (l >> 1) & 1);
                        CPROF_ADD(pWaitF, pt);
                        pt = CPROF_T();
                        tc_fence_after();
                        // ... the rest (rows the other warp's windows needed, this warp's did not) once they have been seen
                        for (; k < start; k++, rr.next(a.nslots)) umma_commit(&empty[rr.slot]);
                        if (q == 0) CTRACE(l * a.nsub + h, 3);
                        JTRACE(Qg, 0);
                        const int winSlot = win.slot;
                        // (bottom -> top layers walk the window rows in reverse, with the kernel rows reversed in their weight
                        // image: the products enter the accumulator in the same order as in a top -> bottom layer, bit for bit)
                        const TcStep *steps = (l & 1) ? a.stepsRev : a.steps;
#pragma unroll 4
                        for (int s = 0; s < a.nsteps; s++) {
                            const TcStep stp = steps[s];
                            // (a_lo is relative to the step's window row, `pad` = that row: the ring wraps without mirror slots)
                            int slot = winSlot + (int)stp.pad;
                            if (slot >= a.nslots) slot -= a.nslots;
                            umma_f16(d, hiA | (uint64_t)(stp.a_lo + rbase16 + (uint32_t)slot * slot16), hiA | (uint64_t)(stp.b_off16 + bconst), a.idesc, stp.accumulate);
                        }
                        if (a.biasFolded) umma_f16(d, hiA | (uint64_t)onesDesc, hiA | (uint64_t)(a.biasB16 + bconst), a.idesc, 1u);
                        CPROF_ADD(pIss, pt);
                        JTRACE(Qg, 1);
                        pt = CPROF_T();
                        umma_commit(&tfull[mw]);
                        CPROF_ADD(pCommit, pt);
                        if (q == nj - 1) CTRACE(l * a.nsub + h, 4);
                    }
                    haveW = true;
                    for (; waited < needTo; waited++) wt.next(a.nslots);
                    for (; released < start; released++) rl.next(a.nslots);
                    __syncwarp();
                }
                rowBase += R;
                Q += nj;
                for (int k = 0; k < R; k++) st.next(a.nslots);
            }
        }
        // rows this warp never needed (the other warp's last windows) still owe its share of the release
        if (elect_one()) {
            for (int k = waited; k < rowBase; k++, wt.next(a.nslots)) mbar_wait(&full[wt.slot], wt.fill & 1);
            for (int k = released; k < rowBase; k++, rl.next(a.nslots)) umma_commit(&empty[rl.slot]);
        }
        __syncwarp();
#ifdef FYN_CHAIN_PROFILE
        if (mw == 0 && elect_one()) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            printf("[chain prof] block %3d xb %d g %2d sm %3u mma0: total %lld waitTempty %lld waitFull %lld issue %lld jobs %d release %lld commit %lld\n", (int)blockIdx.x, xb, g, smid, (long long)(clock64() - pStart), pWaitT, pWaitF,
                   pIss, Q, pRel, pCommit);
        }
#endif
    } else if (warp == pubWarp) {
        // ===================== publisher: completed jobs -> progress counters =====================
        // sCnt[b] / T = jobs on accumulator buffer b whose stores have all been issued (a thread only counts itself into
        // job Q once job Q - 2 is complete, so the quotient is exact); jobs alternate between the buffers, hence the
        // number of jobs completed IN ORDER is 2 * min + (1 if buffer 0 is ahead).
        if (lane == 0) {
            const unsigned long long epochTag = a.epoch << 32;
            const uint32_t T = (uint32_t)E * 32u;
            int total = 0;                                   // jobs of the layers that publish (all but the last)
            for (int h = 0; h < a.nsub; h++) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                total += nj;
            }
            total *= a.nlayers - 1;
            // Consumers wait for two values per strip only: its first mh rows (the halo of the strip above) and all of its
            // rows (its own next layer and the halo of the strip below), so those are the counts worth a release store.
            int l = 0, h = 0, segStart = 0, pubInSeg = 0;    // current segment, its first job, rows of it published so far
            long long t0 = clock64();
            CPROF_DECL(pFence); CPROF_DECL(nPub);
            while (segStart < total) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                if (nj == 0) {
                    if (++h == a.nsub) {
                        h = 0;
                        l++;
                    }
                    continue;
                }
                const uint32_t c0 = ld_acquire_cta_shared(&sCnt[0]) / T, c1 = ld_acquire_cta_shared(&sCnt[1]) / T;
                const int prefix = (int)((c0 > c1) ? 2 * c1 + 1 : 2 * c0);
                const int cur = min(prefix - segStart, nj);
                const int first = min(a.mh, nj);
                const int threshold = pubInSeg < first ? first : nj;
                if (cur < threshold) {
                    __nanosleep(32);
                    if (clock64() - t0 > 20000000000ll) __trap();
                    continue;
                }
                [[maybe_unused]] const long long pf = CPROF_T();
                st_release_gpu(prog(l + 1, t, xb), epochTag | (unsigned)cur);
                CPROF_ADD(pFence, pf);
                CPROF_ON(nPub++);
                pubInSeg = cur;
                if (cur == nj) {
                    CTRACE(l * a.nsub + h, 6);
                    segStart += nj;
                    pubInSeg = 0;
                    if (++h == a.nsub) {
                        h = 0;
                        l++;
                    }
                }
                t0 = clock64();
            }
#ifdef FYN_CHAIN_PROFILE
            if (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) printf("[chain prof] block %d publisher: %lld publishes for %d jobs, %lld cycles in fences\n", (int)blockIdx.x, nPub, total, pFence);
#endif
        }
    } else {
        // ===================== epilogue: warps [0, E) =====================
        // The job loop is kept SMALL on purpose: the first version of this role (run-time octet count, three residual
        // sources, generic activation; ~2k instructions per job) was bound by instruction fetch -- 2-4k cycles per job against
        // 1.1k of tensor-pipe time (role profile, make PROF=1).  Hence the template parameters, the hoisted per-thread
        // constants, and a raw image that also holds the chain input (one residual source).
        const int m = threadIdx.x & 127;
        const int part = warp >> 2;
        const int jx = j0 + m;
        const bool valid = jx < a.Wj;
        const int etid = threadIdx.x;
        const unsigned long long epochTag = a.epoch << 32;
        const long long chunkRow = (long long)a.pitchPx * 8;           // halves per (row, chunk) of an internal image
        const long long rowStride = chunkRow * a.nchunks;              // halves per row
        // this thread's chunks (= accumulator octets part * NOCT + o): offset inside a row of an internal image, or -1
        long long coff[NOCT];
        int plane0[NOCT];
#pragma unroll
        for (int o = 0; o < NOCT; o++) {
            const int c = part * NOCT + o;
            coff[o] = (valid && c < a.nchunks) ? (long long)c * chunkRow + (long long)(jx + a.mh) * 8 : -1;
            plane0[o] = 2 * c;
        }
        // P = 0: reads beyond the left / right edge see the edge pixel (CLAMP_TO_EDGE): the threads of the first and last
        // image column keep the mh border columns of the input images equal to their pixel
        const int edge = (a.P == 0 && valid) ? ((jx == 0 ? 1 : 0) | (jx == a.Wj - 1 ? 2 : 0)) : 0;
        auto store_image = [&](__half *p, uint4 v) {
            *reinterpret_cast<uint4 *>(p) = v;
            if (edge) {
                if (edge & 1) {
#pragma unroll 1
                    for (int e = 1; e <= a.mh; e++) *reinterpret_cast<uint4 *>(p - e * 8) = v;
                }
                if (edge & 2) {
#pragma unroll 1
                    for (int e = 1; e <= a.mh; e++) *reinterpret_cast<uint4 *>(p + e * 8) = v;
                }
            }
        };
        // prefix activation of the consumer, applied to the fp16 values as the loader warps of fyn_conv_tc.cu do.  SIMPLE:
        // every layer has ReLU or none: max(v, floor) with floor = 0 or -inf
        auto activate = [&](uint4 v, const ActParams &act, __half2 floor2) -> uint4 {
            if (SIMPLE) {
                __half2 *q = reinterpret_cast<__half2 *>(&v);
#pragma unroll
                for (int i = 0; i < 4; i++) q[i] = __hmax2(q[i], floor2);
                return v;
            }
            return act_h8(v, act);
        };
        auto floor_of = [](const ActParams &act) { return act.type == 1 ? __float2half2_rn(0.f) : __half2half2(__ushort_as_half((unsigned short)0xfc00)); };
        __half *rawImg = a.R + (long long)n * a.imgElems;
        // ---- chain input (plane layout) -> input image of layer 0 with that layer's prefix activation, and (if layer 1 adds the
        //      chain input) its un-activated copy in the raw image
        {
            const __half *inp = reinterpret_cast<const __half *>(a.in.ptr) + (long long)n * a.in.imageElems + ((long long)a.P * a.in.texW + a.P + jx) * 4;
            __half *img = a.A[0] + (long long)n * a.imgElems;
            const ActParams act0 = a.layer[0].act;
            const __half2 floor0 = floor_of(act0);
            const int parts = E >> 2;
            for (int h = 0; h < a.nsub; h++) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                if (nj == 0) continue;
                if (valid)
                    for (int i = ja; i < ja + nj; i++)
                        for (int c = part; c < a.nchunks; c += parts) {
                            const __half *p = inp + (long long)i * a.in.texW * 4 + (long long)(2 * c) * a.in.planeElems;
                            const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(p));
                            const uint2 hi = (2 * c + 1 < a.nInPlanes) ? __ldg(reinterpret_cast<const uint2 *>(p + a.in.planeElems)) : make_uint2(0u, 0u);
                            const uint4 raw = make_uint4(lo.x, lo.y, hi.x, hi.y);
                            const long long off = (long long)(i + a.P) * rowStride + (long long)c * chunkRow + (long long)(jx + a.mh) * 8;
                            store_image(img + off, activate(raw, act0, floor0));
                            if (a.rawInput) *reinterpret_cast<uint4 *>(rawImg + off) = raw;
                        }
                asm volatile("bar.sync 1, %0;" ::"r"(E * 32) : "memory");
                if (etid == 0) st_release_gpu(prog(0, t, xb), epochTag | (unsigned)nj);
            }
        }
        __half *outp = reinterpret_cast<__half *>(a.out.ptr) + (long long)n * a.out.imageElems + ((long long)a.P * a.out.texW + a.P + jx) * 4;
        int Q = 0;
        CPROF_DECL(pWaitTf); CPROF_DECL(pWaitW); CPROF_DECL(pCnt); CPROF_DECL(pLd); CPROF_DECL(pSt);
        [[maybe_unused]] const long long pStart = CPROF_T();
        [[maybe_unused]] long long pLayer0 = 0;
        for (int l = 0; l < a.nlayers; l++) {
            CPROF_ON(if (l == 1) pLayer0 = clock64() - pStart);
            const bool last = l + 1 == a.nlayers;
            const bool rev = (l & 1) != 0;
            const bool resOn = a.layer[l].resSrc != 0, reluRes = a.layer[l].reluRes != 0, bnRes = a.layer[l].bnRes != 0, writeRaw = a.layer[l].writeRaw != 0;
            const ActParams actNext = a.layer[last ? l : l + 1].act;
            const __half2 floorNext = floor_of(actNext);
            __half *nextImg = a.A[(l + 1) % kChainBufs] + (long long)n * a.imgElems;
            // epilogue parameters ride at the tail of the weight image: copy them out of the region the tensor core streams
            [[maybe_unused]] long long ptw = CPROF_T();
            mbar_wait(&wbar[l & 1], (l >> 1) & 1);
            CPROF_ADD(pWaitW, ptw);
            if (etid < 32) sEpi[(l & 1) * 32 + etid] = reinterpret_cast<const float4 *>(sW + (size_t)(l & 1) * wpad + a.epiOff)[etid];
            asm volatile("bar.sync 1, %0;" ::"r"(E * 32) : "memory");
            // every MMA of layer l-1 has completed (this thread has seen its last accumulator): its weight buffer is free
            if (etid == 0 && l >= 1 && l + 1 < a.nlayers) {
                fence_proxy_async_all();
                mbar_expect_tx(&wbar[(l + 1) & 1], a.wbytes);
                bulk_g2s(sW + (size_t)((l + 1) & 1) * wpad, a.layer[l + 1].wimg, a.wbytes, &wbar[(l + 1) & 1]);
            }
            const float4 *ep = sEpi + (l & 1) * 32;
            // One strip of the layer.  RES / LAST are compile-time so that each variant of the job loop stays short (the role is
            // bound by instruction issue: 12 warps x instructions per job / 4 schedulers).
            auto run_strip = [&](auto resTag, auto lastTag, int ja, int nj, int segIdx) {
                constexpr bool RES = decltype(resTag)::value, LAST = decltype(lastTag)::value;
                // odd layers sweep bottom -> top (see the producer): job q is output row ja + nj - 1 - q there
                const int i0 = rev ? ja + nj - 1 : ja;
                const long long rowStep = rev ? -rowStride : rowStride;
                const long long outStride = rev ? -(long long)a.out.texW * 4 : (long long)a.out.texW * 4;
                __half *nextRow = nextImg + (long long)(i0 + a.P) * rowStride;      // row of the next layer's input image
                __half *rawRow = rawImg + (long long)(i0 + a.P) * rowStride;        // row of the raw image
                __half *outRow = outp + (long long)i0 * a.out.texW * 4;             // row of the chain output (plane layout)
                uint4 rres[NOCT];
                if (RES) {
#pragma unroll
                    for (int o = 0; o < NOCT; o++)
                        if (coff[o] >= 0) rres[o] = ld_global_16(rawRow + coff[o]);
                }
                for (int q = 0; q < nj; q++, nextRow += rowStep, rawRow += rowStep, outRow += outStride) {
                    const int Qg = Q + q, buf = Qg & 1, use = Qg >> 1;
                    [[maybe_unused]] long long pt = CPROF_T();
                    mbar_wait(&tfull[buf], use & 1);
                    CPROF_ADD(pWaitTf, pt);
                    if (etid == 0) JTRACE(Qg, 2);
                    pt = CPROF_T();
                    tc_fence_after();
                    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * 64u + (uint32_t)(part * NOCT) * 8u;
                    uint32_t acc[NOCT][8];
#pragma unroll
                    for (int o = 0; o < NOCT; o++) tmem_ld8(taddr + o * 8, acc[o]);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&tempty[buf]);
                    CPROF_ADD(pLd, pt);
                    pt = CPROF_T();
                    // this job's residual texels were requested one job ago; request the next job's now
                    uint4 rcur[NOCT];
                    if (RES) {
#pragma unroll
                        for (int o = 0; o < NOCT; o++) rcur[o] = rres[o];
                        if (q + 1 < nj) {
#pragma unroll
                            for (int o = 0; o < NOCT; o++)
                                if (coff[o] >= 0) rres[o] = ld_global_16(rawRow + rowStep + coff[o]);
                        }
                    }
#pragma unroll
                    for (int o = 0; o < NOCT; o++) {
                        if (coff[o] < 0) continue;
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            float4 v = make_float4(__uint_as_float(acc[o][4 * k + 0]), __uint_as_float(acc[o][4 * k + 1]), __uint_as_float(acc[o][4 * k + 2]),
                                                   __uint_as_float(acc[o][4 * k + 3]));
                            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
                            if (!a.biasFolded) {
                                const int p = (plane0[o] + k) & 15;
                                const float4 bi = ep[p];
                                sc = ep[16 + p];
                                v = make_float4(fmaf(v.x, sc.x, bi.x), fmaf(v.y, sc.y, bi.y), fmaf(v.z, sc.z, bi.z), fmaf(v.w, sc.w, bi.w));
                            }
                            if (RES) {
                                const uint32_t r0 = k ? rcur[o].z : rcur[o].x, r1 = k ? rcur[o].w : rcur[o].y;
                                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&r0));
                                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&r1));
                                float4 rs = make_float4(f0.x, f0.y, f1.x, f1.y);
                                if (reluRes) rs = make_float4(fmaxf(rs.x, 0.f), fmaxf(rs.y, 0.f), fmaxf(rs.z, 0.f), fmaxf(rs.w, 0.f));
                                if (bnRes) rs = make_float4(rs.x * sc.x, rs.y * sc.y, rs.z * sc.z, rs.w * sc.w);
                                v.x += rs.x;
                                v.y += rs.y;
                                v.z += rs.z;
                                v.w += rs.w;
                            }
                            // (channels that only pad a chunk come out as zero: zero weights, zero bias, zero residual)
                            w[2 * k] = pack_half2(v.x, v.y);
                            w[2 * k + 1] = pack_half2(v.z, v.w);
                        }
                        const uint4 raw = make_uint4(w[0], w[1], w[2], w[3]);
                        if (LAST) {
                            __half *o0 = outRow + (long long)plane0[o] * a.out.planeElems;
                            *reinterpret_cast<uint2 *>(o0) = make_uint2(w[0], w[1]);
                            if (plane0[o] + 1 < a.nOutPlanes) *reinterpret_cast<uint2 *>(o0 + a.out.planeElems) = make_uint2(w[2], w[3]);
                        } else {
                            store_image(nextRow + coff[o], activate(raw, actNext, floorNext));
                            if (writeRaw) *reinterpret_cast<uint4 *>(rawRow + coff[o]) = raw;
                        }
                    }
                    CPROF_ADD(pSt, pt);
                    if (etid == 0) JTRACE(Qg, 3);
                    if (!LAST) {
                        // hand the job's stores to the publisher: count this thread into the job once the buffer's previous
                        // job (two jobs back) is complete -- which it practically always is
                        const uint32_t before = (uint32_t)use * (uint32_t)(E * 32);
                        pt = CPROF_T();
                        while (ld_acquire_cta_shared(&sCnt[buf]) < before) {}
                        red_release_cta_shared_inc(&sCnt[buf]);
                        CPROF_ADD(pCnt, pt);
                        if (etid == 0 && q == nj - 1) CTRACE(segIdx, 5);
                    }
                }
            };
            for (int h = 0; h < a.nsub; h++) {
                int t, ja, nj;
                strip(h, t, ja, nj);
                if (nj == 0) continue;
                const int segIdx = l * a.nsub + h;
                if (last) {
                    if (resOn) run_strip(std::true_type{}, std::true_type{}, ja, nj, segIdx);
                    else run_strip(std::false_type{}, std::true_type{}, ja, nj, segIdx);
                } else {
                    if (resOn) run_strip(std::true_type{}, std::false_type{}, ja, nj, segIdx);
                    else run_strip(std::false_type{}, std::false_type{}, ja, nj, segIdx);
                }
                Q += nj;
            }
        }
#ifdef FYN_CHAIN_PROFILE
        if (tb >= 0 && (etid == 0 || etid == 100 || etid == E * 32 - 1))
            printf("[chain prof] block %d epilogue thread %d: total %lld (first layer incl. input relayout %lld) waitTfull %lld tmemLd %lld math+stores %lld countIn %lld waitWeights %lld jobs %d\n", (int)blockIdx.x, etid,
                   (long long)(clock64() - pStart), pLayer0, pWaitTf, pLd, pSt, pCnt, pWaitW, Q);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == mmaWarp0) tmem_dealloc(tmem, 128);
#ifdef FYN_CHAIN_PROFILE
    if (tb == 1 && threadIdx.x == 0) {
        for (int j = 40; j < 62; j++)
            printf("[chain job] %3d: mma ready %7lld issued %7lld | epilogue sees %7lld stored %7lld (ns)\n", j, (long long)(g_jtrace[j][0] - g_jtrace[40][0]),
                   (long long)(g_jtrace[j][1] - g_jtrace[40][0]), (long long)(g_jtrace[j][2] - g_jtrace[40][0]), (long long)(g_jtrace[j][3] - g_jtrace[40][0]));
    }
    if (tb >= 0 && threadIdx.x == 0) {
        const unsigned long long base = g_ctrace[0][0][0];
        for (int sg = 0; sg < a.nlayers * a.nsub && sg < 64; sg++)
            printf("[chain trace] block %d seg %2d: poll %7lld - %7lld  issued %7lld | mma %7lld - %7lld | epilogue done %7lld  published %7lld (ns)\n", (int)blockIdx.x, sg,
                   (long long)(g_ctrace[tb][sg][0] - base), (long long)(g_ctrace[tb][sg][1] - base), (long long)(g_ctrace[tb][sg][2] - base), (long long)(g_ctrace[tb][sg][3] - base),
                   (long long)(g_ctrace[tb][sg][4] - base), (long long)(g_ctrace[tb][sg][5] - base), (long long)(g_ctrace[tb][sg][6] - base));
    }
#endif
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct fyn_conv_chain {
    fyn_ctx *ctx = nullptr;
    std::vector<fyn_op *> ops;
    std::vector<int> resFrom;
    ChainArgs args{};
    __half *images = nullptr;            // kChainBufs input images + the raw image, one allocation
    unsigned long long *progress = nullptr;
    CUtensorMap *tmaps = nullptr;        // device copy of the tensor maps of the input images
    bool tmapsOk = false;
    int batchAlloc = 0;
    unsigned long long epoch = 0;
    size_t smemBytes = 0;
};

namespace {

int chain_fail_unsupported(const char *why) {
    fyn_set_error("conv chain: %s", why);
    return FYN_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" {

int fyn_conv_chain_create(fyn_ctx *ctx, fyn_op *const *ops, const int *residual_from, int n, fyn_conv_chain **out) {
    if (!ctx || !ops || !out || n < 2) FYN_FAIL(FYN_ERR_INVALID, "conv chain: bad argument");
    *out = nullptr;
    if (n > kMaxChain) return chain_fail_unsupported("too many layers");
    const fyn_op *first = ops[0];
    if (!first || first->kind != FYN_OP_CONV || !first->tc) return chain_fail_unsupported("layers must run on the shallow tcgen05 family");
    const fyn_conv_desc &d0 = first->conv;
    const TcArgs &p0 = first->tc->chainArgs();
    const int K = d0.kernel, mh = (K - 1) / 2;
    if (d0.in_channels != d0.out_channels || d0.downsample != 1 || d0.dilation != 1 || d0.fractional || K < 3)
        return chain_fail_unsupported("layers must be plain stride-1 NxN convolutions with as many outputs as inputs");
    const int P = d0.in_padding;
    if (P > mh || d0.out_padding != P) return chain_fail_unsupported("tensor padding must be the same on both sides and at most the kernel's half width");
    if (p0.mode != 0 || p0.opx != 1 || p0.opy != 1 || p0.nver != 1 || p0.ds != 1 || p0.rowAdvance != 1 || p0.x_lead != mh || p0.nrows != K)
        return chain_fail_unsupported("plan is not a single-phase stride-1 plan");
    for (int i = 0; i < n; i++) {
        const fyn_op *op = ops[i];
        if (!op || op->kind != FYN_OP_CONV || !op->tc || op->ctx != ctx) return chain_fail_unsupported("layers must run on the shallow tcgen05 family");
        const fyn_conv_desc &d = op->conv;
        if (d.width != d0.width || d.height != d0.height || d.in_channels != d0.in_channels || d.out_channels != d0.out_channels || d.kernel != K ||
            d.downsample != 1 || d.dilation != 1 || d.fractional || d.in_padding != P || d.out_padding != P)
            return chain_fail_unsupported("layers differ in geometry");
        if (d.flags & (FYN_FLAG_DEEP | FYN_FLAG_PRE_CLIP)) return chain_fail_unsupported("deep tensors / clip activations are not chained");
        if ((d.flags & FYN_FLAG_POST_BATCHNORM) != (d0.flags & FYN_FLAG_POST_BATCHNORM)) return chain_fail_unsupported("layers differ in post-batchnorm");
        if (op->epilogue != FYN_EPILOGUE_NONE || op->innorm) return chain_fail_unsupported("layers with fused functions are not chained");
        if (!op->tc->d_wimgFlip) return chain_fail_unsupported("a layer has no chain plan");
        const TcArgs &p = op->tc->chainArgs();
        if (p.nsteps != p0.nsteps || p.N != p0.N || p.wbytes != p0.wbytes || p.biasFolded != p0.biasFolded || p.slotBytes != p0.slotBytes || p.rowpx != p0.rowpx ||
            memcmp(p.steps, p0.steps, sizeof(TcStep) * p0.nsteps) != 0)
            return chain_fail_unsupported("layers differ in their tcgen05 plan");
        if (d.flags & FYN_FLAG_RESIDUAL_INPUT) {
            if (d.res_padding != P) return chain_fail_unsupported("residual padding differs");
            if (!residual_from || i < 1 || residual_from[i] != i - 2) return chain_fail_unsupported("a residual must be the input of the previous layer");
        }
    }
    fyn_conv_chain *c = new fyn_conv_chain();
    c->ctx = ctx;
    c->ops.assign(ops, ops + n);
    c->resFrom.assign(n, -2);
    if (residual_from) c->resFrom.assign(residual_from, residual_from + n);
    ChainArgs &a = c->args;
    a.nlayers = n;
    a.wbytes = p0.wbytes;
    a.idesc = p0.idesc;
    a.b_lbo = p0.b_lbo;
    a.nsteps = p0.nsteps;
    {
        // steps are generated window row by window row, the same number for every row of a single-version plan
        if (p0.nsteps % K != 0) {
            delete c;
            return chain_fail_unsupported("steps do not split evenly over the window rows");
        }
        // the chain's tables address a step's operand relative to its window row (TcStep::pad = the row), so that the ring
        // needs no mirror slots; the reversed table walks the rows bottom-up
        const int per = p0.nsteps / K;
        const uint32_t slot16 = (uint32_t)p0.slotBytes >> 4;
        for (int w = 0; w < K; w++)
            for (int j = 0; j < per; j++) {
                TcStep st = p0.steps[w * per + j];
                st.a_lo -= (uint32_t)w * slot16;
                st.pad = (uint32_t)w;
                st.accumulate = (w == 0 && j == 0) ? 0u : 1u;
                a.steps[w * per + j] = st;
                st.accumulate = (w == K - 1 && j == 0) ? 0u : 1u;
                a.stepsRev[(K - 1 - w) * per + j] = st;
            }
    }
    a.N = p0.N;
    a.nchunks = p0.nchunks;
    a.rowpx = p0.rowpx;
    a.K = K;
    a.mh = mh;
    a.slotBytes = (p0.slotBytes + 127) & ~127;     // ring slots start on 128-byte boundaries (destination of the tensor copies)
    if (getenv("FYN_CHAIN_ALIGN") && atoi(getenv("FYN_CHAIN_ALIGN")) == 0) a.slotBytes = p0.slotBytes;   // (measurement knob; implies plain bulk copies)
    a.nmirror = 0;
    a.biasFolded = p0.biasFolded;
    a.biasB16 = p0.biasB16;
    a.onesOff = p0.onesOff;
    a.epiOff = p0.epiOff;
    a.P = P;
    a.Wj = first->Wo;
    a.Hj = first->Ho;
    a.nInPlanes = a.nOutPlanes = (d0.out_channels + 3) / 4;
    a.epiWarps = ((a.N >> 3) % 3 == 0) ? 12 : 8;
    if (const char *e = getenv("FYN_CHAIN_EPI")) {
        const int want = atoi(e);
        if (want == 8 || (want == 12 && (a.N >> 3) % 3 == 0)) a.epiWarps = want;
    }
    if ((a.N >> 3) / (a.epiWarps >> 2) > 4 || (a.N >> 3) % (a.epiWarps >> 2) != 0 || (a.N >> 3) < a.nchunks) {
        delete c;
        return chain_fail_unsupported("accumulator columns do not split over the epilogue warps");
    }
    bool simple = true;
    a.rawInput = 0;
    for (int i = 0; i < n; i++) {
        const fyn_conv_desc &d = ops[i]->conv;
        ChainLayer &L = a.layer[i];
        L.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
        L.resSrc = (d.flags & FYN_FLAG_RESIDUAL_INPUT) ? 2 : 0;
        if (L.resSrc && i == 1) a.rawInput = 1;
        if (L.act.type > 1) simple = false;
        L.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
        L.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
        L.writeRaw = (i + 2 < n && (ops[i + 2]->conv.flags & FYN_FLAG_RESIDUAL_INPUT)) ? 1 : 0;
    }
    a.simpleAct = simple ? 1 : 0;
    // Consumer-side fences after a poll (bit 0: fence.acq_rel.gpu, bit 1: fence.proxy.async).  Off by default: each costs
    // the SM ~2k cycles with the epilogue's stores in flight (measured: trunk 144 us with both, 126 us without), and the rows
    // are then read by bulk copies -- issued only after the branch on the polled value, served by L2 (no L1, nothing to
    // invalidate), where the producer's release store has ordered the rows before the counter.
    a.fenceMode = 0;
    if (const char *e = getenv("FYN_CHAIN_FENCE")) a.fenceMode = atoi(e);
    // ring: as many slots as shared memory holds (two weight images, epilogue parameters, barriers)
    const size_t wpad = ((size_t)a.wbytes + 127) & ~(size_t)127;
    const size_t optin = (size_t)ctx->prop.sharedMemPerBlockOptin;
    auto footprint = [&](int ns) { return 2 * wpad + (size_t)ns * a.slotBytes + 64 * 16 + (2 * (size_t)ns + 8) * 8 + 16; };
    int ns = 24;
    if (const char *e = getenv("FYN_CHAIN_SLOTS")) ns = std::max(1, std::min(24, atoi(e)));
    while (ns > 0 && footprint(ns) > optin) ns--;
    a.nslots = ns;
    if (a.nslots < 2 * K) {
        delete c;
        return chain_fail_unsupported("ring does not fit shared memory");
    }
    c->smemBytes = footprint(a.nslots);
    *out = c;
    return FYN_OK;
}

int fyn_conv_chain_layers(const fyn_conv_chain *c) { return c ? (int)c->ops.size() : 0; }

int fyn_conv_chain_run(fyn_conv_chain *c, const fyn_tensor *in, fyn_tensor *out, void *stream) {
    if (!c || !in || !out) FYN_FAIL(FYN_ERR_INVALID, "conv chain: NULL argument");
    fyn_ctx *ctx = c->ctx;
    const fyn_conv_desc &d0 = c->ops[0]->conv;
    auto fits = [&](const fyn_tensor *t) {
        return t->desc.width == d0.width && t->desc.height == d0.height && t->desc.channels == d0.in_channels && t->desc.padding == c->args.P &&
               t->desc.order == FYN_ORDER_SHALLOW && t->desc.dtype == FYN_F16 && t->geom.packing == 4;
    };
    if (!fits(in) || !fits(out)) return 1;          // tensor formats the chain does not cover: run the layers one by one
    if (in->desc.batch != out->desc.batch || in->dptr == out->dptr) FYN_FAIL(FYN_ERR_INVALID, "conv chain: input and output must be distinct tensors of one batch size");
    FYN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    ChainArgs a = c->args;
    a.batch = in->desc.batch;
    a.nxs = (a.Wj + kTileM - 1) / kTileM;
    const int sms = ctx->prop.multiProcessorCount;
    const long long cols = (long long)a.nxs * a.batch;
    if (cols > sms) return 1;                        // the grid must be co-resident
    // Strips: nG row groups (CTAs per column block) x nsub strips per CTA ("zones" of the image), strip height SH.  More
    // zones hide more of the layer-to-layer dependency latency behind other strips' work but load more halo rows.
    a.nG = std::max(1, (int)(sms / cols));
    a.nsub = 2;
    if (const char *e = getenv("FYN_CHAIN_NSUB")) a.nsub = std::max(1, std::min(8, atoi(e)));
    int minSH = std::max(3, a.mh);               // (a segment reads from its own strip and the two next to it only: SH >= mh)
    if (const char *e = getenv("FYN_CHAIN_SH")) minSH = std::max(std::max(1, atoi(e)), a.mh);
    while (a.nsub > 1 && (a.Hj + a.nsub * a.nG - 1) / (a.nsub * a.nG) < minSH) a.nsub--;
    a.SH = (a.Hj + a.nsub * a.nG - 1) / (a.nsub * a.nG);
    if (a.SH < minSH) {
        a.SH = std::min(minSH, a.Hj);
        a.nG = (a.Hj + a.SH - 1) / a.SH;          // one strip per CTA
    }
    a.NS = a.nsub * a.nG;
    a.pitchPx = a.nxs * kTileM + (a.rowpx - kTileM);
    a.imgRows = a.Hj + 2 * a.P;
    a.imgElems = (long long)a.imgRows * a.nchunks * a.pitchPx * 8;
    const size_t imgBytes = (size_t)a.imgElems * 2 * a.batch;
    const size_t progCount = (size_t)a.nlayers * a.NS * a.batch * a.nxs;
    if (c->batchAlloc != a.batch) {
        // (re)allocation synchronises; it happens on the first run and when the batch size changes
        if (c->images) cudaFree(c->images);
        if (c->progress) cudaFree(c->progress);
        c->images = nullptr;
        c->progress = nullptr;
        FYN_CUDA(cudaMalloc((void **)&c->images, imgBytes * (kChainBufs + 1)));
        FYN_CUDA(cudaMemset(c->images, 0, imgBytes * (kChainBufs + 1)));          // zero borders, never written afterwards
        if (!c->tmaps) FYN_CUDA(cudaMalloc((void **)&c->tmaps, sizeof(CUtensorMap) * kChainBufs));
        c->tmapsOk = false;
        FYN_CUDA(cudaMalloc((void **)&c->progress, progCount * sizeof(unsigned long long)));
        FYN_CUDA(cudaMemset(c->progress, 0, progCount * sizeof(unsigned long long)));
        FYN_CUDA(cudaDeviceSynchronize());
        c->batchAlloc = a.batch;
        c->epoch = 0;
    }
    for (int b = 0; b < kChainBufs; b++) a.A[b] = c->images + (size_t)b * (imgBytes / 2);
    a.R = c->images + (size_t)kChainBufs * (imgBytes / 2);
    a.progress = c->progress;
    a.epoch = ++c->epoch;
    // tensor maps of the input images: encoded once per allocation, kept in device memory
    {
        using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = [] {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
            return reinterpret_cast<EncodeFn>(fn);
        }();
        // FYN_CHAIN_TMA=1: rows through the tensor map (cp.async.bulk.tensor, one instruction per row).  Default: `nchunks` plain
        // bulk copies per row issued by as many lanes -- measured faster (trunk 123.9 us against 128.8 us with the tensor
        // map): the box is a handful of 2 KB runs either way, and the plain copies of a group of rows go out in parallel.
        const bool tmaOn = getenv("FYN_CHAIN_TMA") && atoi(getenv("FYN_CHAIN_TMA")) != 0;      // (read per run: tests switch it)
        const bool want = encode && tmaOn && a.rowpx <= 256 && a.nchunks <= 256 && (a.slotBytes & 127) == 0;
        if (want && !c->tmapsOk) {
            alignas(64) CUtensorMap maps[kChainBufs];
            bool ok = true;
            for (int b = 0; b < kChainBufs && ok; b++) {
                const cuuint64_t dims[4] = {8, (cuuint64_t)a.pitchPx, (cuuint64_t)a.nchunks, (cuuint64_t)a.imgRows * a.batch};
                const cuuint64_t strides[3] = {16, (cuuint64_t)a.pitchPx * 16, (cuuint64_t)a.nchunks * a.pitchPx * 16};
                const cuuint32_t box[4] = {8, (cuuint32_t)a.rowpx, (cuuint32_t)a.nchunks, 1};
                const cuuint32_t estr[4] = {1, 1, 1, 1};
                ok = encode(&maps[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.A[b], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            if (ok) {
                FYN_CUDA(cudaMemcpy(c->tmaps, maps, sizeof(maps), cudaMemcpyHostToDevice));
                c->tmapsOk = true;
            }
        }
        a.useTma = (want && c->tmapsOk) ? 1 : 0;
        a.tmap = c->tmaps;
    }
    // (hot-swapped weights re-pack in place; odd layers sweep bottom -> top and use the image with the kernel rows reversed)
    for (size_t i = 0; i < c->ops.size(); i++) a.layer[i].wimg = c->ops[i]->tc->chainImage((i & 1) != 0);
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    using ChainKernel = void (*)(ChainArgs);
    static const ChainKernel kernels[4][2] = {{k_conv_tc_chain<1, false>, k_conv_tc_chain<1, true>}, {k_conv_tc_chain<2, false>, k_conv_tc_chain<2, true>},
                                              {k_conv_tc_chain<3, false>, k_conv_tc_chain<3, true>}, {k_conv_tc_chain<4, false>, k_conv_tc_chain<4, true>}};
    const int noct = (a.N >> 3) / (a.epiWarps >> 2);
    const ChainKernel fn = kernels[noct - 1][a.simpleAct ? 1 : 0];
    {
        // the attribute is per function and device: keep it at the largest footprint any chain has needed so far
        static size_t cur[64][4][2] = {};
        static std::mutex lock;
        std::lock_guard<std::mutex> guard(lock);
        size_t &have = cur[ctx->device & 63][noct - 1][a.simpleAct ? 1 : 0];
        if (c->smemBytes > have) {
            FYN_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smemBytes));
            have = c->smemBytes;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(cols * a.nG));
    cfg.blockDim = dim3((unsigned)((a.epiWarps + 4) * 32));
    cfg.dynamicSmemBytes = c->smemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const bool noPdl = getenv("FYN_TC_NO_PDL") != nullptr;
    if (noPdl) cfg.numAttrs = 0;
    FYN_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
    FYN_CHECK_LAUNCH(ctx);
    return FYN_OK;
}

int fyn_conv_chain_destroy(fyn_conv_chain *c) {
    if (!c) return FYN_OK;
    cudaSetDevice(c->ctx->device);
    if (c->images) cudaFree(c->images);
    if (c->progress) cudaFree(c->progress);
    if (c->tmaps) cudaFree(c->tmaps);
    delete c;
    return FYN_OK;
}

}  // extern "C"
