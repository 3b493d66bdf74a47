// fyn_conv_tc.cu -- tcgen05 / TMEM implicit-GEMM convolution family for sm_100a.
//
// Replaces the raster-blend channel accumulation of the reference's shallow conv layers
// (fyusenet/gpu/vanilla/convlayerNxN_vanilla.cpp:72-145: one blend pass per (input plane, kernel row)) by
// accumulation in tensor memory; semantics are those of fyn_conv_direct.cu / the oracle.
//
// Formulation ("shifted-window implicit GEMM")
//   D[m][n] = sum_k A[m][k] * B[n][k]     m = 128 consecutive output pixels of one output row
//                                         n = output channel (padded to a multiple of 16)
//                                         k = (kernel tap, input channel)
//   A is never materialised (no im2col): each CTA keeps a ring of INPUT ROWS in shared memory in the UMMA
//   canonical K-major, non-swizzled layout, [8-channel chunk][pixel][8 x fp16] (16 bytes per (chunk, pixel)).
//   The A operand of tap (ky, kx) is then simply the same shared-memory image with the descriptor start
//   address moved by kx*16 bytes (one pixel) and the ring slot chosen by ky: rows of a core matrix stay
//   16 bytes apart, so the canonical layout still holds.  Rows arrive as bulk copies (TMA unit) of raw plane texels;
//   activation-at-fetch (fyusenet/base/layerbase.h:50-59, shaders/activation.inc) and clamp-to-edge addressing
//   (base/buffermanager.cpp:657-670) are applied by the loader warps while they transpose 4-channel planes into
//   8-channel chunks (generic proxy + fence.proxy.async) -- the transform is why TMA cannot write the operand image.
//   For 3-channel inputs (StyleNet conv1 reading the RGB32F upload texture) a chunk is two horizontally
//   adjacent pixels x 4 channels ("pixel-pair" mode), i.e. one 16-byte chunk covers taps kx and kx+1.
//
// Fractional convolutions (fyusenet/gpu/vanilla/fractionalconvlayerNxN_vanilla.cpp, shaders/vanilla/fraconv*.frag,
// fractional.inc) are phase-decomposed: with u = sourceStep*downsample = 1/p, output pixel p*j+phi reads source
// pixel j + delta(phi, tap) with delta = floor(s*(ds*phi + 0.5 + tap)), so every output phase (phi_y, phi_x) is an
// ordinary small convolution over the source image whose taps that land on the same source texel have their
// weights summed at plan time.  The reference's quirks are kept: 3x3 taps at -2s,-s,0, and the prefix
// activation on the first horizontal tap only -- the ring slot then holds two versions of the row (activated and
// raw) and the first-tap weights are routed to the activated version.
//
// Phase stacking.  A tcgen05.mma with M=128, K=16 costs ~(4096 + 32 N)/128 cycles for N <= 64 because its A tile
// (4 KB) comes from shared memory (measured: 46 cycles for N = 16..48, tools/ubench/umma_issue.cu), so thin
// layers are bound by the NUMBER of MMAs, not by their FLOPs.  All output phases that read the same source window
// are therefore stacked along N: n = phase * Cq + co.  One GEMM row then produces p_y x p_x output pixels
// (fractional convs) or four horizontally adjacent pixels (the 3-channel pixel-pair mode), with structural zeros in
// the weight image where a phase does not use a tap.  Where shared memory allows, two job rows are stacked the same
// way (plan_geometry's `ys`), and the bias of layers without post-BN scale is one more MMA step.
//
// Warp roles (576 threads): warps 0-7 epilogue (TMEM -> registers -> bias/BN/residual[/sigmoid] -> fp16 planes; the two
// warps of a TMEM lane quarter split the accumulator columns), warps 8-15 loaders (bulk-copy issue + row finishing in
// 2 or 4 groups that work on different rows), warps 16-17 issue tcgen05.mma for even / odd jobs (one elected lane each).
// Pipelines: landed (bulk copy -> loaders), full/empty mbarriers per ring slot (loaders <-> MMA), tmem_full/tmem_empty
// per accumulator buffer (MMA <-> epilogue, two buffers so the epilogue of job q overlaps the MMAs of job q+1).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "fyn_internal.h"
#include "fyn_tc_common.cuh"

namespace {


// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// dynamic shared memory: [weight image][ring slots + mirror slots][epilogue params][barriers][tmem base]
//
// Work decomposition.  The job space (output rows / opy, output columns / opx) is cut into strips of 128 job
// columns x SH job rows; one CTA per strip.  A job is one GEMM accumulation: 128 job columns of one job row, i.e.
// 128 x opx output pixels of opy output rows, reading a window of `nrows` input rows that only moves forward, so
// the ring of row slots is a FIFO.  The first nmirror slots are mirrored behind the ring so that every window is
// contiguous in shared memory and the A descriptor of a step is (window base + constant).
// MODE: 0 plane-pair chunks, 1 pixel-pair chunks.  ACT: see act_h8_t.  RES: 0 no residual, 1 fp16 shallow residual of a
// single-phase layer (prefetched), 2 any other residual tensor (generic fetch).  EPI: FYN_EPILOGUE_* function fused behind
// the convolution (instantiated for RES == 0 only; other combinations run on the direct kernel).
template <int MODE, int ACT, int RES, int EPI>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW = smem;
    unsigned char *sRing = smem + ((a.wbytes + 127) & ~127u);
    unsigned char *sStage = sRing + (size_t)(a.nslots + a.nmirror) * a.slotBytes;                       // [nstages] raw input rows
    float4 *sEpi = reinterpret_cast<float4 *>(sStage + (size_t)a.nstages * a.stageBytes);                 // [16] bias, [16] scale (copied from the weight image's tail)
    int2 *sTab = reinterpret_cast<int2 *>(sEpi + 32);                                                     // [nitems] row item table
    uint64_t *sZero = reinterpret_cast<uint64_t *>(sTab + a.nitems);                                      // 16 zero bytes (missing second plane)
    uint64_t *bars = sZero + 2;
    uint64_t *full = bars;                     // [nslots] row is ready for the MMA warp        (finishers -> MMA)
    uint64_t *empty = bars + a.nslots;         // [nslots] all MMAs reading the row have retired (MMA -> finishers)
    uint64_t *tfull = bars + 2 * a.nslots;     // [2]
    uint64_t *tempty = tfull + 2;              // [2]
    uint64_t *wbar = tempty + 2;               // weight image landed
    uint64_t *landed = wbar + 1;               // [nstages] raw row copy has landed            (bulk copies -> loaders)
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(landed + a.nstages);

    // (MMA warp 0 owns the TMEM allocation)
    // warp index through a broadcast so the compiler treats the role dispatch (and everything derived from it) as
    // warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    [[maybe_unused]] const long long pK0 = PROF_T();

    // strip decode: blockIdx.x -> (image, column block, row segment)
    int bid = blockIdx.x;
    const int xb = bid % a.nxs;
    bid /= a.nxs;
    const int nseg = (a.Hj + a.SH - 1) / a.SH;
    const int seg = bid % nseg;
    const int n = bid / nseg;
    const int ja = seg * a.SH, jb = min(a.Hj, ja + a.SH);   // job rows [ja, jb)
    const int njobs = jb - ja;
    const int j0 = xb * kTileM;                               // first job column
    const int r0 = a.rowAdvance * ja + a.dyMin;               // first / last input row of the strip (unclamped)
    const int r1 = a.rowAdvance * (jb - 1) + a.dyMin + a.nrows - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nslots; s++) {
            mbar_init(&full[s], 1);              // the leader of the loader group that finished the row
            mbar_init(&empty[s], kMmaWarps);     // both MMA warps must have retired the row
        }
        for (int s = 0; s < a.nstages; s++) mbar_init(&landed[s], 1);   // one arrive.expect_tx + the bytes of the bulk copies
        mbar_init(&tfull[0], 1);
        mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], a.epiWarps * 32);
        mbar_init(&tempty[1], a.epiWarps * 32);
        mbar_init(wbar, 1);
        fence_barrier_init();
        // weight image -> shared memory: one bulk copy, the MMA warp waits for it before its first step
        mbar_expect_tx(wbar, a.wbytes);
        bulk_g2s(sW, a.wimg, a.wbytes, wbar);
    }
    if (warp == kMmaWarp) tmem_alloc(tmemBase, 128);
    // Row geometry (independent of the row): the texel run [xa, xz] of a texture row that this strip reads, and per
    // item the byte offset inside the staged row and inside the ring slot.
    const int esize = (a.in.dtype == FYN_F16) ? 2 : 4, bpp = a.in.packing * esize;   // bytes per texel
    const int gx0 = (MODE == 1) ? 4 * j0 - a.x_lead : ((a.ds == 1) ? j0 - a.x_lead : 2 * j0 - a.x_lead);
    const int npx = (MODE == 1) ? 2 * a.rowpx : a.rowpx;
    const int xa = min(max(gx0 + a.inP, 0), a.in.texW - 1), xz = min(max(gx0 + a.inP + npx - 1, 0), a.in.texW - 1);
    const int ncopy = (MODE == 1) ? 1 : a.nInPlanes;                      // bulk copies per row
    const int stagePlane = a.stageBytes / ncopy;                           // staged bytes per plane (multiple of 16)
    if (threadIdx.x < 2) sZero[threadIdx.x] = 0ull;
    for (int it = threadIdx.x; it < a.nitems; it += kThreads) {
        int2 e;
        if (MODE == 0) {
            // item = (slot pixel, chunk): x = staged offset of the chunk's first plane texel | "second plane present"
            const int c = it / a.rowpx, px = it - c * a.rowpx, half = a.rowpx >> 1;
            // stride 1: slot pixel = image pixel - (j0 - lead); stride 2: slot is [parity][pixel/2]
            const int gx = (a.ds == 1) ? gx0 + px : gx0 + 2 * (px % half) + (px / half);
            const int ix = min(max(gx + a.inP, 0), a.in.texW - 1);      // CLAMP_TO_EDGE
            e.x = (2 * c * stagePlane + (ix - xa) * 8) | ((2 * c + 1 < a.nInPlanes) ? 1 : 0);
            e.y = it * 16;
        } else {
            // item = pixel; chunk = two adjacent pixels x 4 channels, chunks split by parity so that GEMM rows (4 pixels
            // = 2 chunks apart) are 16 bytes apart: slot = [even chunks][odd chunks]
            const int cidx = it >> 1;
            e.x = (min(max(gx0 + it + a.inP, 0), a.in.texW - 1) - xa) * bpp;
            e.y = (cidx & 1) * (a.rowpx >> 1) * 16 + (cidx >> 1) * 16 + (it & 1) * 8;
        }
        sTab[it] = e;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    // Nothing above reads or writes a tensor: under programmatic dependent launch it overlaps the previous layer's
    // tail.  Every role waits for the previous grid before it touches tensor memory (reads AND writes: the buffer
    // pool may hand this layer an output buffer the previous layer still reads).
    [[maybe_unused]] const long long pK1 = PROF_T();
    grid_dep_launch();
    grid_dep_wait();
    [[maybe_unused]] const long long pK2 = PROF_T();

    if (warp >= a.epiWarps && warp < kMmaWarp) {
        // ===================== loaders =====================
        // Input rows travel global -> shared as bulk copies (TMA unit, no registers, no load/store-unit traffic): per
        // row one copy per 4-channel plane of the contiguous texel run the strip needs.  The loader warps then turn a
        // staged row into the UMMA operand layout: clamp-to-edge addressing (base/buffermanager.cpp:657-670),
        // activation at fetch (shaders/activation.inc), fp32 -> fp16, the transpose of two 4-channel planes into one
        // 8-channel chunk, the raw version for taps that bypass the activation and the mirror slots -- and publish
        // the row to the MMA warp.
        // The warps form finGroups groups; group g owns rows g, g + G, ..., the stages s = g (mod G) and the ring slots
        // s = g (mod G), so every barrier is waited on by one group in program order (parity waits stay one phase
        // apart) while G rows are being finished concurrently and nstages rows are in flight.
        const int P = a.inP;
        const int G = a.finGroups, groupThreads = ((kWorkWarps - a.epiWarps) * 32) / G;
        const int tl = threadIdx.x - a.epiWarps * 32, g = tl / groupThreads, tg = tl - g * groupThreads;
        const bool leader = tg < 32;                                  // first warp of the group issues the copies
        const int runBytes = (xz - xa + 1) * bpp;
        const unsigned long long base = reinterpret_cast<unsigned long long>(a.in.ptr) + (unsigned long long)n * a.in.imageElems * esize;
        const unsigned long long planeBytes = (unsigned long long)a.in.planeElems * esize;
        const int R = r1 - r0 + 1;
        const int mirrorOff = a.nslots * a.slotBytes;                 // the first nmirror slots are also written behind the ring
        const uint32_t base15 = (uint32_t)(base & 15ull), plane15 = (uint32_t)(planeBytes & 15ull);
        const int pk = a.in.packing;
        const bool f16in = a.in.dtype == FYN_F16;
        const bool fastIn = !f16in && pk == 3;
        const uint32_t zeroA = smem_u32(sZero);
        PROF_DECL(pLdWait); PROF_DECL(pFinProc); PROF_DECL(pFinFence); PROF_DECL(pFinWaitE);
        [[maybe_unused]] const long long pLdStart = PROF_T();
        // lane q of the leader copies plane q of row r into stage r % nstages
        auto issue_row = [&](int r) {
            const int st = r % a.nstages;
            const int iy = min(max(r0 + r + P, 0), a.in.texH - 1);   // texture row, CLAMP_TO_EDGE
            unsigned long long src = base + (unsigned long long)tg * planeBytes + ((unsigned long long)iy * a.in.texW + xa) * bpp;
            const uint32_t sh = (uint32_t)(src & 15ull);              // copies start on a 16-byte boundary
            const uint32_t bytes = (tg < ncopy) ? ((sh + runBytes + 15u) & ~15u) : 0u;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
            if (tg == 0) mbar_expect_tx(&landed[st], total);
            __syncwarp();
            if (tg < ncopy) bulk_g2s(sStage + (size_t)st * a.stageBytes + (size_t)tg * stagePlane, reinterpret_cast<const void *>(src - sh), bytes, &landed[st]);
        };
        if (leader)
            for (int r = g; r < R && r < a.nstages; r += G) issue_row(r);
        for (int r = g; r < R; r += G) {
            const int st = r % a.nstages, use = r / a.nstages;
            const int slotIdx = r % a.nslots, fill = r / a.nslots;
            const int iy = min(max(r0 + r + P, 0), a.in.texH - 1);
            // a staged plane starts at (address of its run) & 15: in mode 0 that is 0 or 8 bytes, alternating with
            // the plane when a plane holds an odd number of texels
            const uint32_t sh0 = (base15 + (uint32_t)((((unsigned long long)iy * a.in.texW + xa) * (unsigned long long)bpp) & 15ull)) & 15u;
            const uint32_t sh1 = (sh0 + plane15) & 15u;
            const uint32_t hiDelta = (uint32_t)stagePlane + sh1 - sh0;   // first plane texel -> second plane texel of a chunk
            [[maybe_unused]] long long pt = PROF_T();
            mbar_wait(&landed[st], use & 1);
            PROF_ADD(pLdWait, pt);
            if (tg == 0) TRACE(0, r, 0);
            pt = PROF_T();
            mbar_wait(&empty[slotIdx], (fill & 1) ^ 1);
            PROF_ADD(pFinWaitE, pt);
            pt = PROF_T();
            const uint32_t stgA = smem_u32(sStage) + (uint32_t)st * (uint32_t)a.stageBytes + sh0;
            const uint32_t slotA = smem_u32(sRing) + (uint32_t)slotIdx * (uint32_t)a.slotBytes;
            const bool mirror = slotIdx < a.nmirror;
            // The body is written without data-dependent branches (the loaders are bound by instruction latency, not
            // by bandwidth): whole units of groupThreads items in batches of kFinBatch, then single units, then the
            // guarded partial unit.  `nv2` / `mir` are uniform per row and select one of four straight-line variants.
            auto run_row = [&](auto nv2, auto mir) {
                constexpr bool NV2 = decltype(nv2)::value, MIR = decltype(mir)::value;
                auto batch = [&](auto nb, int it0, bool guard) {
                    constexpr int B = decltype(nb)::value;
                    int2 e[B];
                    bool ok[B];
#pragma unroll
                    for (int u = 0; u < B; u++) {
                        const int it = it0 + groupThreads * u;
                        ok[u] = !guard || it < a.nitems;
                        e[u] = ok[u] ? sTab[it] : make_int2(0, 0);
                    }
                    if (MODE == 0) {
                        uint2 lo[B], hi[B];
#pragma unroll
                        for (int u = 0; u < B; u++) {
                            const uint32_t la = stgA + (uint32_t)(e[u].x & ~1);
                            lo[u] = lds64(la);
                            hi[u] = lds64((e[u].x & 1) ? la + hiDelta : zeroA);   // select, not a branch
                        }
#pragma unroll
                        for (int u = 0; u < B; u++) {
                            if (ok[u]) {
                                const uint4 raw = make_uint4(lo[u].x, lo[u].y, hi[u].x, hi[u].y);
                                const uint4 av = act_h8_t<ACT>(raw, a.act);
                                const uint32_t d = slotA + (uint32_t)e[u].y;
                                sts128(d, av);
                                if (NV2) sts128(d + (uint32_t)a.verBytes, raw);          // taps that bypass the activation
                                if (MIR) {
                                    sts128(d + (uint32_t)mirrorOff, av);
                                    if (NV2) sts128(d + (uint32_t)(mirrorOff + a.verBytes), raw);
                                }
                            }
                        }
                    } else {
                        float4 v[B];
#pragma unroll
                        for (int u = 0; u < B; u++) {
                            const uint32_t q = stgA + (uint32_t)e[u].x;
                            if (fastIn) {                                   // fp32 RGB upload texture
                                v[u] = make_float4(lds32f(q), lds32f(q + 4), lds32f(q + 8), 0.f);
                            } else if (f16in) {
                                const __half *hq = reinterpret_cast<const __half *>(__cvta_shared_to_generic(q));
                                v[u] = make_float4(__half2float(hq[0]), pk > 1 ? __half2float(hq[1]) : 0.f, pk > 2 ? __half2float(hq[2]) : 0.f,
                                                   pk > 3 ? __half2float(hq[3]) : 0.f);
                            } else {
                                v[u] = make_float4(lds32f(q), pk > 1 ? lds32f(q + 4) : 0.f, pk > 2 ? lds32f(q + 8) : 0.f, pk > 3 ? lds32f(q + 12) : 0.f);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < B; u++) {
                            if (ok[u]) {
                                const float4 w = act_f4_t<ACT>(v[u], a.act);
                                const uint2 h = make_uint2(pack_half2(w.x, w.y), pack_half2(w.z, w.w));
                                sts64(slotA + (uint32_t)e[u].y, h);
                                if (MIR) sts64(slotA + (uint32_t)(mirrorOff + e[u].y), h);
                            }
                        }
                    }
                };
                const int unitsEnd = (a.nitems / groupThreads) * groupThreads;          // items in whole units
                int it0 = tg;
                for (; it0 + (kFinBatch - 1) * groupThreads < unitsEnd; it0 += kFinBatch * groupThreads) batch(std::integral_constant<int, kFinBatch>{}, it0, false);
                for (; it0 + groupThreads < unitsEnd; it0 += 2 * groupThreads) batch(std::integral_constant<int, 2>{}, it0, false);
                for (; it0 < unitsEnd; it0 += groupThreads) batch(std::integral_constant<int, 1>{}, it0, false);
                if (unitsEnd < a.nitems) batch(std::integral_constant<int, 1>{}, it0, true);
            };
            if (a.nver == 2) {
                if (mirror) run_row(std::true_type{}, std::true_type{});
                else run_row(std::true_type{}, std::false_type{});
            } else {
                if (mirror) run_row(std::false_type{}, std::true_type{});
                else run_row(std::false_type{}, std::false_type{});
            }
            PROF_ADD(pFinProc, pt);
            pt = PROF_T();
            // every thread orders its stores (generic proxy) before the tensor core's reads (async proxy); after the
            // group barrier the leader publishes the row and refills the stage, which nobody reads any more
            fence_proxy_async();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(groupThreads) : "memory");
            if (leader) {
                if (tg == 0) {
                    mbar_arrive(&full[slotIdx]);
                    TRACE(0, r, 1);
                }
                if (r + a.nstages < R) issue_row(r + a.nstages);
            }
            PROF_ADD(pFinFence, pt);
        }
#ifdef FYN_TC_PROFILE
        if (blockIdx.x == 0 && tg == 0)
            printf("[tc prof] loader group %d: total %lld waitLanded %lld waitEmpty %lld process %lld publish %lld\n", g, (long long)(clock64() - pLdStart),
                   pLdWait, pFinWaitE, pFinProc, pFinFence);
#endif
    } else if (warp >= kMmaWarp) {
        // ===================== MMA issuers =====================
        // Two warps: warp w issues jobs w, w + 2, ... into TMEM buffer w, so the per-job bookkeeping of one (barrier
        // waits, commits, ring arithmetic: ~0.9k cycles, measured) hides behind the other's tcgen05.mma stream.  Each
        // warp runs its loop with warp-uniform values (step table in parameter space, ring position derived from
        // block-uniform data) so the descriptors are built in uniform registers; one elected lane waits on the
        // barriers and issues the tcgen05 instructions.
        // A ring slot goes back to the loaders when BOTH warps have retired the row (empty barriers count 2): after
        // its job q a warp releases the rows below the window of its next job q + 2 that it has already waited for.
        // Neither warp can run a full ring ahead of the other: a row can only be published into a slot that both
        // have released, and a warp only releases rows whose "full" phase it has observed.
        const int mw = warp - kMmaWarp;
        const uint32_t rbase16 = smem_u32(sRing) >> 4, slot16 = (uint32_t)a.slotBytes >> 4;
        const uint32_t bconst = (smem_u32(sW) >> 4) | ((a.b_lbo >> 4) << 16);   // B descriptor low word minus the step offset
        const uint64_t hiA = (uint64_t)((128u >> 4) | (1u << 14)) << 32;          // SBO = 128 B, descriptor version 1
        const uint32_t onesDesc = ((smem_u32(sW) + a.onesOff) >> 4) | ((16u >> 4) << 16);   // second K chunk = next row (zero weights)
        const int R = r1 - r0 + 1;
        RingPos win{0, 0};     // ring position of the current window's first row
        win.advance(a.rowAdvance * mw, a.nslots);
        RingPos nxt{0, 0};     // ring position of the first row this warp has not waited for yet
        RingPos rel{0, 0};     // ring position of the first row this warp has not released yet
        int waited = 0, released = 0;
        PROF_DECL(pWaitT); PROF_DECL(pWaitF); PROF_DECL(pIssue); PROF_DECL(pCommit); PROF_DECL(pBlock);
        [[maybe_unused]] const long long pStart = PROF_T();
        for (int q = mw; q < njobs; q += kMmaWarps) {
            const int use = q >> 1;
            const int first = a.rowAdvance * q;          // window start relative to r0
            const int needTo = first + a.nrows;          // rows [waited, needTo) have not been waited for by this warp
            // Rows [released, relTo) retire now: everything below the window of this warp's next job -- but never a row
            // this warp has not waited for yet (windows shorter than two row advances): releasing it could let its
            // slot be refilled, and its barrier pass a second phase, before this warp's parity wait for the first.
            const int relTo = min((q + kMmaWarps < njobs) ? first + kMmaWarps * a.rowAdvance : R, needTo);
            if (elect_one()) {
                [[maybe_unused]] long long pt = PROF_T();
                [[maybe_unused]] const long long ptB = pt;
                mbar_wait(&tempty[mw], (use & 1) ^ 1);
                PROF_ADD(pWaitT, pt);
                pt = PROF_T();
                RingPos w = nxt;
                for (int k = waited; k < needTo; k++) {
                    mbar_wait(&full[w.slot], w.fill & 1);
                    w.advance(1, a.nslots);
                }
                if (q == mw) mbar_wait(wbar, 0);       // weight image
                PROF_ADD(pWaitF, pt);
                pt = PROF_T();
                tc_fence_after();
                TRACE(1, q, 0);
                const uint32_t d = tmem + (uint32_t)mw * 64u;
                const uint32_t winBase = rbase16 + (uint32_t)win.slot * slot16;
#pragma unroll 4
                for (int s = 0; s < a.nsteps; s++) {
                    const TcStep st = a.steps[s];
                    umma_f16(d, hiA | (uint64_t)(st.a_lo + winBase), hiA | (uint64_t)(st.b_off16 + bconst), a.idesc, st.accumulate);
                }
                // folded bias: A = rows of (1, 1, 0, ...), B = (bias_hi, bias_lo, 0, ...) per output column
                if (a.biasFolded) umma_f16(d, hiA | (uint64_t)onesDesc, hiA | (uint64_t)(a.biasB16 + bconst), a.idesc, 1u);
                PROF_ADD(pIssue, pt);
                TRACE(1, q, 1);
                pt = PROF_T();
                umma_commit(&tfull[mw]);
                RingPos rp = rel;
                for (int i = released; i < relTo; i++) {
                    umma_commit(&empty[rp.slot]);
                    rp.advance(1, a.nslots);
                }
                PROF_ADD(pCommit, pt);
                PROF_ADD(pBlock, ptB);
            }
            if (needTo > waited) {
                nxt.advance(needTo - waited, a.nslots);
                waited = needTo;
            }
            if (relTo > released) {
                rel.advance(relTo - released, a.nslots);
                released = relTo;
            }
            win.advance(kMmaWarps * a.rowAdvance, a.nslots);
            __syncwarp();
        }
        // a warp without jobs (single-job strips) still owes its share of every release: R <= nslots then, no slot is reused
        if (mw >= njobs && elect_one())
            for (int i = 0; i < R; i++) mbar_arrive(&empty[i % a.nslots]);
#ifdef FYN_TC_PROFILE
        if (blockIdx.x == 0 && elect_one())
            printf("[tc prof] mma %d: jobs %d steps %d total %lld waitTempty %lld waitFull %lld issue %lld commit %lld electedBlock %lld (cycles)\n", mw, njobs, a.nsteps,
                   (long long)(clock64() - pStart), pWaitT, pWaitF, pIssue, pCommit, pBlock);
#endif
    } else {
        // ===================== epilogue: warps 0-7 =====================
        // Thread = job column (TMEM lane); the two warps of a lane quarter split the accumulator columns evenly in
        // units of 8 columns (an "octet" = two 4-channel texels): warp half h takes octets [h * N/16, (h+1) * N/16).
        // Column order is [fy][plane][fx][channel] (see the weight image), so with 2 or 4 phases along x an octet is two
        // horizontally adjacent output texels of one plane -> one 16-byte store.
        const int m = threadIdx.x & 127;
        const int chalf = warp >> 2;
        const int jx = j0 + m;
        const bool valid = jx < a.Wj;
        const int nOct = (a.N >> 3) / (a.epiWarps >> 2);   // octets per thread (1..4): the warps of a lane quarter split the columns
        const int ppp = a.planesPerPhase;
        // per texel, hoisted out of the job loop: output element offset relative to the job's first output texel
        // (-1 = nothing to store) and the output plane (bias / scale / residual index)
        int poff[4][2], pidx[4][2];
#pragma unroll
        for (int o = 0; o < 4; o++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int pn = (chalf * nOct + o) * 2 + k;   // stacked plane index
                const int fx = pn % a.opx, tq = pn / a.opx;
                const int fy = tq / ppp, p = tq - fy * ppp;
                pidx[o][k] = p;
                poff[o][k] = (o < nOct && fy < a.opy && valid) ? (int)((long long)p * a.out.planeElems + (long long)fy * a.out.texW * 4 + fx * 4) : -1;
            }
        __half *outp = reinterpret_cast<__half *>(a.out.ptr) + (long long)n * a.out.imageElems + ((long long)a.outP * a.out.texW + a.outP + a.opx * jx) * 4;
        const long long outRow = (long long)a.out.texW * 4 * a.opy;   // elements per job row
        // 16-byte stores need every job's first texel on a 16-byte boundary in every output row
        const bool aligned16 = (a.outP & 1) == 0 && (a.out.texW & 1) == 0 && (a.out.planeElems & 7) == 0 && (a.out.imageElems & 7) == 0;
        const bool wide = aligned16 && (a.opx == 2 || a.opx == 4);
        const __half *resp = reinterpret_cast<const __half *>(a.res.ptr) + (long long)n * a.res.imageElems + ((long long)a.resP * a.res.texW + a.resP + jx) * 4;
        const int resPlane = (int)a.res.planeElems;
        PROF_DECL(pEpWait);
        [[maybe_unused]] const long long pEpStart = PROF_T();
        // The epilogue parameters arrive with the weight image (no global load in the prologue); they are moved out of
        // the weight region once: generic loads from the region the tensor core streams its B operand from were
        // measured to slow the epilogue by ~700 cycles per job (conv1: 36.5 -> 50.0 us).
        mbar_wait(wbar, 0);
        if (threadIdx.x < 32) sEpi[threadIdx.x] = reinterpret_cast<const float4 *>(sW + a.epiOff)[threadIdx.x];
        asm volatile("bar.sync 15, %0;" ::"r"(a.epiWarps * 32) : "memory");
        for (int q = 0; q < njobs; q++) {
            const int buf = q & 1, use = q >> 1;
            const int i = ja + q;                    // job row
            // residual texels are fetched before waiting for the accumulator so their latency hides behind the MMAs
            // (requesting the next job's texels one iteration ahead was measured slower: 18.2 vs 17.3 us on res+residual)
            uint2 rres[4][2];
            if (RES == 1) {
                const __half *rp = resp + (long long)i * a.res.texW * 4;
#pragma unroll
                for (int o = 0; o < 4; o++)
#pragma unroll
                    for (int k = 0; k < 2; k++)
                        if (poff[o][k] >= 0) rres[o][k] = __ldg(reinterpret_cast<const uint2 *>(rp + pidx[o][k] * resPlane));
            }
            [[maybe_unused]] const long long pt = PROF_T();
            mbar_wait(&tfull[buf], use & 1);
            PROF_ADD(pEpWait, pt);
            if (threadIdx.x == 0) TRACE(2, q, 0);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * 64u + (uint32_t)(chalf * nOct) * 8u;
            uint32_t acc[4][8];
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (o < nOct) tmem_ld8(taddr + o * 8, acc[o]);
            tmem_ld_wait();
            // accumulators are in registers: hand the TMEM buffer back before the global-memory work
            tc_fence_before();
            mbar_arrive(&tempty[buf]);
            __half *orow = outp + (long long)i * outRow;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                if (o >= nOct) break;
                uint2 t[2];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    t[k] = make_uint2(0u, 0u);
                    if (poff[o][k] >= 0) {
                        const int p = pidx[o][k];
                        float4 v = make_float4(__uint_as_float(acc[o][4 * k + 0]), __uint_as_float(acc[o][4 * k + 1]), __uint_as_float(acc[o][4 * k + 2]),
                                               __uint_as_float(acc[o][4 * k + 3]));
                        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (!a.biasFolded) {               // post-BN scale (or long step tables): bias and scale from shared memory
                            const float4 bi = sEpi[p];
                            sc = sEpi[16 + p];
                            v = make_float4(fmaf(v.x, sc.x, bi.x), fmaf(v.y, sc.y, bi.y), fmaf(v.z, sc.z, bi.z), fmaf(v.w, sc.w, bi.w));
                        }
                        if (RES != 0) {
                            float4 rs;
                            if (RES == 1) {
                                const uint2 raw = rres[o][k];
                                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                                rs = make_float4(f0.x, f0.y, f1.x, f1.y);
                            } else {
                                const int pn = (chalf * nOct + o) * 2 + k, fx = pn % a.opx, fy = (pn / a.opx) / ppp;
                                rs = fyn_fetch(a.res, n, p, a.resP + a.opx * jx + fx, a.resP + a.opy * i + fy);
                            }
                            if (a.reluRes) rs = make_float4(fmaxf(rs.x, 0.f), fmaxf(rs.y, 0.f), fmaxf(rs.z, 0.f), fmaxf(rs.w, 0.f));
                            if (a.bnRes) rs = make_float4(rs.x * sc.x, rs.y * sc.y, rs.z * sc.z, rs.w * sc.w);
                            v.x += rs.x;
                            v.y += rs.y;
                            v.z += rs.z;
                            v.w += rs.w;
                        }
                        t[k] = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
                        if (EPI == FYN_EPILOGUE_SIGMOID) {
                            // fused FunctionLayer, evaluated on the fp16-rounded convolution result like the unfused pair
                            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&t[k].x));
                            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&t[k].y));
                            t[k] = make_uint2(pack_half2(fyn_sigmoid(f0.x), fyn_sigmoid(f0.y)), pack_half2(fyn_sigmoid(f1.x), fyn_sigmoid(f1.y)));
                        }
                    }
                }
                if ((a.debug & 1) && t[0].x != 0x7fff7fffu) continue;   // ablation: no stores (never true for real data paths)
                // with phases along x the two texels are adjacent in memory: one 16-byte store where alignment allows
                if (wide) {
                    if (poff[o][0] >= 0) *reinterpret_cast<uint4 *>(orow + poff[o][0]) = make_uint4(t[0].x, t[0].y, t[1].x, t[1].y);
                } else {
                    if (poff[o][0] >= 0) *reinterpret_cast<uint2 *>(orow + poff[o][0]) = t[0];
                    if (poff[o][1] >= 0) *reinterpret_cast<uint2 *>(orow + poff[o][1]) = t[1];
                }
            }
            if (threadIdx.x == 0) TRACE(2, q, 1);
        }
#ifdef FYN_TC_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0)
            printf("[tc prof] epilogue: total %lld waitTfull %lld\n", (long long)(clock64() - pEpStart), pEpWait);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 128);
#ifdef FYN_TC_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("[tc prof] kernel: prologue %lld depwait %lld body %lld (cycles), grid %d\n", pK1 - pK0, pK2 - pK1, (long long)(clock64() - pK2), (int)gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // event trace of block 0 (cycles since kernel start)
        const int R = r1 - r0 + 1;
        for (int r = 0; r < R && r < 64; r++) printf("[tc trace] row %2d landed %6lld published %6lld\n", r, g_trace[0][r][0], g_trace[0][r][1]);
        for (int q = 0; q < njobs && q < 64; q++)
            printf("[tc trace] job %2d mma issue %6lld - %6lld   epilogue %6lld - %6lld\n", q, g_trace[1][q][0], g_trace[1][q][1], g_trace[2][q][0], g_trace[2][q][1]);
    }
#endif
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side: plan = window + MMA step table + weight image
// ---------------------------------------------------------------------------------------------

namespace {

// one source position of the GEMM: (version, window row dy, source offset dx) with the weights of every phase that
// reads it, w[(phase * Cq + co) * Ci + ci] (merged taps summed)
struct Position {
    int ver, dy, dx;
    std::vector<float> w;
};

struct Geometry {
    int mode = 0, opx = 1, opy = 1, rowAdvance = 1, nver = 1, N = 16, Cq = 4, nchunks = 1, rowpx = 0, x_lead = 0, ds = 1;
    int dxMin = 0, dxMax = 0, dyMin = 0, dyMax = 0;
    int nslots = 0, slotBytes = 0, verBytes = 0, nsteps = 0, stageBytes = 0, nstages = 0, nitems = 0, finGroups = 2, nmirror = 0, epiWarps = 8;
    bool biasFold = false;
    size_t wbytes = 0, smem = 0;
    std::vector<Position> pos;
    bool ok = false;
};

int ifloor(float v) { return (int)floorf(v); }
int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

Position &find_pos(std::vector<Position> &pos, int ver, int dy, int dx, size_t wsize, bool wantW) {
    for (Position &q : pos)
        if (q.ver == ver && q.dy == dy && q.dx == dx) return q;
    pos.push_back({ver, dy, dx, {}});
    if (wantW) pos.back().w.assign(wsize, 0.f);
    return pos.back();
}

// Builds the positions (and stacked, merged weights when wb is given) and the shared-memory geometry.
// `ys` = output job rows stacked along N (1 or 2): fewer, fatter jobs amortise the per-job costs of every role
// (barrier hand-overs, commits, epilogue set-up) and reuse the window rows the stacked outputs share.
Geometry plan_geometry(const fyn_conv_desc *d, const float *wb, int ys) {
    Geometry g;
    if (d->flags & FYN_FLAG_DEEP) return g;          // deep-tiled family: not yet
    if (d->dilation != 1) return g;
    const int K = d->kernel, mh = (K - 1) / 2, Ci = d->in_channels, Co = d->out_channels;
    if (K > 9) return g;
    g.Cq = 4 * ((Co + 3) / 4);
    const bool hasAct = (d->flags & (FYN_FLAG_PRE_RELU | FYN_FLAG_PRE_CLIP)) != 0;
    const float *W = wb ? wb + Co : nullptr;          // [Co][K][K][Ci]
    std::vector<int> tapx(K), tapy(K);
    for (int k = 0; k < K; k++) tapx[k] = tapy[k] = k - mh;
    g.dxMin = g.dyMin = 1 << 20;
    g.dxMax = g.dyMax = -(1 << 20);
    if (d->fractional) {
        if (Ci < 8) return g;
        const float s = d->source_step, u = s * (float)d->downsample;
        int p;
        if (u == 1.0f) p = 1;
        else if (u == 0.5f) p = 2;
        else return g;                                // other ratios: direct kernel
        if (K == 3 && (d->quirks & FYN_QUIRK_FRAC3_ASYM)) {
            tapx[0] = -2;
            tapx[1] = -1;
            tapx[2] = 0;
        }
        const bool actFirstOnly = hasAct && (d->quirks & FYN_QUIRK_FRAC_ACT_FIRST);
        g.opx = p;
        g.opy = p * ys;                               // p output rows per source row, ys source rows per job
        g.rowAdvance = ys;
        g.nver = actFirstOnly ? 2 : 1;
        g.ds = 1;
        g.mode = 0;
        const size_t wsize = (size_t)g.opy * p * g.Cq * Ci;
        for (int fy = 0; fy < g.opy; fy++)
            for (int fx = 0; fx < p; fx++)
                for (int ky = 0; ky < K; ky++)
                    for (int kx = 0; kx < K; kx++) {
                        // delta = floor(s*(ds*phi + 0.5 + tap)) -- gpu/vanilla/convlayerbase_vanilla.cpp:352-371 + fraconv*.frag
                        const int dy = ifloor(s * ((float)(d->downsample * fy) + 0.5f + (float)tapy[ky]));
                        const int dx = ifloor(s * ((float)(d->downsample * fx) + 0.5f + (float)tapx[kx]));
                        const int ver = (actFirstOnly && kx > 0) ? 1 : 0;
                        Position &q = find_pos(g.pos, ver, dy, dx, wsize, wb != nullptr);
                        if (wb) {
                            const int phase = fy * p + fx;
                            for (int o = 0; o < Co; o++)
                                for (int c = 0; c < Ci; c++)
                                    q.w[((size_t)phase * g.Cq + o) * Ci + c] += W[(((size_t)o * K + ky) * K + kx) * Ci + c];
                        }
                    }
    } else if (Ci <= 4 && d->downsample == 1 && K >= 3) {
        // pixel-pair mode: one GEMM row = 4 adjacent output pixels (phases), chunks = pixel pairs.  Position.dx is the
        // chunk offset e relative to chunk 2*j; element (half h, channel c) of chunk e belongs to tap kx = 2e + h - fx + mh.
        g.mode = 1;
        g.opx = 4;
        g.opy = ys;
        g.rowAdvance = ys;
        g.ds = 1;
        const size_t wsize = (size_t)ys * 4 * g.Cq * 8;    // [phase][co][8 chunk elements]
        int e0 = floordiv2(-mh);
        if (e0 & 1) e0 -= 1;   // the slot starts on an even chunk so that chunk parity == (e - e0) parity; extra chunk has zero weights
        for (int fy = 0; fy < ys; fy++)
            for (int ky = 0; ky < K; ky++)
                for (int e = e0; e <= floordiv2(3 + K - 1 - mh); e++) {
                    Position &q = find_pos(g.pos, 0, fy + ky - mh, e, wsize, wb != nullptr);
                    if (wb)
                        for (int fx = 0; fx < 4; fx++)
                            for (int o = 0; o < Co; o++)
                                for (int h = 0; h < 2; h++)
                                    for (int c = 0; c < Ci; c++) {
                                        const int kx = 2 * e + h - fx + mh;
                                        if (kx >= 0 && kx < K)
                                            q.w[((size_t)(fy * 4 + fx) * g.Cq + o) * 8 + h * 4 + c] = W[(((size_t)o * K + ky) * K + kx) * Ci + c];
                                    }
                }
    } else {
        if (d->downsample != 1 && d->downsample != 2) return g;
        // (1x1 kernels -- vanilla::ConvLayer1x1, gpu/vanilla/convlayer1x1_vanilla.cpp:81-150 -- are windows of one row and one
        // tap; inputs with fewer than 8 channels are a single, partly empty chunk)
        g.mode = 0;
        g.opx = 1;
        g.opy = ys;
        g.rowAdvance = ys * d->downsample;
        g.ds = d->downsample;
        const size_t wsize = (size_t)ys * g.Cq * Ci;
        for (int fy = 0; fy < ys; fy++)
            for (int ky = 0; ky < K; ky++)
                for (int kx = 0; kx < K; kx++) {
                    // output row ys*i + fy reads input rows ds*(ys*i + fy) + ky - mh: window row fy*ds + ky - mh
                    Position &q = find_pos(g.pos, 0, fy * d->downsample + ky - mh, kx - mh, wsize, wb != nullptr);
                    if (wb)
                        for (int o = 0; o < Co; o++)
                            for (int c = 0; c < Ci; c++) q.w[((size_t)fy * g.Cq + o) * Ci + c] = W[(((size_t)o * K + ky) * K + kx) * Ci + c];
                }
    }
    const int ntot = g.opx * g.opy * g.Cq;
    g.N = ((ntot + 15) / 16) * 16;
    if (g.N > 64) return g;                           // one 64-column TMEM buffer per job
    for (const Position &q : g.pos) {
        g.dxMin = std::min(g.dxMin, q.dx);
        g.dxMax = std::max(g.dxMax, q.dx);
        g.dyMin = std::min(g.dyMin, q.dy);
        g.dyMax = std::max(g.dyMax, q.dy);
    }
    const int nrows = g.dyMax - g.dyMin + 1, span = g.dxMax - g.dxMin;
    if (nrows > kMaxRows) return g;
    // slot geometry
    if (g.mode == 1) {
        // chunks 2*j + e, e in [dxMin, dxMax]: per parity (128 + span/2 + 1) chunks, rounded up
        const int halfChunks = ((kTileM + span / 2 + 2) + 3) & ~3;
        g.nchunks = 1;
        g.rowpx = 2 * halfChunks;                     // chunks per slot
        g.x_lead = -2 * g.dxMin;                      // pixels (dxMin is even-aligned below)
        if (g.dxMin & 1) return g;
        g.verBytes = g.rowpx * 16;
        if (2 * g.rowpx > kMaxItems) return g;
        g.nitems = 2 * g.rowpx;
        g.stageBytes = (2 * g.rowpx * 16 + 16 + 15) & ~15;             // raw texels: up to 16 bytes each (fp32 RGBA)
    } else {
        g.nchunks = ((Ci + 3) / 4 + 1) / 2;
        g.x_lead = -g.dxMin;
        if (g.ds == 1) g.rowpx = ((kTileM + span + 1) + 3) & ~3;
        else g.rowpx = 2 * (((kTileM + span / 2 + 1) + 3) & ~3);
        g.verBytes = g.nchunks * g.rowpx * 16;
        if (g.rowpx * g.nchunks > kMaxItems || (Ci + 3) / 4 > 32) return g;   // one bulk copy per plane and lane
        g.nitems = g.rowpx * g.nchunks;
        g.stageBytes = ((Ci + 3) / 4) * ((g.rowpx * 8 + 16 + 15) & ~15);   // per plane: the texel run + alignment slack
    }
    g.slotBytes = g.verBytes * g.nver;
    // steps: chunks of the same window row are paired in address order
    int nsteps = 0;
    for (int dy = g.dyMin; dy <= g.dyMax; dy++) {
        int chunks = 0;
        for (const Position &q : g.pos)
            if (q.dy == dy) chunks += g.nchunks;
        nsteps += (chunks + 1) / 2;
    }
    g.nsteps = nsteps;
    if (nsteps > kMaxSteps) return g;
    g.wbytes = (size_t)nsteps * 2 * g.N * 16;
    // Folded bias (layers without post-BN scale whose epilogue, not the MMA stream, is the busy role -- short step
    // tables): one more MMA step adds the bias to the accumulator, so the epilogue only converts and stores.  The
    // weight image grows by the step's B chunks and by the A operand of that step, 130 rows of (1, 1, 0, ..., 0).
    g.biasFold = !(d->flags & FYN_FLAG_POST_BATCHNORM) && nsteps <= 24;
    if (g.biasFold) g.wbytes += (size_t)2 * g.N * 16 + 130 * 16;
    g.wbytes += 32 * 16;                              // epilogue parameters ride along
    // Loader groups, ring slots and staged rows.  The ring holds the window, the rows the next job adds and one more
    // job's worth of slack, rounded up to a multiple of the group count G (slot and stage ownership, see the kernel).
    // Measured on B200 (StyleNet layers, 1524x1856): rows finished concurrently matter more than staged rows per
    // group (res 3x3 40->40: G=1 20.6 us, G=2 18.2 us, G=4 16.1 us), so: four groups if they fit at all, else two.
    const size_t budget = 220 * 1024;
    // (FYN_TC_GROUPS=1|2|4 overrides the choice: tuning knob)
    // (pixel-pair rows are light: eight single-warp groups were measured faster there, conv1 38.1 -> 35.9 us, and slower
    // or not fitting elsewhere)
    int gmax = (g.mode == 1) ? 8 : 4, gmin = 2;
    // Warp split: 12 epilogue + 4 loader warps where the epilogue is the busy role and the rows are light for the
    // loaders -- measured: the 3x3 40-channel conv with residual 17.2 -> 15.3 us; without residual no change, and
    // loader-heavy layers get slower (conv1 36 -> 43 us).  The accumulator columns must split three ways.
    // (FYN_TC_EPI=8|12: tuning knob)
    g.epiWarps = (g.mode == 0 && (d->flags & FYN_FLAG_RESIDUAL_INPUT) && ((g.N >> 3) % 3) == 0 && g.nitems * g.rowAdvance <= 800) ? 12 : 8;
    if (const char *e = getenv("FYN_TC_EPI")) {
        const int want = atoi(e);
        if (want == 8 || (want == 12 && ((g.N >> 3) % 3) == 0)) g.epiWarps = want;
    }
    gmax = std::min(gmax, kWorkWarps - g.epiWarps);
    if (const char *e = getenv("FYN_TC_GROUPS")) gmax = gmin = std::max(1, std::min(8, atoi(e)));
    // Ring size: the window, the rows the next job adds and one more job's worth of slack ("full"); if that does not
    // fit with four groups, a ring with about one job of look-ahead ("tight") is considered as well.  Lower bound
    // for progress with two MMA warps: max(nrows, 2 * rowAdvance) slots (see the release rule in the kernel).
    const int needFull = nrows + 2 * g.rowAdvance;
    const int needTight = std::max(nrows, 2 * g.rowAdvance) + std::max(g.rowAdvance, 3);   // (a single slot of look-ahead was measured slower: res 13.4 -> 16.2 us)
    // Per group count (from gmax down): the full ring if every group still gets two staged rows (a copy in flight while
    // it finishes a row), else the tight ring if that one does, else whichever fits.  Measured on the stacked stride-2
    // conv: tight ring + 8 stages 31.3 us, full ring + 4 stages 37.5 us.
    auto fit = [&](int G, int need, int &nslots, int &nmirror, size_t &fixed) -> int {
        nslots = ((need + G - 1) / G) * G;
        // slots mirrored behind the ring so that every window is contiguous: windows start at multiples of rowAdvance
        nmirror = (nslots % g.rowAdvance == 0) ? std::max(0, nrows - g.rowAdvance) : nrows - 1;
        fixed = ((g.wbytes + 127) & ~(size_t)127) + (size_t)(nslots + nmirror) * g.slotBytes + 32 * 16 + (size_t)g.nitems * 8 +
                (2 * nslots + 5 + kMaxStages) * 8 + 16 + 16;   // + the 16 zero bytes
        if (fixed + (size_t)G * g.stageBytes > budget) return 0;
        return (int)std::min<size_t>(kMaxStages / G, (budget - fixed) / ((size_t)G * g.stageBytes));   // staged rows per group
    };
    for (int G = gmax; G >= gmin && !g.ok; G >>= 1) {
        int nsF, nmF, nsT, nmT;
        size_t fxF, fxT;
        const int mF = fit(G, needFull, nsF, nmF, fxF);
        const int mT = (needTight < needFull) ? fit(G, needTight, nsT, nmT, fxT) : 0;
        const bool tight = (mF < 2 && mT >= 2) || (mF == 0 && mT > 0);
        const int m = tight ? mT : mF;
        if (m == 0) continue;
        g.finGroups = G;
        g.nslots = tight ? nsT : nsF;
        g.nmirror = tight ? nmT : nmF;
        g.nstages = m * G;
        g.smem = (tight ? fxT : fxF) + (size_t)g.nstages * g.stageBytes;
        g.ok = true;
    }
    return g;
}

}  // namespace

namespace {

using TcKernel = void (*)(TcArgs);

// kernel instantiations: [mode][act][res] without fused function, [mode][act] with the fused sigmoid (no residual)
TcKernel tc_kernel(int mode, int act, int res, int epi) {
    static const TcKernel table[2][3][3] = {
        {{k_conv_tc<0, 0, 0, 0>, k_conv_tc<0, 0, 1, 0>, k_conv_tc<0, 0, 2, 0>},
         {k_conv_tc<0, 1, 0, 0>, k_conv_tc<0, 1, 1, 0>, k_conv_tc<0, 1, 2, 0>},
         {k_conv_tc<0, 2, 0, 0>, k_conv_tc<0, 2, 1, 0>, k_conv_tc<0, 2, 2, 0>}},
        {{k_conv_tc<1, 0, 0, 0>, k_conv_tc<1, 0, 2, 0>, k_conv_tc<1, 0, 2, 0>},
         {k_conv_tc<1, 1, 0, 0>, k_conv_tc<1, 1, 2, 0>, k_conv_tc<1, 1, 2, 0>},
         {k_conv_tc<1, 2, 0, 0>, k_conv_tc<1, 2, 2, 0>, k_conv_tc<1, 2, 2, 0>}}};
    static const TcKernel sig[2][3] = {{k_conv_tc<0, 0, 0, 1>, k_conv_tc<0, 1, 0, 1>, k_conv_tc<0, 2, 0, 1>},
                                       {k_conv_tc<1, 0, 0, 1>, k_conv_tc<1, 1, 0, 1>, k_conv_tc<1, 2, 0, 1>}};
    return epi ? sig[mode][act] : table[mode][act][res];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and device, not per launch: keep it at the largest
// footprint any plan has needed so far
int tc_ensure_smem(TcKernel fn, int mode, int act, int res, int epi, int device, size_t bytes) {
    static size_t cur[64][2][3][4] = {};
    static std::mutex lock;                      // ops may be created / run from one thread per context
    std::lock_guard<std::mutex> guard(lock);
    size_t &c = cur[device & 63][mode][act][epi ? 3 : res];
    if (bytes > c) {
        FYN_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        c = bytes;
    }
    return FYN_OK;
}

}  // namespace

int fyn_conv_tc_supported(const fyn_conv_desc *d, int) { return plan_geometry(d, nullptr, 1).ok ? 1 : 0; }

int fyn_conv2d_plan_query(const fyn_conv_desc *desc, int stack_rows, fyn_conv_plan_info *info) {
    if (!desc || !info || stack_rows < 1 || stack_rows > 2) FYN_FAIL(FYN_ERR_INVALID, "plan query: bad argument");
    const Geometry g = plan_geometry(desc, nullptr, stack_rows);
    if (!g.ok) FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 family does not cover this conv");
    *info = fyn_conv_plan_info{};
    info->mode = g.mode;
    info->n = g.N;
    info->steps = g.nsteps;
    info->window_rows = g.dyMax - g.dyMin + 1;
    info->row_advance = g.rowAdvance;
    info->phases_x = g.opx;
    info->phases_y = g.opy;
    info->ring_slots = g.nslots;
    info->mirror_slots = g.nmirror;
    info->slot_bytes = g.slotBytes;
    info->staged_rows = g.nstages;
    info->stage_bytes = g.stageBytes;
    info->row_items = g.nitems;
    info->loader_groups = g.finGroups;
    info->epilogue_warps = g.epiWarps;
    info->bias_folded = g.biasFold ? 1 : 0;
    info->weight_image_bytes = g.wbytes;
    info->shared_bytes = g.smem;
    return FYN_OK;
}

namespace {

// kernel arguments (geometry part) and the weight image of one plan
int build_plan(const Geometry &g, const fyn_conv_desc &d, const float *wb, TcArgs &a, std::vector<__half> &img) {
    const int K = d.kernel, Ci = d.in_channels, Co = d.out_channels, N = g.N;
    const int nphase = g.opx * g.opy;
    a.N = N;
    a.nInPlanes = (Ci + 3) / 4;
    a.mode = g.mode;
    a.opx = g.opx;
    a.opy = g.opy;
    a.planesPerPhase = g.Cq / 4;
    a.rowAdvance = g.rowAdvance;
    a.dyMin = g.dyMin;
    a.nrows = g.dyMax - g.dyMin + 1;
    a.ds = g.ds;
    a.nver = g.nver;
    a.verBytes = g.verBytes;
    a.nchunks = g.nchunks;
    a.rowpx = g.rowpx;
    a.x_lead = g.x_lead;
    a.nslots = g.nslots;
    a.nmirror = g.nmirror;
    a.nsteps = g.nsteps;
    if (const char *e = getenv("FYN_TC_DEBUG_STEPS")) a.nsteps = std::max(1, std::min(g.nsteps, atoi(e)));   // ablation: wrong results, timing only
    a.stageBytes = g.stageBytes;
    a.nstages = g.nstages;
    a.finGroups = g.finGroups;
    a.epiWarps = g.epiWarps;
    a.nitems = g.nitems;
    a.slotBytes = g.slotBytes;
    a.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);  // F32 accum, F16 x F16, K-major A/B
    a.b_lbo = (uint32_t)N * 16u;
    a.wbytes = (uint32_t)g.wbytes;

    img.assign(g.wbytes / 2, __float2half(0.f));
    // byte offset inside a slot of chunk c of the position (GEMM row 0)
    auto aoff = [&](const Position &q, int c) -> uint32_t {
        const int j = q.dx - g.dxMin;   // slot pixel (mode 0) / slot chunk (mode 1) for job column 0
        if (g.mode == 1) return (uint32_t)((j & 1) * (g.rowpx / 2) * 16 + (j >> 1) * 16);
        const int px = (g.ds == 1) ? j : (j & 1) * (g.rowpx / 2) + j / 2;
        return (uint32_t)(q.ver * g.verBytes + (c * g.rowpx + px) * 16);
    };
    struct Chunk { uint32_t off; const Position *pos; int sub; };
    auto fill = [&](size_t chunkIdx, const Chunk &c) {
        for (int ph = 0; ph < nphase; ph++)
            for (int o = 0; o < Co; o++)
                for (int e = 0; e < 8; e++) {
                    float v = 0.f;
                    if (g.mode == 1) {
                        v = c.pos->w[((size_t)ph * g.Cq + o) * 8 + e];
                    } else {
                        const int ci = c.sub * 8 + e;
                        if (ci < Ci) v = c.pos->w[((size_t)ph * g.Cq + o) * Ci + ci];
                    }
                    // column order [fy][plane][fx][channel]: see the epilogue
                    const int fy = ph / g.opx, fx = ph % g.opx, ppp = g.Cq / 4;
                    const size_t col = (((size_t)fy * ppp + o / 4) * g.opx + fx) * 4 + (o & 3);
                    img[(chunkIdx * N + col) * 8 + e] = __float2half_rn(v);
                }
    };
    int s = 0;
    for (int dy = g.dyMin; dy <= g.dyMax; dy++) {
        std::vector<Chunk> chunks;
        for (const Position &q : g.pos)
            if (q.dy == dy)
                for (int c = 0; c < g.nchunks; c++) chunks.push_back({aoff(q, c), &q, c});
        std::sort(chunks.begin(), chunks.end(), [](const Chunk &x, const Chunk &y) { return x.off < y.off; });
        const uint32_t rowOff = (uint32_t)(dy - g.dyMin) * (uint32_t)g.slotBytes;
        for (size_t i = 0; i < chunks.size(); i += 2) {
            TcStep &st = a.steps[s];
            st.accumulate = s == 0 ? 0u : 1u;
            st.pad = 0;
            st.b_off16 = (uint32_t)(((size_t)s * 2 * N * 16) >> 4);
            fill((size_t)s * 2, chunks[i]);
            uint32_t lbo = 16;  // unpaired: second half reads the neighbouring chunk against all-zero weights
            if (i + 1 < chunks.size()) {
                lbo = chunks[i + 1].off - chunks[i].off;
                fill((size_t)s * 2 + 1, chunks[i + 1]);
                if (lbo == 0 || lbo >= (1u << 18)) FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 conv: operand stride %u not encodable", lbo);
            }
            st.a_lo = ((lbo >> 4) << 16) | ((rowOff + chunks[i].off) >> 4);
            s++;
        }
    }
    if (s != g.nsteps) FYN_FAIL(FYN_ERR_INVALID, "tcgen05 conv: internal step count mismatch (%d vs %d)", s, g.nsteps);
    a.biasFolded = g.biasFold ? 1 : 0;
    if (g.biasFold) {
        // bias step: B chunk 0 holds (hi, lo) halves of the fp32 bias in k = 0, 1 for every phase's column, chunk 1 is zero
        const size_t stepBase = (size_t)g.nsteps * 2 * N * 8;           // in halves
        a.biasB16 = (uint32_t)((stepBase * 2) >> 4);
        for (int ph = 0; ph < nphase; ph++)
            for (int o = 0; o < Co; o++) {
                const int fy = ph / g.opx, fx = ph % g.opx, ppp = g.Cq / 4;
                const size_t col = (((size_t)fy * ppp + o / 4) * g.opx + fx) * 4 + (o & 3);
                const __half hi = __float2half_rn(wb[o]);
                img[stepBase + col * 8 + 0] = hi;
                img[stepBase + col * 8 + 1] = __float2half_rn(wb[o] - __half2float(hi));
            }
        const size_t onesBase = stepBase + (size_t)2 * N * 8;
        a.onesOff = (uint32_t)(onesBase * 2);
        for (int r = 0; r < 130; r++) img[onesBase + (size_t)r * 8 + 0] = img[onesBase + (size_t)r * 8 + 1] = __float2half(1.f);
    }

    // epilogue parameters per output plane, [16][4] bias then [16][4] scale, at the tail of the weight image
    {
        float eb[128] = {};
        const float *bn = wb + Co + (size_t)K * K * Ci * Co;
        for (int o = 0; o < Co; o++) {
            float b = wb[o], sc = 1.f;
            if (d.flags & FYN_FLAG_POST_BATCHNORM) {
                sc = bn[o];
                b = b * sc + bn[Co + o];
            }
            eb[o] = b;
            eb[64 + o] = sc;
        }
        a.epiOff = (uint32_t)(g.wbytes - sizeof(eb));
        memcpy(reinterpret_cast<unsigned char *>(img.data()) + a.epiOff, eb, sizeof(eb));
    }
    return FYN_OK;
}

int upload_image(uint4 *&dptr, size_t &have, const std::vector<__half> &img) {
    const size_t bytes = img.size() * sizeof(__half);
    if (dptr && have < bytes) {
        cudaFree(dptr);
        dptr = nullptr;
    }
    if (!dptr) {
        FYN_CUDA(cudaMalloc((void **)&dptr, bytes));
        have = bytes;
    }
    FYN_CUDA(cudaMemcpy(dptr, img.data(), bytes, cudaMemcpyHostToDevice));
    return FYN_OK;
}

}  // namespace

int fyn_conv_tc_create(fyn_op *op, const float *wb) {
    const fyn_conv_desc &d = op->conv;
    // Two job rows stacked along N when the output height allows it and the plan fits (N <= 64, shared memory) without
    // giving up loader groups -- measured on StyleNet 9x9 @1524x1856: deconv3 40.1 -> 30.7 us, deconv1 10.9 -> 9.9 us, but
    // conv2, whose stacked plan only fits with two loader groups, 34.4 -> 44.8 us.
    Geometry g = plan_geometry(&d, wb, 1);
    const Geometry g1 = g;
    {
        Geometry g2 = plan_geometry(&d, wb, 2);
        if (const char *e = getenv("FYN_TC_STACK")) { if (atoi(e) < 2) g2.ok = false; }   // tuning knob
        if (g2.ok && op->Ho % g2.opy == 0 && (!g.ok || g2.finGroups >= g.finGroups)) g = g2;
    }
    if (!g.ok) FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 family does not cover this conv");
    ConvTcPlan *plan = op->tc ? op->tc : new ConvTcPlan();
    op->tc = plan;
    plan->mode = g.mode;
    std::vector<__half> img;
    if (int rc = build_plan(g, d, wb, plan->args, img)) return rc;
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    if (int rc = upload_image(plan->d_wimg, plan->wimgBytes, img)) return rc;
    plan->args.wimg = plan->d_wimg;
    plan->smemBytes = g.smem;
    if (plan->smemBytes > (size_t)op->ctx->prop.sharedMemPerBlockOptin)
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 conv needs %zu bytes of shared memory", plan->smemBytes);
    // The persistent chain kernel (fyn_conv_chain.cu) runs unstacked plans: where this layer runs stacked on its own, and
    // could be part of a chain (stride 1, as many outputs as inputs), the single-row plan is kept beside it -- and, for
    // the layers of a chain that sweep bottom -> top, the single-row image of the kernel with its rows reversed.
    plan->hasRow1 = false;
    const bool chainable = g1.ok && g1.mode == 0 && g1.opx == 1 && g1.nver == 1 && d.downsample == 1 && !d.fractional && d.in_channels == d.out_channels && d.kernel >= 3;
    if (chainable && g.opy > 1) {
        if (int rc = build_plan(g1, d, wb, plan->row1, img)) return rc;
        if (int rc = upload_image(plan->d_wimg1, plan->wimg1Bytes, img)) return rc;
        plan->row1.wimg = plan->d_wimg1;
        plan->hasRow1 = true;
    }
    if (chainable) {
        const int K = d.kernel, Ci = d.in_channels, Co = d.out_channels;
        const size_t nw = (size_t)Co + (size_t)K * K * Ci * Co + ((d.flags & FYN_FLAG_POST_BATCHNORM) ? 2 * (size_t)Co : 0);
        std::vector<float> flip(wb, wb + nw);
        for (int o = 0; o < Co; o++)
            for (int ky = 0; ky < K; ky++)
                memcpy(&flip[(size_t)Co + ((size_t)o * K + ky) * K * Ci], &wb[(size_t)Co + ((size_t)o * K + (K - 1 - ky)) * K * Ci], sizeof(float) * (size_t)K * Ci);
        const Geometry gf = plan_geometry(&d, flip.data(), 1);
        TcArgs scratch{};
        if (int rc = build_plan(gf, d, flip.data(), scratch, img)) return rc;
        if (int rc = upload_image(plan->d_wimgFlip, plan->wimgFlipBytes, img)) return rc;
    } else if (plan->d_wimgFlip) {
        cudaFree(plan->d_wimgFlip);
        plan->d_wimgFlip = nullptr;
        plan->wimgFlipBytes = 0;
    }
    return FYN_OK;
}

int fyn_conv_tc_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t stream) {
    ConvTcPlan *plan = op->tc;
    const fyn_conv_desc &d = op->conv;
    // tensor formats this family reads / writes
    if (out->desc.dtype != FYN_F16 || out->desc.order != FYN_ORDER_SHALLOW || (res && res->desc.dtype != FYN_F16)) return 1;
    if (plan->mode == 0 && (in->desc.dtype != FYN_F16 || in->geom.packing != 4)) return 1;
    if (in->desc.order != FYN_ORDER_SHALLOW && in->desc.channels > 4) return 1;
    // rows are staged with 16-byte-granular bulk copies that may run up to 15 bytes past the last texel of a row:
    // fine inside the tensor and for tensors this library allocated (16 bytes of slack); wrapped memory must end on a
    // 16-byte boundary
    if (!in->owns && (((uintptr_t)in->dptr + in->geom.bytes) & 15u) != 0) return 1;
    if (((uintptr_t)in->dptr & 7u) != 0) return 1;
    TcArgs a = plan->args;
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.res = fyn_make_view(res);
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    if (a.Wo % a.opx || a.Ho % a.opy) return 1;   // phases must tile the output exactly (else: direct kernel)
    a.Wj = a.Wo / a.opx;
    a.Hj = a.Ho / a.opy;
    a.inP = d.in_padding;
    a.outP = d.out_padding;
    a.resP = d.res_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    a.hasRes = (d.flags & FYN_FLAG_RESIDUAL_INPUT) != 0;
    a.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
    a.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
    a.epilogue = op->epilogue;
    if (const char *e = getenv("FYN_TC_DEBUG")) a.debug = atoi(e);
    a.batch = in->desc.batch;
    a.nxs = (a.Wj + kTileM - 1) / kTileM;
    // strip height: one strip per SM (the kernel's register / shared-memory footprint allows one CTA per SM),
    // at least 4 job rows so the prologue (weight image, TMEM allocation) amortises
    const int sms = op->ctx->prop.multiProcessorCount;
    long long cols = (long long)a.nxs * a.batch;
    int segs = (int)std::max<long long>(1, (long long)sms / std::max<long long>(1, cols));
    a.SH = std::max(4, (a.Hj + segs - 1) / segs);
    const int nseg = (a.Hj + a.SH - 1) / a.SH;
    const long long blocks = (long long)a.nxs * nseg * a.batch;
    // programmatic dependent launch: the prologue (barriers, TMEM, weight image) overlaps the previous kernel's tail
    const int actSel = a.act.type == 0 ? 0 : (a.act.type == 1 ? 1 : 2);
    int resSel = 0;
    if (a.hasRes) resSel = (a.mode == 0 && a.opx == 1 && a.opy == 1 && res->desc.dtype == FYN_F16 && res->geom.packing == 4 && res->desc.order == FYN_ORDER_SHALLOW) ? 1 : 2;
    if (op->epilogue != FYN_EPILOGUE_NONE && resSel != 0) return 1;   // fused function + residual: direct kernel
    TcKernel fn = tc_kernel(a.mode, actSel, resSel, op->epilogue);
    if (int rc = tc_ensure_smem(fn, a.mode, actSel, resSel, op->epilogue, op->ctx->device, plan->smemBytes)) return rc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan->smemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const bool noPdl = getenv("FYN_TC_NO_PDL") != nullptr;   // debugging aid: plain stream-ordered launches
    if (noPdl) cfg.numAttrs = 0;
    FYN_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}

void fyn_conv_tc_destroy(fyn_op *op) {
    if (!op->tc) return;
    if (op->tc->d_wimg) cudaFree(op->tc->d_wimg);
    if (op->tc->d_wimg1) cudaFree(op->tc->d_wimg1);
    if (op->tc->d_wimgFlip) cudaFree(op->tc->d_wimgFlip);
    delete op->tc;
    op->tc = nullptr;
}
