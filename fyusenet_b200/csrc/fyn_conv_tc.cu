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
//   16 bytes apart, so the canonical layout still holds.  Activation-at-fetch (fyusenet/base/layerbase.h:50-59,
//   shaders/activation.inc) and clamp-to-edge addressing (base/buffermanager.cpp:657-670) are applied by the
//   loader warps while they transpose 4-channel planes into 8-channel chunks, which is why the operands are
//   staged by threads (generic proxy + fence.proxy.async) instead of TMA.
//   For 3-channel inputs (StyleNet conv1 reading the RGB32F upload texture) a chunk is two horizontally
//   adjacent pixels x 4 channels ("pixel-pair" mode), i.e. one 16-byte chunk covers taps kx and kx+1.
//
// Warp roles (416 threads): warps 0-3 epilogue (TMEM -> registers -> bias/BN/residual -> fp16 planes),
// warps 4-11 loaders (two groups of four warps on alternating input rows, so two rows are always in flight),
// warp 12 issues tcgen05.mma (one elected lane).
// Pipelines: full/empty mbarriers per ring slot (loader <-> MMA), tmem_full/tmem_empty per accumulator
// buffer (MMA <-> epilogue, two buffers so the epilogue of row y overlaps the MMAs of row y+1).
#include <algorithm>
#include <cstring>
#include <vector>

#include "fyn_internal.h"

namespace {

constexpr int kMaxSteps = 48;
constexpr int kLoaderWarps = 8;                       // two groups of four
constexpr int kMmaWarp = 4 + kLoaderWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;       // 416
constexpr int kGroupThreads = kLoaderWarps * 16;    // threads per loader group (128)
constexpr int kUnroll = 6;                           // (pixel, chunk) items in flight per loader thread
constexpr int kTileM = 128;

struct TcStep {
    uint32_t a_off;   // byte offset of the first K-chunk inside its ring slot
    uint32_t a_lbo;   // byte distance to the second K-chunk
    int32_t row;      // ring row relative to the first row of the window (ky)
    uint32_t b_off;   // byte offset inside the weight image
};

struct TcArgs {
    TView in, out, res;
    const uint4 *wimg;
    const float *bias, *scale;
    uint32_t wbytes, idesc, b_lbo;
    int nsteps;
    TcStep steps[kMaxSteps];
    int K, ds, mh;           // kernel, stride, (K-1)/2
    int W, H, Wo, Ho;        // input / output net size
    int inP, outP, resP;
    int nchunks, rowpx;      // chunks per ring slot, pixels per chunk row
    int nslots, slotBytes;
    int SH, nxs;             // output rows per strip, column blocks
    int N, nOutPlanes, nInPlanes;
    int mode;                // 0 = plane-pair chunks (fp16 RGBA planes), 1 = pixel-pair (single plane)
    int x_lead;              // input pixels to the left of the first output pixel's centre held in a slot
    ActParams act;
    int hasRes, reluRes, bnRes;
    int batch;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16 x fp16 -> fp32)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 bytes stored
// contiguously (rows 16 B apart); SBO = distance between 8-row groups, LBO = distance between the two 16-byte
// K chunks of one K=16 instruction; bits [46,48) = 1 (sm_100 descriptor version).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// dynamic shared memory: [weight image][ring slots][barriers][tmem base]
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW = smem;
    unsigned char *sRing = smem + ((a.wbytes + 127) & ~127u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRing + (size_t)a.nslots * a.slotBytes);
    uint64_t *full = bars;                     // [nslots]
    uint64_t *empty = bars + a.nslots;         // [nslots]
    uint64_t *tfull = bars + 2 * a.nslots;     // [2]
    uint64_t *tempty = tfull + 2;              // [2]
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // strip decode: blockIdx.x -> (image, column block, row segment)
    int bid = blockIdx.x;
    const int xb = bid % a.nxs;
    bid /= a.nxs;
    const int nseg = (a.Ho + a.SH - 1) / a.SH;
    const int seg = bid % nseg;
    const int n = bid / nseg;
    const int ya = seg * a.SH, yb = min(a.Ho, ya + a.SH);
    const int x0 = xb * kTileM;
    const int r0 = a.ds * ya - a.mh;                  // first input row of the window (unclamped)
    const int r1 = a.ds * (yb - 1) + a.mh;            // last input row

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nslots; s++) {
            mbar_init(&full[s], kGroupThreads);   // every thread of the loading group arrives
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1);
        mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 128);
        mbar_init(&tempty[1], 128);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmemBase, 128);
    // weight image -> shared memory (all threads, 16-byte copies), visible to the async proxy
    for (uint32_t i = threadIdx.x; i < a.wbytes / 16; i += kThreads) reinterpret_cast<uint4 *>(sW)[i] = __ldg(a.wimg + i);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;

    if (warp >= 4 && warp < kMmaWarp) {
        // ===================== loaders: two groups on alternating rows =====================
        const int grp = (warp - 4) / (kLoaderWarps / 2);
        const int t = (threadIdx.x - 128) % kGroupThreads;   // thread index inside the group
        const int P = a.inP;
        const __half *src = reinterpret_cast<const __half *>(a.in.ptr);
        const __half2 hz = __float2half2_rn(0.f);
        for (int r = r0 + grp; r <= r1; r += 2) {
            const int idx = r - r0, slot = idx % a.nslots, fill = idx / a.nslots;
            mbar_wait(&empty[slot], (fill & 1) ^ 1);
            unsigned char *dst = sRing + (size_t)slot * a.slotBytes;
            const int iy = min(max(r + P, 0), a.in.texH - 1);   // texture row, CLAMP_TO_EDGE
            if (a.mode == 0) {
                // (pixel, chunk) items: two 8-byte plane loads -> one 16-byte chunk store.  All loads of a batch
                // are issued before the first store so that kUnroll*2 requests per thread are in flight.
                const int items = a.rowpx * a.nchunks, half = a.rowpx >> 1;
                const long long rowBase = (long long)n * a.in.imageElems + (long long)iy * a.in.texW * 4;
                for (int base = 0; base < items; base += kGroupThreads * kUnroll) {
                    uint2 lo[kUnroll], hi[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; u++) {
                        const int it = base + u * kGroupThreads + t;
                        lo[u] = make_uint2(0u, 0u);
                        hi[u] = make_uint2(0u, 0u);
                        if (it < items) {
                            const int c = it / a.rowpx, px = it - c * a.rowpx;
                            // stride 1: slot pixel = image pixel - (x0 - lead); stride 2: slot is [parity][pixel/2]
                            const int gx = (a.ds == 1) ? x0 - a.x_lead + px
                                                       : 2 * x0 - a.x_lead + 2 * (px % half) + (px / half);
                            const int ix = min(max(gx + P, 0), a.in.texW - 1);
                            const __half *q = src + rowBase + (long long)ix * 4;
                            if (2 * c < a.nInPlanes) lo[u] = __ldg(reinterpret_cast<const uint2 *>(q + (long long)(2 * c) * a.in.planeElems));
                            if (2 * c + 1 < a.nInPlanes) hi[u] = __ldg(reinterpret_cast<const uint2 *>(q + (long long)(2 * c + 1) * a.in.planeElems));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; u++) {
                        const int it = base + u * kGroupThreads + t;
                        if (it < items) {
                            if (a.act.type == 1) {
                                __half2 *q = reinterpret_cast<__half2 *>(&lo[u]);
                                q[0] = __hmax2(q[0], hz);
                                q[1] = __hmax2(q[1], hz);
                                q = reinterpret_cast<__half2 *>(&hi[u]);
                                q[0] = __hmax2(q[0], hz);
                                q[1] = __hmax2(q[1], hz);
                            } else if (a.act.type != 0) {
                                __half *q = reinterpret_cast<__half *>(&lo[u]);
                                for (int j = 0; j < 4; j++) q[j] = __float2half_rn(fyn_act(__half2float(q[j]), a.act));
                                q = reinterpret_cast<__half *>(&hi[u]);
                                for (int j = 0; j < 4; j++) q[j] = __float2half_rn(fyn_act(__half2float(q[j]), a.act));
                            }
                            *reinterpret_cast<uint4 *>(dst + (size_t)it * 16) = make_uint4(lo[u].x, lo[u].y, hi[u].x, hi[u].y);
                        }
                    }
                }
            } else {
                // pixel-pair mode: chunk(px) = [pixel px | pixel px+1], 4 channels each
                for (int px = t; px <= a.rowpx; px += kGroupThreads) {
                    const int ix = min(max(x0 - a.x_lead + px + P, 0), a.in.texW - 1);
                    float4 v = fyn_act4(fyn_load_texel(a.in, (long long)n * a.in.imageElems + ((long long)iy * a.in.texW + ix) * a.in.packing), a.act);
                    const uint2 h = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
                    if (px < a.rowpx) *reinterpret_cast<uint2 *>(dst + (size_t)px * 16) = h;
                    if (px > 0) *reinterpret_cast<uint2 *>(dst + (size_t)(px - 1) * 16 + 8) = h;
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[slot]);
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t wbase = smem_u32(sW), rbase = smem_u32(sRing);
            for (int y = ya; y < yb; y++) {
                const int i = y - ya, buf = i & 1, use = i >> 1;
                mbar_wait(&tempty[buf], (use & 1) ^ 1);
                // all rows of this output row's window must have landed
                const int first = a.ds * y - a.mh - r0;
                for (int k = 0; k < a.K; k++) {
                    const int idx = first + k;
                    mbar_wait(&full[idx % a.nslots], (idx / a.nslots) & 1);
                }
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)buf * 64u;
                for (int s = 0; s < a.nsteps; s++) {
                    const TcStep st = a.steps[s];
                    const int idx = first + st.row;
                    const uint32_t aaddr = rbase + (uint32_t)(idx % a.nslots) * (uint32_t)a.slotBytes + st.a_off;
                    umma_f16(d, make_desc(aaddr, st.a_lbo, 128), make_desc(wbase + st.b_off, a.b_lbo, 128), a.idesc, s > 0);
                }
                umma_commit(&tfull[buf]);
                // rows that no later output row needs go back to the loaders
                const int keepFrom = (y + 1 < yb) ? a.ds * (y + 1) - a.mh - r0 : (r1 - r0 + 1);
                for (int idx = first; idx < keepFrom; idx++) umma_commit(&empty[idx % a.nslots]);
            }
        }
    } else {
        // ===================== epilogue: warps 0-3, thread = output pixel =====================
        const int m = threadIdx.x;           // 0..127 == TMEM lane
        const int xo = x0 + m;
        const bool valid = xo < a.Wo;
        for (int y = ya; y < yb; y++) {
            const int i = y - ya, buf = i & 1, use = i >> 1;
            mbar_wait(&tfull[buf], use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * 64u;
            for (int g = 0; g < a.N / 16; g++) {
                uint32_t r[16];
                tmem_ld16(taddr + g * 16, r);
                tmem_ld_wait();
                if (g == a.N / 16 - 1) {
                    // accumulator fully read: hand the buffer back before doing the global-memory work
                    tc_fence_before();
                    mbar_arrive(&tempty[buf]);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int p = g * 4 + q;
                    if (p < a.nOutPlanes && valid) {
                        const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.scale) + p);
                        const float4 bi = __ldg(reinterpret_cast<const float4 *>(a.bias) + p);
                        float4 v = make_float4(fmaf(__uint_as_float(r[4 * q + 0]), sc.x, bi.x), fmaf(__uint_as_float(r[4 * q + 1]), sc.y, bi.y),
                                               fmaf(__uint_as_float(r[4 * q + 2]), sc.z, bi.z), fmaf(__uint_as_float(r[4 * q + 3]), sc.w, bi.w));
                        if (a.hasRes) {
                            float4 rs = fyn_fetch(a.res, n, p, a.resP + xo, a.resP + y);
                            if (a.reluRes) rs = make_float4(fmaxf(rs.x, 0.f), fmaxf(rs.y, 0.f), fmaxf(rs.z, 0.f), fmaxf(rs.w, 0.f));
                            if (a.bnRes) rs = make_float4(rs.x * sc.x, rs.y * sc.y, rs.z * sc.z, rs.w * sc.w);
                            v.x += rs.x;
                            v.y += rs.y;
                            v.z += rs.z;
                            v.w += rs.w;
                        }
                        fyn_store_texel(a.out, n, p, a.outP + xo, a.outP + y, v);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 128);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side: plan = MMA step table + weight image
// ---------------------------------------------------------------------------------------------
struct ConvTcPlan {
    TcArgs args{};
    uint4 *d_wimg = nullptr;
    float *d_bias = nullptr;  // [2 * Npad]: bias then scale
    size_t smemBytes = 0;
    int mode = 0;
};

static bool tc_shape_ok(const fyn_conv_desc *d, int *mode) {
    if (d->flags & FYN_FLAG_DEEP) return false;          // deep-tiled family: not yet
    if (d->fractional) return false;                     // fractional family: not yet
    if (d->dilation != 1) return false;
    if (d->out_channels > 64) return false;
    if (d->downsample != 1 && d->downsample != 2) return false;
    if (d->kernel < 3) return false;
    int m;
    if (d->in_channels <= 4 && d->downsample == 1) m = 1;
    else if (d->in_channels >= 8) m = 0;
    else return false;
    // MMA steps per output row and shared-memory footprint must fit
    const int K = d->kernel, mh = (K - 1) / 2, N = ((d->out_channels + 15) / 16) * 16;
    const int nchunks = m == 0 ? ((d->in_channels + 3) / 4 + 1) / 2 : 1;
    const int chunksPerRow = m == 0 ? K * nchunks : (K + 1) / 2;
    const int nsteps = ((chunksPerRow + 1) / 2) * K;
    if (nsteps > kMaxSteps) return false;
    int rowpx;
    if (m == 1) rowpx = ((kTileM + 2 * mh + 2) + 3) & ~3;
    else if (d->downsample == 1) rowpx = ((kTileM + 2 * mh + 1) + 3) & ~3;
    else rowpx = 2 * (((kTileM + mh + 1) + 3) & ~3);
    const size_t smem = (size_t)nsteps * 2 * N * 16 + (size_t)(K + 2 * d->downsample + 1) * nchunks * rowpx * 16 + 1024;
    if (smem > 200 * 1024) return false;
    *mode = m;
    return true;
}

int fyn_conv_tc_supported(const fyn_conv_desc *d, int) {
    int mode;
    return tc_shape_ok(d, &mode) ? 1 : 0;
}

int fyn_conv_tc_create(fyn_op *op, const float *wb) {
    const fyn_conv_desc &d = op->conv;
    int mode = 0;
    if (!tc_shape_ok(&d, &mode)) FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 family does not cover this conv");
    ConvTcPlan *plan = op->tc ? op->tc : new ConvTcPlan();
    op->tc = plan;
    plan->mode = mode;
    TcArgs &a = plan->args;
    const int K = d.kernel, Ci = d.in_channels, Co = d.out_channels, ds = d.downsample, mh = (K - 1) / 2;
    const int N = ((Co + 15) / 16) * 16;
    a.K = K;
    a.ds = ds;
    a.mh = mh;
    a.N = N;
    a.nOutPlanes = (Co + 3) / 4;
    a.nInPlanes = (Ci + 3) / 4;
    a.mode = mode;
    a.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);  // F32 accum, F16 x F16, K-major A/B
    a.b_lbo = (uint32_t)N * 16u;

    // ---- ring-slot geometry and the chunk list of one kernel row -------------------------------------
    // a "chunk" is 16 bytes of K for every pixel; chunkOff[] is its byte offset for output pixel m = 0
    struct Chunk { uint32_t off; int kx; int sub; };   // sub: chunk index inside the tap
    std::vector<Chunk> chunks;
    if (mode == 0) {
        a.nchunks = (a.nInPlanes + 1) / 2;
        a.x_lead = mh;
        if (ds == 1) {
            a.rowpx = ((kTileM + 2 * mh + 1) + 3) & ~3;
            for (int kx = 0; kx < K; kx++)
                for (int c = 0; c < a.nchunks; c++) chunks.push_back({(uint32_t)((c * a.rowpx + kx) * 16), kx, c});
        } else {
            // stride 2: [parity][pixel/2]; slot pixel j = 2m + kx  ->  parity kx&1, index m + kx/2
            const int half = ((kTileM + mh + 1) + 3) & ~3;
            a.rowpx = 2 * half;
            for (int kx = 0; kx < K; kx++)
                for (int c = 0; c < a.nchunks; c++)
                    chunks.push_back({(uint32_t)((c * a.rowpx + (kx & 1) * half + kx / 2) * 16), kx, c});
        }
    } else {
        a.nchunks = 1;
        a.x_lead = mh;
        a.rowpx = ((kTileM + 2 * mh + 2) + 3) & ~3;
        for (int kx = 0; kx < K; kx += 2) chunks.push_back({(uint32_t)(kx * 16), kx, 0});
    }
    a.slotBytes = a.nchunks * a.rowpx * 16;
    a.nslots = K + 2 * ds + 1;

    // ---- pair chunks into K=16 steps (second chunk must lie at a higher address) ----------------------
    // greedy: sort by offset, pair neighbours; an unpaired chunk is paired with itself against zero weights
    std::vector<int> order(chunks.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return chunks[x].off < chunks[y].off; });
    struct Pair { int c0, c1; };
    std::vector<Pair> pairs;
    for (size_t i = 0; i < order.size(); i += 2) {
        if (i + 1 < order.size()) pairs.push_back({order[i], order[i + 1]});
        else pairs.push_back({order[i], -1});
    }
    const int stepsPerRow = (int)pairs.size();
    a.nsteps = stepsPerRow * K;
    if (a.nsteps > kMaxSteps) FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 conv: %d MMA steps exceed the table (%d)", a.nsteps, kMaxSteps);

    // ---- weight image: per step two chunks of [N][8] fp16 ----------------------------------------------
    const size_t wbytes = (size_t)a.nsteps * 2 * N * 16;
    std::vector<__half> img(wbytes / 2, __float2half(0.f));
    const float *W = wb + Co;  // [Co][K][K][Ci]
    auto fill_chunk = [&](size_t chunkIdx, int ky, const Chunk &c) {
        for (int nn = 0; nn < Co; nn++)
            for (int e = 0; e < 8; e++) {
                float v = 0.f;
                if (mode == 0) {
                    const int ci = c.sub * 8 + e;
                    if (ci < Ci) v = W[(((size_t)nn * K + ky) * K + c.kx) * Ci + ci];
                } else {
                    const int kx = c.kx + e / 4, ci = e % 4;
                    if (kx < K && ci < Ci) v = W[(((size_t)nn * K + ky) * K + kx) * Ci + ci];
                }
                img[(chunkIdx * N + nn) * 8 + e] = __float2half_rn(v);
            }
    };
    int s = 0;
    for (int ky = 0; ky < K; ky++)
        for (const Pair &p : pairs) {
            TcStep &st = a.steps[s];
            st.row = ky;
            st.a_off = chunks[p.c0].off;
            st.b_off = (uint32_t)((size_t)s * 2 * N * 16);
            fill_chunk((size_t)s * 2, ky, chunks[p.c0]);
            if (p.c1 >= 0) {
                st.a_lbo = chunks[p.c1].off - chunks[p.c0].off;
                fill_chunk((size_t)s * 2 + 1, ky, chunks[p.c1]);
            } else {
                st.a_lbo = 16;  // second half reads the neighbouring pixel against all-zero weights
            }
            s++;
        }
    a.wbytes = (uint32_t)wbytes;

    FYN_CUDA(cudaSetDevice(op->ctx->device));
    if (!plan->d_wimg) FYN_CUDA(cudaMalloc((void **)&plan->d_wimg, wbytes));
    FYN_CUDA(cudaMemcpy(plan->d_wimg, img.data(), wbytes, cudaMemcpyHostToDevice));
    // epilogue parameters padded to N
    std::vector<float> eb((size_t)2 * N, 0.f);
    const float *bn = wb + Co + (size_t)K * K * Ci * Co;
    for (int o = 0; o < Co; o++) {
        float b = wb[o], sc = 1.f;
        if (d.flags & FYN_FLAG_POST_BATCHNORM) {
            sc = bn[o];
            b = b * sc + bn[Co + o];
        }
        eb[o] = b;
        eb[N + o] = sc;
    }
    if (!plan->d_bias) FYN_CUDA(cudaMalloc((void **)&plan->d_bias, eb.size() * sizeof(float)));
    FYN_CUDA(cudaMemcpy(plan->d_bias, eb.data(), eb.size() * sizeof(float), cudaMemcpyHostToDevice));
    a.wimg = plan->d_wimg;
    a.bias = plan->d_bias;
    a.scale = plan->d_bias + N;
    plan->smemBytes = ((wbytes + 127) & ~(size_t)127) + (size_t)a.nslots * a.slotBytes + (2 * a.nslots + 4) * 8 + 16;
    if (plan->smemBytes > (size_t)op->ctx->prop.sharedMemPerBlockOptin)
        FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 conv needs %zu bytes of shared memory", plan->smemBytes);
    // the attribute is per function, not per launch: keep it at the largest footprint any plan needs
    static size_t maxSmem[64] = {0};
    size_t &cur = maxSmem[op->ctx->device & 63];
    if (plan->smemBytes > cur) {
        FYN_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smemBytes));
        cur = plan->smemBytes;
    }
    return FYN_OK;
}

int fyn_conv_tc_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t stream) {
    ConvTcPlan *plan = op->tc;
    const fyn_conv_desc &d = op->conv;
    // tensor formats this family reads / writes
    if (out->desc.dtype != FYN_F16 || (res && res->desc.dtype != FYN_F16)) return 1;
    if (plan->mode == 0 && (in->desc.dtype != FYN_F16 || in->geom.packing != 4)) return 1;
    if (in->desc.order != FYN_ORDER_SHALLOW && in->desc.channels > 4) return 1;
    TcArgs a = plan->args;
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.res = fyn_make_view(res);
    a.W = d.width;
    a.H = d.height;
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.inP = d.in_padding;
    a.outP = d.out_padding;
    a.resP = d.res_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    a.hasRes = (d.flags & FYN_FLAG_RESIDUAL_INPUT) != 0;
    a.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
    a.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
    a.batch = in->desc.batch;
    a.nxs = (a.Wo + kTileM - 1) / kTileM;
    // strip height: one strip per SM (the kernel's register / shared-memory footprint allows one CTA per SM),
    // at least 4 rows so the prologue (weight image, TMEM allocation) amortises
    const int sms = op->ctx->prop.multiProcessorCount;
    long long target = 1LL * sms;
    long long cols = (long long)a.nxs * a.batch;
    int segs = (int)std::max<long long>(1, target / std::max<long long>(1, cols));
    a.SH = std::max(4, (a.Ho + segs - 1) / segs);
    const int nseg = (a.Ho + a.SH - 1) / a.SH;
    const long long blocks = (long long)a.nxs * nseg * a.batch;
    k_conv_tc<<<(unsigned)blocks, kThreads, plan->smemBytes, stream>>>(a);
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}

void fyn_conv_tc_destroy(fyn_op *op) {
    if (!op->tc) return;
    if (op->tc->d_wimg) cudaFree(op->tc->d_wimg);
    if (op->tc->d_bias) cudaFree(op->tc->d_bias);
    delete op->tc;
    op->tc = nullptr;
}
