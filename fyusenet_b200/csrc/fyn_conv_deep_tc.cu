// fyn_conv_deep_tc.cu -- tcgen05 / TMEM implicit GEMM for the DEEP-tiled convolution layers (ResNet-50).
//
// Replaces the instanced blend passes of DeepConvLayer1x1 / DeepConvLayerNxN / DeepGEMMLayer
// (fyusenet/gpu/deep/deepconvlayer1x1.cpp:67-123, deepconvlayerNxN.cpp:81-211, deepgemmlayer.cpp:66-236: one pass per
// input tile and kernel row, each touching every output tile) by one GEMM per layer:
//   D[m][n] = sum_{tap, ci} act(in[pixel(m) + tap][ci]) * W[n][tap][ci]      m = flattened output pixel (image, y, x)
// Deep images are small (7x7 ... 56x56) and wide (64 ... 2048 channels), so unlike the row-ring kernel of the shallow
// family (fyn_conv_tc.cu) the M dimension runs over flattened pixels of the whole batch and K is walked in stages of
// one kernel tap x 64 input channels:
//   * A stage (128 pixels x 64 channels, 16 KB): loader thread m gathers its pixel's 16 plane texels (8 bytes each; the
//     tiles of the deep layout, fyusenet/gpu/deep/deeptiler.cpp:63-95), applies the activation at fetch
//     (shaders/activation.inc) and writes eight 16-byte chunks in the UMMA K-major SWIZZLE_NONE layout
//     [8-channel chunk][pixel][8 x fp16].  The tile padding of the deep layout supplies the zero border
//     (gpu/gpulayerbase.h:84-90), so 3x3 taps simply read the neighbouring texels.
//   * B stage (N x 64 weights, fp16 TRUNCATED like the reference's RGBA32UI weight texture,
//     deepconvlayerbase.cpp:338-349): pre-packed per (N tile, stage) and fetched with one bulk copy (TMA).
//   * four tcgen05.mma (M = 128, N <= 128, K = 16) per stage accumulate in TMEM; after the last stage the loader
//     warps turn into the epilogue: tcgen05.ld -> *bnScale + bias (fp16-rounded like the reference's RGBA16F bias
//     texture, deepconvlayerbase.cpp:371-394) (+ residual [ReLU] [*bnScale]) -> fp16 texels of the output tiles.
// Four kernels share this arithmetic (the launcher picks by grid size, fyn_conv_deep_tc_run): k_conv_deep_tc_sk (small grids:
// K split over a thread-block cluster, reduce-scatter through distributed shared memory), k_conv_deep_tc_p (persistent, one
// CTA per SM: every other grid), k_conv_deep_tc_h3 (3x3 stride 1: halo tiles) and k_conv_deep_tc, the one-tile-per-CTA
// kernel described here, which the others are tested against (FYN_DEEP_PERSIST=0 selects it).
// One CTA = one (128-pixel, N-tile) output tile; 4 x 4 loader/epilogue warps (stages round robin, so that four gathers
// are in flight) + 1 MMA warp; stage ring of min(4, stages)
// entries, so that the many short-K layers (1x1 convs on 64 channels: a single stage) fit several CTAs per SM.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <mutex>

#include "fyn_internal.h"

namespace {

constexpr int kM = 128;          // pixels per CTA
constexpr int kKC = 64;          // input channels per stage (16 planes)
constexpr int kMaxRing = 4;      // stage ring depth (fewer for layers with fewer stages: more CTAs fit an SM)
constexpr int kLoadSets = 4;     // at most this many sets of four loader warps; they take the stages round robin: that many gathers are in flight per CTA
constexpr int kLoadWarps = 4 * kLoadSets;
constexpr int kThreadsDeep = (kLoadWarps + 1) * 32;
constexpr int kAStageBytes = kM * kKC * 2;   // 16 KB

struct DeepTcArgs {
    TView in, out, res;
    const uint4 *wimg;           // [ntile][stage][8 chunks][NT][8 halfs]
    const float *bias, *scale;   // per output channel (padded to planes), the fp16-rounded parameter set
    const float4 *inNorm;        // fused input batch-norm: scale per input plane, then bias per input plane; or NULL
    int K, ds, mh, Wo, Ho, batch;
    int nInPlanes, Cout4;        // input planes; output channels rounded up to a multiple of 4
    int NT, nstages, kcs;        // columns per N tile, stages = K*K*kcs, kcs = Cin / 64
    int ring;                    // stage ring depth <= min(kMaxRing, nstages)
    int nsets;                   // loader sets of this layer (<= kLoadSets): blockDim = (4 * nsets + 1) warps
    int tapPacked;               // 1: single input plane (<= 4 channels): a stage holds 16 kernel taps x 4 channels (ResNet stem)
    int lastQuads;               // tap-packed: K16 steps (groups of four taps) of the LAST stage that hold kernel taps (7x7: one of four)
    long long Mtotal;            // batch * Ho * Wo
    uint32_t idesc;
    ActParams act;
    int hasRes, reluRes, bnRes;
    int inP, outP, resP;
    // persistent kernel: tiles (m tile, n tile) are numbered n-fastest and taken round robin by the CTAs
    int ntilesN, epiWarps;
    // halo-tile 3x3 kernel: virtual raster width (W + 1), halo pixels per stage (multiple of 8), weight ring depth
    int Wv, HPp, ringB;
    long long totalTiles;
    // split-K cluster kernel (small grids): columns per CTA (a sub-tile of the NT-column operand image), cluster size along K,
    // output planes finished per CTA of a cluster
    int skNT, skSplit, skPpr;
    int skPprLog2, skRingLog2, skSubLog2;   // skPpr, ring and NT / skNT are powers of two in that kernel (shifts instead of divisions)
    int skPer, skRem;                        // K stages per CTA: skPer, the first skRem ranks one more
    float invTxIn, invHW, invWo;             // reciprocals for the exact small-integer divisions of the set-up
    int dbgIndex;                // FYN_SK_TIMELINE builds: launch number (slot of the timeline buffer)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "DWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DWAIT_DONE;\n\t"
        "bra DWAIT_LOOP;\n\t"
        "DWAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint2 act_h4(uint2 v, const ActParams &a) {
    if (a.type == 0) return v;
    if (a.type == 1) {
        const __half2 z = __float2half2_rn(0.f);
        __half2 *q = reinterpret_cast<__half2 *>(&v);
        q[0] = __hmax2(q[0], z);
        q[1] = __hmax2(q[1], z);
        return v;
    }
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
    return make_uint2(pack_half2(fyn_act(f0.x, a), fyn_act(f0.y, a)), pack_half2(fyn_act(f1.x, a), fyn_act(f1.y, a)));
}

template <int ACT>
__device__ __forceinline__ uint2 act_h4_t(uint2 v, const ActParams &a) {
    if (ACT == 0) return v;
    if (ACT == 1) {
        const __half2 z = __float2half2_rn(0.f);
        __half2 *q = reinterpret_cast<__half2 *>(&v);
        q[0] = __hmax2(q[0], z);
        q[1] = __hmax2(q[1], z);
        return v;
    }
    return act_h4(v, a);
}

// dynamic shared memory: [A stages][B stages][plane origin tables][barriers][tmem base]
// NORM: the input batch-norm fusion (its own instantiation, so that the plain kernel keeps its register budget); two CTAs per SM
template <bool NORM>
__global__ void __launch_bounds__(kThreadsDeep, 2) k_conv_deep_tc(const __grid_constant__ DeepTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int bStageBytes = a.NT * kKC * 2;
    unsigned char *sA = smem;
    unsigned char *sB = sA + a.ring * kAStageBytes;
    int *inOrigin = reinterpret_cast<int *>(sB + a.ring * bStageBytes);   // [nInPlanes] element offset of a tile's origin
    int *outOrigin = inOrigin + a.nInPlanes;                               // [NT/4] (output tensor), then [NT/4] (residual tensor)
    int *resOrigin = outOrigin + (a.NT >> 2);
    int *tapTab = resOrigin + (a.NT >> 2);                                 // [64] (ky | kx << 8) of tap-packed stages
    uint64_t *full = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(tapTab + 64) + 7) & ~uintptr_t(7));
    uint64_t *empty = full + kMaxRing;
    uint64_t *done = empty + kMaxRing;
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(done + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int loadWarps = 4 * a.nsets, nthreads = (loadWarps + 1) * 32;
    const int ntile = blockIdx.y;
    const long long m0 = (long long)blockIdx.x * kM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.ring; s++) {
            mbar_init(&full[s], kM + 1);   // the 128 threads of the set that loads the stage + the expect_tx arrival of its weight copy
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    const uint32_t tmemCols = a.NT > 128 ? 256u : 128u;
    if (warp == loadWarps) tmem_alloc(tmemBase, tmemCols);
    // tile origins: plane q sits at tile (q % tx, q / tx), tiles are tileW x tileH texels apart (deeptiler.cpp:91-94)
    for (int q = threadIdx.x; q < a.nInPlanes && !a.tapPacked; q += nthreads) inOrigin[q] = ((q / a.in.tx) * a.in.tileH * a.in.texW + (q % a.in.tx) * a.in.tileW) * 4;
    for (int k = threadIdx.x; k < (a.NT >> 2); k += nthreads) {
        const int p = ntile * (a.NT >> 2) + k;
        outOrigin[k] = ((p / a.out.tx) * a.out.tileH * a.out.texW + (p % a.out.tx) * a.out.tileW) * 4;
        resOrigin[k] = a.hasRes ? ((p / a.res.tx) * a.res.tileH * a.res.texW + (p % a.res.tx) * a.res.tileW) * 4 : 0;
    }
    if (a.tapPacked && threadIdx.x < 64) {
        const int tp = min((int)threadIdx.x, a.K * a.K - 1);                  // taps beyond K*K carry zero weights
        tapTab[threadIdx.x] = (tp / a.K) | ((tp % a.K) << 8);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    // programmatic dependent launch: the set-up above overlapped the tail of the previous layer's kernel; its results (this
    // layer's input / residual) are only touched behind the wait (fetching the first weights ahead of it gained nothing)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < loadWarps) {
        // ===================== loaders (then epilogue): thread = GEMM row = output pixel =====================
        const int t = threadIdx.x & (kM - 1), set = warp >> 2;
        const long long m = m0 + t;
        const bool valid = m < a.Mtotal;
        const int hw = a.Ho * a.Wo;
        const int n = valid ? (int)(m / hw) : 0;
        const int rem = valid ? (int)(m - (long long)n * hw) : 0;
        const int yo = rem / a.Wo, xo = rem - yo * a.Wo;
        const __half *src = reinterpret_cast<const __half *>(a.in.ptr) + (long long)n * a.in.imageElems;
        const uint4 *wsrc = a.wimg + (size_t)ntile * a.nstages * (bStageBytes >> 4);
        for (int s = set; s < a.nstages; s += a.nsets) {
            const int st = s % a.ring, use = s / a.ring;
            const int tap = s / a.kcs, kc = s - tap * a.kcs;
            const int ky = tap / a.K, kx = tap - ky * a.K;
            mbar_wait(&empty[st], (use & 1) ^ 1);
            if (t == 0) {
                mbar_expect_tx(&full[st], (uint32_t)bStageBytes);
                bulk_g2s(sB + (size_t)st * bStageBytes, wsrc + (size_t)s * (bStageBytes >> 4), (uint32_t)bStageBytes, &full[st]);
            }
            uint2 v[kKC / 4];
            if (!a.tapPacked) {
                // texel of this pixel and tap, in tile-local coordinates (origin = the tile's top-left padding texel)
                const int iy = a.inP + a.ds * yo + ky - a.mh, ix = a.inP + a.ds * xo + kx - a.mh;
                const __half *px = src + (iy * a.in.texW + ix) * 4;
                const int *org = inOrigin + kc * (kKC / 4);
#pragma unroll
                for (int j = 0; j < kKC / 4; j++) v[j] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + org[j])) : make_uint2(0u, 0u);
                if (NORM) {
                    // the stand-alone batch-norm layer in front of this convolution, evaluated at the fetch: x * s + b in
                    // fp32, rounded to fp16 like that layer's store (deepbatchnorm.frag:57-58)
                    const float4 *sc = a.inNorm + kc * (kKC / 4), *bi = sc + a.nInPlanes;
#pragma unroll
                    for (int j = 0; j < kKC / 4; j++) {
                        const float4 s4 = __ldg(sc + j), b4 = __ldg(bi + j);
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].x));
                        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].y));
                        v[j] = make_uint2(pack_half2(fmaf(f0.x, s4.x, b4.x), fmaf(f0.y, s4.y, b4.y)), pack_half2(fmaf(f1.x, s4.z, b4.z), fmaf(f1.y, s4.w, b4.w)));
                    }
                }
            } else {
                // one plane, sixteen taps per stage; the texture clamps at its edge (base/buffermanager.cpp:657-670), which is
                // how the under-padded 7x7 stem (P = 1 < 3) still sees a zero border: the outermost texels are padding
                // (the last stage of a 7x7 kernel holds one tap: only its first K16 step is gathered and multiplied)
                const int ntaps = (s == a.nstages - 1) ? 4 * a.lastQuads : kKC / 4;
#pragma unroll
                for (int j = 0; j < kKC / 4; j++) {
                    v[j] = make_uint2(0u, 0u);
                    if (j >= ntaps) continue;
                    const int tt = tapTab[s * (kKC / 4) + j];
                    const int tky = tt & 255, tkx = tt >> 8;
                    const int iy = min(max(a.inP + a.ds * yo + tky - a.mh, 0), a.in.texH - 1);
                    const int ix = min(max(a.inP + a.ds * xo + tkx - a.mh, 0), a.in.texW - 1);
                    if (valid) v[j] = __ldg(reinterpret_cast<const uint2 *>(src + (iy * a.in.texW + ix) * 4));
                }
            }
            unsigned char *dst = sA + (size_t)st * kAStageBytes + t * 16;
#pragma unroll
            for (int c = 0; c < kKC / 8; c++) {
                const uint2 lo = act_h4(v[2 * c], a.act), hi = act_h4(v[2 * c + 1], a.act);
                *reinterpret_cast<uint4 *>(dst + c * (kM * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
            fence_proxy_async();
            mbar_arrive(&full[st]);
        }
        // ===================== epilogue =====================
        __half *outp = reinterpret_cast<__half *>(a.out.ptr) + (long long)n * a.out.imageElems + ((a.outP + yo) * a.out.texW + a.outP + xo) * 4;
        const __half *resp = reinterpret_cast<const __half *>(a.res.ptr) + (long long)n * a.res.imageElems + ((a.resP + yo) * a.res.texW + a.resP + xo) * 4;
        // the residual texels of this thread's first column groups are fetched while the last MMAs are still running
        constexpr int kPre = 2;
        uint2 rpre[kPre][4];
#pragma unroll
        for (int g = 0; g < kPre; g++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int cg = set + g * a.nsets;
                rpre[g][k] = (a.hasRes && valid && cg < (a.NT >> 4) && ntile * a.NT + cg * 16 + 4 * k < a.Cout4)
                                 ? __ldg(reinterpret_cast<const uint2 *>(resp + resOrigin[cg * 4 + k])) : make_uint2(0u, 0u);
            }
        mbar_wait(done, 0);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM lane quarter of this warp
        int g = 0;
        for (int cg = set; cg < (a.NT >> 4); cg += a.nsets, g++) {        // the sets split the column groups
            uint32_t acc[16];
            tmem_ld16(taddr + cg * 16, acc);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c0 = ntile * a.NT + cg * 16 + 4 * k;      // first channel of this texel
                if (!valid || c0 >= a.Cout4) continue;
                const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.scale + c0)), bi = __ldg(reinterpret_cast<const float4 *>(a.bias + c0));
                float4 r = make_float4(fmaf(__uint_as_float(acc[4 * k + 0]), sc.x, bi.x), fmaf(__uint_as_float(acc[4 * k + 1]), sc.y, bi.y),
                                       fmaf(__uint_as_float(acc[4 * k + 2]), sc.z, bi.z), fmaf(__uint_as_float(acc[4 * k + 3]), sc.w, bi.w));
                const int pk = cg * 4 + k;                             // plane inside this N tile
                if (a.hasRes) {
                    const uint2 raw = g == 0 ? rpre[0][k] : (g == 1 ? rpre[1][k] : __ldg(reinterpret_cast<const uint2 *>(resp + resOrigin[pk])));
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                    float4 q = make_float4(f0.x, f0.y, f1.x, f1.y);
                    if (a.reluRes) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
                    if (a.bnRes) q = make_float4(q.x * sc.x, q.y * sc.y, q.z * sc.z, q.w * sc.w);
                    r.x += q.x;
                    r.y += q.y;
                    r.z += q.z;
                    r.w += q.w;
                }
                *reinterpret_cast<uint2 *>(outp + outOrigin[pk]) = make_uint2(pack_half2(r.x, r.y), pack_half2(r.z, r.w));
            }
        }
    } else {
        // ===================== MMA issuer =====================
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
        const uint32_t aLbo = ((uint32_t)(kM * 16) >> 4) << 16, bLbo = ((uint32_t)(a.NT * 16) >> 4) << 16;
        for (int s = 0; s < a.nstages; s++) {
            const int st = s % a.ring, use = s / a.ring;
            if (elect_one()) {
                mbar_wait(&full[st], use & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + (size_t)st * kAStageBytes) >> 4, b0 = smem_u32(sB + (size_t)st * bStageBytes) >> 4;
                const int nq = (a.tapPacked && s == a.nstages - 1) ? a.lastQuads : kKC / 16;
#pragma unroll
                for (int j = 0; j < kKC / 16; j++)
                    if (j < nq)
                        umma_f16(tmem, hi | (uint64_t)(aLbo | (a0 + (uint32_t)(j * 2 * kM))), hi | (uint64_t)(bLbo | (b0 + (uint32_t)(j * 2 * a.NT))), a.idesc,
                                 (s > 0 || j > 0) ? 1u : 0u);
                umma_commit(&empty[st]);
                if (s == a.nstages - 1) umma_commit(done);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == loadWarps) tmem_dealloc(tmem, tmemCols);
}

// ---------------------------------------------------------------------------------------------
// Split-K cluster kernel for SMALL grids (batch 1: a layer is 1 ... 50 output tiles on 148 SMs, and the time of a layer is the
// length of one CTA's chain of K stages -- 2048 -> 512 @7x7 took 42 us as 4 CTAs of 32 stages).  Here the K stages of one output
// tile are split over the CTAs of a thread-block CLUSTER (<= 8, gridDim.z) and the tile is narrowed to skNT <= 64 columns (a
// sub-tile of the NT-column operand image, fetched as eight 16 * skNT-byte bulk copies per stage), so that a layer spreads over
// ~100 CTAs of <= 5 stages, all of them in flight at once.  Every CTA accumulates its K slice in TMEM; the partial tiles
// are then REDUCE-SCATTERED through distributed shared memory: CTA r owns skPpr output planes, every CTA stores its fp32 partials
// of those planes into r's buffer (st.shared::cluster, one float4 = one texel per thread, rows contiguous), a cluster barrier,
// and r sums the partials in rank order (deterministic) and runs the usual epilogue on its planes.  Same operands as the other
// kernels, another fp32 summation order (K slices summed separately): same tolerance against the oracle, not the same bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f4(uint32_t localAddr, uint32_t rank, float4 v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(localAddr), "r"(rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifdef FYN_SK_TIMELINE
// debug build: per launch [min start, max end over the CTAs, then the stamps of CTA (0,0,0)'s thread 0]
__device__ unsigned long long g_skTimeline[2048][16];
#endif

// ACT: activation at the fetch as a compile-time constant (0 none, 1 ReLU, 2 whatever args.act says).  Every instruction of this
// kernel runs once or twice per launch, so its time is largely INSTRUCTION FETCH of cold code: the generic activation inlined
// sixteen times made the conversion of one gathered stage take 1.2 us (in-kernel globaltimer stamps); keep the paths short.
template <bool NORM, int ACT>
__global__ void __launch_bounds__(kThreadsDeep, 1) k_conv_deep_tc_sk(const __grid_constant__ DeepTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int NTs = a.skNT, ks = a.skSplit, ppr = a.skPpr;
    const int bStageBytes = NTs * kKC * 2;
    const int nlocMax = a.skPer + (a.skRem ? 1 : 0);
    unsigned char *sA = smem;
    unsigned char *sB = sA + a.ring * kAStageBytes;
    float4 *red = reinterpret_cast<float4 *>(sB + a.ring * bStageBytes);   // [source rank][plane of this CTA][GEMM row]
    float4 *sScale = red + (size_t)ks * ppr * kM;                          // [ppr] scale, [ppr] bias of this CTA's planes
    float4 *sBias = sScale + ppr;
    float4 *sNorm = sBias + ppr;                                           // NORM: [local stage][16 planes] scale, then bias
    int *inOrigin = reinterpret_cast<int *>(sNorm + (NORM ? 2 * nlocMax * (kKC / 4) : 0));   // [nInPlanes]
    int *outOrigin = inOrigin + a.nInPlanes;                               // [ppr] output tensor, then [ppr] residual tensor
    int *resOrigin = outOrigin + ppr;
    int *stageTab = resOrigin + ppr;                                       // [nlocMax] tap offset (elements) | kc << 24 of this CTA's stages
    uint64_t *full = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(stageTab + nlocMax) + 7) & ~uintptr_t(7));
    uint64_t *empty = full + kMaxRing;
    uint64_t *done = empty + kMaxRing;
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(done + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int loadWarps = 4 * a.nsets, nthreads = (loadWarps + 1) * 32;
    const int rank = ks > 1 ? (int)cluster_ctarank() : 0;                  // cluster = (1, 1, ks): the CTAs of one output tile
#ifdef FYN_SK_TIMELINE
    unsigned long long ts[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define SK_STAMP(i) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[i]))
#else
#define SK_STAMP(i)
#endif
    SK_STAMP(0);
    const int sub = blockIdx.y;                                            // sub-tile of skNT columns
    const int s0 = rank * a.skPer + min(rank, a.skRem), nloc = a.skPer + (rank < a.skRem ? 1 : 0);   // this CTA's K stages
    const int plane0 = sub * (NTs >> 2) + rank * ppr;                      // first output plane this CTA finishes
    const int nOutPlanes = a.Cout4 >> 2;
    // the sub-tile's columns inside the NT-column operand image: [n tile][stage][chunk][NT][8 halfs]
    const int ntile = sub >> a.skSubLog2, colIn = (sub - (ntile << a.skSubLog2)) * NTs;
    const uint4 *wsrc = a.wimg + (size_t)ntile * a.nstages * (a.NT * kKC * 2 >> 4) + colIn;

    // ---- everything that does not depend on the previous layer's output happens ahead of griddepcontrol.wait ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.ring; s++) {
            mbar_init(&full[s], kM + 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    const uint32_t tmemCols = NTs > 64 ? 128u : (NTs > 32 ? 64u : 32u);
    if (warp == loadWarps) tmem_alloc(tmemBase, tmemCols);
    // (q / tx for q < 2^20 through the reciprocal: exact after one correction step)
#pragma unroll 1
    for (int q = threadIdx.x; q < a.nInPlanes; q += nthreads) {
        int ty = (int)(((float)q + 0.5f) * a.invTxIn);
        ty -= (ty * a.in.tx > q) ? 1 : 0;
        inOrigin[q] = (ty * a.in.tileH * a.in.texW + (q - ty * a.in.tx) * a.in.tileW) * 4;
    }
#pragma unroll 1
    for (int i = threadIdx.x; i < nloc; i += nthreads) {
        const int s = s0 + i, tap = s / a.kcs, kc = s - tap * a.kcs, ky = tap / a.K, kx = tap - ky * a.K;
        stageTab[i] = ((ky * a.in.texW + kx) * 4) | (kc << 24);
    }
#pragma unroll 1
    for (int k = threadIdx.x; k < ppr; k += nthreads) {
        const int p = plane0 + k;
        const bool ok = rank * ppr + k < (NTs >> 2) && p < nOutPlanes;
        outOrigin[k] = ((p / a.out.tx) * a.out.tileH * a.out.texW + (p % a.out.tx) * a.out.tileW) * 4;
        resOrigin[k] = a.hasRes ? ((p / a.res.tx) * a.res.tileH * a.res.texW + (p % a.res.tx) * a.res.tileW) * 4 : 0;
        sScale[k] = ok ? __ldg(reinterpret_cast<const float4 *>(a.scale) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
        sBias[k] = ok ? __ldg(reinterpret_cast<const float4 *>(a.bias) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (NORM) {
        // the fused input batch-norm's parameters of this CTA's stages (1x1 layers: stage = 64-channel block)
        for (int q = threadIdx.x; q < nloc * (kKC / 4); q += nthreads) {
            sNorm[q] = __ldg(a.inNorm + s0 * (kKC / 4) + q);
            sNorm[nlocMax * (kKC / 4) + q] = __ldg(a.inNorm + a.nInPlanes + s0 * (kKC / 4) + q);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    // "this CTA runs": nobody stores into a peer's shared memory before every CTA of the cluster has arrived here
    if (ks > 1) cluster_arrive_relaxed();
    SK_STAMP(1);

    if (warp < loadWarps) {
        // ===================== loaders: thread = GEMM row = output pixel =====================
        const int t = threadIdx.x & (kM - 1), set = warp >> 2;
        const unsigned m = blockIdx.x * (unsigned)kM + t;                   // (the launcher keeps Mtotal below 2^31)
        const bool valid = m < (unsigned)a.Mtotal;
        const unsigned hw = (unsigned)(a.Ho * a.Wo);
        unsigned n = valid ? (unsigned)(((float)m + 0.5f) * a.invHW) : 0u;      // m < 2^20 here: exact after the correction
        n -= (n * hw > m) ? 1u : 0u;
        const unsigned rem = valid ? m - n * hw : 0u;
        int yo = (int)(((float)rem + 0.5f) * a.invWo);
        yo -= (yo * a.Wo > (int)rem) ? 1 : 0;
        const int xo = (int)rem - yo * a.Wo;
        const __half *src = reinterpret_cast<const __half *>(a.in.ptr) + (long long)n * a.in.imageElems +
                            ((a.inP + a.ds * yo - a.mh) * a.in.texW + a.inP + a.ds * xo - a.mh) * 4;
        __half *outp = reinterpret_cast<__half *>(a.out.ptr) + (long long)n * a.out.imageElems + ((a.outP + yo) * a.out.texW + a.outP + xo) * 4;
        const __half *resp = reinterpret_cast<const __half *>(a.res.ptr) + (long long)n * a.res.imageElems + ((a.resP + yo) * a.res.texW + a.resP + xo) * 4;
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        SK_STAMP(2);
#pragma unroll 1
        for (int i = set; i < nloc; i += a.nsets) {
            const int s = s0 + i;
            const int st = i & (a.ring - 1), use = i >> a.skRingLog2;
            const int tab = stageTab[i], kc = tab >> 24;
            if (i >= a.ring) {          // (the weights of the first `ring` stages were requested by the MMA warp ahead of the wait)
                mbar_wait(&empty[st], (use & 1) ^ 1);
                if (t < 8) {
                    if (t == 0) mbar_expect_tx(&full[st], (uint32_t)bStageBytes);
                    __syncwarp(0xffu);
                    bulk_g2s(sB + (size_t)st * bStageBytes + t * (NTs * 16), wsrc + (size_t)s * (a.NT * kKC * 2 >> 4) + t * a.NT, (uint32_t)(NTs * 16), &full[st]);
                }
            }
            uint2 v[kKC / 4];
            const __half *px = src + (tab & 0xffffff);
            const int4 *org4 = reinterpret_cast<const int4 *>(inOrigin + kc * (kKC / 4));
            if (i == set) SK_STAMP(7);
#pragma unroll
            for (int j4 = 0; j4 < kKC / 16; j4++) {
                const int4 o = org4[j4];
                v[4 * j4 + 0] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.x)) : make_uint2(0u, 0u);
                v[4 * j4 + 1] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.y)) : make_uint2(0u, 0u);
                v[4 * j4 + 2] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.z)) : make_uint2(0u, 0u);
                v[4 * j4 + 3] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.w)) : make_uint2(0u, 0u);
            }
#ifdef FYN_SK_TIMELINE
            if (i == set) {
                SK_STAMP(8);
                unsigned x = 0;
#pragma unroll
                for (int j = 0; j < kKC / 4; j++) x ^= v[j].x ^ v[j].y;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[9]) : "r"(x));
            }
#endif
            if (NORM) {
                // the stand-alone batch-norm layer in front of this convolution, evaluated at the fetch (see k_conv_deep_tc)
                const float4 *sc = sNorm + i * (kKC / 4), *bi = sc + nlocMax * (kKC / 4);
#pragma unroll
                for (int j = 0; j < kKC / 4; j++) {
                    const float4 s4 = sc[j], b4 = bi[j];
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].y));
                    v[j] = make_uint2(pack_half2(fmaf(f0.x, s4.x, b4.x), fmaf(f0.y, s4.y, b4.y)), pack_half2(fmaf(f1.x, s4.z, b4.z), fmaf(f1.y, s4.w, b4.w)));
                }
            }
            unsigned char *dst = sA + (size_t)st * kAStageBytes + t * 16;
#pragma unroll
            for (int c = 0; c < kKC / 8; c++) {
                const uint2 lo = act_h4_t<ACT>(v[2 * c], a.act), hi = act_h4_t<ACT>(v[2 * c + 1], a.act);
                *reinterpret_cast<uint4 *>(dst + c * (kM * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
            if (i == set) SK_STAMP(10);
            fence_proxy_async();
            if (i == set) SK_STAMP(11);
            mbar_arrive(&full[st]);
            if (i == set) SK_STAMP(12);
        }
        SK_STAMP(3);
        // ===================== partial tile -> owners' buffers =====================
        // the residual texels of this thread's planes travel while the MMAs finish
        constexpr int kPre = 4;
        uint2 rqPre[kPre];
#pragma unroll
        for (int g = 0; g < kPre; g++) {
            const int lp = set + g * a.nsets;
            const bool ok = a.hasRes && valid && lp < ppr && rank * ppr + lp < (NTs >> 2) && plane0 + lp < nOutPlanes;
            rqPre[g] = ok ? __ldg(reinterpret_cast<const uint2 *>(resp + resOrigin[lp])) : make_uint2(0u, 0u);
        }
        mbar_wait(done, 0);
        tc_fence_after();
        SK_STAMP(4);
        if (ks > 1) cluster_wait_acquire();            // every CTA of the cluster is running: its buffer may be written
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM lane quarter of this warp
#pragma unroll 1
        for (int cg = set; cg < (NTs >> 4); cg += a.nsets) {
            uint32_t acc[16];
            tmem_ld16(taddr + cg * 16, acc);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = cg * 4 + k, owner = p >> a.skPprLog2, lp = p - (owner << a.skPprLog2);
                const float4 val = make_float4(__uint_as_float(acc[4 * k + 0]), __uint_as_float(acc[4 * k + 1]), __uint_as_float(acc[4 * k + 2]), __uint_as_float(acc[4 * k + 3]));
                float4 *slot = red + ((size_t)rank * ppr + lp) * kM + t;
                if (ks > 1) st_cluster_f4(smem_u32(slot) & 0xffffffu, (uint32_t)owner, val);
                else *slot = val;
            }
        }
        tc_fence_before();
        if (ks > 1) {
            cluster_arrive_release();
            cluster_wait_acquire();
        } else {
            asm volatile("bar.sync 1, %0;" ::"r"(loadWarps * 32) : "memory");
        }
        SK_STAMP(5);
        // ===================== epilogue of this CTA's planes: partials summed in rank order =====================
        // (the residual texels run four planes ahead of their use: rqPre rotates)
#pragma unroll 1
        for (int lp = set; lp < ppr; lp += a.nsets) {
            const uint2 raw = rqPre[0];
            rqPre[0] = rqPre[1];
            rqPre[1] = rqPre[2];
            rqPre[2] = rqPre[3];
            {
                const int ln = lp + kPre * a.nsets;
                const bool ok = a.hasRes && valid && ln < ppr && rank * ppr + ln < (NTs >> 2) && plane0 + ln < nOutPlanes;
                rqPre[3] = ok ? __ldg(reinterpret_cast<const uint2 *>(resp + resOrigin[ok ? ln : 0])) : make_uint2(0u, 0u);
            }
            if (!valid || rank * ppr + lp >= (NTs >> 2) || plane0 + lp >= nOutPlanes) continue;
            float4 acc = red[(size_t)lp * kM + t];
            for (int r = 1; r < ks; r++) {
                const float4 q = red[((size_t)r * ppr + lp) * kM + t];
                acc.x += q.x;
                acc.y += q.y;
                acc.z += q.z;
                acc.w += q.w;
            }
            const float4 sc = sScale[lp], bi = sBias[lp];
            float4 r = make_float4(fmaf(acc.x, sc.x, bi.x), fmaf(acc.y, sc.y, bi.y), fmaf(acc.z, sc.z, bi.z), fmaf(acc.w, sc.w, bi.w));
            if (a.hasRes) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                float4 q = make_float4(f0.x, f0.y, f1.x, f1.y);
                if (a.reluRes) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
                if (a.bnRes) q = make_float4(q.x * sc.x, q.y * sc.y, q.z * sc.z, q.w * sc.w);
                r.x += q.x;
                r.y += q.y;
                r.z += q.z;
                r.w += q.w;
            }
            *reinterpret_cast<uint2 *>(outp + outOrigin[lp]) = make_uint2(pack_half2(r.x, r.y), pack_half2(r.z, r.w));
        }
#ifdef FYN_SK_TIMELINE
        SK_STAMP(6);
        if ((threadIdx.x & 127) == 0) {
            unsigned long long *row = g_skTimeline[a.dbgIndex & 2047];
            atomicMin(&row[0], ts[0]);
            atomicMax(&row[1], ts[6]);
            atomicMin(&row[9], ts[2]);
            if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
            {
                for (int i = 0; i < 7; i++) row[2 + i] = ts[i];
                for (int i = 7; i < 13; i++) row[3 + i] = ts[i];
            }
        }
#endif
    } else {
        // ===================== MMA issuer (and the weights of the first `ring` stages: constants of the layer, requested
        // while the previous layer's kernel is still running) =====================
        if (elect_one()) {
            const int npre = min(nloc, a.ring);
#pragma unroll 1
            for (int i = 0; i < npre; i++) {
                mbar_expect_tx(&full[i], (uint32_t)bStageBytes);
#pragma unroll 1
                for (int c = 0; c < 8; c++)
                    bulk_g2s(sB + (size_t)i * bStageBytes + c * (NTs * 16), wsrc + (size_t)(s0 + i) * (a.NT * kKC * 2 >> 4) + c * a.NT, (uint32_t)(NTs * 16), &full[i]);
            }
        }
        __syncwarp();
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
        const uint32_t aLbo = ((uint32_t)(kM * 16) >> 4) << 16, bLbo = ((uint32_t)(NTs * 16) >> 4) << 16;
#pragma unroll 1
        for (int i = 0; i < nloc; i++) {
            const int st = i & (a.ring - 1), use = i >> a.skRingLog2;
            if (elect_one()) {
                mbar_wait(&full[st], use & 1);
                tc_fence_after();
                // (in a cluster launch the shared-window address of a CTA carries its rank above bit 24: a matrix descriptor takes
                // the CTA-relative offset, 14 bits of 16-byte units)
                const uint32_t a0 = (smem_u32(sA + (size_t)st * kAStageBytes) & 0x3ffffu) >> 4, b0 = (smem_u32(sB + (size_t)st * bStageBytes) & 0x3ffffu) >> 4;
#pragma unroll
                for (int j = 0; j < kKC / 16; j++)
                    umma_f16(tmem, hi | (uint64_t)(aLbo | (a0 + (uint32_t)(j * 2 * kM))), hi | (uint64_t)(bLbo | (b0 + (uint32_t)(j * 2 * NTs))), a.idesc,
                             (i > 0 || j > 0) ? 1u : 0u);
                umma_commit(&empty[st]);
                if (i == nloc - 1) umma_commit(done);
            }
            __syncwarp();
        }
        if (ks > 1) {   // the cluster barriers count every thread
            cluster_wait_acquire();
            cluster_arrive_release();
            cluster_wait_acquire();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == loadWarps) tmem_dealloc(tmem, tmemCols);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant (every grid above the split-K kernel's regime): ONE CTA per SM walks the output tiles (n tile fastest, round
// robin over the CTAs, so that the CTAs working on the N tiles of one pixel tile run at the same time and share its texels
// through L2).  The roles no longer share threads, so the three phases of a tile overlap with those of its neighbours:
//   * loader sets (4 x 128 threads) gather stage after stage, across tile boundaries, as far ahead as the ring allows;
//   * the MMA warp accumulates tile i into one half of TMEM (2 x NT columns) while
//   * eight epilogue warps (two per TMEM lane quarter, splitting the column groups) drain tile i-1 from the other half:
//     tcgen05.ld -> *scale + bias (+ residual, fetched two column groups ahead) -> fp16 texels.
// Same arithmetic and accumulation order as k_conv_deep_tc: the two kernels produce identical bits (tests compare them).
// ---------------------------------------------------------------------------------------------
constexpr int kWarpsP = 25;          // 4 x loader sets + 1 MMA warp + epilogue warps; the split is chosen per layer (args.nsets / args.epiWarps)
constexpr int kMaxRingP = 8;
constexpr int kThreadsDeepP = kWarpsP * 32;

// tile -> (pixel tile, n tile) of a CTA's tile sequence b, b + G, b + 2G, ... without a division per tile
struct TileWalk {
    int mt, nt, dM, dN, ntilesN;
    unsigned tile, total, G;      // (the launcher keeps totalTiles + gridDim below 2^31)
    __device__ __forceinline__ TileWalk(const DeepTcArgs &a) {
        ntilesN = a.ntilesN;
        G = gridDim.x;
        total = (unsigned)a.totalTiles;
        tile = blockIdx.x;
        mt = (int)(blockIdx.x / (unsigned)ntilesN);
        nt = (int)(blockIdx.x % (unsigned)ntilesN);
        dM = (int)(gridDim.x / (unsigned)ntilesN);
        dN = (int)(gridDim.x % (unsigned)ntilesN);
    }
    __device__ __forceinline__ bool done() const { return tile >= total; }
    __device__ __forceinline__ void next() {
        tile += G;
        mt += dM;
        nt += dN;
        if (nt >= ntilesN) {
            nt -= ntilesN;
            mt++;
        }
    }
};

// position of a GEMM row (output pixel) in the output / residual tensors
struct Px {
    bool valid;
    int plane0;
    __half *outp;
    const __half *resp;
};

// HALO = false: GEMM rows are the flattened output pixels (image, y, x).
// HALO = true (k_conv_deep_tc_h3): rows walk a virtual raster of Wv = W + 1 columns and H + 1 rows per image whose extra column /
// row stand for the zero border; those rows are computed but not stored.
template <bool HALO>
__device__ __forceinline__ Px locate_row(const DeepTcArgs &a, const TileWalk &w, int t) {
    Px p;
    const unsigned m = (unsigned)w.mt * kM + t;
    unsigned n, yo, xo;
    if (HALO) {
        const unsigned vr = m / (unsigned)a.Wv;
        xo = m - vr * (unsigned)a.Wv;
        n = vr / (unsigned)(a.Ho + 1);
        yo = vr - n * (unsigned)(a.Ho + 1);
        p.valid = !w.done() && m < (unsigned)a.Mtotal && xo < (unsigned)a.Wo && yo < (unsigned)a.Ho;
    } else {
        const unsigned hw = (unsigned)(a.Ho * a.Wo);
        p.valid = !w.done() && m < (unsigned)a.Mtotal;
        n = p.valid ? m / hw : 0u;
        const unsigned rem = p.valid ? m - n * hw : 0u;
        yo = rem / (unsigned)a.Wo;
        xo = rem - yo * (unsigned)a.Wo;
    }
    if (!p.valid) n = yo = xo = 0u;
    p.outp = reinterpret_cast<__half *>(a.out.ptr) + (long long)n * a.out.imageElems + ((a.outP + (int)yo) * a.out.texW + a.outP + (int)xo) * 4;
    p.resp = reinterpret_cast<const __half *>(a.res.ptr) + (long long)n * a.res.imageElems + ((a.resP + (int)yo) * a.res.texW + a.resP + (int)xo) * 4;
    p.plane0 = w.nt * (a.NT >> 2);
    return p;
}

// Epilogue warp of the persistent kernels: drains the CTA's tiles from the two TMEM halves.  `e` = index among the epilogue
// warps, `quarter` = TMEM lane quarter of the warp (warp index & 3); the warps of a quarter alternate over the column groups.
// RES: residual handling at compile time (the epilogue is issue-bound: ~70 instructions per texel with run-time flags):
// 0 = no residual, 1 = residual with ReLU (every ResNet bottleneck output), 2 = whatever the arguments say
template <bool HALO, int RES>
__device__ __forceinline__ void epilogue_tiles_t(const DeepTcArgs &a, int e, int quarter, uint32_t tmem, uint64_t *accFull, uint64_t *accEmpty,
                                                 const float4 *sScale, const float4 *sBias, const int *outOrigin, const int *resOrigin) {
    const bool hasRes = RES == 2 ? a.hasRes != 0 : RES == 1, reluRes = RES == 2 ? a.reluRes != 0 : true, bnRes = RES == 2 ? a.bnRes != 0 : false;
    const int part = e >> 2, lane = threadIdx.x & 31, parts = a.epiWarps >> 2;
    const int t = quarter * 32 + lane;
    const int ngroups = a.NT >> 4, nOutPlanes = a.Cout4 >> 2;
    const int myGroups = part < ngroups ? (ngroups - part + parts - 1) / parts : 0;
    uint32_t tcount = 0;
    auto fetch = [&](const Px &p, int cg, uint2 (&rq)[4]) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pk = p.plane0 + cg * 4 + k;
            rq[k] = (hasRes && p.valid && cg < ngroups && pk < nOutPlanes) ? __ldg(reinterpret_cast<const uint2 *>(p.resp + resOrigin[pk])) : make_uint2(0u, 0u);
        }
    };
    auto finish = [&](const Px &p, int cg, const uint32_t (&acc)[16], const uint2 (&rq)[4]) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pk = p.plane0 + cg * 4 + k;
            if (!p.valid || pk >= nOutPlanes) continue;
            const float4 sc = sScale[pk], bi = sBias[pk];
            float4 r = make_float4(fmaf(__uint_as_float(acc[4 * k + 0]), sc.x, bi.x), fmaf(__uint_as_float(acc[4 * k + 1]), sc.y, bi.y),
                                   fmaf(__uint_as_float(acc[4 * k + 2]), sc.z, bi.z), fmaf(__uint_as_float(acc[4 * k + 3]), sc.w, bi.w));
            if (hasRes) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&rq[k].x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&rq[k].y));
                float4 q = make_float4(f0.x, f0.y, f1.x, f1.y);
                if (reluRes) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
                if (bnRes) q = make_float4(q.x * sc.x, q.y * sc.y, q.z * sc.z, q.w * sc.w);
                r.x += q.x;
                r.y += q.y;
                r.z += q.z;
                r.w += q.w;
            }
            *reinterpret_cast<uint2 *>(p.outp + outOrigin[pk]) = make_uint2(pack_half2(r.x, r.y), pack_half2(r.z, r.w));
        }
    };
    TileWalk w(a);
    Px cur = locate_row<HALO>(a, w, t);
    if (myGroups <= 4) {
        // At most four column groups per warp and tile (the usual split): the residual texels of the NEXT tile are fetched while
        // this one is drained -- up to sixteen 8-byte loads in flight per thread, a whole tile ahead of their use.
        uint2 rq[4][4];
#pragma unroll
        for (int g = 0; g < 4; g++)
            if (g < myGroups) fetch(cur, part + g * parts, rq[g]);
        for (; !w.done(); tcount++) {
            const uint32_t buf = tcount & 1;
            TileWalk wn = w;
            wn.next();
            const Px nxt = locate_row<HALO>(a, wn, t);
            mbar_wait(&accFull[buf], (tcount >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)a.NT;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                if (g == myGroups) {   // every column group of this warp has been read: the accumulator half may be overwritten
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accEmpty[buf]);
                }
                if (g >= myGroups) continue;
                const int cg = part + g * parts;
                uint32_t acc[16];
                tmem_ld16(taddr + cg * 16, acc);
                tmem_ld_wait();
                if (g == myGroups - 1 && g == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accEmpty[buf]);
                }
                finish(cur, cg, acc, rq[g]);
                fetch(nxt, cg, rq[g]);
            }
            cur = nxt;
            w = wn;
        }
    } else {
        // more column groups per warp: the residual texels run two groups ahead inside the tile
        for (; !w.done(); w.next(), cur = locate_row<HALO>(a, w, t), tcount++) {
            const uint32_t buf = tcount & 1;
            uint2 rq0[4], rq1[4];
            fetch(cur, part, rq0);
            fetch(cur, part + parts, rq1);
            mbar_wait(&accFull[buf], (tcount >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)a.NT;
            auto do_group = [&](int cg, uint2 (&rq)[4]) {
                uint32_t acc[16];
                tmem_ld16(taddr + cg * 16, acc);
                tmem_ld_wait();
                if (cg + parts >= ngroups) {   // last column group of this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accEmpty[buf]);
                }
                finish(cur, cg, acc, rq);
                fetch(cur, cg + 2 * parts, rq);
            };
            for (int cg = part; cg < ngroups; cg += 2 * parts) {
                do_group(cg, rq0);
                if (cg + parts < ngroups) do_group(cg + parts, rq1);
            }
        }
    }
}


template <bool HALO>
__device__ __forceinline__ void epilogue_tiles(const DeepTcArgs &a, int e, int quarter, uint32_t tmem, uint64_t *accFull, uint64_t *accEmpty,
                                               const float4 *sScale, const float4 *sBias, const int *outOrigin, const int *resOrigin) {
    if (!a.hasRes) epilogue_tiles_t<HALO, 0>(a, e, quarter, tmem, accFull, accEmpty, sScale, sBias, outOrigin, resOrigin);
    else if (a.reluRes && !a.bnRes) epilogue_tiles_t<HALO, 1>(a, e, quarter, tmem, accFull, accEmpty, sScale, sBias, outOrigin, resOrigin);
    else epilogue_tiles_t<HALO, 2>(a, e, quarter, tmem, accFull, accEmpty, sScale, sBias, outOrigin, resOrigin);
}

// ACT: activation at the fetch as a compile-time constant (0 none, 1 ReLU, 2 whatever args.act says)
template <bool NORM, int ACT>
__global__ void __launch_bounds__(kThreadsDeepP, 1) k_conv_deep_tc_p(const __grid_constant__ DeepTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int bStageBytes = a.NT * kKC * 2;
    const int nOutPlanes = a.Cout4 >> 2;
    unsigned char *sA = smem;
    unsigned char *sB = sA + a.ring * kAStageBytes;
    float4 *sScale = reinterpret_cast<float4 *>(sB + a.ring * bStageBytes);   // [nOutPlanes] scale, then [nOutPlanes] bias
    float4 *sBias = sScale + nOutPlanes;
    float4 *sInScale = sBias + nOutPlanes;                                    // fused input batch-norm: [nInPlanes] scale, [nInPlanes] bias (NORM only)
    float4 *sInBias = sInScale + (NORM ? a.nInPlanes : 0);
    int *tapTab = reinterpret_cast<int *>(sInBias + (NORM ? a.nInPlanes : 0));   // [64] (16-byte aligned: read as int4)
    int *tapOff = tapTab + 64;                                                // [64] element offset of a tap-packed stage's taps (interior pixels)
    int *inOrigin = tapOff + 64;                                              // [nInPlanes] (16-byte aligned)
    int *outOrigin = inOrigin + a.nInPlanes;                                  // [nOutPlanes] output tensor, then [nOutPlanes] residual tensor
    int *resOrigin = outOrigin + nOutPlanes;
    int *stageTab = resOrigin + nOutPlanes;                                   // [nstages] ky | kx << 8 | kc << 16
    uint64_t *full = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(stageTab + a.nstages) + 7) & ~uintptr_t(7));
    uint64_t *empty = full + kMaxRingP;
    uint64_t *accFull = empty + kMaxRingP;
    uint64_t *accEmpty = accFull + 2;
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(accEmpty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int loadWarps = 4 * a.nsets, nthreads = kThreadsDeepP;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.ring; s++) {
            mbar_init(&full[s], kM + 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&accFull[b], 1);
            mbar_init(&accEmpty[b], (uint32_t)a.epiWarps);
        }
        fence_barrier_init();
    }
    const uint32_t tmemCols = a.NT > 128 ? 512u : (a.NT > 64 ? 256u : 128u);
    if (warp == loadWarps) tmem_alloc(tmemBase, tmemCols);
    for (int q = threadIdx.x; q < a.nInPlanes && !a.tapPacked; q += nthreads) inOrigin[q] = ((q / a.in.tx) * a.in.tileH * a.in.texW + (q % a.in.tx) * a.in.tileW) * 4;
    for (int p = threadIdx.x; p < nOutPlanes; p += nthreads) {
        outOrigin[p] = ((p / a.out.tx) * a.out.tileH * a.out.texW + (p % a.out.tx) * a.out.tileW) * 4;
        resOrigin[p] = a.hasRes ? ((p / a.res.tx) * a.res.tileH * a.res.texW + (p % a.res.tx) * a.res.tileW) * 4 : 0;
        // (the parameters are constants of the layer, not results of the previous kernel: read ahead of griddepcontrol.wait)
        sScale[p] = __ldg(reinterpret_cast<const float4 *>(a.scale) + p);
        sBias[p] = __ldg(reinterpret_cast<const float4 *>(a.bias) + p);
    }
    if (NORM)
        for (int q = threadIdx.x; q < 2 * a.nInPlanes; q += nthreads) sInScale[q] = __ldg(a.inNorm + q);   // (scale planes, then bias planes)
    if (a.tapPacked && threadIdx.x < 64) {
        const int tp = min((int)threadIdx.x, a.K * a.K - 1);
        tapTab[threadIdx.x] = (tp / a.K) | ((tp % a.K) << 8);
        tapOff[threadIdx.x] = ((tp / a.K) * a.in.texW + (tp % a.K)) * 4;
    }
    for (int st = threadIdx.x; st < a.nstages; st += nthreads) {
        const int tap = st / a.kcs, kc = st - tap * a.kcs;
        stageTab[st] = (tap / a.K) | ((tap % a.K) << 8) | (kc << 16);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned hw = (unsigned)(a.Ho * a.Wo), Mtotal = (unsigned)a.Mtotal;

    if (warp < loadWarps) {
        // ===================== loaders: thread = GEMM row = output pixel; set `set` takes every nsets-th stage of the CTA's stage stream
        const int t = threadIdx.x & (kM - 1), set = warp >> 2;
        int s = set;                                       // tile-local index of the set's next stage
        int slot = set;                                    // ring slot / phase of that stage (nsets <= ring)
        uint32_t phase = 0;
        for (TileWalk w(a); !w.done(); w.next(), s -= a.nstages) {
            if (s >= a.nstages) continue;                  // (tiles of fewer stages than sets: nothing of this tile is ours)
            const unsigned m = (unsigned)w.mt * kM + t;
            const bool valid = m < Mtotal;
            const unsigned n = valid ? m / hw : 0u;
            const unsigned rem = valid ? m - n * hw : 0u;
            const int yo = (int)(rem / (unsigned)a.Wo), xo = (int)rem - yo * a.Wo;
            const __half *src = reinterpret_cast<const __half *>(a.in.ptr) + (long long)n * a.in.imageElems;
            const uint4 *wsrc = a.wimg + (size_t)w.nt * a.nstages * (bStageBytes >> 4);
            const int iy0 = a.inP + a.ds * yo - a.mh, ix0 = a.inP + a.ds * xo - a.mh;
            const __half *px0 = src + (iy0 * a.in.texW + ix0) * 4;
            const bool interior = valid && iy0 >= 0 && ix0 >= 0 && iy0 + a.K <= a.in.texH && ix0 + a.K <= a.in.texW;
            for (; s < a.nstages; s += a.nsets) {
                const int st = slot;
                const int tab = stageTab[s];
                const int ky = tab & 255, kx = (tab >> 8) & 255, kc = tab >> 16;
                mbar_wait(&empty[st], phase ^ 1);
                if (t == 0) {
                    mbar_expect_tx(&full[st], (uint32_t)bStageBytes);
                    bulk_g2s(sB + (size_t)st * bStageBytes, wsrc + (size_t)s * (bStageBytes >> 4), (uint32_t)bStageBytes, &full[st]);
                }
                uint2 v[kKC / 4];
                if (!a.tapPacked) {
                    const __half *px = px0 + (ky * a.in.texW + kx) * 4;
                    const int4 *org4 = reinterpret_cast<const int4 *>(inOrigin + kc * (kKC / 4));
#pragma unroll
                    for (int j4 = 0; j4 < kKC / 16; j4++) {
                        const int4 o = org4[j4];
                        v[4 * j4 + 0] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.x)) : make_uint2(0u, 0u);
                        v[4 * j4 + 1] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.y)) : make_uint2(0u, 0u);
                        v[4 * j4 + 2] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.z)) : make_uint2(0u, 0u);
                        v[4 * j4 + 3] = valid ? __ldg(reinterpret_cast<const uint2 *>(px + o.w)) : make_uint2(0u, 0u);
                    }
                    if (NORM) {
                        const float4 *sc = sInScale + kc * (kKC / 4), *bi = sInBias + kc * (kKC / 4);
#pragma unroll
                        for (int j = 0; j < kKC / 4; j++) {
                            const float4 s4 = sc[j], b4 = bi[j];
                            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].x));
                            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v[j].y));
                            v[j] = make_uint2(pack_half2(fmaf(f0.x, s4.x, b4.x), fmaf(f0.y, s4.y, b4.y)), pack_half2(fmaf(f1.x, s4.z, b4.z), fmaf(f1.y, s4.w, b4.w)));
                        }
                    }
                } else if (interior) {
                    // every tap of this pixel lies inside the texture: base + tap offset, no clamping
                    // (the last stage of a 7x7 kernel holds one tap: only its first K16 step is gathered and multiplied)
                    const int nq = (s == a.nstages - 1) ? a.lastQuads : kKC / 16;
                    const int4 *off4 = reinterpret_cast<const int4 *>(tapOff + s * (kKC / 4));
#pragma unroll
                    for (int j4 = 0; j4 < kKC / 16; j4++) {
                        if (j4 < nq) {
                            const int4 o = off4[j4];
                            v[4 * j4 + 0] = __ldg(reinterpret_cast<const uint2 *>(px0 + o.x));
                            v[4 * j4 + 1] = __ldg(reinterpret_cast<const uint2 *>(px0 + o.y));
                            v[4 * j4 + 2] = __ldg(reinterpret_cast<const uint2 *>(px0 + o.z));
                            v[4 * j4 + 3] = __ldg(reinterpret_cast<const uint2 *>(px0 + o.w));
                        } else {
                            v[4 * j4 + 0] = v[4 * j4 + 1] = v[4 * j4 + 2] = v[4 * j4 + 3] = make_uint2(0u, 0u);
                        }
                    }
                } else {
                    const int ntaps = (s == a.nstages - 1) ? 4 * a.lastQuads : kKC / 4;
#pragma unroll
                    for (int j = 0; j < kKC / 4; j++) {
                        v[j] = make_uint2(0u, 0u);
                        if (j >= ntaps) continue;
                        const int tt = tapTab[s * (kKC / 4) + j];
                        const int tky = tt & 255, tkx = tt >> 8;
                        const int iy = min(max(a.inP + a.ds * yo + tky - a.mh, 0), a.in.texH - 1);
                        const int ix = min(max(a.inP + a.ds * xo + tkx - a.mh, 0), a.in.texW - 1);
                        if (valid) v[j] = __ldg(reinterpret_cast<const uint2 *>(src + (iy * a.in.texW + ix) * 4));
                    }
                }
                unsigned char *dst = sA + (size_t)st * kAStageBytes + t * 16;
#pragma unroll
                for (int c = 0; c < kKC / 8; c++) {
                    const uint2 lo = act_h4_t<ACT>(v[2 * c], a.act), hi = act_h4_t<ACT>(v[2 * c + 1], a.act);
                    *reinterpret_cast<uint4 *>(dst + c * (kM * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                }
                fence_proxy_async();
                mbar_arrive(&full[st]);
                slot += a.nsets;
                if (slot >= a.ring) {
                    slot -= a.ring;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == loadWarps) {
        // ===================== MMA issuer: tile i accumulates into TMEM half (i & 1)
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
        const uint32_t aLbo = ((uint32_t)(kM * 16) >> 4) << 16, bLbo = ((uint32_t)(a.NT * 16) >> 4) << 16;
        if (elect_one()) {
            int slot = 0;
            uint32_t phase = 0, tcount = 0;
            for (TileWalk w(a); !w.done(); w.next(), tcount++) {
                const uint32_t buf = tcount & 1;
                mbar_wait(&accEmpty[buf], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem + buf * (uint32_t)a.NT;
                for (int s = 0; s < a.nstages; s++) {
                    mbar_wait(&full[slot], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)slot * kAStageBytes) >> 4, b0 = smem_u32(sB + (size_t)slot * bStageBytes) >> 4;
                    const int nq = (a.tapPacked && s == a.nstages - 1) ? a.lastQuads : kKC / 16;
#pragma unroll
                    for (int j = 0; j < kKC / 16; j++)
                        if (j < nq)
                            umma_f16(tacc, hi | (uint64_t)(aLbo | (a0 + (uint32_t)(j * 2 * kM))), hi | (uint64_t)(bLbo | (b0 + (uint32_t)(j * 2 * a.NT))), a.idesc,
                                     (s > 0 || j > 0) ? 1u : 0u);
                    umma_commit(&empty[slot]);
                    if (++slot == a.ring) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&accFull[buf]);
            }
        }
        __syncwarp();
    } else if (warp < loadWarps + 1 + a.epiWarps) {
        epilogue_tiles<false>(a, warp - loadWarps - 1, warp & 3, tmem, accFull, accEmpty, sScale, sBias, outOrigin, resOrigin);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == loadWarps) tmem_dealloc(tmem, tmemCols);
}

// ---------------------------------------------------------------------------------------------
// Halo-tile kernel for 3x3 stride-1 layers on large grids.  k_conv_deep_tc_p gathers every input texel nine times (one A
// stage per tap); here a stage holds the HALO of a tile -- its 128 GEMM rows plus one virtual row and one pixel on either
// side -- for 64 input channels, gathered ONCE, and the nine taps are nine tcgen05.mma groups whose A descriptors start
// (ky * Wv + kx) pixels into that stage (SWIZZLE_NONE K-major: a pixel is 16 bytes in every 8-channel chunk, so any
// pixel shift is a legal descriptor start).  For the shift to be the same for every row, the GEMM rows walk a VIRTUAL
// raster: Wv = W + 1 columns and H + 1 rows per image, the extra column / row being the zero border that both neighbours
// share (the deep layout's own idea, deeptiler.cpp:63-95); border rows are computed and dropped by the epilogue
// (W = 56: 3.4 % of the rows, W = 7: 23 %).  Weights travel in their own ring, one (tap, 64 channels) block per entry,
// fetched by a dedicated producer warp.  K order: channel stages outermost, taps inside (k_conv_deep_tc walks taps
// outermost): same products, different fp32 summation order -- not bit-identical to the other kernels, same tolerance.
// ---------------------------------------------------------------------------------------------
constexpr int kThreadsDeepH = (kWarpsP + 1) * 32;

template <int ACT>
__global__ void __launch_bounds__(kThreadsDeepH, 1) k_conv_deep_tc_h3(const __grid_constant__ DeepTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int aStageBytes = a.HPp * 128, bStageBytes = a.NT * kKC * 2;
    const int nOutPlanes = a.Cout4 >> 2;
    unsigned char *sA = smem;
    unsigned char *sB = sA + a.ring * aStageBytes;
    float4 *sScale = reinterpret_cast<float4 *>(sB + a.ringB * bStageBytes);
    float4 *sBias = sScale + nOutPlanes;
    int *inOrigin = reinterpret_cast<int *>(sBias + nOutPlanes);             // [nInPlanes] (16-byte aligned)
    int *outOrigin = inOrigin + a.nInPlanes;
    int *resOrigin = outOrigin + nOutPlanes;
    uint64_t *fullA = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(resOrigin + nOutPlanes) + 7) & ~uintptr_t(7));
    uint64_t *emptyA = fullA + kMaxRingP;
    uint64_t *fullB = emptyA + kMaxRingP;
    uint64_t *emptyB = fullB + kMaxRingP;
    uint64_t *accFull = emptyB + kMaxRingP;
    uint64_t *accEmpty = accFull + 2;
    uint32_t *tmemBase = reinterpret_cast<uint32_t *>(accEmpty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int loadWarps = 4 * a.nsets, nthreads = kThreadsDeepH;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.ring; s++) {
            mbar_init(&fullA[s], kM);
            mbar_init(&emptyA[s], 1);
        }
        for (int s = 0; s < a.ringB; s++) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&accFull[b], 1);
            mbar_init(&accEmpty[b], (uint32_t)a.epiWarps);
        }
        fence_barrier_init();
    }
    const uint32_t tmemCols = a.NT > 128 ? 512u : (a.NT > 64 ? 256u : 128u);
    if (warp == loadWarps) tmem_alloc(tmemBase, tmemCols);
    for (int q = threadIdx.x; q < a.nInPlanes; q += nthreads) inOrigin[q] = ((q / a.in.tx) * a.in.tileH * a.in.texW + (q % a.in.tx) * a.in.tileW) * 4;
    for (int p = threadIdx.x; p < nOutPlanes; p += nthreads) {
        outOrigin[p] = ((p / a.out.tx) * a.out.tileH * a.out.texW + (p % a.out.tx) * a.out.tileW) * 4;
        resOrigin[p] = a.hasRes ? ((p / a.res.tx) * a.res.tileH * a.res.texW + (p % a.res.tx) * a.res.tileW) * 4 : 0;
        sScale[p] = __ldg(reinterpret_cast<const float4 *>(a.scale) + p);
        sBias[p] = __ldg(reinterpret_cast<const float4 *>(a.bias) + p);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemBase;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < loadWarps) {
        // ===================== loaders: a set gathers the halo of one (tile, 64-channel stage): thread = halo pixel t and t + 128
        const int t = threadIdx.x & (kM - 1), set = warp >> 2;
        const int Hv = a.Ho + 1;
        int s = set;                                       // tile-local index of the set's next channel stage
        int slot = set;
        uint32_t phase = 0;
        for (TileWalk w(a); !w.done(); w.next(), s -= a.kcs) {
            if (s >= a.kcs) continue;
            // virtual raster position of halo pixel h: j = tile start - Wv - 1 + h
            const __half *px[2];
            bool ok[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int h = t + r * kM;
                const int j = w.mt * kM - a.Wv - 1 + h;
                ok[r] = h < a.HPp && j >= 0 && j < (int)a.Mtotal;
                const unsigned ju = ok[r] ? (unsigned)j : 0u;
                const unsigned vr = ju / (unsigned)a.Wv, c = ju - vr * (unsigned)a.Wv;
                const unsigned n = vr / (unsigned)Hv, y = vr - n * (unsigned)Hv;
                ok[r] = ok[r] && c < (unsigned)a.Wo && y < (unsigned)a.Ho;
                px[r] = reinterpret_cast<const __half *>(a.in.ptr) + (long long)n * a.in.imageElems + ((a.inP + (int)y) * a.in.texW + a.inP + (int)c) * 4;
            }
            for (; s < a.kcs; s += a.nsets) {
                const int st = slot;
                mbar_wait(&emptyA[st], phase ^ 1);
                const int4 *org4 = reinterpret_cast<const int4 *>(inOrigin + s * (kKC / 4));
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int h = t + r * kM;
                    if (h >= a.HPp) continue;
                    uint2 v[kKC / 4];
#pragma unroll
                    for (int j4 = 0; j4 < kKC / 16; j4++) {
                        const int4 o = org4[j4];
                        v[4 * j4 + 0] = ok[r] ? __ldg(reinterpret_cast<const uint2 *>(px[r] + o.x)) : make_uint2(0u, 0u);
                        v[4 * j4 + 1] = ok[r] ? __ldg(reinterpret_cast<const uint2 *>(px[r] + o.y)) : make_uint2(0u, 0u);
                        v[4 * j4 + 2] = ok[r] ? __ldg(reinterpret_cast<const uint2 *>(px[r] + o.z)) : make_uint2(0u, 0u);
                        v[4 * j4 + 3] = ok[r] ? __ldg(reinterpret_cast<const uint2 *>(px[r] + o.w)) : make_uint2(0u, 0u);
                    }
                    unsigned char *dst = sA + (size_t)st * aStageBytes + h * 16;
#pragma unroll
                    for (int c = 0; c < kKC / 8; c++) {
                        const uint2 lo = act_h4_t<ACT>(v[2 * c], a.act), hi = act_h4_t<ACT>(v[2 * c + 1], a.act);
                        *reinterpret_cast<uint4 *>(dst + c * (a.HPp * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&fullA[st]);
                slot += a.nsets;
                if (slot >= a.ring) {
                    slot -= a.ring;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == loadWarps) {
        // ===================== MMA issuer: per channel stage nine tap groups of four K16 MMAs on the same halo
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
        const uint32_t aLbo = ((uint32_t)(a.HPp * 16) >> 4) << 16, bLbo = ((uint32_t)(a.NT * 16) >> 4) << 16;
        if (elect_one()) {
            int slotA = 0, slotB = 0;
            uint32_t phaseA = 0, phaseB = 0, tcount = 0;
            for (TileWalk w(a); !w.done(); w.next(), tcount++) {
                const uint32_t buf = tcount & 1;
                mbar_wait(&accEmpty[buf], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem + buf * (uint32_t)a.NT;
                for (int kc = 0; kc < a.kcs; kc++) {
                    mbar_wait(&fullA[slotA], phaseA);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)slotA * aStageBytes) >> 4;
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(&fullB[slotB], phaseB);
                        tc_fence_after();
                        const uint32_t b0 = smem_u32(sB + (size_t)slotB * bStageBytes) >> 4;
                        const uint32_t shift = (uint32_t)((tap / 3) * a.Wv + tap % 3);     // pixels = 16-byte units
#pragma unroll
                        for (int j = 0; j < kKC / 16; j++)
                            umma_f16(tacc, hi | (uint64_t)(aLbo | (a0 + shift + (uint32_t)(j * 2 * a.HPp))), hi | (uint64_t)(bLbo | (b0 + (uint32_t)(j * 2 * a.NT))),
                                     a.idesc, (kc > 0 || tap > 0 || j > 0) ? 1u : 0u);
                        umma_commit(&emptyB[slotB]);
                        if (++slotB == a.ringB) {
                            slotB = 0;
                            phaseB ^= 1;
                        }
                    }
                    umma_commit(&emptyA[slotA]);
                    if (++slotA == a.ring) {
                        slotA = 0;
                        phaseA ^= 1;
                    }
                }
                umma_commit(&accFull[buf]);
            }
        }
        __syncwarp();
    } else if (warp < loadWarps + 1 + a.epiWarps) {
        epilogue_tiles<true>(a, warp - loadWarps - 1, warp & 3, tmem, accFull, accEmpty, sScale, sBias, outOrigin, resOrigin);
    } else if (warp == kWarpsP) {
        // ===================== weight producer: one (tap, 64 channels) block per ring entry, in the MMA issuer's order
        if (elect_one()) {
            int slotB = 0;
            uint32_t phaseB = 0;
            for (TileWalk w(a); !w.done(); w.next()) {
                const uint4 *wsrc = a.wimg + (size_t)w.nt * a.nstages * (bStageBytes >> 4);
                for (int kc = 0; kc < a.kcs; kc++)
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(&emptyB[slotB], phaseB ^ 1);
                        mbar_expect_tx(&fullB[slotB], (uint32_t)bStageBytes);
                        bulk_g2s(sB + (size_t)slotB * bStageBytes, wsrc + (size_t)(tap * a.kcs + kc) * (bStageBytes >> 4), (uint32_t)bStageBytes, &fullB[slotB]);
                        if (++slotB == a.ringB) {
                            slotB = 0;
                            phaseB ^= 1;
                        }
                    }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == loadWarps) tmem_dealloc(tmem, tmemCols);
}

}  // namespace

struct DeepTcPlan {
    DeepTcArgs args{};
    uint4 *d_wimg = nullptr;
    size_t wimgBytes = 0;
    size_t smemBytes = 0;
    int ntiles = 0;
    // second operand image with 256 columns per CTA (short-K layers with >= 256 outputs), chosen at run time for large grids
    size_t wideOff = 0, wideSmemBytes = 0;   // offset into d_wimg in uint4 units; 0 = not available
    int wideNtiles = 0;
};

int fyn_conv_deep_tc_supported(const fyn_conv_desc *d) {
    if (!(d->flags & FYN_FLAG_DEEP) || d->fractional || d->dilation != 1) return 0;
    if (d->downsample != 1 && d->downsample != 2) return 0;
    if (d->in_channels <= 4) {
        if (d->kernel > 7) return 0;                          // tap-packed mode (single plane): <= 49 taps = 4 stages
    } else {
        if (d->kernel != 1 && d->kernel != 3) return 0;
        if (d->in_channels % kKC != 0) return 0;
        if (d->in_padding < (d->kernel - 1) / 2) return 0;   // the tile padding must supply the border zeros
    }
    if (d->flags & FYN_FLAG_PRE_CLIP) return 0;
    return 1;
}

int fyn_conv_deep_tc_create(fyn_op *op, const float *wb) {
    const fyn_conv_desc &d = op->conv;
    DeepTcPlan *plan = op->dtc ? op->dtc : new DeepTcPlan();
    op->dtc = plan;
    DeepTcArgs &a = plan->args;
    const int K = d.kernel, Ci = d.in_channels, Co = d.out_channels;
    const int co16 = ((Co + 15) / 16) * 16;
    a.NT = std::min(128, co16);
    // short-K layers with many outputs (1x1 expansions on 64 / 128 channels): 256 columns per CTA halve the number of CTAs and
    // of input gathers; their ring of <= 2 stages (<= 96 KB) still lets two CTAs (2 x 256 TMEM columns) share an SM.  Small
    // grids (batch 1) prefer the 128-column tiles, so both operand images are kept and the run picks one.
    // (FYN_DEEP_WIDE: 0 = never, 1 = short-K layers only, 2 = every layer with >= 256 outputs: multi-stage layers then run one
    // CTA per SM on a ring of four 48 KB stages, but gather every input tile half as often -- +4 % at batch 128)
    static const int wideN = getenv("FYN_DEEP_WIDE") ? atoi(getenv("FYN_DEEP_WIDE")) : 2;
    const bool wide = wideN && Ci > 4 && (K * K * (Ci / kKC) <= 2 || wideN == 2) && co16 >= 256;
    plan->ntiles = (co16 + a.NT - 1) / a.NT;
    a.K = K;
    a.ds = d.downsample;
    a.mh = (K - 1) / 2;
    a.tapPacked = Ci <= 4 ? 1 : 0;
    a.kcs = a.tapPacked ? 1 : Ci / kKC;
    a.nstages = a.tapPacked ? (K * K + 15) / 16 : K * K * a.kcs;
    a.lastQuads = a.tapPacked ? (K * K - 16 * (a.nstages - 1) + 3) / 4 : kKC / 16;
    if (getenv("FYN_DEEP_STEM_FULL")) a.lastQuads = kKC / 16;      // (measurement knob: gather and multiply the zero-weight taps too)
    a.nInPlanes = (Ci + 3) / 4;
    a.Cout4 = ((Co + 3) / 4) * 4;
    a.idesc = (1u << 4) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);   // F32 accum, F16 x F16, K-major A/B
    // weight image [ntile][stage = (tap, kc)][chunk][n][8], fp16 truncated (gpu/floatconversion.cpp:44-58)
    const size_t stageHalfs = (size_t)8 * a.NT * 8;
    std::vector<__half> img((size_t)plan->ntiles * a.nstages * stageHalfs, __float2half(0.f));
    const float *W = wb + Co;   // [Co][K][K][Ci]
    if (a.tapPacked) {
        // k index inside a stage = (tap % 16) * 4 + channel: chunk c holds taps 2c, 2c+1 of the stage
        for (int nt = 0; nt < plan->ntiles; nt++)
            for (int tap = 0; tap < K * K; tap++) {
                __half *dst = img.data() + ((size_t)nt * a.nstages + tap / 16) * stageHalfs;
                const int j = tap % 16;
                for (int n = 0; n < a.NT; n++) {
                    const int o = nt * a.NT + n;
                    if (o >= Co) continue;
                    for (int c = 0; c < Ci; c++)
                        dst[((size_t)(j / 2) * a.NT + n) * 8 + (j & 1) * 4 + c] = __float2half_rz(fyn_half_trunc_host(W[((size_t)o * K * K + tap) * Ci + c]));
                }
            }
    }
    auto packImage = [&](__half *base, int NT, int ntiles) {
        const size_t sh = (size_t)8 * NT * 8;
        for (int nt = 0; nt < ntiles; nt++)
            for (int tap = 0; tap < K * K; tap++)
                for (int kc = 0; kc < a.kcs; kc++) {
                    __half *dst = base + ((size_t)nt * a.nstages + (size_t)tap * a.kcs + kc) * sh;
                    for (int n = 0; n < NT; n++) {
                        const int o = nt * NT + n;
                        if (o >= Co) continue;
                        const float *w = W + ((size_t)o * K * K + tap) * Ci + (size_t)kc * kKC;
                        for (int c = 0; c < 8; c++)
                            for (int e = 0; e < 8; e++) dst[((size_t)c * NT + n) * 8 + e] = __float2half_rz(fyn_half_trunc_host(w[c * 8 + e]));
                    }
                }
    };
    if (!a.tapPacked) packImage(img.data(), a.NT, plan->ntiles);
    plan->wideOff = 0;
    if (wide) {
        plan->wideNtiles = (co16 + 255) / 256;
        plan->wideOff = img.size() / 8;                                  // halfs -> uint4
        img.resize(img.size() + (size_t)plan->wideNtiles * a.nstages * 8 * 256 * 8, __float2half(0.f));
        packImage(img.data() + plan->wideOff * 8, 256, plan->wideNtiles);
    }
    const size_t bytes = img.size() * sizeof(__half);
    FYN_CUDA(cudaSetDevice(op->ctx->device));
    if (plan->d_wimg && plan->wimgBytes < bytes) {
        cudaFree(plan->d_wimg);
        plan->d_wimg = nullptr;
    }
    if (!plan->d_wimg) {
        FYN_CUDA(cudaMalloc((void **)&plan->d_wimg, bytes));
        plan->wimgBytes = bytes;
    }
    FYN_CUDA(cudaMemcpy(plan->d_wimg, img.data(), bytes, cudaMemcpyHostToDevice));
    a.wimg = plan->d_wimg;
    // loader sets / ring depth: registers allow two CTAs of 4 sets, four of 2 sets; a ring entry is 16 KB + NT * 128 bytes
    // (all sets also for single-stage layers: the sets split the epilogue's column groups -- 1x1 64->256 @56x56 with residual
    // takes 68 us with four sets, 98 us with one)
    a.nsets = kLoadSets;
    a.ring = std::min(kMaxRing, a.nstages);
    if (const char *e = getenv("FYN_DEEP_SETS")) a.nsets = std::max(1, std::min(kLoadSets, atoi(e)));
    if (const char *e = getenv("FYN_DEEP_RING")) a.ring = std::min(std::min(kMaxRing, a.nstages), std::max(std::min(a.nsets, a.nstages), atoi(e)));   // ring >= sets: a set must never be two ring uses ahead of the MMA warp (mbarrier parity)
    plan->smemBytes = (size_t)a.ring * kAStageBytes + (size_t)a.ring * a.NT * kKC * 2 + ((size_t)a.nInPlanes + 2 * (a.NT / 4) + 64) * 4 + 8 + (2 * kMaxRing + 1) * 8 + 16;
    plan->wideSmemBytes = plan->smemBytes + (size_t)a.ring * (256 - a.NT) * kKC * 2 + 2 * ((256 - a.NT) / 4) * 4;
    static size_t maxSmem[64] = {0};
    static std::mutex smemLock;                  // ops may be created from one thread per context
    std::lock_guard<std::mutex> guard(smemLock);
    size_t &cur = maxSmem[op->ctx->device & 63];
    if (plan->wideOff && plan->wideSmemBytes > cur) {
        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->wideSmemBytes));
        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->wideSmemBytes));
        cur = plan->wideSmemBytes;
    }
    if (plan->smemBytes > cur) {
        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smemBytes));
        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smemBytes));
        cur = plan->smemBytes;
    }
    return FYN_OK;
}

int fyn_conv_deep_tc_run(fyn_op *op, const fyn_tensor *in, const fyn_tensor *res, fyn_tensor *out, cudaStream_t stream) {
    DeepTcPlan *plan = op->dtc;
    const fyn_conv_desc &d = op->conv;
    if (in->desc.dtype != FYN_F16 || out->desc.dtype != FYN_F16 || (res && res->desc.dtype != FYN_F16)) return 1;   // fp32 storage: direct kernel
    // (a single-plane input is laid out identically as a shallow and as a deep tensor: the stem reads the shallow BN output)
    if ((in->desc.order != FYN_ORDER_DEEP && !plan->args.tapPacked) || out->desc.order != FYN_ORDER_DEEP || in->geom.packing != 4) return 1;
    if (op->epilogue != FYN_EPILOGUE_NONE) return 1;
    DeepTcArgs a = plan->args;
    a.in = fyn_make_view(in);
    a.out = fyn_make_view(out);
    a.res = fyn_make_view(res);
    a.Wo = op->Wo;
    a.Ho = op->Ho;
    a.batch = in->desc.batch;
    a.Mtotal = (long long)a.batch * a.Wo * a.Ho;
    a.inP = d.in_padding;
    a.outP = d.out_padding;
    a.resP = d.res_padding;
    a.act = fyn_act_from_flags(d.flags, d.leaky, d.clip_lo, d.clip_hi);
    a.hasRes = (d.flags & FYN_FLAG_RESIDUAL_INPUT) != 0;
    a.reluRes = (d.flags & FYN_FLAG_RELU_ON_RESIDUAL) != 0;
    a.bnRes = (d.flags & FYN_FLAG_BATCHNORM_ON_RESIDUAL) != 0;
    // the fp16-rounded parameter set of the deep family (fyn_conv.cu: second half of d_bias)
    const int nOut = (d.out_channels + 3) / 4;
    a.bias = op->d_bias + (size_t)nOut * 8;
    a.scale = a.bias + (size_t)nOut * 4;
    a.inNorm = op->innorm ? reinterpret_cast<const float4 *>(op->d_innorm) : nullptr;
    const long long mtiles = (a.Mtotal + kM - 1) / kM;
    static const bool noPdl = getenv("FYN_TC_NO_PDL") != nullptr;   // debugging aid: plain stream-ordered launches
    const char *pe = getenv("FYN_DEEP_PERSIST");
    const int sms = op->ctx->prop.multiProcessorCount;
    // Small grids (batch 1): the split-K cluster kernel -- narrower tiles and the K stages of a tile spread over a cluster, so that
    // a layer of 1 ... 50 tiles still occupies ~100 SMs with <= 5 stages each.  FYN_DEEP_SPLITK: 0 = off, 1 = leave 3x3 stride-1
    // layers to the halo-tile kernel, 2 = every layer (read per run; FYN_DEEP_PERSIST=0 / 2 -- one-tile kernel only / persistent kernel
    // on any grid, the bit-exact pair of the family's tests -- also switch it off)
    const char *ske = getenv("FYN_DEEP_SPLITK");
    const int skMode = ske ? atoi(ske) : 2;
    const bool haloShape = a.K == 3 && a.ds == 1 && !a.inNorm && a.inP >= 1;
    // (regime: fewer tiles than FYN_DEEP_SK_MAXTILES, default half the SM count; the persistent kernels start at two tiles per SM)
    // (a quarter of the SM count, half of it for layers of one or two K stages: measured, see the persistent kernel's threshold below)
    long long skMaxTiles = a.nstages <= 2 ? sms / 2 : sms / 4;
    if (const char *e = getenv("FYN_DEEP_SK_MAXTILES")) skMaxTiles = std::max(1, atoi(e));
    if (skMode && (!pe || atoi(pe) == 1) && !a.tapPacked && a.NT <= 128 && mtiles * plan->ntiles <= skMaxTiles &&
        a.Mtotal < (1ll << 31) - kM && !(skMode == 1 && haloShape)) {
        const int co16 = ((d.out_channels + 15) / 16) * 16;
        int NTs = a.NT % 64 == 0 ? 64 : a.NT, ks = 1;
        long long nsub = 0;
        long long target = sms;                              // CTAs of a layer: at most this many
        if (const char *e = getenv("FYN_DEEP_SK_TARGET")) target = std::max(1, atoi(e));
        // the K split that still fits the target for a given tile width.  Other measured alternatives at batch 1 (ms per forward): 32 columns
        // everywhere 0.316; 64 everywhere 0.318; the K split first and narrower tiles second 0.345; the kernel on mid-size grids of 74 ... 295
        // tiles, where the one-tile kernel runs: batch 8 0.60 -> 0.64 / 0.69 ms, batch 64 1.83 -> 1.89 / 2.00 ms
        static const int minStages = getenv("FYN_DEEP_SK_MINK") ? std::max(1, atoi(getenv("FYN_DEEP_SK_MINK"))) : 1;   // K stages per CTA at least
        auto choose = [&]() {
            nsub = (co16 + NTs - 1) / NTs;
            for (ks = 1; ks * 2 <= 8 && ks * 2 * minStages <= a.nstages && mtiles * nsub * ks * 2 <= target;) ks *= 2;
        };
        // Tile width.  Policy 0: 64 columns, narrowed to 32 when the layer then fills less than half the target.  Policy 1: layers of many
        // pixel tiles (>= 8) whose 32-column grid still fits the target, or whose K is one or two stages, take 32 columns; layers of one or two
        // pixel tiles stay at 64 (fewer, fuller clusters); the others follow policy 0.  Measured on one box (ms per ResNet-50 forward, policy
        // 1 / 0): batch 1 0.309 / 0.327, batch 2 0.399 / 0.392, batch 4 0.508 / 0.469, batch 8 0.730 / 0.609 -- policy 1 keeps the K split
        // from layers with long K on larger grids.  So: policy 1 at batch 1, policy 0 otherwise (FYN_DEEP_SK_POLICY overrides).
        static const int skPolicyEnv = getenv("FYN_DEEP_SK_POLICY") ? atoi(getenv("FYN_DEEP_SK_POLICY")) : -1;
        const int skPolicy = skPolicyEnv >= 0 ? skPolicyEnv : (a.batch == 1 ? 1 : 0);
        const long long nsub32 = (co16 + 31) / 32;
        if (skPolicy == 1 && NTs == 64 && mtiles >= 8 && (mtiles * nsub32 <= target || a.nstages <= 2)) NTs = 32;
        choose();
        if (NTs == 64 && (skPolicy == 0 || mtiles >= 3) && 2 * mtiles * nsub * ks <= target) {
            NTs = 32;
            choose();
        }
        if (const char *e = getenv("FYN_DEEP_SK_NT")) {      // experiments: force the tile width / the split
            const int v = atoi(e);
            if (v >= 16 && v <= a.NT && a.NT % v == 0 && v % 16 == 0) {
                NTs = v;
                choose();
            }
        }
        if (const char *e = getenv("FYN_DEEP_SK_SPLIT")) ks = std::max(1, std::min(std::min(8, a.nstages), atoi(e)));
        auto log2ceil = [](int v) {
            int l = 0;
            while ((1 << l) < v) l++;
            return l;
        };
        if (((a.NT / NTs) & (a.NT / NTs - 1)) != 0) NTs = a.NT;   // sub-tiles per image tile: a power of two (80 outputs: one 80-column tile)
        nsub = (co16 + NTs - 1) / NTs;
        DeepTcArgs k = a;
        k.skNT = NTs;
        k.skSplit = ks;
        k.skPprLog2 = log2ceil(((NTs >> 2) + ks - 1) / ks);
        k.skPpr = 1 << k.skPprLog2;
        k.skSubLog2 = log2ceil(a.NT / NTs);
        k.skPer = a.nstages / ks;
        k.skRem = a.nstages % ks;
        k.idesc = (1u << 4) | ((uint32_t)(NTs >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        // two loader sets (288 threads): two CTAs fit an SM's register file, so the next layer's CTAs are resident and through
        // their set-up while this layer still runs (programmatic dependent launch) -- ResNet-50 batch 1: 0.40 ms with four sets, 0.33 with two
        k.nsets = 2;
        if (const char *e = getenv("FYN_DEEP_SK_SETS")) k.nsets = std::max(1, std::min(kLoadSets, atoi(e)));
        k.skRingLog2 = log2ceil(std::min(kMaxRing, (a.nstages + ks - 1) / ks));
        k.ring = 1 << k.skRingLog2;
        k.invTxIn = 1.0f / (float)k.in.tx;
        k.invHW = 1.0f / (float)(k.Ho * k.Wo);
        k.invWo = 1.0f / (float)k.Wo;
        const size_t nlocMax = (size_t)(a.nstages + ks - 1) / ks;
        const size_t smemK = (size_t)k.ring * (kAStageBytes + (size_t)NTs * kKC * 2) + (size_t)ks * k.skPpr * kM * 16 + (size_t)k.skPpr * 32 +
                             (k.inNorm ? nlocMax * (kKC / 4) * 32 : 0) + ((size_t)a.nInPlanes + 2 * k.skPpr + nlocMax) * 4 + 8 + (2 * kMaxRing + 1) * 8 + 16;
        if (smemK <= (size_t)op->ctx->prop.sharedMemPerBlockOptin) {
            static bool attrSet[64] = {false};
            static std::mutex lockK;
            {
                std::lock_guard<std::mutex> guard(lockK);
                if (!attrSet[op->ctx->device & 63]) {
                    const int optin = (int)op->ctx->prop.sharedMemPerBlockOptin;
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_sk<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                    attrSet[op->ctx->device & 63] = true;
                }
            }
            cudaLaunchAttribute kattr[2];
            int na = 0;
            if (!noPdl) {
                kattr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                kattr[na].val.programmaticStreamSerializationAllowed = 1;
                na++;
            }
            if (ks > 1) {
                kattr[na].id = cudaLaunchAttributeClusterDimension;
                kattr[na].val.clusterDim.x = 1;
                kattr[na].val.clusterDim.y = 1;
                kattr[na].val.clusterDim.z = (unsigned)ks;
                na++;
            }
            cudaLaunchConfig_t kc{};
            kc.gridDim = dim3((unsigned)mtiles, (unsigned)nsub, (unsigned)ks);
            kc.blockDim = dim3((4 * k.nsets + 1) * 32);
            kc.dynamicSmemBytes = smemK;
            kc.stream = stream;
            kc.attrs = kattr;
            kc.numAttrs = (unsigned)na;
#ifdef FYN_SK_TIMELINE
            static int dbgLaunch = 0;
            k.dbgIndex = dbgLaunch++;
#endif
            const int actT = k.act.type <= 1 ? k.act.type : 2;
            if (k.inNorm) {
                if (actT == 0) FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<true, 0>, k));
                else if (actT == 1) FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<true, 1>, k));
                else FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<true, 2>, k));
            } else {
                if (actT == 0) FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<false, 0>, k));
                else if (actT == 1) FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<false, 1>, k));
                else FYN_CUDA(cudaLaunchKernelEx(&kc, k_conv_deep_tc_sk<false, 2>, k));
            }
            op->lastKernel = 13 | (ks << 8) | (NTs << 16);
            FYN_CHECK_LAUNCH(op->ctx);
            return FYN_OK;
        }
    }
    int ntiles = plan->ntiles;
    size_t planSmem = plan->smemBytes;
    // (the 256-column tiles only where the halved grid still runs the persistent kernel: 1x1 1024 -> 2048 on 7x7 at batch 64 is 400 tiles of
    // 128 columns -- persistent -- but 200 of 256 columns, i.e. the one-tile kernel with one 192 KB CTA per SM in two uneven waves;
    // FYN_DEEP_WIDE_MIN: that threshold in tiles per SM, default 2)
    static const long long wideMin = getenv("FYN_DEEP_WIDE_MIN") ? atoll(getenv("FYN_DEEP_WIDE_MIN")) : 2;
    if (plan->wideOff && mtiles * plan->ntiles >= 2ll * op->ctx->prop.multiProcessorCount &&
        (plan->args.nstages <= 2 || mtiles * plan->wideNtiles >= wideMin * op->ctx->prop.multiProcessorCount)) {
        a.NT = 256;
        a.wimg = plan->d_wimg + plan->wideOff;
        a.idesc = (1u << 4) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        ntiles = plan->wideNtiles;
        planSmem = plan->wideSmemBytes;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    // The persistent kernel (one CTA per SM, gathers / MMAs / epilogues of neighbouring tiles overlapped).
    // FYN_DEEP_PERSIST=0 keeps the one-tile-per-CTA kernel (read per run: tests compare the two).
    const char *he = getenv("FYN_DEEP_HALO");          // 0: every layer on the stage-per-tap kernels (read per run)
    // 3x3 stride-1 layers take the halo-tile kernel on ANY grid: on small ones (batch 1) a CTA's chain of K stages is what
    // the layer takes, and one gather per 64 channels instead of nine shortens it ninefold
    const bool halo = a.K == 3 && a.ds == 1 && !a.tapPacked && !a.inNorm && a.inP >= 1 && (!he || atoi(he) != 0) && (!pe || atoi(pe) != 0);
    if (halo) {
        a.ntilesN = ntiles;
        a.totalTiles = mtiles * ntiles;
        {
            DeepTcArgs h = a;
            h.Wv = a.Wo + 1;
            const long long J = (long long)a.batch * (a.Ho + 1) * h.Wv;      // virtual raster rows of the whole batch
            h.Mtotal = J;
            const long long mt3 = (J + kM - 1) / kM;
            h.totalTiles = mt3 * ntiles;
            h.HPp = (kM + 2 * h.Wv + 2 + 7) & ~7;
            const size_t aStage = (size_t)h.HPp * 128, bStage = (size_t)a.NT * kKC * 2;
            const size_t fixedH = (size_t)(a.Cout4 / 4) * 32 + ((size_t)a.nInPlanes + 2 * (size_t)(a.Cout4 / 4)) * 4 + 8 + (4 * kMaxRingP + 4) * 8 + 16;
            const size_t budgetH = (size_t)op->ctx->prop.sharedMemPerBlockOptin - 1024;
            // weight ring: at least four entries (the nine tap groups of a stage stream through it), halo ring: what is left, at most four
            int ringB = 6, ringA = 0;
            for (; ringB >= 3; ringB--) {
                const size_t left = budgetH > fixedH + ringB * bStage ? budgetH - fixedH - ringB * bStage : 0;
                ringA = (int)std::min<size_t>(4, left / aStage);
                if (ringA >= 2) break;
            }
            if (ringB >= 3 && ringA >= 2 && h.HPp <= 2 * kM && h.totalTiles < (1ll << 31) - 2 * sms && J < (1ll << 31) - 4 * kM) {
                h.ring = ringA;
                h.ringB = ringB;
                h.nsets = std::min(ringA, a.NT >= 256 ? 2 : 3);
                if (const char *e = getenv("FYN_DEEP_SETS")) h.nsets = std::max(1, std::min(std::min(4, ringA), atoi(e)));
                h.epiWarps = 24 - 4 * h.nsets;
                if (h.nsets < 2) h.epiWarps = 16;
                const size_t smemH = ringA * aStage + ringB * bStage + fixedH;
                static size_t maxSmemH[64] = {0};
                static std::mutex lockH;
                {
                    std::lock_guard<std::mutex> guard(lockH);
                    size_t &cur = maxSmemH[op->ctx->device & 63];
                    if (smemH > cur) {
                        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_h3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemH));
                        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_h3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemH));
                        FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_h3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemH));
                        cur = smemH;
                    }
                }
                cudaLaunchConfig_t hc{};
                hc.gridDim = dim3((unsigned)std::min<long long>(h.totalTiles, sms));
                hc.blockDim = dim3(kThreadsDeepH);
                hc.dynamicSmemBytes = smemH;
                hc.stream = stream;
                hc.attrs = attr;
                hc.numAttrs = noPdl ? 0 : 1;
                const int actT = h.act.type <= 1 ? h.act.type : 2;
                if (actT == 0) FYN_CUDA(cudaLaunchKernelEx(&hc, k_conv_deep_tc_h3<0>, h));
                else if (actT == 1) FYN_CUDA(cudaLaunchKernelEx(&hc, k_conv_deep_tc_h3<1>, h));
                else FYN_CUDA(cudaLaunchKernelEx(&hc, k_conv_deep_tc_h3<2>, h));
                op->lastKernel = 12;
                FYN_CHECK_LAUNCH(op->ctx);
                return FYN_OK;
            }
        }
    }
    // The persistent kernel on EVERY grid the split-K kernel does not take (FYN_DEEP_PERSIST_MIN: the smallest grid, in tiles, that takes it).
    // Round 2 first kept it for grids of at least two tiles per SM and left 75 ... 295 tiles to the one-tile kernel -- one 128 - 192 KB CTA
    // per SM in uneven waves, no overlap of gather / MMA / epilogue.  Measured (ms per ResNet-50 forward, one-tile kernel below 296 tiles /
    // persistent everywhere, split-K regime 74 tiles / 37 tiles + 74 for one- and two-stage layers): batch 8 0.607 / 0.560, batch 16 0.80 / 0.731,
    // batch 32 1.242 / 1.030, batch 64 1.793 / 1.627, batch 128 3.023 / 2.937; batch 1 ... 4 unchanged or slightly better.
    static const long long persistMinEnv = getenv("FYN_DEEP_PERSIST_MIN") ? atoll(getenv("FYN_DEEP_PERSIST_MIN")) : -1;
    const long long persistMin = persistMinEnv >= 0 ? persistMinEnv : 1;
    if ((!pe || atoi(pe) != 0) && mtiles * ntiles >= (pe && atoi(pe) == 2 ? 1ll : persistMin)) {
        a.ntilesN = ntiles;
        a.totalTiles = mtiles * ntiles;
        const size_t stageBytes = (size_t)kAStageBytes + (size_t)a.NT * kKC * 2;
        const size_t fixed = (size_t)(a.Cout4 / 4) * 32 + (a.inNorm ? (size_t)a.nInPlanes * 32 : 0) + ((size_t)a.nInPlanes + 2 * (size_t)(a.Cout4 / 4) + 128 + a.nstages) * 4 + 8 +
                             (2 * kMaxRingP + 4) * 8 + 16;
        const size_t budget = (size_t)op->ctx->prop.sharedMemPerBlockOptin - 1024;
        int ring = (int)std::min<size_t>(kMaxRingP, fixed < budget ? (budget - fixed) / stageBytes : 0);
        if (const char *e = getenv("FYN_DEEP_PRING")) ring = std::max(1, std::min(ring, atoi(e)));
        if (ring >= 2 && a.totalTiles < (1ll << 31) - 2 * sms && a.Mtotal < (1ll << 31) - 2 * kM) {
            a.ring = ring;
            // 24 worker warps split between gathering and draining: a tile costs the loaders nstages gathers of 16 KB and the
            // epilogue NT / 16 column groups of 4 texels per pixel (plus as many residual texels)
            const double loadWork = (double)a.nstages, epiWork = (a.NT / 16) * (a.hasRes ? 1.5 : 1.0) * 0.5;
            a.nsets = epiWork > 2.0 * loadWork ? 2 : (epiWork > loadWork ? 3 : 4);
            if (const char *e = getenv("FYN_DEEP_SETS")) a.nsets = std::max(1, std::min(5, atoi(e)));
            a.nsets = std::min(a.nsets, ring);
            a.epiWarps = 24 - 4 * a.nsets;
            if (a.nsets < 2) a.epiWarps = 16;   // (a TMEM lane quarter is drained by at most four warps here)
            const size_t smemP = (size_t)ring * stageBytes + fixed;
            static size_t maxSmemP[64] = {0};
            static std::mutex lockP;
            {
                std::lock_guard<std::mutex> guard(lockP);
                size_t &cur = maxSmemP[op->ctx->device & 63];
                if (smemP > cur) {
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    FYN_CUDA(cudaFuncSetAttribute(k_conv_deep_tc_p<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
                    cur = smemP;
                }
            }
            cudaLaunchConfig_t pc{};
            pc.gridDim = dim3((unsigned)std::min<long long>(a.totalTiles, sms));
            pc.blockDim = dim3(kThreadsDeepP);
            pc.dynamicSmemBytes = smemP;
            pc.stream = stream;
            pc.attrs = attr;
            pc.numAttrs = noPdl ? 0 : 1;
            const int actT = a.act.type <= 1 ? a.act.type : 2;
            if (a.inNorm) {
                if (actT == 0) FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<true, 0>, a));
                else if (actT == 1) FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<true, 1>, a));
                else FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<true, 2>, a));
            } else {
                if (actT == 0) FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<false, 0>, a));
                else if (actT == 1) FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<false, 1>, a));
                else FYN_CUDA(cudaLaunchKernelEx(&pc, k_conv_deep_tc_p<false, 2>, a));
            }
            op->lastKernel = 11;
            FYN_CHECK_LAUNCH(op->ctx);
            return FYN_OK;
        }
    }
    dim3 grid((unsigned)mtiles, (unsigned)ntiles);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((4 * a.nsets + 1) * 32);
    size_t smemBytes = planSmem;
    // Large grids of multi-stage layers: three sets and a ring of three (96 KB) let two CTAs share an SM, so that one CTA's
    // set-up and epilogue overlap the other's gathers; small grids keep four sets (more gathers in flight per CTA).
    static const int altMode = getenv("FYN_DEEP_ALT") ? atoi(getenv("FYN_DEEP_ALT")) : 1;
    // (only where the ring of four keeps a second CTA out: narrow tiles, e.g. the 64-column stem, already fit twice)
    if (altMode && a.NT <= 128 && planSmem > 113 * 1024 && a.nstages >= 4 && a.nsets == kLoadSets && a.ring == kMaxRing &&
        mtiles * ntiles >= 2ll * op->ctx->prop.multiProcessorCount) {
        a.nsets = 3;
        a.ring = 3;
        smemBytes -= (size_t)(kMaxRing - 3) * (kAStageBytes + (size_t)a.NT * kKC * 2);
        cfg.blockDim = dim3((4 * a.nsets + 1) * 32);
    }
    cfg.dynamicSmemBytes = smemBytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = noPdl ? 0 : 1;
    if (a.inNorm) FYN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_deep_tc<true>, a));
    else FYN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_deep_tc<false>, a));
    op->lastKernel = 10;
    FYN_CHECK_LAUNCH(op->ctx);
    return FYN_OK;
}

#ifdef FYN_SK_TIMELINE
extern "C" int fyn_debug_sk_timeline(unsigned long long *dst, int reset) {
    if (dst) cudaMemcpyFromSymbol(dst, g_skTimeline, sizeof(unsigned long long) * 2048 * 16);
    if (reset) {
        static unsigned long long init[2048][16];
        for (int i = 0; i < 2048; i++) {
            for (int j = 0; j < 16; j++) init[i][j] = 0;
            init[i][0] = init[i][9] = ~0ull;
        }
        cudaMemcpyToSymbol(g_skTimeline, init, sizeof(init));
    }
    return 0;
}
#endif

void fyn_conv_deep_tc_destroy(fyn_op *op) {
    if (!op->dtc) return;
    if (op->dtc->d_wimg) cudaFree(op->dtc->d_wimg);
    delete op->dtc;
    op->dtc = nullptr;
}
