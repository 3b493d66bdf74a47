// fyn_tc_common.cuh -- PTX wrappers, step table and plan structures shared by the tcgen05 shallow convolution kernels
// (fyn_conv_tc.cu: one layer per launch; fyn_conv_chain.cu: a persistent kernel that walks a chain of layers).
#pragma once

#include "fyn_internal.h"

namespace {

#ifndef FYN_TC_WAIT
#define FYN_TC_WAIT 0   // mbarrier wait flavour: 0 = try_wait with suspend hint, 1 = test_wait polling, 2 = try_wait without hint
#endif
#ifdef FYN_TC_PROFILE
__device__ long long g_trace[4][64][2];   // [role: 0 loader publish, 1 mma issue begin/end, 2 epilogue begin/end][index][begin, end]
#define TRACE(role, idx, which) do { if (blockIdx.x == 0 && (idx) < 64) g_trace[role][idx][which] = clock64() - pK0; } while (0)
#define PROF_DECL(n) long long n = 0
#define PROF_T() clock64()
#define PROF_ADD(acc, t0) acc += clock64() - (t0)
#else
#define PROF_DECL(n)
#define PROF_T() 0
#define TRACE(role, idx, which)
#define PROF_ADD(acc, t0)
#endif

constexpr int kMaxSteps = 96;
constexpr int kWorkWarps = 16;                       // epilogue + loader warps; the split (8 + 8 or 12 + 4) is part of the plan
constexpr int kMmaWarp = kWorkWarps;                 // first of the two MMA warps (even / odd jobs)
constexpr int kMmaWarps = 2;
constexpr int kThreads = (kMmaWarp + kMmaWarps) * 32;   // 576
constexpr int kFinBatch = 4;                         // items a finishing lane keeps in flight
constexpr int kMaxItems = 1152;                      // (pixel, chunk) items / pixels of one input row
constexpr int kMaxStages = 8;                        // staged input rows in flight per CTA (as many as shared memory allows)
constexpr int kTileM = 128;
constexpr int kMaxRows = 11;

// one tcgen05.mma (M=128, N, K=16), pre-encoded for the issuing warp (16 bytes: one constant-bank load)
struct __align__(16) TcStep {
    uint32_t a_lo;       // (LBO >> 4) << 16 | ((window row * slot bytes + chunk offset) >> 4): add the window base >> 4
    uint32_t b_off16;    // byte offset inside the weight image >> 4
    uint32_t accumulate; // 0 = overwrite the accumulator (first step of a job), 1 = accumulate
    uint32_t pad;
};

struct TcArgs {
    TView in, out, res;
    const uint4 *wimg;
    uint32_t wbytes, idesc, b_lbo;
    int nsteps;              // MMA steps per job
    int stageBytes;          // bytes of one staged input row (all planes)
    int nstages;             // staged rows in flight (multiple of finGroups)
    int finGroups;           // loader groups working on different rows; nslots and nstages are multiples of it
    int epiWarps;            // 8 or 12 epilogue warps (the remaining work warps are loaders)
    int nitems;              // entries of the row item table
    int rowAdvance;          // input rows the window moves per job (stride; 1 for fractional)
    int dyMin, nrows;        // window: input rows [rowAdvance*i + dyMin, +nrows)
    TcStep steps[kMaxSteps];
    int opx, opy;            // output phases stacked along N (output pixel = (opx*j + fx, opy*i + fy))
    int planesPerPhase;      // 4-channel planes per phase (Cq / 4)
    int Wo, Ho;              // output net size
    int Hj, Wj;              // job-space size: ceil(Ho / opy) rows, ceil(Wo / opx) columns
    int inP, outP, resP;
    int nchunks, rowpx;      // mode 0: chunks per version and pixels per chunk row; mode 1: rowpx = chunks per slot
    int nver, verBytes;      // slot versions: 0 = activated (or the only one), 1 = raw
    int nslots, slotBytes;   // logical ring slots; the first nmirror slots are mirrored behind the ring
    int nmirror;             // windows start at multiples of rowAdvance: nrows - rowAdvance mirror slots when that divides nslots, else nrows - 1
    int SH, nxs;             // job rows per strip, column blocks
    int N, nInPlanes;
    int mode;                // 0 = plane-pair chunks (fp16 RGBA planes), 1 = pixel-pair chunks (single plane, 4 px per GEMM row)
    int ds;                  // mode 0: horizontal stride of the slot layout (2 = even/odd split)
    int x_lead;              // input pixels to the left of job column 0 held in a slot
    ActParams act;
    int hasRes, reluRes, bnRes;
    int epilogue;            // FYN_EPILOGUE_*: element-wise function fused behind the convolution
    int biasFolded;          // 1: the bias enters the accumulator as one more MMA step (A = ones region behind the weight image)
    uint32_t biasB16, onesOff;   // weight-image offsets: bias step (>> 4) and the ones region (bytes)
    uint32_t epiOff;             // weight-image offset of the epilogue parameters ([16] bias float4, [16] scale float4)
    int debug;               // FYN_TC_DEBUG bits (timing ablations only): 1 = epilogue without global stores
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
// Blocking wait with a large suspend-time hint: the warp sleeps in hardware until the phase completes instead of
// polling (12 polling warps otherwise compete with the tensor core for shared-memory bandwidth).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#if FYN_TC_WAIT == 1
    // experiment: poll (no hardware suspend)
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#elif FYN_TC_WAIT == 2
    // experiment: try_wait without a suspend-time hint
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
#endif
}
// weights: one bulk copy (TMA unit, async proxy) that completes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory accesses through 32-bit shared-window addresses (keeps the loaders' address arithmetic in 32 bits)
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// programmatic dependent launch: everything before grid_dep_wait() overlaps the tail of the previous kernel
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// one deterministic leader lane of a converged warp
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptors (K-major, SWIZZLE_NONE "interleave"): core matrix = 8 rows x 16 bytes stored
// contiguously (rows 16 B apart).  Low word: start address >> 4 in bits [0,14), LBO >> 4 in bits [16,30) (distance
// between the two 16-byte K chunks of one K=16 instruction).  High word: SBO >> 4 in bits [0,14) (distance between
// 8-row groups, 128 B here) and the sm_100 descriptor version 1 in bits [14,16).

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint2 relu_h4(uint2 v) {
    const __half2 z = __float2half2_rn(0.f);
    __half2 *q = reinterpret_cast<__half2 *>(&v);
    q[0] = __hmax2(q[0], z);
    q[1] = __hmax2(q[1], z);
    return v;
}

__device__ __forceinline__ uint4 act_h8(uint4 v, const ActParams &a);

__device__ __forceinline__ uint2 act_h4(uint2 v, const ActParams &a) {
    if (a.type == 1) return relu_h4(v);
    if (a.type == 0) return v;
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
    return make_uint2(pack_half2(fyn_act(f0.x, a), fyn_act(f0.y, a)), pack_half2(fyn_act(f1.x, a), fyn_act(f1.y, a)));
}

__device__ __forceinline__ uint4 act_h8(uint4 v, const ActParams &a) {
    const uint2 l = act_h4(make_uint2(v.x, v.y), a), h = act_h4(make_uint2(v.z, v.w), a);
    return make_uint4(l.x, l.y, h.x, h.y);
}

// compile-time activation selection (ACT: 0 none, 1 ReLU, 2 leaky / clip through fp32): the kernel is instantiated per
// activation and residual flavour so that every role's loop stays small (the instruction caches are 6 KB / 32 KB)
template <int ACT>
__device__ __forceinline__ uint4 act_h8_t(uint4 v, const ActParams &a) {
    if (ACT == 0) return v;
    if (ACT == 1) {
        const uint2 l = relu_h4(make_uint2(v.x, v.y)), h = relu_h4(make_uint2(v.z, v.w));
        return make_uint4(l.x, l.y, h.x, h.y);
    }
    return act_h8(v, a);
}
template <int ACT>
__device__ __forceinline__ float4 act_f4_t(float4 v, const ActParams &a) {
    if (ACT == 0) return v;
    if (ACT == 1) return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    return fyn_act4(v, a);
}

// position in the ring without divisions
struct RingPos {
    int slot, fill;
    __device__ __forceinline__ void advance(int d, int nslots) {
        slot += d;
        while (slot >= nslots) {
            slot -= nslots;
            fill++;
        }
    }
};

}  // namespace

// host-side plan of one tcgen05 convolution (fyn_conv_tc.cu builds it; fyn_conv_chain.cu chains several of them)
struct ConvTcPlan {
    TcArgs args{};
    uint4 *d_wimg = nullptr;
    size_t wimgBytes = 0;
    size_t smemBytes = 0;
    int mode = 0;
    // single-row plan of a layer that runs row-stacked on its own (for the chain kernel, which does not stack rows)
    bool hasRow1 = false;
    TcArgs row1{};
    uint4 *d_wimg1 = nullptr;
    size_t wimg1Bytes = 0;
    // the single-row image with the kernel rows reversed (layers of a chain that sweep bottom -> top); only for layers that
    // can be part of a chain
    uint4 *d_wimgFlip = nullptr;
    size_t wimgFlipBytes = 0;
    const TcArgs &chainArgs() const { return hasRow1 ? row1 : args; }
    const uint4 *chainImage(bool flipped) const { return flipped ? d_wimgFlip : (hasRow1 ? d_wimg1 : d_wimg); }
};
