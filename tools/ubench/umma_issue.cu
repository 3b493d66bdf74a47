// Micro-benchmark: issue rate of tcgen05.mma (M=128, K=16, kind::f16) for several N, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue umma_issue.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int UNROLL>
__global__ void k(int n, int iters, int variant, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tbase;
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tbase;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t base = smem_u32(smem) >> 4;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                // variant 0: same descriptors; variant 1: different A address per MMA; 2: alternate accumulators
                const uint32_t alo = (base + (variant >= 1 ? (uint32_t)(u * 8) : 0u)) | (130u << 16);
                const uint32_t blo = (base + 1024u) | ((uint32_t)n << 16);
                const uint64_t ad = ((uint64_t)hi << 32) | alo, bd = ((uint64_t)hi << 32) | blo;
                const uint32_t d = tm + (variant == 2 ? (uint32_t)((u & 3) * 64) : 0u);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred P1;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 64, U = 16;
    for (int variant = 0; variant < 3; variant++)
        for (int n : {16, 32, 48, 64, 128, 256}) {
            if (variant == 2 && n > 64) continue;
            for (int rep = 0; rep < 2; rep++) {
                k<16><<<1, 128, 64 * 1024>>>(n, iters, variant, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("variant %d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", variant, n, (double)h[0] / (iters * U),
                   (double)h[1] / (iters * U), 128 * n / 256);
        }
    return 0;
}
