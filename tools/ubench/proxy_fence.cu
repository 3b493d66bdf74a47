// proxy_fence.cu -- cost of fence.proxy.async.shared::cta after generic-proxy shared-memory stores (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o proxy_fence proxy_fence.cu && ./proxy_fence
// Variants: stores only / stores + fence / stores + fence while other warps keep cp.async copies in flight /
// fence executed by threads that themselves have cp.async copies in flight.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(const uint4 *src, long long *out, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4 *buf = reinterpret_cast<uint4 *>(smem);
    const int t = threadIdx.x;
    const bool worker = t < 128;
    long long t0 = 0, acc = 0;
    if (MODE >= 4) {
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 90 * 1024);
        if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(128));
    }
    __syncthreads();
    if (worker) {
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (MODE == 3) {   // own copies in flight
                for (int u = 0; u < 4; u++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + 4096 + u * 128 + t)), "l"(src + ((i * 4 + u) * 128 + t) % (1 << 20)) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
#pragma unroll
            for (int u = 0; u < 5; u++) buf[u * 128 + t] = make_uint4(i, u, t, 0);
            if (MODE >= 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        acc = clock64() - t0;
        if (MODE == 3) asm volatile("cp.async.wait_all;" ::: "memory");
        if (t == 0) out[0] = acc;
        if (MODE >= 4) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(smem + 90 * 1024)) : "memory");
    } else if (MODE == 4 || MODE == 5) {
        // other warps block on an mbarrier that only completes when the workers are done (MODE 5: pure spin)
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 90 * 1024);
        unsigned ok = 0;
        while (!ok) {
            if (MODE == 4)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(0), "r"(0x989680u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
        }
    } else if (MODE == 2) {
        // background traffic from other warps: asynchronous copies global -> shared
        for (int i = 0; i < iters * 4; i++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + 4096 + (t - 128))), "l"(src + (i * 128 + t) % (1 << 20)) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if ((i & 7) == 7) asm volatile("cp.async.wait_group 4;" ::: "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
}

int main() {
    uint4 *src;
    long long *out, h;
    cudaMalloc(&src, (size_t)(1 << 20) * 16);
    cudaMemset(src, 1, (size_t)(1 << 20) * 16);
    cudaMalloc(&out, 8);
    const int iters = 2000, smem = 96 * 1024;
    auto run = [&](auto kern, const char *name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<1, 544, smem>>>(src, out, iters);
        kern<<<1, 544, smem>>>(src, out, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("%-58s %8.1f cycles / iteration (%s)\n", name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    };
    run(k<0>, "5 x STS.128 per thread");
    run(k<1>, "5 x STS.128 + fence.proxy.async");
    run(k<2>, "same, other warps keep cp.async in flight");
    run(k<3>, "same, the fencing threads have cp.async in flight");
    run(k<4>, "STS + fence, 13 warps blocked in mbarrier.try_wait");
    run(k<5>, "STS + fence, 13 warps spinning on mbarrier.test_wait");
    return 0;
}
