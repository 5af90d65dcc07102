// Measures the int8 tensor-pipe rate of tcgen05.mma.kind::i8 on this GPU (the roofline denominator of
// csrc/ozaki_syrk.cu): every SM issues back-to-back 128xNx32 MMAs on operands resident in shared memory,
// no loads in the timed region.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/int8_peak tools/int8_peak.cu
// peak2_kernel: the same loop issued as cta_group::2 (M = 256 over a two-CTA cluster) for N = 128 and 256.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + N) * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy smem writes -> async proxy (MMA reads)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tbase;
    // K-major, 32-byte swizzle, 8-row groups 256 B apart (the layout of csrc/ozaki_syrk.cu)
    const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);
    const uint64_t da = ((uint64_t)hi << 32) | (smem_u32(smem) >> 4);
    const uint64_t db = ((uint64_t)hi << 32) | ((smem_u32(smem) + 128 * 32) >> 4);
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N / 8) << 17) | ((128u / 16) << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t acc = t + ((u * N) & 511);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     ::"r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}

template <int N>
static void run(int sms, int iters) {
    unsigned long long *cyc;
    cudaMalloc(&cyc, sms * sizeof(*cyc));
    const int smem = (128 + N) * 32 + 1024;
    cudaFuncSetAttribute(peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<N><<<sms, 128, smem>>>(iters / 10, cyc);
    cudaEventRecord(e0);
    peak_kernel<N><<<sms, 128, smem>>>(iters, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(*cyc), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int i = 0; i < sms; ++i) if (h[i] > mx) mx = h[i];
    const double ops = 2.0 * 128 * N * 32 * 8.0 * iters * sms;
    printf("128x%dx32 int8: %.3f ms, %.1f TOP/s (events), %.1f cycles per MMA (clock64, slowest SM)\n", N, ms,
           ops / (ms * 1e-3) * 1e-12, (double)mx / (8.0 * iters));
    cudaFree(cyc);
}


// ---- two-SM form: cluster of 2, leader issues 256xNx32 MMAs; each CTA holds its 128 rows of A and N/2 rows of B ----
template <int N>
__global__ void __launch_bounds__(128, 1) peak2_kernel(int iters, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (128 + N / 2) * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t t = tbase;
    const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);
    const uint64_t da = ((uint64_t)hi << 32) | (smem_u32(smem) >> 4);
    const uint64_t db = ((uint64_t)hi << 32) | ((smem_u32(smem) + 128 * 32) >> 4);
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N / 8) << 17) | ((256u / 16) << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        t0 = clock64();
        if (rank == 0 && elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t acc = t + ((u * N) & 511);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     ::"r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}

template <int N>
static void run2(int sms, int iters) {
    sms &= ~1;
    unsigned long long *cyc;
    cudaMalloc(&cyc, sms * sizeof(*cyc));
    const int smem = (128 + N / 2) * 32 + 1024;
    cudaFuncSetAttribute(peak2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, peak2_kernel<N>, iters / 10, cyc);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, peak2_kernel<N>, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("2SM N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(*cyc), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int i = 0; i < sms; ++i) if (h[i] > mx) mx = h[i];
    const double ops = 2.0 * 256 * N * 32 * 8.0 * iters * (sms / 2);
    printf("cta_group::2 256x%dx32 int8: %.3f ms, %.1f TOP/s (events), %.1f cycles per MMA (clock64, slowest SM)\n", N, ms,
           ops / (ms * 1e-3) * 1e-12, (double)mx / (8.0 * iters));
    cudaFree(cyc);
}

int main(int argc, char **argv) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    run<128>(sms, iters);
    run<256>(sms, iters);
    run2<128>(sms, iters);
    run2<256>(sms, iters);
    return 0;
}
