// Micro-benchmark (not part of the product): cycles per tcgen05.mma kind::f16 with both operands in shared memory,
// cta_group::1 (M = 128) against cta_group::2 (M = 256 over a CTA pair, B split between the two CTAs), N = 32..256.
// Answers one design question of DESIGN.md section 6: is the N <= 64 convolution MMA bound by shared-memory operand
// reads, and does pairing CTAs relieve it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate profiles/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int swz) {
    const uint64_t layout = swz == 128 ? 2 : 4;
    const uint64_t sbo = (8 * swz) >> 4;
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (sbo << 32) | (uint64_t(1) << 46) | (layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        if (spins > (1u << 22)) __trap();
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int swz, int iters, int taps, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_ptr;
    // four issuing threads (one per warp), each on its own accumulator: the issue cost of one thread (~45-80 cycles
    // with descriptor arithmetic) must not hide the drain rate of the tensor pipe
    if ((threadIdx.x & 31) == 0 && rank == 0) {
        const uint32_t w = threadIdx.x >> 5;
        const uint32_t idesc = make_idesc(CG == 2 ? 256 : 128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 48 * 1024;
        const uint64_t da0 = make_desc(a0, swz), db0 = make_desc(b0, swz);
        const uint32_t d = tm + (N <= 128 ? w * 128 : (w & 1) * 256);
        const uint64_t step = (uint64_t)((swz * 3) >> 4);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const uint64_t da = da0 + t * step, db = db0 + t * 64;
                if (CG == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
            }
        }
        if (CG == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && w == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (threadIdx.x < 32) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
    }
}

template <int CG>
static double run(int N, int swz, int grid) {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 2000, taps = 9;
    const int smem = 97 * 1024;
    cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, swz, iters, taps, d);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return -1; }
    }
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (double)c / (4.0 * iters * taps);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d; cycles per MMA (K = 16), all SMs busy\n", sms);
    const int swzs[2] = {64, 128};
    for (int s = 0; s < 2; ++s)
        for (int N = 32; N <= 256; N *= 2) {
            const double c1 = run<1>(N, swzs[s], sms);
            const double c2 = run<2>(N, swzs[s], sms & ~1);
            printf("swizzle %3d  N %3d   cta_group::1 M=128: %6.1f cyc (math floor %4d)   cta_group::2 M=256: %6.1f cyc per pair-MMA = %6.1f per 128 rows\n",
                   swzs[s], N, c1, 128 * N / 256, c2, c2);
        }
    return 0;
}
