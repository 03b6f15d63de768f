// K5/K7/K9/K10 — Linear layers on the 5th-generation tensor cores.
//   C[M][N] = A[M][K] (fp16, K-major) x W[N][K]^T (fp16, nn.Linear layout is already K-major)
//   fp32 accumulation in TMEM; epilogue: + bias, ReLU, + addend (residual / positional table),
//   stores fp32 and/or fp16 (the fp16 copy is the A operand of the next GEMM).
// Follows Full_model/SubLayers.py:39-57,78-82 and Full_model/Models.py:124-130,411-425.
//
// One 128 x BN output tile per CTA.  Warp 0: TMA producer (SWIZZLE_128B boxes of 64 fp16 = 128 B rows),
// warp 1: TMEM allocator + single-thread tcgen05.mma issuer, warps 2-5: epilogue (tcgen05.ld, one accumulator
// row per thread).  kStages-deep smem ring with full/empty mbarriers; K tails and M/N tails come from TMA
// zero fill.  Two CTAs fit per SM (96 KB smem, 128 TMEM columns each), so one CTA's epilogue overlaps the
// other's MMA stream.
#include "egx_common.cuh"
#include "tc_common.cuh"

namespace egx {

namespace {

using namespace tc;

constexpr int GM = 128;          // tile rows (UMMA M)
constexpr int GK = 64;           // fp16 elements per K block = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kThreads = 192;

template <int BN>
struct GemmSmem {
    static constexpr int kABytes = GM * GK * 2;
    static constexpr int kBBytes = BN * GK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + 128 + 1024;   // barriers + alignment slack
};

struct GemmTcEpi {
    const float* bias;
    const float* addend;
    int addend_rows, addend_ld;
    int relu;
    float* out32; int ld32;
    __half* out16; int ld16;
};

template <int BN>
__global__ void __launch_bounds__(kThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
               int K, GemmTcEpi ep) {
    using S = GemmSmem<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty = full + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * BN;
    const int num_kb = (K + GK - 1) / GK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                unsigned char* a = smem + st * S::kStageBytes;
                mbar_expect_tx(&full[st], S::kStageBytes);
                tma_load_2d(a, &tmA, &full[st], kb * GK, m0);
                tma_load_2d(a + S::kABytes, &tmB, &full[st], kb * GK, n0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(GM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a = smem_u32(smem + st * S::kStageBytes);
                const uint32_t b = a + S::kABytes;
#pragma unroll
                for (int k = 0; k < GK / 16; ++k)
                    umma_f16(tmem_base, make_smem_desc<128>(a + k * 32), make_smem_desc<128>(b + k * 32), idesc,
                             (kb | k) != 0);
                umma_commit(&empty[st]);
            }
            umma_commit(tmem_full);
        }
    } else {
        // epilogue: warp w owns TMEM lanes [32*(w%4), +32)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const float* add_row = nullptr;
        if (ep.addend && m < M) add_row = ep.addend + (size_t)(ep.addend_rows ? m % ep.addend_rows : m) * ep.addend_ld;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
            const int nb = n0 + c * 32;
            if (m < M && nb < N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = nb + j;
                float t = v[j];
                if (n < N) {
                    if (ep.bias) t += __ldg(ep.bias + n);
                    if (ep.relu) t = fmaxf(t, 0.f);
                    if (add_row) t += __ldg(add_row + n);
                }
                v[j] = t;
            }
            const bool full_chunk = nb + 32 <= N;
            if (ep.out32) {
                float* o = ep.out32 + (size_t)m * ep.ld32 + nb;
                if (full_chunk && (ep.ld32 & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    for (int j = 0; j < 32 && nb + j < N; ++j) o[j] = v[j];
                }
            }
            if (ep.out16) {
                __half* o = ep.out16 + (size_t)m * ep.ld16 + nb;
                if (full_chunk && (ep.ld16 & 7) == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u;
                        *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[8 * j], v[8 * j + 1]);
                        *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                        *reinterpret_cast<__half2*>(&u.z) = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
                        *reinterpret_cast<__half2*>(&u.w) = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                        reinterpret_cast<uint4*>(o)[j] = u;
                    }
                } else {
                    for (int j = 0; j < 32 && nb + j < N; ++j) o[j] = __float2half_rn(v[j]);
                }
            }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

__global__ void cvt_pad_kernel(const float* __restrict__ in, int64_t rows, int cols, int ld_in,
                               __half* __restrict__ out, int ld_out) {
    const int64_t total = rows * ld_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld_out;
        const int c = int(i % ld_out);
        out[i] = c < cols ? __float2half_rn(in[r * ld_in + c]) : __half(0.f);
    }
}

}  // namespace

// per-device one-time setup (opt-in shared memory size)
int gemm_tc_init_device() {
    return cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmSmem<128>::kTotal) == cudaSuccess ? 0 : -1;
}

// fp32 [rows][cols] (pitch ld_in) -> fp16 [rows][ld_out], columns >= cols zeroed
int launch_cvt_pad_f16(const float* in, int64_t rows, int cols, int ld_in, __half* out, int ld_out, cudaStream_t s) {
    const int64_t total = rows * ld_out;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    cvt_pad_kernel<<<grid, 256, 0, s>>>(in, rows, cols, ld_in, out, ld_out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// A: [M][K] fp16 with row pitch lda (elements, multiple of 8); W: [N][K] fp16 with row pitch ldw.
int launch_gemm_tc(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const GemmEpi& e,
                   float* out32, int ld32, __half* out16, int ld16, cudaStream_t s) {
    constexpr int BN = 128;
    CUtensorMap ta, tb;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, dB[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t sA[1] = {(uint64_t)lda * 2}, sB[1] = {(uint64_t)ldw * 2};
    const uint32_t bA[2] = {GK, GM}, bB[2] = {GK, BN};
    if (!make_tmap_f16(&ta, A, 2, dA, sA, bA, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_tmap_f16(&tb, W, 2, dB, sB, bB, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    GemmTcEpi ep{e.bias, e.addend, e.addend_rows, e.addend_ld, e.relu, out32, ld32, out16, ld16};
    dim3 grid((M + GM - 1) / GM, (N + BN - 1) / BN);
    gemm_tc_kernel<BN><<<grid, kThreads, GemmSmem<BN>::kTotal, s>>>(ta, tb, M, N, K, ep);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
