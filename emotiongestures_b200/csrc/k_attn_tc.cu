// K8 — short-sequence multi-head attention on tcgen05 (Full_model/Modules.py:13-23 inside
// Full_model/SubLayers.py:43-52): softmax((q / sqrt(d_k)) k^T) v per (clip, head), mask None.
//
// The sequences are 34 (TED) or 60 (BEAT) tokens, far below a 128-row MMA tile, so G = floor(128 / L) clips
// are packed into one tile (3 x 34 = 102 rows, 2 x 60 = 120 rows) and the score matrix is block diagonal.
// Each clip's keys start at a multiple of 16 columns (LP = L rounded up to 16; 48 / 64), so every clip sees
// the same K-step boundaries in the second MMA whatever its slot in the group: a clip's result is
// bit-identical for any batch size, sharding or GPU count.
//   S[128 x NK] = Q_tile[128 x 64] K_tile[NK x 64]^T        (UMMA, both operands K-major)
//   softmax over the L columns of the row's own clip, other columns written as exact zeros
//   O[128 x 64] = P[128 x NK] V_tile[NK x 64]               (UMMA, A = P K-major from smem, B = V MN-major)
// Q, K, V head slices come straight out of the projection GEMM's [rows][3*H*d] (or [rows][H*d] / [rows][2*H*d])
// fp16 output through 2-D TMA boxes {64, G*L}; P is written by the softmax warps into shared memory in the
// SWIZZLE_128B K-major layout the A descriptor expects; 1/sum is applied to O in the epilogue (P stays
// unnormalised in [0,1], which keeps fp16 P accurate).  q/sqrt(d_k) is applied as an exact power-of-two
// scale on the fp32 scores (d_k = 64).
//
// One (clip group, head) tile per iteration, persistent CTAs (2 per SM so one CTA's softmax overlaps the
// other's loads and MMAs).  warp 0: TMA, warp 1: MMA issuer, warps 2-5: softmax + output (one row per thread).
#include "egx_common.cuh"
#include "tc_common.cuh"

namespace egx {

namespace {

using namespace tc;

constexpr int kAttnThreads = 192;
constexpr int kTileBytes = 128 * 128;                 // 128 rows x 64 fp16 (Q tile; one P swizzle atom)
constexpr int kMaxNK = 144;                           // key columns per tile (G * LP)
constexpr int kKVBytes = kMaxNK * 128;                // K / V tiles: one 128-byte row per key
constexpr int kPBytes = 3 * kTileBytes;               // P: up to 192 key columns = three 64-wide swizzle atoms
constexpr int kAttnSmem = kTileBytes + 2 * kKVBytes + kPBytes + 256 + 1024;

struct AttnParams {
    int n_clips, L, G, LP, NK;   // tokens per clip, clips per tile, key slots per clip (16-aligned), G * LP
    int n_head, n_groups;
    int q_col0, k_col0, v_col0;  // column of head 0 inside the q / kv source rows
    __half* out; int ldo;
    float scale_log2e;           // (1 / sqrt(d_k)) * log2(e)
};

// MN-major B operand (V: rows = keys, 128-byte rows of 64 d_v values, SWIZZLE_128B): 8-key groups are 1024 B apart
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AttnParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;
    unsigned char* sK = smem + kTileBytes;
    unsigned char* sV = sK + kKVBytes;
    unsigned char* sP = sV + kKVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
    // Q/K and V have separate barriers: Q and K are free again as soon as the score MMAs are done, so the next
    // tile's Q/K loads (the DRAM latency of the chain) run under this tile's softmax, P.V and output
    uint64_t* full_qk = bars;       // Q, K landed
    uint64_t* empty_qk = bars + 1;  // Q, K consumed (score MMAs done)
    uint64_t* s_full = bars + 2;    // scores in TMEM
    uint64_t* p_full = bars + 3;    // P in smem
    uint64_t* o_full = bars + 4;    // output accumulator in TMEM
    uint64_t* full_v = bars + 5;    // V landed
    uint64_t* empty_v = bars + 6;   // V consumed (P.V MMAs done)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = p.G * p.L;
    const int num_tiles = p.n_groups * p.n_head;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmKV);
        mbar_init(full_qk, 1); mbar_init(empty_qk, 1); mbar_init(full_v, 1); mbar_init(empty_v, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_ptr);
    // the padding key rows of V are never written by TMA but are read by the second MMA: they must be finite (x 0)
    for (int i = threadIdx.x; i < kKVBytes / 16; i += kAttnThreads) reinterpret_cast<uint4*>(sV)[i] = make_uint4(0, 0, 0, 0);
    // P is block diagonal and a row's clip never changes: everything outside a row's own key slot is zeroed ONCE
    // here, the softmax below only ever rewrites the row's own LP columns
    for (int i = threadIdx.x; i < kPBytes / 16; i += kAttnThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 192;   // S: up to kMaxNK (<= 192) columns, O: 64

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t bytes = rows * 128u;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int g = tile / p.n_head, h = tile % p.n_head;
                const int row0 = g * rows;
                mbar_wait(empty_qk, (it & 1) ^ 1);
                mbar_expect_tx(full_qk, 2 * bytes);
                tma_load_2d(sQ, &tmQ, full_qk, p.q_col0 + h * 64, row0);
                for (int cl = 0; cl < p.G; ++cl)          // one box per clip, landing on its 16-aligned key slot
                    tma_load_2d(sK + cl * p.LP * 128, &tmKV, full_qk, p.k_col0 + h * 64, row0 + cl * p.L);
                mbar_wait(empty_v, (it & 1) ^ 1);
                mbar_expect_tx(full_v, bytes);
                for (int cl = 0; cl < p.G; ++cl)
                    tma_load_2d(sV + cl * p.LP * 128, &tmKV, full_v, p.v_col0 + h * 64, row0 + cl * p.L);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_f16(128, p.NK);
            const uint32_t idesc_o = make_idesc_f16(128, 64) | (1u << 16);     // B operand MN-major
            const uint32_t q = smem_u32(sQ), k = smem_u32(sK), v = smem_u32(sV), pp = smem_u32(sP);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(full_qk, it & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_f16(tmem_S, make_smem_desc<128>(q + kk * 32), make_smem_desc<128>(k + kk * 32), idesc_s, kk != 0);
                umma_commit(empty_qk);
                umma_commit(s_full);
                mbar_wait(p_full, it & 1);
                mbar_wait(full_v, it & 1);
                tc_fence_after();
                for (int kk = 0; kk < p.NK / 16; ++kk)
                    umma_f16(tmem_O, make_smem_desc<128>(pp + (kk >> 2) * kTileBytes + (kk & 3) * 32),
                             make_desc_mn128(v + kk * 2048), idesc_o, kk != 0);
                umma_commit(empty_v);
                umma_commit(o_full);
            }
        }
    } else {
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const int cl = r / p.L;                       // clip within the group (>= G for unused rows)
        const int c_lo = cl * p.LP, c_hi = c_lo + p.L; // own key columns
        const bool row_used = r < rows;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        unsigned char* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
        // key columns any row of this warp needs (tcgen05.ld is warp-wide, so the window is the union of the warp's
        // clips: 48 or 96 of the 144 columns for L = 34)
        const int wr0 = q4 * 32, wr1 = min(q4 * 32 + 31, rows - 1);
        const int w_lo = wr0 < rows ? (wr0 / p.L) * p.LP : 0;
        const int w_hi = wr0 < rows ? (wr1 / p.L + 1) * p.LP : 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int g = tile / p.n_head, h = tile % p.n_head;
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            // pass 1: row maximum over the clip's own columns
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = w_lo; c < w_hi; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(tmem_S + lane_addr + c, v);
                if (!row_used || c < c_lo || c >= c_hi) continue;              // not this row's key slot (16-aligned)
                if (c + 16 <= c_hi) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c + j < c_hi) mx = fmaxf(mx, v[j]);
                }
            }
            // pass 2: p = exp((s - max) / sqrt(d_k)), zeros elsewhere, fp16 into the swizzled A-operand layout
            float sum = 0.f;
#pragma unroll 1
            for (int c = w_lo; c < w_hi; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(tmem_S + lane_addr + c, v);
                if (!row_used || c < c_lo || c >= c_lo + p.LP) continue;      // not this row's key slot (16-aligned)
                const float mxs = mx * p.scale_log2e;
                if (c + 16 <= c_hi) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = exp2f(fmaf(v[j], p.scale_log2e, -mxs));
                        sum += v[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e = c + j < c_hi ? exp2f(fmaf(v[j], p.scale_log2e, -mxs)) : 0.f;
                        v[j] = e;
                        sum += e;
                    }
                }
#pragma unroll
                for (int j8 = 0; j8 < 2; ++j8) {
                    const int chunk = (c >> 3) + j8;              // 16-byte chunk index along K
                    uint4 u;
                    *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[8 * j8], v[8 * j8 + 1]);
                    *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[8 * j8 + 2], v[8 * j8 + 3]);
                    *reinterpret_cast<__half2*>(&u.z) = __floats2half2_rn(v[8 * j8 + 4], v[8 * j8 + 5]);
                    *reinterpret_cast<__half2*>(&u.w) = __floats2half2_rn(v[8 * j8 + 6], v[8 * j8 + 7]);
                    *reinterpret_cast<uint4*>(prow + (chunk >> 3) * kTileBytes + (((chunk & 7) ^ (r & 7)) << 4)) = u;
                }
            }
            fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            // output: O / sum
            mbar_wait(o_full, it & 1);
            tc_fence_after();
            const float inv = row_used ? 1.f / sum : 0.f;
            const int clip = g * p.G + cl;
            const bool store = row_used && clip < p.n_clips;
            __half* o = p.out + ((size_t)g * rows + r) * p.ldo + h * 64;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                __syncwarp();
                tmem_ld32(tmem_O + lane_addr + c, v);
                if (store) {
                    if (((reinterpret_cast<uintptr_t>(p.out) | (uintptr_t)(p.ldo * 2)) & 31) == 0) {
                        // 256-bit stores: every instruction writes whole 32-byte sectors
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint32_t u[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const __half2 h2 = __floats2half2_rn(v[16 * j + 2 * e] * inv, v[16 * j + 2 * e + 1] * inv);
                                u[e] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + c + 16 * j),
                                         "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                                         : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[8 * j] * inv, v[8 * j + 1] * inv);
                            *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[8 * j + 2] * inv, v[8 * j + 3] * inv);
                            *reinterpret_cast<__half2*>(&u.z) = __floats2half2_rn(v[8 * j + 4] * inv, v[8 * j + 5] * inv);
                            *reinterpret_cast<__half2*>(&u.w) = __floats2half2_rn(v[8 * j + 6] * inv, v[8 * j + 7] * inv);
                            reinterpret_cast<uint4*>(o + c)[j] = u;
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

int g_attn_sms = 0;

}  // namespace

int attn_tc_init_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&g_attn_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) == cudaSuccess ? 0 : -1;
}

// q rows: (B*L, ldq) with head h at columns q_col0 + 64 h; k / v rows: (B*L, ldkv) at k_col0 / v_col0 + 64 h.
// d_k = d_v = 64, L <= 64.  out: (B*L, ldo) fp16, head h at columns 64 h.
int launch_attention_tc(const __half* q, int ldq, int q_col0, const __half* kv, int ldkv, int k_col0, int v_col0, int B,
                        int L, int n_head, __half* out, int ldo, cudaStream_t s) {
    if (L < 1 || L > 128) return -1;
    AttnParams p;
    p.n_clips = B; p.L = L; p.LP = (L + 15) / 16 * 16;
    p.G = 128 / L;
    while (p.G > 1 && p.G * p.LP > kMaxNK) --p.G;
    p.NK = p.G * p.LP;
    if (p.NK > kMaxNK) return -1;
    p.n_head = n_head; p.n_groups = (B + p.G - 1) / p.G;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.out = out; p.ldo = ldo;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    const int rows = p.G * L;
    CUtensorMap tq, tkv;
    const uint64_t dq[2] = {(uint64_t)ldq, (uint64_t)B * L}, dkv[2] = {(uint64_t)ldkv, (uint64_t)B * L};
    const uint64_t sq[1] = {(uint64_t)ldq * 2}, skv[1] = {(uint64_t)ldkv * 2};
    const uint32_t box[2] = {64, (uint32_t)rows}, box_kv[2] = {64, (uint32_t)L};
    if (!make_tmap_f16(&tq, q, 2, dq, sq, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_tmap_f16(&tkv, kv, 2, dkv, skv, box_kv, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    const int tiles = p.n_groups * n_head;
    const int grid = tiles < 2 * g_attn_sms ? tiles : 2 * g_attn_sms;
    attn_tc_kernel<<<grid, kAttnThreads, kAttnSmem, s>>>(tq, tkv, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
