// K10 — the position-wise feed-forward block as ONE kernel (Full_model/SubLayers.py:74-84):
//     out = LayerNorm(x + W2 relu(W1 x + b1) + b2),  eps 1e-6,   d_model = 256, d_inner = NJ * 128
// The 128 x d_inner hidden tile never leaves the SM: per 128-row tile the hidden units are produced 128 at a time into
// a TMEM accumulator (GEMM1: X[128 x 256] . W1_j^T), pulled through registers (bias, ReLU, fp16) into shared memory in the
// swizzled K-major layout an MMA A operand needs, and immediately consumed by GEMM2 (Y[128 x 256] += H_j . W2_j^T), whose
// fp32 accumulator stays in TMEM for the whole tile; its epilogue adds bias and the fp32 residual and applies LayerNorm
// (one thread owns the 256-wide row, two passes over TMEM).  Against three kernels (w1 GEMM, w2 GEMM, LayerNorm) this
// removes the hidden tensor's round trip (2 x 285 MB per layer at 4096 TED clips) and the pre-LayerNorm one.
//
// TMEM (512 columns): Y 0..255 | H accumulators 256..383 and 384..511 (double buffered).
// Shared memory: X tile 64 KB | two fp16 H buffers 2 x 32 KB | weight ring 3 x 32 KB (a stage is two 128-row x 64-k boxes
// of W1 or one 256-row x 64-k box of W2) — 224 KB.
// Warps: 0 TMA producer, 1 MMA issuer, 2-5 / 6-9 hidden-chunk epilogues (even / odd chunks), 10-13 output epilogue.
// MMA order per tile: G1(0) G1(1) { G2(j) G1(j+2) }: the conversion of chunk j runs under G1(j+1) and G2(j-1).
//
// Every 128-row tile streams all of W1 and W2 (1 MB) from L2.  The kernel can run as thread-block CLUSTERS of CL CTAs
// that walk their row tiles in lock step: each CTA fetches 1/CL of every weight stage and TMA-multicasts it into the
// same ring slot of all CL CTAs; a slot is refilled when the MMA warps of ALL CTAs have released it (tcgen05.commit
// multicast onto every CTA's empty barrier); the X tiles and everything downstream stay private to a CTA.  Measured,
// the weight stream is not what bounds the kernel (see g_ffn_cluster below), so CL = 1 is the default.
#include <cstring>

#include "egx_common.cuh"
#include "tc_common.cuh"

namespace egx {

namespace {

using namespace tc;

constexpr int FM = 128;                 // rows per tile
constexpr int FD = 256;                 // d_model
constexpr int FH = 128;                 // hidden units per chunk
constexpr int FK = 64;                  // fp16 elements per 128-byte swizzle row
constexpr int kFfnThreads = 14 * 32;
constexpr int kXBytes = FM * FD * 2;                // 64 KB: 4 k-blocks of 16 KB
constexpr int kHBytes = FM * FH * 2;                // 32 KB: 2 k-blocks of 16 KB
constexpr int kStageBytes = 32 * 1024;
constexpr int kFfnStages = 3;
constexpr int kXchBytes = 2 * FM * 2 * 4;          // LayerNorm moment exchange between the two column halves (YW = 2)
// the dynamic shared-memory array is declared 1024-byte aligned (swizzle atoms), so no slack is reserved for aligning it
constexpr int kFfnSmem = kXBytes + 2 * kHBytes + kFfnStages * kStageBytes + 256 + kXchBytes;

constexpr int kFfnMaxInner = 2048;
// By value in the kernel's constant bank (7 KB of parameters): the epilogues index the bias / LayerNorm vectors with
// warp-uniform addresses every 32-column chunk, which the constant cache serves without an L2 round trip.
struct FfnParams {
    int M, NJ;                     // rows, hidden chunks (d_inner / 128)
    const float* resid;            // [M][256] fp32 residual (the FFN input)
    float* out32; __half* out16;   // [M][256]
    float b1[kFfnMaxInner];        // [d_inner]
    float b2[FD], ln_g[FD], ln_b[FD];
    unsigned long long* wait_cycles;   // attribution builds: [6] cycles the MMA thread of CTA 0 spent per wait kind + total
    int debug;                     // EGX_FFN_DEBUG (attribution builds only, wrong results): 1 = no weight loads, 2 = no
                                   // hidden-chunk conversion, 4 = no output epilogue work
};

// YW: warp groups of the output epilogue.  1: warps 2-5 / 6-9 convert the even / odd hidden chunks, warps 10-13 run the
// LayerNorm epilogue.  2: warps 2-5 convert all hidden chunks, warps 6-9 / 10-13 each take one 128-column half of the
// output rows and exchange their LayerNorm moments through shared memory — the output accumulator, the resource the next
// tile's second GEMM waits for, is drained twice as fast.
template <int CL, int YW>
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ FfnParams p) {
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    constexpr int kPiece = kStageBytes / CL;          // bytes of a stage one CTA fetches: a {64, 256 / CL} box
    constexpr int kPieceRows = 256 / CL;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    if (smem_u32(smem_raw) & 1023u) __trap();
    unsigned char* sX = smem;
    unsigned char* sH = sX + kXBytes;                       // [2][kHBytes]
    unsigned char* ring = sH + 2 * kHBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kFfnStages * kStageBytes);
    uint64_t* x_full = bars;            // X tile landed
    uint64_t* x_empty = bars + 1;       // last GEMM1 of the tile done
    uint64_t* w_full = bars + 2;        // [3]
    uint64_t* w_empty = bars + 5;       // [3]
    uint64_t* hacc_full = bars + 8;     // [2] GEMM1(j) done: H accumulator in TMEM
    uint64_t* hacc_empty = bars + 10;   // [2] accumulator read back (4 warps)
    uint64_t* hs_full = bars + 12;      // [2] fp16 H chunk in shared memory (4 warps)
    uint64_t* hs_empty = bars + 14;     // [2] GEMM2(j) done
    uint64_t* y_full = bars + 16;
    uint64_t* y_empty = bars + 17;      // output accumulator read back (4 warps)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);
    float* xch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);      // [half][row][mean, M2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.M + FM - 1) / FM;
    const int NJ = p.NJ;
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    // every CTA of a cluster runs the same number of tile rounds (the weight pipeline is shared); a round past the
    // last tile loads zero rows (TMA out-of-bounds fill) and stores nothing
    const int n_iter = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int i = 0; i < kFfnStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], CL); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hacc_full[i], 1); mbar_init(&hacc_empty[i], 4);
            mbar_init(&hs_full[i], 4); mbar_init(&hs_empty[i], 1);
        }
        mbar_init(y_full, 1); mbar_init(y_empty, 4 * YW);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();         // the peers' barriers exist before anything is multicast onto them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_Y = tmem_base, tmem_H = tmem_base + FD;

    if (warp == 0) {
        if (elect_one()) {
            int st = 0;
            uint32_t ph = 0, it = 0;
            auto next_stage = [&]() -> unsigned char* {
                mbar_wait(&w_empty[st], ph ^ 1);
                if (p.debug & 1) { mbar_arrive(&w_full[st]); return nullptr; }
                mbar_expect_tx(&w_full[st], kStageBytes);
                return ring + st * kStageBytes;
            };
            auto advance = [&]() { st = st + 1 == kFfnStages ? 0 : st + 1; ph ^= (st == 0); };
            // a stage is 256 swizzled 128-byte rows: two k-blocks x 128 hidden units of W1, or one hidden k-block x the
            // 256 output rows of W2; this CTA fetches rows [rank * 256 / CL, ...) of it for the whole cluster
            auto load_w1 = [&](int j) {                     // 2 stages: k-blocks {0,1}, {2,3} of W1 rows [128 j, 128 j + 128)
                for (int s2 = 0; s2 < 2; ++s2) {
                    unsigned char* dst = next_stage() + rank * kPiece;
                    if (p.debug & 1) { advance(); continue; }
                    const int row = rank * kPieceRows;      // row of the stage: k-block row / 128, hidden unit row % 128
                    if (CL > 1) tma_load_2d_mc(dst, &tmW1, &w_full[st], (2 * s2 + row / FH) * FK, j * FH + row % FH, kMask);
                    else {
                        tma_load_2d(dst, &tmW1, &w_full[st], (2 * s2) * FK, j * FH);
                        tma_load_2d(dst + 16384, &tmW1, &w_full[st], (2 * s2 + 1) * FK, j * FH);
                    }
                    advance();
                }
            };
            auto load_w2 = [&](int j) {                     // 2 stages: hidden k-blocks 2 j, 2 j + 1 of all 256 output rows
                for (int kb = 0; kb < 2; ++kb) {
                    unsigned char* dst = next_stage() + rank * kPiece;
                    if (p.debug & 1) { advance(); continue; }
                    if (CL > 1) tma_load_2d_mc(dst, &tmW2, &w_full[st], j * FH + kb * FK, rank * kPieceRows, kMask);
                    else tma_load_2d(dst, &tmW2, &w_full[st], j * FH + kb * FK, 0);
                    advance();
                }
            };
            for (int tile = blockIdx.x; it < (uint32_t)n_iter; tile += gridDim.x, ++it) {
                mbar_wait(x_empty, (it & 1) ^ 1);
                mbar_expect_tx(x_full, kXBytes);
                for (int kb = 0; kb < FD / FK; ++kb) tma_load_2d(sX + kb * 16384, &tmX, x_full, kb * FK, tile * FM);
                // the same order the MMA warp consumes them in
                load_w1(0);
                load_w1(1);
                for (int j = 0; j < NJ; ++j) {
                    load_w2(j);
                    if (j + 2 < NJ) load_w1(j + 2);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc1 = make_idesc_f16(FM, FH), idesc2 = make_idesc_f16(FM, FD);
            constexpr uint32_t kDescHi = smem_desc_hi<128>();
            const uint32_t x_lo = smem_desc_lo(smem_u32(sX)), h_lo = smem_desc_lo(smem_u32(sH));
            int st = 0;
            uint32_t ph = 0, it = 0;
            uint32_t n_g1[2] = {0, 0}, n_g2[2] = {0, 0};
#ifdef EGX_ATTRIBUTION
            unsigned long long wc[6] = {0, 0, 0, 0, 0, 0};
            const long long t_begin = clock64();
#define FFN_TIMED_WAIT(kind, bar, par) do { const long long _t = clock64(); mbar_wait(bar, par); wc[kind] += clock64() - _t; } while (0)
#else
#define FFN_TIMED_WAIT(kind, bar, par) mbar_wait(bar, par)
#endif
            auto advance = [&]() { st = st + 1 == kFfnStages ? 0 : st + 1; ph ^= (st == 0); };
            auto gemm1 = [&](int j, bool last) {
                const int b = j & 1;
                FFN_TIMED_WAIT(1, &hacc_empty[b], (n_g1[b] & 1) ^ 1);
                ++n_g1[b];
                tc_fence_after();
                const uint32_t d = tmem_H + b * FH;
                for (int s2 = 0; s2 < 2; ++s2) {
                    FFN_TIMED_WAIT(2, &w_full[st], ph);
                    tc_fence_after();
                    const uint32_t w_lo = smem_desc_lo(smem_u32(ring + st * kStageBytes));
#pragma unroll
                    for (int kbl = 0; kbl < 2; ++kbl) {
                        const int kb = 2 * s2 + kbl;
#pragma unroll
                        for (int k = 0; k < FK / 16; ++k)
                            umma_f16_lo<kDescHi>(d, x_lo + ((kb * 16384) >> 4) + 2 * k, w_lo + ((kbl * 16384) >> 4) + 2 * k, idesc1,
                                                 kb != 0 || k != 0);
                    }
                    if (CL > 1) umma_commit_mc(&w_empty[st], kMask); else umma_commit(&w_empty[st]);
                    advance();
                }
                umma_commit(&hacc_full[b]);
                if (last) umma_commit(x_empty);
            };
            auto gemm2 = [&](int j, bool last) {
                const int b = j & 1;
                FFN_TIMED_WAIT(3, &hs_full[b], n_g2[b] & 1);
                ++n_g2[b];
                tc_fence_after();
                for (int kb = 0; kb < 2; ++kb) {
                    FFN_TIMED_WAIT(2, &w_full[st], ph);
                    tc_fence_after();
                    const uint32_t w_lo = smem_desc_lo(smem_u32(ring + st * kStageBytes));
                    const uint32_t a_lo = h_lo + ((b * kHBytes + kb * 16384) >> 4);
#pragma unroll
                    for (int k = 0; k < FK / 16; ++k)
                        umma_f16_lo<kDescHi>(tmem_Y, a_lo + 2 * k, w_lo + 2 * k, idesc2, j != 0 || kb != 0 || k != 0);
                    if (CL > 1) umma_commit_mc(&w_empty[st], kMask); else umma_commit(&w_empty[st]);
                    advance();
                }
                umma_commit(&hs_empty[b]);
                if (last) umma_commit(y_full);
            };
            for (; it < (uint32_t)n_iter; ++it) {
                FFN_TIMED_WAIT(0, x_full, it & 1);
                tc_fence_after();
                gemm1(0, NJ == 1);
                gemm1(1, NJ == 2);
                for (int j = 0; j < NJ; ++j) {
                    if (j == 0) { FFN_TIMED_WAIT(4, y_empty, (it & 1) ^ 1); tc_fence_after(); }
                    gemm2(j, j == NJ - 1);
                    if (j + 2 < NJ) gemm1(j + 2, j + 3 == NJ);
                }
            }
#ifdef EGX_ATTRIBUTION
            if (p.wait_cycles && blockIdx.x == 0) {
                wc[5] = clock64() - t_begin;
                for (int i = 0; i < 6; ++i) p.wait_cycles[i] = wc[i];
            }
#endif
        }
    } else if (warp < (YW == 2 ? 6 : 10)) {
        // hidden-chunk epilogue: TMEM -> + b1, ReLU, fp16 -> shared memory (SWIZZLE_128B K-major A operand of GEMM2)
        const int b0 = YW == 2 ? 0 : (warp - 2) >> 2;   // YW = 1: chunks j with j & 1 == b0; YW = 2: every chunk
        const int q = warp & 3, r = q * 32 + lane;
        uint32_t n = 0;                                 // uses of buffer b so far = n >> (YW == 2)
        for (int round = 0; round < n_iter; ++round) {
            for (int j = b0; j < NJ; j += (YW == 2 ? 1 : 2), ++n) {
                const int b = j & 1;
                const uint32_t taddr = tmem_H + b * FH + ((uint32_t)(q * 32) << 16);
                unsigned char* hrow = sH + b * kHBytes + (r >> 3) * 1024 + (r & 7) * 128;
                const uint32_t nb_ = YW == 2 ? n >> 1 : n;
#define n nb_
                mbar_wait(&hacc_full[b], n & 1);
                tc_fence_after();
                mbar_wait(&hs_empty[b], (n & 1) ^ 1);
                if (p.debug & 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&hacc_empty[b]); mbar_arrive(&hs_full[b]); }
                    continue;
                }
#pragma unroll 1
                for (int c = 0; c < FH / 32; ++c) {
                    float v[32];
                    __syncwarp();
                    tmem_ld32(taddr + c * 32, v);
                    if (c == FH / 32 - 1) {             // accumulator fully read: GEMM1(j + 2) may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&hacc_empty[b]);
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e] + p.b1[j * FH + c * 32 + e], 0.f);
                    unsigned char* dst = hrow + (c >> 1) * 16384;           // 64-wide k-block of this 32-column chunk
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[8 * i], v[8 * i + 1]);
                        *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                        *reinterpret_cast<__half2*>(&u.z) = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
                        *reinterpret_cast<__half2*>(&u.w) = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                        const int chunk = (c & 1) * 4 + i;                    // 16-byte chunk inside the 128-byte row
                        *reinterpret_cast<uint4*>(dst + ((chunk ^ (r & 7)) << 4)) = u;
                    }
                }
                fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&hs_full[b]);
#undef n
            }
        }
    } else if (YW == 2) {
        // output epilogue, split over two warp groups by column half (see YW above)
        const int half = (warp - 6) >> 2, q = warp & 3, r = q * 32 + lane;
        const uint32_t taddr = tmem_Y + half * (FD / 2) + ((uint32_t)(q * 32) << 16);
        const int col_h = half * (FD / 2);
        uint32_t it = 0;
        for (int tile = blockIdx.x; it < (uint32_t)n_iter; tile += gridDim.x, ++it) {
            const int row0 = tile * FM + q * 32 + (lane & ~3);
            uint32_t rn[32];
            if (tile * FM + r < p.M) {
                const float* own = p.resid + (size_t)(tile * FM + r) * FD + col_h;
#pragma unroll
                for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(own + l * 32));
            }
            load_rows_t(p.resid, FD, row0, p.M, col_h, lane, rn);
            mbar_wait(y_full, it & 1);
            tc_fence_after();
            float v0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < FD / 64; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                seg_transpose4<8>(rn, lane);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += p.b2[col_h + c * 32 + j] + __uint_as_float(rn[j]);
                if (c + 1 < FD / 64) load_rows_t(p.resid, FD, row0, p.M, col_h + (c + 1) * 32, lane, rn);
                if (c == 0) v0 = v[0];
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float d = v[j] - v0; s1 += d; s2 = fmaf(d, d, s2); }
                tmem_st32(taddr + c * 32, v);
            }
            tmem_st_wait();
            // moments of this half about its own first value, combined with the other half's (Chan's update, 128 + 128)
            const float mean_h = v0 + s1 * (2.f / FD), m2_h = fmaxf(s2 - s1 * s1 * (2.f / FD), 0.f);
            xch[(half * FM + r) * 2] = mean_h;
            xch[(half * FM + r) * 2 + 1] = m2_h;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mean_o = xch[((half ^ 1) * FM + r) * 2], m2_o = xch[((half ^ 1) * FM + r) * 2 + 1];
            const float mean = 0.5f * (mean_h + mean_o);
            const float dm = mean_h - mean_o;
            const float rstd = rsqrtf((m2_h + m2_o + dm * dm * (FD / 4)) * (1.f / FD) + 1e-6f);
#pragma unroll 1
            for (int c = 0; c < FD / 64; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                if (c == FD / 64 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(y_empty);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = fmaf((v[j] - mean) * rstd, p.ln_g[col_h + c * 32 + j], p.ln_b[col_h + c * 32 + j]);
                store_rows_t(v, p.out32, FD, p.out16, FD, row0, p.M, col_h + c * 32, lane, (p.debug & 128) != 0);
            }
        }
    } else {
        // output epilogue: Y + b2 + residual -> LayerNorm -> fp32 and fp16 rows.  Global accesses go through the
        // 4-lane transposed layout (tc_common.cuh: seg_transpose4): one full 128-byte line per row and access instead
        // of 32 bytes per line, which is what this epilogue is bound by; b2 and the LayerNorm parameters sit in the
        // kernel's constant bank (no L2 round trip per 32-column chunk).
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t taddr = tmem_Y + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        for (int tile = blockIdx.x; it < (uint32_t)n_iter; tile += gridDim.x, ++it) {
            const int row0 = tile * FM + q * 32 + (lane & ~3);       // first row of this lane's group of four
            uint32_t rn[32];
            // this thread's residual row (1 KB = 8 lines) goes to L2 now, a whole tile of MMA time ahead of its use: the
            // epilogue below runs while the tensor pipe waits for the accumulator and must not pay DRAM latency
            if (tile * FM + r < p.M) {
                const float* own = p.resid + (size_t)(tile * FM + r) * FD;
#pragma unroll
                for (int l = 0; l < 8; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(own + l * 32));
            }
            load_rows_t(p.resid, FD, row0, p.M, 0, lane, rn);
            mbar_wait(y_full, it & 1);
            tc_fence_after();
            if (p.debug & 4) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(y_empty);
                continue;
            }
            float v0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < FD / 32; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                seg_transpose4<8>(rn, lane);                         // -> this thread's own row, columns in order
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += p.b2[c * 32 + j] + __uint_as_float(rn[j]);
                if (c + 1 < FD / 32) load_rows_t(p.resid, FD, row0, p.M, (c + 1) * 32, lane, rn);
                if (c == 0) v0 = v[0];
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float d = v[j] - v0; s1 += d; s2 = fmaf(d, d, s2); }
                tmem_st32(taddr + c * 32, v);
            }
            tmem_st_wait();
            const float mean_d = s1 * (1.f / FD);
            const float rstd = rsqrtf(fmaxf(s2 * (1.f / FD) - mean_d * mean_d, 0.f) + 1e-6f);
            const float mean = v0 + mean_d;
#pragma unroll 1
            for (int c = 0; c < FD / 32; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                if (c == FD / 32 - 1) {                 // accumulator fully read: the next tile's GEMM2 may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(y_empty);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf((v[j] - mean) * rstd, p.ln_g[c * 32 + j], p.ln_b[c * 32 + j]);
                store_rows_t(v, p.out32, FD, p.out16, FD, row0, p.M, c * 32, lane, (p.debug & 128) != 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();         // no CTA leaves while a peer may still multicast into it or signal its barriers
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int g_ffn_sms = 0;
// CTAs per cluster.  Measured at 4096 TED clips (S6 of a step, six launches): 1 -> 5.09 ms, 2 -> 5.16 ms, 4 -> 6.5 ms, and
// with the weight loads removed altogether only 0.08 ms less: the 3-stage ring already hides the weight stream, the
// lock step of a cluster only adds coupling.  What bounds the kernel is the output epilogue (row-per-thread global
// accesses: 0.9 of the 1.25 ms the six launches take beyond their MMA time).  EGX_FFN_CLUSTER selects 2 / 4 in attribution builds.
int g_ffn_cluster = 1;
int g_ffn_yw = 1;          // output-epilogue warp groups (EGX_FFN_YW in attribution builds)

template <int CL, int YW>
int launch_ffn_cl(const CUtensorMap& tx, const CUtensorMap& t1, const CUtensorMap& t2, const FfnParams& p, int tiles, cudaStream_t s) {
    int grid = tiles < g_ffn_sms ? tiles : g_ffn_sms;
    grid = (grid + CL - 1) / CL * CL;
    if (grid > g_ffn_sms) grid = g_ffn_sms / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFfnThreads); cfg.dynamicSmemBytes = kFfnSmem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ffn_tc_kernel<CL, YW>, tx, t1, t2, p) == cudaSuccess ? 1 : -1;
}

}  // namespace

int ffn_tc_init_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&g_ffn_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_ffn_cluster = env_switch("EGX_FFN_CLUSTER", g_ffn_cluster);
    g_ffn_yw = env_switch("EGX_FFN_YW", g_ffn_yw);
    if (cudaFuncSetAttribute(ffn_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ffn_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ffn_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ffn_tc_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem) != cudaSuccess) return -1;
    return 0;
}

// true when launch_ffn_tc can run this geometry (the callers fall back to the three-kernel chain otherwise)
bool ffn_tc_supported(int d_model, int d_inner) {
    return d_model == FD && d_inner % (2 * FH) == 0 && d_inner >= 2 * FH && d_inner <= kFfnMaxInner;
}

// x16 [M][256] fp16 (the FFN input as the GEMM operand), resid [M][256] fp32 (the same input in fp32), w1 [d_inner][256],
// w2 [256][d_inner] fp16 with pitches ldw1 / ldw2; out32 / out16 [M][256], all 32-byte aligned device pointers.
// b1 [d_inner], b2 / ln_g / ln_b [256] are HOST vectors: they travel as kernel parameters.
int launch_ffn_tc(const __half* x16, const float* resid, const __half* w1, int ldw1, const float* b1, const __half* w2, int ldw2,
                  const float* b2, const float* ln_g, const float* ln_b, int M, int d_inner, float* out32, __half* out16,
                  cudaStream_t s) {
    if (!ffn_tc_supported(FD, d_inner) || !b1 || !b2 || !ln_g || !ln_b || !resid || !out32 || !out16) return -1;
    CUtensorMap tx, t1, t2;
    const uint64_t dX[2] = {(uint64_t)FD, (uint64_t)M}, sXp[1] = {(uint64_t)FD * 2};
    const uint64_t d1[2] = {(uint64_t)FD, (uint64_t)d_inner}, s1[1] = {(uint64_t)ldw1 * 2};
    const uint64_t d2[2] = {(uint64_t)d_inner, (uint64_t)FD}, s2[1] = {(uint64_t)ldw2 * 2};
    const int cl = g_ffn_cluster == 4 ? 4 : (g_ffn_cluster == 2 ? 2 : 1);
    // clustered: every CTA fetches a {64, 256 / CL}-row piece of each weight stage
    const uint32_t bX[2] = {FK, FM}, b1x[2] = {FK, cl > 1 ? 256u / cl : (uint32_t)FH}, b2x[2] = {FK, cl > 1 ? 256u / cl : (uint32_t)FD};
    if (!make_tmap_f16(&tx, x16, 2, dX, sXp, bX, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_tmap_f16(&t1, w1, 2, d1, s1, b1x, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_tmap_f16(&t2, w2, 2, d2, s2, b2x, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    FfnParams p;
    p.M = M; p.NJ = d_inner / FH; p.resid = resid; p.out32 = out32; p.out16 = out16;
    std::memcpy(p.b1, b1, sizeof(float) * d_inner);
    std::memcpy(p.b2, b2, sizeof(float) * FD);
    std::memcpy(p.ln_g, ln_g, sizeof(float) * FD);
    std::memcpy(p.ln_b, ln_b, sizeof(float) * FD);
    p.debug = env_switch("EGX_FFN_DEBUG", 0);
    p.wait_cycles = nullptr;
#ifdef EGX_ATTRIBUTION
    {   // MMA-thread wait attribution of the previous launch, printed when EGX_FFN_WAITS=1
        static unsigned long long* dbg = nullptr;
        if (env_switch("EGX_FFN_WAITS", 0)) {
            if (!dbg) cudaMallocManaged(&dbg, 6 * sizeof(unsigned long long));
            else {
                cudaStreamSynchronize(s);
                fprintf(stderr, "ffn waits (cycles): x_full %llu hacc_empty %llu w_full %llu hs_full %llu y_empty %llu total %llu\n", dbg[0],
                        dbg[1], dbg[2], dbg[3], dbg[4], dbg[5]);
            }
            p.wait_cycles = dbg;
        }
    }
#endif
    const int tiles = (M + FM - 1) / FM;
    if (cl == 4) return launch_ffn_cl<4, 1>(tx, t1, t2, p, tiles, s);
    if (cl == 2) return launch_ffn_cl<2, 1>(tx, t1, t2, p, tiles, s);
    if (g_ffn_yw == 2) return launch_ffn_cl<1, 2>(tx, t1, t2, p, tiles, s);
    return launch_ffn_cl<1, 1>(tx, t1, t2, p, tiles, s);
}

}  // namespace egx
