// K3 — 3x3 / 1x1 convolutions of the SE-ResNet trunk as implicit GEMM on tcgen05.
//   M = output pixels (BH x BW spatial patches of one clip, <= 128 rows per tile)
//   N = cout (32 / 64 / 128, final conv 34|60 padded to 48|64),  K = taps * cin
// Follows Full_model/ResNetBlocks.py:24-30 (conv-ReLU-BN / conv-BN), Full_model/ResNetSE34V2.py:43-47
// (1x1 stride-2 downsample + BN) and Full_model/Models.py:121-122 (final conv + BN).
//
// Activations are NHWC fp16.  No im2col buffer exists anywhere: for every (tap, 64-channel chunk) the
// producer warp issues ONE 4-D TMA box load {channels, BW, BH, 1} whose W/H start coordinate is shifted by
// the tap (and may be -1 or run past the edge): TMA's out-of-bounds zero fill IS the convolution padding,
// and its element stride is the convolution stride.  The box lands in shared memory as BH*BW dense rows of
// 128 B (64 B for cin = 32) in the SWIZZLE_128B (64B) pattern the UMMA descriptor expects, i.e. directly as
// the K-major A operand.  Weights [cout][tap*cin] are the K-major B operand (2-D TMA); when they fit
// (<= 80 KB: every conv but the 128->128 ones) they are loaded once per CTA and stay resident.
//
// Persistent, warp-specialised: grid = #SMs, each CTA walks tiles blockIdx.x, +gridDim.x, ...
//   warp 0   TMA producer, runs ahead across tile boundaries through a kStages smem ring
//   warp 1   single-thread tcgen05.mma issuer; accumulators double-buffered in TMEM
//   warps 2-9 two epilogue groups of 4 warps, one per TMEM accumulator buffer (even / odd tiles), one pixel
//            per thread: tcgen05.ld -> bias / ReLU / folded BatchNorm (parameters staged in smem once per
//            CTA) -> fp16 store, overlapping the next tiles' MMAs; optionally the per-(clip, tile, channel)
//            partial sums the SE gate needs (fixed order, no atomics: results do not depend on batch size
//            or GPU count).
#include "egx_common.cuh"
#include "tc_common.cuh"

#include <cstdlib>

namespace egx {

namespace {

using namespace tc;

constexpr int kConvThreads = 320;      // producer + MMA + 2 x 4 epilogue warps
constexpr int kSmemBudget = 200 * 1024;

struct ConvTcParams {
    int Ho, Wo;            // output map
    int BW, BH;            // output patch per tile
    int MW;                // row pitch of the M index: BW, or BW + 2 with halo reuse (MW*BH <= 128)
    int tiles_w, tiles_h;
    int num_tiles;         // B * tiles_h * tiles_w
    int ks, stride, pad;
    int cout;
    int relu_first;
    const float* bias;     // may be null
    const float* scale;
    const float* shift;
    __half* out;           // NHWC (B,Ho,Wo,cout) or, if nchw, (B,cout,Ho*Wo)
    int nchw;
    float* se_part;        // [B][tiles_h*tiles_w][cout] partial channel sums, or null
};

// HALO (stride-1 3x3 with resident weights): instead of nine shifted boxes per tile, ONE box per 64-channel
// chunk brings the (BH+2) x (BW+2) input patch; the M index runs over BH x (BW+2) positions (two junk
// columns per row), so tap (dy,dx) is the SAME shared-memory patch read from a start address shifted by
// (dy*(BW+2)+dx) rows.  TMA and UMMA both apply the 128B/64B swizzle on absolute shared-memory address
// bits, so a row-shifted descriptor still sees the pattern TMA wrote.  L2->SM traffic per tile drops ~7x.
constexpr int kPatchRows = 176;          // >= 128 + 2*(BW+2) + 2 with BW + 2 <= 20

template <int CIN, int NPAD, int TAPS, bool HALO = false>
struct ConvCfg {
    static constexpr int CK = CIN < 64 ? CIN : 64;           // channels per K block
    static constexpr int kSwz = CK * 2;                      // 64 or 128 byte rows
    static constexpr int kChunks = CIN / CK;
    static constexpr int kNumKb = TAPS * kChunks;            // K blocks per tile
    static constexpr int kABytes = 128 * kSwz;
    static constexpr int kBBytes = ((NPAD * kSwz + 1023) / 1024) * 1024;
    static constexpr bool kResidentB = kNumKb * kBBytes <= 80 * 1024;
    // K blocks handled per pipeline stage (one barrier round-trip): keep >= 4 MMAs of work per wait
    static constexpr int kKbPerStage = HALO ? kNumKb : ((CK == 32 && TAPS == 9) ? 3 : 1);
    static constexpr int kPatchBytes = kPatchRows * kSwz;
    static constexpr int kStageBytes = HALO ? kChunks * kPatchBytes
                                            : kKbPerStage * (kABytes + (kResidentB ? 0 : kBBytes));
    static constexpr int kResBytes = kResidentB ? kNumKb * kBBytes : 0;
    static constexpr int kStagesRaw = (kSmemBudget - kResBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kStagesPerTile = kNumKb / kKbPerStage;
    static constexpr int kBarOffset = kResBytes + kStages * kStageBytes;
    static constexpr int kParOffset = kBarOffset + 256;      // bias | scale | shift, 128 floats each
    static constexpr int kRedOffset = kParOffset + 3 * 128 * 4;   // 2 groups x 4 warps x 128 floats (SE sums)
    static constexpr int kTotal = kRedOffset + 2 * 4 * 128 * 4 + 1024;
    static constexpr int kAccStride = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128);
    static constexpr uint32_t kTmemCols = 2 * kAccStride;
    static_assert(kNumKb % kKbPerStage == 0, "stage must divide the K loop");
    static_assert(kStages >= 2, "not enough shared memory for a pipeline");
    static_assert(!HALO || (TAPS == 9 && kResidentB), "halo reuse needs a 3x3 conv with resident weights");
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n)); }

template <int CIN, int NPAD, int TAPS, bool HALO>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcParams p) {
    using S = ConvCfg<CIN, NPAD, TAPS, HALO>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ring = smem + S::kResBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty = full + S::kStages;
    uint64_t* tmem_full = empty + S::kStages;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* b_full = tmem_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);
    float* par = reinterpret_cast<float*>(smem + S::kParOffset);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_clip = p.tiles_w * p.tiles_h;
    constexpr int KS = TAPS == 9 ? 3 : 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < S::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<S::kTmemCols>(tmem_ptr);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {
        const int n = threadIdx.x - 64;
        par[n] = (p.bias && n < p.cout) ? p.bias[n] : 0.f;
        par[128 + n] = n < p.cout ? p.scale[n] : 0.f;
        par[256 + n] = n < p.cout ? p.shift[n] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            if (S::kResidentB) {
                mbar_expect_tx(b_full, (uint32_t)S::kNumKb * NPAD * S::kSwz);
                for (int kb = 0; kb < S::kNumKb; ++kb)
                    tma_load_2d(smem + kb * S::kBBytes, &tmB, b_full, (kb / S::kChunks) * CIN + (kb % S::kChunks) * S::CK, 0);
            }
            const uint32_t a_bytes = (uint32_t)p.BW * p.BH * S::kSwz;
            const uint32_t stage_tx = S::kKbPerStage * (a_bytes + (S::kResidentB ? 0u : (uint32_t)NPAD * S::kSwz));
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int b = tile / tiles_per_clip;
                const int t = tile - b * tiles_per_clip;
                const int wi0 = (t % p.tiles_w) * p.BW * p.stride - p.pad;
                const int hi0 = (t / p.tiles_w) * p.BH * p.stride - p.pad;
                if (HALO) {
                    const int st = it % S::kStages;
                    mbar_wait(&empty[st], ((it / S::kStages) & 1) ^ 1);
                    unsigned char* dst = ring + st * S::kStageBytes;
                    mbar_expect_tx(&full[st], (uint32_t)S::kChunks * p.MW * (p.BH + 2) * S::kSwz);
#pragma unroll
                    for (int ch = 0; ch < S::kChunks; ++ch)
                        tma_load_4d(dst + ch * S::kPatchBytes, &tmA, &full[st], ch * S::CK, wi0, hi0, b);
                    ++it;
                    continue;
                }
                for (int sg = 0; sg < S::kStagesPerTile; ++sg, ++it) {
                    const int st = it % S::kStages;
                    mbar_wait(&empty[st], ((it / S::kStages) & 1) ^ 1);
                    unsigned char* dst = ring + st * S::kStageBytes;
                    mbar_expect_tx(&full[st], stage_tx);
#pragma unroll
                    for (int j = 0; j < S::kKbPerStage; ++j) {
                        const int kb = sg * S::kKbPerStage + j;
                        const int tap = kb / S::kChunks, chunk = kb % S::kChunks;
                        tma_load_4d(dst + j * S::kABytes, &tmA, &full[st], chunk * S::CK, wi0 + tap % KS,
                                    hi0 + tap / KS, b);
                        if (!S::kResidentB)
                            tma_load_2d(dst + S::kKbPerStage * S::kABytes + j * S::kBBytes, &tmB, &full[st],
                                        tap * CIN + chunk * S::CK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, NPAD);
            if (S::kResidentB) { mbar_wait(b_full, 0); tc_fence_after(); }
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * S::kAccStride;
                if (HALO) {
                    const int st = it % S::kStages;
                    mbar_wait(&full[st], (it / S::kStages) & 1);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(ring + st * S::kStageBytes);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t shift = (uint32_t)((tap / 3) * p.MW + (tap % 3)) * S::kSwz;
#pragma unroll
                        for (int ch = 0; ch < S::kChunks; ++ch) {
                            const uint32_t a = a0 + ch * S::kPatchBytes + shift;
                            const uint32_t bb = smem_u32(smem + (tap * S::kChunks + ch) * S::kBBytes);
#pragma unroll
                            for (int k = 0; k < S::CK / 16; ++k)
                                umma_f16(d, make_smem_desc<S::kSwz>(a + k * 32), make_smem_desc<S::kSwz>(bb + k * 32),
                                         idesc, (tap | ch | k) != 0);
                        }
                    }
                    umma_commit(&empty[st]);
                    umma_commit(&tmem_full[acc]);
                    ++it;
                    continue;
                }
                for (int sg = 0; sg < S::kStagesPerTile; ++sg, ++it) {
                    const int st = it % S::kStages;
                    mbar_wait(&full[st], (it / S::kStages) & 1);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(ring + st * S::kStageBytes);
#pragma unroll
                    for (int j = 0; j < S::kKbPerStage; ++j) {
                        const int kb = sg * S::kKbPerStage + j;
                        const uint32_t a = a0 + j * S::kABytes;
                        const uint32_t bb = S::kResidentB ? smem_u32(smem + kb * S::kBBytes)
                                                          : a0 + S::kKbPerStage * S::kABytes + j * S::kBBytes;
#pragma unroll
                        for (int k = 0; k < S::CK / 16; ++k)
                            umma_f16(d, make_smem_desc<S::kSwz>(a + k * 32), make_smem_desc<S::kSwz>(bb + k * 32),
                                     idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[st]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ================= epilogue =================
        const int grp = (warp - 2) >> 2;          // accumulator buffer / tile parity this group drains
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const int ph_ = r / p.MW, pw_ = r % p.MW;
        float* red = reinterpret_cast<float*>(smem + S::kRedOffset) + grp * 512;
        const bool relu_first = p.relu_first != 0;
        uint32_t tcount = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, tcount += 2) {
            const int b = tile / tiles_per_clip;
            const int t = tile - b * tiles_per_clip;
            const int ho = (t / p.tiles_w) * p.BH + ph_, wo = (t % p.tiles_w) * p.BW + pw_;
            const bool valid = ph_ < p.BH && pw_ < p.BW && ho < p.Ho && wo < p.Wo;
            mbar_wait(&tmem_full[grp], (tcount >> 1) & 1);
            tc_fence_after();
            const size_t pix = ((size_t)b * p.Ho + ho) * p.Wo + wo;
            const uint32_t taddr = tmem_base + grp * S::kAccStride + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < (NPAD + 31) / 32; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                const int nb = c * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bi = *reinterpret_cast<const float4*>(par + nb + 4 * j4);
                    const float4 sc = *reinterpret_cast<const float4*>(par + 128 + nb + 4 * j4);
                    const float4 sh = *reinterpret_cast<const float4*>(par + 256 + nb + 4 * j4);
                    const float bv[4] = {bi.x, bi.y, bi.z, bi.w}, sv[4] = {sc.x, sc.y, sc.z, sc.w},
                                hv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float tv = v[4 * j4 + e] + bv[e];
                        if (relu_first) tv = fmaxf(tv, 0.f);
                        tv = fmaf(tv, sv[e], hv[e]);
                        v[4 * j4 + e] = valid ? tv : 0.f;
                    }
                }
                if (valid) {
                    if (!p.nchw) {
                        __half* o = p.out + pix * p.cout + nb;     // cout is a multiple of 32 on this path
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
                        const size_t hw = (size_t)p.Ho * p.Wo;
                        __half* o = p.out + (size_t)b * p.cout * hw + (size_t)ho * p.Wo + wo;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.cout) o[(size_t)(nb + j) * hw] = __float2half_rn(v[j]);
                    }
                }
                if (p.se_part) {
                    // column sums over the warp's 32 rows: butterfly "transpose-reduce", 31 shuffles for 32 columns;
                    // lane l ends up holding the sum of column nb + l
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const bool upper = (lane & step) != 0;
#pragma unroll
                        for (int j = 0; j < step; ++j) {
                            const float send = upper ? v[j] : v[j + step];
                            const float keep = upper ? v[j + step] : v[j];
                            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                        }
                    }
                    red[q * 128 + nb + lane] = v[0];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[grp]);
            if (p.se_part) {
                named_bar_sync(1 + grp, 128);
                const int n = q * 32 + lane;
                if (n < p.cout)
                    p.se_part[((size_t)b * tiles_per_clip + t) * p.cout + n] =
                        (red[n] + red[128 + n]) + (red[256 + n] + red[384 + n]);
                named_bar_sync(1 + grp, 128);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<S::kTmemCols>(tmem_base);
}

// (BW, BH) with BW*BH <= 128 that wastes the fewest MMA rows on an Ho x Wo map
void pick_patch(int Ho, int Wo, int* bw, int* bh) {
    long best = -1;
    *bw = 1; *bh = 1;
    for (int w = 1; w <= 128 && w <= Wo; ++w) {
        const int h = (128 / w) < Ho ? (128 / w) : Ho;
        const long tiles = (long)((Ho + h - 1) / h) * ((Wo + w - 1) / w);
        // fewest tiles; ties go to the wider patch (longer contiguous runs per TMA box row)
        if (best < 0 || tiles <= best) { best = tiles; *bw = w; *bh = h; }
    }
}


// halo variant: BW + 2 <= 20 and BH * (BW + 2) <= 128; fewest tiles wins
void pick_halo_patch(int Ho, int Wo, int* bw, int* bh) {
    long best = -1;
    *bw = 1; *bh = 1;
    for (int w = 1; w <= 18 && w <= Wo; ++w) {
        const int h = (128 / (w + 2)) < Ho ? (128 / (w + 2)) : Ho;
        const long tiles = (long)((Ho + h - 1) / h) * ((Wo + w - 1) / w);
        if (best < 0 || tiles <= best) { best = tiles; *bw = w; *bh = h; }
    }
}

int g_num_sms = 0;
int g_halo = 1;      // EGX_CONV_HALO: 0 = off, 1 = 64->64 convs (default), 2 = also 32->32

template <int CIN, int NPAD, int TAPS, bool HALO>
int launch_one(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw, float* se_part,
               cudaStream_t s) {
    using S = ConvCfg<CIN, NPAD, TAPS, HALO>;
    ConvTcParams p;
    p.ks = c.ks; p.stride = c.stride; p.pad = c.ks / 2;
    p.Ho = (Hin + 2 * p.pad - c.ks) / c.stride + 1;
    p.Wo = (Win + 2 * p.pad - c.ks) / c.stride + 1;
    if (HALO) pick_halo_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    else pick_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    p.MW = HALO ? p.BW + 2 : p.BW;
    p.tiles_w = (p.Wo + p.BW - 1) / p.BW;
    p.tiles_h = (p.Ho + p.BH - 1) / p.BH;
    p.num_tiles = B * p.tiles_w * p.tiles_h;
    p.cout = c.cout; p.relu_first = c.relu_first;
    p.bias = c.bias; p.scale = c.scale; p.shift = c.shift;
    p.out = out; p.nchw = nchw; p.se_part = se_part;

    CUtensorMap ta, tb;
    const uint64_t dA[4] = {(uint64_t)CIN, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    const uint64_t sA[3] = {(uint64_t)CIN * 2, (uint64_t)Win * CIN * 2, (uint64_t)Hin * Win * CIN * 2};
    // with an element (traversal) stride e the box spans boxDim positions and keeps ceil(boxDim / e) of them
    const uint32_t bA[4] = {(uint32_t)S::CK, (uint32_t)(HALO ? p.MW : p.BW * c.stride),
                            (uint32_t)(HALO ? p.BH + 2 : p.BH * c.stride), 1};
    const uint32_t eA[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    const CUtensorMapSwizzle swz = S::kSwz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (!make_tmap_f16(&ta, in, 4, dA, sA, bA, eA, swz)) return -1;
    const int K = TAPS * CIN;
    const uint64_t dB[2] = {(uint64_t)K, (uint64_t)c.cout};
    const uint64_t sB[1] = {(uint64_t)K * 2};
    const uint32_t bB[2] = {(uint32_t)S::CK, (uint32_t)NPAD};
    if (!make_tmap_f16(&tb, c.w16, 2, dB, sB, bB, nullptr, swz)) return -1;
    const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
    conv_tc_kernel<CIN, NPAD, TAPS, HALO><<<grid, kConvThreads, S::kTotal, s>>>(ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int CIN, int NPAD, int TAPS, bool HALO = false>
int set_attr() {
    return cudaFuncSetAttribute(conv_tc_kernel<CIN, NPAD, TAPS, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ConvCfg<CIN, NPAD, TAPS, HALO>::kTotal) == cudaSuccess ? 0 : -1;
}

}  // namespace

int conv_tc_init_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (const char* e = getenv("EGX_CONV_HALO")) g_halo = atoi(e);
    return set_attr<32, 32, 9>() | set_attr<32, 64, 9>() | set_attr<64, 64, 9>() | set_attr<64, 128, 9>() |
           set_attr<128, 128, 9>() | set_attr<128, 48, 9>() | set_attr<128, 64, 9>() | set_attr<32, 64, 1>() |
           set_attr<64, 128, 1>() | set_attr<32, 32, 9, true>() | set_attr<64, 64, 9, true>();
}

// SE partial-sum slots a conv writes per clip (tiles per clip) for an Ho x Wo output map
int conv_tc_tiles_per_clip(int cin, int cout, int Ho, int Wo) {
    int bw, bh;
    if (g_halo && cin == cout && (cin == 64 || (cin == 32 && g_halo > 1))) pick_halo_patch(Ho, Wo, &bw, &bh);
    else pick_patch(Ho, Wo, &bw, &bh);
    return ((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh);
}

// in: NHWC fp16 (B,Hin,Win,cin).  out: NHWC fp16, or (B,cout,Ho*Wo) fp16 when nchw != 0.
// se_part (optional): [B][tiles_per_clip][cout] per-tile channel sums of the fp32 outputs.
int launch_conv_tc(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw, float* se_part,
                   cudaStream_t s) {
    const int npad = c.cout <= 32 ? 32 : (c.cout <= 48 ? 48 : (c.cout <= 64 ? 64 : 128));
    if (c.cout > 128 || (!nchw && c.cout % 32)) return -1;
    if (g_halo && c.ks == 3 && c.stride == 1 && !nchw) {
        // cin = 32 is bound by the MMA's shared-memory operand reads (N = 32), where the halo variant's extra
        // junk columns cost more than the L2 traffic it saves: measured 546 vs 513 us/launch; EGX_CONV_HALO=2 forces it
        if (c.cin == 32 && c.cout == 32 && g_halo > 1) return launch_one<32, 32, 9, true>(c, in, B, Hin, Win, out, nchw, se_part, s);
        if (c.cin == 64 && c.cout == 64) return launch_one<64, 64, 9, true>(c, in, B, Hin, Win, out, nchw, se_part, s);
    }
#define EGX_CONV_CASE(CI, NP, TP) \
    if (c.cin == CI && npad == NP && c.ks * c.ks == TP) return launch_one<CI, NP, TP, false>(c, in, B, Hin, Win, out, nchw, se_part, s);
    EGX_CONV_CASE(32, 32, 9)
    EGX_CONV_CASE(32, 64, 9)
    EGX_CONV_CASE(64, 64, 9)
    EGX_CONV_CASE(64, 128, 9)
    EGX_CONV_CASE(128, 128, 9)
    EGX_CONV_CASE(128, 48, 9)
    EGX_CONV_CASE(128, 64, 9)
    EGX_CONV_CASE(32, 64, 1)
    EGX_CONV_CASE(64, 128, 1)
#undef EGX_CONV_CASE
    return -1;
}

}  // namespace egx
