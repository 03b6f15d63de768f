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

#include <algorithm>
#include <cstdlib>

namespace egx {

namespace {

using namespace tc;

constexpr int kConvThreads = 384;      // producer + 2 MMA issuers + spare + 2 x 4 epilogue warps

struct ConvTcParams {
    int Ho, Wo;            // output map
    int BW, BH;            // output patch per tile
    int MW;                // row pitch of the M index: BW, or BW + 2 with halo reuse (MW*BH <= 128)
    int tiles_w, tiles_h;
    int tiles_per_clip;
    int num_tiles;         // B * tiles_h * tiles_w
    uint32_t magic_tpc, magic_tw;   // ceil(2^40 / d) >> 8 style magics, see fast_div
    int raster;            // halo kernels: tiles are runs of 128 positions of the clip's padded-width raster (see RASTER)
    uint32_t magic_mw;     // fast_div magic of MW (raster)
    int patch_bytes;       // halo kernels: bytes of one channel-chunk patch in shared memory (1024-aligned)
    int n_stages;          // generic halo kernel: ring stages of kChunks patches
    int b_stages;          // conv128 kernel: depth of the weight ring
    int ks, stride, pad;
    int cout;              // output channels of THIS launch (<= NPAD)
    int n_off, ldc;        // first output channel / channel pitch of the output map (cout > 128 runs as 128-wide slices)
    int relu_first;
    const float* bias;     // may be null
    const float* scale;
    const float* shift;
    __half* out;           // NHWC (B,Ho,Wo,cout) or, if nchw, (B,cout,Ho*Wo)
    float* se_part;        // [B][tiles_h*tiles_w][cout] partial channel sums, or null
    const float* gate;     // [B][2][ldc] folded SE epilogue (g*scale | g*shift) of the tile's clip, or null; with it
    const __half* res;     // the residual map (output geometry, pitch ldc): out = relu(acc*gs + gb + res)
    int contig;            // tile walk of a CTA: 1 = one contiguous run of tiles (consecutive tiles share a clip, so the
                           // gated epilogue re-reads its per-clip gate rarely), 0 = blockIdx.x, +gridDim.x, ...
    int debug;             // EGX_CONV_DEBUG (attribution experiments only): 1 = no epilogue work, 2 = no TMA loads
};

// n / d for n*d < 2^32 with magic = ceil(2^32 / d) (d >= 2), exact by the usual round-up argument
__host__ __device__ __forceinline__ uint32_t make_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

// HALO (stride-1 3x3 with resident weights): instead of nine shifted boxes per tile, ONE box per 64-channel
// chunk brings the (BH+2) x (BW+2) input patch; the M index runs over BH x (BW+2) positions (two junk
// columns per row), so tap (dy,dx) is the SAME shared-memory patch read from a start address shifted by
// (dy*(BW+2)+dx) rows.  TMA and UMMA both apply the 128B/64B swizzle on absolute shared-memory address
// bits, so a row-shifted descriptor still sees the pattern TMA wrote.  L2->SM traffic per tile drops ~7x.
constexpr int kPatchRows = 176;          // >= 128 + 2*(BW+2) + 2 with BW + 2 <= 20
// RASTER (halo kernels whose map is narrow enough): instead of BH x BW patches — which leave 12-25% of the 128 MMA rows
// unused because BH * (BW + 2) rarely comes near 128 and the last patch row of a map is mostly empty — a tile is a run
// of 128 consecutive positions of the clip's raster with padded width MW = Wo + 2 (position g -> pixel (g / MW, g % MW),
// columns Wo, Wo + 1 are the junk every halo tile already carries).  The shared-memory patch is R whole padded map rows
// starting one row above the tile's first row (one 4-D TMA box {channels, MW, R, 1} at x = -1: out-of-bounds zero fill is
// still the padding), and tap (dy, dx) is the patch read from row off + dy * MW + dx, off = the tile's first position
// within its first row.  Tiles per TED clip: 80 -> 72 (layer 1), 22 -> 19 (layer 2), 6 -> 5 (layer 3).
inline int raster_map_rows(int mw) { return (128 + 3 * mw) / mw + 1; }        // R: rows a tile's taps can reach

// Output path of the epilogue
enum { OUT_TMA = 0,      // NHWC fp16 through a swizzled shared-memory staging tile and one TMA box store per tile
       OUT_DIRECT = 1,   // NHWC fp16, each thread stores its pixel's channels (>= 128 contiguous bytes)
       OUT_NCHW = 2 };   // (B,cout,Ho*Wo) fp16 (final conv: the A operand of fc1), bias supported

template <int CIN, int NPAD, int TAPS, bool HALO, int OUT>
struct ConvCfg {
    static constexpr int CK = CIN < 64 ? CIN : 64;           // channels per K block
    static constexpr int kSwz = CK * 2;                      // 64 or 128 byte rows
    static constexpr int kChunks = CIN / CK;
    static constexpr int kNumKb = TAPS * kChunks;            // K blocks per tile
    static constexpr int kABytes = 128 * kSwz;
    static constexpr int kBBytes = ((NPAD * kSwz + 1023) / 1024) * 1024;
    static constexpr bool kResidentB = kNumKb * kBBytes <= (HALO ? 112 : 80) * 1024;
    // K blocks handled per pipeline stage (one barrier round-trip): keep >= 4 MMAs of work per wait
    static constexpr int kKbPerStage = HALO ? kNumKb : ((CK == 32 && TAPS == 9) ? 3 : 1);
    static constexpr int kPatchBytes = kPatchRows * kSwz;     // HALO: the launch may size its patches differently (raster)
    static constexpr int kStageBytes = HALO ? kChunks * kPatchBytes
                                            : kKbPerStage * (kABytes + (kResidentB ? 0 : kBBytes));
    static constexpr int kResBytes = kResidentB ? kNumKb * kBBytes : 0;
    // output staging: [group][buffer] tiles of 128 rows x NPAD fp16 (rows of 64 / 128 bytes, TMA-store swizzle)
    static constexpr int kOutRowBytes = NPAD * 2;
    static constexpr int kOutTileBytes = 128 * kOutRowBytes;
    static constexpr int kOutBytes = OUT == OUT_TMA ? 4 * kOutTileBytes : 0;
    static constexpr int kTailBytes = 256 + 3 * 128 * 4 + 2 * 2 * 4 * 128 * 4 + 1024;   // barriers, params, SE sums, align
    static constexpr int kSmemBudget = 227 * 1024 - 512;
    // the ring takes everything the other regions leave; HALO kernels partition it at run time (p.patch_bytes,
    // p.n_stages: the patch size depends on the map width), the others into kStages stages of kStageBytes
    static constexpr int kMaxStages = 8;
    static constexpr int kRingBytes = (kSmemBudget - kResBytes - kOutBytes - kTailBytes) / 1024 * 1024;
    static constexpr int kStagesRaw = kRingBytes / kStageBytes;
    static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
    static constexpr int kStagesPerTile = kNumKb / kKbPerStage;
    static constexpr int kOutOffset = kResBytes + kRingBytes;
    static constexpr int kBarOffset = kOutOffset + kOutBytes;
    static constexpr int kParOffset = kBarOffset + 256;      // bias | scale | shift, 128 floats each
    static constexpr int kRedOffset = kParOffset + 3 * 128 * 4;   // [group][parity][4 warps][128] floats (SE sums)
    static constexpr int kTotal = kRedOffset + 2 * 2 * 4 * 128 * 4 + 1024;
    static constexpr int kAccStride = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128);
    static constexpr uint32_t kTmemCols = 4 * kAccStride;    // four accumulator buffers
    static constexpr bool kTwoIssuers = HALO;                // see the MMA section
    static_assert(kNumKb % kKbPerStage == 0, "stage must divide the K loop");
    static_assert(kStages >= 2, "not enough shared memory for a pipeline");
    static_assert(!HALO || (TAPS == 9 && kResidentB), "halo reuse needs a 3x3 conv with resident weights");
    static_assert(OUT != OUT_TMA || NPAD == 32 || NPAD == 64, "TMA-store staging rows are one swizzle span");
    static_assert((kStageBytes % 1024) == 0 && (kResBytes % 1024) == 0, "swizzle atoms need 1024-byte alignment");
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 32 fp16 channels (64 bytes) of one residual pixel
struct Res32 { uint4 q[4]; };
__device__ __forceinline__ void load_res32(Res32& r, const __half* p) {
    // 256-bit loads: every 32-byte sector is requested by exactly one instruction
#pragma unroll
    for (int i = 0; i < 2; ++i)
        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r.q[2 * i].x), "=r"(r.q[2 * i].y), "=r"(r.q[2 * i].z), "=r"(r.q[2 * i].w),
                       "=r"(r.q[2 * i + 1].x), "=r"(r.q[2 * i + 1].y), "=r"(r.q[2 * i + 1].z), "=r"(r.q[2 * i + 1].w)
                     : "l"(p + 16 * i));
}
// v = relu(v * gs + gb + res) over one 32-channel chunk; gate_u32: shared-memory address of (gs[128] | gb[128])
__device__ __forceinline__ void gated_residual32(float (&v)[32], uint32_t gate_u32, int nb, const Res32& r) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 sc = lds128(gate_u32 + (nb + 4 * j4) * 4);
        const float4 sh = lds128(gate_u32 + (128 + nb + 4 * j4) * 4);
        const uint4& q = r.q[j4 >> 1];
        const uint32_t w0 = (j4 & 1) ? q.z : q.x, w1 = (j4 & 1) ? q.w : q.y;
        const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
        const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
        v[4 * j4] = fmaxf(fmaf(v[4 * j4], sc.x, sh.x) + r0.x, 0.f);
        v[4 * j4 + 1] = fmaxf(fmaf(v[4 * j4 + 1], sc.y, sh.y) + r0.y, 0.f);
        v[4 * j4 + 2] = fmaxf(fmaf(v[4 * j4 + 2], sc.z, sh.z) + r1.x, 0.f);
        v[4 * j4 + 3] = fmaxf(fmaf(v[4 * j4 + 3], sc.w, sh.w) + r1.y, 0.f);
    }
}

// same with the folded gate of all 32 channels in registers (cout = 32)
__device__ __forceinline__ void gated_residual32_reg(float (&v)[32], const float (&gs)[32], const float (&gb)[32], const Res32& r) {
#pragma unroll
    for (int j2 = 0; j2 < 16; ++j2) {
        const uint4& q = r.q[j2 >> 2];
        const uint32_t w = (j2 & 3) == 0 ? q.x : ((j2 & 3) == 1 ? q.y : ((j2 & 3) == 2 ? q.z : q.w));
        const float2 rv = __half22float2(*reinterpret_cast<const __half2*>(&w));
        v[2 * j2] = fmaxf(fmaf(v[2 * j2], gs[2 * j2], gb[2 * j2]) + rv.x, 0.f);
        v[2 * j2 + 1] = fmaxf(fmaf(v[2 * j2 + 1], gs[2 * j2 + 1], gb[2 * j2 + 1]) + rv.y, 0.f);
    }
}

// epilogue flavours
enum { MODE_PLAIN = 0,    // y = BN([relu](acc + bias))
       MODE_SE = 1,       // same, plus the per-tile channel sums of y (conv1 of an SE block)
       MODE_GATED = 2 };  // out = relu(acc * gs + gb + residual) with the clip's folded SE gate (conv2 of an SE block)

// tile walk of this CTA: tile(n) = t0 + n * tstep for n in [0, my_tiles)
struct TileWalk { int t0, tstep, count; };
__device__ __forceinline__ TileWalk tile_walk(int total, int contig) {
    const int G = gridDim.x, c = blockIdx.x;
    TileWalk w;
    if (contig) {
        w.t0 = (int)(((long long)c * total) / G);
        w.tstep = 1;
        w.count = (int)(((long long)(c + 1) * total) / G) - w.t0;
    } else {
        w.t0 = c; w.tstep = G; w.count = c < total ? (total - c + G - 1) / G : 0;
    }
    return w;
}

template <int CIN, int NPAD, int TAPS, bool HALO, int OUT, int MODE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, ConvTcParams p) {
    using S = ConvCfg<CIN, NPAD, TAPS, HALO, OUT>;
    extern __shared__ unsigned char smem_raw[];
    // keep everything an offset from the __shared__ array so that loads/stores stay in the shared window
    const uint32_t raw_u32 = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
    unsigned char* ring = smem + S::kResBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty = full + S::kMaxStages;
    uint64_t* tmem_full = empty + S::kMaxStages;   // [4]
    // ring geometry: compile-time for the streamed kernels, per launch for the halo kernels
    const int n_stages = HALO ? p.n_stages : S::kStages;
    const int patch_bytes = HALO ? p.patch_bytes : S::kPatchBytes;
    const int stage_bytes = HALO ? S::kChunks * p.patch_bytes : S::kStageBytes;
    uint64_t* tmem_empty = tmem_full + 4;          // [4]
    uint64_t* b_full = tmem_empty + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);
    float* par = reinterpret_cast<float*>(smem + S::kParOffset);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_clip = p.tiles_per_clip;
    constexpr int KS = TAPS == 9 ? 3 : 1;
    const TileWalk walk = tile_walk(p.num_tiles, p.contig);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (OUT == OUT_TMA) prefetch_tmap(&tmO);
        for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<S::kTmemCols>(tmem_ptr);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 128) {
        const int n = threadIdx.x - 128;
        par[n] = (p.bias && n < p.cout) ? p.bias[p.n_off + n] : 0.f;
        par[128 + n] = n < p.cout ? p.scale[p.n_off + n] : 0.f;
        par[256 + n] = n < p.cout ? p.shift[p.n_off + n] : 0.f;
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
                    tma_load_2d(smem + kb * S::kBBytes, &tmB, b_full, (kb / S::kChunks) * CIN + (kb % S::kChunks) * S::CK, p.n_off);
            }
            const uint32_t a_bytes = (uint32_t)p.BW * p.BH * S::kSwz;
            const uint32_t stage_tx = S::kKbPerStage * (a_bytes + (S::kResidentB ? 0u : (uint32_t)NPAD * S::kSwz));
            uint32_t it = 0;
            for (int n = 0; n < walk.count; ++n) {
                const int tile = walk.t0 + n * walk.tstep;
                const int b = tile / tiles_per_clip;
                const int t = tile - b * tiles_per_clip;
                int wi0 = (t % p.tiles_w) * p.BW * p.stride - p.pad;
                int hi0 = (t / p.tiles_w) * p.BH * p.stride - p.pad;
                if (HALO && p.raster) { wi0 = -1; hi0 = (t * 128) / p.MW - 1; }
                if (HALO) {
                    const int st = it % n_stages;
                    mbar_wait(&empty[st], ((it / n_stages) & 1) ^ 1);
                    unsigned char* dst = ring + st * stage_bytes;
                    if (p.debug & 2) { mbar_arrive(&full[st]); ++it; continue; }
                    mbar_expect_tx(&full[st], (uint32_t)S::kChunks * p.MW * (p.BH + 2) * S::kSwz);
#pragma unroll
                    for (int ch = 0; ch < S::kChunks; ++ch)
                        tma_load_4d(dst + ch * patch_bytes, &tmA, &full[st], ch * S::CK, wi0, hi0, b);
                    ++it;
                    continue;
                }
                for (int sg = 0; sg < S::kStagesPerTile; ++sg, ++it) {
                    const int st = it % S::kStages;
                    mbar_wait(&empty[st], ((it / S::kStages) & 1) ^ 1);
                    unsigned char* dst = ring + st * S::kStageBytes;
                    if (p.debug & 2) { mbar_arrive(&full[st]); continue; }
                    mbar_expect_tx(&full[st], stage_tx);
#pragma unroll
                    for (int j = 0; j < S::kKbPerStage; ++j) {
                        const int kb = sg * S::kKbPerStage + j;
                        const int tap = kb / S::kChunks, chunk = kb % S::kChunks;
                        tma_load_4d(dst + j * S::kABytes, &tmA, &full[st], chunk * S::CK, wi0 + tap % KS,
                                    hi0 + tap / KS, b);
                        if (!S::kResidentB)
                            tma_load_2d(dst + S::kKbPerStage * S::kABytes + j * S::kBBytes, &tmB, &full[st],
                                        tap * CIN + chunk * S::CK, p.n_off);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ================= MMA issuers =================
        // Four TMEM accumulator buffers (tile n of this CTA -> buffer n & 3), so a tile's MMAs never wait for the
        // epilogue of the tile two before it.  A 128xNx16 MMA with N <= 64 drains in 40-48 cycles (shared-memory
        // operand reads: (128 + N) rows x 32 B at 128 B/clk) but costs its issuing thread ~45 cycles, and every
        // mbarrier wait ~180 cycles even when the phase is already complete, so a single issuer can never build
        // up a queue and each wait is a bubble in the tensor pipe.  Where one tile is one ring stage (HALO) two
        // warps issue: warp 1 the CTA's even tiles, warp 2 the odd ones; their waits overlap the other's MMAs.
        // (Multi-stage tiles keep one issuer: parity waits on a ring stage may not run several phases ahead.)
        const uint32_t issuer = warp - 1;
        if ((S::kTwoIssuers || issuer == 0) && elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, NPAD);
            constexpr uint32_t kDescHi = smem_desc_hi<S::kSwz>();
            const uint32_t b_lo = smem_desc_lo(smem_u32(smem));                 // resident weights start at smem + 0
            const uint32_t row_off[3] = {0u, (uint32_t)(p.MW * S::kSwz) >> 4, (uint32_t)(2 * p.MW * S::kSwz) >> 4};
            if (S::kResidentB) { mbar_wait(b_full, 0); tc_fence_after(); }
            constexpr uint32_t kStep = S::kTwoIssuers ? 2 : 1;
            for (uint32_t tcount = S::kTwoIssuers ? issuer : 0; (int)tcount < walk.count; tcount += kStep) {
                uint32_t it = tcount * (HALO ? 1 : S::kStagesPerTile);
                const uint32_t acc = tcount & 3;
                mbar_wait(&tmem_empty[acc], ((tcount >> 2) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * S::kAccStride;
                if (HALO) {
                    const int st = it % n_stages;
                    mbar_wait(&full[st], (it / n_stages) & 1);
                    tc_fence_after();
                    uint32_t a_lo = smem_desc_lo(smem_u32(ring + st * stage_bytes));
                    if (p.raster) {                   // the tile starts `off` positions into its first map row
                        const uint32_t tile = (uint32_t)(walk.t0 + (int)tcount * walk.tstep);
                        const uint32_t g0 = (tile - fast_div(tile, p.magic_tpc) * tiles_per_clip) * 128u;
                        a_lo += (g0 - fast_div(g0, p.magic_mw) * p.MW) * (S::kSwz >> 4);
                    }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t a_tap = a_lo + row_off[tap / 3] + (tap % 3) * (S::kSwz >> 4);
#pragma unroll
                        for (int ch = 0; ch < S::kChunks; ++ch) {
#pragma unroll
                            for (int k = 0; k < S::CK / 16; ++k)
                                umma_f16_lo<kDescHi>(d, a_tap + ((ch * patch_bytes + k * 32) >> 4),
                                                     b_lo + (((tap * S::kChunks + ch) * S::kBBytes + k * 32) >> 4), idesc,
                                                     (tap | ch | k) != 0);
                        }
                    }
                    umma_commit(&empty[st]);
                    umma_commit(&tmem_full[acc]);
                    continue;
                }
                for (int sg = 0; sg < S::kStagesPerTile; ++sg, ++it) {
                    const int st = it % S::kStages;
                    mbar_wait(&full[st], (it / S::kStages) & 1);
                    tc_fence_after();
                    const uint32_t a_lo = smem_desc_lo(smem_u32(ring + st * S::kStageBytes));
#pragma unroll
                    for (int j = 0; j < S::kKbPerStage; ++j) {
                        const int kb = sg * S::kKbPerStage + j;
                        const uint32_t bb = S::kResidentB ? b_lo + ((kb * S::kBBytes) >> 4)
                                                          : a_lo + ((S::kKbPerStage * S::kABytes + j * S::kBBytes) >> 4);
#pragma unroll
                        for (int k = 0; k < S::CK / 16; ++k)
                            umma_f16_lo<kDescHi>(d, a_lo + ((j * S::kABytes + k * 32) >> 4), bb + ((k * 32) >> 4), idesc,
                                                 (kb | k) != 0);
                    }
                    umma_commit(&empty[st]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int grp = (warp - 4) >> 2;          // accumulator buffer / tile parity this group drains
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;              // accumulator row = M index of this thread's pixel
        const int ph_ = r / p.MW, pw_ = r % p.MW;
        const bool in_patch = ph_ < p.BH && pw_ < p.BW;
        // row of the staged output tile: the dense BH x BW box, or (raster) the tile position itself
        const int orow = p.raster ? r : ph_ * p.BW + pw_;
        const bool relu_first = p.relu_first != 0;
        const bool has_bias = p.bias != nullptr;
        constexpr bool se = MODE == MODE_SE, gated = MODE == MODE_GATED;
        const uint32_t par_u32 = smem_u32(par);
        const uint32_t red_u32 = smem_u32(smem + S::kRedOffset) + grp * (2 * 512 * 4);
        // staging row of this thread with the TMA-store swizzle (16-byte chunk index XOR row bits)
        const uint32_t stage_u32 = smem_u32(smem + S::kOutOffset) + grp * 2 * S::kOutTileBytes + orow * S::kOutRowBytes;
        const uint32_t swz_x = S::kOutRowBytes == 64 ? ((orow >> 1) & 3) : (orow & 7);
        // copy-out assignment of this thread (tile-invariant): chunk id = i * 128 + r -> (row, 16-byte chunk)
        constexpr int kChunksPerRow = S::kOutRowBytes / 16;
        constexpr int kCopyIters = OUT == OUT_TMA ? kChunksPerRow : 1;
        int cp_ph[kCopyIters], cp_pw[kCopyIters];
        uint32_t cp_src[kCopyIters], cp_dst[kCopyIters];
#pragma unroll
        for (int i = 0; i < kCopyIters; ++i) {
            const int id = i * 128 + r, row = id / kChunksPerRow, ck = id % kChunksPerRow;
            cp_ph[i] = p.raster ? row : (row < p.BH * p.BW ? row / p.BW : -1);     // raster: the tile position
            cp_pw[i] = row % p.BW;
            cp_src[i] = row * S::kOutRowBytes + ((ck ^ (S::kOutRowBytes == 64 ? ((row >> 1) & 3) : (row & 7))) << 4);
            cp_dst[i] = (uint32_t)((cp_ph[i] * p.Wo + cp_pw[i]) * S::kOutRowBytes + ck * 16);
        }
        // NPAD == 32: the folded BatchNorm parameters of all channels live in registers
        float sc_r[32], sh_r[32];      // only live for NPAD == 32
        if (NPAD == 32 && !gated) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { sc_r[j] = par[128 + j]; sh_r[j] = par[256 + j]; }
        }
        Res32 rr = {};
        // gated: the folded gate (gs | gb) of the clip being drained lives in registers (cout = 32) or in one of two
        // shared-memory buffers of this group; it is re-read only when the clip changes
        uint32_t cur_b = 0xffffffffu, gsel = 0;
        uint32_t gate_u32 = red_u32;
        for (uint32_t tcount = grp; (int)tcount < walk.count; tcount += 2) {
            const uint32_t n_local = tcount >> 1;
            const int tile = walk.t0 + (int)tcount * walk.tstep;
            const uint32_t b = fast_div((uint32_t)tile, p.magic_tpc);
            const uint32_t t = (uint32_t)tile - b * tiles_per_clip;
            const uint32_t th = fast_div(t, p.magic_tw);
            const uint32_t tw = t - th * p.tiles_w;
            int ho = th * p.BH + ph_, wo = tw * p.BW + pw_;
            bool valid = in_patch && ho < p.Ho && wo < p.Wo;
            if (p.raster) {
                const uint32_t g = t * 128u + r;
                ho = (int)fast_div(g, p.magic_mw); wo = (int)(g - (uint32_t)ho * p.MW);
                valid = ho < p.Ho && wo < p.Wo;
            }
            const uint32_t par_buf = n_local & 1;
            const uint32_t acc = tcount & 3;
            // gated: everything that does not depend on the accumulator is requested before waiting for it
            const __half* res_pix = nullptr;
            if (gated) {
                if (valid && !(p.debug & 32)) {
                    res_pix = p.res + (((size_t)b * p.Ho + ho) * p.Wo + wo) * p.ldc + p.n_off;
                    load_res32(rr, res_pix);
                }
                if (b != cur_b) {                                        // uniform over the group
                    cur_b = b;
                    const float* gp = p.gate + (size_t)b * 2 * p.ldc + p.n_off;
                    if (NPAD == 32) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(gp) + j4);
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(gp + p.ldc) + j4);
                            sc_r[(4 * j4)] = a.x; sc_r[(4 * j4 + 1)] = a.y;
                            sc_r[(4 * j4 + 2)] = a.z; sc_r[(4 * j4 + 3)] = a.w;
                            sh_r[(4 * j4)] = c4.x; sh_r[(4 * j4 + 1)] = c4.y;
                            sh_r[(4 * j4 + 2)] = c4.z; sh_r[(4 * j4 + 3)] = c4.w;
                        }
                    } else {
                        gsel ^= 1;
                        gate_u32 = red_u32 + gsel * 512 * 4;
                        const float g_s = r < p.cout ? __ldg(gp + r) : 0.f, g_b = r < p.cout ? __ldg(gp + p.ldc + r) : 0.f;
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(gate_u32 + r * 4), "f"(g_s) : "memory");
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(gate_u32 + (128 + r) * 4), "f"(g_b) : "memory");
                        named_bar_sync(1 + grp, 128);
                    }
                }
                const uint32_t nt = (uint32_t)tile + 2 * walk.tstep;      // this group's next tile: residual -> L2
                if ((int)tcount + 2 < walk.count && (in_patch || p.raster) && !(p.debug & 64)) {
                    const uint32_t b2 = fast_div(nt, p.magic_tpc), t2 = nt - b2 * tiles_per_clip;
                    const uint32_t th2 = fast_div(t2, p.magic_tw), tw2 = t2 - th2 * p.tiles_w;
                    int ho2 = th2 * p.BH + ph_, wo2 = tw2 * p.BW + pw_;
                    if (p.raster) {
                        const uint32_t g2 = t2 * 128u + r;
                        ho2 = (int)fast_div(g2, p.magic_mw); wo2 = (int)(g2 - (uint32_t)ho2 * p.MW);
                    }
                    if (ho2 < p.Ho && wo2 < p.Wo) {
                        const __half* np = p.res + (((size_t)b2 * p.Ho + ho2) * p.Wo + wo2) * p.ldc + p.n_off;
                        prefetch_l2(np);
                        if (NPAD > 64) prefetch_l2(np + 64);
                    }
                }
            }
            mbar_wait(&tmem_full[acc], (tcount >> 2) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * S::kAccStride + ((uint32_t)(q * 32) << 16);
            if (p.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < (NPAD + 31) / 32; ++c) {
                float v[32];
                Res32 rn = {};
                if (NPAD > 32 && gated && valid && c + 1 < (NPAD + 31) / 32 && !(p.debug & 32)) load_res32(rn, res_pix + (c + 1) * 32);
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                const int nb = c * 32;
                if (gated) {
                    if (NPAD == 32) gated_residual32_reg(v, sc_r, sh_r, rr);
                    else gated_residual32(v, gate_u32, nb, rr);
                    if (NPAD > 32) rr = rn;
                } else {
                if (has_bias) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 bi = lds128(par_u32 + (nb + 4 * j4) * 4);
                        v[4 * j4] += bi.x; v[4 * j4 + 1] += bi.y; v[4 * j4 + 2] += bi.z; v[4 * j4 + 3] += bi.w;
                    }
                }
                if (NPAD == 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float tv = v[j];
                        if (relu_first) tv = fmaxf(tv, 0.f);
                        v[j] = fmaf(tv, sc_r[j], sh_r[j]);
                    }
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 sc = lds128(par_u32 + (128 + nb + 4 * j4) * 4);
                        const float4 sh = lds128(par_u32 + (256 + nb + 4 * j4) * 4);
                        const float sv[4] = {sc.x, sc.y, sc.z, sc.w}, hv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float tv = v[4 * j4 + e];
                            if (relu_first) tv = fmaxf(tv, 0.f);
                            v[4 * j4 + e] = fmaf(tv, sv[e], hv[e]);
                        }
                    }
                }
                }
                if (OUT == OUT_TMA) {
                    if (in_patch || p.raster) {
                        const uint32_t row = stage_u32 + par_buf * S::kOutTileBytes;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts128(row + (((uint32_t)(c * 4 + j) ^ swz_x) << 4), pack_h2(v[8 * j], v[8 * j + 1]),
                                   pack_h2(v[8 * j + 2], v[8 * j + 3]), pack_h2(v[8 * j + 4], v[8 * j + 5]),
                                   pack_h2(v[8 * j + 6], v[8 * j + 7]));
                    }
                } else if (OUT == OUT_DIRECT) {
                    if (valid) {
                        const size_t pix = ((size_t)b * p.Ho + ho) * p.Wo + wo;
                        __half* o = p.out + pix * p.ldc + p.n_off + nb;     // cout is a multiple of 32 on this path
                        // 256-bit stores: every thread writes whole 32-byte sectors (consecutive lanes = consecutive pixels)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 16 * j),
                                         "r"(pack_h2(v[16 * j], v[16 * j + 1])), "r"(pack_h2(v[16 * j + 2], v[16 * j + 3])),
                                         "r"(pack_h2(v[16 * j + 4], v[16 * j + 5])), "r"(pack_h2(v[16 * j + 6], v[16 * j + 7])),
                                         "r"(pack_h2(v[16 * j + 8], v[16 * j + 9])), "r"(pack_h2(v[16 * j + 10], v[16 * j + 11])),
                                         "r"(pack_h2(v[16 * j + 12], v[16 * j + 13])), "r"(pack_h2(v[16 * j + 14], v[16 * j + 15]))
                                         : "memory");
                    }
                } else {
                    if (valid) {
                        const size_t hw = (size_t)p.Ho * p.Wo;
                        __half* o = p.out + (size_t)b * p.cout * hw + (size_t)ho * p.Wo + wo;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.cout) o[(size_t)(nb + j) * hw] = __float2half_rn(v[j]);
                    }
                }
                if (se) {
                    // column sums over the warp's 32 rows: butterfly "transpose-reduce", 31 shuffles for 32 columns;
                    // lane l ends up holding the sum of column nb + l
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
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
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(red_u32 + (par_buf * 512 + q * 128 + nb + lane) * 4), "f"(v[0]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (OUT == OUT_TMA || se) {
                named_bar_sync(1 + grp, 128);
                if (OUT == OUT_TMA && !(p.debug & 16)) {
                    // copy-out: consecutive threads move consecutive 16-byte chunks of the staged tile, so a warp
                    // writes 512 contiguous bytes of an image row (full sectors, no TMA descriptor work)
                    const uint32_t src0 = smem_u32(smem + S::kOutOffset) + (grp * 2 + par_buf) * S::kOutTileBytes;
                    const int h0 = (int)th * p.BH, w0 = (int)tw * p.BW;
                    __half* tile_out = p.out + (((size_t)b * p.Ho + h0) * p.Wo + w0) * NPAD;
                    if (p.raster) {
                        // a tile's pixels are one contiguous run of the output map minus the two junk columns per row
                        __half* clip_out = p.out + (size_t)b * p.Ho * p.Wo * NPAD;
#pragma unroll
                        for (int i = 0; i < kCopyIters; ++i) {
                            const uint32_t g = t * 128u + (uint32_t)cp_ph[i];
                            const uint32_t gy = fast_div(g, p.magic_mw), gx = g - gy * p.MW;
                            if (gy < (uint32_t)p.Ho && gx < (uint32_t)p.Wo) {
                                uint4 u;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(src0 + cp_src[i]));
                                const int ck = (i * 128 + r) % kChunksPerRow;
                                *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(clip_out) +
                                                          (size_t)(gy * p.Wo + gx) * S::kOutRowBytes + ck * 16) = u;
                            }
                        }
                    } else
#pragma unroll
                    for (int i = 0; i < kCopyIters; ++i) {
                        if (cp_ph[i] >= 0 && h0 + cp_ph[i] < p.Ho && w0 + cp_pw[i] < p.Wo) {
                            uint4 u;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(src0 + cp_src[i]));
                            *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(tile_out) + cp_dst[i]) = u;
                        }
                    }
                }
                if (se) {
                    const int n = q * 32 + lane;
                    if (n < p.cout) {
                        const float* red = reinterpret_cast<const float*>(smem + S::kRedOffset) + grp * 1024 + par_buf * 512;
                        p.se_part[(size_t)tile * p.ldc + p.n_off + n] = (red[n] + red[128 + n]) + (red[256 + n] + red[384 + n]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<S::kTmemCols>(tmem_base);
}

// =================================================================================================================
// 128 -> 128 channel 3x3 stride-1 convolutions (layer 3: 11 of the 29 trunk convolutions).
// The generic kernel streams a 16 KB activation box AND a 16 KB weight block for every 4 MMAs: 128 B/clk per SM, far
// above what L2 can deliver to 148 SMs, and its six 256-cycle ring stages cannot cover the TMA round trip.  Here
//   * activations use halo reuse: per tile one (BH+2) x (BW+2) patch per 64-channel half (2 x 22 KB), all nine taps
//     are row-shifted UMMA descriptors over it;
//   * the 288 KB of weights stream through a 7-deep ring of 16 KB (tap, half) blocks, and every block feeds TWO tiles
//     (two accumulators), so weight traffic per tile halves and a ring stage carries 8 MMAs = 512 tensor cycles;
//   * the K loop runs half-major (all taps of channels 0-63, then of 64-127): a tile's first-half patch is free for
//     the next super-tile's load while the second half is still being multiplied — double buffering without the smem.
// warp 0: TMA producer; warps 1,2: one MMA issuer per tile of the pair; warp 3: TMEM alloc; warps 4-11: two epilogue
// groups (one per tile of the pair), accumulators double-buffered across super-tiles (2 x 2 x 128 TMEM columns).
// =================================================================================================================
// Shared-memory layout (runtime: the patch size depends on the map width): [4 patches: tile 2 x half 2][weight ring:
// b_stages x 16 KB][barriers 256 B][bias|scale|shift 1.5 KB][SE sums / gates 8 KB]; everything 1024-byte aligned.
struct C128 {
    static constexpr int kBBytes = 128 * 128;                     // 128 couts x 64 channels of one tap
    static constexpr int kMaxBStages = 7;
    static constexpr int kTailBytes = 256 + 3 * 128 * 4 + 2 * 2 * 4 * 128 * 4 + 1024;
    static constexpr int kBudget = 227 * 1024;
    static constexpr int kTotal = kBudget;                        // opt-in maximum; a launch asks for what it lays out
    static int b_stages(int patch_bytes) {
        const int n = (kBudget - kTailBytes - 4 * patch_bytes) / kBBytes;
        return n > kMaxBStages ? kMaxBStages : n;
    }
    static int total(int patch_bytes, int stages) { return 4 * patch_bytes + stages * kBBytes + kTailBytes; }
};

template <int MODE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv128_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcParams p) {
    using S = C128;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kPatchBytes = p.patch_bytes, kBStages = p.b_stages;
    const int kBOffset = 4 * kPatchBytes, kBarOffset = kBOffset + kBStages * S::kBBytes;
    const int kParOffset = kBarOffset + 256, kRedOffset = kParOffset + 3 * 128 * 4;
    constexpr int kAOffset = 0;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kBarOffset);    // [tile * 2 + half]
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + S::kMaxBStages;
    uint64_t* tmem_full = b_empty + S::kMaxBStages;     // [buf * 2 + tile]
    uint64_t* tmem_empty = tmem_full + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 4);
    float* par = reinterpret_cast<float*>(smem + kParOffset);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileWalk walk = tile_walk((p.num_tiles + 1) >> 1, p.contig);     // over super-tiles (pairs of tiles)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
        for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(tmem_ptr);
    if (threadIdx.x >= 128 && threadIdx.x < 256) {
        const int n = threadIdx.x - 128;
        par[n] = p.bias ? p.bias[n] : 0.f;
        par[128 + n] = p.scale[n];
        par[256 + n] = p.shift[n];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            const uint32_t a_tx = (uint32_t)p.MW * (p.BH + 2) * 128;
            uint32_t it = 0;
            for (uint32_t n = 0; (int)n < walk.count; ++n) {
                const int u = walk.t0 + (int)n * walk.tstep;
                int wi0[2], hi0[2], bb[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const uint32_t tile = 2 * u + t;
                    const uint32_t b = fast_div(tile, p.magic_tpc), tt = tile - b * p.tiles_per_clip;
                    const uint32_t th = fast_div(tt, p.magic_tw), tw = tt - th * p.tiles_w;
                    wi0[t] = (int)tw * p.BW - 1; hi0[t] = (int)th * p.BH - 1; bb[t] = (int)b;   // clip >= B: zero fill
                    if (p.raster) { wi0[t] = -1; hi0[t] = (int)fast_div(tt * 128u, p.magic_mw) - 1; }
                }
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        mbar_wait(&a_empty[t * 2 + ch], (n & 1) ^ 1);
                        mbar_expect_tx(&a_full[t * 2 + ch], a_tx);
                        tma_load_4d(smem + kAOffset + (t * 2 + ch) * kPatchBytes, &tmA, &a_full[t * 2 + ch], ch * 64,
                                    wi0[t], hi0[t], bb[t]);
                    }
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int st = it % kBStages;
                        mbar_wait(&b_empty[st], ((it / kBStages) & 1) ^ 1);
                        mbar_expect_tx(&b_full[st], S::kBBytes);
                        tma_load_2d(smem + kBOffset + st * S::kBBytes, &tmB, &b_full[st], tap * 128 + ch * 64, 0);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ================= MMA issuers: warp 1 -> first tile of the pair, warp 2 -> second =================
        const uint32_t t = warp - 1;
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, 128);
            constexpr uint32_t kDescHi = smem_desc_hi<128>();
            const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem + kBOffset));
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem + kAOffset + t * 2 * kPatchBytes));
            const uint32_t row_off[3] = {0u, (uint32_t)(p.MW * 128) >> 4, (uint32_t)(2 * p.MW * 128) >> 4};
            uint32_t it = 0;
            for (uint32_t n = 0; (int)n < walk.count; ++n) {
                const uint32_t buf = n & 1;
                mbar_wait(&tmem_empty[buf * 2 + t], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (buf * 2 + t) * 128;
                uint32_t off_units = 0;               // raster: the tile starts `off` positions into its first map row
                if (p.raster) {
                    const uint32_t tile = 2u * (uint32_t)(walk.t0 + (int)n * walk.tstep) + t;
                    const uint32_t g0 = (tile - fast_div(tile, p.magic_tpc) * p.tiles_per_clip) * 128u;
                    off_units = (g0 - fast_div(g0, p.magic_mw) * p.MW) * (128u >> 4);
                }
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    mbar_wait(&a_full[t * 2 + ch], n & 1);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + ((ch * kPatchBytes) >> 4) + off_units;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int st = it % kBStages;
                        mbar_wait(&b_full[st], (it / kBStages) & 1);
                        tc_fence_after();
                        const uint32_t a_tap = a_lo + row_off[tap / 3] + (tap % 3) * (128 >> 4);
                        const uint32_t b_lo = b_lo0 + ((st * S::kBBytes) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (tap == 0 && k == 0 && ch == 0) umma_f16_lo<kDescHi>(d, a_tap, b_lo, idesc, false);
                            else umma_f16_lo<kDescHi>(d, a_tap + 2 * k, b_lo + 2 * k, idesc, true);
                        }
                        umma_commit(&b_empty[st]);
                    }
                    umma_commit(&a_empty[t * 2 + ch]);
                }
                umma_commit(&tmem_full[buf * 2 + t]);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: group g drains the accumulator of tile g of every pair =================
        const int grp = (warp - 4) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int ph_ = r / p.MW, pw_ = r % p.MW;
        const bool in_patch = ph_ < p.BH && pw_ < p.BW;
        const bool relu_first = p.relu_first != 0;
        const bool has_bias = p.bias != nullptr;
        constexpr bool se = MODE == MODE_SE, gated = MODE == MODE_GATED;
        const uint32_t par_u32 = smem_u32(par);
        const uint32_t red_u32 = smem_u32(smem + kRedOffset) + grp * (2 * 512 * 4);
        Res32 rr = {};
        uint32_t cur_b = 0xffffffffu, gsel = 0;      // clip whose folded gate sits in shared-memory buffer gsel
        uint32_t gate_u32 = red_u32;
        for (uint32_t n = 0; (int)n < walk.count; ++n) {
            const int u = walk.t0 + (int)n * walk.tstep;
            const uint32_t tile = 2 * u + grp;
            const uint32_t b = fast_div(tile, p.magic_tpc), tt = tile - b * p.tiles_per_clip;
            const uint32_t th = fast_div(tt, p.magic_tw), tw = tt - th * p.tiles_w;
            int ho = th * p.BH + ph_, wo = tw * p.BW + pw_;
            const bool live = tile < (uint32_t)p.num_tiles;
            bool valid = live && in_patch && ho < p.Ho && wo < p.Wo;
            if (p.raster) {
                const uint32_t g = tt * 128u + r;
                ho = (int)fast_div(g, p.magic_mw); wo = (int)(g - (uint32_t)ho * p.MW);
                valid = live && ho < p.Ho && wo < p.Wo;
            }
            const uint32_t buf = n & 1, par_buf = n & 1;
            const __half* res_pix = nullptr;
            if (gated) {
                if (valid && !(p.debug & 32)) {
                    res_pix = p.res + (((size_t)b * p.Ho + ho) * p.Wo + wo) * 128;
                    load_res32(rr, res_pix);
                }
                if (live && b != cur_b) {                                 // uniform over the group
                    cur_b = b;
                    gsel ^= 1;
                    gate_u32 = red_u32 + gsel * 512 * 4;
                    const float* gp = p.gate + (size_t)b * 256 + r;
                    const float g_s = __ldg(gp), g_b = __ldg(gp + 128);
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(gate_u32 + r * 4), "f"(g_s) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(gate_u32 + (128 + r) * 4), "f"(g_b) : "memory");
                    named_bar_sync(1 + grp, 128);
                }
                const uint32_t nt = 2 * (u + walk.tstep) + grp;           // this group's next tile: residual -> L2
                if ((int)n + 1 < walk.count && nt < (uint32_t)p.num_tiles && (in_patch || p.raster) && !(p.debug & 64)) {
                    const uint32_t b2 = fast_div(nt, p.magic_tpc), t2 = nt - b2 * p.tiles_per_clip;
                    const uint32_t th2 = fast_div(t2, p.magic_tw), tw2 = t2 - th2 * p.tiles_w;
                    int ho2 = th2 * p.BH + ph_, wo2 = tw2 * p.BW + pw_;
                    if (p.raster) {
                        const uint32_t g2 = t2 * 128u + r;
                        ho2 = (int)fast_div(g2, p.magic_mw); wo2 = (int)(g2 - (uint32_t)ho2 * p.MW);
                    }
                    if (ho2 < p.Ho && wo2 < p.Wo) {
                        const __half* np = p.res + (((size_t)b2 * p.Ho + ho2) * p.Wo + wo2) * 128;
                        prefetch_l2(np);
                        prefetch_l2(np + 64);
                    }
                }
            }
            mbar_wait(&tmem_full[buf * 2 + grp], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (buf * 2 + grp) * 128 + ((uint32_t)(q * 32) << 16);
            __half* o = p.out + (((size_t)b * p.Ho + ho) * p.Wo + wo) * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                Res32 rn = {};
                if (gated && valid && c < 3 && !(p.debug & 32)) load_res32(rn, res_pix + (c + 1) * 32);
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                const int nb = c * 32;
                if (gated) {
                    gated_residual32(v, gate_u32, nb, rr);
                    rr = rn;
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 sc = lds128(par_u32 + (128 + nb + 4 * j4) * 4);
                        const float4 sh = lds128(par_u32 + (256 + nb + 4 * j4) * 4);
                        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_bias) bi = lds128(par_u32 + (nb + 4 * j4) * 4);
                        const float sv[4] = {sc.x, sc.y, sc.z, sc.w}, hv[4] = {sh.x, sh.y, sh.z, sh.w}, bv[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float tv = v[4 * j4 + e] + bv[e];
                            if (relu_first) tv = fmaxf(tv, 0.f);
                            v[4 * j4 + e] = fmaf(tv, sv[e], hv[e]);
                        }
                    }
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + nb + 16 * j),
                                     "r"(pack_h2(v[16 * j], v[16 * j + 1])), "r"(pack_h2(v[16 * j + 2], v[16 * j + 3])),
                                     "r"(pack_h2(v[16 * j + 4], v[16 * j + 5])), "r"(pack_h2(v[16 * j + 6], v[16 * j + 7])),
                                     "r"(pack_h2(v[16 * j + 8], v[16 * j + 9])), "r"(pack_h2(v[16 * j + 10], v[16 * j + 11])),
                                     "r"(pack_h2(v[16 * j + 12], v[16 * j + 13])), "r"(pack_h2(v[16 * j + 14], v[16 * j + 15]))
                                     : "memory");
                }
                if (se) {
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
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
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(red_u32 + (par_buf * 512 + q * 128 + nb + lane) * 4), "f"(v[0]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf * 2 + grp]);
            if (se) {
                named_bar_sync(1 + grp, 128);
                if (live) {
                    const float* red = reinterpret_cast<const float*>(smem + kRedOffset) + grp * 1024 + par_buf * 512;
                    p.se_part[(size_t)tile * 128 + r] = (red[r] + red[128 + r]) + (red[256 + r] + red[384 + r]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<512>(tmem_base);
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

// Raster tiling applies when the padded rows a tile can touch fit the kernel's patch buffer
// bytes of shared memory a raster tile's patches may take in the kernel that runs (cin -> cin, 3x3): half of the generic
// halo kernel's ring (>= 2 stages), or what leaves conv128 a 3-deep weight ring
int raster_stage_cap(int cin) {
    if (cin == 32) return ConvCfg<32, 32, 9, true, OUT_TMA>::kRingBytes / 2;     // the smaller ring of the two output paths
    if (cin == 64) return ConvCfg<64, 64, 9, true, OUT_TMA>::kRingBytes / 2;
    return 0;
}
bool raster_geometry(int cin, int Ho, int Wo, int* mw, int* rows) {
    *mw = Wo + 2;
    *rows = raster_map_rows(*mw);
    if (*mw > 256 || *rows > 256 || Ho < 1) return false;
    const int ck = cin < 64 ? cin : 64, chunks = cin / ck;
    const int patch = (*rows * *mw * ck * 2 + 1023) / 1024 * 1024;
    if (cin >= 128) return (C128::kBudget - C128::kTailBytes - 4 * patch) / C128::kBBytes >= 3;
    return chunks * patch <= raster_stage_cap(cin);
}
void set_raster(ConvTcParams& p, int B, int mw, int rows) {
    p.raster = 1;
    p.BW = p.Wo; p.MW = mw; p.BH = rows - 2;         // the TMA box is {channels, MW, BH + 2, 1}
    p.tiles_w = 1;
    p.tiles_per_clip = (p.Ho * mw + 127) / 128;
    p.tiles_h = p.tiles_per_clip;
    p.num_tiles = B * p.tiles_per_clip;
    p.magic_mw = make_magic((uint32_t)mw);
}

bool use_raster_fwd(int cin, int Ho, int Wo);

int g_num_sms = 0;
int g_debug = 0;
int g_raster = 1;       // EGX_CONV_RASTER (attribution builds) bits: 1 = raster tiles for the 64 / 128-channel halo kernels, 2 = 32-channel
int g_out_direct = 1;   // EGX_CONV_OUT bit 0: 32->32 halo kernel stores straight from registers, bit 1: 64->64 too, bit 2: 64->64 gated
int g_contig = 0;    // EGX_CONV_CONTIG: 1 = contiguous tile runs per CTA, 0 = strided walk (default: measured faster, the CTAs share halos in L2)
int g_halo = 15;      // EGX_CONV_HALO bits: 1 = 64->64 convs, 2 = 32->32, 4 = 128->128, 8 = final conv (128 -> <= 48, NCHW out)

template <int CIN, int NPAD, int TAPS, bool HALO, int OUT, int MODE>
int launch_one(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, float* se_part, int n_off, cudaStream_t s,
               const float* gate, const __half* res) {
    using S = ConvCfg<CIN, NPAD, TAPS, HALO, OUT>;
    ConvTcParams p;
    p.ks = c.ks; p.stride = c.stride; p.pad = c.ks / 2;
    p.Ho = (Hin + 2 * p.pad - c.ks) / c.stride + 1;
    p.Wo = (Win + 2 * p.pad - c.ks) / c.stride + 1;
    if (HALO) pick_halo_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    else pick_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    p.MW = HALO ? p.BW + 2 : p.BW;
    p.tiles_w = (p.Wo + p.BW - 1) / p.BW;
    p.tiles_h = (p.Ho + p.BH - 1) / p.BH;
    p.tiles_per_clip = p.tiles_w * p.tiles_h;
    p.num_tiles = B * p.tiles_per_clip;
    p.raster = 0; p.magic_mw = 0; p.b_stages = 0;
    {
        int mw, rows;
        if (HALO && OUT != OUT_NCHW && CIN == NPAD && CIN <= 64 && use_raster_fwd(CIN, p.Ho, p.Wo) && raster_geometry(CIN, p.Ho, p.Wo, &mw, &rows))
            set_raster(p, B, mw, rows);
        if (HALO && OUT == OUT_NCHW && use_raster_fwd(CIN, p.Ho, p.Wo)) {
            // final conv (128 -> 34 frames): raster tiles when two stages of whole-row patches fit beside the weights
            mw = p.Wo + 2; rows = raster_map_rows(mw);
            const int patch = (rows * mw * S::kSwz + 1023) / 1024 * 1024;
            if (mw <= 256 && rows <= 256 && 2 * S::kChunks * patch <= S::kRingBytes) set_raster(p, B, mw, rows);
        }
    }
    // rows a tap-shifted 128-row operand can reach: the whole box, and (patch tiles) 128 + 2 MW + 2 rows from its start
    p.patch_bytes = HALO ? (std::max(p.MW * (p.BH + 2), p.raster ? 0 : 128 + 2 * p.MW + 2) * S::kSwz + 1023) / 1024 * 1024
                         : S::kPatchBytes;
    p.n_stages = HALO ? std::min<int>(S::kMaxStages, S::kRingBytes / (S::kChunks * p.patch_bytes)) : S::kStages;
    if (p.n_stages < 2) return -1;
    if ((uint64_t)p.num_tiles * (uint64_t)p.tiles_per_clip >= (1ull << 32)) return -1;   // fast_div exactness
    p.magic_tpc = make_magic((uint32_t)p.tiles_per_clip);
    p.magic_tw = make_magic((uint32_t)p.tiles_w);
    p.cout = c.cout - n_off < 128 ? c.cout - n_off : 128; p.n_off = n_off; p.ldc = c.cout; p.relu_first = c.relu_first;
    p.bias = c.bias; p.scale = c.scale; p.shift = c.shift;
    p.out = out; p.se_part = se_part; p.debug = g_debug;
    p.gate = gate; p.res = res; p.contig = g_contig;

    CUtensorMap ta, tb, to;
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
    if (OUT == OUT_TMA) {
        if (c.cout != NPAD || n_off) return -1;
        const uint64_t dO[4] = {(uint64_t)NPAD, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)B};
        const uint64_t sO[3] = {(uint64_t)NPAD * 2, (uint64_t)p.Wo * NPAD * 2, (uint64_t)p.Ho * p.Wo * NPAD * 2};
        const uint32_t bO[4] = {(uint32_t)NPAD, (uint32_t)p.BW, (uint32_t)p.BH, 1};
        if (!make_tmap_f16(&to, out, 4, dO, sO, bO, nullptr,
                           NPAD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
            return -1;
    } else {
        to = ta;
    }
    const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
    conv_tc_kernel<CIN, NPAD, TAPS, HALO, OUT, MODE><<<grid, kConvThreads, S::kTotal, s>>>(ta, tb, to, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_conv128(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, float* se_part, cudaStream_t s,
                   const float* gate, const __half* res) {
    ConvTcParams p;
    p.ks = 3; p.stride = 1; p.pad = 1;
    p.Ho = Hin; p.Wo = Win;
    pick_halo_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    p.MW = p.BW + 2;
    p.tiles_w = (p.Wo + p.BW - 1) / p.BW;
    p.tiles_h = (p.Ho + p.BH - 1) / p.BH;
    p.tiles_per_clip = p.tiles_w * p.tiles_h;
    p.num_tiles = B * p.tiles_per_clip;
    p.raster = 0; p.magic_mw = 0;
    {
        int mw, rows;
        if (use_raster_fwd(128, p.Ho, p.Wo) && raster_geometry(128, p.Ho, p.Wo, &mw, &rows)) set_raster(p, B, mw, rows);
    }
    p.patch_bytes = ((p.raster ? (p.BH + 2) * p.MW : kPatchRows) * 128 + 1023) / 1024 * 1024;
    p.b_stages = C128::b_stages(p.patch_bytes);
    if (p.b_stages < 3) return -1;
    if ((uint64_t)(p.num_tiles + 1) * (uint64_t)p.tiles_per_clip >= (1ull << 32)) return -1;
    p.magic_tpc = make_magic((uint32_t)p.tiles_per_clip);
    p.magic_tw = make_magic((uint32_t)p.tiles_w);
    p.cout = 128; p.n_off = 0; p.ldc = 128; p.relu_first = c.relu_first;
    p.bias = c.bias; p.scale = c.scale; p.shift = c.shift;
    p.out = out; p.se_part = se_part; p.debug = g_debug;
    p.gate = gate; p.res = res; p.contig = g_contig;
    CUtensorMap ta, tb;
    const uint64_t dA[4] = {128, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    const uint64_t sA[3] = {256, (uint64_t)Win * 256, (uint64_t)Hin * Win * 256};
    const uint32_t bA[4] = {64, (uint32_t)p.MW, (uint32_t)(p.BH + 2), 1};
    if (!make_tmap_f16(&ta, in, 4, dA, sA, bA, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    const uint64_t dB[2] = {9 * 128, 128};
    const uint64_t sB[1] = {9 * 128 * 2};
    const uint32_t bB[2] = {64, 128};
    if (!make_tmap_f16(&tb, c.w16, 2, dB, sB, bB, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    const int n_super = (p.num_tiles + 1) / 2;
    const int grid = n_super < g_num_sms ? n_super : g_num_sms;
    const int smem_bytes = C128::total(p.patch_bytes, p.b_stages);
    if (gate) conv128_tc_kernel<MODE_GATED><<<grid, kConvThreads, smem_bytes, s>>>(ta, tb, p);
    else if (se_part) conv128_tc_kernel<MODE_SE><<<grid, kConvThreads, smem_bytes, s>>>(ta, tb, p);
    else conv128_tc_kernel<MODE_PLAIN><<<grid, kConvThreads, smem_bytes, s>>>(ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int CIN, int NPAD, int TAPS, bool HALO, int OUT, int MODE>
int set_attr() {
    return cudaFuncSetAttribute(conv_tc_kernel<CIN, NPAD, TAPS, HALO, OUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ConvCfg<CIN, NPAD, TAPS, HALO, OUT>::kTotal) == cudaSuccess ? 0 : -1;
}

}  // namespace

// every instantiation: (cin, npad, taps, halo, out)
#define EGX_CONV_INSTANCES(X)                                                                         \
    X(32, 32, 9, true, OUT_TMA) X(32, 32, 9, false, OUT_TMA) X(64, 64, 9, true, OUT_TMA)              \
    X(64, 64, 9, false, OUT_TMA) X(32, 64, 9, false, OUT_TMA) X(32, 64, 1, false, OUT_TMA)            \
    X(64, 128, 9, false, OUT_DIRECT) X(64, 128, 1, false, OUT_DIRECT) X(128, 128, 9, false, OUT_DIRECT) \
    X(32, 32, 9, true, OUT_DIRECT) X(64, 64, 9, true, OUT_DIRECT)                                     \
    X(256, 128, 9, false, OUT_DIRECT) X(128, 128, 1, false, OUT_DIRECT)                               \
    X(128, 48, 9, false, OUT_NCHW) X(128, 64, 9, false, OUT_NCHW) X(128, 48, 9, true, OUT_NCHW)

// conv1 of an SE block (sums its output): stride-1 halo kernels, the stride-2 first blocks of a stage, 256-wide slices
#define EGX_CONV_SE_INSTANCES(X)                                                                       \
    X(32, 32, 9, true, OUT_TMA) X(32, 32, 9, false, OUT_TMA) X(64, 64, 9, true, OUT_TMA)              \
    X(64, 64, 9, false, OUT_TMA) X(32, 64, 9, false, OUT_TMA) X(64, 128, 9, false, OUT_DIRECT)        \
    X(128, 128, 9, false, OUT_DIRECT) X(32, 32, 9, true, OUT_DIRECT) X(64, 64, 9, true, OUT_DIRECT)   \
    X(256, 128, 9, false, OUT_DIRECT)
// conv2 of an SE block (gated residual epilogue): always cin == cout, stride 1
#define EGX_CONV_GATED_INSTANCES(X)                                                                    \
    X(32, 32, 9, true, OUT_TMA) X(32, 32, 9, false, OUT_TMA) X(64, 64, 9, true, OUT_TMA)              \
    X(64, 64, 9, false, OUT_TMA) X(128, 128, 9, false, OUT_DIRECT) X(32, 32, 9, true, OUT_DIRECT)     \
    X(64, 64, 9, true, OUT_DIRECT) X(256, 128, 9, false, OUT_DIRECT)

int conv_tc_init_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_halo = env_switch("EGX_CONV_HALO", g_halo);
    g_debug = env_switch("EGX_CONV_DEBUG", 0);
    g_out_direct = env_switch("EGX_CONV_OUT", g_out_direct);
    g_contig = env_switch("EGX_CONV_CONTIG", g_contig);
    g_raster = env_switch("EGX_CONV_RASTER", g_raster);
    int rc = 0;
    rc |= cudaFuncSetAttribute(conv128_tc_kernel<MODE_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C128::kTotal) == cudaSuccess ? 0 : -1;
    rc |= cudaFuncSetAttribute(conv128_tc_kernel<MODE_SE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C128::kTotal) == cudaSuccess ? 0 : -1;
    rc |= cudaFuncSetAttribute(conv128_tc_kernel<MODE_GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C128::kTotal) == cudaSuccess ? 0 : -1;
#define X(CI, NP, TP, HL, OU) rc |= set_attr<CI, NP, TP, HL, OU, MODE_PLAIN>();
    EGX_CONV_INSTANCES(X)
#undef X
#define X(CI, NP, TP, HL, OU) rc |= set_attr<CI, NP, TP, HL, OU, MODE_SE>();
    EGX_CONV_SE_INSTANCES(X)
#undef X
#define X(CI, NP, TP, HL, OU) rc |= set_attr<CI, NP, TP, HL, OU, MODE_GATED>();
    EGX_CONV_GATED_INSTANCES(X)
#undef X
    return rc;
}

static bool use_halo(int cin, int cout, int ks, int stride, int nchw) {
    // final conv of the TED generator (128 -> 34 frames, N padded to 48): 108 KB of weights stay resident
    if (ks == 3 && stride == 1 && nchw && cin == 128 && cout <= 48 && (g_halo & 8)) return true;
    if (ks != 3 || stride != 1 || nchw || cin != cout) return false;
    return (cin == 64 && g_halo >= 1) || (cin == 32 && (g_halo & 2)) || (cin == 128 && (g_halo & 4));
}

// SE partial-sum slots a conv writes per clip (tiles per clip) for an Ho x Wo output map
// Layer 1 keeps patch tiles: at MW = 72 a raster tile spans 1.8 map rows but loads five (2.8x its outputs against 1.4x for
// a 14 x 8 patch), and that stage is already close to its memory roofline (measured: 8.4 -> 8.9 ms/step with raster tiles)
static bool use_raster(int cin, int Ho, int Wo) {
    int mw, rows;
    return (g_raster & (cin >= 64 ? 1 : 2)) && raster_geometry(cin, Ho, Wo, &mw, &rows);
}

namespace { bool use_raster_fwd(int cin, int Ho, int Wo) { return use_raster(cin, Ho, Wo); } }

int conv_tc_tiles_per_clip(int cin, int cout, int Ho, int Wo) {
    int bw, bh;
    if (use_halo(cin, cout, 3, 1, 0)) {
        if (use_raster(cin, Ho, Wo)) return (Ho * (Wo + 2) + 127) / 128;
        pick_halo_patch(Ho, Wo, &bw, &bh);
    } else pick_patch(Ho, Wo, &bw, &bh);
    return ((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh);
}

// in: NHWC fp16 (B,Hin,Win,cin).  out: NHWC fp16, or (B,cout,Ho*Wo) fp16 when nchw != 0.
// se_part (optional): [B][tiles_per_clip][cout] per-tile channel sums of the fp32 outputs.
int launch_conv_tc(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw, float* se_part,
                   cudaStream_t s, const float* gate, const __half* res) {
    // the gated epilogue replaces conv2's own affine (folded into `gate`), shares shared memory with the SE sums
    if ((gate != nullptr) != (res != nullptr) || (gate && (se_part || nchw || c.relu_first || c.cout % 32))) return -1;
    const int npad = c.cout <= 32 ? 32 : (c.cout <= 48 ? 48 : (c.cout <= 64 ? 64 : 128));
    if ((c.cout > 128 && (nchw || c.cout % 128)) || (!nchw && c.cout % 32)) return -1;
    const bool halo = use_halo(c.cin, c.cout, c.ks, c.stride, nchw);
    if (halo && c.cin == 128 && !nchw) return launch_conv128(c, in, B, Hin, Win, out, se_part, s, gate, res);
    const int out_mode = nchw ? OUT_NCHW : ((npad <= 64 && !(halo && ((c.cin == 32 && (g_out_direct & 1)) || (c.cin == 64 && (g_out_direct & (gate ? 4 : 2)))))) ? OUT_TMA : OUT_DIRECT);
#define X_MODE(CI, NP, TP, HL, OU, MD)                                                          \
    if (c.cin == CI && npad == NP && c.ks * c.ks == TP && halo == HL && out_mode == OU && mode == MD) \
    {                                                                                           \
        int n = 0;                                                                              \
        for (int n_off = 0; n_off < c.cout; n_off += 128) {                                     \
            if (launch_one<CI, NP, TP, HL, OU, MD>(c, in, B, Hin, Win, out, se_part, n_off, s, gate, res) < 0) return -1; \
            ++n;                                                                                \
        }                                                                                       \
        return n;                                                                               \
    }
    const int mode = gate ? MODE_GATED : (se_part ? MODE_SE : MODE_PLAIN);
#define X(CI, NP, TP, HL, OU) X_MODE(CI, NP, TP, HL, OU, MODE_PLAIN)
    EGX_CONV_INSTANCES(X)
#undef X
#define X(CI, NP, TP, HL, OU) X_MODE(CI, NP, TP, HL, OU, MODE_SE)
    EGX_CONV_SE_INSTANCES(X)
#undef X
#define X(CI, NP, TP, HL, OU) X_MODE(CI, NP, TP, HL, OU, MODE_GATED)
    EGX_CONV_GATED_INSTANCES(X)
#undef X
#undef X_MODE
    return -1;
}

}  // namespace egx
