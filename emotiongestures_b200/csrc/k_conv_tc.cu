// K3 — 3x3 / 1x1 convolutions of the SE-ResNet trunk as implicit GEMM on tcgen05.
//   M = output pixels (one BH x BW spatial patch of one clip per CTA, <= 128 rows)
//   N = cout (32 / 64 / 128, final conv 34 padded to 48),  K = taps * cin
// Follows Full_model/ResNetBlocks.py:24-30 (conv-ReLU-BN / conv-BN), Full_model/ResNetSE34V2.py:43-47
// (1x1 stride-2 downsample + BN) and Full_model/Models.py:121-122 (final conv + BN).
//
// Activations are NHWC fp16.  No im2col buffer exists anywhere: for every (tap, 64-channel chunk) the
// producer warp issues ONE 4-D TMA box load {channels, BW, BH, 1} whose W/H start coordinate is shifted by
// the tap (and may be -1 or run past the edge): TMA's out-of-bounds zero fill IS the convolution padding,
// and its element stride is the convolution stride.  The box lands in shared memory as BH*BW dense rows of
// 128 B (64 B for cin = 32) in the SWIZZLE_128B (64B) pattern the UMMA descriptor expects, i.e. directly as
// the K-major A operand.  Weights [cout][tap*cin] are the K-major B operand (2-D TMA).  fp32 accumulators
// live in TMEM; the epilogue (4 warps, one pixel per thread) applies bias / ReLU / folded BatchNorm and
// stores NHWC fp16 (or the (B, F, H*W) layout the fc1 GEMM consumes, for the final conv).
#include "egx_common.cuh"
#include "tc_common.cuh"

namespace egx {

namespace {

using namespace tc;

constexpr int kConvStages = 4;
constexpr int kConvThreads = 192;

struct ConvTcParams {
    int Ho, Wo;            // output map
    int BW, BH;            // output patch per CTA (BW*BH <= 128)
    int tiles_w, tiles_h;
    int ks, stride, pad;
    int cout;
    int relu_first;
    const float* bias;     // may be null
    const float* scale;
    const float* shift;
    __half* out;           // NHWC (B,Ho,Wo,cout) or, if nchw, (B,cout,Ho*Wo)
    int nchw;
};

template <int CIN, int NPAD>
struct ConvSmem {
    static constexpr int CK = CIN < 64 ? CIN : 64;           // channels per K block
    static constexpr int kSwz = CK * 2;                      // 64 or 128 byte rows
    static constexpr int kChunks = CIN / CK;
    static constexpr int kABytes = 128 * kSwz;
    static constexpr int kBBytes = ((NPAD * kSwz + 1023) / 1024) * 1024;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kConvStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + 128 + 1024;
    static constexpr uint32_t kTmemCols = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128);
};

template <int CIN, int NPAD>
__global__ void __launch_bounds__(kConvThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcParams p) {
    using S = ConvSmem<CIN, NPAD>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty = full + kConvStages;
    uint64_t* tmem_full = empty + kConvStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = blockIdx.x % p.tiles_w;
    const int th = (blockIdx.x / p.tiles_w) % p.tiles_h;
    const int b = blockIdx.x / (p.tiles_w * p.tiles_h);
    const int wo0 = tw * p.BW, ho0 = th * p.BH;
    const int num_kb = p.ks * p.ks * S::kChunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < kConvStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<S::kTmemCols>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t a_bytes = (uint32_t)p.BW * p.BH * S::kSwz;
            const uint32_t b_bytes = (uint32_t)NPAD * S::kSwz;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % kConvStages;
                const uint32_t ph = (kb / kConvStages) & 1;
                const int tap = kb / S::kChunks, chunk = kb % S::kChunks;
                const int dy = tap / p.ks, dx = tap % p.ks;
                mbar_wait(&empty[st], ph ^ 1);
                unsigned char* a = smem + st * S::kStageBytes;
                mbar_expect_tx(&full[st], a_bytes + b_bytes);
                tma_load_4d(a, &tmA, &full[st], chunk * S::CK, wo0 * p.stride + dx - p.pad,
                            ho0 * p.stride + dy - p.pad, b);
                tma_load_2d(a + S::kABytes, &tmB, &full[st], tap * CIN + chunk * S::CK, 0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, NPAD);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % kConvStages;
                const uint32_t ph = (kb / kConvStages) & 1;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a = smem_u32(smem + st * S::kStageBytes);
                const uint32_t bb = a + S::kABytes;
#pragma unroll
                for (int k = 0; k < S::CK / 16; ++k)
                    umma_f16(tmem_base, make_smem_desc<S::kSwz>(a + k * 32), make_smem_desc<S::kSwz>(bb + k * 32),
                             idesc, (kb | k) != 0);
                umma_commit(&empty[st]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int ph_ = r / p.BW, pw_ = r % p.BW;
        const int ho = ho0 + ph_, wo = wo0 + pw_;
        const bool valid = r < p.BW * p.BH && ho < p.Ho && wo < p.Wo;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const size_t pix = ((size_t)b * p.Ho + ho) * p.Wo + wo;
#pragma unroll 1
        for (int c = 0; c < (NPAD + 31) / 32; ++c) {
            float v[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
            if (valid) {
                const int nb = c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nb + j;
                    float t = v[j];
                    if (n < p.cout) {
                        if (p.bias) t += __ldg(p.bias + n);
                        if (p.relu_first) t = fmaxf(t, 0.f);
                        t = fmaf(t, __ldg(p.scale + n), __ldg(p.shift + n));
                    }
                    v[j] = t;
                }
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
        // fewest tiles; ties go to the wider patch (longer contiguous stores)
        if (best < 0 || tiles <= best) { best = tiles; *bw = w; *bh = h; }
    }
}

template <int CIN, int NPAD>
int launch_one(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw, cudaStream_t s) {
    using S = ConvSmem<CIN, NPAD>;
    ConvTcParams p;
    p.ks = c.ks; p.stride = c.stride; p.pad = c.ks / 2;
    p.Ho = (Hin + 2 * p.pad - c.ks) / c.stride + 1;
    p.Wo = (Win + 2 * p.pad - c.ks) / c.stride + 1;
    pick_patch(p.Ho, p.Wo, &p.BW, &p.BH);
    p.tiles_w = (p.Wo + p.BW - 1) / p.BW;
    p.tiles_h = (p.Ho + p.BH - 1) / p.BH;
    p.cout = c.cout; p.relu_first = c.relu_first;
    p.bias = c.bias; p.scale = c.scale; p.shift = c.shift;
    p.out = out; p.nchw = nchw;

    CUtensorMap ta, tb;
    const uint64_t dA[4] = {(uint64_t)CIN, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    const uint64_t sA[3] = {(uint64_t)CIN * 2, (uint64_t)Win * CIN * 2, (uint64_t)Hin * Win * CIN * 2};
    // with an element (traversal) stride e the box spans boxDim positions and keeps ceil(boxDim / e) of them
    const uint32_t bA[4] = {(uint32_t)S::CK, (uint32_t)(p.BW * c.stride), (uint32_t)(p.BH * c.stride), 1};
    const uint32_t eA[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    const CUtensorMapSwizzle swz = S::kSwz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (!make_tmap_f16(&ta, in, 4, dA, sA, bA, eA, swz)) return -1;
    const int K = c.ks * c.ks * CIN;
    const uint64_t dB[2] = {(uint64_t)K, (uint64_t)c.cout};
    const uint64_t sB[1] = {(uint64_t)K * 2};
    const uint32_t bB[2] = {(uint32_t)S::CK, (uint32_t)NPAD};
    if (!make_tmap_f16(&tb, c.w16, 2, dB, sB, bB, nullptr, swz)) return -1;
    const long grid = (long)B * p.tiles_w * p.tiles_h;
    conv_tc_kernel<CIN, NPAD><<<(unsigned)grid, kConvThreads, S::kTotal, s>>>(ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int CIN, int NPAD>
int set_attr() {
    return cudaFuncSetAttribute(conv_tc_kernel<CIN, NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ConvSmem<CIN, NPAD>::kTotal) == cudaSuccess ? 0 : -1;
}

}  // namespace

int conv_tc_init_device() {
    return set_attr<32, 32>() | set_attr<32, 64>() | set_attr<64, 64>() | set_attr<64, 128>() |
           set_attr<128, 128>() | set_attr<128, 48>() | set_attr<128, 64>();
}

// in: NHWC fp16 (B,Hin,Win,cin).  out: NHWC fp16, or (B,cout,Ho*Wo) fp16 when nchw != 0.
int launch_conv_tc(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw,
                   cudaStream_t s) {
    const int npad = c.cout <= 32 ? 32 : (c.cout <= 48 ? 48 : (c.cout <= 64 ? 64 : 128));
    if (c.cout > 128 || (!nchw && c.cout % 32)) return -1;
    if (c.cin == 32 && npad == 32) return launch_one<32, 32>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 32 && npad == 64) return launch_one<32, 64>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 64 && npad == 64) return launch_one<64, 64>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 64 && npad == 128) return launch_one<64, 128>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 128 && npad == 128) return launch_one<128, 128>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 128 && npad == 48) return launch_one<128, 48>(c, in, B, Hin, Win, out, nchw, s);
    if (c.cin == 128 && npad == 64) return launch_one<128, 64>(c, in, B, Hin, Win, out, nchw, s);
    return -1;
}

}  // namespace egx
