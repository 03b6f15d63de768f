// SE-ResNet trunk kernels that are not tensor-core work:
//   K2  stem: conv3x3 1->32 (+bias) -> ReLU -> BN        Full_model/ResNetSE34V2.py:64-66
//   K3' direct implicit-GEMM conv on CUDA cores (fp32 arm and fall-back shapes)
//                                                         Full_model/ResNetBlocks.py:24-30
//   K4  SE: global mean -> FC/8 -> ReLU -> FC -> sigmoid -> relu(gate*y + residual)
//                                                         Full_model/ResNetBlocks.py:28-36,92-95
// Activations are NHWC; T is float (fp32 arm) or __half (tensor-core arm).
#include "egx_common.cuh"

namespace egx {

namespace {

template <class T> struct Vec4;
template <> struct Vec4<float> {
    using type = float4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__half> {
    using type = uint2;
    static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
        uint2 t;
        *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
        *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

// packed fp32x2 arithmetic (sm_100: FFMA2 — two independent IEEE fp32 FMAs per issued instruction, bit-identical to two
// FFMAs): the CUDA-core kernels here are instruction-issue bound, not FMA-pipe bound
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ------------------------------------------------------------------------------------------
// K2 stem.  spec (B,H,W) fp32 -> out (B,H,W,32) T.  Persistent CTAs walk (clip, 32-row strip) items.  A thread
// owns 8 output channels whose 72 weights + bias / BN affine stay in registers for the whole kernel; the strip's
// input rows (+1 halo row each side, zero padded) are staged in one of two shared-memory buffers, and the next
// item's rows are already in flight (cp.async, zero fill for the padding) while the current strip is computed.  Inner loop: 9 LDS + 72 FMA
// per 16-byte (8 x fp16) store; 4*H*W bytes in, 2*32*H*W bytes out per clip, a warp writes 512 contiguous bytes.
// ------------------------------------------------------------------------------------------
constexpr int kStemRows = 32;

template <class T>
__global__ void __launch_bounds__(256, 2)
stem_kernel(const float* __restrict__ spec, int H, int W, int strips, int n_items, const float* __restrict__ w,
            const float* __restrict__ bias, const float* __restrict__ scale,
            const float* __restrict__ shift, T* __restrict__ out) {
    extern __shared__ float s_in[];                       // [2][(kStemRows + 2)][W + 2]
    const int PW = W + 2;
    const int n_stage = (kStemRows + 2) * PW;
    const int cg = threadIdx.x & 3;                        // channels [8*cg, 8*cg + 8)
    // weights and bias as channel PAIRS (2p, 2p + 1) for the packed FMAs
    uint64_t wr2[4][9], br2[4];
    float sr[8], tr[8];
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
        const int c = cg * 8 + 2 * p2;
#pragma unroll
        for (int k = 0; k < 9; ++k) wr2[p2][k] = pack2(w[c * 9 + k], w[(c + 1) * 9 + k]);
        br2[p2] = pack2(bias[c], bias[c + 1]);
        sr[2 * p2] = scale[c]; sr[2 * p2 + 1] = scale[c + 1];
        tr[2 * p2] = shift[c]; tr[2 * p2 + 1] = shift[c + 1];
    }
    // pin the 96 parameters in registers (otherwise the compiler re-reads them from global memory inside the loop)
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
#pragma unroll
        for (int k = 0; k < 9; ++k) asm volatile("" : "+l"(wr2[p2][k]));
        asm volatile("" : "+l"(br2[p2]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("" : "+f"(sr[j]), "+f"(tr[j]));
    // stage the padded strip of `item` into buffer sb with 4-byte cp.async copies (src-size 0 = zero fill for the padding)
    auto stage = [&](int item, float* sb) {
        const int b = item / strips, y0 = (item - b * strips) * kStemRows;
        for (int i = threadIdx.x; i < n_stage; i += 256) {
            const int yy = y0 + i / PW - 1, xx = i % PW - 1;
            const bool inb = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float* src = inb ? spec + ((size_t)b * H + yy) * W + xx : spec;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sb + i)),
                         "l"(src), "r"(inb ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int item = blockIdx.x;
    if (item < n_items) stage(item, s_in);
    int buf = 0;
    for (; item < n_items; item += gridDim.x, buf ^= 1) {
        float* sb = s_in + buf * n_stage;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // one barrier per item: the other buffer is two items old
        if (item + (int)gridDim.x < n_items) stage(item + gridDim.x, s_in + (buf ^ 1) * n_stage);
        const int b = item / strips, y0 = (item - b * strips) * kStemRows;
        const int rows = min(kStemRows, H - y0);
        int ly = 0, x = threadIdx.x >> 2;
        while (x >= W) { x -= W; ++ly; }
        for (; ly < rows;) {
            uint64_t tap2[9];                               // (tap, tap)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float t = sb[(ly + dy) * PW + x + dx];
                    tap2[dy * 3 + dx] = pack2(t, t);
                }
            float v[8];
#pragma unroll
            for (int p2 = 0; p2 < 4; ++p2) {
                uint64_t acc = br2[p2];
#pragma unroll
                for (int k = 0; k < 9; ++k) acc = fma2(wr2[p2][k], tap2[k], acc);    // same order as the scalar loop
                float a0, a1;
                unpack2(acc, a0, a1);
                v[2 * p2] = fmaxf(a0, 0.f) * sr[2 * p2] + tr[2 * p2];
                v[2 * p2 + 1] = fmaxf(a1, 0.f) * sr[2 * p2 + 1] + tr[2 * p2 + 1];
            }
            T* o = out + (((size_t)b * H + y0 + ly) * W + x) * 32 + cg * 8;
            const float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
            Vec4<T>::store(o, lo);
            Vec4<T>::store(o + 4, hi);
            x += 64;
            while (x >= W) { x -= W; ++ly; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K3' direct implicit-GEMM convolution (CUDA cores, fp32 accumulate).
//   M = B*Ho*Wo output pixels, N = cout, K = ks*ks*cin.   Tile 64 x 32 x 32, 128 threads,
//   4x4 register micro-tile.  Weights [cout][ks*ks][cin] fp32.
// ------------------------------------------------------------------------------------------
constexpr int CBM = 64, CBN = 32, CBK = 32;

template <class T>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const T* __restrict__ in, int B, int Hin, int Win, int Cin, int Ho, int Wo,
                   int Cout, int ks, int stride, int pad, const float* __restrict__ w,
                   const float* __restrict__ bias, const float* __restrict__ scale,
                   const float* __restrict__ shift, int relu_first, T* __restrict__ out,
                   float* __restrict__ out_nchw) {
    __shared__ __align__(16) float As[CBK][CBM + 4];
    __shared__ __align__(16) float Bs[CBK][CBN + 4];
    const int tid = threadIdx.x;
    const int64_t M = (int64_t)B * Ho * Wo;
    const int64_t m0 = (int64_t)blockIdx.x * CBM;
    const int n0 = blockIdx.y * CBN;
    const int tm = tid & 15, tn = tid >> 4;           // 16 x 8 thread grid

    // A loader: thread -> pixel tid/2, 16 consecutive input channels
    const int lp = tid >> 1, lc = (tid & 1) * 16;
    const int64_t lm = m0 + lp;
    const bool lvalid = lm < M;
    int lb = 0, lho = 0, lwo = 0;
    if (lvalid) {
        lwo = int(lm % Wo);
        lho = int((lm / Wo) % Ho);
        lb = int(lm / ((int64_t)Wo * Ho));
    }
    // B loader: thread -> cout tid/4, 8 consecutive k
    const int bn = tid >> 2, bk = (tid & 3) * 8;
    const int K = ks * ks * Cin;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < ks * ks; ++tap) {
        const int dy = tap / ks, dx = tap % ks;
        const int hi = lho * stride + dy - pad, wi = lwo * stride + dx - pad;
        const bool inb = lvalid && hi >= 0 && hi < Hin && wi >= 0 && wi < Win;
        const T* src = in + (((size_t)lb * Hin + (inb ? hi : 0)) * Win + (inb ? wi : 0)) * Cin;
        for (int c0 = 0; c0 < Cin; c0 += CBK) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (inb) Vec4<T>::load(src + c0 + lc + q * 4, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) As[lc + q * 4 + j][lp] = v[j];
            }
            {
                const int n = n0 + bn;
                const float* wsrc = w + (size_t)(n < Cout ? n : 0) * K + tap * Cin + c0 + bk;
#pragma unroll
                for (int j = 0; j < 8; ++j) Bs[bk + j][bn] = (n < Cout) ? wsrc[j] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CBK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + tm * 4 + i;
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            float t = acc[i][j];
            if (n < Cout) {
                if (bias) t += bias[n];
                if (relu_first) t = fmaxf(t, 0.f);
                t = t * scale[n] + shift[n];
            }
            v[j] = t;
        }
        if (out_nchw) {
            const int wo = int(m % Wo), ho = int((m / Wo) % Ho), b = int(m / ((int64_t)Wo * Ho));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tn * 4 + j;
                if (n < Cout) out_nchw[(((size_t)b * Cout + n) * Ho + ho) * Wo + wo] = v[j];
            }
        } else if (n0 + tn * 4 + 3 < Cout) {
            Vec4<T>::store(out + m * Cout + n0 + tn * 4, v);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tn * 4 + j;
                if (n < Cout) out[m * Cout + n] = T(v[j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4a per-(clip, pixel-chunk, channel) partial sums of y (NHWC).  grid (chunks, B).  Partials
// are stored, not atomically added: the gate kernel sums them in a fixed order, so a clip's
// result is bit-identical whatever the batch size or GPU count (SURVEY.md §4 iv).
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
se_reduce_kernel(const T* __restrict__ y, int HW, int C, int pix_per_block,
                 float* __restrict__ sums) {
    __shared__ float red[256];
    const int b = blockIdx.y;
    const int c4 = C / 4;                       // threads per pixel (4 channels each)
    const int lanes = 256 / c4;                 // pixels processed in parallel
    const int cq = threadIdx.x % c4, pl = threadIdx.x / c4;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const T* base = y + (size_t)b * HW * C;
    for (int p = p0 + pl; p < p1; p += lanes) {
        float v[4];
        Vec4<T>::load(base + (size_t)p * C + cq * 4, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[threadIdx.x] = a[j];
        __syncthreads();
        if (pl == 0) {
            float s = 0.f;
            for (int l = 0; l < lanes; ++l) s += red[l * c4 + cq];
            sums[((size_t)b * gridDim.x + blockIdx.x) * C + cq * 4 + j] = s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K4b gate + residual + ReLU.  grid (chunks, B).  Each CTA recomputes the clip's gate (C*C/4
// MACs) instead of a separate launch: out = relu(gate[c]*y + res).
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
se_apply_kernel(const T* __restrict__ y, const T* __restrict__ res, const float* __restrict__ sums,
                int n_part, int HW, int C, int R, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, int pix_per_block,
                T* __restrict__ out) {
    __shared__ float mean[256], hid[32], gate[256];
    const int b = blockIdx.y;
    if (threadIdx.x < C) {
        float t = 0.f;
        for (int i = 0; i < n_part; ++i) t += sums[((size_t)b * n_part + i) * C + threadIdx.x];
        mean[threadIdx.x] = t / (float)HW;
    }
    __syncthreads();
    if (threadIdx.x < R) {
        float a = b1[threadIdx.x];
        for (int c = 0; c < C; ++c) a = fmaf(w1[threadIdx.x * C + c], mean[c], a);
        hid[threadIdx.x] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float a = b2[threadIdx.x];
        for (int j = 0; j < R; ++j) a = fmaf(w2[threadIdx.x * R + j], hid[j], a);
        gate[threadIdx.x] = 1.f / (1.f + expf(-a));
    }
    __syncthreads();
    const int c4 = C / 4;
    const size_t e0 = (size_t)blockIdx.x * pix_per_block * c4;
    const size_t e1 = min((size_t)HW * c4, e0 + (size_t)pix_per_block * c4);
    const size_t base = (size_t)b * HW * C;
    for (size_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int cq = int(e % c4);
        float vy[4], vr[4], vo[4];
        Vec4<T>::load(y + base + e * 4, vy);
        Vec4<T>::load(res + base + e * 4, vr);
#pragma unroll
        for (int j = 0; j < 4; ++j) vo[j] = fmaxf(fmaf(gate[cq * 4 + j], vy[j], vr[j]), 0.f);
        Vec4<T>::store(out + base + e * 4, vo);
    }
}

// ------------------------------------------------------------------------------------------
// K4c/K4d  SE gate BEFORE conv2 runs (tensor-core arm).  The gate needs mean_{h,w}(BN2(conv2(y1))), and that
// mean is linear in y1: for tap (ky,kx) of the zero-padded stride-1 3x3 conv2 the outputs read, in total, every
// pixel of y1 except one border row and one border column, so
//     mean(conv2(y1))[co] = (1/HW) * sum_{tap,ci} W2[co][tap][ci] * S_tap[ci],
//     S_tap = T - R(excluded row) - C(excluded column) + y1(excluded row, excluded column),
// with T the channel totals (summed by conv1's epilogue, fixed order), R / C border row / column sums.
// K4c builds the nine window means per clip as one fp16 row [9*C] (the A operand of a small tensor-core GEMM
// against conv2's own packed weights); K4d turns the GEMM result into the folded per-clip epilogue of conv2:
//     out = relu(acc * (g*scale2) + (g*shift2) + residual),  g = sigmoid(W2 relu(W1 mean + b1) + b2)
// (Full_model/ResNetBlocks.py:28-36,92-95).  conv2's output is never written unscaled and the separate
// gate*y + residual pass over the map (3 map transfers per block) disappears.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
se_window_kernel(const __half* __restrict__ y, int H, int W, int C, const float* __restrict__ part, int n_part,
                 __half* __restrict__ win) {
    __shared__ float red[4][64][9];              // [class][thread of the class][channel in octet] (padded)
    __shared__ float cls[5][256];                // T, R0 (first row), RL (last row), C0 (first col), CL (last col)
    __shared__ float tot[8][256];                // [slice][channel] partial totals (256 / C slices, C >= 32)
    const int b = blockIdx.x;
    const __half* img = y + (size_t)b * H * W * C;
    // tile totals: 256 / C slices of the partials per channel (summed over the slices in a fixed order below)
    {
        const int c = threadIdx.x % C, sl = threadIdx.x / C, nsl = 256 / C;
        float t = 0.f;
#pragma unroll 4
        for (int i = sl; i < n_part; i += nsl) t += part[((size_t)b * n_part + i) * C + c];
        tot[sl][c] = t;
    }
    // border sums: 64 threads per class (first row, last row, first column, last column); inside a class C/8 threads
    // cover one pixel (16-byte loads) and 512/C pixels are in flight
    const int k = threadIdx.x >> 6, tk = threadIdx.x & 63;
    const int c8 = C / 8, lanes = 64 / c8;
    const int cq = tk % c8, pl = tk / c8;
    const int n_pix = k < 2 ? W : H;
    const size_t base = k == 0 ? 0 : (k == 1 ? (size_t)(H - 1) * W * C : (k == 2 ? 0 : (size_t)(W - 1) * C));
    const size_t pitch = k < 2 ? (size_t)C : (size_t)W * C;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll 4
    for (int i = pl; i < n_pix; i += lanes) {
        const uint4 t = *reinterpret_cast<const uint4*>(img + base + (size_t)i * pitch + cq * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h2[j]); a[2 * j] += f.x; a[2 * j + 1] += f.y; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[k][tk][j] = a[j];
    __syncthreads();
    for (int o = threadIdx.x; o < 5 * C; o += 256) {
        const int kk = o / C, c = o % C;
        float t = 0.f;
        if (kk == 0) {
            for (int l = 0; l < 256 / C; ++l) t += tot[l][c];
        } else {
            for (int l = 0; l < lanes; ++l) t += red[kk - 1][l * c8 + (c >> 3)][c & 7];
        }
        cls[kk][c] = t;
    }
    __syncthreads();
    const float inv = 1.f / (float)(H * W);
    for (int o = threadIdx.x; o < 9 * C; o += 256) {
        const int tap = o / C, c = o % C;
        const int ky = tap / 3, kx = tap % 3;
        float s = cls[0][c];
        int er = -1, ec = -1;                                        // excluded row / column of y1
        if (ky == 0) { s -= cls[2][c]; er = H - 1; } else if (ky == 2) { s -= cls[1][c]; er = 0; }
        if (kx == 0) { s -= cls[4][c]; ec = W - 1; } else if (kx == 2) { s -= cls[3][c]; ec = 0; }
        if (er >= 0 && ec >= 0) s += __half2float(img[((size_t)er * W + ec) * C + c]);
        win[(size_t)b * 9 * C + o] = __float2half_rn(s * inv);
    }
}

__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ mean_raw, int C, int R, const float* __restrict__ bias,
               const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
               float* __restrict__ gate) {
    __shared__ float mean[256], hid[32];
    const int b = blockIdx.x, c = threadIdx.x;
    float sc = 0.f, sh = 0.f;
    if (c < C) {
        sc = scale[c];
        sh = shift[c] + (bias ? bias[c] * sc : 0.f);
        mean[c] = fmaf(mean_raw[(size_t)b * C + c], sc, sh);
    }
    __syncthreads();
    // hidden units: one warp each, lanes stride the channels (coalesced weight rows)
    for (int j = c >> 5; j < R; j += blockDim.x >> 5) {
        float t = 0.f;
        for (int i = c & 31; i < C; i += 32) t = fmaf(w1[j * C + i], mean[i], t);
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((c & 31) == 0) hid[j] = fmaxf(t + b1[j], 0.f);
    }
    __syncthreads();
    if (c < C) {
        float t = b2[c];
        for (int j = 0; j < R; ++j) t = fmaf(w2[c * R + j], hid[j], t);
        const float g = 1.f / (1.f + expf(-t));
        gate[(size_t)b * 2 * C + c] = g * sc;
        gate[(size_t)b * 2 * C + C + c] = g * sh;
    }
}

template <class T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int64_t total, int HW, int C,
                                    float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int p = int(i % HW);
        const int c = int((i / HW) % C);
        const int64_t b = i / ((int64_t)HW * C);
        out[i] = float(in[(b * HW + p) * C + c]);
    }
}

inline bool ok() { return cudaGetLastError() == cudaSuccess; }

}  // namespace

template <class T>
int launch_stem(const ConvW& c, const float* spec, int B, int H, int W, T* out, cudaStream_t s) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    }
    const int strips = (H + kStemRows - 1) / kStemRows;
    const long n_items = (long)B * strips;
    if (n_items > 0x7fffffffL) return -1;
    const int n_stage = (kStemRows + 2) * (W + 2);
    const size_t smem = sizeof(float) * 2 * n_stage;
    if (smem > 48 * 1024) return -1;
    const int grid = (int)std::min<long>(n_items, 2L * sms);
    stem_kernel<T><<<grid, 256, smem, s>>>(spec, H, W, strips, (int)n_items, c.w32, c.bias, c.scale, c.shift, out);
    return ok() ? 1 : -1;
}

template <class T>
int launch_conv_direct(const ConvW& c, const T* in, int B, int Hin, int Win, T* out,
                       float* out_nchw_f32, cudaStream_t s) {
    const int pad = c.ks / 2;
    const int Ho = (Hin + 2 * pad - c.ks) / c.stride + 1;
    const int Wo = (Win + 2 * pad - c.ks) / c.stride + 1;
    const int64_t M = (int64_t)B * Ho * Wo;
    dim3 grid((unsigned)((M + CBM - 1) / CBM), (unsigned)((c.cout + CBN - 1) / CBN));
    conv_direct_kernel<T><<<grid, 128, 0, s>>>(in, B, Hin, Win, c.cin, Ho, Wo, c.cout, c.ks,
                                               c.stride, pad, c.w32, c.bias, c.scale, c.shift,
                                               c.relu_first, out, out_nchw_f32);
    return ok() ? 1 : -1;
}

template <class T>
int launch_se_reduce(const T* y, int B, int HW, int C, float* sums, cudaStream_t s) {
    const int ppb = kSePixPerBlock;
    dim3 grid((HW + ppb - 1) / ppb, B);
    se_reduce_kernel<T><<<grid, 256, 0, s>>>(y, HW, C, ppb, sums);
    return ok() ? 1 : -1;
}

template <class T>
int launch_se_apply(const SEW& se, const T* y, const T* res, const float* sums, int n_part, int B,
                    int HW, T* out, cudaStream_t s) {
    const int ppb = 1024;
    dim3 grid((HW + ppb - 1) / ppb, B);
    se_apply_kernel<T><<<grid, 256, 0, s>>>(y, res, sums, n_part, HW, se.c, se.r, se.w1, se.b1, se.w2,
                                            se.b2, ppb, out);
    return ok() ? 1 : -1;
}

int launch_se_window(const __half* y, int B, int H, int W, int C, const float* part, int n_part, __half* win,
                     cudaStream_t s) {
    if (C % 32 || C > 256 || 256 % C || 64 % (C / 8)) return -1;
    se_window_kernel<<<B, 256, 0, s>>>(y, H, W, C, part, n_part, win);
    return ok() ? 1 : -1;
}

int launch_se_gate(const SEW& se, const ConvW& conv2, const float* mean_raw, int B, float* gate, cudaStream_t s) {
    if (se.c > 256 || se.r > 32 || conv2.relu_first) return -1;
    se_gate_kernel<<<B, 256, 0, s>>>(mean_raw, se.c, se.r, conv2.bias, conv2.scale, conv2.shift, se.w1, se.b1, se.w2,
                                     se.b2, gate);
    return ok() ? 1 : -1;
}

template <class T>
int launch_nhwc_to_nchw_f32(const T* in, int B, int HW, int C, float* out, cudaStream_t s) {
    const int64_t total = (int64_t)B * HW * C;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
    nhwc_to_nchw_kernel<T><<<grid, 256, 0, s>>>(in, total, HW, C, out);
    return ok() ? 1 : -1;
}

#define EGX_INST(T)                                                                              \
    template int launch_stem<T>(const ConvW&, const float*, int, int, int, T*, cudaStream_t);    \
    template int launch_conv_direct<T>(const ConvW&, const T*, int, int, int, T*, float*,        \
                                       cudaStream_t);                                            \
    template int launch_se_reduce<T>(const T*, int, int, int, float*, cudaStream_t);             \
    template int launch_se_apply<T>(const SEW&, const T*, const T*, const float*, int, int, int, \
                                    T*, cudaStream_t);                                               \
    template int launch_nhwc_to_nchw_f32<T>(const T*, int, int, int, float*, cudaStream_t);
EGX_INST(float)
EGX_INST(__half)

}  // namespace egx
