// K1 — fused log-mel front-end: pre-emphasis -> Hann STFT (n_fft 1024, hop 512, centred,
// zero padded) -> 513-bin power -> 128 Slaney mels -> log/dB -> InstanceNorm / ref-max.
//
// Follows: model/utils.py:33-38 (F1), utils/data_utils.py:36 (F2/F3, librosa defaults),
// utils/data_utils.py:37 (F4a), model/ResNetSE34V2.py:96-98 (F4b).
//
// One CTA per clip, one warp per STFT frame (round-robin).  A frame is a 1024-point real
// FFT done as a 512-point complex FFT: three radix-8 Stockham passes whose butterflies live in
// registers (16 points per lane) and exchange through one padded, conflict-free shared-memory
// buffer per warp (fp32, twiddles from a float64-built table), followed by the real-FFT untangling,
// the sparse mel projection (<= 24 taps per mel) and the log.  The raw log values go straight to the
// output tile; once the CTA has written all of it, the per-(clip, mel) statistics are taken from that
// tile (still in L2, 36 KB per clip) and it is normalised in place.  Keeping the tile out of shared
// memory lets three CTAs share an SM (the kernel is latency-bound: every FFT pass is a shared-memory
// round trip).  Algorithmic HBM traffic 4*N + 4*128*W bytes per clip.
#include <algorithm>

#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int kFFT = 1024;
constexpr int kHalf = 512;       // complex FFT length
constexpr int kHop = 512;
constexpr int kMels = 128;
constexpr int kMaxWarps = 12;      // 384 threads x 3 CTAs per SM at <= 56 registers
constexpr int kBufLen = kHalf + kHalf / 8;     // padded: physical index i + (i >> 3)

// Every multiply-add below is spelled out (fmaf / __fmul_rn): left to the compiler, which product of a*b - c*d is fused
// is decided per call site, and the two kernels of this file must agree bit for bit.
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -__fmul_rn(a.y, b.y)), fmaf(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)
__device__ __forceinline__ int pad(int i) { return i + (i >> 3); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 8-point DFT in registers (forward, W8 = exp(-2 pi i / 8)): radix-2 DIT, y[u] = E[u & 3] +- W8^u O[u & 3]
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
    const float h = 0.70710678118654752f;
    const float2 b0 = cadd(a[0], a[4]), b1 = csub(a[0], a[4]), b2 = cadd(a[2], a[6]), b3 = mul_mi(csub(a[2], a[6]));
    const float2 b4 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]), b6 = cadd(a[3], a[7]), b7 = mul_mi(csub(a[3], a[7]));
    const float2 c0 = cadd(b0, b2), c1 = cadd(b1, b3), c2 = csub(b0, b2), c3 = csub(b1, b3);
    const float2 c4 = cadd(b4, b6), o1 = cadd(b5, b7), o2 = csub(b4, b6), o3 = csub(b5, b7);
    const float2 c6 = mul_mi(o2);                                            // * (-i)
    // odd outputs 1/5 and 3/7: c + W8 o and c - W8 o with W8 = (1 - i)/sqrt2, (-1 - i)/sqrt2 — explicit fused forms
    const float s5x = o1.x + o1.y, s5y = o1.y - o1.x, s7x = o3.y - o3.x, s7y = o3.x + o3.y;
    a[0] = cadd(c0, c4); a[4] = csub(c0, c4);
    a[1] = make_float2(fmaf(h, s5x, c1.x), fmaf(h, s5y, c1.y));
    a[5] = make_float2(fmaf(-h, s5x, c1.x), fmaf(-h, s5y, c1.y));
    a[2] = cadd(c2, c6); a[6] = csub(c2, c6);
    a[3] = make_float2(fmaf(h, s7x, c3.x), fmaf(-h, s7y, c3.y));
    a[7] = make_float2(fmaf(-h, s7x, c3.x), fmaf(h, s7y, c3.y));
}

// One Stockham radix-8 pass of the warp's 512-point FFT, in place: every lane reads the 16 points of its two
// butterflies into registers, the warp synchronises, then the results go back in autosort order.
//   in[j + 64 t] * W_{8 Ns}^{k t}  --DFT8-->  out[(j - k) * 8 + k + t * Ns],   k = j mod Ns
template <int NS>
__device__ __forceinline__ void fft_pass(float2* buf, int lane, const float2* __restrict__ tw512) {
    float2 a[2][8];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const int j = lane + 32 * h2;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            a[h2][t] = buf[pad(j + 64 * t)];
            // twiddle W512^(k t 64 / Ns) from the [pass][t][j] table: consecutive lanes, consecutive entries
            if (NS > 1 && t > 0) a[h2][t] = cmul(a[h2][t], __ldg(&tw512[((NS == 8 ? 0 : 8) + t) * 64 + j]));
        }
        dft8(a[h2]);
    }
    __syncwarp();
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const int j = lane + 32 * h2, k = j & (NS - 1);
#pragma unroll
        for (int t = 0; t < 8; ++t) buf[pad((j - k) * 8 + k + t * NS)] = a[h2][t];
    }
    __syncwarp();
}

// raw samples of 8 of the 16 complex points of this lane: point n = lane + 32 (i0 + i) holds samples (base + 2n, base + 2n + 1)
template <bool kAligned>
__device__ __forceinline__ void load_frame(const float* __restrict__ x, int N, int base, int lane, int i0, float (&r0)[8], float (&r1)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int s0 = base + 2 * (lane + 32 * (i0 + i));
        r0[i] = 0.f; r1[i] = 0.f;
        if (s0 >= 0 && s0 + 1 < N) {
            if (kAligned) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(x + s0));
                r0[i] = v.x; r1[i] = v.y;
            } else {
                r0[i] = __ldg(x + s0); r1[i] = __ldg(x + s0 + 1);
            }
        } else if (s0 >= 0 && s0 < N) {
            r0[i] = __ldg(x + s0);
        }
    }
}

__global__ void __launch_bounds__(kMaxWarps * 32, 3)
logmel_kernel(const float* __restrict__ audio, int N, int n_cols, int mode, int preemph,
              float* __restrict__ out, const float* __restrict__ window,
              const float2* __restrict__ tw512, const float2* __restrict__ tw1024,
              const int* __restrict__ mel_start, const int* __restrict__ mel_ptr,
              const float* __restrict__ mel_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw);                 // [warps][kBufLen]
    __shared__ float s_red[kMaxWarps];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const bool log_in = (mode & 0xff) == EGX_LOGMEL_LOG_IN, f16_round = (mode & EGX_LOGMEL_FP16_STORAGE) != 0;
    const float* x = audio + (size_t)blockIdx.x * N;
    float2* buf = fftbuf + warp * kBufLen;
    float* tile = out + (size_t)blockIdx.x * kMels * n_cols;             // raw log values, normalised in place below
    // 8-byte loads need the clip's first sample on an even float (N may be odd): uniform per CTA
    const bool aligned = (((size_t)blockIdx.x * N) & 1) == 0 && (reinterpret_cast<uintptr_t>(audio) & 7) == 0;

    for (int t = warp; t < n_cols; t += n_warps) {
        // ---- framing + pre-emphasis + window; even/odd samples of the frame form one complex point ----
        // all 32 loads of the lane are issued before the first use; x[s0 - 1] comes from the neighbouring lane
        const int base = t * kHop - kFFT / 2;
        float carry = 0.f;                             // x1 of lane 31 of the previous 32-point group
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += 8) {
        float r0[8], r1[8];
        if (aligned) load_frame<true>(x, N, base, lane, i0, r0, r1);
        else load_frame<false>(x, N, base, lane, i0, r0, r1);
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = i0 + ii, n = lane + 32 * i, s0 = base + 2 * n;
            float x0 = r0[ii], x1 = r1[ii];
            if (preemph) {
                // y[t] = x[t] - 0.97 x[t-1], x[-1] := x[1] (reflect pad of 1); samples outside the clip stay zero
                float xm = __shfl_up_sync(0xffffffffu, r1[ii], 1);
                if (lane == 0) xm = i ? carry : ((s0 - 1 >= 0 && s0 - 1 < N) ? __ldg(x + s0 - 1) : 0.f);
                carry = __shfl_sync(0xffffffffu, r1[ii], 31);
                const float y1 = fmaf(-0.97f, x0, x1);
                const float y0 = fmaf(-0.97f, s0 == 0 ? r1[ii] : xm, x0);
                x1 = (s0 + 1 >= 0 && s0 + 1 < N) ? y1 : 0.f;
                x0 = (s0 >= 0 && s0 < N) ? y0 : 0.f;
            }
            const float2 w = __ldg(reinterpret_cast<const float2*>(window + 2 * n));
            buf[pad(n)] = make_float2(__fmul_rn(x0, w.x), __fmul_rn(x1, w.y));
        }
        }
        __syncwarp();
        // ---- 512-point complex FFT: three radix-8 Stockham passes ----
        fft_pass<1>(buf, lane, tw512);
        fft_pass<8>(buf, lane, tw512);
        fft_pass<64>(buf, lane, tw512);
        // ---- untangle to the 513-bin real spectrum; the power overwrites the buffer (513 floats) ----
        float pw[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int k = lane + 32 * i;
            pw[i] = 0.f;
            if (k <= kHalf) {
                const float2 zk = buf[pad(k & (kHalf - 1))];
                const float2 zn = buf[pad((kHalf - k) & (kHalf - 1))];
                const float2 ev = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                const float2 od = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
                const float2 r = cmul(od, __ldg(&tw1024[k]));
                const float re = ev.x + r.x, im = ev.y + r.y;
                pw[i] = fmaf(re, re, __fmul_rn(im, im));
            }
        }
        __syncwarp();
        float* power = reinterpret_cast<float*>(buf);
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int k = lane + 32 * i;
            if (k <= kHalf) power[k] = pw[i];
        }
        __syncwarp();
        // ---- sparse mel projection + log ----
#pragma unroll
        for (int mm = 0; mm < kMels / 32; ++mm) {
            const int m = lane + 32 * mm;
            const int nt = mel_ptr[m], b0 = mel_start[m];
            float acc = 0.f;
            for (int p = 0; p < nt; ++p) acc = fmaf(__ldg(mel_w + p * kMels + m), power[b0 + p], acc);
            float v;
            if (log_in) v = logf(acc + 1e-6f);
            else v = 10.f * log10f(fmaxf(acc, 1e-10f));
            tile[m * n_cols + t] = v;
        }
        __syncwarp();
    }
    __syncthreads();        // CTA-scope visibility of the tile written above (global memory, same CTA)

    float* o = tile;
    if (log_in) {
        // InstanceNorm1d: per (clip, mel) over time, biased variance, eps 1e-5, no affine
        for (int m = warp; m < kMels; m += n_warps) {
            const float* row = tile + m * n_cols;
            float s = 0.f;
            for (int t = lane; t < n_cols; t += 32) s += row[t];
            const float mean = warp_sum(s) / n_cols;
            float q = 0.f;
            for (int t = lane; t < n_cols; t += 32) { const float d = row[t] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(warp_sum(q) / n_cols + 1e-5f);
            for (int t = lane; t < n_cols; t += 32) o[m * n_cols + t] = (row[t] - mean) * rstd;
        }
    } else {
        // power_to_db(ref=np.max): subtract the clip maximum, floor at (max - 80 dB) = -80
        float mx = -INFINITY;
        for (int i = threadIdx.x; i < kMels * n_cols; i += blockDim.x) mx = fmaxf(mx, tile[i]);
        mx = warp_max(mx);
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        mx = s_red[0];
        for (int w = 1; w < n_warps; ++w) mx = fmaxf(mx, s_red[w]);
        // astype('float16') of utils/data_utils.py:38 (the features are stored as fp16 and re-read as fp32,
        // data_loader/lmdb_data_loader_expressive.py:204): optional, applied after the comparison point of the 1e-4 gate
        for (int i = threadIdx.x; i < kMels * n_cols; i += blockDim.x) {
            const float v = fmaxf(tile[i] - mx, -80.f);
            o[i] = f16_round ? __half2float(__float2half_rn(v)) : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Round-2 variant for tiles that fit in shared memory next to the FFT buffers twice per SM (kTileMaxSmem): the SAME
// arithmetic in the same order (bit-identical output), restructured around what bounded the kernel above — the L1/shared
// pipe (82% busy: 658 shared wavefronts + 437 global tag requests per frame) —
//   * the frame goes from global memory straight into the registers of the first radix-8 pass (no staging round trip);
//     interior frames (68 of 71 for a TED clip) skip every bounds check;
//   * the buffer index is XOR-swizzled (i ^ ((i >> 3) & 15)): conflict-free for the stride-8 / stride-64 writes AND the
//     unit-stride reads (the padded layout made every read a 2-way conflict), and needs no padding;
//   * the last pass stays in registers and the real-FFT untangling fetches its mirror bin Z[512 - k] from the lane
//     that holds it with two shuffles instead of a store + two loads;
//   * the (128, n_cols) tile lives in shared memory: the transposed per-frame writes (32 lines per store) and the two
//     re-reads for the statistics never leave the SM, the normalised tile is written once, coalesced.
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t kTileMaxSmem = 113 * 1024;      // two CTAs per SM
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 15); }

// one complex point (two samples) of a frame with every check of the general path
__device__ __forceinline__ float2 load_point_checked(const float* __restrict__ x, int N, int s0, int preemph) {
    float x0 = (s0 >= 0 && s0 < N) ? __ldg(x + s0) : 0.f;
    float x1 = (s0 + 1 >= 0 && s0 + 1 < N) ? __ldg(x + s0 + 1) : 0.f;
    if (preemph) {
        const float xm = s0 == 0 ? x1 : ((s0 - 1 >= 0 && s0 - 1 < N) ? __ldg(x + s0 - 1) : 0.f);
        const float y1 = fmaf(-0.97f, x0, x1);
        const float y0 = fmaf(-0.97f, xm, x0);
        x1 = (s0 + 1 >= 0 && s0 + 1 < N) ? y1 : 0.f;
        x0 = (s0 >= 0 && s0 < N) ? y0 : 0.f;
    }
    return make_float2(x0, x1);
}

// one radix-8 pass with the inputs already in registers: twiddle, butterflies
template <int NS>
__device__ __forceinline__ void twiddle_dft8(float2 (&a)[8], int j, const float2* __restrict__ tw512) {
    if (NS > 1) {
#pragma unroll
        for (int t = 1; t < 8; ++t) a[t] = cmul(a[t], __ldg(&tw512[((NS == 8 ? 0 : 8) + t) * 64 + j]));
    }
    dft8(a);
}

__global__ void __launch_bounds__(kMaxWarps * 32, 2)
logmel_tile_kernel(const float* __restrict__ audio, int N, int n_cols, int mode, int preemph,
                   float* __restrict__ out, const float* __restrict__ window,
                   const float2* __restrict__ tw512, const float2* __restrict__ tw1024,
                   const int* __restrict__ mel_start, const int* __restrict__ mel_ptr,
                   const float* __restrict__ mel_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw);                          // [warps][kHalf + 8]
    __shared__ float s_red[kMaxWarps];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    constexpr int kBuf = kHalf + 8;                                                // 513 power values fit
    float* stile = reinterpret_cast<float*>(fftbuf + n_warps * kBuf);               // [kMels][pitch], pitch odd
    const int pitch = n_cols | 1;
    const bool log_in = (mode & 0xff) == EGX_LOGMEL_LOG_IN, f16_round = (mode & EGX_LOGMEL_FP16_STORAGE) != 0;
    const float* x = audio + (size_t)blockIdx.x * N;
    float2* buf = fftbuf + warp * kBuf;
    // the clip's first sample sits on an even float of the address space (8-byte loads of sample pairs); uniform per CTA
    const bool aligned = (((reinterpret_cast<uintptr_t>(audio) >> 2) + (size_t)blockIdx.x * N) & 1) == 0;

    // pass-2 twiddles W64^((j & 7) t): the same for j = lane and lane + 32, and for every frame — registers
    float2 tw2[7];
#pragma unroll
    for (int tt = 1; tt < 8; ++tt) tw2[tt - 1] = __ldg(&tw512[tt * 64 + lane]);

    for (int t = warp; t < n_cols; t += n_warps) {
        const int base = t * kHop - kFFT / 2;
        // interior frame: every sample (and the one before the frame) exists, and no sample is the clip's first.  A clip
        // that starts on an odd float (N odd: every other clip) reads the aligned pairs (s0 - 1, s0), (s0 + 1, s0 + 2):
        // the first carries the pre-emphasis neighbour, the second touches one sample past the frame
        const bool interior = base >= 2 && base + kFFT + (aligned ? 0 : 1) <= N;
        float2 a[2][8];
        // ---- framing + pre-emphasis + window -> registers of pass 1 (point n = j + 64 tt, j = lane + 32 h2), pass 1 ----
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int j = lane + 32 * h2;
            if (interior && aligned) {
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) a[h2][tt] = __ldg(reinterpret_cast<const float2*>(x + base + 2 * (j + 64 * tt)));
                if (preemph) {
#pragma unroll
                    for (int tt = 0; tt < 8; ++tt) {
                        // x[s0 - 1] is the odd sample of the previous point: the neighbouring lane has it
                        float xm = __shfl_up_sync(0xffffffffu, a[h2][tt].y, 1);
                        if (lane == 0) xm = __ldg(x + base + 2 * (j + 64 * tt) - 1);
                        const float y1 = fmaf(-0.97f, a[h2][tt].x, a[h2][tt].y);
                        const float y0 = fmaf(-0.97f, xm, a[h2][tt].x);
                        a[h2][tt] = make_float2(y0, y1);
                    }
                }
            } else if (interior) {
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) {
                    const float* ps = x + base + 2 * (j + 64 * tt);
                    const float2 lo = __ldg(reinterpret_cast<const float2*>(ps - 1));
                    const float2 hi = __ldg(reinterpret_cast<const float2*>(ps + 1));
                    if (preemph) {
                        const float y1 = fmaf(-0.97f, lo.y, hi.x);
                        const float y0 = fmaf(-0.97f, lo.x, lo.y);
                        a[h2][tt] = make_float2(y0, y1);
                    } else {
                        a[h2][tt] = make_float2(lo.y, hi.x);
                    }
                }
            } else {
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) a[h2][tt] = load_point_checked(x, N, base + 2 * (j + 64 * tt), preemph);
            }
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(window + 2 * (j + 64 * tt)));
                a[h2][tt] = make_float2(__fmul_rn(a[h2][tt].x, w.x), __fmul_rn(a[h2][tt].y, w.y));
            }
            dft8(a[h2]);
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) buf[swz(j * 8 + tt)] = a[h2][tt];        // NS = 1: out[(j - 0) * 8 + 0 + tt]
        }
        __syncwarp();
        // ---- pass 2 (NS = 8) ----
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int j = lane + 32 * h2;
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) a[h2][tt] = buf[swz(j + 64 * tt)];
#pragma unroll
            for (int tt = 1; tt < 8; ++tt) a[h2][tt] = cmul(a[h2][tt], tw2[tt - 1]);
            dft8(a[h2]);
        }
        __syncwarp();
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int j = lane + 32 * h2, k = j & 7;
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) buf[swz((j - k) * 8 + k + tt * 8)] = a[h2][tt];
        }
        __syncwarp();
        // ---- pass 3 (NS = 64): lane holds Z[j + 64 tt] afterwards ----
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int j = lane + 32 * h2;
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) a[h2][tt] = buf[swz(j + 64 * tt)];
            twiddle_dft8<64>(a[h2], j, tw512);
        }
        __syncwarp();                                   // every lane has read its inputs: the buffer becomes `power`
        // ---- untangle to the real spectrum: bin k = j + 64 tt needs Z[512 - k] = Z[(64 - j) + 64 (7 - tt)], held by lane
        //      (32 - lane) & 31 in the other half (lane 0: see below) ----
        float* power = reinterpret_cast<float*>(buf);
        const int src = (32 - lane) & 31;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                // what this lane must OFFER: lanes 1..31 their element (other half, 7 - tt); lane 0 is its own source and
                // needs Z[64 (8 - tt)] (h2 = 0) or Z[32 + 64 (7 - tt)] (h2 = 1)
                const float2 other = a[h2 ^ 1][7 - tt];
                const float2 self0 = h2 == 0 ? a[0][(8 - tt) & 7] : a[1][7 - tt];
                const float2 offer = lane == 0 ? self0 : other;
                float2 zn;
                zn.x = __shfl_sync(0xffffffffu, offer.x, src);
                zn.y = __shfl_sync(0xffffffffu, offer.y, src);
                const float2 zk = a[h2][tt];
                const int k = lane + 32 * h2 + 64 * tt;
                const float2 ev = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                const float2 od = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
                const float2 r = cmul(od, __ldg(&tw1024[k]));
                const float re = ev.x + r.x, im = ev.y + r.y;
                power[k] = fmaf(re, re, __fmul_rn(im, im));
            }
        }
        if (lane == 0) {                                // bin 512: zk = zn = Z[0]
            const float2 z0 = a[0][0];
            const float2 ev = make_float2(0.5f * (z0.x + z0.x), 0.5f * (z0.y - z0.y));
            const float2 od = make_float2(0.5f * (z0.y + z0.y), -0.5f * (z0.x - z0.x));
            const float2 r = cmul(od, __ldg(&tw1024[kHalf]));
            const float re = ev.x + r.x, im = ev.y + r.y;
            power[kHalf] = fmaf(re, re, __fmul_rn(im, im));
        }
        __syncwarp();
        // ---- sparse mel projection + log ----
#pragma unroll
        for (int mm = 0; mm < kMels / 32; ++mm) {
            const int m = lane + 32 * mm;
            const int nt = mel_ptr[m], b0 = mel_start[m];
            float acc = 0.f;
            for (int p = 0; p < nt; ++p) acc = fmaf(__ldg(mel_w + p * kMels + m), power[b0 + p], acc);
            float v;
            if (log_in) v = logf(acc + 1e-6f);
            else v = 10.f * log10f(fmaxf(acc, 1e-10f));
            stile[m * pitch + t] = v;
        }
        __syncwarp();
    }
    __syncthreads();

    float* o = out + (size_t)blockIdx.x * kMels * n_cols;
    if (log_in) {
        for (int m = warp; m < kMels; m += n_warps) {
            const float* row = stile + m * pitch;
            float s = 0.f;
            for (int t = lane; t < n_cols; t += 32) s += row[t];
            const float mean = warp_sum(s) / n_cols;
            float q = 0.f;
            for (int t = lane; t < n_cols; t += 32) { const float d = row[t] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(warp_sum(q) / n_cols + 1e-5f);
            for (int t = lane; t < n_cols; t += 32) o[m * n_cols + t] = (row[t] - mean) * rstd;
        }
    } else {
        float mx = -INFINITY;
        for (int m = warp; m < kMels; m += n_warps)
            for (int t = lane; t < n_cols; t += 32) mx = fmaxf(mx, stile[m * pitch + t]);
        mx = warp_max(mx);
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        mx = s_red[0];
        for (int w = 1; w < n_warps; ++w) mx = fmaxf(mx, s_red[w]);
        for (int m = warp; m < kMels; m += n_warps)
            for (int t = lane; t < n_cols; t += 32) {
                const float v = fmaxf(stile[m * pitch + t] - mx, -80.f);
                o[m * n_cols + t] = f16_round ? __half2float(__float2half_rn(v)) : v;
            }
    }
}

// F5 — make_audio_fixed_length (utils/data_utils.py:69-75): crop to N samples, or extend at the end with
// np.pad(mode='symmetric') (the signal mirrored about its last sample, edge repeated; period 2 n for pads longer than
// the clip).  Ragged clips arrive back to back in `samples`; clip b is samples[offsets[b] .. offsets[b+1]).
__global__ void fixed_length_kernel(const float* __restrict__ samples, const int64_t* __restrict__ offsets, int N,
                                    float* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t lo = offsets[b], n = offsets[b + 1] - lo;
    float* o = out + (size_t)b * N;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
        float v = 0.f;                       // an empty clip has nothing to mirror (np.pad raises; here: silence)
        if (n > 0) {
            const int64_t u = t % (2 * n);
            v = __ldg(samples + lo + (u < n ? u : 2 * n - 1 - u));
        }
        o[t] = v;
    }
}

// 16-bit PCM -> float in [-1, 1): x / 32768, exact — the scaling every wav decoder applies; done on the device so that
// the host->device copy of raw speech carries 2 bytes per sample instead of 4 (extension, see include/egx.h).
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ in, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (float)in[i] * (1.f / 32768.f);
}

}  // namespace

int launch_pcm16_to_f32(const int16_t* in, int64_t n, float* out, cudaStream_t s) {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    pcm16_to_f32_kernel<<<grid > 0 ? grid : 1, 256, 0, s>>>(in, n, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int g_logmel_tile = 1;     // EGX_LOGMEL_TILE (attribution builds): 0 = the global-tile kernel for every width

int launch_logmel(const LogmelTables& t, const float* audio, int B, int N, int n_cols, int mode,
                  int preemph, float* out, cudaStream_t s, bool force_global_tile) {
    // warps per CTA: the count in [9, 12] that wastes the fewest warp slots in the last round of frames
    int warps = kMaxWarps;
    for (int w = kMaxWarps; w >= 9; --w)
        if ((n_cols + w - 1) / w * w < (n_cols + warps - 1) / warps * warps) warps = w;
    const size_t smem_tile = sizeof(float2) * warps * (kHalf + 8) + sizeof(float) * kMels * (size_t)(n_cols | 1);
    const bool tile = !force_global_tile && env_switch("EGX_LOGMEL_TILE", g_logmel_tile) && smem_tile <= kTileMaxSmem;
    const size_t smem = tile ? smem_tile : sizeof(float2) * warps * kBufLen;
    // opt-in shared-memory size is a per-device function attribute (one handle per device may live in one process)
    static size_t configured[2][64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (smem > configured[tile][dev]) {
        const cudaError_t e = tile ? cudaFuncSetAttribute(logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                   : cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1;
        configured[tile][dev] = smem;
    }
    if (tile)
        logmel_tile_kernel<<<B, warps * 32, smem, s>>>(audio, N, n_cols, mode, preemph, out, t.window, t.tw512, t.tw1024,
                                                       t.mel_start, t.mel_ptr, t.mel_w);
    else
        logmel_kernel<<<B, warps * 32, smem, s>>>(audio, N, n_cols, mode, preemph, out, t.window,
                                                    t.tw512, t.tw1024, t.mel_start, t.mel_ptr,
                                                    t.mel_w);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_fixed_length(const float* samples, const int64_t* offsets, int B, int N, float* out, cudaStream_t s) {
    const dim3 grid((unsigned)std::min(cdiv(N, 256 * 4), 64), (unsigned)B);
    fixed_length_kernel<<<grid, 256, 0, s>>>(samples, offsets, N, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
