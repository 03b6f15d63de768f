// K1 — fused log-mel front-end: pre-emphasis -> Hann STFT (n_fft 1024, hop 512, centred,
// zero padded) -> 513-bin power -> 128 Slaney mels -> log/dB -> InstanceNorm / ref-max.
//
// Follows: model/utils.py:33-38 (F1), utils/data_utils.py:36 (F2/F3, librosa defaults),
// utils/data_utils.py:37 (F4a), model/ResNetSE34V2.py:96-98 (F4b).
//
// One CTA per clip, one warp per STFT frame (round-robin).  A frame is a 1024-point real
// FFT done as a 512-point complex Stockham FFT in the warp's private shared-memory ping-pong
// buffers (fp32, twiddles from a float64-built table), followed by the real-FFT untangling,
// the sparse mel projection (<= 24 taps per mel) and the log.  The clip's (128, W) tile stays
// in shared memory until the per-(clip, mel) statistics are known, so the audio is read once
// and the log-mel written once: algorithmic HBM traffic 4*N + 4*128*W bytes per clip.
#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int kFFT = 1024;
constexpr int kHalf = 512;       // complex FFT length
constexpr int kHop = 512;
constexpr int kBins = 513;
constexpr int kMels = 128;
constexpr int kWarps = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

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

__global__ void __launch_bounds__(kWarps * 32)
logmel_kernel(const float* __restrict__ audio, int N, int n_cols, int mode, int preemph,
              float* __restrict__ out, const float* __restrict__ window,
              const float2* __restrict__ tw512, const float2* __restrict__ tw1024,
              const int* __restrict__ mel_start, const int* __restrict__ mel_ptr,
              const float* __restrict__ mel_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw);                 // [kWarps][2][512]
    float* tile = reinterpret_cast<float*>(fftbuf + kWarps * 2 * kHalf);  // [128][n_cols]
    __shared__ float s_red[kWarps];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* x = audio + (size_t)blockIdx.x * N;
    float2* buf0 = fftbuf + warp * 2 * kHalf;
    float2* buf1 = buf0 + kHalf;

    for (int t = warp; t < n_cols; t += kWarps) {
        // ---- framing + pre-emphasis + window; pack even/odd samples as one complex point ----
        const int base = t * kHop - kFFT / 2;
        for (int n = lane; n < kHalf; n += 32) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * n + e;
                const int sidx = base + j;
                float y = 0.f;
                if (sidx >= 0 && sidx < N) {
                    y = x[sidx];
                    if (preemph) {
                        const float prev = (sidx == 0) ? x[1] : x[sidx - 1];  // reflect pad of 1
                        y = y - 0.97f * prev;
                    }
                }
                v[e] = y * window[j];
            }
            buf0[n] = make_float2(v[0], v[1]);
        }
        __syncwarp();
        // ---- 512-point complex Stockham radix-2 FFT (9 passes) ----
        float2* src = buf0;
        float2* dst = buf1;
#pragma unroll 1
        for (int ns = 1; ns < kHalf; ns <<= 1) {
            const int tw_stride = (kHalf / 2) / ns;
            for (int j = lane; j < kHalf / 2; j += 32) {
                const int k = j & (ns - 1);
                const float2 a = src[j];
                const float2 b = cmul(src[j + kHalf / 2], tw512[k * tw_stride]);
                const int o = ((j - k) << 1) + k;
                dst[o] = make_float2(a.x + b.x, a.y + b.y);
                dst[o + ns] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncwarp();
            float2* tmp = src; src = dst; dst = tmp;
        }
        // ---- untangle to the 513-bin real spectrum, power in place of `dst` ----
        float* power = reinterpret_cast<float*>(dst);          // 513 floats fit in 512 float2
        for (int k = lane; k <= kHalf; k += 32) {
            const float2 zk = src[k & (kHalf - 1)];
            const float2 zn = src[(kHalf - k) & (kHalf - 1)];
            const float2 ev = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
            const float2 od = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
            const float2 r = cmul(od, tw1024[k]);
            const float re = ev.x + r.x, im = ev.y + r.y;
            power[k] = re * re + im * im;
        }
        __syncwarp();
        // ---- sparse mel projection + log ----
#pragma unroll
        for (int mm = 0; mm < kMels / 32; ++mm) {
            const int m = lane + 32 * mm;
            const int p0 = mel_ptr[m], p1 = mel_ptr[m + 1], b0 = mel_start[m];
            float acc = 0.f;
            for (int p = p0; p < p1; ++p) acc = fmaf(mel_w[p], power[b0 + (p - p0)], acc);
            float v;
            if (mode == EGX_LOGMEL_LOG_IN) v = logf(acc + 1e-6f);
            else v = 10.f * log10f(fmaxf(acc, 1e-10f));
            tile[m * n_cols + t] = v;
        }
        __syncwarp();
    }
    __syncthreads();

    float* o = out + (size_t)blockIdx.x * kMels * n_cols;
    if (mode == EGX_LOGMEL_LOG_IN) {
        // InstanceNorm1d: per (clip, mel) over time, biased variance, eps 1e-5, no affine
        for (int m = warp; m < kMels; m += kWarps) {
            const float* row = tile + m * n_cols;
            float s = 0.f;
            for (int t = lane; t < n_cols; t += 32) s += row[t];
            const float mean = warp_sum(s) / n_cols;
            float q = 0.f;
            for (int t = lane; t < n_cols; t += 32) { const float d = row[t] - mean; q += d * d; }
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
#pragma unroll
        for (int w = 1; w < kWarps; ++w) mx = fmaxf(mx, s_red[w]);
        for (int i = threadIdx.x; i < kMels * n_cols; i += blockDim.x)
            o[i] = fmaxf(tile[i] - mx, -80.f);
    }
}

}  // namespace

int launch_logmel(const LogmelTables& t, const float* audio, int B, int N, int n_cols, int mode,
                  int preemph, float* out, cudaStream_t s) {
    const size_t smem = sizeof(float2) * kWarps * 2 * kHalf + sizeof(float) * kMels * n_cols;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    logmel_kernel<<<B, kWarps * 32, smem, s>>>(audio, N, n_cols, mode, preemph, out, t.window,
                                                t.tw512, t.tw1024, t.mel_start, t.mel_ptr,
                                                t.mel_w);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
