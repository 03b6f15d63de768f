// fp32 CUDA-core kernels of the projection chain (the EGX_PREC_FP32 arm and the shapes the
// tensor-core GEMM does not take): Linear (GEMM NT + bias/ReLU/addend), LayerNorm, short-sequence
// attention, the prior-pose Conv1d pair, element-wise add.
//
// Follows Full_model/SubLayers.py:30-59,74-84, Full_model/Modules.py:13-23,
// Full_model/Models.py:199-212,411-425.
#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int GBM = 64, GBN = 64, GBK = 16;

// C[M][N] = A[M][K] * W[N][K]^T.  256 threads, 4x4 micro-tile.
template <bool kVec>
__global__ void __launch_bounds__(256)
gemm_nt_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, int M,
                   int N, int K, float* __restrict__ C, int ldc, const float* __restrict__ bias,
                   int relu, const float* __restrict__ addend, int addend_rows, int addend_ld) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int tm = tid & 15, tn = tid >> 4;
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GBK) {
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        const int am = m0 + lr, bn = n0 + lr;
        if (kVec) {
            if (am < M && k0 + lk < K) {
                const float4 t = *reinterpret_cast<const float4*>(A + (size_t)am * lda + k0 + lk);
                a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
            }
            if (bn < N && k0 + lk < K) {
                const float4 t = *reinterpret_cast<const float4*>(Wt + (size_t)bn * K + k0 + lk);
                b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (am < M && k0 + lk + j < K) a[j] = A[(size_t)am * lda + k0 + lk + j];
                if (bn < N && k0 + lk + j < K) b[j] = Wt[(size_t)bn * K + k0 + lk + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[lk + j][lr] = a[j]; Bs[lk + j][lr] = b[j]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 av4 = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
            const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w};
            const float bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
        const int ar = addend ? (addend_rows ? m % addend_rows : m) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= N) continue;
            float t = acc[i][j];
            if (bias) t += bias[n];
            if (relu) t = fmaxf(t, 0.f);
            if (addend) t += addend[(size_t)ar * addend_ld + n];
            C[(size_t)m * ldc + n] = t;
        }
    }
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

// One warp per row; d <= 1024.
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                 const float* __restrict__ b, int rows, int d, float* __restrict__ out,
                 __half* __restrict__ out16) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-6f);
    for (int c = lane; c < d; c += 32) {
        const float y = (xr[c] - mean) * rstd * g[c] + b[c];
        out[(size_t)row * d + c] = y;
        if (out16) out16[(size_t)row * d + c] = __float2half_rn(y);
    }
}

// One CTA per (head, clip).  Everything in shared memory; L <= 64.
template <class T>
__global__ void __launch_bounds__(128)
attention_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk,
                 const T* __restrict__ v, int ldv, int Lq, int Lk, int dk, int dv,
                 float scale, T* __restrict__ out, int ldo) {
    extern __shared__ float sm[];
    float* sq = sm;                         // [Lq][dk]
    float* sk = sq + Lq * dk;               // [Lk][dk+1]
    float* sv = sk + Lk * (dk + 1);         // [Lk][dv]
    float* ss = sv + Lk * dv;               // [Lq][Lk+1]
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x;
    for (int i = tid; i < Lq * dk; i += blockDim.x) {
        const int r = i / dk, c = i % dk;
        sq[i] = float(q[((size_t)b * Lq + r) * ldq + h * dk + c]) * scale;
    }
    for (int i = tid; i < Lk * dk; i += blockDim.x) {
        const int r = i / dk, c = i % dk;
        sk[r * (dk + 1) + c] = float(k[((size_t)b * Lk + r) * ldk + h * dk + c]);
    }
    for (int i = tid; i < Lk * dv; i += blockDim.x) {
        const int r = i / dv, c = i % dv;
        sv[i] = float(v[((size_t)b * Lk + r) * ldv + h * dv + c]);
    }
    __syncthreads();
    for (int i = tid; i < Lq * Lk; i += blockDim.x) {
        const int r = i / Lk, c = i % Lk;
        float a = 0.f;
        for (int t = 0; t < dk; ++t) a = fmaf(sq[r * dk + t], sk[c * (dk + 1) + t], a);
        ss[r * (Lk + 1) + c] = a;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < Lq; r += 4) {
        float* row = ss + r * (Lk + 1);
        float mx = -INFINITY;
        for (int c = lane; c < Lk; c += 32) mx = fmaxf(mx, row[c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < Lk; c += 32) { const float e = expf(row[c] - mx); row[c] = e; sum += e; }
        const float inv = 1.f / warp_sum(sum);
        for (int c = lane; c < Lk; c += 32) row[c] *= inv;
    }
    __syncthreads();
    for (int i = tid; i < Lq * dv; i += blockDim.x) {
        const int r = i / dv, c = i % dv;
        float a = 0.f;
        for (int t = 0; t < Lk; ++t) a = fmaf(ss[r * (Lk + 1) + t], sv[t * dv + c], a);
        out[((size_t)b * Lq + r) * ldo + h * dv + c] = T(a);
    }
}

// Prior-pose encoder front: Conv1d(p->F,k3) -> ReLU -> BN -> Conv1d(F->F,k3) -> ReLU -> BN along
// the pose axis (channels = frames).  One CTA per clip.  The second convolution carries the work (F*F*3 MACs per
// output column): its weights sit in shared memory as [c][k][f] so that a thread producing 4 frames x 1 column
// issues 3 broadcast 16-byte weight loads + 3 activation loads per 12 FMAs.
template <class T>
__global__ void __launch_bounds__(1024)
prior_conv_kernel(const float* __restrict__ prior, int p, int F, int P,
                  const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ s1, const float* __restrict__ t1,
                  const float* __restrict__ w2, const float* __restrict__ b2,
                  const float* __restrict__ s2, const float* __restrict__ t2,
                  T* __restrict__ out, int ldo) {
    extern __shared__ __align__(16) float sm[];
    const int PW = P + 2, F4 = (F + 3) & ~3;
    float* w2s = sm;                  // [F][3][F4]  (input frame c, tap k, output frame f)
    float* sin_ = w2s + F * 3 * F4;   // [p][P+2] zero-padded
    float* mid = sin_ + p * PW;       // [F][P+2] zero-padded
    const int b = blockIdx.x;
    // conv2 weights arrive already in the staged [c][k][f] layout (host-side transpose): a linear 16-byte copy
    for (int i = threadIdx.x; i < F * 3 * F4 / 4; i += blockDim.x)
        reinterpret_cast<float4*>(w2s)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);
    for (int i = threadIdx.x; i < p * PW; i += blockDim.x) {
        const int c = i / PW, x = i % PW - 1;
        sin_[i] = (x >= 0 && x < P) ? prior[((size_t)b * p + c) * P + x] : 0.f;
    }
    for (int i = threadIdx.x; i < F * PW; i += blockDim.x) mid[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < F * P; i += blockDim.x) {
        const int f = i / P, x = i % P;
        float a = b1[f];
        for (int c = 0; c < p; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) a = fmaf(w1[(f * p + c) * 3 + k], sin_[c * PW + x + k], a);
        mid[f * PW + x + 1] = fmaxf(a, 0.f) * s1[f] + t1[f];
    }
    __syncthreads();
    // a thread produces 4 frames x 4 columns: per input frame 3 broadcast weight loads and 6 activations for 48 FMAs
    // (the 4 x 1 version was bound by its shared-memory loads: 6 per 12 FMAs)
    const int XG = (P + 3) / 4;
    for (int i = threadIdx.x; i < (F4 / 4) * XG; i += blockDim.x) {
        const int fg = i / XG, x0 = (i % XG) * 4;
        float a[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) a[j][e] = 0.f;
        for (int c = 0; c < F; ++c) {
            const float* m = mid + c * PW + x0;
            float mv[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) mv[e] = x0 + e < PW ? m[e] : 0.f;
            const float4 wa = *reinterpret_cast<const float4*>(w2s + (c * 3 + 0) * F4 + fg * 4);
            const float4 wb = *reinterpret_cast<const float4*>(w2s + (c * 3 + 1) * F4 + fg * 4);
            const float4 wc = *reinterpret_cast<const float4*>(w2s + (c * 3 + 2) * F4 + fg * 4);
            const float w0[4] = {wa.x, wa.y, wa.z, wa.w}, w1[4] = {wb.x, wb.y, wb.z, wb.w}, w2v[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    a[j][e] = fmaf(w0[j], mv[e], fmaf(w1[j], mv[e + 1], fmaf(w2v[j], mv[e + 2], a[j][e])));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = fg * 4 + j;
            if (f >= F) continue;
            const float bb = b2[f], ss = s2[f], tt = t2[f];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (x0 + e < P) out[((size_t)b * F + f) * ldo + x0 + e] = T(fmaxf(a[j][e] + bb, 0.f) * ss + tt);
        }
    }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i];
        const float4 y = reinterpret_cast<const float4*>(b)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}

__global__ void add_f16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               __half* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i];
        const float4 y = reinterpret_cast<const float4*>(b)[i];
        uint2 u;
        *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(x.x + y.x, x.y + y.y);
        *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(x.z + y.z, x.w + y.w);
        reinterpret_cast<uint2*>(out)[i] = u;
    }
}

inline bool ok() { return cudaGetLastError() == cudaSuccess; }

}  // namespace

int launch_gemm_f32(const float* A, int lda, const float* Wt, int M, int N, int K, float* C,
                    int ldc, const GemmEpi& e, cudaStream_t s) {
    dim3 grid((M + GBM - 1) / GBM, (N + GBN - 1) / GBN);
    const bool vec = (K % 4 == 0) && (lda % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wt)) % 16 == 0);
    if (vec)
        gemm_nt_f32_kernel<true><<<grid, 256, 0, s>>>(A, lda, Wt, M, N, K, C, ldc, e.bias, e.relu,
                                                      e.addend, e.addend_rows, e.addend_ld);
    else
        gemm_nt_f32_kernel<false><<<grid, 256, 0, s>>>(A, lda, Wt, M, N, K, C, ldc, e.bias, e.relu,
                                                       e.addend, e.addend_rows, e.addend_ld);
    return ok() ? 1 : -1;
}

int launch_layernorm(const float* x, const LNW& ln, int rows, int d, float* out, __half* out16,
                     cudaStream_t s) {
    layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, ln.g, ln.b, rows, d, out, out16);
    return ok() ? 1 : -1;
}

template <class T>
int launch_attention(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv,
                     int B, int Lq, int Lk, int n_head, int dk, int dv, T* out, int ldo,
                     cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)Lq * dk + (size_t)Lk * (dk + 1) + (size_t)Lk * dv +
                                         (size_t)Lq * (Lk + 1));
    if (cudaFuncSetAttribute(attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) return -1;
    dim3 grid(n_head, B);
    attention_kernel<T><<<grid, 128, smem, s>>>(q, ldq, k, ldk, v, ldv, Lq, Lk, dk, dv,
                                             1.f / sqrtf((float)dk), out, ldo);
    return ok() ? 1 : -1;
}

template int launch_attention<float>(const float*, int, const float*, int, const float*, int, int, int, int, int,
                                     int, int, float*, int, cudaStream_t);
template int launch_attention<__half>(const __half*, int, const __half*, int, const __half*, int, int, int, int,
                                      int, int, int, __half*, int, cudaStream_t);

template <class T>
int launch_prior_conv(const Weights& w, const float* prior, int B, int p, int F, int P, T* out, int ldo,
                      cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)(p + F) * (P + 2) + (size_t)F * 3 * ((F + 3) & ~3));
    if (cudaFuncSetAttribute(prior_conv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) return -1;
    // one (4-frame, 4-column) item per thread where the geometry allows (TED: 9 x 32 = 288)
    const int items = (((F + 3) & ~3) / 4) * ((P + 3) / 4);
    const int threads = std::min(1024, std::max(128, (items + 31) / 32 * 32));
    prior_conv_kernel<T><<<B, threads, smem, s>>>(prior, p, F, P, w.p_c1w, w.p_c1b, w.p_s1, w.p_t1,
                                              w.p_c2wt, w.p_c2b, w.p_s2, w.p_t2, out, ldo);
    return ok() ? 1 : -1;
}
template int launch_prior_conv<float>(const Weights&, const float*, int, int, int, int, float*, int, cudaStream_t);
template int launch_prior_conv<__half>(const Weights&, const float*, int, int, int, int, __half*, int, cudaStream_t);

int launch_add(const float* a, const float* b, float* out, int64_t n, cudaStream_t s) {
    const int64_t n4 = n / 4;   // callers pass multiples of 4 (d_model % 32 == 0)
    const int grid = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
    add_kernel<<<grid, 256, 0, s>>>(a, b, out, n4);
    return ok() ? 1 : -1;
}

}  // namespace egx

namespace egx {
int launch_add_f16(const float* a, const float* b, __half* out, int64_t n, cudaStream_t s) {
    const int64_t n4 = n / 4;
    const int grid = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
    add_f16_kernel<<<grid, 256, 0, s>>>(a, b, out, n4);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
}  // namespace egx
