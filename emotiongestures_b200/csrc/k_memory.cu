// Prior_MemoryEncoder of the checkpointed generator (Full_model/Models_memory.py:215-345), the part between its
// pred_conv and its post_header.  Per clip b, with x = prior poses (p, P), pred = pred_conv(x) (n_pred, P),
// C = args.chunk and the two chunk encoders already applied to x[p-C:].reshape(C*P) (enc[b] = m_sp | m_tm, 2P):
//   SP_Memory_Net_v1.forward (:233-249): for c < C: s = sigmoid(<m_sp, pred[c]>); pred[c] = s*pred[c] + (1-s)*m_sp
//   TM_Memory_Net.forward (:282-293):    e[b] = temporal_memory_encoder(pred[:C].reshape(C*P))          (C)
//                                        S = m_tm^T e   — a (P, C) matrix summed over the WHOLE BATCH of the call
//                                        soft = softmax(m_tm[b] S);  pred[c] *= 1 + soft[c]  for c < C
//   out[b] = cat(x, pred) (F, P)                                                                    (:339-342)
// All fp32 (the sigmoid / softmax scores are sums over hundreds of products; the tensor-core arm only takes over at
// post_header).  The batch sum runs in a fixed order, so a given batch always produces the same bits; like the
// reference under nn.DataParallel, each replica / call couples only its own clips.
#include "egx_common.cuh"

namespace egx {

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {     // all threads get the sum; fixed order
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// grid B.  pred (B, n_pred, P) in place for rows c < C; enc (B, 2P); tm_w [C][C*P], tm_b [C] -> pred_enc (B, C)
__global__ void __launch_bounds__(256)
mem_spatial_kernel(float* __restrict__ pred, int n_pred, int P, int C, const float* __restrict__ enc,
                   const float* __restrict__ tm_w, const float* __restrict__ tm_b, float* __restrict__ pred_enc) {
    extern __shared__ float sm[];
    float* m_sp = sm;              // [P]
    float* rows = sm + P;          // [C][P] blended rows
    __shared__ float red[8];
    const int b = blockIdx.x;
    float* pb = pred + (size_t)b * n_pred * P;
    for (int x = threadIdx.x; x < P; x += blockDim.x) m_sp[x] = enc[(size_t)b * 2 * P + x];
    __syncthreads();
    for (int c = 0; c < C; ++c) {
        float part = 0.f;
        for (int x = threadIdx.x; x < P; x += blockDim.x) part = fmaf(m_sp[x], pb[c * P + x], part);
        const float score = block_sum(part, red);
        const float s = 1.f / (1.f + expf(-score));
        for (int x = threadIdx.x; x < P; x += blockDim.x) {
            const float v = s * pb[c * P + x] + (1.f - s) * m_sp[x];
            rows[c * P + x] = v;
            pb[c * P + x] = v;
        }
    }
    __syncthreads();
    // temporal_memory_encoder on the blended rows: one warp per output
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < C; j += nw) {
        float t = 0.f;
        const float* wr = tm_w + (size_t)j * C * P;
        for (int i = lane; i < C * P; i += 32) t = fmaf(wr[i], rows[i], t);
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) pred_enc[(size_t)b * C + j] = t + tm_b[j];
    }
}

// S[x][c] = sum_b m_tm[b][x] * pred_enc[b][c], b ascending (one thread per element: deterministic)
__global__ void __launch_bounds__(256)
mem_batch_outer_kernel(const float* __restrict__ enc, const float* __restrict__ pred_enc, int B, int P, int C,
                       float* __restrict__ S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * C) return;
    const int x = i / C, c = i % C;
    float a = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) a = fmaf(enc[(size_t)b * 2 * P + P + x], pred_enc[(size_t)b * C + c], a);
    S[i] = a;
}

// grid B.  soft = softmax_c(m_tm[b] . S[:, c]); out[b] = cat(x[b], pred[b] * (1 + soft[c] for c < C)) as T, pitch ldo
template <class T>
__global__ void __launch_bounds__(256)
mem_temporal_kernel(const float* __restrict__ prior, const float* __restrict__ pred, int p, int n_pred, int P, int C,
                    const float* __restrict__ enc, const float* __restrict__ S, T* __restrict__ out, int ldo) {
    extern __shared__ float sm[];
    float* gain = sm;              // [C] 1 + soft
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* m_tm = enc + (size_t)b * 2 * P + P;
    for (int c = 0; c < C; ++c) {
        float part = 0.f;
        for (int x = threadIdx.x; x < P; x += blockDim.x) part = fmaf(m_tm[x], S[x * C + c], part);
        const float score = block_sum(part, red);
        if (threadIdx.x == 0) gain[c] = score;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = -INFINITY, sum = 0.f;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, gain[c]);
        for (int c = 0; c < C; ++c) { gain[c] = expf(gain[c] - mx); sum += gain[c]; }
        for (int c = 0; c < C; ++c) gain[c] = 1.f + gain[c] / sum;
    }
    __syncthreads();
    const int F = p + n_pred;
    for (int i = threadIdx.x; i < F * P; i += blockDim.x) {
        const int f = i / P, x = i % P;
        float v;
        if (f < p) v = prior[((size_t)b * p + f) * P + x];
        else {
            const int c = f - p;
            v = pred[((size_t)b * n_pred + c) * P + x];
            if (c < C) v *= gain[c];
        }
        out[((size_t)b * F + f) * ldo + x] = T(v);
    }
}

}  // namespace

int launch_mem_spatial(float* pred, int B, int n_pred, int P, int C, const float* enc, const float* tm_w, const float* tm_b,
                       float* pred_enc, cudaStream_t s) {
    const size_t smem = sizeof(float) * (size_t)(P + C * P);
    if (smem > 48 * 1024) return -1;
    mem_spatial_kernel<<<B, 256, smem, s>>>(pred, n_pred, P, C, enc, tm_w, tm_b, pred_enc);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_mem_batch_outer(const float* enc, const float* pred_enc, int B, int P, int C, float* S, cudaStream_t s) {
    mem_batch_outer_kernel<<<(P * C + 255) / 256, 256, 0, s>>>(enc, pred_enc, B, P, C, S);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <class T>
int launch_mem_temporal(const float* prior, const float* pred, int B, int p, int n_pred, int P, int C, const float* enc,
                        const float* S, T* out, int ldo, cudaStream_t s) {
    mem_temporal_kernel<T><<<B, 256, sizeof(float) * C, s>>>(prior, pred, p, n_pred, P, C, enc, S, out, ldo);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
template int launch_mem_temporal<float>(const float*, const float*, int, int, int, int, int, const float*, const float*,
                                        float*, int, cudaStream_t);
template int launch_mem_temporal<__half>(const float*, const float*, int, int, int, int, int, const float*, const float*,
                                         __half*, int, cudaStream_t);

}  // namespace egx
