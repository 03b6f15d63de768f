// K11 — FGD sufficient statistics on the GPU: n, sum(x - shift), sum((x - shift)(x - shift)^T)
// in float64, replacing the D2H copy + np.mean / np.cov of
// test_emotion_gesture_diversity_iterative.py:226-232,251-254 (model/FHD_score.py:240-241).
// The caller all-reduces the packed buffer across ranks and forms mu / Sigma (ddof = 1).
#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int FT = 16;     // gram tile edge
constexpr int FR = 64;     // rows staged per step

__global__ void __launch_bounds__(FT * FT)
fgd_kernel(const float* __restrict__ x, int64_t n, int D, const double* __restrict__ shift,
           int64_t rows_per_split, double* __restrict__ acc) {
    __shared__ double xi[FR][FT + 1], xj[FR][FT + 1];
    const int ti = threadIdx.x / FT, tj = threadIdx.x % FT;
    const int i0 = blockIdx.x * FT, j0 = blockIdx.y * FT;
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r1 = r0 + rows_per_split < n ? r0 + rows_per_split : n;
    double g = 0.0, s = 0.0;
    for (int64_t r = r0; r < r1; r += FR) {
        for (int e = threadIdx.x; e < FR * FT; e += FT * FT) {
            const int rr = e / FT, c = e % FT;
            const int64_t row = r + rr;
            double a = 0.0, b = 0.0;
            if (row < r1) {
                if (i0 + c < D) a = (double)x[row * D + i0 + c] - (shift ? shift[i0 + c] : 0.0);
                if (j0 + c < D) b = (double)x[row * D + j0 + c] - (shift ? shift[j0 + c] : 0.0);
            }
            xi[rr][c] = a;
            xj[rr][c] = b;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < FR; ++rr) {
            g = fma(xi[rr][ti], xj[rr][tj], g);
            if (tj == 0) s += xi[rr][ti];
        }
        __syncthreads();
    }
    if (i0 + ti < D && j0 + tj < D) atomicAdd(&acc[1 + D + (size_t)(i0 + ti) * D + j0 + tj], g);
    if (blockIdx.y == 0 && tj == 0 && i0 + ti < D) atomicAdd(&acc[1 + i0 + ti], s);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&acc[0], (double)(r1 - r0));
}

}  // namespace

int launch_fgd_accumulate(const float* feats, int64_t n, int D, const double* shift, double* acc,
                          cudaStream_t s) {
    if (n <= 0) return 0;
    const int tiles = (D + FT - 1) / FT;
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (148 * 8) / (tiles * tiles) + 1));
    const int64_t rps = (n + splits - 1) / splits;
    splits = (int)((n + rps - 1) / rps);
    dim3 grid(tiles, tiles, splits);
    fgd_kernel<<<grid, FT * FT, 0, s>>>(feats, n, D, shift, rps, acc);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
