// K11 — FGD sufficient statistics on the GPU: n, sum(x - shift), sum((x - shift)(x - shift)^T) in float64, replacing
// the D2H copy + np.mean / np.cov of test_emotion_gesture_diversity_iterative.py:226-232,251-254
// (model/FHD_score.py:240-241).  The caller all-reduces the packed buffer across ranks and forms mu / Sigma (ddof = 1).
//
// The Gram matrix is a float64 syrk, G = X^T X with X = (rows x D) float32 features minus a float64 provisional mean
// (products of two float32 values are exact in float64, so float64 accumulation is what keeps rtol 1e-9 on Sigma).
// It runs on the fp64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA), A = X^T tile, B = X tile — both fragments are the same
// gather X[r0 + lane % 4][c0 + lane / 4].  One CTA owns a 128 x 128 block (i <= j only: the result is symmetric) of
// the Gram matrix over a contiguous slice of rows: 16 warps in a 4 x 4 grid, a 32 x 32 sub-block (16 m8n8 tiles, 32
// accumulator doubles) per warp; 32-row chunks are converted to float64 on their way from global to shared memory
// (row pitch 132 doubles: the 4 x 4 (row, column) pattern of a fragment load hits 16 distinct bank pairs) and double
// buffered, one barrier per chunk.  Partial blocks go to a handle-owned scratch and a second kernel adds them into the
// accumulator in a FIXED order — no atomics, the same rows give the same bits — mirroring the off-diagonal blocks.
#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int GT = 128;            // Gram block edge
constexpr int KC = 32;             // rows per chunk
constexpr int PITCH = GT + 4;      // doubles per shared-memory row
constexpr int kFgdThreads = 512;

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// block pair p -> (bi <= bj) in row-major order of the upper triangle
__device__ __forceinline__ void pair_to_blocks(int p, int nb, int* bi, int* bj) {
    int i = 0;
    while (p >= nb - i) { p -= nb - i; ++i; }
    *bi = i; *bj = i + p;
}

struct Chunk { float4 v[2][2]; };      // [operand][half]: this thread's 2 x 4 columns of the next 32-row chunk

__device__ __forceinline__ float4 load4(const float* __restrict__ x, int64_t row, int64_t r1, int D, int col, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < r1) {
        const float* p = x + row * D + col;
        if (vec && col + 3 < D) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
            if (col < D) v.x = __ldg(p);
            if (col + 1 < D) v.y = __ldg(p + 1);
            if (col + 2 < D) v.z = __ldg(p + 2);
            if (col + 3 < D) v.w = __ldg(p + 3);
        }
    }
    return v;
}

__global__ void __launch_bounds__(kFgdThreads, 1)
fgd_gram_kernel(const float* __restrict__ x, int64_t n, int D, const double* __restrict__ shift, int64_t rows_per_split,
                int nb, double* __restrict__ part_g, double* __restrict__ part_s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);          // [buf 2][operand 2][KC][PITCH]
    __shared__ double colsum[16][GT];

    int bi, bj;
    pair_to_blocks(blockIdx.x, nb, &bi, &bj);
    const bool diag = bi == bj;
    const int split = blockIdx.y, n_split = gridDim.y;
    const int64_t r0 = (int64_t)split * rows_per_split;
    const int64_t r1 = r0 + rows_per_split < n ? r0 + rows_per_split : n;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int wi = warp >> 2, wj = warp & 3;                      // 4 x 4 warps, 32 x 32 per warp
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;

    // loader mapping: float4 column group c4 (fixed per thread: its column sums stay in registers), rows lr and lr + 16
    const int c4 = t & 31, lr = t >> 5;
    const int col_i = bi * GT + 4 * c4, col_j = bj * GT + 4 * c4;
    double sh_i[4], sh_j[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        sh_i[e] = (shift && col_i + e < D) ? shift[col_i + e] : 0.0;
        sh_j[e] = (shift && col_j + e < D) ? shift[col_j + e] : 0.0;
    }
    double csum[4] = {0.0, 0.0, 0.0, 0.0};

    auto fetch = [&](int64_t row0, Chunk& c) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            c.v[0][h] = load4(x, row0 + lr + 16 * h, r1, D, col_i, vec);
            if (!diag) c.v[1][h] = load4(x, row0 + lr + 16 * h, r1, D, col_j, vec);
        }
    };
    auto stage = [&](int64_t row0, const Chunk& c, int buf) {
        double* bi_ = sm + (size_t)(buf * 2 + 0) * KC * PITCH;
        double* bj_ = sm + (size_t)(buf * 2 + 1) * KC * PITCH;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 16 * h;
            const bool live = row0 + r < r1;
            const float vi[4] = {c.v[0][h].x, c.v[0][h].y, c.v[0][h].z, c.v[0][h].w};
            const float vj[4] = {c.v[1][h].x, c.v[1][h].y, c.v[1][h].z, c.v[1][h].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // rows past the slice and columns past D contribute exact zeros
                const double a = (live && col_i + e < D) ? (double)vi[e] - sh_i[e] : 0.0;
                bi_[r * PITCH + 4 * c4 + e] = a;
                csum[e] += a;
                if (!diag) bj_[r * PITCH + 4 * c4 + e] = (live && col_j + e < D) ? (double)vj[e] - sh_j[e] : 0.0;
            }
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    Chunk nxt;
    if (r0 < r1) {
        fetch(r0, nxt);
        stage(r0, nxt, 0);
    }
    __syncthreads();
    int buf = 0;
    for (int64_t row0 = r0; row0 < r1; row0 += KC, buf ^= 1) {
        const bool more = row0 + KC < r1;
        if (more) fetch(row0 + KC, nxt);                         // global loads in flight under the MMAs
        const double* a_s = sm + (size_t)(buf * 2 + 0) * KC * PITCH + 32 * wi + (lane >> 2);
        const double* b_s = sm + (size_t)(buf * 2 + (diag ? 0 : 1)) * KC * PITCH + 32 * wj + (lane >> 2);
#pragma unroll 2
        for (int k = 0; k < KC; k += 4) {
            double fa[4], fb[4];
            const int ro = (k + (lane & 3)) * PITCH;
#pragma unroll
            for (int m = 0; m < 4; ++m) { fa[m] = a_s[ro + 8 * m]; fb[m] = b_s[ro + 8 * m]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b], fa[a], fb[b]);
        }
        if (more) stage(row0 + KC, nxt, buf ^ 1);
        __syncthreads();
    }

    // partial block: row i = 32 wi + 8 a + lane / 4, columns 32 wj + 8 b + 2 (lane % 4) + {0, 1}
    double* pg = part_g + ((size_t)blockIdx.x * n_split + split) * GT * GT;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = 32 * wi + 8 * a + (lane >> 2), j = 32 * wj + 8 * b + 2 * (lane & 3);
            *reinterpret_cast<double2*>(pg + (size_t)i * GT + j) = make_double2(acc[a][b][0], acc[a][b][1]);
        }
    if (diag) {
        // column sums of this slice: 16 loader rows per column group, reduced in a fixed order
#pragma unroll
        for (int e = 0; e < 4; ++e) colsum[lr][4 * c4 + e] = csum[e];
        __syncthreads();
        if (t < GT) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) s += colsum[r][t];
            part_s[((size_t)bi * n_split + split) * GT + t] = s;
        }
    }
}

// acc += sum over the row slices (ascending) of the partial blocks; off-diagonal blocks are mirrored
__global__ void fgd_reduce_kernel(const double* __restrict__ part_g, const double* __restrict__ part_s, int n_split, int nb,
                                  int D, double n_rows, double* __restrict__ acc) {
    const int p = blockIdx.x;
    int bi, bj;
    pair_to_blocks(p, nb, &bi, &bj);
    // blockIdx.y: one of 32 strips of 4 block rows; every thread owns one element, its slices summed in ascending order
    // (four independent loads in flight)
    {
        const int e = blockIdx.y * 512 + threadIdx.x;
        const int i = bi * GT + e / GT, j = bj * GT + e % GT;
        if (i < D && j < D) {
            const double* src = part_g + (size_t)p * n_split * GT * GT + e;
            double s = 0.0;
            int k = 0;
            for (; k + 4 <= n_split; k += 4) {
                const double v0 = src[(size_t)k * GT * GT], v1 = src[(size_t)(k + 1) * GT * GT];
                const double v2 = src[(size_t)(k + 2) * GT * GT], v3 = src[(size_t)(k + 3) * GT * GT];
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; k < n_split; ++k) s += src[(size_t)k * GT * GT];
            acc[1 + D + (size_t)i * D + j] += s;
            if (bi != bj) acc[1 + D + (size_t)j * D + i] += s;
        }
    }
    if (bi == bj && blockIdx.y == 0)
        for (int c = threadIdx.x; c < GT; c += blockDim.x) {
            const int col = bi * GT + c;
            if (col >= D) continue;
            double s = 0.0;
            for (int k = 0; k < n_split; ++k) s += part_s[((size_t)bi * n_split + k) * GT + c];
            acc[1 + col] += s;
        }
    if (p == 0 && blockIdx.y == 0 && threadIdx.x == 0) acc[0] += n_rows;
}

constexpr int kFgdSmem = 2 * 2 * KC * PITCH * (int)sizeof(double);

}  // namespace

// Scratch the two-phase reduction needs for (n rows, D): returned through *n_split_out as well.
size_t fgd_scratch_doubles(int64_t n, int D, int sms, int* n_split_out) {
    const int nb = (D + GT - 1) / GT, pairs = nb * (nb + 1) / 2;
    int64_t splits = std::max<int64_t>(1, (2LL * sms + pairs - 1) / pairs);
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, (n + 8 * KC - 1) / (8 * KC)));    // at least 256 rows per slice
    *n_split_out = (int)splits;
    return (size_t)pairs * splits * GT * GT + (size_t)nb * splits * GT;
}

int launch_fgd_accumulate(const float* feats, int64_t n, int D, const double* shift, double* acc, double* scratch,
                          int n_split, cudaStream_t s) {
    if (n <= 0) return 0;
    const int nb = (D + GT - 1) / GT, pairs = nb * (nb + 1) / 2;
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(fgd_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFgdSmem) != cudaSuccess) return -1;
        configured[dev] = true;
    }
    int64_t rps = (n + n_split - 1) / n_split;
    rps = (rps + KC - 1) / KC * KC;
    double* part_g = scratch;
    double* part_s = scratch + (size_t)pairs * n_split * GT * GT;
    fgd_gram_kernel<<<dim3(pairs, n_split), kFgdThreads, kFgdSmem, s>>>(feats, n, D, shift, rps, nb, part_g, part_s);
    if (cudaGetLastError() != cudaSuccess) return -1;
    fgd_reduce_kernel<<<dim3(pairs, GT * GT / 512), 512, 0, s>>>(part_g, part_s, n_split, nb, D, (double)n, acc);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

}  // namespace egx
