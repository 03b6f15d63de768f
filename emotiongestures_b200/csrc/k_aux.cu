// Small networks either side of the generator.  None of them is GEMM-shaped enough
// to justify tensor cores (a few MFLOP per sample, channel counts 4..64); they are written as fused CUDA-core
// kernels that keep every intermediate of one sample in shared memory, so HBM sees the inputs once and the outputs
// once.  Eval-mode algebra is folded on the host (egx_api.cu): Linear chains without an activation between them
// (nn.Dropout is the identity in eval) collapse into one affine map, BatchNorm becomes scale/shift.
//
//   C4  Full_model/BEAT_CVAE.py:98-136   MLP_Reconstruct.forward / .sample      -> cvae_mlp_kernel
//   E1  CAVE/BEAT_CVAE.py:427-447        MLP_Reconstruct_v3.sample               -> cvae3_sample_kernel
//   D1  model/motion_ae.py:55-62, model/embedding_net.py:67-83  PoseEncoderConv  -> pose_encoder_kernel
#include "egx_common.cuh"

namespace egx {

namespace {

__device__ __forceinline__ float lrelu02(float x) { return x > 0.f ? x : 0.2f * x; }

// ---------------------------------------------------------------------------------------------
// C4: mu = A_mu x + b_mu, lv = A_lv x + b_lv, py = A_y y + c_y, z = eps * exp(lv / 2) + mu (or z given),
//     out = A_d [z ; py] + d.       Dimensions: x, y, out: 90; mu, lv, z, py: 32.
// One thread per sample (lanes = samples), all weights broadcast from shared memory, inputs staged through a
// padded shared tile so that global traffic is coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int CV_IN = 90, CV_Z = 32, CV_TILE = 128, CV_PITCH = 91;

struct CvaeSmem {
    float w_x[CV_IN][2 * CV_Z];      // [k][mu | lv]
    float w_y[CV_IN][CV_Z];
    float w_d[2 * CV_Z][92];         // [k][out], 90 padded to 92
    float b_x[2 * CV_Z], b_y[CV_Z], b_d[92];
    float tile[CV_TILE * CV_PITCH];  // x, then y, then the outputs
};

__global__ void __launch_bounds__(CV_TILE, 1)
cvae_mlp_kernel(CvaeW w, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ noise,
                int noise_is_z, int64_t n, float* __restrict__ out, float* __restrict__ mu_out,
                float* __restrict__ lv_out) {
    extern __shared__ unsigned char cv_raw[];
    CvaeSmem& sm = *reinterpret_cast<CvaeSmem*>(cv_raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < CV_IN * 2 * CV_Z; i += CV_TILE) (&sm.w_x[0][0])[i] = w.w_x[i];
    for (int i = tid; i < CV_IN * CV_Z; i += CV_TILE) (&sm.w_y[0][0])[i] = w.w_y[i];
    for (int i = tid; i < 2 * CV_Z * 92; i += CV_TILE) (&sm.w_d[0][0])[i] = w.w_d[i];
    if (tid < 2 * CV_Z) sm.b_x[tid] = w.b_x[tid];
    if (tid < CV_Z) sm.b_y[tid] = w.b_y[tid];
    if (tid < 92) sm.b_d[tid] = w.b_d[tid];
    for (int64_t base = (int64_t)blockIdx.x * CV_TILE; base < n; base += (int64_t)gridDim.x * CV_TILE) {
        const int cnt = (int)(n - base < CV_TILE ? n - base : CV_TILE);
        const bool live = tid < cnt;
        float zin[2 * CV_Z];      // [z ; py]
        __syncthreads();
        if (!noise_is_z) {
            for (int i = tid; i < cnt * CV_IN; i += CV_TILE) sm.tile[(i / CV_IN) * CV_PITCH + i % CV_IN] = x[base * CV_IN + i];
            __syncthreads();
            float acc[2 * CV_Z];
#pragma unroll
            for (int o = 0; o < 2 * CV_Z; ++o) acc[o] = sm.b_x[o];
            for (int k = 0; k < CV_IN; ++k) {
                const float xv = sm.tile[tid * CV_PITCH + k];
#pragma unroll
                for (int o4 = 0; o4 < 2 * CV_Z / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&sm.w_x[k][4 * o4]);
                    acc[4 * o4] = fmaf(wv.x, xv, acc[4 * o4]);         acc[4 * o4 + 1] = fmaf(wv.y, xv, acc[4 * o4 + 1]);
                    acc[4 * o4 + 2] = fmaf(wv.z, xv, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(wv.w, xv, acc[4 * o4 + 3]);
                }
            }
            if (live) {
                float4* mo = reinterpret_cast<float4*>(mu_out + (base + tid) * CV_Z);
                float4* lo = reinterpret_cast<float4*>(lv_out + (base + tid) * CV_Z);
                const float4* ep = reinterpret_cast<const float4*>(noise + (base + tid) * CV_Z);
#pragma unroll
                for (int o4 = 0; o4 < CV_Z / 4; ++o4) {
                    mo[o4] = make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]);
                    lo[o4] = make_float4(acc[CV_Z + 4 * o4], acc[CV_Z + 4 * o4 + 1], acc[CV_Z + 4 * o4 + 2], acc[CV_Z + 4 * o4 + 3]);
                    const float4 e = ep[o4];
                    const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)      // reparameterize: eps * exp(0.5 * logvar) + mu  (BEAT_CVAE.py:91-94)
                        zin[4 * o4 + j] = fmaf(ev[j], expf(0.5f * acc[CV_Z + 4 * o4 + j]), acc[4 * o4 + j]);
                }
            }
            __syncthreads();
        } else if (live) {
            const float4* zp = reinterpret_cast<const float4*>(noise + (base + tid) * CV_Z);
#pragma unroll
            for (int o4 = 0; o4 < CV_Z / 4; ++o4) {
                const float4 e = zp[o4];
                zin[4 * o4] = e.x; zin[4 * o4 + 1] = e.y; zin[4 * o4 + 2] = e.z; zin[4 * o4 + 3] = e.w;
            }
        }
        for (int i = tid; i < cnt * CV_IN; i += CV_TILE) sm.tile[(i / CV_IN) * CV_PITCH + i % CV_IN] = y[base * CV_IN + i];
        __syncthreads();
        {
            float acc[CV_Z];
#pragma unroll
            for (int o = 0; o < CV_Z; ++o) acc[o] = sm.b_y[o];
            for (int k = 0; k < CV_IN; ++k) {
                const float yv = sm.tile[tid * CV_PITCH + k];
#pragma unroll
                for (int o4 = 0; o4 < CV_Z / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&sm.w_y[k][4 * o4]);
                    acc[4 * o4] = fmaf(wv.x, yv, acc[4 * o4]);         acc[4 * o4 + 1] = fmaf(wv.y, yv, acc[4 * o4 + 1]);
                    acc[4 * o4 + 2] = fmaf(wv.z, yv, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(wv.w, yv, acc[4 * o4 + 3]);
                }
            }
#pragma unroll
            for (int o = 0; o < CV_Z; ++o) zin[CV_Z + o] = acc[o];
        }
        __syncthreads();      // everyone is done reading y from the tile; reuse it for the outputs
#pragma unroll 1
        for (int o0 = 0; o0 < 92; o0 += 4) {
            float4 a = *reinterpret_cast<const float4*>(&sm.b_d[o0]);
#pragma unroll
            for (int k = 0; k < 2 * CV_Z; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(&sm.w_d[k][o0]);
                a.x = fmaf(wv.x, zin[k], a.x); a.y = fmaf(wv.y, zin[k], a.y);
                a.z = fmaf(wv.z, zin[k], a.z); a.w = fmaf(wv.w, zin[k], a.w);
            }
            float* t = &sm.tile[tid * CV_PITCH + o0];
            t[0] = a.x; t[1] = a.y;
            if (o0 + 2 < CV_IN) { t[2] = a.z; t[3] = a.w; }
        }
        __syncthreads();
        for (int i = tid; i < cnt * CV_IN; i += CV_TILE) out[base * CV_IN + i] = sm.tile[(i / CV_IN) * CV_PITCH + i % CV_IN];
    }
}

// ---------------------------------------------------------------------------------------------
// E1: CVAE sampler.  One CTA per sample; every feature map of the decoder lives in shared memory.
//   h0 (4,128) = A_f [z ; A_y y + c_y] + b_f
//   ConvT1d(4->8,k3,s2,p1,op1) LReLU BN -> (8,256); ConvT1d(8->16) LReLU BN -> (16,512);
//   Conv1d(16->32) LReLU BN; Conv1d(32->60) LReLU BN; Conv1d(60->60) -> out (60,512)
// ---------------------------------------------------------------------------------------------
constexpr int E1_THREADS = 256;

// y[co][t] = b[co] + sum_ci sum_k x[ci][i] w[ci][co][k],  t = 2 i - 1 + k  (ConvTranspose1d k3 s2 p1 op1: Lout = 2 Lin)
__device__ void convt_s2(const float* __restrict__ x, int cin, int lin, const float* __restrict__ w,
                         const float* __restrict__ b, const float* __restrict__ sc, const float* __restrict__ sh,
                         int cout, float* __restrict__ yv) {
    const int lout = 2 * lin;
    for (int e = threadIdx.x; e < cout * lout; e += E1_THREADS) {
        const int co = e / lout, t = e % lout;
        float acc = b[co];
        if (t & 1) {              // k = 0 -> i = (t+1)/2 ; k = 2 -> i = (t-1)/2
            const int i0 = (t + 1) >> 1, i2 = (t - 1) >> 1;
            for (int ci = 0; ci < cin; ++ci) {
                const float* wp = w + (ci * cout + co) * 3;
                if (i0 < lin) acc = fmaf(x[ci * lin + i0], wp[0], acc);
                acc = fmaf(x[ci * lin + i2], wp[2], acc);
            }
        } else {                  // k = 1 -> i = t/2
            const int i1 = t >> 1;
            for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci * lin + i1], w[(ci * cout + co) * 3 + 1], acc);
        }
        yv[e] = fmaf(lrelu02(acc), sc[co], sh[co]);
    }
}

// Conv1d k3 p1 over [cin][len] in shared memory; thread tile 4 positions x COB output channels.
template <int COB, bool LAST>
__device__ void conv3_p1(const float* __restrict__ x, int cin, int len, const float* __restrict__ w,
                         const float* __restrict__ b, const float* __restrict__ sc, const float* __restrict__ sh,
                         int cout, float* __restrict__ yv) {
    const int tq = len / 4;
    const int n_items = (cout / COB) * tq;
    for (int e = threadIdx.x; e < n_items; e += E1_THREADS) {
        const int cg = e / tq, t0 = (e % tq) * 4;
        float acc[COB][4];
#pragma unroll
        for (int c = 0; c < COB; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][j] = b[cg * COB + c];
        for (int ci = 0; ci < cin; ++ci) {
            const float* xr = x + ci * len + t0;
            float xv[6];
            xv[0] = t0 > 0 ? xr[-1] : 0.f;
            const float4 m = *reinterpret_cast<const float4*>(xr);
            xv[1] = m.x; xv[2] = m.y; xv[3] = m.z; xv[4] = m.w;
            xv[5] = t0 + 4 < len ? xr[4] : 0.f;
#pragma unroll
            for (int c = 0; c < COB; ++c) {
                const float* wp = w + ((cg * COB + c) * cin + ci) * 3;
                const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[c][j] = fmaf(xv[j], w0, fmaf(xv[j + 1], w1, fmaf(xv[j + 2], w2, acc[c][j])));
            }
        }
#pragma unroll
        for (int c = 0; c < COB; ++c) {
            const int co = cg * COB + c;
            float4 o;
            if (LAST) o = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
            else o = make_float4(fmaf(lrelu02(acc[c][0]), sc[co], sh[co]), fmaf(lrelu02(acc[c][1]), sc[co], sh[co]),
                                 fmaf(lrelu02(acc[c][2]), sc[co], sh[co]), fmaf(lrelu02(acc[c][3]), sc[co], sh[co]));
            *reinterpret_cast<float4*>(yv + co * len + t0) = o;
        }
    }
}

__global__ void __launch_bounds__(E1_THREADS, 1)
cvae3_sample_kernel(Cvae3W w, const float* __restrict__ y, const float* __restrict__ z, int n, float* __restrict__ out) {
    extern __shared__ float e1_sm[];
    float* regA = e1_sm;                  // 60 * 512 floats: a2 (16,512), later a4 (60,512)
    float* regB = e1_sm + 60 * 512;       // 32 * 512 floats: zin, h0, a1, later a3 (32,512)
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
        float* zin = regB;                // 64
        float* h0 = regB + 64;            // 512
        float* a1 = regB + 64 + 512;      // 8 * 256
        __syncthreads();
        if (threadIdx.x < 32) zin[threadIdx.x] = z[(size_t)s * 32 + threadIdx.x];
        else if (threadIdx.x < 64) {
            const int o = threadIdx.x - 32;
            float acc = w.b_y[o];
            for (int k = 0; k < 8; ++k) acc = fmaf(w.w_y[o * 8 + k], y[(size_t)s * 8 + k], acc);
            zin[32 + o] = acc;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < 512; o += E1_THREADS) {
            float acc = w.b_f[o];
            const float4* wr = reinterpret_cast<const float4*>(w.w_f + (size_t)o * 64);
#pragma unroll
            for (int k4 = 0; k4 < 16; ++k4) {
                const float4 wv = __ldg(wr + k4);
                acc = fmaf(wv.x, zin[4 * k4], fmaf(wv.y, zin[4 * k4 + 1], fmaf(wv.z, zin[4 * k4 + 2], fmaf(wv.w, zin[4 * k4 + 3], acc))));
            }
            h0[o] = acc;
        }
        __syncthreads();
        convt_s2(h0, 4, 128, w.t1_w, w.t1_b, w.s1, w.h1, 8, a1);
        __syncthreads();
        convt_s2(a1, 8, 256, w.t2_w, w.t2_b, w.s2, w.h2, 16, regA);
        __syncthreads();
        conv3_p1<4, false>(regA, 16, 512, w.c3_w, w.c3_b, w.s3, w.h3, 32, regB);
        __syncthreads();
        conv3_p1<4, false>(regB, 32, 512, w.c4_w, w.c4_b, w.s4, w.h4, 60, regA);
        __syncthreads();
        conv3_p1<4, true>(regA, 60, 512, w.c5_w, w.c5_b, nullptr, nullptr, 60, out + (size_t)s * 60 * 512);
    }
}

// ---------------------------------------------------------------------------------------------
// D1: PoseEncoderConv.  poses (B, L, P) -> Conv1d(P->32,k3) BN LReLU -> Conv1d(32->64,k3) BN LReLU ->
//     Conv1d(64->64,k4,s2) BN LReLU -> Conv1d(64->32,k3) -> flatten [c][t] -> one affine map (out_net, and fc_mu for
//     embedding_net, collapse: LeakyReLU(True) has slope 1.0, BatchNorm1d is affine in eval).
// One CTA per clip; BatchNorm folded into the conv weights/bias on the host.
// ---------------------------------------------------------------------------------------------
constexpr int PE_THREADS = 256;

__global__ void __launch_bounds__(PE_THREADS, 1)
pose_encoder_kernel(PoseEncW w, const float* __restrict__ poses, int B, float* __restrict__ out) {
    extern __shared__ float pe_sm[];
    const int L = w.L, P = w.P, PP = P | 1;           // odd pitch: lanes walk t with stride PP, conflict-free
    const int L1 = L - 2, L2 = L - 4, L3 = (L2 - 4) / 2 + 1, L4 = L3 - 2;
    float* xin = pe_sm;                   // [L][PP]
    float* a1 = xin + L * PP;             // [32][L1]
    float* a2 = a1 + 32 * L1;             // [64][L2]
    float* a3 = a2 + 64 * L2;             // [64][L3]
    float* a4 = a3 + 64 * L3;             // [32][L4]  (flatten order c * L4 + t)
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < L * P; i += PE_THREADS) xin[(i / P) * PP + i % P] = poses[(size_t)b * L * P + i];
        __syncthreads();
        // conv1: 4 output channels per thread
        for (int e = threadIdx.x; e < 8 * L1; e += PE_THREADS) {
            const int cg = e / L1, t = e % L1;
            float acc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = w.b1[cg * 4 + c];
            for (int p = 0; p < P; ++p) {
                const float x0 = xin[t * PP + p], x1 = xin[(t + 1) * PP + p], x2 = xin[(t + 2) * PP + p];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* wp = w.w1 + ((size_t)(cg * 4 + c) * P + p) * 3;
                    acc[c] = fmaf(x0, __ldg(wp), fmaf(x1, __ldg(wp + 1), fmaf(x2, __ldg(wp + 2), acc[c])));
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) a1[(cg * 4 + c) * L1 + t] = lrelu02(acc[c]);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * L2; e += PE_THREADS) {
            const int cg = e / L2, t = e % L2;
            float acc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = w.b2[cg * 4 + c];
            for (int ci = 0; ci < 32; ++ci) {
                const float x0 = a1[ci * L1 + t], x1 = a1[ci * L1 + t + 1], x2 = a1[ci * L1 + t + 2];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* wp = w.w2 + ((cg * 4 + c) * 32 + ci) * 3;
                    acc[c] = fmaf(x0, __ldg(wp), fmaf(x1, __ldg(wp + 1), fmaf(x2, __ldg(wp + 2), acc[c])));
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) a2[(cg * 4 + c) * L2 + t] = lrelu02(acc[c]);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * L3; e += PE_THREADS) {
            const int co = e / L3, t = e % L3;
            float acc = w.b3[co];
            for (int ci = 0; ci < 64; ++ci) {
                const float* xr = a2 + ci * L2 + 2 * t;
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w.w3 + (co * 64 + ci) * 4));
                acc = fmaf(xr[0], wv.x, fmaf(xr[1], wv.y, fmaf(xr[2], wv.z, fmaf(xr[3], wv.w, acc))));
            }
            a3[co * L3 + t] = lrelu02(acc);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * L4; e += PE_THREADS) {
            const int co = e / L4, t = e % L4;
            float acc = w.b4[co];
            for (int ci = 0; ci < 64; ++ci) {
                const float* wp = w.w4 + (co * 64 + ci) * 3;
                const float* xr = a3 + ci * L3 + t;
                acc = fmaf(xr[0], __ldg(wp), fmaf(xr[1], __ldg(wp + 1), fmaf(xr[2], __ldg(wp + 2), acc)));
            }
            a4[co * L4 + t] = acc;
        }
        __syncthreads();
        // collapsed out_net: one warp per output feature, lanes stride the flattened input
        const int nin = 32 * L4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int o = warp; o < w.n_out; o += PE_THREADS / 32) {
            float acc = 0.f;
            for (int k = lane; k < nin; k += 32) acc = fmaf(a4[k], __ldg(w.w_fc + (size_t)o * nin + k), acc);
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            if (lane == 0) out[(size_t)b * w.n_out + o] = acc + w.b_fc[o];
        }
    }
}

int g_sms = 0;
int num_sms() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

}  // namespace

int launch_cvae_mlp(const CvaeW& w, const float* x, const float* y, const float* noise, int noise_is_z, int64_t n,
                    float* out, float* mu, float* logvar, cudaStream_t s) {
    // the opt-in shared-memory size is a per-device function attribute: set it on every launch (cheap, no global state)
    if (cudaFuncSetAttribute(cvae_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CvaeSmem)) != cudaSuccess) return -1;
    const int grid = (int)std::min<int64_t>((n + CV_TILE - 1) / CV_TILE, (int64_t)num_sms() * 2);
    cvae_mlp_kernel<<<grid, CV_TILE, sizeof(CvaeSmem), s>>>(w, x, y, noise, noise_is_z, n, out, mu, logvar);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_cvae3_sample(const Cvae3W& w, const float* y, const float* z, int n, float* out, cudaStream_t s) {
    const int smem = (60 * 512 + 32 * 512) * 4;
    if (cudaFuncSetAttribute(cvae3_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    const int grid = std::min(n, num_sms());
    cvae3_sample_kernel<<<grid, E1_THREADS, smem, s>>>(w, y, z, n, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

size_t pose_encoder_smem(const PoseEncW& w) {
    const int L = w.L, PP = w.P | 1, L1 = L - 2, L2 = L - 4, L3 = (L2 - 4) / 2 + 1, L4 = L3 - 2;
    return sizeof(float) * ((size_t)L * PP + 32 * L1 + 64 * L2 + 64 * L3 + 32 * L4);
}

int launch_pose_encoder(const PoseEncW& w, const float* poses, int B, float* out, cudaStream_t s) {
    const size_t smem = pose_encoder_smem(w);
    if (smem > 220 * 1024) return -1;
    if (cudaFuncSetAttribute(pose_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const int grid = std::min(B, num_sms() * 2);
    pose_encoder_kernel<<<grid, PE_THREADS, smem, s>>>(w, poses, B, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
