// (f)4 — beat-alignment metric of the evaluation loop (test_emotion_gesture_diversity_iterative.py:243-248), pose half:
// model/Beat_score_v2.py `alignment.load_pose` (:79-127: frame differences of eight 6-coordinate joint groups, their
// L2 norms, strict local minima against `order` neighbours = scipy.signal.argrelextrema(np.less, mode='clip')),
// `motion_frames2time` (:177-180), `GAHR` (:182-196) and `calculate_align` (:198-214, the average of 3 x 8 scores).
// The audio onsets (`load_audio`, :58-77: librosa onset detection, un-vendored) are an input.
//
// One CTA per clip; everything lives in shared memory (a clip is <= 64 frames x 48 coordinates).  Speeds are
// accumulated in float32 in the reference's order with correctly rounded operations (no FMA contraction), so the
// strict comparisons — and therefore the beat indices — are bit-identical to numpy's; the scores are float64.
#include "egx_common.cuh"

namespace egx {

namespace {

constexpr int kBeatMaxFrames = 64;
constexpr int kBeatThreads = 64;
// first of the group's six columns in concat(pose[:, 18:42], pose[:, 150:174]), in load_pose's RETURN order:
// right arm, shoulder, fore-arm, wrist, left arm, shoulder, fore-arm, wrist (:101-127)
__constant__ int kGroupCol0[8] = {6, 0, 12, 18, 30, 24, 36, 42};

__global__ void __launch_bounds__(kBeatThreads)
beat_align_kernel(const float* __restrict__ poses, int F, int P, int lo, int hi, int order, double inv_2sigma2,
                  double pose_fps, const double* __restrict__ onset_t, const int* __restrict__ onset_off,
                  double* __restrict__ scores, unsigned char* __restrict__ beat_mask) {
    __shared__ float speed[8][kBeatMaxFrames];
    __shared__ short beat[8][kBeatMaxFrames];     // compacted beat indices (relative to the group's window)
    __shared__ int n_beat[8];
    __shared__ double part[24];

    const int b = blockIdx.x, t = threadIdx.x;
    const float* pose = poses + (size_t)b * F * P;
    const int T = F - 1;                          // velocity samples
    if (t < T) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int c0 = kGroupCol0[g], col = c0 < 24 ? 18 + c0 : 150 + (c0 - 24);
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float d = __fsub_rn(__ldg(pose + (size_t)(t + 1) * P + col + k), __ldg(pose + (size_t)t * P + col + k));
                const float sq = __fmul_rn(d, d);
                acc = k == 0 ? sq : __fadd_rn(acc, sq);
            }
            speed[g][t] = __fsqrt_rn(acc);
        }
    }
    __syncthreads();
    // windows: the right-side groups (0..3) are searched in [lo, hi), the left-side ones in the whole clip (:114-125)
    const int w_lo = min(max(lo, 0), T), w_hi = min(max(hi, w_lo), T);
    if (t < 8) {
        const int g = t, s0 = g < 4 ? w_lo : 0, n = g < 4 ? w_hi - w_lo : T;
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const float x = speed[g][s0 + i];
            bool ok = true;
            for (int s = 1; s <= order && ok; ++s)
                ok = x < speed[g][s0 + min(i + s, n - 1)] && x < speed[g][s0 + max(i - s, 0)];
            if (ok) beat[g][cnt++] = (short)i;
            if (beat_mask) beat_mask[((size_t)b * 8 + g) * F + i] = ok ? 1 : 0;
        }
        if (beat_mask)
            for (int i = max(n, 0); i < F; ++i) beat_mask[((size_t)b * 8 + g) * F + i] = 0;
        n_beat[g] = cnt;
    }
    __syncthreads();
    // GAHR for the 3 onset lists x 8 groups: mean over audio beats of exp(-(nearest pose beat distance)^2 / (2 sigma^2))
    if (t < 24) {
        const int l = t / 8, g = t % 8;
        const int o0 = onset_off[b * 3 + l], o1 = onset_off[b * 3 + l + 1];
        double total = 0.0;
        for (int j = o0; j < o1; ++j) {
            const double tb = onset_t[j];
            double dmin = INFINITY;
            for (int i = 0; i < n_beat[g]; ++i) dmin = fmin(dmin, fabs((double)beat[g][i] / pose_fps - tb));
            total += exp(-(dmin * dmin) * inv_2sigma2);     // no pose beat: exp(-inf) = 0, as in the reference
        }
        part[t] = total / (double)(o1 - o0);              // no audio beat: 0/0 = NaN (the reference raises ZeroDivisionError)
    }
    __syncthreads();
    if (t == 0) {
        double avg = 0.0;
        for (int i = 0; i < 24; ++i) avg += part[i];      // the reference's summation order
        scores[b] = avg / 24.0;
    }
}

}  // namespace

int launch_beat_align(const float* poses, int B, int F, int P, int lo, int hi, int order, double sigma, double pose_fps,
                      const double* onset_t, const int* onset_off, double* scores, unsigned char* beat_mask,
                      cudaStream_t s) {
    if (F < 2 || F > kBeatMaxFrames || P < 174) return -1;
    beat_align_kernel<<<B, kBeatThreads, 0, s>>>(poses, F, P, lo, hi, order, 1.0 / (2.0 * sigma * sigma), pose_fps,
                                                 onset_t, onset_off, scores, beat_mask);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace egx
