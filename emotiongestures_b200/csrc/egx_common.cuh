// Shared declarations of libegx (internal).  The public C ABI is include/egx.h.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/egx.h"

namespace egx {

// ---------------------------------------------------------------------------------------------
// Host-side containers
// ---------------------------------------------------------------------------------------------
struct HostTensor {
    std::vector<float> v;
    std::vector<int64_t> shape;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// One 3x3 (or 1x1) convolution with its folded eval-mode epilogue:
//   y = (relu_first ? max(acc + bias, 0) : acc + bias) * scale + shift
struct ConvW {
    int cin = 0, cout = 0, ks = 3, stride = 1;
    int relu_first = 0;
    float* w32 = nullptr;     // [cout][ks*ks][cin] fp32 (K-major: k = tap*cin + ci)
    __half* w16 = nullptr;    // same, fp16
    float* bias = nullptr;    // [cout] or null
    float* scale = nullptr;   // [cout]
    float* shift = nullptr;   // [cout]
};

struct LinearW {
    int in = 0, out = 0;
    float* w = nullptr;       // [out][in] fp32 (nn.Linear layout == K-major B operand)
    float* b = nullptr;       // [out] or null
    __half* w16 = nullptr;    // [out][ldw] fp16, rows padded to a multiple of 8 elements (TMA pitch)
    int ldw = 0;
};

struct LNW { float* g = nullptr; float* b = nullptr; };

struct SEW { int c = 0, r = 0; float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr; };

struct BlockW {
    ConvW conv1, conv2, down;
    bool has_down = false;
    SEW se;
};

struct MHAW { LinearW q, k, v, kv, qkv, fc; LNW ln; };
struct FFNW {
    LinearW w1, w2; LNW ln;
    std::vector<float> h_b1, h_b2, h_g, h_b;      // host copies: the fused kernel takes them as kernel parameters
};

// (f)1  Full_model/Models_memory.py Prior_MemoryEncoder (the prior encoder of the checkpointed generator)
struct MemPriorW {
    bool on = false;
    int chunk = 0, n_pred = 0;
    LinearW enc;                      // [2P][chunk*P]: spatial_chunk_encoder | temporal_chunk_encoder, each collapsed
    float *tm_w = nullptr, *tm_b = nullptr;   // temporal_memory_encoder collapsed: [chunk][chunk*P], [chunk]
};

struct Weights {
    ConvW stem;                       // 1->32, bias, relu_first, bn
    std::vector<BlockW> blocks;       // 13 SEBasicBlocks
    ConvW final_conv;                 // 128->F, bias, bn (no relu)
    LinearW a_fc1, a_fc2;
    // prior encoder
    float *p_c1w = nullptr, *p_c1b = nullptr, *p_s1 = nullptr, *p_t1 = nullptr;
    float *p_c2w = nullptr, *p_c2b = nullptr, *p_s2 = nullptr, *p_t2 = nullptr;
    float* p_c2wt = nullptr;          // conv2 weights as the kernel stages them: [cin][tap][cout padded to 4], zero padded
    LinearW p_fc1, p_fc2;             // Prior_ConvEncoder fc1 / fc2, or Prior_MemoryEncoder post_header.0 / .2
    MemPriorW mem;
    LinearW emo0, emo2, sem0, sem2, fus0, fus2;
    LinearW hdr[4];
    LinearW post[4];
    // tensor-core arm: Linear -> Dropout -> Linear chains (no activation between them; Dropout is the identity in
    // eval) collapsed into one affine map, composed in float64 at weight-packing time
    LinearW a_fc, p_fc, emo, sem, post_all;
    float* pos_table = nullptr;       // [n_position][d]
    std::vector<MHAW> enc_attn, dec_attn;
    std::vector<FFNW> enc_ffn, dec_ffn;
};

// ---- small networks either side of the generator (k_aux.cu); all pointers are device memory ----
// C4  Full_model/BEAT_CVAE.py MLP_Reconstruct, Linear chains collapsed (eval: Dropout = identity)
struct CvaeW {
    float* w_x = nullptr;   // [90][64]  x -> (mu | logvar), k-major
    float* b_x = nullptr;   // [64]
    float* w_y = nullptr;   // [90][32]  y -> Posterior_Y_embedding
    float* b_y = nullptr;   // [32]
    float* w_d = nullptr;   // [64][92]  [z ; post_y] -> output (fusion_z_posterior + Decoder), 90 padded to 92
    float* b_d = nullptr;   // [92]
    bool ready = false;
};
// E1  CAVE/BEAT_CVAE.py MLP_Reconstruct_v3 (sampler half)
struct Cvae3W {
    float *w_y = nullptr, *b_y = nullptr;     // [32][8], [32]   Posterior_Y_embedding collapsed
    float *w_f = nullptr, *b_f = nullptr;     // [512][64], [512] fusion_z_posterior collapsed
    float *t1_w = nullptr, *t1_b = nullptr, *s1 = nullptr, *h1 = nullptr;   // ConvT 4->8 (cin,cout,3) + BN after LReLU
    float *t2_w = nullptr, *t2_b = nullptr, *s2 = nullptr, *h2 = nullptr;   // ConvT 8->16
    float *c3_w = nullptr, *c3_b = nullptr, *s3 = nullptr, *h3 = nullptr;   // Conv 16->32 (cout,cin,3)
    float *c4_w = nullptr, *c4_b = nullptr, *s4 = nullptr, *h4 = nullptr;   // Conv 32->60
    float *c5_w = nullptr, *c5_b = nullptr;                                 // Conv 60->60
    bool ready = false;
};
// D1  PoseEncoderConv (model/motion_ae.py:55-62, model/embedding_net.py:67-83), BN folded, out_net collapsed
struct PoseEncW {
    int L = 0, P = 0, n_out = 0;
    float *w1 = nullptr, *b1 = nullptr;       // (32,P,3)
    float *w2 = nullptr, *b2 = nullptr;       // (64,32,3)
    float *w3 = nullptr, *b3 = nullptr;       // (64,64,4), stride 2
    float *w4 = nullptr, *b4 = nullptr;       // (32,64,3)
    float *w_fc = nullptr, *b_fc = nullptr;   // [n_out][32*L4]
    bool ready = false;
};
// D1  per-frame feature MLP (model/FGD.py:26-41 Encoder, three Linears collapsed into one)
struct RowMlpW {
    LinearW lin;                              // in -> out, fp16 copy for the tensor-core GEMM
    bool ready = false;
};

// C3  model/audio_emotion_classifer.py EmotionNet: four-stage SE-ResNet + six Linears (fp16 tensor-core arm only)
struct EmotionNetW {
    ConvW stem;
    std::vector<BlockW> blocks;     // 3 + 4 + 6 + 3
    LinearW fc[6];                  // 65536->4096->2048->512->128->64->8; fc[0] columns permuted to NHWC order
    int flat = 0;
    bool ready = false;
};

// (f)2  skeleton_classifer/Models.py Transformer: the Emotion-ACC classifier run on every generated batch
// (test_emotion_gesture_diversity_iterative.py:158,217): Linear chain -> + sinusoid rows -> n_layers EncoderLayers
// -> flatten -> Linear+ReLU x4 -> Linear.  fp16 tensor-core arm only, d_k = d_v = 64.
struct SkeletonW {
    int T = 0, P = 0, d = 0, d_inner = 0, n_layers = 0, n_head = 0, n_class = 0;
    LinearW prior;                    // prior_seq_encoder.fc1 -> Dropout -> fc2 collapsed
    float* pos_table = nullptr;       // [T][d]
    std::vector<MHAW> attn;
    std::vector<FFNW> ffn;
    LinearW post[5];
    bool ready = false;
};

// Log-mel tables (built in float64 on the host, stored as float32)
struct LogmelTables {
    float* window = nullptr;      // [1024] periodic Hann
    float2* tw512 = nullptr;      // [2][8][64] per-pass twiddles of the radix-8 FFT, W512^((j mod Ns) * t * 64 / Ns)
    float2* tw1024 = nullptr;     // [513] exp(-2*pi*i*k/1024)
    int* mel_start = nullptr;     // [128] first bin with non-zero weight
    int* mel_ptr = nullptr;       // [129] number of taps of every mel filter
    float* mel_w = nullptr;       // [taps][128] non-zero weights, tap-major (zero padded)
};

}  // namespace egx

struct egx_handle {
    egx_cfg cfg{};
    int device = 0;
    std::string err;
    std::map<std::string, egx::HostTensor> staged;   // weights as received
    std::vector<void*> owned;                          // device allocations to free (log-mel tables + generator)
    std::map<std::string, std::vector<void*>> aux_owned;   // per aux-model family
    std::vector<void*>* cur_bucket = nullptr;          // where upload() records allocations (null: `owned`)
    egx::Weights w;
    egx::LogmelTables lm;
    bool finalized = false;                            // generator weights packed
    bool upload_failed = false;                        // a weight upload failed (its list entry is null)
    egx::CvaeW cvae;
    egx::Cvae3W cvae3;
    egx::PoseEncW motion_ae, pose_enc;
    egx::RowMlpW fgd_mlp;
    egx::EmotionNetW emo;
    egx::SkeletonW skel;
    int sms = 0;                                       // multiprocessors of the device (read-only property)
    double* fgd_scratch = nullptr;                     // partial Gram blocks of egx_fgd_accumulate, grown on demand
    size_t fgd_scratch_n = 0;
    int64_t launches = 0;
    // per-launch CUDA-event profiling (egx_profile_enable / egx_profile_read)
    bool profiling = false;
    int stage = 0;
    size_t prof_used = 0;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_stage;
    // trunk geometry
    int H[4] = {0, 0, 0, 0}, W[4] = {0, 0, 0, 0};      // [0]=input/layer1, [1]=layer2, [2]=layer3
};

namespace egx {

#define EGX_CHECK_CUDA(h, expr)                                                            \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

#define EGX_FAIL(h, msg)                                                                   \
    do {                                                                                   \
        (h)->err = (msg);                                                                  \
        return 1;                                                                          \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return int((a + b - 1) / b); }

// Variant / attribution switches (EGX_* environment variables) exist only in builds made with -DEGX_ATTRIBUTION
// (python -m emotiongestures_b200.build --attribution): the shipped library never reads the environment, so a stray
// variable cannot change which kernels run or — for the EGX_*_DEBUG bits, which skip work — corrupt results.
inline int env_switch(const char* name, int dflt) {
#ifdef EGX_ATTRIBUTION
    if (const char* e = getenv(name)) return atoi(e);
#else
    (void)name;
#endif
    return dflt;
}

// ---------------------------------------------------------------------------------------------
// Kernel launchers (each returns the number of kernels it enqueued, or <0 on launch error)
// ---------------------------------------------------------------------------------------------
int launch_logmel(const LogmelTables& t, const float* audio, int B, int N, int n_cols, int mode,
                  int preemph, float* out, cudaStream_t s, bool force_global_tile = false);

int launch_pcm16_to_f32(const int16_t* in, int64_t n, float* out, cudaStream_t s);

// F5: ragged clips (back to back in `samples`, clip b = [offsets[b], offsets[b+1])) -> (B, N), cropped or symmetric-padded
int launch_fixed_length(const float* samples, const int64_t* offsets, int B, int N, float* out, cudaStream_t s);

template <class T>
int launch_stem(const ConvW& c, const float* spec, int B, int H, int W, T* out, cudaStream_t s);

// Direct (CUDA-core) implicit-GEMM convolution over NHWC activations.
//   out_nchw_f32 != nullptr: write fp32 NCHW there instead of NHWC T (final conv).
template <class T>
int launch_conv_direct(const ConvW& c, const T* in, int B, int Hin, int Win, T* out,
                       float* out_nchw_f32, cudaStream_t s);

// sums: [B][se_partials(HW)][C] partial sums in a fixed order (no atomics: deterministic)
constexpr int kSePixPerBlock = 512;
inline int se_partials(int HW) { return (HW + kSePixPerBlock - 1) / kSePixPerBlock; }
template <class T>
int launch_se_reduce(const T* y, int B, int HW, int C, float* sums, cudaStream_t s);

// out = relu(gate(sums) * y + res), gate = sigmoid(W2 relu(W1 mean + b1) + b2)
template <class T>
int launch_se_apply(const SEW& se, const T* y, const T* res, const float* sums, int n_part, int B,
                    int HW, T* out, cudaStream_t s);

// SE gate ahead of conv2 (tensor-core arm): window means of y1 (B,H,W,C) as fp16 rows [B][9*C], then the folded
// per-clip epilogue of conv2, gate [B][2][C] = (g*scale2 | g*(shift2 + bias2*scale2)); see k_trunk.cu K4c/K4d
int launch_se_window(const __half* y, int B, int H, int W, int C, const float* part, int n_part, __half* win,
                     cudaStream_t s);
int launch_se_gate(const SEW& se, const ConvW& conv2, const float* mean_raw, int B, float* gate, cudaStream_t s);

struct GemmEpi {
    const float* bias = nullptr;      // [N]
    int relu = 0;
    const float* addend = nullptr;    // added after bias/relu; row index = row % addend_rows
    int addend_rows = 0;              // 0: row index = row
    int addend_ld = 0;
    // tensor-core arm only: LayerNorm (eps 1e-6) of the finished row in the epilogue when N == 256 (launch_gemm_tc_ln_ok)
    const float* ln_g = nullptr;
    const float* ln_b = nullptr;
};
// C[M][N] (ldc) = A[M][K] (lda) * W[N][K]^T  (+ epilogue), fp32 CUDA cores
int launch_gemm_f32(const float* A, int lda, const float* Wt, int M, int N, int K, float* C,
                    int ldc, const GemmEpi& e, cudaStream_t s);

// out[r] = LayerNorm(x[r]) * g + b   (x already holds the residual sum), eps 1e-6
int launch_layernorm(const float* x, const LNW& ln, int rows, int d, float* out, __half* out16,
                     cudaStream_t s);

// softmax((q/sqrt(dk)) k^T) v per (clip, head).  q rows: (B*Lq, ldq), k/v rows: (B*Lk, ld*)
template <class T>
int launch_attention(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv,
                     int B, int Lq, int Lk, int n_head, int dk, int dv, T* out, int ldo,
                     cudaStream_t s);

template <class T>
int launch_prior_conv(const Weights& w, const float* prior, int B, int p, int F, int P,
                      T* out, int ldo, cudaStream_t s);

int launch_add(const float* a, const float* b, float* out, int64_t n, cudaStream_t s);
int launch_add_f16(const float* a, const float* b, __half* out, int64_t n, cudaStream_t s);

template <class T>
int launch_nhwc_to_nchw_f32(const T* in, int B, int HW, int C, float* out, cudaStream_t s);

// ---- tensor-core arm (tcgen05) ----
int gemm_tc_init_device();
int launch_cvt_pad_f16(const float* in, int64_t rows, int cols, int ld_in, __half* out, int ld_out, cudaStream_t s);
// C = A[M][K] (fp16, pitch lda) x W[N][K]^T (fp16, pitch ldw); outputs fp32 and/or fp16 (either may be null)
int launch_gemm_tc(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const GemmEpi& e,
                   float* out32, int ld32, __half* out16, int ld16, cudaStream_t s);

// fused position-wise feed-forward block (k_ffn_tc.cu): LayerNorm(x + W2 relu(W1 x + b1) + b2), d_model = 256
int ffn_tc_init_device();
bool ffn_tc_supported(int d_model, int d_inner);
int launch_ffn_tc(const __half* x16, const float* resid, const __half* w1, int ldw1, const float* b1, const __half* w2, int ldw2,
                  const float* b2, const float* ln_g, const float* ln_b, int M, int d_inner, float* out32, __half* out16,
                  cudaStream_t s);

int conv_tc_init_device();
// in: NHWC fp16 (B,Hin,Win,cin); out: NHWC fp16 or (B,cout,Ho*Wo) fp16 when nchw != 0
// se_part (optional): [B][conv_tc_tiles_per_clip(Ho,Wo)][cout] per-tile channel sums (fixed order)
// gate / res (optional, together): the epilogue becomes out = relu(acc * gate[b][0][c] + gate[b][1][c] + res), the
// SE-scaled residual sum of Full_model/ResNetBlocks.py:28-36; res is NHWC fp16 with the output's geometry
int launch_conv_tc(const ConvW& c, const __half* in, int B, int Hin, int Win, __half* out, int nchw,
                   float* se_part, cudaStream_t s, const float* gate = nullptr, const __half* res = nullptr);
int conv_tc_tiles_per_clip(int cin, int cout, int Ho, int Wo);

int attn_tc_init_device();
// softmax((q/8) k^T) v per (clip, head) on tcgen05; d_k = d_v = 64.  Head h of q at columns q_col0 + 64 h of
// rows (B*L, ldq); k / v at k_col0 / v_col0 + 64 h of rows (B*L, ldkv); out (B*L, ldo), head h at 64 h.
int launch_attention_tc(const __half* q, int ldq, int q_col0, const __half* kv, int ldkv, int k_col0, int v_col0,
                        int B, int L, int n_head, __half* out, int ldo, cudaStream_t s);

// two-phase, deterministic: partial Gram blocks per row slice into `scratch` (fgd_scratch_doubles), then a fixed-order sum
size_t fgd_scratch_doubles(int64_t n, int D, int sms, int* n_split_out);
int launch_fgd_accumulate(const float* feats, int64_t n, int D, const double* shift, double* acc, double* scratch,
                          int n_split, cudaStream_t s);

// ---- k_memory.cu (Prior_MemoryEncoder between pred_conv and post_header) ----
int launch_mem_spatial(float* pred, int B, int n_pred, int P, int C, const float* enc, const float* tm_w, const float* tm_b,
                       float* pred_enc, cudaStream_t s);
int launch_mem_batch_outer(const float* enc, const float* pred_enc, int B, int P, int C, float* S, cudaStream_t s);
template <class T>
int launch_mem_temporal(const float* prior, const float* pred, int B, int p, int n_pred, int P, int C, const float* enc,
                        const float* S, T* out, int ldo, cudaStream_t s);

// ---- k_beat.cu ((f)4: pose beats + GAHR of model/Beat_score_v2.py) ----
int launch_beat_align(const float* poses, int B, int F, int P, int lo, int hi, int order, double sigma, double pose_fps,
                      const double* onset_t, const int* onset_off, double* scores, unsigned char* beat_mask,
                      cudaStream_t s);

// ---- k_aux.cu ----
int launch_cvae_mlp(const CvaeW& w, const float* x, const float* y, const float* noise, int noise_is_z, int64_t n,
                    float* out, float* mu, float* logvar, cudaStream_t s);
int launch_cvae3_sample(const Cvae3W& w, const float* y, const float* z, int n, float* out, cudaStream_t s);
int launch_pose_encoder(const PoseEncW& w, const float* poses, int B, float* out, cudaStream_t s);

}  // namespace egx
