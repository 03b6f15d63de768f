/*
 * libegx — C ABI of the B200-native EmotionGesture generator-inference path.
 *
 * The reference (XingqunQi-lab/EmotionGestures) is pure PyTorch and defines no
 * FFI of its own; its boundary for this path is one nn.Module call,
 *   pred_pose, _, _, emotion_prediction, _ = generator(in_spec, text, pre_pose, sampled)
 *   (test_emotion_gesture_diversity_iterative.py:205; Full_model/Models.py:389-427;
 *    4th argument Full_model/Models_memory.py:521).
 * The entry points below are what a ctypes binding placed behind that call needs
 * (INTEGRATION.md shows the stub).  Conventions:
 *   - extern "C", int status returns (0 = ok, non-zero = error; text via egx_last_error)
 *   - every data pointer is a DEVICE pointer owned by the caller (e.g. the torch
 *     allocator); the library never frees or keeps them past the call, except weights,
 *     which egx_set_weight copies
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*); the
 *     library does not synchronise it (weight loading is the one exception)
 *   - no mutable global state: one handle per device / per replica thread.  The only process-wide
 *     values are read-only device properties (SM count, opt-in shared-memory sizes) cached by
 *     egx_create; the library never reads the environment (the EGX_* variant / attribution
 *     switches exist only in -DEGX_ATTRIBUTION builds)
 */
#ifndef EGX_H
#define EGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGX_VERSION 1

#if defined(__GNUC__)
#define EGX_API __attribute__((visibility("default")))
#else
#define EGX_API
#endif

typedef struct egx_handle egx_handle;

/* Geometry of the generator: literals of the constructor call
 * (test_emotion_gesture_diversity_iterative.py:135; Full_model/Models.py:298-301). */
typedef struct egx_cfg {
    int32_t frames;        /* F: 34 (TED) / 60 (BEAT)                         */
    int32_t prior_frames;  /* p: 4 / 10                                       */
    int32_t pose_dim;      /* P: 126 / 282                                    */
    int32_t d_model;       /* d: 256 / 512                                    */
    int32_t d_inner;       /* FFN hidden: 1024 / 2048                         */
    int32_t n_layers;      /* 3                                               */
    int32_t n_head;        /* 8                                               */
    int32_t d_k;           /* 64                                              */
    int32_t d_v;           /* 64                                              */
    int32_t n_mels;        /* 128                                             */
    int32_t spec_w;        /* spectrogram columns W: 70 / 124                 */
    int32_t n_position;    /* rows of encoder.position_enc.pos_table: 60      */
    int32_t precision;     /* EGX_PREC_*                                      */
} egx_cfg;

enum { EGX_PREC_FP32 = 0,  /* CUDA-core fp32 everywhere (bit-faithful debugging arm)      */
       EGX_PREC_TC   = 1   /* tcgen05 kind::f16: fp16 operands, fp32 accumulation in TMEM, in the trunk
                              convolutions, the Linear chain and the attention alike; fp32 residual
                              stream, LayerNorm and outputs                                          */ };

enum { EGX_DTYPE_F32 = 0, EGX_DTYPE_I64 = 1 };

enum { EGX_LOGMEL_DB = 0,      /* utils/data_utils.py:36-37 (power_to_db, ref=max)          */
       EGX_LOGMEL_LOG_IN = 1,  /* model/ResNetSE34V2.py:96-98 (log(x+1e-6), InstanceNorm1d) */
       /* OR-ed into EGX_LOGMEL_DB: round the result to fp16 and back, the astype('float16') storage cast of
        * utils/data_utils.py:38 that every feature the reference's checkpoints saw went through */
       EGX_LOGMEL_FP16_STORAGE = 0x100 };

EGX_API int  egx_version(void);

/* Replaces: Transformer.__init__ (Full_model/Models.py:298-383) as far as device state goes. */
EGX_API int  egx_create(const egx_cfg* cfg, int device, egx_handle** out);
EGX_API void egx_destroy(egx_handle* h);
EGX_API const char* egx_last_error(const egx_handle* h);

/* Replaces: nn.Module.load_state_dict for the keys of SURVEY.md §8(b).  `key` is the
 * state_dict key, `data` a device pointer to a contiguous tensor.  Unknown keys (text
 * encoder, never-run modules) are accepted and ignored.  egx_finalize_weights folds
 * eval-mode BatchNorm into scale/shift, repacks conv weights K-major and uploads. */
EGX_API int  egx_set_weight(egx_handle* h, const char* key, const void* data, const int64_t* shape,
                    int ndim, int dtype);
EGX_API int  egx_finalize_weights(egx_handle* h);

/* Replaces: PreEmphasis.forward (model/utils.py:33-38) + librosa.feature.melspectrogram /
 * power_to_db (utils/data_utils.py:35-39) or the log+InstanceNorm recipe
 * (model/ResNetSE34V2.py:96-98).  audio (B,N) f32 -> out (B,128,n_cols) f32. */
EGX_API int  egx_logmel(egx_handle* h, const float* audio, int n_clips, int n_samples, int n_cols,
                int mode, int preemph, float* out, void* stream);

/* 16-bit PCM samples -> float32 in [-1, 1) (x / 32768, exact: the scaling every wav decoder applies).  No counterpart
 * in the reference, whose LMDB holds clips that were decoded offline (data_loader/data_preprocessor_expressive.py:73,
 * 'audio_raw'); an extension for serving raw speech, so that it crosses PCIe at 2 bytes per sample and is widened on the
 * device.  pcm, out: device pointers, n_samples elements. */
EGX_API int  egx_audio_pcm16_to_f32(egx_handle* h, const int16_t* pcm, int64_t n_samples, float* out, void* stream);

/* Replaces: make_audio_fixed_length (utils/data_utils.py:69-75) for a ragged batch.  Clip b is
 * samples[offsets[b] .. offsets[b+1]) (f32 / int64, both device memory, n_clips + 1 offsets); every clip is cropped
 * to n_out samples or extended at its end the way np.pad(mode='symmetric') does -> out (n_clips, n_out) f32, the
 * `audio` argument of egx_logmel. */
EGX_API int  egx_audio_fixed_length(egx_handle* h, const float* samples, const int64_t* offsets, int n_clips,
                            int n_out, float* out, void* stream);

/* Scratch the forward needs for a batch of n_clips (bytes). */
EGX_API size_t egx_workspace_bytes(const egx_handle* h, int n_clips);

/* Replaces: Transformer.forward (Full_model/Models.py:389-427) minus the text encoder.
 * spec (B,128,W) f32; prior (B,p,P) f32; sampled_emotion (B,F,d) f32 or NULL
 * -> poses (B,F,P), emo_feat (B,F,d), sem_feat (B,F,d), emo_logits (B,8), all f32. */
EGX_API int  egx_generator_forward(egx_handle* h, const float* spec, const float* prior,
                           const float* sampled_emotion, int n_clips, float* poses,
                           float* emo_feat, float* sem_feat, float* emo_logits,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Parity probes (tests only).  egx_get_tap copies an intermediate of the LAST forward out of
 * `workspace` in the reference's layout (f32; NCHW for maps).  Names: layer3,
 * spectrum_feature, prior_feature, enc_output, dec_output.  egx_debug_trunk re-runs the
 * trunk on `spec` up to `stage` (0 = stem, 1..3 = layer1..3) and writes that map as f32
 * NCHW.  Both return the element count through *n_out. */
EGX_API int  egx_get_tap(egx_handle* h, const char* name, const void* workspace, int n_clips,
                 float* out, size_t out_capacity, size_t* n_out, void* stream);
EGX_API int  egx_debug_trunk(egx_handle* h, const float* spec, int n_clips, int stage, float* out,
                     size_t out_capacity, size_t* n_out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Replaces: np.mean / np.cov inputs (test_emotion_gesture_diversity_iterative.py:251-254).
 * Adds the sufficient statistics of feats (n,D) f32 into acc = [n | sum (D) | gram (D*D)]
 * (float64, device).  `shift` (D, float64, device, may be NULL) is subtracted from every
 * row first (provisional mean). */
EGX_API int  egx_fgd_accumulate(egx_handle* h, const float* feats, int64_t n_rows, int dim,
                        const double* shift, double* acc, void* stream);

/* ---- small networks either side of the generator ------------------------------
 * Their weights arrive through egx_set_weight under a family prefix followed by the module's own state_dict key:
 *   "cvae."       Full_model/BEAT_CVAE.py   MLP_Reconstruct
 *   "cvae3."      CAVE/BEAT_CVAE.py         MLP_Reconstruct_v3
 *   "motion_ae."  model/motion_ae.py        MotionAE            (encoder half is used)
 *   "pose_enc."   model/embedding_net.py    PoseEncoderConv
 *   "fgd_mlp."    model/FGD.py              MLP_Reconstruct     (Encoder half is used)
 * egx_finalize_weights packs whichever families were staged.  The reference draws its Gaussian noise inside the
 * module (torch.randn / randn_like); here the caller passes the draw in, so results are reproducible and do not
 * depend on how clips are sharded over GPUs. */

/* Replaces: MLP_Reconstruct.forward (Full_model/BEAT_CVAE.py:98-114).  x, y (n,90); eps (n,32) is the randn_like
 * draw of reparameterize (:91-94) -> out (n,90), mu (n,32), logvar (n,32). */
EGX_API int  egx_cvae_forward(egx_handle* h, const float* x, const float* y, const float* eps, int64_t n,
                      float* out, float* mu, float* logvar, void* stream);
/* Replaces: MLP_Reconstruct.sample (Full_model/BEAT_CVAE.py:117-136).  y (n,90); z (n,32) is the torch.randn
 * draw of :130 -> out (n,90). */
EGX_API int  egx_cvae_sample(egx_handle* h, const float* y, const float* z, int64_t n, float* out, void* stream);
/* Replaces: MLP_Reconstruct_v3.sample (CAVE/BEAT_CVAE.py:427-447).  y (n,8) one-hot emotion; z (n,32) the
 * torch.randn draw of :441 -> out (n,60,512), the sampled_emotion_feature of the BEAT generator. */
EGX_API int  egx_cvae3_sample(egx_handle* h, const float* y, const float* z, int n, float* out, void* stream);

/* Replaces: the FGD feature nets of the evaluation loop (test_emotion_gesture_diversity_iterative.py:226-229).
 * kind EGX_POSE_MOTION_AE:     MotionAE.encoder (model/motion_ae.py:55-83,125-130), poses (n,34,P) -> z (n,latent)
 * kind EGX_POSE_EMBEDDING_NET: PoseEncoderConv.forward (model/embedding_net.py:67-83), poses (n,60,P) -> mu (n,32) */
enum { EGX_POSE_MOTION_AE = 0, EGX_POSE_EMBEDDING_NET = 1 };
EGX_API int  egx_pose_features(egx_handle* h, int kind, const float* poses, int n_clips, int n_frames,
                       int pose_dim, float* out, void* stream);
EGX_API int  egx_pose_feature_dim(const egx_handle* h, int kind);
/* Replaces: MLP_Reconstruct.Encoder of model/FGD.py:30-41,66-82 applied per frame: rows (n,282) -> latent (n,512).
 * workspace holds the fp16 copy of the rows (egx_row_features_workspace bytes). */
EGX_API size_t egx_row_features_workspace(const egx_handle* h, int64_t n_rows);
EGX_API int  egx_row_features(egx_handle* h, const float* rows, int64_t n_rows, int dim, float* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: EmotionNet.forward (model/audio_emotion_classifer.py:38-49; trunk model/emotion_ResNetSE34V2.py:57-72),
 * weights under the "emotion_net." prefix.  spec (n,128,W) f32 log-mel, W even with 256*(128/8)*(W/8 rounded up) = 65536
 * for the checkpointed net (W = 124) -> logits (n,8) f32 (the reference also leaves the softmax off). */
EGX_API size_t egx_emotion_net_workspace(const egx_handle* h, int n_clips, int n_mels, int n_cols);
EGX_API int  egx_emotion_net_forward(egx_handle* h, const float* spec, int n_clips, int n_mels, int n_cols,
                             float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: skeleton_classifer/Models.py Transformer.forward (:264-283), the Emotion-ACC classifier the evaluation
 * loop runs on every generated batch (test_emotion_gesture_diversity_iterative.py:158,217); weights under the
 * "skel." prefix, geometry read from their shapes (d_k = d_v = 64 as the evaluation script builds it).
 * poses (n, n_frames, pose_dim) f32 with n_frames == n_position -> logits (n, class_dim) f32 and, when mid_feature is
 * not NULL, the encoder output (n, n_frames, d_model) f32 (the second value the reference returns). */
EGX_API size_t egx_skeleton_workspace(const egx_handle* h, int n_clips);
EGX_API int  egx_skeleton_forward(egx_handle* h, const float* poses, int n_clips, int n_frames, int pose_dim,
                          float* logits, float* mid_feature, void* workspace, size_t workspace_bytes, void* stream);
EGX_API int  egx_skeleton_dims(const egx_handle* h, int* n_frames, int* pose_dim, int* d_model, int* n_class);

/* Replaces: alignment.load_pose + calculate_align of model/Beat_score_v2.py (:79-127, :177-214), called per generated
 * clip at test_emotion_gesture_diversity_iterative.py:243-248.  poses (n, n_frames <= 64, pose_dim >= 174) f32.  The
 * eight joint-group speed series are searched for strict local minima (`order` neighbours each side, clipped ends):
 * the right-side groups inside frames [frame_lo, frame_hi) = [t_start*fps, t_end*fps), the left-side ones over the whole
 * clip, as the reference does.  onset_times (f64, seconds) holds, back to back, the three onset lists of every clip —
 * load_audio's (onset_raw, onset_bt, onset_bt_rms) after librosa.frames_to_time; list l of clip b is
 * onset_times[onset_offsets[3b+l] .. onset_offsets[3b+l+1]) (int32, 3n+1 entries).  scores (n) f64 = the reference's
 * avg_dis_all_b2a per clip (NaN where a list is empty: the reference raises ZeroDivisionError there).  beat_mask
 * (nullable, (n, 8, n_frames) u8) marks the beat frames per group in load_pose's return order, window-relative. */
EGX_API int  egx_beat_align(egx_handle* h, const float* poses, int n_clips, int n_frames, int pose_dim, int frame_lo,
                    int frame_hi, int order, double sigma, double pose_fps, const double* onset_times,
                    const int32_t* onset_offsets, double* scores, unsigned char* beat_mask, void* stream);

/* Parity probe of the tcgen05 Linear kernel alone: out = [relu](A W^T + bias) + addend, A (M,K), W (N,K),
 * out (M,N) f32; operands are rounded to fp16 inside.  Synchronises the stream (test-only). */
EGX_API int  egx_debug_linear_tc(egx_handle* h, const float* A, const float* W, const float* bias,
                         const float* addend, int addend_rows, int M, int N, int K, int relu,
                         float* out32, void* stream);

/* Parity probe of the Linear + residual + LayerNorm epilogue variant (N = 256): out = LayerNorm(A W^T + bias + residual)
 * * ln_g + ln_b, eps 1e-6 (Full_model/SubLayers.py:53-57,80-82); out32 (M,256) f32 and, when not NULL, out16 (M,256)
 * f16.  Synchronises the stream (test-only). */
EGX_API int  egx_debug_linear_ln_tc(egx_handle* h, const float* A, const float* W, const float* bias,
                            const float* residual, const float* ln_g, const float* ln_b, int M, int K,
                            float* out32, void* out16, void* stream);

/* Parity probe of the fused feed-forward kernel alone (Full_model/SubLayers.py:74-84, d_model = 256):
 * out = LayerNorm(x + W2 relu(W1 x + b1) + b2); x (M,256), W1 (d_inner,256), W2 (256,d_inner) f32 (rounded to fp16 inside,
 * the residual stays f32); out32 (M,256) f32, out16 (M,256) f16.  Synchronises the stream (test-only). */
EGX_API int  egx_debug_ffn_tc(egx_handle* h, const float* x, const float* w1, const float* b1, const float* w2,
                      const float* b2, const float* ln_g, const float* ln_b, int M, int d_inner, float* out32,
                      void* out16, void* stream);

/* Test hook of the front end: egx_logmel through the kernel that keeps the (128, n_cols) tile in global memory (the
 * path of spectrograms too wide for two shared-memory tiles per SM, about 140 columns) whatever the width.  Both kernels run the same arithmetic in the same
 * order: tests require bit-identical outputs. */
EGX_API int  egx_debug_logmel_global_tile(egx_handle* h, const float* audio, int n_clips, int n_samples, int n_cols,
                                  int mode, int preemph, float* out, void* stream);

/* Parity probe of the tcgen05 implicit-GEMM convolution alone.  in16: NHWC fp16 (B,H,W,cin); w16: fp16
 * [cout][ks*ks][cin]; y = (relu_first ? relu(acc+bias) : acc+bias)*scale + shift; out16: NHWC fp16, or
 * (B,cout,Ho*Wo) fp16 when nchw != 0; se_part (nullable): [B][tiles][cout] per-tile channel sums. */
EGX_API int  egx_debug_conv_tc(egx_handle* h, const void* in16, int B, int H, int W, int cin,
                       const void* w16, int cout, int ks, int stride, int relu_first,
                       const float* bias, const float* scale, const float* shift, void* out16,
                       int nchw, float* se_part, void* stream);

/* Parity probe of the tcgen05 attention kernel alone (fp16 in/out, d_k = d_v = 64): head h of q at columns
 * q_col0 + 64 h of rows (B*L, ldq); k / v at k_col0 / v_col0 + 64 h of rows (B*L, ldkv). */
EGX_API int  egx_debug_attention_tc(egx_handle* h, const void* q16, int ldq, int q_col0, const void* kv16,
                            int ldkv, int k_col0, int v_col0, int B, int L, int n_head, void* out16,
                            int ldo, void* stream);

/* Measurement hooks (bench.py): with profiling enabled (max_launches > 0) every kernel launch is
 * bracketed by a CUDA-event pair on the launching stream, tagged with its stage of SURVEY.md
 * §8(d) (1 front-end, 2 stem, 3 trunk convolutions of layer 1 — on the tensor-core arm 9 / 10 / 11 are those of
 * layers 2 / 3 (+ final conv) / 4, on the fp32 arm 3 covers all —, 4 SE gate, 5 projection GEMMs,
 * 6 encoder+decoder, 8 FGD statistics, 0 other).  egx_profile_read waits for the recorded
 * events, sums elapsed milliseconds and launch counts per stage and resets the recording. */
EGX_API int  egx_profile_enable(egx_handle* h, int max_launches);
EGX_API int  egx_profile_read(egx_handle* h, double* ms_per_stage, int64_t* launches_per_stage,
                      int n_stages);

/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
EGX_API int64_t egx_launch_count(const egx_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* EGX_H */
