// libegx C ABI (include/egx.h): handle, weight hand-off, workspace planning and the forward
// schedule of the generator (Full_model/Models.py:389-427).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "egx_common.cuh"

using namespace egx;

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr float kBnEps = 1e-5f;    // torch.nn.BatchNorm default; the reference never overrides it

// ---------------------------------------------------------------------------------------------
// device upload helpers
// ---------------------------------------------------------------------------------------------
template <class T>
T* upload(egx_handle* h, const std::vector<T>& v) {
    // a failed allocation / copy is recorded as a null entry of the owning list, which the packers check before they
    // declare a family ready (egx_finalize_weights, BucketScope::ok): no kernel ever sees a null weight pointer
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)) != cudaSuccess) p = nullptr;
    if (p && !v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p);
        p = nullptr;
    }
    if (!p) { cudaGetLastError(); h->upload_failed = true; }
    (h->cur_bucket ? *h->cur_bucket : h->owned).push_back(p);
    return static_cast<T*>(p);
}

const HostTensor* find(egx_handle* h, const std::string& key) {
    auto it = h->staged.find(key);
    if (it == h->staged.end()) {
        h->err = "missing weight: " + key;
        return nullptr;
    }
    return &it->second;
}

bool need(egx_handle* h, const std::string& key, std::initializer_list<int64_t> shape, const HostTensor** out) {
    const HostTensor* t = find(h, key);
    if (!t) return false;
    if (t->shape != std::vector<int64_t>(shape)) {
        std::string got, want;
        for (auto s : t->shape) got += std::to_string(s) + ",";
        for (auto s : shape) want += std::to_string(s) + ",";
        h->err = "weight " + key + " has shape (" + got + ") expected (" + want + ")";
        return false;
    }
    *out = t;
    return true;
}

// eval-mode BatchNorm -> y = x*scale + shift
bool fold_bn(egx_handle* h, const std::string& pre, int c, std::vector<float>& scale, std::vector<float>& shift) {
    const HostTensor *g, *b, *m, *v;
    if (!need(h, pre + ".weight", {c}, &g) || !need(h, pre + ".bias", {c}, &b) ||
        !need(h, pre + ".running_mean", {c}, &m) || !need(h, pre + ".running_var", {c}, &v))
        return false;
    scale.resize(c);
    shift.resize(c);
    for (int i = 0; i < c; ++i) {
        const float s = g->v[i] / std::sqrt(v->v[i] + kBnEps);
        scale[i] = s;
        shift[i] = b->v[i] - m->v[i] * s;
    }
    return true;
}

// torch conv weight (cout, cin, kh, kw) -> [cout][kh*kw][cin]
bool make_conv(egx_handle* h, const std::string& wkey, const std::string* bias_key, const std::string& bnpre,
               int cin, int cout, int ks, int stride, int relu_first, ConvW* out) {
    const HostTensor* w;
    if (!need(h, wkey, {cout, cin, ks, ks}, &w)) return false;
    const int taps = ks * ks;
    std::vector<float> packed((size_t)cout * taps * cin);
    std::vector<__half> packed16(packed.size());
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i)
            for (int t = 0; t < taps; ++t) {
                const float x = w->v[((size_t)o * cin + i) * taps + t];
                packed[((size_t)o * taps + t) * cin + i] = x;
                packed16[((size_t)o * taps + t) * cin + i] = __float2half_rn(x);
            }
    std::vector<float> scale, shift;
    if (!fold_bn(h, bnpre, cout, scale, shift)) return false;
    out->cin = cin; out->cout = cout; out->ks = ks; out->stride = stride; out->relu_first = relu_first;
    out->w32 = upload(h, packed);
    out->w16 = upload(h, packed16);
    out->scale = upload(h, scale);
    out->shift = upload(h, shift);
    out->bias = nullptr;
    if (bias_key) {
        const HostTensor* b;
        if (!need(h, *bias_key, {cout}, &b)) return false;
        out->bias = upload(h, b->v);
    }
    return out->w32 && out->w16 && out->scale && out->shift;
}

// fp16 copy with rows padded to a multiple of 8 elements (16-byte TMA pitch)
void add_f16_copy(egx_handle* h, const std::vector<float>& w, int out_f, int in, LinearW* out) {
    const int ld = (in + 7) / 8 * 8;
    std::vector<__half> v((size_t)out_f * ld, __float2half_rn(0.f));
    for (int o = 0; o < out_f; ++o)
        for (int i = 0; i < in; ++i) v[(size_t)o * ld + i] = __float2half_rn(w[(size_t)o * in + i]);
    out->w16 = upload(h, v);
    out->ldw = ld;
}

bool make_linear(egx_handle* h, const std::string& pre, int in, int out_f, bool bias, LinearW* out) {
    const HostTensor* w;
    if (!need(h, pre + ".weight", {out_f, in}, &w)) return false;
    out->in = in; out->out = out_f;
    out->w = upload(h, w->v);
    add_f16_copy(h, w->v, out_f, in, out);
    out->b = nullptr;
    if (bias) {
        const HostTensor* b;
        if (!need(h, pre + ".bias", {out_f}, &b)) return false;
        out->b = upload(h, b->v);
    }
    return out->w != nullptr;
}

bool make_ln(egx_handle* h, const std::string& pre, int d, LNW* out) {
    const HostTensor *g, *b;
    if (!need(h, pre + ".weight", {d}, &g) || !need(h, pre + ".bias", {d}, &b)) return false;
    out->g = upload(h, g->v);
    out->b = upload(h, b->v);
    return out->g && out->b;
}

bool make_concat_linear(egx_handle* h, std::initializer_list<std::string> pres, int in, int each_out, LinearW* out) {
    std::vector<float> cat;
    for (const auto& p : pres) {
        const HostTensor* w;
        if (!need(h, p + ".weight", {each_out, in}, &w)) return false;
        cat.insert(cat.end(), w->v.begin(), w->v.end());
    }
    out->in = in; out->out = each_out * (int)pres.size();
    out->w = upload(h, cat);
    add_f16_copy(h, cat, out->out, in, out);
    out->b = nullptr;
    return out->w != nullptr;
}

bool make_mha(egx_handle* h, const std::string& pre, const egx_cfg& c, MHAW* m) {
    const int hk = c.n_head * c.d_k, hv = c.n_head * c.d_v;
    if (hk != hv) { h->err = "n_head*d_k must equal n_head*d_v"; return false; }
    return make_linear(h, pre + ".w_qs", c.d_model, hk, false, &m->q) &&
           make_concat_linear(h, {pre + ".w_ks", pre + ".w_vs"}, c.d_model, hk, &m->kv) &&
           make_concat_linear(h, {pre + ".w_qs", pre + ".w_ks", pre + ".w_vs"}, c.d_model, hk, &m->qkv) &&
           make_linear(h, pre + ".fc", hv, c.d_model, false, &m->fc) &&
           make_ln(h, pre + ".layer_norm", c.d_model, &m->ln);
}

bool make_ffn(egx_handle* h, const std::string& pre, const egx_cfg& c, FFNW* f) {
    if (!(make_linear(h, pre + ".w_1", c.d_model, c.d_inner, true, &f->w1) &&
          make_linear(h, pre + ".w_2", c.d_inner, c.d_model, true, &f->w2) &&
          make_ln(h, pre + ".layer_norm", c.d_model, &f->ln)))
        return false;
    f->h_b1 = find(h, pre + ".w_1.bias")->v;
    f->h_b2 = find(h, pre + ".w_2.bias")->v;
    f->h_g = find(h, pre + ".layer_norm.weight")->v;
    f->h_b = find(h, pre + ".layer_norm.bias")->v;
    return true;
}

// SE-ResNet trunk under `fe` (Full_model/ResNetSE34V2.py:13-55; model/emotion_ResNetSE34V2.py adds layer4):
// stem conv+ReLU+BN, then n_layers stages of SEBasicBlocks, filters 32 << stage, depths [3,4,6,3].
bool pack_trunk(egx_handle* h, const std::string& fe, int n_layers, ConvW* stem, std::vector<BlockW>* blocks) {
    {
        const std::string bk = fe + ".conv1.bias";
        if (!make_conv(h, fe + ".conv1.weight", &bk, fe + ".bn1", 1, 32, 3, 1, 1, stem)) return false;
    }
    static const int nblk[4] = {3, 4, 6, 3};
    int cin = 32;
    for (int li = 0; li < n_layers; ++li)
        for (int b = 0; b < nblk[li]; ++b) {
            const std::string pre = fe + ".layer" + std::to_string(li + 1) + "." + std::to_string(b);
            const int cout = 32 << li, stride = (b == 0 && li > 0) ? 2 : 1;
            BlockW bw;
            if (!make_conv(h, pre + ".conv1.weight", nullptr, pre + ".bn1", cin, cout, 3, stride, 1, &bw.conv1)) return false;
            if (!make_conv(h, pre + ".conv2.weight", nullptr, pre + ".bn2", cout, cout, 3, 1, 0, &bw.conv2)) return false;
            bw.has_down = (stride != 1 || cin != cout);
            if (bw.has_down &&
                !make_conv(h, pre + ".downsample.0.weight", nullptr, pre + ".downsample.1", cin, cout, 1, stride, 0, &bw.down))
                return false;
            const int r = cout / 8;
            const HostTensor *w1, *b1, *w2, *b2;
            if (!need(h, pre + ".se.fc.0.weight", {r, cout}, &w1) || !need(h, pre + ".se.fc.0.bias", {r}, &b1) ||
                !need(h, pre + ".se.fc.2.weight", {cout, r}, &w2) || !need(h, pre + ".se.fc.2.bias", {cout}, &b2))
                return false;
            bw.se.c = cout; bw.se.r = r;
            bw.se.w1 = upload(h, w1->v); bw.se.b1 = upload(h, b1->v);
            bw.se.w2 = upload(h, w2->v); bw.se.b2 = upload(h, b2->v);
            blocks->push_back(bw);
            cin = cout;
        }
    return true;
}

// ---------------------------------------------------------------------------------------------
// log-mel tables (float64 on the host)
// ---------------------------------------------------------------------------------------------
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

bool build_logmel_tables(egx_handle* h) {
    const int n_fft = 1024, n_bins = 513, n_mels = 128;
    std::vector<float> window(n_fft);
    for (int i = 0; i < n_fft; ++i) window[i] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * i / n_fft));
    std::vector<float2> tw512(512), tw1024(513);
    for (int k = 0; k < 512; ++k)
        tw512[k] = make_float2((float)std::cos(-2.0 * kPi * k / 512), (float)std::sin(-2.0 * kPi * k / 512));
    for (int k = 0; k <= 512; ++k)
        tw1024[k] = make_float2((float)std::cos(-2.0 * kPi * k / 1024), (float)std::sin(-2.0 * kPi * k / 1024));
    // Slaney mel filterbank, librosa.filters.mel defaults (sr 16000, fmin 0, fmax 8000, norm='slaney')
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<int> start(n_mels), ptr(n_mels + 1, 0);
    std::vector<float> wts;
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        int first = -1, last = -1;
        std::vector<double> row(n_bins);
        for (int k = 0; k < n_bins; ++k) {
            const double f = 8000.0 * k / (n_bins - 1);
            const double lower = (f - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
            const double upper = (mel_f[m + 2] - f) / (mel_f[m + 2] - mel_f[m + 1]);
            row[k] = std::max(0.0, std::min(lower, upper)) * enorm;
            if (row[k] > 0.0) { if (first < 0) first = k; last = k; }
        }
        if (first < 0) { first = 0; last = -1; }
        start[m] = first;
        for (int k = first; k <= last; ++k) wts.push_back((float)row[k]);
        ptr[m + 1] = (int)wts.size();
    }
    // per-pass twiddles of the radix-8 Stockham FFT laid out [pass][t][j] (pass 0: Ns = 8, pass 1: Ns = 64), so that the
    // 32 lanes j of a warp read consecutive entries: tw_pass[pass][t][j] = W512^((j mod Ns) * t * 64 / Ns)
    std::vector<float2> tw_pass(2 * 8 * 64);
    for (int pass = 0; pass < 2; ++pass) {
        const int ns = pass ? 64 : 8;
        for (int t = 0; t < 8; ++t)
            for (int j = 0; j < 64; ++j) tw_pass[(pass * 8 + t) * 64 + j] = tw512[((j & (ns - 1)) * t * (64 / ns)) & 511];
    }
    // mel weights in ELL layout [tap][mel] (zero padded): lanes = mels read consecutive floats
    int max_taps = 0;
    std::vector<int> cnt(n_mels + 1, 0);
    for (int m = 0; m < n_mels; ++m) { cnt[m] = ptr[m + 1] - ptr[m]; max_taps = std::max(max_taps, cnt[m]); }
    std::vector<float> ell((size_t)max_taps * n_mels, 0.f);
    for (int m = 0; m < n_mels; ++m)
        for (int p = 0; p < cnt[m]; ++p) ell[(size_t)p * n_mels + m] = wts[ptr[m] + p];
    h->lm.window = upload(h, window);
    h->lm.tw512 = upload(h, tw_pass);
    h->lm.tw1024 = upload(h, tw1024);
    h->lm.mel_start = upload(h, start);
    h->lm.mel_ptr = upload(h, cnt);
    h->lm.mel_w = upload(h, ell);
    return h->lm.window && h->lm.tw512 && h->lm.tw1024 && h->lm.mel_start && h->lm.mel_ptr && h->lm.mel_w;
}

// ---------------------------------------------------------------------------------------------
// workspace plan: a bump allocator evaluated identically for sizing and for the forward
// ---------------------------------------------------------------------------------------------
struct Plan {
    size_t off = 0;
    char* base = nullptr;
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// fp32 scratch of the Prior_MemoryEncoder front (0 for the plain Prior_ConvEncoder): pred | enc | pred_enc | S
size_t mem_floats(const egx_handle* h, int B) {
    const MemPriorW& m = h->w.mem;
    if (!m.on) return 0;
    const size_t P = h->cfg.pose_dim;
    return (size_t)B * (m.n_pred * P + 2 * P + m.chunk) + P * m.chunk;
}

template <class T>
struct Slots {
    T *act[3], *down;
    float* se_sums;
    float *fcin, *t0, *spec_feat, *pconv, *prior_feat, *h0, *h1, *h2, *fus_in, *x_a, *x_b, *pre;
    float *qkv, *attn_o, *hid, *enc_out, *dec_out, *post0, *post1, *post2;
    float* mem;
};

template <class T>
Slots<T> plan_slots(const egx_handle* h, int B, Plan& p) {
    const egx_cfg& c = h->cfg;
    Slots<T> s;
    const size_t map1 = (size_t)B * h->H[0] * h->W[0] * 32;
    const size_t map2 = (size_t)B * h->H[1] * h->W[1] * 64;
    for (auto& a : s.act) a = p.take<T>(map1);
    s.down = p.take<T>(map2);
    // partial sums per (clip, tile/chunk, channel); the largest case is layer1 with 128-pixel tiles
    s.se_sums = p.take<float>((size_t)B * (size_t)(h->H[0] * h->W[0] / 64 + 8) * 32);
    const size_t R = (size_t)B * c.frames;
    const int hk = c.n_head * c.d_k;
    s.fcin = p.take<float>(R * h->H[2] * h->W[2]);
    s.t0 = p.take<float>(R * c.d_model);
    s.spec_feat = p.take<float>(R * c.d_model);
    s.pconv = p.take<float>(R * c.pose_dim);
    s.mem = p.take<float>(mem_floats(h, B));
    s.prior_feat = p.take<float>(R * c.d_model);
    s.h0 = p.take<float>((size_t)B * c.d_model);
    s.h1 = p.take<float>((size_t)B * 256);
    s.h2 = p.take<float>((size_t)B * 64);
    s.fus_in = p.take<float>(R * c.d_model);
    s.x_a = p.take<float>(R * c.d_model);
    s.x_b = p.take<float>(R * c.d_model);
    s.pre = p.take<float>(R * c.d_model);
    s.qkv = p.take<float>(R * 3 * hk);
    s.attn_o = p.take<float>(R * hk);
    s.hid = p.take<float>(R * c.d_inner);
    s.enc_out = p.take<float>(R * c.d_model);
    s.dec_out = p.take<float>(R * c.d_model);
    s.post0 = p.take<float>(R * 4 * c.d_model);
    s.post1 = p.take<float>(R * c.d_model);
    s.post2 = p.take<float>(R * c.pose_dim);
    return s;
}

// Stage tags follow SURVEY.md §8(d): S1 front-end, S2 stem, S3 trunk convs, S4 SE, S5 projection
// GEMMs outside the transformer, S6 encoder+decoder, S8 FGD statistics; 0 = other.
struct StageScope {
    egx_handle* h;
    int prev;
    StageScope(egx_handle* hh, int stage) : h(hh), prev(hh->stage) { h->stage = stage; }
    ~StageScope() { h->stage = prev; }
};

// When profiling is on, every launch is bracketed by a CUDA-event pair on the launching stream
// (read back by egx_profile_read); otherwise this is a plain counted launch.
#define LAUNCH(h, expr)                                                        \
    do {                                                                       \
        cudaEvent_t _e0 = nullptr, _e1 = nullptr;                              \
        const bool _prof = profile_acquire((h), &_e0, &_e1);                   \
        if (_prof) cudaEventRecord(_e0, s);                                    \
        const int _n = (expr);                                                 \
        if (_prof) cudaEventRecord(_e1, s);                                    \
        if (_n < 0) {                                                          \
            (h)->err = std::string("launch failed: ") + #expr + ": " +         \
                       cudaGetErrorString(cudaGetLastError());                 \
            return 1;                                                          \
        }                                                                      \
        (h)->launches += _n;                                                   \
    } while (0)

bool profile_acquire(egx_handle* h, cudaEvent_t* e0, cudaEvent_t* e1) {
    if (!h->profiling || h->prof_used + 2 > h->prof_events.size()) return false;
    *e0 = h->prof_events[h->prof_used];
    *e1 = h->prof_events[h->prof_used + 1];
    h->prof_stage.push_back(h->stage);
    h->prof_used += 2;
    return true;
}

int linear(egx_handle* h, const LinearW& w, const float* A, int M, float* C, int relu,
           const float* addend, int addend_rows, cudaStream_t s) {
    GemmEpi e;
    e.bias = w.b; e.relu = relu; e.addend = addend; e.addend_rows = addend_rows; e.addend_ld = w.out;
    LAUNCH(h, launch_gemm_f32(A, w.in, w.w, M, w.out, w.in, C, w.out, e, s));
    return 0;
}

// Runs the trunk; returns the buffer holding the output of `upto` (0 stem, 1..3 layers).
// Prior encoder up to the (B, F, P) map its Linear pair consumes: the conv pair alone (Prior_ConvEncoder), or
// pred_conv + spatial / temporal memory + cat(x, pred) (Prior_MemoryEncoder, see k_memory.cu).  scratch: mem_floats().

template <class T>
int run_prior_front(egx_handle* h, const float* prior, int B, float* scratch, T* out, int ldo, cudaStream_t s) {
    const Weights& w = h->w;
    const egx_cfg& c = h->cfg;
    const int p = c.prior_frames, F = c.frames, P = c.pose_dim;
    if (!w.mem.on) {
        LAUNCH(h, launch_prior_conv<T>(w, prior, B, p, F, P, out, ldo, s));
        return 0;
    }
    const MemPriorW& m = w.mem;
    float* pred = scratch;
    float* enc = pred + (size_t)B * m.n_pred * P;
    float* pred_enc = enc + (size_t)B * 2 * P;
    float* S = pred_enc + (size_t)B * m.chunk;
    LAUNCH(h, launch_prior_conv<float>(w, prior, B, p, m.n_pred, P, pred, P, s));
    GemmEpi e;
    e.bias = m.enc.b;
    // both chunk encoders on x[:, p-chunk:, :].reshape(B, chunk*P): a strided view of the prior poses
    LAUNCH(h, launch_gemm_f32(prior + (size_t)(p - m.chunk) * P, p * P, m.enc.w, B, 2 * P, m.chunk * P, enc, 2 * P, e, s));
    LAUNCH(h, launch_mem_spatial(pred, B, m.n_pred, P, m.chunk, enc, m.tm_w, m.tm_b, pred_enc, s));
    LAUNCH(h, launch_mem_batch_outer(enc, pred_enc, B, P, m.chunk, S, s));
    LAUNCH(h, launch_mem_temporal<T>(prior, pred, B, p, m.n_pred, P, m.chunk, enc, S, out, ldo, s));
    return 0;
}

template <class T>
int run_trunk(egx_handle* h, const float* spec, int B, Slots<T>& sl, int upto, T** result, cudaStream_t s) {
    T *x = sl.act[0], *y = sl.act[1], *z = sl.act[2];
    {
        StageScope sc(h, 2);
        LAUNCH(h, launch_stem<T>(h->w.stem, spec, B, h->H[0], h->W[0], x, s));
    }
    *result = x;
    if (upto == 0) return 0;
    static const int nblk[3] = {3, 4, 6};
    int bi = 0;
    int Hc = h->H[0], Wc = h->W[0];
    for (int li = 0; li < 3; ++li) {
        for (int b = 0; b < nblk[li]; ++b, ++bi) {
            const BlockW& bw = h->w.blocks[bi];
            const int Ho = h->H[li], Wo = h->W[li];
            const T* res = x;
            {
                StageScope sc(h, 3);
                LAUNCH(h, launch_conv_direct<T>(bw.conv1, x, B, Hc, Wc, y, nullptr, s));
                LAUNCH(h, launch_conv_direct<T>(bw.conv2, y, B, Ho, Wo, z, nullptr, s));
                if (bw.has_down) {
                    LAUNCH(h, launch_conv_direct<T>(bw.down, x, B, Hc, Wc, sl.down, nullptr, s));
                    res = sl.down;
                }
            }
            StageScope sc(h, 4);
            LAUNCH(h, launch_se_reduce<T>(z, B, Ho * Wo, bw.se.c, sl.se_sums, s));
            LAUNCH(h, launch_se_apply<T>(bw.se, z, res, sl.se_sums, se_partials(Ho * Wo), B, Ho * Wo, y, s));
            std::swap(x, y);
            Hc = Ho; Wc = Wo;
        }
        *result = x;
        if (upto == li + 1) return 0;
    }
    return 0;
}

template <class T>
int forward_impl(egx_handle* h, const float* spec, const float* prior, const float* sampled, int B,
                 float* poses, float* emo_feat, float* sem_feat, float* logits, void* ws, size_t ws_bytes,
                 cudaStream_t s) {
    const egx_cfg& c = h->cfg;
    Plan p;
    p.base = static_cast<char*>(ws);
    Slots<T> sl = plan_slots<T>(h, B, p);
    if (p.off > ws_bytes) EGX_FAIL(h, "workspace too small: need " + std::to_string(p.off));
    const Weights& w = h->w;
    const int R = B * c.frames, d = c.d_model, F = c.frames;
    const int hk = c.n_head * c.d_k;

    // --- audio encoder (Full_model/Models.py:118-133) ---
    T* t3 = nullptr;
    if (run_trunk<T>(h, spec, B, sl, 3, &t3, s)) return 1;
    {
        StageScope sc(h, 3);
        LAUNCH(h, launch_conv_direct<T>(w.final_conv, t3, B, h->H[2], h->W[2], nullptr, sl.fcin, s));
    }
    StageScope sc5(h, 5);
    if (linear(h, w.a_fc1, sl.fcin, R, sl.t0, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.a_fc2, sl.t0, R, sl.spec_feat, 0, nullptr, 0, s)) return 1;
    // --- prior encoder (Full_model/Models.py:199-212) ---
    if (run_prior_front<float>(h, prior, B, sl.mem, sl.pconv, c.pose_dim, s)) return 1;
    if (linear(h, w.p_fc1, sl.pconv, R, sl.t0, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.p_fc2, sl.t0, R, sl.prior_feat, 0, nullptr, 0, s)) return 1;
    // --- emotion / semantic projections, classifier head (Models.py:411-415) ---
    if (linear(h, w.emo0, sl.spec_feat, R, sl.t0, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.emo2, sl.t0, R, emo_feat, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.sem0, sl.spec_feat, R, sl.t0, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.sem2, sl.t0, R, sem_feat, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.hdr[0], emo_feat, B, sl.h0, 1, nullptr, 0, s)) return 1;
    if (linear(h, w.hdr[1], sl.h0, B, sl.h1, 1, nullptr, 0, s)) return 1;
    if (linear(h, w.hdr[2], sl.h1, B, sl.h2, 1, nullptr, 0, s)) return 1;
    if (linear(h, w.hdr[3], sl.h2, B, logits, 0, nullptr, 0, s)) return 1;
    // --- fusion (Models.py:417-418; Models_memory.py:551-555) + positional table (Models.py:46-48) ---
    LAUNCH(h, launch_add(sampled ? sampled : emo_feat, sem_feat, sl.fus_in, (int64_t)R * d, s));
    if (linear(h, w.fus0, sl.fus_in, R, sl.t0, 1, nullptr, 0, s)) return 1;
    if (linear(h, w.fus2, sl.t0, R, sl.x_a, 0, w.pos_table, F, s)) return 1;
    // --- encoder (Models.py:237-260; Layers.py:18-22) ---
    StageScope sc6(h, 6);
    float* x = sl.x_a;
    float* x2 = sl.x_b;
    for (int l = 0; l < c.n_layers; ++l) {
        const MHAW& a = w.enc_attn[l];
        const FFNW& f = w.enc_ffn[l];
        if (linear(h, a.qkv, x, R, sl.qkv, 0, nullptr, 0, s)) return 1;
        LAUNCH(h, launch_attention<float>(sl.qkv, 3 * hk, sl.qkv + hk, 3 * hk, sl.qkv + 2 * hk, 3 * hk, B, F, F,
                                   c.n_head, c.d_k, c.d_v, sl.attn_o, hk, s));
        if (linear(h, a.fc, sl.attn_o, R, sl.pre, 0, x, 0, s)) return 1;
        LAUNCH(h, launch_layernorm(sl.pre, a.ln, R, d, x2, nullptr, s));
        if (linear(h, f.w1, x2, R, sl.hid, 1, nullptr, 0, s)) return 1;
        if (linear(h, f.w2, sl.hid, R, sl.pre, 0, x2, 0, s)) return 1;
        LAUNCH(h, launch_layernorm(sl.pre, f.ln, R, d, l == c.n_layers - 1 ? sl.enc_out : x, nullptr, s));
    }
    // --- decoder (Models.py:279-293; Layers.py:50-58: cross-attention + FFN only) ---
    const float* dx = sl.prior_feat;
    for (int l = 0; l < c.n_layers; ++l) {
        const MHAW& a = w.dec_attn[l];
        const FFNW& f = w.dec_ffn[l];
        if (linear(h, a.q, dx, R, sl.qkv, 0, nullptr, 0, s)) return 1;
        if (linear(h, a.kv, sl.enc_out, R, sl.qkv + (size_t)R * hk, 0, nullptr, 0, s)) return 1;
        const float* kbuf = sl.qkv + (size_t)R * hk;
        LAUNCH(h, launch_attention<float>(sl.qkv, hk, kbuf, 2 * hk, kbuf + hk, 2 * hk, B, F, F, c.n_head, c.d_k,
                                   c.d_v, sl.attn_o, hk, s));
        if (linear(h, a.fc, sl.attn_o, R, sl.pre, 0, dx, 0, s)) return 1;
        LAUNCH(h, launch_layernorm(sl.pre, a.ln, R, d, x2, nullptr, s));
        if (linear(h, f.w1, x2, R, sl.hid, 1, nullptr, 0, s)) return 1;
        if (linear(h, f.w2, sl.hid, R, sl.pre, 0, x2, 0, s)) return 1;
        float* out = (l == c.n_layers - 1) ? sl.dec_out : sl.x_a;
        LAUNCH(h, launch_layernorm(sl.pre, f.ln, R, d, out, nullptr, s));
        dx = out;
    }
    // --- pose head (Models.py:352-360,425) ---
    StageScope sc5b(h, 5);
    if (linear(h, w.post[0], sl.dec_out, R, sl.post0, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.post[1], sl.post0, R, sl.post1, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.post[2], sl.post1, R, sl.post2, 0, nullptr, 0, s)) return 1;
    if (linear(h, w.post[3], sl.post2, R, poses, 0, nullptr, 0, s)) return 1;
    return 0;
}

template <class T>
int get_tap_impl(egx_handle* h, const std::string& name, const void* ws, int B, float* out, size_t cap,
                        size_t* n_out, cudaStream_t s) {
    Plan p;
    p.base = const_cast<char*>(static_cast<const char*>(ws));
    Slots<T> sl = plan_slots<T>(h, B, p);
    const egx_cfg& c = h->cfg;
    const size_t R = (size_t)B * c.frames;
    const float* src = nullptr;
    size_t n = 0;
    if (name == "spectrum_feature") { src = sl.spec_feat; n = R * c.d_model; }
    else if (name == "prior_feature") { src = sl.prior_feat; n = R * c.d_model; }
    else if (name == "enc_output") { src = sl.enc_out; n = R * c.d_model; }
    else if (name == "dec_output") { src = sl.dec_out; n = R * c.d_model; }
    else if (name == "final_conv") { src = sl.fcin; n = R * h->H[2] * h->W[2]; }
    else EGX_FAIL(h, "unknown tap: " + name);
    if (n > cap) EGX_FAIL(h, "tap output buffer too small");
    EGX_CHECK_CUDA(h, cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    *n_out = n;
    return 0;
}

template <class T>
int debug_trunk_impl(egx_handle* h, const float* spec, int B, int stage, float* out, size_t cap, size_t* n_out,
                            void* ws, size_t ws_bytes, cudaStream_t s) {
    Plan p;
    p.base = static_cast<char*>(ws);
    Slots<T> sl = plan_slots<T>(h, B, p);
    if (p.off > ws_bytes) EGX_FAIL(h, "workspace too small");
    T* res = nullptr;
    if (run_trunk<T>(h, spec, B, sl, stage, &res, s)) return 1;
    const int li = stage == 0 ? 0 : stage - 1;
    static const int filt[3] = {32, 64, 128};
    const int HW = h->H[li] * h->W[li], C = filt[li];
    const size_t n = (size_t)B * HW * C;
    if (n > cap) EGX_FAIL(h, "tap output buffer too small");
    LAUNCH(h, launch_nhwc_to_nchw_f32<T>(res, B, HW, C, out, s));
    *n_out = n;
    return 0;
}


// ---------------------------------------------------------------------------------------------
// tensor-core arm: fp16 operands (NHWC fp16 trunk maps, fp16 GEMM A operands), fp32 accumulation,
// fp32 residual stream / LayerNorm / outputs
// ---------------------------------------------------------------------------------------------
constexpr int kSeMaxC = 256;     // widest SE block (EmotionNet stage 4)

struct TcSlots {
    __half *act[3], *down;
    float* se_sums;
    __half* se_win;
    float *se_mean, *se_gate;
    __half *fcin, *t16, *spec16, *pconv16, *prior16, *emo16, *fus16, *h0, *h1, *h2, *x16, *x1_16, *qkv16, *o16,
        *hid16, *enc16, *dec16, *post0, *post1, *post2;
    float *spec_feat, *prior_feat, *x32a, *x32b, *pre, *enc_out, *dec_out, *mem;
    int P8;
};

TcSlots plan_tc(const egx_handle* h, int B, Plan& p) {
    const egx_cfg& c = h->cfg;
    TcSlots s;
    const size_t map1 = (size_t)B * h->H[0] * h->W[0] * 32;
    const size_t map2 = (size_t)B * h->H[1] * h->W[1] * 64;
    for (auto& a : s.act) a = p.take<__half>(map1);
    s.down = p.take<__half>(map2);
    s.se_sums = p.take<float>((size_t)B * (size_t)(h->H[0] * h->W[0] / 64 + 8) * 32);
    s.se_win = p.take<__half>((size_t)B * 9 * kSeMaxC);
    s.se_mean = p.take<float>((size_t)B * kSeMaxC);
    s.se_gate = p.take<float>((size_t)B * 2 * kSeMaxC);
    const size_t R = (size_t)B * c.frames;
    const int hk = c.n_head * c.d_k, d = c.d_model;
    s.P8 = (c.pose_dim + 7) / 8 * 8;
    s.fcin = p.take<__half>(R * h->H[2] * h->W[2]);
    s.t16 = p.take<__half>(R * d);
    s.spec16 = p.take<__half>(R * d);
    s.pconv16 = p.take<__half>(R * s.P8);
    s.mem = p.take<float>(mem_floats(h, B));
    s.prior16 = p.take<__half>(R * d);
    s.emo16 = p.take<__half>(R * d);
    s.fus16 = p.take<__half>(R * d);
    s.h0 = p.take<__half>((size_t)B * d);
    s.h1 = p.take<__half>((size_t)B * 256);
    s.h2 = p.take<__half>((size_t)B * 64);
    s.x16 = p.take<__half>(R * d);
    s.x1_16 = p.take<__half>(R * d);
    s.qkv16 = p.take<__half>(R * 3 * hk);
    s.o16 = p.take<__half>(R * hk);
    s.hid16 = p.take<__half>(R * c.d_inner);
    s.enc16 = p.take<__half>(R * d);
    s.dec16 = p.take<__half>(R * d);
    s.post0 = p.take<__half>(R * 4 * d);
    s.post1 = p.take<__half>(R * d);
    s.post2 = p.take<__half>(R * s.P8);
    s.spec_feat = p.take<float>(R * d);
    s.prior_feat = p.take<float>(R * d);
    s.x32a = p.take<float>(R * d);
    s.x32b = p.take<float>(R * d);
    s.pre = p.take<float>(R * d);
    s.enc_out = p.take<float>(R * d);
    s.dec_out = p.take<float>(R * d);
    return s;
}

// out = [relu](A W^T + b) [+ addend]; A fp16 (pitch lda), outputs fp32 and/or fp16 (pitch ld16)
int linear_tc(egx_handle* h, const LinearW& w, const __half* A, int lda, int M, float* out32, int ld32, __half* out16,
              int ld16, int relu, const float* addend, int addend_rows, cudaStream_t s) {
    GemmEpi e;
    e.bias = w.b; e.relu = relu; e.addend = addend; e.addend_rows = addend_rows; e.addend_ld = w.out;
    LAUNCH(h, launch_gemm_tc(A, lda, w.w16, w.ldw, M, w.out, w.in, e, out32, ld32, out16, ld16, s));
    return 0;
}

// out = LayerNorm(A W^T + b + residual) (Full_model/SubLayers.py:53-57, 80-82).  d_model == 256: residual add and
// LayerNorm run in the GEMM's epilogue (one thread owns the whole row), `pre` is untouched; otherwise the GEMM writes
// the pre-LayerNorm rows to `pre` and the warp-per-row kernel normalises them.
int linear_ln_tc(egx_handle* h, const LinearW& w, const LNW& ln, const __half* A, int lda, int M, const float* residual,
                 float* pre, float* out32, __half* out16, cudaStream_t s) {
    if (w.out == 256) {
        GemmEpi e;
        e.bias = w.b; e.addend = residual; e.addend_ld = w.out; e.ln_g = ln.g; e.ln_b = ln.b;
        LAUNCH(h, launch_gemm_tc(A, lda, w.w16, w.ldw, M, w.out, w.in, e, out32, w.out, out16, w.out, s));
        return 0;
    }
    if (linear_tc(h, w, A, lda, M, pre, w.out, nullptr, 0, 0, residual, 0, s)) return 1;
    LAUNCH(h, launch_layernorm(pre, ln, M, w.out, out32, out16, s));
    return 0;
}

// PositionwiseFeedForward (Full_model/SubLayers.py:74-84): one fused kernel when the geometry allows (d_model = 256),
// else w1 GEMM (+ReLU) -> hid16, w2 GEMM + residual + LayerNorm.  x16 / resid are the block input in fp16 / fp32.
int ffn_block_tc(egx_handle* h, const FFNW& f, const __half* x16, const float* resid, int M, __half* hid16, float* pre,
                 float* out32, __half* out16, cudaStream_t s) {
    static const bool fused = env_switch("EGX_FFN_FUSED", 1) != 0;
    const int d = f.w1.in, d_inner = f.w1.out;
    if (fused && ffn_tc_supported(d, d_inner) && (int)f.h_b1.size() == d_inner && (int)f.h_b2.size() == d) {
        LAUNCH(h, launch_ffn_tc(x16, resid, f.w1.w16, f.w1.ldw, f.h_b1.data(), f.w2.w16, f.w2.ldw, f.h_b2.data(), f.h_g.data(),
                                f.h_b.data(), M, d_inner, out32, out16, s));
        return 0;
    }
    if (linear_tc(h, f.w1, x16, d, M, nullptr, 0, hid16, d_inner, 1, nullptr, 0, s)) return 1;
    return linear_ln_tc(h, f.w2, f.ln, hid16, d_inner, M, resid, pre, out32, out16, s);
}

// Trunk on the tensor-core arm over caller-provided buffers: act[3] hold (B, H0, W0, 32) fp16 maps, `down` a
// (B, H0/2, W0/2, 64) one; returns the buffer and geometry of stage `upto` (0 stem, 1.. layers).
struct TrunkBufs {
    __half *act[3], *down;
    float* se_sums;
    __half* se_win;              // [B][9*C] window means of conv1's output (fp16 GEMM operand)
    float *se_mean, *se_gate;    // [B][C] mean of conv2's raw output; [B][2][C] folded gate
};

// EGX_SE_FUSED=0 (attribution experiments only) restores the separate gate*y + residual pass over the map
bool se_fused() {
    static const bool on = env_switch("EGX_SE_FUSED", 1) != 0;
    return on;
}

int run_trunk_tc(egx_handle* h, const ConvW& stem, const std::vector<BlockW>& blocks, int n_layers, int H0, int W0,
                 const float* spec, int B, const TrunkBufs& tb, int upto, __half** result, int* Hout, int* Wout,
                 cudaStream_t s) {
    __half *x = tb.act[0], *y = tb.act[1], *z = tb.act[2];
    {
        StageScope sc(h, 2);
        LAUNCH(h, launch_stem<__half>(stem, spec, B, H0, W0, x, s));
    }
    *result = x; *Hout = H0; *Wout = W0;
    if (upto == 0) return 0;
    static const int nblk[4] = {3, 4, 6, 3};
    int bi = 0;
    int Hc = H0, Wc = W0;
    for (int li = 0; li < n_layers; ++li) {
        const int Ho = li ? (Hc + 1) / 2 : Hc, Wo = li ? (Wc + 1) / 2 : Wc;
        for (int b = 0; b < nblk[li]; ++b, ++bi) {
            const BlockW& bw = blocks[bi];
            const __half* res = x;
            const int conv_stage = li == 0 ? 3 : 8 + li;      // 3 | 9 | 10 | 11: trunk convolutions of layer li + 1
            if (se_fused()) {
                // SE gate ahead of conv2 (k_trunk.cu K4c/K4d): conv1 sums its output per tile, the window means go
                // through conv2's own weights as a [B x 9C] x [9C x C] GEMM, and conv2's epilogue applies
                // relu(g * BN(acc) + residual) directly: the block's output map is written once, y2 never exists
                const int C = bw.conv2.cout;
                if (C > kSeMaxC || bw.conv2.cin != C || bw.conv2.ks != 3 || bw.conv2.stride != 1) EGX_FAIL(h, "unsupported SE block geometry");
                {
                    StageScope sc(h, conv_stage);
                    LAUNCH(h, launch_conv_tc(bw.conv1, x, B, Hc, Wc, y, 0, tb.se_sums, s));
                    if (bw.has_down) {
                        LAUNCH(h, launch_conv_tc(bw.down, x, B, Hc, Wc, tb.down, 0, nullptr, s));
                        res = tb.down;
                    }
                }
                {
                    StageScope sc(h, 4);
                    LAUNCH(h, launch_se_window(y, B, Ho, Wo, C, tb.se_sums,
                                               conv_tc_tiles_per_clip(bw.conv1.cin, bw.conv1.cout, Ho, Wo), tb.se_win, s));
                    LAUNCH(h, launch_gemm_tc(tb.se_win, 9 * C, bw.conv2.w16, 9 * C, B, C, 9 * C, GemmEpi{}, tb.se_mean, C, nullptr, 0, s));
                    LAUNCH(h, launch_se_gate(bw.se, bw.conv2, tb.se_mean, B, tb.se_gate, s));
                }
                StageScope sc(h, conv_stage);
                LAUNCH(h, launch_conv_tc(bw.conv2, y, B, Ho, Wo, z, 0, nullptr, s, tb.se_gate, res));
                // rotate: the block's output z becomes x; the old x (the residual) is free again
                __half* t = x; x = z; z = t;
                Hc = Ho; Wc = Wo;
                continue;
            }
            {
                StageScope sc(h, 3);
                LAUNCH(h, launch_conv_tc(bw.conv1, x, B, Hc, Wc, y, 0, nullptr, s));
                // conv2's epilogue also emits the per-tile channel sums the SE gate averages
                LAUNCH(h, launch_conv_tc(bw.conv2, y, B, Ho, Wo, z, 0, tb.se_sums, s));
                if (bw.has_down) {
                    LAUNCH(h, launch_conv_tc(bw.down, x, B, Hc, Wc, tb.down, 0, nullptr, s));
                    res = tb.down;
                }
            }
            StageScope sc(h, 4);
            LAUNCH(h, launch_se_apply<__half>(bw.se, z, res, tb.se_sums, conv_tc_tiles_per_clip(bw.conv2.cin, bw.conv2.cout, Ho, Wo), B, Ho * Wo, y,
                                              s));
            std::swap(x, y);
            Hc = Ho; Wc = Wo;
        }
        *result = x; *Hout = Hc; *Wout = Wc;
        if (upto == li + 1) return 0;
    }
    return 0;
}

int run_trunk_tc(egx_handle* h, const float* spec, int B, TcSlots& sl, int upto, __half** result, cudaStream_t s) {
    TrunkBufs tb{{sl.act[0], sl.act[1], sl.act[2]}, sl.down, sl.se_sums, sl.se_win, sl.se_mean, sl.se_gate};
    int Ho, Wo;
    return run_trunk_tc(h, h->w.stem, h->w.blocks, 3, h->H[0], h->W[0], spec, B, tb, upto, result, &Ho, &Wo, s);
}

// Transformer tail of the tensor-core arm for `B` clips whose rows start at r0 of the full-batch taps: encoder-tail
// Linear chain, projections, classifier header, fusion, encoder, decoder, pose head (Full_model/Models.py:124-130,
// 199-212, 237-293, 411-425).  Chunk-local intermediates live in the first rows of the full-size buffers of `sl`.
int forward_tail_tc(egx_handle* h, TcSlots& sl, int B, const __half* fcin, const __half* pconv16, const float* sampled,
                    float* poses, float* emo_feat, float* sem_feat, float* logits, size_t r0, cudaStream_t s) {
    const egx_cfg& c = h->cfg;
    const Weights& w = h->w;
    const int R = B * c.frames, d = c.d_model, F = c.frames, P = c.pose_dim, P8 = sl.P8;
    const int hk = c.n_head * c.d_k, HW3 = h->H[2] * h->W[2];
    float* spec_feat = sl.spec_feat + r0 * d;       // taps keep the whole batch
    float* prior_feat = sl.prior_feat + r0 * d;
    float* enc_out = sl.enc_out + r0 * d;
    float* dec_out = sl.dec_out + r0 * d;
    StageScope sc5(h, 5);
    if (linear_tc(h, w.a_fc, fcin, HW3, R, spec_feat, d, sl.spec16, d, 0, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.p_fc, pconv16, P8, R, prior_feat, d, sl.prior16, d, 0, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.emo, sl.spec16, d, R, emo_feat, d, sl.emo16, d, 0, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.sem, sl.spec16, d, R, sem_feat, d, nullptr, 0, 0, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.hdr[0], sl.emo16, F * d, B, nullptr, 0, sl.h0, d, 1, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.hdr[1], sl.h0, d, B, nullptr, 0, sl.h1, 256, 1, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.hdr[2], sl.h1, 256, B, nullptr, 0, sl.h2, 64, 1, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.hdr[3], sl.h2, 64, B, logits, 8, nullptr, 0, 0, nullptr, 0, s)) return 1;
    LAUNCH(h, launch_add_f16(sampled ? sampled : emo_feat, sem_feat, sl.fus16, (int64_t)R * d, s));
    if (linear_tc(h, w.fus0, sl.fus16, d, R, nullptr, 0, sl.t16, d, 1, nullptr, 0, s)) return 1;
    if (linear_tc(h, w.fus2, sl.t16, d, R, sl.x32a, d, sl.x16, d, 0, w.pos_table, F, s)) return 1;

    StageScope sc6(h, 6);
    float *x32 = sl.x32a, *y32 = sl.x32b;
    for (int l = 0; l < c.n_layers; ++l) {
        const MHAW& a = w.enc_attn[l];
        const FFNW& f = w.enc_ffn[l];
        const bool last = l == c.n_layers - 1;
        if (linear_tc(h, a.qkv, sl.x16, d, R, nullptr, 0, sl.qkv16, 3 * hk, 0, nullptr, 0, s)) return 1;
        LAUNCH(h, launch_attention_tc(sl.qkv16, 3 * hk, 0, sl.qkv16, 3 * hk, hk, 2 * hk, B, F, c.n_head, sl.o16, hk, s));
        if (linear_ln_tc(h, a.fc, a.ln, sl.o16, hk, R, x32, sl.pre, y32, sl.x1_16, s)) return 1;
        if (ffn_block_tc(h, f, sl.x1_16, y32, R, sl.hid16, sl.pre, last ? enc_out : x32, last ? sl.enc16 : sl.x16, s)) return 1;
    }
    const float* dx32 = prior_feat;
    const __half* dx16 = sl.prior16;
    for (int l = 0; l < c.n_layers; ++l) {
        const MHAW& a = w.dec_attn[l];
        const FFNW& f = w.dec_ffn[l];
        const bool last = l == c.n_layers - 1;
        __half* kv = sl.qkv16 + (size_t)R * hk;
        if (linear_tc(h, a.q, dx16, d, R, nullptr, 0, sl.qkv16, hk, 0, nullptr, 0, s)) return 1;
        if (linear_tc(h, a.kv, sl.enc16, d, R, nullptr, 0, kv, 2 * hk, 0, nullptr, 0, s)) return 1;
        LAUNCH(h, launch_attention_tc(sl.qkv16, hk, 0, kv, 2 * hk, 0, hk, B, F, c.n_head, sl.o16, hk, s));
        if (linear_ln_tc(h, a.fc, a.ln, sl.o16, hk, R, dx32, sl.pre, y32, sl.x1_16, s)) return 1;
        float* o32 = last ? dec_out : x32;
        __half* o16 = last ? sl.dec16 : sl.x16;
        if (ffn_block_tc(h, f, sl.x1_16, y32, R, sl.hid16, sl.pre, o32, o16, s)) return 1;
        dx32 = o32; dx16 = o16;
    }
    StageScope sc5b(h, 5);
    if (linear_tc(h, w.post_all, sl.dec16, d, R, poses, P, nullptr, 0, 0, nullptr, 0, s)) return 1;
    return 0;
}

// clips per chunk of the transformer tail: a multiple of 6 (3 TED / 2 BEAT clips share an attention tile)
int tail_chunk_clips() {
    static const int n = [] {
        const int v = env_switch("EGX_TAIL_CHUNK", 0);
        return v <= 0 ? (1 << 30) / 6 * 6 : std::max(6, v / 6 * 6);
    }();
    return n;
}

int forward_tc(egx_handle* h, const float* spec, const float* prior, const float* sampled, int B, float* poses,
               float* emo_feat, float* sem_feat, float* logits, void* ws, size_t ws_bytes, cudaStream_t s) {
    const egx_cfg& c = h->cfg;
    Plan p;
    p.base = static_cast<char*>(ws);
    TcSlots sl = plan_tc(h, B, p);
    if (p.off > ws_bytes) EGX_FAIL(h, "workspace too small: need " + std::to_string(p.off));
    const Weights& w = h->w;
    const int d = c.d_model, F = c.frames, P = c.pose_dim, P8 = sl.P8;
    const int HW3 = h->H[2] * h->W[2];

    __half* t3 = nullptr;
    if (run_trunk_tc(h, spec, B, sl, 3, &t3, s)) return 1;
    {
        StageScope sc(h, 10);
        LAUNCH(h, launch_conv_tc(w.final_conv, t3, B, h->H[2], h->W[2], sl.fcin, 1, nullptr, s));
    }
    {
        StageScope sc5(h, 5);
        // the prior front runs on the whole batch: the memory variant's temporal memory sums over it
        if (run_prior_front<__half>(h, prior, B, sl.mem, sl.pconv16, P8, s)) return 1;
    }
    // Everything after the trunk can run in chunks of clips (EGX_TAIL_CHUNK, off by default).  The idea: its GEMMs are
    // bound by the bytes they write (QKV 428 MB, FFN hidden 285 MB per layer at B = 4096, both larger than the 126 MB
    // L2) and every such tensor is read back by the next kernel, so chunk-sized tensors would stay in L2.  Measured:
    // 34.9 ms/step unchunked, 35.2 at 1536 clips, 36.5 at 768 — the smaller launches lose more (tile quantisation,
    // fixed prologue per launch) than L2 residency returns — so one pass over the whole batch stays the default.
    // Per-clip results do not depend on the chunking (rows are independent, attention groups stay aligned).
    const int chunk = tail_chunk_clips();
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        const size_t r0 = (size_t)b0 * F;
        if (forward_tail_tc(h, sl, nb, sl.fcin + r0 * HW3, sl.pconv16 + r0 * P8, sampled ? sampled + r0 * d : nullptr,
                            poses + r0 * P, emo_feat + r0 * d, sem_feat + r0 * d, logits + (size_t)b0 * 8, r0, s))
            return 1;
    }
    return 0;
}

int get_tap_tc(egx_handle* h, const std::string& name, const void* ws, int B, float* out, size_t cap, size_t* n_out,
               cudaStream_t s) {
    Plan p;
    p.base = const_cast<char*>(static_cast<const char*>(ws));
    TcSlots sl = plan_tc(h, B, p);
    const size_t n = (size_t)B * h->cfg.frames * h->cfg.d_model;
    const float* src = nullptr;
    if (name == "spectrum_feature") src = sl.spec_feat;
    else if (name == "prior_feature") src = sl.prior_feat;
    else if (name == "enc_output") src = sl.enc_out;
    else if (name == "dec_output") src = sl.dec_out;
    else EGX_FAIL(h, "unknown tap: " + name);
    if (n > cap) EGX_FAIL(h, "tap output buffer too small");
    EGX_CHECK_CUDA(h, cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    *n_out = n;
    return 0;
}

int debug_trunk_tc(egx_handle* h, const float* spec, int B, int stage, float* out, size_t cap, size_t* n_out, void* ws,
                   size_t ws_bytes, cudaStream_t s) {
    Plan p;
    p.base = static_cast<char*>(ws);
    TcSlots sl = plan_tc(h, B, p);
    if (p.off > ws_bytes) EGX_FAIL(h, "workspace too small");
    __half* res = nullptr;
    if (run_trunk_tc(h, spec, B, sl, stage, &res, s)) return 1;
    const int li = stage == 0 ? 0 : stage - 1;
    static const int filt[3] = {32, 64, 128};
    const int HW = h->H[li] * h->W[li], C = filt[li];
    const size_t n = (size_t)B * HW * C;
    if (n > cap) EGX_FAIL(h, "tap output buffer too small");
    LAUNCH(h, launch_nhwc_to_nchw_f32<__half>(res, B, HW, C, out, s));
    *n_out = n;
    return 0;
}

}  // namespace


// ---------------------------------------------------------------------------------------------
// small networks: eval-mode algebra folded in float64 on the host
// ---------------------------------------------------------------------------------------------
namespace {

struct Affine {
    int in = 0, out = 0;
    std::vector<double> W, b;     // W[out][in]
};

bool get_affine(egx_handle* h, const std::string& pre, int in, int out, Affine* a) {
    const HostTensor *w, *b;
    if (!need(h, pre + ".weight", {out, in}, &w) || !need(h, pre + ".bias", {out}, &b)) return false;
    a->in = in; a->out = out;
    a->W.assign(w->v.begin(), w->v.end());
    a->b.assign(b->v.begin(), b->v.end());
    return true;
}

// second(first(x))
Affine compose(const Affine& first, const Affine& second) {
    Affine r;
    r.in = first.in; r.out = second.out;
    r.W.assign((size_t)r.out * r.in, 0.0);
    r.b = second.b;
    for (int o = 0; o < second.out; ++o)
        for (int m = 0; m < second.in; ++m) {
            const double s = second.W[(size_t)o * second.in + m];
            r.b[o] += s * first.b[m];
            for (int i = 0; i < first.in; ++i) r.W[(size_t)o * r.in + i] += s * first.W[(size_t)m * first.in + i];
        }
    return r;
}

// y = x * s + t for an eval-mode BatchNorm1d under `pre`
bool bn_scale_shift(egx_handle* h, const std::string& pre, int c, std::vector<double>& s, std::vector<double>& t) {
    const HostTensor *g, *b, *m, *v;
    if (!need(h, pre + ".weight", {c}, &g) || !need(h, pre + ".bias", {c}, &b) ||
        !need(h, pre + ".running_mean", {c}, &m) || !need(h, pre + ".running_var", {c}, &v))
        return false;
    s.resize(c); t.resize(c);
    for (int i = 0; i < c; ++i) {
        s[i] = (double)g->v[i] / std::sqrt((double)v->v[i] + (double)kBnEps);
        t[i] = (double)b->v[i] - (double)m->v[i] * s[i];
    }
    return true;
}

void scale_rows(Affine& a, const std::vector<double>& s, const std::vector<double>& t) {   // BN after a Linear
    for (int o = 0; o < a.out; ++o) {
        for (int i = 0; i < a.in; ++i) a.W[(size_t)o * a.in + i] *= s[o];
        a.b[o] = a.b[o] * s[o] + t[o];
    }
}

std::vector<float> to_f32(const std::vector<double>& v) { return std::vector<float>(v.begin(), v.end()); }

// chain of Linears at Sequential indices idx[] under `pre` (the Dropouts between them are identities in eval)
bool linear_chain(egx_handle* h, const std::string& pre, std::initializer_list<int> idx, std::initializer_list<int> dims, Affine* out) {
    auto d = dims.begin();
    bool first = true;
    for (int i : idx) {
        Affine a;
        if (!get_affine(h, pre + "." + std::to_string(i), d[0], d[1], &a)) return false;
        *out = first ? a : compose(*out, a);
        first = false;
        ++d;
    }
    return true;
}

struct BucketScope {
    egx_handle* h;
    BucketScope(egx_handle* hh, const std::string& family) : h(hh) {
        auto& b = h->aux_owned[family];
        for (void* p : b) cudaFree(p);
        b.clear();
        h->cur_bucket = &b;
    }
    ~BucketScope() { h->cur_bucket = nullptr; }
    bool ok() const {
        for (void* p : *h->cur_bucket) if (!p) return false;
        return true;
    }
};

// C4: Full_model/BEAT_CVAE.py:32-83 (layers), :98-136 (forward / sample)
int pack_cvae(egx_handle* h) {
    BucketScope scope(h, "cvae");
    h->cvae = CvaeW();
    Affine enc, mu, lv, py, fus, dec;
    if (!linear_chain(h, "cvae.Encoder", {0, 2, 4, 6, 8}, {90, 128, 128, 256, 256, 512}, &enc)) return 1;
    if (!get_affine(h, "cvae.fc_mu", 512, 32, &mu) || !get_affine(h, "cvae.fc_var", 512, 32, &lv)) return 1;
    if (!linear_chain(h, "cvae.Posterior_Y_embedding", {0, 2}, {90, 64, 32}, &py)) return 1;
    if (!linear_chain(h, "cvae.fusion_z_posterior", {0, 2}, {64, 256, 512}, &fus)) return 1;
    if (!linear_chain(h, "cvae.Decoder", {0, 2, 4, 6, 8}, {512, 256, 256, 128, 128, 90}, &dec)) return 1;
    const Affine xmu = compose(enc, mu), xlv = compose(enc, lv), zd = compose(fus, dec);
    std::vector<float> w_x(90 * 64), b_x(64), w_y(90 * 32), b_y(32), w_d(64 * 92, 0.f), b_d(92, 0.f);
    for (int k = 0; k < 90; ++k)
        for (int o = 0; o < 32; ++o) {
            w_x[k * 64 + o] = (float)xmu.W[(size_t)o * 90 + k];
            w_x[k * 64 + 32 + o] = (float)xlv.W[(size_t)o * 90 + k];
            w_y[k * 32 + o] = (float)py.W[(size_t)o * 90 + k];
        }
    for (int o = 0; o < 32; ++o) { b_x[o] = (float)xmu.b[o]; b_x[32 + o] = (float)xlv.b[o]; b_y[o] = (float)py.b[o]; }
    for (int k = 0; k < 64; ++k)
        for (int o = 0; o < 90; ++o) w_d[k * 92 + o] = (float)zd.W[(size_t)o * 64 + k];
    for (int o = 0; o < 90; ++o) b_d[o] = (float)zd.b[o];
    CvaeW& c = h->cvae;
    c.w_x = upload(h, w_x); c.b_x = upload(h, b_x); c.w_y = upload(h, w_y); c.b_y = upload(h, b_y);
    c.w_d = upload(h, w_d); c.b_d = upload(h, b_d);
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing cvae weights");
    c.ready = true;
    return 0;
}

bool upload_tensor(egx_handle* h, const std::string& key, std::initializer_list<int64_t> shape, float** out) {
    const HostTensor* t;
    if (!need(h, key, shape, &t)) return false;
    *out = upload(h, t->v);
    return true;
}

bool upload_bn(egx_handle* h, const std::string& pre, int c, float** s_out, float** t_out) {
    std::vector<double> s, t;
    if (!bn_scale_shift(h, pre, c, s, t)) return false;
    *s_out = upload(h, to_f32(s));
    *t_out = upload(h, to_f32(t));
    return true;
}

// E1: CAVE/BEAT_CVAE.py:334-388 (layers), :427-447 (sample)
int pack_cvae3(egx_handle* h) {
    BucketScope scope(h, "cvae3");
    h->cvae3 = Cvae3W();
    Cvae3W& c = h->cvae3;
    Affine py, fus;
    if (!linear_chain(h, "cvae3.Posterior_Y_embedding", {0, 2}, {8, 16, 32}, &py)) return 1;
    if (!linear_chain(h, "cvae3.fusion_z_posterior", {0, 2}, {64, 128, 512}, &fus)) return 1;
    c.w_y = upload(h, to_f32(py.W)); c.b_y = upload(h, to_f32(py.b));
    c.w_f = upload(h, to_f32(fus.W)); c.b_f = upload(h, to_f32(fus.b));
    const std::string d = "cvae3.Decoder.";
    if (!upload_tensor(h, d + "0.weight", {4, 8, 3}, &c.t1_w) || !upload_tensor(h, d + "0.bias", {8}, &c.t1_b) ||
        !upload_bn(h, d + "2", 8, &c.s1, &c.h1) ||
        !upload_tensor(h, d + "3.weight", {8, 16, 3}, &c.t2_w) || !upload_tensor(h, d + "3.bias", {16}, &c.t2_b) ||
        !upload_bn(h, d + "5", 16, &c.s2, &c.h2) ||
        !upload_tensor(h, d + "6.weight", {32, 16, 3}, &c.c3_w) || !upload_tensor(h, d + "6.bias", {32}, &c.c3_b) ||
        !upload_bn(h, d + "8", 32, &c.s3, &c.h3) ||
        !upload_tensor(h, d + "9.weight", {60, 32, 3}, &c.c4_w) || !upload_tensor(h, d + "9.bias", {60}, &c.c4_b) ||
        !upload_bn(h, d + "11", 60, &c.s4, &c.h4) ||
        !upload_tensor(h, d + "12.weight", {60, 60, 3}, &c.c5_w) || !upload_tensor(h, d + "12.bias", {60}, &c.c5_b))
        return 1;
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing cvae3 weights");
    c.ready = true;
    return 0;
}

// conv weight (cout, cin, k) with the following BatchNorm folded in: w' = s w, b' = s b + t
bool upload_conv_bn(egx_handle* h, const std::string& conv, const std::string* bn, int cout, int cin, int k, float** w_out,
                    float** b_out) {
    const HostTensor *w, *b;
    if (!need(h, conv + ".weight", {cout, cin, k}, &w) || !need(h, conv + ".bias", {cout}, &b)) return false;
    std::vector<double> s(cout, 1.0), t(cout, 0.0);
    if (bn && !bn_scale_shift(h, *bn, cout, s, t)) return false;
    std::vector<float> wf(w->v.size()), bf(cout);
    const size_t per = (size_t)cin * k;
    for (int o = 0; o < cout; ++o) {
        for (size_t i = 0; i < per; ++i) wf[o * per + i] = (float)(s[o] * (double)w->v[o * per + i]);
        bf[o] = (float)(s[o] * (double)b->v[o] + t[o]);
    }
    *w_out = upload(h, wf);
    *b_out = upload(h, bf);
    return true;
}

// D1: PoseEncoderConv.  model/motion_ae.py:55-83 (out_net -> latent) / model/embedding_net.py:37-83 (out_net -> fc_mu)
int pack_pose_enc(egx_handle* h, const std::string& family, const std::string& pre, bool with_fc_mu, PoseEncW* out) {
    BucketScope scope(h, family);
    *out = PoseEncW();
    const HostTensor* w1 = find(h, pre + "net.0.0.weight");
    const HostTensor* fc0 = find(h, pre + "out_net.0.weight");
    const HostTensor* fc6 = find(h, pre + "out_net.6.weight");
    if (!w1 || !fc0 || !fc6 || w1->shape.size() != 3 || fc0->shape.size() != 2 || fc6->shape.size() != 2) return 1;
    const int P = (int)w1->shape[1], n_flat = (int)fc0->shape[1], L4 = n_flat / 32, L3 = L4 + 2, L2 = 2 * (L3 - 1) + 4, L = L2 + 4;
    if (n_flat % 32 || L4 < 1) EGX_FAIL(h, "unexpected out_net.0 width for " + family);
    const std::string bn0 = pre + "net.0.1", bn1 = pre + "net.1.1", bn2 = pre + "net.2.1";
    if (!upload_conv_bn(h, pre + "net.0.0", &bn0, 32, P, 3, &out->w1, &out->b1) ||
        !upload_conv_bn(h, pre + "net.1.0", &bn1, 64, 32, 3, &out->w2, &out->b2) ||
        !upload_conv_bn(h, pre + "net.2.0", &bn2, 64, 64, 4, &out->w3, &out->b3) ||
        !upload_conv_bn(h, pre + "net.3", nullptr, 32, 64, 3, &out->w4, &out->b4))
        return 1;
    // out_net: Linear BN LeakyReLU(True) Linear BN LeakyReLU(True) Linear; nn.LeakyReLU(True) sets negative_slope = 1.0,
    // i.e. the identity (model/motion_ae.py:70,73; model/embedding_net.py:57,60), so the whole tail is affine
    const int latent = (int)fc6->shape[0];
    Affine a0, a3, a6;
    std::vector<double> s, t;
    if (!get_affine(h, pre + "out_net.0", n_flat, 256, &a0) || !bn_scale_shift(h, pre + "out_net.1", 256, s, t)) return 1;
    scale_rows(a0, s, t);
    if (!get_affine(h, pre + "out_net.3", 256, 128, &a3) || !bn_scale_shift(h, pre + "out_net.4", 128, s, t)) return 1;
    scale_rows(a3, s, t);
    if (!get_affine(h, pre + "out_net.6", 128, latent, &a6)) return 1;
    Affine tail = compose(compose(a0, a3), a6);
    if (with_fc_mu) {
        Affine mu;
        if (!get_affine(h, pre + "fc_mu", latent, 32, &mu)) return 1;
        tail = compose(tail, mu);
    }
    out->w_fc = upload(h, to_f32(tail.W));
    out->b_fc = upload(h, to_f32(tail.b));
    out->L = L; out->P = P; out->n_out = tail.out;
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing " + family);
    out->ready = true;
    return 0;
}


// C3: model/audio_emotion_classifer.py:17-49 — ResNetSE [3,4,6,3] x [32,64,128,256] + six Linears.
// The first Linear reads the reference's NCHW flatten (c, h, w); our maps are NHWC, so its columns are permuted
// once here and the layer-4 map is the GEMM's A operand as it lies in memory.
int pack_emotion_net(egx_handle* h) {
    BucketScope scope(h, "emotion_net");
    h->emo = EmotionNetW();
    EmotionNetW& e = h->emo;
    if (!pack_trunk(h, "emotion_net.emotion_encoder", 4, &e.stem, &e.blocks)) return 1;
    const HostTensor* w0 = find(h, "emotion_net.emotion_eocder_fc.0.weight");
    if (!w0 || w0->shape.size() != 2 || w0->shape[1] % 256) EGX_FAIL(h, "emotion_net: unexpected first Linear");
    const int flat = (int)w0->shape[1], hw = flat / 256;
    const int dims[7] = {flat, 4096, 2048, 512, 128, 64, 8};
    for (int i = 0; i < 6; ++i) {
        const std::string pre = i < 5 ? "emotion_net.emotion_eocder_fc." + std::to_string(2 * i) : std::string("emotion_net.last_fc");
        const HostTensor *w, *b;
        if (!need(h, pre + ".weight", {dims[i + 1], dims[i]}, &w) || !need(h, pre + ".bias", {dims[i + 1]}, &b)) return 1;
        LinearW& l = e.fc[i];
        l.in = dims[i]; l.out = dims[i + 1]; l.ldw = dims[i]; l.w = nullptr;
        std::vector<__half> w16((size_t)l.out * l.in);
        if (i == 0) {
            for (int o = 0; o < l.out; ++o)
                for (int c = 0; c < 256; ++c)
                    for (int p = 0; p < hw; ++p)
                        w16[(size_t)o * flat + (size_t)p * 256 + c] = __float2half_rn(w->v[(size_t)o * flat + (size_t)c * hw + p]);
        } else {
            for (size_t j = 0; j < w16.size(); ++j) w16[j] = __float2half_rn(w->v[j]);
        }
        l.w16 = upload(h, w16);
        l.b = upload(h, b->v);
    }
    e.flat = flat;
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing emotion_net");
    e.ready = true;
    return 0;
}

struct EmoSlots {
    TrunkBufs tb;
    __half* fc[5];
};

EmoSlots plan_emotion(int B, int H0, int W0, Plan& p) {
    EmoSlots s;
    const size_t map1 = (size_t)B * H0 * W0 * 32;
    for (auto& a : s.tb.act) a = p.take<__half>(map1);
    s.tb.down = p.take<__half>((size_t)B * ((H0 + 1) / 2) * ((W0 + 1) / 2) * 64);
    s.tb.se_sums = p.take<float>((size_t)B * (size_t)(H0 * W0 / 64 + 8) * 32);
    s.tb.se_win = p.take<__half>((size_t)B * 9 * kSeMaxC);
    s.tb.se_mean = p.take<float>((size_t)B * kSeMaxC);
    s.tb.se_gate = p.take<float>((size_t)B * 2 * kSeMaxC);
    static const int dims[5] = {4096, 2048, 512, 128, 64};
    for (int i = 0; i < 5; ++i) s.fc[i] = p.take<__half>((size_t)B * dims[i]);
    return s;
}

// Linear chain `keys` (state_dict prefixes, dims[i] -> dims[i+1]) as ONE LinearW
bool make_collapsed(egx_handle* h, std::initializer_list<std::string> keys, std::initializer_list<int> dims, LinearW* out) {
    auto d = dims.begin();
    Affine acc;
    bool first = true;
    for (const auto& k : keys) {
        Affine a;
        if (!get_affine(h, k, d[0], d[1], &a)) return false;
        acc = first ? a : compose(acc, a);
        first = false;
        ++d;
    }
    out->in = acc.in; out->out = acc.out;
    const std::vector<float> wf = to_f32(acc.W);
    out->w = upload(h, wf);
    out->b = upload(h, to_f32(acc.b));
    add_f16_copy(h, wf, acc.out, acc.in, out);
    return out->w && out->b && out->w16;
}

// D1: per-frame MLP of model/FGD.py:30-41 (Encoder: Linear(282,512) Linear(512,512) Linear(512,512), Dropouts between)
int pack_fgd_mlp(egx_handle* h) {
    BucketScope scope(h, "fgd_mlp");
    h->fgd_mlp = RowMlpW();
    const HostTensor* w0 = find(h, "fgd_mlp.Encoder.0.weight");
    if (!w0 || w0->shape.size() != 2) return 1;
    const int in = (int)w0->shape[1], hid = (int)w0->shape[0];
    Affine enc;
    if (!linear_chain(h, "fgd_mlp.Encoder", {0, 2, 4}, {in, hid, hid, hid}, &enc)) return 1;
    LinearW& l = h->fgd_mlp.lin;
    l.in = in; l.out = hid;
    const std::vector<float> wf = to_f32(enc.W);
    l.w = upload(h, wf);
    l.b = upload(h, to_f32(enc.b));
    add_f16_copy(h, wf, hid, in, &l);
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing fgd_mlp");
    h->fgd_mlp.ready = true;
    return 0;
}

// (f)2: skeleton_classifer/Models.py:199-283 (Transformer), :87-121 (Prior_Encoder), :127-172 (Encoder)
int pack_skeleton(egx_handle* h) {
    BucketScope scope(h, "skel");
    h->skel = SkeletonW();
    SkeletonW& k = h->skel;
    const HostTensor* fc1 = find(h, "skel.prior_seq_encoder.fc1.weight");
    const HostTensor* pt = find(h, "skel.encoder.position_enc.pos_table");
    const HostTensor* wq = find(h, "skel.encoder.layer_stack.0.slf_attn.w_qs.weight");
    const HostTensor* w1 = find(h, "skel.encoder.layer_stack.0.pos_ffn.w_1.weight");
    const HostTensor* last = find(h, "skel.post_projector.8.weight");
    if (!fc1 || !pt || !wq || !w1 || !last || fc1->shape.size() != 2 || pt->shape.size() != 3 || wq->shape.size() != 2 ||
        w1->shape.size() != 2 || last->shape.size() != 2)
        EGX_FAIL(h, "skeleton classifier: missing prior_seq_encoder / encoder / post_projector weights");
    k.d = (int)fc1->shape[0]; k.P = (int)fc1->shape[1];
    k.T = (int)pt->shape[1];
    k.d_inner = (int)w1->shape[0];
    k.n_class = (int)last->shape[0];
    if (wq->shape[0] % 64 || (int)pt->shape[2] != k.d || k.d % 8)
        EGX_FAIL(h, "skeleton classifier: the attention kernel needs d_k = d_v = 64 (as the evaluation script builds it)");
    k.n_head = (int)wq->shape[0] / 64;
    while (find(h, "skel.encoder.layer_stack." + std::to_string(k.n_layers) + ".slf_attn.w_qs.weight")) ++k.n_layers;
    if (!make_collapsed(h, {"skel.prior_seq_encoder.fc1", "skel.prior_seq_encoder.fc2"}, {k.P, k.d, k.d}, &k.prior)) return 1;
    k.pos_table = upload(h, pt->v);
    egx_cfg c{};
    c.d_model = k.d; c.d_inner = k.d_inner; c.n_head = k.n_head; c.d_k = 64; c.d_v = 64;
    k.attn.resize(k.n_layers); k.ffn.resize(k.n_layers);
    for (int l = 0; l < k.n_layers; ++l) {
        const std::string e = "skel.encoder.layer_stack." + std::to_string(l);
        if (!make_mha(h, e + ".slf_attn", c, &k.attn[l]) || !make_ffn(h, e + ".pos_ffn", c, &k.ffn[l])) return 1;
    }
    const int dims[6] = {k.T * k.d, 4 * k.d, k.d, 128, 64, k.n_class};
    for (int i = 0; i < 5; ++i)
        if (!make_linear(h, "skel.post_projector." + std::to_string(2 * i), dims[i], dims[i + 1], true, &k.post[i])) return 1;
    if (!scope.ok()) EGX_FAIL(h, "device allocation failed while packing the skeleton classifier");
    k.ready = true;
    return 0;
}

struct SkelSlots {
    __half *a16, *x16, *x1_16, *qkv16, *o16, *hid16, *p16[4];
    float *x32a, *x32b, *pre;
    int P8;
};

SkelSlots plan_skeleton(const SkeletonW& k, int B, Plan& p) {
    SkelSlots s;
    const size_t R = (size_t)B * k.T;
    const int hk = k.n_head * 64;
    s.P8 = (k.P + 7) / 8 * 8;
    s.a16 = p.take<__half>(R * s.P8);
    s.x16 = p.take<__half>(R * k.d);
    s.x1_16 = p.take<__half>(R * k.d);
    s.qkv16 = p.take<__half>(R * 3 * hk);
    s.o16 = p.take<__half>(R * hk);
    s.hid16 = p.take<__half>(R * k.d_inner);
    const int pd[4] = {4 * k.d, k.d, 128, 64};
    for (int i = 0; i < 4; ++i) s.p16[i] = p.take<__half>((size_t)B * pd[i]);
    s.x32a = p.take<float>(R * k.d);
    s.x32b = p.take<float>(R * k.d);
    s.pre = p.take<float>(R * k.d);
    return s;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int egx_version(void) { return EGX_VERSION; }

int egx_create(const egx_cfg* cfg, int device, egx_handle** out) {
    if (!cfg || !out) return 1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return 2;
    if (cudaSetDevice(device) != cudaSuccess) return 2;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 2;
    if (prop.major != 10) return 3;   // sm_100a only: no other architecture, no fallback
    auto* h = new egx_handle();
    h->cfg = *cfg;
    h->device = device;
    const egx_cfg& c = h->cfg;
    if (c.n_mels != 128 || c.d_model % 32 || c.frames > c.n_position || c.frames > 64 ||
        c.d_k > 64 || c.d_v > 64) {
        delete h;
        return 4;
    }
    if (cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { delete h; return 2; }
    h->H[0] = c.n_mels; h->W[0] = c.spec_w;
    for (int i = 1; i < 3; ++i) { h->H[i] = (h->H[i - 1] + 1) / 2; h->W[i] = (h->W[i - 1] + 1) / 2; }
    if (!build_logmel_tables(h)) { egx_destroy(h); return 5; }
    if (gemm_tc_init_device() != 0 || conv_tc_init_device() != 0 || attn_tc_init_device() != 0 || ffn_tc_init_device() != 0) {
        egx_destroy(h);
        return 6;
    }
    if (c.precision == EGX_PREC_TC && (c.d_k != 64 || c.d_v != 64)) { egx_destroy(h); return 4; }
    *out = h;
    return 0;
}

void egx_destroy(egx_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (void* p : h->owned) cudaFree(p);
    for (auto& kv : h->aux_owned)
        for (void* p : kv.second) cudaFree(p);
    if (h->fgd_scratch) cudaFree(h->fgd_scratch);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    delete h;
}

const char* egx_last_error(const egx_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t egx_launch_count(const egx_handle* h) { return h ? h->launches : 0; }

int egx_set_weight(egx_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
    if (!h || !key || !data) return 1;
    if (dtype != EGX_DTYPE_F32) return 0;   // num_batches_tracked etc.: nothing to keep
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    t.v.resize((size_t)t.numel());
    EGX_CHECK_CUDA(h, cudaMemcpy(t.v.data(), data, t.v.size() * sizeof(float), cudaMemcpyDeviceToHost));
    h->staged[key] = std::move(t);
    h->finalized = false;
    return 0;
}

int egx_finalize_weights(egx_handle* h) {
    if (!h) return 1;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    // Pack every model family whose keys were staged since the last call (generator keys are un-prefixed,
    // the small networks arrive under "cvae.", "cvae3.", "motion_ae.", "pose_enc.", "fgd_mlp.").
    int packed = 0;
    if (h->staged.count("cvae.Encoder.0.weight")) { if (pack_cvae(h)) return 1; ++packed; }
    if (h->staged.count("cvae3.fusion_z_posterior.0.weight")) { if (pack_cvae3(h)) return 1; ++packed; }
    if (h->staged.count("motion_ae.encoder.net.0.0.weight")) { if (pack_pose_enc(h, "motion_ae", "motion_ae.encoder.", false, &h->motion_ae)) return 1; ++packed; }
    if (h->staged.count("pose_enc.net.0.0.weight")) { if (pack_pose_enc(h, "pose_enc", "pose_enc.", true, &h->pose_enc)) return 1; ++packed; }
    if (h->staged.count("fgd_mlp.Encoder.0.weight")) { if (pack_fgd_mlp(h)) return 1; ++packed; }
    if (h->staged.count("emotion_net.emotion_encoder.conv1.weight")) { if (pack_emotion_net(h)) return 1; ++packed; }
    if (h->staged.count("skel.prior_seq_encoder.fc1.weight")) { if (pack_skeleton(h)) return 1; ++packed; }
    if (!h->staged.count("audio_encoder.feat_extractor.conv1.weight")) {
        if (!packed) EGX_FAIL(h, "no known weight family staged");
        h->staged.clear();
        return 0;
    }
    // drop previously packed weights (keep the log-mel tables: first 6 allocations)
    for (size_t i = 6; i < h->owned.size(); ++i) cudaFree(h->owned[i]);
    h->owned.resize(6);
    h->w = Weights();
    const egx_cfg& c = h->cfg;
    Weights& w = h->w;
    if (!pack_trunk(h, "audio_encoder.feat_extractor", 3, &w.stem, &w.blocks)) return 1;
    {
        const std::string bk = "audio_encoder.final_conv1.bias";
        if (!make_conv(h, "audio_encoder.final_conv1.weight", &bk, "audio_encoder.bn1", 128, c.frames, 3, 1, 0, &w.final_conv))
            return 1;
    }
    const int d = c.d_model, F = c.frames, P = c.pose_dim, p = c.prior_frames;
    if (!make_linear(h, "audio_encoder.fc1", h->H[2] * h->W[2], d, true, &w.a_fc1)) return 1;
    if (!make_linear(h, "audio_encoder.fc2", d, d, true, &w.a_fc2)) return 1;
    // Prior encoder: Prior_ConvEncoder (Full_model/Models.py:186-212) or, when the checkpoint carries them, the
    // Prior_MemoryEncoder keys of Full_model/Models_memory.py:296-345 (same conv-ReLU-BN pair, p -> F-p frames)
    const bool memory = find(h, "prior_seq_encoder.pred_conv.0.weight") != nullptr;
    const std::string pe = "prior_seq_encoder.";
    const std::string k_c1 = memory ? pe + "pred_conv.0" : pe + "conv1", k_b1 = memory ? pe + "pred_conv.2" : pe + "bn1";
    const std::string k_c2 = memory ? pe + "pred_conv.3" : pe + "conv2", k_b2 = memory ? pe + "pred_conv.5" : pe + "bn2";
    const std::string k_f1 = memory ? pe + "post_header.0" : pe + "fc1", k_f2 = memory ? pe + "post_header.2" : pe + "fc2";
    const int Fc = memory ? F - p : F;       // frames the conv pair produces
    {
        const HostTensor *c1w, *c1b, *c2w, *c2b;
        if (!need(h, k_c1 + ".weight", {Fc, p, 3}, &c1w) || !need(h, k_c1 + ".bias", {Fc}, &c1b) ||
            !need(h, k_c2 + ".weight", {Fc, Fc, 3}, &c2w) || !need(h, k_c2 + ".bias", {Fc}, &c2b))
            return 1;
        std::vector<float> s1, t1, s2, t2;
        if (!fold_bn(h, k_b1, Fc, s1, t1) || !fold_bn(h, k_b2, Fc, s2, t2)) return 1;
        w.p_c1w = upload(h, c1w->v); w.p_c1b = upload(h, c1b->v); w.p_s1 = upload(h, s1); w.p_t1 = upload(h, t1);
        w.p_c2w = upload(h, c2w->v); w.p_c2b = upload(h, c2b->v); w.p_s2 = upload(h, s2); w.p_t2 = upload(h, t2);
        {
            const int F4c = (Fc + 3) & ~3;
            std::vector<float> wt((size_t)Fc * 3 * F4c, 0.f);
            for (int f = 0; f < Fc; ++f)
                for (int ci = 0; ci < Fc; ++ci)
                    for (int k = 0; k < 3; ++k) wt[((size_t)ci * 3 + k) * F4c + f] = c2w->v[((size_t)f * Fc + ci) * 3 + k];
            w.p_c2wt = upload(h, wt);
        }
    }
    if (memory) {
        const HostTensor* sw = find(h, pe + "spatial_memory.spatial_chunk_encoder.0.weight");
        if (!sw || sw->shape.size() != 2 || sw->shape[0] != P || sw->shape[1] % P)
            EGX_FAIL(h, "prior_seq_encoder.spatial_memory.spatial_chunk_encoder.0.weight: expected (P, chunk*P)");
        MemPriorW& m = w.mem;
        m.chunk = (int)(sw->shape[1] / P);
        m.n_pred = Fc;
        if (m.chunk < 1 || m.chunk > p || m.chunk > Fc) EGX_FAIL(h, "memory chunk must not exceed prior_frames or frames - prior_frames");
        const int CP = m.chunk * P;
        Affine sp, tm, te;
        if (!linear_chain(h, pe + "spatial_memory.spatial_chunk_encoder", {0, 2}, {CP, P, P}, &sp) ||
            !linear_chain(h, pe + "temporal_memory.temporal_chunk_encoder", {0, 2}, {CP, P, P}, &tm) ||
            !linear_chain(h, pe + "temporal_memory.temporal_memory_encoder", {0, 2}, {CP, m.chunk, m.chunk}, &te))
            return 1;
        std::vector<double> W2 = sp.W, b2 = sp.b;
        W2.insert(W2.end(), tm.W.begin(), tm.W.end());
        b2.insert(b2.end(), tm.b.begin(), tm.b.end());
        m.enc.in = CP; m.enc.out = 2 * P;
        m.enc.w = upload(h, to_f32(W2));
        m.enc.b = upload(h, to_f32(b2));
        m.tm_w = upload(h, to_f32(te.W));
        m.tm_b = upload(h, to_f32(te.b));
        m.on = true;
    }
    if (!make_linear(h, k_f1, P, d, true, &w.p_fc1)) return 1;
    if (!make_linear(h, k_f2, d, d, true, &w.p_fc2)) return 1;
    if (!make_linear(h, "emotion_proj.0", d, d, true, &w.emo0) || !make_linear(h, "emotion_proj.2", d, d, true, &w.emo2) ||
        !make_linear(h, "semantic_proj.0", d, d, true, &w.sem0) || !make_linear(h, "semantic_proj.2", d, d, true, &w.sem2) ||
        !make_linear(h, "fusion_proj.0", d, d, true, &w.fus0) || !make_linear(h, "fusion_proj.2", d, d, true, &w.fus2))
        return 1;
    const int hdr_in[4] = {F * d, d, 256, 64}, hdr_out[4] = {d, 256, 64, 8};
    for (int i = 0; i < 4; ++i)
        if (!make_linear(h, "emotion_classifer_header." + std::to_string(2 * i), hdr_in[i], hdr_out[i], true, &w.hdr[i])) return 1;
    const int post_in[4] = {d, 4 * d, d, P}, post_out[4] = {4 * d, d, P, P};
    for (int i = 0; i < 4; ++i)
        if (!make_linear(h, "post_projector." + std::to_string(2 * i), post_in[i], post_out[i], true, &w.post[i])) return 1;
    if (!make_collapsed(h, {"audio_encoder.fc1", "audio_encoder.fc2"}, {h->H[2] * h->W[2], d, d}, &w.a_fc) ||
        !make_collapsed(h, {k_f1, k_f2}, {P, d, d}, &w.p_fc) ||
        !make_collapsed(h, {"emotion_proj.0", "emotion_proj.2"}, {d, d, d}, &w.emo) ||
        !make_collapsed(h, {"semantic_proj.0", "semantic_proj.2"}, {d, d, d}, &w.sem) ||
        !make_collapsed(h, {"post_projector.0", "post_projector.2", "post_projector.4", "post_projector.6"}, {d, 4 * d, d, P, P},
                        &w.post_all))
        return 1;
    {
        const HostTensor* pt;
        if (!need(h, "encoder.position_enc.pos_table", {1, c.n_position, d}, &pt)) return 1;
        w.pos_table = upload(h, pt->v);
    }
    w.enc_attn.resize(c.n_layers); w.enc_ffn.resize(c.n_layers);
    w.dec_attn.resize(c.n_layers); w.dec_ffn.resize(c.n_layers);
    for (int l = 0; l < c.n_layers; ++l) {
        const std::string e = "encoder.layer_stack." + std::to_string(l), dd = "decoder.layer_stack." + std::to_string(l);
        if (!make_mha(h, e + ".slf_attn", c, &w.enc_attn[l]) || !make_ffn(h, e + ".pos_ffn", c, &w.enc_ffn[l]) ||
            !make_mha(h, dd + ".enc_attn", c, &w.dec_attn[l]) || !make_ffn(h, dd + ".pos_ffn", c, &w.dec_ffn[l]))
            return 1;
    }
    for (void* ptr : h->owned)
        if (!ptr) EGX_FAIL(h, "device allocation failed while packing weights");
    h->staged.clear();
    h->finalized = true;
    return 0;
}

static int logmel_checked(egx_handle* h, const float* audio, int n_clips, int n_samples, int n_cols, int mode, int preemph,
                          float* out, void* stream, bool force_global_tile) {
    if (!h) return 1;
    if (n_clips <= 0) return 0;
    if (n_cols < 1 || n_cols > 1 + n_samples / 512) EGX_FAIL(h, "n_cols out of range for n_samples");
    if (n_cols > 256) EGX_FAIL(h, "n_cols > 256 not supported (shared-memory tile)");
    if (n_samples < 2) EGX_FAIL(h, "need at least 2 samples");
    if (mode != EGX_LOGMEL_DB && mode != EGX_LOGMEL_LOG_IN && mode != (EGX_LOGMEL_DB | EGX_LOGMEL_FP16_STORAGE))
        EGX_FAIL(h, "unknown log-mel mode");
    cudaStream_t s = (cudaStream_t)stream;
    StageScope sc(h, 1);
    LAUNCH(h, launch_logmel(h->lm, audio, n_clips, n_samples, n_cols, mode, preemph, out, s, force_global_tile));
    return 0;
}

int egx_logmel(egx_handle* h, const float* audio, int n_clips, int n_samples, int n_cols, int mode, int preemph,
               float* out, void* stream) {
    return logmel_checked(h, audio, n_clips, n_samples, n_cols, mode, preemph, out, stream, false);
}

// Test hook: the same front end through the kernel that keeps the (128, n_cols) tile in global memory (the path of
// spectrograms too wide for two shared-memory tiles per SM), whatever the width: the two kernels must agree bit for bit.
int egx_debug_logmel_global_tile(egx_handle* h, const float* audio, int n_clips, int n_samples, int n_cols, int mode,
                                 int preemph, float* out, void* stream) {
    return logmel_checked(h, audio, n_clips, n_samples, n_cols, mode, preemph, out, stream, true);
}

int egx_audio_pcm16_to_f32(egx_handle* h, const int16_t* pcm, int64_t n_samples, float* out, void* stream) {
    if (!h) return 1;
    if (n_samples <= 0) return 0;
    if (!pcm || !out) EGX_FAIL(h, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    StageScope sc(h, 1);
    LAUNCH(h, launch_pcm16_to_f32(pcm, n_samples, out, s));
    return 0;
}

int egx_audio_fixed_length(egx_handle* h, const float* samples, const int64_t* offsets, int n_clips, int n_out,
                           float* out, void* stream) {
    if (!h) return 1;
    if (n_clips <= 0) return 0;
    if (!samples || !offsets || !out) EGX_FAIL(h, "null pointer argument");
    if (n_out < 1) EGX_FAIL(h, "n_out must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    StageScope sc(h, 1);
    LAUNCH(h, launch_fixed_length(samples, offsets, n_clips, n_out, out, s));
    return 0;
}

size_t egx_workspace_bytes(const egx_handle* h, int n_clips) {
    if (!h || n_clips <= 0) return 0;
    Plan p;
    if (h->cfg.precision == EGX_PREC_FP32) plan_slots<float>(h, n_clips, p);
    else plan_tc(h, n_clips, p);
    return p.off + 256;
}

int egx_generator_forward(egx_handle* h, const float* spec, const float* prior, const float* sampled_emotion,
                          int n_clips, float* poses, float* emo_feat, float* sem_feat, float* emo_logits,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return 1;
    if (!h->finalized) EGX_FAIL(h, "weights not finalized");
    if (n_clips <= 0) return 0;
    if (!spec || !prior || !poses || !emo_feat || !sem_feat || !emo_logits || !workspace) EGX_FAIL(h, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cfg.precision == EGX_PREC_FP32)
        return forward_impl<float>(h, spec, prior, sampled_emotion, n_clips, poses, emo_feat, sem_feat, emo_logits,
                                   workspace, workspace_bytes, s);
    return forward_tc(h, spec, prior, sampled_emotion, n_clips, poses, emo_feat, sem_feat, emo_logits, workspace,
                      workspace_bytes, s);
}

int egx_get_tap(egx_handle* h, const char* name, const void* workspace, int n_clips, float* out, size_t out_capacity,
                size_t* n_out, void* stream) {
    if (!h || !name || !workspace || !out || !n_out) return 1;
    if (h->cfg.precision == EGX_PREC_FP32)
        return get_tap_impl<float>(h, name, workspace, n_clips, out, out_capacity, n_out, (cudaStream_t)stream);
    return get_tap_tc(h, name, workspace, n_clips, out, out_capacity, n_out, (cudaStream_t)stream);
}

int egx_debug_trunk(egx_handle* h, const float* spec, int n_clips, int stage, float* out, size_t out_capacity,
                    size_t* n_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !spec || !out || !n_out || !workspace) return 1;
    if (!h->finalized) EGX_FAIL(h, "weights not finalized");
    if (stage < 0 || stage > 3) EGX_FAIL(h, "stage must be 0..3");
    if (h->cfg.precision == EGX_PREC_FP32)
        return debug_trunk_impl<float>(h, spec, n_clips, stage, out, out_capacity, n_out, workspace, workspace_bytes,
                                       (cudaStream_t)stream);
    return debug_trunk_tc(h, spec, n_clips, stage, out, out_capacity, n_out, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int egx_debug_linear_tc(egx_handle* h, const float* A, const float* W, const float* bias, const float* addend,
                        int addend_rows, int M, int N, int K, int relu, float* out32, void* stream) {
    if (!h || !A || !W || !out32) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    const int ldk = (K + 7) / 8 * 8;
    __half *a16 = nullptr, *w16 = nullptr, *o16 = nullptr;
    EGX_CHECK_CUDA(h, cudaMalloc(&a16, (size_t)M * ldk * 2));
    EGX_CHECK_CUDA(h, cudaMalloc(&w16, (size_t)N * ldk * 2));
    EGX_CHECK_CUDA(h, cudaMalloc(&o16, (size_t)M * N * 2));
    LAUNCH(h, launch_cvt_pad_f16(A, M, K, K, a16, ldk, s));
    LAUNCH(h, launch_cvt_pad_f16(W, N, K, K, w16, ldk, s));
    GemmEpi e;
    e.bias = bias; e.relu = relu; e.addend = addend; e.addend_rows = addend_rows; e.addend_ld = N;
    LAUNCH(h, launch_gemm_tc(a16, ldk, w16, ldk, M, N, K, e, out32, N, o16, N, s));
    EGX_CHECK_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(a16); cudaFree(w16); cudaFree(o16);
    return 0;
}

int egx_debug_linear_ln_tc(egx_handle* h, const float* A, const float* W, const float* bias, const float* residual,
                           const float* ln_g, const float* ln_b, int M, int K, float* out32, void* out16, void* stream) {
    if (!h || !A || !W || !ln_g || !ln_b || !out32) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    const int N = 256, ldk = (K + 7) / 8 * 8;
    __half *a16 = nullptr, *w16 = nullptr;
    EGX_CHECK_CUDA(h, cudaMalloc(&a16, (size_t)M * ldk * 2));
    EGX_CHECK_CUDA(h, cudaMalloc(&w16, (size_t)N * ldk * 2));
    LAUNCH(h, launch_cvt_pad_f16(A, M, K, K, a16, ldk, s));
    LAUNCH(h, launch_cvt_pad_f16(W, N, K, K, w16, ldk, s));
    GemmEpi e;
    e.bias = bias; e.addend = residual; e.addend_ld = N; e.ln_g = ln_g; e.ln_b = ln_b;
    LAUNCH(h, launch_gemm_tc(a16, ldk, w16, ldk, M, N, K, e, out32, N, static_cast<__half*>(out16), N, s));
    EGX_CHECK_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(a16); cudaFree(w16);
    return 0;
}

int egx_debug_ffn_tc(egx_handle* h, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* ln_g, const float* ln_b, int M, int d_inner, float* out32, void* out16, void* stream) {
    if (!h || !x || !w1 || !b1 || !w2 || !b2 || !ln_g || !ln_b || !out32 || !out16) return 1;
    if (!ffn_tc_supported(256, d_inner)) EGX_FAIL(h, "fused FFN needs d_model = 256 and d_inner a multiple of 256");
    cudaStream_t s = (cudaStream_t)stream;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    __half *x16 = nullptr, *w1h = nullptr, *w2h = nullptr;
    EGX_CHECK_CUDA(h, cudaMalloc(&x16, (size_t)M * 256 * 2));
    EGX_CHECK_CUDA(h, cudaMalloc(&w1h, (size_t)d_inner * 256 * 2));
    EGX_CHECK_CUDA(h, cudaMalloc(&w2h, (size_t)256 * d_inner * 2));
    LAUNCH(h, launch_cvt_pad_f16(x, M, 256, 256, x16, 256, s));
    LAUNCH(h, launch_cvt_pad_f16(w1, d_inner, 256, 256, w1h, 256, s));
    LAUNCH(h, launch_cvt_pad_f16(w2, 256, d_inner, d_inner, w2h, d_inner, s));
    std::vector<float> hb1(d_inner), hb2(256), hg(256), hb(256);
    EGX_CHECK_CUDA(h, cudaMemcpy(hb1.data(), b1, sizeof(float) * d_inner, cudaMemcpyDeviceToHost));
    EGX_CHECK_CUDA(h, cudaMemcpy(hb2.data(), b2, sizeof(float) * 256, cudaMemcpyDeviceToHost));
    EGX_CHECK_CUDA(h, cudaMemcpy(hg.data(), ln_g, sizeof(float) * 256, cudaMemcpyDeviceToHost));
    EGX_CHECK_CUDA(h, cudaMemcpy(hb.data(), ln_b, sizeof(float) * 256, cudaMemcpyDeviceToHost));
    LAUNCH(h, launch_ffn_tc(x16, x, w1h, 256, hb1.data(), w2h, d_inner, hb2.data(), hg.data(), hb.data(), M, d_inner, out32,
                            static_cast<__half*>(out16), s));
    EGX_CHECK_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(x16); cudaFree(w1h); cudaFree(w2h);
    return 0;
}

int egx_debug_conv_tc(egx_handle* h, const void* in16, int B, int H, int W, int cin, const void* w16, int cout,
                      int ks, int stride, int relu_first, const float* bias, const float* scale,
                      const float* shift, void* out16, int nchw, float* se_part, void* stream) {
    if (!h || !in16 || !w16 || !scale || !shift || !out16) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    ConvW c;
    c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride; c.relu_first = relu_first;
    c.w16 = const_cast<__half*>(static_cast<const __half*>(w16));
    c.bias = const_cast<float*>(bias); c.scale = const_cast<float*>(scale); c.shift = const_cast<float*>(shift);
    LAUNCH(h, launch_conv_tc(c, static_cast<const __half*>(in16), B, H, W, static_cast<__half*>(out16), nchw, se_part, s));
    return 0;
}

int egx_debug_attention_tc(egx_handle* h, const void* q16, int ldq, int q_col0, const void* kv16, int ldkv, int k_col0,
                           int v_col0, int B, int L, int n_head, void* out16, int ldo, void* stream) {
    if (!h || !q16 || !kv16 || !out16) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_attention_tc(static_cast<const __half*>(q16), ldq, q_col0, static_cast<const __half*>(kv16), ldkv,
                                  k_col0, v_col0, B, L, n_head, static_cast<__half*>(out16), ldo, s));
    return 0;
}

int egx_profile_enable(egx_handle* h, int max_launches) {
    if (!h) return 1;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    h->profiling = max_launches > 0;
    h->prof_used = 0;
    h->prof_stage.clear();
    while (h->profiling && h->prof_events.size() < (size_t)max_launches * 2) {
        cudaEvent_t e;
        EGX_CHECK_CUDA(h, cudaEventCreate(&e));
        h->prof_events.push_back(e);
    }
    return 0;
}

int egx_profile_read(egx_handle* h, double* ms_per_stage, int64_t* launches_per_stage, int n_stages) {
    if (!h || !ms_per_stage || !launches_per_stage) return 1;
    for (int i = 0; i < n_stages; ++i) { ms_per_stage[i] = 0.0; launches_per_stage[i] = 0; }
    for (size_t i = 0; i < h->prof_stage.size(); ++i) {
        cudaEvent_t e0 = h->prof_events[2 * i], e1 = h->prof_events[2 * i + 1];
        EGX_CHECK_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        EGX_CHECK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const int st = h->prof_stage[i];
        if (st >= 0 && st < n_stages) { ms_per_stage[st] += ms; launches_per_stage[st] += 1; }
    }
    h->prof_used = 0;
    h->prof_stage.clear();
    return 0;
}

int egx_fgd_accumulate(egx_handle* h, const float* feats, int64_t n_rows, int dim, const double* shift, double* acc,
                       void* stream) {
    if (!h) return 1;
    if (n_rows == 0) return 0;      // empty shard: nothing to add
    if (!feats || !acc) EGX_FAIL(h, "null pointer argument");
    if (dim <= 0 || n_rows < 0) EGX_FAIL(h, "dim must be positive and n_rows non-negative");
    cudaStream_t s = (cudaStream_t)stream;
    EGX_CHECK_CUDA(h, cudaSetDevice(h->device));
    int n_split = 1;
    const size_t need = fgd_scratch_doubles(n_rows, dim, h->sms, &n_split);
    if (need > h->fgd_scratch_n) {
        // first use (or a larger problem): cudaFree waits for work still reading the old scratch
        if (h->fgd_scratch) cudaFree(h->fgd_scratch);
        h->fgd_scratch = nullptr; h->fgd_scratch_n = 0;
        EGX_CHECK_CUDA(h, cudaMalloc(&h->fgd_scratch, need * sizeof(double)));
        h->fgd_scratch_n = need;
    }
    StageScope sc(h, 8);
    LAUNCH(h, launch_fgd_accumulate(feats, n_rows, dim, shift, acc, h->fgd_scratch, n_split, s));
    return 0;
}


int egx_cvae_forward(egx_handle* h, const float* x, const float* y, const float* eps, int64_t n, float* out, float* mu,
                     float* logvar, void* stream) {
    if (!h) return 1;
    if (!h->cvae.ready) EGX_FAIL(h, "cvae weights not loaded");
    if (n == 0) return 0;
    if (n < 0 || !x || !y || !eps || !out || !mu || !logvar) EGX_FAIL(h, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_cvae_mlp(h->cvae, x, y, eps, 0, n, out, mu, logvar, s));
    return 0;
}

int egx_cvae_sample(egx_handle* h, const float* y, const float* z, int64_t n, float* out, void* stream) {
    if (!h) return 1;
    if (!h->cvae.ready) EGX_FAIL(h, "cvae weights not loaded");
    if (n == 0) return 0;
    if (n < 0 || !y || !z || !out) EGX_FAIL(h, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_cvae_mlp(h->cvae, nullptr, y, z, 1, n, out, nullptr, nullptr, s));
    return 0;
}

int egx_cvae3_sample(egx_handle* h, const float* y, const float* z, int n, float* out, void* stream) {
    if (!h) return 1;
    if (!h->cvae3.ready) EGX_FAIL(h, "cvae3 weights not loaded");
    if (n == 0) return 0;
    if (n < 0 || !y || !z || !out) EGX_FAIL(h, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_cvae3_sample(h->cvae3, y, z, n, out, s));
    return 0;
}

int egx_pose_features(egx_handle* h, int kind, const float* poses, int n_clips, int n_frames, int pose_dim, float* out,
                      void* stream) {
    if (!h) return 1;
    if (kind != EGX_POSE_MOTION_AE && kind != EGX_POSE_EMBEDDING_NET) EGX_FAIL(h, "unknown pose feature net");
    const PoseEncW& w = kind == EGX_POSE_MOTION_AE ? h->motion_ae : h->pose_enc;
    if (!w.ready) EGX_FAIL(h, "pose feature net weights not loaded");
    if (n_clips == 0) return 0;
    if (n_clips < 0 || !poses || !out) EGX_FAIL(h, "null pointer argument");
    if (n_frames != w.L || pose_dim != w.P)
        EGX_FAIL(h, "pose clip shape (" + std::to_string(n_frames) + "," + std::to_string(pose_dim) + ") does not match the loaded net (" +
                        std::to_string(w.L) + "," + std::to_string(w.P) + ")");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_pose_encoder(w, poses, n_clips, out, s));
    return 0;
}

int egx_pose_feature_dim(const egx_handle* h, int kind) {
    if (!h) return 0;
    const PoseEncW& w = kind == EGX_POSE_MOTION_AE ? h->motion_ae : h->pose_enc;
    return w.ready ? w.n_out : 0;
}

int egx_beat_align(egx_handle* h, const float* poses, int n_clips, int n_frames, int pose_dim, int frame_lo, int frame_hi,
                   int order, double sigma, double pose_fps, const double* onset_times, const int32_t* onset_offsets,
                   double* scores, unsigned char* beat_mask, void* stream) {
    if (!h) return 1;
    if (n_clips == 0) return 0;
    if (n_clips < 0 || !poses || !onset_times || !onset_offsets || !scores) EGX_FAIL(h, "null pointer argument");
    if (n_frames < 2 || n_frames > 64) EGX_FAIL(h, "beat alignment supports 2..64 frames per clip");
    if (pose_dim < 174) EGX_FAIL(h, "beat alignment reads pose columns 18:42 and 150:174 (BEAT skeleton); pose_dim is too small");
    if (order < 1 || !(sigma > 0.0) || !(pose_fps > 0.0)) EGX_FAIL(h, "order >= 1, sigma > 0 and pose_fps > 0 required");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(h, launch_beat_align(poses, n_clips, n_frames, pose_dim, frame_lo, frame_hi, order, sigma, pose_fps, onset_times,
                                onset_offsets, scores, beat_mask, s));
    return 0;
}

size_t egx_row_features_workspace(const egx_handle* h, int64_t n_rows) {
    if (!h || !h->fgd_mlp.ready || n_rows <= 0) return 0;
    return (size_t)n_rows * h->fgd_mlp.lin.ldw * sizeof(__half) + 256;
}

int egx_row_features(egx_handle* h, const float* rows, int64_t n_rows, int dim, float* out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (!h) return 1;
    if (!h->fgd_mlp.ready) EGX_FAIL(h, "fgd_mlp weights not loaded");
    if (n_rows == 0) return 0;
    const LinearW& l = h->fgd_mlp.lin;
    if (n_rows < 0 || !rows || !out || !workspace) EGX_FAIL(h, "null pointer argument");
    if (dim != l.in) EGX_FAIL(h, "row width does not match the loaded net");
    if (n_rows > (int64_t)0x7fffffff) EGX_FAIL(h, "too many rows for one call");
    if (workspace_bytes < egx_row_features_workspace(h, n_rows)) EGX_FAIL(h, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    __half* a16 = static_cast<__half*>(workspace);
    LAUNCH(h, launch_cvt_pad_f16(rows, n_rows, l.in, l.in, a16, l.ldw, s));
    GemmEpi e;
    e.bias = l.b;
    LAUNCH(h, launch_gemm_tc(a16, l.ldw, l.w16, l.ldw, (int)n_rows, l.out, l.in, e, out, l.out, nullptr, 0, s));
    return 0;
}


size_t egx_emotion_net_workspace(const egx_handle* h, int n_clips, int n_mels, int n_cols) {
    if (!h || n_clips <= 0) return 0;
    Plan p;
    plan_emotion(n_clips, n_mels, n_cols, p);
    return p.off + 256;
}

int egx_emotion_net_forward(egx_handle* h, const float* spec, int n_clips, int n_mels, int n_cols, float* logits,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return 1;
    if (!h->emo.ready) EGX_FAIL(h, "emotion_net weights not loaded");
    if (n_clips == 0) return 0;
    if (n_clips < 0 || !spec || !logits || !workspace) EGX_FAIL(h, "null pointer argument");
    const EmotionNetW& e = h->emo;
    int Hc = n_mels, Wc = n_cols;
    for (int i = 0; i < 3; ++i) { Hc = (Hc + 1) / 2; Wc = (Wc + 1) / 2; }
    if (n_cols % 2 || Hc * Wc * 256 != e.flat)
        EGX_FAIL(h, "spectrogram (" + std::to_string(n_mels) + "," + std::to_string(n_cols) + ") does not flatten to the " +
                        std::to_string(e.flat) + " inputs of emotion_eocder_fc.0");
    cudaStream_t s = (cudaStream_t)stream;
    Plan p;
    p.base = static_cast<char*>(workspace);
    EmoSlots sl = plan_emotion(n_clips, n_mels, n_cols, p);
    if (p.off > workspace_bytes) EGX_FAIL(h, "workspace too small: need " + std::to_string(p.off));
    __half* feat = nullptr;
    int Ho, Wo;
    if (run_trunk_tc(h, e.stem, e.blocks, 4, n_mels, n_cols, spec, n_clips, sl.tb, 4, &feat, &Ho, &Wo, s)) return 1;
    StageScope sc(h, 5);
    const __half* a = feat;
    for (int i = 0; i < 6; ++i) {
        const LinearW& l = e.fc[i];
        if (linear_tc(h, l, a, l.in, n_clips, i == 5 ? logits : nullptr, 8, i < 5 ? sl.fc[i] : nullptr, l.out, i < 5, nullptr, 0, s))
            return 1;
        if (i < 5) a = sl.fc[i];
    }
    return 0;
}

size_t egx_skeleton_workspace(const egx_handle* h, int n_clips) {
    if (!h || !h->skel.ready || n_clips <= 0) return 0;
    Plan p;
    plan_skeleton(h->skel, n_clips, p);
    return p.off + 256;
}

int egx_skeleton_dims(const egx_handle* h, int* n_frames, int* pose_dim, int* d_model, int* n_class) {
    if (!h || !h->skel.ready) return 1;
    if (n_frames) *n_frames = h->skel.T;
    if (pose_dim) *pose_dim = h->skel.P;
    if (d_model) *d_model = h->skel.d;
    if (n_class) *n_class = h->skel.n_class;
    return 0;
}

int egx_skeleton_forward(egx_handle* h, const float* poses, int n_clips, int n_frames, int pose_dim, float* logits,
                         float* mid_feature, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return 1;
    const SkeletonW& k = h->skel;
    if (!k.ready) EGX_FAIL(h, "skeleton classifier weights not loaded");
    if (n_clips == 0) return 0;
    if (n_clips < 0 || !poses || !logits || !workspace) EGX_FAIL(h, "null pointer argument");
    if (n_frames != k.T || pose_dim != k.P)
        EGX_FAIL(h, "poses (n," + std::to_string(n_frames) + "," + std::to_string(pose_dim) + ") do not match the classifier's (" +
                        std::to_string(k.T) + "," + std::to_string(k.P) + ") input (post_projector flattens n_position frames)");
    cudaStream_t s = (cudaStream_t)stream;
    Plan p;
    p.base = static_cast<char*>(workspace);
    SkelSlots sl = plan_skeleton(k, n_clips, p);
    if (p.off > workspace_bytes) EGX_FAIL(h, "workspace too small: need " + std::to_string(p.off));
    const int B = n_clips, R = B * k.T, d = k.d, hk = k.n_head * 64;
    StageScope sc(h, 6);
    LAUNCH(h, launch_cvt_pad_f16(poses, R, k.P, k.P, sl.a16, sl.P8, s));
    // Prior_Encoder, then PositionalEncoding.forward: x + pos_table[:, :T] (skeleton_classifer/Models.py:49-50,109-113)
    if (linear_tc(h, k.prior, sl.a16, sl.P8, R, sl.x32a, d, sl.x16, d, 0, k.pos_table, k.T, s)) return 1;
    float *x32 = sl.x32a, *y32 = sl.x32b;
    for (int l = 0; l < k.n_layers; ++l) {
        const MHAW& a = k.attn[l];
        const FFNW& f = k.ffn[l];
        if (linear_tc(h, a.qkv, sl.x16, d, R, nullptr, 0, sl.qkv16, 3 * hk, 0, nullptr, 0, s)) return 1;
        LAUNCH(h, launch_attention_tc(sl.qkv16, 3 * hk, 0, sl.qkv16, 3 * hk, hk, 2 * hk, B, k.T, k.n_head, sl.o16, hk, s));
        if (linear_ln_tc(h, a.fc, a.ln, sl.o16, hk, R, x32, sl.pre, y32, sl.x1_16, s)) return 1;
        if (ffn_block_tc(h, f, sl.x1_16, y32, R, sl.hid16, sl.pre, x32, sl.x16, s)) return 1;
    }
    if (mid_feature) EGX_CHECK_CUDA(h, cudaMemcpyAsync(mid_feature, x32, (size_t)R * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    // enc_output.reshape(B, -1) -> post_projector (:277-281): rows of x16 are already (clip, frame)-major
    const __half* a = sl.x16;
    int lda = k.T * d;
    for (int i = 0; i < 5; ++i) {
        const LinearW& l = k.post[i];
        if (linear_tc(h, l, a, lda, B, i == 4 ? logits : nullptr, l.out, i < 4 ? sl.p16[i] : nullptr, l.out, i < 4, nullptr, 0, s))
            return 1;
        if (i < 4) { a = sl.p16[i]; lda = l.out; }
    }
    return 0;
}

}  // extern "C"
