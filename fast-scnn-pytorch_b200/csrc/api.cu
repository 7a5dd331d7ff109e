// api.cu -- the C ABI of libfscnn_b200.so (see include/fscnn_b200.h): context, state_dict manifest,
// BN folding / weight packing plan, workspace plan, and the stage sequencing of the forward path
// (reference models/fast_scnn.py:33-46).  Host-only logic; every device operation is a kernel from
// the sibling .cu files enqueued on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fscnn_b200.h"
#include "kernels.h"

using namespace fscnn;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

struct Param { std::string name; int64_t numel; };

struct BneckPlan { const char* name; int cin, cout, stride; };
const BneckPlan kBnecks[9] = {   // GlobalFeatureExtractor._make_layer, models/fast_scnn.py:170-180
    {"bottleneck1.0", 64, 64, 2}, {"bottleneck1.1", 64, 64, 1}, {"bottleneck1.2", 64, 64, 1},
    {"bottleneck2.0", 64, 96, 2}, {"bottleneck2.1", 96, 96, 1}, {"bottleneck2.2", 96, 96, 1},
    {"bottleneck3.0", 96, 128, 1}, {"bottleneck3.1", 128, 128, 1}, {"bottleneck3.2", 128, 128, 1},
};

enum Stage { kStem = 0, kDs1, kDs2, kB0, kPpm = kB0 + 9, kFfm, kCls1, kCls2Head, kAux, kNumStages };
const char* kStageNames[kNumStages] = {
    "stem", "l2d.dsconv1", "l2d.dsconv2", "gfe.bottleneck1.0", "gfe.bottleneck1.1", "gfe.bottleneck1.2",
    "gfe.bottleneck2.0", "gfe.bottleneck2.1", "gfe.bottleneck2.2", "gfe.bottleneck3.0", "gfe.bottleneck3.1",
    "gfe.bottleneck3.2", "gfe.ppm", "ffm", "cls.dsconv1", "cls.dsconv2+head", "aux"};

// Spatial sizes of every stage for an h x w input.
struct Dims {
    int h, w;      // input
    int h1, w1;    // stem: 3x3 s2 p0
    int h2, w2;    // dsconv1: s2 p1
    int h3, w3;    // dsconv2: s2 p1 (higher_res, /8)
    int h4, w4;    // bottleneck1 (/16)
    int h5, w5;    // bottleneck2,3 / ppm (/32)
    bool ok;
};

Dims make_dims(int h, int w) {
    Dims d{};
    d.h = h; d.w = w;
    d.h1 = (h - 3) / 2 + 1; d.w1 = (w - 3) / 2 + 1;
    d.h2 = (d.h1 - 1) / 2 + 1; d.w2 = (d.w1 - 1) / 2 + 1;
    d.h3 = (d.h2 - 1) / 2 + 1; d.w3 = (d.w2 - 1) / 2 + 1;
    d.h4 = (d.h3 - 1) / 2 + 1; d.w4 = (d.w3 - 1) / 2 + 1;
    d.h5 = (d.h4 - 1) / 2 + 1; d.w5 = (d.w4 - 1) / 2 + 1;
    d.ok = h >= 3 && w >= 3 && d.h5 >= 1 && d.w5 >= 1;
    return d;
}

// Workspace layout for a micro-batch of `mb` images: byte offsets of every stage tensor.
struct WsPlan {
    size_t stem, ds1, higher, b[9], ppm, rowsum, z, z16, ppm_r, ffm, cls1, logits, aux_logits, total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

struct fscnn_ctx {
    int nc = 0, ncp = 0, aux = 0, prec = 0;
    std::vector<Param> params;
    size_t packed_floats = 0;
    bool loaded = false;
    int64_t launches = 0;
    int micro_batch = 0;
    bool in_dirty = false; // uint8 input: the normalisation-folded stem image must be rebuilt before the next launch
    int fuse_front = 1;   // bf16: stem + dsconv1 in one kernel whenever both stages are requested
    StemIn in{FSCNN_IN_F32_NCHW, {0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
    // offsets (floats) into the packed buffer, fixed at create time
    struct Off {
        size_t stem_w, stem_b;
        size_t ds_wd[4], ds_bd[4], ds_wp[4], ds_bp[4];     // l2d.dsconv1, l2d.dsconv2, cls.dsconv1, cls.dsconv2
        size_t bn_we[9], bn_be[9], bn_wd[9], bn_bd[9], bn_wp[9], bn_bp[9];
        size_t ppm_wc[4], ppm_bc[4], ppm_wo, ppm_bo;
        size_t ffm_wd, ffm_bd, ffm_wcat, ffm_bcat;
        size_t head_w, head_b;
        size_t aux_w, aux_b, auxh_w, auxh_b;
        size_t bn_weT_img[9], bn_wpT_img[9], bn_tabT_img[9]; // bf16 tcgen05 operand images + constant tables of the transposed-expand kernels (offsets in floats)
        size_t ppm_wx_img;   // PPM out conv, x part + branch part as [128 x 128] operand images
        size_t ffm_tabT;   // FFM, transposed kernel: per-channel depthwise records + fused bias
        size_t ds_wp_img[4], head_img, ffm_img, stem_img, stem_img_u8, stem_b_u8, stem_imgx, stem_imgx_u8;
    } off{};
    // device pointers resolved by load_weights
    StemW stem{};
    DsW ds[4]{};
    BneckW bn[9]{};
    const bf16* bn_weT_img[9]{};
    const bf16* bn_wpT_img[9]{};
    const unsigned char* bn_tabT_img[9]{};
    const bf16* ds_wp_img[4]{};
    const bf16* head_img = nullptr;
    const bf16* ffm_img = nullptr;
    const unsigned char* ffm_tabT = nullptr;
    const bf16* ppm_wx_img = nullptr;
    int front_transposed = 1;  // bf16 fused front kernel with the transposed stem (0 = l2d_front_tc.cu)
    const bf16* stem_img = nullptr;
    bf16* stem_img_u8 = nullptr;
    float* stem_b_u8 = nullptr;
    bf16* stem_imgx = nullptr;      // fused front kernel: RGBX images (fp32 input / uint8 input with folded normalisation)
    bf16* stem_imgx_u8 = nullptr;
    PpmW ppm{};
    FfmW ffm{};
    HeadW head{};
    AuxW auxw{};

    int esize() const { return prec == FSCNN_PREC_BF16 ? 2 : 4; }
    int eff_mb(int n, int h, int w) const {
        int mb = micro_batch;
        if (mb <= 0) {   // default: about 192 Mpixel of input per micro-batch (96 images at 1024x2048): the small late
                         // stages then launch many full waves and the per-launch prologues / tails of the 19 persistent
                         // kernels amortise (measured: 8 -> 32 images +12 %, 37 -> 111 images +5 %)
            long long px = (long long)h * w;
            mb = (int)((192ll << 20) / (px > 0 ? px : 1));
            if (mb < 1) mb = 1;
            if (mb > 128) mb = 128;
            // ... nudged (within +-25 %) to the count whose 8x16-pixel tiles at the coarsest level (/32: seven of the nine
            // bottlenecks + PPM, a handful of tiles per image) fill whole waves of persistent CTAs: 37 images at 1024x2048
            // (16 tiles each = 4 x 148) and its multiples instead of 32 (3.46 waves, the last one 46 % full)
            if (n > mb * 3 / 4) {
                const Dims d = make_dims(h, w);
                const int sms = num_sms();
                if (d.ok && sms > 0) {
                    const long long t5 = (long long)((d.h5 + 7) / 8) * ((d.w5 + 15) / 16);
                    int best = mb;
                    double best_eff = 0.0;
                    const int hi = (mb * 5 / 4 < 128 ? mb * 5 / 4 : 128) < n ? (mb * 5 / 4 < 128 ? mb * 5 / 4 : 128) : n;
                    for (int c = mb * 3 / 4 > 1 ? mb * 3 / 4 : 1; c <= hi; ++c) {
                        const long long tiles = t5 * c, waves = (tiles + sms - 1) / sms;
                        const double eff = (double)tiles / (double)(waves * sms);
                        if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && abs(c - mb) < abs(best - mb))) { best_eff = eff; best = c; }
                    }
                    mb = best;
                }
            }
        }
        return mb < n ? mb : n;
    }
    WsPlan plan(int mb, const Dims& d) const {
        WsPlan p{};
        size_t o = 0;
        const size_t es = esize();
        auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
        p.stem = take((size_t)mb * d.h1 * d.w1 * 32 * es);
        p.ds1 = take((size_t)mb * d.h2 * d.w2 * 48 * es);
        p.higher = take((size_t)mb * d.h3 * d.w3 * 64 * es);
        for (int i = 0; i < 9; ++i) {
            const bool lvl4 = i < 3;
            p.b[i] = take((size_t)mb * (lvl4 ? d.h4 * d.w4 : d.h5 * d.w5) * kBnecks[i].cout * es);
        }
        p.ppm = take((size_t)mb * d.h5 * d.w5 * 128 * es);
        p.rowsum = take((size_t)mb * d.h5 * 12 * 128 * 4);
        p.z = take((size_t)mb * 50 * 128 * 4);
        p.z16 = take((size_t)mb * 128 * 64 * 2);
        p.ppm_r = take((size_t)d.h5 * d.w5 * 64 * 2);
        p.ffm = take((size_t)mb * d.h3 * d.w3 * 128 * es);
        p.cls1 = take((size_t)mb * d.h3 * d.w3 * 128 * es);
        p.logits = take((size_t)mb * d.h3 * d.w3 * ncp * 4);
        p.aux_logits = aux ? take((size_t)mb * d.h3 * d.w3 * ncp * 4) : 0;
        p.total = o;
        return p;
    }
};

namespace {

void add_conv(fscnn_ctx* c, const std::string& name, int64_t wn, int64_t bias_n = 0) {
    c->params.push_back({name + ".weight", wn});
    if (bias_n) c->params.push_back({name + ".bias", bias_n});
}
void add_bn(fscnn_ctx* c, const std::string& name, int64_t ch) {
    for (const char* k : {".weight", ".bias", ".running_mean", ".running_var"}) c->params.push_back({name + k, ch});
}

void build_manifest_and_offsets(fscnn_ctx* c) {
    size_t o = 0;
    auto take = [&](size_t floats) { size_t at = o; o += (floats + 63) & ~(size_t)63; return at; };   // 256-byte granules
    auto& f = c->off;
    // LearningToDownsample (models/fast_scnn.py:148-161)
    add_conv(c, "learning_to_downsample.conv.conv.0", 32 * 27);
    add_bn(c, "learning_to_downsample.conv.conv.1", 32);
    f.stem_w = take(27 * 32); f.stem_b = take(32);
    const struct { const char* p; int cin, cout; } dss[4] = {{"learning_to_downsample.dsconv1", 32, 48},
                                                              {"learning_to_downsample.dsconv2", 48, 64},
                                                              {"classifier.dsconv1", 128, 128},
                                                              {"classifier.dsconv2", 128, 128}};
    auto add_ds = [&](int i) {
        const std::string p = dss[i].p;
        add_conv(c, p + ".conv.0", dss[i].cin * 9); add_bn(c, p + ".conv.1", dss[i].cin);
        add_conv(c, p + ".conv.3", (int64_t)dss[i].cout * dss[i].cin); add_bn(c, p + ".conv.4", dss[i].cout);
        f.ds_wd[i] = take(9 * dss[i].cin); f.ds_bd[i] = take(dss[i].cin);
        f.ds_wp[i] = take((size_t)dss[i].cin * dss[i].cout); f.ds_bp[i] = take(dss[i].cout);
    };
    add_ds(0); add_ds(1);
    // GlobalFeatureExtractor (:164-187)
    for (int i = 0; i < 9; ++i) {
        const std::string p = std::string("global_feature_extractor.") + kBnecks[i].name + ".block";
        const int ci = kBnecks[i].cin, ce = 6 * ci, co = kBnecks[i].cout;
        add_conv(c, p + ".0.conv.0", (int64_t)ce * ci); add_bn(c, p + ".0.conv.1", ce);
        add_conv(c, p + ".1.conv.0", ce * 9); add_bn(c, p + ".1.conv.1", ce);
        add_conv(c, p + ".2", (int64_t)co * ce); add_bn(c, p + ".3", co);
        f.bn_we[i] = take((size_t)ci * ce); f.bn_be[i] = take(ce);
        f.bn_wd[i] = take(9 * ce); f.bn_bd[i] = take(ce);
        f.bn_wp[i] = take((size_t)ce * co); f.bn_bp[i] = take(co);
    }
    for (int i = 0; i < 4; ++i) {
        const std::string p = "global_feature_extractor.ppm.conv" + std::to_string(i + 1) + ".conv";
        add_conv(c, p + ".0", 32 * 128); add_bn(c, p + ".1", 32);
        f.ppm_wc[i] = take(128 * 32); f.ppm_bc[i] = take(32);
    }
    add_conv(c, "global_feature_extractor.ppm.out.conv.0", 128 * 256);
    add_bn(c, "global_feature_extractor.ppm.out.conv.1", 128);
    f.ppm_wo = take(256 * 128); f.ppm_bo = take(128);
    // FeatureFusionModule (:190-218)
    add_conv(c, "feature_fusion.dwconv.conv.0", 128 * 9); add_bn(c, "feature_fusion.dwconv.conv.1", 128);
    add_conv(c, "feature_fusion.conv_lower_res.0", 128 * 128, 128); add_bn(c, "feature_fusion.conv_lower_res.1", 128);
    add_conv(c, "feature_fusion.conv_higher_res.0", 128 * 64, 128); add_bn(c, "feature_fusion.conv_higher_res.1", 128);
    f.ffm_wd = take(9 * 128); f.ffm_bd = take(128); f.ffm_wcat = take(192 * 128); f.ffm_bcat = take(128);
    // Classifer (:221-237)
    add_ds(2); add_ds(3);
    add_conv(c, "classifier.conv.1", (int64_t)c->nc * 128, c->nc);
    f.head_w = take((size_t)128 * c->ncp); f.head_b = take(c->ncp);
    if (c->aux) {   // aux head (:24-31)
        add_conv(c, "auxlayer.0", 32 * 64 * 9); add_bn(c, "auxlayer.1", 32);
        add_conv(c, "auxlayer.4", (int64_t)c->nc * 32, c->nc);
        f.aux_w = take(576 * 32); f.aux_b = take(32);
        f.auxh_w = take((size_t)32 * c->ncp); f.auxh_b = take(c->ncp);
    }
    if (c->prec == FSCNN_PREC_BF16)
        for (int i = 0; i < 9; ++i) {
            const int ci = kBnecks[i].cin, ce = 6 * ci, co = kBnecks[i].cout;
            f.bn_weT_img[i] = take((bottleneck_s1t_we_bytes(ci) + 3) / 4);
            f.bn_wpT_img[i] = take((bottleneck_s1t_wp_bytes(ci, co) + 3) / 4);
            f.bn_tabT_img[i] = take((bottleneck_s1t_tab_bytes(ci, co) + 3) / 4);
        }
    if (c->prec == FSCNN_PREC_BF16) {
        for (int i = 0; i < 4; ++i) f.ds_wp_img[i] = take((size_t)dss[i].cin * dss[i].cout / 2);
        f.head_img = take((size_t)((c->nc + 15) & ~15) * 128 / 2);
        f.ffm_img = take((size_t)128 * 192 / 2);
        f.ffm_tabT = take((128 * 32 + 128 * 4) / 4);
        f.ppm_wx_img = take((size_t)128 * 256 / 2);
        f.stem_img = take((size_t)32 * 32 / 2);
        f.stem_img_u8 = take((size_t)32 * 32 / 2);
        f.stem_b_u8 = take(32);
        f.stem_imgx = take((size_t)3 * 32 * 16 / 2);
        f.stem_imgx_u8 = take((size_t)3 * 32 * 16 / 2);
    }
    c->packed_floats = o;
}

struct Loader {
    std::unordered_map<std::string, const fscnn_tensor*> map;
    cudaStream_t s;
    int err = 0;
    const float* get(const std::string& name, int64_t numel) {
        auto it = map.find(name);
        if (it == map.end()) { err = fail(FSCNN_ENOENT, "state_dict tensor '%s' is missing", name.c_str()); return nullptr; }
        if (it->second->numel != numel || !it->second->d_data) {
            err = fail(FSCNN_EINVAL, "state_dict tensor '%s': numel %lld, expected %lld", name.c_str(),
                       (long long)it->second->numel, (long long)numel);
            return nullptr;
        }
        return it->second->d_data;
    }
    // conv `conv` (+ optional bias) followed by BN `bn` ("" = none) -> out_w (k-major, ld_out) / out_b
    void fold(const std::string& conv, bool has_bias, const std::string& bn, int cout, int kdim, int taps, int tap_major,
              float* out_w, int ld_out, float* out_b, int accumulate = 0) {
        if (err) return;
        const float* w = get(conv + ".weight", (int64_t)cout * kdim);
        const float* cb = has_bias ? get(conv + ".bias", cout) : nullptr;
        const float *g = nullptr, *b = nullptr, *m = nullptr, *v = nullptr;
        if (!bn.empty()) {
            g = get(bn + ".weight", cout); b = get(bn + ".bias", cout);
            m = get(bn + ".running_mean", cout); v = get(bn + ".running_var", cout);
        }
        if (err) return;
        if (launch_fold(w, cb, g, b, m, v, cout, kdim, taps, tap_major, out_w, ld_out, out_b, accumulate, s) != cudaSuccess)
            err = fail(FSCNN_ECUDA, "fold kernel launch failed for '%s': %s", conv.c_str(),
                       cudaGetErrorString(cudaGetLastError()));
    }
    // bf16 contexts only: depthwise tables become bf16-representable with error-diffused rounding (fold.cu)
    void round_dw(const fscnn_ctx* c, float* wd, int ch) {
        if (err || c->prec != FSCNN_PREC_BF16) return;
        if (launch_dw_round_bf16(wd, ch, s) != cudaSuccess)
            err = fail(FSCNN_ECUDA, "depthwise rounding launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    void fold_umma(const std::string& conv, const std::string& bn, int nrows, int kdim, int nc, int kc, bf16* out) {
        if (err) return;
        const float* w = get(conv + ".weight", (int64_t)nrows * kdim);
        const float* g = bn.empty() ? nullptr : get(bn + ".weight", nrows);
        const float* v = bn.empty() ? nullptr : get(bn + ".running_var", nrows);
        if (err) return;
        if (launch_fold_umma(w, g, v, nrows, kdim, nc, kc, out, s) != cudaSuccess)
            err = fail(FSCNN_ECUDA, "umma fold launch failed for '%s': %s", conv.c_str(), cudaGetErrorString(cudaGetLastError()));
    }
};

template <typename T>
cudaError_t bottleneck_dispatch(fscnn_ctx* c, int i, const T* in, T* out, int m, int hi, int wi, int ho, int wo, cudaStream_t s);

template <typename T>
cudaError_t dsconv_dispatch(fscnn_ctx* c, int i, int cin, int cout, int stride, const T* in, T* out, bool head, float* logits,
                            int m, int hi, int wi, int ho, int wo, cudaStream_t s);

template <typename T>
cudaError_t ppm_dispatch(fscnn_ctx* c, const T* in, float* rowsum, float* z, bf16* z16, bf16* r_img, T* out, int m, int h, int w, cudaStream_t s);
template <typename T>
cudaError_t ffm_dispatch(fscnn_ctx* c, const T* higher, const T* lower, T* out, int m, int hh, int wh, int hl, int wl, cudaStream_t s);

template <typename T>
cudaError_t stem_dispatch(fscnn_ctx* c, const void* x, T* out, int m, const Dims& d, cudaStream_t s);
// returns true (and sets *e) if stem + dsconv1 were issued as one fused launch
template <typename T>
bool front_fused(fscnn_ctx* c, const void* x, T* out_ds1, int m, const Dims& d, cudaStream_t s, cudaError_t* e);

template <typename T>
int run_stages(fscnn_ctx* c, const void* x, int m, const Dims& d, const WsPlan& p, char* ws, int first, int last,
               cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    auto at = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
    auto atf = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    // programmatic dependent launch for launch sets of up to 160 Mpixel of input (76 Cityscapes images; measured +3.1 % at 16 images
    // per launch, +1.5 % at 32, +0.1 % at 64, -0.7 % at 111); the tail kernel launched after this call for the same images follows
    // the same hint
    pdl_hint() = (long long)m * d.h * d.w <= 160ll * 1000 * 1000;
    for (int st = first; st <= last && e == cudaSuccess; ++st) {
        if (st == kStem && last >= kDs1 && front_fused<T>(c, x, at(p.ds1), m, d, s, &e)) {
            ++st;   // dsconv1 is done too
        } else if (st == kStem) {
            e = stem_dispatch<T>(c, x, at(p.stem), m, d, s);
        } else if (st == kDs1) {
            e = dsconv_dispatch<T>(c, 0, 32, 48, 2, at(p.stem), at(p.ds1), false, nullptr, m, d.h1, d.w1, d.h2, d.w2, s);
        } else if (st == kDs2) {
            e = dsconv_dispatch<T>(c, 1, 48, 64, 2, at(p.ds1), at(p.higher), false, nullptr, m, d.h2, d.w2, d.h3, d.w3, s);
        } else if (st >= kB0 && st < kB0 + 9) {
            const int i = st - kB0;
            const T* in = i == 0 ? at(p.higher) : at(p.b[i - 1]);
            const int hi = i == 0 ? d.h3 : (i <= 3 ? d.h4 : d.h5), wi = i == 0 ? d.w3 : (i <= 3 ? d.w4 : d.w5);
            const int ho = i < 3 ? d.h4 : d.h5, wo = i < 3 ? d.w4 : d.w5;
            e = bottleneck_dispatch<T>(c, i, in, at(p.b[i]), m, hi, wi, ho, wo, s);
        } else if (st == kPpm) {
            e = ppm_dispatch<T>(c, at(p.b[8]), atf(p.rowsum), atf(p.z), reinterpret_cast<bf16*>(ws + p.z16), reinterpret_cast<bf16*>(ws + p.ppm_r),
                                at(p.ppm), m, d.h5, d.w5, s);
            c->launches += 2;
        } else if (st == kFfm) {
            e = ffm_dispatch<T>(c, at(p.higher), at(p.ppm), at(p.ffm), m, d.h3, d.w3, d.h5, d.w5, s);
        } else if (st == kCls1) {
            e = dsconv_dispatch<T>(c, 2, 128, 128, 1, at(p.ffm), at(p.cls1), false, nullptr, m, d.h3, d.w3, d.h3, d.w3, s);
        } else if (st == kCls2Head) {
            e = dsconv_dispatch<T>(c, 3, 128, 128, 1, at(p.cls1), nullptr, true, atf(p.logits), m, d.h3, d.w3, d.h3, d.w3, s);
        } else if (st == kAux) {
            if (!c->aux) continue;
            e = launch_aux<T>(at(p.higher), c->auxw, atf(p.aux_logits), m, d.h3, d.w3, s);
        }
        c->launches += 1;
    }
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

template <>
cudaError_t bottleneck_dispatch<float>(fscnn_ctx* c, int i, const float* in, float* out, int m, int hi, int wi, int ho, int wo,
                                       cudaStream_t s) {
    return launch_bottleneck<float>(kBnecks[i].cin, kBnecks[i].cout, kBnecks[i].stride, in, c->bn[i], out, m, hi, wi, ho, wo, s);
}
template <>
cudaError_t bottleneck_dispatch<bf16>(fscnn_ctx* c, int i, const bf16* in, bf16* out, int m, int hi, int wi, int ho, int wo,
                                      cudaStream_t s) {
    if (kBnecks[i].stride == 2)
        return launch_bottleneck_s2t_tc(kBnecks[i].cin, kBnecks[i].cout, in, c->bn_tabT_img[i], c->bn_weT_img[i], c->bn_wpT_img[i], out,
                                        m, hi, wi, ho, wo, s);
    return launch_bottleneck_s1t_tc(kBnecks[i].cin, kBnecks[i].cout, in, c->bn_tabT_img[i], c->bn_weT_img[i], c->bn_wpT_img[i], out,
                                    m, hi, wi, s);
}

template <>
cudaError_t dsconv_dispatch<float>(fscnn_ctx* c, int i, int cin, int cout, int stride, const float* in, float* out, bool head,
                                   float* logits, int m, int hi, int wi, int ho, int wo, cudaStream_t s) {
    return launch_dsconv<float>(cin, cout, stride, in, c->ds[i], out, head ? &c->head : nullptr, logits, m, hi, wi, ho, wo, s);
}
template <>
cudaError_t dsconv_dispatch<bf16>(fscnn_ctx* c, int i, int cin, int cout, int stride, const bf16* in, bf16* out, bool head,
                                  float* logits, int m, int hi, int wi, int ho, int wo, cudaStream_t s) {
    return launch_dsconv_tc(cin, cout, stride, in, c->ds[i], c->ds_wp_img[i], out, head ? &c->head : nullptr, c->head_img, logits,
                            m, hi, wi, ho, wo, s);
}

template <>
cudaError_t ppm_dispatch<float>(fscnn_ctx* c, const float* in, float* rowsum, float* z, bf16*, bf16*, float* out, int m, int h, int w, cudaStream_t s) {
    return launch_ppm<float>(in, c->ppm, rowsum, z, out, m, h, w, s);
}
template <>
cudaError_t ppm_dispatch<bf16>(fscnn_ctx* c, const bf16* in, float* rowsum, float* z, bf16* z16, bf16* r_img, bf16* out, int m, int h, int w, cudaStream_t s) {
    return launch_ppm_tc(in, c->ppm, c->ppm_wx_img, rowsum, z, z16, r_img, out, m, h, w, s);
}
template <>
cudaError_t ffm_dispatch<float>(fscnn_ctx* c, const float* higher, const float* lower, float* out, int m, int hh, int wh, int hl,
                                int wl, cudaStream_t s) {
    return launch_ffm<float>(higher, lower, c->ffm, out, m, hh, wh, hl, wl, s);
}
template <>
cudaError_t ffm_dispatch<bf16>(fscnn_ctx* c, const bf16* higher, const bf16* lower, bf16* out, int m, int hh, int wh, int hl,
                               int wl, cudaStream_t s) {
    // the network's own resize ratio ((h/4 - 1) / (h - 1) <= 1/4 at every input size) always fits the kernel's 5 x 8 source patch
    if (!ffm_t_supported(hh, wh, hl, wl)) return cudaErrorInvalidValue;
    return launch_ffm_t_tc(higher, lower, c->ffm_tabT, c->ffm_img, out, m, hh, wh, hl, wl, s);
}

template <>
cudaError_t stem_dispatch<float>(fscnn_ctx* c, const void* x, float* out, int m, const Dims& d, cudaStream_t s) {
    return launch_stem<float>(x, c->in, c->stem, out, m, d.h, d.w, d.h1, d.w1, s);
}
template <>
cudaError_t stem_dispatch<bf16>(fscnn_ctx* c, const void* x, bf16* out, int m, const Dims& d, cudaStream_t s) {
    return launch_stem_tc(x, c->in, c->stem_img, c->stem.b, out, m, d.h, d.w, d.h1, d.w1, s);
}

template <>
bool front_fused<float>(fscnn_ctx*, const void*, float*, int, const Dims&, cudaStream_t, cudaError_t*) { return false; }
template <>
bool front_fused<bf16>(fscnn_ctx* c, const void* x, bf16* out_ds1, int m, const Dims& d, cudaStream_t s, cudaError_t* e) {
    if (!c->fuse_front) return false;
    const bool u8 = c->in.format == FSCNN_IN_U8_NHWC;
    if (u8 && c->in_dirty) {
        *e = launch_stem_refold(c->stem.w, c->stem.b, c->in, c->stem_img_u8, c->stem_b_u8, s);
        if (*e == cudaSuccess) *e = launch_stem_pack_rgbx(c->stem.w, c->stem.b, c->in, 1, c->stem_imgx_u8, c->stem_b_u8, s);
        if (*e != cudaSuccess) return true;
        c->in_dirty = false;
        c->launches += 2;
    }
    if (c->front_transposed) {   // transposed stem (l2d_front_t_tc.cu); inputs without 16-byte aligned rows keep the previous kernel
        *e = launch_l2d_front_t_tc(x, c->in, u8 ? c->stem_imgx_u8 : c->stem_imgx, c->ds[0], c->ds_wp_img[0], out_ds1, m, d.h, d.w, d.h1,
                                   d.w1, d.h2, d.w2, s);
        if (*e != cudaErrorNotSupported) return true;
    }
    *e = launch_l2d_front_tc(x, c->in, u8 ? c->stem_imgx_u8 : c->stem_imgx, u8 ? c->stem_b_u8 : c->stem.b, c->ds[0], c->ds_wp_img[0],
                             out_ds1, m, d.h, d.w, d.h1, d.w1, d.h2, d.w2, s);
    return true;
}

int dispatch_stages(fscnn_ctx* c, const void* x, int m, const Dims& d, const WsPlan& p, char* ws, int first, int last,
                    cudaStream_t s) {
    return c->prec == FSCNN_PREC_BF16 ? run_stages<bf16>(c, x, m, d, p, ws, first, last, s)
                                      : run_stages<float>(c, x, m, d, p, ws, first, last, s);
}

int check_forward_args(const fscnn_ctx* c, const void* x, int n, int h, int w, const void* ws, size_t ws_bytes, Dims* d,
                       int* mb, WsPlan* plan) {
    if (!c) return fail(FSCNN_EINVAL, "null context");
    if (!c->loaded) return fail(FSCNN_ESTATE, "fscnn_load_weights has not been called");
    if (!x || !ws) return fail(FSCNN_EINVAL, "null input or workspace pointer");
    if (n < 1) return fail(FSCNN_EINVAL, "batch size %d < 1", n);
    *d = make_dims(h, w);
    if (!d->ok) return fail(FSCNN_EINVAL, "input %dx%d is too small for five stride-2 stages", h, w);
    if (((uintptr_t)x & 15) || ((uintptr_t)ws & 255)) return fail(FSCNN_EINVAL, "input must be 16-byte and workspace 256-byte aligned");
    *mb = c->eff_mb(n, h, w);
    *plan = c->plan(*mb, *d);
    if (ws_bytes < plan->total)
        return fail(FSCNN_ENOMEM, "workspace of %zu bytes is smaller than the %zu needed", ws_bytes, plan->total);
    return FSCNN_OK;
}

}  // namespace

extern "C" {

int fscnn_abi_version(void) { return FSCNN_ABI_VERSION; }
const char* fscnn_last_error(void) { return g_err.c_str(); }
int64_t fscnn_conf_len(int num_classes) { return (int64_t)(num_classes + 1) * (num_classes + 1) + 2; }

int fscnn_create(fscnn_ctx** out, int num_classes, int aux, int precision) {
    if (!out) return fail(FSCNN_EINVAL, "null output pointer");
    if (precision != FSCNN_PREC_FP32 && precision != FSCNN_PREC_BF16) return fail(FSCNN_EINVAL, "unknown precision %d", precision);
    // capability limits, stated here instead of at the first forward: the x8 upsample stages 960 bytes of shared memory per class
    // (227 KB: 240 classes), the bf16 head kernel keeps 2 x round_up(nc, 16) accumulator columns next to 256 others in TMEM (128)
    const int max_nc = precision == FSCNN_PREC_BF16 ? 128 : 240;
    if (num_classes < 1 || num_classes > max_nc)
        return fail(FSCNN_EINVAL, "num_classes %d outside [1, %d] for the %s path", num_classes, max_nc,
                    precision == FSCNN_PREC_BF16 ? "bf16" : "fp32");
    fscnn_ctx* c = new fscnn_ctx();
    c->nc = num_classes;
    c->ncp = (num_classes + 3) & ~3;
    c->aux = aux ? 1 : 0;
    c->prec = precision;
    build_manifest_and_offsets(c);
    *out = c;
    return FSCNN_OK;
}

void fscnn_destroy(fscnn_ctx* ctx) { delete ctx; }

int fscnn_param_count(const fscnn_ctx* ctx) { return ctx ? (int)ctx->params.size() : 0; }
const char* fscnn_param_name(const fscnn_ctx* ctx, int i) {
    return (ctx && i >= 0 && i < (int)ctx->params.size()) ? ctx->params[i].name.c_str() : nullptr;
}
int64_t fscnn_param_numel(const fscnn_ctx* ctx, int i) {
    return (ctx && i >= 0 && i < (int)ctx->params.size()) ? ctx->params[i].numel : -1;
}

int fscnn_packed_weight_bytes(const fscnn_ctx* ctx, size_t* out_bytes) {
    if (!ctx || !out_bytes) return fail(FSCNN_EINVAL, "null argument");
    *out_bytes = ctx->packed_floats * sizeof(float);
    return FSCNN_OK;
}

int fscnn_load_weights(fscnn_ctx* c, const fscnn_tensor* tensors, int n_tensors, void* d_packed, size_t packed_bytes,
                       void* stream) {
    if (!c || !tensors || !d_packed) return fail(FSCNN_EINVAL, "null argument");
    if (packed_bytes < c->packed_floats * sizeof(float))
        return fail(FSCNN_ENOMEM, "packed buffer of %zu bytes is smaller than the %zu needed", packed_bytes,
                    c->packed_floats * sizeof(float));
    if ((uintptr_t)d_packed & 255) return fail(FSCNN_EINVAL, "packed buffer must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    Loader L;
    L.s = s;
    for (int i = 0; i < n_tensors; ++i)
        if (tensors[i].name) L.map[tensors[i].name] = &tensors[i];
    float* P = reinterpret_cast<float*>(d_packed);
    if (cudaMemsetAsync(P, 0, c->packed_floats * sizeof(float), s) != cudaSuccess)
        return fail(FSCNN_ECUDA, "memset of the packed buffer failed: %s", cudaGetErrorString(cudaGetLastError()));
    const auto& f = c->off;
    c->loaded = false;

    L.fold("learning_to_downsample.conv.conv.0", false, "learning_to_downsample.conv.conv.1", 32, 27, 1, 0, P + f.stem_w, 32,
           P + f.stem_b);
    c->stem = {P + f.stem_w, P + f.stem_b};
    if (c->prec == FSCNN_PREC_BF16) {   // [32 x 27] padded to K = 32 (columns 27..31 stay zero)
        bf16* img = reinterpret_cast<bf16*>(P + f.stem_img);
        L.fold_umma("learning_to_downsample.conv.conv.0", "learning_to_downsample.conv.conv.1", 32, 27, 32, 32, img);
        c->stem_img = img;
        c->stem_img_u8 = reinterpret_cast<bf16*>(P + f.stem_img_u8);
        c->stem_b_u8 = P + f.stem_b_u8;
        c->stem_imgx = reinterpret_cast<bf16*>(P + f.stem_imgx);
        c->stem_imgx_u8 = reinterpret_cast<bf16*>(P + f.stem_imgx_u8);
        if (!L.err && launch_stem_pack_rgbx(c->stem.w, c->stem.b, c->in, 0, c->stem_imgx, nullptr, L.s) != cudaSuccess)
            L.err = fail(FSCNN_ECUDA, "stem pack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->in_dirty = true;
    }
    const struct { const char* p; int cin, cout; } dss[4] = {{"learning_to_downsample.dsconv1", 32, 48},
                                                              {"learning_to_downsample.dsconv2", 48, 64},
                                                              {"classifier.dsconv1", 128, 128},
                                                              {"classifier.dsconv2", 128, 128}};
    for (int i = 0; i < 4; ++i) {
        const std::string p = dss[i].p;
        L.fold(p + ".conv.0", false, p + ".conv.1", dss[i].cin, 9, 1, 0, P + f.ds_wd[i], dss[i].cin, P + f.ds_bd[i]);
        L.round_dw(c, P + f.ds_wd[i], dss[i].cin);
        L.fold(p + ".conv.3", false, p + ".conv.4", dss[i].cout, dss[i].cin, 1, 0, P + f.ds_wp[i], dss[i].cout, P + f.ds_bp[i]);
        c->ds[i] = {P + f.ds_wd[i], P + f.ds_bd[i], P + f.ds_wp[i], P + f.ds_bp[i]};
        if (c->prec == FSCNN_PREC_BF16) {
            bf16* img = reinterpret_cast<bf16*>(P + f.ds_wp_img[i]);
            L.fold_umma(p + ".conv.3", p + ".conv.4", dss[i].cout, dss[i].cin, dss[i].cout, dss[i].cin, img);
            c->ds_wp_img[i] = img;
        }
    }
    for (int i = 0; i < 9; ++i) {
        const std::string p = std::string("global_feature_extractor.") + kBnecks[i].name + ".block";
        const int ci = kBnecks[i].cin, ce = 6 * ci, co = kBnecks[i].cout;
        L.fold(p + ".0.conv.0", false, p + ".0.conv.1", ce, ci, 1, 0, P + f.bn_we[i], ce, P + f.bn_be[i]);
        L.fold(p + ".1.conv.0", false, p + ".1.conv.1", ce, 9, 1, 0, P + f.bn_wd[i], ce, P + f.bn_bd[i]);
        L.round_dw(c, P + f.bn_wd[i], ce);
        L.fold(p + ".2", false, p + ".3", co, ce, 1, 0, P + f.bn_wp[i], co, P + f.bn_bp[i]);
        c->bn[i] = {P + f.bn_we[i], P + f.bn_be[i], P + f.bn_wd[i], P + f.bn_bd[i], P + f.bn_wp[i], P + f.bn_bp[i]};
        if (c->prec == FSCNN_PREC_BF16) {
            {
                bf16* weT = reinterpret_cast<bf16*>(P + f.bn_weT_img[i]);
                bf16* wpT = reinterpret_cast<bf16*>(P + f.bn_wpT_img[i]);
                unsigned char* tabT = reinterpret_cast<unsigned char*>(P + f.bn_tabT_img[i]);
                if (!L.err && cudaMemsetAsync(wpT, 0, bottleneck_s1t_wp_bytes(ci, co), L.s) != cudaSuccess)
                    L.err = fail(FSCNN_ECUDA, "memset failed: %s", cudaGetErrorString(cudaGetLastError()));
                L.fold_umma(p + ".2", p + ".3", co, ce, co, 128, wpT);              // project: chunks of 128 columns
                if (!L.err && launch_pack_s1t(c->bn[i], ci, co, weT, tabT, L.s) != cudaSuccess)
                    L.err = fail(FSCNN_ECUDA, "s1t pack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                c->bn_weT_img[i] = weT;
                c->bn_wpT_img[i] = wpT;
                c->bn_tabT_img[i] = tabT;
            }
        }
    }
    for (int i = 0; i < 4; ++i) {
        const std::string p = "global_feature_extractor.ppm.conv" + std::to_string(i + 1) + ".conv";
        L.fold(p + ".0", false, p + ".1", 32, 128, 1, 0, P + f.ppm_wc[i], 32, P + f.ppm_bc[i]);
        c->ppm.wc[i] = P + f.ppm_wc[i];
        c->ppm.bc[i] = P + f.ppm_bc[i];
        c->ppm.wo_s[i] = P + f.ppm_wo + (size_t)(128 + 32 * i) * 128;   // cat order: x, feat1..feat4 (:143)
    }
    L.fold("global_feature_extractor.ppm.out.conv.0", false, "global_feature_extractor.ppm.out.conv.1", 128, 256, 1, 0,
           P + f.ppm_wo, 128, P + f.ppm_bo);
    c->ppm.wo_x = P + f.ppm_wo;
    if (c->prec == FSCNN_PREC_BF16) {
        bf16* img = reinterpret_cast<bf16*>(P + f.ppm_wx_img);
        L.fold_umma("global_feature_extractor.ppm.out.conv.0", "global_feature_extractor.ppm.out.conv.1", 128, 256, 128, 128, img);
        c->ppm_wx_img = img;   // chunk 0 = the 128 input channels of x (cat order: x first, :143)
    }
    c->ppm.bo = P + f.ppm_bo;
    L.fold("feature_fusion.dwconv.conv.0", false, "feature_fusion.dwconv.conv.1", 128, 9, 1, 0, P + f.ffm_wd, 128, P + f.ffm_bd);
    L.round_dw(c, P + f.ffm_wd, 128);
    L.fold("feature_fusion.conv_higher_res.0", true, "feature_fusion.conv_higher_res.1", 128, 64, 1, 0, P + f.ffm_wcat, 128,
           P + f.ffm_bcat, 0);
    L.fold("feature_fusion.conv_lower_res.0", true, "feature_fusion.conv_lower_res.1", 128, 128, 1, 0,
           P + f.ffm_wcat + (size_t)64 * 128, 128, P + f.ffm_bcat, 1);
    c->ffm = {P + f.ffm_wd, P + f.ffm_bd, P + f.ffm_wcat, P + f.ffm_bcat};
    if (c->prec == FSCNN_PREC_BF16) {   // [128 x 192] image, k-block major: columns 0..63 = higher, 64..191 = lower
        bf16* img = reinterpret_cast<bf16*>(P + f.ffm_img);
        L.fold_umma("feature_fusion.conv_higher_res.0", "feature_fusion.conv_higher_res.1", 128, 64, 128, 64, img);
        L.fold_umma("feature_fusion.conv_lower_res.0", "feature_fusion.conv_lower_res.1", 128, 128, 128, 128, img + 128 * 64);
        c->ffm_img = img;
        unsigned char* ftab = reinterpret_cast<unsigned char*>(P + f.ffm_tabT);
        if (!L.err && launch_pack_dw_tab(c->ffm.wd, c->ffm.bd, c->ffm.bcat, 128, 128, ftab, L.s) != cudaSuccess)
            L.err = fail(FSCNN_ECUDA, "ffm table pack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->ffm_tabT = ftab;
    }
    L.fold("classifier.conv.1", true, "", c->nc, 128, 1, 0, P + f.head_w, c->ncp, P + f.head_b);
    c->head = {P + f.head_w, P + f.head_b, c->nc, c->ncp};
    if (c->prec == FSCNN_PREC_BF16) {
        bf16* img = reinterpret_cast<bf16*>(P + f.head_img);
        L.fold_umma("classifier.conv.1", "", c->nc, 128, (c->nc + 15) & ~15, 128, img);   // rows >= nc stay zero
        c->head_img = img;
    }
    if (c->aux) {
        L.fold("auxlayer.0", false, "auxlayer.1", 32, 576, 9, 1, P + f.aux_w, 32, P + f.aux_b);
        L.fold("auxlayer.4", true, "", c->nc, 32, 1, 0, P + f.auxh_w, c->ncp, P + f.auxh_b);
        c->auxw = {P + f.aux_w, P + f.aux_b, {P + f.auxh_w, P + f.auxh_b, c->nc, c->ncp}};
    }
    if (L.err) return L.err;
    c->loaded = true;
    return FSCNN_OK;
}

int fscnn_workspace_bytes(const fscnn_ctx* c, int n, int h, int w, size_t* out_bytes) {
    if (!c || !out_bytes) return fail(FSCNN_EINVAL, "null argument");
    const Dims d = make_dims(h, w);
    if (!d.ok || n < 1) return fail(FSCNN_EINVAL, "bad shape n=%d h=%d w=%d", n, h, w);
    *out_bytes = c->plan(c->eff_mb(n, h, w), d).total;
    return FSCNN_OK;
}

int fscnn_set_input_format(fscnn_ctx* c, int format, const float* mean3, const float* std3) {
    if (!c) return fail(FSCNN_EINVAL, "null context");
    if (format != FSCNN_IN_F32_NCHW && format != FSCNN_IN_U8_NHWC) return fail(FSCNN_EINVAL, "unknown input format %d", format);
    c->in.format = format;
    for (int i = 0; i < 3; ++i) {
        c->in.mean[i] = mean3 ? mean3[i] : 0.f;
        const float sd = std3 ? std3[i] : 1.f;
        if (!(sd > 0.f)) return fail(FSCNN_EINVAL, "std[%d] must be positive", i);
        c->in.inv_std[i] = 1.f / sd;
    }
    c->in_dirty = true;
    return FSCNN_OK;
}

int fscnn_set_option(fscnn_ctx* c, const char* key, int value) {
    if (!c || !key) return fail(FSCNN_EINVAL, "null argument");
    if (!strcmp(key, "fuse_front")) { c->fuse_front = value ? 1 : 0; return FSCNN_OK; }
    if (!strcmp(key, "front_transposed")) { c->front_transposed = value ? 1 : 0; return FSCNN_OK; }
    if (!strcmp(key, "micro_batch")) return fscnn_set_micro_batch(c, value);
    return fail(FSCNN_ENOENT, "unknown option '%s'", key);
}

int fscnn_set_micro_batch(fscnn_ctx* c, int images) {
    if (!c || images < 0) return fail(FSCNN_EINVAL, "bad argument");
    c->micro_batch = images;
    return FSCNN_OK;
}

int64_t fscnn_launch_count(const fscnn_ctx* c) { return c ? c->launches : 0; }
int fscnn_stage_count(const fscnn_ctx* c) { return c ? (c->aux ? kNumStages : kNumStages - 1) : 0; }
const char* fscnn_stage_name(const fscnn_ctx* c, int stage) {
    return (c && stage >= 0 && stage < fscnn_stage_count(c)) ? kStageNames[stage] : nullptr;
}

int fscnn_tap_info(const fscnn_ctx* c, int n, int h, int w, const char* tap, fscnn_tap* out) {
    if (!c || !tap || !out) return fail(FSCNN_EINVAL, "null argument");
    const Dims d = make_dims(h, w);
    if (!d.ok || n < 1) return fail(FSCNN_EINVAL, "bad shape n=%d h=%d w=%d", n, h, w);
    const int mb = c->eff_mb(n, h, w);
    const WsPlan p = c->plan(mb, d);
    const std::string t = tap;
    const int es = c->esize();
    auto set = [&](size_t off, int hh, int ww, int ch, int cs, int eb) {
        *out = {off, mb, hh, ww, ch, cs, eb};
        return FSCNN_OK;
    };
    if (t == "l2d.conv") return set(p.stem, d.h1, d.w1, 32, 32, es);
    if (t == "l2d.dsconv1") return set(p.ds1, d.h2, d.w2, 48, 48, es);
    if (t == "l2d.dsconv2") return set(p.higher, d.h3, d.w3, 64, 64, es);
    for (int i = 0; i < 9; ++i)
        if (t == std::string("gfe.") + kBnecks[i].name)
            return set(p.b[i], i < 3 ? d.h4 : d.h5, i < 3 ? d.w4 : d.w5, kBnecks[i].cout, kBnecks[i].cout, es);
    if (t == "gfe.ppm") return set(p.ppm, d.h5, d.w5, 128, 128, es);
    if (t == "ffm") return set(p.ffm, d.h3, d.w3, 128, 128, es);
    if (t == "cls.dsconv1") return set(p.cls1, d.h3, d.w3, 128, 128, es);
    if (t == "cls.logits_lowres") return set(p.logits, d.h3, d.w3, c->nc, c->ncp, 4);
    if (t == "aux.logits_lowres" && c->aux) return set(p.aux_logits, d.h3, d.w3, c->nc, c->ncp, 4);
    return fail(FSCNN_ENOENT, "unknown tap '%s'", tap);
}

int fscnn_forward_range(fscnn_ctx* c, const void* d_x, int n, int h, int w, int first, int last, void* ws, size_t ws_bytes,
                        void* stream) {
    Dims d; int mb; WsPlan p;
    int rc = check_forward_args(c, d_x, n, h, w, ws, ws_bytes, &d, &mb, &p);
    if (rc) return rc;
    if (n > mb) return fail(FSCNN_EINVAL, "forward_range handles one micro-batch (%d images), got %d", mb, n);
    if (first < 0 || last >= fscnn_stage_count(c) || first > last) return fail(FSCNN_EINVAL, "bad stage range [%d, %d]", first, last);
    return dispatch_stages(c, d_x, n, d, p, (char*)ws, first, last, (cudaStream_t)stream);
}

int fscnn_forward_logits(fscnn_ctx* c, const void* d_x, int n, int h, int w, float* d_logits, float* d_aux, void* ws,
                         size_t ws_bytes, void* stream) {
    Dims d; int mb; WsPlan p;
    int rc = check_forward_args(c, d_x, n, h, w, ws, ws_bytes, &d, &mb, &p);
    if (rc) return rc;
    if (!d_logits) return fail(FSCNN_EINVAL, "null logits pointer");
    if (((uintptr_t)d_logits & 15) || ((uintptr_t)d_aux & 15)) return fail(FSCNN_EINVAL, "logits must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t img = (size_t)3 * h * w * (c->in.format == FSCNN_IN_U8_NHWC ? 1 : 4), lg = (size_t)c->nc * h * w;
    const bool want_aux = c->aux && d_aux;
    for (int i0 = 0; i0 < n; i0 += mb) {
        const int m = n - i0 < mb ? n - i0 : mb;
        rc = dispatch_stages(c, (const char*)d_x + i0 * img, m, d, p, (char*)ws, kStem, want_aux ? kAux : kCls2Head, s);
        if (rc) return rc;
        cudaError_t e = launch_up_logits(reinterpret_cast<float*>((char*)ws + p.logits), c->nc, c->ncp, d_logits + i0 * lg, m,
                                         d.h3, d.w3, h, w, s);
        c->launches += 1;
        if (e == cudaSuccess && want_aux) {
            e = launch_up_logits(reinterpret_cast<float*>((char*)ws + p.aux_logits), c->nc, c->ncp, d_aux + i0 * lg, m, d.h3,
                                 d.w3, h, w, s);
            c->launches += 1;
        }
        if (e != cudaSuccess) return fail(FSCNN_ECUDA, "upsample launch failed: %s", cudaGetErrorString(e));
    }
    return FSCNN_OK;
}

static bool valid_label_dtype(int dt) { return dt == FSCNN_U8 || dt == FSCNN_I32 || dt == FSCNN_I64; }
// class maps are read / written four elements at a time (uchar4 / int4 / 2 x longlong2) when the width is a multiple of 4
static bool vec4_aligned(const void* p, int dt) {
    const uintptr_t a = dt == FSCNN_U8 ? 4 : 16;
    return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0;
}

static int forward_mask_impl(fscnn_ctx* c, const void* d_x, const void* d_labels, int label_dtype, int n, int h, int w,
                             long long* d_conf, void* d_mask, int mask_dtype, void* ws, size_t ws_bytes, void* stream) {
    Dims d; int mb; WsPlan p;
    int rc = check_forward_args(c, d_x, n, h, w, ws, ws_bytes, &d, &mb, &p);
    if (rc) return rc;
    if (!valid_label_dtype(mask_dtype)) return fail(FSCNN_EINVAL, "bad mask dtype %d", mask_dtype);
    if (d_labels && !valid_label_dtype(label_dtype)) return fail(FSCNN_EINVAL, "bad label dtype %d", label_dtype);
    if (!vec4_aligned(d_mask, mask_dtype)) return fail(FSCNN_EINVAL, "mask must be aligned to 4 elements (4 bytes for uint8, 16 bytes for int32 / int64)");
    if (d_labels && !vec4_aligned(d_labels, label_dtype))
        return fail(FSCNN_EINVAL, "labels must be aligned to 4 elements (4 bytes for uint8, 16 bytes for int32 / int64)");
    const size_t msz = mask_dtype == FSCNN_U8 ? 1 : (mask_dtype == FSCNN_I32 ? 4 : 8);
    const size_t lsz = label_dtype == FSCNN_U8 ? 1 : (label_dtype == FSCNN_I32 ? 4 : 8);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t img = (size_t)3 * h * w * (c->in.format == FSCNN_IN_U8_NHWC ? 1 : 4), px = (size_t)h * w;
    for (int i0 = 0; i0 < n; i0 += mb) {
        const int m = n - i0 < mb ? n - i0 : mb;
        rc = dispatch_stages(c, (const char*)d_x + i0 * img, m, d, p, (char*)ws, kStem, kCls2Head, s);
        if (rc) return rc;
        cudaError_t e = launch_up_argmax(reinterpret_cast<float*>((char*)ws + p.logits), c->nc, c->ncp,
                                         d_mask ? (char*)d_mask + i0 * px * msz : nullptr, mask_dtype,
                                         d_labels ? (const char*)d_labels + i0 * px * lsz : nullptr, label_dtype,
                                         reinterpret_cast<unsigned long long*>(d_conf), m, d.h3, d.w3, h, w, s);
        c->launches += 1;
        if (e != cudaSuccess) return fail(FSCNN_ECUDA, "upsample+argmax launch failed: %s", cudaGetErrorString(e));
    }
    return FSCNN_OK;
}

int fscnn_forward_mask(fscnn_ctx* c, const void* d_x, int n, int h, int w, void* d_mask, int mask_dtype, void* ws,
                       size_t ws_bytes, void* stream) {
    if (!d_mask) return fail(FSCNN_EINVAL, "null mask pointer");
    return forward_mask_impl(c, d_x, nullptr, 0, n, h, w, nullptr, d_mask, mask_dtype, ws, ws_bytes, stream);
}

int fscnn_forward_confusion(fscnn_ctx* c, const void* d_x, const void* d_labels, int label_dtype, int n, int h, int w,
                            long long* d_conf, void* d_mask, int mask_dtype, void* ws, size_t ws_bytes, void* stream) {
    if (!d_labels || !d_conf) return fail(FSCNN_EINVAL, "null labels or confusion pointer");
    return forward_mask_impl(c, d_x, d_labels, label_dtype, n, h, w, d_conf, d_mask, d_mask ? mask_dtype : FSCNN_U8, ws,
                             ws_bytes, stream);
}

int fscnn_confusion_from_mask(const void* d_pred, int pred_dtype, const void* d_label, int label_dtype, int64_t n_pixels,
                              int num_classes, long long* d_conf, void* stream) {
    if (!d_conf || n_pixels < 0) return fail(FSCNN_EINVAL, "bad argument");
    if (n_pixels == 0) return FSCNN_OK;
    if (!d_pred || !d_label) return fail(FSCNN_EINVAL, "null class map");
    if (num_classes < 1 || num_classes > 1 << 15) return fail(FSCNN_EINVAL, "num_classes %d out of range", num_classes);
    for (int dt : {pred_dtype, label_dtype})
        if (dt != FSCNN_U8 && dt != FSCNN_I32 && dt != FSCNN_I64) return fail(FSCNN_EINVAL, "bad dtype %d", dt);
    cudaError_t e = launch_confusion(d_pred, pred_dtype, d_label, label_dtype, n_pixels, num_classes,
                                     reinterpret_cast<unsigned long long*>(d_conf), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "confusion launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_colorize(const void* d_mask, int mask_dtype, int64_t n_pixels, const unsigned char* h_palette768, unsigned char* d_rgb,
                   void* stream) {
    if (n_pixels < 0 || !h_palette768) return fail(FSCNN_EINVAL, "bad argument");
    if (n_pixels == 0) return FSCNN_OK;
    if (!d_mask || !d_rgb) return fail(FSCNN_EINVAL, "null device pointer");
    if (mask_dtype != FSCNN_U8 && mask_dtype != FSCNN_I32 && mask_dtype != FSCNN_I64) return fail(FSCNN_EINVAL, "bad dtype %d", mask_dtype);
    if ((uintptr_t)d_rgb & 3) return fail(FSCNN_EINVAL, "rgb output must be 4-byte aligned");
    cudaError_t e = launch_colorize(d_mask, mask_dtype, n_pixels, h_palette768, d_rgb, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "colorize launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_overlay(const unsigned char* d_image, const void* d_mask, int mask_dtype, int64_t n_pixels, const unsigned char* h_palette768,
                  const unsigned int* h_draw_classes8, double alpha, unsigned char* d_out, void* stream) {
    if (n_pixels < 0 || !h_palette768 || !h_draw_classes8 || !(alpha >= 0.0 && alpha <= 1.0)) return fail(FSCNN_EINVAL, "bad argument");
    if (n_pixels == 0) return FSCNN_OK;
    if (!d_image || !d_mask || !d_out) return fail(FSCNN_EINVAL, "null device pointer");
    if (!valid_label_dtype(mask_dtype)) return fail(FSCNN_EINVAL, "bad dtype %d", mask_dtype);
    cudaError_t e = launch_overlay(d_image, d_mask, mask_dtype, n_pixels, h_palette768, h_draw_classes8, alpha, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "overlay launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_e2e_preprocess(const void* d_frames, int frame_dtype, int n, int h, int w, int base_size, const float* h_mean3,
                         const float* h_std3, float* d_out, void* stream) {
    if (!d_frames || !d_out) return fail(FSCNN_EINVAL, "null device pointer");
    if (frame_dtype != FSCNN_U8 && frame_dtype != FSCNN_F32) return fail(FSCNN_EINVAL, "frames must be uint8 or float32, got dtype %d", frame_dtype);
    if ((h_mean3 == nullptr) != (h_std3 == nullptr)) return fail(FSCNN_EINVAL, "mean and std go together");
    if (n < 1 || h < 1 || w < 1 || base_size < 1) return fail(FSCNN_EINVAL, "bad shape n=%d h=%d w=%d base=%d", n, h, w, base_size);
    cudaError_t e = launch_e2e_preprocess(d_frames, frame_dtype == FSCNN_U8, n, h, w, base_size, h_mean3, h_std3, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "e2e preprocess launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_e2e_postprocess(const float* d_low_logits, int num_classes, int padded_classes, int n, int hl, int wl, int base_h, int base_w,
                          int out_h, int out_w, int apply_softmax, float* d_out, void* stream) {
    if (!d_low_logits || !d_out) return fail(FSCNN_EINVAL, "null device pointer");
    if (num_classes < 1 || num_classes > 32 || padded_classes < num_classes)
        return fail(FSCNN_EINVAL, "e2e postprocess supports 1..32 classes (got %d, padded %d)", num_classes, padded_classes);
    if (n < 1 || hl < 1 || wl < 1 || base_h < 1 || base_w < 1 || out_h < 1 || out_w < 1) return fail(FSCNN_EINVAL, "bad shape");
    cudaError_t e = launch_e2e_postprocess(d_low_logits, num_classes, padded_classes, n, hl, wl, base_h, base_w, out_h, out_w, apply_softmax,
                                           d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "e2e postprocess launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_upsample_argmax(const float* d_low_logits, int nc, int ncp, int n, int hl, int wl, int h, int w, void* d_mask, int mask_dtype,
                          const void* d_labels, int label_dtype, long long* d_conf, int flags, void* stream) {
    if (!d_low_logits || (!d_mask && !d_labels)) return fail(FSCNN_EINVAL, "null device pointer");
    if (nc < 1 || nc > 256 || ncp < nc || (ncp & 3)) return fail(FSCNN_EINVAL, "bad class count %d (padded %d)", nc, ncp);
    if (n < 1 || hl < 1 || wl < 1 || h < 1 || w < 1) return fail(FSCNN_EINVAL, "bad shape");
    if ((double)(hl - 1) * 7.3 > (double)(h - 1) || (double)(wl - 1) * 7.3 > (double)(w - 1))
        return fail(FSCNN_EINVAL, "upsample ratio must be >= 7.3 (%dx%d -> %dx%d)", hl, wl, h, w);
    if (d_mask && !valid_label_dtype(mask_dtype)) return fail(FSCNN_EINVAL, "bad mask dtype %d", mask_dtype);
    if (d_labels && (!valid_label_dtype(label_dtype) || !d_conf)) return fail(FSCNN_EINVAL, "labels need a valid dtype and d_conf");
    if ((reinterpret_cast<uintptr_t>(d_low_logits) & 15) || !vec4_aligned(d_mask, mask_dtype) || !vec4_aligned(d_labels, label_dtype))
        return fail(FSCNN_EINVAL, "d_low_logits must be 16-byte aligned, d_mask / d_labels aligned to 4 elements");
    cudaError_t e = launch_up_argmax(d_low_logits, nc, ncp, d_mask, mask_dtype, d_labels, label_dtype,
                                     reinterpret_cast<unsigned long long*>(d_conf), n, hl, wl, h, w, (cudaStream_t)stream,
                                     !(flags & FSCNN_TAIL_EXHAUSTIVE));
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "upsample+argmax launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_conf_to_totals(const long long* h_conf, int nc, long long* h_inter, long long* h_union, long long* h_correct,
                         long long* h_label) {
    if (!h_conf || !h_inter || !h_union || !h_correct || !h_label || nc < 1) return fail(FSCNN_EINVAL, "bad argument");
    const int w = nc + 1;
    for (int k = 0; k < nc; ++k) {
        long long area_pred = 0, area_lab = 0;
        for (int r = 0; r < w; ++r) area_pred += h_conf[r * w + k];   // every labeled pixel predicted k
        for (int q = 0; q < w; ++q) area_lab += h_conf[k * w + q];    // every pixel labeled k
        h_inter[k] = h_conf[k * w + k];
        h_union[k] = area_pred + area_lab - h_inter[k];
    }
    *h_label = h_conf[w * w];
    *h_correct = h_conf[w * w + 1];
    return FSCNN_OK;
}

// ---- training step, first slice: see include/fscnn_b200.h ----
static int train_ws_ok(const void* ws, size_t have, size_t need) {
    if (!ws) return fail(FSCNN_EINVAL, "null workspace");
    if (have < need) return fail(FSCNN_ENOMEM, "training workspace too small: %zu < %zu bytes", have, need);
    if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(FSCNN_EINVAL, "workspace must be 16-byte aligned");
    return FSCNN_OK;
}

int fscnn_train_workspace_bytes(int max_channels, int max_cout, int max_cin, size_t* out) {
    if (!out || max_channels < 1 || max_cout < 1 || max_cin < 1) return fail(FSCNN_EINVAL, "bad argument");
    *out = train_workspace_bytes(max_channels, max_cout, max_cin);
    return FSCNN_OK;
}

int fscnn_train_dwconv3x3_forward(const float* d_x, const float* d_w, float* d_y, int n, int c, int h, int w, int stride, void* stream) {
    if (!d_x || !d_w || !d_y) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || h < 1 || w < 1 || (stride != 1 && stride != 2)) return fail(FSCNN_EINVAL, "bad shape n=%d c=%d h=%d w=%d stride=%d", n, c, h, w, stride);
    cudaError_t e = launch_train_dw_fwd(d_x, d_w, d_y, n, c, h, w, stride, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "depthwise forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_dwconv3x3_backward(const float* d_x, const float* d_w, const float* d_dy, float* d_dx, float* d_dw, void* d_ws,
                                   size_t ws_bytes, int n, int c, int h, int w, int stride, void* stream) {
    if (!d_x || !d_w || !d_dy || (!d_dx && !d_dw)) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || h < 1 || w < 1 || (stride != 1 && stride != 2)) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(c, 1, 1));
    if (rc) return rc;
    cudaError_t e = launch_train_dw_bwd(d_x, d_w, d_dy, d_dx, d_dw, d_ws, n, c, h, w, stride, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "depthwise backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_pwconv_forward(const float* d_x, const float* d_w, float* d_y, int n, int cin, int cout, int hw, void* stream) {
    if (!d_x || !d_w || !d_y) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || n > 65535 || cin < 1 || cout < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad shape");
    cudaError_t e = launch_train_pw_fwd(d_x, d_w, d_y, n, cin, cout, hw, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "pointwise forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_pwconv_backward(const float* d_x, const float* d_w, const float* d_dy, float* d_dx, float* d_dw, void* d_ws,
                                size_t ws_bytes, int n, int cin, int cout, int hw, void* stream) {
    if (!d_x || !d_w || !d_dy || (!d_dx && !d_dw)) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || n > 64 || cin < 1 || cout < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad shape (the weight gradient takes at most 64 images per call)");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(1, cout, cin));
    if (rc) return rc;
    cudaError_t e = launch_train_pw_bwd(d_x, d_w, d_dy, d_dx, d_dw, d_ws, n, cin, cout, hw, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "pointwise backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_batchnorm_forward(const float* d_x, const float* d_gamma, const float* d_beta, float* d_running_mean,
                                  float* d_running_var, float* d_y, float* d_save_mean, float* d_save_rstd, void* d_ws, size_t ws_bytes,
                                  int n, int c, int hw, float eps, float momentum, int relu, void* stream) {
    if (!d_x || !d_gamma || !d_beta || !d_y || !d_save_mean || !d_save_rstd) return fail(FSCNN_EINVAL, "null device pointer");
    if ((d_running_mean == nullptr) != (d_running_var == nullptr)) return fail(FSCNN_EINVAL, "running mean and var go together");
    if (n < 1 || c < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(c, 1, 1));
    if (rc) return rc;
    cudaError_t e = launch_train_bn_fwd(d_x, d_gamma, d_beta, d_running_mean, d_running_var, d_y, d_save_mean, d_save_rstd, d_ws, n, c, hw,
                                        eps, momentum, relu, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "batchnorm forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_batchnorm_backward(const float* d_x, const float* d_dy, const float* d_gamma, const float* d_beta, const float* d_save_mean,
                                   const float* d_save_rstd, float* d_dx, float* d_dgamma, float* d_dbeta, void* d_ws, size_t ws_bytes,
                                   int n, int c, int hw, int relu, void* stream) {
    if (!d_x || !d_dy || !d_gamma || !d_save_mean || !d_save_rstd || !d_dx || !d_dgamma || !d_dbeta || (relu && !d_beta))
        return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(c, 1, 1));
    if (rc) return rc;
    cudaError_t e = launch_train_bn_bwd(d_x, d_dy, d_gamma, d_beta, d_save_mean, d_save_rstd, d_dx, d_dgamma, d_dbeta, d_ws, n, c, hw, relu,
                                        (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "batchnorm backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_ohem_workspace_bytes(size_t* out) {
    if (!out) return fail(FSCNN_EINVAL, "bad argument");
    *out = train_ohem_workspace_bytes();
    return FSCNN_OK;
}

int fscnn_train_ohem_forward(const float* d_logits, const long long* d_label, const float* d_class_weight, float* d_prob, float* d_out3,
                             void* d_ws, size_t ws_bytes, int n, int c, int hw, long long ignore_label, float thresh, int min_kept,
                             void* stream) {
    if (!d_logits || !d_label || !d_prob || !d_out3) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || hw < 1 || min_kept < 0) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_ohem_workspace_bytes());
    if (rc) return rc;
    cudaError_t e = launch_train_ohem_fwd(d_logits, d_label, d_class_weight, d_prob, d_out3, d_ws, n, c, hw, ignore_label, thresh, min_kept,
                                          (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "OHEM forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_ohem_backward(const float* d_logits, const long long* d_label, const float* d_class_weight, const float* d_prob,
                              const float* d_out3, const float* d_grad_out, float* d_dlogits, const void* d_ws, int n, int c, int hw,
                              long long ignore_label, void* stream) {
    if (!d_logits || !d_label || !d_prob || !d_out3 || !d_grad_out || !d_dlogits || !d_ws) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad shape");
    cudaError_t e = launch_train_ohem_bwd(d_logits, d_label, d_class_weight, d_prob, d_out3, d_grad_out, d_dlogits, d_ws, n, c, hw,
                                          ignore_label, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "OHEM backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_ohem_upsampled_forward(const float* d_low_logits, const long long* d_label, const float* d_class_weight, float* d_prob,
                                       float* d_nll, float* d_out3, void* d_ws, size_t ws_bytes, int n, int c, int hl, int wl, int h, int w,
                                       long long ignore_label, float thresh, int min_kept, void* stream) {
    if (!d_low_logits || !d_label || !d_prob || !d_out3) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || hl < 1 || wl < 1 || h < 1 || w < 1 || min_kept < 0) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_ohem_workspace_bytes());
    if (rc) return rc;
    cudaError_t e = launch_train_ohem_up_fwd(d_low_logits, d_label, d_class_weight, d_prob, d_nll, d_out3, d_ws, n, c, hl, wl, h, w, ignore_label,
                                             thresh, min_kept, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "fused upsample + OHEM forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_ohem_upsampled_backward(const float* d_low_logits, const long long* d_label, const float* d_class_weight, const float* d_prob,
                                        const float* d_out3, const float* d_grad_out, float* d_dlow, const void* d_ws, int n, int c, int hl,
                                        int wl, int h, int w, long long ignore_label, void* stream) {
    if (!d_low_logits || !d_label || !d_prob || !d_out3 || !d_grad_out || !d_dlow || !d_ws) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || hl < 1 || wl < 1 || h < 1 || w < 1) return fail(FSCNN_EINVAL, "bad shape");
    if ((double)(hl - 1) * 7.0 > (double)(h - 1) || (double)(wl - 1) * 7.0 > (double)(w - 1) || c > 128)
        return fail(FSCNN_EINVAL, "the fused backward needs an upsampling ratio >= 7 and <= 128 classes (%dx%d -> %dx%d, %d classes)", hl, wl, h, w, c);
    cudaError_t e = launch_train_ohem_up_bwd(d_low_logits, d_label, d_class_weight, d_prob, d_out3, d_grad_out, d_dlow, d_ws, n, c, hl, wl, h, w,
                                             ignore_label, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "fused upsample + OHEM backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_criterion_workspace_bytes(size_t* out) {
    if (!out) return fail(FSCNN_EINVAL, "bad argument");
    *out = train_criterion_workspace_bytes();
    return FSCNN_OK;
}

static int criterion_args_ok(int kind, int n, int c, int hl, int wl, int h, int w, float gamma) {
    if (kind < FSCNN_CRITERION_CE || kind > FSCNN_CRITERION_FOCAL_DICE) return fail(FSCNN_EINVAL, "unknown criterion %d", kind);
    if (n < 1 || c < 1 || hl < 1 || wl < 1 || h < 1 || w < 1 || hl > h || wl > w) return fail(FSCNN_EINVAL, "bad shape");
    if (kind == FSCNN_CRITERION_FOCAL_DICE && !(gamma >= 1.f)) return fail(FSCNN_EINVAL, "focal gamma %g < 1 is not supported", (double)gamma);
    return FSCNN_OK;
}

int fscnn_train_criterion_forward(const float* d_logits, const long long* d_label, double* d_out6, void* d_ws, size_t ws_bytes, int kind,
                                  int n, int c, int hl, int wl, int h, int w, long long ignore_label, float smooth, float alpha, float gamma,
                                  float dice_weight, void* stream) {
    if (!d_logits || !d_label || !d_out6) return fail(FSCNN_EINVAL, "null device pointer");
    int rc = criterion_args_ok(kind, n, c, hl, wl, h, w, gamma);
    if (rc) return rc;
    rc = train_ws_ok(d_ws, ws_bytes, train_criterion_workspace_bytes());
    if (rc) return rc;
    cudaError_t e = launch_train_criterion_fwd(d_logits, d_label, d_out6, d_ws, kind, n, c, hl, wl, h, w, ignore_label, smooth, alpha, gamma,
                                               dice_weight, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "criterion forward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

int fscnn_train_criterion_backward(const float* d_logits, const long long* d_label, const double* d_out6, const float* d_grad_out,
                                   float* d_dlogits, int kind, int n, int c, int hl, int wl, int h, int w, long long ignore_label, float smooth,
                                   float alpha, float gamma, float dice_weight, void* stream) {
    if (!d_logits || !d_label || !d_out6 || !d_grad_out || !d_dlogits) return fail(FSCNN_EINVAL, "null device pointer");
    int rc = criterion_args_ok(kind, n, c, hl, wl, h, w, gamma);
    if (rc) return rc;
    if ((hl != h || wl != w) && ((double)(hl - 1) * 7.0 > (double)(h - 1) || (double)(wl - 1) * 7.0 > (double)(w - 1) || c > 128))
        return fail(FSCNN_EINVAL, "the fused backward needs an upsampling ratio >= 7 and <= 128 classes (%dx%d -> %dx%d, %d classes)", hl, wl, h, w, c);
    cudaError_t e = launch_train_criterion_bwd(d_logits, d_label, d_out6, d_grad_out, d_dlogits, kind, n, c, hl, wl, h, w, ignore_label, smooth,
                                               alpha, gamma, dice_weight, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FSCNN_ECUDA, "criterion backward launch failed: %s", cudaGetErrorString(e));
    return FSCNN_OK;
}

#define FSCNN_TRAIN_CALL(expr, what)                                                                       \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess) return fail(FSCNN_ECUDA, what " launch failed: %s", cudaGetErrorString(e_)); \
        return FSCNN_OK;                                                                                   \
    } while (0)

int fscnn_train_im2col3x3(const float* d_x, float* d_cols, int n, int c, int h, int w, int stride, int pad, void* stream) {
    if (!d_x || !d_cols) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || (stride != 1 && stride != 2) || (pad != 0 && pad != 1) || h + 2 * pad < 3 || w + 2 * pad < 3) return fail(FSCNN_EINVAL, "bad shape");
    FSCNN_TRAIN_CALL(launch_train_im2col(d_x, d_cols, n, c, h, w, stride, pad, (cudaStream_t)stream), "im2col");
}
int fscnn_train_col2im3x3(const float* d_dcols, float* d_dx, int n, int c, int h, int w, int stride, int pad, void* stream) {
    if (!d_dcols || !d_dx) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || c < 1 || (stride != 1 && stride != 2) || (pad != 0 && pad != 1) || h + 2 * pad < 3 || w + 2 * pad < 3) return fail(FSCNN_EINVAL, "bad shape");
    FSCNN_TRAIN_CALL(launch_train_col2im(d_dcols, d_dx, n, c, h, w, stride, pad, (cudaStream_t)stream), "col2im");
}
int fscnn_train_bias_add(float* d_y, const float* d_bias, int n, int c, int hw, void* stream) {
    if (!d_y || !d_bias || n < 1 || c < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_bias_add(d_y, d_bias, n, c, hw, (cudaStream_t)stream), "bias add");
}
int fscnn_train_bias_grad(const float* d_dy, float* d_dbias, void* d_ws, size_t ws_bytes, int n, int c, int hw, void* stream) {
    if (!d_dy || !d_dbias || n < 1 || c < 1 || hw < 1) return fail(FSCNN_EINVAL, "bad argument");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(c, 1, 1));
    if (rc) return rc;
    FSCNN_TRAIN_CALL(launch_train_bias_grad(d_dy, d_dbias, d_ws, n, c, hw, (cudaStream_t)stream), "bias gradient");
}
int fscnn_train_bilinear(const float* d_in, float* d_out, int planes, int hi, int wi, int ho, int wo, int backward, void* stream) {
    if (!d_in || !d_out || planes < 1 || hi < 1 || wi < 1 || ho < 1 || wo < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_bilinear(d_in, d_out, planes, hi, wi, ho, wo, backward, (cudaStream_t)stream), "bilinear resize");
}
int fscnn_train_adaptive_avg_pool(const float* d_in, float* d_out, int planes, int h, int w, int bins, int backward, void* stream) {
    if (!d_in || !d_out || planes < 1 || h < 1 || w < 1 || bins < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_adaptive_pool(d_in, d_out, planes, h, w, bins, backward, (cudaStream_t)stream), "adaptive pool");
}
int fscnn_train_dropout(const float* d_x, float* d_y, float p, unsigned long long seed, const unsigned long long* d_step, int64_t numel,
                        void* stream) {
    if (!d_x || !d_y || numel < 1 || !(p >= 0.f && p < 1.f)) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_dropout(d_x, d_y, p, seed, d_step, numel, (cudaStream_t)stream), "dropout");
}
int fscnn_train_add_relu(const float* d_a, const float* d_b, float* d_y, int relu, int64_t numel, void* stream) {
    if (!d_a || !d_b || !d_y || numel < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_add_relu(d_a, d_b, d_y, relu, numel, (cudaStream_t)stream), "add");
}
int fscnn_train_relu_backward(const float* d_y, const float* d_dy, float* d_dx, int64_t numel, void* stream) {
    if (!d_y || !d_dy || !d_dx || numel < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_relu_bwd(d_y, d_dy, d_dx, numel, (cudaStream_t)stream), "relu backward");
}
int fscnn_train_stem_forward(const float* d_x, const float* d_w, float* d_y, int n, int h, int w, void* stream) {
    if (!d_x || !d_w || !d_y) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || h < 3 || w < 3) return fail(FSCNN_EINVAL, "bad shape");
    FSCNN_TRAIN_CALL(launch_train_stem_fwd(d_x, d_w, d_y, n, h, w, (cudaStream_t)stream), "stem forward");
}
int fscnn_train_stem_weight_grad(const float* d_x, const float* d_dy, float* d_dw, void* d_ws, size_t ws_bytes, int n, int h, int w,
                                 void* stream) {
    if (!d_x || !d_dy || !d_dw) return fail(FSCNN_EINVAL, "null device pointer");
    if (n < 1 || h < 3 || w < 3) return fail(FSCNN_EINVAL, "bad shape");
    int rc = train_ws_ok(d_ws, ws_bytes, train_workspace_bytes(32, 1, 1));
    if (rc) return rc;
    FSCNN_TRAIN_CALL(launch_train_stem_wgrad(d_x, d_dy, d_dw, d_ws, n, h, w, (cudaStream_t)stream), "stem weight gradient");
}

int fscnn_train_set_math(int mode) {
    if (train_set_math(mode)) return fail(FSCNN_EINVAL, "math mode must be 0 (fp32) or 1 (TF32)");
    return FSCNN_OK;
}
int fscnn_train_get_math(void) { return train_get_math(); }

int fscnn_train_sgd_step(float* d_param, const float* d_grad, float* d_momentum_buf, float lr, float momentum, float weight_decay,
                         float grad_scale, int first_step, int64_t numel, void* stream) {
    if (!d_param || !d_grad || !d_momentum_buf || numel < 1) return fail(FSCNN_EINVAL, "bad argument");
    FSCNN_TRAIN_CALL(launch_train_sgd(d_param, d_grad, d_momentum_buf, lr, momentum, weight_decay, grad_scale, first_step, numel,
                                      (cudaStream_t)stream), "SGD");
}

int fscnn_train_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, float lr, double beta1, double beta2,
                           float eps, float weight_decay, float grad_scale, int64_t step, int64_t numel, void* stream) {
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq || numel < 1 || step < 1) return fail(FSCNN_EINVAL, "bad argument");
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0)) return fail(FSCNN_EINVAL, "betas must lie in [0, 1)");
    FSCNN_TRAIN_CALL(launch_train_adamw(d_param, d_grad, d_exp_avg, d_exp_avg_sq, lr, beta1, beta2, eps, weight_decay, grad_scale, step, numel,
                                        (cudaStream_t)stream), "AdamW");
}

}  // extern "C"
