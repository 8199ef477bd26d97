// Topology, memory plan and the forward / backward schedules of the AMS student network.
// Replaces the TF1 graph executor behind tf.Session.run for the subgraph features -> student_logits of
// checkpoints/*/model.meta plus the head/loss/optimizer nodes added by utils/graph_utils.py:338-533.
#include "net.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

namespace ams {

namespace {

struct BlockSpec { int t, c, s, d; };
// MobileNetV2 (depth multiplier 1, output stride 16: stride of block 13 turned into dilation 2 for blocks 14-16),
// as decoded from the shipped model.meta.
const BlockSpec kBlocks[17] = {
    {1, 16, 1, 1},  {6, 24, 2, 1},  {6, 24, 1, 1},  {6, 32, 2, 1},  {6, 32, 1, 1},  {6, 32, 1, 1},
    {6, 64, 2, 1},  {6, 64, 1, 1},  {6, 64, 1, 1},  {6, 64, 1, 1},  {6, 96, 1, 1},  {6, 96, 1, 1},
    {6, 96, 1, 1},  {6, 160, 1, 1}, {6, 160, 1, 2}, {6, 160, 1, 2}, {6, 320, 1, 2}};

void same_pad(int in, int k, int s, int d, int* out, int* pad_before) {
    *out = (in + s - 1) / s;
    const int total = std::max((*out - 1) * s + (k - 1) * d + 1 - in, 0);
    *pad_before = total / 2;
}

template <typename T>
int dev_alloc(T** p, size_t count, std::vector<void*>* track = nullptr) {
    void* q = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    *p = static_cast<T*>(q);
    if (track) track->push_back(q);
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- profiler
cudaEvent_t Profiler::get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void Profiler::begin(cudaStream_t s, const char* tag, double algo_bytes) {
    if (!enabled) return;
    ProfEntry e;
    e.tag = tag; e.algo_bytes = algo_bytes;
    if (per_layer && layer) { e.tag += "@"; e.tag += layer; } e.e0 = get_event(); e.e1 = get_event();
    cudaEventRecord(e.e0, s);
    pending.push_back(e);
}
void Profiler::end(cudaStream_t s) {
    if (!enabled || pending.empty()) return;
    cudaEventRecord(pending.back().e1, s);
}
void Profiler::collect() {
    for (ProfEntry& e : pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.e0, e.e1) == cudaSuccess) {
            if (!agg.count(e.tag)) order.push_back(e.tag);
            ProfAgg& a = agg[e.tag];
            a.launches += 1; a.ms += ms; a.algo_bytes += e.algo_bytes;
        }
        pool.push_back(e.e0); pool.push_back(e.e1);
    }
    pending.clear();
}
void Profiler::reset() { collect(); agg.clear(); order.clear(); }

#define PROF(tag, bytes, call)                                   \
    do {                                                         \
        net->prof.begin(s, tag, static_cast<double>(bytes));     \
        const int _rc = (call);                                  \
        net->prof.end(s);                                        \
        if (_rc) return -1;                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------- topology
int net_build_topology(Net* net) {
    const ams_config& c = net->cfg;
    AMS_REQUIRE(c.num_classes > 0 && c.num_classes <= 21, "num_classes must be in [1,21]");
    const float eps_backbone = 0.001f, eps_aspp = 1.001e-05f;
    const float k_default = 1.0f - 0.9f;                     // graph computes 1 - decay in fp32
    const float k_dw = c.graph_variant == 1 ? (1.0f - 0.98f) : k_default;   // VOC graph: depthwise BN decay 0.98
    auto& L = net->layers;
    L.clear();
    net->vars.clear();
    long long t_off = 0, m_off = 0, bn_off = 0, w16_off = 0;

    auto add_var = [&](const std::string& name, std::initializer_list<int> shape, bool trainable) -> long long {
        VarInfo v;
        v.name = name;
        v.ndim = static_cast<int>(shape.size());
        long long cnt = 1;
        int i = 0;
        for (int s : shape) { v.shape[i++] = s; cnt *= s; }
        v.count = cnt;
        v.trainable = trainable;
        long long& off = trainable ? t_off : m_off;
        v.offset = off;
        off += cnt;
        net->var_index[name] = static_cast<int>(net->vars.size());
        if (trainable) net->trainable_order.push_back(static_cast<int>(net->vars.size()));
        net->vars.push_back(v);
        return v.offset;
    };
    auto add_layer = [&](const std::string& name, int kind, int cin, int cout, int stride, int dil, int act, bool bn,
                         float eps, float k, int input, int residual) -> int {
        LayerDef d;
        d.name = name; d.kind = kind; d.cin = cin; d.cout = cout; d.stride = stride; d.dil = dil; d.act = act;
        d.has_bn = bn; d.eps = eps; d.one_minus_decay = k; d.input = input; d.residual = residual;
        const std::string wname = (kind == kDepthwise) ? name + "/depthwise_weights:0" : name + "/weights:0";
        if (kind == kStem) d.w_off = add_var(wname, {3, 3, cin, cout}, true);
        else if (kind == kDepthwise) d.w_off = add_var(wname, {3, 3, cin, 1}, true);
        else d.w_off = add_var(wname, {1, 1, cin, cout}, true);
        if (bn) {
            d.gamma_off = add_var(name + "/BatchNorm/gamma:0", {cout}, true);
            d.beta_off = add_var(name + "/BatchNorm/beta:0", {cout}, true);
            d.mm_off = add_var(name + "/BatchNorm/moving_mean:0", {cout}, false);
            d.mv_off = add_var(name + "/BatchNorm/moving_variance:0", {cout}, false);
            d.bn_off = bn_off;
            bn_off += 6LL * cout;
        } else {
            d.bias_off = add_var(name + "/biases:0", {cout}, true);
        }
        L.push_back(d);
        return static_cast<int>(L.size()) - 1;
    };

    int prev = add_layer("MobilenetV2/Conv", kStem, 3, 32, 2, 1, 2, true, eps_backbone, k_default, -1, -1);
    int cin = 32;
    for (int b = 0; b < 17; ++b) {
        const BlockSpec& bs = kBlocks[b];
        const std::string base = "MobilenetV2/expanded_conv" + (b == 0 ? std::string("") : "_" + std::to_string(b));
        const int block_in = prev;
        int cur = prev, cexp = cin * bs.t;
        if (bs.t != 1) cur = add_layer(base + "/expand", kConv1x1, cin, cexp, 1, 1, 2, true, eps_backbone, k_default, cur, -1);
        // graph: depthwise node is ".../depthwise/depthwise", variable scope ".../depthwise"
        cur = add_layer(base + "/depthwise", kDepthwise, cexp, cexp, bs.s, bs.d, 2, true, eps_backbone, k_dw, cur, -1);
        const bool res = (bs.s == 1 && cin == bs.c);
        cur = add_layer(base + "/project", kConv1x1, cexp, bs.c, 1, 1, 0, true, eps_backbone, k_default, cur, res ? block_in : -1);
        prev = cur;
        cin = bs.c;
    }
    const int feat = prev;
    const int ip = add_layer("image_pooling", kImagePool, 320, 256, 1, 1, 1, true, eps_aspp, k_default, feat, -1);
    const int aspp = add_layer("aspp0", kConv1x1, 320, 256, 1, 1, 1, true, eps_aspp, k_default, feat, -1);
    const int cp = add_layer("concat_projection", kConv1x1, 512, 256, 1, 1, 1, true, eps_aspp, k_default, aspp, -1);
    add_layer("logits/semantic", kLogits, 256, c.num_classes, 1, 1, 0, false, 0.f, 0.f, cp, -1);
    (void)ip;
    // gradient buckets (data parallel): the late bucket starts at the first expand conv that runs at the final resolution
    for (size_t i = 0; i < L.size(); ++i)
        if (L[i].name == "MobilenetV2/expanded_conv_7/expand") { net->bucket_layer = static_cast<int>(i); net->bucket_split = L[i].w_off; }
    net->n_train = t_off;
    net->n_moving = m_off;
    net->n_bnpool = bn_off;

    // geometry + bf16 weight copies
    const int Hp = c.height + 1, Wp = c.width + 1;            // graph pads 1 px bottom/right with the mean pixel
    for (size_t i = 0; i < L.size(); ++i) {
        LayerDef& d = L[i];
        if (d.kind == kStem) { d.in_h = Hp; d.in_w = Wp; }
        else if (d.kind == kImagePool) { d.in_h = 1; d.in_w = 1; }
        else { d.in_h = L[d.input].out_h; d.in_w = L[d.input].out_w; }
        if (d.kind == kStem || d.kind == kDepthwise) {
            same_pad(d.in_h, 3, d.stride, d.dil, &d.out_h, &d.pad_top);
            same_pad(d.in_w, 3, d.stride, d.dil, &d.out_w, &d.pad_left);
        } else { d.out_h = d.in_h; d.out_w = d.in_w; }
        if (d.kind == kConv1x1 || d.kind == kLogits) {
            d.k_rows0 = (d.name == "concat_projection") ? 256 : 0;
            d.k_rows = d.cin - d.k_rows0;
            d.ld_fwd = d.k_rows;                                   // [cout][k_rows]
            d.ld_bwd = (d.kind == kLogits) ? 32 : d.cout;          // [k_rows][cout] (logits: K padded to 32)
            d.wfwd_off = w16_off; w16_off += static_cast<long long>(d.cout) * d.ld_fwd;
            w16_off = (w16_off + 63) & ~63LL;
            d.wbwd_off = w16_off; w16_off += static_cast<long long>(d.k_rows) * d.ld_bwd;
            w16_off = (w16_off + 63) & ~63LL;
            if (d.cout <= kSplitWeightMaxCout) {
                d.wlo_off = w16_off; w16_off += static_cast<long long>(d.cout) * d.ld_fwd;
                w16_off = (w16_off + 63) & ~63LL;
            }
        }
    }
    net->n_bf16 = w16_off;
    return 0;
}

// ---------------------------------------------------------------------------------------------- plan
static BnLayer bn_layer(Net* net, const LayerDef& d, long long M) {
    BnLayer b;
    b.C = d.cout; b.M = M; b.eps = d.eps; b.one_minus_decay = d.one_minus_decay;
    b.gamma = net->params + d.gamma_off; b.beta = net->params + d.beta_off;
    b.moving_mean = net->moving + d.mm_off; b.moving_var = net->moving + d.mv_off;
    float* pool = net->bnpool + d.bn_off;
    b.scale = pool; b.shift = pool + d.cout; b.mean = pool + 2 * d.cout; b.rstd = pool + 3 * d.cout;
    if (net->sync_active) {
        // 4 words per channel and direction; layers laid out in bnpool order, backward region after the forward one
        b.sync = net->syncbn_dev;
        b.xoff_fwd = d.bn_off / 6 * 4;
        b.xoff_bwd = net->n_bnpool / 6 * 4 + b.xoff_fwd;
    }
    return b;
}
static float* fscale(Net* net, const LayerDef& d) { return net->bnpool + d.bn_off + 4 * d.cout; }
static float* fshift(Net* net, const LayerDef& d) { return net->bnpool + d.bn_off + 5 * d.cout; }

static int build_plan(Net* net, Plan* p, bool need_backward);

Plan* net_get_plan(Net* net, int N, bool need_backward) {
    auto it = net->plans.find(N);
    if (it == net->plans.end()) {
        std::unique_ptr<Plan> p(new Plan());
        p->N = N;
        if (build_plan(net, p.get(), need_backward)) return nullptr;
        it = net->plans.emplace(N, std::move(p)).first;
    } else if (need_backward && !it->second->have_backward) {
        if (build_plan(net, it->second.get(), true)) return nullptr;
    }
    return it->second.get();
}

int net_half_plans(Net* net, Plan* p, Plan** a, Plan** b) {
    if (p->N < 2 || (p->N & 1)) { set_last_error("split inference needs an even batch"); return -1; }
    for (int k = 0; k < 2; ++k) {
        if (p->half[k]) continue;
        std::unique_ptr<Plan> h(new Plan());
        h->N = p->N / 2; h->parent = p; h->part = k;
        if (build_plan(net, h.get(), false)) return -1;
        p->half[k] = std::move(h);
    }
    *a = p->half[0].get(); *b = p->half[1].get();
    return 0;
}

static int build_plan(Net* net, Plan* p, bool need_backward) {
    const auto& L = net->layers;
    const int N = p->N;
    const ams_config& c = net->cfg;
    const bool first = p->buf.empty();
    if (first) {
        p->buf.resize(L.size());
        p->fwd_frozen.resize(L.size()); p->fwd_train.resize(L.size()); p->dgrad.resize(L.size()); p->wgrad.resize(L.size());
        p->has_fwd.assign(L.size(), 0); p->has_dgrad.assign(L.size(), 0); p->has_wgrad.assign(L.size(), 0);
        p->dw_fused.assign(L.size(), 0); p->lazy_y.assign(L.size(), 0);
        const char* nofuse = getenv("AMS_NO_DW_FUSION");
        for (size_t i = 0; i < L.size(); ++i) {
            const LayerDef& d = L[i];
            if (d.kind != kDepthwise || (nofuse && nofuse[0] == '1')) continue;
            Conv2dGeom g{N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
            if (!dw_tiled_supported(g) || !L[d.input].has_bn) continue;
            p->dw_fused[i] = 1;
            p->lazy_y[d.input] = 1;
        }
        if (dev_alloc(&p->coef[0], 3 * 1024, &p->allocations)) return -1;
        if (dev_alloc(&p->coef[1], 3 * 1024, &p->allocations)) return -1;
        // a half-batch view (p->parent) shares the parent's per-image buffers at image offset part * N
        const Plan* par = p->parent;
        const size_t img0 = par ? static_cast<size_t>(p->part) * N : 0;
        const size_t px_img = static_cast<size_t>(c.height) * c.width;
        if (par) {
            p->in_labels = par->in_labels + img0 * px_img;
            p->pred = par->pred + img0 * px_img;            // in_frames depends on the frame dtype: set per call (infer_common)
        } else {
            if (dev_alloc(reinterpret_cast<float**>(&p->in_frames), static_cast<size_t>(N) * px_img * 3, &p->allocations)) return -1;
            if (dev_alloc(&p->in_labels, static_cast<size_t>(N) * px_img, &p->allocations)) return -1;
            if (dev_alloc(&p->pred, static_cast<size_t>(N) * px_img, &p->allocations)) return -1;
        }
        if (dev_alloc(&p->loss_dev, 4, &p->allocations)) return -1;
        size_t bn_ws = 0;
        for (size_t i = 0; i < L.size(); ++i) {
            const LayerDef& d = L[i];
            if (d.kind == kImagePool) continue;
            const long long M = static_cast<long long>(N) * d.out_h * d.out_w;
            if (d.kind == kLogits) {
                if (par) { p->logits = par->logits + static_cast<size_t>(M) * p->part * 32; continue; }
                if (dev_alloc(&p->logits, static_cast<size_t>(M) * 32, &p->allocations)) return -1;
                AMS_CUDA_CHECK(cudaMemset(p->logits, 0, static_cast<size_t>(M) * 32 * sizeof(float)));
                continue;
            }
            if (par) {
                p->buf[i].y = par->buf[i].y + static_cast<size_t>(M) * p->part * d.cout;
                p->buf[i].z = par->buf[i].z + static_cast<size_t>(M) * p->part * d.cout;
            } else {
                if (dev_alloc(&p->buf[i].y, static_cast<size_t>(M) * d.cout, &p->allocations)) return -1;
                if (dev_alloc(&p->buf[i].z, static_cast<size_t>(M) * d.cout, &p->allocations)) return -1;
            }
            bn_ws = std::max(bn_ws, bn_workspace_doubles(M, d.cout));
            if (d.kind == kDepthwise) {
                Conv2dGeom g{N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
                bn_ws = std::max(bn_ws, static_cast<size_t>(dw_tiled_stats_rows(g)) * 2 * d.cout + 3 * static_cast<size_t>(d.cout));
                bn_ws = std::max(bn_ws, static_cast<size_t>(dw_bwd_fused_rows(g)) * 2 * d.cout + 3 * static_cast<size_t>(d.cout));
            }
        }
        if (dev_alloc(&p->bn_ws, bn_ws, &p->allocations)) return -1;
        const LayerDef& ipd = L[L.size() - 4];
        if (dev_alloc(&p->pooled, static_cast<size_t>(N) * ipd.cin, &p->allocations)) return -1;
        if (dev_alloc(&p->ip_z, static_cast<size_t>(N) * ipd.cout, &p->allocations)) return -1;
        if (dev_alloc(&p->ip_act, static_cast<size_t>(N) * ipd.cout, &p->allocations)) return -1;
        if (dev_alloc(&p->bias_img, static_cast<size_t>(N) * 256, &p->allocations)) return -1;
        if (dev_alloc(&p->ip_dbias, static_cast<size_t>(N) * (256 + ipd.cout), &p->allocations)) return -1;
        if (dev_alloc(&p->dfeat_rowbias, static_cast<size_t>(N) * ipd.cin, &p->allocations)) return -1;
        if (dev_alloc(&p->small_ws, colsum_workspace_doubles(std::max(N, 1), 512), &p->allocations)) return -1;
        // forward GEMM plans
        for (size_t i = 0; i < L.size(); ++i) {
            const LayerDef& d = L[i];
            if (d.kind != kConv1x1 && d.kind != kLogits) continue;
            const long long M = static_cast<long long>(N) * d.out_h * d.out_w;
            GemmDesc g;
            g.A = p->buf[d.input].y; g.lda = L[d.input].cout;
            g.B = net->wpool + d.wfwd_off; g.ldb = d.ld_fwd;
            if (d.wlo_off >= 0) g.B_lo = net->wpool + d.wlo_off;
            g.M = static_cast<int>(M); g.N = d.cout; g.K = d.k_rows;
            if (d.name == "concat_projection") { g.rowbias = p->bias_img; g.rows_per_image = d.out_h * d.out_w; }
            if (d.kind == kLogits) {
                g.out = p->logits; g.ldc = 32; g.out_fp32 = 1; g.shift = net->params + d.bias_off;
                if (gemm_plan(g, net->num_sms, &p->fwd_frozen[i])) return -1;
                p->fwd_train[i] = p->fwd_frozen[i];
            } else {
                GemmDesc f = g;
                f.out = p->buf[i].y; f.ldc = d.cout; f.scale = fscale(net, d); f.shift = fshift(net, d); f.act = d.act;
                if (d.residual >= 0) { f.residual = p->buf[d.residual].y; f.ldr = d.cout; }
                if (gemm_plan(f, net->num_sms, &p->fwd_frozen[i])) return -1;
                GemmDesc t = g;
                t.out = p->buf[i].z; t.ldc = d.cout;
                t.stats_partial = p->bn_ws;                  // BatchNorm batch statistics fused into the epilogue
                if (gemm_plan(t, net->num_sms, &p->fwd_train[i])) return -1;
            }
            p->has_fwd[i] = 1;
        }
        // block-fused frozen inference: expand (i) -> depthwise (i+1) -> project (i+2) of every stride-1 block
        p->fused_plan.resize(L.size()); p->fused.assign(L.size(), 0);
        for (size_t i = 0; i + 2 < L.size(); ++i) {
            if (i >= net->fused_params.size() || !net->fused_params[i]) continue;
            const LayerDef& e = L[i]; const LayerDef& dw = L[i + 1]; const LayerDef& pr = L[i + 2];
            FusedBlockDesc f;
            f.N = N; f.H = e.out_h; f.W = e.out_w; f.Cin = e.cin; f.Cexp = e.cout; f.Cout = pr.cout; f.stride = dw.stride; f.dil = dw.dil;
            f.pad_top = dw.pad_top; f.pad_left = dw.pad_left;
            f.x = p->buf[e.input].y;
            f.We = net->wpool + e.wfwd_off; f.We_lo = e.wlo_off >= 0 ? net->wpool + e.wlo_off : nullptr; f.ld_we = e.ld_fwd;
            f.Wp = net->wpool + pr.wfwd_off; f.Wp_lo = pr.wlo_off >= 0 ? net->wpool + pr.wlo_off : nullptr; f.ld_wp = pr.ld_fwd;
            f.params = net->fused_params[i];
            f.s3 = fscale(net, pr); f.t3 = fshift(net, pr);
            f.residual = pr.residual >= 0 ? p->buf[pr.residual].y : nullptr;
            f.out = p->buf[i + 2].y;
            if (fused_block_plan(f, net->num_sms, &p->fused_plan[i])) return -1;
            p->fused[i] = 1;
        }
    }
    if (need_backward && !p->have_backward) {
        size_t red_ws = 0, wg_ws = 0;
        const LayerDef& lg = L.back();
        const long long M16 = static_cast<long long>(N) * lg.out_h * lg.out_w;
        if (dev_alloc(&p->dlogits_f32, static_cast<size_t>(M16) * 32, &p->allocations)) return -1;
        if (dev_alloc(&p->dlogits_bf16, static_cast<size_t>(M16) * 32, &p->allocations)) return -1;
        HeadGeom hg = net->head; hg.N = N;
        if (dev_alloc(&p->rowbuf, head_rowbuf_floats(hg), &p->allocations)) return -1;
        for (size_t i = 0; i < L.size(); ++i) {
            const LayerDef& d = L[i];
            if (d.kind == kImagePool || d.kind == kLogits) continue;
            const long long M = static_cast<long long>(N) * d.out_h * d.out_w;
            if (dev_alloc(&p->buf[i].g, static_cast<size_t>(M) * d.cout, &p->allocations)) return -1;
            if (d.residual >= 0) { if (dev_alloc(&p->buf[i].gz, static_cast<size_t>(M) * d.cout, &p->allocations)) return -1; }
            else p->buf[i].gz = p->buf[i].g;
            if (d.kind == kDepthwise) {
                Conv2dGeom g{N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
                red_ws = std::max(red_ws, dw_bwd_workspace_floats(g));
                red_ws = std::max(red_ws, static_cast<size_t>(dw_bwd_fused_rows(g)) * 9 * d.cin);
            }
            if (d.kind == kStem) red_ws = std::max(red_ws, stem_bwd_workspace_floats(N, d.out_h, d.out_w));
            if (d.kind == kConv1x1) wg_ws = std::max(wg_ws, wgrad_workspace_floats(d.k_rows, d.cout, M, net->num_sms));
        }
        wg_ws = std::max(wg_ws, wgrad_workspace_floats(lg.cin, lg.cout, M16, net->num_sms));
        if (dev_alloc(&p->red_ws, red_ws, &p->allocations)) return -1;
        p->red_ws_floats = red_ws;
        if (dev_alloc(&p->wgrad_ws, wg_ws, &p->allocations)) return -1;
        p->wgrad_ws_floats = wg_ws;
        for (size_t i = 0; i < L.size(); ++i) {
            const LayerDef& d = L[i];
            if (d.kind != kConv1x1 && d.kind != kLogits) continue;
            const long long M = static_cast<long long>(N) * d.out_h * d.out_w;
            // weight gradient: dW[k_rows][cout] = A^T * dz
            WgradDesc w;
            w.X = p->buf[d.input].y; w.ldx = L[d.input].cout; w.Cin = d.k_rows;
            if (d.kind == kLogits) { w.dZ = p->dlogits_bf16; w.ldz = 32; }
            else { w.dZ = p->buf[i].gz; w.ldz = d.cout; }
            w.Cout = d.cout; w.M = M;
            w.dW = net->grads + d.w_off + static_cast<long long>(d.k_rows0) * d.cout; w.lddw = d.cout;
            w.workspace = p->wgrad_ws; w.workspace_floats = p->wgrad_ws_floats;
            if (wgrad_plan(w, net->num_sms, &p->wgrad[i])) return -1;
            p->has_wgrad[i] = 1;
            // data gradient: g[input] = dz * W^T (+ skip gradient / image-pool gradient)
            GemmDesc g;
            if (d.kind == kLogits) { g.A = p->dlogits_bf16; g.lda = 32; g.K = 32; }
            else { g.A = p->buf[i].gz; g.lda = d.cout; g.K = d.cout; }
            g.a_fp16 = 0; g.b_fp16 = 0; g.out_fp16 = 0;      // bf16 gradient x bf16 transposed weights -> bf16 gradient
            g.B = net->wpool + d.wbwd_off; g.ldb = d.ld_bwd;
            g.M = static_cast<int>(M); g.N = d.k_rows;
            g.out = p->buf[d.input].g; g.ldc = L[d.input].cout;
            // expand conv of a residual block: the block input also receives the block output's gradient
            if (i + 2 < L.size() && L[i + 2].kind == kConv1x1 && L[i + 2].residual == d.input && L[i + 1].kind == kDepthwise) {
                g.residual = p->buf[i + 2].g; g.ldr = L[i + 2].cout;
            }
            if (d.name == "aspp0") { g.rowbias = p->dfeat_rowbias; g.rows_per_image = d.out_h * d.out_w; }
            if (gemm_plan(g, net->num_sms, &p->dgrad[i])) return -1;
            p->has_dgrad[i] = 1;
        }
        p->have_backward = true;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- weights
// Frozen (client) graph: trim_graph_frozen(kill_norms=True) replaces every FusedBatchNormV3 by the `_patch` op that
// tf.layers.batch_normalization(training=False) created (reference utils/graph_utils.py:362-369, :52-76), whose epsilon
// is the layer default 1e-3 for all 54 layers -- including image_pooling / aspp0 / concat_projection, which TRAIN with
// epsilon 1.001e-5.
constexpr float kFrozenBnEps = 1e-3f;
int net_prepare_weights(Net* net, bool frozen) {
    if (net->weights_dirty) {
        if (cast_weights(net->cast_table, net->cast_layers, net->cast_max, net->stream)) return -1;
        net->weights_dirty = false;
    }
    if (frozen && net->fold_dirty) {
        for (const LayerDef& d : net->layers) {
            if (!d.has_bn || d.kind == kImagePool) continue;
            if (bn_fold_frozen(net->params + d.gamma_off, net->params + d.beta_off, net->moving + d.mm_off,
                               net->moving + d.mv_off, kFrozenBnEps, fscale(net, d), fshift(net, d), d.cout, net->stream)) return -1;
        }
        for (size_t i = 0; i + 2 < net->layers.size() && i < net->fused_params.size(); ++i) {
            if (!net->fused_params[i]) continue;
            const LayerDef& e = net->layers[i]; const LayerDef& dw = net->layers[i + 1]; const LayerDef& pr = net->layers[i + 2];
            FusedBlockDesc f;
            f.Cin = e.cin; f.Cexp = e.cout; f.Cout = pr.cout; f.stride = dw.stride; f.dil = dw.dil; f.params = net->fused_params[i];
            if (fused_block_fill_params(f, fscale(net, e), fshift(net, e), net->params + dw.w_off, fscale(net, dw), fshift(net, dw),
                                        net->stream)) return -1;
        }
        net->fold_dirty = false;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- forward
static ImgPoolFwd imgpool_desc(Net* net, Plan* p, bool frozen, bool update_moving) {
    const auto& L = net->layers;
    const LayerDef& d = L[L.size() - 4];
    const LayerDef& cp = L[L.size() - 2];
    ImgPoolFwd a;
    a.N = p->N; a.HW = L[d.input].out_h * L[d.input].out_w; a.Cin = d.cin; a.Cmid = d.cout; a.Cout = cp.cout;
    a.feat = p->buf[d.input].y;
    a.w_pool = net->params + d.w_off;
    a.w_proj_top = net->params + cp.w_off;
    a.bn = bn_layer(net, d, p->N);
    a.frozen = frozen; a.update_moving = update_moving;
    if (frozen) a.bn.eps = kFrozenBnEps;
    a.pooled = p->pooled; a.z = p->ip_z; a.act = p->ip_act; a.bias_img = p->bias_img; a.ws = p->small_ws;
    return a;
}

int net_forward(Net* net, Plan* p, int bn_mode, bool update_moving, const StreamSet* ss) {
    const auto& L = net->layers;
    const ams_config& c = net->cfg;
    const bool frozen = (bn_mode == AMS_BN_MOVING);
    cudaStream_t s = ss ? ss->main : net->stream;
    cudaStream_t side_stream = ss ? ss->side : net->side_stream;
    cudaEvent_t ev_fork = ss ? ss->fork : net->ev_fork, ev_join = ss ? ss->join : net->ev_join;
    if (net_prepare_weights(net, frozen)) return -1;
    bool pool_pending = false;
    for (size_t i = 0; i < L.size(); ++i) {
        const LayerDef& d = L[i];
        net->prof.layer = d.name.c_str();
        const long long M = static_cast<long long>(p->N) * d.out_h * d.out_w;
        LayerBuf& b = p->buf[i];
        const double out_bytes = 2.0 * M * d.cout;
        if (d.kind == kImagePool) {
            // six tiny latency-bound kernels: off the chain (side stream) while aspp0 runs; joined before concat_projection
            if (!net->prof.enabled && side_stream) {
                AMS_CUDA_CHECK(cudaEventRecord(ev_fork, s));
                AMS_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
                if (imgpool_forward(imgpool_desc(net, p, frozen, update_moving), side_stream)) return -1;
                AMS_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
                pool_pending = true;
            } else {
                PROF("imgpool_fwd", 2.0 * p->N * L[d.input].out_h * L[d.input].out_w * d.cin, imgpool_forward(imgpool_desc(net, p, frozen, update_moving), s));
            }
            continue;
        }
        if (frozen && net->block_fusion && p->fused[i]) {
            // one kernel for expand -> depthwise -> project (+ skip): the 6C-wide tensors stay in TMEM / shared memory
            const LayerDef& pr = L[i + 2];
            const double px = static_cast<double>(p->N) * d.out_h * d.out_w;
            PROF("fused_block", 2.0 * px * (d.cin + pr.cout * (pr.residual >= 0 ? 2 : 1)) + 2.0 * (d.cin * d.cout + d.cout * pr.cout) + 36.0 * d.cout,
                 fused_block_launch(p->fused_plan[i], s));
            i += 2;
            continue;
        }
        if (pool_pending && d.name == "concat_projection") {
            AMS_CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
            pool_pending = false;
        }
        if (d.kind == kLogits) {
            PROF("gemm_logits", 2.0 * M * d.cin + 4.0 * M * 32, gemm_launch(p->fwd_frozen[i], s));
            continue;
        }
        act_t* conv_out = frozen ? b.y : b.z;
        int dw_rows = 0;
        const float* sc = frozen ? fscale(net, d) : nullptr;
        const float* sh = frozen ? fshift(net, d) : nullptr;
        if (d.kind == kStem) {
            const double in_bytes = static_cast<double>(p->N) * c.height * c.width * 3 * (p->in_dtype == AMS_FRAMES_U8 ? 1 : 4);
            PROF("stem_fwd", in_bytes + out_bytes,
                 stem_conv_fwd(p->in_frames, p->in_dtype == AMS_FRAMES_U8, p->N, c.height, c.width, d.in_h, d.in_w, d.out_h,
                               d.out_w, d.pad_top, d.pad_left, 127.5f, 0.007843137718737125f, 1.0f, net->params + d.w_off,
                               sc, sh, conv_out, s));
        } else if (d.kind == kDepthwise) {
            Conv2dGeom g{p->N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
            // training: batch statistics of the stored z come out of the same kernel (per-tile partials)
            const bool lazy_in = !frozen && p->lazy_y[d.input];      // producer's BN + act applied while staging its raw z
            const LayerDef& pd = L[d.input];
            const BnLayer pbl = bn_layer(net, pd, 0);
            PROF("dw_fwd", 2.0 * p->N * d.in_h * d.in_w * d.cin + out_bytes,
                 dw_conv_fwd_tiled(lazy_in ? p->buf[d.input].z : p->buf[d.input].y, net->params + d.w_off, g,
                                   lazy_in ? pbl.scale : nullptr, lazy_in ? pbl.shift : nullptr, pd.act, sc, sh, d.act, conv_out,
                                   frozen ? nullptr : p->bn_ws, &dw_rows, s));
        } else {
            const double gb = 2.0 * M * d.k_rows + 2.0 * d.k_rows * d.cout + out_bytes + ((frozen && d.residual >= 0) ? out_bytes : 0.0);
            PROF("gemm_fwd", gb, gemm_launch(frozen ? p->fwd_frozen[i] : p->fwd_train[i], s));
        }
        // TIMING EXPERIMENTS ONLY (results are wrong): upper bound of what removing a launch group would buy
        static const bool x_skip_fin = [] { const char* e = getenv("AMS_X_SKIP_FWD_FINALIZE"); return e && e[0] == '1'; }();
        static const bool x_skip_dwapply = [] { const char* e = getenv("AMS_X_SKIP_DW_BN_APPLY"); return e && e[0] == '1'; }();
        if (!frozen) {
            BnLayer bl = bn_layer(net, d, M);
            if (x_skip_fin && d.kind != kStem) { /* skipped */ }
            else if (d.kind == kConv1x1 && p->fwd_train[i].d.stats_partial)
                PROF("bn_finalize", 16.0 * p->fwd_train[i].grid * d.cout, bn_finalize_partials(p->bn_ws, p->fwd_train[i].grid, bl, update_moving ? 1 : 0, s));
            else if (d.kind == kDepthwise)
                PROF("bn_finalize", 16.0 * dw_rows * d.cout, bn_finalize_partials(p->bn_ws, dw_rows, bl, update_moving ? 1 : 0, s));
            else
                PROF("bn_stats", out_bytes, bn_forward_stats(b.z, bl, update_moving ? 1 : 0, p->bn_ws, s));
            if (!p->lazy_y[i] && !(x_skip_dwapply && d.kind == kDepthwise))
                PROF("bn_apply", (d.residual >= 0 ? 3.0 : 2.0) * out_bytes,
                     bn_apply(b.z, bl.scale, bl.shift, d.act, d.residual >= 0 ? p->buf[d.residual].y : nullptr, b.y, M, d.cout, s));
        }
    }
    p->last_was_train = !frozen;
    return 0;
}

// ---------------------------------------------------------------------------------------------- backward
int net_backward(Net* net, Plan* p, bool normalize) {
    const auto& L = net->layers;
    const ams_config& c = net->cfg;
    cudaStream_t s = net->stream;
    const int nl = static_cast<int>(L.size());
    p->pending_bn_rows = 0;
    // Filter gradients of the 1x1 convs: nothing downstream in the backward chain reads them, so outside profiling
    // they go to a side stream (fork after their dz is ready, one join at the end) and overlap with the chain.
    const bool side = !net->prof.enabled && net->side_stream != nullptr;
    bool forked = false, pool_bwd_pending = false, dw_reduce_pending = false;
    static const bool x_skip_wgrad = [] { const char* e = getenv("AMS_X_SKIP_WGRAD"); return e && e[0] == '1'; }();   // TIMING EXPERIMENT ONLY
    auto wgrad_on_side = [&](const WgradPlan& wp) -> int {
        if (x_skip_wgrad) return 0;
        AMS_CUDA_CHECK(cudaEventRecord(net->ev_fork, s));
        AMS_CUDA_CHECK(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
        forked = true;
        return wgrad_launch(wp, net->side_stream);
    };
    // head: loss + d low-res logits
    HeadGeom hg = net->head; hg.N = p->N; hg.normalize = normalize ? 1 : 0;
    const LayerDef& lg = L[nl - 1];
    const long long M16 = static_cast<long long>(p->N) * lg.out_h * lg.out_w;
    net->prof.layer = lg.name.c_str();
    PROF("head_loss_bwd", static_cast<double>(p->N) * c.height * c.width + 4.0 * M16 * 32 * 2,
         head_loss_backward(p->logits, hg, p->in_labels, p->rowbuf, p->dlogits_f32, p->dlogits_bf16, net->head_st, p->loss_dev, s));
    if (side) {
        AMS_CUDA_CHECK(cudaEventRecord(net->ev_fork, s));
        AMS_CUDA_CHECK(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
        forked = true;
        if (colsum_groups(p->dlogits_f32, nullptr, nullptr, 32, M16, 1, lg.cout, 1.f, net->grads + lg.bias_off, p->small_ws, net->side_stream)) return -1;
    } else {
        PROF("bias_grad", 4.0 * M16 * 32, colsum_groups(p->dlogits_f32, nullptr, nullptr, 32, M16, 1, lg.cout, 1.f, net->grads + lg.bias_off, p->small_ws, s));
    }
    if (side) { if (wgrad_on_side(p->wgrad[nl - 1])) return -1; }
    else PROF("gemm_wgrad", 2.0 * M16 * (lg.cin + 32), wgrad_launch(p->wgrad[nl - 1], s));
    PROF("gemm_dgrad", 2.0 * M16 * (lg.cin + 32), gemm_launch(p->dgrad[nl - 1], s));
    for (int i = nl - 2; i >= 0; --i) {
        const LayerDef& d = L[i];
        if (d.kind == kImagePool) continue;          // handled together with concat_projection
        net->prof.layer = d.name.c_str();
        const long long M = static_cast<long long>(p->N) * d.out_h * d.out_w;
        LayerBuf& b = p->buf[i];
        BnLayer bl = bn_layer(net, d, M);
        const double tb = 2.0 * M * d.cout;
        if (d.kind == kDepthwise && p->dw_fused[i]) {
            // BN-backward column sums of this layer, then ONE kernel: BN-backward apply while staging (g, z), filter +
            // data gradient, the producer's activation mask and the column sums of ITS BN backward
            const LayerDef& pd = L[d.input];
            const BnLayer pbl = bn_layer(net, pd, 0);
            Conv2dGeom g{p->N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
            PROF("bn_bwd_reduce", 2.0 * tb, bn_backward_reduce(b.g, b.z, bl, d.act, p->coef[0], net->grads + d.gamma_off, net->grads + d.beta_off, p->bn_ws, s));
            DwBwdFused f;
            f.g = b.g; f.z = b.z; f.scale = bl.scale; f.shift = bl.shift; f.act = d.act; f.coef = p->coef[0];
            f.zin = p->buf[d.input].z; f.in_scale = pbl.scale; f.in_shift = pbl.shift; f.in_act = pd.act;
            f.w = net->params + d.w_off; f.gout = p->buf[d.input].g; f.dw = net->grads + d.w_off;
            f.dw_partial = p->red_ws; f.dw_partial_floats = p->red_ws_floats; f.bn_partial = p->bn_ws;
            const double ib = 2.0 * p->N * d.in_h * d.in_w * d.cin;
            if (side) {
                // the per-tile filter-gradient partials are reduced on the side stream (nothing in the chain reads dW); the
                // NEXT fused depthwise backward reuses red_ws, so it waits for this reduction first
                if (dw_reduce_pending) { AMS_CUDA_CHECK(cudaStreamWaitEvent(s, net->ev_dwred, 0)); dw_reduce_pending = false; }
                if (dw_conv_bwd_fused(f, g, &p->pending_bn_rows, s, true)) return -1;
                AMS_CUDA_CHECK(cudaEventRecord(net->ev_fork, s));
                AMS_CUDA_CHECK(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
                forked = true;
                if (dw_conv_bwd_reduce(f, g, net->side_stream)) return -1;
                AMS_CUDA_CHECK(cudaEventRecord(net->ev_dwred, net->side_stream));
                dw_reduce_pending = true;
            } else {
                PROF("dw_bwd_fused", 2.0 * tb + 2.0 * ib, dw_conv_bwd_fused(f, g, &p->pending_bn_rows, s));
            }
            continue;
        }
        static const bool x_skip_bfin = [] { const char* e = getenv("AMS_X_SKIP_BWD_FINALIZE"); return e && e[0] == '1'; }();
        if (p->pending_bn_rows > 0) {
            // the fused depthwise backward left the masked gradient in b.g and the column sums in bn_ws
            if (!x_skip_bfin)
            PROF("bn_bwd_finalize", 16.0 * p->pending_bn_rows * d.cout,
                 bn_backward_finalize_partials(p->bn_ws, p->pending_bn_rows, bl, p->coef[1], net->grads + d.gamma_off, net->grads + d.beta_off, s));
            PROF("bn_bwd_apply", 3.0 * tb, bn_backward_apply(b.g, b.z, bl, p->coef[1], b.gz, s));
            p->pending_bn_rows = 0;
        } else {
            // reduce pass reads dy,z; apply pass reads dy,z and writes dz
            PROF("bn_bwd", 5.0 * tb, bn_backward(b.g, nullptr, b.z, bl, d.act, b.gz, net->grads + d.gamma_off, net->grads + d.beta_off, p->bn_ws, s));
        }
        if (d.kind == kConv1x1) {
            const bool is_cp = d.name == "concat_projection";
            ImgPoolBwd ib;
            if (is_cp) {
                const LayerDef& ipd = L[nl - 4];
                ib.f = imgpool_desc(net, p, false, false);
                ib.dz_proj = b.gz;
                ib.dbias = p->ip_dbias;
                ib.d_w_proj_top = net->grads + d.w_off;
                ib.d_w_pool = net->grads + ipd.w_off;
                ib.d_gamma = net->grads + ipd.gamma_off; ib.d_beta = net->grads + ipd.beta_off;
                ib.dfeat_rowbias = p->dfeat_rowbias;
            }
            if (side) {
                if (is_cp) {
                    // pooled-branch backward first on the side stream; the chain needs its row bias only at the aspp0 dgrad
                    AMS_CUDA_CHECK(cudaEventRecord(net->ev_fork, s));
                    AMS_CUDA_CHECK(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
                    forked = true;
                    if (imgpool_backward(ib, net->side_stream)) return -1;
                    AMS_CUDA_CHECK(cudaEventRecord(net->ev_pool, net->side_stream));
                    pool_bwd_pending = true;
                }
                if (wgrad_on_side(p->wgrad[i])) return -1;
            } else {
                PROF("gemm_wgrad", 2.0 * M * d.k_rows + tb, wgrad_launch(p->wgrad[i], s));
                if (is_cp) PROF("imgpool_bwd", tb, imgpool_backward(ib, s));
            }
            if (pool_bwd_pending && d.name == "aspp0") {
                AMS_CUDA_CHECK(cudaStreamWaitEvent(s, net->ev_pool, 0));
                pool_bwd_pending = false;
            }
            PROF("gemm_dgrad", tb + 2.0 * M * d.k_rows + 2.0 * d.k_rows * d.cout, gemm_launch(p->dgrad[i], s));
        } else if (d.kind == kDepthwise) {
            Conv2dGeom g{p->N, d.in_h, d.in_w, d.cin, d.out_h, d.out_w, d.stride, d.dil, d.pad_top, d.pad_left};
            const double ib = 2.0 * p->N * d.in_h * d.in_w * d.cin;
            if (dw_reduce_pending) { AMS_CUDA_CHECK(cudaStreamWaitEvent(s, net->ev_dwred, 0)); dw_reduce_pending = false; }   // red_ws reuse
            PROF("dw_bwd_filter", ib + tb, dw_conv_bwd_filter(p->buf[d.input].y, b.gz, g, net->grads + d.w_off, p->red_ws, p->red_ws_floats, s));
            PROF("dw_bwd_data", ib + tb, dw_conv_bwd_data(b.gz, net->params + d.w_off, g, p->buf[d.input].g, s));
        } else if (d.kind == kStem) {
            if (dw_reduce_pending) { AMS_CUDA_CHECK(cudaStreamWaitEvent(s, net->ev_dwred, 0)); dw_reduce_pending = false; }   // red_ws reuse
            PROF("stem_bwd_filter", tb + static_cast<double>(p->N) * c.height * c.width * 3,
                 stem_conv_bwd_filter(p->in_frames, p->in_dtype == AMS_FRAMES_U8, p->N, c.height, c.width, d.in_h, d.in_w,
                                      d.out_h, d.out_w, d.pad_top, d.pad_left, 127.5f, 0.007843137718737125f, 1.0f, b.gz,
                                      net->grads + d.w_off, p->red_ws, p->red_ws_floats, s));
        }
        if (i == net->bucket_layer && net->ev_bucket) {
            // every gradient of the late bucket has been enqueued: main-stream parts up to here, filter gradients on the side
            // stream.  One event after both; external-record flavour when this runs under stream capture.
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            AMS_CUDA_CHECK(cudaStreamIsCapturing(s, &cap));
            const unsigned flags = cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
            cudaStream_t rec = s;
            if (side) {
                AMS_CUDA_CHECK(cudaEventRecord(net->ev_bucket_main, s));
                AMS_CUDA_CHECK(cudaStreamWaitEvent(net->side_stream, net->ev_bucket_main, 0));
                forked = true;
                rec = net->side_stream;
            }
            AMS_CUDA_CHECK(cudaEventRecordWithFlags(net->ev_bucket, rec, flags));
        }
    }
    if (forked) {
        AMS_CUDA_CHECK(cudaEventRecord(net->ev_join, net->side_stream));
        AMS_CUDA_CHECK(cudaStreamWaitEvent(s, net->ev_join, 0));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- graph replay
int net_train_fwd_bwd(Net* net, Plan* p, bool normalize) {
    static const bool no_graph = [] { const char* e = getenv("AMS_NO_GRAPH"); return e && e[0] == '1'; }();
    const int k = normalize ? 1 : 0;
    auto eager = [&]() -> int {
        // data parallel with global-batch BatchNorm: bump the exchange epoch, then every BN finalize of this step
        // (forward and backward) sums its statistics over the ranks through NVLink peer memory
        net->sync_active = net->syncbn_enabled;
        int rc = 0;
        if (net->sync_active) rc = syncbn_begin_step(net->syncbn_dev, net->stream);
        if (!rc) rc = net_forward(net, p, AMS_BN_BATCH, true);
        if (!rc) rc = net_backward(net, p, normalize);
        net->sync_active = false;
        return rc;
    };
    if (no_graph || net->prof.enabled || p->train_runs < 0) return eager();
    if (p->train_graph[k] && p->train_graph_dtype[k] == p->in_dtype) {
        AMS_CUDA_CHECK(cudaGraphLaunch(p->train_graph[k], net->stream));
        count_launches(p->train_graph_kernels[k]);
        net->weights_dirty = false;              // the captured sequence starts with the bf16 weight refresh
        p->last_was_train = true;
        return 0;
    }
    if (p->train_runs == 0) { p->train_runs = 1; return eager(); }      // first step on this plan: sets kernel attributes
    if (p->train_graph[k]) { cudaGraphExecDestroy(p->train_graph[k]); p->train_graph[k] = nullptr; }
    net->weights_dirty = true;
    cudaGraph_t graph = nullptr;
    const long long launches0 = launches_so_far();
    AMS_CUDA_CHECK(cudaStreamBeginCapture(net->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = eager();
    const cudaError_t ec = cudaStreamEndCapture(net->stream, &graph);
    if (rc || ec != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        p->train_runs = -1;                      // this plan stays eager
        net->weights_dirty = true;
        return eager();
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !exec) { cudaGetLastError(); p->train_runs = -1; net->weights_dirty = true; return eager(); }
    p->train_graph[k] = exec;
    p->train_graph_dtype[k] = p->in_dtype;
    p->train_graph_kernels[k] = launches_so_far() - launches0;       // counted while capturing: replays add the same number
    AMS_CUDA_CHECK(cudaGraphLaunch(exec, net->stream));
    net->weights_dirty = false;
    p->last_was_train = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------- queue
int net_dequeue(Net* net, Plan** plan_out, bool need_backward) {
    int slot = -1;
    {
        // TF's QueueDequeueV2 blocks until a batch is there (the reference's feeder thread runs concurrently);
        // give up after 60 s so that a missing ams_enqueue is an error, not a hang
        std::unique_lock<std::mutex> lk(net->qmu);
        const bool got = net->qcv.wait_for(lk, std::chrono::seconds(60), [&] { return !net->filled.empty(); });
        AMS_REQUIRE(got, "input queue stayed empty for 60 s: call ams_enqueue first");
        slot = net->filled.front();
        net->filled.pop_front();
    }
    QueueSlot& q = net->slots[slot];
    Plan* p = net_get_plan(net, q.n, need_backward);
    if (!p) return -1;
    const size_t px = static_cast<size_t>(q.n) * net->cfg.height * net->cfg.width;
    const size_t fbytes = px * 3 * (q.dtype == AMS_FRAMES_U8 ? 1 : 4);
    AMS_CUDA_CHECK(cudaMemcpyAsync(p->in_frames, q.frames, fbytes, cudaMemcpyDeviceToDevice, net->stream));
    if (q.has_labels) AMS_CUDA_CHECK(cudaMemcpyAsync(p->in_labels, q.labels, px, cudaMemcpyDeviceToDevice, net->stream));
    else AMS_CUDA_CHECK(cudaMemsetAsync(p->in_labels, 255, px, net->stream));
    p->in_dtype = q.dtype;
    p->in_has_labels = q.has_labels;
    AMS_CUDA_CHECK(cudaEventRecord(q.consumed, net->stream));
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        q.consumed_pending = true;
        net->free_slots.push_back(slot);
    }
    net->qcv.notify_all();
    net->last_n = q.n;
    *plan_out = p;
    return 0;
}

}  // namespace ams
