// extern "C" surface of libams_b200 (see include/ams_b200.h for the reference call site each entry replaces).
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>

#include "net.cuh"
#include "fused_block.cuh"

namespace ams {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() { static const bool on = [] { const char* e = getenv("AMS_NO_PDL"); return !(e && e[0] == '1'); }(); return on; }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace ams

using namespace ams;

namespace {

int fill_head_geom(const ams_config& c, int h, int w, HeadGeom* g) {
    AMS_REQUIRE(c.class_count > 0 && c.class_count <= 21, "class_count must be in [1,21]");
    g->N = 0; g->h = h; g->w = w; g->ldl = 32; g->H = c.height; g->W = c.width; g->class_count = c.class_count;
    g->normalize = 1;
    for (int i = 0; i < kMaxClasses; ++i) g->cls_idx[i] = 0;
    for (int i = 0; i < 256; ++i) g->label_lut[i] = -1;
    const int depth = c.label_depth > 0 ? c.label_depth : 19;
    for (int j = 0; j < c.class_count; ++j) {
        const int ch = c.class_indices[j];
        AMS_REQUIRE(ch >= 0 && ch < c.num_classes, "class index outside the logits");
        AMS_REQUIRE(j == 0 || ch > c.class_indices[j - 1], "class_indices must be ascending");
        g->cls_idx[j] = ch;
        if (ch < depth) g->label_lut[ch] = j;      // one_hot(labels, depth) -> gather(class_indices)
    }
    return 0;
}

int find_var(Net* net, const char* name) {
    auto it = net->var_index.find(name);
    return it == net->var_index.end() ? -1 : it->second;
}

int ensure_slot(QueueSlot& q, size_t fbytes, size_t lbytes) {
    if (q.frames_cap < fbytes) {
        if (q.frames) cudaFree(q.frames);
        AMS_CUDA_CHECK(cudaMalloc(&q.frames, fbytes));
        q.frames_cap = fbytes;
    }
    if (q.labels_cap < lbytes) {
        if (q.labels) cudaFree(q.labels);
        AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&q.labels), lbytes));
        q.labels_cap = lbytes;
    }
    return 0;
}

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

extern "C" {

const char* ams_last_error(void) { return g_last_error.c_str(); }
int ams_abi_version(void) { return 1; }
long long ams_launch_count(void) { return g_launches.load(); }

ams_net* ams_create(const ams_config* cfg) {
    if (!cfg) { set_last_error("null config"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_last_error("libams_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
        return nullptr;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_last_error("bad device ordinal"); return nullptr; }
    if (cfg->height <= 0 || cfg->width <= 0) { set_last_error("bad frame size"); return nullptr; }
    Net* net = new Net();
    net->cfg = *cfg;
    auto fail = [&](const char* what) -> ams_net* {
        if (g_last_error.empty()) set_last_error(what);
        delete net;
        return nullptr;
    };
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
    if (prop.major != 10) {
        set_last_error(std::string("libams_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
        delete net;
        return nullptr;
    }
    net->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&net->own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&net->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&net->side_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&net->ev_pool, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&net->ev_dwred, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&net->ev_bucket, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&net->ev_bucket_main, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaStreamCreateWithFlags(&net->split_ss.main, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&net->split_ss.side, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    for (cudaEvent_t* e : {&net->split_ss.fork, &net->split_ss.join, &net->ev_split_fork, &net->ev_split_join})
        if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    { const char* e = getenv("AMS_NO_INFER_SPLIT"); if (e && e[0] == '1') net->infer_split = false; }
    net->stream = net->own_stream;
    if (net_build_topology(net)) return fail("topology");
    const LayerDef& lg = net->layers.back();
    if (fill_head_geom(net->cfg, lg.out_h, lg.out_w, &net->head)) return fail("head");
    const size_t nt = static_cast<size_t>(net->n_train);
    bool ok = true;
    auto A = [&](auto** p, size_t count) {
        void* q = nullptr;
        if (ok && cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(**p)) != cudaSuccess) ok = false;
        *p = static_cast<std::remove_reference_t<decltype(*p)>>(q);
        if (ok) cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(**p));
    };
    A(&net->params, nt); A(&net->grads, nt); A(&net->adam_m, nt); A(&net->adam_v, nt); A(&net->before, nt);
    A(&net->delta_scratch, nt); A(&net->mask, nt);
    A(&net->moving, static_cast<size_t>(net->n_moving));
    A(&net->bnpool, static_cast<size_t>(net->n_bnpool));
    A(&net->wpool, static_cast<size_t>(net->n_bf16));
    A(&net->select_sc, 1); A(&net->head_st, 1);
    if (!ok) return fail("device allocation failed");
    cudaMemset(net->mask, 1, nt);
    // weight-cast table and per-variable bit segments
    std::vector<WeightCast> table;
    for (const LayerDef& d : net->layers) {
        if (d.kind != kConv1x1 && d.kind != kLogits) continue;
        WeightCast t;
        t.w = net->params + d.w_off; t.w_fwd = reinterpret_cast<act_t*>(net->wpool + d.wfwd_off); t.w_bwd = reinterpret_cast<bf16*>(net->wpool + d.wbwd_off);
        t.w_lo = d.wlo_off >= 0 ? reinterpret_cast<act_t*>(net->wpool + d.wlo_off) : nullptr;
        t.Cin = d.cin; t.Cout = d.cout; t.ld_fwd = d.ld_fwd; t.ld_bwd = d.ld_bwd; t.row0 = d.k_rows0; t.rows = d.k_rows;
        table.push_back(t);
        net->cast_max = std::max(net->cast_max, d.k_rows * d.cout);
    }
    net->cast_layers = static_cast<int>(table.size());
    A(&net->cast_table, table.size());
    std::vector<VarSeg> segs;
    long long boff = 0;
    for (int vi : net->trainable_order) {
        const VarInfo& v = net->vars[vi];
        segs.push_back(VarSeg{v.offset, v.count, boff});
        boff += (v.count + 7) / 8;
    }
    net->mask_bytes = boff;
    A(&net->segs_dev, segs.size());
    A(&net->pack_bits, static_cast<size_t>(boff));
    A(&net->pack_vals, nt);
    A(&net->pack_counts, static_cast<size_t>(pack_delta_blocks(net->n_train)));
    A(&net->pack_kept, 1);
    if (!ok) return fail("device allocation failed");
    cudaMemcpy(net->cast_table, table.data(), table.size() * sizeof(WeightCast), cudaMemcpyHostToDevice);
    cudaMemcpy(net->segs_dev, segs.data(), segs.size() * sizeof(VarSeg), cudaMemcpyHostToDevice);
    // block-fused frozen inference (fused_block.cu): padded per-channel parameter vectors of every eligible block
    net->fused_params.assign(net->layers.size(), nullptr);
    {
        // ON by default (env AMS_BLOCK_FUSION=0: one kernel per layer): same fp16 rounding points as the per-layer schedule
        const char* nf = getenv("AMS_BLOCK_FUSION");
        net->block_fusion = !(nf && nf[0] == '0');
        for (size_t i = 0; i + 2 < net->layers.size(); ++i) {
            const LayerDef& e = net->layers[i]; const LayerDef& dw = net->layers[i + 1]; const LayerDef& pr = net->layers[i + 2];
            if (e.kind != kConv1x1 || dw.kind != kDepthwise || pr.kind != kConv1x1 || dw.input != static_cast<int>(i) ||
                pr.input != static_cast<int>(i + 1)) continue;
            FusedBlockDesc f;
            f.N = 1; f.H = e.out_h; f.W = e.out_w; f.Cin = e.cin; f.Cexp = e.cout; f.Cout = pr.cout; f.stride = dw.stride; f.dil = dw.dil;
            if (!fused_block_supported(f)) continue;
            if (cudaMalloc(reinterpret_cast<void**>(&net->fused_params[i]), fused_block_param_floats(f) * sizeof(float)) != cudaSuccess)
                return fail("device allocation failed");
        }
    }
    const int cap = cfg->queue_capacity > 0 ? cfg->queue_capacity : 4;
    net->slots.resize(cap);
    for (int i = 0; i < cap; ++i) {
        if (cudaEventCreateWithFlags(&net->slots[i].consumed, cudaEventDisableTiming) != cudaSuccess) return fail("event");
        net->free_slots.push_back(i);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("device sync failed");
    return reinterpret_cast<ams_net*>(net);
}

void ams_destroy(ams_net* h) {
    if (!h) return;
    Net* net = reinterpret_cast<Net*>(h);
    cudaSetDevice(net->cfg.device);
    cudaDeviceSynchronize();
    for (auto& kv : net->plans) {
        for (cudaGraphExec_t g : kv.second->train_graph) if (g) cudaGraphExecDestroy(g);
        for (cudaGraphExec_t g : kv.second->infer_graph) if (g) cudaGraphExecDestroy(g);
        for (auto& hv : kv.second->half) if (hv) for (void* p : hv->allocations) cudaFree(p);
        for (void* p : kv.second->allocations) cudaFree(p);
    }
    for (auto& q : net->slots) { if (q.frames) cudaFree(q.frames); if (q.labels) cudaFree(q.labels); if (q.consumed) cudaEventDestroy(q.consumed); }
    if (net->raw_frames) cudaFree(net->raw_frames);
    if (net->raw_labels) cudaFree(net->raw_labels);
    for (void* m : net->syncbn_mapped) if (m) cudaIpcCloseMemHandle(m);
    if (net->syncbn_recv) cudaFree(net->syncbn_recv);
    if (net->syncbn_dev) cudaFree(net->syncbn_dev);
    void* ptrs[] = {net->params, net->grads, net->adam_m, net->adam_v, net->before, net->delta_scratch, net->mask, net->moving,
                    net->bnpool, net->wpool, net->select_sc, net->head_st, net->cast_table, net->segs_dev, net->pack_bits,
                    net->pack_vals, net->pack_counts, net->pack_kept};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (float* p : net->fused_params) if (p) cudaFree(p);
    if (net->own_stream) cudaStreamDestroy(net->own_stream);
    if (net->copy_stream) cudaStreamDestroy(net->copy_stream);
    if (net->side_stream) cudaStreamDestroy(net->side_stream);
    if (net->ev_fork) cudaEventDestroy(net->ev_fork);
    if (net->ev_join) cudaEventDestroy(net->ev_join);
    if (net->ev_pool) cudaEventDestroy(net->ev_pool);
    if (net->ev_dwred) cudaEventDestroy(net->ev_dwred);
    if (net->ev_bucket) cudaEventDestroy(net->ev_bucket);
    if (net->ev_bucket_main) cudaEventDestroy(net->ev_bucket_main);
    if (net->split_ss.main) cudaStreamDestroy(net->split_ss.main);
    if (net->split_ss.side) cudaStreamDestroy(net->split_ss.side);
    for (cudaEvent_t e : {net->split_ss.fork, net->split_ss.join, net->ev_split_fork, net->ev_split_join}) if (e) cudaEventDestroy(e);
    delete net;
}

#define NET(h) Net* net = reinterpret_cast<Net*>(h); if (!net) { set_last_error("null handle"); return -1; } cudaSetDevice(net->cfg.device)

int ams_set_stream(ams_net* h, void* stream) {
    NET(h);
    net->stream = stream ? as_stream(stream) : net->own_stream;
    return 0;
}
int ams_synchronize(ams_net* h) {
    NET(h);
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}

int ams_num_tensors(const ams_net* h) { return h ? static_cast<int>(reinterpret_cast<const Net*>(h)->vars.size()) : -1; }

int ams_tensor_info(const ams_net* h, int index, char* name, int cap, int shape4[4], int* ndim, int* trainable, long long* off) {
    const Net* net = reinterpret_cast<const Net*>(h);
    if (!net || index < 0 || index >= static_cast<int>(net->vars.size())) { set_last_error("bad tensor index"); return -1; }
    const VarInfo& v = net->vars[index];
    if (name && cap > 0) { std::strncpy(name, v.name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = v.shape[i];
    if (ndim) *ndim = v.ndim;
    if (trainable) *trainable = v.trainable ? 1 : 0;
    if (off) *off = v.offset;
    return 0;
}

static int resolve(Net* net, const char* name, float** dev, long long* count, bool* is_param) {
    std::string s(name ? name : "");
    *is_param = false;
    if (s == "beta1_power:0" || s == "beta2_power:0") { *dev = nullptr; *count = 1; return 0; }
    float* base_t = net->params; float* base_m = net->moving;
    std::string var = s;
    auto ends_with = [&](const std::string& suf) { return s.size() > suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; };
    bool slot = false;
    if (ends_with("/Adam_1:0")) { var = s.substr(0, s.size() - 9) + ":0"; base_t = net->adam_v; slot = true; }
    else if (ends_with("/Adam:0")) { var = s.substr(0, s.size() - 7) + ":0"; base_t = net->adam_m; slot = true; }
    const int vi = find_var(net, var.c_str());
    AMS_REQUIRE(vi >= 0, std::string("KeyError: no variable named '") + s + "'");
    const VarInfo& v = net->vars[vi];
    AMS_REQUIRE(!slot || v.trainable, "optimizer slots exist for trainable variables only");
    *dev = (v.trainable ? base_t : base_m) + v.offset;
    *count = v.count;
    *is_param = !slot;
    return 0;
}

int ams_set_tensor(ams_net* h, const char* name, const float* host, long long count) {
    NET(h);
    std::string s(name ? name : "");
    if (s == "beta1_power:0") { net->beta1_power = host[0]; return 0; }
    if (s == "beta2_power:0") { net->beta2_power = host[0]; return 0; }
    float* dev; long long n; bool is_param;
    if (resolve(net, name, &dev, &n, &is_param)) return -1;
    AMS_REQUIRE(n == count, "element count does not match the variable's shape");
    AMS_CUDA_CHECK(cudaMemcpyAsync(dev, host, n * sizeof(float), cudaMemcpyHostToDevice, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    if (is_param) { net->weights_dirty = true; net->fold_dirty = true; }
    return 0;
}

int ams_get_tensor(ams_net* h, const char* name, float* host, long long count) {
    NET(h);
    std::string s(name ? name : "");
    if (s == "beta1_power:0") { host[0] = net->beta1_power; return 0; }
    if (s == "beta2_power:0") { host[0] = net->beta2_power; return 0; }
    float* dev; long long n; bool is_param;
    if (resolve(net, name, &dev, &n, &is_param)) return -1;
    AMS_REQUIRE(n == count, "element count does not match the variable's shape");
    AMS_CUDA_CHECK(cudaMemcpyAsync(host, dev, n * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}


// =============================================================================================== frozen hand-off
// Container of the client model (the reference writes a TF GraphDef with constants, SemanticNetwork.py:706-714 ->
// utils/graph_utils.py:79-126 trim_graph_frozen(kill_norms=True); this library has no GraphDef, so the file holds what
// that graph holds: every variable of the student with inference-mode BatchNorm semantics, i.e. gamma / beta and the
// MOVING mean / variance, in tf.global_variables() order):
//   "AMSFRZ01" | i32 num_classes | i32 graph_variant | i32 n_tensors |
//   per tensor: u16 name_len | name | i32 ndim | i32 shape[4] | i64 count | count x f32
static const char kFrozenMagic[8] = {'A', 'M', 'S', 'F', 'R', 'Z', '0', '1'};

int ams_export_frozen(ams_net* h, const char* path) {
    NET(h);
    AMS_REQUIRE(path && path[0], "empty path");
    std::vector<float> tr(static_cast<size_t>(net->n_train)), mv(static_cast<size_t>(net->n_moving));
    AMS_CUDA_CHECK(cudaMemcpyAsync(tr.data(), net->params, tr.size() * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaMemcpyAsync(mv.data(), net->moving, mv.size() * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    FILE* f = std::fopen(path, "wb");
    AMS_REQUIRE(f != nullptr, std::string("cannot open ") + path + " for writing");
    bool ok = std::fwrite(kFrozenMagic, 1, 8, f) == 8;
    const int32_t hdr[3] = {net->cfg.num_classes, net->cfg.graph_variant, static_cast<int32_t>(net->vars.size())};
    ok = ok && std::fwrite(hdr, sizeof(int32_t), 3, f) == 3;
    for (const VarInfo& v : net->vars) {
        const uint16_t nl = static_cast<uint16_t>(v.name.size());
        const int32_t nd = v.ndim;
        const int32_t shp[4] = {v.shape[0], v.shape[1], v.shape[2], v.shape[3]};
        const int64_t cnt = v.count;
        const float* src = (v.trainable ? tr.data() : mv.data()) + v.offset;
        ok = ok && std::fwrite(&nl, 2, 1, f) == 1 && std::fwrite(v.name.data(), 1, nl, f) == nl && std::fwrite(&nd, 4, 1, f) == 1 &&
             std::fwrite(shp, 4, 4, f) == 4 && std::fwrite(&cnt, 8, 1, f) == 1 &&
             std::fwrite(src, sizeof(float), static_cast<size_t>(cnt), f) == static_cast<size_t>(cnt);
    }
    ok = (std::fclose(f) == 0) && ok;
    AMS_REQUIRE(ok, std::string("short write to ") + path);
    return 0;
}

ams_net* ams_create_frozen(const char* path, const ams_config* cfg) {
    if (!path || !cfg) { set_last_error("null path / config"); return nullptr; }
    FILE* f = std::fopen(path, "rb");
    if (!f) { set_last_error(std::string("cannot open frozen model ") + path); return nullptr; }
    char magic[8];
    int32_t hdr[3] = {0, 0, 0};
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, kFrozenMagic, 8) != 0 || std::fread(hdr, 4, 3, f) != 3) {
        std::fclose(f);
        set_last_error(std::string(path) + " is not an ams_b200 frozen model (a TF GraphDef .pb cannot be loaded)");
        return nullptr;
    }
    ams_config c = *cfg;
    if ((c.num_classes != 0 && c.num_classes != hdr[0]) || (c.num_classes != 0 && c.graph_variant != hdr[1])) {
        std::fclose(f);
        set_last_error("frozen model was exported from a different graph (num_classes / graph_variant mismatch)");
        return nullptr;
    }
    c.num_classes = hdr[0]; c.graph_variant = hdr[1];
    ams_net* h = ams_create(&c);
    if (!h) { std::fclose(f); return nullptr; }
    Net* net = reinterpret_cast<Net*>(h);
    auto fail = [&](const std::string& msg) -> ams_net* { std::fclose(f); ams_destroy(h); set_last_error(msg); return nullptr; };
    if (hdr[2] != static_cast<int32_t>(net->vars.size())) return fail("frozen model: wrong number of variables");
    std::vector<float> buf;
    std::string name;
    for (int i = 0; i < hdr[2]; ++i) {
        uint16_t nl = 0; int32_t nd = 0, shp[4]; int64_t cnt = 0;
        if (std::fread(&nl, 2, 1, f) != 1) return fail("frozen model: truncated file");
        name.resize(nl);
        if (std::fread(&name[0], 1, nl, f) != nl || std::fread(&nd, 4, 1, f) != 1 || std::fread(shp, 4, 4, f) != 4 ||
            std::fread(&cnt, 8, 1, f) != 1) return fail("frozen model: truncated file");
        const int vi = find_var(net, name.c_str());
        if (vi < 0) return fail("frozen model: KeyError: no variable named '" + name + "'");
        const VarInfo& v = net->vars[vi];
        if (cnt != v.count || nd != v.ndim) return fail("frozen model: shape mismatch for " + name);
        buf.resize(static_cast<size_t>(cnt));
        if (std::fread(buf.data(), sizeof(float), buf.size(), f) != buf.size()) return fail("frozen model: truncated file");
        if (ams_set_tensor(h, name.c_str(), buf.data(), cnt)) { std::fclose(f); ams_destroy(h); return nullptr; }
    }
    std::fclose(f);
    net->frozen = true;
    return h;
}
int ams_is_frozen(const ams_net* h) { return (h && reinterpret_cast<const Net*>(h)->frozen) ? 1 : 0; }

long long ams_trainable_count(const ams_net* h) { return h ? reinterpret_cast<const Net*>(h)->n_train : -1; }
int ams_get_trainable(ams_net* h, float* host) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemcpyAsync(host, net->params, net->n_train * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}
int ams_set_trainable(ams_net* h, const float* host) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemcpyAsync(net->params, host, net->n_train * sizeof(float), cudaMemcpyHostToDevice, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->weights_dirty = true; net->fold_dirty = true;
    return 0;
}
int ams_reset_optimizer(ams_net* h) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemsetAsync(net->adam_m, 0, net->n_train * sizeof(float), net->stream));
    AMS_CUDA_CHECK(cudaMemsetAsync(net->adam_v, 0, net->n_train * sizeof(float), net->stream));
    net->beta1_power = 0.9f; net->beta2_power = 0.999f;
    return 0;
}

int ams_enqueue(ams_net* h, const void* frames, int dtype, const uint8_t* labels, int n) {
    NET(h);
    AMS_REQUIRE(frames && n > 0, "frames must be non-empty");
    AMS_REQUIRE(dtype == AMS_FRAMES_U8 || dtype == AMS_FRAMES_F32, "frames_dtype must be u8 or f32");
    int slot = -1;
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        // FIFOQueue semantics: block while the queue is full -- but a full queue nobody drains is an error, not a hang
        const bool got = net->qcv.wait_for(lk, std::chrono::seconds(60), [&] { return !net->free_slots.empty(); });
        AMS_REQUIRE(got, "input queue stayed full for 60 s: nothing is consuming it (queue_capacity too small?)");
        slot = net->free_slots.front();
        net->free_slots.pop_front();
    }
    QueueSlot& q = net->slots[slot];
    if (q.consumed_pending) { AMS_CUDA_CHECK(cudaEventSynchronize(q.consumed)); q.consumed_pending = false; }
    const size_t px = static_cast<size_t>(n) * net->cfg.height * net->cfg.width;
    const size_t fbytes = px * 3 * (dtype == AMS_FRAMES_U8 ? 1 : 4);
    if (ensure_slot(q, fbytes, px)) return -1;
    AMS_CUDA_CHECK(cudaMemcpyAsync(q.frames, frames, fbytes, cudaMemcpyHostToDevice, net->copy_stream));
    if (labels) AMS_CUDA_CHECK(cudaMemcpyAsync(q.labels, labels, px, cudaMemcpyHostToDevice, net->copy_stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->copy_stream));
    q.n = n; q.dtype = dtype; q.has_labels = labels != nullptr;
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        net->filled.push_back(slot);
    }
    net->qcv.notify_all();
    return 0;
}
int ams_enqueue_raw(ams_net* h, const uint8_t* frames, int src_h, int src_w, int bgr_to_rgb, const uint8_t* labels, int lab_h,
                    int lab_w, int n) {
    NET(h);
    AMS_REQUIRE(frames && n > 0 && src_h > 0 && src_w > 0, "frames must be non-empty");
    AMS_REQUIRE(!labels || (lab_h > 0 && lab_w > 0), "label size missing");
    int slot = -1;
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        const bool got = net->qcv.wait_for(lk, std::chrono::seconds(60), [&] { return !net->free_slots.empty(); });
        AMS_REQUIRE(got, "input queue stayed full for 60 s: nothing is consuming it (queue_capacity too small?)");
        slot = net->free_slots.front();
        net->free_slots.pop_front();
    }
    QueueSlot& q = net->slots[slot];
    if (q.consumed_pending) { AMS_CUDA_CHECK(cudaEventSynchronize(q.consumed)); q.consumed_pending = false; }
    const size_t px = static_cast<size_t>(n) * net->cfg.height * net->cfg.width;
    if (ensure_slot(q, px * 3, px)) return -1;
    const size_t fraw = static_cast<size_t>(n) * src_h * src_w * 3;
    if (fraw > net->raw_frames_cap) {
        if (net->raw_frames) cudaFree(net->raw_frames);
        AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&net->raw_frames), fraw));
        net->raw_frames_cap = fraw;
    }
    cudaStream_t cs = net->copy_stream;
    AMS_CUDA_CHECK(cudaMemcpyAsync(net->raw_frames, frames, fraw, cudaMemcpyHostToDevice, cs));
    // cv2.resize(frame, (W, H)) + cv2.cvtColor(BGR2RGB)  -- run.py:415-416
    if (resize_u8(net->raw_frames, n, src_h, src_w, 3, static_cast<uint8_t*>(q.frames), net->cfg.height, net->cfg.width, 0,
                  bgr_to_rgb ? 1 : 0, cs)) return -1;
    if (labels) {
        const size_t lraw = static_cast<size_t>(n) * lab_h * lab_w;
        if (lraw > net->raw_labels_cap) {
            if (net->raw_labels) cudaFree(net->raw_labels);
            AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&net->raw_labels), lraw));
            net->raw_labels_cap = lraw;
        }
        AMS_CUDA_CHECK(cudaMemcpyAsync(net->raw_labels, labels, lraw, cudaMemcpyHostToDevice, cs));
        // cv2.resize(label, (W, H), interpolation=cv2.INTER_NEAREST)  -- run.py:183, :421
        if (resize_u8(net->raw_labels, n, lab_h, lab_w, 1, q.labels, net->cfg.height, net->cfg.width, 1, 0, cs)) return -1;
    }
    AMS_CUDA_CHECK(cudaStreamSynchronize(cs));
    q.n = n; q.dtype = AMS_FRAMES_U8; q.has_labels = labels != nullptr;
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        net->filled.push_back(slot);
    }
    net->qcv.notify_all();
    return 0;
}
int ams_set_block_fusion(ams_net* h, int on) {
    NET(h);
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->block_fusion = on != 0;
    for (auto& kv : net->plans)                      // captured inference graphs contain the old schedule
        for (int k = 0; k < 2; ++k)
            if (kv.second->infer_graph[k]) { cudaGraphExecDestroy(kv.second->infer_graph[k]); kv.second->infer_graph[k] = nullptr; }
    return 0;
}
int ams_set_infer_split(ams_net* h, int on) {
    NET(h);
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->infer_split = on != 0;
    for (auto& kv : net->plans)
        for (int k = 0; k < 2; ++k)
            if (kv.second->infer_graph[k]) { cudaGraphExecDestroy(kv.second->infer_graph[k]); kv.second->infer_graph[k] = nullptr; }
    return 0;
}
int ams_queue_size(ams_net* h) {
    NET(h);
    std::unique_lock<std::mutex> lk(net->qmu);
    return static_cast<int>(net->filled.size());
}
int ams_queue_clear(ams_net* h) {
    NET(h);
    int dropped = 0;
    {
        std::unique_lock<std::mutex> lk(net->qmu);
        while (!net->filled.empty()) {
            net->free_slots.push_back(net->filled.front());       // fully staged (ams_enqueue synchronises its copy stream): reusable at once
            net->filled.pop_front();
            ++dropped;
        }
    }
    net->qcv.notify_all();
    return dropped;
}

static int infer_common(Net* net, int bn_mode, bool metric, int32_t* out_labels, int64_t* out_cm, float* out_loss) {
    Plan* p = nullptr;
    if (net_dequeue(net, &p, false)) return -1;
    HeadStats hs;
    // Frozen inference of an even batch >= 4 runs as two half batches on two streams (one graph): the layers at 1/16
    // resolution launch fewer CTAs than the GPU has SMs and every layer ends in a tail wave, so two independent chains
    // fill each other's gaps.  Frames are independent in moving-statistics mode and the metric accumulators are integer
    // atomics, so the result is bit-identical to the unsplit run.
    const bool split = net->infer_split && bn_mode == AMS_BN_MOVING && !net->prof.enabled && p->N >= 4 && (p->N & 1) == 0;
    Plan *ha = nullptr, *hb = nullptr;
    if (split) {
        if (net_half_plans(net, p, &ha, &hb)) return -1;
        const size_t el = p->in_dtype == AMS_FRAMES_U8 ? 1 : 4;
        const size_t half_bytes = static_cast<size_t>(ha->N) * net->cfg.height * net->cfg.width * 3 * el;
        ha->in_frames = p->in_frames; hb->in_frames = static_cast<char*>(p->in_frames) + half_bytes;
        ha->in_dtype = hb->in_dtype = p->in_dtype;
    }
    auto head = [&](Plan* q, cudaStream_t s) -> int {
        HeadGeom g = net->head; g.N = q->N;
        const double px_ = static_cast<double>(q->N) * net->cfg.height * net->cfg.width;
        net->prof.begin(s, "head_infer", px_ * (metric ? 5.0 : 4.0) + 4.0 * q->N * g.h * g.w * 32);
        const int rc = head_infer(q->logits, g, metric ? q->in_labels : nullptr, q->pred, net->head_st, s);
        net->prof.end(s);
        return rc ? -1 : 0;
    };
    auto body = [&]() -> int {
        if (!split) {
            if (net_forward(net, p, bn_mode, false)) return -1;
            if (metric) { if (head_reset(net->head_st, net->stream)) return -1; }
            return head(p, net->stream);
        }
        if (net_prepare_weights(net, true)) return -1;
        if (metric) { if (head_reset(net->head_st, net->stream)) return -1; }
        AMS_CUDA_CHECK(cudaEventRecord(net->ev_split_fork, net->stream));
        AMS_CUDA_CHECK(cudaStreamWaitEvent(net->split_ss.main, net->ev_split_fork, 0));
        if (net_forward(net, ha, bn_mode, false)) return -1;
        if (head(ha, net->stream)) return -1;
        if (net_forward(net, hb, bn_mode, false, &net->split_ss)) return -1;
        if (head(hb, net->split_ss.main)) return -1;
        AMS_CUDA_CHECK(cudaEventRecord(net->ev_split_join, net->split_ss.main));
        AMS_CUDA_CHECK(cudaStreamWaitEvent(net->stream, net->ev_split_join, 0));
        p->last_was_train = false;
        return 0;
    };
    // frozen client inference: the ~60 launches of forward + head are captured once per plan and replayed
    static const bool no_graph = [] { const char* e = getenv("AMS_NO_GRAPH"); return e && e[0] == '1'; }();
    const int k = metric ? 1 : 0;
    if (bn_mode != AMS_BN_MOVING || no_graph || net->prof.enabled || p->infer_runs < 0) {
        if (body()) return -1;
    } else {
        if (net_prepare_weights(net, true)) return -1;           // conditional work stays outside the graph
        if (p->infer_graph[k] && p->infer_graph_dtype[k] == p->in_dtype) {
            AMS_CUDA_CHECK(cudaGraphLaunch(p->infer_graph[k], net->stream));
            count_launches(p->infer_graph_kernels[k]);
            p->last_was_train = false;
        } else if (p->infer_runs == 0) {
            p->infer_runs = 1;
            if (body()) return -1;
        } else {
            if (p->infer_graph[k]) { cudaGraphExecDestroy(p->infer_graph[k]); p->infer_graph[k] = nullptr; }
            const long long launches0 = launches_so_far();
            cudaGraph_t graph = nullptr;
            AMS_CUDA_CHECK(cudaStreamBeginCapture(net->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = body();
            const cudaError_t ec = cudaStreamEndCapture(net->stream, &graph);
            cudaGraphExec_t exec = nullptr;
            if (rc || ec != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess || !exec) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                p->infer_runs = -1;                               // this plan stays eager
                if (body()) return -1;
            } else {
                cudaGraphDestroy(graph);
                p->infer_graph[k] = exec;
                p->infer_graph_dtype[k] = p->in_dtype;
                p->infer_graph_kernels[k] = launches_so_far() - launches0;
                AMS_CUDA_CHECK(cudaGraphLaunch(exec, net->stream));
            }
        }
    }
    const size_t px = static_cast<size_t>(p->N) * net->cfg.height * net->cfg.width;
    if (out_labels) AMS_CUDA_CHECK(cudaMemcpyAsync(out_labels, p->pred, px * sizeof(int32_t), cudaMemcpyDeviceToHost, net->stream));
    if (metric) AMS_CUDA_CHECK(cudaMemcpyAsync(&hs, net->head_st, sizeof(HeadStats), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    if (metric) {
        const int cc = net->cfg.class_count;
        if (out_cm) for (int i = 0; i < cc; ++i) for (int j = 0; j < cc; ++j) out_cm[i * cc + j] = hs.confmat[i * kMaxClasses + j];
        if (out_loss) *out_loss = hs.n_valid > 0 ? static_cast<float>((static_cast<double>(hs.loss_fixed) / kLossFixedScale) / static_cast<double>(hs.n_valid)) : NAN;
    }
    return 0;
}

int ams_infer(ams_net* h, int bn_mode, int32_t* out_labels) {
    NET(h);
    return infer_common(net, bn_mode, false, out_labels, nullptr, nullptr);
}
int ams_infer_metric(ams_net* h, int bn_mode, int32_t* out_labels, int64_t* out_cm, float* out_loss) {
    NET(h);
    return infer_common(net, bn_mode, true, out_labels, out_cm, out_loss);
}

int ams_confmat_labels(ams_net* h, const uint8_t* before, const uint8_t* after, long long n, int64_t* out_cm) {
    NET(h);
    AMS_REQUIRE(n > 0, "empty label maps");
    uint8_t* dev = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&dev), 2 * n));
    int rc = 0;
    do {
        if (cudaMemcpyAsync(dev, before, n, cudaMemcpyHostToDevice, net->stream) != cudaSuccess ||
            cudaMemcpyAsync(dev + n, after, n, cudaMemcpyHostToDevice, net->stream) != cudaSuccess) { set_last_error("H2D failed"); rc = -1; break; }
        if (head_reset(net->head_st, net->stream) || head_label_confmat(dev, dev + n, n, net->head, net->head_st, net->stream)) { rc = -1; break; }
        HeadStats hs;
        if (cudaMemcpyAsync(&hs, net->head_st, sizeof(HeadStats), cudaMemcpyDeviceToHost, net->stream) != cudaSuccess ||
            cudaStreamSynchronize(net->stream) != cudaSuccess) { set_last_error("D2H failed"); rc = -1; break; }
        const int cc = net->cfg.class_count;
        for (int i = 0; i < cc; ++i) for (int j = 0; j < cc; ++j) out_cm[i * cc + j] = hs.confmat[i * kMaxClasses + j];
    } while (0);
    cudaFree(dev);
    return rc;
}

static int apply_optimizer(Net* net, float lr, int masked, float grad_scale, const double* scale_terms = nullptr) {
    // TF1 Adam: alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t), all in fp32
    const float alpha = lr * std::sqrt(1.0f - net->beta2_power) / (1.0f - net->beta1_power);
    const uint8_t* mask = (masked && !net->mask_all_ones) ? net->mask : nullptr;
    net->prof.begin(net->stream, "adam_masked", (mask ? 29.0 : 28.0) * net->n_train);
    const int rc_adam = adam_masked(net->params, net->grads, grad_scale, scale_terms, net->adam_m, net->adam_v, mask, net->n_train, alpha,
                                    1.0f - 0.9f, 1.0f - 0.999f, 1e-8f, net->stream);
    net->prof.end(net->stream);
    if (rc_adam) return -1;
    net->beta1_power *= 0.9f;
    net->beta2_power *= 0.999f;
    net->weights_dirty = true; net->fold_dirty = true;
    return 0;
}

int ams_train_forward_backward(ams_net* h, long long* out_n_valid, double* out_loss_sum) {
    NET(h);
    AMS_REQUIRE(!net->frozen, "Can't train frozen graph!!! (handle built by ams_create_frozen is inference-only)");
    Plan* p = nullptr;
    if (net_dequeue(net, &p, true)) return -1;
    if (net_train_fwd_bwd(net, p, false)) return -1;
    if (!out_n_valid && !out_loss_sum) return 0;          // asynchronous form: the terms stay on the device (ams_step_terms_device)
    HeadStats hs;
    AMS_CUDA_CHECK(cudaMemcpyAsync(&hs, net->head_st, sizeof(HeadStats), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    if (out_n_valid) *out_n_valid = hs.n_valid;
    if (out_loss_sum) *out_loss_sum = hs.loss_sum;
    return 0;
}
void* ams_gradient_arena(ams_net* h, long long* count) {
    Net* net = reinterpret_cast<Net*>(h);
    if (!net) return nullptr;
    if (count) *count = net->n_train;
    return net->grads;
}
long long ams_gradient_bucket_split(const ams_net* h) { return h ? reinterpret_cast<const Net*>(h)->bucket_split : -1; }
int ams_gradient_bucket_wait(ams_net* h, void* cuda_stream) {
    NET(h);
    AMS_REQUIRE(cuda_stream != nullptr, "a stream is needed");
    AMS_CUDA_CHECK(cudaStreamWaitEvent(as_stream(cuda_stream), net->ev_bucket, 0));
    return 0;
}
int ams_apply_optimizer(ams_net* h, float lr, int masked, float grad_scale) {
    NET(h);
    return apply_optimizer(net, lr, masked, grad_scale);
}

void* ams_step_terms_device(ams_net* h) {
    Net* net = reinterpret_cast<Net*>(h);
    if (!net) return nullptr;
    return reinterpret_cast<char*>(net->head_st) + offsetof(HeadStats, terms);
}
int ams_apply_optimizer_device(ams_net* h, float lr, int masked, float* out_loss_pinned) {
    NET(h);
    const double* terms = reinterpret_cast<const double*>(reinterpret_cast<const char*>(net->head_st) + offsetof(HeadStats, terms));
    if (apply_optimizer(net, lr, masked, 1.0f, terms)) return -1;
    if (out_loss_pinned) {
        if (head_mean_loss_from_terms(net->head_st, net->stream)) return -1;
        AMS_CUDA_CHECK(cudaMemcpyAsync(out_loss_pinned, reinterpret_cast<const char*>(net->head_st) + offsetof(HeadStats, dp_loss),
                                       sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    }
    return 0;
}
int ams_train_step_async(ams_net* h, float lr, int masked, float* out_loss_pinned) {
    NET(h);
    AMS_REQUIRE(!net->frozen, "Can't train frozen graph!!! (handle built by ams_create_frozen is inference-only)");
    Plan* p = nullptr;
    if (net_dequeue(net, &p, true)) return -1;
    if (net_train_fwd_bwd(net, p, true)) return -1;
    if (apply_optimizer(net, lr, masked, 1.0f)) return -1;
    if (out_loss_pinned) AMS_CUDA_CHECK(cudaMemcpyAsync(out_loss_pinned, p->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    return 0;
}

// ---- global-batch BatchNorm for data parallel: statistics exchanged through NVLink peer memory (kernels.cuh SyncBn)
static void drop_train_graphs(Net* net) {
    // the captured steps hold the BN kernels' parameters (with / without the exchange context): capture again
    cudaStreamSynchronize(net->stream);
    for (auto& kv : net->plans) {
        Plan* p = kv.second.get();
        for (int k = 0; k < 2; ++k) {
            if (p->train_graph[k]) { cudaGraphExecDestroy(p->train_graph[k]); p->train_graph[k] = nullptr; }
            p->train_graph_dtype[k] = -1;
        }
        if (p->train_runs < 0) p->train_runs = 1;
    }
}
int ams_syncbn_init(ams_net* h, int world, int rank, void* out_ipc_handle, int handle_capacity) {
    NET(h);
    AMS_REQUIRE(world >= 1 && world <= kSyncBnMaxWorld && rank >= 0 && rank < world, "world must be 1..8, rank inside it");
    AMS_REQUIRE(handle_capacity >= static_cast<int>(sizeof(cudaIpcMemHandle_t)), "handle buffer too small (64 bytes)");
    AMS_REQUIRE(!net->syncbn_recv, "ams_syncbn_init was already called on this handle");
    SyncBn& sb = net->syncbn_host;
    sb = SyncBn{};
    sb.world = world; sb.rank = rank; sb.epoch = 0; sb.error = 0;
    sb.words_per_src = net->n_bnpool / 6 * 8;
    sb.timeout_ns = 10000000000ull;
    if (const char* e = getenv("AMS_SYNCBN_TIMEOUT_MS")) sb.timeout_ns = static_cast<unsigned long long>(atoll(e)) * 1000000ull;
    const size_t bytes = static_cast<size_t>(2) * world * sb.words_per_src * sizeof(unsigned long long);
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&net->syncbn_recv), bytes));
    AMS_CUDA_CHECK(cudaMemset(net->syncbn_recv, 0, bytes));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&net->syncbn_dev), sizeof(SyncBn)));
    AMS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    AMS_CUDA_CHECK(cudaIpcGetMemHandle(&hd, net->syncbn_recv));
    memcpy(out_ipc_handle, &hd, sizeof(hd));
    return 0;
}
int ams_syncbn_connect(ams_net* h, const void* all_handles, int count) {
    NET(h);
    AMS_REQUIRE(net->syncbn_recv, "call ams_syncbn_init first");
    SyncBn& sb = net->syncbn_host;
    AMS_REQUIRE(count == sb.world, "one handle per rank");
    for (int r = 0; r < sb.world; ++r) {
        if (r == sb.rank) { sb.peer[r] = net->syncbn_recv; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, static_cast<const char*>(all_handles) + static_cast<size_t>(r) * sizeof(hd), sizeof(hd));
        void* ptr = nullptr;
        AMS_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        net->syncbn_mapped[r] = ptr;
        sb.peer[r] = static_cast<unsigned long long*>(ptr);
    }
    AMS_CUDA_CHECK(cudaMemcpy(net->syncbn_dev, &sb, sizeof(sb), cudaMemcpyHostToDevice));
    drop_train_graphs(net);
    net->syncbn_enabled = true;
    return 0;
}
int ams_syncbn_enable(ams_net* h, int on) {
    NET(h);
    AMS_REQUIRE(!on || net->syncbn_dev, "ams_syncbn_connect has not been called");
    if (net->syncbn_enabled != (on != 0)) { drop_train_graphs(net); net->syncbn_enabled = on != 0; }
    return 0;
}
int ams_syncbn_status(ams_net* h, unsigned int* out_epoch, unsigned int* out_error) {
    NET(h);
    SyncBn sb{};
    if (net->syncbn_dev) {
        AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
        AMS_CUDA_CHECK(cudaMemcpy(&sb, net->syncbn_dev, sizeof(sb), cudaMemcpyDeviceToHost));
    }
    if (out_epoch) *out_epoch = sb.epoch;
    if (out_error) *out_error = sb.error;
    return 0;
}

int ams_train_step(ams_net* h, float lr, int masked, float* out_loss) {
    NET(h);
    AMS_REQUIRE(!net->frozen, "Can't train frozen graph!!! (handle built by ams_create_frozen is inference-only)");
    Plan* p = nullptr;
    if (net_dequeue(net, &p, true)) return -1;
    if (net_train_fwd_bwd(net, p, true)) return -1;
    if (apply_optimizer(net, lr, masked, 1.0f)) return -1;
    if (out_loss) {
        AMS_CUDA_CHECK(cudaMemcpyAsync(out_loss, p->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, net->stream));
        AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    }
    return 0;
}

int ams_set_mask(ams_net* h, const uint8_t* host_mask) {
    NET(h);
    if (!host_mask) {
        AMS_CUDA_CHECK(cudaMemsetAsync(net->mask, 1, net->n_train, net->stream));
        net->mask_all_ones = true;
    } else {
        AMS_CUDA_CHECK(cudaMemcpyAsync(net->mask, host_mask, net->n_train, cudaMemcpyHostToDevice, net->stream));
        AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
        net->mask_all_ones = false;
    }
    return 0;
}
int ams_get_mask(ams_net* h, uint8_t* host_mask) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemcpyAsync(host_mask, net->mask, net->n_train, cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}
int ams_snapshot_before(ams_net* h) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemcpyAsync(net->before, net->params, net->n_train * sizeof(float), cudaMemcpyDeviceToDevice, net->stream));
    return 0;
}

static void percentile_rank(double coord_frac, long long n, long long* lo, double* w_hi) {
    // NumPy 1.19: q = 100*(1-frac); q /= 100; idx = q*(n-1); lo = floor(idx); w_hi = idx - lo
    const double q = 100 * (1 - coord_frac);
    const double idx = (q / 100.0) * static_cast<double>(n - 1);
    long long l = static_cast<long long>(std::floor(idx));
    if (l > n - 1) l = n - 1;
    if (l < 0) l = 0;
    *lo = l;
    *w_hi = idx - static_cast<double>(l);
}

int ams_select_topk(ams_net* h, double coord_frac, long long* out_kept, float* out_thr) {
    NET(h);
    long long lo; double w_hi;
    percentile_rank(coord_frac, net->n_train, &lo, &w_hi);
    if (select_coordinates(net->params, net->before, net->delta_scratch, net->mask, net->n_train, lo, w_hi, net->select_sc, net->stream)) return -1;
    SelectScratch sc;
    AMS_CUDA_CHECK(cudaMemcpyAsync(&sc, net->select_sc, sizeof(sc), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->mask_all_ones = false;
    net->weights_dirty = true; net->fold_dirty = true;
    if (out_kept) *out_kept = static_cast<long long>(sc.kept);
    if (out_thr) *out_thr = sc.threshold;
    return 0;
}

int ams_pack_delta(ams_net* h, uint8_t* out, long long cap, long long* out_len) {
    NET(h);
    if (pack_delta(net->params, net->mask, net->segs_dev, static_cast<int>(net->trainable_order.size()), net->n_train,
                   net->mask_bytes, net->pack_bits, net->pack_vals, net->pack_counts, pack_delta_blocks(net->n_train),
                   net->pack_kept, net->stream)) return -1;
    unsigned long long kept = 0;
    AMS_CUDA_CHECK(cudaMemcpyAsync(&kept, net->pack_kept, sizeof(kept), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    const long long len = net->mask_bytes + 2 * static_cast<long long>(kept);
    if (out_len) *out_len = len;
    if (out && cap >= len) {
        AMS_CUDA_CHECK(cudaMemcpyAsync(out, net->pack_bits, net->mask_bytes, cudaMemcpyDeviceToHost, net->stream));
        if (kept) AMS_CUDA_CHECK(cudaMemcpyAsync(out + net->mask_bytes, net->pack_vals, 2 * kept, cudaMemcpyDeviceToHost, net->stream));
        AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    }
    return 0;
}

int ams_apply_delta(ams_net* h, const uint8_t* blob, long long len, long long* out_updated) {
    NET(h);
    AMS_REQUIRE(blob && len >= net->mask_bytes, "delta shorter than its mask section");
    const int nblocks = pack_delta_blocks(net->n_train);
    AMS_CUDA_CHECK(cudaMemcpyAsync(net->pack_bits, blob, net->mask_bytes, cudaMemcpyHostToDevice, net->stream));
    // the bits are unpacked into a SCRATCH byte map and the length is validated before any state of the handle changes:
    // a truncated or corrupt delta must leave the training mask and the parameters as they were
    uint8_t* scratch_mask = reinterpret_cast<uint8_t*>(net->delta_scratch);          // n_train floats >= n_train bytes
    if (unpack_delta_mask(net->pack_bits, net->segs_dev, static_cast<int>(net->trainable_order.size()), net->n_train, net->mask_bytes,
                          scratch_mask, net->pack_counts, nblocks, net->pack_kept, net->stream)) return -1;
    unsigned long long kept = 0;
    AMS_CUDA_CHECK(cudaMemcpyAsync(&kept, net->pack_kept, sizeof(kept), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    AMS_REQUIRE(len == net->mask_bytes + 2 * static_cast<long long>(kept), "delta length does not match its mask (" +
                std::to_string(len) + " bytes for " + std::to_string(kept) + " selected coordinates)");
    AMS_CUDA_CHECK(cudaMemcpyAsync(net->mask, scratch_mask, net->n_train, cudaMemcpyDeviceToDevice, net->stream));
    net->mask_all_ones = false;
    if (kept) {
        AMS_CUDA_CHECK(cudaMemcpyAsync(net->pack_vals, blob + net->mask_bytes, 2 * kept, cudaMemcpyHostToDevice, net->stream));
        if (unpack_delta_values(net->params, net->mask, net->n_train, net->pack_counts, nblocks, net->pack_vals, net->stream)) return -1;
        AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    }
    net->weights_dirty = true; net->fold_dirty = true;
    if (out_updated) *out_updated = static_cast<long long>(kept);
    return 0;
}

int ams_get_logits(ams_net* h, float* host, long long count) {
    NET(h);
    auto it = net->plans.find(net->last_n);
    AMS_REQUIRE(it != net->plans.end(), "no forward pass has run yet");
    Plan* p = it->second.get();
    const LayerDef& lg = net->layers.back();
    const long long M = static_cast<long long>(p->N) * lg.out_h * lg.out_w;
    AMS_REQUIRE(count == M * lg.cout, "logits element count mismatch");
    AMS_CUDA_CHECK(cudaMemcpy2DAsync(host, lg.cout * sizeof(float), p->logits, 32 * sizeof(float), lg.cout * sizeof(float), M,
                                     cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}
int ams_get_gradients(ams_net* h, float* host) {
    NET(h);
    AMS_CUDA_CHECK(cudaMemcpyAsync(host, net->grads, net->n_train * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}
int ams_num_layers(const ams_net* h) { return h ? static_cast<int>(reinterpret_cast<const Net*>(h)->layers.size()) : -1; }
int ams_layer_info(const ams_net* h, int index, char* name, int cap, int* kind, int* cin, int* cout, int* stride, int* dil,
                   int* act, float* eps, float* k, int* residual_from) {
    const Net* net = reinterpret_cast<const Net*>(h);
    if (!net || index < 0 || index >= static_cast<int>(net->layers.size())) { set_last_error("bad layer index"); return -1; }
    const LayerDef& d = net->layers[index];
    if (name && cap > 0) { std::strncpy(name, d.name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (kind) *kind = d.kind; if (cin) *cin = d.cin; if (cout) *cout = d.cout; if (stride) *stride = d.stride;
    if (dil) *dil = d.dil; if (act) *act = d.act; if (eps) *eps = d.eps; if (k) *k = d.one_minus_decay;
    if (residual_from) *residual_from = d.residual;
    return 0;
}
int ams_get_activation(ams_net* h, int index, int which, uint16_t* host, long long count) {
    NET(h);
    auto it = net->plans.find(net->last_n);
    AMS_REQUIRE(it != net->plans.end(), "no forward pass has run yet");
    AMS_REQUIRE(index >= 0 && index < static_cast<int>(net->layers.size()), "bad layer index");
    Plan* p = it->second.get();
    const LayerDef& d = net->layers[index];
    const void* src = which == 0 ? static_cast<const void*>(p->buf[index].y)
                                 : (which == 1 ? static_cast<const void*>(p->buf[index].z) : static_cast<const void*>(p->buf[index].g));
    AMS_REQUIRE(src != nullptr, "layer has no such buffer");
    const long long M = static_cast<long long>(p->N) * d.out_h * d.out_w;
    if (which == 0 && p->last_was_train && p->lazy_y[index]) {
        // batch-statistics passes never materialise this layer's normalised output (its consumer applies the BN while
        // staging): produce it on demand for the parity hooks, with the same arithmetic
        const float* pool = net->bnpool + d.bn_off;
        if (bn_apply(p->buf[index].z, pool, pool + d.cout, d.act, nullptr, p->buf[index].y, M, d.cout, net->stream)) return -1;
    }
    AMS_REQUIRE(count == M * d.cout, "activation element count mismatch");
    AMS_CUDA_CHECK(cudaMemcpyAsync(host, src, count * 2, cudaMemcpyDeviceToHost, net->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    return 0;
}

// =============================================================================================== profiler
int ams_profile_enable(ams_net* h, int on) {
    NET(h);
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->prof.reset();
    net->prof.enabled = on != 0;
    net->prof.per_layer = on == 2;
    return 0;
}
int ams_profile_report(ams_net* h, char* buf, int cap) {
    NET(h);
    AMS_CUDA_CHECK(cudaStreamSynchronize(net->stream));
    net->prof.collect();
    std::string out;
    for (const std::string& tag : net->prof.order) {
        const ProfAgg& a = net->prof.agg[tag];
        char line[256];
        snprintf(line, sizeof(line), "%s %lld %.6f %.0f\n", tag.c_str(), a.launches, a.ms, a.algo_bytes);
        out += line;
    }
    if (buf && cap > 0) { std::strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0; }
    return static_cast<int>(out.size());
}

// =============================================================================================== host-only layout
static Net* layout_net(int num_classes, int variant) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::unique_ptr<Net>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(num_classes, variant);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::unique_ptr<Net> n(new Net());
        n->cfg.num_classes = num_classes; n->cfg.graph_variant = variant; n->cfg.height = 512; n->cfg.width = 1024;
        if (net_build_topology(n.get())) return nullptr;
        it = cache.emplace(key, std::move(n)).first;
    }
    return it->second.get();
}
int ams_layout_num_tensors(int nc, int variant) {
    Net* n = layout_net(nc, variant);
    return n ? static_cast<int>(n->vars.size()) : -1;
}
int ams_layout_tensor_info(int nc, int variant, int index, char* name, int cap, int shape4[4], int* ndim, int* trainable, long long* off) {
    Net* n = layout_net(nc, variant);
    if (!n) return -1;
    return ams_tensor_info(reinterpret_cast<const ams_net*>(n), index, name, cap, shape4, ndim, trainable, off);
}
int ams_layout_num_layers(int nc, int variant) {
    Net* n = layout_net(nc, variant);
    return n ? static_cast<int>(n->layers.size()) : -1;
}
int ams_layout_layer_info(int nc, int variant, int index, char* name, int cap, int* kind, int* cin, int* cout, int* stride,
                          int* dil, int* act, float* eps, float* k, int* residual_from) {
    Net* n = layout_net(nc, variant);
    if (!n) return -1;
    return ams_layer_info(reinterpret_cast<const ams_net*>(n), index, name, cap, kind, cin, cout, stride, dil, act, eps, k, residual_from);
}

// =============================================================================================== op-level hooks
int ams_op_conv1x1(const void* a, const void* w, int M, int N, int K, const float* scale, const float* shift,
                   const float* rowbias, int rows_per_image, const void* residual, int act, void* out, int out_fp32, int ldc,
                   int grad_types, const void* w_lo, void* stream) {
    GemmDesc d;
    d.A = a; d.lda = K; d.B = w; d.ldb = K; d.B_lo = w_lo;
    if (grad_types) { d.a_fp16 = 0; d.b_fp16 = 0; d.out_fp16 = 0; }     // data-gradient GEMM: bf16 gradient x bf16 weights -> bf16
    d.M = M; d.N = N; d.K = K; d.out = out; d.ldc = ldc; d.out_fp32 = out_fp32; d.scale = scale; d.shift = shift;
    d.rowbias = rowbias; d.rows_per_image = rows_per_image > 0 ? rows_per_image : 1;
    d.residual = residual; d.ldr = N; d.act = act;
    int dev = 0; cudaGetDevice(&dev);
    int sms = kNumSMs; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    GemmPlan pl;
    if (gemm_plan(d, sms, &pl)) return -1;
    return gemm_launch(pl, as_stream(stream));
}

static void* g_fused_debug_timeline = nullptr;
/* diagnostics (tools/micro/fused_block_run.py): device buffer of 3 x 64 x 4 u64 that the next ams_op_fused_block fills with the
 * timeline (ns) of CTA 0: [MMA issuer | group E | group W][chunk][event] */
int ams_debug_fused_timeline(void* device_buffer) { g_fused_debug_timeline = device_buffer; return 0; }

int ams_op_fused_block(const void* x, int n, int h, int w_, int cin, int cexp, int cout, int dil, int stride, const void* we, const void* we_lo,
                       const float* s1, const float* t1, const float* wd, const float* s2, const float* t2, const void* wp, const void* wp_lo,
                       const float* s3, const float* t3, int residual, void* out, void* stream) {
    int dev = 0; cudaGetDevice(&dev);
    int sms = kNumSMs; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    FusedBlockDesc d;
    d.N = n; d.H = h; d.W = w_; d.Cin = cin; d.Cexp = cexp; d.Cout = cout; d.stride = stride; d.dil = dil;
    d.x = x; d.We = we; d.We_lo = we_lo; d.ld_we = cin; d.Wp = wp; d.Wp_lo = wp_lo; d.ld_wp = cexp;
    d.s3 = s3; d.t3 = t3; d.residual = residual ? x : nullptr; d.out = out;
    d.debug_timeline = g_fused_debug_timeline;
    AMS_REQUIRE(fused_block_supported(d), "fused block: unsupported geometry");
    float* params = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&params), fused_block_param_floats(d) * sizeof(float)));
    d.params = params;
    FusedBlockPlan pl;
    int rc = fused_block_fill_params(d, s1, t1, wd, s2, t2, as_stream(stream));
    if (!rc) rc = fused_block_plan(d, sms, &pl);
    if (!rc) rc = fused_block_launch(pl, as_stream(stream));
    const cudaError_t e = cudaStreamSynchronize(as_stream(stream));
    cudaFree(params);
    if (!rc && e != cudaSuccess) { set_last_error(std::string("fused block kernel: ") + cudaGetErrorString(e)); rc = -1; }
    return rc;
}

int ams_op_wgrad(const void* x, int cin, const void* dz, int cout, long long M, float* dw, void* stream) {
    int dev = 0; cudaGetDevice(&dev);
    int sms = kNumSMs; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgradDesc d;
    d.X = x; d.ldx = cin; d.Cin = cin; d.dZ = dz; d.ldz = cout; d.Cout = cout;      // X fp16 activations, dZ bf16 gradients
    d.M = M; d.dW = dw; d.lddw = cout;
    const size_t wsf = wgrad_workspace_floats(cin, cout, M, sms);
    float* ws = nullptr;
    if (wsf) AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), wsf * sizeof(float)));
    d.workspace = ws; d.workspace_floats = wsf;
    WgradPlan pl;
    int rc = wgrad_plan(d, sms, &pl);
    if (!rc) rc = wgrad_launch(pl, as_stream(stream));
    if (ws) { cudaStreamSynchronize(as_stream(stream)); cudaFree(ws); }
    return rc;
}

static Conv2dGeom make_geom(int n, int h, int w, int c, int stride, int dil) {
    Conv2dGeom g;
    g.N = n; g.H = h; g.W = w; g.C = c; g.stride = stride; g.dil = dil;
    g.Ho = (h + stride - 1) / stride; g.Wo = (w + stride - 1) / stride;
    g.pad_top = std::max((g.Ho - 1) * stride + 2 * dil + 1 - h, 0) / 2;
    g.pad_left = std::max((g.Wo - 1) * stride + 2 * dil + 1 - w, 0) / 2;
    return g;
}

int ams_op_depthwise(const void* in, const float* w, int n, int h, int w_, int c, int stride, int dil, const float* scale,
                     const float* shift, int act, void* out, void* stream) {
    return dw_conv_fwd_tiled(static_cast<const act_t*>(in), w, make_geom(n, h, w_, c, stride, dil), nullptr, nullptr, 0, scale,
                             shift, act, static_cast<act_t*>(out), nullptr, nullptr, as_stream(stream));
}

__global__ void sum_rows_kernel(const double* __restrict__ partial, int rows, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0;
    for (int r = 0; r < rows; ++r) a += partial[static_cast<long long>(r) * n + i];
    out[i] = a;
}

int ams_op_depthwise_fused(const void* in, const float* w, int n, int h, int w_, int c, int stride, int dil,
                           const float* in_scale, const float* in_shift, int in_act, void* out, double* stats_out,
                           void* stream) {
    const Conv2dGeom g = make_geom(n, h, w_, c, stride, dil);
    double* ws = nullptr;
    const long long rows = dw_tiled_stats_rows(g);
    if (stats_out) AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), static_cast<size_t>(rows) * 2 * c * sizeof(double)));
    int got = 0;
    int rc = dw_conv_fwd_tiled(static_cast<const act_t*>(in), w, g, in_scale, in_shift, in_act, nullptr, nullptr, 0,
                               static_cast<act_t*>(out), ws, &got, as_stream(stream));
    if (!rc && stats_out) {
        sum_rows_kernel<<<ceil_div(2 * c, 128), 128, 0, as_stream(stream)>>>(ws, got, 2 * c, stats_out);
        if (cudaGetLastError() != cudaSuccess) rc = -1;
    }
    cudaStreamSynchronize(as_stream(stream));
    if (ws) cudaFree(ws);
    return rc;
}

int ams_op_resize_u8(const void* src, int n, int src_h, int src_w, int channels, void* dst, int dst_h, int dst_w, int nearest,
                     int swap_rb, void* stream) {
    return resize_u8(static_cast<const uint8_t*>(src), n, src_h, src_w, channels, static_cast<uint8_t*>(dst), dst_h, dst_w, nearest,
                     swap_rb, as_stream(stream));
}

int ams_op_depthwise_bwd_fused(const void* g, const void* z, const float* scale, const float* shift, int act, const float* coef,
                               const void* zin, const float* in_scale, const float* in_shift, int in_act, const float* w,
                               int n, int h, int w_, int c, int stride, int dil, void* gout, float* dw, double* bn_sums,
                               void* stream) {
    const Conv2dGeom gm = make_geom(n, h, w_, c, stride, dil);
    const long long rows = dw_bwd_fused_rows(gm);
    float* wsf = nullptr; double* wsd = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&wsf), static_cast<size_t>(rows) * 9 * c * sizeof(float)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&wsd), static_cast<size_t>(rows) * 2 * c * sizeof(double)));
    DwBwdFused f;
    f.g = static_cast<const bf16*>(g); f.z = static_cast<const act_t*>(z); f.scale = scale; f.shift = shift; f.act = act; f.coef = coef;
    f.zin = static_cast<const act_t*>(zin); f.in_scale = in_scale; f.in_shift = in_shift; f.in_act = in_act; f.w = w;
    f.gout = static_cast<bf16*>(gout); f.dw = dw; f.dw_partial = wsf; f.dw_partial_floats = static_cast<size_t>(rows) * 9 * c;
    f.bn_partial = wsd;
    int got = 0;
    int rc = dw_conv_bwd_fused(f, gm, &got, as_stream(stream));
    if (!rc && bn_sums && in_scale) {
        sum_rows_kernel<<<ceil_div(2 * c, 128), 128, 0, as_stream(stream)>>>(wsd, got, 2 * c, bn_sums);
        if (cudaGetLastError() != cudaSuccess) rc = -1;
    }
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(wsf); cudaFree(wsd);
    return rc;
}

int ams_op_depthwise_bwd(const void* x, const void* dz, const float* w, int n, int h, int w_, int c, int stride, int dil,
                         void* dx, float* dw, void* stream) {
    const Conv2dGeom g = make_geom(n, h, w_, c, stride, dil);
    if (dx && dw_conv_bwd_data(static_cast<const bf16*>(dz), w, g, static_cast<bf16*>(dx), as_stream(stream))) return -1;
    if (dw) {
        const size_t wsf = dw_bwd_workspace_floats(g);
        float* ws = nullptr;
        AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), wsf * sizeof(float)));
        int rc = dw_conv_bwd_filter(static_cast<const act_t*>(x), static_cast<const bf16*>(dz), g, dw, ws, wsf, as_stream(stream));
        cudaStreamSynchronize(as_stream(stream));
        cudaFree(ws);
        return rc;
    }
    return 0;
}

static void stem_geom(int h, int w, int* Hp, int* Wp, int* Ho, int* Wo, int* pt, int* pl) {
    *Hp = h + 1; *Wp = w + 1;
    *Ho = (*Hp + 1) / 2; *Wo = (*Wp + 1) / 2;
    *pt = std::max((*Ho - 1) * 2 + 3 - *Hp, 0) / 2;
    *pl = std::max((*Wo - 1) * 2 + 3 - *Wp, 0) / 2;
}

int ams_op_stem(const void* frames, int dtype, int n, int h, int w_, const float* w, const float* scale, const float* shift,
                void* out, void* stream) {
    int Hp, Wp, Ho, Wo, pt, pl;
    stem_geom(h, w_, &Hp, &Wp, &Ho, &Wo, &pt, &pl);
    return stem_conv_fwd(frames, dtype == AMS_FRAMES_U8, n, h, w_, Hp, Wp, Ho, Wo, pt, pl, 127.5f, 0.007843137718737125f, 1.0f, w,
                         scale, shift, static_cast<act_t*>(out), as_stream(stream));
}

int ams_op_stem_bwd(const void* frames, int dtype, int n, int h, int w_, const void* dz, float* dw, void* stream) {
    int Hp, Wp, Ho, Wo, pt, pl;
    stem_geom(h, w_, &Hp, &Wp, &Ho, &Wo, &pt, &pl);
    const size_t wsf = stem_bwd_workspace_floats(n, Ho, Wo);
    float* ws = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), wsf * sizeof(float)));
    int rc = stem_conv_bwd_filter(frames, dtype == AMS_FRAMES_U8, n, h, w_, Hp, Wp, Ho, Wo, pt, pl, 127.5f, 0.007843137718737125f,
                                  1.0f, static_cast<const bf16*>(dz), dw, ws, wsf, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(ws);
    return rc;
}

int ams_op_bn_train(const void* z, long long M, int C, const float* gamma, const float* beta, float eps, int act,
                    const void* residual, void* y, float* mean, float* rstd, void* stream) {
    float* tmp = nullptr; double* ws = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tmp), 4 * C * sizeof(float)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), bn_workspace_doubles(M, C) * sizeof(double)));
    BnLayer L;
    L.C = C; L.M = M; L.eps = eps; L.one_minus_decay = 0.f; L.gamma = gamma; L.beta = beta;
    L.moving_mean = tmp + 2 * C; L.moving_var = tmp + 3 * C; L.mean = mean; L.rstd = rstd; L.scale = tmp; L.shift = tmp + C;
    int rc = bn_forward_stats(static_cast<const act_t*>(z), L, 0, ws, as_stream(stream));
    if (!rc) rc = bn_apply(static_cast<const act_t*>(z), L.scale, L.shift, act, static_cast<const act_t*>(residual),
                           static_cast<act_t*>(y), M, C, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(tmp); cudaFree(ws);
    return rc;
}

int ams_op_bn_backward(const void* dy, const void* z, long long M, int C, const float* gamma, const float* beta, float eps,
                       int act, void* dz, float* dgamma, float* dbeta, void* stream) {
    float* tmp = nullptr; double* ws = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tmp), 6 * C * sizeof(float)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ws), bn_workspace_doubles(M, C) * sizeof(double)));
    BnLayer L;
    L.C = C; L.M = M; L.eps = eps; L.one_minus_decay = 0.f; L.gamma = gamma; L.beta = beta;
    L.moving_mean = tmp + 4 * C; L.moving_var = tmp + 5 * C; L.mean = tmp + 2 * C; L.rstd = tmp + 3 * C; L.scale = tmp; L.shift = tmp + C;
    int rc = bn_forward_stats(static_cast<const act_t*>(z), L, 0, ws, as_stream(stream));
    if (!rc) rc = bn_backward(static_cast<const bf16*>(dy), nullptr, static_cast<const act_t*>(z), L, act, static_cast<bf16*>(dz),
                              dgamma, dbeta, ws, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(tmp); cudaFree(ws);
    return rc;
}

static int op_head_geom(int n, int h, int w, int ldl, int H, int W, int cc, const int* cls, int depth, HeadGeom* g) {
    ams_config c{};
    c.num_classes = 32; c.height = H; c.width = W; c.class_count = cc; c.label_depth = depth;
    for (int i = 0; i < cc && i < AMS_MAX_CLASSES; ++i) c.class_indices[i] = cls[i];
    if (fill_head_geom(c, h, w, g)) return -1;
    g->N = n; g->ldl = ldl;
    return 0;
}

int ams_op_head_infer(const float* logits, int n, int h, int w_, int ldl, int H, int W, int cc, const int* cls, int depth,
                      const uint8_t* labels, int32_t* pred, int64_t* confmat, double* loss_sum, long long* n_valid, void* stream) {
    HeadGeom g;
    if (op_head_geom(n, h, w_, ldl, H, W, cc, cls, depth, &g)) return -1;
    HeadStats* st = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&st), sizeof(HeadStats)));
    int rc = head_reset(st, as_stream(stream));
    if (!rc) rc = head_infer(logits, g, labels, pred, st, as_stream(stream));
    HeadStats hs;
    cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(st);
    if (confmat) for (int i = 0; i < cc; ++i) for (int j = 0; j < cc; ++j) confmat[i * cc + j] = hs.confmat[i * kMaxClasses + j];
    if (loss_sum) *loss_sum = static_cast<double>(hs.loss_fixed) / kLossFixedScale;
    if (n_valid) *n_valid = hs.n_valid;
    return rc;
}

int ams_op_head_backward(const float* logits, int n, int h, int w_, int H, int W, int cc, const int* cls, int depth,
                         const uint8_t* labels, float* dlogits, float* loss, void* stream) {
    HeadGeom g;
    if (op_head_geom(n, h, w_, 32, H, W, cc, cls, depth, &g)) return -1;
    g.normalize = 1;
    HeadStats* st = nullptr; float* rowbuf = nullptr; bf16* d16 = nullptr; float* loss_dev = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&st), sizeof(HeadStats)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&rowbuf), head_rowbuf_floats(g) * sizeof(float)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d16), static_cast<size_t>(n) * h * w_ * 32 * sizeof(bf16)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&loss_dev), sizeof(float)));
    int rc = head_loss_backward(logits, g, labels, rowbuf, dlogits, d16, st, loss_dev, as_stream(stream));
    if (loss) cudaMemcpyAsync(loss, loss_dev, sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(st); cudaFree(rowbuf); cudaFree(d16); cudaFree(loss_dev);
    return rc;
}

int ams_op_select(float* after, const float* before, long long n, double coord_frac, uint8_t* mask, long long* kept,
                  float* threshold, void* stream) {
    long long lo; double w_hi;
    percentile_rank(coord_frac, n, &lo, &w_hi);
    float* d = nullptr; SelectScratch* sc = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d), n * sizeof(float)));
    AMS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sc), sizeof(SelectScratch)));
    int rc = select_coordinates(after, before, d, mask, n, lo, w_hi, sc, as_stream(stream));
    SelectScratch hs;
    cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(d); cudaFree(sc);
    if (kept) *kept = static_cast<long long>(hs.kept);
    if (threshold) *threshold = hs.threshold;
    return rc;
}

int ams_op_adam(float* p, const float* g, float* m, float* v, const uint8_t* mask, long long n, float lr, float b1p, float b2p,
                void* stream) {
    const float alpha = lr * std::sqrt(1.0f - b2p) / (1.0f - b1p);
    return adam_masked(p, g, 1.0f, nullptr, m, v, mask, n, alpha, 1.0f - 0.9f, 1.0f - 0.999f, 1e-8f, as_stream(stream));
}

}  // extern "C"
