// The network handle behind the C ABI: topology (restated from the shipped model.meta, checked against
// ams_b200/graphs/*.json by tests/test_layout.py), flat parameter arenas, per-batch-size execution plans.
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ams_b200.h"
#include "fused_block.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace ams {

enum LayerKind { kStem = 0, kConv1x1 = 1, kDepthwise = 2, kImagePool = 3, kLogits = 4 };
// Forward 1x1 convs with at most this many output channels (one N tile: every project conv, the head, the early expand
// convs) run on split fp16 weights W = hi + lo (two MMAs per k-step against the same A tile).  Rounding the PROJECT
// weights to plain fp16 was the largest single term of the end-to-end logit error (oracle ablation); for
// the wide expand convs the second plane buys nothing measurable and would cost their 2-CTAs-per-SM shared-memory budget.
constexpr int kSplitWeightMaxCout = 256;

struct VarInfo {
    std::string name;
    int shape[4] = {0, 0, 0, 0};
    int ndim = 0;
    bool trainable = false;
    long long offset = 0;      // floats into the trainable arena (trainable) or the moving arena
    long long count = 0;
};

struct LayerDef {
    std::string name;
    int kind = 0;
    int cin = 0, cout = 0, stride = 1, dil = 1, act = 0;
    bool has_bn = true;
    float eps = 0.f, one_minus_decay = 0.f;
    int input = -1;            // producer layer (-1 = frame)
    int residual = -1;         // layer whose output is added after BN (-1 = none)
    // parameter offsets (floats): trainable arena / moving arena
    long long w_off = -1, gamma_off = -1, beta_off = -1, bias_off = -1, mm_off = -1, mv_off = -1;
    long long bn_off = -1;     // offset into the per-layer BN vector pool (units of floats, 6 vectors of cout)
    // bf16 weight copies (elements into the bf16 pool), 1x1 layers only
    long long wfwd_off = -1, wbwd_off = -1;
    long long wlo_off = -1;    // low plane of the split forward weights (layers with cout <= kSplitWeightMaxCout), else -1
    int ld_fwd = 0, ld_bwd = 0, k_rows0 = 0, k_rows = 0;   // HWIO row window used by the GEMM (concat_projection: 256..511)
    // geometry at the configured frame size (per image)
    int in_h = 0, in_w = 0, out_h = 0, out_w = 0, pad_top = 0, pad_left = 0;
};

struct LayerBuf { act_t* z = nullptr; act_t* y = nullptr; bf16* g = nullptr; bf16* gz = nullptr; };   // forward fp16, gradients bf16

struct Plan {
    int N = 0;
    std::vector<LayerBuf> buf;
    std::vector<void*> allocations;
    void* in_frames = nullptr; uint8_t* in_labels = nullptr; int in_dtype = 0; bool in_has_labels = false;
    float* logits = nullptr;          // [N*h*w][32] fp32
    float* dlogits_f32 = nullptr; bf16* dlogits_bf16 = nullptr; float* rowbuf = nullptr;
    int32_t* pred = nullptr;
    float* loss_dev = nullptr;
    // image pooling scratch
    float *pooled = nullptr, *ip_z = nullptr, *ip_act = nullptr, *bias_img = nullptr, *ip_dbias = nullptr, *dfeat_rowbias = nullptr;
    double* bn_ws = nullptr; double* small_ws = nullptr;
    float* red_ws = nullptr; size_t red_ws_floats = 0;
    float* wgrad_ws = nullptr; size_t wgrad_ws_floats = 0;       // split-K partials of the side-stream filter gradients
    std::vector<GemmPlan> fwd_frozen, fwd_train, dgrad;
    std::vector<WgradPlan> wgrad;
    std::vector<char> has_fwd, has_dgrad, has_wgrad;
    // frozen inference: fused[i] != 0 for the EXPAND layer i of a block that runs as one kernel (fused_block.cu); the
    // expand / depthwise outputs of such a block are never materialised
    std::vector<FusedBlockPlan> fused_plan; std::vector<char> fused;
    // training-mode fusion (dw_tiled.cu): dw_fused[i] = depthwise layer i runs the fused backward and reads its
    // producer's RAW output; lazy_y[j] = layer j's normalised output is never materialised in training mode
    std::vector<char> dw_fused, lazy_y;
    float* coef[2] = {nullptr, nullptr};       // BN-backward coefficients [3][Cmax]: [0] depthwise layer, [1] its producer
    int pending_bn_rows = 0;
    // whole forward+backward captured once per (loss normalisation, frame dtype) and replayed (net_train_fwd_bwd)
    cudaGraphExec_t train_graph[2] = {nullptr, nullptr}; int train_graph_dtype[2] = {-1, -1}; int train_runs = 0;
    long long train_graph_kernels[2] = {0, 0};
    // frozen inference (forward + head), keyed by [with metric]
    cudaGraphExec_t infer_graph[2] = {nullptr, nullptr}; int infer_graph_dtype[2] = {-1, -1}; int infer_runs = 0;
    long long infer_graph_kernels[2] = {0, 0};                   // > 0: bn_ws holds the fused kernel's column sums for the next layer
    bool have_backward = false;
    bool last_was_train = false;
    // split frozen inference (net_half_plans): two half-batch VIEWS of this plan -- same activation / logits / prediction
    // buffers at an image offset, own GEMM tensor maps and small scratch -- run on two streams inside one graph
    Plan* parent = nullptr; int part = 0;
    std::unique_ptr<Plan> half[2];
};

// The streams and events one forward pass is enqueued on (default: the handle's main + side stream).
struct StreamSet { cudaStream_t main = nullptr, side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };

// Per-launch-group device timing (CUDA events on the launching stream), aggregated by tag.  Off by default.
struct ProfEntry { std::string tag; double algo_bytes = 0; cudaEvent_t e0 = nullptr, e1 = nullptr; };
struct ProfAgg { long long launches = 0; double ms = 0, algo_bytes = 0; };
struct Profiler {
    bool enabled = false;
    bool per_layer = false;            // tags become "group@layer" (diagnostics: tools/profile_layers.py)
    const char* layer = nullptr;
    std::vector<ProfEntry> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, ProfAgg> agg;
    std::vector<std::string> order;
    cudaEvent_t get_event();
    void begin(cudaStream_t s, const char* tag, double algo_bytes);
    void end(cudaStream_t s);
    void collect();          // after a stream sync: fold pending pairs into agg
    void reset();
};

struct QueueSlot {
    void* frames = nullptr; size_t frames_cap = 0;
    uint8_t* labels = nullptr; size_t labels_cap = 0;
    int n = 0, dtype = 0; bool has_labels = false;
    cudaEvent_t consumed = nullptr; bool consumed_pending = false;
};

struct Net {
    ams_config cfg{};
    int num_sms = kNumSMs;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    // filter gradients of the 1x1 convs run on a side stream, concurrently with the backward chain that does not need them
    cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pool = nullptr, ev_dwred = nullptr;
    // data-parallel bucketed gradient exchange: the late bucket = arena floats [bucket_split, n_train) (every layer of the
    // final-resolution stage, ASPP, logits); ev_bucket is recorded (as an external event node when the step is a CUDA graph)
    // once all of its gradients are complete, long before the backward pass ends
    int bucket_layer = -1; long long bucket_split = 0; cudaEvent_t ev_bucket = nullptr, ev_bucket_main = nullptr;
    std::vector<VarInfo> vars;
    std::unordered_map<std::string, int> var_index;
    std::vector<int> trainable_order;       // indices into vars, tf.trainable_variables() order
    std::vector<LayerDef> layers;
    long long n_train = 0, n_moving = 0, n_bnpool = 0, n_bf16 = 0;
    float *params = nullptr, *moving = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr, *before = nullptr;
    float* delta_scratch = nullptr;
    uint8_t* mask = nullptr; bool mask_all_ones = true;
    float beta1_power = 0.9f, beta2_power = 0.999f;
    float* bnpool = nullptr;                // per layer: scale, shift, mean, rstd (training) + fscale, fshift (frozen)
    uint16_t* wpool = nullptr;              // 16-bit GEMM operand copies of the 1x1 weights: fp16 [Cout][Cin] (forward), bf16 [Cin][Cout] (dgrad)
    WeightCast* cast_table = nullptr; int cast_layers = 0, cast_max = 0;
    VarSeg* segs_dev = nullptr; long long mask_bytes = 0;
    SelectScratch* select_sc = nullptr;
    HeadStats* head_st = nullptr;
    // cross-GPU BatchNorm statistics (data parallel, ams_syncbn_*): device context, local receive buffer, mapped peers
    SyncBn* syncbn_dev = nullptr; SyncBn syncbn_host{}; unsigned long long* syncbn_recv = nullptr;
    void* syncbn_mapped[kSyncBnMaxWorld] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool syncbn_enabled = false;        // connected and switched on
    bool sync_active = false;           // true only while a TRAINING step is being enqueued (inference never exchanges)
    uint8_t* pack_bits = nullptr; __half* pack_vals = nullptr; unsigned int* pack_counts = nullptr; unsigned long long* pack_kept = nullptr;
    bool weights_dirty = true, fold_dirty = true;
    // frozen inference of a batch >= 4 as two concurrent half batches (ams_set_infer_split): the second half's stream set
    bool infer_split = true; StreamSet split_ss{}; cudaEvent_t ev_split_fork = nullptr, ev_split_join = nullptr;
    bool block_fusion = true;          // frozen inference: stride-1 inverted-residual blocks as one kernel each (ams_set_block_fusion)
    std::vector<float*> fused_params;  // per expand layer: [13][cpad] padded per-channel vectors of the fused kernel (or null)
    bool frozen = false;               // built by ams_create_frozen: inference only ("Can't train frozen graph", SemanticNetwork.py:217)
    HeadGeom head{};
    std::map<int, std::unique_ptr<Plan>> plans;
    // input queue
    std::mutex qmu; std::condition_variable qcv;
    std::vector<QueueSlot> slots; std::deque<int> filled; std::deque<int> free_slots;
    int last_n = 0;
    // staging for ams_enqueue_raw (camera-size frames / label maps before the on-device resize), owned by the feeder thread
    uint8_t* raw_frames = nullptr; size_t raw_frames_cap = 0; uint8_t* raw_labels = nullptr; size_t raw_labels_cap = 0;
    Profiler prof;
};

int net_build_topology(Net* net);
Plan* net_get_plan(Net* net, int N, bool need_backward);
int net_prepare_weights(Net* net, bool frozen);
int net_forward(Net* net, Plan* p, int bn_mode, bool update_moving, const StreamSet* ss = nullptr);
// the two half-batch views of an even-sized plan (built on first use); null on failure
int net_half_plans(Net* net, Plan* p, Plan** a, Plan** b);
int net_backward(Net* net, Plan* p, bool normalize);
// forward (batch statistics, moving-average update) + backward of one distillation step: eager the first time on a plan,
// then a CUDA graph replay (~450 launches, two streams) unless profiling or AMS_NO_GRAPH=1
int net_train_fwd_bwd(Net* net, Plan* p, bool normalize);
int net_dequeue(Net* net, Plan** plan_out, bool need_backward);

}  // namespace ams
