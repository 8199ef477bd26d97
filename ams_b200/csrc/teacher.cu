// Teacher network (config C5): DeepLabv3+ with the Xception-65 backbone, inference only, for the label-extraction path
// of the reference (extract_labels.py:51-99; utils/graph_utils.py:129-152 `create_teacher` imports the teacher's
// .meta and runs images:0 -> predictions:0).  The teacher's graph and weights are NOT in the reference repository
// (external download, README.md:45-46), so the topology below restates the PUBLIC model-zoo definition the checkpoint
// was exported from (tensorflow/models research/deeplab: core/xception.py `xception_65`, model.py with
// atrous_rates 6/12/18, output_stride 16, decoder_output_stride 4, separable ASPP / decoder convs -- the
// `xception65_cityscapes_trainfine` configuration); variable names follow that definition so a re-obtained checkpoint
// dict loads by name.  Arithmetic is unpinned and unsourced here: parity is against
// oracle/teacher_oracle.py, a torch restatement of the same public definition.
//
// Kernels: the 1x1 convolutions (96 % of the 0.84 TFLOP of a 1025x2049 frame; K and N up to 2048: the only tensor-bound
// GEMMs of the repository) run on the tcgen05 GEMM of gemm.cu with the folded BN / ReLU / residual epilogue; the second
// stem conv (3x3, 32 -> 64) is an implicit GEMM on tcgen05 fed by nine shifted 4-D TMA boxes per tile (zero padding = TMA
// out-of-bounds fill); depthwise 3x3 (stride 1/2, dilation 1..18, optional ReLU on load), stride-2 subsampling for the
// shortcut convs and the align-corners bilinear resize are small bandwidth kernels; the final upsample + argmax is the
// student's head kernel.
#include "net.cuh"
#include "tcgen05.cuh"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ams {
namespace {

// ============================================================================================ small kernels
// depthwise 3x3, NHWC fp16, C % 8 == 0, runtime stride / dilation; thread = one output pixel x 8 channels.
// pre_relu: ReLU applied to the input on load (Xception applies the activation BEFORE each separable conv of the entry /
// middle flow, while the shortcut branch reads the un-activated tensor); out = act(conv * scale + shift).
__global__ void __launch_bounds__(256)
t_dw_kernel(const act_t* __restrict__ in, const float* __restrict__ w /*[3][3][C]*/, const float* __restrict__ scale,
            const float* __restrict__ shift, act_t* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo, int stride, int dil,
            int pad_top, int pad_left, int pre_relu, int act) {
    pdl_entry();
    const int c8n = C >> 3;
    const long long total = static_cast<long long>(N) * Ho * Wo * c8n;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int c0 = static_cast<int>(tid % c8n) * 8;
    long long pix = tid / c8n;
    const int ox = static_cast<int>(pix % Wo); pix /= Wo;
    const int oy = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - pad_top + ky * dil;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * stride - pad_left + kx * dil;
            if (ix < 0 || ix >= W) continue;
            float v[8];
            unpack8h(__ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * H + iy) * W + ix) * C + c0)), v);
            const float4 w0 = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c0);
            const float4 w1 = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c0 + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(pre_relu ? fmaxf(v[q], 0.f) : v[q], wv[q], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float y = fmaf(acc[q], scale[c0 + q], shift[c0 + q]);
        acc[q] = act ? fmaxf(y, 0.f) : y;
    }
    stg_stream(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + c0, pack8h(acc));
}

// out[n, y, x, :] = in[n, 2y, 2x, :]  (input of a 1x1 stride-2 'SAME' convolution)
__global__ void __launch_bounds__(256)
t_subsample2_kernel(const act_t* __restrict__ in, act_t* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo) {
    pdl_entry();
    const int c8n = C >> 3;
    const long long total = static_cast<long long>(N) * Ho * Wo * c8n;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int c0 = static_cast<int>(tid % c8n) * 8;
    long long pix = tid / c8n;
    const int ox = static_cast<int>(pix % Wo); pix /= Wo;
    const int oy = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    const uint4 v = ldg_stream(in + ((static_cast<long long>(n) * H + 2 * oy) * W + 2 * ox) * C + c0);
    stg_stream(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + c0, v);
}

// ResizeBilinear(align_corners=True) on NHWC fp16, written into a channel slice of a wider tensor (ld_out >= C)
__global__ void __launch_bounds__(256)
t_resize_kernel(const act_t* __restrict__ in, act_t* __restrict__ out, int N, int h, int w, int C, int Ho, int Wo, int ld_out, float sy, float sx) {
    pdl_entry();
    const int c8n = C >> 3;
    const long long total = static_cast<long long>(N) * Ho * Wo * c8n;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int c0 = static_cast<int>(tid % c8n) * 8;
    long long pix = tid / c8n;
    const int ox = static_cast<int>(pix % Wo); pix /= Wo;
    const int oy = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    const float fy = static_cast<float>(oy) * sy, fx = static_cast<float>(ox) * sx;
    const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
    const int y1 = min(static_cast<int>(ceilf(fy)), h - 1), x1 = min(static_cast<int>(ceilf(fx)), w - 1);
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const act_t* base = in + static_cast<long long>(n) * h * w * C + c0;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8h(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * w + x0) * C)), a);
    unpack8h(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * w + x1) * C)), b);
    unpack8h(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * w + x0) * C)), c);
    unpack8h(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * w + x1) * C)), d);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float top = a[q] + (b[q] - a[q]) * lx, bot = c[q] + (d[q] - c[q]) * lx;      // lerp x then y, like the TF kernel
        o[q] = top + (bot - top) * ly;
    }
    stg_stream(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ld_out + c0, pack8h(o));
}

// fp32 HWIO [3][3][32][64] -> fp16 [tap][cout][cin]  (B operand of the implicit GEMM)
__global__ void t_cast_conv3_kernel(const float* __restrict__ w, act_t* __restrict__ out, int Cin, int Cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * Cin * Cout) return;
    const int tap = i / (Cin * Cout), rem = i - tap * Cin * Cout;
    const int ci = rem / Cout, co = rem - ci * Cout;
    out[(static_cast<long long>(tap) * Cout + co) * Cin + ci] = __float2half_rn(w[i]);
}

// ============================================================================================ 3x3 dense conv, 32 -> 64, stride 1
// Implicit GEMM on tcgen05: M tile = 128 consecutive pixels of one image row, N = 64 output channels, K = 9 taps x 32
// input channels.  The A operand of tap (ky, kx) is the same row segment shifted by (ky-1, kx-1): one 4-D TMA box
// [32 ch, 128 px, 1, 1] per tap, zero padding = out-of-bounds fill.  The 9 weight tiles [64 x 32] stay resident.
constexpr int kC3Threads = 192;               // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue
constexpr int kC3Stages = 9;                  // one stage per tap: a whole tile in flight
struct Conv3Params {
    int N, H, W, tiles_x, num_tiles;
    const float* scale; const float* shift;   // folded BN [64]
    act_t* out;
};
__device__ __forceinline__ void t_tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(t5::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(t5::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__global__ void __launch_bounds__(kC3Threads, 1)
t_conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Conv3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kATile = 128 * 64, kBTile = 64 * 64;      // 32 channels = 64-byte rows (SWIZZLE_64B)
    uint8_t* smA = smem;                                     // [stages][128][64 B]
    uint8_t* smB = smem + kC3Stages * kATile;                // [9][64][64 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + 9 * kBTile);
    uint64_t* empty_bar = full_bar + kC3Stages;
    uint64_t* tfull_bar = empty_bar + kC3Stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmX); t5::tma_prefetch_desc(&tmW);
        for (int s = 0; s < kC3Stages; ++s) { t5::mbar_init(&full_bar[s], 1); t5::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { t5::mbar_init(&tfull_bar[s], 1); t5::mbar_init(&tempty_bar[s], 128); }
        t5::mbar_init(w_bar, 1);
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, 128); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 0) {
        if (lane == 0) {
            t5::mbar_arrive_expect_tx(w_bar, 9 * kBTile);
            for (int tap = 0; tap < 9; ++tap) t5::tma_load_2d(smB + tap * kBTile, &tmW, w_bar, 0, tap * 64);
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                int r = t;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int y = r % p.H;
                const int n = r / p.H;
                for (int tap = 0; tap < 9; ++tap) {
                    t5::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                    t5::mbar_arrive_expect_tx(&full_bar[stage], kATile);
                    t_tma_load_4d(smA + stage * kATile, &tmX, &full_bar[stage], 0, tx * 128 + (tap % 3) - 1, y + (tap / 3) - 1, n);
                    if (++stage == kC3Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = t5::make_idesc_f16(128, 64, 0, 0, 0, 0);
        int stage = 0; uint32_t phase = 0; int it = 0;
        t5::mbar_wait_relaxed(w_bar, 0);
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
            t5::mbar_wait_relaxed(&tempty_bar[as], aphase ^ 1);
            t5::fence_after_thread_sync();
            const uint32_t tmem_d = tmem_base + as * 64;
            for (int tap = 0; tap < 9; ++tap) {
                t5::mbar_wait_relaxed(&full_bar[stage], phase);
                t5::fence_after_thread_sync();
                {
                    // whole converged warp, one elected lane per instruction (see gemm.cu)
                    const uint64_t da0 = t5::make_smem_desc(t5::smem_u32(smA + stage * kATile), 16, 512, 4);
                    const uint64_t db0 = t5::make_smem_desc(t5::smem_u32(smB + tap * kBTile), 16, 512, 4);
#pragma unroll
                    for (int k = 0; k < 2; ++k) t5::mma_f16_ss_warp(tmem_d, da0 + 2 * k, db0 + 2 * k, idesc, (tap | k) != 0);
                    t5::mma_commit_warp(&empty_bar[stage]);
                    if (tap == 8) t5::mma_commit_warp(&tfull_bar[as]);
                }
                if (++stage == kC3Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            int r = t;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int y = r % p.H;
            const int n = r / p.H;
            const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
            t5::mbar_wait(&tfull_bar[as], aphase);
            t5::fence_after_thread_sync();
            const int x = tx * 128 + row;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 64;
            act_t* o = p.out + ((static_cast<long long>(n) * p.H + y) * p.W + x) * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t rr[16];
                t5::tmem_ld16(taddr + c0, rr);
                t5::tmem_ld_wait();
                if (x < p.W) {
                    float v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = fmaxf(fmaf(__uint_as_float(rr[k]), __ldg(p.scale + c0 + k), __ldg(p.shift + c0 + k)), 0.f);
                    stg_stream(o + c0, pack8h(v));
                    stg_stream(o + c0 + 8, pack8h(v + 8));
                }
            }
            t5::fence_before_thread_sync();
            t5::mbar_arrive(&tempty_bar[as]);
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) { t5::fence_after_thread_sync(); t5::tmem_dealloc(tmem_base, 128); }
}

PFN_cuTensorMapEncodeTiled_v12000 t_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
    }
    return fn;
}

// ============================================================================================ the network
struct TVar { std::string name; int shape[4] = {0, 0, 0, 0}; int ndim = 0; long long off = 0, count = 0; };
enum TKind { T_STEM, T_CONV3, T_DW, T_PW, T_SUBSAMPLE, T_RESIZE, T_IMGPOOL, T_LOGITS };
struct TOp {
    int kind = 0;
    std::string name;
    int in = -1, out = -1, res = -1;            // tensor ids (res: residual added in the PW epilogue)
    int cin = 0, cout = 0, stride = 1, dil = 1;
    int pre_relu = 0, act = 0;                  // act: 0 none, 1 ReLU
    long long w_off = -1, bias_off = -1;        // into params
    long long gamma = -1, beta = -1, mean = -1, var = -1; float eps = 1e-3f;
    long long fold_off = -1;                    // scale / shift in the fold pool (2 * cout floats)
    long long w16_off = -1, w16lo_off = -1;     // fp16 operand copies
    int k_rows0 = 0, k_rows = 0;                // PW: HWIO row window (concat_projection skips the image-pooling rows)
    int out_ch0 = 0;                            // PW / RESIZE: first channel of the output tensor this op writes (concat slices)
};
struct TTensor { int h = 0, w = 0, c = 0; act_t* p = nullptr; };

struct Teacher {
    int device = 0, num_sms = kNumSMs, num_classes = 19;
    int N = 0, H = 0, W = 0;                    // planned input size (0 = no plan yet)
    cudaStream_t stream = nullptr;
    std::vector<TVar> vars; std::unordered_map<std::string, int> var_index;
    std::vector<TOp> ops;
    std::vector<int> t_c, t_scale;              // per tensor: channels, cumulative stride (1 = input resolution)
    long long n_params = 0, n_fold = 0, n_w16 = 0;
    float* params = nullptr; float* fold = nullptr; uint16_t* w16 = nullptr; float* ones = nullptr; float* zeros = nullptr;
    bool dirty = true;
    // plan
    std::vector<TTensor> tens; std::vector<void*> allocs;
    std::vector<GemmPlan> gemm; CUtensorMap tmX3, tmW3;
    uint8_t* frames = nullptr; float* logits = nullptr; int32_t* pred = nullptr; HeadStats* head_st = nullptr;
    float *pooled = nullptr, *ip_z = nullptr, *ip_act = nullptr, *bias_img = nullptr; double* small_ws = nullptr;
    WeightCast* cast_table = nullptr; int cast_n = 0, cast_max = 0;
};

struct Builder {
    Teacher* t;
    int new_tensor(int c, int scale) { t->t_c.push_back(c); t->t_scale.push_back(scale); return static_cast<int>(t->t_c.size()) - 1; }
    long long add_var(const std::string& name, std::initializer_list<int> shape) {
        TVar v; v.name = name; v.ndim = static_cast<int>(shape.size());
        long long cnt = 1; int i = 0;
        for (int s : shape) { v.shape[i++] = s; cnt *= s; }
        v.count = cnt; v.off = t->n_params; t->n_params += cnt;
        t->var_index[name] = static_cast<int>(t->vars.size());
        t->vars.push_back(v);
        return v.off;
    }
    void add_bn(TOp& o, const std::string& scope, float eps) {
        o.gamma = add_var(scope + "/BatchNorm/gamma:0", {o.cout});
        o.beta = add_var(scope + "/BatchNorm/beta:0", {o.cout});
        o.mean = add_var(scope + "/BatchNorm/moving_mean:0", {o.cout});
        o.var = add_var(scope + "/BatchNorm/moving_variance:0", {o.cout});
        o.eps = eps; o.fold_off = t->n_fold; t->n_fold += 2LL * o.cout;
    }
    // 1x1 conv + BN (+ ReLU) (+ residual); k_rows0: first HWIO row used (concat_projection); out_t / out_ch0: concat slice
    int pw(const std::string& scope, int in, int cout, int act, int res, float eps, int out_t = -1, int out_ch0 = 0, int k_rows0 = 0,
           int k_total = -1) {
        TOp o; o.kind = T_PW; o.name = scope; o.in = in; o.cin = t->t_c[in]; o.cout = cout; o.act = act; o.res = res;
        const int ktot = k_total > 0 ? k_total : o.cin;
        o.k_rows0 = k_rows0; o.k_rows = o.cin;
        o.w_off = add_var(scope + "/weights:0", {1, 1, ktot, cout});
        add_bn(o, scope, eps);
        o.out = out_t >= 0 ? out_t : new_tensor(cout, t->t_scale[in]); o.out_ch0 = out_ch0;
        o.w16_off = t->n_w16; t->n_w16 += (static_cast<long long>(cout) * o.cin + 63) & ~63LL;
        if (cout <= kSplitWeightMaxCout) { o.w16lo_off = t->n_w16; t->n_w16 += (static_cast<long long>(cout) * o.cin + 63) & ~63LL; }
        t->ops.push_back(o);
        return o.out;
    }
    int dw(const std::string& scope, int in, int stride, int dil, int pre_relu, int act, float eps) {
        TOp o; o.kind = T_DW; o.name = scope; o.in = in; o.cin = o.cout = t->t_c[in]; o.stride = stride; o.dil = dil; o.pre_relu = pre_relu; o.act = act;
        o.w_off = add_var(scope + "/depthwise_weights:0", {3, 3, o.cin, 1});
        add_bn(o, scope, eps);
        o.out = new_tensor(o.cout, t->t_scale[in] * stride);
        t->ops.push_back(o);
        return o.out;
    }
    // Xception separable conv: [ReLU on load] depthwise + BN [+ ReLU] -> pointwise + BN [+ ReLU]
    int sep(const std::string& scope, int in, int cout, int stride, int dil, bool act_inside, int res = -1) {
        const int d = dw(scope + "_depthwise", in, stride, dil, act_inside ? 0 : 1, act_inside ? 1 : 0, 1e-3f);
        return pw(scope + "_pointwise", d, cout, act_inside ? 1 : 0, res, 1e-3f);
    }
    int subsample(int in) {
        TOp o; o.kind = T_SUBSAMPLE; o.name = "subsample"; o.in = in; o.cin = o.cout = t->t_c[in]; o.stride = 2;
        o.out = new_tensor(o.cout, t->t_scale[in] * 2);
        t->ops.push_back(o);
        return o.out;
    }
    // one xception_module: three separable convs (the last with the stride), skip 'conv' | 'sum' | 'none'
    int module(const std::string& scope, int in, const int depth[3], int skip /*0 conv, 1 sum, 2 none*/, int stride, int dil, bool act_inside,
               int* low_level = nullptr) {
        int shortcut = -1;
        if (skip == 0) {
            const int src = stride == 2 ? subsample(in) : in;
            shortcut = pw(scope + "/shortcut", src, depth[2], 0, -1, 1e-3f);
        } else if (skip == 1) shortcut = in;
        int x = sep(scope + "/separable_conv1", in, depth[0], 1, dil, act_inside);
        x = sep(scope + "/separable_conv2", x, depth[1], 1, dil, act_inside);
        if (low_level) *low_level = x;
        x = sep(scope + "/separable_conv3", x, depth[2], stride, dil, act_inside, shortcut);
        return x;
    }
};

int build_topology(Teacher* t) {
    Builder b{t};
    const int in = b.new_tensor(3, 1);
    (void)in;
    // entry flow
    TOp s; s.kind = T_STEM; s.name = "xception_65/entry_flow/conv1_1"; s.in = 0; s.cin = 3; s.cout = 32; s.stride = 2; s.act = 1;
    s.w_off = b.add_var(s.name + "/weights:0", {3, 3, 3, 32}); b.add_bn(s, s.name, 1e-3f);
    s.out = b.new_tensor(32, 2); t->ops.push_back(s);
    TOp c; c.kind = T_CONV3; c.name = "xception_65/entry_flow/conv1_2"; c.in = s.out; c.cin = 32; c.cout = 64; c.act = 1;
    c.w_off = b.add_var(c.name + "/weights:0", {3, 3, 32, 64}); b.add_bn(c, c.name, 1e-3f);
    c.out = b.new_tensor(64, 2); c.w16_off = t->n_w16; t->n_w16 += 9 * 64 * 32; t->ops.push_back(c);
    int x = c.out, low = -1;
    const int d1[3] = {128, 128, 128}, d2[3] = {256, 256, 256}, d3[3] = {728, 728, 728}, e1[3] = {728, 1024, 1024}, e2[3] = {1536, 1536, 2048};
    x = b.module("xception_65/entry_flow/block1/unit_1/xception_module", x, d1, 0, 2, 1, false);
    x = b.module("xception_65/entry_flow/block2/unit_1/xception_module", x, d2, 0, 2, 1, false, &low);
    x = b.module("xception_65/entry_flow/block3/unit_1/xception_module", x, d3, 0, 2, 1, false);
    for (int u = 1; u <= 16; ++u)
        x = b.module("xception_65/middle_flow/block1/unit_" + std::to_string(u) + "/xception_module", x, d3, 1, 1, 1, false);
    // output stride 16 is reached: the stride of exit_flow/block1 becomes the dilation of what follows
    x = b.module("xception_65/exit_flow/block1/unit_1/xception_module", x, e1, 0, 1, 1, false);
    x = b.module("xception_65/exit_flow/block2/unit_1/xception_module", x, e2, 2, 1, 2, true);
    const int feat = x;                                                      // [N, H/16, W/16, 2048]
    // ASPP: concat order [image_pooling, aspp0, aspp1, aspp2, aspp3]; the (broadcast) pooling branch becomes a per-image bias
    const int cat = b.new_tensor(1024, t->t_scale[feat]);
    TOp ip; ip.kind = T_IMGPOOL; ip.name = "image_pooling"; ip.in = feat; ip.cin = 2048; ip.cout = 256; ip.act = 1;
    ip.w_off = b.add_var("image_pooling/weights:0", {1, 1, 2048, 256}); b.add_bn(ip, "image_pooling", 1e-5f);
    t->ops.push_back(ip);
    b.pw("aspp0", feat, 256, 1, -1, 1e-5f, cat, 0);
    const int rates[3] = {6, 12, 18};
    for (int i = 0; i < 3; ++i) {
        const std::string sc = "aspp" + std::to_string(i + 1);
        const int d = b.dw(sc + "_depthwise", feat, 1, rates[i], 0, 1, 1e-5f);
        b.pw(sc + "_pointwise", d, 256, 1, -1, 1e-5f, cat, 256 * (i + 1));
    }
    const int cp = b.pw("concat_projection", cat, 256, 1, -1, 1e-5f, -1, 0, 256, 1280);
    // decoder (output stride 4): low-level features of entry_flow/block2 (separable_conv2_pointwise) -> 48 channels
    const int dcat = b.new_tensor(304, t->t_scale[low]);
    TOp rz; rz.kind = T_RESIZE; rz.name = "decoder/upsample"; rz.in = cp; rz.cin = rz.cout = 256; rz.out = dcat; rz.out_ch0 = 0;
    t->ops.push_back(rz);
    b.pw("decoder/feature_projection0", low, 48, 1, -1, 1e-5f, dcat, 256);
    int y = b.dw("decoder/decoder_conv0_depthwise", dcat, 1, 1, 0, 1, 1e-5f);
    y = b.pw("decoder/decoder_conv0_pointwise", y, 256, 1, -1, 1e-5f);
    y = b.dw("decoder/decoder_conv1_depthwise", y, 1, 1, 0, 1, 1e-5f);
    y = b.pw("decoder/decoder_conv1_pointwise", y, 256, 1, -1, 1e-5f);
    TOp lg; lg.kind = T_LOGITS; lg.name = "logits/semantic"; lg.in = y; lg.cin = 256; lg.cout = t->num_classes; lg.k_rows = 256;
    lg.w_off = b.add_var("logits/semantic/weights:0", {1, 1, 256, t->num_classes});
    lg.bias_off = b.add_var("logits/semantic/biases:0", {t->num_classes});
    lg.w16_off = t->n_w16; t->n_w16 += (static_cast<long long>(lg.cout) * 256 + 63) & ~63LL;
    lg.w16lo_off = t->n_w16; t->n_w16 += (static_cast<long long>(lg.cout) * 256 + 63) & ~63LL;
    t->ops.push_back(lg);
    return 0;
}

int out_size(int in, int stride) { return (in + stride - 1) / stride; }

void free_plan(Teacher* t) {
    for (void* p : t->allocs) cudaFree(p);
    t->allocs.clear(); t->tens.clear(); t->gemm.clear();
    t->frames = nullptr; t->logits = nullptr; t->pred = nullptr;
    t->N = t->H = t->W = 0;
}

template <typename T>
int t_alloc(Teacher* t, T** p, size_t count) {
    void* q = nullptr;
    AMS_CUDA_CHECK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    *p = static_cast<T*>(q);
    t->allocs.push_back(q);
    return 0;
}

int build_plan(Teacher* t, int N, int H, int W) {
    free_plan(t);
    const size_t nt = t->t_c.size();
    t->tens.resize(nt);
    for (size_t i = 0; i < nt; ++i) {
        TTensor& x = t->tens[i];
        int h = H, w = W;
        for (int s = t->t_scale[i]; s > 1; s >>= 1) { h = out_size(h, 2); w = out_size(w, 2); }
        x.h = h; x.w = w; x.c = t->t_c[i];
        if (i == 0) continue;
        if (t_alloc(t, &x.p, static_cast<size_t>(N) * h * w * x.c)) return -1;
    }
    if (t_alloc(t, &t->frames, static_cast<size_t>(N) * H * W * 3)) return -1;
    if (t_alloc(t, &t->pred, static_cast<size_t>(N) * H * W)) return -1;
    t->gemm.resize(t->ops.size());
    for (size_t i = 0; i < t->ops.size(); ++i) {
        const TOp& o = t->ops[i];
        if (o.kind == T_PW || o.kind == T_LOGITS) {
            const TTensor& a = t->tens[o.in];
            GemmDesc g;
            g.A = a.p; g.lda = a.c; g.M = N * a.h * a.w; g.K = o.k_rows; g.N = o.cout;
            g.B = t->w16 + o.w16_off; g.ldb = o.k_rows;
            if (o.w16lo_off >= 0) g.B_lo = t->w16 + o.w16lo_off;
            if (o.kind == T_LOGITS) {
                if (t_alloc(t, &t->logits, static_cast<size_t>(g.M) * 32)) return -1;
                AMS_CUDA_CHECK(cudaMemset(t->logits, 0, static_cast<size_t>(g.M) * 32 * sizeof(float)));
                g.out = t->logits; g.ldc = 32; g.out_fp32 = 1; g.shift = t->params + o.bias_off;
            } else {
                const TTensor& y = t->tens[o.out];
                g.out = y.p + o.out_ch0; g.ldc = y.c; g.scale = t->fold + o.fold_off; g.shift = t->fold + o.fold_off + o.cout; g.act = o.act;
                if (o.res >= 0) { g.residual = t->tens[o.res].p; g.ldr = t->tens[o.res].c; }
                if (o.name == "concat_projection") { g.rowbias = t->bias_img; g.rows_per_image = a.h * a.w; }
            }
            if (o.name == "concat_projection" && !t->bias_img) {
                if (t_alloc(t, &t->pooled, static_cast<size_t>(N) * 2048) || t_alloc(t, &t->ip_z, static_cast<size_t>(N) * 256) ||
                    t_alloc(t, &t->ip_act, static_cast<size_t>(N) * 256) || t_alloc(t, &t->bias_img, static_cast<size_t>(N) * 256) ||
                    t_alloc(t, &t->small_ws, colsum_workspace_doubles(N, 2048))) return -1;
                g.rowbias = t->bias_img;
            }
            if (gemm_plan(g, t->num_sms, &t->gemm[i])) return -1;
        } else if (o.kind == T_CONV3) {
            auto fn = t_encode_fn();
            AMS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available");
            const TTensor& a = t->tens[o.in];
            cuuint64_t dims[4] = {32, static_cast<cuuint64_t>(a.w), static_cast<cuuint64_t>(a.h), static_cast<cuuint64_t>(N)};
            cuuint64_t strides[3] = {64, static_cast<cuuint64_t>(a.w) * 64, static_cast<cuuint64_t>(a.h) * a.w * 64};
            cuuint32_t box[4] = {32, 128, 1, 1}, estr[4] = {1, 1, 1, 1};
            CUresult r = fn(&t->tmX3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            AMS_REQUIRE(r == CUDA_SUCCESS, "tensor map (conv1_2 input) failed");
            cuuint64_t wd[2] = {32, 9 * 64}; cuuint64_t ws[1] = {64}; cuuint32_t wb[2] = {32, 64};
            r = fn(&t->tmW3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t->w16 + o.w16_off, wd, ws, wb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            AMS_REQUIRE(r == CUDA_SUCCESS, "tensor map (conv1_2 weights) failed");
        }
    }
    // image-pooling scratch may have been allocated after its first use in a GEMM descriptor: rebuild that plan
    for (size_t i = 0; i < t->ops.size(); ++i)
        if (t->ops[i].name == "concat_projection") { GemmDesc g = t->gemm[i].d; g.rowbias = t->bias_img; if (gemm_plan(g, t->num_sms, &t->gemm[i])) return -1; }
    t->N = N; t->H = H; t->W = W;
    return 0;
}

int prepare_weights(Teacher* t) {
    if (!t->dirty) return 0;
    cudaStream_t s = t->stream;
    for (const TOp& o : t->ops)
        if (o.fold_off >= 0 && o.kind != T_IMGPOOL)
            if (bn_fold_frozen(t->params + o.gamma, t->params + o.beta, t->params + o.mean, t->params + o.var, o.eps, t->fold + o.fold_off,
                               t->fold + o.fold_off + o.cout, o.cout, s)) return -1;
    if (cast_weights(t->cast_table, t->cast_n, t->cast_max, s)) return -1;
    for (const TOp& o : t->ops)
        if (o.kind == T_CONV3) {
            t_cast_conv3_kernel<<<ceil_div(9 * 32 * 64, 256), 256, 0, s>>>(t->params + o.w_off, reinterpret_cast<act_t*>(t->w16 + o.w16_off), 32, 64);
            AMS_CUDA_CHECK(cudaGetLastError());
        }
    t->dirty = false;
    return 0;
}

int forward(Teacher* t) {
    cudaStream_t s = t->stream;
    const int N = t->N;
    if (prepare_weights(t)) return -1;
    for (size_t i = 0; i < t->ops.size(); ++i) {
        const TOp& o = t->ops[i];
        const TTensor& a = t->tens[o.in];
        switch (o.kind) {
        case T_STEM: {
            const TTensor& y = t->tens[o.out];
            const int pt = 1, pl = 1;                   // conv2d_same: explicit padding (k_eff - 1) / 2 in front, whatever the size
            if (stem_conv_fwd(t->frames, 1, N, t->H, t->W, t->H, t->W, y.h, y.w, pt, pl, 0.f, 0.007843137718737125f, 1.0f, t->params + o.w_off,
                              t->fold + o.fold_off, t->fold + o.fold_off + 32, y.p, s, 1)) return -1;
            break;
        }
        case T_CONV3: {
            const TTensor& y = t->tens[o.out];
            Conv3Params p{N, a.h, a.w, ceil_div(a.w, 128), N * a.h * ceil_div(a.w, 128), t->fold + o.fold_off, t->fold + o.fold_off + 64, y.p};
            const size_t smem = 1024 + kC3Stages * 128 * 64 + 9 * 64 * 64 + (2 * kC3Stages + 5) * 8 + 16;
            static bool attr = false;
            if (!attr) { AMS_CUDA_CHECK(cudaFuncSetAttribute(t_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
            AMS_LAUNCH((t_conv3x3_kernel), std::min(p.num_tiles, t->num_sms), kC3Threads, smem, s, t->tmX3, t->tmW3, p);
            break;
        }
        case T_DW: {
            const TTensor& y = t->tens[o.out];
            // separable_conv2d_same: stride 1 -> 'SAME' (dil on each side); stride 2 -> explicit padding (k_eff - 1) / 2 = dil in
            // front followed by a 'VALID' conv (== 'SAME' at the odd sizes the teacher sees: 1025 x 2049 and its halvings)
            const int pt = o.dil, pl = o.dil;
            Conv2dGeom cg{N, a.h, a.w, a.c, y.h, y.w, o.stride, o.dil, pt, pl};
            if (dw_tiled_supported(cg)) {
                // the student's shared-memory tiled depthwise kernel (dw_tiled.cu): ReLU-on-load = its producer-activation hook
                // with an identity affine, folded BN (+ ReLU) on the way out
                if (dw_conv_fwd_tiled(a.p, t->params + o.w_off, cg, o.pre_relu ? t->ones : nullptr, o.pre_relu ? t->zeros : nullptr, 1,
                                      t->fold + o.fold_off, t->fold + o.fold_off + o.cout, o.act ? 1 : 0, y.p, nullptr, nullptr, s)) return -1;
                break;
            }
            const long long total = static_cast<long long>(N) * y.h * y.w * (a.c / 8);
            AMS_LAUNCH((t_dw_kernel), static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, s, a.p, t->params + o.w_off, t->fold + o.fold_off,
                       t->fold + o.fold_off + o.cout, y.p, N, a.h, a.w, a.c, y.h, y.w, o.stride, o.dil, pt, pl, o.pre_relu, o.act);
            break;
        }
        case T_SUBSAMPLE: {
            const TTensor& y = t->tens[o.out];
            const long long total = static_cast<long long>(N) * y.h * y.w * (a.c / 8);
            AMS_LAUNCH((t_subsample2_kernel), static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, s, a.p, y.p, N, a.h, a.w, a.c, y.h, y.w);
            break;
        }
        case T_RESIZE: {
            const TTensor& y = t->tens[o.out];
            const float sy = y.h > 1 ? static_cast<float>(a.h - 1) / static_cast<float>(y.h - 1) : 0.f;
            const float sx = y.w > 1 ? static_cast<float>(a.w - 1) / static_cast<float>(y.w - 1) : 0.f;
            const long long total = static_cast<long long>(N) * y.h * y.w * (a.c / 8);
            AMS_LAUNCH((t_resize_kernel), static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, s, a.p, y.p + o.out_ch0, N, a.h, a.w, a.c, y.h, y.w, y.c, sy, sx);
            break;
        }
        case T_IMGPOOL: {
            // Mean -> 1x1 conv -> BN -> ReLU -> (broadcast) -> rows 0..255 of concat_projection: a per-image fp32 bias
            const TOp* cp = nullptr;
            for (const TOp& q : t->ops) if (q.name == "concat_projection") cp = &q;
            ImgPoolFwd f;
            f.N = N; f.HW = a.h * a.w; f.Cin = o.cin; f.Cmid = o.cout; f.Cout = cp->cout; f.feat = a.p;
            f.w_pool = t->params + o.w_off; f.w_proj_top = t->params + cp->w_off;
            BnLayer bn; bn.C = o.cout; bn.M = N; bn.eps = o.eps; bn.one_minus_decay = 0.f; bn.gamma = t->params + o.gamma; bn.beta = t->params + o.beta;
            bn.moving_mean = t->params + o.mean; bn.moving_var = t->params + o.var;
            bn.mean = t->fold + o.fold_off; bn.rstd = t->fold + o.fold_off + o.cout; bn.scale = t->fold + o.fold_off; bn.shift = t->fold + o.fold_off + o.cout;
            f.bn = bn; f.frozen = 1; f.update_moving = 0;
            f.pooled = t->pooled; f.z = t->ip_z; f.act = t->ip_act; f.bias_img = t->bias_img; f.ws = t->small_ws;
            if (imgpool_forward(f, s)) return -1;
            break;
        }
        case T_PW:
        case T_LOGITS:
            if (gemm_launch(t->gemm[i], s)) return -1;
            break;
        }
    }
    return 0;
}

}  // namespace
}  // namespace ams

using namespace ams;

extern "C" {

struct ams_teacher;

ams_teacher* ams_teacher_create(int num_classes, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_last_error("libams_b200 needs a CUDA device (sm_100a); there is no CPU fallback"); return nullptr; }
    if (device < 0 || device >= ndev || num_classes < 1 || num_classes > 21) { set_last_error("bad device ordinal / class count"); return nullptr; }
    Teacher* t = new Teacher();
    t->device = device; t->num_classes = num_classes;
    auto fail = [&](const char* what) -> ams_teacher* { set_last_error(what); delete t; return nullptr; };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return fail("libams_b200 is built for sm_100a only");
    t->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (build_topology(t)) return fail("topology");
    bool ok = cudaMalloc(reinterpret_cast<void**>(&t->params), t->n_params * sizeof(float)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&t->fold), t->n_fold * sizeof(float)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&t->w16), t->n_w16 * sizeof(uint16_t)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&t->head_st), sizeof(HeadStats)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&t->ones), 2 * 2048 * sizeof(float)) == cudaSuccess;
    if (!ok) return fail("device allocation failed");
    {
        std::vector<float> oz(2 * 2048, 0.f);
        for (int i = 0; i < 2048; ++i) oz[i] = 1.f;
        cudaMemcpy(t->ones, oz.data(), oz.size() * sizeof(float), cudaMemcpyHostToDevice);
        t->zeros = t->ones + 2048;
    }
    cudaMemset(t->params, 0, t->n_params * sizeof(float));
    cudaMemset(t->w16, 0, t->n_w16 * sizeof(uint16_t));
    std::vector<WeightCast> table;
    for (const TOp& o : t->ops) {
        if (o.kind != T_PW && o.kind != T_LOGITS) continue;
        WeightCast c;
        c.w = t->params + o.w_off; c.w_fwd = reinterpret_cast<act_t*>(t->w16 + o.w16_off);
        c.w_lo = o.w16lo_off >= 0 ? reinterpret_cast<act_t*>(t->w16 + o.w16lo_off) : nullptr; c.w_bwd = nullptr;
        c.Cin = o.k_rows0 + o.k_rows; c.Cout = o.cout; c.ld_fwd = o.k_rows; c.ld_bwd = 0; c.row0 = o.k_rows0; c.rows = o.k_rows;
        table.push_back(c);
        t->cast_max = std::max(t->cast_max, o.k_rows * o.cout);
    }
    t->cast_n = static_cast<int>(table.size());
    if (cudaMalloc(reinterpret_cast<void**>(&t->cast_table), table.size() * sizeof(WeightCast)) != cudaSuccess) return fail("device allocation failed");
    cudaMemcpy(t->cast_table, table.data(), table.size() * sizeof(WeightCast), cudaMemcpyHostToDevice);
    return reinterpret_cast<ams_teacher*>(t);
}

void ams_teacher_destroy(ams_teacher* h) {
    if (!h) return;
    Teacher* t = reinterpret_cast<Teacher*>(h);
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    free_plan(t);
    void* ptrs[] = {t->params, t->fold, t->w16, t->head_st, t->cast_table, t->ones};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

int ams_teacher_num_tensors(const ams_teacher* h) { return h ? static_cast<int>(reinterpret_cast<const Teacher*>(h)->vars.size()) : -1; }

int ams_teacher_tensor_info(const ams_teacher* h, int index, char* name, int cap, int shape4[4], int* ndim) {
    const Teacher* t = reinterpret_cast<const Teacher*>(h);
    if (!t || index < 0 || index >= static_cast<int>(t->vars.size())) { set_last_error("bad tensor index"); return -1; }
    const TVar& v = t->vars[index];
    if (name && cap > 0) { std::strncpy(name, v.name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = v.shape[i];
    if (ndim) *ndim = v.ndim;
    return 0;
}

int ams_teacher_set_tensor(ams_teacher* h, const char* name, const float* host, long long count) {
    Teacher* t = reinterpret_cast<Teacher*>(h);
    if (!t) { set_last_error("null handle"); return -1; }
    cudaSetDevice(t->device);
    auto it = t->var_index.find(name ? name : "");
    AMS_REQUIRE(it != t->var_index.end(), std::string("KeyError: no variable named '") + (name ? name : "") + "'");
    const TVar& v = t->vars[it->second];
    AMS_REQUIRE(v.count == count, "element count does not match the variable's shape");
    AMS_CUDA_CHECK(cudaMemcpyAsync(t->params + v.off, host, count * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(t->stream));
    t->dirty = true;
    return 0;
}

/* frames: [n,h,w,3] u8 RGB (the caller has applied the reference's 1-px symmetric top/left pad); out_labels int32 [n,h,w];
 * out_logits (optional) fp32 [n, ceil(h/4), ceil(w/4), num_classes]: logits/semantic before the final upsample */
int ams_teacher_predict(ams_teacher* h, const uint8_t* frames, int n, int height, int width, int32_t* out_labels, float* out_logits) {
    Teacher* t = reinterpret_cast<Teacher*>(h);
    if (!t) { set_last_error("null handle"); return -1; }
    cudaSetDevice(t->device);
    AMS_REQUIRE(frames && n > 0 && height >= 33 && width >= 33, "frames must be non-empty and at least 33 x 33");
    if (t->N != n || t->H != height || t->W != width)
        if (build_plan(t, n, height, width)) return -1;
    cudaStream_t s = t->stream;
    AMS_CUDA_CHECK(cudaMemcpyAsync(t->frames, frames, static_cast<size_t>(n) * height * width * 3, cudaMemcpyHostToDevice, s));
    if (forward(t)) return -1;
    const TOp& lg = t->ops.back();
    const TTensor& a = t->tens[lg.in];
    HeadGeom g{};
    g.N = n; g.h = a.h; g.w = a.w; g.ldl = 32; g.H = height; g.W = width; g.class_count = t->num_classes; g.normalize = 1;
    for (int i = 0; i < kMaxClasses; ++i) g.cls_idx[i] = i < t->num_classes ? i : 0;
    for (int i = 0; i < 256; ++i) g.label_lut[i] = -1;
    if (head_reset(t->head_st, s)) return -1;
    if (head_infer(t->logits, g, nullptr, t->pred, t->head_st, s)) return -1;
    if (out_labels) AMS_CUDA_CHECK(cudaMemcpyAsync(out_labels, t->pred, static_cast<size_t>(n) * height * width * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (out_logits)
        AMS_CUDA_CHECK(cudaMemcpy2DAsync(out_logits, t->num_classes * sizeof(float), t->logits, 32 * sizeof(float), t->num_classes * sizeof(float),
                                         static_cast<size_t>(n) * a.h * a.w, cudaMemcpyDeviceToHost, s));
    AMS_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

/* host-only layout queries (no device needed): the variable table of the restated teacher graph */
static Teacher* teacher_layout(int num_classes) {
    static std::unordered_map<int, Teacher*> cache;
    auto it = cache.find(num_classes);
    if (it != cache.end()) return it->second;
    if (num_classes < 1 || num_classes > 21) { set_last_error("bad class count"); return nullptr; }
    Teacher* t = new Teacher();
    t->num_classes = num_classes;
    if (build_topology(t)) { delete t; return nullptr; }
    cache[num_classes] = t;
    return t;
}
int ams_teacher_layout_num_tensors(int num_classes) {
    Teacher* t = teacher_layout(num_classes);
    return t ? static_cast<int>(t->vars.size()) : -1;
}
int ams_teacher_layout_tensor_info(int num_classes, int index, char* name, int cap, int shape4[4], int* ndim) {
    Teacher* t = teacher_layout(num_classes);
    if (!t) return -1;
    return ams_teacher_tensor_info(reinterpret_cast<const ams_teacher*>(t), index, name, cap, shape4, ndim);
}

/* device time of `reps` forward passes at the planned size (CUDA events on the teacher's stream), ms per pass */
int ams_teacher_time_forward(ams_teacher* h, int reps, float* out_ms) {
    Teacher* t = reinterpret_cast<Teacher*>(h);
    if (!t || t->N == 0) { set_last_error("no plan: call ams_teacher_predict first"); return -1; }
    cudaSetDevice(t->device);
    cudaEvent_t e0, e1;
    AMS_CUDA_CHECK(cudaEventCreate(&e0)); AMS_CUDA_CHECK(cudaEventCreate(&e1));
    if (forward(t)) return -1;
    AMS_CUDA_CHECK(cudaEventRecord(e0, t->stream));
    for (int i = 0; i < reps; ++i) if (forward(t)) return -1;
    AMS_CUDA_CHECK(cudaEventRecord(e1, t->stream));
    AMS_CUDA_CHECK(cudaStreamSynchronize(t->stream));
    float ms = 0.f;
    AMS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (out_ms) *out_ms = ms / std::max(reps, 1);
    return 0;
}

}  // extern "C"
