// Head kernels: the bilinear x15.5 upsample (ResizeBilinear, align_corners=True) is never materialised.
//   head_infer          upsample + class-subset argmax (+ integer-atomic confusion matrix + softmax-CE loss)
//   head_label_confmat  label-vs-label confusion matrix (calc_cross_miou)
//   head_loss_backward  upsample + softmax-CE forward and its gradient w.r.t. the LOW-RES logits, computed
//                       as two separable, gather-style (deterministic) passes: rows, then columns.
// Arithmetic order follows the TF kernel exactly (lerp in x, then y; separate mul/add roundings), so that the
// argmax is bit-identical to the oracle's given identical low-res logits.
// Replaces: ResizeBilinear_1/2 of model.meta + utils/graph_utils.py:373-408 (gather, argmax, one_hot,
// tf.metrics.mean_iou, softmax_cross_entropy_with_logits, boolean_mask, reduce_mean) and their gradients.
#include "kernels.cuh"

namespace ams {
namespace {

constexpr int kMaxC = 21;     // selected classes held in registers (Cityscapes 19, VOC 21)

struct ResizeAxis { float scale; int in_size; };

__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& lo, int& hi, float& lerp) {
    const float s = __fmul_rn(static_cast<float>(dst), scale);
    const float f = floorf(s);
    lo = static_cast<int>(f);
    hi = min(static_cast<int>(ceilf(s)), in_size - 1);
    lerp = __fsub_rn(s, f);
}

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {      // a + (b - a) * t, no contraction
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

struct HeadConst {
    int N, h, w, ldl, H, W, cc, normalize;
    float sy, sx;
    int cls[kMaxC];
    signed char lut[256];            // teacher label id -> reduced class (-1 = ignored); kernel-parameter resident
    signed char ch2c[kMaxClasses];   // logits channel -> reduced class (-1 = not selected)
};

__device__ __forceinline__ void warp_agg_hist(int* s_hist, int bin, bool active) {
    // warp-aggregated shared-memory histogram increment (uniform regions make most lanes hit one bin)
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(act, bin);
    if ((__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(&s_hist[bin], __popc(peers));
}

// block = 256 threads = 256 consecutive x; each thread walks RY rows
constexpr int kRY = 16;
template <int CM>      // class slots held in registers: 8 / 16 / 21 (loops over unused slots would still cost issue slots)
__global__ void __launch_bounds__(256)
head_infer_kernel(const float* __restrict__ logits, const __grid_constant__ HeadConst g,
                  const uint8_t* __restrict__ labels, int32_t* __restrict__ pred, HeadStats* __restrict__ st) {
    pdl_entry();
    __shared__ int s_hist[kMaxC * kMaxC];
    __shared__ int s_lut[256];
    __shared__ double s_loss[8];
    __shared__ int s_valid[8];
    const bool want_stats = (labels != nullptr);
    if (want_stats) {
        for (int i = threadIdx.x; i < g.cc * g.cc; i += blockDim.x) s_hist[i] = 0;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = g.lut[i];
        __syncthreads();
    }
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * kRY;
    const int n = blockIdx.z;
    const bool x_ok = x < g.W;
    int xlo = 0, xhi = 0; float xl = 0.f;
    src_index(x_ok ? x : 0, g.sx, g.w, xlo, xhi, xl);
    const float* base = logits + static_cast<long long>(n) * g.h * g.w * g.ldl;
    float top[CM], bot[CM];
    int cur_lo = -1, cur_hi = -1;
    double loss_acc = 0.0; int valid_acc = 0;
    for (int y = y0; y < min(y0 + kRY, g.H); ++y) {
        int ylo, yhi; float yl;
        src_index(y, g.sy, g.h, ylo, yhi, yl);
        if (ylo != cur_lo || yhi != cur_hi) {
            cur_lo = ylo; cur_hi = yhi;
            const float* rt = base + static_cast<long long>(ylo) * g.w * g.ldl;
            const float* rbp = base + static_cast<long long>(yhi) * g.w * g.ldl;
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                if (c < g.cc) {
                    const int ch = g.cls[c];
                    top[c] = lerp_rn(__ldg(rt + xlo * g.ldl + ch), __ldg(rt + xhi * g.ldl + ch), xl);
                    bot[c] = lerp_rn(__ldg(rbp + xlo * g.ldl + ch), __ldg(rbp + xhi * g.ldl + ch), xl);
                }
            }
        }
        float best = 0.f; int arg = 0;
        float v[CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            if (c < g.cc) {
                v[c] = lerp_rn(top[c], bot[c], yl);
                if (c == 0 || v[c] > best) { best = v[c]; arg = c; }
            }
        }
        const long long o = (static_cast<long long>(n) * g.H + y) * g.W + x;
        if (x_ok && pred) pred[o] = arg;
        if (want_stats) {
            int lab = -1;
            if (x_ok) lab = s_lut[labels[o]];
            const bool valid = lab >= 0;
            warp_agg_hist(s_hist, valid ? lab * g.cc + arg : 0, valid);
            if (valid) {
                float sum = 0.f, picked = 0.f;
#pragma unroll
                for (int c = 0; c < CM; ++c) {
                    if (c < g.cc) {
                        sum += expf(v[c] - best);
                        if (c == lab) picked = v[c];
                    }
                }
                loss_acc += static_cast<double>(logf(sum) + best - picked);
                ++valid_acc;
            }
        }
    }
    if (want_stats) {
        loss_acc = warp_sum_d(loss_acc);
        valid_acc = __reduce_add_sync(0xffffffffu, valid_acc);
        if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss_acc; s_valid[threadIdx.x >> 5] = valid_acc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double l = 0.0; int vv = 0;
            for (int i = 0; i < 8; ++i) { l += s_loss[i]; vv += s_valid[i]; }
            if (vv) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&st->loss_fixed),
                          static_cast<unsigned long long>(__double2ll_rn(l * kLossFixedScale)));
                atomicAdd(reinterpret_cast<unsigned long long*>(&st->n_valid), static_cast<unsigned long long>(vv));
            }
        }
        for (int i = threadIdx.x; i < g.cc * g.cc; i += blockDim.x)
            if (s_hist[i]) atomicAdd(reinterpret_cast<unsigned long long*>(&st->confmat[(i / g.cc) * kMaxClasses + (i % g.cc)]),
                                     static_cast<unsigned long long>(s_hist[i]));
    }
}

__global__ void __launch_bounds__(256)
label_confmat_kernel(const uint8_t* __restrict__ before, const uint8_t* __restrict__ after, long long n,
                     const __grid_constant__ HeadConst g, HeadStats* __restrict__ st) {
    pdl_entry();
    const int cc = g.cc;
    __shared__ int s_hist[kMaxC * kMaxC];
    __shared__ int s_lut[256];
    for (int i = threadIdx.x; i < cc * cc; i += blockDim.x) s_hist[i] = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = g.lut[i];
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n_round = (n + 31) / 32 * 32;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int a = -1, b = -1;
        if (i < n) { b = s_lut[before[i]]; a = s_lut[after[i]]; }
        const bool valid = (a >= 0) && (b >= 0);
        warp_agg_hist(s_hist, valid ? b * cc + a : 0, valid);       // rows = labels_before, cols = labels_after
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cc * cc; i += blockDim.x)
        if (s_hist[i]) atomicAdd(reinterpret_cast<unsigned long long*>(&st->confmat[(i / cc) * kMaxClasses + (i % cc)]),
                                 static_cast<unsigned long long>(s_hist[i]));
}

// ------------------------------------------------------------------------------------------ training head
// pass 1: one block per output row (n, y).  g[x][c] = softmax - onehot at valid pixels (else 0);
//         rowbuf[n][y][ix][c] = sum_x wx(x, ix) * g[x][c];  row_loss / row_valid partials.
template <int CM>
__global__ void __launch_bounds__(256)
head_rows_kernel(const float* __restrict__ logits, const __grid_constant__ HeadConst g,
                 const uint8_t* __restrict__ labels, float* __restrict__ rowbuf, double* __restrict__ row_loss,
                 int* __restrict__ row_valid) {
    pdl_entry();
    extern __shared__ float s_g[];                 // [W][cc], then the per-x source taps: int xlo[W], float xl[W]
    int* s_xlo = reinterpret_cast<int*>(s_g + static_cast<size_t>(g.W) * g.cc);
    float* s_xl = reinterpret_cast<float*>(s_xlo + g.W);
    __shared__ int s_lut[256];
    __shared__ double s_loss[8];
    __shared__ int s_valid[8];
    const int y = blockIdx.x % g.H, n = blockIdx.x / g.H;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = g.lut[i];
    __syncthreads();
    int ylo, yhi; float yl;
    src_index(y, g.sy, g.h, ylo, yhi, yl);
    const float* base = logits + static_cast<long long>(n) * g.h * g.w * g.ldl;
    const float* rt = base + static_cast<long long>(ylo) * g.w * g.ldl;
    const float* rbp = base + static_cast<long long>(yhi) * g.w * g.ldl;
    double loss_acc = 0.0; int valid_acc = 0;
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int lab = s_lut[labels[(static_cast<long long>(n) * g.H + y) * g.W + x]];
        float* gx = s_g + x * g.cc;
        int xlo, xhi; float xl;
        src_index(x, g.sx, g.w, xlo, xhi, xl);
        s_xlo[x] = xlo | (xhi << 16);
        s_xl[x] = xl;
        if (lab < 0) {
            for (int c = 0; c < g.cc; ++c) gx[c] = 0.f;
            continue;
        }
        float v[CM];
        float best = -INFINITY;
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            if (c < g.cc) {
                const int ch = g.cls[c];
                const float t = lerp_rn(__ldg(rt + xlo * g.ldl + ch), __ldg(rt + xhi * g.ldl + ch), xl);
                const float b = lerp_rn(__ldg(rbp + xlo * g.ldl + ch), __ldg(rbp + xhi * g.ldl + ch), xl);
                v[c] = lerp_rn(t, b, yl);
                best = fmaxf(best, v[c]);
            }
        }
        float sum = 0.f, picked = 0.f;
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            if (c < g.cc) {
                v[c] = expf(v[c] - best);
                sum += v[c];
            }
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            if (c < g.cc) {
                const float p = v[c] * inv;
                if (c == lab) picked = p;
                gx[c] = p - (c == lab ? 1.f : 0.f);
            }
        }
        loss_acc += static_cast<double>(-logf(picked));
        ++valid_acc;
    }
    loss_acc = warp_sum_d(loss_acc);
    valid_acc = __reduce_add_sync(0xffffffffu, valid_acc);
    if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss_acc; s_valid[threadIdx.x >> 5] = valid_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0; int vv = 0;
        for (int i = 0; i < 8; ++i) { l += s_loss[i]; vv += s_valid[i]; }
        row_loss[blockIdx.x] = l;
        row_valid[blockIdx.x] = vv;
    }
    // x-direction transpose-interpolation (gather form, fixed order)
    const float inv_sx = 1.f / g.sx;
    for (int i = threadIdx.x; i < g.w * g.cc; i += blockDim.x) {
        const int ix = i / g.cc, c = i % g.cc;
        int xa = static_cast<int>(floorf((ix - 1) * inv_sx)) - 1, xb = static_cast<int>(ceilf((ix + 1) * inv_sx)) + 1;
        xa = max(xa, 0); xb = min(xb, g.W - 1);
        float acc = 0.f;
        for (int x = xa; x <= xb; ++x) {
            const int packed = s_xlo[x];
            const int xlo = packed & 0xffff, xhi = packed >> 16;
            const float xl = s_xl[x];
            float wgt = 0.f;
            if (xlo == ix) wgt += 1.f - xl;
            if (xhi == ix) wgt += xl;
            if (wgt != 0.f) acc = fmaf(wgt, s_g[x * g.cc + c], acc);
        }
        rowbuf[(static_cast<long long>(blockIdx.x) * g.w + ix) * g.cc + c] = acc;
    }
}

__global__ void head_finalize_kernel(const double* __restrict__ row_loss, const int* __restrict__ row_valid, int rows,
                                     HeadStats* st, float* loss_out) {
    pdl_entry();
    __shared__ double s_l[256];
    __shared__ long long s_v[256];
    double l = 0.0; long long v = 0;
    // fixed partition of the rows across the 256 threads, then a fixed-order tree: deterministic
    for (int i = threadIdx.x; i < rows; i += 256) { l += row_loss[i]; v += row_valid[i]; }
    s_l[threadIdx.x] = l; s_v[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_v[threadIdx.x] += s_v[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st->loss_sum = s_l[0];
        st->n_valid = s_v[0];
        st->terms[0] = static_cast<double>(s_v[0]);
        st->terms[1] = s_l[0];
        *loss_out = s_v[0] > 0 ? static_cast<float>(s_l[0] / static_cast<double>(s_v[0])) : __int_as_float(0x7fc00000);
    }
}

// pass 2: thread per (n, iy, ix, channel of the padded 32-wide logits row)
__global__ void __launch_bounds__(256)
head_cols_kernel(const float* __restrict__ rowbuf, const __grid_constant__ HeadConst g,
                 const HeadStats* __restrict__ st, float* __restrict__ dl_f32, bf16* __restrict__ dl_bf16) {
    pdl_entry();
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.N) * g.h * g.w * g.ldl;
    if (tid >= total) return;
    const int ch = static_cast<int>(tid % g.ldl);
    long long r = tid / g.ldl;
    const int ix = static_cast<int>(r % g.w); r /= g.w;
    const int iy = static_cast<int>(r % g.h);
    const int n = static_cast<int>(r / g.h);
    const int c = g.ch2c[ch];
    float acc = 0.f;
    if (c >= 0 && st->n_valid > 0) {
        const float inv_sy = 1.f / g.sy;
        int ya = static_cast<int>(floorf((iy - 1) * inv_sy)) - 1, yb = static_cast<int>(ceilf((iy + 1) * inv_sy)) + 1;
        ya = max(ya, 0); yb = min(yb, g.H - 1);
        for (int y = ya; y <= yb; ++y) {
            int ylo, yhi; float yl;
            src_index(y, g.sy, g.h, ylo, yhi, yl);
            float wgt = 0.f;
            if (ylo == iy) wgt += 1.f - yl;
            if (yhi == iy) wgt += yl;
            if (wgt != 0.f) acc = fmaf(wgt, rowbuf[((static_cast<long long>(n) * g.H + y) * g.w + ix) * g.cc + c], acc);
        }
        if (g.normalize) acc /= static_cast<float>(st->n_valid);
    }
    dl_f32[tid] = acc;
    dl_bf16[tid] = __float2bfloat16_rn(acc);
}

__global__ void head_reset_kernel(HeadStats* st) {
    pdl_entry();
    for (int i = threadIdx.x; i < kMaxClasses * kMaxClasses; i += blockDim.x) st->confmat[i] = 0;
    if (threadIdx.x == 0) { st->loss_sum = 0.0; st->n_valid = 0; st->loss_fixed = 0; st->terms[0] = 0.0; st->terms[1] = 0.0; }
}

__global__ void head_mean_loss_kernel(HeadStats* st) {
    pdl_entry();
    if (threadIdx.x == 0)
        st->dp_loss[0] = st->terms[0] > 0.0 ? static_cast<float>(st->terms[1] / st->terms[0]) : __int_as_float(0x7fc00000);
}

int make_const(const HeadGeom& g, HeadConst* c) {
    AMS_REQUIRE(g.class_count > 0 && g.class_count <= kMaxC, "class_count must be in [1,21]");
    AMS_REQUIRE(g.ldl <= kMaxClasses, "logits row too wide");
    c->normalize = g.normalize; c->N = g.N; c->h = g.h; c->w = g.w; c->ldl = g.ldl; c->H = g.H; c->W = g.W; c->cc = g.class_count;
    c->sy = g.H > 1 ? static_cast<float>(static_cast<double>(g.h - 1) / static_cast<double>(g.H - 1)) : 0.f;
    c->sx = g.W > 1 ? static_cast<float>(static_cast<double>(g.w - 1) / static_cast<double>(g.W - 1)) : 0.f;
    for (int i = 0; i < kMaxC; ++i) c->cls[i] = i < g.class_count ? g.cls_idx[i] : 0;
    for (int i = 0; i < 256; ++i) c->lut[i] = static_cast<signed char>(g.label_lut[i]);
    for (int i = 0; i < kMaxClasses; ++i) c->ch2c[i] = -1;
    for (int j = 0; j < g.class_count; ++j) {
        AMS_REQUIRE(g.cls_idx[j] >= 0 && g.cls_idx[j] < g.ldl, "class index outside the logits row");
        c->ch2c[g.cls_idx[j]] = static_cast<signed char>(j);
    }
    return 0;
}

}  // namespace

int head_reset(HeadStats* st, cudaStream_t s) {
    AMS_LAUNCH((head_reset_kernel), 1, 256, 0, s, st);
    return 0;
}

int head_mean_loss_from_terms(HeadStats* st, cudaStream_t s) {
    AMS_LAUNCH((head_mean_loss_kernel), 1, 32, 0, s, st);
    return 0;
}

int head_infer(const float* logits, const HeadGeom& g, const uint8_t* labels, int32_t* pred, HeadStats* st,
               cudaStream_t s) {
    HeadConst c;
    if (make_const(g, &c)) return -1;
    dim3 grid(ceil_div(g.W, 256), ceil_div(g.H, kRY), g.N);
    if (g.class_count <= 8) AMS_LAUNCH((head_infer_kernel<8>), grid, 256, 0, s, logits, c, labels, pred, st);
    else if (g.class_count <= 16) AMS_LAUNCH((head_infer_kernel<16>), grid, 256, 0, s, logits, c, labels, pred, st);
    else AMS_LAUNCH((head_infer_kernel<kMaxC>), grid, 256, 0, s, logits, c, labels, pred, st);
    return 0;
}

int head_label_confmat(const uint8_t* before, const uint8_t* after, long long n, const HeadGeom& g, HeadStats* st,
                       cudaStream_t s) {
    HeadGeom g2 = g;
    if (g2.ldl <= 0) g2.ldl = kMaxClasses;
    HeadConst c;
    if (make_const(g2, &c)) return -1;
    const int blocks = static_cast<int>(std::min<long long>(ceil_div_ll(n, 256 * 8), 4 * kNumSMs));
    AMS_LAUNCH((label_confmat_kernel), std::max(blocks, 1), 256, 0, s, before, after, n, c, st);
    return 0;
}

size_t head_rowbuf_floats(const HeadGeom& g) {
    // rowbuf [N,H,w,cc] floats, then row_loss [N*H] doubles and row_valid [N*H] ints
    const size_t rows = static_cast<size_t>(g.N) * g.H;
    return rows * g.w * g.class_count + rows * 2 + rows + 64;
}

int head_loss_backward(const float* logits, const HeadGeom& g, const uint8_t* labels, float* rowbuf, float* dl_f32,
                       bf16* dl_bf16, HeadStats* st, float* loss_out, cudaStream_t s) {
    HeadConst c;
    if (make_const(g, &c)) return -1;
    AMS_REQUIRE(g.ldl == 32, "training head expects 32-wide padded logits rows");
    const size_t rows = static_cast<size_t>(g.N) * g.H;
    size_t off = rows * g.w * g.class_count;
    off = (off + 1) & ~size_t(1);                                  // 8-byte align the doubles
    double* row_loss = reinterpret_cast<double*>(rowbuf + off);
    int* row_valid = reinterpret_cast<int*>(row_loss + rows);
    const size_t smem = static_cast<size_t>(g.W) * (g.class_count + 2) * sizeof(float);     // gradients + cached x taps
    AMS_REQUIRE(smem <= 200 * 1024 && g.w < 65536, "row too wide for the head kernel");
    static size_t smem_set = 0;
    if (smem > smem_set) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(head_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        AMS_CUDA_CHECK(cudaFuncSetAttribute(head_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        AMS_CUDA_CHECK(cudaFuncSetAttribute(head_rows_kernel<kMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        smem_set = smem;
    }
    if (g.class_count <= 8) AMS_LAUNCH((head_rows_kernel<8>), static_cast<int>(rows), 256, smem, s, logits, c, labels, rowbuf, row_loss, row_valid);
    else if (g.class_count <= 16) AMS_LAUNCH((head_rows_kernel<16>), static_cast<int>(rows), 256, smem, s, logits, c, labels, rowbuf, row_loss, row_valid);
    else AMS_LAUNCH((head_rows_kernel<kMaxC>), static_cast<int>(rows), 256, smem, s, logits, c, labels, rowbuf, row_loss, row_valid);
    AMS_LAUNCH((head_finalize_kernel), 1, 256, 0, s, row_loss, row_valid, static_cast<int>(rows), st, loss_out);
    const long long total = static_cast<long long>(g.N) * g.h * g.w * g.ldl;
    AMS_LAUNCH((head_cols_kernel), static_cast<int>(ceil_div_ll(total, 256)), 256, 0, s, rowbuf, c, st, dl_f32, dl_bf16);
    return 0;
}

}  // namespace ams
