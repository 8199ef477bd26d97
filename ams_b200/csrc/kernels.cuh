// Launch interfaces of the non-GEMM kernels (HBM-bound byte/elementwise/reduction work).
// All tensors NHWC; forward activations fp16 (act_t), activation gradients bf16, statistics / parameters / parameter
// gradients fp32, counters int64.
#pragma once
#include "common.cuh"

namespace ams {

typedef __nv_bfloat16 bf16;     // activation GRADIENTS (range matters)
typedef __half act_t;           // forward ACTIVATIONS and 1x1 weight operands (precision matters: DESIGN.md 3)

struct Conv2dGeom {          // one image-plane geometry; TF 'SAME' pads resolved on the host
    int N, H, W, C;          // input
    int Ho, Wo;              // output
    int stride, dil;
    int pad_top, pad_left;
};

// ---- stem: pad(127.5) + (x*2/255-1) + 3x3 s2 conv 3->32, fused (SURVEY K1+K2)
// in: u8 or f32 [N,H,W,3] un-padded frame; the graph's 1-px bottom/right mean-pixel pad is virtual.
// out fp16 [N,Ho,Wo,32]; scale/shift != null => y = relu6(acc*scale+shift) (frozen BN fold), else raw z.
int stem_conv_fwd(const void* in, int in_is_u8, int N, int H, int W, int Hp, int Wp, int Ho, int Wo, int pad_top,
                  int pad_left, float pad_value, float norm_scale, float norm_shift, const float* w /*[3,3,3,32]*/,
                  const float* scale, const float* shift, act_t* out, cudaStream_t s, int act = 2 /* with scale/shift: 2 ReLU6, 1 ReLU */);
// dW[3,3,3,32] = sum_pixels patch(in) (x) dz ; partials [chunks][864] then fixed-order reduce.
int stem_conv_bwd_filter(const void* in, int in_is_u8, int N, int H, int W, int Hp, int Wp, int Ho, int Wo,
                         int pad_top, int pad_left, float pad_value, float norm_scale, float norm_shift,
                         const bf16* dz, float* dw, float* workspace, size_t workspace_floats, cudaStream_t s);
size_t stem_bwd_workspace_floats(int N, int Ho, int Wo);

// ---- depthwise 3x3
int dw_conv_fwd(const act_t* in, const float* w /*[3,3,C]*/, const Conv2dGeom& g, const float* scale,
                const float* shift, int act, act_t* out, cudaStream_t s);
int dw_conv_bwd_data(const bf16* dz, const float* w, const Conv2dGeom& g, bf16* dx, cudaStream_t s);
int dw_conv_bwd_filter(const act_t* x, const bf16* dz, const Conv2dGeom& g, float* dw, float* workspace,
                       size_t workspace_floats, cudaStream_t s);
size_t dw_bwd_workspace_floats(const Conv2dGeom& g);
// shared-memory tiled forward (dw_tiled.cu): optional BN+act of the PRODUCER applied while staging the input
// (in_scale/in_shift), optional folded BN + act on the result (out_scale/out_shift), optional per-tile column sums
// stats[tile][2][C] (sum, sum of squares of the stored bf16 values) for bn_finalize_partials.
bool dw_tiled_supported(const Conv2dGeom& g);
long long dw_tiled_stats_rows(const Conv2dGeom& g);
int dw_conv_fwd_tiled(const act_t* in, const float* w, const Conv2dGeom& g, const float* in_scale, const float* in_shift,
                      int in_act, const float* out_scale, const float* out_shift, int out_act, act_t* out, double* stats,
                      int* stats_rows, cudaStream_t s);

// fused backward (dw_tiled.cu): BN backward of the depthwise layer applied while (g, z) is staged, data gradient +
// filter gradient from the same staged tile, the producer's BN+act recomputed from its raw output `zin`, activation
// mask applied to the stored gradient, column sums for the producer's BN backward.  coef = [3][C] (A, B, Cc) of
// bn_backward_reduce.  in_scale == null: `zin` is the (materialised) input activation and bn_partial is not written.
struct DwBwdFused {
    const bf16* g; const act_t* z; const float* scale; const float* shift; int act; const float* coef;
    const act_t* zin; const float* in_scale; const float* in_shift; int in_act;
    const float* w;
    bf16* gout; float* dw;                  // [N,H,W,C] gradient wrt the producer's BN output (masked); dW [9][C]
    float* dw_partial; size_t dw_partial_floats;    // >= dw_bwd_fused_rows * 9 * C
    double* bn_partial;                     // [rows][2][C]
};
long long dw_bwd_fused_rows(const Conv2dGeom& g);
// reduce_stream != null: the fixed-order reduction of the per-tile filter-gradient partials is NOT launched; the caller runs
// dw_conv_bwd_reduce(a, g, stream) itself (on a side stream: nothing in the backward chain reads the filter gradient)
int dw_conv_bwd_fused(const DwBwdFused& a, const Conv2dGeom& g, int* rows_out, cudaStream_t s, bool defer_reduce = false);
int dw_conv_bwd_reduce(const DwBwdFused& a, const Conv2dGeom& g, cudaStream_t s);

// ---- cross-GPU BatchNorm statistics (SURVEY 8e caveat 2: the reference normalises over the WHOLE batch in one process)
// One exchange per BatchNorm layer and direction, inside the finalize kernels: every rank pushes its per-channel fp64
// sum pair straight into every peer's receive buffer over NVLink (peer memory mapped through CUDA IPC), as 8-byte words
// that carry 32 bits of payload and the step's epoch as the arrival flag -- no fence, no separate flag, no host or NCCL
// call inside the step (the kernels are captured in the step's CUDA graph like all others).  Each rank then adds the
// world's pairs in rank order, so every rank computes bit-identical statistics.
constexpr int kSyncBnMaxWorld = 8;
struct SyncBn {                              // device-resident; kernels get a pointer to it (null = per-replica statistics)
    int world, rank;
    unsigned int epoch;                      // bumped once per training step (syncbn_begin_step); flags compare against it
    unsigned int error;                      // != 0: a wait ran into the timeout (a peer never arrived); later exchanges skip the wait
    long long words_per_src;                 // 8 * sum of BN channels: forward region, then backward region
    unsigned long long timeout_ns;
    unsigned long long* peer[kSyncBnMaxWorld];   // receive buffers [2 epoch parities][world sources][words_per_src]; peer[rank] is local
};
int syncbn_begin_step(SyncBn* sb, cudaStream_t s);

// ---- BatchNorm, training mode
struct BnLayer {            // device pointers into the parameter / state arenas, all [C] fp32
    SyncBn* sync = nullptr; // != null: batch statistics are summed over the ranks of the data-parallel job
    long long xoff_fwd = 0, xoff_bwd = 0;      // this layer's word offsets inside a source's region of the receive buffer
    int C;
    long long M;            // N*H*W reduction length
    float eps, one_minus_decay;
    const float* gamma; const float* beta;
    float* moving_mean; float* moving_var;     // updated in place when update_moving != 0
    float* mean; float* rstd;                  // saved batch statistics
    float* scale; float* shift;                // y = z*scale + shift
};
size_t bn_workspace_doubles(long long M, int C);
// batch statistics of z (fp16 [M,C]) -> scale/shift/mean/rstd (+ moving-average update)
int bn_forward_stats(const act_t* z, const BnLayer& L, int update_moving, double* workspace, cudaStream_t s);
// same, from per-chunk partial sums a producer kernel already wrote: partial[chunk][0][C] = sum, [chunk][1][C] = sum sq
int bn_finalize_partials(const double* partial, int chunks, const BnLayer& L, int update_moving, cudaStream_t s);
// y = act(z*scale+shift) (+ residual)
int bn_apply(const act_t* z, const float* scale, const float* shift, int act, const act_t* residual, act_t* y,
             long long M, int C, cudaStream_t s);
// frozen fold: scale = gamma*rsqrt(mv_var+eps), shift = beta - mv_mean*scale
int bn_fold_frozen(const float* gamma, const float* beta, const float* mv_mean, const float* mv_var, float eps,
                   float* scale, float* shift, int C, cudaStream_t s);
// backward: given dy (grad wrt act(BN(z))) and z; writes dz (dz_out may alias dy), d_gamma, d_beta.
// If dy2 != null the incoming gradient is dy + dy2 (two consumers).
int bn_backward(const bf16* dy, const bf16* dy2, const act_t* z, const BnLayer& L, int act, bf16* dz_out,
                float* d_gamma, float* d_beta, double* workspace, cudaStream_t s);

// the three pieces of bn_backward, for callers that fuse the apply pass (or the reduce pass) into another kernel:
// column sums of (dy masked, dy masked * z) -> coef [3][C] (dz = A*g + B*z + Cc), d_gamma, d_beta
int bn_backward_reduce(const bf16* dy, const act_t* z, const BnLayer& L, int act, float* coef, float* d_gamma, float* d_beta,
                       double* workspace, cudaStream_t s);
// same from per-tile partial sums [rows][2][C] a producer kernel wrote (dy already masked)
int bn_backward_finalize_partials(const double* partial, int rows, const BnLayer& L, float* coef, float* d_gamma,
                                  float* d_beta, cudaStream_t s);
// dz = A*dy + B*z + Cc for an already-masked dy (dz_out may alias dy)
int bn_backward_apply(const bf16* dy_masked, const act_t* z, const BnLayer& L, const float* coef, bf16* dz_out, cudaStream_t s);

// ---- ASPP image-pooling branch folded into a per-image bias
struct ImgPoolFwd {
    int N, HW, Cin /*320*/, Cmid /*256*/, Cout /*256*/;
    const act_t* feat;           // [N,HW,Cin]
    const float* w_pool;         // [Cin][Cmid]
    const float* w_proj_top;     // concat_projection rows 0..Cmid-1: [Cmid][Cout]
    BnLayer bn;                  // image_pooling BN (M = N)
    int frozen, update_moving;
    float* pooled;               // [N][Cin]   saved
    float* z;                    // [N][Cmid]  saved (pre-BN)
    float* act;                  // [N][Cmid]  saved (post relu)
    float* bias_img;             // [N][Cout]  -> rowbias of concat_projection
    double* ws;                  // colsum_workspace_doubles(N, max(Cin, Cout))
};
int imgpool_forward(const ImgPoolFwd& a, cudaStream_t s);
struct ImgPoolBwd {
    ImgPoolFwd f;
    const bf16* dz_proj;         // [N,HW,Cout] gradient wrt concat_projection pre-BN output
    float* dbias;                // [N][Cout] scratch
    float* d_w_proj_top;         // [Cmid][Cout]
    float* d_w_pool;             // [Cin][Cmid]
    float* d_gamma; float* d_beta;
    float* dfeat_rowbias;        // [N][Cin]  = d pooled / HW  -> rowbias of the aspp0 dgrad GEMM
};
int imgpool_backward(const ImgPoolBwd& a, cudaStream_t s);

// ---- head: bilinear upsample (align_corners) fused with argmax / confusion matrix / softmax-CE
constexpr int kMaxClasses = 32;
struct HeadGeom {
    int N, h, w, ldl;            // low-res logits fp32 [N,h,w,ldl]
    int H, W;                    // output size
    int class_count;             // selected classes
    int cls_idx[kMaxClasses];    // logits channel of reduced class j
    int label_lut[256];          // teacher label id -> reduced class, -1 = ignored pixel
    int normalize;               // training head: 1 = gradient of the MEAN loss (divide by n_valid), 0 = of the SUM
};
struct HeadStats {               // device, zeroed by head_reset()
    long long confmat[kMaxClasses * kMaxClasses];
    double loss_sum;             // training head: written by the fixed-order finalize
    long long n_valid;
    long long loss_fixed;        // inference metric: sum of per-CTA losses in 2^-36 fixed point (integer atomics: deterministic)
    double terms[2];             // training head: (n_valid, loss_sum) as doubles -- the pair a data-parallel job sums over ranks in place
    float dp_loss[2];            // [0] = terms[1] / terms[0] (global mean loss), written by head_mean_loss_from_terms()
};
constexpr double kLossFixedScale = 68719476736.0;   // 2^36
int head_reset(HeadStats* st, cudaStream_t s);
// st->dp_loss[0] = st->terms[1] / st->terms[0] (NaN when no valid pixel): the data-parallel loss after the terms were summed
int head_mean_loss_from_terms(HeadStats* st, cudaStream_t s);
// pred int32 [N,H,W] (may be null); labels u8 [N,H,W] or null; st accumulates confmat / loss / n_valid
int head_infer(const float* logits, const HeadGeom& g, const uint8_t* labels, int32_t* pred, HeadStats* st,
               cudaStream_t s);
// label-vs-label confusion matrix (calc_cross_miou): weight = both valid
int head_label_confmat(const uint8_t* before, const uint8_t* after, long long n, const HeadGeom& g, HeadStats* st,
                       cudaStream_t s);
// training: loss + d loss / d low-res logits, never materialising full-res logits.
// dlogits_f32 [N,h,w,ldl] fp32 (bias grad source), dlogits_bf16 [N*h*w, 32] bf16 (GEMM operand);
// rowbuf: [N,H,w,class_count] fp32 scratch.
size_t head_rowbuf_floats(const HeadGeom& g);
int head_loss_backward(const float* logits, const HeadGeom& g, const uint8_t* labels, float* rowbuf,
                       float* dlogits_f32, bf16* dlogits_bf16, HeadStats* st, float* loss_out /*device*/,
                       cudaStream_t s);

// ---- frame ingest (ingest_kernels.cu): cv2.resize INTER_LINEAR / INTER_NEAREST on uint8, optional BGR<->RGB swap
int resize_u8(const uint8_t* src, int n, int sh, int sw, int cn, uint8_t* dst, int dh, int dw, int nearest, int swap_rb,
              cudaStream_t s);

// ---- generic reductions
// colsum[g][c] = sum over rows of group g (rows_per_group consecutive rows) of x[row][c]   (deterministic)
// exactly one of x_f32 / x_bf16 / x_fp16 is non-null
int colsum_groups(const float* x_f32, const bf16* x_bf16, const act_t* x_fp16, int ld, long long rows_per_group, int groups, int C,
                  float scale, float* out, double* workspace, cudaStream_t s);
size_t colsum_workspace_doubles(int groups, int C);

// ---- optimizer / selection / delta
// scale_terms != null: grad_scale = 1 / scale_terms[0] read on the device (0 when scale_terms[0] <= 0)
int adam_masked(float* p, const float* g, float grad_scale, const double* scale_terms, float* m, float* v, const uint8_t* mask, long long n,
                float alpha, float one_minus_b1, float one_minus_b2, float eps, cudaStream_t s);
struct SelectScratch {           // device
    unsigned int hist[256];
    unsigned int prefix, rank, v_lo, count_le, next_gt, pad;
    float threshold;
    unsigned long long kept;
};
// mask[i] = |after-before| > thr  with thr = float32(a[lo]*(1-w_hi) + a[lo+1]*w_hi) over the n deltas
// (NumPy-1.19 percentile + value-based float32 demotion, see oracle/student_oracle.py);
// unselected coordinates are reverted: after[i] = before[i].
int select_coordinates(float* after, const float* before, float* delta_scratch, uint8_t* mask, long long n,
                       long long lo, double w_hi, SelectScratch* sc, cudaStream_t s);
struct VarSeg { long long offset; long long size; long long bit_byte_offset; };
// packbits per variable (MSB first) then fp16 values of masked coordinates in order
int pack_delta(const float* params, const uint8_t* mask, const VarSeg* segs_dev, int nseg, long long n,
               long long mask_bytes, uint8_t* out_bits, __half* out_vals, unsigned int* block_counts,
               int nblocks_alloc, unsigned long long* kept_out, cudaStream_t s);
int pack_delta_blocks(long long n);
// inverse (client side): bits -> byte mask (+ per-block offsets and the kept count), then fp16 values -> params[mask]
int unpack_delta_mask(const uint8_t* bits, const VarSeg* segs_dev, int nseg, long long n, long long mask_bytes, uint8_t* mask,
                      unsigned int* block_counts, int nblocks, unsigned long long* kept_out, cudaStream_t s);
int unpack_delta_values(float* params, const uint8_t* mask, long long n, const unsigned int* block_offsets, int nblocks,
                        const __half* vals, cudaStream_t s);

// fp32 HWIO 1x1 weights -> fp16 [Cout][Cin] (forward B operand, multiplies fp16 activations) and bf16 [Cin][ldb] (dgrad B
// operand, multiplies bf16 gradients)
// w_lo != null: the forward operand is split, w_fwd = fp16(w), w_lo = fp16(w - w_fwd)
struct WeightCast { const float* w; act_t* w_fwd; act_t* w_lo; bf16* w_bwd; int Cin, Cout, ld_fwd, ld_bwd; int row0, rows; };
int cast_weights(const WeightCast* table_dev, int n_layers, int max_elems, cudaStream_t s);

}  // namespace ams
