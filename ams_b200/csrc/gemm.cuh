// Host-side interface of the tcgen05 GEMM kernels (gemm.cu).
#pragma once
#include "common.cuh"

namespace ams {

// One 1x1-convolution-shaped contraction  OUT[M,N] = epilogue( A[M,K] * B[N,K]^T )
//   A: row-major [M, lda] 16-bit (NHWC pixels x channels), K-major: fp16 activations (forward) or bf16 gradients (dgrad)
//   B: weights,     row-major [N, ldb], K-major, same element type as A (forward: fp16 [Cout][Cin]; dgrad: bf16 [Cin][Cout])
// tcgen05.mma kind::f16 needs ONE element format for both operands (mixed fp16 x bf16 descriptors raise an illegal
// instruction on sm_100a -- measured), so the forward GEMM runs fp16 x fp16 and the data-gradient GEMM bf16 x bf16 on a
// bf16 copy of the transposed weights; accumulation is fp32 in TMEM either way.
// epilogue, in this order (fp32):  v = acc (+ rowbias[m / rows_per_image][n]) ; v = v*scale[n]+shift[n] ;
//   v = act(v) ; v += residual[m][n] ; store bf16 or fp32.
struct GemmDesc {
    const void* A = nullptr; int lda = 0; int a_fp16 = 1;       // 1 = fp16, 0 = bf16
    const void* B = nullptr; int ldb = 0; int b_fp16 = 1;
    // optional second weight plane (same shape / stride / type as B): OUT = A*B^T + A*B_lo^T in the same fp32 accumulator.
    // The forward convs pass B = fp16(W) and B_lo = fp16(W - B): ~21 significant bits of the fp32 master weight reach
    // the tensor core for one more (tiny, L2-resident) B tile per k-block -- weight rounding was the dominant term of
    // the end-to-end logit error with plain fp16 weights.
    const void* B_lo = nullptr;
    int M = 0, N = 0, K = 0;
    void* out = nullptr; int ldc = 0; int out_fp32 = 0;
    int out_fp16 = 1;                  // 16-bit output (and residual) element type: 1 = fp16 (activations), 0 = bf16 (gradients)
    const float* scale = nullptr;      // [N] or null (=1)
    const float* shift = nullptr;      // [N] or null (=0)
    const float* rowbias = nullptr;    // [M / rows_per_image][N] or null
    int rows_per_image = 1;
    const void* residual = nullptr; int ldr = 0;
    int act = 0;                       // 0 none, 1 relu, 2 relu6
    // optional BatchNorm batch statistics of the STORED (16-bit-rounded) output, fused into the epilogue:
    // stats_partial[cta][0][N] = column sums, [cta][1][N] = column sums of squares (one slab per CTA of the grid)
    double* stats_partial = nullptr;
};

// A prepared launch: tensor maps encoded once, reused every step (buffers are static in the plan).
struct GemmPlan {
    CUtensorMap tmA, tmB, tmB2, tmC;           // tmB2: the low weight plane (== tmB when there is none)
    GemmDesc d;
    int v2 = 0, stage_bufs = 1, acc_stages = 2, block_k = 64, linear_out = 0, pitch = 0, cbuf_bytes = 0, b_resident = 0;
    int block_n = 0, n_tiles = 0, m_tiles = 0, k_blocks = 0, stages = 0, tmem_cols = 0;
    size_t smem_bytes = 0;
    int grid = 0;
};
int gemm_plan(const GemmDesc& d, int num_sms, GemmPlan* plan);
int gemm_launch(const GemmPlan& plan, cudaStream_t stream);

// Weight gradient  dW[Cin,Cout] = sum_m X[m,Cin] * dZ[m,Cout]   (both operands MN-major, split over m)
struct WgradDesc {
    const void* X = nullptr; int ldx = 0; int Cin = 0; int x_fp16 = 1;       // activations (fp16: converted to bf16 tile by tile inside the kernel)
    const void* dZ = nullptr; int ldz = 0; int Cout = 0; int z_fp16 = 0;     // gradients (bf16)
    long long M = 0;
    float* dW = nullptr; int lddw = 0;          // fp32 [Cin][lddw]
    float* workspace = nullptr; size_t workspace_floats = 0;   // split-K partials
};
struct WgradPlan {
    CUtensorMap tmX, tmZ;
    WgradDesc d;
    int ci_tiles = 0, co_tiles = 0, block_n = 0, boxes_b = 0, splits = 0, kb_per_split = 0, k_blocks = 0, stages = 0;
    int tmem_cols = 0;
    size_t smem_bytes = 0;
};
size_t wgrad_workspace_floats(int Cin, int Cout, long long M, int num_sms);
int wgrad_plan(const WgradDesc& d, int num_sms, WgradPlan* plan);
int wgrad_launch(const WgradPlan& plan, cudaStream_t stream);

}  // namespace ams
