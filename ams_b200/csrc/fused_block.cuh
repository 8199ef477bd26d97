// Host interface of the block-fused inverted-residual kernel (fused_block.cu): frozen inference, stride-1 blocks.
#pragma once
#include "common.cuh"

namespace ams {

struct FusedBlockDesc {
    int N = 0, H = 0, W = 0;                 // INPUT spatial size (stride 1: input == output; stride 2: output = ceil(in / 2))
    int Cin = 0, Cexp = 0, Cout = 0, stride = 1, dil = 1;
    int pad_top = -1, pad_left = -1;         // stride 2: the depthwise conv's top / left padding (-1: TensorFlow 'SAME')
    const void* x = nullptr;                 // [N,H,W,Cin] fp16 block input
    const void* We = nullptr; const void* We_lo = nullptr; int ld_we = 0;     // expand weights fp16 [Cexp][ld_we] (+ low plane or null)
    const void* Wp = nullptr; const void* Wp_lo = nullptr; int ld_wp = 0;     // project weights fp16 [Cout][ld_wp] (+ low plane or null)
    const float* params = nullptr;           // [13][cpad]: s1 t1 wd[9] s2 t2, filled by fused_block_fill_params()
    const float* s3 = nullptr; const float* t3 = nullptr;                      // folded BN of the project conv [Cout]
    const void* residual = nullptr;          // block input when the block has a skip connection, else null
    void* out = nullptr;                     // [N,Ho,Wo,Cout] fp16
    void* debug_timeline = nullptr;          // optional: 3 x 64 x 4 u64 device words, timeline of CTA 0 in ns (diagnostics)
};

struct FusedBlockPlan {
    CUtensorMap tmX, tmWe, tmWeLo, tmWp, tmWpLo;
    FusedBlockDesc d;
    alignas(16) unsigned char params[320];   // the kernel's parameter block (FusedParams, private to fused_block.cu)
    size_t smem_bytes = 0;
    int grid = 0;
};

bool fused_block_supported(const FusedBlockDesc& d);
size_t fused_block_param_floats(const FusedBlockDesc& d);
int fused_block_plan(const FusedBlockDesc& d, int num_sms, FusedBlockPlan* plan);
// (re)computes the padded per-channel parameter vectors from the folded BN vectors and the depthwise filter
int fused_block_fill_params(const FusedBlockDesc& d, const float* s1, const float* t1, const float* wd, const float* s2, const float* t2,
                            cudaStream_t s);
int fused_block_launch(const FusedBlockPlan& plan, cudaStream_t s);

}  // namespace ams
