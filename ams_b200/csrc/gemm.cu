// tcgen05 / TMEM / TMA GEMM kernels for the 1x1 convolutions of the AMS student (sm_100a only).
//
//  gemm_kmajor_kernel : OUT[M,N] = epi(A[M,K] * B[N,K]^T)   forward 1x1 conv and its data gradient.
//      Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread
//      tcgen05.mma issuer, warps 2..5 = epilogue (TMEM -> registers -> fused BN/act/residual -> HBM).
//      Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
//      These layers are HBM-bound (AI = K*N/(K+N) flop/B): the kernel's job is to stream
//      A once and write OUT once; weights stay L2-resident.
//  gemm_wgrad_kernel  : dW[Cin,Cout] = sum_m X[m,Cin] * dZ[m,Cout]  (filter gradient).
//      Both operands are consumed MN-major straight from their NHWC layout (no transpose pass);
//      the pixel dimension is split across CTAs, partials are reduced in a fixed order (deterministic).
//
// Replaces: cuDNN implicit-GEMM calls behind tf.Session.run for every 1x1 `Conv2D` node of
// checkpoints/*/model.meta and their gradients (reference SemanticNetwork.py:260, :179).
#include "gemm.cuh"
#include "tcgen05.cuh"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdlib>

namespace ams {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;    // 6 warps
constexpr int kMaxStages = 8;
constexpr size_t kSmemBudget = 200 * 1024;
constexpr size_t kWgradSmemBudget = 150 * 1024;    // measured: 100 / 150 / 200 KB -> 152.3 / 155.5 / 153.2 steps/s

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 2-D bf16 tensor [outer][inner] with a 128B-swizzled box; out-of-bounds elements read as zero.
int encode_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                   uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128) {
    auto fn = get_encode_fn();
    AMS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    AMS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    AMS_REQUIRE(row_stride_bytes % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
    AMS_REQUIRE(box_inner * 2 <= static_cast<uint32_t>(swizzle_bytes) && box_outer <= 256, "TMA box too large");
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return 0;
}

uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while (c < static_cast<uint32_t>(n)) c <<= 1;
    return c;
}

// =============================================================================================
//                                   K-major GEMM (forward / dgrad)
// =============================================================================================
struct GemmKParams {
    int M, N, K;
    int block_n, n_tiles, num_tiles, k_blocks, stages, n_alloc;
    int b_split;                       // two weight planes per k-block (hi at the stage base, lo one tile further)
    uint32_t tmem_cols;
    void* out; int ldc; int out_fp32; int out_fp16, a_bf16, b_bf16;
    const float* scale; const float* shift; const float* rowbias; int rows_per_image;
    const void* residual; int ldr;
    int act;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmB2, const GemmKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_a = BLOCK_M * BLOCK_K * 2;
    const int tile_b = p.block_n * BLOCK_K * 2;
    const int stage_b = p.b_split ? 2 * tile_b : tile_b;
    uint8_t* smA = smem;
    uint8_t* smB = smem + p.stages * stage_a;
    float* s_scale = reinterpret_cast<float*>(smB + p.stages * stage_b);
    float* s_shift = s_scale + p.n_alloc;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + p.n_alloc);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmA);
        t5::tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { t5::mbar_init(&full_bar[s], 1); t5::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { t5::mbar_init(&tfull_bar[s], 1); t5::mbar_init(&tempty_bar[s], 128); }
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, p.tmem_cols); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    for (int n = threadIdx.x; n < p.n_alloc; n += kGemmThreads) {
        s_scale[n] = (n < p.N) ? (p.scale ? p.scale[n] : 1.f) : 0.f;
        s_shift[n] = (n < p.N && p.shift) ? p.shift[n] : 0.f;
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int m_tile = t / p.n_tiles, n_tile = t % p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    t5::mbar_wait(&empty_bar[stage], phase ^ 1);
                    t5::mbar_arrive_expect_tx(&full_bar[stage], stage_a + stage_b);
                    t5::tma_load_2d(smA + stage * stage_a, &tmA, &full_bar[stage], kb * BLOCK_K, m_tile * BLOCK_M);
                    t5::tma_load_2d(smB + stage * stage_b, &tmB, &full_bar[stage], kb * BLOCK_K, n_tile * p.block_n);
                    if (p.b_split) t5::tma_load_2d(smB + stage * stage_b + tile_b, &tmB2, &full_bar[stage], kb * BLOCK_K, n_tile * p.block_n);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = t5::make_idesc_f16(BLOCK_M, p.block_n, 0, 0, p.a_bf16, p.b_bf16);
        int stage = 0; uint32_t phase = 0; int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
            t5::mbar_wait(&tempty_bar[as], aphase ^ 1);
            t5::fence_after_thread_sync();
            const uint32_t tmem_d = tmem_base + as * p.block_n;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                t5::mbar_wait(&full_bar[stage], phase);
                t5::fence_after_thread_sync();
                {
                    // whole converged warp, one elected lane per instruction (see the v2 kernel); K-major SWIZZLE_128B: 8-row
                    // groups 1024 B apart (SBO); +32 B (+2 in descriptor units) per UMMA_K step
                    const uint64_t da0 = t5::make_smem_desc_sw128(t5::smem_u32(smA + stage * stage_a), 16, 1024);
                    const uint64_t db0 = t5::make_smem_desc_sw128(t5::smem_u32(smB + stage * stage_b), 16, 1024);
                    const uint64_t dl0 = db0 + (tile_b >> 4);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        t5::mma_f16_ss_warp(tmem_d, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                        if (p.b_split) t5::mma_f16_ss_warp(tmem_d, da0 + 2 * k, dl0 + 2 * k, idesc, 1u);
                    }
                    t5::mma_commit_warp(&empty_bar[stage]);                    // frees the smem slot when MMAs retire
                    if (kb == p.k_blocks - 1) t5::mma_commit_warp(&tfull_bar[as]);   // accumulator ready
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const bool st256 = (!p.out_fp32) && (p.ldc % 16 == 0);
        int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int m_tile = t / p.n_tiles, n_tile = t % p.n_tiles;
            const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
            t5::mbar_wait(&tfull_bar[as], aphase);
            t5::fence_after_thread_sync();
            const long long m = static_cast<long long>(m_tile) * BLOCK_M + row;
            const bool row_ok = m < p.M;
            const float* rb = nullptr;
            if (p.rowbias && row_ok) rb = p.rowbias + (m / p.rows_per_image) * p.N;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.block_n;
            for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                uint32_t r[16];
                t5::tmem_ld16(taddr0 + c0, r);
                t5::tmem_ld_wait();
                const int n0 = n_tile * p.block_n + c0;
                if (!row_ok) continue;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (rb) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n0 + j < p.N) v[j] += rb[n0 + j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = act_apply(fmaf(v[j], s_scale[n0 + j], s_shift[n0 + j]), p.act);
                if (p.out_fp32) {
                    float* o = reinterpret_cast<float*>(p.out) + m * p.ldc + n0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (n0 + j < p.ldc) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    const int valid = p.N - n0;       // N % 8 == 0 for every bf16 output
                    if (valid <= 0) continue;
                    if (p.residual) {
                        const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.residual) + m * p.ldr + n0;
                        float f[8];
                        if (p.out_fp16) unpack8h(ldg_stream(rp), f); else unpack8(ldg_stream(rp), f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += f[j];
                        if (valid >= 16) {
                            if (p.out_fp16) unpack8h(ldg_stream(rp + 8), f); else unpack8(ldg_stream(rp + 8), f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 + j] += f[j];
                        }
                    }
                    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + m * p.ldc + n0;
                    const uint4 lo = p.out_fp16 ? pack8h(v) : pack8(v), hi = p.out_fp16 ? pack8h(v + 8) : pack8(v + 8);
                    if (valid >= 16) {
                        if (st256) stg256(o, lo, hi);
                        else { stg_stream(o, lo); stg_stream(o + 8, hi); }
                    } else {
                        stg_stream(o, lo);
                    }
                }
            }
            t5::fence_before_thread_sync();
            t5::mbar_arrive(&tempty_bar[as]);
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        t5::fence_after_thread_sync();
        t5::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// =============================================================================================
//                     K-major GEMM v2: shared-memory staged epilogue, TMA store, fused BN statistics
// =============================================================================================
// 10 warps: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue (two warps per TMEM
// lane quarter, alternating 16-column chunks).  The bf16 output tile is staged in shared memory in the
// SWIZZLE_128B layout of the output tensor map (bank-conflict-free 16-byte writes), optionally reduced column-wise
// for the BatchNorm batch statistics, and written with cp.async.bulk.tensor (full-line coalesced HBM writes, M/N
// tails clipped by the TMA unit).
constexpr int kGemm2Threads = 320;
constexpr int kEpiThreads = 256;

struct Gemm2Params {
    int M, N, K;
    int block_n, n_tiles, num_tiles, k_blocks, stages, n_alloc, stage_bufs, nboxes, acc_stages, block_k;
    int b_resident;                              // single B tile (k_blocks == 1, n_tiles == 1): loaded once per CTA
    int b_split;                                 // two weight planes per B slot: hi at the slot base, lo one tile further
    int linear_out, pitch, cbuf_bytes;          // linear_out: dense padded staging + coalesced copy-out (n_tiles == 1)
    uint16_t* out; int ldc;             // fp16 (kEpiOutF16) or bf16 elements
    int a_bf16, b_bf16;                 // operand element formats of the instruction descriptor
    uint32_t tmem_cols;
    const float* scale; const float* shift; const float* rowbias; int rows_per_image;
    const uint16_t* residual; int ldr;  // same element type as out
    int act;
    double* stats_partial;
};

enum : int { kEpiAffine = 1, kEpiResidual = 2, kEpiRowBias = 4, kEpiStats = 8, kEpiOutF16 = 16 };

__device__ __forceinline__ void lds8(const float* p, float* o) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ void fadd2(float2& d, const float2& a) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)));
}

template <int F>
__global__ void __launch_bounds__(kGemm2Threads, 2)
gemm_kmajor_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC, const Gemm2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_a = BLOCK_M * p.block_k * 2;
    const int tile_b = p.block_n * p.block_k * 2;
    const int stage_b = p.b_split ? 2 * tile_b : tile_b;
    const int cbuf_bytes = p.cbuf_bytes;    // TMA mode: nboxes x [128 rows][128 B] swizzled; linear mode: [128 rows][pitch]
    uint8_t* smA = smem;
    uint8_t* smB = smA + p.stages * stage_a;
    uint8_t* smC = smB + (p.b_resident ? 1 : p.stages) * stage_b;                          // 1024-aligned (stage sizes are multiples of 2 KB)
    float* s_scale = reinterpret_cast<float*>(smC + p.stage_bufs * cbuf_bytes);
    float* s_shift = s_scale + p.n_alloc;
    // statistics scratch only in the kernels that compute statistics: 16 B per column + 128 B per tile column are the
    // difference between two and three pipeline stages for the K, N >= 728 GEMMs of the teacher
    constexpr bool kHasStats = (F & kEpiStats) != 0;
    double* s_run = reinterpret_cast<double*>(s_shift + p.n_alloc);   // [2][n_alloc] running column sums of this CTA
    float2* s_part = reinterpret_cast<float2*>(s_run + (kHasStats ? 2 * p.n_alloc : 0));   // [16 slices][block_n]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_part + (kHasStats ? 16 * p.block_n : 0));
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* bres_bar = tempty_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmA);
        t5::tma_prefetch_desc(&tmB);
        t5::tma_prefetch_desc(&tmC);
        for (int s = 0; s < p.stages; ++s) { t5::mbar_init(&full_bar[s], 1); t5::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { t5::mbar_init(&tfull_bar[s], 1); t5::mbar_init(&tempty_bar[s], kEpiThreads); }
        t5::mbar_init(bres_bar, 1);
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, p.tmem_cols); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    for (int n = threadIdx.x; n < p.n_alloc; n += kGemm2Threads) {
        s_scale[n] = (n < p.N) ? (p.scale ? p.scale[n] : 1.f) : 0.f;
        s_shift[n] = (n < p.N && p.shift) ? p.shift[n] : 0.f;
        if (kHasStats) { s_run[n] = 0.0; s_run[p.n_alloc + n] = 0.0; }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (p.b_resident) {
                t5::mbar_arrive_expect_tx(bres_bar, stage_b);
                t5::tma_load_2d(smB, &tmB, bres_bar, 0, 0);
                if (p.b_split) t5::tma_load_2d(smB + tile_b, &tmB2, bres_bar, 0, 0);
            }
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int m_tile = t / p.n_tiles, n_tile = t % p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    t5::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                    t5::mbar_arrive_expect_tx(&full_bar[stage], stage_a + (p.b_resident ? 0 : stage_b));
                    t5::tma_load_2d(smA + stage * stage_a, &tmA, &full_bar[stage], kb * p.block_k, m_tile * BLOCK_M);
                    if (!p.b_resident) {
                        t5::tma_load_2d(smB + stage * stage_b, &tmB, &full_bar[stage], kb * p.block_k, n_tile * p.block_n);
                        if (p.b_split) t5::tma_load_2d(smB + stage * stage_b + tile_b, &tmB2, &full_bar[stage], kb * p.block_k, n_tile * p.block_n);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = t5::make_idesc_f16(BLOCK_M, p.block_n, 0, 0, p.a_bf16, p.b_bf16);
        int stage = 0; uint32_t phase = 0; int it = 0;
        // K-major swizzled rows of block_k*2 bytes (128/64/32-byte swizzle): 8-row groups SBO apart
        const uint64_t desc_fixed = t5::make_smem_desc(0, 16, 16u * p.block_k, p.block_k == 64 ? 2u : (p.block_k == 32 ? 4u : 6u));
        const int ksteps = p.block_k / UMMA_K;
        if (p.b_resident) t5::mbar_wait_relaxed(bres_bar, 0);
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int as = (p.acc_stages == 2) ? (it & 1) : 0;
            const uint32_t aphase = (p.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
            t5::mbar_wait_relaxed(&tempty_bar[as], aphase ^ 1);
            t5::fence_after_thread_sync();
            const uint32_t tmem_d = tmem_base + as * p.block_n;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                t5::mbar_wait_relaxed(&full_bar[stage], phase);
                t5::fence_after_thread_sync();
                {
                    // Issued by the whole converged warp with warp-uniform operands, one elected lane per instruction, and
                    // descriptors that differ only by an add in the address field (16-byte units: a UMMA_K step of 32 bytes is
                    // +2).  Under `if (lane == 0)` with the descriptors rebuilt per instruction this warp needed 200-800 cycles
                    // per tcgen05.mma while the epilogue warps of its scheduler were busy (tools/micro/mma_contention.cu).
                    const uint64_t da0 = desc_fixed | ((t5::smem_u32(smA + stage * stage_a) & 0x3FFFFu) >> 4);
                    const uint64_t db0 = desc_fixed | ((t5::smem_u32(smB + (p.b_resident ? 0 : stage) * stage_b) & 0x3FFFFu) >> 4);
                    const uint64_t dl0 = db0 + (tile_b >> 4);         // low weight plane: same tile shape, tile_b bytes further
                    for (int k = 0; k < ksteps; ++k) {
                        t5::mma_f16_ss_warp(tmem_d, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                        if (p.b_split) t5::mma_f16_ss_warp(tmem_d, da0 + 2 * k, dl0 + 2 * k, idesc, 1u);
                    }
                    t5::mma_commit_warp(&empty_bar[stage]);
                    if (kb == p.k_blocks - 1) t5::mma_commit_warp(&tfull_bar[as]);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        constexpr bool OH = (F & kEpiOutF16) != 0;    // 16-bit output / residual element type: fp16 or bf16
        const int et = threadIdx.x - 64;              // 0..255
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;             // which of the two warps of this quarter
        const int row = q * 32 + lane;
        const int nchunks = p.block_n >> 4;
        // statistics slicing: nslices * (block_n / 8) <= 256 threads, nslices a power of two <= 16
        const int ncg = p.block_n >> 3;
        int nslices = 1;
        while (nslices < 16 && nslices * 2 * ncg <= kEpiThreads) nslices <<= 1;
        const int rows_per_slice = BLOCK_M / nslices;
        // division-free copy-out: this thread's first 16-byte chunk (row, chunk) and its step per 256 threads
        const int cpr = max(p.N >> 3, 1);
        const int co_r0 = et / cpr, co_c0 = et - co_r0 * cpr;
        const int co_dr = kEpiThreads / cpr, co_dc = kEpiThreads - co_dr * cpr;
        int it = 0, buf = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int m_tile = t / p.n_tiles, n_tile = t % p.n_tiles;
            const int as = (p.acc_stages == 2) ? (it & 1) : 0;
            const uint32_t aphase = (p.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
            uint8_t* cbuf = smC + buf * cbuf_bytes;
            const uint32_t cbuf_s = t5::smem_u32(cbuf);
            // the TMA store that last read this staging buffer must have finished reading it
            if (!p.linear_out && et == 0) { if (p.stage_bufs == 2) t5::tma_store_wait_read<1>(); else t5::tma_store_wait_read<0>(); }
            t5::named_barrier_sync(1, kEpiThreads);
            t5::mbar_wait(&tfull_bar[as], aphase);
            t5::fence_after_thread_sync();
            const long long m = static_cast<long long>(m_tile) * BLOCK_M + row;
            const bool row_ok = m < p.M;
            const float* rb = nullptr;
            if ((F & kEpiRowBias) && row_ok) rb = p.rowbias + (m / p.rows_per_image) * p.N;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.block_n;
            const float act_lo = p.act ? 0.f : -INFINITY, act_hi = (p.act == 2) ? 6.f : INFINITY;
            auto process = [&](const uint32_t* r, int j) {
                const int c0 = j << 4;
                const int n0 = n_tile * p.block_n + c0;
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
                if (row_ok) {
                    if (F & kEpiRowBias) {
                        if (n0 + 16 <= p.N) {
                            float b[8];
                            lds8(rb + n0, b);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += b[k];
                            lds8(rb + n0 + 8, b);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[8 + k] += b[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k) if (n0 + k < p.N) v[k] += rb[n0 + k];
                        }
                    }
                    if (F & kEpiAffine) {
#pragma unroll
                        for (int h8 = 0; h8 < 16; h8 += 8) {
                            float sc[8], sh[8];
                            lds8(s_scale + n0 + h8, sc);
                            lds8(s_shift + n0 + h8, sh);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[h8 + k] = fminf(fmaxf(fmaf(v[h8 + k], sc[k], sh[k]), act_lo), act_hi);
                        }
                    }
                    if (F & kEpiResidual) {
                        const uint16_t* rp = p.residual + m * p.ldr + n0;
                        float f[8];
                        if (n0 < p.N) {
                            unpack8t<OH>(ldg_stream(rp), f);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += f[k];
                        }
                        if (n0 + 8 < p.N) {
                            unpack8t<OH>(ldg_stream(rp + 8), f);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[8 + k] += f[k];
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = 0.f;
                }
                // swizzled staging write: box = 64 columns, 16-byte chunk index XOR (row & 7)
                if (p.linear_out) {
                    // dense rows of `pitch` bytes (pitch/16 odd => 8 consecutive rows hit 8 different 16-byte bank groups)
                    const uint32_t rowp = cbuf_s + row * p.pitch + c0 * 2;
                    if (c0 < p.N) sts128(rowp, pack8t<OH>(v));
                    if (c0 + 8 < p.N) sts128(rowp + 16, pack8t<OH>(v + 8));
                } else {
                    const uint32_t rowp = cbuf_s + (c0 >> 6) * (BLOCK_M * 128) + row * 128;
                    const int ci = (c0 & 63) >> 3;
                    sts128(rowp + ((ci ^ (row & 7)) << 4), pack8t<OH>(v));
                    sts128(rowp + (((ci + 1) ^ (row & 7)) << 4), pack8t<OH>(v + 8));
                }
            };
            for (int j = half; j < nchunks; j += 4) {
                // two TMEM loads in flight per wait
                uint32_t r0[16], r1[16];
                const bool two = (j + 2) < nchunks;
                t5::tmem_ld16(taddr0 + (j << 4), r0);
                if (two) t5::tmem_ld16(taddr0 + ((j + 2) << 4), r1);
                t5::tmem_ld_wait();
                process(r0, j);
                if (two) process(r1, j + 2);
            }
            t5::fence_before_thread_sync();
            t5::mbar_arrive(&tempty_bar[as]);                    // TMEM stage free for the next MMAs
            if (!p.linear_out) t5::fence_proxy_async_smem();     // generic-proxy writes -> visible to the TMA unit
            t5::named_barrier_sync(2, kEpiThreads);
            if (F & kEpiStats) {
                // column statistics of the stored bf16 tile: thread = (slice of rows, group of 8 columns), 128-bit
                // shared-memory reads, packed fp32x2 accumulation
                const int g8 = et % ncg, slice = et / ncg;
                if (slice < nslices) {
                    float2 s2[4], q2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) s2[k] = q2[k] = make_float2(0.f, 0.f);
                    const int r0 = slice * rows_per_slice;
                    if (p.linear_out) {
                        uint32_t a = cbuf_s + r0 * p.pitch + (g8 << 4);
#pragma unroll 4
                        for (int rr = 0; rr < rows_per_slice; ++rr, a += p.pitch) {
                            const uint4 v = lds128(a);
                            const float2 f0 = unpack2t<OH>(v.x), f1 = unpack2t<OH>(v.y), f2 = unpack2t<OH>(v.z), f3 = unpack2t<OH>(v.w);
                            fadd2(s2[0], f0); fadd2(s2[1], f1); fadd2(s2[2], f2); fadd2(s2[3], f3);
                            ffma2(q2[0], f0, f0); ffma2(q2[1], f1, f1); ffma2(q2[2], f2, f2); ffma2(q2[3], f3, f3);
                        }
                    } else {
                        const uint32_t a0 = cbuf_s + (g8 >> 3) * (BLOCK_M * 128);
                        const int ci = g8 & 7;
#pragma unroll 4
                        for (int rr = r0; rr < r0 + rows_per_slice; ++rr) {
                            const uint4 v = lds128(a0 + rr * 128 + ((ci ^ (rr & 7)) << 4));
                            const float2 f0 = unpack2t<OH>(v.x), f1 = unpack2t<OH>(v.y), f2 = unpack2t<OH>(v.z), f3 = unpack2t<OH>(v.w);
                            fadd2(s2[0], f0); fadd2(s2[1], f1); fadd2(s2[2], f2); fadd2(s2[3], f3);
                            ffma2(q2[0], f0, f0); ffma2(q2[1], f1, f1); ffma2(q2[2], f2, f2); ffma2(q2[3], f3, f3);
                        }
                    }
                    float2* dstp = s_part + slice * p.block_n + (g8 << 3);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        dstp[2 * k] = make_float2(s2[k].x, q2[k].x);
                        dstp[2 * k + 1] = make_float2(s2[k].y, q2[k].y);
                    }
                }
                t5::named_barrier_sync(3, kEpiThreads);
                if (et < p.block_n) {
                    // 128 rows of one tile: fp32 is enough (fixed order); the running sums over tiles are fp64
                    float s = 0.f, sq = 0.f;
#pragma unroll 4
                    for (int sl = 0; sl < nslices; ++sl) { const float2 v2 = s_part[sl * p.block_n + et]; s += v2.x; sq += v2.y; }
                    s_run[n_tile * p.block_n + et] += static_cast<double>(s);
                    s_run[p.n_alloc + n_tile * p.block_n + et] += static_cast<double>(sq);
                }
            }
            if (p.linear_out) {
                // the tile is one contiguous run of global memory (all N columns, ldc == N): fully coalesced 16-byte stores
                const int rows_valid = static_cast<int>(min(static_cast<long long>(BLOCK_M), p.M - static_cast<long long>(m_tile) * BLOCK_M));
                const int total = rows_valid * cpr;
                uint4* gdst = reinterpret_cast<uint4*>(p.out + static_cast<long long>(m_tile) * BLOCK_M * p.ldc);
                int rr = co_r0, cc = co_c0;
                for (int i = et; i < total; i += kEpiThreads) {
                    stg_stream(gdst + i, lds128(cbuf_s + rr * p.pitch + (cc << 4)));
                    cc += co_dc; rr += co_dr;
                    if (cc >= cpr) { cc -= cpr; ++rr; }
                }
            } else if (et == 0) {
                for (int b = 0; b < p.nboxes; ++b)
                    t5::tma_store_2d(&tmC, cbuf + b * (BLOCK_M * 128), n_tile * p.block_n + b * 64, m_tile * BLOCK_M);
                t5::tma_store_commit();
            }
            if (p.stage_bufs == 2) buf ^= 1;
        }
        if (et == 0) t5::tma_store_wait_all<0>();
        if (F & kEpiStats) {
            t5::named_barrier_sync(1, kEpiThreads);
            double* dst = p.stats_partial + static_cast<long long>(blockIdx.x) * 2 * p.N;
            for (int n = et; n < p.N; n += kEpiThreads) { dst[n] = s_run[n]; dst[p.N + n] = s_run[p.n_alloc + n]; }
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        t5::fence_after_thread_sync();
        t5::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// =============================================================================================
//                                   MN-major split-K GEMM (wgrad)
// =============================================================================================
struct WgradKParams {
    int Cin, Cout, block_n, boxes_b, co_tiles, ci_tiles, splits, kb_per_split, k_blocks, stages;
    uint32_t tmem_cols;
    int convert_x;         // X tiles arrive as fp16 and are rewritten in place as bf16 before the MMAs (see the kernel)
    float* out;            // splits==1: dW [Cin][lddw]; else workspace [split][Cin][Cout]
    int ld_out;
    long long split_stride;
};

constexpr int kWgradBoxBytes = BLOCK_K * 128;      // one TMA box: 64 pixels x 64 channels bf16 = 8 KB

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ,
                  const WgradKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_a = 2 * kWgradBoxBytes;                 // 128 input channels
    const int stage_b = p.boxes_b * kWgradBoxBytes;
    uint8_t* smA = smem;
    uint8_t* smB = smem + p.stages * stage_a;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + p.stages * stage_b);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* done_bar = empty_bar + kMaxStages;
    uint64_t* conv_bar = done_bar + 1;                      // [kMaxStages]: X tile of the stage converted to bf16
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(conv_bar + kMaxStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // work item
    const int split = blockIdx.x % p.splits;
    const int tile = blockIdx.x / p.splits;
    const int co_tile = tile % p.co_tiles, ci_tile = tile / p.co_tiles;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
    const int nkb = max(kb1 - kb0, 0);

    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmX);
        t5::tma_prefetch_desc(&tmZ);
        for (int s = 0; s < p.stages; ++s) { t5::mbar_init(&full_bar[s], 1); t5::mbar_init(&empty_bar[s], 1); t5::mbar_init(&conv_bar[s], 128); }
        t5::mbar_init(done_bar, 1);
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, p.tmem_cols); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                t5::mbar_wait(&empty_bar[stage], phase ^ 1);
                t5::mbar_arrive_expect_tx(&full_bar[stage], stage_a + stage_b);
                uint8_t* a = smA + stage * stage_a;
                uint8_t* b = smB + stage * stage_b;
                // box = [64 channels (inner), 64 pixels (outer)]; channels beyond the tensor read as zero
                t5::tma_load_2d(a, &tmX, &full_bar[stage], ci_tile * 128, kb * BLOCK_K);
                t5::tma_load_2d(a + kWgradBoxBytes, &tmX, &full_bar[stage], ci_tile * 128 + 64, kb * BLOCK_K);
                for (int j = 0; j < p.boxes_b; ++j)
                    t5::tma_load_2d(b + j * kWgradBoxBytes, &tmZ, &full_bar[stage], co_tile * p.block_n + j * 64,
                                    kb * BLOCK_K);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // both operands bf16: tcgen05.mma kind::f16 rejects mixed A/B element formats (illegal instruction on sm_100a)
        const uint32_t idesc = t5::make_idesc_f16(BLOCK_M, p.block_n, 1, 1, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < nkb; ++i) {
            t5::mbar_wait(p.convert_x ? &conv_bar[stage] : &full_bar[stage], phase);
            t5::fence_after_thread_sync();
            {
                // whole converged warp, one elected lane per instruction (see gemm_bn_act_kernel_v2).  MN-major SWIZZLE_128B:
                // 64-channel atoms one TMA box apart (LBO), 8-pixel groups 1024 B apart (SBO); a UMMA_K = 16 step advances two
                // pixel groups = 2048 B (+128 in the 16-byte units of the descriptor's address field).
                const uint64_t da0 = t5::make_smem_desc_sw128(t5::smem_u32(smA + stage * stage_a), kWgradBoxBytes, 1024);
                const uint64_t db0 = t5::make_smem_desc_sw128(t5::smem_u32(smB + stage * stage_b), kWgradBoxBytes, 1024);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                    t5::mma_f16_ss_warp(tmem_base, da0 + 128 * k, db0 + 128 * k, idesc, (i | k) != 0);
                t5::mma_commit_warp(&empty_bar[stage]);
                if (i == nkb - 1) t5::mma_commit_warp(done_bar);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else {
        if (p.convert_x) {
            // The activations X are stored in fp16, the gradients dZ in bf16, and the tensor core wants one element format for
            // both operands: the four (otherwise idle) epilogue warps rewrite each X tile in place as bf16 once its TMA has
            // landed -- elementwise on 16-byte chunks, so the 128-byte swizzle of the tile is irrelevant -- and hand the stage
            // to the MMA issuer through conv_bar.  Only the FILTER gradient sees the 8-bit mantissa (a sum over >= 2,145
            // pixels per weight, so the rounding averages out); the forward path and the data gradients never do.
            const int et = threadIdx.x - 64;              // 0..127
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                t5::mbar_wait(&full_bar[stage], phase);
                const uint32_t a_addr = t5::smem_u32(smA + stage * stage_a);
#pragma unroll
                for (int j = 0; j < (2 * kWgradBoxBytes) / (128 * 16); ++j) {
                    const uint32_t addr = a_addr + (j * 128 + et) * 16;
                    float f[8];
                    unpack8h(lds128(addr), f);
                    sts128(addr, pack8(f));
                }
                t5::fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core's smem reads
                t5::mbar_arrive(&conv_bar[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ci = ci_tile * 128 + row;
        float* o = p.out + static_cast<long long>(split) * p.split_stride + static_cast<long long>(ci) * p.ld_out;
        if (nkb > 0) {
            t5::mbar_wait(done_bar, 0);
            t5::fence_after_thread_sync();
        }
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int c0 = 0; c0 < p.block_n; c0 += 16) {
            uint32_t r[16];
            if (nkb > 0) { t5::tmem_ld16(taddr0 + c0, r); t5::tmem_ld_wait(); }
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            const int n0 = co_tile * p.block_n + c0;
            if (ci < p.Cin) {
                if (n0 + 16 <= p.Cout && (reinterpret_cast<uintptr_t>(o + n0) & 15) == 0) {
                    // 64 contiguous bytes per thread: four 128-bit stores (two full sectors) instead of 16 scalar ones
                    float4* o4 = reinterpret_cast<float4*>(o + n0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                            __uint_as_float(r[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < p.Cout) o[n0 + j] = __uint_as_float(r[j]);
                }
            }
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        t5::fence_after_thread_sync();
        t5::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// Fixed split order => deterministic.  Few splits: one thread per 4 outputs walks the splits; many splits (the
// high-resolution layers, ~300 splits of a tiny dW): one warp per 4 outputs, lanes stride the splits, fixed shuffle tree.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cin, int Cout, int lddw, int splits,
                    long long split_stride, int warp_per_item) {
    pdl_entry();
    const long long n = static_cast<long long>(Cin) * Cout;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long i4 = (warp_per_item ? (tid >> 5) : tid) * 4;
    if (i4 >= n) return;
    const int s0 = warp_per_item ? lane : 0, ds = warp_per_item ? 32 : 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if ((Cout & 3) == 0) {
#pragma unroll 8
        for (int s = s0; s < splits; s += ds) {
            const float4 v = *reinterpret_cast<const float4*>(ws + s * split_stride + i4);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
        }
    } else {
        for (int s = s0; s < splits; s += ds)
            for (int k = 0; k < 4; ++k) if (i4 + k < n) acc[k] += ws[s * split_stride + i4 + k];
    }
    if (warp_per_item) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
        if (lane != 0) return;
    }
    for (int k = 0; k < 4; ++k)
        if (i4 + k < n) dw[((i4 + k) / Cout) * lddw + ((i4 + k) % Cout)] = acc[k];
}

}  // namespace

// ============================================================================================= host
static bool use_v1() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("AMS_GEMM_V1"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int gemm_plan(const GemmDesc& d, int num_sms, GemmPlan* plan) {
    AMS_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "empty GEMM");
    AMS_REQUIRE(d.a_fp16 == d.b_fp16, "tcgen05.mma kind::f16 needs the same element format for A and B");
    AMS_REQUIRE(d.lda % 8 == 0 && d.ldb % 8 == 0, "bf16 operand strides must be multiples of 8 elements");
    AMS_REQUIRE(d.out_fp32 || (d.N % 8 == 0 && d.ldc % 8 == 0), "bf16 output needs N, ldc multiples of 8");
    AMS_REQUIRE(!d.out_fp32 || d.ldc % 4 == 0, "fp32 output needs ldc multiple of 4");
    AMS_REQUIRE(!d.residual || d.ldr % 8 == 0, "residual stride must be a multiple of 8");
    plan->d = d;
    plan->v2 = (!d.out_fp32 && !use_v1()) ? 1 : 0;
    AMS_REQUIRE(plan->v2 || !d.stats_partial, "fused statistics need the bf16 (v2) epilogue");
    int npad = ceil_div(d.N, 16) * 16;
    if (npad <= 256) {
        plan->block_n = npad;
        plan->n_tiles = 1;
    } else {
        // several N tiles: whole 64-column TMA boxes per tile (the last tile is clipped by the tensor extent)
        int t = ceil_div(npad, 256);
        plan->block_n = ceil_div(ceil_div(npad, t), 64) * 64;
        plan->n_tiles = ceil_div(npad, plan->block_n);
        if (d.K >= 512) {
            // big-K GEMMs run one persistent CTA per SM, so whole waves of tiles count: N = 728 at M = 8,385 is 198 tiles of 256
            // columns (two rounds on 148 SMs, the second one third full) or 264 tiles of 192 (two rounds, 0.75x the work each)
            const int m_tiles = ceil_div(d.M, BLOCK_M);
            long long best = -1;
            for (int tt = t; tt <= t + 2; ++tt) {
                const int bn = ceil_div(ceil_div(npad, tt), 64) * 64;
                const int nt = ceil_div(npad, bn);
                const long long cost = ceil_div_ll(static_cast<long long>(m_tiles) * nt, num_sms) * (bn + 64);
                if (best < 0 || cost < best) { best = cost; plan->block_n = bn; plan->n_tiles = nt; }
            }
        }
    }
    npad = plan->block_n * plan->n_tiles;                       // = n_alloc of the kernels
    plan->m_tiles = ceil_div(d.M, BLOCK_M);
    // v2 runs two CTAs per SM (more epilogue warps in flight): each gets <= 256 TMEM columns and <= ~110 KB smem
    plan->acc_stages = (!plan->v2 || 2 * plan->block_n <= 256) ? 2 : 1;
    plan->tmem_cols = tmem_cols_for(plan->acc_stages * plan->block_n);
    AMS_REQUIRE(plan->tmem_cols <= (plan->v2 ? 256u : 512), "TMEM overflow");
    // small-K layers: a TMA box as wide as the row (32/64-byte swizzle) instead of a mostly out-of-bounds 128-byte box;
    // a narrower k-block is also the way out when two stages of the widest box do not fit (split weights on N = 256)
    int block_k = (plan->v2 && d.K <= 16) ? 16 : ((plan->v2 && d.K <= 32) ? 32 : BLOCK_K);
    for (;; block_k /= 2) {
        plan->block_k = block_k;
        plan->k_blocks = ceil_div(d.K, plan->block_k);
        plan->b_resident = (plan->v2 && plan->k_blocks == 1 && plan->n_tiles == 1) ? 1 : 0;
        const size_t b_tile = size_t(plan->block_n) * plan->block_k * 2 * (d.B_lo ? 2 : 1);
        AMS_REQUIRE(!d.B_lo || (size_t(plan->block_n) * plan->block_k * 2) % 512 == 0, "split weights: B tile must be a multiple of 512 bytes");
        const size_t stage_bytes = size_t(BLOCK_M) * plan->block_k * 2 + (plan->b_resident ? 0 : b_tile);
        size_t fixed = 1024 /*align slack*/ + size_t(npad) * 8 + (2 * kMaxStages + 6) * 8 + 16 + (plan->b_resident ? b_tile : 0);
        size_t budget = kSmemBudget;
        if (plan->v2) {
            const int nboxes = ceil_div(plan->block_n, 64);
            plan->linear_out = (plan->n_tiles == 1 && d.ldc == d.N) ? 1 : 0;
            plan->pitch = d.N * 2 + (((d.N / 8) % 2 == 0) ? 16 : 0);
            plan->cbuf_bytes = plan->linear_out ? ((BLOCK_M * plan->pitch + 1023) / 1024) * 1024 : nboxes * BLOCK_M * 128;
            plan->stage_bufs = (plan->cbuf_bytes <= 16 * 1024) ? 2 : 1;
            fixed += size_t(plan->stage_bufs) * plan->cbuf_bytes + (d.stats_partial ? size_t(npad) * 16 + size_t(plan->block_n) * 16 * 8 : 0);
            // two CTAs per SM where two pipeline stages fit next to the fixed part in half an SM; else the whole SM for one CTA:
            // a K >= 728 mainloop is bound by the bytes in flight (stages x 48 KB against ~1.4 us of TMA latency)
            budget = (fixed + 2 * stage_bytes <= 110 * 1024) ? 110 * 1024 : 224 * 1024;
        }
        int stages = budget > fixed ? int((budget - fixed) / stage_bytes) : 0;
        stages = std::max(2, std::min(stages, kMaxStages));
        plan->stages = stages;
        plan->smem_bytes = fixed + stages * stage_bytes;
        if (plan->smem_bytes <= 227 * 1024 || block_k == 16 || !plan->v2) break;
    }
    AMS_REQUIRE(plan->smem_bytes <= 227 * 1024, "GEMM shared memory overflow");
    // a plan that fills more than half of the SM's shared memory runs ONE CTA per SM anyway: give it the whole tensor memory,
    // i.e. two accumulator stages also for 256-column tiles, so that the epilogue of tile i overlaps the MMAs of tile i+1
    // (the K = N = 728 .. 2048 pointwise convs of the teacher network: the tensor-bound GEMMs of the repository)
    if (plan->v2 && plan->acc_stages == 1 && plan->smem_bytes > 113 * 1024) {
        plan->acc_stages = 2;
        plan->tmem_cols = tmem_cols_for(2 * plan->block_n);
        AMS_REQUIRE(plan->tmem_cols <= 512, "TMEM overflow");
    }
    const int tiles = plan->m_tiles * plan->n_tiles;
    plan->grid = std::min(tiles, (plan->v2 && plan->smem_bytes <= 113 * 1024) ? 2 * num_sms : num_sms);
    if (encode_2d_bf16(&plan->tmA, d.A, d.K, d.M, size_t(d.lda) * 2, plan->block_k, BLOCK_M, plan->block_k * 2)) return -1;
    if (encode_2d_bf16(&plan->tmB, d.B, d.K, d.N, size_t(d.ldb) * 2, plan->block_k, plan->block_n, plan->block_k * 2)) return -1;
    if (encode_2d_bf16(&plan->tmB2, d.B_lo ? d.B_lo : d.B, d.K, d.N, size_t(d.ldb) * 2, plan->block_k, plan->block_n, plan->block_k * 2)) return -1;
    if (plan->v2 && encode_2d_bf16(&plan->tmC, d.out, d.N, d.M, size_t(d.ldc) * 2, 64, BLOCK_M)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(gemm_kmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#define AMS_G2A(FL) AMS_CUDA_CHECK(cudaFuncSetAttribute(gemm_kmajor_v2_kernel<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_G2A(0) AMS_G2A(1) AMS_G2A(2) AMS_G2A(3) AMS_G2A(4) AMS_G2A(5) AMS_G2A(6) AMS_G2A(7)
        AMS_G2A(8) AMS_G2A(9) AMS_G2A(10) AMS_G2A(11) AMS_G2A(12) AMS_G2A(13) AMS_G2A(14) AMS_G2A(15)
        AMS_G2A(16) AMS_G2A(17) AMS_G2A(18) AMS_G2A(19) AMS_G2A(20) AMS_G2A(21) AMS_G2A(22) AMS_G2A(23)
        AMS_G2A(24) AMS_G2A(25) AMS_G2A(26) AMS_G2A(27) AMS_G2A(28) AMS_G2A(29) AMS_G2A(30) AMS_G2A(31)
#undef AMS_G2A
        attr_set = true;
    }
    return 0;
}

int gemm_launch(const GemmPlan& pl, cudaStream_t stream) {
    const GemmDesc& d = pl.d;
    if (pl.v2) {
        Gemm2Params p;
        p.M = d.M; p.N = d.N; p.K = d.K;
        p.block_n = pl.block_n; p.n_tiles = pl.n_tiles; p.num_tiles = pl.m_tiles * pl.n_tiles;
        p.k_blocks = pl.k_blocks; p.stages = pl.stages; p.n_alloc = pl.n_tiles * pl.block_n;
        p.stage_bufs = pl.stage_bufs; p.nboxes = ceil_div(pl.block_n, 64); p.acc_stages = pl.acc_stages; p.block_k = pl.block_k; p.b_resident = pl.b_resident;
        p.b_split = d.B_lo ? 1 : 0;
        p.linear_out = pl.linear_out; p.pitch = pl.pitch; p.cbuf_bytes = pl.cbuf_bytes;
        p.out = static_cast<uint16_t*>(d.out); p.ldc = d.ldc;
        p.a_bf16 = d.a_fp16 ? 0 : 1; p.b_bf16 = d.b_fp16 ? 0 : 1;
        p.tmem_cols = pl.tmem_cols;
        p.scale = d.scale; p.shift = d.shift; p.rowbias = d.rowbias; p.rows_per_image = d.rows_per_image;
        p.residual = static_cast<const uint16_t*>(d.residual); p.ldr = d.ldr; p.act = d.act; p.stats_partial = d.stats_partial;
        const int flags = ((d.scale || d.shift || d.act) ? kEpiAffine : 0) | (d.residual ? kEpiResidual : 0) |
                          (d.rowbias ? kEpiRowBias : 0) | (d.stats_partial ? kEpiStats : 0) | (d.out_fp16 ? kEpiOutF16 : 0);
        switch (flags) {
#define AMS_G2(FL) case FL: AMS_LAUNCH((gemm_kmajor_v2_kernel<FL>), pl.grid, kGemm2Threads, pl.smem_bytes, stream, pl.tmA, pl.tmB, pl.tmB2, pl.tmC, p); break;
            AMS_G2(0) AMS_G2(1) AMS_G2(2) AMS_G2(3) AMS_G2(4) AMS_G2(5) AMS_G2(6) AMS_G2(7)
            AMS_G2(8) AMS_G2(9) AMS_G2(10) AMS_G2(11) AMS_G2(12) AMS_G2(13) AMS_G2(14) AMS_G2(15)
            AMS_G2(16) AMS_G2(17) AMS_G2(18) AMS_G2(19) AMS_G2(20) AMS_G2(21) AMS_G2(22) AMS_G2(23)
            AMS_G2(24) AMS_G2(25) AMS_G2(26) AMS_G2(27) AMS_G2(28) AMS_G2(29) AMS_G2(30) AMS_G2(31)
#undef AMS_G2
        }
        return 0;
    }
    GemmKParams p;
    p.M = d.M; p.N = d.N; p.K = d.K;
    p.block_n = pl.block_n; p.n_tiles = pl.n_tiles; p.num_tiles = pl.m_tiles * pl.n_tiles;
    p.k_blocks = pl.k_blocks; p.stages = pl.stages; p.n_alloc = pl.n_tiles * pl.block_n;
    p.tmem_cols = pl.tmem_cols; p.b_split = d.B_lo ? 1 : 0;
    p.out = d.out; p.ldc = d.ldc; p.out_fp32 = d.out_fp32; p.out_fp16 = d.out_fp16;
    p.a_bf16 = d.a_fp16 ? 0 : 1; p.b_bf16 = d.b_fp16 ? 0 : 1;
    p.scale = d.scale; p.shift = d.shift; p.rowbias = d.rowbias; p.rows_per_image = d.rows_per_image;
    p.residual = d.residual; p.ldr = d.ldr; p.act = d.act;
    AMS_LAUNCH((gemm_kmajor_kernel), pl.grid, kGemmThreads, pl.smem_bytes, stream, pl.tmA, pl.tmB, pl.tmB2, p);
    return 0;
}

static void wgrad_shape(int Cin, int Cout, long long M, int num_sms, int* ci_tiles, int* co_tiles, int* block_n,
                        int* boxes_b, int* splits, int* kb_per_split, int* k_blocks) {
    *ci_tiles = ceil_div(Cin, 128);
    const int cpad = ceil_div(Cout, 16) * 16;
    int t = ceil_div(cpad, 256);
    // block_n must be a whole number of 64-channel TMA boxes unless it is the only tile
    int bn = (t == 1) ? cpad : ceil_div(ceil_div(cpad, t), 64) * 64;
    *co_tiles = ceil_div(cpad, bn);
    *block_n = bn;
    *boxes_b = ceil_div(bn, 64);
    *k_blocks = int(ceil_div_ll(M, BLOCK_K));
    const int tiles = (*ci_tiles) * (*co_tiles);
    static const int wg_mult = [] { const char* e = getenv("AMS_WGRAD_CTAS_PER_SM"); return e ? atoi(e) : 1; }();
    int want = std::max(1, wg_mult * num_sms / tiles);     // about one CTA per SM: fewer, longer split-K runs
    int max_splits = std::max(1, *k_blocks / 8);           // at least 8 k-blocks (512 pixels) per CTA
    int s = std::min(want, max_splits);
    *kb_per_split = ceil_div(*k_blocks, s);
    *splits = ceil_div(*k_blocks, *kb_per_split);
}

size_t wgrad_workspace_floats(int Cin, int Cout, long long M, int num_sms) {
    int a, b, c, d, s, e, f;
    wgrad_shape(Cin, Cout, M, num_sms, &a, &b, &c, &d, &s, &e, &f);
    return s > 1 ? size_t(s) * Cin * Cout : 0;
}

int wgrad_plan(const WgradDesc& d, int num_sms, WgradPlan* plan) {
    AMS_REQUIRE(d.ldx % 8 == 0 && d.ldz % 8 == 0, "bf16 operand strides must be multiples of 8 elements");
    plan->d = d;
    wgrad_shape(d.Cin, d.Cout, d.M, num_sms, &plan->ci_tiles, &plan->co_tiles, &plan->block_n, &plan->boxes_b,
                &plan->splits, &plan->kb_per_split, &plan->k_blocks);
    AMS_REQUIRE(plan->splits == 1 || d.workspace_floats >= size_t(plan->splits) * d.Cin * d.Cout,
                "wgrad workspace too small");
    plan->tmem_cols = tmem_cols_for(plan->block_n);
    const size_t stage_bytes = size_t(2 + plan->boxes_b) * kWgradBoxBytes;
    AMS_REQUIRE(!d.z_fp16, "the filter-gradient GEMM takes bf16 gradients");
    const size_t fixed = 1024 + (3 * kMaxStages + 2) * 8 + 16;
    // below the full 227 KB: the filter gradients run on a side stream and must leave room for the chain CTAs on the same SM
    static const size_t wg_budget = [] { const char* e = getenv("AMS_WGRAD_SMEM_KB"); return e ? size_t(atoi(e)) * 1024 : kWgradSmemBudget; }();
    int stages = int((wg_budget - fixed) / stage_bytes);
    plan->stages = std::max(2, std::min(stages, kMaxStages));
    plan->smem_bytes = fixed + plan->stages * stage_bytes;
    if (encode_2d_bf16(&plan->tmX, d.X, d.Cin, d.M, size_t(d.ldx) * 2, 64, BLOCK_K)) return -1;
    if (encode_2d_bf16(&plan->tmZ, d.dZ, d.Cout, d.M, size_t(d.ldz) * 2, 64, BLOCK_K)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    return 0;
}

int wgrad_launch(const WgradPlan& pl, cudaStream_t stream) {
    const WgradDesc& d = pl.d;
    WgradKParams p;
    p.Cin = d.Cin; p.Cout = d.Cout; p.block_n = pl.block_n; p.boxes_b = pl.boxes_b;
    p.co_tiles = pl.co_tiles; p.ci_tiles = pl.ci_tiles; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split;
    p.k_blocks = pl.k_blocks; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
    p.convert_x = d.x_fp16 ? 1 : 0;
    if (pl.splits == 1) { p.out = d.dW; p.ld_out = d.lddw; p.split_stride = 0; }
    else { p.out = d.workspace; p.ld_out = d.Cout; p.split_stride = static_cast<long long>(d.Cin) * d.Cout; }
    const int grid = pl.ci_tiles * pl.co_tiles * pl.splits;
    AMS_LAUNCH((gemm_wgrad_kernel), grid, kGemmThreads, pl.smem_bytes, stream, pl.tmX, pl.tmZ, p);
    if (pl.splits > 1) {
        const long long n = static_cast<long long>(d.Cin) * d.Cout;
        const int wpi = pl.splits > 16 ? 1 : 0;
        const long long threads = ceil_div_ll(n, 4) * (wpi ? 32 : 1);
        AMS_LAUNCH((wgrad_reduce_kernel), int(ceil_div_ll(threads, 256)), 256, 0, stream, d.workspace, d.dW, d.Cin, d.Cout, d.lddw,
                   pl.splits, p.split_stride, wpi);
    }
    return 0;
}

}  // namespace ams
