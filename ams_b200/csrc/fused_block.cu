// Block-fused inverted-residual kernel for FROZEN inference (stride-1 blocks of the MobileNetV2 backbone):
//
//     y = BN3( project_1x1( relu6(BN2( depthwise_3x3( relu6(BN1( expand_1x1(x) )) ) )) ) ) [+ x]
//
// in ONE kernel: the 6C-wide expanded tensors -- 84 % of the layer-boundary traffic of a frame (336.8 MB vs
// 55.2 MB block-boundary) and, in the 33x65 stage, two of the three launches of every block -- never exist in HBM.
//
// Second design ("channels on the lanes").  The first one put the halo PIXELS on the UMMA M axis and walked the
// expanded channels in chunks of 32/64 columns: two M tiles per chunk, a shared-memory round trip between the expand
// epilogue and the depthwise threads, and 4x more (small-N) tcgen05.mma instructions than the tensor pipe retires in the
// time the CUDA cores need -- it measured SLOWER than one kernel per layer.  Here the expand GEMM is transposed:
//
//   tcgen05  : D1[128 expanded channels x halo pixels] = We[128 x Cin] * X^T      (M = channels, N = 180/240 halo pixels)
//              so that TMEM LANE = channel and TMEM COLUMN = halo pixel (row-major 18- or 20-wide halo rows).
//   CUDA core: a thread owns ONE channel (its TMEM lane): it reads its halo rows with tcgen05.ld, applies the folded
//              BN1 + ReLU6 with per-thread constants, zeroes positions outside the image (the depthwise conv pads its
//              INPUT), rounds to fp16, and runs the 3x3 depthwise conv entirely in registers on a sliding window of
//              2d+1 halo rows -- no shared-memory round trip, no per-element address arithmetic; then folded BN2 +
//              ReLU6 -> fp16 -> 16-byte stores into the MN-major (pixel-contiguous) swizzled A operand of the project GEMM.
//   tcgen05  : D2[128 pixels x Cout] += A2[128 px x 128 ch] * Wp^T (+ Wp_lo^T), accumulated over the channel chunks
//   epilogue : D2 -> folded BN3 (+ residual x) -> fp16 -> HBM
// A persistent CTA (four warpgroups, setmaxnreg per role) owns one output tile at a time: {TMA producer, MMA issuer} | 2 x 4
// compute warps (TMEM lane quarter x upper/lower half of the tile rows) | 4 epilogue warps, one tile behind.  Stride-1
// blocks (dilation 1 | 2): 8x16 output pixels, halo (8+2d) x (16+2d).  Stride-2 blocks: 2x16 output pixels from a 5x33
// input halo (32 of the project GEMM's 128 rows carry pixels; the tensor pipe is idle most of the time anyway).
// X tile: 4-D TMA with zero fill outside the image, K-major SWIZZLE_64B (32-channel k-blocks), two buffers where they fit;
// weights stream through two rings of 32-k units; D1 / A2 / D2 are double-buffered where TMEM / shared memory allow; where
// N <= 256 both planes of the split project weights are ONE stacked operand (D2 = two column halves, added in the epilogue).
// Every stored value goes through the same fp16 rounding points as the unfused path (expand output, depthwise output,
// block output), so the two paths agree up to the accumulation order.
//
// Replaces, for the frozen client graph (reference utils/graph_utils.py:52-126; SemanticNetwork.py:173): the nodes
// expanded_conv_k/{expand, depthwise, project} (+ BatchNorm, Relu6, add) of checkpoints/*/model.meta for every
// stride-1 block k.
#include "fused_block.cuh"
#include "tcgen05.cuh"

#include <cudaTypedefs.h>
#include <algorithm>

namespace ams {
namespace {

constexpr int kTH = 8, kTW = 16;              // output tile: 128 pixels = one UMMA M tile of the project GEMM
constexpr int kKB = 32;                       // k-block of every operand ring: 32 channels = 64-byte swizzled rows
constexpr int kChunk = 128;                   // expanded channels per chunk = UMMA M of the expand GEMM = TMEM lanes
constexpr int kThreads = 512;                 // four warpgroups: {TMA producer, MMA issuer + TMEM owner, 2 spare}, 2 x compute, epilogue
constexpr int kCompute = 256;                 // warps 4..11: TMEM lane quarter x upper/lower half of the tile rows
constexpr int kEpilogue = 128;                // warps 12..15: one TMEM lane quarter (32 output pixels) each
constexpr int kRegsLight = 64, kRegsEpilogue = 80, kRegsCompute = 176;     // setmaxnreg: 128*(64+80) + 256*176 <= 64K registers
constexpr int kMaxWe = 8, kMaxWp = 8;         // ring slots
constexpr uint32_t kSpinLimit = 1u << 28;     // a barrier that never completes traps instead of hanging the GPU

struct FusedParams {
    int N, H, W, Cin, Cexp, Cout, dil;
    int Ho, Wo, pad_top, pad_left;    // stride-2 blocks: output size and the depthwise conv's 'SAME' padding (stride 1: Ho = H, Wo = W, pad = dil)
    int k_blocks;                 // ceil(Cin / 32)
    int n_chunks;                 // ceil(Cexp / 128)
    int we_split, wp_split;       // low weight planes present
    int np_mma;                   // Cout rounded up to the UMMA N granule (16)
    int wp_stack;                 // split project weights as ONE operand of 2*np_mma rows (hi rows, then lo rows): half the MMAs, D2 = two column halves
    int tiles_x, tiles_y, num_tiles;
    int halo_w, halo_h, halo_rows, n1;        // n1 = halo_rows rounded up to 16 = UMMA N of the expand GEMM
    int d1_bufs, a2_bufs, we_slots, wp_slots, x_bufs, d2_bufs;
    int resident;                 // the whole block's weights stay in the rings (each ring = exactly one tile's units): loaded once per CTA
    int wp_group;                 // 32-channel k-blocks of project weights per ring unit / barrier (4 = a whole chunk, else 1)
    uint32_t tmem_cols, d2_col, d2_cols;       // first D2 column; columns per D2 buffer
    const float* par;             // [n_chunks][13][128]: s1 t1 wd[9] s2 t2, zero beyond Cexp
    const float* s3; const float* t3;           // [Cout]
    const __half* residual;                      // block input (same geometry) or null
    __half* out;
    uint32_t off_x, off_we, off_wp, off_a2, off_bar;       // shared-memory offsets (bytes, from the 1024-aligned base)
    uint32_t x_slab, x_buf, we_unit, we_plane, wp_unit, wp_plane, wp_kb;  // bytes: one k-block of X; one ring unit; one plane of a k-block; wp_kb = all planes of one k-block
    unsigned long long* dbg;                     // optional timeline of CTA 0 (tools/micro/fused_block_run.py): [role][chunk][4] ns
};
__device__ __forceinline__ void dbg_mark(const FusedParams& p, int role, long long g, int k) {
    if (p.dbg && blockIdx.x == 0 && g < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[(role * 64 + g) * 4 + k] = t;
    }
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(t5::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(t5::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void wait_bar_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!t5::mbar_try_wait(bar, parity)) { __nanosleep(32); if (++spins > (kSpinLimit >> 4)) __trap(); }
}
// long waits (the epilogue warps wait a whole tile for D2): a spinning warp takes issue slots from the two compute warps of
// its scheduler -- the spin loop was 15 % of all executed instructions in the first ncu capture of this kernel
__device__ __forceinline__ void wait_bar_idle(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!t5::mbar_try_wait(bar, parity)) { __nanosleep(256); if (++spins > (kSpinLimit >> 6)) __trap(); }
}
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {          // d += a * b (per lane), packed fp32x2
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 32 lanes x 32 bit, N consecutive columns (N = 16 + 2 or 16 + 4: one halo row): thread i gets TMEM lane base + i
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

// barrier slots inside the shared-memory barrier area
enum : int { B_XFULL = 0, B_XEMPTY = 2, B_WEFULL = 4, B_WEEMPTY = B_WEFULL + kMaxWe, B_WPFULL = B_WEEMPTY + kMaxWe, B_WPEMPTY = B_WPFULL + kMaxWp,
             B_D1FULL = B_WPEMPTY + kMaxWp, B_D1EMPTY = B_D1FULL + 2, B_A2FULL = B_D1EMPTY + 2, B_A2EMPTY = B_A2FULL + 2,
             B_D2FULL = B_A2EMPTY + 2, B_D2EMPTY = B_D2FULL + 2, B_COUNT = B_D2EMPTY + 2 };

// <D, S>: depthwise dilation (1 | 2, stride 1) and stride (1 | 2, dilation 1).  Stride 1: 8x16 output tile, halo (8+2D) x (16+2D).
// Stride 2: 2x16 output tile (32 of the 128 project-GEMM rows carry pixels), input halo 5 x 33.
template <int D, int S>
__global__ void __launch_bounds__(kThreads, 1)
fused_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                   const __grid_constant__ CUtensorMap tmWeLo, const __grid_constant__ CUtensorMap tmWp,
                   const __grid_constant__ CUtensorMap tmWpLo, const FusedParams p) {
    constexpr int TH = S == 1 ? kTH : 2;       // output tile rows
    constexpr int HW = S == 1 ? kTW + 2 * D : 2 * kTW + 1;      // halo row length = TMEM columns per halo row
    constexpr int WIN = 2 * D + 1;             // halo rows one output row needs
    constexpr int ROWS = TH / 2;               // output rows per compute warp
    constexpr int HR = S == 1 ? ROWS + 2 * D : 3;               // halo rows per compute warp
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smX = smem + p.off_x;        // [k_blocks][n1 rows][64 B]                 K-major, swizzle 64B (row = halo pixel)
    uint8_t* smWe = smem + p.off_we;      // [we_slots][planes][128 rows][64 B]        K-major, swizzle 64B (row = expanded channel)
    uint8_t* smWp = smem + p.off_wp;      // [wp_slots][planes][np_mma rows][64 B]     K-major, swizzle 64B (row = output channel)
    uint8_t* smA2 = smem + p.off_a2;      // [a2_bufs][2 atoms][128 k rows][128 B]     MN-major, swizzle 128B (64 pixels per row)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmX); t5::tma_prefetch_desc(&tmWe); t5::tma_prefetch_desc(&tmWp);
        for (int b = 0; b < B_COUNT; ++b) {
            const bool by_compute = (b >= B_D1EMPTY && b < B_D1EMPTY + 2) || (b >= B_A2FULL && b < B_A2FULL + 2);
            t5::mbar_init(&bars[b], by_compute ? kCompute : ((b >= B_D2EMPTY && b < B_D2EMPTY + 2) ? kEpilogue : 1));
        }
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, p.tmem_cols); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_d2 = tmem_base + p.d2_col;

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; chunks are numbered globally (g = tile_local * n_chunks + j)
    // so that the weight prefetch and the expand GEMMs run ahead ACROSS tile boundaries
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int G = my_tiles * p.n_chunks;

    // register budget per warpgroup (setmaxnreg at the top of each role's branch, so that ptxas allocates the branch against
    // it): the compute threads hold a (2d+1) x (16+2d) fp32 window + 16 accumulators
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegsLight));
    if (warp == 0) {
        // ================================================================== TMA producer: loads in the order the MMA issuer consumes them
        if (lane == 0) {
            // ring bookkeeping without divisions: slot index + number of completed passes over the ring
            int we_s = 0, wp_s = 0; uint32_t we_pass = 0, wp_pass = 0;
            int tl_e = 0, j_e = 0, j_p = 0;             // tile / chunk of the next expand-weight issue; chunk of the next project-weight issue
            int xb = 0; uint32_t x_pass = 0;            // X buffer of that tile and completed passes over the X buffers
            auto issue_we = [&]() {
                if (j_e == 0) {
                    int r = static_cast<int>(blockIdx.x) + tl_e * static_cast<int>(gridDim.x);
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y;
                    const int n = r / p.tiles_y;
                    if (x_pass > 0) wait_bar_relaxed(&bars[B_XEMPTY + xb], (x_pass - 1) & 1);   // every expand GEMM of the tile that used this buffer retired
                    t5::mbar_arrive_expect_tx(&bars[B_XFULL + xb], static_cast<uint32_t>(p.k_blocks * p.halo_rows * 64));
                    for (int kb = 0; kb < p.k_blocks; ++kb)
                        tma_load_4d(smX + xb * p.x_buf + kb * p.x_slab, &tmX, &bars[B_XFULL + xb], kb * kKB, tx * kTW * S - p.pad_left, ty * TH * S - p.pad_top, n);
                    if (++xb == p.x_bufs) { xb = 0; ++x_pass; }
                }
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    if (!(p.resident && we_pass > 0)) {                                  // resident weights: first tile only
                        if (we_pass > 0) wait_bar_relaxed(&bars[B_WEEMPTY + we_s], (we_pass - 1) & 1);
                        uint8_t* dst = smWe + we_s * p.we_unit;
                        t5::mbar_arrive_expect_tx(&bars[B_WEFULL + we_s], p.we_unit);
                        t5::tma_load_2d(dst, &tmWe, &bars[B_WEFULL + we_s], kb * kKB, j_e * kChunk);
                        if (p.we_split) t5::tma_load_2d(dst + p.we_plane, &tmWeLo, &bars[B_WEFULL + we_s], kb * kKB, j_e * kChunk);
                    }
                    if (++we_s == p.we_slots) { we_s = 0; ++we_pass; }
                }
                if (++j_e == p.n_chunks) { j_e = 0; ++tl_e; }
            };
            auto issue_wp = [&]() {
                const int kbs = min(kChunk / kKB, (p.Cexp - j_p * kChunk + kKB - 1) / kKB);      // k-blocks of this chunk that hold channels
                for (int u0 = 0; u0 < kbs; u0 += p.wp_group) {
                    const int nkb = min(p.wp_group, kbs - u0);
                    if (!(p.resident && wp_pass > 0)) {
                        if (wp_pass > 0) wait_bar_relaxed(&bars[B_WPEMPTY + wp_s], (wp_pass - 1) & 1);
                        uint8_t* dst = smWp + wp_s * p.wp_unit;
                        t5::mbar_arrive_expect_tx(&bars[B_WPFULL + wp_s], nkb * p.wp_kb);
                        for (int u = 0; u < nkb; ++u) {
                            t5::tma_load_2d(dst + u * p.wp_kb, &tmWp, &bars[B_WPFULL + wp_s], j_p * kChunk + (u0 + u) * kKB, 0);
                            if (p.wp_split) t5::tma_load_2d(dst + u * p.wp_kb + p.wp_plane, &tmWpLo, &bars[B_WPFULL + wp_s], j_p * kChunk + (u0 + u) * kKB, 0);
                        }
                    }
                    if (++wp_s == p.wp_slots) { wp_s = 0; ++wp_pass; }
                }
                if (++j_p == p.n_chunks) j_p = 0;
            };
            if (G > 0) issue_we();
            for (int g = 0; g < G; ++g) {
                if (g + 1 < G) issue_we();
                issue_wp();
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one elected lane issues)
        const uint32_t idesc1 = t5::make_idesc_f16(128, p.n1, 0, 0, 0, 0);            // A = We (K-major), B = X (K-major)
        const uint32_t idesc2 = t5::make_idesc_f16(128, p.wp_stack ? 2 * p.np_mma : p.np_mma, 1, 0, 0, 0);   // A = A2 (MN-major: pixels contiguous), B = Wp (K-major)
        int we_s = 0, wp_s = 0; uint32_t we_pass = 0, wp_pass = 0;
        // expand side: tile / chunk / D1 buffer / pass over the D1 buffers of the NEXT expand GEMM
        int tl_e = 0, j_e = 0, b_e = 0; uint32_t d1_pass = 0;
        int xb = 0; uint32_t x_pass = 0;                // X buffer of the tile being expanded
        auto expand = [&](int g) {
            if (j_e == 0) wait_bar_relaxed(&bars[B_XFULL + xb], x_pass & 1);
            if (d1_pass > 0) wait_bar_relaxed(&bars[B_D1EMPTY + b_e], (d1_pass - 1) & 1);          // the compute warps have read the previous chunk out of this buffer
            t5::fence_after_thread_sync();
            const uint32_t d = tmem_base + b_e * p.n1;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                if (!(p.resident && we_pass > 0)) wait_bar_relaxed(&bars[B_WEFULL + we_s], we_pass & 1);     // resident weights: arrived during the first tile
                t5::fence_after_thread_sync();
                if (lane == 0 && kb == 0) dbg_mark(p, 0, g, 0);
                {
                    // The whole (converged) warp runs this with warp-uniform operands and one elected lane issues each
                    // instruction: the issue path stays on the uniform datapath (no R2UR / BRA.U.ANY sequences), which
                    // matters because this warp shares its scheduler with two FFMA-bound compute warps.
                    // Descriptors: the start-address field is bits [0,14) in 16-byte units, so a k-step of 32 bytes is +2.
                    const uint64_t da0 = t5::make_smem_desc(t5::smem_u32(smWe) + we_s * p.we_unit, 16, 512, 4);
                    const uint64_t db0 = t5::make_smem_desc(t5::smem_u32(smX) + xb * p.x_buf + kb * p.x_slab, 16, 512, 4);
                    const uint64_t lo = p.we_plane >> 4;
#pragma unroll
                    for (int k = 0; k < kKB / 16; ++k) {
                        t5::mma_f16_ss_warp(d, da0 + 2 * k, db0 + 2 * k, idesc1, (kb | k) != 0);
                        if (p.we_split) t5::mma_f16_ss_warp(d, da0 + lo + 2 * k, db0 + 2 * k, idesc1, 1u);
                    }
                    if (!p.resident) t5::mma_commit_warp(&bars[B_WEEMPTY + we_s]);
                }
                if (lane == 0 && kb == p.k_blocks - 1) dbg_mark(p, 0, g, 2);
                if (++we_s == p.we_slots) { we_s = 0; ++we_pass; }
            }
            t5::mma_commit_warp(&bars[B_D1FULL + b_e]);
            if (j_e == p.n_chunks - 1) t5::mma_commit_warp(&bars[B_XEMPTY + xb]);
            if (++b_e == p.d1_bufs) { b_e = 0; ++d1_pass; }
            if (++j_e == p.n_chunks) { j_e = 0; ++tl_e; if (++xb == p.x_bufs) { xb = 0; ++x_pass; } }
        };
        if (G > 0) expand(0);
        int tl = 0, j = 0, a = 0; uint32_t a2_pass = 0;
        int db = 0; uint32_t d2_pass = 0;               // D2 buffer of the tile being projected
        for (int g = 0; g < G; ++g) {
            if (g + 1 < G) expand(g + 1);
            if (j == 0 && d2_pass > 0) wait_bar_relaxed(&bars[B_D2EMPTY + db], (d2_pass - 1) & 1);   // the epilogue of the tile that used this D2 buffer drained it
            const uint32_t tmem_d2b = tmem_d2 + db * p.d2_cols;
            wait_bar_relaxed(&bars[B_A2FULL + a], a2_pass & 1);
            t5::fence_after_thread_sync();
            const int kbs = min(kChunk / kKB, (p.Cexp - j * kChunk + kKB - 1) / kKB);
            for (int u0 = 0; u0 < kbs; u0 += p.wp_group) {
                const int nkb = min(p.wp_group, kbs - u0);
                if (!(p.resident && wp_pass > 0)) wait_bar_relaxed(&bars[B_WPFULL + wp_s], wp_pass & 1);
                t5::fence_after_thread_sync();
                if (lane == 0 && u0 == 0) dbg_mark(p, 0, g, 1);
                {
                    // A2 is MN-major SWIZZLE_128B: 64-pixel atoms 16 KB apart (LBO), 8-channel groups 1024 B apart (SBO); one
                    // UMMA_K = 16 channels = 2048 B (+128 in the 16-byte units of the descriptor's address field)
                    const uint64_t da0 = t5::make_smem_desc_sw128(t5::smem_u32(smA2) + a * (2 * 16384) + u0 * 4096, 16384, 1024);
                    const uint64_t db0 = t5::make_smem_desc(t5::smem_u32(smWp) + wp_s * p.wp_unit, 16, 512, 4);
                    const uint64_t lo = p.wp_plane >> 4, kbstep = p.wp_kb >> 4;
                    for (int u = 0; u < nkb; ++u) {
#pragma unroll
                        for (int k = 0; k < kKB / 16; ++k) {
                            t5::mma_f16_ss_warp(tmem_d2b, da0 + 256 * u + 128 * k, db0 + kbstep * u + 2 * k, idesc2, (j | (u0 + u) | k) != 0);
                            if (p.wp_split && !p.wp_stack) t5::mma_f16_ss_warp(tmem_d2b, da0 + 256 * u + 128 * k, db0 + kbstep * u + lo + 2 * k, idesc2, 1u);
                        }
                    }
                    if (!p.resident) t5::mma_commit_warp(&bars[B_WPEMPTY + wp_s]);
                }
                if (lane == 0 && u0 + nkb >= kbs) dbg_mark(p, 0, g, 3);
                if (++wp_s == p.wp_slots) { wp_s = 0; ++wp_pass; }
            }
            t5::mma_commit_warp(&bars[B_A2EMPTY + a]);
            if (j == p.n_chunks - 1) t5::mma_commit_warp(&bars[B_D2FULL + db]);
            if (++a == p.a2_bufs) { a = 0; ++a2_pass; }
            if (++j == p.n_chunks) { j = 0; ++tl; if (++db == p.d2_bufs) { db = 0; ++d2_pass; } }
        }
    }
    } else if (warp < 12) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegsCompute));
        // ================================================================== 256 compute threads: thread = one expanded channel x half of the tile rows
        const int q = warp & 3;                            // TMEM lane quarter of this warp
        const int hh = (warp - 4) >> 2;                    // 0: output rows 0..3, 1: rows 4..7
        const int ch = q * 32 + lane;                      // channel within the chunk = TMEM lane = K row of A2
        const uint32_t a2_s = t5::smem_u32(smA2);
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        int g = 0, b = 0, a = 0; uint32_t d1_pass = 0, a2_pass = 0;
        for (int t = blockIdx.x, tl = 0; t < p.num_tiles; t += gridDim.x, ++tl) {
            int r = t;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y;
            const int y0 = ty * TH * S - p.pad_top, x0 = tx * kTW * S - p.pad_left;     // image coordinates of halo position (0, 0)
            const int col_hi = p.W - x0;                    // halo columns >= col_hi lie right of the image
            for (int j = 0; j < p.n_chunks; ++j, ++g) {
                // per-channel constants of this chunk (coalesced over the lanes; zero beyond Cexp)
                const float* pp = p.par + static_cast<long long>(j) * 13 * kChunk + ch;
                const float s1 = __ldg(pp), t1 = __ldg(pp + kChunk);
                float wd[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) wd[k] = __ldg(pp + (2 + k) * kChunk);
                const float s2 = __ldg(pp + 11 * kChunk), t2 = __ldg(pp + 12 * kChunk);
                if (threadIdx.x == 128) dbg_mark(p, 1, g, 0);
                wait_bar_relaxed(&bars[B_D1FULL + b], d1_pass & 1);
                t5::fence_after_thread_sync();
                if (threadIdx.x == 128) dbg_mark(p, 1, g, 1);
                if (threadIdx.x == 128) dbg_mark(p, 1, g, 3);
                const uint32_t taddr = tmem_base + lane_addr + b * p.n1 + (hh * ROWS * S) * HW;
                // K row `ch` of both 64-pixel atoms; 16-byte chunk index XOR (ch mod 8)
                const uint32_t a2_row = a2_s + a * (2 * 16384) + ch * 128;
                if constexpr (S == 1) {
                // window rows as aligned fp32 pairs (columns 2i, 2i+1): taps with an even column offset (all of them at
                // dilation 2; kx = 0, 2 at dilation 1) run as packed fp32x2 FMAs on two adjacent output pixels
                float2 win[WIN][HW / 2];
                uint32_t raw[HW];
                const float2 s1p = make_float2(s1, s1), t1p = make_float2(t1, t1), s2p = make_float2(s2, s2), t2p = make_float2(t2, t2);
                float2 wp2[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) wp2[k] = make_float2(wd[k], wd[k]);
                t5::tmem_ld16(taddr, raw);
                if (D == 1) tmem_ld2(taddr + 16, raw + 16); else tmem_ld4(taddr + 16, raw + 16);
#pragma unroll
                for (int hr = 0; hr < HR; ++hr) {
                    t5::tmem_ld_wait();
                    const int gy = y0 + hh * ROWS + hr;
                    float2* wr = win[hr % WIN];
                    if (gy >= 0 && gy < p.H) {
#pragma unroll
                        for (int i = 0; i < HW / 2; ++i) {
                            float2 v = t1p;
                            ffma2(v, make_float2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1])), s1p);
                            v.x = fminf(fmaxf(v.x, 0.f), 6.f); v.y = fminf(fmaxf(v.y, 0.f), 6.f);
                            wr[i] = h16x2(pack_h16(v.x, v.y));                          // the value the unfused path stores (fp16)
                        }
                        if (x0 < 0) {                                                   // columns left of the image (x0 = -D)
                            if (D == 1) wr[0].x = 0.f; else wr[0] = make_float2(0.f, 0.f);
                        }
                        if (col_hi < HW) {                                              // columns right of the image (last tile column only)
#pragma unroll
                            for (int i = 0; i < HW / 2; ++i) { if (2 * i >= col_hi) wr[i].x = 0.f; if (2 * i + 1 >= col_hi) wr[i].y = 0.f; }
                        }
                    } else {                                                            // row above / below the image: the depthwise conv's zero padding
#pragma unroll
                        for (int i = 0; i < HW / 2; ++i) wr[i] = make_float2(0.f, 0.f);
                    }
                    if (hr + 1 < HR) {                                                  // next halo row: in flight during the FMAs below
                        t5::tmem_ld16(taddr + (hr + 1) * HW, raw);
                        if (D == 1) tmem_ld2(taddr + (hr + 1) * HW + 16, raw + 16); else tmem_ld4(taddr + (hr + 1) * HW + 16, raw + 16);
                    } else {
                        t5::fence_before_thread_sync();
                        t5::mbar_arrive(&bars[B_D1EMPTY + b]);                         // this thread has read its part of D1
                    }
                    if (hr >= 2 * D) {
                        const int orow = hr - 2 * D;                                    // output row within this warp's half
                        float2 acc[kTW / 2];
#pragma unroll
                        for (int i = 0; i < kTW / 2; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const float2* xr = win[(orow + ky * D) % WIN];
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int o = kx * D;
                                if ((o & 1) == 0) {
#pragma unroll
                                    for (int i = 0; i < kTW / 2; ++i) ffma2(acc[i], xr[i + o / 2], wp2[ky * 3 + kx]);
                                } else {                                                // odd offset: (x[2i+o], x[2i+o+1]) straddles two pairs
                                    const float w = wd[ky * 3 + kx];
#pragma unroll
                                    for (int i = 0; i < kTW / 2; ++i) {
                                        acc[i].x = fmaf(xr[i + (o - 1) / 2].y, w, acc[i].x);
                                        acc[i].y = fmaf(xr[i + (o + 1) / 2].x, w, acc[i].y);
                                    }
                                }
                            }
                        }
                        float o[kTW];
#pragma unroll
                        for (int i = 0; i < kTW / 2; ++i) {
                            float2 v = t2p;
                            ffma2(v, acc[i], s2p);
                            o[2 * i] = fminf(fmaxf(v.x, 0.f), 6.f); o[2 * i + 1] = fminf(fmaxf(v.y, 0.f), 6.f);
                        }
                        // output pixel m = (hh*4 + orow)*16 + c: atom m / 64, 16-byte chunk (m % 64) / 8
                        const int m0 = (hh * ROWS + orow) * kTW;
                        // first store of the chunk: the project GEMM that last read this A2 buffer must have retired (waited for
                        // HERE, after the BN1 work of 2d+1 halo rows, so that its latency hides behind that work)
                        if (orow == 0 && a2_pass > 0) wait_bar_relaxed(&bars[B_A2EMPTY + a], (a2_pass - 1) & 1);
                        const uint32_t base = a2_row + (m0 >> 6) * 16384;
                        const int ck = (m0 & 63) >> 3;
                        sts128(base + (((ck) ^ (ch & 7)) << 4), pack8h(o));
                        sts128(base + (((ck + 1) ^ (ch & 7)) << 4), pack8h(o + 8));
                    }
                }
                } else {
                // stride 2: this warp computes ONE output row (16 pixels) from 3 halo rows x 33 columns held as scalars
                float win[3][HW];
                uint32_t raw[HW];
                auto load_row = [&](int hr) {
                    t5::tmem_ld16(taddr + hr * HW, raw);
                    t5::tmem_ld16(taddr + hr * HW + 16, raw + 16);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(raw[32]) : "r"(taddr + hr * HW + 32) : "memory");
                };
                load_row(0);
#pragma unroll
                for (int hr = 0; hr < HR; ++hr) {
                    t5::tmem_ld_wait();
                    const int gy = y0 + hh * ROWS * 2 + hr;
                    float* wr = win[hr];
                    if (gy >= 0 && gy < p.H) {
#pragma unroll
                        for (int i = 0; i < HW / 2; ++i) {
                            const float v0 = fminf(fmaxf(fmaf(__uint_as_float(raw[2 * i]), s1, t1), 0.f), 6.f);
                            const float v1 = fminf(fmaxf(fmaf(__uint_as_float(raw[2 * i + 1]), s1, t1), 0.f), 6.f);
                            const float2 f = h16x2(pack_h16(v0, v1));                  // the value the unfused path stores (fp16)
                            wr[2 * i] = f.x; wr[2 * i + 1] = f.y;
                        }
                        {
                            const float v0 = fminf(fmaxf(fmaf(__uint_as_float(raw[HW - 1]), s1, t1), 0.f), 6.f);
                            wr[HW - 1] = h16x2(pack_h16(v0, 0.f)).x;
                        }
                        if (x0 < 0) wr[0] = 0.f;                                        // column left of the image (pad_left = 1)
                        if (col_hi < HW) {                                              // columns right of the image
#pragma unroll
                            for (int c = 0; c < HW; ++c) if (c >= col_hi) wr[c] = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < HW; ++c) wr[c] = 0.f;
                    }
                    if (hr + 1 < HR) {
                        load_row(hr + 1);
                    } else {
                        t5::fence_before_thread_sync();
                        t5::mbar_arrive(&bars[B_D1EMPTY + b]);                         // this thread has read its part of D1
                    }
                }
                {
                    float acc[kTW];
#pragma unroll
                    for (int c = 0; c < kTW; ++c) acc[c] = 0.f;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int c = 0; c < kTW; ++c) acc[c] = fmaf(win[ky][2 * c + kx], wd[ky * 3 + kx], acc[c]);
                    }
                    float o[kTW];
#pragma unroll
                    for (int c = 0; c < kTW; ++c) o[c] = fminf(fmaxf(fmaf(acc[c], s2, t2), 0.f), 6.f);
                    // output pixel m = hh*16 + c (rows 32..127 of the project GEMM's M tile carry no pixels)
                    if (a2_pass > 0) wait_bar_relaxed(&bars[B_A2EMPTY + a], (a2_pass - 1) & 1);
                    const int ck = (hh * kTW) >> 3;
                    sts128(a2_row + (((ck) ^ (ch & 7)) << 4), pack8h(o));
                    sts128(a2_row + (((ck + 1) ^ (ch & 7)) << 4), pack8h(o + 8));
                }
                }
                t5::fence_proxy_async_smem();                   // generic-proxy writes of A2 -> visible to the tensor core
                t5::mbar_arrive(&bars[B_A2FULL + a]);
                if (threadIdx.x == 128) dbg_mark(p, 1, g, 2);
                if (++b == p.d1_bufs) { b = 0; ++d1_pass; }
                if (++a == p.a2_bufs) { a = 0; ++a2_pass; }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegsEpilogue));
        // ================================================================== epilogue warpgroup: D2 -> BN3 (+ x) -> fp16 -> HBM, one tile behind
        // the compute warps (which go straight on to the next tile: its first expand GEMM has already run)
        const int q = warp & 3;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int prow = q * 32 + lane;
        const int ncols = p.Cout;
        int db = 0; uint32_t d2_pass = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            int r = t;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y;
            const int n = r / p.tiles_y;
            const int oy = ty * TH + prow / kTW, ox = tx * kTW + prow % kTW;
            const bool ok = prow < TH * kTW && oy < p.Ho && ox < p.Wo;
            const long long pix = (static_cast<long long>(n) * p.Ho + oy) * p.Wo + ox;
            // the skip connection's first 32 channels are in flight while the last project GEMM of the tile retires
            uint4 res[4];
            if (p.residual && ok) {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k * 8 < ncols) res[k] = ldg_stream(p.residual + pix * p.Cout + k * 8);
            }
            wait_bar_idle(&bars[B_D2FULL + db], d2_pass & 1);
            t5::fence_after_thread_sync();
            const uint32_t taddr = tmem_d2 + db * p.d2_cols + lane_addr;
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t ra[16], rb[16];
                t5::tmem_ld16(taddr + c0, ra);
                if (c0 + 16 < ncols) t5::tmem_ld16(taddr + c0 + 16, rb);
                if (p.wp_stack) {
                    // hi-plane and lo-plane products were accumulated in two column halves of D2: add them here
                    uint32_t la[16], lb[16];
                    t5::tmem_ld16(taddr + p.np_mma + c0, la);
                    if (c0 + 16 < ncols) t5::tmem_ld16(taddr + p.np_mma + c0 + 16, lb);
                    t5::tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        ra[k] = __float_as_uint(__uint_as_float(ra[k]) + __uint_as_float(la[k]));
                        if (c0 + 16 < ncols) rb[k] = __float_as_uint(__uint_as_float(rb[k]) + __uint_as_float(lb[k]));
                    }
                }
                if (c0 > 0 && p.residual && ok) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (c0 + k * 8 < ncols) res[k] = ldg_stream(p.residual + pix * p.Cout + c0 + k * 8);
                }
                t5::tmem_ld_wait();
                if (c0 + 32 >= ncols) {                                      // last read of D2: the next tile's project GEMMs may start
                    t5::fence_before_thread_sync();
                    t5::mbar_arrive(&bars[B_D2EMPTY + db]);
                }
                if (!ok) continue;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int cb = c0 + h2 * 16;
                    if (cb >= ncols) break;
                    const uint32_t* rr = h2 ? rb : ra;
                    float v[16];
#pragma unroll
                    for (int k = 0; k < 16; k += 4) {
                        if (cb + k < ncols) {                                 // Cout is a multiple of 8: groups of 4 are all-in or all-out
                            const float4 s3 = __ldg(reinterpret_cast<const float4*>(p.s3 + cb + k)), t3 = __ldg(reinterpret_cast<const float4*>(p.t3 + cb + k));
                            v[k] = fmaf(__uint_as_float(rr[k]), s3.x, t3.x); v[k + 1] = fmaf(__uint_as_float(rr[k + 1]), s3.y, t3.y);
                            v[k + 2] = fmaf(__uint_as_float(rr[k + 2]), s3.z, t3.z); v[k + 3] = fmaf(__uint_as_float(rr[k + 3]), s3.w, t3.w);
                        } else { v[k] = v[k + 1] = v[k + 2] = v[k + 3] = 0.f; }
                    }
                    if (p.residual) {
                        float f[8];
                        unpack8h(res[2 * h2], f);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] += f[k];
                        if (cb + 8 < ncols) {
                            unpack8h(res[2 * h2 + 1], f);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[8 + k] += f[k];
                        }
                    }
                    __half* o = p.out + pix * p.Cout + cb;
                    stg_stream(o, pack8h(v));
                    if (cb + 8 < ncols) stg_stream(o + 8, pack8h(v + 8));
                }
            }
            if (++db == p.d2_bufs) { db = 0; ++d2_pass; }
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        t5::fence_after_thread_sync();
        t5::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
    }
    return fn;
}

int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
           int swizzle_bytes) {
    auto fn = encode_fn();
    AMS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available");
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return 0;
}

// geometry-only part of the plan (shared by fused_block_supported / _param_floats / _plan)
struct FusedGeom { int halo_w, halo_h, halo_rows, n1, np_mma, k_blocks, n_chunks, d1_bufs, wp_stack, d2_bufs; bool ok; };
FusedGeom fused_geom(const FusedBlockDesc& d) {
    FusedGeom g{};
    if (d.stride == 2) { g.halo_w = 2 * kTW + 1; g.halo_h = 5; }                   // 2x16 output tile: input rows 2r..2r+2, columns 2c..2c+2
    else { g.halo_w = kTW + 2 * d.dil; g.halo_h = kTH + 2 * d.dil; }
    g.halo_rows = g.halo_w * g.halo_h;
    g.n1 = ceil_div(g.halo_rows, 16) * 16;
    g.np_mma = ceil_div(d.Cout, 16) * 16;
    g.k_blocks = ceil_div(d.Cin, kKB);
    g.n_chunks = ceil_div(d.Cexp, kChunk);
    g.d1_bufs = (2 * g.n1 + g.np_mma <= 512) ? 2 : 1;
    g.ok = g.n1 <= 256 && g.np_mma <= 256 && g.n1 + g.np_mma <= 512;
    // D2: (a) both planes of the split project weights as one 2*np-row operand (half the project MMAs, D2 = two column halves
    // added in the epilogue) and (b) two D2 buffers, so that the epilogue of a tile overlaps the project GEMMs of the next one
    // -- what bounds tiles of one or two chunks.  Whatever fits in the 512 columns next to the D1 stages; (b) first for short tiles.
    const int room = 512 - g.d1_bufs * g.n1;
    const bool short_tiles = g.n_chunks <= 2;
    g.wp_stack = 0; g.d2_bufs = 1;
    if (2 * g.np_mma <= 256 && 4 * g.np_mma <= room) { g.wp_stack = 1; g.d2_bufs = 2; }
    else if (short_tiles && 2 * g.np_mma <= room) { g.d2_bufs = 2; }
    else if (2 * g.np_mma <= 256 && 2 * g.np_mma <= room) { g.wp_stack = 1; }
    else if (2 * g.np_mma <= room) { g.d2_bufs = 2; }
    return g;
}

}  // namespace

bool fused_block_supported(const FusedBlockDesc& d) {
    if (d.stride == 2) { if (d.dil != 1 || d.residual) return false; }
    else if (d.stride != 1 || (d.dil != 1 && d.dil != 2)) return false;
    if (d.Cin % 8 || d.Cexp % 8 || d.Cout % 8) return false;
    return fused_geom(d).ok;
}

size_t fused_block_param_floats(const FusedBlockDesc& d) {
    return static_cast<size_t>(13) * kChunk * ceil_div(d.Cexp, kChunk);
}

int fused_block_plan(const FusedBlockDesc& d, int num_sms, FusedBlockPlan* plan) {
    AMS_REQUIRE(fused_block_supported(d), "fused block: unsupported geometry");
    plan->d = d;
    FusedParams& p = *reinterpret_cast<FusedParams*>(plan->params);
    static_assert(sizeof(FusedParams) <= sizeof(plan->params), "FusedBlockPlan::params too small");
    p = FusedParams{};
    const FusedGeom g = fused_geom(d);
    p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cexp = d.Cexp; p.Cout = d.Cout; p.dil = d.dil;
    p.k_blocks = g.k_blocks; p.n_chunks = g.n_chunks;
    p.we_split = d.We_lo ? 1 : 0; p.wp_split = d.Wp_lo ? 1 : 0;
    p.np_mma = g.np_mma;
    p.wp_stack = (g.wp_stack && d.Wp_lo) ? 1 : 0;
    p.d2_cols = (p.wp_stack ? 2 : 1) * p.np_mma;
    p.d2_bufs = (g.d2_bufs == 2 || (!p.wp_stack && g.d1_bufs * g.n1 + 2 * g.np_mma <= 512)) ? 2 : 1;
    if (d.stride == 2) {
        p.Ho = ceil_div(d.H, 2); p.Wo = ceil_div(d.W, 2);
        // TensorFlow 'SAME': total padding max((out-1)*2 + 3 - in, 0), the smaller half first
        p.pad_top = d.pad_top >= 0 ? d.pad_top : std::max((p.Ho - 1) * 2 + 3 - d.H, 0) / 2;
        p.pad_left = d.pad_left >= 0 ? d.pad_left : std::max((p.Wo - 1) * 2 + 3 - d.W, 0) / 2;
    } else { p.Ho = d.H; p.Wo = d.W; p.pad_top = p.pad_left = d.dil; }
    p.tiles_x = ceil_div(p.Wo, kTW); p.tiles_y = ceil_div(p.Ho, d.stride == 2 ? 2 : kTH); p.num_tiles = d.N * p.tiles_x * p.tiles_y;
    p.halo_w = g.halo_w; p.halo_h = g.halo_h; p.halo_rows = g.halo_rows; p.n1 = g.n1;
    p.s3 = d.s3; p.t3 = d.t3;
    p.residual = static_cast<const __half*>(d.residual); p.out = static_cast<__half*>(d.out);
    p.dbg = static_cast<unsigned long long*>(d.debug_timeline);
    p.par = d.params;
    p.d1_bufs = g.d1_bufs;
    p.d2_col = p.d1_bufs * p.n1;
    p.tmem_cols = 512;
    p.x_slab = p.n1 * 64;
    p.we_plane = kChunk * 64; p.we_unit = p.we_plane * (p.we_split ? 2 : 1);
    p.wp_plane = p.np_mma * 64; p.wp_kb = p.wp_plane * (p.wp_split ? 2 : 1);
    // Rings as deep as shared memory allows.  Priorities: (1) the expand ring holds at least one whole chunk (k_blocks units) --
    // otherwise every expand GEMM stalls on a TMA round trip in its middle; (2) two A2 buffers, so that the project GEMM of a
    // chunk is off the compute warps' critical path; (3) a whole chunk of project weights in flight; (4) deeper rings.
    // Project weights come as whole chunks (4 k-blocks behind ONE barrier) where two such units fit: the MMA-issuing warp
    // shares its scheduler with two compute warps and every wait / commit / descriptor build costs it ~0.1 us.
    bool fits = false;
    int best_score = -1;
    p.x_buf = (p.k_blocks * p.x_slab + 1023u) & ~1023u;
    static const int grp_hi = [] { const char* e = getenv("AMS_FUSED_WP_GROUP"); return (e && e[0] == '4') ? 4 : 1; }();
    for (int grp = grp_hi; grp >= 1; grp -= 3)
    for (int xbf = 1; xbf <= 2; ++xbf)
    for (int a2 = 1; a2 <= 2; ++a2)
        for (int we = 2; we <= kMaxWe; ++we)
            for (int wp = 2; wp <= kMaxWp; ++wp) {
                if (grp == 4 && wp > 3) continue;
                const size_t total = 1024 + size_t(xbf) * p.x_buf + size_t(we) * p.we_unit + size_t(wp) * grp * p.wp_kb + size_t(a2) * 2 * 16384 +
                                     4096 /* alignment slack + barriers */;
                if (total > 227 * 1024) continue;
                // a second X buffer (the next tile's input lands while this tile computes) matters when a tile has few chunks
                const int score = (we >= p.k_blocks ? 1000 : 0) + (a2 == 2 ? 400 : 0) + (wp * grp >= 4 ? 200 : 0) + (xbf == 2 ? (p.n_chunks <= 2 ? 300 : 100) : 0) +
                                  (grp == 4 ? 150 : 0) + std::min(we, 2 * p.k_blocks) * 8 + std::min(wp * grp, 8) * 4;
                if (score > best_score) { best_score = score; p.we_slots = we; p.wp_slots = wp; p.a2_bufs = a2; p.x_bufs = xbf; p.wp_group = grp; fits = true; }
            }
    p.resident = 0;
    {
        // short blocks: ALL weights of the block stay resident (each ring = exactly the units of one tile, so the slot of a unit
        // is the same in every tile): no weight TMA, no FULL polls and no EMPTY commits after the first tile -- the barrier
        // polls are what the MMA-issuing warp spends most of its time on for tiles of one or two chunks
        static const bool off = [] { const char* e = getenv("AMS_FUSED_NO_RESIDENT"); return e && e[0] == '1'; }();
        const int we_all = p.n_chunks * p.k_blocks, wp_all = ceil_div(d.Cexp, kKB);
        if (!off && p.n_chunks <= 2 && we_all <= kMaxWe && wp_all <= kMaxWp) {
            for (int xbf = 2; xbf >= 1 && !p.resident; --xbf)
                for (int a2 = 2; a2 >= 1 && !p.resident; --a2) {
                    const size_t total = 1024 + size_t(xbf) * p.x_buf + size_t(we_all) * p.we_unit + size_t(wp_all) * p.wp_kb + size_t(a2) * 2 * 16384 + 4096;
                    if (total <= 227 * 1024) { p.resident = 1; p.we_slots = we_all; p.wp_slots = wp_all; p.a2_bufs = a2; p.x_bufs = xbf; p.wp_group = 1; fits = true; }
                }
        }
    }
    p.wp_unit = p.wp_group * p.wp_kb;
    if (fits) {
        uint32_t off = 0;
        auto take = [&](uint32_t bytes) { const uint32_t at = off; off = (off + bytes + 1023u) & ~1023u; return at; };
        p.off_x = take(p.x_bufs * p.x_buf);
        p.off_we = take(p.we_unit * p.we_slots);
        p.off_wp = take(p.wp_unit * p.wp_slots);
        p.off_a2 = take(p.a2_bufs * 2 * 16384);
        p.off_bar = take(B_COUNT * 8 + 16);
        plan->smem_bytes = std::max<size_t>(off + 1024, 116 * 1024);      // > half an SM: one CTA per SM (each allocates all 512 TMEM columns)
        fits = plan->smem_bytes <= 227 * 1024;
    }
    AMS_REQUIRE(fits, "fused block: shared memory overflow (" + std::to_string(plan->smem_bytes) + " bytes)");
    AMS_REQUIRE((p.x_slab % 512) == 0 && (p.we_plane % 512) == 0 && (p.wp_plane % 512) == 0, "fused block: operand tile alignment");
    plan->grid = std::min(p.num_tiles, num_sms);
    // tensor maps
    {
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cin), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.N)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.Cin) * 2, static_cast<cuuint64_t>(d.W) * d.Cin * 2,
                                 static_cast<cuuint64_t>(d.H) * d.W * d.Cin * 2};
        cuuint32_t box[4] = {kKB, static_cast<cuuint32_t>(p.halo_w), static_cast<cuuint32_t>(p.halo_h), 1};
        if (encode(&plan->tmX, d.x, 4, dims, strides, box, 64)) return -1;
    }
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Cin), static_cast<cuuint64_t>(d.Cexp)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.ld_we) * 2};
        cuuint32_t box[2] = {kKB, kChunk};
        if (encode(&plan->tmWe, d.We, 2, dims, strides, box, 64)) return -1;
        if (encode(&plan->tmWeLo, d.We_lo ? d.We_lo : d.We, 2, dims, strides, box, 64)) return -1;
    }
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Cexp), static_cast<cuuint64_t>(d.Cout)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.ld_wp) * 2};
        cuuint32_t box[2] = {kKB, static_cast<cuuint32_t>(p.np_mma)};
        if (encode(&plan->tmWp, d.Wp, 2, dims, strides, box, 64)) return -1;
        if (encode(&plan->tmWpLo, d.Wp_lo ? d.Wp_lo : d.Wp, 2, dims, strides, box, 64)) return -1;
    }
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<1, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<2, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<1, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    return 0;
}

int fused_block_launch(const FusedBlockPlan& plan, cudaStream_t s) {
    const FusedParams& p = *reinterpret_cast<const FusedParams*>(plan.params);
    if (plan.d.stride == 2)
        AMS_LAUNCH((fused_block_kernel<1, 2>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    else if (p.dil == 1)
        AMS_LAUNCH((fused_block_kernel<1, 1>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    else
        AMS_LAUNCH((fused_block_kernel<2, 1>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    return 0;
}

// params[n_chunks][13][128]: s1 t1 (folded BN of the expand conv), wd[9] (depthwise filter, [3,3,C] fp32), s2 t2 (folded BN
// of the depthwise conv), chunk-major so that a warp's 32 channels read 13 coalesced lines; channels >= Cexp are zero so
// that a padded chunk contributes exactly nothing.
__global__ void fused_fill_params_kernel(const float* __restrict__ s1, const float* __restrict__ t1, const float* __restrict__ wd,
                                         const float* __restrict__ s2, const float* __restrict__ t2, int C, int cpad, int chunk,
                                         float* __restrict__ out) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 13 * cpad) return;
    // out index = ((chunk index * 13) + which) * chunk + channel within the chunk
    const int j = i / (13 * chunk), rem = i - j * (13 * chunk);
    const int which = rem / chunk, c = j * chunk + (rem - which * chunk);
    float v = 0.f;
    if (c < C) {
        if (which == 0) v = s1[c];
        else if (which == 1) v = t1[c];
        else if (which < 11) v = wd[(which - 2) * C + c];
        else if (which == 11) v = s2[c];
        else v = t2[c];
    }
    out[i] = v;
}

int fused_block_fill_params(const FusedBlockDesc& d, const float* s1, const float* t1, const float* wd, const float* s2, const float* t2,
                            cudaStream_t s) {
    const int cpad = kChunk * ceil_div(d.Cexp, kChunk);
    AMS_LAUNCH((fused_fill_params_kernel), ceil_div(13 * cpad, 256), 256, 0, s, s1, t1, wd, s2, t2, d.Cexp, cpad, kChunk, const_cast<float*>(d.params));
    return 0;
}

}  // namespace ams
