// Block-fused inverted-residual kernel for FROZEN inference (stride-1 blocks of the MobileNetV2 backbone):
//
//     y = BN3( project_1x1( relu6(BN2( depthwise_3x3( relu6(BN1( expand_1x1(x) )) ) )) ) ) [+ x]
//
// in ONE kernel: the 6C-wide expanded tensors -- 84 % of the layer-boundary traffic of a frame (336.8 MB vs
// 55.2 MB block-boundary) and, in the 33x65 stage, two of the three launches of every block -- never exist in HBM.
//
// A persistent CTA owns one 8x16-pixel OUTPUT tile at a time (128 pixels = one UMMA M tile) and walks the expanded
// channels in chunks of 32 or 64:
//   TMA      : the fp16 input tile with its depthwise halo ((8+2d) x (16+2d) pixels x Cin, zero-filled outside the image)
//              lands once per tile as a K-major SWIZZLE_64B operand (32-channel k-blocks); per chunk the expand weights
//              [chunk x Cin] and the project weights [Cout x chunk] (hi and, where the layer uses split weights, lo plane)
//   tcgen05  : D1[halo pixels x chunk] = X * We^T              (two M tiles, fp32 accumulators in TMEM, double-buffered)
//   epilogue : 8 warps read D1 from TMEM, apply the folded BN1 + ReLU6, force pixels outside the image to zero (the
//              depthwise conv zero-pads its INPUT, i.e. the activated expand output), round to fp16 and write the halo
//              tile [pixel][chunk] to shared memory
//   CUDA core: depthwise 3x3 (dilation d) on that tile with packed fp32x2 FMAs, folded BN2 + ReLU6, fp16 rounding, written
//              straight into the K-major swizzled layout the tensor core reads as the A operand of the project GEMM
//   tcgen05  : D2[128 x Cout] += A2 * Wp^T (+ A2 * Wp_lo^T)    (accumulated over the chunks in TMEM)
//   epilogue : D2 -> folded BN3 (+ residual x) -> fp16 -> HBM
// Every stored value goes through the same fp16 rounding points as the unfused path (expand output, depthwise output,
// block output), so the two paths agree up to the accumulation order of the tensor core.
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

constexpr int kTH = 8, kTW = 16;              // output tile: 128 pixels = one UMMA M tile
constexpr int kKB = 32;                       // expand k-block: 32 input channels = 64-byte swizzled rows
constexpr int kThreads = 576;                 // warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 group E, warps 10..17 group W
constexpr int kGroup = 256;                   // threads of group E: TMEM -> BN1/ReLU6 -> halo tile (+ the final epilogue)
constexpr int kGroupW = 256;                  // threads of group W: depthwise -> A2 (16 warps with 2-pixel strips measured 8 % SLOWER).
                                              // The groups work on DIFFERENT chunks at the same time (double-buffered halo tile), so the
                                              // TMEM / shared-memory latencies of one overlap the FFMA work of the other.
constexpr int kXRows = 256;                   // halo pixels padded to two M tiles
constexpr uint32_t kSpinLimit = 1u << 28;     // a barrier that never completes traps instead of hanging the GPU

struct FusedParams {
    int N, H, W, Cin, Cexp, Cout, dil;
    int k_blocks;                 // ceil(Cin / 32)
    int chunk, n_chunks;          // expanded channels per chunk (32 | 64)
    int we_split, wp_split;       // low weight planes present
    int np_tiles, np, np_mma;     // project N tiles (Cout > 256: 2 x Cout/2); np_mma = np rounded up to the UMMA granule (16)
    int tiles_x, tiles_y, num_tiles;
    int halo_w, halo_h, halo_rows;
    int dw_pitch;                 // bytes per pixel row of the depthwise input tile (chunk*2 + 16)
    uint32_t tmem_cols, d1_cols;  // d1_cols = 2 * chunk (two M tiles) per stage
    // per-channel vectors, chunk-major [n_chunks][13][chunk] (s1 t1 wd[9] s2 t2), zero beyond Cexp: one bulk copy per chunk
    const float* par;
    const float* s3; const float* t3;           // [Cout]
    int cpad;
    const __half* residual;                      // block input (same geometry) or null
    __half* out;
    // shared-memory offsets (bytes, from the 1024-aligned base)
    uint32_t off_x, off_we, off_wp, off_dw, off_a2, off_par, off_bar;
    uint32_t we_tile, wp_tile, a2_tile;          // bytes of one plane / one A2 buffer
    int we_stages, wp_stages;                    // weight rings: the TMA round trip (~1.7 us) + the retire latency of the GEMM that
    uint32_t we_stage, wp_stage;                 // frees a slot (~0.9 us) exceed a chunk's compute time, so the loads run 1-2 chunks ahead
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
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!t5::mbar_try_wait(bar, parity)) { if (++spins > kSpinLimit) __trap(); }
}
__device__ __forceinline__ void wait_bar_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!t5::mbar_try_wait(bar, parity)) { __nanosleep(32); if (++spins > (kSpinLimit >> 4)) __trap(); }
}
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {          // explicit ld.shared: a generic-pointer load costs a long-scoreboard round trip
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts64(uint32_t a, const uint2& v) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// barrier slots inside the shared-memory barrier area
enum : int { B_XFULL = 0, B_XEMPTY, B_WEFULL0, B_WEFULL1, B_WEFULL2, B_WEEMPTY0, B_WEEMPTY1, B_WEEMPTY2, B_WPFULL0, B_WPFULL1, B_WPEMPTY0, B_WPEMPTY1,
             B_D1FULL0, B_D1FULL1, B_D1EMPTY0, B_D1EMPTY1, B_A2FULL0, B_A2FULL1, B_A2EMPTY0, B_A2EMPTY1, B_D2FULL, B_D2EMPTY,
             B_PARFULL0, B_PARFULL1, B_DWFULL0, B_DWFULL1, B_DWEMPTY0, B_DWEMPTY1, B_COUNT };

// 1-D bulk copy global -> shared (the per-chunk parameter block), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(t5::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(t5::smem_u32(bar)) : "memory");
}

template <int D, int CH>
__global__ void __launch_bounds__(kThreads, 1)
fused_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                   const __grid_constant__ CUtensorMap tmWeLo, const __grid_constant__ CUtensorMap tmWp,
                   const __grid_constant__ CUtensorMap tmWpLo, const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smX = smem + p.off_x;        // [k_blocks][256 rows][64 B]  swizzle 64B
    uint8_t* smWe = smem + p.off_we;      // [we_stages][planes][k_blocks][chunk rows][64 B]
    uint8_t* smWp = smem + p.off_wp;      // [wp_stages][planes][Cout rows][chunk*2 B]
    uint8_t* smDw = smem + p.off_dw;      // [2][halo_rows][dw_pitch]
    uint8_t* smA2 = smem + p.off_a2;      // [2][128 rows][chunk*2 B]    swizzle = row bytes
    float* smPar = reinterpret_cast<float*>(smem + p.off_par);     // [2][13][chunk]: s1 t1 wd[9] s2 t2
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int chunk = CH;

    if (threadIdx.x == 0) {
        t5::tma_prefetch_desc(&tmX); t5::tma_prefetch_desc(&tmWe); t5::tma_prefetch_desc(&tmWp);
        for (int b = 0; b < B_COUNT; ++b) {
            const bool e_arrives = (b == B_D1EMPTY0 || b == B_D1EMPTY1 || b == B_D2EMPTY || b == B_DWFULL0 || b == B_DWFULL1);
            const bool w_arrives = (b == B_DWEMPTY0 || b == B_DWEMPTY1);
            t5::mbar_init(&bars[b], e_arrives ? kGroup : (w_arrives ? kGroupW : 1));
        }
        t5::fence_barrier_init();
    }
    if (warp == 1) { t5::tmem_alloc(tmem_ptr, p.tmem_cols); t5::tmem_relinquish(); pdl_launch_dependents(); }
    pdl_wait();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_d2 = tmem_base + 2 * p.d1_cols;

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; chunks are numbered globally (g = tile_local * n_chunks + j)
    // so that the weight prefetch and the expand GEMMs run one chunk ahead ACROSS tile boundaries
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const long long G = static_cast<long long>(my_tiles) * p.n_chunks;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            uint32_t ph_x = 0, ph_we[3] = {0, 0, 0}, ph_wp[2] = {0, 0};     // parities of the *_EMPTY barriers this thread waits on next
            uint32_t ph_a2[2] = {0, 0};
            auto tile_of = [&](long long g, int* j) { const int tl = static_cast<int>(g / p.n_chunks); *j = static_cast<int>(g - static_cast<long long>(tl) * p.n_chunks); return tl; };
            // X tile (first chunk of a tile) + expand weights of global chunk g into ring slot g % we_stages
            auto issue_we = [&](long long g) {
                int j; const int tl = tile_of(g, &j);
                if (j == 0) {
                    int r = static_cast<int>(blockIdx.x) + tl * static_cast<int>(gridDim.x);
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y;
                    const int n = r / p.tiles_y;
                    if (tl > 0) { wait_bar_relaxed(&bars[B_XEMPTY], ph_x); ph_x ^= 1; }                  // every expand GEMM of the previous tile retired
                    t5::mbar_arrive_expect_tx(&bars[B_XFULL], static_cast<uint32_t>(p.k_blocks * p.halo_rows * 64));
                    for (int kb = 0; kb < p.k_blocks; ++kb)
                        tma_load_4d(smX + kb * (kXRows * 64), &tmX, &bars[B_XFULL], kb * kKB, tx * kTW - D, ty * kTH - D, n);
                }
                const int ws = static_cast<int>(g % p.we_stages);
                if (g >= p.we_stages) { wait_bar_relaxed(&bars[B_WEEMPTY0 + ws], ph_we[ws]); ph_we[ws] ^= 1; }   // expand GEMM g - we_stages retired
                uint8_t* dst = smWe + ws * p.we_stage;
                t5::mbar_arrive_expect_tx(&bars[B_WEFULL0 + ws], p.we_stage);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    t5::tma_load_2d(dst + kb * (chunk * 64), &tmWe, &bars[B_WEFULL0 + ws], kb * kKB, j * chunk);
                    if (p.we_split) t5::tma_load_2d(dst + p.we_tile + kb * (chunk * 64), &tmWeLo, &bars[B_WEFULL0 + ws], kb * kKB, j * chunk);
                }
            };
            // per-channel parameters of global chunk g; the buffer was last read by the depthwise phase two chunks ago
            auto issue_par = [&](long long g) {
                int j; tile_of(g, &j);
                const int sb = static_cast<int>(g & 1);
                if (g >= 2) { wait_bar_relaxed(&bars[B_A2FULL0 + sb], ph_a2[sb]); ph_a2[sb] ^= 1; }
                t5::mbar_arrive_expect_tx(&bars[B_PARFULL0 + sb], 13u * chunk * 4u);
                bulk_load_1d(smPar + sb * (13 * chunk), p.par + static_cast<long long>(j) * 13 * chunk, 13u * chunk * 4u, &bars[B_PARFULL0 + sb]);
            };
            auto issue_wp = [&](long long g) {
                int j; tile_of(g, &j);
                const int ws = static_cast<int>(g % p.wp_stages);
                if (g >= p.wp_stages) { wait_bar_relaxed(&bars[B_WPEMPTY0 + ws], ph_wp[ws]); ph_wp[ws] ^= 1; }   // project GEMM g - wp_stages retired
                uint8_t* dst = smWp + ws * p.wp_stage;
                t5::mbar_arrive_expect_tx(&bars[B_WPFULL0 + ws], p.wp_stage);
                for (int nt = 0; nt < p.np_tiles; ++nt) {
                    t5::tma_load_2d(dst + nt * (p.np_mma * chunk * 2), &tmWp, &bars[B_WPFULL0 + ws], j * chunk, nt * p.np);
                    if (p.wp_split) t5::tma_load_2d(dst + p.wp_tile + nt * (p.np_mma * chunk * 2), &tmWpLo, &bars[B_WPFULL0 + ws], j * chunk, nt * p.np);
                }
            };
            // everything is issued as far ahead as its ring allows; the loop body keeps the order of first use
            const int we_ahead = p.we_stages - 1, wp_ahead = p.wp_stages - 1;
            for (long long g = 0; g <= we_ahead && g < G; ++g) issue_we(g);
            if (G > 0) issue_par(0);
            for (long long g = 0; g <= wp_ahead && g < G; ++g) issue_wp(g);
            for (long long g = 0; g < G; ++g) {
                if (g + 1 < G) issue_par(g + 1);
                if (g + 1 + wp_ahead < G) issue_wp(g + 1 + wp_ahead);
                if (g + 1 + we_ahead < G) issue_we(g + 1 + we_ahead);
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one elected lane issues)
        const uint32_t idesc1 = t5::make_idesc_f16(128, chunk, 0, 0, 0, 0);
        const uint32_t idesc2 = t5::make_idesc_f16(128, p.np_mma, 0, 0, 0, 0);
        constexpr uint32_t lt2 = chunk == 64 ? 2u : 4u, sbo2 = chunk == 64 ? 1024u : 512u;       // A2 / Wp rows are chunk*2 bytes wide
        uint32_t ph_xfull = 0, ph_wefull[3] = {0, 0, 0}, ph_wpfull[2] = {0, 0}, ph_d2empty = 0;
        uint32_t ph_d1empty[2] = {0, 0}, ph_a2full[2] = {0, 0};
        auto expand = [&](long long g) {
            const int j = static_cast<int>(g % p.n_chunks);
            const int s = static_cast<int>(g & 1);
            const int ws = static_cast<int>(g % p.we_stages);
            if (j == 0) { wait_bar_relaxed(&bars[B_XFULL], ph_xfull); ph_xfull ^= 1; }
            wait_bar_relaxed(&bars[B_WEFULL0 + ws], ph_wefull[ws]); ph_wefull[ws] ^= 1;
            if (g >= 2) { wait_bar_relaxed(&bars[B_D1EMPTY0 + s], ph_d1empty[s]); ph_d1empty[s] ^= 1; }    // epilogue-1 of chunk g-2 drained the stage
            t5::fence_after_thread_sync();
            if (lane == 0) {
                dbg_mark(p, 0, g, 0);
                const uint32_t x_addr = t5::smem_u32(smX), w_addr = t5::smem_u32(smWe) + ws * p.we_stage;
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t d = tmem_base + s * p.d1_cols + mt * chunk;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
#pragma unroll
                        for (int k = 0; k < kKB / 16; ++k) {
                            const uint64_t da = t5::make_smem_desc(x_addr + kb * (kXRows * 64) + mt * (128 * 64) + k * 32, 16, 512, 4);
                            const uint64_t db = t5::make_smem_desc(w_addr + kb * (chunk * 64) + k * 32, 16, 512, 4);
                            t5::mma_bf16_ss(d, da, db, idesc1, (kb | k) != 0);
                            if (p.we_split) {
                                const uint64_t dl = t5::make_smem_desc(w_addr + p.we_tile + kb * (chunk * 64) + k * 32, 16, 512, 4);
                                t5::mma_bf16_ss(d, da, dl, idesc1, 1u);
                            }
                        }
                    }
                }
                t5::mma_commit(&bars[B_WEEMPTY0 + ws]);
                t5::mma_commit(&bars[B_D1FULL0 + s]);
                if (j == p.n_chunks - 1) t5::mma_commit(&bars[B_XEMPTY]);
            }
            __syncwarp();
        };
        if (G > 0) expand(0);
        for (long long g = 0; g < G; ++g) {
            if (g + 1 < G) expand(g + 1);
            const int j = static_cast<int>(g % p.n_chunks);
            const int b = static_cast<int>(g & 1);
            const int wps = static_cast<int>(g % p.wp_stages);
            if (j == 0 && g > 0) { wait_bar_relaxed(&bars[B_D2EMPTY], ph_d2empty); ph_d2empty ^= 1; }     // final epilogue of the previous tile drained D2
            wait_bar_relaxed(&bars[B_WPFULL0 + wps], ph_wpfull[wps]); ph_wpfull[wps] ^= 1;
            wait_bar_relaxed(&bars[B_A2FULL0 + b], ph_a2full[b]); ph_a2full[b] ^= 1;
            t5::fence_after_thread_sync();
            if (lane == 0) {
                dbg_mark(p, 0, g, 1);
                const uint32_t a_addr = t5::smem_u32(smA2) + b * p.a2_tile, w_addr = t5::smem_u32(smWp) + wps * p.wp_stage;
                for (int nt = 0; nt < p.np_tiles; ++nt) {
                    const uint32_t d = tmem_d2 + nt * p.np_mma;
#pragma unroll
                    for (int k = 0; k < chunk / 16; ++k) {
                        const uint64_t da = t5::make_smem_desc(a_addr + k * 32, 16, sbo2, lt2);
                        const uint64_t db = t5::make_smem_desc(w_addr + nt * (p.np_mma * chunk * 2) + k * 32, 16, sbo2, lt2);
                        t5::mma_bf16_ss(d, da, db, idesc2, (j | k) != 0);
                        if (p.wp_split) {
                            const uint64_t dl = t5::make_smem_desc(w_addr + p.wp_tile + nt * (p.np_mma * chunk * 2) + k * 32, 16, sbo2, lt2);
                            t5::mma_bf16_ss(d, da, dl, idesc2, 1u);
                        }
                    }
                }
                t5::mma_commit(&bars[B_WPEMPTY0 + wps]);
                t5::mma_commit(&bars[B_A2EMPTY0 + b]);
                if (j == p.n_chunks - 1) t5::mma_commit(&bars[B_D2FULL]);
            }
            __syncwarp();
        }
    } else {
        // ================================================================== worker groups
        const uint32_t dw_s = t5::smem_u32(smDw), a2_s = t5::smem_u32(smA2), par_s = t5::smem_u32(smPar);
        const uint32_t dw_tile = static_cast<uint32_t>(p.halo_rows * p.dw_pitch + 15) & ~15u;
        if (warp < 10) {
            // -------------------------------------------------------------- group E: D1 -> BN1 + ReLU6 -> fp16 halo tile; D2 -> HBM
            const int q = warp & 3;                            // TMEM lane quarter of this warp
            const int mt = (warp - 2) >> 2;                    // M tile of the halo rows this warp owns in epilogue-1
            uint32_t ph_d1full[2] = {0, 0}, ph_par[2] = {0, 0}, ph_dwempty[2] = {0, 0}, ph_d2full = 0;
            long long g = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                int r = t;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int ty = r % p.tiles_y;
                const int n = r / p.tiles_y;
                const int y0 = ty * kTH, x0 = tx * kTW;
                const int hp = mt * 128 + q * 32 + lane;       // this thread's halo pixel
                const int hy = hp / p.halo_w, hx = hp - hy * p.halo_w;
                const int gy = y0 - D + hy, gx = x0 - D + hx;
                const bool hp_in_tile = hp < p.halo_rows;
                const bool hp_in_image = hp_in_tile && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                const bool warp_has_rows = (mt * 128 + q * 32) < p.halo_rows;
                for (int j = 0; j < p.n_chunks; ++j, ++g) {
                    const int s = static_cast<int>(g & 1);
                    const uint32_t par = par_s + s * (13 * chunk * 4);                      // shared-memory byte address
                    wait_bar(&bars[B_PARFULL0 + s], ph_par[s]); ph_par[s] ^= 1;
                    if (g >= 2) { wait_bar(&bars[B_DWEMPTY0 + s], ph_dwempty[s]); ph_dwempty[s] ^= 1; }      // W has left this halo buffer (chunk g-2)
                    if (threadIdx.x == 64) dbg_mark(p, 1, g, 0);
                    wait_bar(&bars[B_D1FULL0 + s], ph_d1full[s]); ph_d1full[s] ^= 1;
                    t5::fence_after_thread_sync();
                    if (threadIdx.x == 64) dbg_mark(p, 1, g, 1);
                    if (warp_has_rows) {
                        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * p.d1_cols + mt * chunk;
                        const uint32_t rowp = dw_s + s * dw_tile + hp * p.dw_pitch;
#pragma unroll
                        for (int c0 = 0; c0 < chunk; c0 += 32) {
                            uint32_t ra[16], rb[16];
                            t5::tmem_ld16(taddr + c0, ra);
                            t5::tmem_ld16(taddr + c0 + 16, rb);
                            t5::tmem_ld_wait();
                            if (hp_in_tile) {
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const uint32_t* rr = hh ? rb : ra;
                                    const int cb = c0 + hh * 16;
                                    float v[16];
#pragma unroll
                                    for (int k = 0; k < 16; k += 4) {
                                        const float4 sc = lds_f4(par + (cb + k) * 4);
                                        const float4 sh = lds_f4(par + (chunk + cb + k) * 4);
                                        v[k] = fminf(fmaxf(fmaf(__uint_as_float(rr[k]), sc.x, sh.x), 0.f), 6.f);
                                        v[k + 1] = fminf(fmaxf(fmaf(__uint_as_float(rr[k + 1]), sc.y, sh.y), 0.f), 6.f);
                                        v[k + 2] = fminf(fmaxf(fmaf(__uint_as_float(rr[k + 2]), sc.z, sh.z), 0.f), 6.f);
                                        v[k + 3] = fminf(fmaxf(fmaf(__uint_as_float(rr[k + 3]), sc.w, sh.w), 0.f), 6.f);
                                    }
                                    if (!hp_in_image) {
#pragma unroll
                                        for (int k = 0; k < 16; ++k) v[k] = 0.f;
                                    }
                                    sts128(rowp + cb * 2, pack8h(v));
                                    sts128(rowp + cb * 2 + 16, pack8h(v + 8));
                                }
                            }
                        }
                    }
                    t5::fence_before_thread_sync();
                    t5::mbar_arrive(&bars[B_D1EMPTY0 + s]);     // TMEM stage free for the expand GEMM of chunk g + 2
                    t5::mbar_arrive(&bars[B_DWFULL0 + s]);      // (release) this thread's part of the halo tile is written
                    if (threadIdx.x == 64) dbg_mark(p, 1, g, 2);
                }
                // ---------------------------------------------------------- final epilogue: D2 -> BN3 (+ x) -> fp16 -> HBM
                wait_bar(&bars[B_D2FULL], ph_d2full); ph_d2full ^= 1;
                t5::fence_after_thread_sync();
                {
                    const int prow = q * 32 + lane;
                    const int oy = y0 + prow / kTW, ox = x0 + prow % kTW;
                    const bool ok = oy < p.H && ox < p.W;
                    const long long pix = (static_cast<long long>(n) * p.H + oy) * p.W + ox;
                    const int ncols = p.Cout;
                    const uint32_t taddr = tmem_d2 + (static_cast<uint32_t>(q * 32) << 16);
                    for (int c0 = mt * 16; c0 < ncols; c0 += 32) {              // the two warps of a quarter interleave 16-column groups
                        uint32_t rr[16];
                        t5::tmem_ld16(taddr + c0, rr);
                        t5::tmem_ld_wait();
                        if (!ok) continue;
                        float v[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = (c0 + k < ncols) ? fmaf(__uint_as_float(rr[k]), __ldg(p.s3 + c0 + k), __ldg(p.t3 + c0 + k)) : 0.f;
                        if (p.residual) {
                            const __half* rp = p.residual + pix * p.Cout + c0;
                            float f[8];
                            unpack8h(ldg_stream(rp), f);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += f[k];
                            if (c0 + 8 < ncols) {
                                unpack8h(ldg_stream(rp + 8), f);
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[8 + k] += f[k];
                            }
                        }
                        __half* o = p.out + pix * p.Cout + c0;
                        stg_stream(o, pack8h(v));
                        if (c0 + 8 < ncols) stg_stream(o + 8, pack8h(v + 8));
                    }
                }
                t5::fence_before_thread_sync();
                t5::mbar_arrive(&bars[B_D2EMPTY]);
            }
        } else {
            // -------------------------------------------------------------- group W: depthwise 3x3 + BN2 + ReLU6 -> A2 (UMMA K-major, swizzled)
            const int wt = threadIdx.x - 320;                  // 0..255
            constexpr int CV4 = chunk / 4;                     // threads per pixel
            constexpr int SL = 4;                              // strip length (pixels along x)
            constexpr int NSTRIP = kTW / SL;
            constexpr int NPT = kGroupW / CV4;                 // pixel-threads: 16 (chunk 64) or 32 (chunk 32)
            constexpr int ITEMS = kTH * NSTRIP;                // 32 (row, strip) items
            const int l4 = wt % CV4, pt = wt / CV4;
            constexpr int row_bytes2 = chunk * 2;
            uint32_t ph_dwfull[2] = {0, 0}, ph_a2empty[2] = {0, 0}, ph_par[2] = {0, 0};
            for (long long g = 0; g < G; ++g) {
                const int b = static_cast<int>(g & 1);
                const uint32_t par = par_s + b * (13 * chunk * 4);
                wait_bar(&bars[B_PARFULL0 + b], ph_par[b]); ph_par[b] ^= 1;
                const int c0 = l4 * 4;
                float2 wk[9][2];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4 a = lds_f4(par + ((2 + k) * chunk + c0) * 4);
                    wk[k][0] = make_float2(a.x, a.y); wk[k][1] = make_float2(a.z, a.w);
                }
                const float4 s2v = lds_f4(par + (11 * chunk + c0) * 4);
                const float4 t2v = lds_f4(par + (12 * chunk + c0) * 4);
                if (wt == 0) dbg_mark(p, 2, g, 0);
                wait_bar(&bars[B_DWFULL0 + b], ph_dwfull[b]); ph_dwfull[b] ^= 1;                 // (acquire) E finished this chunk's halo tile
                if (wt == 0) dbg_mark(p, 2, g, 1);
                if (g >= 2) { wait_bar(&bars[B_A2EMPTY0 + b], ph_a2empty[b]); ph_a2empty[b] ^= 1; }  // project GEMM g-2 retired: A2[b] reusable
                const uint32_t dwb = dw_s + b * dw_tile, a2b = a2_s + b * p.a2_tile;
                for (int it = pt; it < ITEMS; it += NPT) {
                    const int oy = it / NSTRIP, sx = it - oy * NSTRIP;
                    float2 acc[SL][2];
#pragma unroll
                    for (int a = 0; a < SL; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
                    constexpr int NCOLS = (SL - 1) + 2 * D + 1;      // columns that feed a strip of SL outputs
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t rp = dwb + ((oy + ky * D) * p.halo_w + sx * SL) * p.dw_pitch + c0 * 2;
#pragma unroll
                        for (int jx = 0; jx < NCOLS; ++jx) {
                            bool used = false;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) used = used || ((jx - kx * D) >= 0 && (jx - kx * D) < SL);
                            if (!used) continue;
                            const uint2 raw = lds64(rp + jx * p.dw_pitch);
                            const float2 v0 = h16x2(raw.x), v1 = h16x2(raw.y);
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int tt = jx - kx * D;
                                if (tt >= 0 && tt < SL) { ffma2(acc[tt][0], v0, wk[ky * 3 + kx][0]); ffma2(acc[tt][1], v1, wk[ky * 3 + kx][1]); }
                            }
                        }
                    }
#pragma unroll
                    for (int a = 0; a < SL; ++a) {
                        const int prow = oy * kTW + sx * SL + a;                 // A2 row = output pixel
                        const float f0 = fminf(fmaxf(fmaf(acc[a][0].x, s2v.x, t2v.x), 0.f), 6.f);
                        const float f1 = fminf(fmaxf(fmaf(acc[a][0].y, s2v.y, t2v.y), 0.f), 6.f);
                        const float f2 = fminf(fmaxf(fmaf(acc[a][1].x, s2v.z, t2v.z), 0.f), 6.f);
                        const float f3 = fminf(fmaxf(fmaf(acc[a][1].y, s2v.w, t2v.w), 0.f), 6.f);
                        // K-major swizzled row: 16-byte chunk index XOR (row mod 8) [128B rows] or ((row / 2) mod 4) [64B rows]
                        const int ck = c0 >> 3;
                        constexpr int nck = row_bytes2 >> 4;
                        const int sw = chunk == 64 ? (prow & 7) : ((prow >> 1) & 3);
                        const uint32_t addr = a2b + prow * row_bytes2 + (((ck ^ sw) & (nck - 1)) << 4) + (c0 & 7) * 2;
                        sts64(addr, make_uint2(pack_h16(f0, f1), pack_h16(f2, f3)));
                    }
                }
                t5::mbar_arrive(&bars[B_DWEMPTY0 + b]);         // this thread has left the halo buffer
                t5::fence_proxy_async_smem();                   // generic-proxy writes of A2 -> visible to the tensor core
                t5::named_barrier_sync(1, kGroupW);
                if (wt == 0) { t5::mbar_arrive(&bars[B_A2FULL0 + b]); dbg_mark(p, 2, g, 2); }
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

}  // namespace

// expanded channels per chunk: 64 where the resident X tile, the double-buffered halo tile and at least a (2, 1) weight ring fit
// next to it, else 32.  Also fixes the chunk-major layout of the parameter block.
static int fused_chunk(const FusedBlockDesc& d) {
    for (int chunk = 64; chunk >= 32; chunk /= 2) {
        const int kb = ceil_div(d.Cin, kKB);
        const int halo_rows = (kTW + 2 * d.dil) * (kTH + 2 * d.dil);
        const int np_tiles = d.Cout > 256 ? 2 : 1, np_mma = ceil_div(d.Cout / np_tiles, 16) * 16;
        // geometry only (callers size the parameter block before the weight pointers exist): both weights assumed split
        const size_t we = size_t(kb) * chunk * 64 * 2, wp = size_t(np_tiles) * np_mma * chunk * 2 * (d.Cout > 256 ? 1 : 2);
        const size_t total = size_t(kb) * kXRows * 64 + 2 * we + 2 * wp + 2 * size_t(halo_rows) * (chunk * 2 + 16) + 2 * 128 * chunk * 2 +
                             2 * 13 * chunk * 4 + 8 * 1024;
        if (total <= 227 * 1024 || chunk == 32) return chunk;
    }
    return 32;
}

bool fused_block_supported(const FusedBlockDesc& d) {
    if (d.stride != 1 || (d.dil != 1 && d.dil != 2)) return false;
    if (d.Cin % 8 || d.Cexp % 8 || d.Cout % 8 || d.Cout > 320) return false;
    if (d.Cout > 256 && (d.Cout % 32)) return false;
    return true;
}

size_t fused_block_param_floats(const FusedBlockDesc& d) {
    const int chunk = fused_chunk(d);
    const int cpad = ((d.Cexp + chunk - 1) / chunk) * chunk;
    return static_cast<size_t>(13) * cpad;
}

int fused_block_plan(const FusedBlockDesc& d, int num_sms, FusedBlockPlan* plan) {
    AMS_REQUIRE(fused_block_supported(d), "fused block: unsupported geometry");
    plan->d = d;
    FusedParams& p = *reinterpret_cast<FusedParams*>(plan->params);
    static_assert(sizeof(FusedParams) <= sizeof(plan->params), "FusedBlockPlan::params too small");
    p = FusedParams{};
    p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cexp = d.Cexp; p.Cout = d.Cout; p.dil = d.dil;
    p.k_blocks = ceil_div(d.Cin, kKB);
    p.we_split = d.We_lo ? 1 : 0; p.wp_split = d.Wp_lo ? 1 : 0;
    p.np_tiles = d.Cout > 256 ? 2 : 1; p.np = d.Cout / p.np_tiles; p.np_mma = ceil_div(p.np, 16) * 16;
    p.tiles_x = ceil_div(d.W, kTW); p.tiles_y = ceil_div(d.H, kTH); p.num_tiles = d.N * p.tiles_x * p.tiles_y;
    p.halo_w = kTW + 2 * d.dil; p.halo_h = kTH + 2 * d.dil; p.halo_rows = p.halo_w * p.halo_h;
    AMS_REQUIRE(p.halo_rows <= kXRows, "fused block: halo tile exceeds two M tiles");
    p.s3 = d.s3; p.t3 = d.t3;
    p.residual = static_cast<const __half*>(d.residual); p.out = static_cast<__half*>(d.out);
    p.dbg = static_cast<unsigned long long*>(d.debug_timeline);
    p.par = d.params;
    // the chunk width is fixed by the layout of the parameter block (fused_chunk()); the weight rings take what is left
    p.chunk = fused_chunk(d);
    p.n_chunks = ceil_div(d.Cexp, p.chunk);
    p.cpad = p.n_chunks * p.chunk;
    p.dw_pitch = p.chunk * 2 + 16;
    p.d1_cols = 2 * p.chunk;
    const uint32_t need = 2 * p.d1_cols + p.np_tiles * p.np_mma;
    AMS_REQUIRE(need <= 512, "fused block: TMEM overflow");
    p.tmem_cols = 512;
    p.we_tile = p.k_blocks * p.chunk * 64;
    p.we_stage = p.we_tile * (p.we_split ? 2 : 1);
    p.wp_tile = p.np_tiles * p.np_mma * p.chunk * 2;
    p.wp_stage = p.wp_tile * (p.wp_split ? 2 : 1);
    p.a2_tile = 128 * p.chunk * 2;
    const int ring_options[4][2] = {{3, 2}, {2, 2}, {2, 1}, {1, 1}};
    bool fits = false;
    for (int o = 0; o < 4 && !fits; ++o) {
        p.we_stages = ring_options[o][0]; p.wp_stages = ring_options[o][1];
        uint32_t off = 0;
        auto take = [&](uint32_t bytes) { const uint32_t at = off; off = (off + bytes + 1023u) & ~1023u; return at; };
        p.off_x = take(p.k_blocks * kXRows * 64);
        p.off_we = take(p.we_stage * p.we_stages);
        p.off_wp = take(p.wp_stage * p.wp_stages);
        p.off_dw = take(2 * ((p.halo_rows * p.dw_pitch + 15) & ~15));
        p.off_a2 = take(2 * p.a2_tile);
        p.off_par = take(2 * 13 * p.chunk * 4);
        p.off_bar = take(B_COUNT * 8 + 16);
        plan->smem_bytes = off + 1024;
        fits = plan->smem_bytes <= 227 * 1024;
    }
    AMS_REQUIRE(fits, "fused block: shared memory overflow (" + std::to_string(plan->smem_bytes) + " bytes)");
    AMS_REQUIRE((p.we_tile % 512) == 0 && (p.wp_tile % 1024) == 0 && ((p.np_mma * p.chunk * 2) % 1024) == 0 && (p.we_stage % 1024) == 0,
                "fused block: weight tile alignment");
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
        cuuint32_t box[2] = {kKB, static_cast<cuuint32_t>(p.chunk)};
        if (encode(&plan->tmWe, d.We, 2, dims, strides, box, 64)) return -1;
        if (encode(&plan->tmWeLo, d.We_lo ? d.We_lo : d.We, 2, dims, strides, box, 64)) return -1;
    }
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Cexp), static_cast<cuuint64_t>(d.Cout)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.ld_wp) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(p.chunk), static_cast<cuuint32_t>(p.np_mma)};
        if (encode(&plan->tmWp, d.Wp, 2, dims, strides, box, p.chunk * 2)) return -1;
        if (encode(&plan->tmWpLo, d.Wp_lo ? d.Wp_lo : d.Wp, 2, dims, strides, box, p.chunk * 2)) return -1;
    }
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<1, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<2, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<1, 32>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute((fused_block_kernel<2, 32>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    return 0;
}

int fused_block_launch(const FusedBlockPlan& plan, cudaStream_t s) {
    const FusedParams& p = *reinterpret_cast<const FusedParams*>(plan.params);
    if (p.dil == 1 && p.chunk == 64)
        AMS_LAUNCH((fused_block_kernel<1, 64>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    else if (p.dil == 2 && p.chunk == 64)
        AMS_LAUNCH((fused_block_kernel<2, 64>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    else if (p.dil == 1)
        AMS_LAUNCH((fused_block_kernel<1, 32>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    else
        AMS_LAUNCH((fused_block_kernel<2, 32>), plan.grid, kThreads, plan.smem_bytes, s, plan.tmX, plan.tmWe, plan.tmWeLo, plan.tmWp, plan.tmWpLo, p);
    return 0;
}

// params[n_chunks][13][chunk]: s1 t1 (folded BN of the expand conv), wd[9] (depthwise filter, [3,3,C] fp32), s2 t2 (folded BN
// of the depthwise conv), chunk-major so that one bulk copy brings a chunk's block; channels >= Cexp are zero so that a
// padded chunk contributes exactly nothing.
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
    const int chunk = fused_chunk(d);
    const int cpad = ((d.Cexp + chunk - 1) / chunk) * chunk;
    AMS_LAUNCH((fused_fill_params_kernel), ceil_div(13 * cpad, 256), 256, 0, s, s1, t1, wd, s2, t2, d.Cexp, cpad, chunk, const_cast<float*>(d.params));
    return 0;
}

}  // namespace ams
