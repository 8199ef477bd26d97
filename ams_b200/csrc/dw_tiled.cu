// Shared-memory tiled depthwise 3x3 kernels (forward, and the fused backward) for the 17 depthwise layers.
//
// A CTA owns one spatial tile x one channel chunk (32/48/64 channels) of one image.  The input tile with its halo is
// staged ONCE in shared memory by 128-bit coalesced loads (conv zero padding resolved there), optionally passing
// through the producer's BatchNorm + ReLU6 on the way in (training: the producer wrote its raw conv output z and the
// normalised tensor is never materialised).  Compute threads own 4 channels x a strip of 4 pixels and walk the tile
// with a register sliding window over 64-bit shared-memory reads; the multiply-accumulates are packed fp32x2
// (FFMA2: two channels per instruction).  On B200 these layers are bound by instruction issue as much as by HBM
// (6.5 TB/s leaves ~22 thread-instructions per bf16 element moved), so everything that can be a compile-time
// constant is one (channel chunk, stride, dilation) and shared memory is addressed with 32-bit offsets.
//
// Forward: optional folded (frozen) BatchNorm + activation on the way out, and/or the per-tile column sums
// (sum z, sum z^2 of the STORED bf16 values) that the BatchNorm finalize reduces in a fixed order.
// Backward (dw_bwd_fused): see the kernel.  No atomics anywhere: deterministic.
//
// Replaces: the `DepthwiseConv2dNative` nodes (+ SpaceToBatchND/BatchToSpaceND atrous wrappers), the
// FusedBatchNormV3/Relu6 that follow the expand convs, and their TF-generated gradients
// (checkpoints/*/model.meta; reference SemanticNetwork.py:260 runs them through tf.Session.run).
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace ams {
namespace {

constexpr int kStrip = 4;     // pixels per thread strip
constexpr int kCh = 4;        // channels per compute thread (64-bit shared-memory reads, two fp32x2 lanes)

__host__ __device__ constexpr int dw_threads(int CB) { return (256 / (CB / kCh)) * (CB / kCh); }

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {          // d += a * b (per lane)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ void fadd2(float2& d, const float2& a) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)));
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float2 unpack2(uint32_t v) { return make_float2(bf16_lo(v), bf16_hi(v)); }

// ------------------------------------------------------------------------------------------------ forward
struct DwFwdParams {
    const bf16* in; bf16* out; const float* w;            // w [9][C] fp32
    int N, H, W, C, Ho, Wo, pad_top, pad_left;
    const float* in_scale; const float* in_shift; int in_act;      // != null: act(in*scale+shift) applied while staging
    const float* out_scale; const float* out_shift; int out_act;   // != null: folded BN + act applied to the result
    double* stats;                                                  // != null: [tile][2][C] column sums of the stored tile
    int th, twt, ntx, nty, chunks, nstrips, ih, iwp;
};

// Stage rows [0,ih) x cols [0,iwp) x CB channels of `img` (image-relative origin gy0,gx0; outside the image = 0) as bf16
// [ih][iwp][CB]; optional per-channel affine + activation on the way (exactly the bn_apply arithmetic, rounded to bf16).
template <int CB, int THREADS>
__device__ __forceinline__ void stage_tile(uint32_t sbase, const bf16* __restrict__ img /* + channel chunk */, int C, int H,
                                           int W, int gy0, int gx0, int ih, int iwp, const float* __restrict__ scale,
                                           const float* __restrict__ shift, int act, int c_chunk0) {
    constexpr int CV8 = CB / 8, PXT = THREADS / CV8, U = 4;
    const int c8 = threadIdx.x % CV8, lane_px = threadIdx.x / CV8;
    float sc[8], sh[8];
    if (scale) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { sc[q] = scale[c_chunk0 + c8 * 8 + q]; sh[q] = shift[c_chunk0 + c8 * 8 + q]; }
    }
    int ly = lane_px / iwp, lx = lane_px - ly * iwp;
    const int step_y = PXT / iwp, step_x = PXT - step_y * iwp;
    uint32_t sdst = sbase + (lane_px * CB + c8 * 8) * 2;
    const bf16* src = img + c8 * 8;
    // software pipeline: the loads of batch k+1 are in flight while batch k is transformed and stored
    uint4 v[U], vn[U];
    int st[U], stn[U];                                  // 0 = past the tile, 1 = outside the image (zero), 2 = loaded
    auto issue = [&](uint4* dst, int* state) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int gy = gy0 + ly, gx = gx0 + lx;
            state[u] = ly < ih ? ((gy >= 0 && gy < H && gx >= 0 && gx < W) ? 2 : 1) : 0;
            dst[u] = make_uint4(0u, 0u, 0u, 0u);
            if (state[u] == 2) dst[u] = ldg_stream(src + (static_cast<long long>(gy) * W + gx) * C);
            lx += step_x; ly += step_y;
            if (lx >= iwp) { lx -= iwp; ++ly; }
        }
    };
    issue(vn, stn);
    while (stn[0]) {
#pragma unroll
        for (int u = 0; u < U; ++u) { v[u] = vn[u]; st[u] = stn[u]; }
        if (st[U - 1]) issue(vn, stn); else stn[0] = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (st[u]) {
                if (scale && st[u] == 2) {
                    float f[8];
                    unpack8(v[u], f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) f[q] = act_apply(fmaf(f[q], sc[q], sh[q]), act);
                    v[u] = pack8(f);
                }
                sts128(sdst + u * (PXT * CB * 2), v[u]);
            }
        }
        sdst += U * (PXT * CB * 2);
    }
}

template <int S, int D, int CB>
__global__ void __launch_bounds__(dw_threads(CB), 3)
dw_fwd_tiled_kernel(const DwFwdParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int THREADS = dw_threads(CB), CV4 = CB / kCh, NPT = THREADS / CV4;
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int chunk = blockIdx.x % p.chunks;
    int t = blockIdx.x / p.chunks;
    const int tx = t % p.ntx; t /= p.ntx;
    const int ty = t % p.nty;
    const int n = t / p.nty;
    const int c_base = chunk * CB;
    const int oy0 = ty * p.th, ox0 = tx * p.twt;

    stage_tile<CB, THREADS>(sbase, p.in + static_cast<long long>(n) * p.H * p.W * p.C + c_base, p.C, p.H, p.W,
                            oy0 * S - p.pad_top, ox0 * S - p.pad_left, p.ih, p.iwp, p.in_scale, p.in_shift, p.in_act, c_base);
    __syncthreads();

    // ---------------------------------------------------------------- compute: thread = 4 channels x strips of 4 pixels
    const int l4 = threadIdx.x % CV4, pt = threadIdx.x / CV4;
    const int c0 = c_base + l4 * kCh;
    float2 wk[9][2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(p.w + k * p.C + c0);
        wk[k][0] = make_float2(a.x, a.y); wk[k][1] = make_float2(a.z, a.w);
    }
    float osc[kCh], osh[kCh];
    if (p.out_scale) {
#pragma unroll
        for (int q = 0; q < kCh; ++q) { osc[q] = p.out_scale[c0 + q]; osh[q] = p.out_shift[c0 + q]; }
    }
    float2 ssum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, ssq[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    constexpr int NCOLS = (kStrip - 1) * S + 2 * D + 1;
    const uint32_t row_bytes = static_cast<uint32_t>(p.iwp) * CB * 2;
    int r = pt / p.nstrips, s = pt - r * p.nstrips;
    const int step_r = NPT / p.nstrips, step_s = NPT - step_r * p.nstrips;
    for (; r < p.th; r += step_r, s += step_s) {
        if (s >= p.nstrips) { s -= p.nstrips; ++r; if (r >= p.th) break; }
        float2 acc[kStrip][2];
#pragma unroll
        for (int a = 0; a < kStrip; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
        const uint32_t base = sbase + static_cast<uint32_t>(r * S) * row_bytes + (s * (kStrip * S) * CB + l4 * kCh) * 2;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const uint32_t rowp = base + static_cast<uint32_t>(ky * D) * row_bytes;
#pragma unroll
            for (int j = 0; j < NCOLS; ++j) {
                bool used = false;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) used = used || ((j - kx * D) >= 0 && (j - kx * D) % S == 0 && (j - kx * D) / S < kStrip);
                if (!used) continue;
                const uint2 raw = lds64(rowp + j * (CB * 2));
                const float2 v0 = unpack2(raw.x), v1 = unpack2(raw.y);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int tt = j - kx * D;
                    if (tt >= 0 && tt % S == 0 && tt / S < kStrip) {
                        ffma2(acc[tt / S][0], v0, wk[ky * 3 + kx][0]);
                        ffma2(acc[tt / S][1], v1, wk[ky * 3 + kx][1]);
                    }
                }
            }
        }
        const int oy = oy0 + r;
        if (oy >= p.Ho) continue;
        bf16* orow = p.out + ((static_cast<long long>(n) * p.Ho + oy) * p.Wo) * p.C + c0;
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            const int lx = s * kStrip + a, ox = ox0 + lx;
            if (lx >= p.twt || ox >= p.Wo) break;
            float f[4] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y};
            if (p.out_scale) {
#pragma unroll
                for (int q = 0; q < kCh; ++q) f[q] = act_apply(fmaf(f[q], osc[q], osh[q]), p.out_act);
            }
            const uint2 pk = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
            *reinterpret_cast<uint2*>(orow + static_cast<long long>(ox) * p.C) = pk;
            if (p.stats) {
                const float2 r0 = unpack2(pk.x), r1 = unpack2(pk.y);
                fadd2(ssum[0], r0); fadd2(ssum[1], r1);
                ffma2(ssq[0], r0, r0); ffma2(ssq[1], r1, r1);
            }
        }
    }
    if (p.stats) {
        // fixed-order block reduction: red[pt][2][CB] fp32, then one thread per (stat, channel) walks the pixel-threads
        __syncthreads();
        float* red = reinterpret_cast<float*>(smem);
        float* mine = red + (pt * 2) * CB + l4 * kCh;
        mine[0] = ssum[0].x; mine[1] = ssum[0].y; mine[2] = ssum[1].x; mine[3] = ssum[1].y;
        mine[CB + 0] = ssq[0].x; mine[CB + 1] = ssq[0].y; mine[CB + 2] = ssq[1].x; mine[CB + 3] = ssq[1].y;
        __syncthreads();
        if (threadIdx.x < 2 * CB) {
            double a = 0.0;
            for (int k = 0; k < NPT; ++k) a += static_cast<double>(red[k * 2 * CB + threadIdx.x]);
            const int which = threadIdx.x / CB, c = threadIdx.x - which * CB;
            const long long tile = blockIdx.x / p.chunks;
            p.stats[(tile * 2 + which) * p.C + c_base + c] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------ tile selection
struct DwTile { int th, twt, ntx, nty, cb, chunks, nstrips, ih, iwp, threads; size_t smem; long long blocks; };

int pick_cb(int C) {
    if (C % 64 == 0) return 64;
    if (C % 48 == 0) return 48;
    if (C % 32 == 0) return 32;
    return 0;
}

// Minimise (padded work + staged pixels) over tile heights / widths; shared memory capped so that 3 CTAs fit an SM.
DwTile pick_tile(const Conv2dGeom& g, size_t smem_cap) {
    DwTile best{};
    double best_cost = 1e300;
    const int cb = pick_cb(g.C);
    if (!cb) return best;
    const int S = g.stride, D = g.dil;
    const int cv4 = cb / kCh;
    const int npt = 256 / cv4;
    for (int th = 2; th <= 24; ++th) {
        const int nty = ceil_div(g.Ho, th);
        for (int ntx = 1; ntx <= 32; ++ntx) {
            const int twt = ceil_div(g.Wo, ntx);
            if (twt < 8 && ntx > 1) break;
            const int nstrips = ceil_div(twt, kStrip);
            const int ih = (th - 1) * S + 2 * D + 1;
            const int iwp = (nstrips * kStrip - 1) * S + 2 * D + 1;
            size_t smem = static_cast<size_t>(ih) * iwp * cb * 2;
            if (smem > smem_cap) continue;
            smem = std::max(smem, static_cast<size_t>(npt) * 2 * cb * sizeof(float));        // statistics scratch
            const int items = th * nstrips;
            const int iters = ceil_div(items, npt);
            const long long blocks = static_cast<long long>(g.N) * nty * ntx * (g.C / cb);
            // cost: whole waves of CTAs (3 per SM) x per-CTA time (fixed + staging + compute iterations), in pixel-slots
            const double compute = static_cast<double>(iters) * npt * kStrip;
            const double stage = static_cast<double>(ih) * iwp * 0.6;
            const double waves = static_cast<double>(ceil_div_ll(blocks, kNumSMs * 3LL));
            const double cost = waves * (compute + stage + 150.0);
            if (cost < best_cost) {
                best_cost = cost;
                best = DwTile{th, twt, ntx, nty, cb, g.C / cb, nstrips, ih, iwp, cv4 * npt, smem, blocks};
            }
        }
    }
    return best;
}

constexpr size_t kFwdSmemCap = 72 * 1024;

template <int S, int D, int CB>
int launch_fwd(const DwFwdParams& p, const DwTile& t, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_fwd_tiled_kernel<S, D, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr = true;
    }
    dw_fwd_tiled_kernel<S, D, CB><<<static_cast<unsigned>(t.blocks), dw_threads(CB), t.smem, s>>>(p);
    AMS_LAUNCH_CHECK();
    return 0;
}
template <int S, int D>
int launch_fwd_cb(const DwFwdParams& p, const DwTile& t, cudaStream_t s) {
    if (t.cb == 64) return launch_fwd<S, D, 64>(p, t, s);
    if (t.cb == 48) return launch_fwd<S, D, 48>(p, t, s);
    return launch_fwd<S, D, 32>(p, t, s);
}

}  // namespace

// diagnostics (tools/): the tile the planner picks for a geometry
extern "C" int ams_debug_dw_tile(int N, int H, int W, int C, int Ho, int Wo, int stride, int dil, int* out8) {
    Conv2dGeom g{N, H, W, C, Ho, Wo, stride, dil, dil, dil};
    const DwTile t = pick_tile(g, kFwdSmemCap);
    out8[0] = t.th; out8[1] = t.twt; out8[2] = t.ntx; out8[3] = t.nty; out8[4] = t.cb; out8[5] = static_cast<int>(t.smem);
    out8[6] = static_cast<int>(t.blocks); out8[7] = t.threads;
    return 0;
}

bool dw_tiled_supported(const Conv2dGeom& g) {
    return pick_cb(g.C) != 0 && ((g.stride == 1 && (g.dil == 1 || g.dil == 2)) || (g.stride == 2 && g.dil == 1));
}

long long dw_tiled_stats_rows(const Conv2dGeom& g) {
    const DwTile t = pick_tile(g, kFwdSmemCap);
    return static_cast<long long>(g.N) * t.nty * t.ntx;
}

int dw_conv_fwd_tiled(const bf16* in, const float* w, const Conv2dGeom& g, const float* in_scale, const float* in_shift,
                      int in_act, const float* out_scale, const float* out_shift, int out_act, bf16* out, double* stats,
                      int* stats_rows, cudaStream_t s) {
    AMS_REQUIRE(dw_tiled_supported(g), "tiled depthwise: unsupported channels / stride / dilation");
    const DwTile t = pick_tile(g, kFwdSmemCap);
    AMS_REQUIRE(t.blocks > 0, "tiled depthwise: no tile fits shared memory");
    DwFwdParams p;
    p.in = in; p.out = out; p.w = w;
    p.N = g.N; p.H = g.H; p.W = g.W; p.C = g.C; p.Ho = g.Ho; p.Wo = g.Wo; p.pad_top = g.pad_top; p.pad_left = g.pad_left;
    p.in_scale = in_scale; p.in_shift = in_shift; p.in_act = in_act;
    p.out_scale = out_scale; p.out_shift = out_shift; p.out_act = out_act;
    p.stats = stats;
    p.th = t.th; p.twt = t.twt; p.ntx = t.ntx; p.nty = t.nty; p.chunks = t.chunks; p.nstrips = t.nstrips;
    p.ih = t.ih; p.iwp = t.iwp;
    if (stats_rows) *stats_rows = static_cast<int>(static_cast<long long>(g.N) * t.nty * t.ntx);
    if (g.stride == 1 && g.dil == 1) return launch_fwd_cb<1, 1>(p, t, s);
    if (g.stride == 2 && g.dil == 1) return launch_fwd_cb<2, 1>(p, t, s);
    return launch_fwd_cb<1, 2>(p, t, s);
}

}  // namespace ams
