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
__device__ __forceinline__ float2 unpack2(uint32_t v) { return make_float2(bf16_lo(v), bf16_hi(v)); }      // bf16 pair (gradients)
__device__ __forceinline__ float2 unpack2h(uint32_t v) { return h16x2(v); }                                 // fp16 pair (activations)

// ------------------------------------------------------------------------------------------------ forward
struct DwFwdParams {
    const act_t* in; act_t* out; const float* w;          // w [9][C] fp32
    int N, H, W, C, Ho, Wo, pad_top, pad_left;
    const float* in_scale; const float* in_shift; int in_act;      // != null: act(in*scale+shift) applied while staging
    const float* out_scale; const float* out_shift; int out_act;   // != null: folded BN + act applied to the result
    double* stats;                                                  // != null: [tile][2][C] column sums of the stored tile
    int th, twt, ntx, nty, chunks, nstrips, ih, iwp;
    int mg_w, mg_s, st_y, st_x, cs_r, cs_s;       // host-computed: division magics (n*mg>>16 == n/d, n < 256) and loop steps
};

// Stage rows [0,ih) x cols [0,iwp) x CB channels of `img` (image-relative origin gy0,gx0; outside the image = 0) as bf16
// [ih][iwp][CB]; optional per-channel affine + activation on the way (exactly the bn_apply arithmetic, rounded to bf16).
template <int CB, int THREADS>
__device__ __forceinline__ void stage_tile(uint32_t sbase, const act_t* __restrict__ img /* + channel chunk */, int C, int H,
                                           int W, int gy0, int gx0, int ih, int iwp, const float* __restrict__ scale,
                                           const float* __restrict__ shift, int act, int c_chunk0, int mg_w, int step_y,
                                           int step_x) {
    constexpr int CV8 = CB / 8, PXT = THREADS / CV8, U = 4;
    const int c8 = threadIdx.x % CV8, lane_px = threadIdx.x / CV8;
    float sc[8], sh[8];
    if (scale) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { sc[q] = scale[c_chunk0 + c8 * 8 + q]; sh[q] = shift[c_chunk0 + c8 * 8 + q]; }
    }
    int ly = (lane_px * mg_w) >> 16, lx = lane_px - ly * iwp;
    uint32_t sdst = sbase + (lane_px * CB + c8 * 8) * 2;
    const act_t* src = img + c8 * 8;
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
                    unpack8h(v[u], f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) f[q] = act_apply(fmaf(f[q], sc[q], sh[q]), act);
                    v[u] = pack8h(f);
                }
                sts128(sdst + u * (PXT * CB * 2), v[u]);
            }
        }
        sdst += U * (PXT * CB * 2);
    }
}

template <int S, int D, int CB, int MINB>
__global__ void __launch_bounds__(dw_threads(CB), MINB)
dw_fwd_tiled_kernel(const DwFwdParams p) {
    pdl_entry();
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
                            oy0 * S - p.pad_top, ox0 * S - p.pad_left, p.ih, p.iwp, p.in_scale, p.in_shift, p.in_act, c_base,
                            p.mg_w, p.st_y, p.st_x);
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
    int r = (pt * p.mg_s) >> 16, s = pt - r * p.nstrips;
    const int step_r = p.cs_r, step_s = p.cs_s;
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
                const float2 v0 = unpack2h(raw.x), v1 = unpack2h(raw.y);
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
        act_t* orow = p.out + ((static_cast<long long>(n) * p.Ho + oy) * p.Wo) * p.C + c0;
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            const int lx = s * kStrip + a, ox = ox0 + lx;
            if (lx >= p.twt || ox >= p.Wo) break;
            float f[4] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y};
            if (p.out_scale) {
#pragma unroll
                for (int q = 0; q < kCh; ++q) f[q] = act_apply(fmaf(f[q], osc[q], osh[q]), p.out_act);
            }
            const uint2 pk = make_uint2(pack_h16(f[0], f[1]), pack_h16(f[2], f[3]));
            *reinterpret_cast<uint2*>(orow + static_cast<long long>(ox) * p.C) = pk;
            if (p.stats) {
                const float2 r0 = unpack2h(pk.x), r1 = unpack2h(pk.y);
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
    if (C % 56 == 0) return 56;      // 728 = 13 x 56: the Xception middle flow of the teacher (forward kernel only)
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

static size_t env_kb(const char* name, size_t dflt_kb) { const char* e = getenv(name); return (e ? size_t(atoi(e)) : dflt_kb) * 1024; }
static const size_t kFwdSmemCap = env_kb("AMS_DWF_SMEM_KB", 72);

// resident CTAs per SM the forward kernel is compiled for: 3 (80 registers, a few spills) or 2 (128 registers, none)
static const int kFwdMinBlocks = [] { const char* e = getenv("AMS_DWF_MINB"); return (e && atoi(e) == 2) ? 2 : 3; }();

template <int S, int D, int CB, int MINB>
int launch_fwd_mb(const DwFwdParams& p, const DwTile& t, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_fwd_tiled_kernel<S, D, CB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr = true;
    }
    AMS_LAUNCH((dw_fwd_tiled_kernel<S, D, CB, MINB>), static_cast<unsigned>(t.blocks), dw_threads(CB), t.smem, s, p);
    return 0;
}
template <int S, int D, int CB>
int launch_fwd(const DwFwdParams& p, const DwTile& t, cudaStream_t s) {
    return kFwdMinBlocks == 2 ? launch_fwd_mb<S, D, CB, 2>(p, t, s) : launch_fwd_mb<S, D, CB, 3>(p, t, s);
}
template <int S, int D>
int launch_fwd_cb(const DwFwdParams& p, const DwTile& t, cudaStream_t s) {
    if (t.cb == 64) return launch_fwd<S, D, 64>(p, t, s);
    if (t.cb == 48) return launch_fwd<S, D, 48>(p, t, s);
    if (t.cb == 56) return launch_fwd<S, D, 56>(p, t, s);
    return launch_fwd<S, D, 32>(p, t, s);
}


// ------------------------------------------------------------------------------------------------ fused backward
// One kernel for everything between "gradient wrt the depthwise layer's activated output" and "gradient wrt the raw
// output of the layer that feeds it":
//   staging : gz = bf16(A*mask(g) + B*z + Cc)   BatchNorm backward of the depthwise layer itself, applied while the
//             tile of (g, z) with its halo is loaded (A, B, Cc from the column sums of a previous reduce pass)
//   compute : per INPUT pixel i (tile owner) and tap k:  o = (i + pad - k*D) / S
//             dX[i]   = sum_k w[k] * gz[o]               (data gradient)
//             dW[k]  += x[i] * gz[o]                     (filter gradient; every (o,k) pair is owned by exactly one i)
//             x[i]    = bf16(act(z_in*sc+sh))            the producer's BN+act recomputed from its raw output
//             gm      = act'(..) ? bf16(dX[i]) : 0       stored as the gradient wrt the producer's BN output
//             S1 += gm, S2 += gm * z_in                  column sums for the producer's BatchNorm backward
// so the normalised input, the depthwise dz and the unmasked dX never exist in HBM: 3 tensor reads + 1 write
// instead of 9 reads + 4 writes (bn_bwd reduce/apply + dw_bwd_filter + dw_bwd_data + next bn_bwd reduce).
struct DwBwdParams {
    const bf16* g; const act_t* z;                // [N,Ho,Wo,C] gradient wrt act(BN(z)) (bf16) and the raw depthwise output (fp16)
    const float* scale; const float* shift; int act;          // the depthwise layer's BN (activation mask)
    const float* coef;                            // [3][C]: A, B, Cc of its BN backward
    const act_t* zin;                             // [N,H,W,C] raw output of the producer (or its activation if in_scale == null)
    const float* in_scale; const float* in_shift; int in_act;
    const float* w;                               // [9][C]
    bf16* gout;                                   // [N,H,W,C]
    float* dw_partial;                            // [tile][9][C]
    double* bn_partial;                           // [tile][2][C]  (null if in_scale == null)
    int N, H, W, C, Ho, Wo, pad_top, pad_left;
    int th, twt, ntx, nty, chunks, nstrips, oh, owp;
    int mg_w, mg_s, st_y, st_x, cs_r, cs_s;
};

template <int S, int D, int PADX>
struct BwdCols {
    // staged column (relative to the strip's first staged column) read by input pixel t of the strip for tap kx, or -1
    __host__ __device__ static constexpr int col(int t, int kx) {
        return S == 1 ? (t + 2 * D - kx * D)
                      : (((t + PADX - kx) % 2 != 0) ? -1 : ((t + PADX - kx + 2) / 2 - PADX));
    }
    static constexpr int NJ = S == 1 ? (kStrip + 2 * D) : 3;
    static constexpr int STRIP_COLS = kStrip / S;       // staged columns a strip advances by
};

template <int S, int D, int CB, int PADX>
__global__ void __launch_bounds__(dw_threads(CB), 2)
dw_bwd_fused_kernel(const DwBwdParams p) {
    pdl_entry();
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int THREADS = dw_threads(CB), CV4 = CB / kCh, NPT = THREADS / CV4, CV8 = CB / 8, PXT = THREADS / CV8;
    typedef BwdCols<S, D, PADX> Cols;
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int chunk = blockIdx.x % p.chunks;
    int t_ = blockIdx.x / p.chunks;
    const long long tile = t_;
    const int tx = t_ % p.ntx; t_ /= p.ntx;
    const int ty = t_ % p.nty;
    const int n = t_ / p.nty;
    const int c_base = chunk * CB;
    const int iy0 = ty * p.th, ix0 = tx * p.twt;
    // first staged output row / column
    const int oy_lo = S == 1 ? (iy0 + p.pad_top - 2 * D) : (iy0 / 2 - 1 + p.pad_top);
    const int ox_lo = S == 1 ? (ix0 + p.pad_left - 2 * D) : (ix0 / 2 - 1 + p.pad_left);
    const uint32_t w_smem = sbase;                                        // [9][CB] fp32
    const uint32_t tile_smem = sbase + 9 * CB * 4;

    // ---------------------------------------------------------------- stage gz = BN-backward(g, z) (+ halo)
    {
        const int c8 = threadIdx.x % CV8, lane_px = threadIdx.x / CV8;
        const int c0 = c_base + c8 * 8;
        float2 sc2[4], sh2[4], ca2[4], cb2[4], cc2[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            sc2[h] = *reinterpret_cast<const float2*>(p.scale + c0 + 2 * h); sh2[h] = *reinterpret_cast<const float2*>(p.shift + c0 + 2 * h);
            ca2[h] = *reinterpret_cast<const float2*>(p.coef + c0 + 2 * h);
            cb2[h] = *reinterpret_cast<const float2*>(p.coef + p.C + c0 + 2 * h);
            cc2[h] = *reinterpret_cast<const float2*>(p.coef + 2 * p.C + c0 + 2 * h);
        }
        for (int i = threadIdx.x; i < 9 * CB; i += THREADS) {
            const int k = i / CB, c = i - k * CB;
            reinterpret_cast<float*>(smem)[i] = p.w[k * p.C + c_base + c];
        }
        constexpr int U = 4;
        int ly = (lane_px * p.mg_w) >> 16, lx = lane_px - ly * p.owp;
        const int step_y = p.st_y, step_x = p.st_x;
        uint32_t sdst = tile_smem + (lane_px * CB + c8 * 8) * 2;
        const long long img = static_cast<long long>(n) * p.Ho * p.Wo * p.C + c0;
        uint4 vg[U], vz[U], ng[U], nz[U];
        int st[U], stn[U];
        auto issue = [&](uint4* dg, uint4* dz, int* state) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int gy = oy_lo + ly, gx = ox_lo + lx;
                state[u] = ly < p.oh ? ((gy >= 0 && gy < p.Ho && gx >= 0 && gx < p.Wo) ? 2 : 1) : 0;
                if (state[u] == 2) {
                    const long long off = img + (static_cast<long long>(gy) * p.Wo + gx) * p.C;
                    dg[u] = ldg_stream(p.g + off);
                    dz[u] = ldg_stream(p.z + off);
                }
                lx += step_x; ly += step_y;
                if (lx >= p.owp) { lx -= p.owp; ++ly; }
            }
        };
        issue(ng, nz, stn);
        while (stn[0]) {
#pragma unroll
            for (int u = 0; u < U; ++u) { vg[u] = ng[u]; vz[u] = nz[u]; st[u] = stn[u]; }
            if (st[U - 1]) issue(ng, nz, stn); else stn[0] = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (st[u]) {
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (st[u] == 2) {
                        const uint32_t gw[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
                        const uint32_t zw[4] = {vz[u].x, vz[u].y, vz[u].z, vz[u].w};
                        float g[8];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            // two channels at a time: packed fp32x2 FMAs, scalar compare/select for the activation mask
                            const float2 z2 = unpack2h(zw[h]);
                            float2 gm = unpack2(gw[h]);
                            float2 yh = sh2[h];
                            ffma2(yh, z2, sc2[h]);
                            if (p.act == 1) { gm.x = yh.x > 0.f ? gm.x : 0.f; gm.y = yh.y > 0.f ? gm.y : 0.f; }
                            else if (p.act == 2) {
                                gm.x = (yh.x > 0.f && yh.x < 6.f) ? gm.x : 0.f;
                                gm.y = (yh.y > 0.f && yh.y < 6.f) ? gm.y : 0.f;
                            }
                            float2 t2 = cc2[h];
                            ffma2(t2, cb2[h], z2);
                            ffma2(t2, ca2[h], gm);
                            g[2 * h] = t2.x; g[2 * h + 1] = t2.y;
                        }
                        o = pack8(g);
                    }
                    sts128(sdst + u * (PXT * CB * 2), o);
                }
            }
            sdst += U * (PXT * CB * 2);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- compute: thread = 4 channels x strips of 4 input pixels
    const int l4 = threadIdx.x % CV4, pt = threadIdx.x / CV4;
    const int c0 = c_base + l4 * kCh;
    float isc[kCh], ish[kCh];
    if (p.in_scale) {
#pragma unroll
        for (int q = 0; q < kCh; ++q) { isc[q] = p.in_scale[c0 + q]; ish[q] = p.in_shift[c0 + q]; }
    }
    float2 dW[9][2];
#pragma unroll
    for (int k = 0; k < 9; ++k) dW[k][0] = dW[k][1] = make_float2(0.f, 0.f);
    float2 s1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    const uint32_t row_bytes = static_cast<uint32_t>(p.owp) * CB * 2;
    const uint32_t w_mine = w_smem + l4 * kCh * 4;
    int r = (pt * p.mg_s) >> 16, s = pt - r * p.nstrips;
    const int step_r = p.cs_r, step_s = p.cs_s;
    // The producer's raw output at the strip's 4 pixels comes through a two-deep per-thread cp.async ring: the loads of
    // strip i+1 are in flight while strip i is computed (no registers held, no exposed global-memory latency).
    const uint32_t zring = tile_smem + static_cast<uint32_t>(p.oh) * row_bytes + threadIdx.x * 8;     // [2][kStrip][THREADS] x 8 bytes
    auto strip_ok = [&](int rr, int ss) { return rr < p.th && iy0 + rr < p.H && ss < p.nstrips; };
    auto prefetch = [&](int rr, int ss, int stage) {
        const int lx0 = ss * kStrip;
        const act_t* zrow = p.zin + ((static_cast<long long>(n) * p.H + iy0 + rr) * p.W + ix0 + lx0) * p.C + c0;
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            if ((lx0 + a < p.twt) && (ix0 + lx0 + a < p.W))
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(zring + (stage * kStrip + a) * (THREADS * 8)),
                             "l"(zrow + static_cast<long long>(a) * p.C) : "memory");
        }
    };
    if (strip_ok(r, s)) prefetch(r, s, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int stage = 0; strip_ok(r, s); stage ^= 1) {
        int rn = r + step_r, sn = s + step_s;
        if (sn >= p.nstrips) { sn -= p.nstrips; ++rn; }
        if (strip_ok(rn, sn)) prefetch(rn, sn, stage ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const int iy = iy0 + r;
        const int lx0 = s * kStrip;
        const int s_cur = s;
        r = rn; s = sn;
        uint2 zraw[kStrip];
        bool pvalid[kStrip];
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            pvalid[a] = (lx0 + a < p.twt) && (ix0 + lx0 + a < p.W);
            zraw[a] = make_uint2(0u, 0u);
            if (pvalid[a]) zraw[a] = lds64(zring + (stage * kStrip + a) * (THREADS * 8));
        }
        float2 x[kStrip][2];
        uint32_t mask = 0;                                  // bit a*4+q: gradient passes at pixel a, channel q
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            float f[4] = {h16_lo(zraw[a].x), h16_hi(zraw[a].x), h16_lo(zraw[a].y), h16_hi(zraw[a].y)};
            if (p.in_scale) {
#pragma unroll
                for (int q = 0; q < kCh; ++q) {
                    const float pre = fmaf(f[q], isc[q], ish[q]);
                    const bool pass = p.in_act == 2 ? (pre > 0.f && pre < 6.f) : (p.in_act == 1 ? pre > 0.f : true);
                    if (pass && pvalid[a]) mask |= 1u << (a * 4 + q);
                    f[q] = act_apply(pre, p.in_act);
                }
                const uint2 pk = make_uint2(pack_h16(f[0], f[1]), pack_h16(f[2], f[3]));      // x as the forward stored it (fp16)
                x[a][0] = unpack2h(pk.x); x[a][1] = unpack2h(pk.y);
            } else {
                if (pvalid[a]) mask |= 0xfu << (a * 4);
                x[a][0] = make_float2(f[0], f[1]); x[a][1] = make_float2(f[2], f[3]);
            }
            if (!pvalid[a]) x[a][0] = x[a][1] = make_float2(0.f, 0.f);
        }
        float2 gx[kStrip][2];
#pragma unroll
        for (int a = 0; a < kStrip; ++a) gx[a][0] = gx[a][1] = make_float2(0.f, 0.f);
        const uint32_t strip_base = tile_smem + (s_cur * Cols::STRIP_COLS * CB + l4 * kCh) * 2;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ay = iy + p.pad_top - ky * D;
            if (S == 2 && (ay & 1)) continue;
            const int orow = (S == 1 ? ay : (ay >> 1)) - oy_lo;            // in [0, oh): rows outside the image are staged zeros
            const uint32_t rowp = strip_base + static_cast<uint32_t>(orow) * row_bytes;
            float2 wv[3][2];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float4 wf;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(wf.x), "=f"(wf.y), "=f"(wf.z), "=f"(wf.w)
                             : "r"(w_mine + (ky * 3 + kx) * (CB * 4)));
                wv[kx][0] = make_float2(wf.x, wf.y); wv[kx][1] = make_float2(wf.z, wf.w);
            }
#pragma unroll
            for (int j = 0; j < Cols::NJ; ++j) {
                const uint2 raw = lds64(rowp + j * (CB * 2));
                const float2 v0 = unpack2(raw.x), v1 = unpack2(raw.y);
#pragma unroll
                for (int a = 0; a < kStrip; ++a) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        if (Cols::col(a, kx) == j) {
                            ffma2(gx[a][0], v0, wv[kx][0]); ffma2(gx[a][1], v1, wv[kx][1]);
                            ffma2(dW[ky * 3 + kx][0], x[a][0], v0); ffma2(dW[ky * 3 + kx][1], x[a][1], v1);
                        }
                    }
                }
            }
        }
        bf16* orow_p = p.gout + ((static_cast<long long>(n) * p.H + iy) * p.W + ix0 + lx0) * p.C + c0;
#pragma unroll
        for (int a = 0; a < kStrip; ++a) {
            if (!pvalid[a]) continue;
            float f[4] = {gx[a][0].x, gx[a][0].y, gx[a][1].x, gx[a][1].y};
#pragma unroll
            for (int q = 0; q < kCh; ++q) if (!((mask >> (a * 4 + q)) & 1u)) f[q] = 0.f;
            const uint2 pk = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
            *reinterpret_cast<uint2*>(orow_p + static_cast<long long>(a) * p.C) = pk;
            if (p.bn_partial) {
                const float2 g0 = unpack2(pk.x), g1 = unpack2(pk.y);
                const float2 z0 = unpack2h(zraw[a].x), z1 = unpack2h(zraw[a].y);
                fadd2(s1[0], g0); fadd2(s1[1], g1);
                ffma2(s2[0], g0, z0); ffma2(s2[1], g1, z1);
            }
        }
    }

    // ---------------------------------------------------------------- fixed-order block reductions -> per-tile partials
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);                // [NPT][11][CB]
    {
        float* mine = red + static_cast<size_t>(pt) * 11 * CB + l4 * kCh;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            mine[k * CB + 0] = dW[k][0].x; mine[k * CB + 1] = dW[k][0].y; mine[k * CB + 2] = dW[k][1].x; mine[k * CB + 3] = dW[k][1].y;
        }
        mine[9 * CB + 0] = s1[0].x; mine[9 * CB + 1] = s1[0].y; mine[9 * CB + 2] = s1[1].x; mine[9 * CB + 3] = s1[1].y;
        mine[10 * CB + 0] = s2[0].x; mine[10 * CB + 1] = s2[0].y; mine[10 * CB + 2] = s2[1].x; mine[10 * CB + 3] = s2[1].y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 11 * CB; i += THREADS) {
        const int k = i / CB, c = i - k * CB;
        if (k < 9) {
            float a = 0.f;
            for (int q = 0; q < NPT; ++q) a += red[static_cast<size_t>(q) * 11 * CB + i];
            p.dw_partial[(tile * 9 + k) * p.C + c_base + c] = a;
        } else if (p.bn_partial) {
            double a = 0.0;
            for (int q = 0; q < NPT; ++q) a += static_cast<double>(red[static_cast<size_t>(q) * 11 * CB + i]);
            p.bn_partial[(tile * 2 + (k - 9)) * p.C + c_base + c] = a;
        }
    }
}

// out[i] = sum over rows of partial[row][i]; block = 32 outputs x 32 row-lanes (coalesced), fixed order => deterministic
__global__ void __launch_bounds__(1024)
dw_reduce_rows_kernel(const float* __restrict__ partial, int rows, int n, float* __restrict__ out) {
    pdl_entry();
    __shared__ double s_s[32][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cl;
    double a = 0.0;
    if (i < n) {
#pragma unroll 4
        for (int r = rl; r < rows; r += 32) a += static_cast<double>(partial[static_cast<long long>(r) * n + i]);
    }
    s_s[rl][cl] = a;
    __syncthreads();
    if (rl != 0 || i >= n) return;
    for (int k = 1; k < 32; ++k) a += s_s[k][cl];
    out[i] = static_cast<float>(a);
}

struct DwBwdTile { int th, twt, ntx, nty, cb, chunks, nstrips, oh, owp; size_t smem; long long blocks; };

DwBwdTile pick_bwd_tile(const Conv2dGeom& g, size_t smem_cap) {
    DwBwdTile best{};
    double best_cost = 1e300;
    const int cb = pick_cb(g.C);
    if (!cb) return best;
    const int S = g.stride, D = g.dil;
    const int cv4 = cb / kCh;
    const int npt = 256 / cv4;
    const size_t fixed = 9 * static_cast<size_t>(cb) * 4;
    for (int th = 2; th <= 32; th += (S == 2 ? 2 : 1)) {
        const int nty = ceil_div(g.H, th);
        for (int ntx = 1; ntx <= 40; ++ntx) {
            int twt = ceil_div(g.W, ntx);
            if (S == 2) twt += twt & 1;                         // even tile origins keep the tap parity compile-time
            if (twt < 8 && ntx > 1) break;
            if (ceil_div(g.W, twt) != ntx) continue;
            const int nstrips = ceil_div(twt, kStrip);
            const int oh = S == 1 ? th + 2 * D : th / 2 + 1;
            const int owp = S == 1 ? nstrips * kStrip + 2 * D : nstrips * 2 + 1;
            size_t smem = fixed + static_cast<size_t>(oh) * owp * cb * 2 + 2 * kStrip * dw_threads(cb) * 8;      // + the cp.async ring
            if (smem > smem_cap) continue;
            smem = std::max(smem, static_cast<size_t>(npt) * 11 * cb * sizeof(float));
            const int iters = ceil_div(th * nstrips, npt);
            const long long blocks = static_cast<long long>(g.N) * nty * ntx * (g.C / cb);
            const double compute = static_cast<double>(iters) * npt * kStrip * 2.0;
            const double stage = static_cast<double>(oh) * owp * 1.2;
            const double waves = static_cast<double>(ceil_div_ll(blocks, kNumSMs * 2LL));
            const double cost = waves * (compute + stage + 400.0);
            if (cost < best_cost) {
                best_cost = cost;
                best = DwBwdTile{th, twt, ntx, nty, cb, g.C / cb, nstrips, oh, owp, smem, blocks};
            }
        }
    }
    return best;
}

static const size_t kBwdSmemCap = env_kb("AMS_DWB_SMEM_KB", 112);

template <int S, int D, int CB, int PADX>
int launch_bwd(const DwBwdParams& p, const DwBwdTile& t, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_bwd_fused_kernel<S, D, CB, PADX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        attr = true;
    }
    AMS_LAUNCH((dw_bwd_fused_kernel<S, D, CB, PADX>), static_cast<unsigned>(t.blocks), dw_threads(CB), t.smem, s, p);
    return 0;
}
template <int S, int D, int PADX>
int launch_bwd_cb(const DwBwdParams& p, const DwBwdTile& t, cudaStream_t s) {
    if (t.cb == 64) return launch_bwd<S, D, 64, PADX>(p, t, s);
    if (t.cb == 48) return launch_bwd<S, D, 48, PADX>(p, t, s);
    AMS_REQUIRE(t.cb == 32, "fused depthwise backward: channel chunk not instantiated");
    return launch_bwd<S, D, 32, PADX>(p, t, s);
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

extern "C" int ams_debug_dw_bwd_tile(int N, int H, int W, int C, int Ho, int Wo, int stride, int dil, int* out8) {
    Conv2dGeom g{N, H, W, C, Ho, Wo, stride, dil, dil, dil};
    const DwBwdTile t = pick_bwd_tile(g, kBwdSmemCap);
    out8[0] = t.th; out8[1] = t.twt; out8[2] = t.ntx; out8[3] = t.nty; out8[4] = t.cb; out8[5] = static_cast<int>(t.smem);
    out8[6] = static_cast<int>(t.blocks); out8[7] = t.nstrips;
    return 0;
}

bool dw_tiled_supported(const Conv2dGeom& g) {
    return pick_cb(g.C) != 0 && ((g.stride == 1 && (g.dil == 1 || g.dil == 2)) || (g.stride == 2 && g.dil == 1));
}

long long dw_tiled_stats_rows(const Conv2dGeom& g) {
    const DwTile t = pick_tile(g, kFwdSmemCap);
    return static_cast<long long>(g.N) * t.nty * t.ntx;
}

int dw_conv_fwd_tiled(const act_t* in, const float* w, const Conv2dGeom& g, const float* in_scale, const float* in_shift,
                      int in_act, const float* out_scale, const float* out_shift, int out_act, act_t* out, double* stats,
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
    {
        const int pxt = dw_threads(t.cb) / (t.cb / 8), npt = dw_threads(t.cb) / (t.cb / kCh);
        p.mg_w = 65536 / t.iwp + 1; p.mg_s = 65536 / t.nstrips + 1;
        p.st_y = pxt / t.iwp; p.st_x = pxt - p.st_y * t.iwp;
        p.cs_r = npt / t.nstrips; p.cs_s = npt - p.cs_r * t.nstrips;
    }
    if (stats_rows) *stats_rows = static_cast<int>(static_cast<long long>(g.N) * t.nty * t.ntx);
    if (g.stride == 1 && g.dil == 1) return launch_fwd_cb<1, 1>(p, t, s);
    if (g.stride == 2 && g.dil == 1) return launch_fwd_cb<2, 1>(p, t, s);
    return launch_fwd_cb<1, 2>(p, t, s);
}


long long dw_bwd_fused_rows(const Conv2dGeom& g) {
    const DwBwdTile t = pick_bwd_tile(g, kBwdSmemCap);
    return static_cast<long long>(g.N) * t.nty * t.ntx;
}

int dw_conv_bwd_reduce(const DwBwdFused& a, const Conv2dGeom& g, cudaStream_t s) {
    const DwBwdTile t = pick_bwd_tile(g, kBwdSmemCap);
    const long long rows = static_cast<long long>(g.N) * t.nty * t.ntx;
    AMS_LAUNCH((dw_reduce_rows_kernel), ceil_div(9 * g.C, 32), 1024, 0, s, a.dw_partial, static_cast<int>(rows), 9 * g.C, a.dw);
    return 0;
}

int dw_conv_bwd_fused(const DwBwdFused& a, const Conv2dGeom& g, int* rows_out, cudaStream_t s, bool defer_reduce) {
    AMS_REQUIRE(dw_tiled_supported(g), "fused depthwise backward: unsupported channels / stride / dilation");
    const DwBwdTile t = pick_bwd_tile(g, kBwdSmemCap);
    AMS_REQUIRE(t.blocks > 0, "fused depthwise backward: no tile fits shared memory");
    const long long rows = static_cast<long long>(g.N) * t.nty * t.ntx;
    AMS_REQUIRE(a.dw_partial_floats >= static_cast<size_t>(rows) * 9 * g.C, "fused depthwise backward: dW workspace too small");
    DwBwdParams p;
    p.g = a.g; p.z = a.z; p.scale = a.scale; p.shift = a.shift; p.act = a.act; p.coef = a.coef;
    p.zin = a.zin; p.in_scale = a.in_scale; p.in_shift = a.in_shift; p.in_act = a.in_act; p.w = a.w;
    p.gout = a.gout; p.dw_partial = a.dw_partial; p.bn_partial = a.in_scale ? a.bn_partial : nullptr;
    p.N = g.N; p.H = g.H; p.W = g.W; p.C = g.C; p.Ho = g.Ho; p.Wo = g.Wo; p.pad_top = g.pad_top; p.pad_left = g.pad_left;
    p.th = t.th; p.twt = t.twt; p.ntx = t.ntx; p.nty = t.nty; p.chunks = t.chunks; p.nstrips = t.nstrips; p.oh = t.oh; p.owp = t.owp;
    {
        const int pxt = dw_threads(t.cb) / (t.cb / 8), npt = dw_threads(t.cb) / (t.cb / kCh);
        p.mg_w = 65536 / t.owp + 1; p.mg_s = 65536 / t.nstrips + 1;
        p.st_y = pxt / t.owp; p.st_x = pxt - p.st_y * t.owp;
        p.cs_r = npt / t.nstrips; p.cs_s = npt - p.cs_r * t.nstrips;
    }
    int rc;
    if (g.stride == 1 && g.dil == 1) rc = launch_bwd_cb<1, 1, 0>(p, t, s);
    else if (g.stride == 1) rc = launch_bwd_cb<1, 2, 0>(p, t, s);
    else if (g.pad_left & 1) rc = launch_bwd_cb<2, 1, 1>(p, t, s);
    else rc = launch_bwd_cb<2, 1, 0>(p, t, s);
    if (rc) return rc;
    if (!defer_reduce) AMS_LAUNCH((dw_reduce_rows_kernel), ceil_div(9 * g.C, 32), 1024, 0, s, a.dw_partial, static_cast<int>(rows), 9 * g.C, a.dw);
    if (rows_out) *rows_out = static_cast<int>(rows);
    return 0;
}

}  // namespace ams
