// Direct (CUDA-core) convolution kernels: the 3x3 stem (Cin=3: K=27 is too small for tensor cores) and the
// 17 depthwise 3x3 layers with their gradients.  All are HBM-bound (AI 1.8-19.7 flop/B):
// 128-bit NHWC channel-vector accesses, register sliding windows, no materialised padding.
// Replaces: `MobilenetV2/Conv/Conv2D`, the `DepthwiseConv2dNative` nodes (+ SpaceToBatchND/BatchToSpaceND
// atrous wrappers) of checkpoints/*/model.meta and their TF-generated gradients.
#include "kernels.cuh"

namespace ams {
namespace {

// ============================================================================================ stem
template <typename T> __device__ __forceinline__ float load_px(const T* p);
template <> __device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) { return static_cast<float>(*p); }
template <> __device__ __forceinline__ float load_px<float>(const float* p) { return *p; }

struct StemGeom {
    int N, H, W, Hp, Wp, Ho, Wo, pad_top, pad_left;
    float pad_value, norm_scale, norm_shift;
};

// value of the padded + normalised input at padded coordinate (y, x) of image n, channel c
template <typename T>
__device__ __forceinline__ void stem_pixel(const T* in, const StemGeom& g, int n, int y, int x, float* v3) {
    if (y < 0 || x < 0 || y >= g.Hp || x >= g.Wp) { v3[0] = v3[1] = v3[2] = 0.f; return; }   // conv zero pad
    if (y >= g.H || x >= g.W) {                                                               // graph mean-pixel pad
        const float t = __fsub_rn(__fmul_rn(g.norm_scale, g.pad_value), g.norm_shift);
        v3[0] = v3[1] = v3[2] = t;
        return;
    }
    const T* p = in + (static_cast<long long>(n) * g.H * g.W + static_cast<long long>(y) * g.W + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v3[c] = __fsub_rn(__fmul_rn(g.norm_scale, load_px<T>(p + c)), g.norm_shift);
}

// one thread = one output pixel x all 32 output channels.  The kernel is bound by instruction issue, not by its 80 MB of
// traffic: the 27 input values are fetched and normalised once per pixel, and the 27 x 32 FMAs run as packed fp32x2
// instructions (two output channels each; every lane is the same IEEE fma, in the same tap order, as a scalar loop).
template <typename T>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const T* __restrict__ in, StemGeom g, const float* __restrict__ w, const float* __restrict__ scale,
                const float* __restrict__ shift, act_t* __restrict__ out, float act_hi) {
    pdl_entry();
    __shared__ __align__(16) float sw[27 * 32];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.N) * g.Ho * g.Wo;
    if (pix >= total) return;
    const int ox = static_cast<int>(pix % g.Wo);
    const int oy = static_cast<int>((pix / g.Wo) % g.Ho);
    const int n = static_cast<int>(pix / (static_cast<long long>(g.Wo) * g.Ho));
    float2 acc2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc2[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            float v[3];
            stem_pixel<T>(in, g, n, oy * 2 - g.pad_top + ky, ox * 2 - g.pad_left + kx, v);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4* wr = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + c) * 32);
                const float2 vv = make_float2(v[c], v[c]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w4 = wr[j];
                    const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&acc2[2 * j]))
                        : "l"(*reinterpret_cast<const unsigned long long*>(&vv)), "l"(*reinterpret_cast<const unsigned long long*>(&wa)));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&acc2[2 * j + 1]))
                        : "l"(*reinterpret_cast<const unsigned long long*>(&vv)), "l"(*reinterpret_cast<const unsigned long long*>(&wb)));
                }
            }
        }
    }
    float acc[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    if (scale) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            acc[j] = fminf(fmaxf(fmaf(acc[j], __ldg(scale + j), __ldg(shift + j)), 0.f), act_hi);      // ReLU6 (6) or ReLU (+inf)
    }
    act_t* o = out + pix * 32;
    stg256(o, pack8h(acc), pack8h(acc + 8));
    stg256(o + 16, pack8h(acc + 16), pack8h(acc + 24));
}

// filter gradient: a block owns `rows_per_block` output rows; per segment of 64 output pixels it stages the three
// normalised input rows the segment touches ONCE (coalesced, pad values resolved there) plus the 64 x 32 dz values,
// then thread (tap k, 4 output channels) walks the pixels: one broadcast word + one 128-bit dz read per 4 FMAs.
constexpr int kStemBwdPix = 64;
constexpr int kStemBwdThreads = 27 * 8;
constexpr int kStemRowFloats = (2 * kStemBwdPix + 1) * 3 + 1;       // 129 input pixels x 3 channels (+1: odd stride)
template <typename T>
__global__ void __launch_bounds__(kStemBwdThreads)
stem_bwd_filter_kernel(const T* __restrict__ in, StemGeom g, const bf16* __restrict__ dz, float* __restrict__ partial,
                       int rows_per_block) {
    pdl_entry();
    __shared__ float s_in[3][kStemRowFloats];
    __shared__ __align__(16) float s_dz[kStemBwdPix][32];
    __shared__ float s_red[6][864];
    // thread = (pixel lane of 6, filter row ky, input channel c, group of 8 output channels): per pixel 3 input words + two
    // 128-bit dz reads feed 24 FMAs (the first version read one word + one 128-bit vector per 4 FMAs and kept the
    // shared-memory pipe 87 % busy: ncu "Mem Busy")
    const int cg = threadIdx.x & 3, c = (threadIdx.x >> 2) % 3, ky = (threadIdx.x / 12) % 3, pl = threadIdx.x / 36;
    float acc[3][8];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[kx][q] = 0.f;
    const int total_rows = g.N * g.Ho;
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(r_begin + rows_per_block, total_rows);
    const float pad_t = __fsub_rn(__fmul_rn(g.norm_scale, g.pad_value), g.norm_shift);
    for (int row = r_begin; row < r_end; ++row) {
        const int n = row / g.Ho, oy = row - n * g.Ho;
        for (int ox0 = 0; ox0 < g.Wo; ox0 += kStemBwdPix) {
            const int np = min(kStemBwdPix, g.Wo - ox0);
            __syncthreads();
            // input rows 2*oy - pad + {0,1,2}, columns 2*ox0 - pad .. + 2*np
            const int ncol = 2 * np + 1;
            for (int i = threadIdx.x; i < 3 * ncol; i += kStemBwdThreads) {
                const int r3 = i / ncol, cx = i - r3 * ncol;
                const int y = oy * 2 - g.pad_top + r3, x = ox0 * 2 - g.pad_left + cx;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f;                     // conv zero pad
                if (y >= 0 && x >= 0 && y < g.Hp && x < g.Wp) {
                    if (y >= g.H || x >= g.W) { v0 = v1 = v2 = pad_t; }   // graph mean-pixel pad
                    else {
                        const T* px = in + (static_cast<long long>(n) * g.H * g.W + static_cast<long long>(y) * g.W + x) * 3;
                        v0 = __fsub_rn(__fmul_rn(g.norm_scale, load_px<T>(px)), g.norm_shift);
                        v1 = __fsub_rn(__fmul_rn(g.norm_scale, load_px<T>(px + 1)), g.norm_shift);
                        v2 = __fsub_rn(__fmul_rn(g.norm_scale, load_px<T>(px + 2)), g.norm_shift);
                    }
                }
                s_in[r3][cx * 3] = v0; s_in[r3][cx * 3 + 1] = v1; s_in[r3][cx * 3 + 2] = v2;
            }
            const long long p0 = (static_cast<long long>(n) * g.Ho + oy) * g.Wo + ox0;
            for (int i = threadIdx.x; i < np * 4; i += kStemBwdThreads) {
                const int pp = i >> 2, c8 = i & 3;
                float f[8];
                unpack8(ldg_stream(dz + (p0 + pp) * 32 + c8 * 8), f);
                *reinterpret_cast<float4*>(&s_dz[pp][c8 * 8]) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(&s_dz[pp][c8 * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
            }
            __syncthreads();
            const float* a_ptr = &s_in[ky][c];
#pragma unroll 2
            for (int pp = pl; pp < np; pp += 6) {
                const float a0 = a_ptr[pp * 6], a1 = a_ptr[pp * 6 + 3], a2 = a_ptr[pp * 6 + 6];
                const float4 d0 = *reinterpret_cast<const float4*>(&s_dz[pp][cg * 8]);
                const float4 d1 = *reinterpret_cast<const float4*>(&s_dz[pp][cg * 8 + 4]);
                const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    acc[0][q] = fmaf(a0, d[q], acc[0][q]);
                    acc[1][q] = fmaf(a1, d[q], acc[1][q]);
                    acc[2][q] = fmaf(a2, d[q], acc[2][q]);
                }
            }
        }
    }
    // the six pixel lanes are summed in a fixed order; partial[block][k = ky*9 + kx*3 + c][32]
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int q = 0; q < 8; ++q) s_red[pl][(ky * 9 + kx * 3 + c) * 32 + cg * 8 + q] = acc[kx][q];
    __syncthreads();
    for (int i = threadIdx.x; i < 864; i += kStemBwdThreads) {
        float t = s_red[0][i];
#pragma unroll
        for (int l = 1; l < 6; ++l) t += s_red[l][i];
        partial[static_cast<long long>(blockIdx.x) * 864 + i] = t;
    }
}

// out[i] = sum over chunks of partial[chunk][i]; one warp per output, fixed lane assignment (deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int chunks) {
    pdl_entry();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    double acc = 0.0;
    for (int c = lane; c < chunks; c += 32) acc += static_cast<double>(partial[static_cast<long long>(c) * n + i]);
    acc = warp_sum_d(acc);
    if (lane == 0) out[i] = static_cast<float>(acc);
}

// ============================================================================================ depthwise
// thread = 8 channels x TW consecutive output pixels along W; sliding window over the needed input columns
template <int S, int D, int TW>
__global__ void __launch_bounds__(256)
dw_fwd_kernel(const act_t* __restrict__ in, const float* __restrict__ w, Conv2dGeom g, const float* __restrict__ scale,
              const float* __restrict__ shift, int act, act_t* __restrict__ out) {
    pdl_entry();
    const int C8 = g.C >> 3;
    const int WG = (g.Wo + TW - 1) / TW;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.N) * g.Ho * WG * C8;
    if (tid >= total) return;
    const int c8 = static_cast<int>(tid % C8);
    long long r = tid / C8;
    const int xg = static_cast<int>(r % WG); r /= WG;
    const int oy = static_cast<int>(r % g.Ho);
    const int n = static_cast<int>(r / g.Ho);
    const int ox0 = xg * TW;
    const int c0 = c8 * 8;
    constexpr int NCOLS = (TW - 1) * S + 2 * D + 1;
    float acc[TW][8];
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    const int ix0 = ox0 * S - g.pad_left;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * S - g.pad_top + ky * D;
        if (iy < 0 || iy >= g.H) continue;
        float wk[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4 a = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * g.C + c0);
            const float4 b = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * g.C + c0 + 4);
            wk[kx][0] = a.x; wk[kx][1] = a.y; wk[kx][2] = a.z; wk[kx][3] = a.w;
            wk[kx][4] = b.x; wk[kx][5] = b.y; wk[kx][6] = b.z; wk[kx][7] = b.w;
        }
        const act_t* row = in + ((static_cast<long long>(n) * g.H + iy) * g.W) * g.C + c0;
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) {
            // does any (t, kx) use column j?  t*S + kx*D == j
            bool used = false;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) used = used || ((j - kx * D) >= 0 && (j - kx * D) % S == 0 && (j - kx * D) / S < TW);
            if (!used) continue;
            const int ix = ix0 + j;
            if (ix < 0 || ix >= g.W) continue;
            float v[8];
            unpack8h(__ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ix) * g.C)), v);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int tt = j - kx * D;
                if (tt >= 0 && tt % S == 0 && tt / S < TW) {
                    const int t = tt / S;
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[t][q] = fmaf(v[q], wk[kx][q], acc[t][q]);
                }
            }
        }
    }
    float sc[8], sh[8];
    if (scale) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { sc[q] = scale[c0 + q]; sh[q] = shift[c0 + q]; }
    }
#pragma unroll
    for (int t = 0; t < TW; ++t) {
        const int ox = ox0 + t;
        if (ox >= g.Wo) break;
        if (scale) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[t][q] = act_apply(fmaf(acc[t][q], sc[q], sh[q]), act);
        }
        stg_stream(out + ((static_cast<long long>(n) * g.Ho + oy) * g.Wo + ox) * g.C + c0, pack8h(acc[t]));
    }
}

// dx[iy,ix,c] = sum_{ky,kx} dz[(iy+pt-ky*D)/S, (ix+pl-kx*D)/S, c] * w[ky,kx,c]   (where divisible and in range)
template <int S, int D>
__global__ void __launch_bounds__(256)
dw_bwd_data_kernel(const bf16* __restrict__ dz, const float* __restrict__ w, Conv2dGeom g, bf16* __restrict__ dx) {
    pdl_entry();
    const int C8 = g.C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.N) * g.H * g.W * C8;
    if (tid >= total) return;
    const int c0 = static_cast<int>(tid % C8) * 8;
    long long r = tid / C8;
    const int ix = static_cast<int>(r % g.W); r /= g.W;
    const int iy = static_cast<int>(r % g.H);
    const int n = static_cast<int>(r / g.H);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int ty = iy + g.pad_top - ky * D;
        if (ty < 0 || ty % S != 0) continue;
        const int oy = ty / S;
        if (oy >= g.Ho) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int tx = ix + g.pad_left - kx * D;
            if (tx < 0 || tx % S != 0) continue;
            const int ox = tx / S;
            if (ox >= g.Wo) continue;
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dz + ((static_cast<long long>(n) * g.Ho + oy) * g.Wo + ox) * g.C + c0)), v);
            const float4 a = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * g.C + c0);
            const float4 b = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * g.C + c0 + 4);
            acc[0] = fmaf(v[0], a.x, acc[0]); acc[1] = fmaf(v[1], a.y, acc[1]);
            acc[2] = fmaf(v[2], a.z, acc[2]); acc[3] = fmaf(v[3], a.w, acc[3]);
            acc[4] = fmaf(v[4], b.x, acc[4]); acc[5] = fmaf(v[5], b.y, acc[5]);
            acc[6] = fmaf(v[6], b.z, acc[6]); acc[7] = fmaf(v[7], b.w, acc[7]);
        }
    }
    stg_stream(dx + ((static_cast<long long>(n) * g.H + iy) * g.W + ix) * g.C + c0, pack8(acc));
}

// dW[ky,kx,c] = sum_{n,oy,ox} x[iy,ix,c] * dz[oy,ox,c].
// thread = one 8-channel group x a strip of TW consecutive output pixels (register sliding window over the input
// columns the strip touches); a block holds 256/min(C/8,256) strips side by side, each block owns a contiguous run
// of strips, partial [block][9][C] is reduced in fixed order afterwards (deterministic).
constexpr int kDwfTW = 4;
template <int S, int D>
__global__ void __launch_bounds__(256, 2)
dw_bwd_filter_kernel(const act_t* __restrict__ x, const bf16* __restrict__ dz, Conv2dGeom g, float* __restrict__ partial,
                     int strips_per_block) {
    pdl_entry();
    extern __shared__ float s_red[];                       // [rows_in_block][tpr][73]
    constexpr int TW = kDwfTW;
    constexpr int NCOLS = (TW - 1) * S + 2 * D + 1;
    const int C8 = g.C >> 3;
    const int tpr = min(C8, 256);
    const int rows_in_block = 256 / tpr;
    const int lr = threadIdx.x / tpr, lc = threadIdx.x % tpr;
    const int WG = (g.Wo + TW - 1) / TW;
    const long long total = static_cast<long long>(g.N) * g.Ho * WG;
    const long long s_begin = static_cast<long long>(blockIdx.x) * strips_per_block;
    const long long s_end = min(s_begin + strips_per_block, total);
    const int c0 = lc * 8;
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[t][q] = 0.f;
    if (lr < rows_in_block) {
        for (long long sidx = s_begin + lr; sidx < s_end; sidx += rows_in_block) {
            const int xg = static_cast<int>(sidx % WG);
            const int oy = static_cast<int>((sidx / WG) % g.Ho);
            const int n = static_cast<int>(sidx / (static_cast<long long>(WG) * g.Ho));
            const int ox0 = xg * TW;
            float d[TW][8];
#pragma unroll
            for (int t = 0; t < TW; ++t) {
                if (ox0 + t < g.Wo) unpack8(ldg_stream(dz + ((static_cast<long long>(n) * g.Ho + oy) * g.Wo + ox0 + t) * g.C + c0), d[t]);
                else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[t][q] = 0.f;
                }
            }
            const int ix0 = ox0 * S - g.pad_left;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = oy * S - g.pad_top + ky * D;
                if (iy < 0 || iy >= g.H) continue;
                const act_t* row = x + ((static_cast<long long>(n) * g.H + iy) * g.W) * g.C + c0;
#pragma unroll
                for (int j = 0; j < NCOLS; ++j) {
                    bool used = false;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) used = used || ((j - kx * D) >= 0 && (j - kx * D) % S == 0 && (j - kx * D) / S < TW);
                    if (!used) continue;
                    const int ix = ix0 + j;
                    if (ix < 0 || ix >= g.W) continue;
                    float v[8];
                    unpack8h(__ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ix) * g.C)), v);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int tt = j - kx * D;
                        if (tt >= 0 && tt % S == 0 && tt / S < TW) {
                            const int t = tt / S;
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc[ky * 3 + kx][q] = fmaf(v[q], d[t][q], acc[ky * 3 + kx][q]);
                        }
                    }
                }
            }
        }
        float* dst = s_red + (static_cast<size_t>(lr) * tpr + lc) * 73;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[t * 8 + q] = acc[t][q];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < tpr * 72; e += 256) {
        const int cc = e / 72, k = e % 72;
        float sum = 0.f;
        for (int l = 0; l < rows_in_block; ++l) sum += s_red[(static_cast<size_t>(l) * tpr + cc) * 73 + k];
        partial[(static_cast<long long>(blockIdx.x) * 9 + k / 8) * g.C + cc * 8 + (k % 8)] = sum;
    }
}

int blocks_for(long long total, int threads) { return static_cast<int>((total + threads - 1) / threads); }

constexpr int kStemRowsPerBlock = 1;        // ~2k small CTAs, 9 resident per SM: the per-segment staging round trips overlap
int dw_strips_per_block(const Conv2dGeom& g) {
    const long long total = static_cast<long long>(g.N) * g.Ho * ceil_div(g.Wo, kDwfTW);
    const int rows_in_block = 256 / std::min(g.C / 8, 256);
    const long long want_blocks = 4LL * kNumSMs;
    long long spb = std::max<long long>(rows_in_block * 4LL, ceil_div_ll(total, want_blocks));
    return static_cast<int>(spb);
}

}  // namespace

// ============================================================================================ host
int stem_conv_fwd(const void* in, int in_is_u8, int N, int H, int W, int Hp, int Wp, int Ho, int Wo, int pad_top,
                  int pad_left, float pad_value, float norm_scale, float norm_shift, const float* w, const float* scale,
                  const float* shift, act_t* out, cudaStream_t s, int act) {
    StemGeom g{N, H, W, Hp, Wp, Ho, Wo, pad_top, pad_left, pad_value, norm_scale, norm_shift};
    const long long total = static_cast<long long>(N) * Ho * Wo;
    const float act_hi = act == 1 ? INFINITY : 6.f;
    if (in_is_u8)
        AMS_LAUNCH((stem_fwd_kernel<uint8_t>), blocks_for(total, 256), 256, 0, s, static_cast<const uint8_t*>(in), g, w, scale, shift, out, act_hi);
    else
        AMS_LAUNCH((stem_fwd_kernel<float>), blocks_for(total, 256), 256, 0, s, static_cast<const float*>(in), g, w, scale, shift, out, act_hi);
    return 0;
}

size_t stem_bwd_workspace_floats(int N, int Ho, int Wo) {
    (void)Wo;
    return static_cast<size_t>(ceil_div(N * Ho, kStemRowsPerBlock)) * 864;
}

int stem_conv_bwd_filter(const void* in, int in_is_u8, int N, int H, int W, int Hp, int Wp, int Ho, int Wo, int pad_top,
                         int pad_left, float pad_value, float norm_scale, float norm_shift, const bf16* dz, float* dw,
                         float* workspace, size_t workspace_floats, cudaStream_t s) {
    StemGeom g{N, H, W, Hp, Wp, Ho, Wo, pad_top, pad_left, pad_value, norm_scale, norm_shift};
    const int chunks = ceil_div(N * Ho, kStemRowsPerBlock);
    AMS_REQUIRE(workspace_floats >= static_cast<size_t>(chunks) * 864, "stem bwd workspace too small");
    if (in_is_u8)
        AMS_LAUNCH((stem_bwd_filter_kernel<uint8_t>), chunks, kStemBwdThreads, 0, s, static_cast<const uint8_t*>(in), g, dz, workspace, kStemRowsPerBlock);
    else
        AMS_LAUNCH((stem_bwd_filter_kernel<float>), chunks, kStemBwdThreads, 0, s, static_cast<const float*>(in), g, dz, workspace, kStemRowsPerBlock);
    AMS_LAUNCH((reduce_partials_kernel), ceil_div(864 * 32, 256), 256, 0, s, workspace, dw, 864, chunks);
    return 0;
}

int dw_conv_fwd(const act_t* in, const float* w, const Conv2dGeom& g, const float* scale, const float* shift, int act,
                act_t* out, cudaStream_t s) {
    AMS_REQUIRE(g.C % 8 == 0, "depthwise channels must be a multiple of 8");
    constexpr int TW = 4;
    const long long total = static_cast<long long>(g.N) * g.Ho * ceil_div(g.Wo, TW) * (g.C / 8);
    const int nb = blocks_for(total, 256);
    if (g.stride == 1 && g.dil == 1) AMS_LAUNCH((dw_fwd_kernel<1, 1, TW>), nb, 256, 0, s, in, w, g, scale, shift, act, out);
    else if (g.stride == 2 && g.dil == 1) AMS_LAUNCH((dw_fwd_kernel<2, 1, TW>), nb, 256, 0, s, in, w, g, scale, shift, act, out);
    else if (g.stride == 1 && g.dil == 2) AMS_LAUNCH((dw_fwd_kernel<1, 2, TW>), nb, 256, 0, s, in, w, g, scale, shift, act, out);
    else AMS_REQUIRE(false, "unsupported depthwise stride/dilation");
    return 0;
}

int dw_conv_bwd_data(const bf16* dz, const float* w, const Conv2dGeom& g, bf16* dx, cudaStream_t s) {
    const long long total = static_cast<long long>(g.N) * g.H * g.W * (g.C / 8);
    const int nb = blocks_for(total, 256);
    if (g.stride == 1 && g.dil == 1) AMS_LAUNCH((dw_bwd_data_kernel<1, 1>), nb, 256, 0, s, dz, w, g, dx);
    else if (g.stride == 2 && g.dil == 1) AMS_LAUNCH((dw_bwd_data_kernel<2, 1>), nb, 256, 0, s, dz, w, g, dx);
    else if (g.stride == 1 && g.dil == 2) AMS_LAUNCH((dw_bwd_data_kernel<1, 2>), nb, 256, 0, s, dz, w, g, dx);
    else AMS_REQUIRE(false, "unsupported depthwise stride/dilation");
    return 0;
}

size_t dw_bwd_workspace_floats(const Conv2dGeom& g) {
    const long long total = static_cast<long long>(g.N) * g.Ho * ceil_div(g.Wo, kDwfTW);
    return static_cast<size_t>(ceil_div_ll(total, dw_strips_per_block(g))) * 9 * g.C;
}

int dw_conv_bwd_filter(const act_t* x, const bf16* dz, const Conv2dGeom& g, float* dw, float* workspace,
                       size_t workspace_floats, cudaStream_t s) {
    AMS_REQUIRE(g.C % 8 == 0 && g.C / 8 <= 256, "depthwise channels must be a multiple of 8, at most 2048");
    const long long total = static_cast<long long>(g.N) * g.Ho * ceil_div(g.Wo, kDwfTW);
    const int spb = dw_strips_per_block(g);
    const int chunks = static_cast<int>(ceil_div_ll(total, spb));
    AMS_REQUIRE(workspace_floats >= static_cast<size_t>(chunks) * 9 * g.C, "depthwise bwd workspace too small");
    const int tpr = std::min(g.C / 8, 256);
    const size_t smem = static_cast<size_t>(256 / tpr) * tpr * 73 * sizeof(float);
    static bool attr = false;
    if (!attr) {
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_bwd_filter_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_bwd_filter_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        AMS_CUDA_CHECK(cudaFuncSetAttribute(dw_bwd_filter_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        attr = true;
    }
    if (g.stride == 1 && g.dil == 1) AMS_LAUNCH((dw_bwd_filter_kernel<1, 1>), chunks, 256, smem, s, x, dz, g, workspace, spb);
    else if (g.stride == 2 && g.dil == 1) AMS_LAUNCH((dw_bwd_filter_kernel<2, 1>), chunks, 256, smem, s, x, dz, g, workspace, spb);
    else if (g.stride == 1 && g.dil == 2) AMS_LAUNCH((dw_bwd_filter_kernel<1, 2>), chunks, 256, smem, s, x, dz, g, workspace, spb);
    else AMS_REQUIRE(false, "unsupported depthwise stride/dilation");
    AMS_LAUNCH((reduce_partials_kernel), ceil_div(9 * g.C * 32, 256), 256, 0, s, workspace, dw, 9 * g.C, chunks);
    return 0;
}

}  // namespace ams
