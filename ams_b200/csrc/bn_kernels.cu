// BatchNorm kernels (training-mode statistics, apply, backward) and the small fp32 kernels of the ASPP
// image-pooling branch.  Everything here is a streaming reduction or elementwise pass over bf16 NHWC
// activations: 128-bit loads, per-thread fp32 runs flushed into fp64 accumulators, fixed-order (deterministic)
// cross-block reduction -- no floating-point atomics anywhere (SURVEY 7 'bit-exact top-k' hard part).
// Replaces: the 54 `FusedBatchNormV3(is_training=True)` nodes + 108 `AssignSub` moving-average updates of
// checkpoints/*/model.meta and their gradients; inference-mode patch BN of utils/graph_utils.py:362-369.
#include "kernels.cuh"

namespace ams {
namespace {

constexpr int kRedThreads = 256;
constexpr int kFlush = 16;          // fp32 run length before flushing into fp64

int red_chunks(long long M, int C) {
    // enough blocks to fill the machine, each with at least ~64 rows per thread-row
    const int c8 = C / 8;
    const int rows_per_block = std::max(1, kRedThreads / c8);
    long long want = std::max<long long>(1, std::min<long long>(4 * kNumSMs, M / (static_cast<long long>(rows_per_block) * 32)));
    return static_cast<int>(want);
}

// partial[chunk][0][c] = sum z, partial[chunk][1][c] = sum z^2 over the rows of the chunk
__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const bf16* __restrict__ z, long long M, int C, long long rows_per_chunk, double* __restrict__ partial) {
    extern __shared__ double s_acc[];                     // [rows_in_block][2][C]
    const int c8n = C >> 3;
    const int tpr = min(c8n, kRedThreads);                // threads per row
    const int rows_in_block = kRedThreads / tpr;
    const int lr = threadIdx.x / tpr, lc = threadIdx.x % tpr;
    const long long r_begin = static_cast<long long>(blockIdx.x) * rows_per_chunk;
    const long long r_end = min(r_begin + rows_per_chunk, M);
    for (int c8 = lc; c8 < c8n; c8 += tpr) {
        double ds[8], dq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { ds[q] = 0.0; dq[q] = 0.0; }
        if (lr < rows_in_block) {
            long long r = r_begin + lr;
            while (r < r_end) {
                float fs[8], fq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { fs[q] = 0.f; fq[q] = 0.f; }
#pragma unroll 4
                for (int i = 0; i < kFlush && r < r_end; ++i, r += rows_in_block) {
                    float v[8];
                    unpack8(ldg_stream(z + r * C + c8 * 8), v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) { fs[q] += v[q]; fq[q] = fmaf(v[q], v[q], fq[q]); }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) { ds[q] += fs[q]; dq[q] += fq[q]; }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                s_acc[(lr * 2 + 0) * C + c8 * 8 + q] = ds[q];
                s_acc[(lr * 2 + 1) * C + c8 * 8 + q] = dq[q];
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kRedThreads) {
        double a = 0.0;
        for (int l = 0; l < rows_in_block; ++l) a += s_acc[l * 2 * C + i];
        partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = a;
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ partial, int chunks, BnLayer L, int update_moving) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.C) return;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < chunks; ++k) {
        s += partial[static_cast<long long>(k) * 2 * L.C + c];
        q += partial[static_cast<long long>(k) * 2 * L.C + L.C + c];
    }
    const double n = static_cast<double>(L.M);
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float meanf = static_cast<float>(mean), varf = static_cast<float>(var);
    const float rstd = rsqrtf(varf + L.eps);
    const float sc = L.gamma[c] * rstd;
    L.mean[c] = meanf;
    L.rstd[c] = rstd;
    L.scale[c] = sc;
    L.shift[c] = L.beta[c] - meanf * sc;
    if (update_moving) {
        // AssignSub(mv, (mv - batch) * (1 - decay)); batch variance is Bessel-corrected (FusedBatchNormV3 output 2)
        const float unb = static_cast<float>(var * (n / fmax(n - 1.0, 1.0)));
        const float mm = L.moving_mean[c], mv = L.moving_var[c];
        L.moving_mean[c] = __fsub_rn(mm, __fmul_rn(__fsub_rn(mm, meanf), L.one_minus_decay));
        L.moving_var[c] = __fsub_rn(mv, __fmul_rn(__fsub_rn(mv, unb), L.one_minus_decay));
    }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const bf16* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                const bf16* __restrict__ residual, bf16* __restrict__ y, long long total8, int C) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total8) return;
    const int c0 = static_cast<int>((i * 8) % C);
    float v[8];
    unpack8(ldg_stream(z + i * 8), v);
    const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = act_apply(fmaf(v[q], sc[q], sh[q]), act);
    if (residual) {
        float r[8];
        unpack8(ldg_stream(residual + i * 8), r);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += r[q];
    }
    stg_stream(y + i * 8, pack8(v));
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mm, const float* mv, float eps,
                               float* scale, float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] * rsqrtf(mv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - mm[c] * sc;
}

// backward pass 1: partial[chunk][0][c] = sum g, [1][c] = sum g*z, g = (dy [+ dy2]) masked by the activation
__device__ __forceinline__ float act_mask(float g, float yhat, int act) {
    if (act == 1) return yhat > 0.f ? g : 0.f;
    if (act == 2) return (yhat > 0.f && yhat < 6.f) ? g : 0.f;
    return g;
}

__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2, const bf16* __restrict__ z,
                     const float* __restrict__ scale, const float* __restrict__ shift, int act, long long M, int C,
                     long long rows_per_chunk, double* __restrict__ partial) {
    extern __shared__ double s_acc[];
    const int c8n = C >> 3;
    const int tpr = min(c8n, kRedThreads);
    const int rows_in_block = kRedThreads / tpr;
    const int lr = threadIdx.x / tpr, lc = threadIdx.x % tpr;
    const long long r_begin = static_cast<long long>(blockIdx.x) * rows_per_chunk;
    const long long r_end = min(r_begin + rows_per_chunk, M);
    for (int c8 = lc; c8 < c8n; c8 += tpr) {
        double ds[8], dq[8];
        float sc[8], sh[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { ds[q] = 0.0; dq[q] = 0.0; sc[q] = scale[c8 * 8 + q]; sh[q] = shift[c8 * 8 + q]; }
        if (lr < rows_in_block) {
            long long r = r_begin + lr;
            while (r < r_end) {
                float fs[8], fq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { fs[q] = 0.f; fq[q] = 0.f; }
#pragma unroll 2
                for (int i = 0; i < kFlush && r < r_end; ++i, r += rows_in_block) {
                    float g[8], v[8];
                    unpack8(ldg_stream(dy + r * C + c8 * 8), g);
                    if (dy2) {
                        float g2[8];
                        unpack8(ldg_stream(dy2 + r * C + c8 * 8), g2);
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] += g2[q];
                    }
                    unpack8(ldg_stream(z + r * C + c8 * 8), v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float gm = act_mask(g[q], fmaf(v[q], sc[q], sh[q]), act);
                        fs[q] += gm;
                        fq[q] = fmaf(gm, v[q], fq[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) { ds[q] += fs[q]; dq[q] += fq[q]; }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                s_acc[(lr * 2 + 0) * C + c8 * 8 + q] = ds[q];
                s_acc[(lr * 2 + 1) * C + c8 * 8 + q] = dq[q];
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kRedThreads) {
        double a = 0.0;
        for (int l = 0; l < rows_in_block; ++l) a += s_acc[l * 2 * C + i];
        partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = a;
    }
}

// coef[0][c]=A, [1][c]=B, [2][c]=Cc with dz = A*g + B*z + Cc ; also d_gamma, d_beta
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partial, int chunks, BnLayer L, float* d_gamma,
                                       float* d_beta, float* coef) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.C) return;
    double s1 = 0.0, sz = 0.0;
    for (int k = 0; k < chunks; ++k) {
        s1 += partial[static_cast<long long>(k) * 2 * L.C + c];
        sz += partial[static_cast<long long>(k) * 2 * L.C + L.C + c];
    }
    const double mean = L.mean[c], rstd = L.rstd[c], gamma = L.gamma[c];
    const double n = static_cast<double>(L.M);
    const double s2 = rstd * (sz - mean * s1);           // sum g * xhat
    d_gamma[c] = static_cast<float>(s2);
    d_beta[c] = static_cast<float>(s1);
    const double A = gamma * rstd;
    const double B = -gamma * rstd * rstd * s2 / n;
    const double Cc = -gamma * rstd * (s1 / n - mean * rstd * s2 / n);
    coef[c] = static_cast<float>(A);
    coef[L.C + c] = static_cast<float>(B);
    coef[2 * L.C + c] = static_cast<float>(Cc);
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* dy, const bf16* __restrict__ dy2, const bf16* __restrict__ z,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
                    int act, long long total8, int C, bf16* dz_out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total8) return;
    const int c0 = static_cast<int>((i * 8) % C);
    float g[8], v[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + i * 8), g);      // may alias dz_out: plain load
    if (dy2) {
        float g2[8];
        unpack8(ldg_stream(dy2 + i * 8), g2);
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] += g2[q];
    }
    unpack8(ldg_stream(z + i * 8), v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float gm = act_mask(g[q], fmaf(v[q], scale[c0 + q], shift[c0 + q]), act);
        g[q] = fmaf(coef[c0 + q], gm, fmaf(coef[C + c0 + q], v[q], coef[2 * C + c0 + q]));
    }
    stg_stream(dz_out + i * 8, pack8(g));
}

// ------------------------------------------------------------------------------------ generic column sums
// out[g][c] = sum over rows [g*rpg, (g+1)*rpg) of x[row][c]; one block per (group, 64-channel slab); fixed order
__global__ void __launch_bounds__(256)
colsum_groups_kernel(const float* __restrict__ xf, const bf16* __restrict__ xb, int ld, long long rpg, int C,
                     float* __restrict__ out) {
    __shared__ double s_red[4][64];
    const int g = blockIdx.x;
    const int c = blockIdx.y * 64 + (threadIdx.x & 63);
    const int lr = threadIdx.x >> 6;
    double acc = 0.0;
    if (c < C) {
        const long long r0 = static_cast<long long>(g) * rpg;
        for (long long r = r0 + lr; r < r0 + rpg; r += 4)
            acc += xf ? static_cast<double>(xf[r * ld + c]) : static_cast<double>(__bfloat162float(xb[r * ld + c]));
    }
    s_red[lr][threadIdx.x & 63] = acc;
    __syncthreads();
    if (lr == 0 && c < C)
        out[static_cast<long long>(g) * C + c] = static_cast<float>(s_red[0][threadIdx.x] + s_red[1][threadIdx.x] +
                                                                    s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}

// ------------------------------------------------------------------------------------ image pooling branch
// single block; everything is tiny (N <= 64 images, 320 -> 256 -> 256)
__global__ void __launch_bounds__(256)
imgpool_fwd_kernel(ImgPoolFwd a) {
    const int N = a.N, Cin = a.Cin, Cm = a.Cmid, Co = a.Cout;
    // z[n][co] = sum_c pooled[n][c] * w_pool[c][co]
    for (int i = threadIdx.x; i < N * Cm; i += blockDim.x) {
        const int n = i / Cm, co = i % Cm;
        float acc = 0.f;
        for (int c = 0; c < Cin; ++c) acc = fmaf(a.pooled[n * Cin + c], a.w_pool[c * Cm + co], acc);
        a.z[i] = acc;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < Cm; co += blockDim.x) {
        float sc, sh;
        if (a.frozen) {
            sc = a.bn.gamma[co] * rsqrtf(a.bn.moving_var[co] + a.bn.eps);
            sh = a.bn.beta[co] - a.bn.moving_mean[co] * sc;
        } else {
            double s = 0.0, q = 0.0;
            for (int n = 0; n < N; ++n) { const double v = a.z[n * Cm + co]; s += v; }
            const double mean = s / N;
            for (int n = 0; n < N; ++n) { const double d = a.z[n * Cm + co] - mean; q += d * d; }
            const double var = q / N;
            const float meanf = static_cast<float>(mean), varf = static_cast<float>(var);
            const float rstd = rsqrtf(varf + a.bn.eps);
            sc = a.bn.gamma[co] * rstd;
            sh = a.bn.beta[co] - meanf * sc;
            a.bn.mean[co] = meanf;
            a.bn.rstd[co] = rstd;
            if (a.update_moving) {
                const float unb = static_cast<float>(var * (static_cast<double>(N) / fmax(N - 1.0, 1.0)));
                const float mm = a.bn.moving_mean[co], mv = a.bn.moving_var[co];
                a.bn.moving_mean[co] = __fsub_rn(mm, __fmul_rn(__fsub_rn(mm, meanf), a.bn.one_minus_decay));
                a.bn.moving_var[co] = __fsub_rn(mv, __fmul_rn(__fsub_rn(mv, unb), a.bn.one_minus_decay));
            }
        }
        a.bn.scale[co] = sc;
        a.bn.shift[co] = sh;
        for (int n = 0; n < N; ++n) a.act[n * Cm + co] = fmaxf(fmaf(a.z[n * Cm + co], sc, sh), 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * Co; i += blockDim.x) {
        const int n = i / Co, co = i % Co;
        float acc = 0.f;
        for (int c = 0; c < Cm; ++c) acc = fmaf(a.act[n * Cm + c], a.w_proj_top[c * Co + co], acc);
        a.bias_img[i] = acc;
    }
}

__global__ void __launch_bounds__(256)
imgpool_bwd_kernel(ImgPoolBwd b, float* __restrict__ dact /*[N][Cmid] scratch*/) {
    const ImgPoolFwd& a = b.f;
    const int N = a.N, Cin = a.Cin, Cm = a.Cmid, Co = a.Cout;
    // d act[n][c] = sum_co dbias[n][co] * w_proj_top[c][co]
    for (int i = threadIdx.x; i < N * Cm; i += blockDim.x) {
        const int n = i / Cm, c = i % Cm;
        float acc = 0.f;
        for (int co = 0; co < Co; ++co) acc = fmaf(b.dbias[n * Co + co], a.w_proj_top[c * Co + co], acc);
        dact[i] = a.act[i] > 0.f ? acc : 0.f;                 // through the ReLU
    }
    // d w_proj_top[c][co] = sum_n act[n][c] * dbias[n][co]
    for (int i = threadIdx.x; i < Cm * Co; i += blockDim.x) {
        const int c = i / Co, co = i % Co;
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(a.act[n * Cm + c], b.dbias[n * Co + co], acc);
        b.d_w_proj_top[i] = acc;
    }
    __syncthreads();
    // BN backward over the batch dimension (training statistics): dact -> dz (in place)
    for (int co = threadIdx.x; co < Cm; co += blockDim.x) {
        const double mean = a.bn.mean[co], rstd = a.bn.rstd[co], gamma = a.bn.gamma[co];
        double s1 = 0.0, s2 = 0.0;
        for (int n = 0; n < N; ++n) {
            const double g = dact[n * Cm + co];
            s1 += g;
            s2 += g * (a.z[n * Cm + co] - mean) * rstd;
        }
        b.d_gamma[co] = static_cast<float>(s2);
        b.d_beta[co] = static_cast<float>(s1);
        for (int n = 0; n < N; ++n) {
            const double xh = (a.z[n * Cm + co] - mean) * rstd;
            dact[n * Cm + co] = static_cast<float>(gamma * rstd * (dact[n * Cm + co] - s1 / N - xh * s2 / N));
        }
    }
    __syncthreads();
    // d w_pool[c][co] = sum_n pooled[n][c] * dz[n][co]
    for (int i = threadIdx.x; i < Cin * Cm; i += blockDim.x) {
        const int c = i / Cm, co = i % Cm;
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(a.pooled[n * Cin + c], dact[n * Cm + co], acc);
        b.d_w_pool[i] = acc;
    }
    // d pooled[n][c] / HW
    const float inv_hw = 1.f / static_cast<float>(a.HW);
    for (int i = threadIdx.x; i < N * Cin; i += blockDim.x) {
        const int n = i / Cin, c = i % Cin;
        float acc = 0.f;
        for (int co = 0; co < Cm; ++co) acc = fmaf(dact[n * Cm + co], a.w_pool[c * Cm + co], acc);
        b.dfeat_rowbias[i] = acc * inv_hw;
    }
}

__global__ void scale_rows_kernel(float* x, int n, float s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= s;
}

}  // namespace

// ============================================================================================ host
size_t bn_workspace_doubles(long long M, int C) {
    return static_cast<size_t>(red_chunks(M, C)) * 2 * C + 3 * static_cast<size_t>(C);   // partials + coef (as floats inside)
}

static size_t red_smem(int C) {
    const int tpr = std::min(C / 8, kRedThreads);
    return static_cast<size_t>(kRedThreads / tpr) * 2 * C * sizeof(double);
}

int bn_forward_stats(const bf16* z, const BnLayer& L, int update_moving, double* ws, cudaStream_t s) {
    AMS_REQUIRE(L.C % 8 == 0, "BN channels must be a multiple of 8");
    const int chunks = red_chunks(L.M, L.C);
    const long long rpc = ceil_div_ll(L.M, chunks);
    const size_t smem = red_smem(L.C);
    AMS_REQUIRE(smem <= 48 * 1024, "BN reduction shared memory");
    bn_stats_kernel<<<chunks, kRedThreads, smem, s>>>(z, L.M, L.C, rpc, ws);
    AMS_LAUNCH_CHECK();
    bn_finalize_kernel<<<ceil_div(L.C, 128), 128, 0, s>>>(ws, chunks, L, update_moving);
    AMS_LAUNCH_CHECK();
    return 0;
}

int bn_apply(const bf16* z, const float* scale, const float* shift, int act, const bf16* residual, bf16* y, long long M,
             int C, cudaStream_t s) {
    const long long total8 = M * C / 8;
    bn_apply_kernel<<<static_cast<int>(ceil_div_ll(total8, 256)), 256, 0, s>>>(z, scale, shift, act, residual, y, total8, C);
    AMS_LAUNCH_CHECK();
    return 0;
}

int bn_fold_frozen(const float* gamma, const float* beta, const float* mm, const float* mv, float eps, float* scale,
                   float* shift, int C, cudaStream_t s) {
    bn_fold_kernel<<<ceil_div(C, 128), 128, 0, s>>>(gamma, beta, mm, mv, eps, scale, shift, C);
    AMS_LAUNCH_CHECK();
    return 0;
}

int bn_backward(const bf16* dy, const bf16* dy2, const bf16* z, const BnLayer& L, int act, bf16* dz_out, float* d_gamma,
                float* d_beta, double* ws, cudaStream_t s) {
    const int chunks = red_chunks(L.M, L.C);
    const long long rpc = ceil_div_ll(L.M, chunks);
    const size_t smem = red_smem(L.C);
    float* coef = reinterpret_cast<float*>(ws + static_cast<size_t>(chunks) * 2 * L.C);
    bn_bwd_reduce_kernel<<<chunks, kRedThreads, smem, s>>>(dy, dy2, z, L.scale, L.shift, act, L.M, L.C, rpc, ws);
    AMS_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<ceil_div(L.C, 128), 128, 0, s>>>(ws, chunks, L, d_gamma, d_beta, coef);
    AMS_LAUNCH_CHECK();
    const long long total8 = L.M * L.C / 8;
    bn_bwd_apply_kernel<<<static_cast<int>(ceil_div_ll(total8, 256)), 256, 0, s>>>(dy, dy2, z, L.scale, L.shift, coef, act, total8, L.C, dz_out);
    AMS_LAUNCH_CHECK();
    return 0;
}

int colsum_groups(const float* xf, const bf16* xb, int ld, long long rows_per_group, int groups, int C, float* out,
                  cudaStream_t s) {
    dim3 grid(groups, ceil_div(C, 64));
    colsum_groups_kernel<<<grid, 256, 0, s>>>(xf, xb, ld, rows_per_group, C, out);
    AMS_LAUNCH_CHECK();
    return 0;
}

int imgpool_forward(const ImgPoolFwd& a, cudaStream_t s) {
    // pooled[n][c] = mean over HW of feat
    if (colsum_groups(nullptr, a.feat, a.Cin, a.HW, a.N, a.Cin, a.pooled, s)) return -1;
    scale_rows_kernel<<<ceil_div(a.N * a.Cin, 256), 256, 0, s>>>(a.pooled, a.N * a.Cin, 1.f / static_cast<float>(a.HW));
    AMS_LAUNCH_CHECK();
    imgpool_fwd_kernel<<<1, 256, 0, s>>>(a);
    AMS_LAUNCH_CHECK();
    return 0;
}

int imgpool_backward(const ImgPoolBwd& b, cudaStream_t s) {
    const ImgPoolFwd& a = b.f;
    if (colsum_groups(nullptr, b.dz_proj, a.Cout, a.HW, a.N, a.Cout, b.dbias, s)) return -1;
    // dact scratch lives right after dbias (caller allocates N*(Cout+Cmid) floats)
    imgpool_bwd_kernel<<<1, 256, 0, s>>>(b, b.dbias + static_cast<size_t>(a.N) * a.Cout);
    AMS_LAUNCH_CHECK();
    return 0;
}

}  // namespace ams
