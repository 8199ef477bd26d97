// BatchNorm kernels (training-mode statistics, apply, backward) and the small fp32 kernels of the ASPP
// image-pooling branch.  Everything here is a streaming reduction or elementwise pass over bf16 NHWC
// activations: 128-bit loads, per-thread fp32 runs flushed into fp64 accumulators, fixed-order (deterministic)
// cross-block reduction -- no floating-point atomics anywhere (SURVEY 7 'bit-exact top-k' hard part).
// Replaces: the 54 `FusedBatchNormV3(is_training=True)` nodes + 108 `AssignSub` moving-average updates of
// checkpoints/*/model.meta and their gradients; inference-mode patch BN of utils/graph_utils.py:362-369.
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace ams {
namespace {

constexpr int kRedThreads = 256;
constexpr int kFlush = 16;          // fp32 run length before flushing into fp64

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
// Measured (tools/ab_step.py, gpurun_out/ab_r1*.jsonl): 2 chunks per SM with 4 rows of each tensor in flight per thread
// beats 6 chunks x 2 rows by 0.2 ms per step (5.90 -> 5.71 ms) -- every chunk costs a [2][C] fp64 partial row that is
// written here and read again by the finalize, so memory-level parallelism has to come from the loads, not from CTAs.
static const int kRedChunksPerSM = env_int("AMS_RED_CHUNKS_PER_SM", 2);
static const int kBwdRedUnroll = env_int("AMS_BWD_RED_UNROLL", 4);

int red_chunks(long long M, int C) {
    // enough blocks to fill the machine while every thread still walks >= 8 rows (see kRedChunksPerSM)
    const int c8 = C / 8;
    const int rows_per_block = std::max(1, kRedThreads / std::min(c8, kRedThreads));
    const long long want = std::min<long long>(static_cast<long long>(kRedChunksPerSM) * kNumSMs, M / (static_cast<long long>(rows_per_block) * 8));
    return static_cast<int>(std::max<long long>(1, want));
}

// partial[chunk][0][c] = sum z, partial[chunk][1][c] = sum z^2 over the rows of the chunk
__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const act_t* __restrict__ z, long long M, int C, long long rows_per_chunk, double* __restrict__ partial) {
    pdl_entry();
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
            const long long RB = rows_in_block;
            while (r < r_end) {
                float fs[8], fq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { fs[q] = 0.f; fq[q] = 0.f; }
                // 4 x 4 rows per fp32 run; 4 independent 128-bit loads in flight per thread
                for (int it = 0; it < kFlush / 4 && r < r_end; ++it) {
                    uint4 raw[4];
                    int nv = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long rr = r + u * RB;
                        if (rr < r_end) { raw[u] = ldg_stream(z + rr * C + c8 * 8); nv = u + 1; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u < nv) {
                            float v[8];
                            unpack8h(raw[u], v);
#pragma unroll
                            for (int q = 0; q < 8; ++q) { fs[q] += v[q]; fq[q] = fmaf(v[q], v[q], fq[q]); }
                        }
                    }
                    r += 4 * RB;
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

// block = 32 channels x kFinRows row-lanes: coalesced reads of the partial rows, fixed lane assignment and a
// fixed-order shared-memory tree => deterministic
constexpr int kFinRows = 32;
__device__ __forceinline__ bool finalize_sums(const double* __restrict__ partial, int chunks, int C, int c, double* s_out,
                                              double* q_out) {
    __shared__ double s_s[kFinRows][33], s_q[kFinRows][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    double s = 0.0, q = 0.0;
    if (c < C) {
        // 16 independent row loads in flight per thread and pass (the rows are a latency chain otherwise); the ragged
        // tail is one more predicated pass, not a serial loop
        for (int k = rl; k < chunks; k += 8 * kFinRows) {
            double a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int kk = k + u * kFinRows;
                const bool ok = kk < chunks;
                a[u] = ok ? partial[static_cast<long long>(kk) * 2 * C + c] : 0.0;
                b[u] = ok ? partial[static_cast<long long>(kk) * 2 * C + C + c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s += a[u]; q += b[u]; }
        }
    }
    s_s[rl][cl] = s; s_q[rl][cl] = q;
    __syncthreads();
    if (rl != 0 || c >= C) return false;
    for (int k = 1; k < kFinRows; ++k) { s += s_s[k][cl]; q += s_q[k][cl]; }
    *s_out = s; *q_out = q;
    return true;
}

// ------------------------------------------------------------------------------------ cross-GPU sum of a (s, q) pair
// Called by ONE thread per channel.  word_off = the layer's offset + 4 * channel.  Push: 4 words (payload 32 bits |
// epoch << 32) into slot [parity][my rank] of every rank's receive buffer (8-byte stores are single NVLink writes, so a
// word is either old or complete).  Gather: poll the local slots of every source until their flag equals the epoch, add
// in rank order.  A peer that never arrives trips the timeout: the error flag is raised (the host checks it) and the
// local values are kept -- the kernel never hangs.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void syncbn_push(SyncBn* sb, int peer, long long slot, unsigned int ep, double s, double q) {
    const unsigned long long tag = static_cast<unsigned long long>(ep) << 32;
    unsigned long long* dst = sb->peer[peer] + slot;
    const unsigned long long w0 = tag | static_cast<unsigned int>(__double2loint(s)), w1 = tag | static_cast<unsigned int>(__double2hiint(s));
    const unsigned long long w2 = tag | static_cast<unsigned int>(__double2loint(q)), w3 = tag | static_cast<unsigned int>(__double2hiint(q));
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst), "l"(w0) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst + 1), "l"(w1) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst + 2), "l"(w2) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst + 3), "l"(w3) : "memory");
}
// poll the 4 words of one source in the LOCAL receive buffer (all four loads in flight per round); false = timed out
__device__ __forceinline__ bool syncbn_poll(SyncBn* sb, long long slot, unsigned int ep, unsigned long long t0, double* s, double* q) {
    const unsigned long long* from = sb->peer[sb->rank] + slot;
    unsigned long long v0, v1, v2, v3;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v0) : "l"(from) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v1) : "l"(from + 1) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v2) : "l"(from + 2) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v3) : "l"(from + 3) : "memory");
        if (static_cast<unsigned int>(v0 >> 32) == ep && static_cast<unsigned int>(v1 >> 32) == ep &&
            static_cast<unsigned int>(v2 >> 32) == ep && static_cast<unsigned int>(v3 >> 32) == ep) break;
        if ((++spins & 63u) == 0u && global_timer_ns() - t0 > sb->timeout_ns) { atomicExch(&sb->error, 1u); return false; }
    }
    *s = __hiloint2double(static_cast<int>(static_cast<unsigned int>(v1)), static_cast<int>(static_cast<unsigned int>(v0)));
    *q = __hiloint2double(static_cast<int>(static_cast<unsigned int>(v3)), static_cast<int>(static_cast<unsigned int>(v2)));
    return true;
}
// slot of (source rank, word offset) in the receive buffer of the step's epoch parity
__device__ __forceinline__ long long syncbn_slot(const SyncBn* sb, unsigned int ep, int src, long long word_off) {
    return (static_cast<long long>(ep & 1u) * sb->world + src) * sb->words_per_src + word_off;
}
// one thread per channel does everything (image-pooling BN: two exchanges per step, nothing to parallelise);
// a[src], b[src] = the pair of every source rank (own values on a timeout: error flag raised)
__device__ __forceinline__ void syncbn_gather(SyncBn* sb, long long word_off, double s, double q, double* a, double* b) {
    const int world = sb->world, rank = sb->rank;
    const unsigned int ep = sb->epoch;
    for (int p = 0; p < world; ++p) syncbn_push(sb, (rank + p) % world, syncbn_slot(sb, ep, rank, word_off), ep, s, q);
    for (int src = 0; src < world; ++src) { a[src] = s; b[src] = q; }
    if (*reinterpret_cast<volatile unsigned int*>(&sb->error)) return;
    const unsigned long long t0 = global_timer_ns();
    for (int src = 0; src < world; ++src)
        if (!syncbn_poll(sb, syncbn_slot(sb, ep, src, word_off), ep, t0, &a[src], &b[src])) return;
}
// Block form for the finalize kernels (32 channels x 32 row lanes, every thread of the block calls it): the lead lanes
// (row lane 0) hold the channel's pair; row lane r < world pushes it to peer r and polls source r, so the exchange costs
// one NVLink store + one polling round whatever the world size; the lead lanes then add the sources in rank order.
__device__ __forceinline__ void syncbn_exchange_block(SyncBn* sb, long long layer_off, int c, int C, bool lead, double* s, double* q) {
    __shared__ double x_s[kSyncBnMaxWorld + 1][32], x_q[kSyncBnMaxWorld + 1][32];
    __shared__ int x_fail;
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int world = sb->world, rank = sb->rank;
    const unsigned int ep = sb->epoch;
    if (threadIdx.x == 0) x_fail = *reinterpret_cast<volatile unsigned int*>(&sb->error) ? 1 : 0;
    if (lead) { x_s[kSyncBnMaxWorld][cl] = *s; x_q[kSyncBnMaxWorld][cl] = *q; }
    __syncthreads();
    if (rl < world && c < C) {
        const long long off = layer_off + 4LL * c;
        syncbn_push(sb, rl, syncbn_slot(sb, ep, rank, off), ep, x_s[kSyncBnMaxWorld][cl], x_q[kSyncBnMaxWorld][cl]);
        if (!x_fail) {
            double a = 0.0, b = 0.0;
            if (!syncbn_poll(sb, syncbn_slot(sb, ep, rl, off), ep, global_timer_ns(), &a, &b)) atomicExch(&x_fail, 1);
            x_s[rl][cl] = a; x_q[rl][cl] = b;
        }
    }
    __syncthreads();
    if (lead && !x_fail) {
        double S = 0.0, Q = 0.0;
        for (int src = 0; src < world; ++src) { S += x_s[src][cl]; Q += x_q[src][cl]; }
        *s = S; *q = Q;
    }
}
__global__ void syncbn_begin_step_kernel(SyncBn* sb) {
    pdl_entry();
    if (threadIdx.x == 0) {
        unsigned int e = sb->epoch + 1u;
        if (e == 0u) e = 1u;                  // 0 is the "never written" flag of a fresh buffer
        sb->epoch = e;
    }
}

__global__ void __launch_bounds__(32 * kFinRows)
bn_finalize_kernel(const double* __restrict__ partial, int chunks, BnLayer L, int update_moving) {
    pdl_entry();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double s, q;
    const bool lead = finalize_sums(partial, chunks, L.C, c, &s, &q);
    double n = static_cast<double>(L.M);
    if (L.sync) { syncbn_exchange_block(L.sync, L.xoff_fwd, c, L.C, lead, &s, &q); n *= L.sync->world; }
    if (!lead) return;
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

// Elementwise passes: a thread owns one 8-channel group (its per-channel parameters live in registers) and walks kEwRows
// rows of the block's row range with all loads issued before the arithmetic; block = min(C/8, 256) channel groups x rows.
static const int kEwRowsKnob = env_int("AMS_EW_ROWS", 4);      // rows per thread of the elementwise passes: 2, 4 or 8
template <int kEwRows>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const act_t* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                const act_t* __restrict__ residual, act_t* __restrict__ y, int M, int C) {
    pdl_entry();
    const int c8n = C >> 3;
    const int tpr = min(c8n, 256), rib = 256 / tpr;
    const int lr = threadIdx.x / tpr, lc = threadIdx.x - lr * tpr;
    if (lr >= rib) return;
    for (int c8 = lc; c8 < c8n; c8 += tpr) {
        const int c0 = c8 * 8;
        float sc[8], sh[8];
        {
            const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
            sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
            sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
        }
        const int r0 = blockIdx.x * (rib * kEwRows) + lr;
        uint4 vz[kEwRows], vr[kEwRows];
#pragma unroll
        for (int u = 0; u < kEwRows; ++u) {
            const int r = r0 + u * rib;
            if (r < M) {
                const long long off = static_cast<long long>(r) * C + c0;
                vz[u] = ldg_stream(z + off);
                if (residual) vr[u] = ldg_stream(residual + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwRows; ++u) {
            const int r = r0 + u * rib;
            if (r < M) {
                float v[8];
                unpack8h(vz[u], v);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = act_apply(fmaf(v[q], sc[q], sh[q]), act);
                if (residual) {
                    float f[8];
                    unpack8h(vr[u], f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += f[q];
                }
                stg_stream(y + static_cast<long long>(r) * C + c0, pack8h(v));
            }
        }
    }
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mm, const float* mv, float eps,
                               float* scale, float* shift, int C) {
    pdl_entry();
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

// U = rows in flight per thread and tensor (2 or 4 independent 128-bit loads each of dy and z); HAS2 = a second gradient
template <int U, bool HAS2>
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2, const act_t* __restrict__ z,
                     const float* __restrict__ scale, const float* __restrict__ shift, int act, long long M, int C,
                     long long rows_per_chunk, double* __restrict__ partial) {
    pdl_entry();
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
            const long long RB = rows_in_block;
            while (r < r_end) {
                float fs[8], fq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { fs[q] = 0.f; fq[q] = 0.f; }
                for (int it = 0; it < kFlush / U && r < r_end; ++it) {
                    uint4 rg[U], rz[U], rg2[HAS2 ? U : 1];
                    int nv = 0;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long rr = r + u * RB;
                        if (rr < r_end) {
                            rg[u] = ldg_stream(dy + rr * C + c8 * 8);
                            rz[u] = ldg_stream(z + rr * C + c8 * 8);
                            if (HAS2) rg2[HAS2 ? u : 0] = ldg_stream(dy2 + rr * C + c8 * 8);
                            nv = u + 1;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (u < nv) {
                            float g[8], v[8];
                            unpack8(rg[u], g);
                            unpack8h(rz[u], v);
                            if (HAS2) {
                                float g2[8];
                                unpack8(rg2[HAS2 ? u : 0], g2);
#pragma unroll
                                for (int q = 0; q < 8; ++q) g[q] += g2[q];
                            }
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float gm = act_mask(g[q], fmaf(v[q], sc[q], sh[q]), act);
                                fs[q] += gm;
                                fq[q] = fmaf(gm, v[q], fq[q]);
                            }
                        }
                    }
                    r += U * RB;
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
__global__ void __launch_bounds__(32 * kFinRows)
bn_bwd_finalize_kernel(const double* __restrict__ partial, int chunks, BnLayer L, float* d_gamma, float* d_beta, float* coef) {
    pdl_entry();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double s1, sz;
    const bool lead = finalize_sums(partial, chunks, L.C, c, &s1, &sz);
    const double s1_local = s1, sz_local = sz;
    double n = static_cast<double>(L.M);
    if (L.sync) { syncbn_exchange_block(L.sync, L.xoff_bwd, c, L.C, lead, &s1, &sz); n *= L.sync->world; }
    if (!lead) return;
    const double mean = L.mean[c], rstd = L.rstd[c], gamma = L.gamma[c];
    d_gamma[c] = static_cast<float>(rstd * (sz_local - mean * s1_local));    // LOCAL sums: the gradient allreduce adds the ranks
    d_beta[c] = static_cast<float>(s1_local);
    const double s2 = rstd * (sz - mean * s1);           // sum g * xhat over the (global) batch
    const double A = gamma * rstd;
    const double B = -gamma * rstd * rstd * s2 / n;
    const double Cc = -gamma * rstd * (s1 / n - mean * rstd * s2 / n);
    coef[c] = static_cast<float>(A);
    coef[L.C + c] = static_cast<float>(B);
    coef[2 * L.C + c] = static_cast<float>(Cc);
}

__device__ __forceinline__ void load8f(const float* p, float* o) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

template <int kEwRows>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* dy, const bf16* __restrict__ dy2, const act_t* __restrict__ z,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
                    int act, int M, int C, bf16* dz_out) {
    pdl_entry();
    const int c8n = C >> 3;
    const int tpr = min(c8n, 256), rib = 256 / tpr;
    const int lr = threadIdx.x / tpr, lc = threadIdx.x - lr * tpr;
    if (lr >= rib) return;
    for (int c8 = lc; c8 < c8n; c8 += tpr) {
        const int c0 = c8 * 8;
        float sc[8], sh[8], ca[8], cb[8], cc[8];
        load8f(scale + c0, sc); load8f(shift + c0, sh);
        load8f(coef + c0, ca); load8f(coef + C + c0, cb); load8f(coef + 2 * C + c0, cc);
        const int r0 = blockIdx.x * (rib * kEwRows) + lr;
        uint4 vg[kEwRows], vz[kEwRows], vg2[kEwRows];
#pragma unroll
        for (int u = 0; u < kEwRows; ++u) {
            const int r = r0 + u * rib;
            if (r < M) {
                const long long off = static_cast<long long>(r) * C + c0;
                vg[u] = *reinterpret_cast<const uint4*>(dy + off);           // may alias dz_out: plain load
                vz[u] = ldg_stream(z + off);
                if (dy2) vg2[u] = ldg_stream(dy2 + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwRows; ++u) {
            const int r = r0 + u * rib;
            if (r < M) {
                float g[8], v[8];
                unpack8(vg[u], g);
                unpack8h(vz[u], v);
                if (dy2) {
                    float g2[8];
                    unpack8(vg2[u], g2);
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] += g2[q];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float gm = act_mask(g[q], fmaf(v[q], sc[q], sh[q]), act);
                    g[q] = fmaf(ca[q], gm, fmaf(cb[q], v[q], cc[q]));
                }
                stg_stream(dz_out + static_cast<long long>(r) * C + c0, pack8(g));
            }
        }
    }
}

// ------------------------------------------------------------------------------------ generic column sums
// partial[g][split][c] = sum over the split's rows of group g of x[row][c]; thread = (row lane, channel), 64 channels
// per block; a second tiny kernel adds the splits in fixed order (deterministic).
constexpr int kColsumSplits = 32;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ xf, const bf16* __restrict__ xb, const act_t* __restrict__ xh, int ld, long long rpg, int C,
                      double* __restrict__ partial) {
    pdl_entry();
    __shared__ double s_red[4][64];
    const int g = blockIdx.x, split = blockIdx.z;
    const int c = blockIdx.y * 64 + (threadIdx.x & 63);
    const int lr = threadIdx.x >> 6;
    const long long rows_per_split = (rpg + kColsumSplits - 1) / kColsumSplits;
    const long long r0 = static_cast<long long>(g) * rpg + split * rows_per_split;
    const long long r1 = min(r0 + rows_per_split, static_cast<long long>(g + 1) * rpg);
    double acc = 0.0;
    if (c < C) {
        float run = 0.f;
        int cnt = 0;
        for (long long r = r0 + lr; r < r1; r += 4) {
            run += xf ? xf[r * ld + c] : (xb ? __bfloat162float(xb[r * ld + c]) : __half2float(xh[r * ld + c]));
            if (++cnt == 32) { acc += run; run = 0.f; cnt = 0; }
        }
        acc += run;
    }
    s_red[lr][threadIdx.x & 63] = acc;
    __syncthreads();
    if (lr == 0 && c < C)
        partial[(static_cast<long long>(g) * kColsumSplits + split) * C + c] =
            s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x];
}
__global__ void colsum_final_kernel(const double* __restrict__ partial, int groups, int C, float scale, float* __restrict__ out) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * C) return;
    const int g = i / C, c = i % C;
    double acc = 0.0;
    for (int s = 0; s < kColsumSplits; ++s) acc += partial[(static_cast<long long>(g) * kColsumSplits + s) * C + c];
    out[i] = static_cast<float>(acc) * scale;
}

// ------------------------------------------------------------------------------------ image pooling branch
// Everything here is tiny (N <= 64 images, 320 -> 256 -> 256) and fp32; several small multi-block kernels.
// out[n][co] = sum_c in[n][c] * w[c][co].  Block = 32 outputs (coalesced over co) x 32 slices of the input channels; the
// slice partials are summed in a fixed order (deterministic).  A single thread per output walked Cin dependent FMAs:
// 160 us for the teacher's 2048 -> 256 image-pooling conv, on the critical path of its ASPP.
__global__ void __launch_bounds__(1024)
small_fc_kernel(const float* __restrict__ in, const float* __restrict__ w, int N, int Cin, int Cout, float* __restrict__ out) {
    pdl_entry();
    __shared__ float part[32][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int tiles = (Cout + 31) / 32;
    const int n = blockIdx.x / tiles, co = (blockIdx.x % tiles) * 32 + lane;
    float acc = 0.f;
    if (co < Cout) {
        const float* x = in + static_cast<long long>(n) * Cin;
        for (int c = slice; c < Cin; c += 32) acc = fmaf(x[c], w[static_cast<long long>(c) * Cout + co], acc);
    }
    part[slice][lane] = acc;
    __syncthreads();
    if (slice == 0 && co < Cout) {
        float t = part[0][lane];
#pragma unroll
        for (int k = 1; k < 32; ++k) t += part[k][lane];
        out[static_cast<long long>(n) * Cout + co] = t;
    }
}
// out[n][c] = scale * sum_co g[n][co] * w[c][co]   (warp per output, lanes over co), optional ReLU gate
__global__ void __launch_bounds__(256)
small_fc_t_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ gate, int N, int C,
                  int Cout, float scale, float* __restrict__ out) {
    pdl_entry();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N * C) return;
    const int n = warp / C, c = warp % C;
    float acc = 0.f;
    for (int co = lane; co < Cout; co += 32) acc = fmaf(g[n * Cout + co], w[c * Cout + co], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[warp] = (gate && !(gate[warp] > 0.f)) ? 0.f : acc * scale;
}
// out[c][co] = sum_n a[n][c] * b[n][co]
__global__ void __launch_bounds__(256)
small_outer_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int C, int Cout, float* __restrict__ out) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * Cout) return;
    const int c = i / Cout, co = i % Cout;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(a[n * C + c], b[n * Cout + co], acc);
    out[i] = acc;
}
// BN over the batch dimension + ReLU (training statistics or frozen)
__global__ void __launch_bounds__(256)
imgpool_bn_kernel(ImgPoolFwd a) {
    pdl_entry();
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = a.N, Cm = a.Cmid;
    if (co >= Cm) return;
    float sc, sh;
    if (a.frozen) {
        sc = a.bn.gamma[co] * rsqrtf(a.bn.moving_var[co] + a.bn.eps);
        sh = a.bn.beta[co] - a.bn.moving_mean[co] * sc;
    } else {
        double s = 0.0, q = 0.0, cnt = N;
        for (int n = 0; n < N; ++n) s += a.z[n * Cm + co];
        double mean, var;
        if (a.bn.sync) {
            // global batch: every rank's (sum, centred sum of squares) -> pooled mean / variance (Chan et al.)
            const double mean_l = s / N;
            for (int n = 0; n < N; ++n) { const double d = a.z[n * Cm + co] - mean_l; q += d * d; }
            double ss[kSyncBnMaxWorld], qq[kSyncBnMaxWorld];
            syncbn_gather(a.bn.sync, a.bn.xoff_fwd + 4LL * co, s, q, ss, qq);
            const int world = a.bn.sync->world;
            cnt *= world;
            double S = 0.0, M2 = 0.0;
            for (int r = 0; r < world; ++r) S += ss[r];
            mean = S / cnt;
            for (int r = 0; r < world; ++r) { const double d = ss[r] / N - mean; M2 += qq[r] + N * d * d; }
            var = M2 / cnt;
        } else {
            mean = s / N;
            for (int n = 0; n < N; ++n) { const double d = a.z[n * Cm + co] - mean; q += d * d; }
            var = q / N;
        }
        const float meanf = static_cast<float>(mean), varf = static_cast<float>(var);
        const float rstd = rsqrtf(varf + a.bn.eps);
        sc = a.bn.gamma[co] * rstd;
        sh = a.bn.beta[co] - meanf * sc;
        a.bn.mean[co] = meanf;
        a.bn.rstd[co] = rstd;
        if (a.update_moving) {
            const float unb = static_cast<float>(var * (cnt / fmax(cnt - 1.0, 1.0)));
            const float mm = a.bn.moving_mean[co], mv = a.bn.moving_var[co];
            a.bn.moving_mean[co] = __fsub_rn(mm, __fmul_rn(__fsub_rn(mm, meanf), a.bn.one_minus_decay));
            a.bn.moving_var[co] = __fsub_rn(mv, __fmul_rn(__fsub_rn(mv, unb), a.bn.one_minus_decay));
        }
    }
    a.bn.scale[co] = sc;
    a.bn.shift[co] = sh;
    for (int n = 0; n < N; ++n) a.act[n * Cm + co] = fmaxf(fmaf(a.z[n * Cm + co], sc, sh), 0.f);
}
// BN backward over the batch dimension: dact -> dz in place, d_gamma, d_beta
__global__ void __launch_bounds__(256)
imgpool_bn_bwd_kernel(ImgPoolFwd a, float* __restrict__ dact, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
    pdl_entry();
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = a.N, Cm = a.Cmid;
    if (co >= Cm) return;
    const double mean = a.bn.mean[co], rstd = a.bn.rstd[co], gamma = a.bn.gamma[co];
    double s1 = 0.0, s2 = 0.0;
    for (int n = 0; n < N; ++n) {
        const double g = dact[n * Cm + co];
        s1 += g;
        s2 += g * (a.z[n * Cm + co] - mean) * rstd;
    }
    d_gamma[co] = static_cast<float>(s2);                // LOCAL sums: the gradient allreduce adds the ranks
    d_beta[co] = static_cast<float>(s1);
    double cnt = N;
    if (a.bn.sync) {
        double ss[kSyncBnMaxWorld], qq[kSyncBnMaxWorld];
        syncbn_gather(a.bn.sync, a.bn.xoff_bwd + 4LL * co, s1, s2, ss, qq);
        const int world = a.bn.sync->world;
        cnt *= world;
        s1 = 0.0; s2 = 0.0;
        for (int r = 0; r < world; ++r) { s1 += ss[r]; s2 += qq[r]; }
    }
    for (int n = 0; n < N; ++n) {
        const double xh = (a.z[n * Cm + co] - mean) * rstd;
        dact[n * Cm + co] = static_cast<float>(gamma * rstd * (dact[n * Cm + co] - s1 / cnt - xh * s2 / cnt));
    }
}

}  // namespace

// ============================================================================================ host
int syncbn_begin_step(SyncBn* sb, cudaStream_t s) {
    AMS_LAUNCH((syncbn_begin_step_kernel), 1, 32, 0, s, sb);
    return 0;
}

size_t bn_workspace_doubles(long long M, int C) {
    // partials (reduction chunks, or one slab per CTA of a producer kernel with fused statistics) + coef
    const size_t chunks = std::max<size_t>(red_chunks(M, C), 2 * kNumSMs);
    return chunks * 2 * C + 3 * static_cast<size_t>(C);
}

static size_t red_smem(int C) {
    const int tpr = std::min(C / 8, kRedThreads);
    return static_cast<size_t>(kRedThreads / tpr) * 2 * C * sizeof(double);
}

static int launch_bwd_apply(const bf16* dy, const bf16* dy2, const act_t* z, const BnLayer& L, const float* coef, int act, int rib,
                            bf16* dz_out, cudaStream_t s) {
    const int rows = kEwRowsKnob == 2 ? 2 : (kEwRowsKnob == 8 ? 8 : 4);
    const int grid = static_cast<int>(ceil_div_ll(L.M, rib * rows)), M = static_cast<int>(L.M);
    if (rows == 2) AMS_LAUNCH((bn_bwd_apply_kernel<2>), grid, 256, 0, s, dy, dy2, z, L.scale, L.shift, coef, act, M, L.C, dz_out);
    else if (rows == 8) AMS_LAUNCH((bn_bwd_apply_kernel<8>), grid, 256, 0, s, dy, dy2, z, L.scale, L.shift, coef, act, M, L.C, dz_out);
    else AMS_LAUNCH((bn_bwd_apply_kernel<4>), grid, 256, 0, s, dy, dy2, z, L.scale, L.shift, coef, act, M, L.C, dz_out);
    return 0;
}

static int launch_bwd_reduce(const bf16* dy, const bf16* dy2, const act_t* z, const BnLayer& L, int act, int chunks, long long rpc,
                             size_t smem, double* ws, cudaStream_t s) {
    if (dy2) AMS_LAUNCH((bn_bwd_reduce_kernel<2, true>), chunks, kRedThreads, smem, s, dy, dy2, z, L.scale, L.shift, act, L.M, L.C, rpc, ws);
    else if (kBwdRedUnroll == 8) AMS_LAUNCH((bn_bwd_reduce_kernel<8, false>), chunks, kRedThreads, smem, s, dy, dy2, z, L.scale, L.shift, act, L.M, L.C, rpc, ws);
    else if (kBwdRedUnroll == 4) AMS_LAUNCH((bn_bwd_reduce_kernel<4, false>), chunks, kRedThreads, smem, s, dy, dy2, z, L.scale, L.shift, act, L.M, L.C, rpc, ws);
    else AMS_LAUNCH((bn_bwd_reduce_kernel<2, false>), chunks, kRedThreads, smem, s, dy, dy2, z, L.scale, L.shift, act, L.M, L.C, rpc, ws);
    return 0;
}

int bn_forward_stats(const act_t* z, const BnLayer& L, int update_moving, double* ws, cudaStream_t s) {
    AMS_REQUIRE(L.C % 8 == 0, "BN channels must be a multiple of 8");
    const int chunks = red_chunks(L.M, L.C);
    const long long rpc = ceil_div_ll(L.M, chunks);
    const size_t smem = red_smem(L.C);
    AMS_REQUIRE(smem <= 48 * 1024, "BN reduction shared memory");
    AMS_LAUNCH((bn_stats_kernel), chunks, kRedThreads, smem, s, z, L.M, L.C, rpc, ws);
    AMS_LAUNCH((bn_finalize_kernel), ceil_div(L.C, 32), 32 * kFinRows, 0, s, ws, chunks, L, update_moving);
    return 0;
}

int bn_finalize_partials(const double* partial, int chunks, const BnLayer& L, int update_moving, cudaStream_t s) {
    AMS_LAUNCH((bn_finalize_kernel), ceil_div(L.C, 32), 32 * kFinRows, 0, s, partial, chunks, L, update_moving);
    return 0;
}

int bn_apply(const act_t* z, const float* scale, const float* shift, int act, const act_t* residual, act_t* y, long long M,
             int C, cudaStream_t s) {
    const int rib = 256 / std::min(C / 8, 256);
    const int rows = kEwRowsKnob == 2 ? 2 : (kEwRowsKnob == 8 ? 8 : 4);
    const int grid = static_cast<int>(ceil_div_ll(M, rib * rows));
    if (rows == 2) AMS_LAUNCH((bn_apply_kernel<2>), grid, 256, 0, s, z, scale, shift, act, residual, y, static_cast<int>(M), C);
    else if (rows == 8) AMS_LAUNCH((bn_apply_kernel<8>), grid, 256, 0, s, z, scale, shift, act, residual, y, static_cast<int>(M), C);
    else AMS_LAUNCH((bn_apply_kernel<4>), grid, 256, 0, s, z, scale, shift, act, residual, y, static_cast<int>(M), C);
    return 0;
}

int bn_fold_frozen(const float* gamma, const float* beta, const float* mm, const float* mv, float eps, float* scale,
                   float* shift, int C, cudaStream_t s) {
    AMS_LAUNCH((bn_fold_kernel), ceil_div(C, 128), 128, 0, s, gamma, beta, mm, mv, eps, scale, shift, C);
    return 0;
}

int bn_backward(const bf16* dy, const bf16* dy2, const act_t* z, const BnLayer& L, int act, bf16* dz_out, float* d_gamma,
                float* d_beta, double* ws, cudaStream_t s) {
    const int chunks = red_chunks(L.M, L.C);
    const long long rpc = ceil_div_ll(L.M, chunks);
    const size_t smem = red_smem(L.C);
    float* coef = reinterpret_cast<float*>(ws + static_cast<size_t>(chunks) * 2 * L.C);
    if (launch_bwd_reduce(dy, dy2, z, L, act, chunks, rpc, smem, ws, s)) return -1;
    AMS_LAUNCH((bn_bwd_finalize_kernel), ceil_div(L.C, 32), 32 * kFinRows, 0, s, ws, chunks, L, d_gamma, d_beta, coef);
    const int rib = 256 / std::min(L.C / 8, 256);
    if (launch_bwd_apply(dy, dy2, z, L, coef, act, rib, dz_out, s)) return -1;
    return 0;
}

int bn_backward_reduce(const bf16* dy, const act_t* z, const BnLayer& L, int act, float* coef, float* d_gamma, float* d_beta,
                       double* ws, cudaStream_t s) {
    const int chunks = red_chunks(L.M, L.C);
    const long long rpc = ceil_div_ll(L.M, chunks);
    if (launch_bwd_reduce(dy, nullptr, z, L, act, chunks, rpc, red_smem(L.C), ws, s)) return -1;
    AMS_LAUNCH((bn_bwd_finalize_kernel), ceil_div(L.C, 32), 32 * kFinRows, 0, s, ws, chunks, L, d_gamma, d_beta, coef);
    return 0;
}

int bn_backward_finalize_partials(const double* partial, int rows, const BnLayer& L, float* coef, float* d_gamma,
                                  float* d_beta, cudaStream_t s) {
    AMS_LAUNCH((bn_bwd_finalize_kernel), ceil_div(L.C, 32), 32 * kFinRows, 0, s, partial, rows, L, d_gamma, d_beta, coef);
    return 0;
}

int bn_backward_apply(const bf16* dy_masked, const act_t* z, const BnLayer& L, const float* coef, bf16* dz_out, cudaStream_t s) {
    const int rib = 256 / std::min(L.C / 8, 256);
    if (launch_bwd_apply(dy_masked, nullptr, z, L, coef, 0, rib, dz_out, s)) return -1;
    return 0;
}

int colsum_groups(const float* xf, const bf16* xb, const act_t* xh, int ld, long long rows_per_group, int groups, int C, float scale,
                  float* out, double* workspace, cudaStream_t s) {
    dim3 grid(groups, ceil_div(C, 64), kColsumSplits);
    AMS_LAUNCH((colsum_partial_kernel), grid, 256, 0, s, xf, xb, xh, ld, rows_per_group, C, workspace);
    AMS_LAUNCH((colsum_final_kernel), ceil_div(groups * C, 128), 128, 0, s, workspace, groups, C, scale, out);
    return 0;
}
size_t colsum_workspace_doubles(int groups, int C) { return static_cast<size_t>(groups) * kColsumSplits * C; }

int imgpool_forward(const ImgPoolFwd& a, cudaStream_t s) {
    // pooled[n][c] = mean over HW of feat
    if (colsum_groups(nullptr, nullptr, a.feat, a.Cin, a.HW, a.N, a.Cin, 1.f / static_cast<float>(a.HW), a.pooled, a.ws, s)) return -1;
    AMS_LAUNCH((small_fc_kernel), a.N * ceil_div(a.Cmid, 32), 1024, 0, s, a.pooled, a.w_pool, a.N, a.Cin, a.Cmid, a.z);
    AMS_LAUNCH((imgpool_bn_kernel), ceil_div(a.Cmid, 256), 256, 0, s, a);
    AMS_LAUNCH((small_fc_kernel), a.N * ceil_div(a.Cout, 32), 1024, 0, s, a.act, a.w_proj_top, a.N, a.Cmid, a.Cout, a.bias_img);
    return 0;
}

int imgpool_backward(const ImgPoolBwd& b, cudaStream_t s) {
    const ImgPoolFwd& a = b.f;
    float* dact = b.dbias + static_cast<size_t>(a.N) * a.Cout;       // caller allocates N*(Cout+Cmid) floats
    if (colsum_groups(nullptr, b.dz_proj, nullptr, a.Cout, a.HW, a.N, a.Cout, 1.f, b.dbias, a.ws, s)) return -1;
    // d act = relu'(act) * dbias * w_proj_top^T ;  d w_proj_top = act^T dbias
    AMS_LAUNCH((small_fc_t_kernel), ceil_div(a.N * a.Cmid * 32, 256), 256, 0, s, b.dbias, a.w_proj_top, a.act, a.N, a.Cmid, a.Cout, 1.f, dact);
    AMS_LAUNCH((small_outer_kernel), ceil_div(a.Cmid * a.Cout, 256), 256, 0, s, a.act, b.dbias, a.N, a.Cmid, a.Cout, b.d_w_proj_top);
    AMS_LAUNCH((imgpool_bn_bwd_kernel), ceil_div(a.Cmid, 256), 256, 0, s, a, dact, b.d_gamma, b.d_beta);
    AMS_LAUNCH((small_outer_kernel), ceil_div(a.Cin * a.Cmid, 256), 256, 0, s, a.pooled, dact, a.N, a.Cin, a.Cmid, b.d_w_pool);
    AMS_LAUNCH((small_fc_t_kernel), ceil_div(a.N * a.Cin * 32, 256), 256, 0, s, dact, a.w_pool, nullptr, a.N, a.Cin, a.Cmid, 1.f / static_cast<float>(a.HW), b.dfeat_rowbias);
    return 0;
}

}  // namespace ams
