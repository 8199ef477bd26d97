// Optimizer / coordinate-selection / model-delta kernels over the flat fp32 parameter arena (2.1 M floats).
// All integer work is exact; the selection is an exact function of its input array (radix select of the two
// order statistics NumPy's percentile interpolates between), so the mask is bit-identical to the oracle's.
// Replaces: 164x ApplyAdam + backup/`tf.where(mask,new,backup)` assigns (utils/graph_utils.py:459-496), the host
// NumPy selection at SemanticNetwork.py:263-288, and the packbits/fp16 delta writer at run.py:316-328.
#include "kernels.cuh"

namespace ams {
namespace {

__global__ void __launch_bounds__(256)
adam_masked_kernel(float* __restrict__ p, const float* __restrict__ g, float grad_scale,
                   const double* __restrict__ scale_terms, float* __restrict__ m, float* __restrict__ v,
                   const uint8_t* __restrict__ mask, long long n, float alpha, float omb1, float omb2, float eps) {
    pdl_entry();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (scale_terms) {                       // data parallel: 1 / (n_valid summed over ranks), still on the device
        const double nv = scale_terms[0];
        grad_scale = nv > 0.0 ? static_cast<float>(1.0 / nv) : 0.f;
    }
    // TF1 ApplyAdam functor: m += (g-m)(1-b1); v += (g^2-v)(1-b2); var -= (m*alpha)/(sqrt(v)+eps)
    const float gi = __fmul_rn(g[i], grad_scale);
    const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gi, m[i]), omb1));
    const float vi = __fadd_rn(v[i], __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), v[i]), omb2));
    m[i] = mi;
    v[i] = vi;
    if (!mask || mask[i]) p[i] = __fsub_rn(p[i], __fdiv_rn(__fmul_rn(mi, alpha), __fadd_rn(__fsqrt_rn(vi), eps)));
}

__global__ void __launch_bounds__(256)
abs_delta_kernel(const float* __restrict__ after, const float* __restrict__ before, float* __restrict__ d, long long n) {
    pdl_entry();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = fabsf(__fsub_rn(after[i], before[i]));
}

__global__ void select_init_kernel(SelectScratch* sc, unsigned int rank) {
    pdl_entry();
    if (threadIdx.x < 256) sc->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sc->prefix = 0; sc->rank = rank; sc->count_le = 0; sc->next_gt = 0xffffffffu; sc->kept = 0; }
}

// histogram of byte (key >> shift) & 0xff over keys whose bits above shift+8 equal the current prefix
__global__ void __launch_bounds__(256)
radix_hist_kernel(const float* __restrict__ d, long long n, int shift, SelectScratch* sc) {
    pdl_entry();
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned int prefix = sc->prefix;
    const unsigned int hi_mask = (shift >= 24) ? 0u : (0xffffffffu << (shift + 8));
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned int key = __float_as_uint(d[i]);
        if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&s_h[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&sc->hist[threadIdx.x], s_h[threadIdx.x]);
}

__global__ void radix_pick_kernel(SelectScratch* sc, int shift) {
    pdl_entry();
    if (threadIdx.x == 0) {
        unsigned int rank = sc->rank, cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
            const unsigned int h = sc->hist[b];
            if (rank < cum + h) break;
            cum += h;
        }
        if (b > 255) b = 255;
        sc->rank = rank - cum;
        sc->prefix |= static_cast<unsigned int>(b) << shift;
    }
    __syncthreads();
    if (threadIdx.x < 256) sc->hist[threadIdx.x] = 0;
}

// count of keys <= v_lo and the smallest key > v_lo
__global__ void __launch_bounds__(256)
select_neighbors_kernel(const float* __restrict__ d, long long n, SelectScratch* sc) {
    pdl_entry();
    const unsigned int v = sc->prefix;
    unsigned int cnt = 0, nxt = 0xffffffffu;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned int key = __float_as_uint(d[i]);
        if (key <= v) ++cnt; else nxt = min(nxt, key);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    nxt = __reduce_min_sync(0xffffffffu, nxt);
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&sc->count_le, cnt);
        atomicMin(&sc->next_gt, nxt);
    }
}

__global__ void select_threshold_kernel(SelectScratch* sc, unsigned int lo, unsigned int n, double w_lo, double w_hi) {
    pdl_entry();
    if (threadIdx.x != 0) return;
    const unsigned int v_lo = sc->prefix;
    sc->v_lo = v_lo;
    unsigned int v_hi = v_lo;
    if (lo + 1 <= n - 1 && !(sc->count_le > lo + 1)) v_hi = sc->next_gt;     // a[lo+1] is the next distinct value
    const double a_lo = static_cast<double>(__uint_as_float(v_lo));
    const double a_hi = static_cast<double>(__uint_as_float(v_hi));
    // NumPy 1.19: x1 = a[lo]*w_below ; x2 = a[hi]*w_above ; r = x1 + x2 (float64), then float32 for the compare
    const double thr = __dadd_rn(__dmul_rn(a_lo, w_lo), __dmul_rn(a_hi, w_hi));
    sc->threshold = static_cast<float>(thr);
}

__global__ void __launch_bounds__(256)
select_apply_kernel(float* __restrict__ after, const float* __restrict__ before, const float* __restrict__ d,
                    uint8_t* __restrict__ mask, long long n, SelectScratch* sc) {
    pdl_entry();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const float thr = sc->threshold;
    bool keep = false;
    if (i < n) {
        keep = d[i] > thr;
        mask[i] = keep ? 1 : 0;
        if (!keep) after[i] = before[i];
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sc->kept, static_cast<unsigned long long>(__popc(b)));
}

// ------------------------------------------------------------------------------------------ delta packing
__global__ void __launch_bounds__(256)
pack_bits_kernel(const uint8_t* __restrict__ mask, const VarSeg* __restrict__ segs, int nseg, long long total_bytes,
                 uint8_t* __restrict__ out) {
    pdl_entry();
    const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= total_bytes) return;
    int lo = 0, hi = nseg - 1;                       // last segment with bit_byte_offset <= b
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].bit_byte_offset <= b) lo = mid; else hi = mid - 1;
    }
    const VarSeg s = segs[lo];
    const long long e0 = (b - s.bit_byte_offset) * 8;
    unsigned int byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const long long e = e0 + k;
        if (e < s.size && mask[s.offset + e]) byte |= (0x80u >> k);     // np.packbits: big-endian bit order
    }
    out[b] = static_cast<uint8_t>(byte);
}

constexpr int kPackBlock = 1024;
__global__ void __launch_bounds__(256)
pack_count_kernel(const uint8_t* __restrict__ mask, long long n, unsigned int* __restrict__ counts) {
    pdl_entry();
    __shared__ unsigned int s_c[8];
    const long long base = static_cast<long long>(blockIdx.x) * kPackBlock;
    unsigned int c = 0;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        if (i < n && mask[i]) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int i = 0; i < 8; ++i) t += s_c[i];
        counts[blockIdx.x] = t;
    }
}
__global__ void pack_scan_kernel(unsigned int* counts, int nblocks, unsigned long long* kept_out) {
    pdl_entry();
    // single thread exclusive scan (<= ~2100 blocks)
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < nblocks; ++i) { const unsigned int c = counts[i]; counts[i] = static_cast<unsigned int>(run); run += c; }
        *kept_out = run;
    }
}
__global__ void __launch_bounds__(256)
pack_scatter_kernel(const float* __restrict__ params, const uint8_t* __restrict__ mask, long long n,
                    const unsigned int* __restrict__ offsets, __half* __restrict__ out) {
    pdl_entry();
    __shared__ unsigned int s_w[8];
    const long long base = static_cast<long long>(blockIdx.x) * kPackBlock;
    bool keep[4]; unsigned int c = 0;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        keep[k] = (i < n) && mask[i];
        c += keep[k];
    }
    // exclusive scan of c across the block (thread order == element order)
    unsigned int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned int wbase = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
    unsigned int pos = offsets[blockIdx.x] + wbase + inc - c;
    for (int k = 0; k < 4; ++k) {
        if (keep[k]) out[pos++] = __float2half_rn(params[base + threadIdx.x * 4 + k]);
    }
}

// fp32 HWIO [Cin][Cout] -> fp16 [Cout][ld_fwd] (transposed, hi and optional lo plane) and bf16 [Cin][ld_bwd].
// One 32x32 tile per block through shared memory: coalesced reads along Cout, coalesced transposed writes along Cin
// (a thread per element wrote the transposed planes 2 bytes at a time with a stride of ld_fwd: 58 us per step).
__global__ void __launch_bounds__(256)
cast_weights_kernel(const WeightCast* __restrict__ table) {
    pdl_entry();
    __shared__ float tile[32][33];
    const WeightCast t = table[blockIdx.y];
    const int tiles_co = (t.Cout + 31) >> 5, tiles_ci = (t.rows + 31) >> 5;
    if (static_cast<int>(blockIdx.x) >= tiles_co * tiles_ci) return;
    const int ci0 = (blockIdx.x / tiles_co) << 5, co0 = (blockIdx.x % tiles_co) << 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        float w = 0.f;
        if (ci < t.rows && co < t.Cout) {
            w = t.w[static_cast<long long>(t.row0 + ci) * t.Cout + co];
            if (t.w_bwd) t.w_bwd[static_cast<long long>(ci) * t.ld_bwd + co] = __float2bfloat16_rn(w);
        }
        tile[r][tx] = w;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        if (co < t.Cout && ci < t.rows) {
            const float w = tile[tx][r];
            const act_t hi = __float2half_rn(w);                                                      // |w| << 65504
            if (t.w_fwd) t.w_fwd[static_cast<long long>(co) * t.ld_fwd + ci] = hi;
            if (t.w_lo) t.w_lo[static_cast<long long>(co) * t.ld_fwd + ci] = __float2half_rn(__fsub_rn(w, __half2float(hi)));
        }
    }
}

}  // namespace

int adam_masked(float* p, const float* g, float grad_scale, const double* scale_terms, float* m, float* v, const uint8_t* mask,
                long long n, float alpha, float omb1, float omb2, float eps, cudaStream_t s) {
    AMS_LAUNCH((adam_masked_kernel), static_cast<int>(ceil_div_ll(n, 256)), 256, 0, s, p, g, grad_scale, scale_terms, m, v, mask, n, alpha, omb1, omb2, eps);
    return 0;
}

int select_coordinates(float* after, const float* before, float* d, uint8_t* mask, long long n, long long lo,
                       double w_hi, SelectScratch* sc, cudaStream_t s) {
    AMS_REQUIRE(n > 0 && n < (1LL << 32) && lo >= 0 && lo < n, "selection size / rank out of range");
    const int nb = static_cast<int>(ceil_div_ll(n, 256));
    const int nbr = std::min(nb, 8 * kNumSMs);
    AMS_LAUNCH((abs_delta_kernel), nb, 256, 0, s, after, before, d, n);
    AMS_LAUNCH((select_init_kernel), 1, 256, 0, s, sc, static_cast<unsigned int>(lo));
    for (int shift = 24; shift >= 0; shift -= 8) {
        AMS_LAUNCH((radix_hist_kernel), nbr, 256, 0, s, d, n, shift, sc);
        AMS_LAUNCH((radix_pick_kernel), 1, 256, 0, s, sc, shift);
    }
    AMS_LAUNCH((select_neighbors_kernel), nbr, 256, 0, s, d, n, sc);
    AMS_LAUNCH((select_threshold_kernel), 1, 32, 0, s, sc, static_cast<unsigned int>(lo), static_cast<unsigned int>(n), 1.0 - w_hi, w_hi);
    AMS_LAUNCH((select_apply_kernel), nb, 256, 0, s, after, before, d, mask, n, sc);
    return 0;
}

int pack_delta_blocks(long long n) { return static_cast<int>(ceil_div_ll(n, kPackBlock)); }

int pack_delta(const float* params, const uint8_t* mask, const VarSeg* segs_dev, int nseg, long long n,
               long long mask_bytes, uint8_t* out_bits, __half* out_vals, unsigned int* block_counts, int nblocks_alloc,
               unsigned long long* kept_out, cudaStream_t s) {
    const int nblocks = pack_delta_blocks(n);
    AMS_REQUIRE(nblocks <= nblocks_alloc, "pack_delta scratch too small");
    AMS_LAUNCH((pack_bits_kernel), static_cast<int>(ceil_div_ll(mask_bytes, 256)), 256, 0, s, mask, segs_dev, nseg, mask_bytes, out_bits);
    AMS_LAUNCH((pack_count_kernel), nblocks, 256, 0, s, mask, n, block_counts);
    AMS_LAUNCH((pack_scan_kernel), 1, 32, 0, s, block_counts, nblocks, kept_out);
    AMS_LAUNCH((pack_scatter_kernel), nblocks, 256, 0, s, params, mask, n, block_counts, out_vals);
    return 0;
}

// ---- inverse of pack_delta: the client side of the model stream (the reference only SIZES the delta, run.py:316-336)
__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint8_t* __restrict__ bits, const VarSeg* __restrict__ segs, int nseg, long long total_bytes,
                   uint8_t* __restrict__ mask) {
    pdl_entry();
    const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= total_bytes) return;
    int lo = 0, hi = nseg - 1;                       // last segment with bit_byte_offset <= b
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].bit_byte_offset <= b) lo = mid; else hi = mid - 1;
    }
    const VarSeg s = segs[lo];
    const long long e0 = (b - s.bit_byte_offset) * 8;
    const unsigned int byte = bits[b];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const long long e = e0 + k;
        if (e < s.size) mask[s.offset + e] = (byte >> (7 - k)) & 1u;      // np.unpackbits: big-endian bit order
    }
}
__global__ void __launch_bounds__(256)
unpack_scatter_kernel(float* __restrict__ params, const uint8_t* __restrict__ mask, long long n,
                      const unsigned int* __restrict__ offsets, const __half* __restrict__ vals) {
    pdl_entry();
    __shared__ unsigned int s_w[8];
    const long long base = static_cast<long long>(blockIdx.x) * kPackBlock;
    bool keep[4]; unsigned int c = 0;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        keep[k] = (i < n) && mask[i];
        c += keep[k];
    }
    unsigned int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned int wbase = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
    unsigned int pos = offsets[blockIdx.x] + wbase + inc - c;
    for (int k = 0; k < 4; ++k) {
        if (keep[k]) params[base + threadIdx.x * 4 + k] = __half2float(vals[pos++]);
    }
}

int unpack_delta_mask(const uint8_t* bits, const VarSeg* segs_dev, int nseg, long long n, long long mask_bytes, uint8_t* mask,
                      unsigned int* block_counts, int nblocks, unsigned long long* kept_out, cudaStream_t s) {
    AMS_LAUNCH((unpack_bits_kernel), static_cast<int>(ceil_div_ll(mask_bytes, 256)), 256, 0, s, bits, segs_dev, nseg, mask_bytes, mask);
    AMS_LAUNCH((pack_count_kernel), nblocks, 256, 0, s, mask, n, block_counts);
    AMS_LAUNCH((pack_scan_kernel), 1, 32, 0, s, block_counts, nblocks, kept_out);
    return 0;
}
int unpack_delta_values(float* params, const uint8_t* mask, long long n, const unsigned int* block_offsets, int nblocks,
                        const __half* vals, cudaStream_t s) {
    AMS_LAUNCH((unpack_scatter_kernel), nblocks, 256, 0, s, params, mask, n, block_offsets, vals);
    return 0;
}

int cast_weights(const WeightCast* table_dev, int n_layers, int max_elems, cudaStream_t s) {
    // upper bound of the 32x32 tiles of any layer: rows*Cout/1024 plus the ragged edges (rows, Cout <= 4096)
    dim3 grid(ceil_div(max_elems, 1024) + 2 * 128 + 1, n_layers);
    AMS_LAUNCH((cast_weights_kernel), grid, 256, 0, s, table_dev);
    return 0;
}

}  // namespace ams
