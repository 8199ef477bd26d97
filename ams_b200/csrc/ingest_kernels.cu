// Frame ingest on the device: the host-side `cv2.resize` + BGR->RGB that the reference runs on every
// camera frame before the student (run.py:181-183, :263-264, :415-421), bit-exact with OpenCV's uint8 paths
// (imgproc/src/resize.cpp): INTER_LINEAR in 11-bit fixed point (x taps and weights clamped at the borders, y taps
// clipped only), the exact-2x decimation rerouted to INTER_AREA, INTER_NEAREST for teacher label maps.
// Restated and pinned against cv2 itself in oracle/cv2_resize_oracle.py.  HBM-bound byte work: one thread per output
// pixel, the four taps of neighbouring pixels overlap in L1/L2.
#include "kernels.cuh"

namespace ams {
namespace {

struct ResizeGeom { int n, sh, sw, dh, dw; double scale_x, scale_y; int swap_rb, area2x; };

__device__ __forceinline__ void lin_coeff(int d, double scale, int sn, bool clamp_weight, int* i0, int* i1, int* c0, int* c1) {
    float f = static_cast<float>(__dsub_rn(__dmul_rn(d + 0.5, scale), 0.5));      // two roundings like the host code, no FMA
    int i = static_cast<int>(floorf(f));
    f -= static_cast<float>(i);
    if (clamp_weight) {
        if (i < 0) { i = 0; f = 0.f; }
        if (i >= sn - 1) { i = sn - 1; f = 0.f; }
    }
    *c1 = __float2int_rn(f * 2048.f);
    *c0 = __float2int_rn((1.f - f) * 2048.f);
    *i0 = min(max(i, 0), sn - 1);
    *i1 = min(max(i + 1, 0), sn - 1);
}

template <int CN>
__global__ void __launch_bounds__(256)
resize_linear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, ResizeGeom g) {
    pdl_entry();
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.n) * g.dh * g.dw;
    if (t >= total) return;
    const int dx = static_cast<int>(t % g.dw);
    const int dy = static_cast<int>((t / g.dw) % g.dh);
    const int n = static_cast<int>(t / (static_cast<long long>(g.dw) * g.dh));
    const uint8_t* img = src + static_cast<long long>(n) * g.sh * g.sw * CN;
    uint8_t* o = dst + t * CN;
    if (g.area2x) {
        const uint8_t* p0 = img + (static_cast<long long>(2 * dy) * g.sw + 2 * dx) * CN;
        const uint8_t* p1 = p0 + static_cast<long long>(g.sw) * CN;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const int v = (p0[c] + p0[CN + c] + p1[c] + p1[CN + c] + 2) >> 2;
            o[(g.swap_rb && CN == 3) ? 2 - c : c] = static_cast<uint8_t>(v);
        }
        return;
    }
    int x0, x1, a0, a1, y0, y1, b0, b1;
    lin_coeff(dx, g.scale_x, g.sw, true, &x0, &x1, &a0, &a1);
    lin_coeff(dy, g.scale_y, g.sh, false, &y0, &y1, &b0, &b1);
    const uint8_t* r0 = img + static_cast<long long>(y0) * g.sw * CN;
    const uint8_t* r1 = img + static_cast<long long>(y1) * g.sw * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        const int s0 = r0[x0 * CN + c] * a0 + r0[x1 * CN + c] * a1;
        const int s1 = r1[x0 * CN + c] * a0 + r1[x1 * CN + c] * a1;
        int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        o[(g.swap_rb && CN == 3) ? 2 - c : c] = static_cast<uint8_t>(v);
    }
}

__global__ void __launch_bounds__(256)
resize_nearest_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, ResizeGeom g) {
    pdl_entry();
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(g.n) * g.dh * g.dw;
    if (t >= total) return;
    const int dx = static_cast<int>(t % g.dw);
    const int dy = static_cast<int>((t / g.dw) % g.dh);
    const int n = static_cast<int>(t / (static_cast<long long>(g.dw) * g.dh));
    const int sx = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(dx), g.scale_x))), g.sw - 1);
    const int sy = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(dy), g.scale_y))), g.sh - 1);
    dst[t] = src[(static_cast<long long>(n) * g.sh + sy) * g.sw + sx];
}

}  // namespace

int resize_u8(const uint8_t* src, int n, int sh, int sw, int cn, uint8_t* dst, int dh, int dw, int nearest, int swap_rb,
              cudaStream_t s) {
    AMS_REQUIRE(n > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, "empty resize");
    AMS_REQUIRE(cn == 1 || cn == 3, "resize: 1 or 3 channels");
    ResizeGeom g;
    g.n = n; g.sh = sh; g.sw = sw; g.dh = dh; g.dw = dw; g.swap_rb = swap_rb;
    g.scale_x = 1.0 / (static_cast<double>(dw) / sw);           // resize.cpp: scale = 1. / inv_scale
    g.scale_y = 1.0 / (static_cast<double>(dh) / sh);
    g.area2x = (!nearest && sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
    const long long total = static_cast<long long>(n) * dh * dw;
    const int blocks = static_cast<int>(ceil_div_ll(total, 256));
    if (nearest) {
        AMS_REQUIRE(cn == 1, "nearest resize is for single-channel label maps");
        AMS_LAUNCH((resize_nearest_u8_kernel), blocks, 256, 0, s, src, dst, g);
    } else if (cn == 3) {
        AMS_LAUNCH((resize_linear_u8_kernel<3>), blocks, 256, 0, s, src, dst, g);
    } else {
        AMS_LAUNCH((resize_linear_u8_kernel<1>), blocks, 256, 0, s, src, dst, g);
    }
    return 0;
}

}  // namespace ams
