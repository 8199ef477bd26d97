// Shared device/host helpers for the sm_100a kernels of libams_b200.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ams {

// ----------------------------------------------------------------------------- errors
void set_last_error(const std::string& msg);
#define AMS_CUDA_CHECK(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ams::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +    \
                                __FILE__ + ":" + std::to_string(__LINE__));                      \
            return -1;                                                                           \
        }                                                                                        \
    } while (0)
#define AMS_REQUIRE(cond, msg)                                                                   \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ams::set_last_error(std::string("requirement failed: ") + #cond + " -- " + (msg) +  \
                                " at " + __FILE__ + ":" + std::to_string(__LINE__));             \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)
void count_launch();
void count_launches(long long n);     // kernels inside a replayed CUDA graph
long long launches_so_far();
#define AMS_LAUNCH_CHECK()                      \
    do {                                        \
        ams::count_launch();                    \
        AMS_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the step starts with pdl_entry(): it lets the NEXT kernel of the stream be scheduled as soon as all
// CTAs of this one are resident (its CTAs then park in griddepcontrol.wait), and waits for the PREVIOUS kernel to
// complete and flush before touching global memory.  Launch latency and prologues overlap the predecessor's tail.
// Kernels that allocate tensor memory trigger only after their tcgen05.alloc (a parked dependent holding TMEM
// columns must never be what a not-yet-allocated CTA of the running kernel is waiting for).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// AMS_LAUNCH((kernel<T>), grid, block, smem, stream, args...): PDL launch + launch counter + error check
#define AMS_STRIP_PARENS(...) __VA_ARGS__
#define AMS_LAUNCH(kernel, grid, block, smem, stream, ...)                                                       \
    do {                                                                                                         \
        ams::count_launch();                                                                                     \
        AMS_CUDA_CHECK(ams::launch_pdl(AMS_STRIP_PARENS kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)); \
    } while (0)

constexpr int kNumSMs = 148;   // B200; grids of the persistent kernels are sized from the runtime value

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- bf16 pack / unpack
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
// ----------------------------------------------------------------------------- fp16 pack / unpack
// Forward activations and the 1x1-conv weight operands are stored in IEEE fp16 (10-bit mantissa: 8x finer than
// bf16 for the same bytes; post-BN/ReLU6 values live in [0, 6], raw conv outputs stay far below 65504 and the
// conversion saturates instead of producing inf).  Gradient tensors keep bf16 (range matters there, not precision).
__device__ __forceinline__ float h16_lo(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xffffu))); }
__device__ __forceinline__ float h16_hi(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16))); }
__device__ __forceinline__ float2 h16x2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ uint32_t pack_h16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
    const float2 a = h16x2(v.x), b = h16x2(v.y), c = h16x2(v.z), d = h16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
    return make_uint4(pack_h16(f[0], f[1]), pack_h16(f[2], f[3]), pack_h16(f[4], f[5]), pack_h16(f[6], f[7]));
}
// compile-time choice of the 16-bit storage type: H = true fp16 (activations), false bf16 (gradients)
template <bool H> __device__ __forceinline__ float2 unpack2t(uint32_t v) { return H ? h16x2(v) : make_float2(bf16_lo(v), bf16_hi(v)); }
template <bool H> __device__ __forceinline__ uint32_t pack2t(float lo, float hi) { return H ? pack_h16(lo, hi) : pack_bf16(lo, hi); }
template <bool H> __device__ __forceinline__ void unpack8t(const uint4& v, float* f) { if (H) unpack8h(v, f); else unpack8(v, f); }
template <bool H> __device__ __forceinline__ uint4 pack8t(const float* f) { return H ? pack8h(f) : pack8(f); }

// streaming 128-bit loads/stores (read-once / write-once tensors: keep them out of L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 256-bit store (sm_100+): one full 32-byte sector per thread
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

__device__ __forceinline__ float act_apply(float v, int act) {   // 0 none, 1 relu, 2 relu6
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ams
