// Microbenchmark: tcgen05.mma (M = 128, N = 64, same accumulator, K-major SWIZZLE_64B) retire rate on one SM while 8 other
// warps of the CTA (a) idle, (b) run dependent-free FFMA streams, (c) stream tcgen05.ld from other TMEM columns,
// (d) stream st.shared, (e) spin on an mbarrier with try_wait.  Tells which resource the MMA stream shares with them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ams_b200/csrc -I include -o /tmp/mma_cont tools/micro/mma_contention.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace ams;

template <int MODE, int WARP>
__global__ void __launch_bounds__(320, 1) cont_kernel(int n, int reps, long long* out, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, never;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { t5::mbar_init(&bar, 1); t5::mbar_init(&never, 1); t5::fence_barrier_init(); stop = 0; }
    if (threadIdx.x < 32) { t5::tmem_alloc(&tmem_slot, 512); t5::tmem_relinquish(); }
    t5::fence_proxy_async_smem();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) {
        if (WARP || threadIdx.x == 32) {
            const uint32_t a_addr = t5::smem_u32(smem), b_addr = a_addr + 48 * 1024;
            const uint32_t idesc = t5::make_idesc_f16(128, n, 0, 0, 0, 0);
            const uint64_t da0 = t5::make_smem_desc(a_addr, 16, 512, 4), db0 = t5::make_smem_desc(b_addr, 16, 512, 4);
            const long long t0 = clock64();
            for (int i = 0; i < reps; i += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (WARP) t5::mma_f16_ss_warp(tmem, da0 + 2 * (k & 1), db0 + 2 * (k & 1), idesc, (i | k) != 0);
                    else t5::mma_bf16_ss(tmem, da0 + 2 * (k & 1), db0 + 2 * (k & 1), idesc, (i | k) != 0);
                }
            }
            const long long t1 = clock64();
            if (WARP) t5::mma_commit_warp(&bar); else t5::mma_commit(&bar);
            while (!t5::mbar_try_wait(&bar, 0)) { }
            const long long t2 = clock64();
            if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop = 1; }
        }
    } else if (warp >= 2) {
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x * 0.001f + k;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t my_smem = t5::smem_u32(smem) + 64 * 1024 + threadIdx.x * 16;
        while (!stop) {
            if (MODE == 1) {
#pragma unroll
                for (int r = 0; r < 64; ++r)
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], 1.0001f, 0.5f);
            } else if (MODE == 5) {          // FFMA only in warps 2..5: one compute warp on the MMA warp's scheduler
                if (warp <= 5) {
#pragma unroll
                    for (int r = 0; r < 64; ++r)
#pragma unroll
                        for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], 1.0001f, 0.5f);
                } else __nanosleep(200);
            } else if (MODE == 6) {          // FFMA with a zero-length nanosleep every 64 instructions
#pragma unroll
                for (int r = 0; r < 64; ++r) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], 1.0001f, 0.5f);
                    if ((r & 3) == 3) asm volatile("nanosleep.u32 0;");
                }
            } else if (MODE == 7) {          // FFMA, 4 independent chains only (latency-bound: ~1 issue per cycle per 4 cycles)
#pragma unroll
                for (int r = 0; r < 256; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] = fmaf(acc[k], 1.0001f, 0.5f);
            } else if (MODE == 8) {          // FFMA only in warps 2,3,6,7: NO compute warp on the MMA warp's scheduler
                if ((warp & 3) >= 2) {
#pragma unroll
                    for (int r = 0; r < 64; ++r)
#pragma unroll
                        for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], 1.0001f, 0.5f);
                } else __nanosleep(200);
            } else if (MODE == 2) {
                uint32_t raw[16];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    t5::tmem_ld16(tmem + lane_addr + 256 + (r & 7) * 16, raw);
                    t5::tmem_ld_wait();
                    acc[r & 15] += __uint_as_float(raw[r & 15]);
                }
            } else if (MODE == 3) {
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(my_smem), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
            } else if (MODE == 4) {
                for (int r = 0; r < 16; ++r) if (t5::mbar_try_wait(&never, 0)) acc[0] += 1.f;
            } else {
                __nanosleep(200);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += acc[k];
        if (s == 123.456f) sink[threadIdx.x] = s;
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (threadIdx.x < 32) { t5::fence_after_thread_sync(); t5::tmem_dealloc(tmem, 512); }
}

template <int MODE, int WARP>
void run(const char* what, int n, long long* out, float* sink) {
    cudaFuncSetAttribute(cont_kernel<MODE, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    cont_kernel<MODE, WARP><<<1, 320, 120 * 1024>>>(n, 512, out, sink);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    printf("N %3d  %-10s  other 8 warps: %-28s issue %6.1f  retire %6.1f cycles per MMA\n", n, WARP ? "warp+elect" : "lane0", what, h[0] / 512.0, h[1] / 512.0);
}

int main() {
    long long* out; float* sink; cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    for (int n : {64}) {
        run<0, 0>("sleeping", n, out, sink);             run<0, 1>("sleeping", n, out, sink);
        run<1, 0>("FFMA streams", n, out, sink);         run<1, 1>("FFMA streams", n, out, sink);
        run<2, 0>("tcgen05.ld streams", n, out, sink);   run<2, 1>("tcgen05.ld streams", n, out, sink);
        run<3, 0>("st.shared streams", n, out, sink);    run<3, 1>("st.shared streams", n, out, sink);
        run<4, 0>("mbarrier try_wait spin", n, out, sink); run<4, 1>("mbarrier try_wait spin", n, out, sink);
        run<5, 0>("FFMA in warps 2-5 only", n, out, sink); run<5, 1>("FFMA in warps 2-5 only", n, out, sink);
        run<6, 0>("FFMA + nanosleep(0)/64", n, out, sink); run<6, 1>("FFMA + nanosleep(0)/64", n, out, sink);
        run<7, 0>("FFMA 4 chains (ILP 4)", n, out, sink); run<7, 1>("FFMA 4 chains (ILP 4)", n, out, sink);
        run<8, 0>("FFMA in warps 2,3,6,7 only", n, out, sink); run<8, 1>("FFMA in warps 2,3,6,7 only", n, out, sink);
    }
    return 0;
}
