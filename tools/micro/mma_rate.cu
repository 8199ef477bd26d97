// Microbenchmark: how fast does one SM retire a chain of tcgen05.mma (kind::f16, M = 128) instructions as a function of N,
// of the operand layouts this repo uses, of whether consecutive instructions accumulate into the same TMEM columns, and
// of HOW the instruction is issued (one lane under `if (lane == 0)` vs the whole warp + elect.sync)?  Descriptors are
// precomputed and the inner 8 instructions are unrolled, so the loop itself costs nothing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ams_b200/csrc -I include -o /tmp/mma_rate tools/micro/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace ams;

struct Cfg { int n, a_mn, sw128, reps, commit_every; };

template <int ACCS, int WARP, int CE>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint64_t side_bar[8];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { t5::mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) t5::mbar_init(&side_bar[i], 1); t5::fence_barrier_init(); }
    if (threadIdx.x < 32) { t5::tmem_alloc(&tmem_slot, 512); t5::tmem_relinquish(); }
    t5::fence_proxy_async_smem();
    t5::fence_before_thread_sync();
    __syncthreads();
    t5::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (WARP ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
        const uint32_t a_addr = t5::smem_u32(smem), b_addr = a_addr + 48 * 1024;
        const uint32_t idesc = t5::make_idesc_f16(128, c.n, c.a_mn, 0, 0, 0);
        uint64_t da[8], db[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c.a_mn) da[k] = t5::make_smem_desc_sw128(a_addr + k * 2048, 16384, 1024);
            else if (c.sw128) da[k] = t5::make_smem_desc_sw128(a_addr + (k & 3) * 32, 16, 1024);
            else da[k] = t5::make_smem_desc(a_addr + (k & 1) * 32, 16, 512, 4);
            if (c.sw128) db[k] = t5::make_smem_desc_sw128(b_addr + (k & 3) * 32, 16, 1024);
            else db[k] = t5::make_smem_desc(b_addr + (k & 1) * 32, 16, 512, 4);
        }
        for (int round = 0; round < 2; ++round) {
            const long long t0 = clock64();
            for (int i = 0; i < c.reps; i += 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (WARP) t5::mma_f16_ss_warp(tmem + (k % ACCS) * c.n, da[k], db[k], idesc, (i | (k >= ACCS)) != 0);
                    else t5::mma_bf16_ss(tmem + (k % ACCS) * c.n, da[k], db[k], idesc, (i | (k >= ACCS)) != 0);
                    if (CE > 0 && (k % (CE > 0 ? CE : 1)) == CE - 1) {
                        if (WARP) t5::mma_commit_warp(&side_bar[k]); else t5::mma_commit(&side_bar[k]);
                    }
                }
            }
            const long long t1 = clock64();
            if (WARP) t5::mma_commit_warp(&bar); else t5::mma_commit(&bar);
            while (!t5::mbar_try_wait(&bar, round & 1)) { }
            const long long t2 = clock64();
            if (threadIdx.x == 0) { out[round * 2] = t1 - t0; out[round * 2 + 1] = t2 - t0; }
        }
    }
    t5::fence_before_thread_sync();
    __syncthreads();
    if (threadIdx.x < 32) { t5::fence_after_thread_sync(); t5::tmem_dealloc(tmem, 512); }
}

template <int ACCS, int WARP, int CE>
void run(const Cfg& c, long long* out) {
    cudaFuncSetAttribute(rate_kernel<ACCS, WARP, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    rate_kernel<ACCS, WARP, CE><<<1, 128, 100 * 1024>>>(c, out);
    long long h[4];
    cudaError_t e = cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    printf("%-10s commit/%d %4d %5s %6s %5d | %13.1f %14.1f %13.1f\n", WARP ? "warp+elect" : "lane0", CE, c.n, c.a_mn ? "MN" : "K", c.sw128 ? "128B" : "64B", ACCS,
           h[2] / double(c.reps), h[3] / double(c.reps), 128.0 * c.n * 16 * 2 / 8192.0);
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    printf("%-10s %4s %5s %6s %5s | %13s %14s %13s\n", "issue", "N", "A", "swz", "accs", "issue cyc/mma", "retire cyc/mma", "ideal cyc/mma");
    const int ns[] = {32, 64, 128, 192, 256};
    for (int layout = 0; layout < 3; ++layout)
        for (int n : ns) {
            if (layout != 2 && n != 64) continue;
            Cfg c{n, layout == 2, layout >= 1, 256, 0};
            run<1, 0, 0>(c, out); run<1, 1, 0>(c, out);
            run<1, 0, 4>(c, out); run<1, 1, 4>(c, out);
            run<1, 0, 1>(c, out); run<1, 1, 1>(c, out);
        }
    return 0;
}
