// Micro-benchmark: cost of one dependent kernel in a stream-captured CUDA graph on this GPU, with and without
// programmatic dependent launch, for an empty kernel and for one that touches a little memory.  (tools/ only.)
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_empty(float* p, int pdl) {
    if (pdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    if (threadIdx.x == 0 && blockIdx.x == 0 && p[0] < 0) p[1] = 1.f;
}
__global__ void k_small(float* p, int n, int pdl) {
    if (pdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
template <typename F> float run(int chain, bool graph, F launch) {
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphExec_t exec = nullptr;
    if (graph) {
        cudaGraph_t g; cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < chain; ++i) launch(s);
        cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&exec, g, 0); cudaGraphDestroy(g);
    }
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        if (graph) cudaGraphLaunch(exec, s); else for (int i = 0; i < chain; ++i) launch(s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    if (exec) cudaGraphExecDestroy(exec);
    cudaStreamDestroy(s);
    return best * 1000.f / chain;
}
int main() {
    float* p; const int n = 1 << 20; cudaMalloc(&p, n * 4); cudaMemset(p, 0, n * 4);
    const int chain = 1000;
    for (int pdl = 0; pdl < 2; ++pdl) {
        auto cfgl = [&](auto kern, dim3 g, dim3 b, cudaStream_t s, auto... args) {
            cudaLaunchConfig_t c = {}; c.gridDim = g; c.blockDim = b; c.stream = s;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            c.attrs = at; c.numAttrs = pdl ? 1 : 0;
            cudaLaunchKernelEx(&c, kern, args...);
        };
        for (int graph = 0; graph < 2; ++graph) {
            float a = run(chain, graph, [&](cudaStream_t s) { cfgl(k_empty, dim3(1), dim3(32), s, p, pdl); });
            float b = run(chain, graph, [&](cudaStream_t s) { cfgl(k_empty, dim3(296), dim3(256), s, p, pdl); });
            float c = run(chain, graph, [&](cudaStream_t s) { cfgl(k_small, dim3(n / 256), dim3(256), s, p, n, pdl); });
            printf("pdl=%d graph=%d: empty 1 CTA %.2f us | empty 296 CTAs %.2f us | 4 MB rmw 4096 CTAs %.2f us per kernel\n", pdl, graph, a, b, c);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
