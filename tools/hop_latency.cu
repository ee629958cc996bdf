// dev probe: latency of one producer->consumer hop through a 16-byte L2 message between SMs, alone and with
// background pollers.  nvcc -arch=sm_100a -O3 -o hop_latency hop_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ld_msg(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_msg(float4 *p, float4 v) {
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// ring of `players` blocks passing a token; other blocks poll a word that changes only at the end
__global__ void k_ring(float4 *msg, int players, int rounds, int sleep_ns, volatile unsigned int *stop, int poll_threads) {
    if ((int) blockIdx.x < players) {
        if (threadIdx.x != 0) return;
        const int me = blockIdx.x;
        for (int r = 0; r < rounds; ++r) {
            const unsigned int expect = (unsigned int) (r * players + me);
            float4 v;
            do { v = ld_msg(&msg[0]); } while (__float_as_uint(v.w) != expect);
            v.x += 1.0f;
            v.w = __uint_as_float(expect + 1u);
            st_msg(&msg[0], v);
        }
        if (me == players - 1) *stop = 1u;
    } else {
        if ((int) threadIdx.x >= poll_threads) return;
        const float4 *p = &msg[1 + (blockIdx.x * blockDim.x + threadIdx.x) % 4096];
        float acc = 0.f;
        while (*stop == 0u) {
            float4 v = ld_msg(p);
            acc += v.x;
            if (sleep_ns > 0) __nanosleep(sleep_ns);
        }
        if (acc == 12345.f) printf("x");
    }
}
int main() {
    float4 *msg; unsigned int *stop;
    cudaMalloc(&msg, 8192 * sizeof(float4)); cudaMalloc(&stop, 4);
    const int rounds = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int players : {2, 8}) for (int pollers : {0, 146, 146 * 2, 146 * 4}) for (int pt : {32, 256}) for (int sl : {0, 100, 1000}) {
        if (pollers == 0 && (pt != 32 || sl != 0)) continue;
        cudaMemset(msg, 0, 8192 * sizeof(float4)); cudaMemset(stop, 0, 4);
        cudaEventRecord(a);
        k_ring<<<players + pollers, 256>>>(msg, players, rounds, sl, stop, pt);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("players %d  poll blocks %4d x %3d thr  sleep %4d ns : %.3f us per hop  (%s)\n", players, pollers, pt, sl,
               ms * 1e3 / (rounds * players), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
