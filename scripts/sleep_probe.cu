// sleep_probe.cu -- how long nanosleep / mbarrier.try_wait (with a suspend-time hint) really hold a thread (measurement only)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__global__ void probe(long long *out, int mode, unsigned ns) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    __syncthreads();
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        for (int it = 0; it < 1000; ++it) {
            if (mode == 0) __nanosleep(ns);
            else {
                uint32_t ok;
                if (mode == 1)
                    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(0u), "r"(ns) : "memory");
                else
                    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(0u) : "memory");
                if (ok) break;
            }
        }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        out[0] = clock64() - t0;
        out[1] = (long long)(g1 - g0);
    }
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    for (int mode = 0; mode < 3; ++mode)
        for (unsigned ns : {100u, 1000u, 10000u, 1000000u, 10000000u}) {
            probe<<<1, 32>>>(d, mode, ns);
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s arg %8u ns: %9.1f cycles, %9.1f ns per call\n", mode == 0 ? "nanosleep        " : mode == 1 ? "try_wait + hint  " : "try_wait         ", ns, h[0] / 1000.0, h[1] / 1000.0);
            if (mode == 2) break;
        }
    return 0;
}
