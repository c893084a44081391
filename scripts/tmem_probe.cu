// tmem_probe.cu -- what tensor-memory traffic costs a warp on sm_100a (measurement only):
// cycles per operation of tcgen05.ld / tcgen05.st / tcgen05.wait patterns with 4, 8 or 12 warps of
// one CTA per SM active (the fused-net kernel's layouts: 3 warps share a 32-lane quadrant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_probe scripts/tmem_probe.cu && /tmp/tmem_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void ld16(uint32_t ta, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta));
}
__device__ __forceinline__ void ld4(uint32_t ta, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
}
__device__ __forceinline__ void ld2(uint32_t ta, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ta));
}
__device__ __forceinline__ void st2(uint32_t ta, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(ta), "r"(a), "r"(b));
}
__device__ __forceinline__ void st16(uint32_t ta, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;"); }

// MODE 0: 7 x ld16 in flight + one wait (a 112-column map set)      -> cycles per iteration
// MODE 1: ld16 + wait, 7 times (one load in flight)
// MODE 2: 56 x st2 + wait::st
// MODE 3: 7 x st16 + wait::st
// MODE 4: ld2 + wait, 56 times (latency of the smallest load)
// MODE 5: wait::ld alone, 56 times (nothing outstanding)
// MODE 6: 28 x ld4 in flight + one wait
// MODE 7: MODE 0 interleaved with 448 independent FFMA per iteration (does the read hide under arithmetic?)
// MODE 8: the 448 FFMA alone
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long *out, int iters, int active_warps) {
    __shared__ uint32_t tmem_word;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_word)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_word + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 160);
    uint32_t r[7][16];
    uint32_t acc = 0;
    float f[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) f[q] = threadIdx.x * 1e-3f + q;
#pragma unroll
    for (int g = 0; g < 7; ++g)
#pragma unroll
        for (int q = 0; q < 16; ++q) r[g][q] = threadIdx.x + g * 16 + q;
    if (warp < active_warps) {  // initialise the columns
#pragma unroll
        for (int g = 0; g < 7; ++g) st16(base + g * 16, r[g]);
        wait_st();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp < active_warps) {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0 || MODE == 7) {
#pragma unroll
                for (int g = 0; g < 7; ++g) ld16(base + g * 16, r[g]);
                if (MODE == 7) {
#pragma unroll
                    for (int u = 0; u < 28; ++u)
#pragma unroll
                        for (int q = 0; q < 16; ++q) f[q] = fmaf(f[q], 1.0001f, 0.5f);
                }
                wait_ld();
#pragma unroll
                for (int g = 0; g < 7; ++g) {
                    asm volatile("" : "+r"(r[g][0]), "+r"(r[g][15]));
                    acc += r[g][0] ^ r[g][15];
                }
            } else if (MODE == 8) {
#pragma unroll
                for (int u = 0; u < 28; ++u)
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = fmaf(f[q], 1.0001f, 0.5f);
            } else if (MODE == 1) {
#pragma unroll
                for (int g = 0; g < 7; ++g) {
                    ld16(base + g * 16, r[g]);
                    wait_ld();
                    asm volatile("" : "+r"(r[g][0]), "+r"(r[g][15]));
                    acc += r[g][0] ^ r[g][15];
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int g = 0; g < 7; ++g)
#pragma unroll
                    for (int q = 0; q < 8; ++q) st2(base + g * 16 + 2 * q, r[g][2 * q] + it, r[g][2 * q + 1]);
                wait_st();
            } else if (MODE == 3) {
#pragma unroll
                for (int g = 0; g < 7; ++g) { r[g][0] += it; st16(base + g * 16, r[g]); }
                wait_st();
            } else if (MODE == 4) {
#pragma unroll
                for (int g = 0; g < 56; ++g) {
                    uint32_t t[2];
                    ld2(base + 2 * g, t);
                    wait_ld();
                    asm volatile("" : "+r"(t[0]), "+r"(t[1]));
                    acc += t[0] ^ t[1];
                }
            } else if (MODE == 5) {
#pragma unroll
                for (int g = 0; g < 56; ++g) { wait_ld(); acc += g; }
            } else if (MODE == 6) {
                uint32_t t[28][4];
#pragma unroll
                for (int g = 0; g < 28; ++g) ld4(base + 4 * g, t[g]);
                wait_ld();
#pragma unroll
                for (int g = 0; g < 28; ++g) {
                    asm volatile("" : "+r"(t[g][0]), "+r"(t[g][3]));
                    acc += t[g][0] ^ t[g][3];
                }
            }
        }
    }
    const long long t1 = clock64();
    float fs = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) fs += f[q];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u || fs == 1.2345f) out[blockIdx.x + 1000] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_word), "r"(512));
    }
}

template <int MODE>
void run(const char *name, long long *d_out) {
    const int iters = 2000;
    for (int aw : {1, 4, 8, 12}) {
        probe<MODE><<<148, 512>>>(d_out, 10, aw);
        probe<MODE><<<148, 512>>>(d_out, iters, aw);
        long long h = 0;
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%-48s warps %2d: %8.1f cycles / iteration%s\n", name, aw, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 8 * 4096);
    run<0>("7 x ld16 in flight + wait (14 KB per warp)", d_out);
    run<1>("7 x (ld16 + wait)", d_out);
    run<6>("28 x ld4 in flight + wait", d_out);
    run<4>("56 x (ld2 + wait)", d_out);
    run<5>("56 x wait::ld, nothing outstanding", d_out);
    run<2>("56 x st2 + wait::st", d_out);
    run<3>("7 x st16 + wait::st", d_out);
    run<8>("448 FFMA alone", d_out);
    run<7>("7 x ld16 + 448 FFMA + wait", d_out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
