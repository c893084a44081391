// microbench.cu -- pipe-throughput probes used as roofline denominators by bench.py
// (MEASURED_PEAKS.json has no FP32 figure).  Built into libcnngp_bench.so; measurement
// infrastructure, not part of the product ABI.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace {

template <int KIND>
__global__ void __launch_bounds__(256) probe(float *out, int iters, float a, float b) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = threadIdx.x * 1e-3f + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (KIND == 0) {  // FFMA, 8 independent chains
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = fmaf(r[k], a, b);
            } else if (KIND == 1) {  // packed fma.rn.f32x2 (sm_100+)
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    asm volatile(
                        "{ .reg .b64 x, y, z;\n"
                        "  mov.b64 x, {%0, %1};\n"
                        "  mov.b64 y, {%2, %2};\n"
                        "  mov.b64 z, {%3, %3};\n"
                        "  fma.rn.f32x2 x, x, y, z;\n"
                        "  mov.b64 {%0, %1}, x; }\n"
                        : "+f"(r[k]), "+f"(r[k + 1])
                        : "f"(a), "f"(b));
                }
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    asm volatile(
                        "{ .reg .b64 x, y, z;\n"
                        "  mov.b64 x, {%0, %1};\n"
                        "  mov.b64 y, {%2, %2};\n"
                        "  mov.b64 z, {%3, %3};\n"
                        "  fma.rn.f32x2 x, x, y, z;\n"
                        "  mov.b64 {%0, %1}, x; }\n"
                        : "+f"(r[k]), "+f"(r[k + 1])
                        : "f"(a), "f"(b));
                }
            } else if (KIND == 2) {  // MUFU.SQRT
#pragma unroll
                for (int k = 0; k < 8; ++k) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(r[k]));
            } else if (KIND == 3) {  // MUFU.RCP
#pragma unroll
                for (int k = 0; k < 8; ++k) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[k]));
            } else if (KIND == 4) {  // 7 FFMA : 1 MUFU
#pragma unroll
                for (int k = 0; k < 7; ++k) r[k] = fmaf(r[k], a, b);
                asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(r[7]));
            } else if (KIND == 5) {  // FADD
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a);
            } else if (KIND == 6) {  // 6 FFMA : 2 FMNMX (alu pipe)
#pragma unroll
                for (int k = 0; k < 6; ++k) r[k] = fmaf(r[k], a, b);
                r[6] = fmaxf(r[6], a); r[7] = fminf(r[7], b);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += r[k];
    if (s == 123.456f) out[0] = s;
}

template <int VEC>
__global__ void __launch_bounds__(256) probe_lds(float *out, int iters) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (VEC == 4) {
                float4 v = buf[(idx + u * 32) & 1023];
                acc += v.x + v.y + v.z + v.w;
            } else {
                acc += reinterpret_cast<float *>(buf)[(idx + u * 32) & 4095];
            }
        }
        idx += 7;
    }
    if (acc == 123.456f) out[0] = acc;
}

}  // namespace

// returns ops-per-second of the probed instruction (lane-ops, i.e. 32 x warp instructions),
// counting a packed f32x2 fma as 2 and KIND 4/6 as the total of all 8 lane-ops per group.
extern "C" double mb_probe(int kind, int blocks_per_sm, int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *d = nullptr;
    cudaMalloc(&d, 16);
    const int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](int n) {
        switch (kind) {
            case 0: probe<0><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 1: probe<1><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 2: probe<2><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 3: probe<3><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 4: probe<4><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 5: probe<5><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 6: probe<6><<<grid, 256>>>(d, n, 1.0001f, 1e-3f); break;
            case 10: probe_lds<1><<<grid, 256>>>(d, n); break;
            case 11: probe_lds<4><<<grid, 256>>>(d, n); break;
        }
    };
    run(iters / 4 + 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        run(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    double per_thread = (kind >= 10) ? 8.0 * iters * (kind == 11 ? 4 : 1) : 8.0 * 4 * iters * (kind == 1 ? 2 : 1);
    return per_thread * 256.0 * grid / (best * 1e-3);
}
