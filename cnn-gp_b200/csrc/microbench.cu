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

// FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), CH independent accumulator chains per warp
template <int CH>
__global__ void __launch_bounds__(256) probe_dmma(float *out, int iters) {
    double acc[CH][2];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k][0] = acc[k][1] = threadIdx.x * 1e-3 + k;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < CH; ++k)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[k][0]), "+d"(acc[k][1])
                             : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += acc[k][0] + acc[k][1];
    if (s == 123.456) out[0] = (float)s;
}

// DFMA, 8 independent chains
__global__ void __launch_bounds__(256) probe_dfma(float *out, int iters, double a, double b) {
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = fma(r[k], a, b);
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += r[k];
    if (s == 123.456) out[0] = (float)s;
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
            case 20: probe_dmma<8><<<grid, 256>>>(d, n); break;
            case 21: probe_dfma<<<grid, 256>>>(d, n, 1.0000001, 1e-9); break;
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
    // kind 20: FMA-equivalents of the DMMA chain (8 x 4 mma per iteration per warp, 256 FMA each);
    // kind 21: DFMA lane-ops
    if (kind == 20) return 8.0 * 4 * iters * 256.0 * (256 / 32) * grid / (best * 1e-3);
    double per_thread = (kind >= 10) ? 8.0 * iters * (kind == 11 ? 4 : 1) : 8.0 * 4 * iters * (kind == 1 ? 2 : 1);
    return per_thread * 256.0 * grid / (best * 1e-3);
}

// ---- ReLU-step probes: the inner loop of gram_fused.cu in isolation --------------------------
namespace {

__device__ __forceinline__ float mb_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float mb_h2(float e) {
    float h = 7.577116048e-05f;
    h = fmaf(h, e, -5.945927478e-05f);
    h = fmaf(h, e, 2.400144585e-04f);
    h = fmaf(h, e, 5.259375321e-04f);
    h = fmaf(h, e, 2.417275915e-03f);
    h = fmaf(h, e, 1.500489842e-02f);
    h = fmaf(h, e, 3.001054525e-01f);
    return h;
}
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

constexpr int RS = 28, RP = RS * RS;

// V = 0: scalar ReLU as in gram_fused.cu.  V = 1: packed f32x2, pairs (i0j0,i1j1) and (i0j1,i1j0),
// variance maps pair-interleaved (s_a, s_b, r_a, r_b) per pixel.
template <int V>
__global__ void __launch_bounds__(256, 1) probe_relu(float *out, int iters) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < 12 * RP * 2; i += 256) smem[i] = 1.0f + 1e-3f * (i & 7);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = lane < RS ? lane : RS - 1;
    const int wi = warp >> 2, wj = warp & 3;
    float m[4][RS];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < RS; ++r) m[q][r] = 0.3f + 1e-3f * (lane + r + q);
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {
            const float2 *sb = reinterpret_cast<const float2 *>(smem) + lx;
            const float2 *ai0 = sb + (wi * 2 + 0) * RP, *ai1 = sb + (wi * 2 + 1) * RP;
            const float2 *bj0 = sb + (4 + wj * 2 + 0) * RP, *bj1 = sb + (4 + wj * 2 + 1) * RP;
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                const float2 A0 = ai0[r * RS], A1 = ai1[r * RS], B0 = bj0[r * RS], B1 = bj1[r * RS];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 A = (q & 2) ? A1 : A0, B = (q & 1) ? B1 : B0;
                    const float s = A.x * B.x, rr = A.y * B.y;
                    const float c = m[q][r];
                    const float d = s - fabsf(c);
                    const float e = d * rr;
                    const float w = d * mb_sqrt(fabsf(e));
                    m[q][r] = fmaf(fmaf(w, mb_h2(e), fmaxf(c, 0.f)), 0.45f, 0.1f);
                }
            }
        } else {
            const float4 *sb = reinterpret_cast<const float4 *>(smem) + lx;
            const float4 *ai = sb + wi * RP, *bj = sb + (2 + wj) * RP;
            const u64 C6 = pk(7.577116048e-05f, 7.577116048e-05f), C5 = pk(-5.945927478e-05f, -5.945927478e-05f),
                      C4 = pk(2.400144585e-04f, 2.400144585e-04f), C3 = pk(5.259375321e-04f, 5.259375321e-04f),
                      C2 = pk(2.417275915e-03f, 2.417275915e-03f), C1 = pk(1.500489842e-02f, 1.500489842e-02f),
                      C0 = pk(3.001054525e-01f, 3.001054525e-01f), K0 = pk(0.45f, 0.45f), K1 = pk(0.1f, 0.1f);
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                const float4 A = ai[r * RS], B = bj[r * RS];
                const u64 SA = pk(A.x, A.y), RA = pk(A.z, A.w);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // h = 0: maps (i0j0, i1j1) = m[0], m[3];  h = 1: maps (i0j1, i1j0) = m[1], m[2]
                    const u64 SB = h ? pk(B.y, B.x) : pk(B.x, B.y);
                    const u64 RB = h ? pk(B.w, B.z) : pk(B.z, B.w);
                    float &c0 = h ? m[1][r] : m[0][r];
                    float &c1 = h ? m[2][r] : m[3][r];
                    const u64 NC = pk(-fabsf(c0), -fabsf(c1));
                    const u64 D = fma2(SA, SB, NC);
                    const u64 E = mul2(D, mul2(RA, RB));
                    float e0, e1;
                    upk(E, e0, e1);
                    const u64 Q = pk(mb_sqrt(fabsf(e0)), mb_sqrt(fabsf(e1)));
                    u64 H = fma2(C6, E, C5);
                    H = fma2(H, E, C4);
                    H = fma2(H, E, C3);
                    H = fma2(H, E, C2);
                    H = fma2(H, E, C1);
                    H = fma2(H, E, C0);
                    const u64 W = mul2(D, Q);
                    u64 O = fma2(W, H, pk(fmaxf(c0, 0.f), fmaxf(c1, 0.f)));
                    O = fma2(O, K0, K1);
                    upk(O, c0, c1);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < RS; ++r) s += m[q][r];
    if (s == 123.456f) out[0] = s;
}

}  // namespace

// pixel-pairs per second (4 maps x 28 x 28 lanes-used per warp-iteration) of the ReLU step
extern "C" double mb_relu(int variant, int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *d = nullptr;
    cudaMalloc(&d, 16);
    const size_t smem = 12 * RP * 2 * sizeof(float);
    cudaFuncSetAttribute(probe_relu<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe_relu<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](int n) {
        if (variant == 0) probe_relu<0><<<sms, 256, smem>>>(d, n);
        else probe_relu<1><<<sms, 256, smem>>>(d, n);
    };
    run(iters / 4 + 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
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
    return 4.0 * RP * 8.0 * sms * iters / (best * 1e-3);
}
