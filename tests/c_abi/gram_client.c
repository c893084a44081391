/* A plain-C client of include/cnngp.h: no Python, no torch -- CUDA runtime allocations and the
 * five calls a binding needs.  Evaluates the README model of the reference
 * (README.md:33-46: Conv2d k3, ReLU, Conv2d k3 stride 2, ReLU, Conv2d k14 padding 0) on
 * deterministic pseudo-random images and writes Kxz [N1, N2] and Kxx [N1, N1] as raw float32;
 * tests/test_gpu_gram.py::test_plain_c_client compares them with the Python front end.
 *   usage: gram_client OUT_FILE                                                             */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "cnngp.h"

#define N1 10
#define N2 7
#define C 3
#define S 28

#define CK(call)                                                                             \
    do {                                                                                     \
        if ((call) != 0) {                                                                   \
            fprintf(stderr, "%s failed: %s\n", #call, cnngp_last_error());                  \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)
#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));              \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

static uint32_t lcg_state = 12345u;
static float next_uniform(void) { /* the test regenerates exactly this stream */
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) / 16777216.0f;
}

static cnngp_op conv(int k, int stride, int pad) {
    cnngp_op o = {CNNGP_OP_CONV, 0, 0, k, 0, stride, pad, 1, (double)(float)(1.0 / (k * k)), 0.0};
    return o;
}
static cnngp_op relu(void) {
    cnngp_op o = {CNNGP_OP_RELU, 0, 0, 0, 0, 1, 0, 1, 1.0, 0.0};
    return o;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (cnngp_abi_version() != CNNGP_ABI_VERSION) return 3;
    cnngp_op ops[5];
    ops[0] = conv(3, 1, 1); ops[1] = relu(); ops[2] = conv(3, 2, 1); ops[3] = relu(); ops[4] = conv(14, 1, 0);
    cnngp_plan *plan = NULL;
    CK(cnngp_plan_create(ops, 5, 1, S, S, CNNGP_F32, &plan));
    const int64_t aux = cnngp_plan_aux_elems(plan);

    const size_t px = (size_t)C * S * S;
    float *hx = (float *)malloc(sizeof(float) * N1 * px), *hz = (float *)malloc(sizeof(float) * N2 * px);
    for (size_t i = 0; i < N1 * px; ++i) hx[i] = next_uniform();
    for (size_t i = 0; i < N2 * px; ++i) hz[i] = next_uniform();

    float *dx, *dz, *ax, *az, *kd, *kxz, *kxx;
    CU(cudaMalloc((void **)&dx, sizeof(float) * N1 * px));
    CU(cudaMalloc((void **)&dz, sizeof(float) * N2 * px));
    CU(cudaMalloc((void **)&ax, sizeof(float) * (size_t)(N1 + N1 % 2) * (size_t)aux)); /* an even number of rows */
    CU(cudaMalloc((void **)&az, sizeof(float) * (size_t)(N2 + N2 % 2) * (size_t)aux));
    CU(cudaMalloc((void **)&kd, sizeof(float) * N1));
    CU(cudaMalloc((void **)&kxz, sizeof(float) * N1 * N2));
    CU(cudaMalloc((void **)&kxx, sizeof(float) * N1 * N1));
    CU(cudaMemcpy(dx, hx, sizeof(float) * N1 * px, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dz, hz, sizeof(float) * N2 * px, cudaMemcpyHostToDevice));

    cudaStream_t stream;
    CU(cudaStreamCreate(&stream));
    CK(cnngp_variances(plan, dx, NULL, N1, C, ax, NULL, kd, stream));
    CK(cnngp_variances(plan, dz, NULL, N2, C, az, NULL, NULL, stream));
    /* model(X, Z) */
    CK(cnngp_gram(plan, dx, N1, dz, N2, C, ax, az, NULL, 0, 0, 0, kxz, N2, 0, stream));
    const int path_xz = cnngp_last_path();
    /* model(X): same images on both sides -> j >= i computed and mirrored, diagonal from kd */
    CK(cnngp_gram(plan, dx, N1, dx, N1, C, ax, ax, kd, 1, 0, 1, kxx, N1, 0, stream));
    CU(cudaStreamSynchronize(stream));

    float *out = (float *)malloc(sizeof(float) * (N1 * N2 + N1 * N1));
    CU(cudaMemcpy(out, kxz, sizeof(float) * N1 * N2, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out + N1 * N2, kxx, sizeof(float) * N1 * N1, cudaMemcpyDeviceToHost));
    FILE *f = fopen(argv[1], "wb");
    if (!f || fwrite(out, sizeof(float), N1 * N2 + N1 * N1, f) != (size_t)(N1 * N2 + N1 * N1)) return 4;
    fclose(f);
    printf("ok path=%d kernel_family=%d aux=%lld\n", path_xz, cnngp_plan_has_fused(plan), (long long)aux);
    cnngp_plan_destroy(plan);
    return 0;
}
