/* A plain-C client of include/cnngp_h5.h (no Python, no C++): the save_K layout of
 * cnn_gp/kernel_save_tools.py:7-23 written block by block, reopened, read back, merged.
 * Built and run by tests/test_h5store.py::test_plain_c_client. Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cnngp_h5.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        if ((call) != 0) {                                                               \
            fprintf(stderr, "%s failed: %s\n", #call, cnngp_h5_last_error());           \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

static int make_worker_file(const char *path, int rank, int n_workers, int N, int bs) {
    cnngp_h5 *f = NULL;
    int ds = -1;
    const float nan_fill = NAN;
    int64_t shape[3] = {1, N, N}, maxshape[3] = {CNNGP_H5_UNLIMITED, N, N}, chunks[3] = {1, bs, bs};
    float *block = (float *)malloc(sizeof(float) * (size_t)bs * (size_t)bs);
    CHECK(cnngp_h5_open(path, "w", &f));
    CHECK(cnngp_h5_create_dataset(f, "Kxx", 3, shape, maxshape, chunks, 0, &nan_fill, &ds));
    /* upper block triangle, tiles dealt out round-robin to the workers */
    int tile = 0;
    for (int i = 0; i < N; i += bs)
        for (int j = i; j < N; j += bs, ++tile) {
            if (tile % n_workers != rank) continue;
            const int n = i + bs <= N ? bs : N - i, m = j + bs <= N ? bs : N - j;
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < m; ++b) block[a * m + b] = (float)(1000 * (i + a) + (j + b));
            int64_t start[3] = {0, i, j}, count[3] = {1, n, m};
            CHECK(cnngp_h5_write(f, ds, start, count, block));
        }
    free(block);
    CHECK(cnngp_h5_close(f));
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const int N = 23, bs = 5, W = 3;
    char path[3][512];
    for (int r = 0; r < W; ++r) {
        snprintf(path[r], sizeof path[r], "%s/w%d.h5", argv[1], r);
        if (make_worker_file(path[r], r, W, N, bs)) return 1;
    }
    cnngp_h5 *dest = NULL;
    CHECK(cnngp_h5_open(path[0], "a", &dest));
    for (int r = 1; r < W; ++r) {
        cnngp_h5 *src = NULL;
        CHECK(cnngp_h5_open(path[r], "r", &src));
        CHECK(cnngp_h5_merge_nan(dest, cnngp_h5_find(dest, "Kxx"), src, cnngp_h5_find(src, "Kxx")));
        CHECK(cnngp_h5_close(src));
    }
    CHECK(cnngp_h5_close(dest));

    cnngp_h5 *f = NULL;
    cnngp_h5_info info;
    CHECK(cnngp_h5_open(path[0], "r", &f));
    if (cnngp_h5_count(f) != 1 || cnngp_h5_find(f, "Kxx") != 0 || cnngp_h5_find(f, "nope") != -1) return 3;
    CHECK(cnngp_h5_dataset_info(f, 0, &info));
    if (info.rank != 3 || info.dtype != 0 || !info.chunked || !info.has_fill || !isnan(info.fill) ||
        info.shape[1] != N || info.maxshape[0] != CNNGP_H5_UNLIMITED || info.chunks[2] != bs ||
        info.n_chunks_stored != 15)
        return 4;
    float *K = (float *)malloc(sizeof(float) * N * N);
    int64_t start[3] = {0, 0, 0}, count[3] = {1, N, N};
    CHECK(cnngp_h5_read(f, 0, start, count, K));
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const float v = K[i * N + j];
            if (j / bs >= i / bs ? v != (float)(1000 * i + j) : !isnan(v)) {
                fprintf(stderr, "entry (%d, %d) = %g\n", i, j, v);
                return 5;
            }
        }
    free(K);
    int64_t bad[3] = {0, 0, N - 1}, two[3] = {1, 1, 2};
    float tmp[2];
    if (cnngp_h5_read(f, 0, bad, two, tmp) == 0 || strlen(cnngp_h5_last_error()) == 0) return 6; /* out of range */
    CHECK(cnngp_h5_close(f));
    puts("ok");
    return 0;
}
