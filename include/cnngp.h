/*
 * cnngp.h -- C ABI of the B200-native cnn-gp hot path (libcnngp.so, sm_100a).
 *
 * The reference (waleedbinkhalid74/cnn-gp) is pure Python on PyTorch and has no FFI of its
 * own; its "operator interface" for this path is the Python call protocol
 *     model(x, y=None, same=None, diag=False)                cnn_gp/kernels.py:18-57
 *     module.propagate(kp)                                    cnn_gp/kernels.py:92-98,134-165,
 *                                                             184-187,221-225,252-254
 *     save_K(f, kern, name, X, X2, diag, batch_size, ...)     cnn_gp/kernel_save_tools.py:26-58
 *     solve_system(Kxx, Y) / print_accuracy(A, Kxvx, Y, key)  exp_mnist_resnet/classify_gp.py:17-42
 * The entry points below are what a binding for that path needs; every one cites the
 * reference code it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer owned by the caller
 *     (PyTorch allocates; pass tensor.data_ptr()).  The library allocates nothing that
 *     outlives a call except the small plan object, its per-device copy of the op list and
 *     one 8-byte tile counter per (device, stream) the persistent Gram kernels have run on.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), on the current
 *     device; no hidden synchronisation.
 *   - return value 0 = success, non-zero = error; cnngp_last_error() returns a thread-local
 *     message.  Nothing throws or exits across the ABI.
 *   - dtype: 0 = float32, 1 = float64 (a .double() model, parity 1e-10).
 */
#ifndef CNNGP_H
#define CNNGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNNGP_ABI_VERSION 1

enum { CNNGP_F32 = 0, CNNGP_F64 = 1 };

/* One step of the linearised layer program.  A program is the module tree
 * (Sequential / Sum / Mixture of Conv2d and ReLU, kernels.py:178-254) flattened by the host
 * into three-address ops over numbered map slots; slot 0 holds the initial per-pixel
 * covariance map (kernels.py:43-49) and the last op's dst holds the final 1x1 map. */
enum {
    CNNGP_OP_CONV  = 1, /* dst = box_conv(src)*scale + bias          kernels.py:92-98   */
    CNNGP_OP_RELU  = 2, /* dst = arccos expectation of src           kernels.py:134-165 */
    CNNGP_OP_COPY  = 3, /* dst = src                 (Sum keeps its input alive)        */
    CNNGP_OP_ADD   = 4, /* dst = dst + src           Sum.propagate, kernels.py:252-254  */
    CNNGP_OP_SCALE = 5  /* dst = src * scale         Mixture.propagate, kernels.py:221-225 */
};

typedef struct cnngp_op {
    int32_t opcode;
    int32_t src, dst;   /* slot indices, 0 <= slot < n_slots */
    int32_t ke;         /* CONV: kernel extent = kernel_size (+1 when zero_first)           */
    int32_t zero_first; /* CONV: row 0 / col 0 of the kernel are zero (kernels.py:73-84)    */
    int32_t stride;     /* CONV */
    int32_t pad;        /* CONV: zero padding on every side                                 */
    int32_t dil;        /* CONV: dilation                                                   */
    double scale;       /* CONV: tap = float32(var_weight/kernel_size^2) (kernels.py:87-88);
                           SCALE: mixture proportion                                        */
    double bias;        /* CONV: var_bias                                                   */
} cnngp_op;

typedef struct cnngp_plan cnngp_plan; /* opaque, immutable after creation, thread-shareable */

/* Which implementation cnngp_gram used for the last call on this thread. */
enum { CNNGP_PATH_NONE = 0, CNNGP_PATH_GENERIC = 1, CNNGP_PATH_FUSED = 2, CNNGP_PATH_FUSED_NET = 3 };

int cnngp_abi_version(void);
const char *cnngp_last_error(void);

/* Build a plan for `ops` on H x W input maps.  Fails if any conv output would be empty, a
 * slot is read before it is written, ADD shapes differ, or the final map is not 1x1 (the
 * reference fails there with a view error, kernels.py:54-57). */
int cnngp_plan_create(const cnngp_op *ops, int32_t n_ops, int32_t n_slots, int32_t H, int32_t W,
                      int32_t dtype, cnngp_plan **out);
void cnngp_plan_destroy(cnngp_plan *plan);

/* Elements per image row of the variance buffers cnngp_variances fills: the `xx` / `yy`
 * operand of every ReLU (kernels.py:146), followed -- when the fused kernel covers the program
 * -- by the same maps as (sqrt, 1/sqrt) pairs in the layout that kernel stages. */
int64_t cnngp_plan_aux_elems(const cnngp_plan *plan);
/* Algorithmic flop per image pair under SURVEY.md 8(d)'s counting convention. */
double cnngp_plan_flops_per_pair(const cnngp_plan *plan, int32_t C);
/* Which register-resident fused kernel covers this program: CNNGP_PATH_FUSED (straight-line 28x28
 * programs), CNNGP_PATH_FUSED_NET (Sum / stride / several map sizes, skip maps stashed in tensor
 * memory), or 0 (generic kernel only). */
int cnngp_plan_has_fused(const cnngp_plan *plan);

/* Human-readable description of how the plan will be executed (kernel family and, for the fused
 * kernels, the register-level op list the host translator produced), written NUL-terminated into
 * buf (at most cap bytes).  Returns the number of bytes the full text needs.  For tests and
 * debugging of the host-side translation; no GPU is involved. */
int64_t cnngp_plan_describe(const cnngp_plan *plan, char *buf, int64_t cap);
/* The same translation with every numeric field, one op per line ("KIND key=value ..."), so that a
 * CPU interpreter can re-evaluate what the fused kernels will compute (tests/test_host.py checks the
 * translators that way against the recursive tree walk on random programs).  Same calling convention. */
int64_t cnngp_plan_dump(const cnngp_plan *plan, char *buf, int64_t cap);

/* Per-image variance recursion: the xx / yy maps of kernels.py:48-49 pushed through the
 * program (Conv2d acts on them as on xy, kernels.py:98; ReLU halves them, kernels.py:154,164).
 *   d_x        [N, C, H, W] images
 *   d_z        NULL, or [N, C, H, W] partner images for the literal same=True-with-different-
 *              data semantics (after each ReLU yy := xx, kernels.py:155-156)
 *   d_aux_x    [N rounded up to even, aux_elems] out: xx at the input of every ReLU (+ the fused kernels'
 *              operands, which interleave images 2k and 2k+1 over rows 2k and 2k+1)
 *   d_aux_z    [N, aux_elems] out (only when d_z != NULL): yy at the input of every ReLU
 *   d_kdiag    [N] out, may be NULL: the final 1x1 value of the xx recursion, i.e.
 *              model(x, diag=True) (kernels.py:155-158) */
int cnngp_variances(const cnngp_plan *plan, const void *d_x, const void *d_z, int64_t N, int32_t C,
                    void *d_aux_x, void *d_aux_z, void *d_kdiag, void *stream);

/* The Gram tile: model(x, z, same, diag) of kernels.py:18-57.
 *   d_x [N1,C,H,W], d_z [N2,C,H,W]; d_aux_x / d_aux_z from cnngp_variances.  Sub-blocks of a
 *            larger array may be passed (offset image / row pointers) as long as they start at an
 *            EVEN image index: the fused maps interleave images 2k and 2k+1 over rows 2k, 2k+1,
 *            and the variance buffers must hold an even number of rows
 *   d_kdiag  NULL, or [N1] from cnngp_variances: with `symmetric` the diagonal entries are
 *            copied from it, so that diag(model(X)) == model(X, diag=True) bit for bit
 *   same   != 0: entries with i == j follow the variance recursion (kernels.py:155-162)
 *   diag   != 0: only pairs (n, n) are evaluated, out is [N1]  (requires N1 == N2)
 *   symmetric != 0: caller asserts d_x and d_z hold the same images (model(X)): only
 *                j >= i is computed and mirrored
 *   d_out  [N1, ld_out] (row stride ld_out elements) or [N1] when diag
 *   path   0 = auto (a fused kernel when one covers the program), 1 = force generic,
 *          2 = force fused (fails if neither fused kernel covers the program) */
int cnngp_gram(const cnngp_plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
               int32_t C, const void *d_aux_x, const void *d_aux_z, const void *d_kdiag, int32_t same,
               int32_t diag, int32_t symmetric, void *d_out, int64_t ld_out, int32_t path, void *stream);
int cnngp_last_path(void);
/* kernels the last cnngp_gram / cnngp_gram_symmetric_to_host call of this thread launched (programs with a
 * folded phase run as two launches per chunk of super-tiles, everything else as one) */
int cnngp_last_launches(void);

/* A band of block rows of model(X) in ONE launch: rows [0, N1) x columns [0, N2) of the symmetric Gram of the N2
 * images at d_x (N1 <= N2; d_aux / d_kdiag are the variance rows / diagonal values of the same images, from
 * cnngp_variances).  This is what a worker's slice of the reference's tile list (cnn_gp/data.py:11-29) amounts to for
 * whole block rows: the same=True tile on the diagonal and every tile to its right (kernel_save_tools.py:49-58).
 * Entries with j >= i are computed; inside each diagonal block of `block` rows (0: everywhere, i.e. N1 == N2
 * gives model(X)) they are mirrored, so that out holds exactly what the reference's tiles (i, j >= i) of these block
 * rows hold; entries below the diagonal blocks are left untouched (the reference leaves them NaN).  d_x must start at an
 * even image index of the array cnngp_variances saw.  Fused kernels only (error 4 otherwise). */
int cnngp_gram_band(const cnngp_plan *plan, const void *d_x, int64_t N1, int64_t N2, int32_t C, const void *d_aux,
                    const void *d_kdiag, int64_t block, void *d_out, int64_t ld_out, void *stream);

/* model(X) for a caller that wants the result in HOST memory (what save_kernel's `kern` does with
 * .cpu(), exp_mnist_resnet/save_kernel.py:21-24): the symmetric Gram of d_x is computed into
 * d_out [N, ld_out] as by cnngp_gram(symmetric = 1), and bands of finished rows are copied to the
 * pinned host array h_out [N, ld_host] on `copy_stream` WHILE the launch is still running -- the
 * kernel counts finished tiles per band in d_scratch (device memory, scratch_bytes >= 4 bytes per
 * 48 rows -- bands are 504 rows unless the program's variance maps are large; zeroed here), and the copy stream waits on those counters with stream memory
 * operations.  Both streams must differ; the caller synchronises `copy_stream` before reading
 * h_out.  Returns 4 if the plan's kernel family does not report progress (use cnngp_gram + a
 * copy then). */
int cnngp_gram_symmetric_to_host(const cnngp_plan *plan, const void *d_x, int64_t N, int32_t C, const void *d_aux,
                                 const void *d_kdiag, void *d_out, int64_t ld_out, void *h_out, int64_t ld_host,
                                 void *d_scratch, int64_t scratch_bytes, void *stream, void *copy_stream);

/* Map-level steps behind module.propagate(kp): a stack of M maps [M, Hi, Wi]. */
int cnngp_conv_maps(const void *d_in, int64_t M, int32_t Hi, int32_t Wi, const cnngp_op *conv,
                    int32_t dtype, void *d_out, void *stream);                 /* kernels.py:94-97 */
int cnngp_relu_maps(void *d_xy, const void *d_xx, const void *d_yy, int64_t Nx, int64_t Ny,
                    int64_t P, int32_t same, int32_t diag, int32_t dtype, void *stream); /* :146-162 */

/* exp_mnist_resnet/classify_gp.py:17-27: scipy.linalg.solve(Kxx, Y, assume_a='pos',
 * lower=False) == LAPACK dposv('U').  Column-major is not assumed: A is row-major [n, lda]
 * and only its upper triangle (j >= i) is read, exactly the blocks save_K writes.
 * potrf overwrites the upper triangle with U (A = U^T U); *d_info = 0 on success, k > 0 if
 * the leading minor of order k is not positive definite.  potrs solves U^T U X = B in place,
 * B row-major [n, ldb] with nrhs columns. */
int cnngp_potrf_upper_f64(double *d_A, int64_t n, int64_t lda, int32_t *d_info, void *stream);
/* Building blocks of the same factorisation for drivers that keep block rows of A on several GPUs
 * (cnn-gp_b200/cnn_gp/linalg_dist.py).
 * panel: factorise one block row of up to 256 rows in place.  d_P points at its diagonal element:
 *        rows i < min(256, width), columns i <= j < width, row stride ldp; on return they hold
 *        U[k, k:].  *d_info is NOT cleared; the first non-positive pivot sets it to
 *        info_base + (row within the panel) + 1 if it is still 0.  d_work: 128 * 128 doubles of
 *        caller-owned scratch (the inverse of the current diagonal block).
 * syrk:  C[i][j] -= sum_k X[k][i] X[k][j] for j >= i, i, j < m and i in the 128-row tiles
 *        [ib_lo, ib_hi); X is the factored panel [K <= 256, ldx] restricted to the trailing columns and
 *        C is addressed as d_C + i*ldc + j, so a rank that holds only some rows passes a shifted base. */
int cnngp_potrf_panel_f64(double *d_P, int64_t ldp, int64_t width, int64_t info_base, int32_t *d_info,
                          double *d_work, void *stream);
int cnngp_syrk_upper_f64(const double *d_X, int64_t ldx, int32_t K, double *d_C, int64_t ldc, int64_t m,
                         int32_t ib_lo, int32_t ib_hi, void *stream);
/* syrk for a rank that owns every `stride`-th 256-row block of the trailing matrix, in ONE launch:
 * trailing blocks ti0, ti0 + stride, ... (n_blocks of them; trailing block t = trailing rows
 * [256 t, 256 t + 256)) are stored stacked, 256 rows apart, from local block q0 of d_C_local, whose
 * column 0 is the trailing matrix's column 0. */
int cnngp_syrk_upper_strided_f64(const double *d_X, int64_t ldx, int32_t K, double *d_C_local, int64_t ldc,
                                 int64_t m, int32_t ti0, int32_t stride, int32_t q0, int32_t n_blocks,
                                 void *stream);
int cnngp_potrs_upper_f64(const double *d_U, int64_t n, int64_t lda, double *d_B, int32_t nrhs,
                          int64_t ldb, void *stream);
/* The two triangular sweeps of potrs for drivers that keep block rows of U on several GPUs
 * (linalg_dist.py: no rank ever holds all of U; classify_gp.py:17-27 on a matrix spread over GPUs).
 * A panel is one block row of U as its owner stores it: d_P points at its diagonal element,
 * rows <= 256, `width` columns to the end of the matrix.
 * fwd_panel: d_B[0:rows] (right-hand sides with the updates of all earlier blocks, already summed
 *            over the ranks) becomes y = U_kk^-T (.) in place, and the caller's accumulator rows
 *            d_B[rows:width] take  -= U[panel rows, later columns]^T y.
 * bwd_diag:  d_B[0:rows] (y minus the updates of all later blocks) becomes x = U_kk^-1 (.) in place.
 * rows_update: d_Y[i] -= d_U[i][0:nb] d_X for the nrows stacked local rows above the solved block;
 *            d_U points at that block's columns, d_X [nb, ldx] is its (broadcast) solution. */
int cnngp_trsm_fwd_panel_f64(const double *d_P, int64_t ldp, int64_t rows, int64_t width, double *d_B,
                             int32_t nrhs, int64_t ldb, void *stream);
int cnngp_trsm_bwd_diag_f64(const double *d_P, int64_t ldp, int64_t rows, double *d_B, int32_t nrhs,
                            int64_t ldb, void *stream);
int cnngp_rows_update_f64(const double *d_U, int64_t ldu, int64_t nrows, int32_t nb, const double *d_X,
                          int64_t ldx, double *d_Y, int64_t ldy, int32_t nrhs, void *stream);
/* classify_gp.py:39-41: pred[r] = argmax_c (K[r,:] . A[:,c]); K row-major float32 [R, ldk]
 * (as stored by save_K), A row-major float64 [n, nrhs]; accumulates in float64. */
int cnngp_predict_argmax(const float *d_K, int64_t R, int64_t n, int64_t ldk, const double *d_A,
                         int32_t nrhs, int64_t *d_pred, double *d_scores, void *stream);
/* the same for a kernel block that is genuinely float64 (classify_gp.py:63 widens what it loads; a caller
 * may also hand over float64 kernels it computed itself) */
int cnngp_predict_argmax_f64(const double *d_K, int64_t R, int64_t n, int64_t ldk, const double *d_A,
                             int32_t nrhs, int64_t *d_pred, double *d_scores, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CNNGP_H */
