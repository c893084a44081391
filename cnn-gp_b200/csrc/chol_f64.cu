// chol_f64.cu -- blocked fp64 Cholesky (upper, row-major), triangular solves and the
// prediction step of the GP classifier, for sm_100a.
//
// Reference (paths relative to /root/reference):
//   exp_mnist_resnet/classify_gp.py:17-27   scipy.linalg.solve(Kxx, Y, assume_a='pos', lower=False)
//                                           == LAPACK dposv('U'): A = U^T U, then two triangular solves
//   exp_mnist_resnet/classify_gp.py:39-41   (Kxvx @ A).argmax(dim=1)
//
// Layout.  A is row-major [n, lda]; only j >= i is ever read or written (save_K never writes the
// strictly lower block triangle of Kxx, it stays NaN: cnn_gp/kernel_save_tools.py:21-23,
// cnn_gp/data.py:22-29).  U overwrites the upper triangle.
//
// potrf, right-looking, two levels.  A block row of 2 NB = 256 rows is factorised as a panel:
//   potf2_inv     one CTA: U_kk = chol(A_kk) in shared memory, then W = U_kk^{-1} (workspace)
//   tn_kernel<1>  X = W^T A[k, k+1:]                       (the triangular solve as a DMMA GEMM)
//   tn_kernel<0>  rank-128 update of the panel's second half, then potf2_inv / tn_kernel<1> again
// and the trailing matrix takes one rank-256 update per panel,
//   tn_kernel<0>  A22 -= X^T X, j >= i    (SYRK, the one dense contraction: mma.sync.m8n8k4.f64 =
//                 SASS DMMA.8x8x4, the FP64 tensor pipe of sm_100a),
// which halves the read-modify-write traffic of A22 against rank-128 updates.  Look-ahead: the
// update's first 256 rows run first; the next panel is then factorised on a high-priority side
// stream underneath the rest of the update (it touches only those 256 rows).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>

#include "plan.h"

namespace cnngp {
namespace {

constexpr int NB = 128;        // block column width
constexpr int TI = 128, TJ = 64;  // CTA tile of the DMMA kernel (rows i x columns j)
constexpr int KC = 16;         // k-chunk per pipeline stage
constexpr int STAGES = 3;
constexpr int PP = TI + 4, PQ = TJ + 4;  // smem pitches (doubles), == 4 mod 16: a half-warp's fragment load
                                         // (k = lane%4, row = lane/4) covers 16 distinct 8-byte banks
constexpr int TN_THREADS = 256;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async8(void *dst, const void *src, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src, int valid_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

constexpr int kMaxRowTiles = 512;

struct TnParams {
    const double *P;  // [K, ldp]: operand giving the rows of C (C row i <- column i of P)
    const double *Q;  // [K, ldq]: operand giving the columns of C
    double *C;        // [M, ldc]
    long long ldp, ldq, ldc;
    int M, N, K;
    int Tj;            // column tiles
    int ib_lo;         // first row tile of this launch
    long long t_off;   // linear index of tile (ib_lo, first column) in the triangular enumeration
    int vec_ok;        // every row of P, Q, C is 16-byte aligned at even columns
    // strided row-tile list (a rank that owns every rt_stride-th 256-row block of the trailing matrix,
    // stored stacked from local block rt_q0): row tile k of the launch is trailing row tile
    // 2 (rt_ti0 + rt_stride (k / 2)) + k % 2 and lives at local row 256 (rt_q0 + k / 2) + 128 (k % 2)
    int rt_n;          // 0 = contiguous mode (ib_lo / t_off above)
    int rt_ti0, rt_stride, rt_q0;
    int rt_prefix[kMaxRowTiles + 1];  // tiles before row tile k
};

// C (op)= P^T Q on TI x TJ tiles.  MODE 0: C -= P^T Q restricted to j >= i (SYRK; tiles are
// enumerated over the upper block triangle only).  MODE 1: C = P^T Q, a single row tile
// (M <= TI); Q and C may alias because a CTA owns whole columns.
template <int MODE>
__global__ void __launch_bounds__(TN_THREADS, 2) tn_kernel(const TnParams g) {
    extern __shared__ __align__(16) double tn_smem[];
    double *Ps = tn_smem;                       // [STAGES][KC][PP]
    double *Qs = tn_smem + STAGES * KC * PP;    // [STAGES][KC][PQ]

    int ib, jb;
    long long crow0 = -1;  // row of C holding the tile's first row (-1: the trailing row index itself)
    if (MODE == 0 && g.rt_n > 0) {
        const int t = (int)blockIdx.x;
        int lo = 0, hi = g.rt_n - 1;  // last k with rt_prefix[k] <= t
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (g.rt_prefix[mid] <= t) lo = mid; else hi = mid - 1;
        }
        ib = 2 * (g.rt_ti0 + g.rt_stride * (lo >> 1)) + (lo & 1);
        jb = 2 * ib + (t - g.rt_prefix[lo]);
        crow0 = 256LL * (g.rt_q0 + (lo >> 1)) + 128 * (lo & 1);
    } else if (MODE == 0) {
        // row tile ib holds column tiles jb >= 2 ib: prefix(ib) = ib*Tj - ib*(ib-1)
        const long long t = g.t_off + blockIdx.x;
        const double b = (double)g.Tj + 1.0;
        long long r = (long long)floor((b - sqrt(fmax(b * b - 4.0 * (double)t, 0.0))) * 0.5);
        if (r < 0) r = 0;
        while (r > 0 && r * g.Tj - r * (r - 1) > t) --r;
        while ((r + 1) * g.Tj - (r + 1) * r <= t) ++r;
        ib = (int)r;
        jb = 2 * ib + (int)(t - (r * g.Tj - r * (r - 1)));
    } else {
        ib = 0;
        jb = blockIdx.x;
    }
    const int i0 = ib * TI, j0 = jb * TJ;
    if (crow0 < 0) crow0 = i0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wi = warp >> 1, wj = warp & 1;  // 4 x 2 warps of 32 x 32
    const int nchunks = (g.K + KC - 1) / KC;

    auto load_chunk = [&](int chunk, int stage) {
        const int k0 = chunk * KC;
        double *ps = Ps + stage * KC * PP, *qs = Qs + stage * KC * PQ;
        if (g.vec_ok) {
            // P: KC rows x TI/2 16-byte pieces; Q: KC rows x TJ/2
            for (int e = tid; e < KC * (TI / 2); e += TN_THREADS) {
                const int t = e / (TI / 2), c = (e % (TI / 2)) * 2;
                const int col = i0 + c;
                int bytes = (k0 + t < g.K) ? 8 * max(0, min(2, g.M - col)) : 0;
                const double *src = g.P + (long long)(bytes ? k0 + t : 0) * g.ldp + (bytes ? col : 0);
                cp_async16(ps + t * PP + c, src, bytes);
            }
            for (int e = tid; e < KC * (TJ / 2); e += TN_THREADS) {
                const int t = e / (TJ / 2), c = (e % (TJ / 2)) * 2;
                const int col = j0 + c;
                int bytes = (k0 + t < g.K) ? 8 * max(0, min(2, g.N - col)) : 0;
                const double *src = g.Q + (long long)(bytes ? k0 + t : 0) * g.ldq + (bytes ? col : 0);
                cp_async16(qs + t * PQ + c, src, bytes);
            }
        } else {
            for (int e = tid; e < KC * TI; e += TN_THREADS) {
                const int t = e / TI, c = e % TI;
                const bool ok = (k0 + t < g.K) && (i0 + c < g.M);
                cp_async8(ps + t * PP + c, g.P + (ok ? (long long)(k0 + t) * g.ldp + i0 + c : 0), ok);
            }
            for (int e = tid; e < KC * TJ; e += TN_THREADS) {
                const int t = e / TJ, c = e % TJ;
                const bool ok = (k0 + t < g.K) && (j0 + c < g.N);
                cp_async8(qs + t * PQ + c, g.Q + (ok ? (long long)(k0 + t) * g.ldq + j0 + c : 0), ok);
            }
        }
    };

    if (MODE == 0) {
        // the epilogue reads the C tile (128 rows x 512 bytes) after the last chunk: pull it into L2
        // now so that those loads do not pay HBM latency with the tensor pipe idle
        const int row = i0 + (tid >> 1);
        if (row < g.M) {
            const double *c = g.C + (crow0 + (tid >> 1)) * g.ldc + j0 + (tid & 1) * 32;
            if (j0 + (tid & 1) * 32 < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(c));
            if (j0 + (tid & 1) * 32 + 16 < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + 16));
        }
    }
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) load_chunk(s, s);
        cp_commit();
    }
    const int lr = lane >> 2, lk = lane & 3;
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_wait<STAGES - 2>();
        __syncthreads();
        if (ch + STAGES - 1 < nchunks) load_chunk(ch + STAGES - 1, (ch + STAGES - 1) % STAGES);
        cp_commit();
        const double *ps = Ps + (ch % STAGES) * KC * PP + wi * 32 + lr;
        const double *qs = Qs + (ch % STAGES) * KC * PQ + wj * 32 + lr;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                a[f] = ps[(k4 * 4 + lk) * PP + f * 8];
                b[f] = qs[(k4 * 4 + lk) * PQ + f * 8];
            }
#pragma unroll
            for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                for (int fj = 0; fj < 4; ++fj) dmma(acc[fi][fj][0], acc[fi][fj][1], a[fi], b[fj]);
        }
    }
    cp_wait<0>();

    // epilogue: fragment (fi, fj) holds C[i][j], C[i][j+1] with i = i0 + wi*32 + fi*8 + lane/4,
    // j = j0 + wj*32 + fj*8 + 2*(lane%4)
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
        const int i = i0 + wi * 32 + fi * 8 + lr;
        if (i >= g.M) continue;
        double *crow = g.C + (crow0 + (i - i0)) * g.ldc;
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) {
            const int j = j0 + wj * 32 + fj * 8 + 2 * lk;
            if (MODE == 0) {
                if (j + 1 < g.N && j >= i && g.vec_ok) {
                    double2 c = *reinterpret_cast<double2 *>(crow + j);
                    c.x -= acc[fi][fj][0];
                    c.y -= acc[fi][fj][1];
                    *reinterpret_cast<double2 *>(crow + j) = c;
                } else {
                    if (j < g.N && j >= i) crow[j] -= acc[fi][fj][0];
                    if (j + 1 < g.N && j + 1 >= i) crow[j + 1] -= acc[fi][fj][1];
                }
            } else {
                if (j + 1 < g.N && g.vec_ok) {
                    *reinterpret_cast<double2 *>(crow + j) = make_double2(acc[fi][fj][0], acc[fi][fj][1]);
                } else {
                    if (j < g.N) crow[j] = acc[fi][fj][0];
                    if (j + 1 < g.N) crow[j + 1] = acc[fi][fj][1];
                }
            }
        }
    }
}

constexpr size_t kTnSmem = (size_t)STAGES * KC * (PP + PQ) * sizeof(double);

// ---- diagonal block: unblocked Cholesky + triangular inverse in shared memory ---------------
constexpr int DP = NB + 1;  // pitch
constexpr int POTF2_THREADS = 512;

// D = A[kb:kb+nb, kb:kb+nb] (upper).  Writes U_kk back over D and W = U_kk^{-1} (upper, zero
// elsewhere, full NB x NB) to the workspace.  *info = kb + c + 1 for the first non-positive
// pivot (LAPACK dpotrf convention).
//
// Blocked in shared memory with 32-wide panels so that the 128 dependent pivot steps need no
// block-wide barrier: the 32 x 32 diagonal blocks are factorised / inverted by single warps
// (__syncwarp only), the row-panel solves keep a column in registers, and the block updates
// are small products spread over all 512 threads.
constexpr int QB = 32;

// W = U^{-1} in place for the upper-triangular NB x NB block s[NB][DP] (identity-padded), rinv = reciprocals of its
// diagonal, tmp = [QB][QB + 1] scratch; all POTF2_THREADS threads of the CTA call it, a barrier has ordered s and rinv.
__device__ __forceinline__ void tri_inverse_inplace(double *s, const double *rinv, double *tmp, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < NB / QB) {  // diagonal blocks: lane j solves U w = e_j for column j of the inverse, in registers
        const int o = warp * QB;
        double w[QB];
#pragma unroll
        for (int i = QB - 1; i >= 0; --i) {
            double acc = 0.0;
#pragma unroll
            for (int t = i + 1; t < QB; ++t) acc += s[(o + i) * DP + o + t] * (t <= lane ? w[t] : 0.0);
            w[i] = i == lane ? rinv[o + i] : (i < lane ? -acc * rinv[o + i] : 0.0);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < QB; ++i)
            if (lane >= i) s[(o + i) * DP + o + lane] = w[i];
    }
    __syncthreads();
    // off-diagonal blocks: W_ij = -W_ii * sum_{k=i+1..j} U_ik W_kj, block columns right to left (the
    // U_ik of the columns left of j are still intact), block rows bottom to top
    for (int bj = NB / QB - 1; bj >= 1; --bj) {
        for (int bi = bj - 1; bi >= 0; --bi) {
            double t2[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = tid + q * POTF2_THREADS, r = e / QB, c = e % QB;
                double acc = 0.0;
                for (int k = (bi + 1) * QB; k < (bj + 1) * QB; ++k) acc += s[(bi * QB + r) * DP + k] * s[k * DP + bj * QB + c];
                t2[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = tid + q * POTF2_THREADS;
                tmp[(e / QB) * (QB + 1) + e % QB] = t2[q];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = tid + q * POTF2_THREADS, r = e / QB, c = e % QB;
                double acc = 0.0;
                for (int t = r; t < QB; ++t) acc += s[(bi * QB + r) * DP + bi * QB + t] * tmp[t * (QB + 1) + c];
                t2[q] = -acc;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = tid + q * POTF2_THREADS;
                s[(bi * QB + e / QB) * DP + bj * QB + e % QB] = t2[q];
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(POTF2_THREADS, 1) potf2_inv_kernel(double *A, long long lda, int kb, int nb,
                                                                     double *W, int *info, long long info_base) {
    extern __shared__ __align__(16) double ps_smem[];
    double *s = ps_smem;                 // [NB][DP]
    double *rinv = ps_smem + NB * DP;    // [NB] reciprocals of the diagonal of U
    double *tmp = rinv + NB;             // [QB][QB + 1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *D = A + (long long)kb * lda + kb;
#pragma unroll 8
    for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
        const int i = e / NB, j = e % NB;
        s[i * DP + j] = (i < nb && j < nb && j >= i) ? D[(long long)i * lda + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();

    // ---- U = chol(D), right-looking over 32-wide panels -------------------------------------
    for (int o = 0; o < NB; o += QB) {
        if (warp == 0) {  // diagonal block in registers: lane j holds column j, pivots travel by shuffle
            double a[QB];
#pragma unroll
            for (int i = 0; i < QB; ++i) a[i] = s[(o + i) * DP + o + lane];
#pragma unroll
            for (int c = 0; c < QB; ++c) {
                const double d = __shfl_sync(0xffffffffu, a[c], c);
                const bool bad = !(d > 0.0);
                if (bad && lane == 0 && *info == 0) *info = (int)(info_base + kb + o + c + 1);
                const double r = bad ? nan("") : sqrt(d), ri = 1.0 / r;
                const double u = lane == c ? r : a[c] * ri;  // row c of U at this lane's column (valid for lane >= c)
                a[c] = u;
                if (lane == c) rinv[o + c] = ri;
#pragma unroll
                for (int i = c + 1; i < QB; ++i) {
                    const double ui = __shfl_sync(0xffffffffu, u, i);
                    a[i] -= ui * u;  // only entries with lane >= i are ever used
                }
            }
#pragma unroll
            for (int i = 0; i < QB; ++i)
                if (lane >= i) s[(o + i) * DP + o + lane] = a[i];
        }
        __syncthreads();
        const int ncol = NB - o - QB;  // columns right of the panel
        if (tid < ncol) {  // row panel: X = U_oo^{-T} A[o:o+32, j], one column per thread, in registers
            const int j = o + QB + tid;
            double x[QB];
#pragma unroll
            for (int c = 0; c < QB; ++c) x[c] = s[(o + c) * DP + j];
#pragma unroll
            for (int c = 0; c < QB; ++c) {
                x[c] *= rinv[o + c];
#pragma unroll
                for (int c2 = c + 1; c2 < QB; ++c2) x[c2] -= s[(o + c) * DP + o + c2] * x[c];
            }
#pragma unroll
            for (int c = 0; c < QB; ++c) s[(o + c) * DP + j] = x[c];
        }
        __syncthreads();
        for (int e = tid; e < ncol * ncol; e += POTF2_THREADS) {  // trailing block -= X^T X (j >= i)
            const int ii = e / ncol, jj = e % ncol;
            if (jj < ii) continue;
            const int i = o + QB + ii, j = o + QB + jj;
            double acc = 0.0;
#pragma unroll 8
            for (int t = 0; t < QB; ++t) acc += s[(o + t) * DP + i] * s[(o + t) * DP + j];
            s[i * DP + j] -= acc;
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += POTF2_THREADS) {
        const int i = e / nb, j = e % nb;
        if (j >= i) D[(long long)i * lda + j] = s[i * DP + j];
    }
    __syncthreads();

    // ---- W = U^{-1} in place --------------------------------------------------------------------
    tri_inverse_inplace(s, rinv, tmp, tid);
    for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
        const int i = e / NB, j = e % NB;
        W[e] = (i < nb && j < nb && j >= i) ? s[i * DP + j] : 0.0;
    }
}

constexpr size_t kPotf2Smem = (size_t)(NB * DP + NB + QB * (QB + 1)) * sizeof(double);

// ---- triangular solves with a few right-hand sides ------------------------------------------
constexpr int NR = 16;  // right-hand sides per pass

// One CTA: solve with the diagonal block.  FWD: U_kk^T y = b (forward);  else U_kk x = y (backward).
template <bool FWD>
__global__ void __launch_bounds__(512, 1) trsv_block_kernel(const double *U, long long lda, int kb, int nb,
                                                           double *B, long long ldb, int c0, int nr) {
    extern __shared__ __align__(16) double tr_smem[];
    double *s = tr_smem;             // [NB][DP]
    double *bs = tr_smem + NB * DP;  // [NR][DP]: right-hand side c of row r at bs[c * DP + r] (lanes = rows: conflict-free)
    const int tid = threadIdx.x;
    const double *D = U + (long long)kb * lda + kb;
#pragma unroll 8
    for (int e = tid; e < nb * nb; e += 512) {
        const int i = e / nb, j = e % nb;
        if (j >= i) s[i * DP + j] = D[(long long)i * lda + j];
    }
    for (int e = tid; e < nb * nr; e += 512) {
        const int r = e / nr, c = e % nr;
        bs[c * DP + r] = B[(long long)(kb + r) * ldb + c0 + c];
    }
    // reciprocals of the diagonal once (LAPACK's dtrsm scales by the reciprocal too), so that the
    // 128 dependent steps below carry a multiply instead of a division
    double *invd = bs + NR * DP;
    __syncthreads();
    if (tid < nb) invd[tid] = 1.0 / s[tid * DP + tid];
    const int r = tid & (NB - 1), cg = tid >> 7;  // 4 column groups
    if (FWD) {
        for (int t = 0; t < nb; ++t) {
            __syncthreads();
            if (r > t && r < nb) {
                const double f = s[t * DP + r] * invd[t];
                for (int c = cg; c < nr; c += 4) bs[c * DP + r] -= f * bs[c * DP + t];
            }
        }
    } else {
        for (int t = nb - 1; t >= 0; --t) {
            __syncthreads();
            if (r < t) {
                const double f = s[r * DP + t] * invd[t];
                for (int c = cg; c < nr; c += 4) bs[c * DP + r] -= f * bs[c * DP + t];
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < nb * nr; e += 512) {
        const int rr = e / nr, c = e % nr;
        B[(long long)(kb + rr) * ldb + c0 + c] = bs[c * DP + rr] * invd[rr];
    }
}

constexpr size_t kTrsvSmem = (size_t)(NB * DP + NR * DP + NB) * sizeof(double);

// forward update: B[j] -= sum_t U[kb+t][j] * Y[t]  for j >= kb + nb   (thread per row j of B)
__global__ void __launch_bounds__(256) fwd_update_kernel(const double *U, long long lda, int kb, int nb, int n,
                                                         double *B, long long ldb, int c0, int nr) {
    __shared__ double ys[NB * NR];
    for (int e = threadIdx.x; e < nb * NR; e += 256) {
        const int t = e / NR, c = e % NR;
        ys[e] = c < nr ? B[(long long)(kb + t) * ldb + c0 + c] : 0.0;
    }
    __syncthreads();
    const long long j = (long long)kb + nb + (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    double acc[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = 0.0;
    const double *u = U + (long long)kb * lda + j;
#pragma unroll 8
    for (int t = 0; t < nb; ++t) {
        const double v = u[(long long)t * lda];
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[c] += v * ys[t * NR + c];
    }
    double *b = B + j * ldb + c0;
#pragma unroll
    for (int c = 0; c < NR; ++c)
        if (c < nr) b[c] -= acc[c];
}

// backward update: Y[i] -= sum_t U[i][kb+t] * X[t]  for i < kb   (64 rows per CTA, staged)
__global__ void __launch_bounds__(256) bwd_update_kernel(const double *U, long long lda, int kb, int nb, double *B,
                                                         long long ldb, int c0, int nr) {
    extern __shared__ __align__(16) double bw_smem[];
    double *us = bw_smem;             // [64][DP]
    double *xs = bw_smem + 64 * DP;   // [NB][NR]
    const int i0 = blockIdx.x * 64;
    for (int e = threadIdx.x; e < 64 * nb; e += 256) {
        const int i = e / nb, t = e % nb;
        us[i * DP + t] = (i0 + i < kb) ? U[(long long)(i0 + i) * lda + kb + t] : 0.0;
    }
    for (int e = threadIdx.x; e < nb * NR; e += 256) {
        const int t = e / NR, c = e % NR;
        xs[e] = c < nr ? B[(long long)(kb + t) * ldb + c0 + c] : 0.0;
    }
    __syncthreads();
    const int i = threadIdx.x & 63, cg = threadIdx.x >> 6;  // columns cg, cg+4, cg+8, cg+12
    if (i0 + i >= kb) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < nb; ++t) {
        const double v = us[i * DP + t];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += v * xs[t * NR + cg + 4 * q];
    }
    double *b = B + (long long)(i0 + i) * ldb + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (cg + 4 * q < nr) b[cg + 4 * q] -= acc[q];
}

constexpr size_t kBwdSmem = (size_t)(64 * DP + NB * NR) * sizeof(double);

// ---- the two triangular sweeps as persistent dataflow kernels -------------------------------------
// U^T y = b and U x = y with a few right-hand sides are chains of n / NB dependent block steps.  One
// launch per step (the first version: a one-CTA substitution + an update launch per block, 1 024
// launches at n = 32 768) is latency-bound: 39 ms for 8.6 GB of traffic.  Here ONE launch runs the
// whole sweep: CTA c owns the right-hand-side blocks c, c + G, ... (G co-resident CTAs).  For its
// block j it applies the updates of all earlier blocks k as their solutions are published
// (flag[k], release / acquire through global memory), then solves with the diagonal block and
// publishes flag[j].  Only "last update + substitution" of each block is on the critical path; all
// other updates run ahead of it.  Each U block is read once, coalesced, staged through shared memory.
constexpr int SW_THREADS = 512;
constexpr int SW_NR = 16;   // right-hand sides per pass
// pitches of the update stage: a staged U block and the published solution are read as mma.m8n8k4 fragments
// (lane = 4 x fragment row + k): a pitch of 4 mod 16 doubles puts the sixteen lanes of a half-warp on sixteen bank pairs
constexpr int SW_UP = NB + 4;     // staged U block (the diagonal block reuses the buffer at pitch DP, see below)
constexpr int SW_YP = SW_NR + 4;  // published solution of block k
constexpr size_t kSweepSmem = (size_t)(NB * SW_UP + SW_NR * DP + NB * SW_YP + NB) * sizeof(double);

__device__ __forceinline__ void flag_wait(const int *f) {
    if (threadIdx.x == 0) {
        int v;
        do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory"); } while (v == 0);
    }
    __syncthreads();
}
__device__ __forceinline__ void flag_set(int *f) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
    }
}

// FWD: U^T y = b (blocks ascending, update with U[k-block rows, j-block columns]^T);
// else U x = y (blocks descending, update with U[j-block rows, k-block columns]).
// Thread (r, cg): row r of the block, right-hand sides 4 cg .. 4 cg + 3.
template <bool FWD>
__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(const double *__restrict__ U, long long lda, int n, double *B,
                                                             long long ldb, int c0, int nr, int *flags) {
    extern __shared__ __align__(16) double sw_smem[];
    double *d = sw_smem;                    // [NB][SW_UP]: a U block (updates: staged ahead of the flag); then [NB][DP]: the diagonal block
    double *bs = d + NB * SW_UP;            // [SW_NR][DP]: this block's right-hand sides, bs[c * DP + r]
    double *ys = bs + SW_NR * DP;           // [NB][SW_YP]: the published solution of block k
    double *invd = ys + NB * SW_YP;         // [NB]
    const int tid = threadIdx.x;
    const int nblk = (n + NB - 1) / NB;
    // the updates run on the FP64 tensor pipe (mma.m8n8k4: 4.4 x the DFMA rate of one SM, and the last update
    // of a block is on the critical path of the whole sweep): warp w owns rows 8 w .. 8 w + 7 of the block and
    // all sixteen right-hand sides = two 8 x 8 accumulator fragments, element (row lr, columns 8 nt + 2 lk + {0, 1})
    const int wq = tid >> 5, lr = (tid & 31) >> 2, lk = tid & 3;
    for (int q = blockIdx.x; q < nblk; q += gridDim.x) {
        const int j = FWD ? q : nblk - 1 - q;
        const int jb = j * NB, nbj = n - jb < NB ? n - jb : NB;
        double acc[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int rr = 8 * wq + lr, c = nt * 8 + 2 * lk + e;
                acc[nt][e] = (rr < nbj && c < nr) ? B[(long long)(jb + rr) * ldb + c0 + c] : 0.0;
            }
        // W = inverse of the diagonal block, computed HERE -- before the updates, i.e. while the blocks this one
        // depends on are still being solved (only the first blocks of a sweep have no such slack) -- and kept in
        // registers (32 doubles per thread) under the update loop: the solve with the diagonal block is then one
        // more tensor-pipe product on the critical path instead of 128 dependent substitution steps
        double dg[NB * NB / SW_THREADS];
#pragma unroll 8
        for (int e = tid; e < NB * NB; e += SW_THREADS) {
            const int i = e >> 7, t = e & (NB - 1);
            d[i * DP + t] = (i < nbj && t < nbj && t >= i) ? __ldg(U + (long long)(jb + i) * lda + jb + t) : (i == t ? 1.0 : 0.0);
        }
        __syncthreads();
        if (tid < NB) invd[tid] = 1.0 / d[tid * DP + tid];
        __syncthreads();
        tri_inverse_inplace(d, invd, bs, tid);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NB * NB / SW_THREADS; ++u) {
            const int e = tid + u * SW_THREADS, i = e >> 7, t = e & (NB - 1);
            dg[u] = t >= i ? d[i * DP + t] : 0.0;
        }
        __syncthreads();
        const int k_lo = FWD ? 0 : j + 1, k_hi = FWD ? j : nblk;
        for (int kk = k_lo; kk < k_hi; ++kk) {
            const int k = FWD ? kk : nblk - 1 - (kk - k_lo);  // backward: the last block first
            const int kb = k * NB, nbk = n - kb < NB ? n - kb : NB;
            // stage the U block BEFORE looking at the flag (it does not depend on it): all 32 loads of a
            // thread are in flight together, and the DRAM latency hides under the wait for block k.
            // FWD: rows kb.., columns jb..; else rows jb.., columns kb.. (both read along their rows: coalesced)
            {
                const long long row0 = FWD ? kb : jb, col0 = FWD ? jb : kb;
                const int rows = FWD ? nbk : nbj, cols = FWD ? nbj : nbk;
#pragma unroll 8
                for (int e = tid; e < NB * NB; e += SW_THREADS) {
                    const int i = e >> 7, t = e & (NB - 1);
                    d[i * SW_UP + t] = (i < rows && t < cols) ? __ldg(U + (row0 + i) * lda + col0 + t) : 0.0;
                }
            }
            flag_wait(flags + k);
            for (int e = tid; e < NB * SW_NR; e += SW_THREADS) {
                const int t = e / SW_NR, c = e % SW_NR;
                // published by another SM: read through L2
                ys[t * SW_YP + c] = (t < nbk && c < nr) ? __ldcg(B + (long long)(kb + t) * ldb + c0 + c) : 0.0;
            }
            __syncthreads();
            {   // FWD: acc[r] -= sum_t U[kb + t][jb + r] y[t]; else acc[r] -= sum_t U[jb + r][kb + t] x[t]
                const double *pa = FWD ? d + lk * SW_UP + 8 * wq + lr : d + (8 * wq + lr) * SW_UP + lk;
                const double *pb = ys + lk * SW_YP + lr;
#pragma unroll 8
                for (int k4 = 0; k4 < NB / 4; ++k4) {
                    const double a = -(FWD ? pa[k4 * 4 * SW_UP] : pa[k4 * 4]);
                    const double b0 = pb[k4 * 4 * SW_YP], b1 = pb[k4 * 4 * SW_YP + 8];
                    dmma(acc[0][0], acc[0][1], a, b0);
                    dmma(acc[1][0], acc[1][1], a, b1);
                }
            }
            __syncthreads();
        }
        // y_j = W^T b (FWD) / x_j = W y (else): W into the fragment-friendly buffer, the updated right-hand sides
        // into the B-operand buffer, one product over the non-zero half of the k range
#pragma unroll
        for (int u = 0; u < NB * NB / SW_THREADS; ++u) {
            const int e = tid + u * SW_THREADS;
            d[(e >> 7) * SW_UP + (e & (NB - 1))] = dg[u];
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) ys[(8 * wq + lr) * SW_YP + nt * 8 + 2 * lk + e] = acc[nt][e];
        __syncthreads();
        {
            double out[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const double *pa = FWD ? d + lk * SW_UP + 8 * wq + lr : d + (8 * wq + lr) * SW_UP + lk;
            const double *pb = ys + lk * SW_YP + lr;
            // W is upper triangular: FWD uses W[t][r], t <= r; else W[r][t], t >= r (rows 8 wq .. 8 wq + 7 of this warp)
            const int k4_lo = FWD ? 0 : 2 * wq, k4_hi = FWD ? 2 * wq + 2 : NB / 4;
#pragma unroll 4
            for (int k4 = k4_lo; k4 < k4_hi; ++k4) {
                const double a = FWD ? pa[k4 * 4 * SW_UP] : pa[k4 * 4];
                const double b0 = pb[k4 * 4 * SW_YP], b1 = pb[k4 * 4 * SW_YP + 8];
                dmma(out[0][0], out[0][1], a, b0);
                dmma(out[1][0], out[1][1], a, b1);
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int rr = 8 * wq + lr, c = nt * 8 + 2 * lk + e;
                    if (rr < nbj && c < nr) B[(long long)(jb + rr) * ldb + c0 + c] = out[nt][e];
                }
        }
        flag_set(flags + j);
        __syncthreads();
    }
}

// Y[i][:] -= sum_t U[i][t] X[t][:] for nrows stacked rows of a rank's block rows (distributed backward
// sweep: X is the freshly broadcast solution block, U points at its columns): 64 rows per CTA, staged
__global__ void __launch_bounds__(256) rows_update_kernel(const double *U, long long ldu, int nrows, int nb, const double *X,
                                                          long long ldx, double *Y, long long ldy, int c0, int nr) {
    extern __shared__ __align__(16) double bw_smem[];
    double *us = bw_smem;             // [64][DP]
    double *xs = bw_smem + 64 * DP;   // [NB][NR]
    const int i0 = blockIdx.x * 64;
    for (int e = threadIdx.x; e < 64 * nb; e += 256) {
        const int i = e / nb, t = e % nb;
        us[i * DP + t] = (i0 + i < nrows) ? U[(long long)(i0 + i) * ldu + t] : 0.0;
    }
    for (int e = threadIdx.x; e < nb * NR; e += 256) {
        const int t = e / NR, c = e % NR;
        xs[e] = c < nr ? X[(long long)t * ldx + c0 + c] : 0.0;
    }
    __syncthreads();
    const int i = threadIdx.x & 63, cg = threadIdx.x >> 6;
    if (i0 + i >= nrows) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < nb; ++t) {
        const double v = us[i * DP + t];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += v * xs[t * NR + cg + 4 * q];
    }
    double *b = Y + (long long)(i0 + i) * ldy + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (cg + 4 * q < nr) b[cg + 4 * q] -= acc[q];
}

// ---- prediction: scores = K A (float32 K widened to float64), pred = argmax ------------------
// classify_gp.py:39-41.  The product is skinny ([R, n] x [n, <= 16]) and its arithmetic is float64: on the
// plain FP64 pipe it is compute-bound at 4 x the time HBM needs for K (1.3 G DFMA at n = 32 768, R = 4 096), so
// it runs on the FP64 tensor pipe.  A warp owns 32 rows of K (four 8-row mma.m8n8k4 fragments that share every
// B fragment) and one chunk of the n training points; lane (lr, lk) reads four consecutive entries of its
// rows per 16 points (one 16-byte load where the row is aligned) and feeds them to four mma steps, the
// B fragments take the matching rows of A (L2-resident) -- the order of k inside a step is free as long as
// both operands agree.  Partial scores per chunk go to a scratch buffer and are summed in chunk order by
// predict_reduce_kernel (deterministic), which also takes the argmax.
constexpr int PD_ROWS = 32;  // rows of K per warp
constexpr int PD_NC = 16;    // right-hand sides per pass (two 8-column fragments)

__device__ __forceinline__ void load4(const float *p, bool vec, long long left, double (&a)[4]) {
    if (vec && left >= 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = j < left ? (double)__ldg(p + j) : 0.0;
    }
}
__device__ __forceinline__ void load4(const double *p, bool vec, long long left, double (&a)[4]) {
    if (vec && left >= 4) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p)), w = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        a[0] = v.x; a[1] = v.y; a[2] = w.x; a[3] = w.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = j < left ? __ldg(p + j) : 0.0;
    }
}

constexpr int PD_KT = 128;         // training points per shared-memory tile of A
constexpr int PD_PB = PD_NC + 1;   // its pitch: odd, so that the B fragments (lane = 4 x column + k) are conflict-free

template <typename KT>
__global__ void __launch_bounds__(128) predict_dmma_kernel(const KT *K, long long R, long long n, long long ldk,
                                                           const double *A, int nrhs, int c0, long long kc, int nchunks,
                                                           double *part, int vec) {
    __shared__ double sB[PD_KT * PD_PB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
    const long long groups = (R + 4 * PD_ROWS - 1) / (4 * PD_ROWS);  // a CTA: 4 warps x 32 rows, one chunk
    const long long rg = blockIdx.x % groups, ch = blockIdx.x / groups;
    const long long r0 = rg * 4 * PD_ROWS + warp * PD_ROWS, k_begin = ch * kc, k_end = k_begin + kc < n ? k_begin + kc : n;
    const int nc = nrhs - c0 < PD_NC ? nrhs - c0 : PD_NC;
    double acc[4][2][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;
    const KT *rowp[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const long long r = r0 + 8 * m + lr;
        rowp[m] = K + (r < R ? r : R - 1) * ldk;
    }
    for (int e = tid; e < PD_KT * PD_PB; e += 128) sB[e] = 0.0;  // columns >= nc stay zero
    for (long long kt = k_begin; kt < k_end; kt += PD_KT) {
        __syncthreads();
        const int rows = (int)(k_end - kt < PD_KT ? k_end - kt : PD_KT);
        for (int e = tid; e < PD_KT * nc; e += 128) {  // rows of A are contiguous when nc == nrhs: coalesced
            const int k = e / nc, c = e - k * nc;
            sB[k * PD_PB + c] = k < rows ? __ldg(A + (kt + k) * nrhs + c0 + c) : 0.0;
        }
        __syncthreads();
        if (r0 >= R) continue;
#pragma unroll 2
        for (int k16 = 0; k16 < PD_KT; k16 += 16) {
            if (k16 >= rows) break;
            const long long kk = kt + k16 + 4 * lk, left = k_end - kk;  // this lane's four points
            double a[4][4];
#pragma unroll
            for (int m = 0; m < 4; ++m) load4(rowp[m] + kk, vec != 0, left, a[m]);
            const double *pb = sB + (k16 + 4 * lk) * PD_PB + lr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double b0 = pb[j * PD_PB], b1 = pb[j * PD_PB + 8];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    dmma(acc[m][0][0], acc[m][0][1], a[m][j], b0);
                    dmma(acc[m][1][0], acc[m][1][1], a[m][j], b1);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const long long r = r0 + 8 * m + lr;
        if (r >= R) continue;
        double *o = part + (ch * R + r) * PD_NC + 2 * lk;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) *reinterpret_cast<double2 *>(o + 8 * nt) = make_double2(acc[m][nt][0], acc[m][nt][1]);
    }
}

// one thread per row: partial scores summed in chunk order, then the running argmax over the column passes
__global__ void __launch_bounds__(128) predict_reduce_kernel(const double *part, long long R, int nrhs, int c0, int nchunks,
                                                             long long *pred, double *scores, double *best_val) {
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= R) return;
    const int nc = nrhs - c0 < PD_NC ? nrhs - c0 : PD_NC;
    double bv = c0 == 0 ? -INFINITY : best_val[r];
    long long bi = c0 == 0 ? 0 : pred[r];
    bool any = c0 != 0;
    for (int c = 0; c < nc; ++c) {
        double v = 0.0;
        for (int ch = 0; ch < nchunks; ++ch) v += part[((long long)ch * R + r) * PD_NC + c];
        if (scores) scores[r * nrhs + c0 + c] = v;
        // first maximum wins, NaN propagates like torch.argmax (a NaN score is the maximum)
        if (!any || v > bv || (v != v && bv == bv)) { bv = v; bi = c0 + c; any = true; }
    }
    pred[r] = bi;
    if (best_val) best_val[r] = bv;
}

bool check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// Tj column tiles; row tile ib holds jb in [2 ib, Tj)
long long tri_prefix(long long ib, long long Tj) { return ib * Tj - ib * (ib - 1); }

bool ensure_attrs() {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && attr_done[dev]) return true;
    if (!check(cudaFuncSetAttribute(tn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTnSmem), "attr") ||
        !check(cudaFuncSetAttribute(tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTnSmem), "attr") ||
        !check(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotf2Smem), "attr"))
        return false;
    if (dev < 64) attr_done[dev] = true;
    return true;
}

// C[i][j] -= sum_k X[k][i] X[k][j] for j >= i, i, j < m, row tiles [ib_lo, ib_hi) of TI rows; C + i*ldc + j
void launch_syrk(const double *X, int64_t ldx, int K, double *C, int64_t ldc, int64_t m, int ib_lo, int ib_hi,
                 cudaStream_t st) {
    if (m <= 0) return;
    TnParams g{};
    g.P = X; g.ldp = ldx; g.Q = X; g.ldq = ldx; g.C = C; g.ldc = ldc;
    g.M = (int)m; g.N = (int)m; g.K = K;
    g.vec_ok = (ldx % 2 == 0) && (ldc % 2 == 0) && aligned16(X) && aligned16(C);
    g.Tj = (int)((m + TJ - 1) / TJ);
    const int Ti = (int)((m + TI - 1) / TI);
    if (ib_hi > Ti) ib_hi = Ti;
    if (ib_lo >= ib_hi) return;
    g.ib_lo = ib_lo; g.t_off = tri_prefix(ib_lo, g.Tj);
    const long long tiles = tri_prefix(ib_hi, g.Tj) - g.t_off;
    tn_kernel<0><<<(unsigned)tiles, TN_THREADS, kTnSmem, st>>>(g);
}

// the same update for a rank that owns every `stride`-th 256-row block of the trailing matrix:
// blocks ti0, ti0 + stride, ... (nblk of them), stored stacked in C_local from local block q0
void launch_syrk_strided(const double *X, int64_t ldx, int K, double *C_local, int64_t ldc, int64_t m, int ti0,
                         int stride, int q0, int nblk, cudaStream_t st) {
    if (m <= 0) return;
    const int Tj = (int)((m + TJ - 1) / TJ), Ti = (int)((m + TI - 1) / TI);
    for (int b0 = 0; b0 < nblk; b0 += kMaxRowTiles / 2) {
        const int nb = nblk - b0 < kMaxRowTiles / 2 ? nblk - b0 : kMaxRowTiles / 2;
        TnParams g{};
        g.P = X; g.ldp = ldx; g.Q = X; g.ldq = ldx; g.C = C_local; g.ldc = ldc;
        g.M = (int)m; g.N = (int)m; g.K = K;
        g.vec_ok = (ldx % 2 == 0) && (ldc % 2 == 0) && aligned16(X) && aligned16(C_local);
        g.Tj = Tj;
        g.rt_n = 2 * nb; g.rt_ti0 = ti0 + stride * b0; g.rt_stride = stride; g.rt_q0 = q0 + b0;
        int acc = 0;
        for (int k = 0; k < 2 * nb; ++k) {
            const int ib = 2 * (g.rt_ti0 + stride * (k >> 1)) + (k & 1);
            g.rt_prefix[k] = acc;
            if (ib < Ti && Tj - 2 * ib > 0) acc += Tj - 2 * ib;
        }
        g.rt_prefix[2 * nb] = acc;
        // trailing empty row tiles would break the search for the last k with prefix <= t: drop them
        while (g.rt_n > 0 && g.rt_prefix[g.rt_n - 1] == acc) --g.rt_n;
        if (acc > 0) tn_kernel<0><<<(unsigned)acc, TN_THREADS, kTnSmem, st>>>(g);
    }
}

// X = W^T X in place: X [nb, ldx] with m columns
void launch_trsm(const double *W, double *X, int64_t ldx, int nb, int64_t m, cudaStream_t st) {
    if (m <= 0) return;
    TnParams g{};
    g.P = W; g.ldp = NB; g.Q = X; g.ldq = ldx; g.C = X; g.ldc = ldx;
    g.M = nb; g.N = (int)m; g.K = nb;
    g.vec_ok = (ldx % 2 == 0) && aligned16(X) && aligned16(W);
    g.Tj = (int)((m + TJ - 1) / TJ);
    tn_kernel<1><<<(unsigned)g.Tj, TN_THREADS, kTnSmem, st>>>(g);
}

// factorise the block row P[0:rows, 0:width] (rows <= 2 NB; P[0][0] is its diagonal element): two
// diagonal blocks and their row panels, with the rank-128 update of the second half in between
void launch_panel(double *P, int64_t ldp, int64_t width, double *W, int *info, int64_t info_base, cudaStream_t st) {
    const int nb1 = (int)(width < NB ? width : NB);
    potf2_inv_kernel<<<1, POTF2_THREADS, kPotf2Smem, st>>>(P, ldp, 0, nb1, W, info, info_base);
    if (width <= NB) return;
    launch_trsm(W, P + NB, ldp, nb1, width - NB, st);
    launch_syrk(P + NB, ldp, nb1, P + NB * ldp + NB, ldp, width - NB, 0, 1, st);  // the second half only
    const int nb2 = (int)(width - NB < NB ? width - NB : NB);
    potf2_inv_kernel<<<1, POTF2_THREADS, kPotf2Smem, st>>>(P, ldp, NB, nb2, W, info, info_base);
    launch_trsm(W, P + NB * ldp + 2 * NB, ldp, nb2, width - 2 * NB, st);
}

}  // namespace
}  // namespace cnngp

using namespace cnngp;

template <typename KT>
static int predict_any(const KT *d_K, int64_t R, int64_t n, int64_t ldk, const double *d_A, int32_t nrhs,
                       int64_t *d_pred, double *d_scores, void *stream_, const char *who) {
    if (!d_K || !d_A || !d_pred || R < 0 || n < 0 || ldk < n || nrhs < 1) { set_error(std::string(who) + ": bad arguments"); return 1; }
    if (R == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_;
    double *best = nullptr, *part = nullptr;
    if (nrhs > PD_NC && !check(cudaMallocAsync((void **)&best, sizeof(double) * R, s), "predict workspace")) return 6;
    const long long groups = (R + 4 * PD_ROWS - 1) / (4 * PD_ROWS);
    // chunks of whole shared-memory tiles, chosen so that the CTAs fill whole waves (four resident per SM): the cost of
    // a split is waves x (points per chunk + a fixed per-CTA share)
    long long nchunks = 1, kc = (n + PD_KT - 1) / PD_KT * PD_KT;
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long resident = 4LL * sms, max_chunks = (n + 2 * PD_KT - 1) / (2 * PD_KT);
        double best = 1e300;
        for (long long c = 1; c <= max_chunks && c <= 512; ++c) {
            const long long k = ((n + c - 1) / c + PD_KT - 1) / PD_KT * PD_KT, ch = (n + k - 1) / k;
            const long long waves = (groups * ch + resident - 1) / resident;
            const double cost = (double)waves * (double)(k + 2 * PD_KT);
            if (cost < best) { best = cost; kc = k; nchunks = ch; }
        }
    }
    if (nchunks < 1) nchunks = 1;
    pool_keep((size_t)64 << 20);
    if (!check(cudaMallocAsync((void **)&part, sizeof(double) * nchunks * R * PD_NC, s), "predict workspace")) {
        if (best) cudaFreeAsync(best, s);
        return 6;
    }
    // 16-byte loads of four consecutive entries need aligned rows
    const int vec = ((uintptr_t)d_K % 16 == 0) && (ldk * sizeof(KT)) % 16 == 0;
    const unsigned grid = (unsigned)(groups * nchunks), rgrid = (unsigned)((R + 127) / 128);
    for (int c0 = 0; c0 < nrhs; c0 += PD_NC) {
        predict_dmma_kernel<KT><<<grid, 128, 0, s>>>(d_K, R, n, ldk, d_A, nrhs, c0, kc, (int)nchunks, part, vec);
        predict_reduce_kernel<<<rgrid, 128, 0, s>>>(part, R, nrhs, c0, (int)nchunks, (long long *)d_pred, d_scores, best);
    }
    cudaFreeAsync(part, s);
    if (best) cudaFreeAsync(best, s);
    return check(cudaGetLastError(), who) ? 0 : 9;
}

extern "C" {

// ---- building blocks of a right-looking Cholesky, for drivers that keep block rows on several GPUs
int cnngp_potrf_panel_f64(double *d_P, int64_t ldp, int64_t width, int64_t info_base, int32_t *d_info,
                          double *d_work, void *stream_) {
    if (!d_P || width < 1 || ldp < width || !d_info || !d_work) { set_error("cnngp_potrf_panel_f64: bad arguments"); return 1; }
    if (!ensure_attrs()) return 7;
    launch_panel(d_P, ldp, width, d_work, d_info, info_base, (cudaStream_t)stream_);
    return check(cudaGetLastError(), "cnngp_potrf_panel_f64") ? 0 : 9;
}

int cnngp_syrk_upper_f64(const double *d_X, int64_t ldx, int32_t K, double *d_C, int64_t ldc, int64_t m, int32_t ib_lo,
                         int32_t ib_hi, void *stream_) {
    if (!d_X || !d_C || K < 1 || K > 2 * NB || m < 0 || ldx < m || ib_lo < 0) { set_error("cnngp_syrk_upper_f64: bad arguments"); return 1; }
    if (!ensure_attrs()) return 7;
    launch_syrk(d_X, ldx, K, d_C, ldc, m, ib_lo, ib_hi, (cudaStream_t)stream_);
    return check(cudaGetLastError(), "cnngp_syrk_upper_f64") ? 0 : 9;
}

int cnngp_syrk_upper_strided_f64(const double *d_X, int64_t ldx, int32_t K, double *d_C_local, int64_t ldc, int64_t m,
                                 int32_t ti0, int32_t stride, int32_t q0, int32_t n_blocks, void *stream_) {
    if (!d_X || !d_C_local || K < 1 || K > 2 * NB || m < 0 || ldx < m || ti0 < 0 || stride < 1 || q0 < 0 || n_blocks < 0) {
        set_error("cnngp_syrk_upper_strided_f64: bad arguments");
        return 1;
    }
    if (!ensure_attrs()) return 7;
    launch_syrk_strided(d_X, ldx, K, d_C_local, ldc, m, ti0, stride, q0, n_blocks, (cudaStream_t)stream_);
    return check(cudaGetLastError(), "cnngp_syrk_upper_strided_f64") ? 0 : 9;
}

int cnngp_potrf_upper_f64(double *d_A, int64_t n, int64_t lda, int32_t *d_info, void *stream_) {
    if (!d_A || n < 0 || lda < n || !d_info) { set_error("cnngp_potrf_upper_f64: bad arguments"); return 1; }
    if (n > 2000000000LL) { set_error("cnngp_potrf_upper_f64: n too large"); return 1; }
    cudaStream_t s = (cudaStream_t)stream_;
    if (!check(cudaMemsetAsync(d_info, 0, sizeof(int32_t), s), "potrf memset")) return 5;
    if (n == 0) return 0;
    if (!ensure_attrs()) return 7;
    double *W = nullptr;
    if (!check(cudaMallocAsync((void **)&W, sizeof(double) * NB * NB, s), "potrf workspace")) return 6;
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_head = nullptr, ev_panel = nullptr;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (!check(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio_hi), "potrf stream") ||
        !check(cudaEventCreateWithFlags(&ev_head, cudaEventDisableTiming), "potrf event") ||
        !check(cudaEventCreateWithFlags(&ev_panel, cudaEventDisableTiming), "potrf event")) {
        if (s2) cudaStreamDestroy(s2);
        if (ev_head) cudaEventDestroy(ev_head);
        cudaFreeAsync(W, s);
        return 6;
    }
    constexpr int NBO = 2 * NB;
    // trailing update with the block row at kb: C = A[r0:, r0:], X = A[kb:kb+NBO, r0:]
    auto update = [&](int64_t kb, int ib_lo, int ib_hi, cudaStream_t st) {
        const int64_t r0 = kb + NBO;
        launch_syrk(d_A + kb * lda + r0, lda, NBO, d_A + r0 * lda + r0, lda, n - r0, ib_lo, ib_hi, st);
    };
    auto panel = [&](int64_t kb, cudaStream_t st) {
        launch_panel(d_A + kb * lda + kb, lda, n - kb, W, d_info, kb, st);
    };

    panel(0, s);
    for (int64_t kb = 0; kb + NBO < n; kb += NBO) {
        update(kb, 0, NBO / TI, s);  // look-ahead: the next panel's rows first
        cudaEventRecord(ev_head, s);
        cudaStreamWaitEvent(s2, ev_head, 0);
        panel(kb + NBO, s2);
        cudaEventRecord(ev_panel, s2);
        update(kb, NBO / TI, 1 << 30, s);
        cudaStreamWaitEvent(s, ev_panel, 0);
    }
    cudaFreeAsync(W, s);
    cudaEventDestroy(ev_head);
    cudaEventDestroy(ev_panel);
    cudaStreamDestroy(s2);
    return check(cudaGetLastError(), "cnngp_potrf_upper_f64") ? 0 : 9;
}

static bool ensure_solve_attrs() {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && attr_done[dev]) return true;
    if (!check(cudaFuncSetAttribute(trsv_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsvSmem), "attr") ||
        !check(cudaFuncSetAttribute(trsv_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsvSmem), "attr") ||
        !check(cudaFuncSetAttribute(bwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem), "attr") ||
        !check(cudaFuncSetAttribute(rows_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem), "attr") ||
        !check(cudaFuncSetAttribute(sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem), "attr") ||
        !check(cudaFuncSetAttribute(sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem), "attr"))
        return false;
    if (dev < 64) attr_done[dev] = true;
    return true;
}

int cnngp_potrs_upper_f64(const double *d_U, int64_t n, int64_t lda, double *d_B, int32_t nrhs, int64_t ldb,
                          void *stream_) {
    if (!d_U || !d_B || n < 0 || lda < n || nrhs < 0 || ldb < nrhs) { set_error("cnngp_potrs_upper_f64: bad arguments"); return 1; }
    if (n == 0 || nrhs == 0) return 0;
    if (n > 2000000000LL) { set_error("cnngp_potrs_upper_f64: n too large"); return 1; }
    cudaStream_t s = (cudaStream_t)stream_;
    if (!ensure_solve_attrs()) return 7;
    const int nblk = (int)((n + NB - 1) / NB);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the sweep's CTAs wait for each other: all of them must be resident (one per SM: 150 KB of shared memory each)
    const unsigned grid = (unsigned)(nblk < sms ? nblk : sms);
    int *flags = nullptr;
    if (!check(cudaMallocAsync((void **)&flags, sizeof(int) * nblk, s), "potrs flags")) return 6;
    // cooperative launches: the runtime guarantees that all CTAs are resident together (or fails the launch),
    // which the flag dataflow relies on
    long long lda_ = lda, ldb_ = ldb;
    int n_ = (int)n;
    bool ok = true;
    for (int c0 = 0; c0 < nrhs && ok; c0 += SW_NR) {
        int nr = nrhs - c0 < SW_NR ? nrhs - c0 : SW_NR, c0_ = c0;
        void *args[] = {(void *)&d_U, (void *)&lda_, (void *)&n_, (void *)&d_B, (void *)&ldb_, (void *)&c0_, (void *)&nr, (void *)&flags};
        cudaMemsetAsync(flags, 0, sizeof(int) * nblk, s);
        ok = check(cudaLaunchCooperativeKernel((const void *)sweep_kernel<true>, dim3(grid), dim3(SW_THREADS), args, kSweepSmem, s),
                   "cnngp_potrs_upper_f64 (forward sweep)");  // U^T y = b
        if (!ok) break;
        cudaMemsetAsync(flags, 0, sizeof(int) * nblk, s);
        ok = check(cudaLaunchCooperativeKernel((const void *)sweep_kernel<false>, dim3(grid), dim3(SW_THREADS), args, kSweepSmem, s),
                   "cnngp_potrs_upper_f64 (backward sweep)");  // U x = y
    }
    cudaFreeAsync(flags, s);
    if (!ok) return 9;
    return check(cudaGetLastError(), "cnngp_potrs_upper_f64") ? 0 : 9;
}

// ---- building blocks of the two sweeps for drivers that keep block rows of U on several GPUs --------
// (cnn_gp/linalg_dist.py).  A "panel" is one block row of U as its owner stores it: d_P points at
// the diagonal element, `rows` (<= 256) rows of `width` columns each.
//
// forward: the rows' right-hand sides d_B[0:rows] (already reduced over the ranks) become y; the
// rank's accumulator d_B[rows:width] takes  -= U[panel rows, later columns]^T y.
int cnngp_trsm_fwd_panel_f64(const double *d_P, int64_t ldp, int64_t rows, int64_t width, double *d_B, int32_t nrhs,
                             int64_t ldb, void *stream_) {
    if (!d_P || !d_B || rows < 1 || width < rows || ldp < width || nrhs < 1 || ldb < nrhs) {
        set_error("cnngp_trsm_fwd_panel_f64: bad arguments");
        return 1;
    }
    cudaStream_t s = (cudaStream_t)stream_;
    if (!ensure_solve_attrs()) return 7;
    for (int c0 = 0; c0 < nrhs; c0 += NR) {
        const int nr = nrhs - c0 < NR ? nrhs - c0 : NR;
        for (int64_t kb = 0; kb < rows; kb += NB) {
            const int nb = (int)(rows - kb < NB ? rows - kb : NB);
            trsv_block_kernel<true><<<1, 512, kTrsvSmem, s>>>(d_P, ldp, (int)kb, nb, d_B, ldb, c0, nr);
            const int64_t m = width - kb - nb;
            if (m > 0)
                fwd_update_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(d_P, ldp, (int)kb, nb, (int)width, d_B, ldb, c0, nr);
        }
    }
    return check(cudaGetLastError(), "cnngp_trsm_fwd_panel_f64") ? 0 : 9;
}

// backward, the diagonal block only: d_B[0:rows] (y minus the updates of all later blocks) becomes x
int cnngp_trsm_bwd_diag_f64(const double *d_P, int64_t ldp, int64_t rows, double *d_B, int32_t nrhs, int64_t ldb,
                            void *stream_) {
    if (!d_P || !d_B || rows < 1 || ldp < rows || nrhs < 1 || ldb < nrhs) { set_error("cnngp_trsm_bwd_diag_f64: bad arguments"); return 1; }
    cudaStream_t s = (cudaStream_t)stream_;
    if (!ensure_solve_attrs()) return 7;
    for (int c0 = 0; c0 < nrhs; c0 += NR) {
        const int nr = nrhs - c0 < NR ? nrhs - c0 : NR;
        const int64_t last = ((rows - 1) / NB) * NB;
        for (int64_t kb = last; kb >= 0; kb -= NB) {
            const int nb = (int)(rows - kb < NB ? rows - kb : NB);
            trsv_block_kernel<false><<<1, 512, kTrsvSmem, s>>>(d_P, ldp, (int)kb, nb, d_B, ldb, c0, nr);
            if (kb > 0)
                bwd_update_kernel<<<(unsigned)((kb + 63) / 64), 256, kBwdSmem, s>>>(d_P, ldp, (int)kb, nb, d_B, ldb, c0, nr);
        }
    }
    return check(cudaGetLastError(), "cnngp_trsm_bwd_diag_f64") ? 0 : 9;
}

// backward update of a rank's stacked rows: d_Y[i] -= d_U[i][0:nb] d_X for i < nrows (d_U points at the
// solved block's columns inside the rank's local rows, d_X [nb, ldx] is the broadcast solution block)
int cnngp_rows_update_f64(const double *d_U, int64_t ldu, int64_t nrows, int32_t nb, const double *d_X, int64_t ldx,
                          double *d_Y, int64_t ldy, int32_t nrhs, void *stream_) {
    if (!d_U || !d_X || !d_Y || nrows < 0 || nb < 1 || nb > 2 * NB || ldu < nb || nrhs < 1 || ldx < nrhs || ldy < nrhs) {
        set_error("cnngp_rows_update_f64: bad arguments");
        return 1;
    }
    if (nrows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_;
    if (!ensure_solve_attrs()) return 7;
    for (int c0 = 0; c0 < nrhs; c0 += NR)
        for (int t0 = 0; t0 < nb; t0 += NB) {  // the staged tile is NB columns wide
            const int nbt = nb - t0 < NB ? nb - t0 : NB;
            rows_update_kernel<<<(unsigned)((nrows + 63) / 64), 256, kBwdSmem, s>>>(
                d_U + t0, ldu, (int)nrows, nbt, d_X + (int64_t)t0 * ldx, ldx, d_Y, ldy, c0, nrhs - c0 < NR ? nrhs - c0 : NR);
        }
    return check(cudaGetLastError(), "cnngp_rows_update_f64") ? 0 : 9;
}

int cnngp_predict_argmax(const float *d_K, int64_t R, int64_t n, int64_t ldk, const double *d_A, int32_t nrhs,
                         int64_t *d_pred, double *d_scores, void *stream_) {
    return predict_any<float>(d_K, R, n, ldk, d_A, nrhs, d_pred, d_scores, stream_, "cnngp_predict_argmax");
}

int cnngp_predict_argmax_f64(const double *d_K, int64_t R, int64_t n, int64_t ldk, const double *d_A, int32_t nrhs,
                             int64_t *d_pred, double *d_scores, void *stream_) {
    return predict_any<double>(d_K, R, n, ldk, d_A, nrhs, d_pred, d_scores, stream_, "cnngp_predict_argmax_f64");
}

}  // extern "C"
