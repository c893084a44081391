// gram_fused.cu -- register-resident fused Gram kernel for sm_100a.
//
// One warp owns a 2 x 2 block of image pairs (twelve consumer warps per CTA, one CTA per SM, a
// 4 x 12-image tile) and keeps their four covariance maps in registers for the whole layer stack: lane = one coordinate of the map, register index = the
// other.  A box convolution is a sliding-window sum along the register axis, a transposition
// through the warp's private shared-memory tile, and a second sliding-window sum; every layer
// therefore flips the map between "lane = column" and "lane = row" layout.  The ReLU step is
// element-wise in registers and reads the per-image standard deviations (s, 1/s) of the eight
// warps' twelve images from a double-buffered shared-memory stage that a producer warp fills
// one layer ahead with bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP).  The dense
// last layer is a register sum plus a warp-shuffle reduction.  Nothing but the final kernel
// entry is written to HBM.
//
// Call forms: rectangular K(X, Z); symmetric model(X) (only j >= i computed, mirrored); a BAND of block rows
// of model(X) -- rows [0, N1) x columns [0, N2 >= N1), mirrored inside the diagonal blocks only, which is a
// worker's run of whole block rows of the reference's tile list (cnngp_gram_band) --; and the symmetric
// call with progress counters per band of rows (template parameter PROG), behind which a copy stream
// moves finished bands to host memory while the launch runs (cnngp_gram_symmetric_to_host).
//
// Reference semantics (paths relative to /root/reference):
//   init   cnn_gp/kernels.py:43-49     conv  cnn_gp/kernels.py:92-98
//   relu   cnn_gp/kernels.py:146-152 rewritten as
//            xy' = relu(c)/2 + s * e^1.5 * H(e),  s = sqrt(xx*yy), e = 1 - |c|/s
//          where H is analytic on [0,1] (degree-6 minimax fit, error below float32 rounding);
//          this needs one MUFU.SQRT per pixel instead of rsqrt + sqrt + acos + divide.
//          The factor 1/2 is folded into the next convolution's tap (exact: power of two).
#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fused_common.cuh"
#include "plan.h"

namespace cnngp {

namespace {

// Geometry of one kernel variant: NW consumer warps as 2 x NW/2 warps of 2 x 2 image pairs,
// plus a fourth/third warpgroup that holds the producer warp and donates registers (setmaxnreg).
// NG > 1: the consumer warps form NG independent groups, each with its own tile stream, staging ring
// and producer warp -- like NG small CTAs sharing the SM.  Groups drift apart, so that one group's
// ReLU (FP32 pipe) runs under another group's transposition (shared-memory pipe) instead of all
// warps of the SM hitting the same pipe in the same phase.
template <int NW, int NG = 1>
struct Geo {
    static constexpr int kWarps = NW;
    static constexpr int kGroups = NG, kGW = NW / NG;     // consumer warps per group
    static constexpr int kTileI = 4, kTileJ = NW / NG;    // images per group tile along i and j
    static constexpr int kImgs = kTileI + kTileJ;
    static constexpr int kPairs = kImgs / 2;              // image pairs whose interleaved variance maps are staged
    static constexpr int kThreads = (NW + 4) * 32;
    static constexpr int kRegsProducer = NW == 8 ? 24 : 32;  // what the pool holds: launch registers x threads
    static constexpr int kRegsConsumer = NW == 8 ? 240 : 160;  // 8*32*240 + 4*32*32 = 12*32*160 + 4*32*32 = 65536
};
constexpr int kMaxOps = 40;
constexpr int kSuperEdge = 504;              // super-tile edge in images (L2-resident variance maps); 4 | 504, 8 | 504, 12 | 504

enum { F_CONV = 0, F_RELU = 1, F_DENSE = 2 };

struct FOp {
    int kind;
    int lo, hi;     // F_CONV: window offsets [-lo, +hi] along each axis
    float pre_bias; // F_CONV with a window: added inside the second sliding sum (two adds per map, see below)
    float scale, bias;  // F_CONV: applied as one FMA pass only when scale != 1 or the conv is pointwise; F_DENSE
    int aux_off;    // F_RELU: offset (floats) of this layer inside the fused section of a row
};

struct FParams {
    FOp ops[kMaxOps];
    int n_ops, n_relu;
    const float *x, *z;          // images [N, C, S, S]
    const float *aux_x, *aux_z;  // per-image rows; the fused maps start at aux_f_off floats; rows 2k and
                                 // 2k+1 together hold pair k's (s_2k, s_2k+1, 1/s_2k, 1/s_2k+1) maps
    long long aux_stride;        // floats per image row
    int aux_f_off;
    int N1, N2, C;
    float *out;
    long long ld_out;
    int symmetric;
    int mirror_bs;     // symmetric with N2 > N1 (a band of block rows): mirror only inside diagonal blocks of this many rows
    const float *kdiag;
    int nbi, nbj;      // CTA tiles along i / j
    int sti, stj;      // super-tile size in CTA tiles
    int nst_j;         // super-tiles along j (non-symmetric)
    int nst;           // super-tiles per side (symmetric)
    long long n_tiles; // CTA tiles enumerated (super-tile padded)
    unsigned long long *tile_ctr;  // zeroed before the launch: the next tile index to hand out
    unsigned *row_done;            // optional: finished (tile, warp) units per super-row (RowProgress)
    float inv_c;       // 1 / C
};

using namespace fusedk;

// Box sum along the register axis with zero padding, out[y] = sum_{t=-LO..HI} v[y+t], on two
// maps at once (packed lanes), as two sliding windows that start at the two ends and meet in
// the middle: independent dependency chains (ILP) and half the rounding-error accumulation of
// one long slide.
// BIAS: both windows start from v + B, so every output carries the constant B (the folded conv bias).
template <int S, int LO, int HI, bool BIAS>
__device__ __forceinline__ void box_pass(u64 (&v)[S], u64 B) {
    if (LO == 0 && HI == 0) return;
    constexpr int MID = S / 2;
    u64 o[S];
    u64 top = v[0], bot = v[S - 1];
    if (BIAS) { top = add2(top, B); bot = add2(bot, B); }
#pragma unroll
    for (int t = 1; t <= HI && t < S; ++t) top = add2(top, v[t]);
#pragma unroll
    for (int t = 1; t <= LO && t < S; ++t) bot = add2(bot, v[S - 1 - t]);
    o[0] = top;
    o[S - 1] = bot;
#pragma unroll
    for (int y = 1; y < MID; ++y) {
        if (y + HI < S) top = add2(top, v[y + HI]);
        if (y - LO - 1 >= 0) top = sub2(top, v[y - LO - 1]);
        o[y] = top;
    }
#pragma unroll
    for (int y = S - 2; y >= MID; --y) {
        if (y - LO >= 0) bot = add2(bot, v[y - LO]);
        if (y + HI + 1 < S) bot = sub2(bot, v[y + HI + 1]);
        o[y] = bot;
    }
#pragma unroll
    for (int y = 0; y < S; ++y) v[y] = o[y];
}

// LO < 0 selects the window at run time (programs that mix window shapes).
template <int S, int LO, int HI, bool BIAS>
__device__ __forceinline__ void box_any(u64 (&v)[S], int lo, int hi, u64 B) {
    if (LO >= 0) {
        box_pass<S, (LO >= 0 ? LO : 0), HI, BIAS>(v, B);
    } else {
        if (lo == 3 && hi == 3) box_pass<S, 3, 3, BIAS>(v, B);
        else if (lo == 1 && hi == 1) box_pass<S, 1, 1, BIAS>(v, B);
        else if (lo == 1 && hi == 2) box_pass<S, 1, 2, BIAS>(v, B);
        else if (lo == 2 && hi == 2) box_pass<S, 2, 2, BIAS>(v, B);
    }
}

template <int S>
__device__ __forceinline__ void tile_store(u64 *tile, const u64 (&a)[S], int lane) {
    constexpr int PITCH = S + 1;  // odd: row-wise writes and column-wise reads are both conflict-free
    // lanes >= S hold nothing: they all write the pad column (no branch around the stores, so the
    // compiler can sink them into the sliding sums that produce the values)
    const int col = lane < S ? lane : S;
#pragma unroll
    for (int r = 0; r < S; ++r) tile[r * PITCH + col] = a[r];
}
template <int S>
__device__ __forceinline__ void tile_load_t(const u64 *tile, u64 (&a)[S], int lx) {
    constexpr int PITCH = S + 1;
#pragma unroll
    for (int r = 0; r < S; ++r) a[r] = tile[lx * PITCH + r];
}

// Warp w of a CTA tile owns images i0 = 2*(2*ib + wi) + {0,1} and j0 = 2*((NW/2)*jb + wj) + {0,1}.
// Its four maps live in two packed arrays:  M[0][r] = (i0j0, i1j1),  M[1][r] = (i0j1, i1j0),
// so that every packed operation multiplies the (i0, i1) pair of one image pair with the
// (j0, j1) pair -- or its swap -- of another: no broadcasts are needed.
//
// Staging: a ring of NST stages.  A ReLU layer's pair-interleaved variance maps arrive in NSPLIT
// row bands (one stage each), a channel of the tile's images in IMG_PARTS bands.
// PROG: the launch reports finished tiles per super-row (RowProgress).  A separate instantiation: the consumer
// warps run at the register limit, and even the few instructions of the reporting path cost the plain launch
// 0.7 % through a different register allocation.
template <int S, int LO, int HI, int NW, int NSPLIT, int NST, int NG = 1, bool PROG = false>
__global__ void __launch_bounds__(Geo<NW, NG>::kThreads, 1) fused_kernel(const __grid_constant__ FParams p) {
    using G = Geo<NW, NG>;
    constexpr int kWarps = G::kWarps, kTileI = G::kTileI, kTileJ = G::kTileJ, kImgs = G::kImgs, kPairs = G::kPairs;
    constexpr int kGW = G::kGW;
    constexpr int P = S * S;
    constexpr int PITCH = S + 1;
    constexpr int IMG_PARTS = NSPLIT == 4 ? 2 : 1;
    constexpr int BAND = P / NSPLIT;                  // pixels of one ReLU band
    constexpr int STAGE_F4 = kPairs * BAND;           // float4 per stage
    constexpr int IBAND = P / IMG_PARTS;              // pixels of one image band
    static_assert(S % NSPLIT == 0 && (NSPLIT == 1 || NSPLIT == 2 || NSPLIT == 4), "band split");
    static_assert(kImgs * IBAND * 4 <= STAGE_F4 * 16, "an image band must fit a stage");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [NG][NST][kPairs][BAND] float4 stages | [kWarps][S*PITCH] u64 transpose tiles | barriers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp < kWarps ? warp / kGW : (warp - kWarps < NG ? warp - kWarps : 0);  // consumer group / the group a producer serves
    float4 *stage = reinterpret_cast<float4 *>(smem_raw) + grp * NST * STAGE_F4;
    u64 *tiles = reinterpret_cast<u64 *>(reinterpret_cast<float4 *>(smem_raw) + NG * NST * STAGE_F4);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + kWarps * S * PITCH) + grp * 3 * NST;
    uint64_t *full = bars, *empty = bars + NST;
    // tile index each stage belongs to (-1: no more tiles).  Tiles are handed out by a global
    // atomic counter, not by a fixed stride: all CTAs then work on consecutive tiles of one
    // super-tile at any time (static strides let CTAs drift apart over the triangular
    // enumeration, and the variance rows of several super-tiles competed for L2), and the tail
    // of the launch is balanced to one tile.
    long long *stage_tile = reinterpret_cast<long long *>(bars + 2 * NST);
    // RowProgress: the consumers arrive on done[k & 1] after writing the entries of their k-th tile; the producer
    // lane -- which has the time and the registers -- makes them visible device-wide and moves the counter, one
    // fence per CTA tile.  (A fence + atomic per tile in every consumer warp cost 1.8 % of the launch: 197.8 ms
    // against 194.3 ms without counters.)
    uint64_t *done = reinterpret_cast<uint64_t *>(tiles + kWarps * S * PITCH) + NG * 3 * NST + grp * 2;

    if (threadIdx.x == 0) {
        for (int g = 0; g < NG; ++g) {
            for (int b = 0; b < NST; ++b) { mbar_init(&full[g * 3 * NST + b], 1); mbar_init(&empty[g * 3 * NST + b], kGW); }
            for (int b = 0; b < 2; ++b) mbar_init(&done[(g - grp) * 2 + b], kGW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (PROG && threadIdx.x < kWarps) reinterpret_cast<unsigned *>(done + (NG - grp) * 2)[threadIdx.x] = 0u;  // tiles per consumer warp
    __syncthreads();

    const int per_st = p.sti * p.stj;
    // CTA tile t -> (ib, jb); false when the tile is outside the matrix or below the diagonal
    auto decode = [&](long long t, int &ib, int &jb) -> bool {
        const int st = (int)(t / per_st), w_in = (int)(t - (long long)st * per_st);
        int si, sj;
        if (p.symmetric) {  // upper-triangular enumeration of square super-tiles
            int r = 0, rem = st;
            while (rem >= p.nst - r) { rem -= p.nst - r; ++r; }
            si = r; sj = r + rem;
        } else {
            si = st / p.nst_j; sj = st - si * p.nst_j;
        }
        ib = si * p.sti + w_in / p.stj;
        jb = sj * p.stj + w_in % p.stj;
        if (ib >= p.nbi || jb >= p.nbj) return false;
        if (p.symmetric && jb * kTileJ + (kTileJ - 1) < ib * kTileI) return false;
        return true;
    };

    if (warp >= kWarps) {
        // last warpgroup: hand its registers to the consumers; only its first warp works
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::kRegsProducer));
        if (warp >= kWarps + NG) return;
        // ---- producer warp: stages, up to NST - 1 stages ahead, first the tile's images (per
        // channel) and then each ReLU layer's pair-interleaved (s_a, s_b, 1/s_a, 1/s_b) maps of the
        // tile's image pairs.  It runs on across tile boundaries, so the next tile's first stages
        // are in flight while the consumers finish.  Lane 0 owns the ring protocol; the copies of a
        // stage are issued by the lanes in parallel, each from a base pointer worked out once per
        // tile (lane u < 2 kPairs: image row 2 pr + (u & 1) of tile pair u >> 1; lane s < kImgs:
        // tile image s).
        {
            unsigned l = 0;  // running stage counter over all tiles of this CTA
            const int last_pi = (p.N1 - 1) >> 1, last_pj = (p.N2 - 1) >> 1;
            long long t = 0;
            auto acquire = [&](unsigned bytes) -> float4 * {
                const unsigned buf = l % NST;
                if (lane == 0) {
                    if (l >= NST) mbar_wait_relaxed(&empty[buf], ((l / NST) - 1) & 1);
                    stage_tile[buf] = t;  // published by the release of the arrive below
                    mbar_arrive_expect_tx(&full[buf], bytes);
                }
                __syncwarp();
                return stage + buf * STAGE_F4;
            };
            // tile_ctr == NULL: fixed stride (tile = blockIdx.x + k * gridDim.x), kept for comparison
            const bool dyn = p.tile_ctr != nullptr;
            auto next_index = [&](long long stat) -> long long {
                if (!dyn) return stat;
                long long v = 0;
                if (lane == 0) v = (long long)atomicAdd(p.tile_ctr, 1ull);
                return __shfl_sync(0xffffffffu, v, 0);
            };
            const long long stride = (long long)gridDim.x * NG;  // fixed-stride order: groups interleave
            long long t_raw = next_index((long long)blockIdx.x * NG + grp);
            // RowProgress: tile k of this group is complete when all consumer warps have arrived on done[k & 1]
            // (phase k >> 1).  When the ring lets tile k + 2 start, every consumer is deep inside tile k + 1, so
            // the wait for tile k returns at once; only the last two tiles are really waited for.
            unsigned seq = 0;            // tiles of this group so far
            int row_a = 0, row_b = 0;    // super-rows of tiles seq - 2 and seq - 1
            auto report = [&](unsigned k, int row) {
                if (lane == 0) {
                    mbar_wait_relaxed(&done[k & 1], (k >> 1) & 1);
                    __threadfence();  // cumulative: the consumers' entries, observed through the barrier
                    atomicAdd(&p.row_done[row], (unsigned)kGW);
                }
            };
            for (;;) {
                int ib, jb;
                t = t_raw;
                while (t < p.n_tiles && !decode(t, ib, jb)) t = next_index(t + stride);
                if (t >= p.n_tiles) break;
                // the report of tile seq - 2 waits until this tile's first layers are staged: at a tile boundary the
                // ring is at its tightest (the image stage is consumed in no time), later the producer only waits
                bool owe = PROG && seq >= 2;
                const int row_old = row_a;
                if (PROG) { row_a = row_b; row_b = ib / p.sti; }
                ++seq;
                // the next index is requested now and first looked at when this tile's stages are out
                t_raw = next_index(t + stride);
                const int i_base = ib * kTileI, j_base = jb * kTileJ;
                const float *img = nullptr, *var = nullptr;  // this lane's sources for the whole tile
                if (lane < kImgs)
                    img = lane < kTileI ? p.x + (long long)min(i_base + lane, p.N1 - 1) * p.C * P
                                        : p.z + (long long)min(j_base + lane - kTileI, p.N2 - 1) * p.C * P;
                const int vs = lane >> 1, vh = lane & 1;  // pair of the tile, image row of the pair
                if (lane < 2 * kPairs) {
                    const long long pr = vs < kTileI / 2 ? min((i_base >> 1) + vs, last_pi)
                                                         : min((j_base >> 1) + vs - kTileI / 2, last_pj);
                    var = (vs < kTileI / 2 ? p.aux_x : p.aux_z) + (2 * pr + vh) * p.aux_stride + p.aux_f_off;
                }
                for (int c = 0; c < p.C; ++c) {
                    for (int ip = 0; ip < IMG_PARTS; ++ip) {
                        float *dst = reinterpret_cast<float *>(acquire(kImgs * IBAND * 4));
                        if (lane < kImgs) bulk_g2s(dst + lane * IBAND, img + (long long)c * P + ip * IBAND, IBAND * 4, &full[l % NST]);
                        ++l;
                    }
                }
                for (int k = 0; k < p.n_ops; ++k) {
                    if (p.ops[k].kind != F_RELU) continue;
                    const int off = p.ops[k].aux_off;
                    for (int q = 0; q < NSPLIT; ++q) {
                        float4 *dst = acquire(STAGE_F4 * 16);
                        uint64_t *bar = &full[l % NST];
                        // the float4 map of a pair is split over the two images' rows (first half of the pixels in row 2k)
                        if (lane < 2 * kPairs) {
                            if (NSPLIT == 1) bulk_g2s(dst + vs * P + vh * (P / 2), var + off, P * 8, bar);
                            else if (NSPLIT == 2) { if (vh == q) bulk_g2s(dst + vs * BAND, var + off, BAND * 16, bar); }
                            else if (vh == (q >> 1)) bulk_g2s(dst + vs * BAND, var + off + (q & 1) * BAND * 4, BAND * 16, bar);
                        }
                        ++l;
                    }
                    if (owe) { report(seq - 3, row_old); owe = false; }
                }
                if (owe) report(seq - 3, row_old);
            }
            t = -1;  // end marker: one empty stage
            acquire(0);
            if (PROG) {
                if (seq >= 2) report(seq - 2, row_a);
                if (seq >= 1) report(seq - 1, row_b);
            }
        }
        return;
    }

    // ---- consumers ------------------------------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::kRegsConsumer));
    const int wi = (warp % kGW) / (kGW / 2), wj = (warp % kGW) % (kGW / 2);
    const int lx = lane < S ? lane : S - 1;  // clamped lane for loads
    u64 *tile = tiles + warp * S * PITCH;
    unsigned stage_l = 0;  // running stage counter, in step with the producer's
    // 2*H(e), degree-5 minimax fit on [0,1]: |error| < 2.1e-7 relative (one float32 ulp is 1.2e-7)
    const u64 C5 = pk(1.678542030e-04f, 1.678542030e-04f), C4 = pk(-1.571319990e-05f, -1.571319990e-05f),
              C3 = pk(6.585370866e-04f, 6.585370866e-04f), C2 = pk(2.386197913e-03f, 2.386197913e-03f),
              C1 = pk(1.500756294e-02f, 1.500756294e-02f), C0 = pk(3.001053929e-01f, 3.001053929e-01f),
              ONE = pk(1.f, 1.f);

    for (;;) {
        // the tile this CTA works on next travels with its first stage
        mbar_wait(&full[stage_l % NST], (stage_l / NST) & 1);
        const long long t = stage_tile[stage_l % NST];
        if (t < 0) break;
        int ib, jb;
        decode(t, ib, jb);
        const int i_base = ib * kTileI, j_base = jb * kTileJ;

        u64 M[2][S];  // lane = column, register = row after init
        {
            // init, kernels.py:43-49: channel c of the tile's images arrives through the stage
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int r = 0; r < S; ++r) M[h][r] = 0ull;
            for (int c = 0; c < p.C; ++c) {
#pragma unroll
                for (int ip = 0; ip < IMG_PARTS; ++ip) {
                    const unsigned buf = stage_l % NST;
                    mbar_wait(&full[buf], (stage_l / NST) & 1);
                    const float *sb = reinterpret_cast<const float *>(stage + buf * STAGE_F4) + lx;
                    const float *x0 = sb + (wi * 2 + 0) * IBAND, *x1 = sb + (wi * 2 + 1) * IBAND;
                    const float *z0 = sb + (kTileI + wj * 2 + 0) * IBAND, *z1 = sb + (kTileI + wj * 2 + 1) * IBAND;
                    constexpr int R = S / IMG_PARTS;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        const int r = ip * R + rr;
                        const float a0 = x0[rr * S], a1 = x1[rr * S], b0 = z0[rr * S], b1 = z1[rr * S];
                        const u64 A = pk(a0, a1);
                        M[0][r] = fma2(A, pk(b0, b1), M[0][r]);
                        M[1][r] = fma2(A, pk(b1, b0), M[1][r]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[buf]);
                    ++stage_l;
                }
            }
            if (p.C > 1) {
                const u64 IC = pk(p.inv_c, p.inv_c);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int r = 0; r < S; ++r) M[h][r] = mul2(M[h][r], IC);
            }
        }

        for (int k = 0; k < p.n_ops; ++k) {
            const FOp o = p.ops[k];
            if (o.kind == F_CONV) {
                const bool window = o.lo != 0 || o.hi != 0;
                if (window) {
                    // pass 1, transposition, pass 2 -- software-pipelined over the two packed
                    // arrays so that the smem traffic of one overlaps the adds of the other.
                    // The maps are kept divided by the product of the conv taps so far (the host
                    // scales each ReLU's variance maps to match: the arccos kernel is homogeneous),
                    // so "* tap + bias" is only "+ bias / taps", and a constant added to the two
                    // starting windows of the second sliding sum is in every output: two adds per
                    // map instead of one FMA per pixel.
                    const u64 PB = pk(o.pre_bias, o.pre_bias);
                    box_any<S, LO, HI, false>(M[0], o.lo, o.hi, 0ull);
                    tile_store<S>(tile, M[0], lane);
                    __syncwarp();
                    tile_load_t<S>(tile, M[0], lx);
                    box_any<S, LO, HI, false>(M[1], o.lo, o.hi, 0ull);
                    __syncwarp();
                    tile_store<S>(tile, M[1], lane);
                    box_any<S, LO, HI, true>(M[0], o.lo, o.hi, PB);
                    __syncwarp();
                    tile_load_t<S>(tile, M[1], lx);
                    __syncwarp();
                    box_any<S, LO, HI, true>(M[1], o.lo, o.hi, PB);
                }
                if (!window || o.scale != 1.f) {  // pointwise convs, and the host's rare explicit rescale
                    const u64 SC = pk(o.scale, o.scale), BI = pk(o.bias, o.bias);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int r = 0; r < S; ++r) M[h][r] = fma2(M[h][r], SC, BI);
                }
            } else if (o.kind == F_RELU) {
#pragma unroll
                for (int q = 0; q < NSPLIT; ++q) {
                    const unsigned buf = stage_l % NST;
                    mbar_wait(&full[buf], (stage_l / NST) & 1);
                    const float4 *sb = stage + buf * STAGE_F4 + lx;
                    const float4 *ai = sb + wi * BAND, *bj = sb + (kTileI / 2 + wj) * BAND;
                    constexpr int R = S / NSPLIT;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        const int r = q * R + rr;
                        const float4 A = ai[rr * S], B = bj[rr * S];
                        const u64 SA = pk(A.x, A.y), RA = pk(A.z, A.w);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const u64 SB = h ? pk(B.y, B.x) : pk(B.x, B.y);
                            const u64 RB = h ? pk(B.w, B.z) : pk(B.z, B.w);
                            float c0, c1;
                            upk(M[h][r], c0, c1);
                            const u64 NC = pk(neg_abs(c0), neg_abs(c1));
                            const u64 D = fma2(SA, SB, NC);               // s - |c|
                            const u64 E = fma2(NC, mul2(RA, RB), ONE);    // e = 1 - |c|/s, independent of D
                            float e0, e1;
                            upk(E, e0, e1);
                            const u64 W = mul2(D, pk(sqrt_approx(fabsf(e0)), sqrt_approx(fabsf(e1))));
                            u64 H = fma2(C5, E, C4);
                            H = fma2(H, E, C3);
                            H = fma2(H, E, C2);
                            H = fma2(H, E, C1);
                            H = fma2(H, E, C0);
                            M[h][r] = fma2(W, H, pk(fmaxf(c0, 0.f), fmaxf(c1, 0.f)));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[buf]);
                    ++stage_l;
                }
            } else {  // F_DENSE: whole-map sum, scale, bias -> the kernel entry
                float tot[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    u64 acc = M[h][0];
#pragma unroll
                    for (int r = 1; r < S; ++r) acc = add2(acc, M[h][r]);
                    float a0, a1;
                    upk(acc, a0, a1);
                    if (lane >= S) { a0 = 0.f; a1 = 0.f; }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, d);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, d);
                    }
                    // M[0] = (i0j0, i1j1), M[1] = (i0j1, i1j0); tot index = 2a + b
                    tot[h == 0 ? 0 : 1] = fmaf(a0, o.scale, o.bias);
                    tot[h == 0 ? 3 : 2] = fmaf(a1, o.scale, o.bias);
                }
                if (lane < 4) {
                    const int a = lane >> 1, b = lane & 1;
                    const int i = i_base + wi * 2 + a, j = j_base + wj * 2 + b;
                    const float v = lane == 0 ? tot[0] : lane == 1 ? tot[1] : lane == 2 ? tot[2] : tot[3];
                    if (i < p.N1 && j < p.N2) {
                        if (!p.symmetric) {
                            p.out[(long long)i * p.ld_out + j] = v;
                        } else if (j > i) {
                            p.out[(long long)i * p.ld_out + j] = v;
                            // a band of block rows (N2 > N1): the mirror image stays inside the diagonal block, as in
                            // the reference's same=True tile; rows below the band belong to other launches
                            if (j < p.N1 && (p.mirror_bs == 0 || j / p.mirror_bs == i / p.mirror_bs))
                                p.out[(long long)j * p.ld_out + i] = v;
                        } else if (j == i) {
                            // i == j follows the variance recursion (kernels.py:155-162)
                            p.out[(long long)i * p.ld_out + i] = p.kdiag ? p.kdiag[i] : v;
                        }
                    }
                }
                // this warp's entries of its k-th tile are written (release at CTA scope; the producer lane reports)
                // (the tile count lives in shared memory: the consumers have no register to spare across a tile)
                if (PROG) {
                    __syncwarp();
                    if (lane == 0) {
                        unsigned *kw = reinterpret_cast<unsigned *>(done + (NG - grp) * 2) + warp;
                        const unsigned k = *kw;
                        mbar_arrive(&done[k & 1]);
                        *kw = k + 1;
                    }
                }
            }
        }
    }
}

template <int NW, int NSPLIT, int NST, int NG = 1>
constexpr size_t fused_smem(int S) {
    return (size_t)NG * NST * (Geo<NW, NG>::kPairs * S * S / NSPLIT) * 16 + (size_t)NW * S * (S + 1) * 8 + (size_t)NG * 3 * NST * 8 +
           (size_t)NG * 2 * 8 + (size_t)NW * 4;  // + the two tile-done barriers of every group, a tile count per consumer warp
}

}  // namespace

struct FusedPlan {
    int S = 0;
    int lo = -1, hi = -1;  // the single non-trivial window shape of the program, or -1 if mixed
    int n_ops = 0, n_relu = 0;
    FOp ops[kMaxOps];
    size_t smem = 0;
};

// Decide whether the program is in the fused kernel's set and, if so, translate it.  Also marks
// which ReLU layers consume their variance maps in transposed (lane = row) layout.
FusedPlan *fused_plan_create(const Plan *plan_const) {
    Plan *plan = const_cast<Plan *>(plan_const);
    if (plan->dtype != CNNGP_F32) return nullptr;
    if (plan->H != plan->W || plan->H != 28) return nullptr;
    const int S = plan->H;
    FusedPlan fp;
    fp.S = S;
    int cur = 0;             // the slot the straight-line program lives in
    bool transposed = false;  // current register layout: lane = row?
    double rho = 1.0;        // true map = rho * the kernel's registers: conv taps and the 1/2 of the ReLU's
                             // doubled output are never applied to the maps, only carried here
    bool after_relu = false;
    const bool fold = getenv("CNNGP_NO_FOLD") == nullptr;  // measurement aid: one explicit FMA pass per conv
    bool done = false;
    int n_windows = 0;
    for (size_t k = 0; k < plan->ops.size(); ++k) {
        DevOp &o = plan->ops[k];
        if (done || fp.n_ops >= kMaxOps) return nullptr;
        if (o.src != cur) return nullptr;
        FOp f{};
        if (o.opcode == CNNGP_OP_CONV) {
            if (o.dil != 1 || o.stride != 1) return nullptr;
            if (o.Hi != S || o.Wi != S) return nullptr;
            const int lo = o.pad - o.t0, hi = o.ke - 1 - o.pad;
            if (o.Ho == S && o.Wo == S) {
                if (lo < 0 || hi < 0) return nullptr;
                const bool known = (lo == 0 && hi == 0) || (lo == 3 && hi == 3) || (lo == 1 && hi == 1) ||
                                   (lo == 1 && hi == 2) || (lo == 2 && hi == 2);
                if (!known) return nullptr;
                f.kind = F_CONV; f.lo = lo; f.hi = hi;
                if (lo || hi) {
                    transposed = !transposed;
                    if (n_windows == 0) { fp.lo = lo; fp.hi = hi; }
                    else if (fp.lo != lo || fp.hi != hi) { fp.lo = -1; fp.hi = -1; }
                    ++n_windows;
                }
            } else if (o.Ho == 1 && o.Wo == 1 && o.pad == 0 && o.t0 == 0 && o.ke == S) {
                f.kind = F_DENSE;
                done = true;
            } else {
                return nullptr;
            }
            const double alpha = o.scale_d * rho;  // true output = alpha * (box(registers) + bias / alpha)
            if (f.kind == F_DENSE) {
                f.scale = (float)alpha; f.bias = (float)o.bias_d;
            } else if (fold && alpha > 1e-30 && alpha < 1e30 && alpha == alpha) {
                const float b = (float)(o.bias_d / alpha);
                if (f.lo || f.hi) { f.pre_bias = b; f.scale = 1.f; f.bias = 0.f; }
                else { f.pre_bias = 0.f; f.scale = 1.f; f.bias = b; }
                rho = alpha;
            } else {  // the carried factor would leave float range: apply it once, explicitly
                f.pre_bias = 0.f; f.scale = (float)alpha; f.bias = (float)o.bias_d;
                rho = 1.0;
            }
            after_relu = false;
        } else if (o.opcode == CNNGP_OP_RELU) {
            if (after_relu) return nullptr;  // ReLU directly after ReLU: not in the set
            if (o.Hi != S || o.Wi != S) return nullptr;
            f.kind = F_RELU;
            f.aux_off = o.aux_foff;
            o.aux_t = transposed ? 1 : 0;
            // relu_k(rho * m; s) = rho * relu_k(m; s / rho): each image's s map carries 1 / sqrt(rho)
            o.aux_scale = (float)(1.0 / std::sqrt(rho));
            rho *= 0.5;  // the kernel's ReLU output is doubled
            after_relu = true;
            ++fp.n_relu;
        } else {
            return nullptr;
        }
        cur = o.dst;
        fp.ops[fp.n_ops++] = f;
    }
    if (!done) {
        for (DevOp &o : plan->ops) { o.aux_t = 0; o.aux_scale = 1.f; }
        return nullptr;
    }
    return new FusedPlan(fp);
}

void fused_plan_destroy(FusedPlan *fp) { delete fp; }

std::string fused_plan_describe(const FusedPlan *fp) {
    std::string t = "fused S=" + std::to_string(fp->S) + " window=" + std::to_string(fp->lo) + "," + std::to_string(fp->hi) + " :";
    for (int k = 0; k < fp->n_ops; ++k) {
        const FOp &o = fp->ops[k];
        if (o.kind == F_CONV) t += " CONV(" + std::to_string(o.lo) + "," + std::to_string(o.hi) + ")";
        else if (o.kind == F_RELU) t += " RELU";
        else t += " DENSE";
    }
    return t;
}

// one line per op, every field; RELU lines carry the factor the host put on that layer's s maps
std::string fused_plan_dump(const FusedPlan *fp, const Plan *plan) {
    std::string t = "fused S=" + std::to_string(fp->S) + "\n";
    char line[256];
    std::vector<const DevOp *> relus;
    for (const DevOp &o : plan->ops)
        if (o.opcode == CNNGP_OP_RELU) relus.push_back(&o);
    size_t r = 0;
    for (int k = 0; k < fp->n_ops; ++k) {
        const FOp &o = fp->ops[k];
        if (o.kind == F_RELU) {
            snprintf(line, sizeof line, "RELU aux=%d aux_scale=%.9g transposed=%d\n", o.aux_off,
                     (double)relus[r]->aux_scale, relus[r]->aux_t);
            ++r;
        } else {
            snprintf(line, sizeof line, "%s lo=%d hi=%d pre_bias=%.9g scale=%.9g bias=%.9g\n", o.kind == F_CONV ? "CONV" : "DENSE",
                     o.lo, o.hi, (double)o.pre_bias, (double)o.scale, (double)o.bias);
        }
        t += line;
    }
    return t;
}

namespace {

struct Variant { int nw, nsplit, nst; };

template <int NW, int NSPLIT, int NST, int NG = 1>
int launch_variant(const FusedPlan *fp, FParams &p, int64_t N1, int64_t N2, cudaStream_t stream, RowProgress *prog) {
    using G = Geo<NW, NG>;
    p.nbi = (int)((N1 + G::kTileI - 1) / G::kTileI);
    p.nbj = (int)((N2 + G::kTileJ - 1) / G::kTileJ);
    long long n_super;
    int edge = kSuperEdge;
    if (const char *e = getenv("CNNGP_SUPER_EDGE")) { const int v = atoi(e); if (v >= 24 && v % 24 == 0) edge = v; }
    const int super_i = edge / G::kTileI, super_j = edge / G::kTileJ;
    if (p.nbi <= super_i && p.nbj <= super_j) {  // one (possibly small) super-tile
        p.sti = p.nbi; p.stj = p.nbj; p.nst_j = 1; p.nst = 1;
        n_super = 1;
    } else {
        p.sti = super_i; p.stj = super_j;
        const int nsi = (p.nbi + super_i - 1) / super_i, nsj = (p.nbj + super_j - 1) / super_j;
        p.nst_j = nsj;
        // symmetric: super-row r holds the super-tiles (r, r .. nsj - 1); N2 > N1 (a band of block rows) has
        // fewer super-rows than super-columns
        p.nst = p.symmetric ? nsj : (nsi > nsj ? nsi : nsj);
        n_super = p.symmetric ? (long long)nsi * nsj - (long long)nsi * (nsi - 1) / 2 : (long long)nsi * nsj;
    }
    p.n_tiles = n_super * p.sti * p.stj;
    if (prog) {  // what the kernel will count per super-row: valid tiles x consumer warps (decode() below)
        prog->n_super_rows = (p.nbi + p.sti - 1) / p.sti;
        prog->rows_per_super = (int64_t)p.sti * G::kTileI;
        if (!p.symmetric || N1 != N2 || prog->n_super_rows > prog->capacity) { set_error("fused kernel: progress counters too few"); return 8; }
        prog->expected.assign(prog->n_super_rows, 0u);
        for (int ib = 0; ib < p.nbi; ++ib) {
            const int si = ib / p.sti;
            // jb >= si * stj (upper-triangular super-tiles) and the tile reaches the diagonal or beyond
            long long jb_lo = (long long)si * p.stj;
            const long long need = ((long long)ib * G::kTileI - (G::kTileJ - 1) + G::kTileJ - 1) / G::kTileJ;
            if (need > jb_lo) jb_lo = need;
            if (jb_lo < p.nbj) prog->expected[si] += (unsigned)((p.nbj - jb_lo) * G::kGW);
        }
        p.row_done = prog->d_done;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);  // persistent: one CTA per SM
    void (*kern)(const FParams) = nullptr;
    constexpr bool kDefault = NW == 12 && NSPLIT == 2 && NST == 3 && NG == 1;
    if (prog) {  // the reporting instantiations exist for the default variant only
        if (!kDefault) return -1;
        if (fp->lo == 3 && fp->hi == 3) kern = fused_kernel<28, 3, 3, 12, 2, 3, 1, true>;
        else if (fp->lo == 1 && fp->hi == 1) kern = fused_kernel<28, 1, 1, 12, 2, 3, 1, true>;
        else if (fp->lo == 1 && fp->hi == 2) kern = fused_kernel<28, 1, 2, 12, 2, 3, 1, true>;
        else if (fp->lo == 2 && fp->hi == 2) kern = fused_kernel<28, 2, 2, 12, 2, 3, 1, true>;
        else kern = fused_kernel<28, -1, -1, 12, 2, 3, 1, true>;
    }
    else if (fp->lo == 3 && fp->hi == 3) kern = fused_kernel<28, 3, 3, NW, NSPLIT, NST, NG>;
    else if (kDefault && fp->lo == 1 && fp->hi == 1) kern = fused_kernel<28, 1, 1, 12, 2, 3>;
    else if (kDefault && fp->lo == 1 && fp->hi == 2) kern = fused_kernel<28, 1, 2, 12, 2, 3>;
    else if (kDefault && fp->lo == 2 && fp->hi == 2) kern = fused_kernel<28, 2, 2, 12, 2, 3>;
    else if (kDefault) kern = fused_kernel<28, -1, -1, 12, 2, 3>;
    else return -1;  // experimental variants exist for the 7x7 window only: caller falls back to the default
    const size_t smem = fused_smem<NW, NSPLIT, NST, NG>(28);
    cudaError_t e = cudaSuccess;
    const char *order = getenv("CNNGP_TILE_ORDER");  // "static": fixed stride instead of the counter
    if (order && !strcmp(order, "static")) {
        p.tile_ctr = nullptr;
    } else {
        p.tile_ctr = tile_counter_for(stream);
        if (!p.tile_ctr) return 7;
        e = cudaMemsetAsync(p.tile_ctr, 0, sizeof(unsigned long long), stream);
    }
    if (e != cudaSuccess) { set_error(std::string("fused tile counter: ") + cudaGetErrorString(e)); return 7; }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error(std::string("fused cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 7; }
    kern<<<grid, G::kThreads, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("fused kernel launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

}  // namespace

int launch_fused_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
                      int32_t C, const void *d_aux_x, const void *d_aux_z, int32_t same, int32_t diag,
                      int32_t symmetric, const void *d_kdiag, void *d_out, int64_t ld_out, void *stream,
                      RowProgress *progress, int64_t mirror_block) {
    (void)same;
    const FusedPlan *fp = plan->fused;
    if (!fp || diag) { set_error("fused kernel: unsupported call"); return 4; }
    if (N1 > 2000000000LL || N2 > 2000000000LL) { set_error("fused kernel: too many images"); return 8; }
    FParams p{};
    memcpy(p.ops, fp->ops, sizeof(FOp) * fp->n_ops);
    p.n_ops = fp->n_ops; p.n_relu = fp->n_relu;
    p.x = (const float *)d_x; p.z = (const float *)d_z;
    p.aux_x = (const float *)d_aux_x; p.aux_z = (const float *)d_aux_z;
    p.aux_stride = plan->aux_elems; p.aux_f_off = plan->aux_f_off;
    p.N1 = (int)N1; p.N2 = (int)N2; p.C = C;
    p.out = (float *)d_out; p.ld_out = ld_out;
    p.symmetric = symmetric ? 1 : 0;
    if (symmetric && N2 < N1) { set_error("fused kernel: a symmetric band needs N2 >= N1"); return 4; }
    p.mirror_bs = (int)mirror_block;
    p.kdiag = (const float *)d_kdiag;
    p.inv_c = 1.0f / (float)C;
    // kernel variant: consumer warps, ReLU bands per layer, ring depth.  Default 12 warps (three
    // per SM sub-partition at 160 registers), two bands per layer, three stages (one and a half
    // layers of variance maps in flight: 247.6 M pairs/s against 237.7 M with two stages on the same
    // box; 228 M for the 8-warp / 240-register variant).
    // CNNGP_FUSED_VARIANT=nw,nsplit,nst selects the others (7x7 window only) for comparison.
    Variant v{12, 2, 3};
    if (const char *e = getenv("CNNGP_FUSED_VARIANT")) sscanf(e, "%d,%d,%d", &v.nw, &v.nsplit, &v.nst);
    int rc = -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (v.nw == 12 && v.nsplit == 4 && v.nst == 4) rc = launch_variant<12, 4, 4>(fp, p, N1, N2, st, progress);
    else if (v.nw == 8 && v.nsplit == 1 && v.nst == 2) rc = launch_variant<8, 1, 2>(fp, p, N1, N2, st, progress);
    else if (v.nw == 8 && v.nsplit == 2 && v.nst == 4) rc = launch_variant<8, 2, 4>(fp, p, N1, N2, st, progress);
    else if (v.nw == 12 && v.nsplit == 2 && v.nst == 3) rc = launch_variant<12, 2, 3>(fp, p, N1, N2, st, progress);
    else if (v.nw == 12 && v.nsplit == 2 && v.nst == 2) rc = launch_variant<12, 2, 2>(fp, p, N1, N2, st, progress);
    else if (v.nw == 122) rc = launch_variant<12, 2, 2, 2>(fp, p, N1, N2, st, progress);  // two independent groups of six warps
    if (rc < 0) rc = launch_variant<12, 2, 3>(fp, p, N1, N2, st, progress);
    return rc;
}

}  // namespace cnngp
