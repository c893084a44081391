// gram_fnet.cu -- register-resident fused Gram kernel for layer programs with Sum (residual)
// branches, strided convolutions and several map sizes: the ResNet GPs of configs/mnist.py,
// mnist_as_tf.py, cifar10.py, the residual CNN GP and the README model.
//
// Same execution model as gram_fused.cu: a persistent CTA per SM, consumer warps that each keep the
// four covariance maps of a 2 x 2 block of image pairs in registers (lane = one map coordinate,
// register index = the other, two maps per packed f32x2 register), a producer warp that stages
// images and per-layer (s, 1/s) variance maps with bulk async copies, box convolutions as sliding
// sums along the register axis + a shared-memory transposition.
//
// What is specific to this kernel:
//   * a two-slot program needs a second live map per pair (the skip connection of a Sum,
//     reference cnn_gp/kernels.py:246-254).  It does not fit the register file next to the
//     working map, and shared memory is taken by the staging ring -- so it is stashed in TENSOR
//     MEMORY: each warp owns a window of its 32-lane TMEM quadrant and moves a whole map set with
//     tcgen05.st / tcgen05.ld (SASS STTM / LDTM), off the shared-memory port.
//   * stride-2 convolutions (kernels.py:92-98 with stride 2) subsample along the register axis
//     in both passes; maps shrink to S/2 and S/4.  Maps of edge <= S/2 are FOLDED: the second
//     packed array moves into lanes 16..31 of the first, so every later op runs on one array.
//   * the op loop has three phases with their own register budget: phase A (full-size maps, the
//     whole M[2][S0] set is live), phase B (folded maps: S0/2 registers are live, the rest of the
//     register file is free for instruction-level parallelism), and the scalar tail after the
//     global pooling convolution (1 x 1 maps).
//   * ops are 16-byte descriptors in shared memory (one LDS.128 per op, fetched one op ahead) and
//     whole residual blocks are single dispatch cases (`IDBLOCK`: STASH RELU CONV RELU CONV ADD,
//     `RESBLOCK`: STASH CONV RELU TRANSPOSE ADD): the per-op dispatch of the first version of this
//     kernel (a 36-byte descriptor in the constant bank, ~40 instructions and two spilled
//     descriptors per op) was 27 % of the stall samples of an mnist_as_tf launch.
//   * conv taps and the doubled ReLU outputs are never applied to the maps: every slot carries the
//     factor it owes (`pend`), the host scales each ReLU layer's per-image (s, 1/s) maps to match
//     (the arccos kernel is positively homogeneous), a Sum folds the ratio of its operands'
//     factors into its FMA, and a conv bias rides on the two starting windows of the second
//     sliding sum (see gram_fused.cu).  An explicit scale pass is emitted only when the carried
//     factor would leave a safe range.
//   * every windowed convolution flips the register layout (lane = column <-> lane = row); the
//     translator tracks the layout of every slot and inserts a transposition where a Sum would
//     add maps of different layouts, and records for every ReLU the layout its variance maps
//     must be stored in.
//   * the tail after the global pooling convolution (1 x 1 maps: kernels.py:134-165 on one
//     pixel, 1 x 1 convolutions) runs on the four scalars of the warp.
#include <cuda_runtime.h>

#include <atomic>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "fused_common.cuh"
#include "plan.h"

namespace cnngp {

namespace {
using namespace fusedk;

// NW consumer warps as 2 x NW/2 warps of 2 x 2 image pairs, plus a warpgroup holding the producer
template <int NW>
struct NGeo {
    static constexpr int kWarps = NW;
    static constexpr int kTileI = 4, kTileJ = NW;
    static constexpr int kImgs = kTileI + kTileJ;
    static constexpr int kPairs = kImgs / 2;
    static constexpr int kThreads = (NW + 4) * 32;
    static constexpr int kRegsProducer = NW == 12 ? 32 : 24;  // what the pool holds: launch registers x threads
    // setmaxnreg moves registers inside the CTA's launch allocation (ptxas' count x threads): 16 + 4 warps are
    // launched with 96 registers, 20 x 32 x 96 = 16 x 32 x 112 + 4 x 32 x 24 (asking for more blocks forever)
    static constexpr int kRegsConsumer = NW == 8 ? 240 : (NW == 12 ? 160 : 112);
};
constexpr int kMaxNOps = 192;   // register-level ops of a translated program
constexpr int kMaxKOps = 224;   // descriptors the kernel sees (ops + phase sentinels)
constexpr int kMaxRelu = 96;
constexpr int kTmemCols = 512;

// register-level op kinds of the translator
enum { N_CONV = 0, N_AFFINE, N_RELU, N_STASH, N_UNSTASH, N_ADD, N_TRANSPOSE, N_DENSE, T_RELU, T_AFFINE };

// dispatch cases.  Phase A works on the full register set M[2][S0]: kind x size class (0: S0,
// 1: S0/2, 2: S0/4; folded maps live in M[0]), convolutions by variant.  Phase B works on the
// folded array F[S0/2] (size classes 1 and 2 only).  The tail works on the four scalars.
enum { A_CONV = 0 /* + 10: 3 x 3 with dilation 2 at full size */, A_AFFINE = 11, A_TRANSPOSE = 14, A_STASH = 17, A_UNSTASH = 20,
       A_ADD = 23, A_DENSE = 26, A_RELU = 29, A_IDBLOCK = 32, A_RESBLOCK = 33, A_END = 34, A_CASES = 35 };
enum { B_CONV = 0 /* +0: S/2 s1, +1: S/2 -> S/4 k3 s2, +2: S/2 -> S/4 k1 s2, +3: S/4 s1 */, B_AFFINE = 4, B_TRANSPOSE = 6,
       B_STASH = 8, B_UNSTASH = 10, B_ADD = 12, B_DENSE = 14, B_RELU = 16, B_IDBLOCK = 18, B_END = 20, B_CASES = 21 };
enum { T_CASE_AFFINE = 0, T_CASE_RELU = 1, T_END = 2 };

struct NOp {
    int kind;
    short si, so;       // map edge before / after the op
    short lo, hi, st;   // N_CONV: window offsets [-lo, +hi] (in taps) and stride
    short dil;          // N_CONV: distance between taps (reference Conv2d(dilation=), kernels.py:61,95-96)
    short slot;         // N_STASH / N_UNSTASH / N_ADD: tensor-memory slot (0 or 1)
    float scale, bias;  // N_CONV / N_AFFINE / N_DENSE / T_AFFINE: explicit scale-and-bias pass (1, 0: none);
                        // N_ADD: factor applied to the stashed map, constant added
    float pre_bias;     // N_CONV (stride 1, windowed): constant carried by the second sliding sum
    float aux_scale;    // N_RELU: factor the host puts on this layer's per-image s maps (dump only)
    int aux;            // N_RELU: float offset inside the fused section; T_RELU: float offset of xx in the row
    int half;           // N_RELU: pixels held by the first row of the pair
};

// what the kernel reads: 16 bytes per op in shared memory.  code: bits 0..7 dispatch case, 8..9
// tensor-memory slot, 16..31 `half`.  aux: N_RELU / T_RELU offset; N_CONV: pre_bias (float bits).
struct __align__(16) KOp { int code; float scale; float bias; int aux; };

struct NParams {
    KOp ops[kMaxKOps];
    int relu_aux[kMaxRelu];   // producer's list of staged ReLU layers, program order: offset in the fused section
    int relu_half[kMaxRelu];  //   pixels in the first row of the pair; bit 30: full-size layer (NSPLIT bands)
    int n_ops, n_relu;
    const float *x, *z;
    const float *aux_x, *aux_z;
    long long aux_stride;
    int aux_f_off;
    int N1, N2, C;
    float *out;
    long long ld_out;
    int symmetric;
    int mirror_bs;  // symmetric with N2 > N1 (a band of block rows): mirror only inside diagonal blocks of this many rows
    const float *kdiag;
    int nbi, nbj, sti, stj, nst_j, nst;
    long long n_tiles;
    unsigned long long *tile_ctr;  // zeroed before the launch: the next tile index to hand out
    // split launches (phase A and phase B as two kernels, see fnet_kernel): this launch covers tiles
    // [t_begin, n_tiles); phase A leaves every 2 x 2 block's folded map and tensor-memory slot 1 in
    // `handoff` (one record of 2 x S0/2 x 32 packed entries per block, `rec_per_st` records per
    // super-tile, `rec_ld` per block row; super-tile `st_begin` is the first of the buffer)
    long long t_begin;
    unsigned long long *handoff;
    int st_begin, rec_ld, rec_per_st;
    int n_relu_a;  // staged ReLU layers that belong to phase A
    int k_b;       // first descriptor of phase B
    unsigned *row_done;            // optional: finished (tile, warp) units per super-row (RowProgress, plan.h)
    float inv_c;
};

// ---- tensor memory ---------------------------------------------------------------------------
// tcgen05.st / tcgen05.ld move N consecutive 32-bit registers of every thread to / from N
// consecutive columns of the thread's own TMEM lane.  The b32 halves of the packed maps are passed
// as the instruction's own operands (no staging copies); a load is only complete after
// tcgen05.wait::ld, which is tied to the loaded registers through "+r" operands so that the
// compiler cannot move their first use above it.
__device__ __forceinline__ void split64(u64 v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ u64 join64(uint32_t lo, uint32_t hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, const uint32_t (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_landed4(uint32_t (&r)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}
__device__ __forceinline__ void tmem_landed2(uint32_t (&r)[2]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]) : : "memory");
}
__device__ __forceinline__ void tmem_landed16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_landed8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one packed array (first S entries of a[NA]) -> S * 2 consecutive tensor-memory columns at `ta`, one
// tcgen05.st.x2 per entry: a packed map entry IS an aligned register pair, so nothing has to be
// marshalled (the x16 form wants sixteen consecutive registers: 112 register moves per full-size
// stash, which the profile showed as 15 % of all executed instructions).  The caller waits (tmem_wait_st).
template <int NA, int S>
__device__ __forceinline__ void stash_store_arr(uint32_t ta, const u64 (&a)[NA]) {
#pragma unroll
    for (int q = 0; q < S; ++q) {
        uint32_t r[2];
        split64(a[q], r[0], r[1]);
        tmem_st2(ta + 2 * q, r);
    }
}

// a = stash (ADD == false) or a = stash * alpha + a (ADD == true); the loads of the array are in
// flight together, two 16-register groups at a time (LIGHT: the 12-warp variant runs phase A at 160
// registers with 112 of them holding the maps) or all of them (phase B: registers are plentiful)
template <int NA, int S, bool ADD, bool ONE_GROUP = false>
__device__ __forceinline__ void stash_load_arr(uint32_t ta, u64 (&a)[NA], u64 alpha) {
    if (!ADD) {  // plain reload: x2 loads land in the map's own register pairs, all in flight, one wait
        uint32_t r[S][2];
#pragma unroll
        for (int q = 0; q < S; ++q) tmem_ld2(ta + 2 * q, r[q]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < S; ++q) {
            asm volatile("" : "+r"(r[q][0]), "+r"(r[q][1]));  // first use stays below the wait
            a[q] = join64(r[q][0], r[q][1]);
        }
        return;
    }
    constexpr int G8 = S / 8, R8 = S % 8, B4 = G8 * 8, B2 = B4 + (R8 & 4), B1 = B2 + (R8 & 2);
    if (ONE_GROUP) {
        // one 16-register group at a time: with twelve warps reading (three per lane quadrant) the loads of the
        // other warps keep tensor memory busy anyway (scripts/tmem_probe.cu: 7 x (ld16 + wait) = 205 cycles
        // against 203 with all seven in flight), and the 40 staging registers of the deeper form do not
        // exist next to a full-size map set in a 160-register warp
#pragma unroll
        for (int c = 0; c < G8; ++c) {
            uint32_t t[16];
            tmem_ld16(ta + c * 16, t);
            tmem_landed16(t);
#pragma unroll
            for (int q = 0; q < 8; ++q) a[c * 8 + q] = fma2(join64(t[2 * q], t[2 * q + 1]), alpha, a[c * 8 + q]);
        }
        if (R8 & 4) {
            uint32_t t[8];
            tmem_ld8(ta + 2 * B4, t);
            tmem_landed8(t);
#pragma unroll
            for (int q = 0; q < 4; ++q) a[(B4 + q) % NA] = fma2(join64(t[2 * q], t[2 * q + 1]), alpha, a[(B4 + q) % NA]);
        }
        if (R8 & 2) {
            uint32_t t[4];
            tmem_ld4(ta + 2 * B2, t);
            tmem_landed4(t);
#pragma unroll
            for (int q = 0; q < 2; ++q) a[(B2 + q) % NA] = fma2(join64(t[2 * q], t[2 * q + 1]), alpha, a[(B2 + q) % NA]);
        }
        if (R8 & 1) {
            uint32_t t[2];
            tmem_ld2(ta + 2 * B1, t);
            tmem_landed2(t);
            a[B1 % NA] = fma2(join64(t[0], t[1]), alpha, a[B1 % NA]);
        }
        return;
    }
    uint32_t t8[8], t4[4], t2[2];
    if (R8 & 4) tmem_ld8(ta + 2 * B4, t8);
    if (R8 & 2) tmem_ld4(ta + 2 * B2, t4);
    if (R8 & 1) tmem_ld2(ta + 2 * B1, t2);
#pragma unroll
    for (int c = 0; c < G8; c += 2) {
        uint32_t ta0[16], ta1[16];
        tmem_ld16(ta + c * 16, ta0);
        if (c + 1 < G8) tmem_ld16(ta + (c + 1) * 16, ta1);
        tmem_landed16(ta0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const u64 v = join64(ta0[2 * q], ta0[2 * q + 1]);
            a[c * 8 + q] = ADD ? fma2(v, alpha, a[c * 8 + q]) : v;
        }
        if (c + 1 < G8) {
            tmem_landed16(ta1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const u64 v = join64(ta1[2 * q], ta1[2 * q + 1]);
                a[(c + 1) * 8 + q] = ADD ? fma2(v, alpha, a[(c + 1) * 8 + q]) : v;
            }
        }
    }
    if (R8 & 4) {
        tmem_landed8(t8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u64 v = join64(t8[2 * q], t8[2 * q + 1]);
            a[(B4 + q) % NA] = ADD ? fma2(v, alpha, a[(B4 + q) % NA]) : v;
        }
    }
    if (R8 & 2) {
        tmem_landed4(t4);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const u64 v = join64(t4[2 * q], t4[2 * q + 1]);
            a[(B2 + q) % NA] = ADD ? fma2(v, alpha, a[(B2 + q) % NA]) : v;
        }
    }
    if (R8 & 1) {
        tmem_landed2(t2);
        const u64 v = join64(t2[0], t2[1]);
        a[B1 % NA] = ADD ? fma2(v, alpha, a[B1 % NA]) : v;
    }
}

template <int NA, int S>
__device__ __forceinline__ void add_const(u64 (&a)[NA], float beta_f) {
    if (beta_f != 0.f) {  // aliased 1 x 1 convolution with a bias (uniform branch)
        const u64 beta = pk(beta_f, beta_f);
#pragma unroll
        for (int r = 0; r < S; ++r) a[r] = add2(a[r], beta);
    }
}

// ---- box sums along the register axis --------------------------------------------------------
// stride 1, zero padding: out[y] = B + sum_{t=-LO..HI} v[y+t] on the first S entries, two sliding
// windows from both ends (as gram_fused.cu); BIAS: both windows start from v + B, so every output
// carries the constant B (the folded conv bias)
template <int NA, int S, int LO, int HI, bool BIAS>
__device__ __forceinline__ void box_s1(u64 (&v)[NA], u64 B) {
    if (LO == 0 && HI == 0) return;
    if (LO == 1 && HI == 1) {
        // 3 taps: the direct sum costs the same two adds per output as a sliding window, but every output is
        // independent (no 14-step dependency chain from either end) and each input dies two outputs after it
        // is first used (fewer register moves: the sliding form keeps the leaving element alive)
        u64 prev = v[0];
        v[0] = add2(v[0], v[1]);
#pragma unroll
        for (int y = 1; y < S; ++y) {
            const u64 cur = v[y];
            const u64 two = add2(prev, cur);
            v[y] = y + 1 < S ? add2(two, v[y + 1]) : two;
            prev = cur;
        }
        if (BIAS) {
            float b0, b1;
            upk(B, b0, b1);
            if (b0 != 0.f) {  // uniform: the ResNet GPs have no bias
#pragma unroll
                for (int y = 0; y < S; ++y) v[y] = add2(v[y], B);
            }
        }
        return;
    }
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

// dilated taps, stride 1, zero padding: out[y] = B + sum_{t=-LO..HI} v[y + t DIL] -- a direct sum (the taps of
// neighbouring outputs do not overlap, so there is nothing to slide)
template <int NA, int S, int LO, int HI, int DIL, bool BIAS>
__device__ __forceinline__ void box_dil(u64 (&v)[NA], u64 B) {
    u64 o[S];
#pragma unroll
    for (int y = 0; y < S; ++y) {
        o[y] = BIAS ? add2(v[y], B) : v[y];
#pragma unroll
        for (int t = -LO; t <= HI; ++t) {
            const int idx = y + t * DIL;
            if (t != 0 && idx >= 0 && idx < S) o[y] = add2(o[y], v[idx]);
        }
    }
#pragma unroll
    for (int y = 0; y < S; ++y) v[y] = o[y];
}

// stride 2: out[y] = sum_{t=-LO..HI} v[2y+t], SI entries -> SO entries
template <int NA, int SI, int SO, int LO, int HI>
__device__ __forceinline__ void box_s2(u64 (&v)[NA]) {
    u64 o[SO];
#pragma unroll
    for (int y = 0; y < SO; ++y) {
        bool first = true;
#pragma unroll
        for (int t = -LO; t <= HI; ++t) {
            const int idx = 2 * y + t;
            if (idx >= 0 && idx < SI) {
                o[y] = first ? v[idx] : add2(o[y], v[idx]);
                first = false;
            }
        }
    }
#pragma unroll
    for (int y = 0; y < SO; ++y) v[y] = o[y];
}

template <int S0, int R, int L>
__device__ __forceinline__ void tstore(u64 *tile, const u64 (&a)[S0], int lane) {
    constexpr int PITCH = S0 + 1;  // odd: row-wise writes and column-wise reads are both conflict-free
    // lanes >= L hold nothing: they all write the pad column S0, which is never read (no branch
    // around the stores, so the compiler can sink them into the sums that produce the values)
    const int col = lane < L ? lane : S0;
#pragma unroll
    for (int r = 0; r < R; ++r) tile[r * PITCH + col] = a[r];
}
template <int S0, int R>
__device__ __forceinline__ void tload(const u64 *tile, u64 (&a)[S0], int lx) {
    constexpr int PITCH = S0 + 1;
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = tile[lx * PITCH + r];
}

template <int NA, int S>
__device__ __forceinline__ void affine_arr(u64 (&a)[NA], float scale, float bias) {
    const u64 SC = pk(scale, scale), BI = pk(bias, bias);
#pragma unroll
    for (int r = 0; r < S; ++r) a[r] = fma2(a[r], SC, BI);
}

// box convolution SI x SI -> SO x SO on the full register set: pass, transposition, pass (flips the
// register layout), software-pipelined over the two packed arrays.  Stride 1: the second pass
// carries `pre_bias`.  The explicit tap * sum + bias pass runs only when the translator asks for it
// (scale != 1 or bias != 0: strided convolutions with a bias, and the rare reset of the carried factor).
template <int S0, int SI, int SO, int LO, int HI, int ST, int DIL = 1>
__device__ __forceinline__ void conv_op(u64 (&M)[2][S0], u64 *tile, int lane, float pre_bias, float scale, float bias) {
    static_assert(ST == 1 ? SI == SO : SI == 2 * SO, "conv geometry");
    static_assert(DIL == 1 || ST == 1, "dilated windows are stride 1 here");
    const u64 PB = pk(pre_bias, pre_bias);
    auto pass1 = [](u64 (&v)[S0]) {
        if (DIL > 1) box_dil<S0, SI, LO, HI, DIL, false>(v, 0ull);
        else if (ST == 1) box_s1<S0, SI, LO, HI, false>(v, 0ull);
        else box_s2<S0, SI, SO, LO, HI>(v);
    };
    auto pass2 = [&](u64 (&v)[S0]) {
        if (DIL > 1) box_dil<S0, SI, LO, HI, DIL, true>(v, PB);
        else if (ST == 1) box_s1<S0, SI, LO, HI, true>(v, PB);
        else box_s2<S0, SI, SO, LO, HI>(v);
    };
    const int lx = lane < SO ? lane : SO - 1;
    pass1(M[0]);
    tstore<S0, SO, SI>(tile, M[0], lane);
    __syncwarp();
    tload<S0, SI>(tile, M[0], lx);
    pass1(M[1]);
    __syncwarp();
    tstore<S0, SO, SI>(tile, M[1], lane);
    pass2(M[0]);
    __syncwarp();
    tload<S0, SI>(tile, M[1], lx);
    __syncwarp();
    pass2(M[1]);
    if (scale != 1.f || bias != 0.f) {
        affine_arr<S0, SO>(M[0], scale, bias);
        affine_arr<S0, SO>(M[1], scale, bias);
    }
}

template <int S0, int S>
__device__ __forceinline__ void transpose_op(u64 (&M)[2][S0], u64 *tile, int lane) {
    const int lx = lane < S ? lane : S - 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        tstore<S0, S, S>(tile, M[h], lane);
        __syncwarp();
        tload<S0, S>(tile, M[h], lx);
        __syncwarp();
    }
}

// ---- folded mode: maps of edge <= S0 / 2 ------------------------------------------------------
// After the first stride-2 convolution a map needs at most 16 lanes, so the second packed array is
// folded into lanes 16..31 of the first (lane = 16 b + l: array b, column l).  Every later op then
// runs on ONE packed array a[NA] (M[0] while full-size ops are still to come, the small array F of
// phase B afterwards): half the instructions for the S0/2 and S0/4 stages of a ResNet.
__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    uint32_t lo, hi;
    split64(v, lo, hi);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return join64(lo, hi);
}

template <int S0, int S>
__device__ __forceinline__ void fold_op(u64 (&M)[2][S0], int lane) {
#pragma unroll
    for (int r = 0; r < S; ++r) {
        const u64 v = shfl64(M[1][r], lane & 15);
        if (lane >= 16) M[0][r] = v;
    }
}

// transposition tile in folded mode: row r holds array 0's columns at [0, L) and array 1's at [L, 2L)
template <int NA, int S0, int R, int L>
__device__ __forceinline__ void tstore_f(u64 *tile, const u64 (&a)[NA], int lane) {
    constexpr int PITCH = S0 + 1;
    static_assert(2 * L <= S0, "the pad column must be free");
    const int b = lane >> 4, l = lane & 15;
    // lanes that hold nothing write the pad column S0, which is never read (no divergent region around the stores)
    const int col = l < L ? b * L + l : S0;
#pragma unroll
    for (int r = 0; r < R; ++r) tile[r * PITCH + col] = a[r];
}
template <int NA, int S0, int R, int NEWL>
__device__ __forceinline__ void tload_f(const u64 *tile, u64 (&a)[NA], int lane) {
    constexpr int PITCH = S0 + 1;
    const int b = lane >> 4, l = lane & 15;
    const int lx = l < NEWL ? l : NEWL - 1;
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = tile[lx * PITCH + b * R + r];
}

template <int NA, int S0, int SI, int SO, int LO, int HI, int ST>
__device__ __forceinline__ void conv_op_f(u64 (&a)[NA], u64 *tile, int lane, float pre_bias, float scale, float bias) {
    static_assert(ST == 1 ? SI == SO : SI == 2 * SO, "conv geometry");
    static_assert(2 * SI <= S0 + 1, "both halves must fit a tile row");
    if (ST == 1) box_s1<NA, SI, LO, HI, false>(a, 0ull);
    else box_s2<NA, SI, SO, LO, HI>(a);
    tstore_f<NA, S0, SO, SI>(tile, a, lane);   // SO rows of SI columns per half
    __syncwarp();
    tload_f<NA, S0, SI, SO>(tile, a, lane);    // new lane = old row (< SO), registers = old columns (SI)
    __syncwarp();
    if (ST == 1) box_s1<NA, SI, LO, HI, true>(a, pk(pre_bias, pre_bias));
    else box_s2<NA, SI, SO, LO, HI>(a);
    if (scale != 1.f || bias != 0.f) affine_arr<NA, SO>(a, scale, bias);
}

template <int NA, int S0, int S>
__device__ __forceinline__ void transpose_op_f(u64 (&a)[NA], u64 *tile, int lane) {
    tstore_f<NA, S0, S, S>(tile, a, lane);
    __syncwarp();
    tload_f<NA, S0, S, S>(tile, a, lane);
    __syncwarp();
}

// 2 * H(e), degree-5 minimax fit on [0,1] (gram_fused.cu)
#define FNET_C5 1.678542030e-04f
#define FNET_C4 -1.571319990e-05f
#define FNET_C3 6.585370866e-04f
#define FNET_C2 2.386197913e-03f
#define FNET_C1 1.500756294e-02f
#define FNET_C0 3.001053929e-01f

// rows [R0, R1) of the ReLU step; ai / bj point at this lane's pixel of row 0 of the staged
// (s_a, s_b, 1/s_a, 1/s_b) maps of the warp's i-pair and j-pair
template <int S0, int S, int R0, int R1>
__device__ __forceinline__ void relu_rows(u64 (&M)[2][S0], const float4 *ai, const float4 *bj) {
    const u64 C5 = pk(FNET_C5, FNET_C5), C4 = pk(FNET_C4, FNET_C4), C3 = pk(FNET_C3, FNET_C3),
              C2 = pk(FNET_C2, FNET_C2), C1 = pk(FNET_C1, FNET_C1), C0 = pk(FNET_C0, FNET_C0), ONE = pk(1.f, 1.f);
#pragma unroll
    for (int r = R0; r < R1; ++r) {
        const float4 A = ai[r * S], B = bj[r * S];
        const u64 SA = pk(A.x, A.y), RA = pk(A.z, A.w);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const u64 SB = h ? pk(B.y, B.x) : pk(B.x, B.y);
            const u64 RB = h ? pk(B.w, B.z) : pk(B.z, B.w);
            float c0, c1;
            upk(M[h][r], c0, c1);
            const u64 NC = pk(neg_abs(c0), neg_abs(c1));
            const u64 D = fma2(SA, SB, NC);             // s - |c|
            const u64 RR = mul2(RA, RB);
            const u64 E = fma2(NC, RR, ONE);  // e = 1 - |c|/s
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
}

// folded: lanes 16..31 hold (i0 j1, i1 j0), i.e. they need the j-pair's operands swapped
template <int NA, int S>
__device__ __forceinline__ void relu_rows_f(u64 (&a)[NA], const float4 *ai, const float4 *bj, int lane) {
    const u64 C5 = pk(FNET_C5, FNET_C5), C4 = pk(FNET_C4, FNET_C4), C3 = pk(FNET_C3, FNET_C3),
              C2 = pk(FNET_C2, FNET_C2), C1 = pk(FNET_C1, FNET_C1), C0 = pk(FNET_C0, FNET_C0), ONE = pk(1.f, 1.f);
    const bool sw = lane >= 16;
#pragma unroll
    for (int r = 0; r < S; ++r) {
        const float4 A = ai[r * S], B = bj[r * S];
        const u64 SA = pk(A.x, A.y), RA = pk(A.z, A.w);
        const u64 SB = pk(sw ? B.y : B.x, sw ? B.x : B.y);
        const u64 RB = pk(sw ? B.w : B.z, sw ? B.z : B.w);
        float c0, c1;
        upk(a[r], c0, c1);
        const u64 NC = pk(neg_abs(c0), neg_abs(c1));
        const u64 D = fma2(SA, SB, NC);
        const u64 E = fma2(NC, mul2(RA, RB), ONE);
        float e0, e1;
        upk(E, e0, e1);
        const u64 W = mul2(D, pk(sqrt_approx(fabsf(e0)), sqrt_approx(fabsf(e1))));
        u64 H = fma2(C5, E, C4);
        H = fma2(H, E, C3);
        H = fma2(H, E, C2);
        H = fma2(H, E, C1);
        H = fma2(H, E, C0);
        a[r] = fma2(W, H, pk(fmaxf(c0, 0.f), fmaxf(c1, 0.f)));
    }
}

// the same step on one pixel in scalar code (1 x 1 maps after the pooling convolution); returns
// the doubled value like the packed version
__device__ __forceinline__ float relu_scalar(float c, float vx, float vy) {
    const float tiny = 1.0842021724855044e-19f;  // sqrt(f32_tiny), as cnngp_variances stores it
    const float sx = sqrtf(vx) + tiny, sy = sqrtf(vy) + tiny;
    const float s = sx * sy;
    const float nc = -fabsf(c);
    const float d = s + nc;
    const float e = fmaf(nc, (1.0f / sx) * (1.0f / sy), 1.0f);
    const float w = d * sqrtf(fabsf(e));
    float h = fmaf(FNET_C5, e, FNET_C4);
    h = fmaf(h, e, FNET_C3);
    h = fmaf(h, e, FNET_C2);
    h = fmaf(h, e, FNET_C1);
    h = fmaf(h, e, FNET_C0);
    return fmaf(w, h, fmaxf(c, 0.f));
}

template <int S0, int S>
__device__ __forceinline__ void dense_op(const u64 (&M)[2][S0], int lane, float scale, float bias, float (&tot)[4]) {
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
        tot[h == 0 ? 0 : 1] = fmaf(a0, scale, bias);
        tot[h == 0 ? 3 : 2] = fmaf(a1, scale, bias);
    }
}

template <int NA, int S>
__device__ __forceinline__ void dense_op_f(const u64 (&a)[NA], int lane, float scale, float bias, float (&tot)[4]) {
    u64 acc = a[0];
#pragma unroll
    for (int r = 1; r < S; ++r) acc = add2(acc, a[r]);
    float a0, a1;
    upk(acc, a0, a1);
    if ((lane & 15) >= S) { a0 = 0.f; a1 = 0.f; }
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) {  // within each 16-lane half
        a0 += __shfl_xor_sync(0xffffffffu, a0, d);
        a1 += __shfl_xor_sync(0xffffffffu, a1, d);
    }
    // half 0 = (i0j0, i1j1), half 1 = (i0j1, i1j0); tot index = 2a + b
    tot[0] = fmaf(__shfl_sync(0xffffffffu, a0, 0), scale, bias);
    tot[3] = fmaf(__shfl_sync(0xffffffffu, a1, 0), scale, bias);
    tot[1] = fmaf(__shfl_sync(0xffffffffu, a0, 16), scale, bias);
    tot[2] = fmaf(__shfl_sync(0xffffffffu, a1, 16), scale, bias);
}

// NSPLIT: row bands (= stages) a full-size ReLU layer's variance maps arrive in (2 or 4); a channel
// of the tile's images arrives in NSPLIT / 2 bands.
// PH: 3 = the whole program in one launch.  1 / 2 = the program split at the end of phase A into two
// launches with their own geometry: phase A needs the full register set (eight warps x 240
// registers: the maps alone are 112 / 128), phase B keeps S0 / 2 registers per thread and is bound by
// dependency latency, so it wants many warps per scheduler (sixteen x 120 registers).  Launch 1 ends
// every tile by writing each 2 x 2 block's folded map and tensor-memory slot 1 to global memory
// (7 / 8 KB per block, coalesced), launch 2 starts from there (other tile shape, same block grid).
template <int S0, int NW, int NST, int NSPLIT, int PH = 3>
__global__ void __launch_bounds__(NGeo<NW>::kThreads, 1) fnet_kernel(const __grid_constant__ NParams p) {
    using G = NGeo<NW>;
    constexpr int kWarps = G::kWarps, kTileI = G::kTileI, kTileJ = G::kTileJ, kImgs = G::kImgs, kPairs = G::kPairs;
    constexpr int P0 = S0 * S0;
    constexpr int PITCH = S0 + 1;
    constexpr int SF = S0 / 2;                                     // registers of a folded map
    constexpr int IMG_PARTS = NSPLIT / 2, IBAND = P0 / IMG_PARTS;  // pixels of one image band
    constexpr int BAND = P0 / NSPLIT;                              // pixels of one full-size ReLU band
    // bytes of a stage: one image band of kImgs images == kPairs ReLU bands of float4; phase B alone: one folded layer
    constexpr int STAGE = PH == 2 ? kPairs * ((SF * SF + 1) / 2) * 32 : kImgs * IBAND * 4;
    constexpr int TILE_ELEMS = PH == 2 ? SF * PITCH : S0 * PITCH;  // transposition tile of a warp (folded maps: SF rows)
    constexpr int REC = 2 * SF * 32;                               // packed entries of a hand-off record
    static_assert(NSPLIT == 2 || NSPLIT == 4, "band split");
    static_assert(PH == 2 || (kPairs * BAND * 16 == STAGE && S0 % NSPLIT == 0), "stage geometry");
    static_assert(PH >= 1 && PH <= 3 && (NW != 16 || PH == 2), "sixteen warps: phase B only");
    // tensor-memory window of a warp: 8 warps: 256 columns, slots at 0 / 128 with 64 columns per
    // array; 12 warps (three per lane quadrant): 5 S0 columns, slot 0 (full size) at 0 with 2 S0
    // per array, slot 1 (a folded map of at most half the edge: one array of S0 columns) at 4 S0
    // (12 warps, aligned: 160 columns, the two arrays of slot 0 at 0 and 64, slot 1 at 128 -- every x16 group
    // then starts at a multiple of 16 columns)
    // (16 warps, folded maps only: 128 columns, slots at 0 / 64)
    constexpr int TM_WARP = NW == 8 ? 256 : (NW == 12 ? 160 : 128), TM_SLOT1 = NW == 16 ? 64 : 128;
    constexpr int TM_A0 = 64, TM_A1 = NW == 8 ? 64 : 32;
    static_assert(2 * S0 <= 64 && (NW == 8 || S0 <= 32), "a packed array of S0 entries takes 2 S0 columns");
    static_assert((NW / 4) * TM_WARP <= kTmemCols, "tensor-memory budget");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *stage = smem_raw;
    u64 *tiles = reinterpret_cast<u64 *>(smem_raw + (size_t)NST * STAGE);
    int4 *ops_s = reinterpret_cast<int4 *>(tiles + kWarps * TILE_ELEMS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ops_s + kMaxKOps);
    uint64_t *full = bars, *empty = bars + NST;
    // tile index each stage belongs to (-1: no more tiles); tiles are handed out by a global atomic
    // counter so that all CTAs stay on consecutive tiles of one super-tile (see gram_fused.cu)
    long long *stage_tile = reinterpret_cast<long long *>(bars + 2 * NST);
    uint32_t *tmem_word = reinterpret_cast<uint32_t *>(bars + 3 * NST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the op descriptors move from the parameter bank to shared memory: one LDS.128 per op
    for (int q = threadIdx.x; q < p.n_ops; q += blockDim.x) {
        const KOp o = p.ops[q];
        ops_s[q] = make_int4(o.code, __float_as_int(o.scale), __float_as_int(o.bias), o.aux);
    }
    if (warp == 0) {  // one warp allocates all of this SM's tensor memory (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_word;

    const int per_st = p.sti * p.stj;
    int dec_st = 0, dec_w = 0;  // super-tile and position inside it of the last decoded tile
    auto decode = [&](long long t, int &ib, int &jb) -> bool {
        const int st = (int)(t / per_st), w_in = (int)(t - (long long)st * per_st);
        dec_st = st; dec_w = w_in;
        int si, sj;
        if (p.symmetric) {
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
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::kRegsProducer));
        if (warp != kWarps) return;
        // ---- producer warp ----------------------------------------------------------------
        // Lane 0 owns the ring protocol (waits for a free stage, publishes the tile index, posts the
        // expected byte count); the copies of a stage are issued by the lanes in parallel, each
        // from a base pointer it worked out ONCE per tile: lane u < 2 kPairs serves image row
        // 2 pr + (u & 1) of tile pair u >> 1 (variance maps), lane s < kImgs serves tile image s.
        // A single lane doing all of it was the bottleneck of the small (folded) layers, whose
        // sixteen copies cost more to set up than the layer costs to compute.
        {
            unsigned l = 0;
            const int last_pi = (p.N1 - 1) >> 1, last_pj = (p.N2 - 1) >> 1;
            long long t = 0;
            auto acquire = [&](unsigned bytes) -> unsigned char * {
                const unsigned buf = l % NST;
                if (lane == 0) {
                    if (l >= NST) mbar_wait_relaxed(&empty[buf], ((l / NST) - 1) & 1);
                    stage_tile[buf] = t;  // published by the release of the arrive below
                    mbar_arrive_expect_tx(&full[buf], bytes);
                }
                __syncwarp();
                return stage + (size_t)buf * STAGE;
            };
            // tile_ctr == NULL: fixed stride (tile = blockIdx.x + k * gridDim.x), kept for comparison
            const bool dyn = p.tile_ctr != nullptr;
            auto next_index = [&](long long stat) -> long long {
                if (!dyn) return stat;
                long long v = 0;
                if (lane == 0) v = p.t_begin + (long long)atomicAdd(p.tile_ctr, 1ull);
                return __shfl_sync(0xffffffffu, v, 0);
            };
            long long t_raw = next_index(p.t_begin + (long long)blockIdx.x);
            const int relu_lo = PH == 2 ? p.n_relu_a : 0, relu_hi = PH == 1 ? p.n_relu_a : p.n_relu;
            for (;;) {
                int ib, jb;
                t = t_raw;
                while (t < p.n_tiles && !decode(t, ib, jb)) t = next_index(t + gridDim.x);
                if (t >= p.n_tiles) break;
                // the next index is requested now and first looked at when this tile's stages are out
                t_raw = next_index(t + gridDim.x);
                const int i_base = ib * kTileI, j_base = jb * kTileJ;
                // this lane's sources for the whole tile
                const float *img = nullptr, *var = nullptr;
                if (lane < kImgs)
                    img = lane < kTileI ? p.x + (long long)min(i_base + lane, p.N1 - 1) * p.C * P0
                                        : p.z + (long long)min(j_base + lane - kTileI, p.N2 - 1) * p.C * P0;
                const int vs = lane >> 1, vh = lane & 1;  // pair of the tile, image row of the pair
                if (lane < 2 * kPairs) {
                    const long long pr = vs < kTileI / 2 ? min((i_base >> 1) + vs, last_pi)
                                                         : min((j_base >> 1) + vs - kTileI / 2, last_pj);
                    var = (vs < kTileI / 2 ? p.aux_x : p.aux_z) + (2 * pr + vh) * p.aux_stride + p.aux_f_off;
                }
                for (int c = 0; c < (PH == 2 ? 0 : p.C); ++c) {
                    for (int ip = 0; ip < IMG_PARTS; ++ip) {
                        float *dst = reinterpret_cast<float *>(acquire(kImgs * IBAND * 4));
                        if (lane < kImgs) bulk_g2s(dst + lane * IBAND, img + (long long)c * P0 + ip * IBAND, IBAND * 4, &full[l % NST]);
                        ++l;
                    }
                }
                for (int k = relu_lo; k < relu_hi; ++k) {
                    const int off = p.relu_aux[k];
                    const int half = p.relu_half[k] & 0xffff;
                    if (PH != 2 && (p.relu_half[k] >> 30)) {  // full-size layers: NSPLIT row bands, one stage each
                        for (int part = 0; part < NSPLIT; ++part) {
                            float4 *dst = reinterpret_cast<float4 *>(acquire((unsigned)(kPairs * BAND * 16)));
                            // band `part`: pixels [part * BAND, +BAND); the first half of the pixels is in row 2k
                            if (lane < 2 * kPairs && vh == part / (NSPLIT / 2))
                                bulk_g2s(dst + vs * BAND, var + off + (part % (NSPLIT / 2)) * BAND * 4, BAND * 16, &full[l % NST]);
                            ++l;
                        }
                    } else {  // folded maps: both halves of the layer in one stage
                        float4 *dst = reinterpret_cast<float4 *>(acquire((unsigned)(kPairs * half * 32)));
                        if (lane < 2 * kPairs) bulk_g2s(dst + vs * 2 * half + vh * half, var + off, half * 16, &full[l % NST]);
                        ++l;
                    }
                }
            }
            t = -1;  // end marker: one empty stage
            acquire(0);
        }
        return;
    }

    // ---- consumers ------------------------------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::kRegsConsumer));
    const int wi = warp / (NW / 2), wj = warp % (NW / 2);
    u64 *tile = tiles + warp * TILE_ELEMS;
    // this warp's tensor-memory window: lanes of quadrant warp % 4
    const uint32_t tm_warp = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TM_WARP);
    unsigned stage_l = 0;

    // descriptor fields
    auto f_scale = [](const int4 &o) { return __int_as_float(o.y); };
    auto f_bias = [](const int4 &o) { return __int_as_float(o.z); };
    auto f_pre = [](const int4 &o) { return __int_as_float(o.w); };
    auto f_slot = [](const int4 &o) { return (o.x >> 8) & 3; };
    auto f_half = [](const int4 &o) { return (int)((unsigned)o.x >> 16); };

    for (;;) {
        // the tile this CTA works on next travels with its first stage
        mbar_wait(&full[stage_l % NST], (stage_l / NST) & 1);
        const long long t = stage_tile[stage_l % NST];
        if (t < 0) break;
        int ib, jb;
        decode(t, ib, jb);
        const int i_base = ib * kTileI, j_base = jb * kTileJ;

        float tot[4] = {0.f, 0.f, 0.f, 0.f};
        u64 F[SF];
        int k = PH == 2 ? p.k_b : 0;
        int4 o = ops_s[k];
        // this warp's hand-off record (split launches): block row / column inside the super-tile
        u64 *rec = nullptr;
        if (PH != 3)
            rec = p.handoff + ((size_t)(dec_st - p.st_begin) * p.rec_per_st + (size_t)(dec_w / p.stj * (kTileI / 2) + wi) * p.rec_ld +
                               (size_t)(dec_w % p.stj) * (kTileJ / 2) + wj) * REC + lane;
        if constexpr (PH != 2) {   // ================= phase A: the full register set =================================
            u64 M[2][S0];
            {   // init, kernels.py:43-49
                const int lx = lane < S0 ? lane : S0 - 1;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int r = 0; r < S0; ++r) M[h][r] = 0ull;
                for (int c = 0; c < p.C; ++c) {
#pragma unroll
                    for (int ip = 0; ip < IMG_PARTS; ++ip) {
                        const unsigned buf = stage_l % NST;
                        mbar_wait(&full[buf], (stage_l / NST) & 1);
                        const float *sb = reinterpret_cast<const float *>(stage + (size_t)buf * STAGE) + lx;
                        const float *x0 = sb + (wi * 2 + 0) * IBAND, *x1 = sb + (wi * 2 + 1) * IBAND;
                        const float *z0 = sb + (kTileI + wj * 2 + 0) * IBAND, *z1 = sb + (kTileI + wj * 2 + 1) * IBAND;
                        constexpr int R = S0 / IMG_PARTS;
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) {
                            const int r = ip * R + rr;
                            const float a0 = x0[rr * S0], a1 = x1[rr * S0], b0 = z0[rr * S0], b1 = z1[rr * S0];
                            const u64 A = pk(a0, a1);
                            M[0][r] = fma2(A, pk(b0, b1), M[0][r]);
                            M[1][r] = fma2(A, pk(b1, b0), M[1][r]);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[buf]);
                        ++stage_l;
                    }
                }
                if (p.C > 1) { affine_arr<S0, S0>(M[0], p.inv_c, 0.f); affine_arr<S0, S0>(M[1], p.inv_c, 0.f); }
            }

            // ---- the ops of phase A as callable pieces (single cases and block cases share them)
            auto relu_full = [&]() {  // NSPLIT stages, one row band each
                const int lx = lane < S0 ? lane : S0 - 1;
                auto band = [&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    const unsigned buf = stage_l % NST;
                    mbar_wait(&full[buf], (stage_l / NST) & 1);
                    // the stage holds pixels [q * BAND, +BAND) of every pair: bias the pointer so that r * S0 + lx indexes it
                    const float4 *sb = reinterpret_cast<const float4 *>(stage + (size_t)buf * STAGE) + lx - q * BAND;
                    relu_rows<S0, S0, q * (S0 / NSPLIT), (q + 1) * (S0 / NSPLIT)>(M, sb + wi * BAND, sb + (kTileI / 2 + wj) * BAND);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[buf]);
                    ++stage_l;
                };
                band(std::integral_constant<int, 0>{});
                band(std::integral_constant<int, 1>{});
                if (NSPLIT == 4) {
                    band(std::integral_constant<int, NSPLIT == 4 ? 2 : 0>{});
                    band(std::integral_constant<int, NSPLIT == 4 ? 3 : 1>{});
                }
            };
            auto relu_fold = [&](auto SZ, const int4 &d) {  // folded map in M[0]: one stage holds the whole layer
                constexpr int S = decltype(SZ)::value;
                const unsigned buf = stage_l % NST;
                mbar_wait(&full[buf], (stage_l / NST) & 1);
                const float4 *st4 = reinterpret_cast<const float4 *>(stage + (size_t)buf * STAGE);
                const int l = lane & 15, half = f_half(d);
                const float4 *sb = st4 + (l < S ? l : S - 1);
                relu_rows_f<S0, S>(M[0], sb + wi * 2 * half, sb + (kTileI / 2 + wj) * 2 * half, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[buf]);
                ++stage_l;
            };
            auto stash_full = [&](const int4 &d) {
                const int s = f_slot(d);
                const uint32_t ta = tm_warp + s * TM_SLOT1;
                stash_store_arr<S0, S0>(ta, M[0]);
                stash_store_arr<S0, S0>(ta + (s ? TM_A1 : TM_A0), M[1]);
                tmem_wait_st();
            };
            auto unstash_full = [&](const int4 &d) {
                const int s = f_slot(d);
                const uint32_t ta = tm_warp + s * TM_SLOT1;
                stash_load_arr<S0, S0, false>(ta, M[0], 0ull);
                stash_load_arr<S0, S0, false>(ta + (s ? TM_A1 : TM_A0), M[1], 0ull);
            };
            auto add_full = [&](const int4 &d) {
                const int s = f_slot(d);
                const uint32_t ta = tm_warp + s * TM_SLOT1;
                const float al = f_scale(d);
                const u64 alpha = pk(al, al);
                stash_load_arr<S0, S0, true, NW == 12>(ta, M[0], alpha);
                stash_load_arr<S0, S0, true, NW == 12>(ta + (s ? TM_A1 : TM_A0), M[1], alpha);
                add_const<S0, S0>(M[0], f_bias(d));
                add_const<S0, S0>(M[1], f_bias(d));
            };

            for (bool more = true; more;) {
                int4 nxt;
#define FETCH(LEN) nxt = ops_s[k + (LEN)]; k += (LEN)
#define FOLDED_A(CODE, BODY)                                                               \
    case CODE + 1: { constexpr int S = S0 / 2; FETCH(1); BODY; break; }                    \
    case CODE + 2: { constexpr int S = S0 / 4; FETCH(1); BODY; break; }
                switch (o.x & 0xff) {
                    case A_CONV + 0: FETCH(1); conv_op<S0, S0, S0, 1, 1, 1>(M, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_CONV + 1: FETCH(1); conv_op<S0, S0, S0, 1, 2, 1>(M, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_CONV + 2: FETCH(1); conv_op<S0, S0, S0, 2, 2, 1>(M, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_CONV + 3: FETCH(1); conv_op<S0, S0, S0, 3, 3, 1>(M, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    // stride 2 out of the full size: unfolded op, then the result is folded
                    case A_CONV + 4: FETCH(1); conv_op<S0, S0, S0 / 2, 1, 1, 2>(M, tile, lane, 0.f, f_scale(o), f_bias(o)); fold_op<S0, S0 / 2>(M, lane); break;
                    case A_CONV + 5: FETCH(1); conv_op<S0, S0, S0 / 2, 0, 0, 2>(M, tile, lane, 0.f, f_scale(o), f_bias(o)); fold_op<S0, S0 / 2>(M, lane); break;
                    case A_CONV + 6: FETCH(1); conv_op_f<S0, S0, S0 / 2, S0 / 2, 1, 1, 1>(M[0], tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_CONV + 7: FETCH(1); conv_op_f<S0, S0, S0 / 2, S0 / 4, 1, 1, 2>(M[0], tile, lane, 0.f, f_scale(o), f_bias(o)); break;
                    case A_CONV + 8: FETCH(1); conv_op_f<S0, S0, S0 / 2, S0 / 4, 0, 0, 2>(M[0], tile, lane, 0.f, f_scale(o), f_bias(o)); break;
                    case A_CONV + 9: FETCH(1); conv_op_f<S0, S0, S0 / 4, S0 / 4, 1, 1, 1>(M[0], tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_CONV + 10: FETCH(1); conv_op<S0, S0, S0, 1, 1, 1, 2>(M, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case A_AFFINE + 0: FETCH(1); affine_arr<S0, S0>(M[0], f_scale(o), f_bias(o)); affine_arr<S0, S0>(M[1], f_scale(o), f_bias(o)); break;
                    FOLDED_A(A_AFFINE, (affine_arr<S0, S>(M[0], f_scale(o), f_bias(o))))
                    case A_TRANSPOSE + 0: FETCH(1); transpose_op<S0, S0>(M, tile, lane); break;
                    FOLDED_A(A_TRANSPOSE, (transpose_op_f<S0, S0, S>(M[0], tile, lane)))
                    case A_STASH + 0: FETCH(1); stash_full(o); break;
                    FOLDED_A(A_STASH, (stash_store_arr<S0, S>(tm_warp + f_slot(o) * TM_SLOT1, M[0]), tmem_wait_st()))
                    case A_UNSTASH + 0: FETCH(1); unstash_full(o); break;
                    FOLDED_A(A_UNSTASH, (stash_load_arr<S0, S, false>(tm_warp + f_slot(o) * TM_SLOT1, M[0], 0ull)))
                    case A_ADD + 0: FETCH(1); add_full(o); break;
                    FOLDED_A(A_ADD, (stash_load_arr<S0, S, true>(tm_warp + f_slot(o) * TM_SLOT1, M[0], pk(f_scale(o), f_scale(o))),
                                     add_const<S0, S>(M[0], f_bias(o))))
                    case A_DENSE + 0: FETCH(1); dense_op<S0, S0>(M, lane, f_scale(o), f_bias(o), tot); break;
                    FOLDED_A(A_DENSE, (dense_op_f<S0, S>(M[0], lane, f_scale(o), f_bias(o), tot)))
                    case A_RELU + 0: FETCH(1); relu_full(); break;
                    case A_RELU + 1: FETCH(1); relu_fold(std::integral_constant<int, S0 / 2>{}, o); break;
                    case A_RELU + 2: FETCH(1); relu_fold(std::integral_constant<int, S0 / 4>{}, o); break;
                    case A_IDBLOCK: {  // STASH RELU CONV(3x3) RELU CONV(3x3) ADD: an identity residual block
                        const int4 c1 = ops_s[k + 2], c2 = ops_s[k + 4], ad = ops_s[k + 5];
                        FETCH(6);
                        stash_full(o);
                        relu_full();
                        conv_op<S0, S0, S0, 1, 1, 1>(M, tile, lane, f_pre(c1), f_scale(c1), f_bias(c1));
                        relu_full();
                        conv_op<S0, S0, S0, 1, 1, 1>(M, tile, lane, f_pre(c2), f_scale(c2), f_bias(c2));
                        add_full(ad);
                        break;
                    }
                    case A_RESBLOCK: {  // STASH CONV(4x4 "same") RELU TRANSPOSE ADD: mnist_paper_residual_cnn_gp
                        const int4 c1 = ops_s[k + 1], ad = ops_s[k + 4];
                        FETCH(5);
                        stash_full(o);
                        conv_op<S0, S0, S0, 1, 2, 1>(M, tile, lane, f_pre(c1), f_scale(c1), f_bias(c1));
                        relu_full();
                        transpose_op<S0, S0>(M, tile, lane);
                        add_full(ad);
                        break;
                    }
                    default: FETCH(1); more = false; break;  // A_END
                }
#undef FOLDED_A
                // twelve warps: the next descriptor is fetched now, not one op ahead (four registers less across the op)
                o = NW == 12 ? ops_s[k] : nxt;
            }
#pragma unroll
            for (int r = 0; r < SF; ++r) F[r] = M[0][r];
            if constexpr (PH == 1) {  // hand the folded map and tensor-memory slot 1 over to the phase-B launch
                u64 K[SF];
                stash_load_arr<SF, SF, false>(tm_warp + TM_SLOT1, K, 0ull);
#pragma unroll
                for (int r = 0; r < SF; ++r) {
                    rec[r * 32] = F[r];
                    rec[(SF + r) * 32] = K[r];
                }
                continue;
            }
        }

        if constexpr (PH != 1) {   // ================= phase B: folded maps, S0 / 2 live registers ====================
            // single launch: tensor-memory slot 1 lives in registers here (K): phase B has them to spare at 160 / 240
            // registers per thread, and its residual blocks then never touch tensor memory
            // (the phase-B launch keeps it in tensor memory: its sixteen warps run at 112 registers, where the 28 / 32
            // registers of K were spilled -- 193 against 200 ms per 6 000 x 6 000 mnist_as_tf Gram)
            constexpr bool KREGS = PH != 2;
            u64 K[KREGS ? SF : 1];
            if constexpr (PH == 2) {
#pragma unroll
                for (int r = 0; r < SF; ++r) F[r] = rec[r * 32];
                if constexpr (KREGS) {
#pragma unroll
                    for (int r = 0; r < SF; ++r) K[r] = rec[(SF + r) * 32];
                } else {
                    u64 T[SF];
#pragma unroll
                    for (int r = 0; r < SF; ++r) T[r] = rec[(SF + r) * 32];
                    stash_store_arr<SF, SF>(tm_warp + TM_SLOT1, T);
                    tmem_wait_st();
                }
            } else if constexpr (KREGS) {
                stash_load_arr<SF, SF, false>(tm_warp + TM_SLOT1, K, 0ull);
            }
            auto b_stash = [&](auto SZ, const int4 &d) {
                constexpr int S = decltype(SZ)::value;
                if constexpr (!KREGS) {
                    stash_store_arr<SF, S>(tm_warp + f_slot(d) * TM_SLOT1, F);
                    tmem_wait_st();
                } else if (f_slot(d)) {
#pragma unroll
                    for (int r = 0; r < S; ++r) K[r] = F[r];
                } else {
                    stash_store_arr<SF, S>(tm_warp, F);
                    tmem_wait_st();
                }
            };
            auto b_unstash = [&](auto SZ, const int4 &d) {
                constexpr int S = decltype(SZ)::value;
                if constexpr (!KREGS) {
                    stash_load_arr<SF, S, false>(tm_warp + f_slot(d) * TM_SLOT1, F, 0ull);
                } else if (f_slot(d)) {
#pragma unroll
                    for (int r = 0; r < S; ++r) F[r] = K[r];
                } else {
                    stash_load_arr<SF, S, false>(tm_warp, F, 0ull);
                }
            };
            auto b_add = [&](auto SZ, const int4 &d) {
                constexpr int S = decltype(SZ)::value;
                const u64 alpha = pk(f_scale(d), f_scale(d));
                if constexpr (!KREGS) {
                    stash_load_arr<SF, S, true>(tm_warp + f_slot(d) * TM_SLOT1, F, alpha);
                } else if (f_slot(d)) {
#pragma unroll
                    for (int r = 0; r < S; ++r) F[r] = fma2(K[r], alpha, F[r]);
                } else {
                    stash_load_arr<SF, S, true>(tm_warp, F, alpha);
                }
                add_const<SF, S>(F, f_bias(d));
            };
            auto relu_fold = [&](auto SZ, const int4 &d) {
                constexpr int S = decltype(SZ)::value;
                const unsigned buf = stage_l % NST;
                mbar_wait(&full[buf], (stage_l / NST) & 1);
                const float4 *st4 = reinterpret_cast<const float4 *>(stage + (size_t)buf * STAGE);
                const int l = lane & 15, half = f_half(d);
                const float4 *sb = st4 + (l < S ? l : S - 1);
                relu_rows_f<SF, S>(F, sb + wi * 2 * half, sb + (kTileI / 2 + wj) * 2 * half, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[buf]);
                ++stage_l;
            };
            auto idblock = [&](auto SZ, const int4 &st, const int4 &r1, const int4 &c1, const int4 &r2, const int4 &c2, const int4 &ad) {
                constexpr int S = decltype(SZ)::value;
                b_stash(SZ, st);
                relu_fold(SZ, r1);
                conv_op_f<SF, S0, S, S, 1, 1, 1>(F, tile, lane, f_pre(c1), f_scale(c1), f_bias(c1));
                relu_fold(SZ, r2);
                conv_op_f<SF, S0, S, S, 1, 1, 1>(F, tile, lane, f_pre(c2), f_scale(c2), f_bias(c2));
                b_add(SZ, ad);
            };
            for (bool more = true; more;) {
                int4 nxt;
#define FOLDED_B(CODE, BODY)                                                               \
    case CODE + 0: { constexpr int S = S0 / 2; FETCH(1); BODY; break; }                    \
    case CODE + 1: { constexpr int S = S0 / 4; FETCH(1); BODY; break; }
                switch (o.x & 0xff) {
                    case B_CONV + 0: FETCH(1); conv_op_f<SF, S0, S0 / 2, S0 / 2, 1, 1, 1>(F, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    case B_CONV + 1: FETCH(1); conv_op_f<SF, S0, S0 / 2, S0 / 4, 1, 1, 2>(F, tile, lane, 0.f, f_scale(o), f_bias(o)); break;
                    case B_CONV + 2: FETCH(1); conv_op_f<SF, S0, S0 / 2, S0 / 4, 0, 0, 2>(F, tile, lane, 0.f, f_scale(o), f_bias(o)); break;
                    case B_CONV + 3: FETCH(1); conv_op_f<SF, S0, S0 / 4, S0 / 4, 1, 1, 1>(F, tile, lane, f_pre(o), f_scale(o), f_bias(o)); break;
                    FOLDED_B(B_AFFINE, (affine_arr<SF, S>(F, f_scale(o), f_bias(o))))
                    FOLDED_B(B_TRANSPOSE, (transpose_op_f<SF, S0, S>(F, tile, lane)))
                    FOLDED_B(B_STASH, (b_stash(std::integral_constant<int, S>{}, o)))
                    FOLDED_B(B_UNSTASH, (b_unstash(std::integral_constant<int, S>{}, o)))
                    FOLDED_B(B_ADD, (b_add(std::integral_constant<int, S>{}, o)))
                    FOLDED_B(B_DENSE, (dense_op_f<SF, S>(F, lane, f_scale(o), f_bias(o), tot)))
                    case B_RELU + 0: FETCH(1); relu_fold(std::integral_constant<int, S0 / 2>{}, o); break;
                    case B_RELU + 1: FETCH(1); relu_fold(std::integral_constant<int, S0 / 4>{}, o); break;
                    case B_IDBLOCK + 0: {
                        const int4 r1 = ops_s[k + 1], c1 = ops_s[k + 2], r2 = ops_s[k + 3], c2 = ops_s[k + 4], ad = ops_s[k + 5];
                        FETCH(6);
                        idblock(std::integral_constant<int, S0 / 2>{}, o, r1, c1, r2, c2, ad);
                        break;
                    }
                    case B_IDBLOCK + 1: {
                        const int4 r1 = ops_s[k + 1], c1 = ops_s[k + 2], r2 = ops_s[k + 3], c2 = ops_s[k + 4], ad = ops_s[k + 5];
                        FETCH(6);
                        idblock(std::integral_constant<int, S0 / 4>{}, o, r1, c1, r2, c2, ad);
                        break;
                    }
                    default: FETCH(1); more = false; break;  // B_END
                }
#undef FOLDED_B
                o = nxt;
            }
        }

        // ================= tail: the four scalars of the warp ==================================
        for (bool more = true; more;) {
            int4 nxt;
            switch (o.x & 0xff) {
                case T_CASE_AFFINE:
                    FETCH(1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) tot[q] = fmaf(tot[q], f_scale(o), f_bias(o));
                    break;
                case T_CASE_RELU: {
                    FETCH(1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = min(i_base + wi * 2 + (q >> 1), p.N1 - 1), j = min(j_base + wj * 2 + (q & 1), p.N2 - 1);
                        const float vx = __ldg(p.aux_x + (long long)i * p.aux_stride + o.w);
                        const float vy = __ldg(p.aux_z + (long long)j * p.aux_stride + o.w);
                        tot[q] = relu_scalar(tot[q], vx, vy);
                    }
                    break;
                }
                default: more = false; break;  // T_END
            }
            if (more) o = nxt;
        }
#undef FETCH

        if (lane < 4) {
            const int a = lane >> 1, b = lane & 1;
            const int i = i_base + wi * 2 + a, j = j_base + wj * 2 + b;
            const float v = lane == 0 ? tot[0] : lane == 1 ? tot[1] : lane == 2 ? tot[2] : tot[3];
            if (i < p.N1 && j < p.N2) {
                if (!p.symmetric) {
                    p.out[(long long)i * p.ld_out + j] = v;
                } else if (j > i) {
                    p.out[(long long)i * p.ld_out + j] = v;
                    // a band of block rows (N2 > N1): the mirror image stays inside the diagonal block (gram_fused.cu)
                    if (j < p.N1 && (p.mirror_bs == 0 || j / p.mirror_bs == i / p.mirror_bs))
                        p.out[(long long)j * p.ld_out + i] = v;
                } else if (j == i) {
                    // i == j follows the variance recursion (kernels.py:155-162)
                    p.out[(long long)i * p.ld_out + i] = p.kdiag ? p.kdiag[i] : v;
                }
            }
        }
        if (p.row_done) {  // the entries above are visible device-wide before the count moves
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(&p.row_done[ib / p.sti], 1u);
        }
    }

    // all consumers are done with tensor memory before the allocating warp frees it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

// ---- host: translate the two-slot plan into the kernel's op list --------------------------------
struct FNetPlan {
    int S0 = 0;
    int n_ops = 0;            // register-level ops (what dump() lists)
    NOp ops[kMaxNOps];
    int n_kops = 0;           // kernel descriptors: ops with dispatch cases (blocks grouped) + phase sentinels
    KOp kops[kMaxKOps];
    int n_relu = 0;           // staged ReLU layers in program order (the producer's list)
    int relu_aux[kMaxRelu], relu_half[kMaxRelu];
    int n_blocks = 0;         // residual blocks dispatched as one case
    bool folded_phase = false;  // the program has ops on folded (half / quarter size) maps
    size_t smem = 0;
    int nst = 0;
    int fused_row_floats = 0;  // floats per image the Gram kernel reads (super-tile sizing)
    int nw = 8;                // consumer warps of the kernel variant
    // split launches: phase A (8 warps) and phase B (16 warps) as two kernels with a global hand-off
    bool split = false;
    int n_relu_a = 0;          // staged ReLU layers consumed by phase A
    int k_b = 0;               // first kernel descriptor of phase B
    size_t smem_a = 0, smem_b = 0;
    int nw_a = 8;              // consumer warps of the phase-A launch (8 or 12)
};

namespace {

bool conv_supported(int S0, int si, int lo, int hi, int st, int dil) {
    if (dil != 1) return dil == 2 && si == S0 && st == 1 && lo == 1 && hi == 1;  // 3 x 3, dilation 2, "same"
    if (si == S0) {
        if (st == 1) return (lo == 1 && hi == 1) || (lo == 1 && hi == 2) || (lo == 2 && hi == 2) || (lo == 3 && hi == 3);
        return (lo == 1 && hi == 1) || (lo == 0 && hi == 0);
    }
    if (si == S0 / 2) {
        if (st == 1) return lo == 1 && hi == 1;
        return (lo == 1 && hi == 1) || (lo == 0 && hi == 0);
    }
    if (si == S0 / 4) return st == 1 && lo == 1 && hi == 1;
    return false;
}

struct Translator {
    const std::vector<DevOp> &ops;
    int S0;
    bool carry;  // carry conv taps / ReLU doublings as per-slot factors instead of applying them
    std::vector<NOp> out;
    bool ok = true;
    // per program slot
    // `pend`: the stored values are pend x the true ones.  A slot may be an ALIAS of a stashed map:
    // true(slot) = a_scale * true(stash) + a_bias, where the stash holds a_pend x its true values
    // (a 1 x 1 convolution, Mixture scaling or copy of a map that stays live is never materialised).
    struct Slot {
        bool valid = false; int size = 0; int orient = 0; float pend = 1.f; bool in_regs = false; int tm = -1;
        bool alias = false; float a_scale = 1.f, a_bias = 0.f;
    };
    std::vector<Slot> slot;
    int tm_owner[2] = {-1, -1};
    int slot1_max = 0;  // largest map edge ever stashed in tensor-memory slot 1

    Translator(const std::vector<DevOp> &o, int n_slots, int s0, bool carry_)
        : ops(o), S0(s0), carry(carry_), slot(n_slots) {}

    bool reads(const DevOp &o, int s) const { return o.src == s || (o.opcode == CNNGP_OP_ADD && o.dst == s); }
    // is the value now in slot s needed by an op after index k?
    bool live_after(int k, int s) const {
        for (size_t q = k + 1; q < ops.size(); ++q) {
            if (reads(ops[q], s)) return true;
            if (ops[q].dst == s) return false;  // overwritten without being read
        }
        return false;
    }
    int reg_slot() const {
        for (size_t s = 0; s < slot.size(); ++s)
            if (slot[s].valid && slot[s].in_regs) return (int)s;
        return -1;
    }
    void emit(NOp n) { out.push_back(n); }
    static NOp blank(int kind, int si, int so) {
        NOp n{};
        n.kind = kind; n.si = (short)si; n.so = (short)so; n.scale = 1.f; n.bias = 0.f; n.pre_bias = 0.f; n.aux_scale = 1.f;
        n.dil = 1;
        return n;
    }
    void affine(int size, float scale, float bias) {
        NOp n = blank(size == 1 ? T_AFFINE : N_AFFINE, size, size);
        n.scale = scale; n.bias = bias;
        emit(n);
    }
    // the factor a slot may carry: a power-of-two-ish safe range around 1, so that stored values
    // (pend x true) and the scaled variance maps stay far from the ends of the float32 range
    static bool pend_ok(double pend) { return pend == pend && pend > 9.5e-7 && pend < 1.05e6; }

    void free_tm(int s) {
        if (slot[s].tm >= 0) { tm_owner[slot[s].tm] = -1; slot[s].tm = -1; }
    }
    bool stash(int s) {  // copy the register-resident slot s to a free tensor-memory slot
        if (slot[s].tm >= 0) return true;
        if (slot[s].size == 1) return false;  // 1 x 1 maps live in scalars: no stash
        // full-size maps prefer slot 0, smaller ones slot 1 (the 12-warp kernel's slot 1 is half size)
        const int first = slot[s].size == S0 ? 0 : 1;
        for (int q = 0; q < 2; ++q) {
            const int t = first ^ q;
            if (tm_owner[t] < 0) {
                NOp n = blank(N_STASH, slot[s].size, slot[s].size);
                n.slot = (short)t;
                emit(n);
                tm_owner[t] = s; slot[s].tm = t;
                if (t == 1 && slot[s].size > slot1_max) slot1_max = slot[s].size;
                return true;
            }
        }
        return false;
    }
    // make slot s the register-resident one; k = index of the op about to run
    bool to_regs(int s, int k) {
        if (slot[s].in_regs) return true;
        if (slot[s].tm < 0) return false;
        const int r = reg_slot();
        if (r >= 0) {
            // the op at k still counts as a reader of r
            const bool needed = reads(ops[k], r) || (ops[k].dst != r && live_after(k, r));
            if (needed && !stash(r)) return false;
            slot[r].in_regs = false;
            if (!needed) { slot[r].valid = false; free_tm(r); }
        }
        NOp n = blank(N_UNSTASH, slot[s].size, slot[s].size);
        n.slot = (short)slot[s].tm;
        emit(n);
        slot[s].in_regs = true;
        if (slot[s].alias) {  // materialise: registers hold pend x true(stash); keep that factor
            if (slot[s].a_scale != 1.f || slot[s].a_bias != 0.f) affine(slot[s].size, slot[s].a_scale, slot[s].a_bias * slot[s].pend);
            slot[s].alias = false;
            free_tm(s);  // the stash holds the un-aliased map, not this slot's value
        }
        return true;
    }
    bool run(std::vector<DevOp> &mutable_ops, int final_slot) {
        slot[0].valid = true; slot[0].size = S0; slot[0].in_regs = true;
        for (size_t k = 0; k < ops.size(); ++k) {
            const DevOp &o = ops[k];
            if (!slot[o.src].valid) return false;
            if (o.opcode == CNNGP_OP_ADD) {  // dst += src
                if (!slot[o.dst].valid || slot[o.dst].size != slot[o.src].size || o.src == o.dst) return false;
                if (slot[o.dst].size == 1) return false;  // Sum on 1 x 1 maps: not in the set
                // registers must hold one operand, tensor memory the other
                if (!slot[o.dst].in_regs && !slot[o.src].in_regs && !to_regs(o.dst, (int)k)) return false;
                const bool src_live = live_after((int)k, o.src);
                const int r = slot[o.dst].in_regs ? o.dst : o.src, other = r == o.dst ? o.src : o.dst;
                if (slot[r].alias) return false;  // cannot happen: aliases are never register-resident
                if (r == o.src && src_live && !stash(o.src)) return false;
                if (slot[other].tm < 0) return false;
                if (slot[r].orient != slot[other].orient) emit(blank(N_TRANSPOSE, slot[r].size, slot[r].size));
                NOp n = blank(N_ADD, slot[r].size, slot[r].size);
                n.slot = (short)slot[other].tm;
                // registers (pend_r x true) += pend_r x true(other); the stash holds pend_o x true(stash)
                n.scale = slot[r].pend / slot[other].pend * (slot[other].alias ? slot[other].a_scale : 1.f);
                n.bias = slot[other].alias ? slot[r].pend * slot[other].a_bias : 0.f;
                if (!(n.scale == n.scale) || std::fabs(n.scale) > 1e30f || std::fabs(n.bias) > 1e30f) return false;
                emit(n);
                Slot d = slot[r];
                d.orient = slot[other].orient; d.in_regs = true; d.tm = -1; d.valid = true; d.alias = false;
                // the sum lives in registers and belongs to dst; stale copies of dst go
                free_tm(o.dst);
                slot[o.src].in_regs = false;
                if (!src_live) { free_tm(o.src); slot[o.src].valid = false; }
                slot[o.dst] = d;
                continue;
            }
            // unary ops: CONV, RELU, COPY, SCALE
            {   // a pointwise op (1 x 1 stride-1 convolution, scaling, copy) whose input stays live is
                // not materialised: dst becomes an alias of the stashed input
                bool pointwise = o.opcode == CNNGP_OP_COPY || o.opcode == CNNGP_OP_SCALE;
                float a_scale = o.opcode == CNNGP_OP_SCALE ? o.scale_f : 1.f, a_bias = 0.f;
                if (o.opcode == CNNGP_OP_CONV && o.dil == 1 && o.stride == 1 && o.pad == o.t0 && o.ke - 1 == o.pad &&
                    o.Ho == o.Hi && o.Wo == o.Wi) {
                    pointwise = true; a_scale = o.scale_f; a_bias = o.bias_f;
                }
                Slot &src = slot[o.src];
                if (pointwise && o.dst != o.src && src.size > 1 && !src.alias && live_after((int)k, o.src) &&
                    (src.in_regs || src.tm >= 0)) {
                    if (src.tm < 0 && !stash(o.src)) return false;
                    if (src.in_regs) {  // the stash moves to dst; src keeps living in registers
                        free_tm(o.dst);
                        Slot d = src;
                        d.in_regs = false; d.alias = true; d.a_scale = a_scale; d.a_bias = a_bias;
                        tm_owner[src.tm] = o.dst;
                        src.tm = -1;
                        slot[o.dst] = d;
                        continue;
                    }
                }
            }
            if (!to_regs(o.src, (int)k)) return false;
            if (o.dst != o.src && live_after((int)k, o.src)) {
                if (!stash(o.src)) return false;
            }
            Slot src = slot[o.src];
            if (o.dst != o.src) {
                slot[o.src].in_regs = false;
                if (!live_after((int)k, o.src)) { free_tm(o.src); slot[o.src].valid = false; }
                free_tm(o.dst);
            } else {
                free_tm(o.src);  // in-place update: a stashed copy would be stale
            }
            Slot dst = src;
            dst.in_regs = true; dst.tm = -1; dst.valid = true;
            switch (o.opcode) {
                case CNNGP_OP_COPY: break;
                case CNNGP_OP_SCALE: {
                    const double np = (double)src.pend / (double)o.scale_f;  // carried: stored stays, the factor moves
                    if (carry && src.size > 1 && o.scale_f > 0.f && pend_ok(np)) dst.pend = (float)np;
                    else affine(src.size, o.scale_f, 0.f);
                    break;
                }
                case CNNGP_OP_RELU: {
                    if (o.Hi != o.Wi || o.Hi != src.size) return false;
                    NOp n = blank(src.size == 1 ? T_RELU : N_RELU, src.size, src.size);
                    if (src.size == 1) {
                        if (src.pend != 1.f) affine(1, 1.f / src.pend, 0.f);
                        n.aux = o.aux_off;
                        dst.pend = 2.f;
                    } else {
                        // relu_k(pend * m; sqrt(pend) s) = pend * relu_k(m; s): the layer's per-image s maps
                        // carry sqrt(pend) (applied once per image by cnngp_variances)
                        float pend = src.pend;
                        if (!carry || !pend_ok(pend)) {
                            if (pend != 1.f) affine(src.size, 1.f / pend, 0.f);
                            pend = 1.f;
                        }
                        n.aux = o.aux_foff; n.half = o.aux_half;
                        n.aux_scale = (float)std::sqrt((double)pend);
                        mutable_ops[k].aux_t = src.orient;
                        mutable_ops[k].aux_scale = n.aux_scale;
                        dst.pend = 2.f * pend;  // the kernel's ReLU output is doubled
                    }
                    emit(n);
                    break;
                }
                case CNNGP_OP_CONV: {
                    if (o.Hi != o.Wi || o.Ho != o.Wo || o.Hi != src.size) return false;
                    // taps t0 .. ke-1 at distance dil, the first at offset dil * t0 - pad: window [-lo, +hi] in taps
                    if (o.dil < 1 || (o.pad - o.dil * o.t0) % o.dil != 0) return false;
                    const int lo = (o.pad - o.dil * o.t0) / o.dil, hi = o.ke - 1 - o.t0 - lo;
                    if (lo < 0 || hi < 0) return false;
                    if (o.dil != 1 && (src.size == 1 || (lo == 0 && hi == 0))) return false;
                    NOp n = blank(N_CONV, src.size, o.Ho);
                    // true_out = scale_f * box(true_in) + bias_f; stored_in = pend * true_in
                    const double np = (double)src.pend / (double)o.scale_f;  // carried: stored_out = box(stored_in) + bias_f * np
                    bool carried = false;
                    if (src.size == 1) {
                        if (lo != 0 || hi != 0 || o.Ho != 1) return false;
                        n.kind = T_AFFINE;
                    } else if (o.Ho == 1 && lo == 0 && hi == src.size - 1 && o.dil == 1) {
                        n.kind = N_DENSE;
                    } else if (lo == 0 && hi == 0 && o.stride == 1) {
                        n.kind = N_AFFINE;
                        if (carry && o.scale_f > 0.f && pend_ok(np)) {
                            carried = true;
                            n.scale = 1.f; n.bias = (float)(o.bias_d * np);
                        }
                    } else {
                        if (o.stride != 1 && o.stride != 2) return false;
                        if (!conv_supported(S0, src.size, lo, hi, o.stride, o.dil)) return false;
                        if (o.Ho != (o.stride == 1 ? src.size : src.size / 2)) return false;
                        n.lo = (short)lo; n.hi = (short)hi; n.st = (short)o.stride; n.dil = (short)o.dil;
                        dst.orient = src.orient ^ 1;
                        if (carry && o.scale_f > 0.f && pend_ok(np)) {
                            carried = true;
                            n.scale = 1.f;
                            const float b = (float)(o.bias_d * np);
                            if (o.stride == 1) { n.pre_bias = b; n.bias = 0.f; }  // rides on the second sliding sum
                            else { n.pre_bias = 0.f; n.bias = b; }                // strided: explicit pass when non-zero
                        }
                    }
                    if (!carried) { n.scale = o.scale_f / src.pend; n.bias = o.bias_f; n.pre_bias = 0.f; }
                    if (n.kind == N_AFFINE && n.scale == 1.f && n.bias == 0.f) { /* nothing to do */ }
                    else emit(n);
                    dst.size = o.Ho;
                    dst.pend = carried ? (float)np : 1.f;
                    break;
                }
                default: return false;
            }
            slot[o.dst] = dst;
            if ((int)out.size() > kMaxNOps - 8) return false;
        }
        Slot &f = slot[final_slot];
        if (!f.valid || !f.in_regs || f.size != 1) return false;
        if (f.pend != 1.f) affine(1, 1.f / f.pend, 0.f);
        return (int)out.size() <= kMaxNOps;
    }
};

template <int S0, int NW, int NST, int NSPLIT, int PH = 3>
constexpr size_t fnet_smem() {
    constexpr int SF = S0 / 2;
    return (PH == 2 ? (size_t)NST * NGeo<NW>::kPairs * ((SF * SF + 1) / 2) * 32 + (size_t)NW * SF * (S0 + 1) * 8
                    : (size_t)NST * NGeo<NW>::kImgs * S0 * S0 * 4 / (NSPLIT / 2) + (size_t)NW * S0 * (S0 + 1) * 8) +
           (size_t)kMaxKOps * 16 + (size_t)3 * NST * 8 + 16;
}

const char *kNames[] = {"CONV", "AFFINE", "RELU", "STASH", "UNSTASH", "ADD", "TRANSPOSE", "DENSE", "T_RELU", "T_AFFINE"};

// dispatch case of one register-level op in phase A (full = false: folded map held in M[0]) or B
int case_of(const NOp &n, int S0, bool phase_b) {
    const int sc = n.si == S0 ? 0 : (n.si == S0 / 2 ? 1 : 2);  // size class of the op's input
    if (!phase_b) {
        switch (n.kind) {
            case N_CONV: {
                if (n.dil == 2) return A_CONV + 10;
                const int v = n.st == 1 ? (n.lo == 1 && n.hi == 1 ? 0 : n.lo == 1 && n.hi == 2 ? 1 : n.lo == 2 ? 2 : 3)
                                        : (n.lo == 1 ? 4 : 5);
                return A_CONV + (sc == 0 ? v : sc == 1 ? (n.st == 1 ? 6 : (n.lo == 1 ? 7 : 8)) : 9);
            }
            case N_AFFINE: return A_AFFINE + sc;
            case N_TRANSPOSE: return A_TRANSPOSE + sc;
            case N_STASH: return A_STASH + sc;
            case N_UNSTASH: return A_UNSTASH + sc;
            case N_ADD: return A_ADD + sc;
            case N_DENSE: return A_DENSE + sc;
            default: return A_RELU + sc;
        }
    }
    const int q = sc - 1;  // 0: S0/2, 1: S0/4
    switch (n.kind) {
        case N_CONV: return B_CONV + (sc == 1 ? (n.st == 1 ? 0 : (n.lo == 1 ? 1 : 2)) : 3);
        case N_AFFINE: return B_AFFINE + q;
        case N_TRANSPOSE: return B_TRANSPOSE + q;
        case N_STASH: return B_STASH + q;
        case N_UNSTASH: return B_UNSTASH + q;
        case N_ADD: return B_ADD + q;
        case N_DENSE: return B_DENSE + q;
        default: return B_RELU + q;
    }
}

KOp make_kop(const NOp &n, int code) {
    KOp k;
    k.code = code | ((n.slot & 3) << 8) | ((n.half & 0xffff) << 16);
    k.scale = n.scale; k.bias = n.bias;
    k.aux = n.aux;
    if (n.kind == N_CONV) memcpy(&k.aux, &n.pre_bias, 4);
    return k;
}

// ops[i..] starts an identity residual block  STASH RELU CONV(3x3, s1) RELU CONV(3x3, s1) ADD  on one map size
bool is_idblock(const NOp *o, int n_left) {
    if (n_left < 6) return false;
    const int s = o[0].si;
    auto conv33 = [&](const NOp &c) { return c.kind == N_CONV && c.si == s && c.so == s && c.st == 1 && c.lo == 1 && c.hi == 1 && c.dil == 1; };
    return o[0].kind == N_STASH && o[1].kind == N_RELU && o[1].si == s && conv33(o[2]) && o[3].kind == N_RELU && o[3].si == s &&
           conv33(o[4]) && o[5].kind == N_ADD && o[5].si == s && o[5].slot == o[0].slot;
}
// STASH CONV(4x4 "same") RELU TRANSPOSE ADD at full size (mnist_paper_residual_cnn_gp)
bool is_resblock(const NOp *o, int n_left, int S0) {
    if (n_left < 5) return false;
    return o[0].kind == N_STASH && o[0].si == S0 && o[1].kind == N_CONV && o[1].si == S0 && o[1].so == S0 && o[1].st == 1 &&
           o[1].lo == 1 && o[1].hi == 2 && o[1].dil == 1 && o[2].kind == N_RELU && o[2].si == S0 && o[3].kind == N_TRANSPOSE && o[3].si == S0 &&
           o[4].kind == N_ADD && o[4].si == S0 && o[4].slot == o[0].slot;
}

// register-level ops -> kernel descriptors: phase split, block grouping, sentinels, producer list
bool build_kops(FNetPlan *fp, bool blocks, bool blocks_a = true) {
    const int S0 = fp->S0, n = fp->n_ops;
    int first_tail = n;
    for (int k = 0; k < n; ++k)
        if (fp->ops[k].kind == T_RELU || fp->ops[k].kind == T_AFFINE) { first_tail = k; break; }
    for (int k = first_tail; k < n; ++k)
        if (fp->ops[k].kind != T_RELU && fp->ops[k].kind != T_AFFINE) return false;  // map ops after the pooling: not in the set
    int end_a = 0;  // one past the last op that works on a full-size map
    for (int k = 0; k < first_tail; ++k)
        if (fp->ops[k].si == S0) end_a = k + 1;
    int nk = 0;
    auto push = [&](KOp k) { if (nk < kMaxKOps) fp->kops[nk] = k; ++nk; };
    KOp sentinel{};
    sentinel.scale = 1.f;
    for (int k = 0; k < end_a;) {
        const NOp *o = fp->ops + k;
        int len = 1, code = case_of(*o, S0, false);
        if (blocks && blocks_a && o->si == S0 && is_idblock(o, end_a - k)) { len = 6; code = A_IDBLOCK; }
        else if (blocks && blocks_a && is_resblock(o, end_a - k, S0)) { len = 5; code = A_RESBLOCK; }
        if (len > 1) ++fp->n_blocks;
        push(make_kop(o[0], code));
        for (int q = 1; q < len; ++q) push(make_kop(o[q], case_of(o[q], S0, false)));
        k += len;
    }
    sentinel.code = A_END; push(sentinel);
    fp->k_b = nk;
    fp->folded_phase = end_a < first_tail;
    {   // can the program be cut at the end of phase A?  Phase B must stage at least one ReLU layer (a tile's
        // index travels with its first stage) and only tensor-memory slot 1 may carry a map across the cut
        int relu_b = 0;
        fp->n_relu_a = 0;
        for (int k = 0; k < first_tail; ++k)
            if (fp->ops[k].kind == N_RELU) (k < end_a ? fp->n_relu_a : relu_b) += 1;
        bool slot0_live = false;
        for (int k = end_a; k < first_tail; ++k) {
            const NOp &o = fp->ops[k];
            if ((o.kind == N_UNSTASH || o.kind == N_ADD) && o.slot == 0) { slot0_live = true; break; }
            if (o.kind == N_STASH && o.slot == 0) break;
        }
        fp->split = fp->folded_phase && relu_b > 0 && !slot0_live && end_a > 0;
    }
    for (int k = end_a; k < first_tail;) {
        const NOp *o = fp->ops + k;
        if (o->si == S0 || o->si == 1) return false;
        int len = 1, code = case_of(*o, S0, true);
        if (blocks && is_idblock(o, first_tail - k)) { len = 6; code = B_IDBLOCK + (o->si == S0 / 2 ? 0 : 1); }
        if (len > 1) ++fp->n_blocks;
        push(make_kop(o[0], code));
        for (int q = 1; q < len; ++q) push(make_kop(o[q], case_of(o[q], S0, true)));
        k += len;
    }
    sentinel.code = B_END; push(sentinel);
    for (int k = first_tail; k < n; ++k) push(make_kop(fp->ops[k], fp->ops[k].kind == T_RELU ? T_CASE_RELU : T_CASE_AFFINE));
    sentinel.code = T_END; push(sentinel);
    push(sentinel);  // the one-ahead fetch of the last op reads one descriptor further
    if (nk > kMaxKOps) return false;
    fp->n_kops = nk;
    fp->n_relu = 0;
    for (int k = 0; k < n; ++k) {
        if (fp->ops[k].kind != N_RELU) continue;
        if (fp->n_relu >= kMaxRelu) return false;
        fp->relu_aux[fp->n_relu] = fp->ops[k].aux;
        fp->relu_half[fp->n_relu] = fp->ops[k].half | (fp->ops[k].si == S0 ? (1 << 30) : 0);
        ++fp->n_relu;
    }
    return true;
}

}  // namespace

FNetPlan *fnet_plan_create(const Plan *plan_const) {
    Plan *plan = const_cast<Plan *>(plan_const);
    if (plan->dtype != CNNGP_F32) return nullptr;
    if (plan->H != plan->W || (plan->H != 28 && plan->H != 32)) return nullptr;
    if (plan->n_slots > 3) return nullptr;  // one map in registers + two tensor-memory slots
    const int S0 = plan->H;
    std::vector<DevOp> saved = plan->ops;
    // measurement aids: CNNGP_FNET_NOCARRY=1 applies every tap explicitly (one FMA pass per conv),
    // CNNGP_FNET_NOBLOCKS=1 dispatches every op on its own
    Translator tr(saved, plan->n_slots, S0, getenv("CNNGP_FNET_NOCARRY") == nullptr);
    if (!tr.run(plan->ops, plan->final_slot)) {
        plan->ops = saved;
        return nullptr;
    }
    // the scalar tail needs the pooled value: exactly one N_DENSE, and nothing map-shaped after it
    int n_dense = 0;
    for (const NOp &n : tr.out) n_dense += n.kind == N_DENSE;
    FNetPlan *fp = new FNetPlan();
    fp->S0 = S0;
    fp->n_ops = (int)tr.out.size();
    for (int k = 0; k < fp->n_ops; ++k) fp->ops[k] = tr.out[k];
    if (n_dense != 1 || !build_kops(fp, getenv("CNNGP_FNET_NOBLOCKS") == nullptr)) {
        delete fp;
        plan->ops = saved;
        return nullptr;
    }
    // twelve consumer warps when the second tensor-memory slot only ever holds folded maps of at most
    // half the edge (3 warps share a 512-column lane quadrant: 3 x 5 S0 columns)
    // ... and when the program has a folded phase: its short, latency-bound ops want three warps per
    // scheduler (mnist_as_tf: 75.8 M pairs/s with twelve warps, 72.8 M with eight), whereas full-size layers
    // alone run better on eight warps with 240 registers (mnist_paper_residual_cnn_gp: 123.8 M against 114.8 M)
    const bool twelve = tr.slot1_max <= S0 / 2 && (fp->folded_phase || getenv("CNNGP_FNET_12WARPS")) && !getenv("CNNGP_FNET_8WARPS");
    // three ring stages (1.5 layers of variance maps in flight): + 4 % over two on the straight-line kernel
    const bool deep = getenv("CNNGP_FNET_NST2") == nullptr;
    if (S0 == 28 && twelve && deep) { fp->nw = 12; fp->nst = 3; fp->smem = fnet_smem<28, 12, 3, 2>(); }
    else if (S0 == 28 && twelve) { fp->nw = 12; fp->nst = 2; fp->smem = fnet_smem<28, 12, 2, 2>(); }
    else if (S0 == 28) { fp->nw = 8; fp->nst = 4; fp->smem = fnet_smem<28, 8, 4, 2>(); }
    // 32 x 32: the maps alone are 128 registers per thread, which leaves a 160-register warp nothing to
    // work with (measured: 21 M pairs/s with twelve spilling warps against 62 M with eight) -- eight warps
    else { fp->nw = 8; fp->nst = 3; fp->smem = fnet_smem<32, 8, 3, 2>(); }
    // programs with a folded phase run as two launches (phase A on eight warps, phase B on sixteen) unless
    // CNNGP_FNET_NOSPLIT=1 asks for the single-launch kernels above
    if (getenv("CNNGP_FNET_NOSPLIT")) fp->split = false;
    if (fp->split) {
        // phase A of 28 x 28 programs on twelve warps (without the phase-B code and the output bookkeeping the
        // 160-register warps hardly spill: 207 ms against 214 ms on eight warps, mnist_as_tf at 6 000 images;
        // CNNGP_FNET_A8=1 asks for eight), full-size ops dispatched one by one (the block cases cost registers a
        // 160-register warp does not have: 1.35 against 1.23 ns per pair and block)
        if (S0 == 28 && tr.slot1_max <= S0 / 2 && !getenv("CNNGP_FNET_A8")) {
            fp->nw_a = 12;
            fp->n_blocks = 0;
            if (!build_kops(fp, getenv("CNNGP_FNET_NOBLOCKS") == nullptr, false)) { delete fp; plan->ops = saved; return nullptr; }
            fp->split = true;
        }
        fp->smem_a = S0 == 28 ? (fp->nw_a == 12 ? fnet_smem<28, 12, 3, 2, 1>() : fnet_smem<28, 8, 4, 2, 1>()) : fnet_smem<32, 8, 3, 2, 1>();
        fp->smem_b = S0 == 28 ? fnet_smem<28, 16, 4, 2, 2>() : fnet_smem<32, 16, 3, 2, 2>();
    }
    for (const DevOp &o : plan->ops)
        if (o.opcode == CNNGP_OP_RELU) fp->fused_row_floats += 4 * o.aux_half;
    return fp;
}

void fnet_plan_destroy(FNetPlan *fp) { delete fp; }

std::string fnet_plan_describe(const FNetPlan *fp) {
    std::string t = "fused_net S0=" + std::to_string(fp->S0) +
                    (fp->split ? " warps=" + std::to_string(fp->nw_a) + "+16 (two launches)" : " warps=" + std::to_string(fp->nw) + " stages=" + std::to_string(fp->nst)) +
                    " blocks=" + std::to_string(fp->n_blocks) + " :";
    for (int k = 0; k < fp->n_ops; ++k) {
        const NOp &o = fp->ops[k];
        t += std::string(" ") + kNames[o.kind] + "(" + std::to_string(o.si);
        if (o.kind == N_CONV) t += "," + std::to_string(o.lo) + "," + std::to_string(o.hi) + ",s" + std::to_string(o.st) +
                                   (o.dil != 1 ? ",d" + std::to_string(o.dil) : "");
        if (o.kind == N_STASH || o.kind == N_UNSTASH || o.kind == N_ADD) t += ",t" + std::to_string(o.slot);
        t += ")";
    }
    return t;
}

// one line per register-level op, every field, floats with nine significant digits (exact for float32)
std::string fnet_plan_dump(const FNetPlan *fp) {
    std::string t = "fused_net S0=" + std::to_string(fp->S0) + "\n";
    char line[320];
    for (int k = 0; k < fp->n_ops; ++k) {
        const NOp &o = fp->ops[k];
        snprintf(line, sizeof line,
                 "%s si=%d so=%d lo=%d hi=%d st=%d dil=%d slot=%d scale=%.9g bias=%.9g pre_bias=%.9g aux_scale=%.9g aux=%d half=%d\n",
                 kNames[o.kind], (int)o.si, (int)o.so, (int)o.lo, (int)o.hi, (int)o.st, (int)o.dil, (int)o.slot, (double)o.scale,
                 (double)o.bias, (double)o.pre_bias, (double)o.aux_scale, o.aux, o.half);
        t += line;
    }
    return t;
}

namespace {

// tile grid of one kernel geometry (tiles of 4 x tile_j images, super-tiles of `edge` images)
void fnet_geometry(NParams &p, int64_t N1, int64_t N2, int tile_j, int edge, long long *n_super) {
    const int kTileI = 4;
    p.nbi = (int)((N1 + kTileI - 1) / kTileI);
    p.nbj = (int)((N2 + tile_j - 1) / tile_j);
    const int super_i = edge / kTileI, super_j = edge / tile_j;
    if (p.nbi <= super_i && p.nbj <= super_j) {
        p.sti = p.nbi; p.stj = p.nbj; p.nst_j = 1; p.nst = 1;
        *n_super = 1;
    } else {
        p.sti = super_i; p.stj = super_j;
        const int nsi = (p.nbi + super_i - 1) / super_i, nsj = (p.nbj + super_j - 1) / super_j;
        p.nst_j = nsj;
        // symmetric: super-row r holds the super-tiles (r, r .. nsj - 1); a band of block rows (N2 > N1) has fewer
        // super-rows than super-columns
        p.nst = p.symmetric ? nsj : (nsi > nsj ? nsi : nsj);
        *n_super = p.symmetric ? (long long)nsi * nsj - (long long)nsi * (nsi - 1) / 2 : (long long)nsi * nsj;
    }
}

// what the kernel will count per super-row: valid tiles x consumer warps (see gram_fused.cu)
int fnet_progress(NParams &p, RowProgress *prog, int tile_j, int nw) {
    const int kTileI = 4;
    prog->n_super_rows = (p.nbi + p.sti - 1) / p.sti;
    prog->rows_per_super = (int64_t)p.sti * kTileI;
    if (!p.symmetric || p.N1 != p.N2 || prog->n_super_rows > prog->capacity) { set_error("fused-net kernel: progress counters too few"); return 8; }
    prog->expected.assign(prog->n_super_rows, 0u);
    for (int ib = 0; ib < p.nbi; ++ib) {
        const int si = ib / p.sti;
        long long jb_lo = (long long)si * p.stj;
        const long long need = ((long long)ib * kTileI) / tile_j;  // first tile that reaches the diagonal
        if (need > jb_lo) jb_lo = need;
        if (jb_lo < p.nbj) prog->expected[si] += (unsigned)((p.nbj - jb_lo) * nw);
    }
    p.row_done = prog->d_done;
    return 0;
}

int fnet_launch_one(void (*kern)(const NParams), NParams &p, unsigned threads, size_t smem, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long todo = p.n_tiles - p.t_begin;
    const unsigned grid = (unsigned)(todo < sms ? todo : sms);
    cudaError_t e = cudaSuccess;
    const char *order = getenv("CNNGP_TILE_ORDER");  // "static": fixed stride instead of the counter
    if (order && !strcmp(order, "static")) {
        p.tile_ctr = nullptr;
    } else {
        p.tile_ctr = tile_counter_for(stream);
        if (!p.tile_ctr) return 7;
        e = cudaMemsetAsync(p.tile_ctr, 0, sizeof(unsigned long long), stream);
    }
    if (e != cudaSuccess) { set_error(std::string("fused-net tile counter: ") + cudaGetErrorString(e)); return 7; }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error(std::string("fused-net cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 7; }
    kern<<<grid, threads, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("fused-net kernel launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

}  // namespace

int launch_fnet_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2, int32_t C,
                     const void *d_aux_x, const void *d_aux_z, int32_t symmetric, const void *d_kdiag,
                     void *d_out, int64_t ld_out, void *stream_, RowProgress *prog, int64_t mirror_block) {
    const FNetPlan *fp = plan->fnet;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!fp) { set_error("fused-net kernel: unsupported call"); return 4; }
    if (N1 > 2000000000LL || N2 > 2000000000LL) { set_error("fused-net kernel: too many images"); return 8; }
    NParams p;  // ~4.5 KB, passed by value at launch
    memset(&p, 0, sizeof p);
    memcpy(p.ops, fp->kops, sizeof(KOp) * fp->n_kops);
    p.n_ops = fp->n_kops;
    memcpy(p.relu_aux, fp->relu_aux, sizeof(int) * fp->n_relu);
    memcpy(p.relu_half, fp->relu_half, sizeof(int) * fp->n_relu);
    p.n_relu = fp->n_relu;
    p.n_relu_a = fp->n_relu_a;
    p.k_b = fp->k_b;
    p.x = (const float *)d_x; p.z = (const float *)d_z;
    p.aux_x = (const float *)d_aux_x; p.aux_z = (const float *)d_aux_z;
    p.aux_stride = plan->aux_elems; p.aux_f_off = plan->aux_f_off;
    p.N1 = (int)N1; p.N2 = (int)N2; p.C = C;
    p.out = (float *)d_out; p.ld_out = ld_out;
    p.symmetric = symmetric ? 1 : 0;
    if (symmetric && N2 < N1) { set_error("fused-net kernel: a symmetric band needs N2 >= N1"); return 4; }
    p.mirror_bs = (int)mirror_block;
    p.kdiag = (const float *)d_kdiag;
    p.inv_c = 1.0f / (float)C;
    const size_t row_bytes = (size_t)fp->fused_row_floats * 4;
    if (!fp->split) {
        // super-tiles of side `edge` images: the 2 * edge variance rows a wave of CTAs shares stay in L2
        int edge = 504;  // divisible by 4, 8 and 12
        while (edge > 72 && (size_t)2 * edge * row_bytes > ((size_t)72 << 20)) edge = edge / 2 / 24 * 24;
        if (const char *e = getenv("CNNGP_SUPER_EDGE")) { const int v = atoi(e); if (v >= 24 && v % 24 == 0) edge = v; }
        long long n_super;
        fnet_geometry(p, N1, N2, fp->nw, edge, &n_super);
        p.n_tiles = n_super * p.sti * p.stj;
        if (prog) { const int rc = fnet_progress(p, prog, fp->nw, fp->nw); if (rc) return rc; }
        void (*kern)(const NParams) = fp->S0 == 32 ? fnet_kernel<32, 8, 3, 2>
                                                   : (fp->nw == 12 ? (fp->nst == 3 ? fnet_kernel<28, 12, 3, 2> : fnet_kernel<28, 12, 2, 2>)
                                                                   : fnet_kernel<28, 8, 4, 2>);
        return fnet_launch_one(kern, p, (fp->nw + 4) * 32, fp->smem, stream);
    }
    // ---- two launches per chunk of super-tiles: phase A (4 x 8 tiles, eight warps), phase B (4 x 16, sixteen) ----
    int edge = 480;  // divisible by 4, 8 and 16; both launches see the same super-tiles and 2 x 2 blocks
    while (edge > 48 && (size_t)2 * edge * row_bytes > ((size_t)72 << 20)) edge = edge > 96 ? edge / 2 / 48 * 48 : 48;
    if (const char *e = getenv("CNNGP_SUPER_EDGE")) { const int v = atoi(e); if (v >= 48 && v % 48 == 0) edge = v; }
    NParams pa = p, pb = p;
    long long n_super = 0, n_super_b = 0;
    const int nw_a = fp->nw_a;
    fnet_geometry(pa, N1, N2, nw_a, edge, &n_super);
    fnet_geometry(pb, N1, N2, 16, edge, &n_super_b);
    if (n_super != n_super_b) { set_error("fused-net kernel: the two launches disagree on the super-tiles"); return 9; }
    if (prog) { const int rc = fnet_progress(pb, prog, 16, 16); if (rc) return rc; }
    const int blocks_i = pa.sti * 2, blocks_j = pa.stj * (nw_a / 2) > pb.stj * 8 ? pa.stj * (nw_a / 2) : pb.stj * 8;
    const size_t rec_bytes = (size_t)2 * (fp->S0 / 2) * 32 * 8;
    const size_t st_bytes = (size_t)blocks_i * blocks_j * rec_bytes;
    size_t budget = (size_t)2 << 30;  // hand-off buffer per chunk (1 GB: 201.4 ms, 2 GB: 200.3 ms, 128 MB: 230 ms per 6000 x 6000 mnist_as_tf Gram)
    if (const char *e = getenv("CNNGP_FNET_HANDOFF_MB")) { const long v = atol(e); if (v > 0) budget = (size_t)v << 20; }
    long long per_chunk = (long long)(budget / st_bytes);
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > n_super) per_chunk = n_super;
    void *handoff = nullptr;
    pool_keep(budget);  // the hand-off buffer stays mapped between calls
    cudaError_t e = cudaMallocAsync(&handoff, (size_t)per_chunk * st_bytes, stream);
    if (e != cudaSuccess) { set_error(std::string("fused-net hand-off buffer: ") + cudaGetErrorString(e)); return 7; }
    pa.handoff = pb.handoff = (unsigned long long *)handoff;
    pa.rec_ld = pb.rec_ld = blocks_j;
    pa.rec_per_st = pb.rec_per_st = blocks_i * blocks_j;
    pa.row_done = nullptr;
    void (*kern_a)(const NParams) = fp->S0 == 32 ? fnet_kernel<32, 8, 3, 2, 1>
                                                 : (nw_a == 12 ? fnet_kernel<28, 12, 3, 2, 1> : fnet_kernel<28, 8, 4, 2, 1>);
    void (*kern_b)(const NParams) = fp->S0 == 32 ? fnet_kernel<32, 16, 3, 2, 2> : fnet_kernel<28, 16, 4, 2, 2>;
    int rc = 0;
    for (long long s0 = 0; s0 < n_super && !rc; s0 += per_chunk) {
        const long long s1 = s0 + per_chunk < n_super ? s0 + per_chunk : n_super;
        pa.st_begin = pb.st_begin = (int)s0;
        pa.t_begin = s0 * pa.sti * pa.stj; pa.n_tiles = s1 * pa.sti * pa.stj;
        pb.t_begin = s0 * pb.sti * pb.stj; pb.n_tiles = s1 * pb.sti * pb.stj;
        rc = fnet_launch_one(kern_a, pa, (nw_a + 4) * 32, fp->smem_a, stream);
        if (!rc) rc = fnet_launch_one(kern_b, pb, (16 + 4) * 32, fp->smem_b, stream);
    }
    cudaFreeAsync(handoff, stream);
    note_launches((int)(2 * ((n_super + per_chunk - 1) / per_chunk)));
    return rc;
}

}  // namespace cnngp
