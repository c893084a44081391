// gram_variance.cu -- per-image variance rows of straight-line 28 x 28 programs at HBM rate.
//
// cnngp_variances is O(N) work next to the O(N^2) Gram, but it is replicated on every GPU and sits in
// front of every Gram call.  The generic interpreter (gram_generic.cu, MODE 1) walks the program with
// all maps in shared memory and writes the fused kernel's operands four bytes at a time: 0.77 ms per
// 10 000 images of the headline program, 0.86 TB/s.  Here one warp owns an image PAIR (2k, 2k+1) --
// the unit the fused kernel's float4 (s_2k, s_2k+1, 1/s_2k, 1/s_2k+1) operands interleave -- and keeps
// both variance maps in registers as one packed array (lane = one map coordinate, register = the
// other), so the recursion is register arithmetic plus two transpositions per layer through the
// warp's private shared-memory tile, and every operand row leaves as whole 16-byte stores.
//
// Bit-identical to the generic kernel by construction: the same IEEE operations in the same order
// (direct window sums over ascending taps, rows first, then columns; separate mul / add for
// tap * sum + bias; sqrt.rn / div.rn for the operands), only packed two images at a time
// (add.rn.f32x2 / mul.rn.f32x2 round each half like the scalar instruction).
//
// Reference semantics (paths relative to /root/reference): cnn_gp/kernels.py:48-49 (xx = mean_c x^2),
// :98 (Conv2d acts on xx as on xy), :154,164 (ReLU halves xx), :155-158 (diag value).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "fused_common.cuh"
#include "plan.h"

namespace cnngp {

namespace {

using namespace fusedk;

constexpr int S = 28, P = S * S, PITCH = S + 1;
constexpr int kVarWarps = 4;  // 4 x 6.5 KB of transposition tiles: eight CTAs per SM by shared memory
constexpr int kVarMaxOps = 48;

enum { V_CONV = 0, V_RELU = 1, V_DENSE = 2 };

struct VOp {
    int kind;
    int lo, hi;          // V_CONV: window offsets [-lo, +hi] (0, 0: pointwise)
    float scale, bias;   // V_CONV / V_DENSE
    int aux_off;         // V_RELU: offset of the plain map in a row
    int aux_foff;        // V_RELU: offset of the (s, 1/s) section (floats, relative to aux_f_off)
    int aux_t;           // V_RELU: operands stored transposed
    float aux_scale;     // V_RELU: factor on s
};

struct VParams {
    VOp ops[kVarMaxOps];
    int n_ops;
    const float *x;
    long long N, n_pairs;
    int C;
    float *aux;
    long long aux_elems;
    int aux_f_off;
    float *kdiag;
};

// out[y] = sum over ascending taps t = -LO .. HI of v[y + t], zero padded: what the generic kernel's
// "acc = 0; for t: if in range acc += src" computes (0 + v is exact)
template <int LO, int HI>
__device__ __forceinline__ void box_direct(u64 (&v)[S]) {
    u64 o[S];
#pragma unroll
    for (int y = 0; y < S; ++y) {
        const int first = y - LO < 0 ? 0 : y - LO;
        u64 acc = v[first];
#pragma unroll
        for (int t = first + 1; t <= y + HI && t < S; ++t) acc = add2(acc, v[t]);
        o[y] = acc;
    }
#pragma unroll
    for (int y = 0; y < S; ++y) v[y] = o[y];
}

__device__ __forceinline__ void box_window(u64 (&v)[S], int lo, int hi) {
    if (lo == 3 && hi == 3) box_direct<3, 3>(v);
    else if (lo == 1 && hi == 1) box_direct<1, 1>(v);
    else if (lo == 1 && hi == 2) box_direct<1, 2>(v);
    else box_direct<2, 2>(v);
}

__device__ __forceinline__ void transpose(u64 *tile, u64 (&v)[S], int lane, int lx) {
    const int col = lane < S ? lane : S;  // idle lanes write the pad column
#pragma unroll
    for (int r = 0; r < S; ++r) tile[r * PITCH + col] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < S; ++r) v[r] = tile[lx * PITCH + r];
    __syncwarp();
}

__global__ void __launch_bounds__(kVarWarps * 32) variance_kernel(const __grid_constant__ VParams p) {
    __shared__ __align__(16) u64 tiles[kVarWarps][S * PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = lane < S ? lane : S - 1;
    u64 *tile = tiles[warp];

    for (long long pr = (long long)blockIdx.x * kVarWarps + warp; pr < p.n_pairs; pr += (long long)gridDim.x * kVarWarps) {
        const long long n0 = 2 * pr;
        const bool two = n0 + 1 < p.N;  // odd N: the last pair's second image is a copy of the first
        const float *a0 = p.x + n0 * p.C * P, *a1 = p.x + (two ? n0 + 1 : n0) * p.C * P;
        float *row0 = p.aux + n0 * p.aux_elems, *row1 = row0 + p.aux_elems;

        // init, kernels.py:48-49: lane = column, register = row ("natural")
        u64 M[S];
        {
            const float cnt = (float)p.C;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                float s0 = 0.f, s1 = 0.f;
                for (int c = 0; c < p.C; ++c) {
                    const float u = a0[c * P + r * S + lx], w = a1[c * P + r * S + lx];
                    s0 += u * u;
                    s1 += w * w;
                }
                M[r] = pk(__fdiv_rn(s0, cnt), __fdiv_rn(s1, cnt));
            }
        }
        bool natural = true;  // lane = column; false: lane = row, register = column

        for (int k = 0; k < p.n_ops; ++k) {
            const VOp o = p.ops[k];
            if (o.kind == V_CONV) {
                if (o.lo || o.hi) {
                    // rows first (sum over x: the register axis when lane = row), then columns
                    if (natural) transpose(tile, M, lane, lx);
                    box_window(M, o.lo, o.hi);
                    transpose(tile, M, lane, lx);
                    box_window(M, o.lo, o.hi);
                    natural = true;
                }
                // scalar mul.rn / add.rn: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (the explicit
                // rounding modifier protects only the scalar forms), and the interpreter rounds twice
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    float v0, v1;
                    upk(M[r], v0, v1);
                    M[r] = pk(__fadd_rn(__fmul_rn(v0, o.scale), o.bias), __fadd_rn(__fmul_rn(v1, o.scale), o.bias));
                }
            } else if (o.kind == V_RELU) {
                if (!natural) { transpose(tile, M, lane, lx); natural = true; }
                if (lane < S) {
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        float v0, v1;
                        upk(M[r], v0, v1);
                        row0[o.aux_off + r * S + lane] = v0;
                        if (two) row1[o.aux_off + r * S + lane] = v1;
                    }
                }
                // the fused kernel's operands in the layout it is in at this layer
                if (o.aux_t) { transpose(tile, M, lane, lx); natural = false; }
                if (p.aux_f_off > 0 && lane < S) {
                    float *f0 = row0 + p.aux_f_off + o.aux_foff, *f1 = row1 + p.aux_f_off + o.aux_foff;
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        float v0, v1;
                        upk(M[r], v0, v1);
                        const float ss0 = __fmul_rn(__fadd_rn(__fsqrt_rn(v0), 1.0842021724855044e-19f), o.aux_scale);
                        const float ss1 = __fmul_rn(__fadd_rn(__fsqrt_rn(v1), 1.0842021724855044e-19f), o.aux_scale);
                        const float4 q = make_float4(ss0, ss1, __fdiv_rn(1.f, ss0), __fdiv_rn(1.f, ss1));
                        // pixels [0, P/2) of the layer live in row 2k, the rest in row 2k+1
                        float *dst = r < S / 2 ? f0 + 4 * (r * S + lane) : f1 + 4 * ((r - S / 2) * S + lane);
                        *reinterpret_cast<float4 *>(dst) = q;
                    }
                }
                const u64 HALF = pk(0.5f, 0.5f);  // kernels.py:154 (x / 2 == x * 0.5 exactly)
#pragma unroll
                for (int r = 0; r < S; ++r) M[r] = mul2(M[r], HALF);
            } else {  // V_DENSE: the S x S window without padding, one output
                if (natural) { transpose(tile, M, lane, lx); natural = false; }
                u64 acc = M[0];
#pragma unroll
                for (int r = 1; r < S; ++r) acc = add2(acc, M[r]);
                float r0, r1;
                upk(acc, r0, r1);
                float t0 = __shfl_sync(0xffffffffu, r0, 0), t1 = __shfl_sync(0xffffffffu, r1, 0);
#pragma unroll
                for (int y = 1; y < S; ++y) {
                    t0 = __fadd_rn(t0, __shfl_sync(0xffffffffu, r0, y));
                    t1 = __fadd_rn(t1, __shfl_sync(0xffffffffu, r1, y));
                }
                if (p.kdiag && lane == 0) {
                    p.kdiag[n0] = __fadd_rn(__fmul_rn(t0, o.scale), o.bias);
                    if (two) p.kdiag[n0 + 1] = __fadd_rn(__fmul_rn(t1, o.scale), o.bias);
                }
            }
        }
    }
}

}  // namespace

// Straight-line 28 x 28 float32 programs (the fused kernel's set); -1: not covered, the caller uses
// the generic kernel.
int launch_fused_variances(const Plan *plan, const void *d_x, int64_t N, int32_t C, void *d_aux_x,
                           void *d_kdiag, void *stream) {
    if (!plan->fused || plan->dtype != CNNGP_F32 || plan->H != S || plan->W != S) return -1;
    if (getenv("CNNGP_VARIANCE_GENERIC")) return -1;  // measurement / test aid
    if (plan->ops.size() > (size_t)kVarMaxOps) return -1;
    VParams p{};
    int cur = 0;
    bool done = false;
    for (const DevOp &o : plan->ops) {
        if (done || o.src != cur) return -1;
        VOp v{};
        if (o.opcode == CNNGP_OP_CONV) {
            if (o.dil != 1 || o.stride != 1 || o.Hi != S || o.Wi != S) return -1;
            if (o.Ho == S && o.Wo == S) {
                v.kind = V_CONV;
                v.lo = o.pad - o.t0;
                v.hi = o.ke - 1 - o.pad;
                const bool known = (v.lo == 0 && v.hi == 0) || (v.lo == 3 && v.hi == 3) || (v.lo == 1 && v.hi == 1) ||
                                   (v.lo == 1 && v.hi == 2) || (v.lo == 2 && v.hi == 2);
                if (!known) return -1;
            } else if (o.Ho == 1 && o.Wo == 1 && o.pad == 0 && o.t0 == 0 && o.ke == S) {
                v.kind = V_DENSE;
                done = true;
            } else {
                return -1;
            }
            v.scale = o.scale_f;
            v.bias = o.bias_f;
        } else if (o.opcode == CNNGP_OP_RELU) {
            if (o.Hi != S || o.Wi != S || o.aux_half != P / 2) return -1;
            v.kind = V_RELU;
            v.aux_off = o.aux_off;
            v.aux_foff = o.aux_foff;
            v.aux_t = o.aux_t;
            v.aux_scale = o.aux_scale;
        } else {
            return -1;
        }
        cur = o.dst;
        p.ops[p.n_ops++] = v;
    }
    if (!done || cur != plan->final_slot) return -1;
    p.x = (const float *)d_x;
    p.N = N;
    p.n_pairs = (N + 1) / 2;
    p.C = C;
    p.aux = (float *)d_aux_x;
    p.aux_elems = plan->aux_elems;
    p.aux_f_off = plan->aux_f_off;
    p.kdiag = (float *)d_kdiag;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, variance_kernel, kVarWarps * 32, 0);
    if (per_sm < 1) per_sm = 1;
    // a whole number of waves: every resident warp gets the same number of pairs (to within one)
    const long long ctas = (p.n_pairs + kVarWarps - 1) / kVarWarps, resident = (long long)sms * per_sm;
    const unsigned grid = (unsigned)(ctas < resident ? ctas : resident);
    variance_kernel<<<grid, kVarWarps * 32, 0, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("variance kernel launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

}  // namespace cnngp
