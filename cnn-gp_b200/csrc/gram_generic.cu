// gram_generic.cu -- the general Gram-recursion engine: any layer program, any map size that
// fits shared memory, float32 and float64.  One CTA evaluates G entries (image pairs, or
// image variance planes) with every map resident in shared memory; nothing but the final
// 1x1 value (and, in variance mode, the per-ReLU variance maps) is written to HBM.
//
// It is the first correct CUDA path and the float64 / odd-shape route; the register-resident
// kernel in gram_fused.cu takes over for the shapes it covers.
//
// Reference semantics restated here (paths relative to /root/reference):
//   init        cnn_gp/kernels.py:43-49    xy = mean_c x*y ; xx = mean_c x^2
//   conv        cnn_gp/kernels.py:92-98    box kernel of taps var_weight/k^2 (+ zero first row/col
//                                          for even "same" kernels, :73-84), zero padding, + var_bias
//   relu        cnn_gp/kernels.py:146-164  literal op order, f32_tiny included
//   add/scale   cnn_gp/kernel_patch.py:43-63 (Sum / Mixture element-wise combination)
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <string>

#include "plan.h"

namespace cnngp {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxG = 16;

// ---- arithmetic that must round where the reference's separate tensor ops round ----------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float acos_t(float a) { return acosf(a); }
__device__ __forceinline__ double acos_t(double a) { return acos(a); }

// idx / d for 0 <= idx < 2^22 and 1 <= d <= idx_max without an integer division (a runtime divisor costs
// ~20 instructions, and the interpreter's loops did two to four of them per map element): (idx + 0.5) / d
// is at least 0.5 / d away from an integer, far more than the float32 rounding of the product
__device__ __forceinline__ int fdiv(int idx, float inv_d) { return (int)(((float)idx + 0.5f) * inv_d); }

// kernels.py:146-152, one pixel.
template <typename T>
__device__ __forceinline__ T relu_literal(T c, T vx, T vy) {
    const T tiny = (T)FLT_MIN;  // np.finfo(np.float32).tiny, also in float64 mode (kernels.py:133)
    const T pi = (T)3.14159265358979323846;
    const T two_pi = (T)(2.0 * 3.14159265358979323846);
    T m = add_rn(mul_rn(vx, vy), tiny);
    T cs = mul_rn(c, div_rn((T)1, sqrt_rn(m)));
    cs = cs < (T)-1 ? (T)-1 : (cs > (T)1 ? (T)1 : cs);
    T d = sub_rn(m, mul_rn(c, c));
    d = d < (T)0 ? (T)0 : d;
    T sn = sqrt_rn(d);
    T th = acos_t(cs);
    T t = add_rn(sn, mul_rn(sub_rn(pi, th), c));
    return div_rn(t, two_pi);
}

template <typename T>
struct GenericParams {
    const DevOp *ops;
    int n_ops, n_slots, max_map, G, NP;
    int H, W, C;
    const T *x, *z;
    int64_t N1, N2, Q;  // Q = number of entries
    const T *aux_x, *aux_z;
    int64_t aux_elems;
    int aux_f_off;  // variance mode: where the fused (s, 1/s) maps start in a row (0 = none)
    T *out;
    int64_t ld_out;
    int same, diag, sym, final_slot;
    T *aux_x_out, *aux_z_out, *kdiag;  // variance mode
};

// upper-triangular (incl. diagonal) entry q of an N x N matrix -> (i, j), row-major
__device__ __forceinline__ void tri_decode(int64_t q, int64_t N, int64_t &i, int64_t &j) {
    double b = 2.0 * (double)N + 1.0;
    int64_t r = (int64_t)floor((b - sqrt(b * b - 8.0 * (double)q)) * 0.5);
    if (r < 0) r = 0;
    if (r > N - 1) r = N - 1;
    auto off = [N](int64_t a) { return a * N - (a * (a - 1)) / 2; };
    while (off(r) > q) --r;
    while (r + 1 <= N - 1 && off(r + 1) <= q) ++r;
    i = r;
    j = r + (q - off(r));
}

// MODE 0: image pairs.  MODE 1: per-image variance planes.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) generic_kernel(GenericParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    const int G = p.G, MM = p.max_map;
    // slot s of entry g: smem + (s*G + g)*MM ; scratch of entry g: smem + (n_slots*G + g)*MM
    auto slot = [&](int s, int g) -> T * { return smem + ((size_t)s * G + g) * MM; };
    T *scratch0 = smem + (size_t)p.n_slots * G * MM;
    __shared__ int64_t ei[kMaxG], ej[kMaxG];
    __shared__ int evalid[kMaxG];
    const int tid = threadIdx.x;

    if (tid < G) {
        int64_t q = (int64_t)blockIdx.x * G + tid;
        int valid = q < p.Q;
        int64_t i = 0, j = 0;
        if (valid) {
            if (MODE == 1) { i = q / p.NP; j = q % p.NP; }      // image, plane
            else if (p.diag) { i = q; j = q; }
            else if (p.sym) tri_decode(q, p.N1, i, j);
            else { i = q / p.N2; j = q % p.N2; }
        }
        ei[tid] = i; ej[tid] = j; evalid[tid] = valid;
    }
    __syncthreads();

    // ---- init, kernels.py:43-49 -----------------------------------------------------------
    {
        const int P = p.H * p.W;
        const T cnt = (T)p.C;
        const float inv_P = 1.0f / (float)P;
        for (int idx = tid; idx < G * P; idx += kThreads) {
            const int g = fdiv(idx, inv_P), px = idx - g * P;
            T s = (T)0;
            if (evalid[g]) {
                const T *a, *b;
                if (MODE == 1) {
                    const T *img = (ej[g] == 0 ? p.x : p.z) + (size_t)ei[g] * p.C * P;
                    a = img; b = img;
                } else {
                    a = p.x + (size_t)ei[g] * p.C * P;
                    b = p.z + (size_t)ej[g] * p.C * P;
                }
                for (int c = 0; c < p.C; ++c) s += a[(size_t)c * P + px] * b[(size_t)c * P + px];
                s = div_rn(s, cnt);
            }
            slot(0, g)[px] = s;
        }
    }
    __syncthreads();

    for (int k = 0; k < p.n_ops; ++k) {
        const DevOp o = p.ops[k];
        switch (o.opcode) {
            case CNNGP_OP_CONV: {
                // separable box sum: rows first (into scratch), then columns
                const int n1 = o.Hi * o.Wo;
                const float inv_n1 = 1.0f / (float)n1, inv_Wo = 1.0f / (float)o.Wo;
                for (int idx = tid; idx < G * n1; idx += kThreads) {
                    const int g = fdiv(idx, inv_n1), r = idx - g * n1;
                    const int y = fdiv(r, inv_Wo), xo = r - y * o.Wo;
                    const T *src = slot(o.src, g) + y * o.Wi;
                    T acc = (T)0;
                    int xi = xo * o.stride - o.pad + o.dil * o.t0;
                    for (int t = o.t0; t < o.ke; ++t, xi += o.dil)
                        if (xi >= 0 && xi < o.Wi) acc += src[xi];
                    scratch0[(size_t)g * MM + r] = acc;
                }
                __syncthreads();
                const int n2 = o.Ho * o.Wo;
                const T scale = sizeof(T) == 4 ? (T)o.scale_f : (T)o.scale_d;
                const T bias = sizeof(T) == 4 ? (T)o.bias_f : (T)o.bias_d;
                const float inv_n2 = 1.0f / (float)n2;
                for (int idx = tid; idx < G * n2; idx += kThreads) {
                    const int g = fdiv(idx, inv_n2), r = idx - g * n2;
                    const int yo = fdiv(r, inv_Wo), xo = r - yo * o.Wo;
                    const T *src = scratch0 + (size_t)g * MM + xo;
                    T acc = (T)0;
                    int yi = yo * o.stride - o.pad + o.dil * o.t0;
                    for (int t = o.t0; t < o.ke; ++t, yi += o.dil)
                        if (yi >= 0 && yi < o.Hi) acc += src[yi * o.Wo];
                    slot(o.dst, g)[r] = add_rn(mul_rn(acc, scale), bias);
                }
                break;
            }
            case CNNGP_OP_RELU: {
                const int P = o.Hi * o.Wi;
                const float inv_P = 1.0f / (float)P, inv_Wi = 1.0f / (float)o.Wi;
                if (MODE == 0) {
                    for (int idx = tid; idx < G * P; idx += kThreads) {
                        const int g = fdiv(idx, inv_P), px = idx - g * P;
                        T r = (T)0;
                        if (evalid[g]) {
                            const T c = slot(o.src, g)[px];
                            const T vx = p.aux_x[(size_t)ei[g] * p.aux_elems + o.aux_off + px];
                            if (p.same && ei[g] == ej[g]) {
                                r = div_rn(vx, (T)2);  // kernels.py:155-162
                            } else {
                                const T vy = p.aux_z[(size_t)ej[g] * p.aux_elems + o.aux_off + px];
                                r = relu_literal(c, vx, vy);
                            }
                        }
                        slot(o.dst, g)[px] = r;
                    }
                } else {
                    const int GE = G / p.NP;  // images in this CTA
                    for (int idx = tid; idx < GE * P; idx += kThreads) {
                        const int ge = fdiv(idx, inv_P), px = idx - ge * P;
                        const int g0 = ge * p.NP;
                        if (!evalid[g0]) continue;
                        const size_t arow = (size_t)ei[g0] * p.aux_elems + o.aux_off + px;
                        const T v0 = slot(o.src, g0)[px];
                        p.aux_x_out[arow] = v0;
                        if (p.aux_f_off > 0) {
                            // operands of the fused kernel's ReLU: s = sqrt(xx) (+ sqrt(f32_tiny), the
                            // separable stand-in for kernels.py:146's "+ f32_tiny") and 1/s, in the
                            // register layout that kernel is in at this layer, interleaved with the
                            // partner image of the pair (2k, 2k+1): float4 (s_2k, s_2k+1, 1/s_2k, 1/s_2k+1)
                            // per pixel, first half of the pixels in row 2k, second half in row 2k+1
                            const int py = fdiv(px, inv_Wi);
                            const int tp = o.aux_t ? (px - py * o.Wi) * o.Hi + py : px;
                            const int half = o.aux_half;
                            const int hi = tp >= half ? 1 : 0;
                            const int64_t n = ei[g0];
                            const T sd = add_rn(sqrt_rn(v0), (T)1.0842021724855044e-19);
                            T *f = p.aux_x_out + (size_t)((n & ~(int64_t)1) + hi) * p.aux_elems + p.aux_f_off +
                                   (size_t)o.aux_foff + 4 * (size_t)(tp - hi * half) + (n & 1);
                            const T ss = mul_rn(sd, (T)o.aux_scale);  // 1 unless the fused kernel folds conv taps
                            f[0] = ss;
                            f[2] = div_rn((T)1, ss);
                        }
                        const T h0 = div_rn(v0, (T)2);  // kernels.py:154
                        if (p.NP == 2) {
                            p.aux_z_out[arow] = slot(o.src, g0 + 1)[px];
                            slot(o.dst, g0 + 1)[px] = h0;  // same: yy = xx (kernels.py:155-156)
                        }
                        slot(o.dst, g0)[px] = h0;
                    }
                }
                break;
            }
            case CNNGP_OP_COPY: {
                const int P = o.Hi * o.Wi;
                for (int idx = tid; idx < G * P; idx += kThreads) {
                    const int g = fdiv(idx, 1.0f / (float)P), px = idx - g * P;
                    slot(o.dst, g)[px] = slot(o.src, g)[px];
                }
                break;
            }
            case CNNGP_OP_ADD: {
                const int P = o.Hi * o.Wi;
                for (int idx = tid; idx < G * P; idx += kThreads) {
                    const int g = fdiv(idx, 1.0f / (float)P), px = idx - g * P;
                    slot(o.dst, g)[px] = add_rn(slot(o.dst, g)[px], slot(o.src, g)[px]);
                }
                break;
            }
            case CNNGP_OP_SCALE: {
                const int P = o.Hi * o.Wi;
                const T scale = sizeof(T) == 4 ? (T)o.scale_f : (T)o.scale_d;
                for (int idx = tid; idx < G * P; idx += kThreads) {
                    const int g = fdiv(idx, 1.0f / (float)P), px = idx - g * P;
                    slot(o.dst, g)[px] = mul_rn(slot(o.src, g)[px], scale);
                }
                break;
            }
            default: break;
        }
        __syncthreads();
    }

    if (tid < G && evalid[tid]) {
        const T v = slot(p.final_slot, tid)[0];
        if (MODE == 1) {
            if (p.kdiag && ej[tid] == 0) p.kdiag[ei[tid]] = v;
        } else if (p.diag) {
            p.out[ei[tid]] = v;
        } else {
            p.out[(size_t)ei[tid] * p.ld_out + ej[tid]] = v;
            if (p.sym && ei[tid] != ej[tid]) p.out[(size_t)ej[tid] * p.ld_out + ei[tid]] = v;
        }
    }
}

template <typename T, int MODE>
int launch(const Plan *plan, GenericParams<T> &gp, int NP, cudaStream_t st) {
    const DevOp *d_ops = plan_device_ops(plan);
    if (!d_ops) return 5;
    gp.ops = d_ops;
    gp.n_ops = (int)plan->ops.size();
    gp.n_slots = plan->n_slots;
    gp.max_map = plan->max_map;
    gp.H = plan->H; gp.W = plan->W;
    gp.final_slot = plan->final_slot;
    gp.aux_elems = plan->aux_elems;
    gp.aux_f_off = plan->aux_f_off;
    gp.NP = NP;
    int dev = 0;
    cudaGetDevice(&dev);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t per_entry = (size_t)(plan->n_slots + 1) * plan->max_map * sizeof(T);
    // aim for two CTAs per SM: half the opt-in budget, minus the static arrays
    size_t budget = (size_t)max_smem / 2 - 1024;
    int G = (int)(budget / per_entry);
    if (G < NP) { budget = (size_t)max_smem - 1024; G = (int)(budget / per_entry); }
    if (G < NP) { set_error("generic kernel: maps do not fit shared memory"); return 6; }
    if (G > kMaxG) G = kMaxG;
    if ((int64_t)G > gp.Q) G = (int)((gp.Q + NP - 1) / NP * NP);
    {   // small problems (the variance rows of one 200-image tile): fewer entries per CTA so that at
        // least two CTAs per SM exist -- with the shared-memory maximum only ~10 SMs would work
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t g_par = (gp.Q + 2 * sms - 1) / (2 * sms);
        if (g_par < NP) g_par = NP;
        if ((int64_t)G > g_par) G = (int)g_par;
    }
    G -= G % NP;
    if (G < NP) G = NP;
    if (MODE == 1 && gp.Q > G) {
        // variance rows (one launch per dataset, O(N) work): pick the entries per CTA that minimise
        // waves x entries -- the time of the launch if a CTA's time is proportional to its entries.  With
        // the largest G, 10 000 images are 625 CTAs on 296 resident ones: three waves, the last 11 % full.
        int sms = 148, smem_sm = 228 * 1024;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        int best = G;
        double best_cost = 1e300;
        for (int g = G; g >= NP; g -= NP) {
            int per_sm = (int)((size_t)smem_sm / (per_entry * g + 2048));
            if (per_sm > 2048 / kThreads) per_sm = 2048 / kThreads;
            if (per_sm < 1) continue;
            const int64_t ctas = (gp.Q + g - 1) / g, resident = (int64_t)per_sm * sms;
            const double cost = (double)((ctas + resident - 1) / resident) * g;
            if (cost < best_cost - 1e-9) { best_cost = cost; best = g; }
        }
        G = best;
    }
    gp.G = G;
    const size_t smem = per_entry * G;
    auto kern = generic_kernel<T, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 7; }
    const int64_t blocks = (gp.Q + G - 1) / G;
    if (blocks > 2147483647LL) { set_error("generic kernel: too many entries for one launch"); return 8; }
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(gp);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("generic kernel launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

template <typename T>
int variances_t(const Plan *plan, const void *d_x, const void *d_z, int64_t N, int32_t C, void *d_aux_x,
                void *d_aux_z, void *d_kdiag, cudaStream_t st) {
    GenericParams<T> gp{};
    const int NP = d_z ? 2 : 1;
    gp.C = C;
    gp.x = (const T *)d_x; gp.z = (const T *)d_z;
    gp.N1 = N; gp.N2 = N; gp.Q = N * NP;
    gp.aux_x_out = (T *)d_aux_x; gp.aux_z_out = (T *)d_aux_z; gp.kdiag = (T *)d_kdiag;
    return launch<T, 1>(plan, gp, NP, st);
}

template <typename T>
int gram_t(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2, int32_t C,
           const void *d_aux_x, const void *d_aux_z, int32_t same, int32_t diag, int32_t symmetric,
           void *d_out, int64_t ld_out, cudaStream_t st) {
    GenericParams<T> gp{};
    gp.C = C;
    gp.x = (const T *)d_x; gp.z = (const T *)d_z;
    gp.N1 = N1; gp.N2 = N2;
    gp.aux_x = (const T *)d_aux_x; gp.aux_z = (const T *)d_aux_z;
    gp.out = (T *)d_out; gp.ld_out = ld_out;
    gp.same = same ? 1 : 0; gp.diag = diag ? 1 : 0; gp.sym = (symmetric && !diag) ? 1 : 0;
    gp.Q = diag ? N1 : (gp.sym ? N1 * (N1 + 1) / 2 : N1 * N2);
    return launch<T, 0>(plan, gp, 1, st);
}

// ---- map-level kernels behind module.propagate(kp) -----------------------------------------
template <typename T>
__global__ void conv_maps_kernel(const T *in, int64_t M, int Hi, int Wi, int ke, int t0, int stride, int pad,
                                 int dil, T scale, T bias, T *out, int Ho, int Wo) {
    const int64_t total = M * Ho * Wo;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / (Ho * Wo);
        const int r = (int)(idx - m * Ho * Wo);
        const int yo = r / Wo, xo = r - yo * Wo;
        const T *src = in + m * Hi * Wi;
        T acc = (T)0;
        for (int ty = t0; ty < ke; ++ty) {
            const int yi = yo * stride - pad + dil * ty;
            if (yi < 0 || yi >= Hi) continue;
            for (int tx = t0; tx < ke; ++tx) {
                const int xi = xo * stride - pad + dil * tx;
                if (xi < 0 || xi >= Wi) continue;
                acc += src[yi * Wi + xi];
            }
        }
        out[idx] = add_rn(mul_rn(acc, scale), bias);
    }
}

template <typename T>
__global__ void relu_maps_kernel(T *xy, const T *xx, const T *yy, int64_t Nx, int64_t Ny, int64_t P, int same,
                                 int diag) {
    const int64_t rows = diag ? Nx : Nx * Ny;
    const int64_t total = rows * P;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / P, px = idx - r * P;
        const int64_t i = diag ? r : r / Ny, j = diag ? r : r % Ny;
        const T vx = xx[i * P + px];
        xy[idx] = (same && i == j) ? div_rn(vx, (T)2) : relu_literal(xy[idx], vx, yy[j * P + px]);
    }
}

int grid_for(int64_t total) {
    int64_t b = (total + 255) / 256;
    return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace

int launch_generic_variances(const Plan *plan, const void *d_x, const void *d_z, int64_t N, int32_t C,
                             void *d_aux_x, void *d_aux_z, void *d_kdiag, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    return plan->dtype == CNNGP_F32 ? variances_t<float>(plan, d_x, d_z, N, C, d_aux_x, d_aux_z, d_kdiag, st)
                                    : variances_t<double>(plan, d_x, d_z, N, C, d_aux_x, d_aux_z, d_kdiag, st);
}

int launch_generic_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
                        int32_t C, const void *d_aux_x, const void *d_aux_z, int32_t same, int32_t diag,
                        int32_t symmetric, void *d_out, int64_t ld_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    return plan->dtype == CNNGP_F32
               ? gram_t<float>(plan, d_x, N1, d_z, N2, C, d_aux_x, d_aux_z, same, diag, symmetric, d_out, ld_out, st)
               : gram_t<double>(plan, d_x, N1, d_z, N2, C, d_aux_x, d_aux_z, same, diag, symmetric, d_out, ld_out, st);
}

int launch_conv_maps(const void *d_in, int64_t M, int32_t Hi, int32_t Wi, const cnngp_op *c, int32_t dtype,
                     void *d_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (c->opcode != CNNGP_OP_CONV || c->ke < 1 || c->stride < 1 || c->dil < 1 || c->pad < 0) {
        set_error("cnngp_conv_maps: not a valid CONV op");
        return 1;
    }
    auto osz = [&](int n) { int num = n + 2 * c->pad - c->dil * (c->ke - 1) - 1; return num < 0 ? 0 : num / c->stride + 1; };
    const int Ho = osz(Hi), Wo = osz(Wi);
    if (Ho < 1 || Wo < 1) { set_error("cnngp_conv_maps: conv output would be empty"); return 2; }
    const int t0 = c->zero_first ? 1 : 0;
    const int64_t total = M * Ho * Wo;
    if (dtype == CNNGP_F32)
        conv_maps_kernel<float><<<grid_for(total), 256, 0, st>>>((const float *)d_in, M, Hi, Wi, c->ke, t0, c->stride,
                                                                 c->pad, c->dil, (float)c->scale, (float)c->bias,
                                                                 (float *)d_out, Ho, Wo);
    else
        conv_maps_kernel<double><<<grid_for(total), 256, 0, st>>>((const double *)d_in, M, Hi, Wi, c->ke, t0, c->stride,
                                                                  c->pad, c->dil, c->scale, c->bias, (double *)d_out,
                                                                  Ho, Wo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("conv_maps launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

int launch_relu_maps(void *d_xy, const void *d_xx, const void *d_yy, int64_t Nx, int64_t Ny, int64_t P,
                     int32_t same, int32_t diag, int32_t dtype, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (diag ? Nx : Nx * Ny) * P;
    if (dtype == CNNGP_F32)
        relu_maps_kernel<float><<<grid_for(total), 256, 0, st>>>((float *)d_xy, (const float *)d_xx, (const float *)d_yy,
                                                                 Nx, Ny, P, same, diag);
    else
        relu_maps_kernel<double><<<grid_for(total), 256, 0, st>>>((double *)d_xy, (const double *)d_xx,
                                                                  (const double *)d_yy, Nx, Ny, P, same, diag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("relu_maps launch: ") + cudaGetErrorString(e)); return 9; }
    return 0;
}

}  // namespace cnngp
