// plan.cu -- plan construction (shape inference + validation) and the C ABI entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "plan.h"

namespace cnngp {

static thread_local std::string g_err;
static thread_local int g_last_path = CNNGP_PATH_NONE;
static thread_local int g_last_launches = 0;

void set_error(const std::string &msg) { g_err = msg; }
void note_launches(int n) { g_last_launches = n; }
void pool_keep(size_t bytes) {
    static std::mutex mu;
    static unsigned long long kept[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(mu);
    if (kept[dev] >= bytes) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
    unsigned long long keep = (unsigned long long)bytes;
    if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess) kept[dev] = keep;
}

static int conv_out(int n, int ke, int stride, int pad, int dil) {
    // torch: floor((n + 2 pad - dil (ke-1) - 1) / stride) + 1, empty when the numerator < 0
    int num = n + 2 * pad - dil * (ke - 1) - 1;
    if (num < 0) return 0;
    return num / stride + 1;
}


int build_plan(const cnngp_op *ops, int32_t n_ops, int32_t n_slots, int32_t H, int32_t W,
               int32_t dtype, Plan **out) {
    if (!ops || n_ops < 0 || n_slots < 1 || H < 1 || W < 1 || (dtype != CNNGP_F32 && dtype != CNNGP_F64)) {
        set_error("cnngp_plan_create: bad arguments");
        return 1;
    }
    struct Shape { int h = -1, w = -1; };
    std::vector<Shape> slot(n_slots);
    slot[0].h = H;
    slot[0].w = W;
    Plan *p = new Plan();
    p->n_ops = n_ops; p->n_slots = n_slots; p->H = H; p->W = W; p->dtype = dtype;
    p->max_map = H * W;
    p->final_slot = 0;
    char buf[256];
    auto fail = [&](const char *m, int idx) {
        snprintf(buf, sizeof buf, "cnngp_plan_create: op %d: %s", idx, m);
        set_error(buf);
        delete p;
        return 2;
    };
    for (int k = 0; k < n_ops; ++k) {
        const cnngp_op &o = ops[k];
        if (o.src < 0 || o.src >= n_slots || o.dst < 0 || o.dst >= n_slots) return fail("slot out of range", k);
        if (slot[o.src].h < 0) return fail("source slot read before it is written", k);
        DevOp d{};
        d.opcode = o.opcode; d.src = o.src; d.dst = o.dst;
        d.Hi = slot[o.src].h; d.Wi = slot[o.src].w; d.Ho = d.Hi; d.Wo = d.Wi;
        d.aux_off = 0; d.relu_index = -1; d.aux_t = 0; d.aux_foff = 0; d.aux_half = 0; d.aux_scale = 1.f;
        d.scale_d = o.scale; d.bias_d = o.bias; d.scale_f = (float)o.scale; d.bias_f = (float)o.bias;
        switch (o.opcode) {
            case CNNGP_OP_CONV: {
                if (o.ke < 1 || o.stride < 1 || o.dil < 1 || o.pad < 0) return fail("bad conv geometry", k);
                if (o.zero_first && o.ke < 2) return fail("zero_first needs ke >= 2", k);
                d.ke = o.ke; d.t0 = o.zero_first ? 1 : 0; d.stride = o.stride; d.pad = o.pad; d.dil = o.dil;
                d.Ho = conv_out(d.Hi, o.ke, o.stride, o.pad, o.dil);
                d.Wo = conv_out(d.Wi, o.ke, o.stride, o.pad, o.dil);
                if (d.Ho < 1 || d.Wo < 1) return fail("conv output would be empty", k);
                p->max_map = std::max(p->max_map, d.Hi * d.Wo);  // separable intermediate
                break;
            }
            case CNNGP_OP_RELU:
                d.aux_off = (int32_t)p->aux_elems;
                d.relu_index = p->n_relu++;
                p->aux_elems += (int64_t)d.Hi * d.Wi;
                break;
            case CNNGP_OP_COPY:
            case CNNGP_OP_SCALE:
                break;
            case CNNGP_OP_ADD:
                if (slot[o.dst].h != d.Hi || slot[o.dst].w != d.Wi) return fail("ADD of maps with different shapes", k);
                break;
            default:
                return fail("unknown opcode", k);
        }
        slot[o.dst].h = d.Ho;
        slot[o.dst].w = d.Wo;
        p->max_map = std::max(p->max_map, d.Ho * d.Wo);
        p->final_slot = o.dst;
        p->ops.push_back(d);
    }
    if (slot[p->final_slot].h != 1 || slot[p->final_slot].w != 1) {
        snprintf(buf, sizeof buf, "cnngp_plan_create: final map is %dx%d, not 1x1", slot[p->final_slot].h,
                 slot[p->final_slot].w);
        set_error(buf);
        delete p;
        return 3;
    }
    p->relu_elems = (int32_t)p->aux_elems;
    {   // layout of the fused kernels' (s, 1/s) section: per ReLU 4 * ceil(pixels / 2) floats per row
        int32_t foff = 0;
        for (DevOp &d : p->ops) {
            if (d.opcode != CNNGP_OP_RELU) continue;
            d.aux_half = (d.Hi * d.Wi + 1) / 2;
            d.aux_foff = foff;
            foff += 4 * d.aux_half;
        }
        // CNNGP_NO_FUSED=1 (measurement aid): straight-line programs run on the fused-net kernel too
        p->fused = getenv("CNNGP_NO_FUSED") ? nullptr : fused_plan_create(p);
        if (!p->fused) {
            for (DevOp &d : p->ops) { d.aux_t = 0; d.aux_scale = 1.f; }  // whatever a partial translation left
            p->fnet = fnet_plan_create(p);
        }
        if (p->fused || p->fnet) {  // rows grow by the pair-interleaved maps, 16-byte aligned
            p->aux_f_off = (p->relu_elems + 3) / 4 * 4;
            p->aux_elems = (int64_t)p->aux_f_off + foff;
        }
    }
    *out = p;
    return 0;
}

const DevOp *plan_device_ops(const Plan *plan) {
    std::lock_guard<std::mutex> lk(plan->mu);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed (no CUDA device?)"); return nullptr; }
    for (const auto &e : plan->d_ops)
        if (e.first == dev) return e.second;
    DevOp *d = nullptr;
    size_t bytes = std::max<size_t>(1, plan->ops.size()) * sizeof(DevOp);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc(plan ops): ") + cudaGetErrorString(e)); return nullptr; }
    e = cudaMemcpy(d, plan->ops.data(), plan->ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error(std::string("cudaMemcpy(plan ops): ") + cudaGetErrorString(e)); cudaFree(d); return nullptr; }
    plan->d_ops.emplace_back(dev, d);
    return d;
}

unsigned long long *tile_counter_for(void *stream) {
    static std::mutex mu;
    static std::vector<std::pair<std::pair<int, void *>, unsigned long long *>> table;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed (no CUDA device?)"); return nullptr; }
    std::lock_guard<std::mutex> lk(mu);
    for (const auto &e : table)
        if (e.first.first == dev && e.first.second == stream) return e.second;
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(unsigned long long));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc(tile counter): ") + cudaGetErrorString(e)); return nullptr; }
    table.push_back({{dev, stream}, d});
    return d;
}

// init 2*C*H*W; conv as separable box sum (k-1)*Hi*Wo + (k-1)*Wo*Ho + 2*Wo*Ho;
// ReLU 15 per input pixel; Sum / Mixture 1 per pixel per extra branch.
double plan_flops_per_pair(const Plan *plan, int32_t C) {
    double f = 2.0 * C * plan->H * plan->W;
    for (const DevOp &o : plan->ops) {
        switch (o.opcode) {
            case CNNGP_OP_CONV: {
                double k = o.ke - o.t0;
                f += (k - 1) * o.Hi * o.Wo + (k - 1) * o.Wo * o.Ho + 2.0 * o.Wo * o.Ho;
                break;
            }
            case CNNGP_OP_RELU: f += 15.0 * o.Hi * o.Wi; break;
            case CNNGP_OP_ADD:
            case CNNGP_OP_SCALE: f += 1.0 * o.Hi * o.Wi; break;
            default: break;
        }
    }
    return f;
}

}  // namespace cnngp

using namespace cnngp;

extern "C" {

int cnngp_abi_version(void) { return CNNGP_ABI_VERSION; }
const char *cnngp_last_error(void) { return g_err.c_str(); }
int cnngp_last_path(void) { return g_last_path; }
int cnngp_last_launches(void) { return g_last_launches; }

int cnngp_plan_create(const cnngp_op *ops, int32_t n_ops, int32_t n_slots, int32_t H, int32_t W,
                      int32_t dtype, cnngp_plan **out) {
    if (!out) { set_error("cnngp_plan_create: out is NULL"); return 1; }
    Plan *p = nullptr;
    int rc = build_plan(ops, n_ops, n_slots, H, W, dtype, &p);
    if (rc) return rc;
    *out = reinterpret_cast<cnngp_plan *>(p);
    return 0;
}

void cnngp_plan_destroy(cnngp_plan *plan) {
    Plan *p = reinterpret_cast<Plan *>(plan);
    if (!p) return;
    for (const auto &e : p->d_ops) cudaFree(e.second);
    if (p->fused) fused_plan_destroy(p->fused);
    if (p->fnet) fnet_plan_destroy(p->fnet);
    delete p;
}

int64_t cnngp_plan_aux_elems(const cnngp_plan *plan) { return reinterpret_cast<const Plan *>(plan)->aux_elems; }
double cnngp_plan_flops_per_pair(const cnngp_plan *plan, int32_t C) {
    return plan_flops_per_pair(reinterpret_cast<const Plan *>(plan), C);
}
int cnngp_plan_has_fused(const cnngp_plan *plan) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    return p->fused ? CNNGP_PATH_FUSED : (p->fnet ? CNNGP_PATH_FUSED_NET : 0);
}

int64_t cnngp_plan_describe(const cnngp_plan *plan, char *buf, int64_t cap) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    std::string t;
    if (!p) t = "null plan";
    else if (p->fused) t = fused_plan_describe(p->fused);
    else if (p->fnet) t = fnet_plan_describe(p->fnet);
    else t = "generic: " + std::to_string(p->ops.size()) + " ops, " + std::to_string(p->n_slots) + " slots";
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>((size_t)cap - 1, t.size());
        memcpy(buf, t.data(), n);
        buf[n] = 0;
    }
    return (int64_t)t.size() + 1;
}

int64_t cnngp_plan_dump(const cnngp_plan *plan, char *buf, int64_t cap) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    std::string t;
    if (!p) t = "null plan\n";
    else if (p->fused) t = fused_plan_dump(p->fused, p);
    else if (p->fnet) t = fnet_plan_dump(p->fnet);
    else t = "generic\n";
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>((size_t)cap - 1, t.size());
        memcpy(buf, t.data(), n);
        buf[n] = 0;
    }
    return (int64_t)t.size() + 1;
}

int cnngp_variances(const cnngp_plan *plan, const void *d_x, const void *d_z, int64_t N, int32_t C,
                    void *d_aux_x, void *d_aux_z, void *d_kdiag, void *stream) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    if (!p || !d_x || N < 0 || C < 1) { set_error("cnngp_variances: bad arguments"); return 1; }
    if (p->aux_elems > 0 && !d_aux_x) { set_error("cnngp_variances: d_aux_x is NULL"); return 1; }
    if (d_z && p->aux_elems > 0 && !d_aux_z) { set_error("cnngp_variances: d_aux_z is NULL"); return 1; }
    if (N == 0) return 0;
    if (!d_z) {  // straight-line 28 x 28 programs: the streaming kernel (bit-identical to the interpreter)
        const int rc = launch_fused_variances(p, d_x, N, C, d_aux_x, d_kdiag, stream);
        if (rc >= 0) return rc;
    }
    return launch_generic_variances(p, d_x, d_z, N, C, d_aux_x, d_aux_z, d_kdiag, stream);
}

int cnngp_gram(const cnngp_plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
               int32_t C, const void *d_aux_x, const void *d_aux_z, const void *d_kdiag, int32_t same,
               int32_t diag, int32_t symmetric, void *d_out, int64_t ld_out, int32_t path, void *stream) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    if (!p || !d_x || !d_z || !d_out || N1 < 0 || N2 < 0 || C < 1) { set_error("cnngp_gram: bad arguments"); return 1; }
    if (p->aux_elems > 0 && (!d_aux_x || !d_aux_z)) { set_error("cnngp_gram: variance maps missing"); return 1; }
    if (diag && N1 != N2) { set_error("cnngp_gram: diag needs N1 == N2 (kernels.py:28)"); return 1; }
    if ((same || symmetric) && !diag && N1 != N2) { set_error("cnngp_gram: same needs N1 == N2 (kernels.py:161)"); return 1; }
    if (!diag && ld_out < N2) { set_error("cnngp_gram: ld_out < N2"); return 1; }
    if (N1 == 0 || N2 == 0) return 0;
    const bool have = p->fused || p->fnet;
    bool use_fused = false;
    if (path == CNNGP_PATH_FUSED) {
        if (!have) { set_error("cnngp_gram: no fused kernel covers this program"); return 4; }
        use_fused = true;
    } else if (path == CNNGP_PATH_NONE) {
        use_fused = have;
    }
    // the fused kernels assume i == j entries equal the variance recursion, which the literal
    // reference only guarantees when x and z hold the same images
    if (use_fused && same && !symmetric) {
        if (path == CNNGP_PATH_FUSED) { set_error("cnngp_gram: fused path needs symmetric when same"); return 4; }
        use_fused = false;
    }
    if (use_fused && diag) use_fused = false;  // O(N) work: generic is enough
    g_last_path = !use_fused ? CNNGP_PATH_GENERIC : (p->fused ? CNNGP_PATH_FUSED : CNNGP_PATH_FUSED_NET);
    g_last_launches = 1;
    if (use_fused && p->fused)
        return launch_fused_gram(p, d_x, N1, d_z, N2, C, d_aux_x, d_aux_z, same, diag, symmetric, d_kdiag, d_out, ld_out, stream);
    if (use_fused)
        return launch_fnet_gram(p, d_x, N1, d_z, N2, C, d_aux_x, d_aux_z, symmetric, d_kdiag, d_out, ld_out, stream);
    return launch_generic_gram(p, d_x, N1, d_z, N2, C, d_aux_x, d_aux_z, same, diag, symmetric, d_out, ld_out, stream);
}

// A band of block rows of model(X) in one launch (see cnngp.h).
int cnngp_gram_band(const cnngp_plan *plan, const void *d_x, int64_t N1, int64_t N2, int32_t C, const void *d_aux,
                    const void *d_kdiag, int64_t block, void *d_out, int64_t ld_out, void *stream) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    if (!p || !d_x || !d_aux || !d_out || N1 < 0 || N2 < N1 || C < 1 || block < 0 || ld_out < N2) {
        set_error("cnngp_gram_band: bad arguments");
        return 1;
    }
    if (!p->fused && !p->fnet) { set_error("cnngp_gram_band: only the fused kernels evaluate bands"); return 4; }
    if (N1 == 0) return 0;
    g_last_path = p->fused ? CNNGP_PATH_FUSED : CNNGP_PATH_FUSED_NET;
    g_last_launches = 1;
    return p->fused ? launch_fused_gram(p, d_x, N1, d_x, N2, C, d_aux, d_aux, 1, 0, 1, d_kdiag, d_out, ld_out, stream, nullptr, block)
                    : launch_fnet_gram(p, d_x, N1, d_x, N2, C, d_aux, d_aux, 1, d_kdiag, d_out, ld_out, stream, nullptr, block);
}

// model(X) with the result streamed to host memory while the launch is still running: the kernel
// counts finished tiles per band of rows (RowProgress), the copy stream waits on each band's counter
// with a stream memory operation (cuStreamWaitValue32) and copies the band out -- mirrored entries
// of a band are written by the bands before it, which the in-order waits cover.
int cnngp_gram_symmetric_to_host(const cnngp_plan *plan, const void *d_x, int64_t N, int32_t C, const void *d_aux,
                                 const void *d_kdiag, void *d_out, int64_t ld_out, void *h_out, int64_t ld_host,
                                 void *d_scratch, int64_t scratch_bytes, void *stream, void *copy_stream) {
    const Plan *p = reinterpret_cast<const Plan *>(plan);
    if (!p || !d_x || !d_aux || !d_out || !h_out || !d_scratch || N < 1 || C < 1 || ld_out < N || ld_host < N) {
        set_error("cnngp_gram_symmetric_to_host: bad arguments");
        return 1;
    }
    if (!p->fused && !p->fnet) { set_error("cnngp_gram_symmetric_to_host: only the fused kernels report progress"); return 4; }
    typedef int (*WaitValue32)(void *, unsigned long long, unsigned int, unsigned int);
    static WaitValue32 wait_value = nullptr;
    if (!wait_value) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            set_error("cnngp_gram_symmetric_to_host: cuStreamWaitValue32 unavailable");
            return 7;
        }
        wait_value = reinterpret_cast<WaitValue32>(fn);
    }
    cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
    RowProgress prog;
    prog.d_done = (unsigned *)d_scratch;
    prog.capacity = scratch_bytes / (int64_t)sizeof(unsigned);
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, (size_t)scratch_bytes, st);
    cudaEvent_t ev = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ev, 0);  // no band is waited for before the counters are zero
    if (ev) cudaEventDestroy(ev);
    if (e != cudaSuccess) { set_error(std::string("cnngp_gram_symmetric_to_host: ") + cudaGetErrorString(e)); return 7; }
    g_last_path = p->fused ? CNNGP_PATH_FUSED : CNNGP_PATH_FUSED_NET;
    g_last_launches = 1;
    int rc = p->fused ? launch_fused_gram(p, d_x, N, d_x, N, C, d_aux, d_aux, 1, 0, 1, d_kdiag, d_out, ld_out, stream, &prog)
                      : launch_fnet_gram(p, d_x, N, d_x, N, C, d_aux, d_aux, 1, d_kdiag, d_out, ld_out, stream, &prog);
    if (rc) return rc;  // nothing was queued on the copy stream yet: it cannot wait for a launch that never ran
    // behind the launch every counter is raised to 0x7f7f7f7f (>= any expected count in the signed
    // comparison the wait uses): whatever happens, the copy stream's waits end when the kernel has
    // (a band can then only leave late, never hang the stream)
    e = cudaMemsetAsync(d_scratch, 0x7F, (size_t)scratch_bytes, st);
    if (e != cudaSuccess) { set_error(std::string("cnngp_gram_symmetric_to_host: ") + cudaGetErrorString(e)); return 7; }
    const size_t esz = sizeof(float);
    const char *dbg = getenv("CNNGP_E2E_DEBUG");  // measurement aid: "nowait" = counters only, "nocopy" = waits without copies
    if (dbg && !strcmp(dbg, "nowait")) return 0;
    for (int b = 0; b < prog.n_super_rows; ++b) {
        const int64_t r0 = (int64_t)b * prog.rows_per_super;
        const int64_t rows = std::min<int64_t>(prog.rows_per_super, N - r0);
        if (rows <= 0) break;
        // CU_STREAM_WAIT_VALUE_GEQ = 0: waits until (int32_t)(*addr - value) >= 0
        const int wr = wait_value(cs, (unsigned long long)(uintptr_t)(prog.d_done + b), prog.expected[b], 0u);
        if (wr != 0) { set_error("cnngp_gram_symmetric_to_host: cuStreamWaitValue32 failed (" + std::to_string(wr) + ")"); return 7; }
        if (dbg && !strcmp(dbg, "nocopy")) continue;
        e = cudaMemcpy2DAsync((char *)h_out + (size_t)r0 * ld_host * esz, (size_t)ld_host * esz,
                              (const char *)d_out + (size_t)r0 * ld_out * esz, (size_t)ld_out * esz, (size_t)N * esz,
                              (size_t)rows, cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) { set_error(std::string("cnngp_gram_symmetric_to_host: ") + cudaGetErrorString(e)); return 7; }
    }
    return 0;
}

int cnngp_conv_maps(const void *d_in, int64_t M, int32_t Hi, int32_t Wi, const cnngp_op *conv,
                    int32_t dtype, void *d_out, void *stream) {
    if (!d_in || !d_out || !conv || M < 0) { set_error("cnngp_conv_maps: bad arguments"); return 1; }
    if (M == 0) return 0;
    return launch_conv_maps(d_in, M, Hi, Wi, conv, dtype, d_out, stream);
}

int cnngp_relu_maps(void *d_xy, const void *d_xx, const void *d_yy, int64_t Nx, int64_t Ny, int64_t P,
                    int32_t same, int32_t diag, int32_t dtype, void *stream) {
    if (!d_xy || !d_xx || !d_yy) { set_error("cnngp_relu_maps: bad arguments"); return 1; }
    if (Nx == 0 || Ny == 0 || P == 0) return 0;
    return launch_relu_maps(d_xy, d_xx, d_yy, Nx, Ny, P, same, diag, dtype, stream);
}

}  // extern "C"
