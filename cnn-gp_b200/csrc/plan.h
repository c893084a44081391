// plan.h -- host-side plan: the linearised layer program with inferred shapes.
// Reference semantics: cnn_gp/kernels.py:18-57 (forward), :61-98 (Conv2d), :128-165 (ReLU),
// :246-254 (Sum), :203-225 (Mixture).  See include/cnngp.h for the op encoding.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cnngp.h"

namespace cnngp {

// One op with resolved geometry, laid out for the device interpreter.
struct DevOp {
    int32_t opcode, src, dst;
    int32_t ke, t0, stride, pad, dil;  // t0 = first non-zero tap (1 when zero_first)
    int32_t Hi, Wi, Ho, Wo;            // input / output map size of this op
    int32_t aux_off;                   // RELU: offset of its variance map in the per-image aux row
    int32_t relu_index;                // RELU: ordinal
    int32_t aux_t;                     // RELU: fused (s, 1/s) map stored transposed (lane = row layout)
    int32_t aux_foff;                  // RELU: offset (floats) of its pair-interleaved (s, 1/s) maps inside the
                                       // fused section of a row; the section of this layer is 4*aux_half floats
    int32_t aux_half;                  // RELU: ceil(pixels / 2): pixels [0, half) live in row 2k, the rest in 2k+1
    float aux_scale;                   // RELU: factor on the fused s map (1/s gets its inverse): the straight-line
                                       // fused kernel keeps its maps divided by the product of the conv taps so far
    float scale_f, bias_f;
    double scale_d, bias_d;
};

struct FusedPlan;  // gram_fused.cu
struct FNetPlan;   // gram_fnet.cu

struct Plan {
    int32_t n_ops = 0, n_slots = 0, H = 0, W = 0, dtype = 0;
    std::vector<DevOp> ops;
    int64_t aux_elems = 0;    // floats per image row: xx maps [+ fused (s, 1/s) maps]
    int32_t relu_elems = 0;   // sum of ReLU input map sizes
    int32_t aux_f_off = 0;    // offset of the fused maps inside a row (0 = none)
    int32_t n_relu = 0;
    int32_t max_map = 0;      // largest map (or separable-conv intermediate) in elements
    int32_t final_slot = 0;
    // lazily uploaded device copies of `ops`, one per device the plan has been used on
    mutable std::mutex mu;
    mutable std::vector<std::pair<int, DevOp *>> d_ops;
    // fused-kernel description (nullptr when the program is outside the fused kernel's set)
    FusedPlan *fused = nullptr;
    // fused kernel for programs with Sum / stride / several map sizes (nullptr when not covered)
    FNetPlan *fnet = nullptr;
};

void set_error(const std::string &msg);
// stream-ordered scratch (cudaMallocAsync): let the device's default pool keep at least `bytes` between calls
// (its default gives everything back at the next synchronisation: a map / unmap per call)
void pool_keep(size_t bytes);
void note_launches(int n);  // kernels launched by the current Gram call (cnngp_last_launches)
int build_plan(const cnngp_op *ops, int32_t n_ops, int32_t n_slots, int32_t H, int32_t W,
               int32_t dtype, Plan **out);
const DevOp *plan_device_ops(const Plan *plan);  // uploads on first use; nullptr on CUDA error
// One tile counter per (device, stream) for the persistent Gram kernels: launches on one stream are
// ordered, so zeroing the stream's counter on that stream ahead of each launch cannot disturb a
// launch that is still running (a round-robin pool shared by all streams could).  nullptr on error.
unsigned long long *tile_counter_for(void *stream);
double plan_flops_per_pair(const Plan *plan, int32_t C);

// gram_generic.cu
int launch_generic_variances(const Plan *plan, const void *d_x, const void *d_z, int64_t N, int32_t C,
                             void *d_aux_x, void *d_aux_z, void *d_kdiag, void *stream);
int launch_generic_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
                        int32_t C, const void *d_aux_x, const void *d_aux_z, int32_t same, int32_t diag,
                        int32_t symmetric, void *d_out, int64_t ld_out, void *stream);
int launch_conv_maps(const void *d_in, int64_t M, int32_t Hi, int32_t Wi, const cnngp_op *conv,
                     int32_t dtype, void *d_out, void *stream);
int launch_relu_maps(void *d_xy, const void *d_xx, const void *d_yy, int64_t Nx, int64_t Ny, int64_t P,
                     int32_t same, int32_t diag, int32_t dtype, void *stream);

// gram_variance.cu: variance rows of the straight-line fused kernel's programs, one warp per image pair;
// -1 when the program is not covered (the caller falls back to launch_generic_variances)
int launch_fused_variances(const Plan *plan, const void *d_x, int64_t N, int32_t C, void *d_aux_x,
                           void *d_kdiag, void *stream);

// Optional progress reporting of a symmetric fused launch, for streaming the result out while the
// launch is still running: the kernel counts finished (tile, warp) units per SUPER-ROW (a band of
// `rows_per_super` image rows, enumerated first to last); band b of the output is final -- mirrored
// entries included -- once bands 0..b have reached their expected counts.
struct RowProgress {
    unsigned *d_done = nullptr;       // in: device counters, zeroed on the launch stream before the launch
    int64_t capacity = 0;             // in: number of counters available
    int n_super_rows = 0;             // out
    int64_t rows_per_super = 0;       // out
    std::vector<unsigned> expected;   // out: count at which super-row b is complete
};

// gram_fused.cu
FusedPlan *fused_plan_create(const Plan *plan);  // nullptr if not covered
void fused_plan_destroy(FusedPlan *fp);
std::string fused_plan_describe(const FusedPlan *fp);
std::string fused_plan_dump(const FusedPlan *fp, const Plan *plan);
int launch_fused_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2,
                      int32_t C, const void *d_aux_x, const void *d_aux_z, int32_t same, int32_t diag,
                      int32_t symmetric, const void *d_kdiag, void *d_out, int64_t ld_out, void *stream,
                      RowProgress *progress = nullptr, int64_t mirror_block = 0);

// gram_fnet.cu
FNetPlan *fnet_plan_create(const Plan *plan);  // nullptr if not covered
void fnet_plan_destroy(FNetPlan *fp);
std::string fnet_plan_describe(const FNetPlan *fp);
std::string fnet_plan_dump(const FNetPlan *fp);
int launch_fnet_gram(const Plan *plan, const void *d_x, int64_t N1, const void *d_z, int64_t N2, int32_t C,
                     const void *d_aux_x, const void *d_aux_z, int32_t symmetric, const void *d_kdiag,
                     void *d_out, int64_t ld_out, void *stream, RowProgress *progress = nullptr, int64_t mirror_block = 0);

}  // namespace cnngp
