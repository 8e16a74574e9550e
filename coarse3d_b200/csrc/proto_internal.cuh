// Internal seams between proto_loss.cu, proto_ema.cu and proto_step.cu (not part of the ABI).
#pragma once
#include "common.cuh"
#include "labelsplit.cuh"

namespace c3d {

enum ProtoPhase { kPhaseSplit = 1, kPhaseSample = 2, kPhaseRows = 4 };

// proto_loss.cu
size_t loss_ws_bytes(int B, int C, int HW, int D, int M, int A);
SplitWs loss_ws_split(void* base, int B, int C, int HW, int D, int M, int A);
int proto_loss_forward_impl(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* proto_queue, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, uint64_t seed, int need_grad, int phases, void* workspace,
    float* loss_out, float* zero_buf, int zero_n, void* stream,
    const float* raw_rows /* [slots, D] rows left by the EMA kernel (fused step), or null */, int raw_cap,
    const float* bank_n /* [C, M, D] F.normalize'd bank, or null: normalised here */,
    uint64_t* seed_dev /* [2] device step counters; [0] is added to `seed` and advanced by the sampler */,
    int rows_mode /* 0: register-tiled FFMA products, 1: tensor cores (mma.sync, 3xTF32) */,
    FillShare fill /* carried zero fill: given to the split (kPhaseSplit) or to the rows kernel (kPhaseRows) */);

// proto_ema.cu
struct DenseRows { const float* out_feat; const float* nearest; const float* sim; };
size_t ema_extra_bytes(int C, int D, int M, long long max_rows);
// shared_split != nullptr: the label split has been done (same labels) and `workspace` holds
// only the EMA arrays (ema_extra_bytes); otherwise `workspace` is a full EMA workspace.
int proto_ema_accumulate_impl(
    const float* embedding, const DenseRows* dense, const int64_t* label, const float* prototypes,
    const float* ln_d_w, const float* ln_d_b, const float* ln_c_w, const float* ln_c_b, float ln_eps,
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos, int ignore_label,
    int64_t max_rows, const float* gumbel, int assign_mode, uint64_t seed, void* workspace,
    const SplitWs* shared_split, float* packed, float* proto_target, void* stream,
    float* raw_rows /* [max_rows, D] un-normalised gathered rows, or null */, int rows_v1 /* A/B: warp-per-row kernel */,
    const float* bank_n /* [C, M, D] l2-normalised bank, or null: normalised here */,
    const uint64_t* seed_dev /* [2] device step counters added to `seed` ([1] is this operator's), or null */,
    FillShare fill /* a share of the carried zero fill for the rows kernel, or {null, 0} */);

}  // namespace c3d
