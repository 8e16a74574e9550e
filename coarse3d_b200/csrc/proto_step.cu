// The prototype step of a training iteration as one unit: EMA prototype update
// (salsanext_proto.py:497-527, inside model.forward) followed by the contrastive loss on the
// UPDATED bank (trainer.py:675-686), both on the same label image.
//
// When both operators see the same labels (the weak-label image; `entropy_selection` off, or
// the step pipeline of bench.py), everything they derive from the labels is shared: ONE label
// split (labelsplit.cuh) feeds the loss's anchor sampler and the EMA's row kernels.  The caller
// drives the phases, so that it can place the all-reduce of the packed sums and
// c3d_proto_ema_apply between the accumulation and the loss rows:
//
//   phase 1  split        labels -> class-major slots + entropy weights         (2 launches)
//   phase 2  sample       anchor sampling per (scan, class) segment             (1)
//   phase 4  accumulate   EMA rows, Sinkhorn, segmented sums -> packed          (4)
//   [ all-reduce(packed); c3d_proto_ema_apply(prototypes, packed) -> prototypes ]
//   phase 8  loss rows    loss + gradient rows against the updated bank         (2)
//   [ c3d_proto_loss_backward on the same workspace ]
//
// The workspace starts with the loss workspace (c3d_proto_loss_backward / _info / _rows work on
// it unchanged), followed by the EMA arrays.
#include "proto_internal.cuh"

using namespace c3d;

static size_t step_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t c3d_proto_step_workspace_bytes(int batch, int n_classes, int hw, int dim, int sub_protos,
                                                 int num_anchor, int64_t max_rows) {
  if (batch <= 0 || n_classes < 2 || hw <= 0 || dim <= 0 || sub_protos <= 0 || num_anchor <= 0 || max_rows <= 0)
    return 0;
  return step_align(loss_ws_bytes(batch, n_classes, hw, dim, sub_protos, num_anchor)) +
         step_align(ema_extra_bytes(n_classes, dim, sub_protos, max_rows)) +
         (dim % 32 == 0 && dim <= 256 ? (size_t)max_rows * dim * 4 : 0);   // raw gathered rows
}

extern "C" int c3d_proto_step(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* prototypes, const float* ln_d_w, const float* ln_d_b, const float* ln_c_w,
    const float* ln_c_b, float ln_eps, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, const float* gumbel, int assign_mode, uint64_t seed,
    int64_t max_rows, int need_grad, int phases, const float* bank_n, uint64_t* seed_counters,
    void* workspace, float* packed, float* proto_target, float* loss_out, void* cofill_ptr,
    size_t cofill_bytes, void* stream) {
  C3D_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256 B aligned");
  C3D_REQUIRE(phases > 0 && phases < 16, "phases: bit mask of 1 split, 2 sample, 4 accumulate, 8 loss rows");
  C3D_REQUIRE(batch > 0 && n_classes >= 2 && dim > 0 && sub_protos > 0 && num_anchor > 0 && max_rows > 0 &&
              proj_h > 0 && proj_w > 0, "bad shape");
  C3D_REQUIRE((cofill_ptr == nullptr) == (cofill_bytes == 0) && cofill_bytes % 16 == 0 &&
              (reinterpret_cast<uintptr_t>(cofill_ptr) & 15) == 0, "carried fill: 16 B aligned pointer and size");
  // The carried fill goes to ONE kernel of the call: the rows kernel of the latest phase asked
  // for (loss rows, else EMA rows), else the label split.
  FillShare fill{reinterpret_cast<char*>(cofill_ptr), cofill_bytes};
  const FillShare none{nullptr, 0};
  const int carrier = (phases & 8) ? 8 : ((phases & 4) ? 4 : ((phases & 1) ? 1 : 0));
  if (cofill_bytes && carrier == 0) { int rcf = launch_fill(cofill_ptr, cofill_bytes, (cudaStream_t)stream); if (rcf) return rcf; }
  const int HW = proj_h * proj_w;
  int rc;
  char* extra = reinterpret_cast<char*>(workspace) +
                step_align(loss_ws_bytes(batch, n_classes, HW, dim, sub_protos, num_anchor));
  // un-normalised rows gathered once by the EMA kernel and re-read, contiguous, by the loss rows
  float* raw_rows = (dim % 32 == 0 && dim <= 256)
      ? reinterpret_cast<float*>(extra + step_align(ema_extra_bytes(n_classes, dim, sub_protos, max_rows))) : nullptr;
  const int loss_phases = ((phases & 1) ? kPhaseSplit : 0) | ((phases & 2) ? kPhaseSample : 0);
  if (loss_phases) {
    rc = proto_loss_forward_impl(nullptr, probs, labels, keep_mask, nullptr, batch, dim, proj_h, proj_w,
                                 n_classes, sub_protos, ignore_label, temperature, base_temperature,
                                 num_anchor, keep, keep_rows, seed, need_grad, loss_phases, workspace, nullptr,
                                 nullptr, 0, stream, nullptr, 0, nullptr, seed_counters, 0, carrier == 1 ? fill : none);
    if (rc) return rc;
  }
  if (phases & 4) {
    const SplitWs s = loss_ws_split(workspace, batch, n_classes, HW, dim, sub_protos, num_anchor);
    rc = proto_ema_accumulate_impl(feats, nullptr, nullptr, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b, ln_eps,
                                   batch, dim, proj_h, proj_w, n_classes, sub_protos, ignore_label, max_rows,
                                   gumbel, assign_mode, seed, extra, &s, packed, proto_target, stream, raw_rows, 0, bank_n,
                                   seed_counters, carrier == 4 ? fill : none);
    if (rc) return rc;
  }
  if (phases & 8) {
    rc = proto_loss_forward_impl(feats, nullptr, nullptr, nullptr, prototypes, batch, dim, proj_h, proj_w,
                                 n_classes, sub_protos, ignore_label, temperature, base_temperature,
                                 num_anchor, nullptr, 0, seed, need_grad, kPhaseRows, workspace, loss_out,
                                 nullptr, 0, stream, raw_rows, (int)(max_rows > 0x7fffffff ? 0x7fffffff : max_rows), bank_n,
                                 nullptr, need_grad >> 1, carrier == 8 ? fill : none);
    if (rc) return rc;
  }
  return C3D_OK;
}
