/*
 * coarse3d_b200 -- C ABI of the B200-native COARSE3D per-scan hot path.
 *
 * The reference (astra-vision/COARSE3D) is pure Python: it has no FFI layer.
 * The boundary it offers is the `pc_processor` operator surface; each entry
 * point below is what a binding for one of those operators calls, and cites
 * the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless marked "host";
 *   - no hidden allocation: scratch is passed in, sized by *_workspace_bytes;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; all work is
 *     enqueued on it and nothing synchronises;
 *   - no process-global state decides what a call does: no environment variables are read,
 *     every option is an argument (the only globals are the launch counter and the optional
 *     profiler below, both diagnostics); calls on different streams are independent;
 *   - every function returns a c3d_status; on failure c3d_last_error() gives a
 *     thread-local message; nothing throws across the ABI;
 *   - a scan batch is CSR: `offsets[b] .. offsets[b+1]` are scan b's points.
 */
#ifndef COARSE3D_B200_H_
#define COARSE3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  C3D_OK = 0,
  C3D_INVALID_ARGUMENT = 1,
  C3D_CUDA_ERROR = 2,
  C3D_UNSUPPORTED = 3
} c3d_status;

/* ABI version (major*100 + minor). */
int c3d_version(void);
/* Message of the last failure on this thread ("" if none). */
const char* c3d_last_error(void);
/* Number of kernel launches enqueued by this library since load (all threads). */
long long c3d_launch_count(void);

/* Optional per-kernel device timing: CUDA events recorded on the launching
 * stream right before and after each kernel launch.  kernel_name "" = every
 * kernel, a kernel's name = only that kernel, NULL = off (default).  Not
 * usable during CUDA-graph capture.  c3d_profile_read waits for the recorded
 * events and returns the summed elapsed ms and the number of launches of
 * `kernel_name` ("" = all); c3d_profile_names lists the recorded names,
 * comma separated; c3d_profile_reset drops the records. */
int c3d_profile_enable(const char* kernel_name);
int c3d_profile_read(const char* kernel_name, double* total_ms, long long* count);
int c3d_profile_names(char* buf, int buf_len);
/* "name,start_us,end_us\n" per recorded launch, relative to the first record. */
int c3d_profile_timeline(char* buf, int buf_len);
int c3d_profile_reset(void);

/* ---------------------------------------------------------------- a1 ----
 * RangeProjection.doProjection, pc_processor/dataset/preprocess/projection.py:43-115,
 * for a CSR batch of scans.
 *
 * fov arguments are the Python floats RangeProjection.__init__ stores
 * (projection.py:29-35): abs(fov_left), fov_hori, abs(fov_down), fov_vert in
 * radians; they are rounded to float32 exactly where numpy does.
 * z-buffer rule: minimum depth, then minimum point index (projection.py:94 is
 * an unstable sort, so ties are otherwise undefined).
 * status_flags[0] bit 0 is set when a pixel coordinate is NaN (depth == 0), the
 * case in which the reference crashes.
 */
size_t c3d_project_workspace_bytes(int batch, int proj_h, int proj_w);
/* 1 if workspace_flags bit 3 (cluster form: the scan's z-buffer in the distributed shared memory
 * of an 8-CTA thread-block cluster, workspace untouched) applies to this shape. */
int c3d_project_cluster_supported(int c_in, int proj_h, int proj_w);

int c3d_project_batch(
    const float* points,          /* [total_points, c_in] x,y,z,(intensity,...)  */
    int c_in,                     /* >= 3; 4 is the fast path                    */
    const int32_t* offsets,       /* [batch+1]                                   */
    int batch,
    int64_t total_points,
    const float* depth_override,  /* [total_points] or NULL (projection.py:46)   */
    double abs_fov_left, double fov_hori, double abs_fov_down, double fov_vert,
    int proj_h, int proj_w,
    float* proj_range,            /* [batch, H, W]        -1 where empty         */
    float* proj_pointcloud,       /* [batch, H, W, c_in]  -1 where empty         */
    int32_t* proj_idx,            /* [batch, H, W]        -1 where empty         */
    int32_t* proj_mask,           /* [batch, H, W]        proj_idx > 0 (:113)    */
    int32_t* uproj_x_idx,         /* [total_points]  cached_data (:87-89)        */
    int32_t* uproj_y_idx,         /* [total_points]                              */
    float* uproj_depth,           /* [total_points]                              */
    void* workspace,              /* c3d_project_workspace_bytes                 */
    int workspace_flags,          /* bit 0: workspace is as a previous call OF THE SAME
                                     FORM (bit 2) left it
                                     (all 0xFF), so the 8 B/pixel memset is
                                     skipped (0: fresh memory); bit 1: evaluate the
                                     angles of every point in fp64 (default: only
                                     inside the guard band of a pixel boundary --
                                     same pixels); bit 2: the fused
                                     persistent kernel instead of the two-kernel form
                                     (same results; kept for A/B timing, slower); bit 3: the
                                     cluster form where c3d_project_cluster_supported(): the
                                     scan's z-buffer in the distributed shared memory of an
                                     8-CTA thread-block cluster, workspace untouched (same
                                     results; kept for A/B timing, slower) */
    int32_t* status_flags,        /* [1], caller-zeroed                          */
    void* cofill_ptr,             /* carried fill (see c3d_knn_batch): a 16 B aligned buffer
                                     zeroed meanwhile by the two passes' CTAs, or NULL       */
    size_t cofill_bytes,          /* multiple of 16                                        */
    void* stream);

/* ---------------------------------------------------------------- f1 ----
 * Projection fused with its caller: instead of the (H,W,4) projected
 * point cloud, the resolve pass writes what the data path builds from it --
 * pc_processor/dataset/semantic_kitti/wss_sem_kitti_loader.py:124-132 (eval / train label
 * images = per-point labels of the winning points), :159-172 (input
 * [range, x, y, z, intensity*(intensity != -1)], channel-major) and
 * tasks/weak_segmentation/trainer.py:600-608 (int64 labels; with img_mean/img_std:
 * (feature - mean) / std * (eval_label > 0)).  Points are [total, 4] x,y,z,intensity.
 * Nullable: depth_override, sem_label, weak_label, img_mean+img_std, train_label,
 * eval_label, proj_range, proj_idx.
 */
int c3d_project_assemble_batch(
    const float* points, const int32_t* offsets, int batch, int64_t total_points,
    const float* depth_override,
    const void* sem_label,        /* [total_points] full labels (mapped to [0, C))        */
    const void* weak_label,       /* [total_points] weak labels, 0 = unlabelled            */
    int label_is_u8,              /* 0: labels are int32 (the loaders' dtype); 1: uint8 --
                                     a quarter of the host->device bytes                   */
    const float* img_mean, const float* img_std,   /* [5] each (config sensor.img_mean)   */
    double abs_fov_left, double fov_hori, double abs_fov_down, double fov_vert,
    int proj_h, int proj_w,
    float* feature,               /* [batch, 5, H, W]                                      */
    int64_t* train_label,         /* [batch, H, W]                                         */
    int64_t* eval_label,          /* [batch, H, W]                                         */
    float* proj_range,            /* [batch, H, W]                                         */
    int32_t* proj_idx,            /* [batch, H, W]                                         */
    int32_t* uproj_x_idx, int32_t* uproj_y_idx, float* uproj_depth,   /* [total_points]   */
    void* workspace, int workspace_flags, int32_t* status_flags,
    void* cofill_ptr, size_t cofill_bytes,   /* carried fill, as in c3d_project_batch */
    void* stream);

/* ---------------------------------------------------------------- a4 ----
 * KNN.forward, pc_processor/postproc/knn.py:54-142, for a CSR batch of scans.
 *
 * inv_gauss is (1 - get_gaussian_kernel(search, sigma)) flattened row-major
 * (knn.py:11-33,102-104), computed by the host with the reference's formula.
 * Tie rule: k smallest by (distance, window slot); vote argmax = first maximum.
 * pxy_is_i64 selects the dtype of px / py, label_is_i64 (bit 0) that of proj_argmax /
 * out_labels (int64 = the reference's dtypes, int32 = this library's
 * projection outputs).  label_is_i64 bit 1: out_labels is uint8 whatever proj_argmax is
 * (class ids < 256: an eighth of the device->host bytes of the reference's int64).
 *
 * Co-scheduled fill: the vote is ALU bound and leaves HBM idle, so the caller may hand it
 * an unrelated zero fill -- in the hot-path step the dense gradient buffer of
 * c3d_proto_loss_backward (then called with grad_is_zeroed=1).  [cofill_ptr, +cofill_bytes)
 * is zeroed by the time the kernel completes; pass NULL / 0 for the plain vote.
 */
int c3d_knn_batch(
    const float* proj_range,      /* [batch, H, W]                               */
    const void* proj_argmax,      /* [batch, H, W] i64 or i32                    */
    const float* unproj_range,    /* [total_points]                              */
    const void* px, const void* py, /* [total_points] i64 or i32                 */
    const int32_t* offsets,       /* [batch+1]                                   */
    int batch, int64_t total_points,
    int proj_h, int proj_w,
    int knn, int search, float cutoff, int nclasses,
    const float* inv_gauss,       /* [search*search]                             */
    int pxy_is_i64, int label_is_i64,
    void* out_labels,             /* [total_points] i64 or i32, in [1, C-1]      */
    void* cofill_ptr,             /* 16 B aligned buffer to zero meanwhile, or NULL */
    size_t cofill_bytes,          /* multiple of 16                              */
    void* stream);

/* ---------------------------------------------------------------- f2 ----
 * Un-projection gather + confusion matrix:
 * `argmax_2d[ii, uproj_y_idx[ii], uproj_x_idx[ii]]` per scan,
 * tasks/weak_segmentation/trainer.py:714-724, and IOUEval.addBatch,
 * pc_processor/metrics/iou_eval.py:35-58 (conf[pred, gt] += 1; rows = prediction,
 * columns = ground truth).  conf_matrix is ACCUMULATED into (int64, caller-zeroed at
 * reset); either output may be NULL.  status_flags bit 0: pixel index outside the
 * image, bit 1: class outside [0, C).
 */
int c3d_unproject_confusion_batch(
    const void* proj_argmax,      /* [batch, H, W] i64 or i32                    */
    const void* px, const void* py, /* [total_points] i64 or i32                 */
    const void* labels,           /* [total_points] i64 or i32, or NULL          */
    const int32_t* offsets, int batch, int64_t total_points, int proj_h, int proj_w,
    int nclasses, int argmax_is_i64, int pxy_is_i64, int label_is_i64,
    void* unproj_argmax,          /* [total_points], dtype of proj_argmax, or NULL */
    int64_t* conf_matrix,         /* [C, C] accumulated, or NULL                 */
    int32_t* status_flags, void* stream);

/* ---------------------------------------------------------------- f3 ----
 * Trainer.entropy_based_selection, tasks/weak_segmentation/trainer.py:447-518 (SURVEY.md
 * 8f-3): per (scan, class present in train_label) int(count * select_ratio) of the pixels
 * predicted as that class are drawn without replacement with weights exp(-entropy); the
 * result is the pseudo-label image (weak ground truth kept) and its mask.
 * `noise` [B, C, H*W] injects the Exp(1) draws torch.multinomial would make in iteration
 * (scan, class) -- selection is topk(weight / noise) -- or NULL for Philox draws from `seed`.
 */
size_t c3d_entropy_select_workspace_bytes(int batch, int n_classes, int hw);

int c3d_entropy_select_batch(
    const float* probs,           /* [B, C, H, W] softmax output                 */
    const int64_t* train_label,   /* [B, H, W]                                   */
    const uint8_t* wss_mask,      /* [B, H, W] bool: weak label present          */
    const uint8_t* eval_mask,     /* [B, H, W] bool                              */
    int batch, int n_classes, int proj_h, int proj_w, int ignore_cls, float select_ratio,
    const float* noise, uint64_t seed,
    void* workspace,              /* c3d_entropy_select_workspace_bytes, 256 B aligned */
    int64_t* out_label,           /* [B, H, W] pseudo label (:512-515)           */
    uint8_t* out_mask,            /* [B, H, W] bool: label != ignore_cls (:516)  */
    void* stream);

/* ---------------------------------------------------------------- f4 ----
 * Lovasz-softmax loss, pc_processor/loss/lovasz_softmax.py:51-157 (lovasz_grad,
 * lovasz_softmax_flat, flatten_probas), as the trainer calls it
 * (tasks/weak_segmentation/trainer.py:362-364,650): class probabilities, weak labels,
 * `ignore`, classes = "present" (classes_all = 0), "all" (1) or a list (2), per_image = False.
 * Forward keeps per-element gradient entries in the workspace; backward zero-fills the
 * dense (B,C,H,W) gradient (unless grad_is_zeroed) and scatters P x C entries scaled by
 * grad_out.  The rank pass is quadratic in the number of valid pixels P, so max_valid is
 * capped at 32768 (weak labels: ~1e3 per batch); more valid pixels set flag bit 0 and are
 * dropped, max_valid above the cap returns C3D_UNSUPPORTED.  Ties in the errors rank by
 * ascending pixel index (torch.sort is unstable).  ignore < 0: no ignored label.
 */
size_t c3d_lovasz_workspace_bytes(int n_classes, int64_t max_valid);

int c3d_lovasz_forward(
    const float* probs,           /* [B, C, H, W] class probabilities            */
    const int64_t* labels,        /* [B, H, W]                                   */
    int batch, int n_classes, int proj_h, int proj_w, int ignore,
    int classes_all,              /* 0: classes = 'present', 1: 'all', 2: the list in class_mask */
    uint64_t class_mask,          /* bit c set: class c is in the list (:117-122)        */
    int64_t max_valid,
    void* workspace,              /* c3d_lovasz_workspace_bytes, 256 B aligned   */
    float* loss_out,              /* [1]                                         */
    void* stream);

int c3d_lovasz_backward(
    int batch, int n_classes, int proj_h, int proj_w, int classes_all, uint64_t class_mask,
    int64_t max_valid,
    void* workspace,              /* as left by c3d_lovasz_forward               */
    const float* grad_out,        /* [1] upstream gradient                       */
    float* grad_probs,            /* [B, C, H, W], 16 B aligned                  */
    int grad_is_zeroed, void* stream);

/* Synchronous: host_info4 = {valid pixels, classes averaged, flags (1: more than max_valid
 * valid pixels, 2: none), 0}. */
int c3d_lovasz_info(const void* workspace, int32_t* host_info4, void* stream);

/* ---------------------------------------------------------------- a2 ----
 * ContrastMEMLoss.forward, pc_processor/loss/contrast_pixel_loss.py:27-195,
 * and its autograd (gradient w.r.t. feats only; the bank is detached at
 * tasks/weak_segmentation/trainer.py:675-678).
 *
 * Anchors: with keep == NULL the A anchors per (scan, class) segment are drawn
 * on the device (entropy-weighted, with replacement, Philox keyed by `seed`),
 * the statistical equivalent of torch.multinomial (:114-116).  With keep != NULL
 * the caller injects the sampled pixel indices, shape [keep_rows, num_anchor],
 * rows in the reference's X_ptr order (scan ascending, class ascending); every
 * index must be a kept pixel of that segment's class (what multinomial returns).
 * Duplicate anchors are evaluated once and weighted by their multiplicity.
 * The sub-prototype randperm (:142-143) is not applied (it only reorders sums).
 *
 * With need_grad the forward also evaluates, per distinct sampled pixel, the
 * gradient row d loss / d feats[:, pixel] for a unit upstream gradient; backward
 * is then the dense zero fill plus D strided stores per row, scaled by grad_out.
 * The workspace carries those rows from forward to backward; it must stay
 * untouched in between.  ws[0..2] (int32) = {T segments, labelled pixels,
 * flags}; flags: 1 no anchor (loss is NaN; the reference crashes), 2 injected
 * index not in its segment, 4 keep_rows != T, 8 label outside [0, C), 32 backward
 * called after a forward without need_grad.
 */
size_t c3d_proto_loss_workspace_bytes(int batch, int n_classes, int hw, int dim, int sub_protos,
                                      int num_anchor);

int c3d_proto_loss_forward(
    const float* feats,           /* [B, D, H, W]                                */
    const float* probs,           /* [B, C, H, W] softmax output (:46-49)        */
    const int64_t* labels,        /* [B, H, W]                                   */
    const uint8_t* keep_mask,     /* [B, H, W] bool, or NULL (:36-38)            */
    const float* proto_queue,     /* [C, M, D]                                   */
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos,
    int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep,          /* [keep_rows, num_anchor] or NULL             */
    int keep_rows, uint64_t seed,
    int need_grad,                /* bit 0: also evaluate the gradient rows (training);
                                     bit 1: run the two row x bank products on the tensor
                                     cores (mma.sync m16n8k8, 3xTF32 split: fp32-level
                                     accuracy) instead of the FFMA form -- see DESIGN.md  */
    void* workspace,              /* c3d_proto_loss_workspace_bytes, 256 B aligned */
    float* loss_out,              /* [1]                                         */
    void* stream);

/* The same forward in two parts, so that a scheduler can place other work between them:
 * phases = 1 runs the selection (label split, anchor sampling), 2 the loss / gradient rows
 * (needs a preceding phase 1 on the same workspace), 3 both. */
int c3d_proto_loss_forward_phase(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* proto_queue, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, uint64_t seed, int need_grad, int phases,
    void* workspace, float* loss_out, void* stream);

int c3d_proto_loss_backward(
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos, int num_anchor,
    void* workspace,              /* as left by c3d_proto_loss_forward(need_grad=1) */
    const float* grad_out,        /* [1] upstream gradient of the scalar loss    */
    float* grad_feats,            /* [B, D, H, W] dense, fully written           */
    int grad_is_zeroed,           /* 1: caller already zero-filled grad_feats
                                     (c3d_zero_fill, e.g. on a side stream)      */
    void* stream);

/* The dense-gradient zero fill (128-bit streaming stores) as its own entry point. */
int c3d_zero_fill(void* dst, size_t nbytes, void* stream);

/* The same fill as a background ("daemon") kernel: one warp per CTA, ctas_per_sm CTAs per
 * SM, resident for as long as the fill takes, so that it can run UNDER the other kernels of
 * a step (launch it first, on its own stream) instead of holding every SM slot.
 * mode 0: one thread per CTA issues cp.async.bulk shared->global copies of a zero page of
 * page_bytes (multiple of 1024, <= 65536), at most `inflight` (1, 2, 4, 8, 16) copies in
 * flight per CTA; mode 1: the warp issues 128-bit streaming stores (page_bytes / inflight
 * ignored).  nbytes must be a multiple of 16. */
int c3d_zero_fill_background(void* dst, size_t nbytes, int mode, int ctas_per_sm, int page_bytes,
                             int inflight, void* stream);

/* Placement-proof daemon: launch_per_sm * 148 one-warp CTAs are launched, each claims the SM
 * it lands on and retires at once if that SM already runs max_per_sm of them; pages (TMA
 * bulk stores of a zero page of page_bytes) are handed out chunk_pages at a time by a global
 * counter.  ctrl_ws: 1 KB of device scratch (zeroed here); debug: NULL, or
 * [launch_per_sm * 148][4] int64 {smid, first ns, last ns, pages} per CTA that worked. */
int c3d_zero_fill_daemon(void* dst, size_t nbytes, int max_per_sm, int launch_per_sm, int page_bytes,
                         int chunk_pages, void* ctrl_ws, void* debug, void* stream);

/* Holds `stream` for ns nanoseconds (one spinning thread). */
int c3d_delay(unsigned long long ns, void* stream);

/* Synchronous: copies {T, labelled pixels, flags, 0} to host_info4 (host). */
int c3d_proto_loss_info(const void* workspace, int32_t* host_info4, void* stream);

/* Exports the first `capacity` labelled-pixel slots of the last forward, sorted
 * by (class, scan, pixel): pix = scan*H*W + pixel, cls = class, cnt = number of
 * anchors that hit the slot (sums to num_anchor per segment). */
int c3d_proto_loss_rows(const void* workspace, int batch, int dim, int hw, int n_classes,
                        int sub_protos, int num_anchor, int64_t capacity, int32_t* pix,
                        int32_t* cls, int32_t* cnt, void* stream);

/* ---------------------------------------------------------------- a3 ----
 * EMA prototype update: the pre-step of SalsaNextProto.forward,
 * pc_processor/models/salsanext_proto.py:497-510, prototype_learning :337-402
 * (same logic in rangenet_proto.py:460-567, squeezesegv3_Proto.py:253-351),
 * momentum_update :19-31 and distributed_sinkhorn, models/sinkhorn.py:5-33 --
 * restricted to the pixels with label != ignore_label, the only rows that
 * influence the update.
 *
 * c3d_proto_ema_accumulate produces the all-reduce payload
 *     packed = [ C*M*D feature sums | C*M counts ]   (float32)
 * i.e. per (class, sub-prototype) the sum of the LayerNorm+L2-normalised
 * features of correctly predicted pixels assigned to it, and how many.  Ranks
 * sum `packed` (one NCCL all-reduce) and then each calls c3d_proto_ema_apply,
 * which normalises the sums, applies the EMA where count != 0 and renormalises
 * every prototype (:379-394).  On one rank this equals the reference.
 *
 * assign_mode: 0 = one_hot(argmax Q) (sinkhorn.py:30, deterministic),
 *              1 = F.gumbel_softmax(hard) with the caller's noise `gumbel`
 *                  [rows, M], rows in (class, global pixel) order (sinkhorn.py:31),
 *              2 = the same with Philox noise drawn on the device from `seed`.
 * ws[0..2] (int32) = {non-empty (class, scan) segments, labelled rows, flags};
 * flags: 1 no labelled pixel, 8 label outside [0, C), 16 rows > max_rows (the
 * update is then skipped entirely: packed is all zero).
 */
size_t c3d_proto_ema_workspace_bytes(int batch, int n_classes, int hw, int dim, int sub_protos,
                                     int64_t max_rows);

int c3d_proto_ema_accumulate(
    const float* embedding,       /* [B, D, H, W] (feat_2d)                      */
    const int64_t* label,         /* [B, H, W]                                   */
    const float* prototypes,      /* [C, M, D] current bank (any norm)           */
    const float* ln_d_w, const float* ln_d_b, /* feat_norm = LayerNorm(D) (:327) */
    const float* ln_c_w, const float* ln_c_b, /* mask_norm = LayerNorm(C) (:328) */
    float ln_eps,
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos,
    int ignore_label, int64_t max_rows,
    const float* gumbel, int assign_mode, uint64_t seed,
    void* workspace,              /* c3d_proto_ema_workspace_bytes, 256 B aligned */
    float* packed,                /* [C*M*D + C*M]                               */
    float* proto_target,          /* [B*H*W] or NULL (:346,390-392)              */
    void* stream);

/* The same accumulation from the DENSE tensors the reference's forward has already built
 * (salsanext_proto.py:497-510), i.e. with the exact arguments of
 * `prototype_learning(out_feat, nearest_proto_distance, label, eval_mask, feat_proto_sim)`
 * (:337-339; eval_mask is unused by the reference): out_feat [n, D] LayerNorm+L2-normalised
 * rows (n = B*H*W, pixel-major), nearest [B, C, H, W], feat_proto_sim [n, M, C].  Only the
 * labelled rows are read.  Workspace, packed, proto_target, assign_mode as above. */
int c3d_proto_ema_accumulate_dense(
    const float* out_feat, const float* nearest, const int64_t* label, const float* feat_proto_sim,
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos, int ignore_label,
    int64_t max_rows, const float* gumbel, int assign_mode, uint64_t seed, void* workspace,
    float* packed, float* proto_target, void* stream);

int c3d_proto_ema_apply(
    const float* prototypes_in,   /* [C, M, D]                                   */
    const float* packed,          /* [C*M*D + C*M], summed over ranks            */
    int n_classes, int sub_protos, int dim, int ignore_label, double momentum,
    float* prototypes_out,        /* [C, M, D] (may alias prototypes_in)         */
    float* normalised_out,        /* [C, M, D] or NULL: F.normalize(prototypes_out), the form
                                     the bank's readers use (contrast_pixel_loss.py:167;
                                     salsanext_proto.py:502) -- pass it as `bank_n` to
                                     c3d_proto_step and no normalise kernel is launched   */
    uint64_t* seed_counters,      /* [2] or NULL: device step counters of c3d_proto_step;
                                     [1] (the EMA's noise stream) is advanced here        */
    void* stream);

/* Multi-GPU form of c3d_proto_ema_apply WITHOUT a collective call ("only the K x D
 * prototype sums and counts combined ... so every rank applies an identical EMA update"): the
 * all-reduce of `packed` and the EMA are ONE kernel over peer memory.  It replaces, for ranks of
 * one NVLink / NVSwitch box, `dist.all_reduce(...)` of salsanext_proto.py:397-400 plus the EMA.
 * Every rank owns an exchange buffer of c3d_peer_exchange_bytes() allocated by c3d_peer_alloc
 * (a dedicated cudaMalloc, zeroed), exports it (c3d_peer_export -> 64-byte CUDA IPC handle, sent
 * to the other processes by any means, e.g. torch.distributed.all_gather_object) and opens the
 * others' (c3d_peer_import).  Per call each rank PUSHES its payload as 8-byte {value, step flag}
 * pairs into its slot of every rank's buffer (posted NVLink stores: no fence, no round trip),
 * then polls its own buffer until every rank's pairs carry this step's flag (bounded by
 * timeout_s; a missing peer sets bit r of state[3] instead of hanging), sums the values in rank
 * order 0..world-1 -- identical order, hence bit-identical banks, on every rank -- writes the
 * sum back to `packed` and applies the EMA.  `state` = c3d_peer_state_bytes() of zero-initialised
 * device memory ([0] step counter advanced by the kernel, so a captured CUDA graph replays
 * correctly; [1] ticket; [3] error bits; then scratch).  All ranks must make the same sequence of
 * calls; world <= 8 (one NVSwitch box). */
size_t c3d_peer_exchange_bytes(int n_classes, int sub_protos, int dim, int world);
size_t c3d_peer_state_bytes(int n_classes, int sub_protos);
int c3d_peer_alloc(size_t bytes, void** ptr);
int c3d_peer_free(void* ptr);
int c3d_peer_export(void* ptr, unsigned char* handle64);
int c3d_peer_import(const unsigned char* handle64, void** ptr);
int c3d_peer_close(void* ptr);
int c3d_proto_ema_apply_peers(
    const float* prototypes_in,   /* [C, M, D]                                               */
    float* packed,                /* [C*M*D + C*M] this rank's sums and counts; holds the sum
                                     over ranks afterwards                                    */
    void* const* peer_bufs,       /* HOST array [world]: every rank's exchange buffer as mapped
                                     in this process (own buffer at index `rank`)             */
    int rank, int world, int32_t* state,
    int n_classes, int sub_protos, int dim, int ignore_label, double momentum,
    float* prototypes_out, float* normalised_out, uint64_t* seed_counters,
    double timeout_s,             /* spin bound per call (<= 0: 2 s)                         */
    void* stream);

/* Optional pre-pass of c3d_knn_batch: bins the points of every scan by (row, 32-pixel column
 * segment) -- a per-scan counting sort, three small kernels -- and writes one 16-byte record
 * {range f32, x i32, y i32, original point index i32} per point in binned, still scan-major
 * order.  Passing the records to c3d_knn_batch as `px` with pxy_is_i64 = 2 (unproj_range and py
 * are then ignored) makes the lanes of a warp gather from shared cache lines; the labels come
 * out at the original point indices, bit-identical to the plain call.  Measured at batch 64: the
 * vote + fill kernel 673 -> 612 us, the three binning kernels 201 us -- a net loss inside the
 * step, so the step pipeline does not use it; it pays when the same points are voted on
 * repeatedly (e.g. several argmax images per projection). */
size_t c3d_knn_sort_workspace_bytes(int batch, int64_t total_points, int proj_h, int proj_w);
int c3d_knn_sort_points(const float* unproj_range, const void* px, const void* py,
                        const int32_t* offsets, int batch, int64_t total_points, int proj_h, int proj_w,
                        int pxy_is_i64, void* workspace, void* sorted_records /* [total_points] x 16 B */,
                        void* stream);

/* -------------------------------------------------------- a2 + a3 fused ----
 * The prototype step of a training iteration: the EMA update inside model.forward
 * (salsanext_proto.py:520-527) followed by ContrastMEMLoss on the UPDATED bank
 * (tasks/weak_segmentation/trainer.py:675-686), for callers whose two operators see the same
 * label image (weak labels; `entropy_selection` off).  Everything derived from the labels is
 * shared: one label split feeds the loss's anchor sampler and the EMA's row kernels.
 * phases (bit mask, each phase needs the earlier ones on the same workspace):
 *   1 split, 2 anchor sampling, 4 EMA accumulation -> packed, 8 loss + gradient rows.
 * Between 4 and 8 the caller all-reduces `packed` (multi-GPU) and calls c3d_proto_ema_apply
 * (in place on `prototypes`); after 8, c3d_proto_loss_backward / _info / _rows take the same
 * workspace (it begins with a loss workspace of the same shape).  Arguments as in
 * c3d_proto_loss_forward and c3d_proto_ema_accumulate; pointers a phase does not use may be NULL.
 * With keep_mask the masked labels are what BOTH operators see. */
size_t c3d_proto_step_workspace_bytes(int batch, int n_classes, int hw, int dim, int sub_protos,
                                      int num_anchor, int64_t max_rows);

int c3d_proto_step(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* prototypes, const float* ln_d_w, const float* ln_d_b, const float* ln_c_w,
    const float* ln_c_b, float ln_eps, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, const float* gumbel, int assign_mode, uint64_t seed,
    int64_t max_rows, int need_grad, int phases,
    const float* bank_n,          /* [C, M, D] or NULL: F.normalize(prototypes) as left by
                                     c3d_proto_ema_apply(normalised_out); NULL = normalised here */
    uint64_t* seed_counters,      /* [2] or NULL: device-side step counters added to `seed`
                                     ([0] anchor sampling, advanced by phase 2; [1] Gumbel noise,
                                     advanced by c3d_proto_ema_apply), so that replays of a
                                     captured CUDA graph draw fresh randomness              */
    void* workspace, float* packed, float* proto_target, float* loss_out,
    void* cofill_ptr, size_t cofill_bytes,   /* carried fill (see c3d_knn_batch) for ONE kernel of the
                                                call: the rows kernel of the latest phase asked for,
                                                else the label split; NULL / 0 for none          */
    void* stream);

/* F.normalize(x, p=2, dim=-1) of `rows` bank rows (eps 1e-12): the `bank_n` of c3d_proto_step for a
 * bank that did not come out of c3d_proto_ema_apply (e.g. the initial one). */
int c3d_proto_bank_normalise(const float* prototypes, int rows, int dim, float* normalised_out, void* stream);

/* Synchronous: copies {segments, labelled rows, flags, 0} to host_info4 (host). */
int c3d_proto_ema_info(const void* workspace, int32_t* host_info4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COARSE3D_B200_H_ */
