"""The per-scan hot path as one device-resident step, eager or CUDA-graph captured.

A step is what BASELINE.json's metric counts: for one batch of scans,
  projection  ->  prototype loss forward + backward  ->  EMA prototype update
  (one all-reduce of the packed sums when world_size > 1)  ->  KNN vote.
The range-image CNN between projection and loss is out of scope (SURVEY.md 2,
row 9), so its outputs (features, softmax probabilities, argmax image) are
synthetic resident tensors, exactly as BASELINE.json's configs prescribe.
"""
import os as _os

import numpy as np
import torch

from . import distributed, ops, synth


class StepInputs:
    """One set of resident inputs for a batch of `batch` scans of `shape`."""

    def __init__(self, shape: synth.ScanShape, batch, dim, sub_protos, seed0, device,
                 feats=None, share_probs=None, sensor_order=False):
        self.shape, self.batch, self.dim = shape, batch, dim
        pts, offs, _full, weak = synth.make_batch(shape, batch, seed0, sensor_order=sensor_order)
        H, W, C = shape.proj_h, shape.proj_w, shape.n_classes
        self.host_points, self.host_offsets, self.host_weak = pts, offs, weak
        self.points = torch.from_numpy(pts).to(device)
        self.offsets = torch.from_numpy(offs).to(device)
        self.weak = torch.from_numpy(weak).to(device)
        self.full = torch.from_numpy(_full).to(device)
        self.seed0 = seed0
        self.n_points = int(pts.shape[0])
        g = torch.Generator(device=device).manual_seed(seed0)
        self.feats = feats if feats is not None else \
            torch.randn((batch, dim, H, W), device=device, generator=g)
        if share_probs is not None:
            self.probs = share_probs
        else:
            self.probs = torch.softmax(torch.randn((batch, C, H, W), device=device, generator=g), 1)
        self.argmax = torch.randint(0, C, (batch, H, W), device=device, generator=g)  # int64
        self.labels = None      # (B,H,W) int64: weak labels of the winning points
        self.keep_mask = None   # (B,H,W) bool
        self.labels_loss = None     # the contrastive loss's labels when they differ from the EMA's
        self.keep_mask_loss = None  # (pseudo-label regime, trainer.py:654-686)

    def derive_labels(self, proj, loss_label_frac=None):
        """Projected weak-label image, like the reference's loader builds it
        (wss_sem_kitti_loader.py:124-132): label of the point that won the pixel."""
        B = self.batch
        gidx = proj.proj_idx.long() + self.offsets[:-1].long().view(B, 1, 1)
        lab = torch.zeros_like(gidx)
        valid = proj.proj_idx >= 0
        lab[valid] = self.weak[gidx[valid]]
        self.labels = lab.contiguous()
        self.keep_mask = (lab > 0).contiguous()
        if loss_label_frac is not None:
            # stand-in for entropy_based_selection's output: the weak labels plus the ground-truth
            # class on a random `loss_label_frac` of the occupied pixels
            g = torch.Generator(device=lab.device).manual_seed(self.seed0 + 17)
            pick = valid & (torch.rand(lab.shape, device=lab.device, generator=g) < loss_label_frac)
            full_img = torch.zeros_like(gidx)
            full_img[valid] = self.full[gidx[valid]]
            ll = torch.where(pick, full_img, lab)
            self.labels_loss = ll.contiguous()
            self.keep_mask_loss = (ll > 0).contiguous()


def fill_partition(n_floats, shares, chains=True, page=2048):
    """Split a flat buffer of `n_floats` float32 into five contiguous ranges (lo, hi) for the
    carriers (projection, label split, EMA rows, loss rows, KNN vote): the first four take
    `shares[j]` of the buffer rounded DOWN to whole 8 KB pages (`page` floats), the vote takes the
    rest.  With chains=False the prototype chains' shares (indices 1..3) stay with the vote.
    The ranges tile [0, n_floats) exactly; every boundary but the last is page-aligned."""
    out, lo = [], 0
    for j, f in enumerate(shares):
        if j > 0 and not chains:
            f = 0.0
        hi = min(n_floats, lo + int(n_floats * max(f, 0.0)) // page * page)
        out.append((lo, hi))
        lo = hi
    out.append((lo, n_floats))
    return out


class HotPathStep:
    """Pre-allocated, allocation-free step over rotating input sets."""

    def __init__(self, shape, batch, dim=128, sub_protos=20, num_anchor=512, temperature=0.07,
                 momentum=0.999, n_sets=3, seed0=1000, device="cuda", group=None,
                 knn=(5, 5, 1.0, 1.0), concurrent=True, parts=None, bank_seed=7, sensor_order=False,
                 loss_label_frac=None):
        self.shape, self.batch, self.dim, self.M = shape, batch, dim, sub_protos
        self.device, self.group = torch.device(device), group
        self.knn_k, self.knn_s, self.knn_sigma, self.knn_cutoff = knn
        H, W, C = shape.proj_h, shape.proj_w, shape.n_classes
        self.fov = ops.Fov.from_degrees(shape.fov_up, shape.fov_down)
        self.cfg = ops.ProtoLossConfig(0, temperature, 0.07, num_anchor)
        self.momentum = momentum
        self.sets = []
        for i in range(n_sets):
            self.sets.append(StepInputs(shape, batch, dim, sub_protos, seed0 + 100 * i, self.device,
                                        sensor_order=sensor_order))
        n = self.sets[0].n_points
        assert all(s.n_points == n for s in self.sets)
        self.n_points = n
        self.proj_bufs = [ops.ProjectionBuffers(batch, n, 4, H, W, self.device) for _ in self.sets]
        for s, b in zip(self.sets, self.proj_bufs):
            s.derive_labels(ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b), loss_label_frac)
        # the bank is a model parameter: the SAME initial value on every rank (seed0 differs per
        # rank, it seeds the scans), kept identical afterwards by the summed update
        g = torch.Generator(device=self.device).manual_seed(bank_seed)
        self.protos = torch.nn.functional.normalize(
            torch.randn((C, sub_protos, dim), device=self.device, generator=g), dim=-1)
        self.ln_d = (torch.ones(dim, device=self.device), torch.zeros(dim, device=self.device))
        self.ln_c = (torch.ones(C, device=self.device), torch.zeros(C, device=self.device))
        self.max_rows = min(batch * H * W, 1 << 17)
        # fused prototype step (one label split for the EMA update and the loss); its workspace
        # begins with a loss workspace, so the unfused calls can run on it too
        # pseudo-label regime: the EMA sees the weak labels (model.forward, trainer.py:625-630), the
        # loss the expanded ones (:680-686): two label splits, so not the fused step
        self.loss_label_frac = loss_label_frac
        self.fused_step = _os.environ.get("C3D_FUSED_STEP", "1") == "1" and loss_label_frac is None
        self.loss_ws = ops.proto_step_workspace(batch, C, H * W, dim, sub_protos, num_anchor, self.max_rows,
                                                self.device)
        self.loss = torch.zeros((), device=self.device)
        self.grad_out = torch.ones((), device=self.device)
        self.grad = torch.empty((batch, dim, H, W), device=self.device)
        nws = ops.lib.c3d_proto_ema_workspace_bytes(batch, C, H * W, dim, sub_protos, self.max_rows)
        self.ema_ws = torch.empty((nws,), dtype=torch.uint8, device=self.device)
        K = C * sub_protos
        self.packed = torch.empty((K * dim + K,), device=self.device)
        # F.normalize(bank): written by every EMA apply, read by the loss (this step) and by the
        # similarity pre-step of the EMA (next step); device step counters feed the Philox streams
        self.bank_n = ops.bank_normalise(self.protos)
        self.seed_counters = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.device_seeds = _os.environ.get("C3D_DEVICE_SEEDS", "1") == "1"   # 0: `seed` alone decides the draws
        self.knn_out = torch.empty((n,), dtype=torch.int64, device=self.device)
        self.inv_gauss = (1 - ops.gaussian_kernel(self.knn_s, self.knn_sigma)).reshape(-1).to(self.device)
        self.daemon_ctrl = torch.zeros(256, dtype=torch.int32, device=self.device)
        self.daemon_dbg = None
        self.daemon_lead_ns = int(_os.environ.get("C3D_DAEMON_LEAD_NS", "0"))
        self.ev_fork0 = torch.cuda.Event()
        self.graphs = None
        self.concurrent = concurrent
        self.schedule = _os.environ.get("C3D_SCHEDULE", "fill_spread")
        # fill daemon (schedule "fill_daemon"): mode, CTAs per SM, zero-page bytes, copies in flight
        self.daemon = tuple(int(v) for v in _os.environ.get("C3D_DAEMON", "0,1,8192,4").split(","))
        # ablation switch for tools/ablate.py: which chains run (default: all)
        self.parts = set(parts) if parts else {"proj", "knn", "fill", "loss", "ema"}
        # Priorities: the latency-bound chains (loss, EMA) high, so their small CTAs
        # are placed first whenever the short CTAs of the fill / KNN retire.
        lo, hi = 0, -1
        fill_prio = hi if (_os.environ.get("C3D_FILL_PRIO", "0") == "1" or self.schedule == "fill_daemon") else lo
        self.side = [torch.cuda.Stream(self.device, priority=fill_prio),   # fill
                     torch.cuda.Stream(self.device, priority=lo),   # projection -> KNN
                     torch.cuda.Stream(self.device, priority=hi),   # EMA chain
                     torch.cuda.Stream(self.device, priority=hi)]   # loss chain
        (self.ev_fork, self.ev_fill, self.ev_proj, self.ev_ema, self.ev_loss,
         self.ev_resolved, self.ev_selected) = (torch.cuda.Event() for _ in range(7))
        self.ev_split = torch.cuda.Event()
        self.knn_after_select = _os.environ.get("C3D_KNN_AFTER_SELECT", "0") == "1"
        # the vote (+ its share of the fill) released only when the loss rows are done: the rows
        # kernels of both chains need whole SMs (shared memory) and, run next to the vote's
        # 30 k CTAs, stall it for longer than they take ("auto": batches >= 48, measured)
        kar = _os.environ.get("C3D_KNN_AFTER_ROWS", "auto")
        self.knn_after_rows = (batch >= 20) if kar == "auto" else kar == "1"
        self.ev_rows = torch.cuda.Event()
        # the vote on points binned by (row, 32-pixel segment) (c3d_knn_sort_points): measured a net
        # loss inside the step (vote + fill 673 -> 612 us at batch 64, binning 201 us), off by default
        self.knn_binned = _os.environ.get("C3D_KNN_BINNED", "0") == "1"
        self.knn_sort_ws = ops.knn_sort_workspace(batch, n, H, W, self.device)
        self.knn_records = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        # N > 1: the all-reduce of the packed sums fused with the EMA over peer memory (default),
        # or ncclAllReduce + c3d_proto_ema_apply (C3D_PEER_EXCHANGE=0)
        self.peer = None
        if distributed.world(group)[1] > 1 and _os.environ.get("C3D_PEER_EXCHANGE", "1") == "1":
            self.peer = distributed.PeerExchange(C, sub_protos, dim, self.device, group)
            if not self.peer.available:     # e.g. no CUDA IPC in this container: all ranks fall back together
                self.peer = None
        # scans in the first of two KNN launches.  With the vote held back until the rows kernels
        # are done (below), the first launch is NOT held: a few scans' votes fit between the
        # projection and the loss rows (-26 us per step for 4 scans at batch 20..40, nothing from
        # batch 48 on: profiles/r2/fill_spread_early*.txt)
        ks = _os.environ.get("C3D_KNN_SPLIT", "auto")
        self.knn_split = (4 if 20 <= batch < 48 else 0) if ks == "auto" else int(ks)
        # schedule "fill_spread": fractions of the dense-gradient zero fill carried by the
        # projection's two passes, the label split, the EMA rows kernel and the loss rows kernel
        # (the KNN vote carries the rest).  Defaults from the sweeps in profiles/r2/fill_spread_*.txt:
        # only the loss rows kernel's carrier warp pays -- half of the fill at batch 8, where the
        # vote is short, a fifth once the vote is held back until the rows kernels are done.
        shares = _os.environ.get("C3D_FILL_SHARES", "auto")
        if shares == "auto":
            self.fill_shares = (0.0, 0.0, 0.0, 0.5 if batch <= 12 else (0.3 if batch < 20 else 0.2))
        else:
            self.fill_shares = tuple(float(v) for v in shares.split(","))
        torch.cuda.synchronize(self.device)

    def _fill_slices(self, chains=True):
        """The gradient buffer as five flat 8 KB-aligned slices: (proj, split, ema, loss, knn);
        without the prototype chains their shares stay with the vote."""
        flat = self.grad.view(-1)
        return [flat[lo:hi] if hi > lo else None
                for lo, hi in fill_partition(flat.numel(), self.fill_shares, chains)]

    # bytes the reference dtypes move per step
    def algorithmic_bytes(self):
        HW = self.shape.proj_h * self.shape.proj_w
        B, N, D = self.batch, self.n_points, self.dim
        return {
            "project": 28 * N + 28 * B * HW,
            "knn": 12 * B * HW + 28 * N,
            "loss": (4 * D + 9) * B * HW,
            "loss_grad_fill": 4 * D * B * HW,
        }

    def fill_bytes_by_carrier(self):
        """Bytes of the dense-gradient zero fill each kernel of the current schedule carries."""
        names = ("project", "label_split", "ema_rows", "loss_rows", "knn_vote")
        total = self.grad.numel() * 4
        if self.schedule == "fill_in_knn":
            return dict(zip(names, (0, 0, 0, 0, total)))
        if self.schedule != "fill_spread":
            return {"fill_kernel": total}
        sl = self._fill_slices(chains=self.fused_step and {"loss", "ema"} <= self.parts)
        return {k: (0 if v is None else v.numel() * 4) for k, v in zip(names, sl)}

    def run(self, i, seed=0):
        """Enqueue one step (no allocation, no sync).  The four independent chains
        -- projection -> KNN, the dense-gradient zero fill, loss forward -> backward,
        EMA update -- are forked onto side streams and joined before returning, so
        the latency-bound small kernels overlap the bandwidth-bound fill; under
        CUDA-graph capture they become parallel branches of the graph."""
        s, b = self.sets[i % len(self.sets)], self.proj_bufs[i % len(self.sets)]
        H, W, C = self.shape.proj_h, self.shape.proj_w, self.shape.n_classes
        cur = torch.cuda.current_stream(self.device)
        spread = (self.schedule == "fill_spread" and {"proj", "knn", "fill"} <= self.parts and
                  not self.knn_after_select)
        fused = (self.schedule == "fill_in_knn" or spread) and {"knn", "fill"} <= self.parts
        self._fs = [None, None, None, None, self.grad]
        if spread:
            self._fs = self._fill_slices(chains=self.fused_step and {"loss", "ema"} <= self.parts)
        if not self.concurrent:
            pr = ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b, cofill=self._fs[0])
            self._bin(pr, s)
            if self.fused_step:
                self._step(ops.STEP_SPLIT, s.labels, s, seed, cofill=self._fs[1])
                self._step(ops.STEP_SAMPLE | ops.STEP_ACCUMULATE, s.labels, s, seed, cofill=self._fs[2])
                self._ema_finish()
                self._step(ops.STEP_LOSS_ROWS, s.labels, s, seed, cofill=self._fs[3])
            else:
                self._ema(s, seed)            # the bank is updated in the model forward, before the loss
                self._loss_fwd(s, seed)
            if fused:
                self._knn(s, pr, C, cofill=self._fs[4])
            ops.proto_loss_backward_raw(s.feats.shape, self.cfg, C, self.M, self.loss_ws, self.grad_out,
                                        self.grad, grad_is_zeroed=fused)
            if not fused:
                self._knn(s, pr, C)
            return pr
        return self._run_concurrent(s, b, seed, fused, cur)

    def _run_concurrent(self, s, b, seed, fused, cur):
        H, W, C = self.shape.proj_h, self.shape.proj_w, self.shape.n_classes
        st_fill, st_proj, st_ema, st_loss = self.side
        P = self.parts
        pr = None
        sched = self.schedule
        if sched == "fill_daemon" and self.daemon_lead_ns > 0 and "fill" in P:
            # The daemon first, alone: launched next to other grids its one-warp CTAs are packed
            # onto a handful of SMs (measured: 148 CTAs on 6-13 SMs) and the fill crawls at the
            # per-SM copy-engine rate.  A short hold lets it become resident on every SM.
            self.ev_fork0.record(cur)
            st_fill.wait_event(self.ev_fork0)
            with torch.cuda.stream(st_fill):
                self._daemon()
                self.ev_fill.record(st_fill)
            ops.delay(self.daemon_lead_ns)
        self.ev_fork.record(cur)
        for st in self.side:
            st.wait_event(self.ev_fork)
        hold = fused and self.knn_after_select and "loss" in P
        defer_knn = (fused and self.knn_after_rows and self.fused_step and {"loss", "ema"} <= P and not hold)
        if hold:
            # selection part of the loss first; its event releases the KNN + fill kernel
            with torch.cuda.stream(st_loss):
                self._loss_fwd(s, seed, phases=1)
                self.ev_selected.record(st_loss)
        if fused:
            # The KNN vote (ALU bound) carries the zero fill (HBM bound): one kernel, both
            # pipes busy.  As two kernels they serialise: the fill's CTAs occupy every SM slot.
            with torch.cuda.stream(st_proj):
                if "proj" in P:
                    pr = ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b, cofill=self._fs[0])
                self._bin(pr if pr is not None else self._last_proj(b), s)
                if hold:
                    # The vote's 3750 CTAs keep every SM full until its grid is drained, and the
                    # loss-rows kernel needs a whole SM (217 KB of shared memory): released
                    # together, the high-priority rows CTAs are placed first.
                    st_proj.wait_event(self.ev_selected)
                if not defer_knn:
                    self._knn(s, pr if pr is not None else self._last_proj(b), C, cofill=self._fs[4])
                    self.ev_proj.record(st_proj)
                else:   # the first knn_split scans are voted while the chains are between rows kernels
                    self._knn(s, pr if pr is not None else self._last_proj(b), C, cofill=self._fs[4], part="early")
            if not defer_knn:
                st_fill.wait_event(self.ev_proj)
                self.ev_fill.record(st_fill)
        elif sched == "fill_daemon":
            # The fill as a minimal-footprint persistent kernel (one warp per SM), launched first
            # and running UNDER everything else of the step.
            if not (self.daemon_lead_ns > 0 and "fill" in P):
                with torch.cuda.stream(st_fill):
                    if "fill" in P:
                        self._daemon()
                    self.ev_fill.record(st_fill)
            with torch.cuda.stream(st_proj):
                if "proj" in P:
                    pr = ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b)
                self._bin(pr if pr is not None else self._last_proj(b), s)
                if "knn" in P:
                    self._knn(s, pr if pr is not None else self._last_proj(b), C)
                self.ev_proj.record(st_proj)
        elif sched == "fill_first":
            # fill (optionally throttled, C3D_FILL_PERSISTENT) together with the latency-bound
            # loss / EMA chains from t = 0; projection -> KNN afterwards
            with torch.cuda.stream(st_fill):
                if "fill" in P:
                    ops.zero_fill(self.grad)
                self.ev_fill.record(st_fill)
            with torch.cuda.stream(st_proj):
                st_proj.wait_event(self.ev_fill)
                if "proj" in P:
                    pr = ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b)
                self._bin(pr if pr is not None else self._last_proj(b), s)
                if "knn" in P:
                    self._knn(s, pr if pr is not None else self._last_proj(b), C)
                self.ev_proj.record(st_proj)
        else:
            with torch.cuda.stream(st_proj):
                if "proj" in P:
                    pr = ops.project_batch(s.points, s.offsets, self.fov, H, W, buffers=b)
                self._bin(pr if pr is not None else self._last_proj(b), s)
                self.ev_resolved.record(st_proj)
                if "knn" in P:
                    self._knn(s, pr if pr is not None else self._last_proj(b), C)  # ALU-bound
                self.ev_proj.record(st_proj)
            with torch.cuda.stream(st_fill):
                # The fill saturates HBM and streams 512 MB through L2, which slows the
                # z-buffer atomics and scatter/gather of the projection far more than the
                # overlap gains (profiles/timeline_r1.txt); it therefore starts after the
                # projection (default) or after the KNN vote ("fill_last").
                st_fill.wait_event(self.ev_proj if sched == "fill_last" else self.ev_resolved)
                if "fill" in P:
                    ops.zero_fill(self.grad)
                self.ev_fill.record(st_fill)
        fused_step = self.fused_step and {"loss", "ema"} <= P and not hold
        if fused_step:
            # one label split for both operators, then the sampler next to the EMA chain
            with torch.cuda.stream(st_loss):
                self._step(ops.STEP_SPLIT, s.labels, s, seed, cofill=self._fs[1])
                self.ev_split.record(st_loss)
                self._step(ops.STEP_SAMPLE, s.labels, s, seed)
            with torch.cuda.stream(st_ema):
                st_ema.wait_event(self.ev_split)
                self._step(ops.STEP_ACCUMULATE, s.labels, s, seed, cofill=self._fs[2])
                self._ema_finish()
                self.ev_ema.record(st_ema)
        else:
            with torch.cuda.stream(st_ema):
                if "ema" in P:
                    self._ema(s, seed)
                self.ev_ema.record(st_ema)
        with torch.cuda.stream(st_loss):
            # The selection phase (label split, anchor sampling) does not read the bank and runs
            # next to the EMA chain; the rows phase reads the bank the EMA has just updated
            # (salsanext_proto.py:520-527 runs inside model.forward, trainer.py:675-686 after it).
            if "loss" in P and not hold and not fused_step:
                self._loss_fwd(s, seed, phases=1)
            st_loss.wait_event(self.ev_ema)
            if fused_step:
                self._step(ops.STEP_LOSS_ROWS, s.labels, s, seed, cofill=self._fs[3])
            elif "loss" in P:
                self._loss_fwd(s, seed, phases=2)
            if defer_knn:
                self.ev_rows.record(st_loss)
                with torch.cuda.stream(st_proj):
                    st_proj.wait_event(self.ev_rows)
                    self._knn(s, pr if pr is not None else self._last_proj(b), C, cofill=self._fs[4], part="late")
                    self.ev_proj.record(st_proj)
                st_fill.wait_event(self.ev_proj)
                self.ev_fill.record(st_fill)
            st_loss.wait_event(self.ev_fill)
            if "loss" in P:
                ops.proto_loss_backward_raw(s.feats.shape, self.cfg, C, self.M, self.loss_ws,
                                            self.grad_out, self.grad, grad_is_zeroed=True)
            self.ev_loss.record(st_loss)
        for ev in (self.ev_proj, self.ev_ema, self.ev_loss):
            cur.wait_event(ev)
        return pr

    def run_inputs(self, points, offsets, weak, bufs, set_index=0, seed=0):
        """One step on externally supplied device inputs -- raw points (sum N, 4), CSR offsets and
        per-point weak labels (int32), e.g. just copied from the host: the projection is the
        fused f-1 call, whose weak-label image feeds the loss and the EMA (so those chains start
        after the projection), then the same chains and schedule as `run`.  Returns
        (loss 0-dim, per-point KNN labels int64, Assembled); both tensors are reused by the
        next call."""
        return self._run_inputs(points, offsets, weak, bufs, set_index, seed)

    def _run_inputs(self, points, offsets, weak, bufs, set_index, seed):
        s = self.sets[set_index % len(self.sets)]
        H, W, C = self.shape.proj_h, self.shape.proj_w, self.shape.n_classes
        cur = torch.cuda.current_stream(self.device)
        _, st_proj, st_ema, st_loss = self.side
        self.ev_fork.record(cur)
        for st in (st_proj, st_ema, st_loss):
            st.wait_event(self.ev_fork)
        with torch.cuda.stream(st_proj):
            asm = ops.project_assemble_batch(points, offsets, self.fov, H, W, weak_label=weak, buffers=bufs)
            self.ev_resolved.record(st_proj)
            ops.knn_batch(asm.proj_range, s.argmax, asm.uproj_depth, asm.uproj_x_idx, asm.uproj_y_idx,
                          offsets, self.knn_k, self.knn_s, self.knn_sigma, self.knn_cutoff, C,
                          inv_gauss=self.inv_gauss, out=self.knn_out, cofill=self.grad)
            self.ev_proj.record(st_proj)
        labels = asm.train_label
        with torch.cuda.stream(st_loss):
            st_loss.wait_event(self.ev_resolved)
            self._step(ops.STEP_SPLIT, labels, s, seed)
            self.ev_split.record(st_loss)
            self._step(ops.STEP_SAMPLE, labels, s, seed)
        with torch.cuda.stream(st_ema):
            st_ema.wait_event(self.ev_split)
            self._step(ops.STEP_ACCUMULATE, labels, s, seed)
            self._ema_finish()
            self.ev_ema.record(st_ema)
        with torch.cuda.stream(st_loss):
            st_loss.wait_event(self.ev_ema)           # the rows phase reads the updated bank
            self._step(ops.STEP_LOSS_ROWS, labels, s, seed)
            st_loss.wait_event(self.ev_proj)          # the vote has zero-filled self.grad
            ops.proto_loss_backward_raw(s.feats.shape, self.cfg, C, self.M, self.loss_ws, self.grad_out,
                                        self.grad, grad_is_zeroed=True)
            self.ev_loss.record(st_loss)
        for ev in (self.ev_proj, self.ev_ema, self.ev_loss):
            cur.wait_event(ev)
        return self.loss, self.knn_out, asm

    def _daemon(self):
        if self.daemon[0] == 2:     # (2, max_per_sm, launch_per_sm, page, chunk): claim form
            ops.zero_fill_daemon(self.grad, self.daemon_ctrl, *self.daemon[1:], debug=self.daemon_dbg)
        else:
            ops.zero_fill_background(self.grad, *self.daemon)

    def set_schedule(self, schedule, daemon=None, parts=None, fill_priority=None, fill_shares=None,
                     knn_after_rows=None, knn_split=None):
        """Switch the schedule of an existing step (drops captured graphs)."""
        self.schedule = schedule
        if knn_split is not None:
            self.knn_split = int(knn_split)
        if knn_after_rows is not None:
            self.knn_after_rows = bool(knn_after_rows)
        if fill_shares is not None:
            self.fill_shares = tuple(fill_shares)
        if daemon is not None:
            self.daemon = tuple(daemon)
        if parts is not None:
            self.parts = set(parts)
        self.graphs = None
        if fill_priority is None:
            fill_priority = -1 if schedule == "fill_daemon" else 0
        self.side[0] = torch.cuda.Stream(self.device, priority=fill_priority)

    def _last_proj(self, b):
        return ops.Projection(b.proj_pointcloud, b.proj_range, b.proj_idx, b.proj_mask,
                              b.uproj_x_idx, b.uproj_y_idx, b.uproj_depth, b.flags)

    def _step(self, phases, labels, s, seed, cofill=None):
        ops.proto_step_raw(phases, s.feats, s.probs, labels, None, self.protos, *self.ln_d, *self.ln_c,
                           self.cfg, self.loss_ws, self.packed, self.loss, self.max_rows,
                           assign_mode=ops.ASSIGN_GUMBEL_DEVICE, seed=seed, bank_n=self.bank_n,
                           seed_counters=self.seed_counters if self.device_seeds else None, cofill=cofill)

    def _ema_finish(self):
        """all-reduce of the packed sums (N > 1) + the EMA itself, in place on the bank"""
        sc = self.seed_counters if self.device_seeds else None
        if self.peer is not None:
            self.peer.apply(self.protos, self.packed, self.momentum, 0, out=self.protos,
                            normalised_out=self.bank_n, seed_counters=sc)
            return
        distributed.allreduce_packed(self.packed, self.group)
        ops.proto_ema_apply(self.protos, self.packed, self.momentum, 0, out=self.protos,
                            normalised_out=self.bank_n, seed_counters=sc)

    def _loss_fwd(self, s, seed, phases=3):
        labels = s.labels if s.labels_loss is None else s.labels_loss
        keep = s.keep_mask if s.keep_mask_loss is None else s.keep_mask_loss
        ops.proto_loss_forward_raw(s.feats, s.probs, labels, keep, self.protos, self.cfg,
                                   None, seed, self.loss_ws, self.loss, phases=phases)

    def _ema(self, s, seed):
        distributed.prototype_update(
            s.feats, s.labels, self.protos, *self.ln_d, *self.ln_c, self.momentum,
            assign_mode=ops.ASSIGN_GUMBEL_DEVICE, seed=seed, max_rows=self.max_rows, group=self.group,
            workspace=self.ema_ws, packed=self.packed, out=self.protos)

    def _bin(self, pr, s):
        """Binning pre-pass of the vote (right after the projection, on its stream)."""
        if self.knn_binned and "knn" in self.parts:
            ops.knn_sort_points(pr.uproj_depth, pr.uproj_x_idx, pr.uproj_y_idx, s.offsets,
                                self.shape.proj_h, self.shape.proj_w, workspace=self.knn_sort_ws,
                                out=self.knn_records)

    def _knn(self, s, pr, C, cofill=None, part=None):
        """The vote (carrying `cofill`).  With knn_split = b1 scans it is two launches over scans
        [0, b1) and [b1, B), each with its proportional share of the fill; `part` = "early" /
        "late" issues only one of them (the pipeline releases them at different points)."""
        rec = self.knn_records if self.knn_binned else None
        b1 = self.knn_split if rec is None else 0      # records are not sliced by scan: one launch
        if b1 <= 0 or b1 >= self.batch:
            if part != "early":
                ops.knn_batch(pr.proj_range, s.argmax, pr.uproj_depth, pr.uproj_x_idx, pr.uproj_y_idx,
                              s.offsets, self.knn_k, self.knn_s, self.knn_sigma, self.knn_cutoff, C,
                              inv_gauss=self.inv_gauss, out=self.knn_out, cofill=cofill, records=rec)
            return
        n1 = int(s.host_offsets[b1])
        if not hasattr(s, "offsets_tail") or s.offsets_tail_b1 != b1:
            s.offsets_tail = (s.offsets[b1:] - n1).contiguous()
            s.offsets_tail_b1 = b1
        f1 = f2 = None
        if cofill is not None:
            flat = cofill.view(-1)
            cut = flat.numel() * b1 // self.batch // 2048 * 2048      # 8 KB pages
            f1, f2 = (flat[:cut] if cut else None), flat[cut:]
        # binned records carry ORIGINAL (whole-batch) point indices: both launches write into knn_out
        if part != "late":
            ops.knn_batch(pr.proj_range[:b1], s.argmax[:b1], pr.uproj_depth[:n1], pr.uproj_x_idx[:n1],
                          pr.uproj_y_idx[:n1], s.offsets[:b1 + 1], self.knn_k, self.knn_s, self.knn_sigma,
                          self.knn_cutoff, C, inv_gauss=self.inv_gauss,
                          out=self.knn_out if rec is not None else self.knn_out[:n1], cofill=f1,
                          records=None if rec is None else rec[:n1])
        if part != "early":
            ops.knn_batch(pr.proj_range[b1:], s.argmax[b1:], pr.uproj_depth[n1:], pr.uproj_x_idx[n1:],
                          pr.uproj_y_idx[n1:], s.offsets_tail, self.knn_k, self.knn_s, self.knn_sigma,
                          self.knn_cutoff, C, inv_gauss=self.inv_gauss,
                          out=self.knn_out if rec is not None else self.knn_out[n1:], cofill=f2,
                          records=None if rec is None else rec[n1:])

    def capture(self):
        """Capture one CUDA graph per input set.  Returns False if capture fails
        (e.g. a collective that cannot be captured); the step then stays eager."""
        try:
            graphs = []
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for i in range(len(self.sets)):
                    self.run(i)  # warm-up on the capture stream
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            for i in range(len(self.sets)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.run(i)
                graphs.append(g)
            self.graphs = graphs
            return True
        except Exception as e:  # noqa: BLE001
            self.graphs = None
            self.capture_error = repr(e)
            torch.cuda.synchronize(self.device)
            return False

    def step(self, i):
        if self.graphs is not None:
            self.graphs[i % len(self.graphs)].replay()
        else:
            self.run(i, seed=i)
