"""The training step: ``batch_processor`` (mono/apis/trainer.py:30-56) plus the optimizer hook
(mono/core/utils/dist_utils.py:34-60) re-hosted on flat device buffers.

Reference per iteration: every input ``.float().cuda()`` -> model -> sum of *all* loss_dict entries ->
``.item()`` on each entry (~20 syncs) -> zero_grad -> backward (DDP bucket all-reduce) -> a second flat
all-reduce of all gradients -> clip_grad_norm_(35) over 466 tensors -> Adam over 466 tensors.

Here: parameters and gradients live in two flat fp32 buffers (views keep the per-tensor shapes and the
channels-last weight layout); one NCCL all-reduce of the gradient buffer, one ``jpb_sumsq`` and one
``jpb_adam_step`` launch (1/world scaling + clipping + Adam fused); all loss scalars travel to the host
in one copy.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
import torch.distributed as dist

from .. import _lib
from .env import get_dist_info

HOST_KEYS = ("odometry_K", "Tr_cam2_velo")


def change_input_variable(data, device=None, non_blocking=True):
    """Move a batch dict to the GPU as float32 (trainer.py:20-27).  Small calibration tensors keep a host
    copy under ``("_host", name)`` so host-side caches (the static scale-label quad) need no D2H."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    out = {}
    for k, v in data.items():
        if isinstance(k, tuple) and k and k[0] == "bev_path":
            out[k] = v
            continue
        if isinstance(k, tuple) and "kp" in k:
            out[k] = v
            continue
        t = torch.as_tensor(v)
        if isinstance(k, tuple) and k and k[0] in HOST_KEYS and not t.is_cuda:
            out[("_host", k[0])] = t.float()
        out[k] = t.to(device=device, dtype=torch.float32, non_blocking=non_blocking)
    return out


def loss_scalars(losses):
    """Ordered names + one stacked device tensor [each entry..., total]; total = sum of ALL entries
    (the layout terms are double-counted, exactly as trainer.py:44 does)."""
    names, vals = [], []
    for name, v in losses.items():
        if isinstance(v, torch.Tensor):
            vals.append(v.mean() if v.dim() else v)
        elif isinstance(v, list):
            vals.append(sum(x.mean() for x in v))
        else:
            raise TypeError("{} is not a tensor or list of tensors".format(name))
        names.append(str(name))
    total = vals[0]
    for v in vals[1:]:
        total = total + v
    return names, vals, total


def batch_processor(model, data, train_mode):
    """Reference contract: returns dict(loss=tensor, log_vars=OrderedDict of floats, num_samples=int)."""
    data = change_input_variable(data)
    model_out, losses = model(data)
    names, vals, total = loss_scalars(losses)
    host = torch.stack([v.detach() for v in vals] + [total.detach()]).tolist()   # ONE device->host copy
    log_vars = OrderedDict(zip(names + ["loss"], host))
    return dict(loss=total, log_vars=log_vars, num_samples=len(data[("color", 0, 0)]))


class FlatParameters:
    """Re-home every parameter (and its gradient) of ``model`` as a view into one flat fp32 buffer."""

    def __init__(self, model):
        params = list(model.parameters())
        frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
        if frozen:
            # checkpoint optimizer state is indexed over model.parameters() (torch.optim.Adam layout, as the reference's); a
            # frozen parameter would shift every index after it
            raise NotImplementedError("frozen parameters are not supported by the flat optimizer (%s ...)" % frozen[0])
        dev = params[0].device
        # every tensor starts on a 256-byte boundary: TMA / cp.async / vectorised loads need >= 16-byte alignment
        n = sum(self._padded(p.numel()) for p in params)
        self.numel = n
        self.param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in params:
            pv, gv = self._view(self.param, off, p), self._view(self.grad, off, p)
            pv.copy_(p.data)
            p.data = pv
            p.grad = gv
            self.views.append((off, p.numel()))
            off += self._padded(p.numel())
        self.params = params

    ALIGN = 64  # floats

    @classmethod
    def _padded(cls, n):
        return (n + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN

    @staticmethod
    def _view(flat, off, p):
        sl = flat[off:off + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            o, i, h, w = p.shape
            return sl.view(o, h, w, i).permute(0, 3, 1, 2)
        return sl.view(p.shape)

    def permute(self, order, extras=()):
        """Re-lay the flat buffers so that parameters appear in ``order`` (indices into ``self.params``); ``extras`` are other
        flat tensors of the same layout (optimizer moments) to permute alike — the new tensors are returned in the same order.
        Per-parameter views, ``views`` offsets and values are preserved; every cached pointer into the old buffers is stale."""
        assert sorted(order) == list(range(len(self.params)))
        new_off, off = [0] * len(self.params), 0
        for i in order:
            new_off[i] = off
            off += self._padded(self.params[i].numel())
        olds = [self.param, self.grad] + list(extras)
        news = [torch.zeros_like(t) for t in olds]
        for i, p in enumerate(self.params):
            o, n = self.views[i]
            for a, b in zip(olds, news):
                b[new_off[i]:new_off[i] + n].copy_(a[o:o + n])
        for i, p in enumerate(self.params):
            pv, gv = self._view(news[0], new_off[i], p), self._view(news[1], new_off[i], p)
            p.data = pv
            p.grad = gv
            self.views[i] = (new_off[i], p.numel())
        self.param, self.grad = news[0], news[1]
        return news[2:]

    def zero_grad(self):
        self.grad.zero_()
        for p, (off, n) in zip(self.params, self.views):   # autograd may have been told to drop .grad
            if p.grad is None:
                p.grad = self._view(self.grad, off, p)


class FusedAdam:
    """Adam on the flat buffers via ``jpb_sumsq`` + ``jpb_adam_step`` (clip + 1/world + update fused)."""

    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=None):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.initial_lr = lr          # mmcv LrUpdaterHook's 'initial_lr': base of the schedule, saved with every checkpoint
        self.max_norm = max_norm
        dev = flat.param.device
        self.exp_avg = torch.zeros_like(flat.param)
        self.exp_avg_sq = torch.zeros_like(flat.param)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.normsq = torch.zeros(1, dtype=torch.float64, device=dev)

    def step(self, world_size=1):
        f = self.flat
        lib = _lib.lib()
        st = _lib.stream_of(f.param)
        a = _lib.AdamArgs()
        a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay
        a.grad_scale = 1.0 / world_size
        a.max_norm = float(self.max_norm) if self.max_norm else 0.0
        a.step = _lib.ptr(self.step_count)
        if self.max_norm:
            self.normsq.zero_()
            _lib.check(lib.jpb_sumsq(_lib.ptr(f.grad), f.numel, _lib.ptr(self.normsq), st), "jpb_sumsq")
            a.normsq = _lib.ptr(self.normsq)
        _lib.check(lib.jpb_adam_step(_lib.ptr(f.param), _lib.ptr(f.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                     f.numel, C.byref(a), st), "jpb_adam_step")

    def grad_norm(self, world_size=1):
        """Host float of the (averaged) gradient norm of the last step — a sync; for logging/tests only."""
        return float(self.normsq.sqrt().item()) / world_size


def build_optimizer(model, optimizer_cfg, grad_clip=None):
    """``build_optimizer`` of trainer.py:76-143 for the configs' ``dict(type='Adam', lr=..., weight_decay=0)``."""
    cfg = dict(optimizer_cfg)
    if cfg.pop("type", "Adam") != "Adam":
        raise NotImplementedError("the B200 path implements the optimizer the reference configs use (Adam)")
    if cfg.pop("paramwise_options", None) is not None:
        raise NotImplementedError("paramwise_options are not used by the reference configs")
    flat = getattr(model, "_jpb_flat", None)
    if flat is None:
        flat = model._jpb_flat = FlatParameters(model)
    max_norm = None
    if grad_clip:
        if grad_clip.get("norm_type", 2) != 2:
            raise NotImplementedError("only the L2 gradient-norm clip of the reference configs is implemented")
        max_norm = grad_clip.get("max_norm")
    return FusedAdam(flat, lr=cfg.get("lr", 1e-3), betas=tuple(cfg.get("betas", (0.9, 0.999))), eps=cfg.get("eps", 1e-8),
                     weight_decay=cfg.get("weight_decay", 0.0), max_norm=max_norm)


class GradExchange:
    """The path's only collective, overlapped with backward: the flat gradient buffer is all-reduced in buckets, each as soon as
    every gradient in it has been written (the reference gets the same overlap from DDP's buckets, trainer.py:167).

    Step 1 of a multi-rank engine TRACES the order in which backward finishes the parameters' gradients (hook calls from the
    backward kernels' host code and from autograd's accumulate nodes), exchanges the gradient in one piece, then re-lays the flat
    buffers in that completion order and cuts them into buckets.  From step 2 on, a bucket's all-reduce is issued (async, on the
    process group's stream, after the streams that wrote into it) a few hook events after its last gradient was produced; the
    end of backward issues the rest and makes the caller's stream wait for all of them.  Every rank issues the same buckets in the
    same order.  Works on NCCL (inside or outside a CUDA graph capture) and on gloo (CPU tests)."""

    LAG = 6   # hook events between "the last gradient of a bucket is about to be written" and the bucket's launch: the writing
              # launch follows its hook call inside the same autograd function (at most 4 parameters per function), so six events
              # later it has certainly been issued

    def __init__(self, engine, bucket_bytes=24 << 20, tail_bytes=4 << 20):
        self.engine = engine
        self.bucket_bytes, self.tail_bytes = bucket_bytes, tail_bytes
        self.mode = "trace"
        self.counter = 0
        self.last = {}          # id(param) -> last event index (trace)
        self.buckets = []       # (lo, hi, ready_at)
        self.bucket_of = {}     # id(param) -> bucket index
        self.step_streams = []  # per bucket: streams that wrote into it during the current step
        self.main = None
        self.next = 0
        self.works = []
        self.index = {id(p): i for i, p in enumerate(engine.flat.params)}
        for p in engine.flat.params:   # gradients that arrive through autograd's accumulate nodes (not written by our kernels directly)
            p.register_post_accumulate_grad_hook(self._accumulated)

    def _accumulated(self, p):
        from .. import functional as JF
        if JF.GRAD_EVENT is not None:
            JF.GRAD_EVENT(p)

    # ---- hook target
    def event(self, p):
        self.counter += 1
        if self.mode == "trace":
            self.last[id(p)] = self.counter
            return
        if p.is_cuda:   # the streams that write into a bucket are those of THIS step (eager, or the streams of a graph capture)
            self.step_streams[self.bucket_of[id(p)]].add(torch.cuda.current_stream(p.device))
        while self.next < len(self.buckets) and self.buckets[self.next][2] + self.LAG <= self.counter:
            self._launch(self.next)
            self.next += 1

    def begin(self):
        self.counter, self.next, self.works = 0, 0, []
        self.step_streams = [set() for _ in self.buckets]
        g = self.engine.flat.grad
        self.main = torch.cuda.current_stream(g.device) if g.is_cuda else None   # zero-fill of the gradient buffer runs here

    def _launch(self, i):
        lo, hi, _ = self.buckets[i]
        grad = self.engine.flat.grad
        if grad.is_cuda:
            comm = self._comm_stream(grad.device)
            streams = set(self.step_streams[i])
            streams.add(self.main)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                self.works.append(dist.all_reduce(grad[lo:hi], async_op=True))
        else:
            self.works.append(dist.all_reduce(grad[lo:hi], async_op=True))

    def _comm_stream(self, device):
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device)
        return self._comm

    def finish(self):
        """End of backward: issue the buckets that are still pending, then make the caller's stream wait for all of them."""
        while self.next < len(self.buckets):
            self._launch(self.next)
            self.next += 1
        for w in self.works:
            w.wait()
        self.works = []

    # ---- after the traced step
    def build(self):
        eng = self.engine
        flat = eng.flat
        last = [self.last.get(id(p), 0) for p in flat.params]            # 0 = never written (unused parameters): first bucket
        order = sorted(range(len(flat.params)), key=lambda i: (last[i], i))
        opt = eng.optimizer
        opt.exp_avg, opt.exp_avg_sq = flat.permute(order, extras=(opt.exp_avg, opt.exp_avg_sq))
        buckets, members, lo, ready = [], [], None, 0
        total_bytes = flat.numel * 4
        for i in order:
            off, n = flat.views[i]
            if lo is None:
                lo = off
            ready = max(ready, last[i])
            members.append(i)
            hi = off + flat._padded(n)
            size = (hi - lo) * 4
            limit = self.tail_bytes if (total_bytes - hi * 4) < 2 * self.tail_bytes else self.bucket_bytes   # small buckets at the tail
            if size >= limit:
                for m in members:
                    self.bucket_of[id(flat.params[m])] = len(buckets)
                buckets.append((lo, hi, ready))
                lo, members = None, []
        if lo is not None:
            for m in members:
                self.bucket_of[id(flat.params[m])] = len(buckets)
            buckets.append((lo, flat.numel, ready))
        self.buckets = buckets
        self.mode = "run"
        from .. import conv as JC
        JC.WT.entries.clear()          # keyed by the weights' (old) addresses
        JC.WT.table = None
        eng._graph = None


class TrainEngine:
    """One object per process (= per GPU): forward, backward, gradient exchange, optimizer step."""

    DEFAULT_CLIP = dict(max_norm=35, norm_type=2)     # optimizer_config of every reference config

    def __init__(self, model, optimizer_cfg=None, grad_clip="default"):
        """``grad_clip``: ``dict(max_norm=..., norm_type=2)``; ``None`` / ``{}`` = no clipping (DistOptimizerHook with
        grad_clip=None, dist_utils.py:56-58); omitted = the reference configs' ``max_norm=35``."""
        self.model = model
        if isinstance(grad_clip, str):
            grad_clip = dict(self.DEFAULT_CLIP)
        self.optimizer = build_optimizer(model, optimizer_cfg or dict(type="Adam", lr=1e-4, weight_decay=0), grad_clip)
        self.flat = model._jpb_flat
        self.rank, self.world = get_dist_info()
        self.exchange = None
        if self.world > 1:
            # what MMDistributedDataParallel does at construction: every rank starts from rank 0's parameters and buffers
            dist.broadcast(self.flat.param, src=0)
            self.sync_buffers()
            import os as _os
            # Overlapped, bucketed exchange: implemented and correct (tests/test_trainer.py on gloo, tests/test_gpu_training.py on
            # NCCL), but MEASURED SLOWER than the one-piece exchange on 2 and 8 B200s (profiles/r2_scaling_n8.md: 27.29 vs 26.89 ms
            # per step at N = 8) — the backward kernels already fill the GPU and NCCL's kernels take SMs from them — so it is opt-in.
            if _os.environ.get("JPB_OVERLAP_ALLREDUCE", "0") not in ("", "0"):
                self.exchange = GradExchange(self)
        self.last_names = None
        if hasattr(model, "step_counter"):
            model.step_counter = self.optimizer.step_count   # fresh automask noise per step, also under graph replay
        self._graph = None

    def sync_buffers(self):
        """Broadcast every module buffer (BatchNorm running statistics, ``num_batches_tracked``) from rank 0.  Statistics stay
        per-GPU DURING training (no SyncBN, as the reference); rank 0's are the ones a checkpoint holds, so they are what
        distributed validation must score on every rank."""
        if self.world > 1:
            for b in self.model.buffers():
                dist.broadcast(b, src=0)

    def exchange_gradients(self):
        """The path's only collective: all-reduce(sum) of the flat gradient buffer (dist_utils.py:27);
        the division by world size is fused into the optimizer kernel."""
        if self.world > 1:
            dist.all_reduce(self.flat.grad)

    def forward_backward(self, data):
        """Forward + ``compute_losses`` + backward into the flat gradient buffer.  Returns (names, values, total)."""
        from .. import functional as JF, conv as JC
        self.flat.zero_grad()
        JC.WT.enabled = True
        # one launch: flipped/transposed weights of every convolution for this step's data gradients.  Only the backward reads
        # them, so on the GPU the launch runs on a side stream beside the forward (it used to sit alone in front of the step)
        wt_event = None
        if self.flat.grad.is_cuda:
            cur = torch.cuda.current_stream(self.flat.grad.device)
            if getattr(self, "_wt_stream", None) is None:
                self._wt_stream = torch.cuda.Stream(self.flat.grad.device)
            self._wt_stream.wait_stream(cur)       # the weights are final: the previous optimizer step is ordered before this point
            with torch.cuda.stream(self._wt_stream):
                JC.WT.refresh()
                wt_event = torch.cuda.Event()
                wt_event.record(self._wt_stream)
        else:
            JC.WT.refresh()
        try:
            _, losses = self.model(data)
            names, vals, total = loss_scalars(losses)
            JF.DIRECT_GRAD = True  # backward kernels add parameter gradients straight into the flat gradient buffer
            if wt_event is not None:
                torch.cuda.current_stream(self.flat.grad.device).wait_event(wt_event)   # every backward node is ordered after this
            total.backward()
            # branch-concurrent model: its side streams ran backward nodes that wrote parameter gradients straight into the flat
            # buffer (no AccumulateGrad node, so autograd's end-of-backward stream sync does not cover them)
            if hasattr(self.model, "side_streams") and self.flat.grad.is_cuda:
                cur = torch.cuda.current_stream(self.flat.grad.device)
                for st in self.model.side_streams():
                    cur.wait_stream(st)
        finally:
            JF.DIRECT_GRAD = False
            JC.WT.enabled = JC.WT.fresh = False
        return names, vals, total

    def step(self, data, need_log=True):
        """data: dict of device tensors (see ``change_input_variable``).  Returns the stacked loss tensor
        ``[entries..., total]`` on the device (names in ``self.last_names``)."""
        from .. import functional as JF
        ex = self.exchange
        if ex is None:
            names, vals, total = self.forward_backward(data)
            self.exchange_gradients()
        else:
            ex.begin()
            JF.GRAD_EVENT = ex.event
            try:
                names, vals, total = self.forward_backward(data)
            finally:
                JF.GRAD_EVENT = None
            if ex.mode == "trace":       # first step: one exchange of the whole buffer, then lay the buffers out in completion order
                self.exchange_gradients()
                ex.build()
            else:
                ex.finish()
        self.optimizer.step(self.world)
        self.last_names = names + ["loss"]
        if need_log:
            return torch.stack([v.detach() for v in vals] + [total.detach()])
        return total.detach()


    # ------------------------------------------------------------------ CUDA-graph execution of the whole step
    def capture(self, data, warmup=3):
        """Capture forward + losses + backward + gradient exchange + optimizer into one CUDA graph.  ``data`` provides the
        shapes; its tensors become the static input buffers (``replay`` copies new batches into them).  Shapes are static
        in this workload, so one graph serves the whole run; every kernel is launched on the capture stream by the same
        code path as the eager step."""
        from .. import functional as JF
        self._static_in = {k: (v.clone() if torch.is_tensor(v) and v.is_cuda else v) for k, v in data.items()}
        # everything that touches autograd before / during capture runs on ONE non-default stream: a leaf whose gradient
        # accumulator was bound to the legacy default stream would make the capture depend on it (cudaErrorStreamCaptureImplicit)
        side = self._capture_stream = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # also creates chunk tables, sets kernel attributes, fills host-side caches
                self.step(self._static_in, need_log=True)
        side.synchronize()
        prof, JF.PROFILE_ON = JF.PROFILE_ON, False
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):
            self._static_out = self.step(self._static_in, need_log=True)
        JF.PROFILE_ON = prof
        torch.cuda.current_stream().wait_stream(side)
        return self

    def release_graph(self):
        """Drop the captured graph and its output buffers (before tearing down a process group whose collectives it holds, or
        to return the graph's private memory pool)."""
        self._graph = None
        self._static_out = None
        self._stage = None

    def _refresh_host_caches(self, data):
        """Host-side caches whose device buffers were baked into the captured graph: the static scale label's quad mask is
        rasterised on the host from sample 0's calibration (net.py:292-306).  When a replayed batch carries a different
        calibration (KITTI-Odometry sequences differ in K / Tr) the mask is rasterised again INTO THE SAME device buffer, so the
        graph reads the right one; batches without the host copies (``("_host", name)``) are taken to share the captured
        calibration."""
        model = self.model
        cache = getattr(model, "_quad_cache", None)
        hK, hT = data.get(("_host", "odometry_K")), data.get(("_host", "Tr_cam2_velo"))
        if not cache or hK is None or hT is None:
            return
        (key, buf), = cache.items()
        new = (hK[0].numpy().tobytes(), hT[0].numpy().tobytes(), key[2], key[3])
        if new != key:
            from ..model.mono_baseline.net import static_quad_mask_host
            m = static_quad_mask_host(hK[0].numpy(), hT[0].numpy(), model.opt.split, model.opt.occ_map_size, key[2], key[3])
            buf.copy_(torch.from_numpy(m).to(buf.dtype), non_blocking=False)
            model._quad_cache = {new: buf}

    def replay(self, data=None):
        """Run the captured step; ``data`` (device or pinned-host tensors) is copied into the static inputs first."""
        if data is not None:
            self._refresh_host_caches(data)
            for k, v in data.items():
                dst = self._static_in.get(k)
                if torch.is_tensor(dst) and dst.is_cuda and torch.is_tensor(v) and v is not dst:
                    dst.copy_(v, non_blocking=True)
        self._graph.replay()
        return self._static_out


    def replay_pipelined(self, next_batch):
        """Input-pipelined replay: run the captured step on the batch staged by the PREVIOUS call while ``next_batch`` (pinned
        host tensors) is copied host->device into a staging copy on a side stream — what a prefetching data loader does.
        Per call: one H2D of a whole batch (overlapped with the step), one on-device move of the staged batch into the graph's
        static inputs (~40 us for 120 MB), one graph replay.  The first call stages ``next_batch`` synchronously."""
        if getattr(self, "_stage", None) is None:
            self._stage = {k: torch.empty_like(v) for k, v in self._static_in.items() if torch.is_tensor(v) and v.is_cuda}
            self._copy_stream = torch.cuda.Stream()
            self._staged_event = None
        cur = torch.cuda.current_stream()
        if self._staged_event is None:
            self._stage_batch(next_batch, after=None)
            self._staged_batch = next_batch
        cur.wait_event(self._staged_event)
        self._refresh_host_caches(self._staged_batch)        # the batch that runs now is the one staged by the previous call
        for k, st in self._stage.items():
            self._static_in[k].copy_(st, non_blocking=True)
        moved = torch.cuda.Event()
        moved.record(cur)
        self._stage_batch(next_batch, after=moved)
        self._staged_batch = next_batch
        self._graph.replay()
        return self._static_out

    def _stage_batch(self, batch, after):
        with torch.cuda.stream(self._copy_stream):
            if after is not None:
                self._copy_stream.wait_event(after)          # the previous staged batch has been moved out
            for k, st in self._stage.items():
                v = batch.get(k)
                if torch.is_tensor(v):
                    st.copy_(v, non_blocking=True)
            self._staged_event = torch.cuda.Event()
            self._staged_event.record(self._copy_stream)


def train_mono(model, dataset_train, dataset_val, cfg, args=None, distributed=False, validate=False, logger=None):
    """``mono.apis.train_mono`` (trainer.py:58-73, 146-199): build the runner, register the config's hooks, resume / load,
    run ``cfg.total_epochs``.  ``dataset_train`` is a map-style dataset with a ``flag`` array (what ``get_dataset`` returns): it
    is wrapped by ``build_dataloader(dataset, cfg.imgs_per_gpu, cfg.workers_per_gpu, dist=distributed)`` exactly as
    ``_dist_train`` / ``_non_dist_train`` do (trainer.py:148-151, 202-208) — epoch-seeded, group-pure, one contiguous slice of
    the plan per rank; an iterable of already collated batch dicts is used as it is.  ``validate=True`` registers the device-side ``DistEvalMonoHook`` over ``dataset_val`` every
    ``cfg.validate_interval`` epochs (trainer.py:186-190)."""
    from .runner import Runner
    if torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    elif _lib.is_emulated():                  # tests only: the kernels' host emulation was installed explicitly
        dev = torch.device("cpu")
    else:
        raise _lib.JpbError("train_mono needs a CUDA device (there is no CPU fallback)")
    model.to(dev).train()
    runner = Runner(model, cfg.optimizer, cfg.get("optimizer_config", {}), cfg.get("work_dir"), cfg.get("log_level", "INFO"), logger)
    if validate:
        from ..core.evaluation import DistEvalMonoHook
        runner.register_hook(DistEvalMonoHook(dataset_val, cfg.get("validate_interval", 1), cfg))
    runner.register_training_hooks(cfg.get("lr_config"), cfg.get("optimizer_config"), cfg.get("checkpoint_config"), cfg.get("log_config"))
    if cfg.get("resume_from"):
        runner.resume(cfg.resume_from)
    elif cfg.get("load_from"):
        runner.load_checkpoint(cfg.load_from)
    loader = dataset_train
    if hasattr(dataset_train, "__getitem__") and hasattr(dataset_train, "flag"):
        from ..datasets.loader import build_dataloader
        gpus = cfg.get("gpus", [0])
        loader = build_dataloader(dataset_train, cfg.imgs_per_gpu, cfg.get("workers_per_gpu", 0),
                                  num_gpus=1 if distributed else max(len(gpus) if hasattr(gpus, "__len__") else int(gpus), 1),
                                  dist=distributed)
    runner.run([loader], cfg.get("workflow", [("train", 1)]), cfg.get("total_epochs", 1))
    return runner.engine
