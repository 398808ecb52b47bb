"""Host side of the training step: the step maths of the reference ``Solver.train`` loop
(solver_encoder.py:227-243 losses, :293-300 zero_grad/backward/Adam) and the data-parallel
gradient exchange the B200 build adds.

The reference ``Solver`` class itself runs unchanged on ``autovc_b200.Generator`` (its loop only
calls ``self.G(...)``, ``F.mse_loss``/``F.l1_loss``, ``backward`` and the optimizer).  This module
offers the same step as a function — with the three losses going through the fused CUDA loss
kernels instead of ATen — plus ``GradBucketReducer`` for one-process-per-GPU data parallelism:
gradients are all-reduced in ~25 MB buckets in reverse registration order (postnet -> decoder ->
encoder) as soon as each bucket's gradients are final, on a side stream, overlapped with the rest
of backward.  BatchNorm statistics stay rank-local (DDP semantics, ``broadcast_buffers=False``).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops


def generator_losses(G, x_real, emb_org, lambda_cd: float = 1.0):
    """solver_encoder.py:228-243 with the fused loss kernels."""
    x_identic, x_identic_psnt, code_real = G(x_real, emb_org, emb_org)          # :228
    g_loss_id = ops.mse_loss(x_real, x_identic.squeeze(1))                      # :230
    g_loss_id_psnt = ops.mse_loss(x_real, x_identic_psnt.squeeze(1))            # :233
    code_reconst = G(x_identic_psnt, emb_org, None)                             # :235
    g_loss_cd = ops.l1_loss(code_real, code_reconst)                            # :236
    g_loss = g_loss_id + g_loss_id_psnt + lambda_cd * g_loss_cd                 # :243
    return g_loss, (g_loss_id, g_loss_id_psnt, g_loss_cd), (x_identic, x_identic_psnt, code_real, code_reconst)


def generator_losses_wav(G, x_real, emb_org, lambda_cd: float = 1.0, lambda_SISNR: float = 1.0):
    """The 'wav' branch of the step, solver_encoder.py:264-290 (GeneratorWav): reconstruction MSE on the waveform, MSE
    between the analysis filterbank's output and the decoder's estimate of it, L1 content-code loss on a second encoder
    pass over the synthesised waveform, and the SI-SNR term -- each through a fused CUDA loss kernel."""
    from . import ops_wav
    x_convtas, x_identic, gen_outputs, code_real = G(x_real, emb_org, emb_org)  # :265
    g_loss_id = ops.mse_loss(x_real, x_identic)                                 # :268
    # both are transposed views of channels-last storage: compare them in that layout (the mean is layout-independent)
    g_loss_gen = ops.mse_loss(x_convtas.transpose(1, 2), gen_outputs.transpose(1, 2))   # :271
    code_reconst = G(x_identic, emb_org, None)                                  # :274
    g_loss_cd = ops.l1_loss(code_real, code_reconst)                            # :275
    g_loss_SISNR = ops_wav.sisnr_loss(x_identic, x_real)                        # :281-288
    g_loss = g_loss_id + lambda_SISNR * g_loss_SISNR + g_loss_gen + lambda_cd * g_loss_cd   # :291
    return g_loss, (g_loss_id, g_loss_gen, g_loss_cd, g_loss_SISNR), (x_convtas, x_identic, gen_outputs, code_real, code_reconst)


def train_step_wav(G, optimizer, x_real, emb_org, lambda_cd: float = 1.0, lambda_SISNR: float = 1.0,
                   reducer: "Optional[GradBucketReducer]" = None, return_outputs: bool = False):
    """One 'wav' training iteration (solver_encoder.py:264-300).  Returns g_loss and the four terms as floats."""
    g_loss, terms, outs = generator_losses_wav(G, x_real, emb_org, lambda_cd, lambda_SISNR)
    optimizer.zero_grad()
    if reducer is not None:
        reducer.begin_backward()
    g_loss.backward()
    if reducer is not None:
        reducer.finish()
    result: Dict[str, object] = {}
    if return_outputs:
        result["grads"] = {n: p.grad.detach().clone() for n, p in G.named_parameters()}
        for k, v in zip(("x_convtas", "x_identic", "gen_outputs", "code_real", "code_reconst"), outs):
            result[k] = v.detach()
    optimizer.step()
    vals = torch.stack([g_loss.detach()] + [t.detach() for t in terms]).tolist()
    result.update({"g_loss": vals[0], "L_id": vals[1], "L_gen": vals[2], "L_cd": vals[3], "L_SISNR": vals[4]})
    return result


def train_step(G, optimizer, x_real, emb_org, lambda_cd: float = 1.0, reducer: "Optional[GradBucketReducer]" = None,
               return_outputs: bool = False, sync_losses: bool = True):
    """One training iteration (solver_encoder.py:228-300).  Returns the three loss terms the
    reference logs with ``.item()`` (:315-317) — as floats when ``sync_losses`` (one device sync,
    like the reference), else as device tensors."""
    g_loss, (l_id, l_id_psnt, l_cd), outs = generator_losses(G, x_real, emb_org, lambda_cd)
    optimizer.zero_grad()                                                       # :293
    if reducer is not None:
        reducer.begin_backward()
    g_loss.backward()                                                           # :294
    if reducer is not None:
        reducer.finish()
    result: Dict[str, object] = {}
    if return_outputs:
        result["grads"] = {n: p.grad.detach().clone() for n, p in G.named_parameters()}
        for k, v in zip(("x_identic", "x_identic_psnt", "code_real", "code_reconst"), outs):
            result[k] = v.detach()
    optimizer.step()                                                            # :300
    if sync_losses == "async":
        # the four scalars go to pinned host memory without blocking the host; the caller reads them through the handle
        # (typically one step later, when it logs) -- the GPU queue then never runs dry between steps
        result["losses"] = AsyncLosses(torch.stack([g_loss.detach(), l_id.detach(), l_id_psnt.detach(), l_cd.detach()]))
        return result
    if sync_losses:
        vals = torch.stack([g_loss.detach(), l_id.detach(), l_id_psnt.detach(), l_cd.detach()]).tolist()
    else:
        vals = [g_loss.detach(), l_id.detach(), l_id_psnt.detach(), l_cd.detach()]
    result.update({"g_loss": vals[0], "L_id": vals[1], "L_id_psnt": vals[2], "L_cd": vals[3]})
    return result


class AsyncLosses:
    """Device -> pinned-host copy of a step's loss scalars (solver_encoder.py:315-317 reads them with ``.item()``) that does
    not stall the host: ``values()`` waits for the copy and returns [g_loss, L_id, L_id_psnt, L_cd] as floats."""

    _pool: List[torch.Tensor] = []

    def __init__(self, dev_vals: torch.Tensor):
        self.host = AsyncLosses._pool.pop() if AsyncLosses._pool else torch.empty(4, dtype=torch.float32).pin_memory()
        self.host.copy_(dev_vals, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def values(self) -> List[float]:
        self.event.synchronize()
        out = self.host.tolist()
        AsyncLosses._pool.append(self.host)
        return out


class HostBatchPrefetcher:
    """Double-buffered host -> device staging of the (x_real, emb_org) batches a loader hands over in pinned memory
    (the reference copies them synchronously with ``.to(self.device)``, solver_encoder.py:203-204).

    ``put(x_pin, e_pin)`` enqueues the copy of the NEXT batch on a side stream; ``get()`` makes the compute stream wait
    for the oldest enqueued batch and returns its device tensors.  With one batch in flight the copy of step n+1 runs
    under the kernels of step n instead of in front of them."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]      # device buffers, reused
        self._queue = []                # (slot, event)
        self._next = 0

    def put(self, x_pin: torch.Tensor, e_pin: torch.Tensor):
        slot = self._next
        self._next ^= 1
        bufs = self._slots[slot]
        if bufs is None or bufs[0].shape != x_pin.shape or bufs[1].shape != e_pin.shape:
            bufs = (torch.empty(x_pin.shape, dtype=x_pin.dtype, device=self.device),
                    torch.empty(e_pin.shape, dtype=e_pin.dtype, device=self.device))
            self._slots[slot] = bufs
        # the slot's previous contents may still be read by kernels of the step that used it
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            bufs[0].copy_(x_pin, non_blocking=True)
            bufs[1].copy_(e_pin, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._queue.append((slot, ev))

    def get(self):
        slot, ev = self._queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return self._slots[slot]


def _capped_nccl_group(max_ctas: int):
    """A process group over all ranks whose NCCL communicator is limited to ``max_ctas`` CTAs per collective, so that a gradient
    all-reduce can never take the SMs a cooperative recurrence launch is waiting for (see ``nccl_env_defaults``: 33.7 ms per
    step with NCCL's default CTA count on 4 x B200, 14.1-14.4 ms with 4-16 CTAs).  The reducer owns this group, so the cap
    does not depend on the caller having set NCCL_MAX_CTAS before ``init_process_group``.  Collective: every rank constructs its
    reducer.  Returns (group, how the cap was applied)."""
    import os
    import warnings
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        return dist.new_group(backend="nccl", pg_options=opts), f"ncclConfig.maxCTAs={int(max_ctas)} on the reducer's own communicator"
    except Exception as exc:          # a torch build without ncclConfig: fall back to the environment, and say so
        if os.environ.get("NCCL_MAX_CTAS") is None:
            warnings.warn("autovc_b200.GradBucketReducer: could not cap NCCL's CTA count on its own communicator "
                          f"({type(exc).__name__}: {exc}) and NCCL_MAX_CTAS is not set -- all-reduces with NCCL's default CTA count "
                          "convoy with the cooperative recurrence launches (2.4x slower steps measured on 4 GPUs); call "
                          "solver.nccl_env_defaults() before init_process_group", RuntimeWarning)
            return None, "UNCAPPED (NCCL default CTA count)"
        return None, f"NCCL_MAX_CTAS={os.environ['NCCL_MAX_CTAS']} from the environment"


class GradBucketReducer:
    """Bucketed, backward-overlapped gradient all-reduce (mean) over a process group.

    * parameters are bucketed in REVERSE registration order so buckets fill in the order backward
      produces gradients; encoder gradients (sum of the two passes of a step) finalise last;
    * each parameter's post-accumulate-grad hook copies its gradient into the bucket's flat
      buffer and re-points ``param.grad`` at that slice; when a bucket is complete its all-reduce
      is issued asynchronously on ``comm_stream`` (CUDA) so it overlaps the remaining backward;
    * ``finish()`` waits for all buckets (and divides by world size when the backend has no AVG).
    Works with NCCL on GPUs and with gloo on CPU tensors (used by the world_size-2 CPU tests).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, bucket_mb: float = 25.0,
                 nccl_max_ctas: int = 16):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        self.cta_cap = None
        if process_group is None and self.backend == "nccl" and self.world > 1:
            process_group, self.cta_cap = _capped_nccl_group(nccl_max_ctas)
        self.group = process_group
        cap = int(bucket_mb * 1024 * 1024)
        self.buckets: List[dict] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and cur_bytes + nbytes > cap:
                self._close_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close_bucket(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b["params"]:
                self._where[p] = (bi, off)
                off += p.numel()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        dev = self.params[0].device
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._pending: List = []
        self.launch_order: List[int] = []

    def _close_bucket(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=plist[0].dtype, device=plist[0].device)
        views, off = [], 0
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.buckets.append({"params": list(plist), "flat": flat, "views": views, "ready": 0})

    def begin_backward(self):
        for b in self.buckets:
            b["ready"] = 0
        self._pending.clear()
        self.launch_order.clear()

    def _on_grad(self, p):
        bi, _ = self._where[p]
        b = self.buckets[bi]
        b["ready"] += 1
        if b["ready"] == len(b["params"]):
            self._launch(bi)

    @torch.no_grad()
    def _launch(self, bi):
        """The bucket's gradients are final: gather them into the flat buffer with ONE multi-tensor copy (74 per-parameter
        copies with their stream switches cost the host ~2 ms per step), re-point ``param.grad`` at the slices, and issue
        the all-reduce."""
        b = self.buckets[bi]
        self.launch_order.append(bi)
        params, views = b["params"], b["views"]
        grads = [p.grad for p in params]
        dev = b["flat"].device
        side = ops.wgrad_stream(dev) if dev.type == "cuda" else None
        if side is None:
            torch._foreach_copy_(views, grads)
        else:
            # gradients may have been produced on the weight-gradient side stream (ops._wgrad_side): copy there, after
            # everything queued on the main stream, so that the main stream never waits for a weight-gradient GEMM
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                torch._foreach_copy_(views, grads)
            ops.keep_until_join(grads)      # read on the side stream: released when the main stream has re-joined it
        for p, v in zip(params, views):
            p.grad = v
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self.backend == "nccl" else dist.ReduceOp.SUM
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            if side is not None:
                self.comm_stream.wait_stream(side)
            with torch.cuda.stream(self.comm_stream):
                work = dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True)
        self._pending.append((work, bi))

    def finish(self):
        missing = [bi for bi, b in enumerate(self.buckets) if b["ready"] != len(b["params"])]
        if missing:
            raise RuntimeError(f"GradBucketReducer: buckets {missing} never filled (a parameter received no gradient)")
        for work, bi in self._pending:
            work.wait()
            if self.backend != "nccl":
                self.buckets[bi]["flat"].div_(self.world)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            dev = self.params[0].device
            if ops.wgrad_stream(dev) is not None:      # buckets were gathered on the side stream: join it, release the kept gradients
                ops._side_join(dev.index if dev.index is not None else torch.cuda.current_device())
        self._pending.clear()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


# --------------------------------------------------------------------------------------
# model_EMA + checkpoint                                       solver_encoder.py:168-177, :333-346, :147-153
# --------------------------------------------------------------------------------------
class _ParamTable:
    """Device pointer table + chunk map over a parameter list (the layout avc_adam_step / avc_ema_blend walk)."""

    _cache: Dict[tuple, "_ParamTable"] = {}

    def __init__(self, plist):
        import ctypes
        from ._lib import query
        chunk = query("avc_adam_chunk_elems")
        dev = plist[0].device
        rows, pairs = [], []
        for ti, p in enumerate(plist):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("autovc_b200: parameters must be contiguous float32 CUDA tensors (there is no CPU fallback)")
            rows.append([p.data_ptr(), 0, 0, 0, p.numel()])
            pairs += [(ti, c) for c in range((p.numel() + chunk - 1) // chunk)]
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.chunks = torch.tensor(pairs, dtype=torch.int32).to(dev)
        self.nchunks = len(pairs)
        self.ptr = lambda t: ctypes.c_void_p(t.data_ptr())

    @classmethod
    def of(cls, plist):
        key = tuple((p.data_ptr(), p.numel()) for p in plist)
        hit = cls._cache.get(key)
        if hit is None:
            cls._cache.clear()
            hit = cls._cache[key] = cls(plist)
        return hit


@torch.no_grad()
def model_EMA(G, ema: float):
    """``Solver.model_EMA`` (solver_encoder.py:168-177): overwrite every parameter with ``ema * p + (1 - ema) * p`` -- the
    near-identity the reference applies before each checkpoint, reproduced bit for bit (two fp32 products and
    one sum per element, in that order) in ONE launch instead of a 113 MB concatenation, three ATen passes and 74 copies."""
    import ctypes
    from ._lib import call
    plist = [p for p in G.parameters()]
    if not plist:
        return
    tab = _ParamTable.of(plist)
    call("avc_ema_blend", tab.ptr(tab.table), tab.ptr(tab.chunks), tab.nchunks, float(ema),
         ctypes.c_void_p(torch.cuda.current_stream(plist[0].device).cuda_stream))
    ops._GLOBAL_CACHE.begin_step()          # the parameters changed behind autograd's version counters


class AsyncCheckpoint:
    """Handle of a checkpoint being written in the background; ``wait()`` blocks until the file is on disk."""

    def __init__(self, thread, path):
        self._thread, self.path = thread, path

    def wait(self):
        self._thread.join()
        return self.path


def save_checkpoint(G, optimizer, epoch: int, loss: dict, path: str, ema: Optional[float] = None, blocking: bool = False):
    """The checkpoint of solver_encoder.py:333-346: ``model_EMA()`` first (when ``ema`` is given), then
    ``torch.save({'epoch', 'state_dict', 'optimizer', 'loss'}, path)`` with the reference's layout, so that the reference's
    resume code (:147-153) and ``load_checkpoint`` below read it.

    The reference stalls the training loop for the device->host copy and the pickling.  Here the step only pays a
    device-side snapshot (one multi-tensor copy, ~0.3 ms for 340 MB of parameters + Adam state); the snapshot goes to pinned
    host memory on a copy stream and a background thread writes the file.  Returns an ``AsyncCheckpoint``."""
    import copy
    import threading
    if ema is not None:
        model_EMA(G, ema)
    dev = next(G.parameters()).device
    sd = G.state_dict()
    osd = optimizer.state_dict()
    src: List[torch.Tensor] = [v for v in sd.values() if isinstance(v, torch.Tensor) and v.is_cuda]
    for st in osd["state"].values():
        src += [v for v in st.values() if isinstance(v, torch.Tensor) and v.is_cuda]
    snap = [torch.empty_like(t) for t in src]
    if snap:
        torch._foreach_copy_(snap, src)                                   # device-side snapshot on the compute stream
    copy_stream = torch.cuda.Stream(device=dev)
    copy_stream.wait_stream(torch.cuda.current_stream(dev))
    host = []
    with torch.cuda.stream(copy_stream):
        for t in snap:
            h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            h.copy_(t, non_blocking=True)
            host.append(h)
        done = torch.cuda.Event()
        done.record(copy_stream)
    it = iter(host)

    def swap(v):
        # everything the file will hold is detached from the live training state NOW: device tensors through the snapshot,
        # host tensors (Adam's per-parameter ``step`` counters, which the next optimizer step increments in place while the
        # writer thread may not have pickled them yet) and plain values by copy
        if isinstance(v, torch.Tensor):
            return next(it) if v.is_cuda else v.detach().clone()
        return copy.deepcopy(v)
    sd_h = type(sd)((k, swap(v)) for k, v in sd.items())
    osd_h = {"state": {k: {kk: swap(vv) for kk, vv in st.items()} for k, st in osd["state"].items()},
             "param_groups": copy.deepcopy(osd["param_groups"])}
    state = {"epoch": int(epoch), "state_dict": sd_h, "optimizer": osd_h, "loss": copy.deepcopy(dict(loss))}

    def write():
        done.synchronize()
        del snap[:]
        tmp = path + ".tmp"
        torch.save(state, tmp)
        import os
        os.replace(tmp, path)
    th = threading.Thread(target=write, daemon=False)
    th.start()
    handle = AsyncCheckpoint(th, path)
    if blocking:
        handle.wait()
    return handle


def load_checkpoint(path: str, G, optimizer, map_location=None):
    """solver_encoder.py:147-153: restore model and optimizer from a checkpoint (ours or the reference's); returns
    (epoch, loss)."""
    ckpt = torch.load(path, map_location=map_location or next(G.parameters()).device, weights_only=False)
    G.load_state_dict(ckpt["state_dict"])
    optimizer.load_state_dict(ckpt["optimizer"])
    ops._GLOBAL_CACHE.begin_step()
    return ckpt["epoch"], ckpt["loss"]


def nccl_env_defaults():
    """Call BEFORE ``dist.init_process_group("nccl")``.  The persistent recurrence kernels are cooperative launches of 128
    CTAs (one per SM, ~220 KB of shared memory each) that spin on each other: they start only when 128 SMs are free at
    once.  An all-reduce kernel with NCCL's default CTA count takes more than the 20 SMs that are left, so every overlap
    of a bucket with a recurrence turned into a convoy across ranks (measured on 4 x B200: 33.7 ms per step with NCCL's
    default, 14.1-14.4 ms with 4, 8 or 16 CTAs -- 113 MB of gradients per step do not need more over NVLink 5 / NVLS).
    Explicit settings in the environment win."""
    import os
    os.environ.setdefault("NCCL_MAX_CTAS", "16")


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers (done once)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
