"""Drop-in for the optimizer of the training loop: ``torch.optim.Adam(self.G.parameters(), lr)``
(solver_encoder.py:130) stepped at :300.

``FusedAdam`` IS a ``torch.optim.Adam`` (same constructor, ``param_groups``, per-parameter state
``step`` / ``exp_avg`` / ``exp_avg_sq`` and therefore the same ``state_dict()`` layout a reference
checkpoint holds under ``optimizer``, solver_encoder.py:334-339), but ``step()`` runs ONE CUDA
launch over every parameter tensor (``avc_adam_step``) instead of torch's eight multi-tensor passes.
Only the configuration the reference uses is supported: no weight decay, no amsgrad, no maximize.
No CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import List

import torch

from . import _lib, ops
from ._lib import call, query


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise _lib.AvcError("FusedAdam supports the reference configuration only (weight_decay=0, amsgrad=False)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        self._plans = {}
        self.grad_scale = 1.0          # set to 1/world_size by a caller whose all-reduce SUMS the gradients

    def _plan(self, gi: int, plist: List[torch.Tensor]):
        """Static part per group (valid while the parameter tensors stay where they are): chunk map and the
        (param, exp_avg, exp_avg_sq, numel) columns of the pointer table."""
        key = tuple((p.data_ptr(), p.numel()) for p in plist)
        hit = self._plans.get(gi)
        if hit is not None and hit["key"] == key:
            return hit
        dev = plist[0].device
        chunk = query("avc_adam_chunk_elems")
        pairs = []
        rows = []
        for ti, p in enumerate(plist):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.AvcError("FusedAdam needs contiguous float32 CUDA parameters (there is no CPU fallback)")
            st = self.state[p]
            if len(st) == 0:          # torch.optim.Adam._init_group's layout
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                    raise _lib.AvcError("FusedAdam: optimizer state must be contiguous float32 on the parameter's device")
            rows.append([p.data_ptr(), 0, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            pairs += [(ti, c) for c in range((p.numel() + chunk - 1) // chunk)]
        plan = {
            "key": key,
            "chunks": torch.tensor(pairs, dtype=torch.int32).to(dev),
            "nchunks": len(pairs),
            "rows": rows,
            "steps": [self.state[p]["step"] for p in plist],
            "state_ids": [(id(self.state[p]["exp_avg"]), id(self.state[p]["exp_avg_sq"])) for p in plist],
            "gptrs": None,
            "table": torch.empty(len(plist), 5, dtype=torch.int64, device=dev),
        }
        self._plans[gi] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad") or group.get("maximize"):
                raise _lib.AvcError("FusedAdam: weight_decay / amsgrad / maximize are not on the supported path")
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            plan = self._plan(gi, plist)
            # load_state_dict() replaces the state tensors: rebuild the static columns when that happened
            if any((id(self.state[p]["exp_avg"]), id(self.state[p]["exp_avg_sq"])) != ids for p, ids in zip(plist, plan["state_ids"])):
                self._plans.pop(gi)
                plan = self._plan(gi, plist)
            gptrs = []
            for p in plist:
                g = p.grad
                if g.dtype != torch.float32 or g.is_sparse or not g.is_contiguous() or g.device != p.device:
                    raise _lib.AvcError("FusedAdam needs dense contiguous float32 gradients on the parameter's device")
                gptrs.append(g.data_ptr())
            steps = set()
            for p in plist:
                st = self.state[p]["step"]
                st += 1
                steps.add(int(st))
            if len(steps) != 1:
                raise _lib.AvcError("FusedAdam: parameters of one group are at different step counts")
            if gptrs != plan["gptrs"]:
                # the caching allocator usually hands the gradients the same blocks every step: upload only on change.
                # (pageable source + non_blocking: the driver stages the bytes before returning, no stream sync)
                for row, gp in zip(plan["rows"], gptrs):
                    row[1] = gp
                plan["table"].copy_(torch.tensor(plan["rows"], dtype=torch.int64), non_blocking=True)
                plan["gptrs"] = gptrs
            b1, b2 = group["betas"]
            call("avc_adam_step", ctypes.c_void_p(plan["table"].data_ptr()), ctypes.c_void_p(plan["chunks"].data_ptr()),
                 plan["nchunks"], float(group["lr"]), float(b1), float(b2), float(group["eps"]), steps.pop(),
                 float(self.grad_scale), ctypes.c_void_p(torch.cuda.current_stream(plist[0].device).cuda_stream))
        # the kernel writes the parameters through raw pointers, so their autograd version counters do not move: drop the
        # packed-weight cache explicitly (torch.optim.Adam invalidates it through the counters) -- an encoder-only call or a
        # direct sub-module call right after this step must not read packs of the previous weights
        ops._GLOBAL_CACHE.begin_step()
        return loss
