"""Adam on a flat parameter buffer, fused with the data-parallel gradient average.

ref: train_stage_rays_auto.py:200-210 (`torch.optim.Adam([{'params': trainable_parameters}], lr=cfg.optimizer.lr)`),
     :494-509 (`optimizer.step(); optimizer.zero_grad()`, then lr = lr0 * decay_factor ** (i / (lr_decay * 1000))).

`FlatAdam` is a `torch.optim.Optimizer`: same constructor keywords (lr, betas, eps), `param_groups[..]["lr"]` can be
overwritten every step as the script does, `state_dict()` holds `exp_avg` / `exp_avg_sq` / `step` per parameter.  On
construction it re-homes every parameter into ONE flat fp32 buffer (the parameters become views of it), and keeps the
two moments and a gradient staging buffer in the same layout.  `step()` is then three launches instead of a walk over
124 tensors: gather the gradients into the flat buffer, (optionally) one sum all-reduce of that buffer, one
`sahs_adam_step` kernel that folds the 1 / world-size average into its gradient read.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import lib as L


def exp_lr(lr0: float, decay_factor: float, decay_steps: float, iteration: int) -> float:
    """Learning rate of iteration i (ref: train_stage_rays_auto.py:503-507): lr0 * decay_factor ** (i / decay_steps),
    decay_steps = cfg.scheduler.lr_decay * 1000."""
    return lr0 * (decay_factor ** (iteration / decay_steps))


class FlatAdam(torch.optim.Optimizer):
    """`capturable=True`: the step count and the exponential learning-rate schedule (`schedule=(decay_factor,
    decay_steps)`, ref: train_stage_rays_auto.py:503-509; None = constant lr) live in device memory, `step()` is
    `sahs_adam_step_dev` + `sahs_adam_advance` with no host-side state, so a whole training step -- sampler, forward,
    loss, backward, all-reduce, this step -- can be recorded in a CUDA graph (sahs_b200.train.GraphedStep).  In this mode
    `param_groups[..]["lr"]` is the INITIAL rate; the running one is `current_lr()` (reads the device)."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, group=None,
                 capturable: bool = False, schedule=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("bad Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        L.load()                                             # fail loudly without the CUDA library
        self._group = group
        self._segments = []                                  # (param_group index, begin, end) in the flat buffers
        plist = [p for g in self.param_groups for p in g["params"]]
        if not plist:
            raise ValueError("no parameters")
        dev = plist[0].device
        for p in plist:
            if p.device != dev or not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FlatAdam needs fp32 CUDA parameters on one device (there is no CPU path)")
        # every parameter starts on a 16-byte boundary so the kernel's float4 accesses never straddle two tensors' ends
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self._n = total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._steps = 0
        self._step_t = torch.tensor(0.0)                     # one counter object shared by every parameter's state
        self._capturable = bool(capturable)
        self._hyper = None
        if capturable:
            if len(self.param_groups) != 1:
                raise ValueError("capturable FlatAdam supports one parameter group")
            factor, steps = schedule if schedule is not None else (1.0, 1.0)
            self._hyper = torch.tensor([lr, float(factor), float(steps), betas[0], betas[1], eps, 1.0, 0.0],
                                       dtype=torch.float64, device=dev)
        self._grad_views = []
        with torch.no_grad():
            for p, off in zip(plist, offs):
                n = p.numel()
                view = self.flat_param[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # the parameter now lives in the flat buffer
                self._grad_views.append(self.flat_grad[off:off + n].view(p.shape))
                self.state[p] = {"step": self._step_t, "exp_avg": self.flat_exp_avg[off:off + n].view(p.shape),
                                 "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view(p.shape)}
        i = 0
        for gi, g in enumerate(self.param_groups):
            b = offs[i] if g["params"] else 0
            i += len(g["params"])
            e = offs[i] if i < len(offs) else total
            self._segments.append((gi, b, e))
        self._params = plist        # (moving the storage changes data_ptr(), which the packed-weight cache keys on)
        self._offs = offs

    def load_state_dict(self, state_dict) -> None:
        """Resume (ref: train_stage_rays_auto.py:253, :705 `optimizer.load_state_dict(checkpoint["optimizer_state_dict"])`;
        the checkpoint may come from this class or from the reference's torch.optim.Adam -- same per-parameter layout).
        The base class replaces `state[p]` with fresh tensors; the moments are copied back into the flat buffers the
        kernel reads, the per-parameter entries are re-pointed at views of them and the step counter is restored."""
        super().load_state_dict(state_dict)
        steps = 0
        with torch.no_grad():
            for p, off in zip(self._params, self._offs):
                st = self.state.get(p)
                n = p.numel()
                views = {"exp_avg": self.flat_exp_avg[off:off + n].view(p.shape),
                         "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view(p.shape)}
                if not st:                                   # no state saved for this parameter: starts from zero
                    for v in views.values():
                        v.zero_()
                    self.state[p] = {"step": self._step_t, **views}
                    continue
                for name, v in views.items():
                    if name in st:
                        v.copy_(st[name].to(v.device, torch.float32))
                    else:
                        v.zero_()
                    st[name] = v
                steps = max(steps, int(float(st.get("step", 0))))
                st["step"] = self._step_t
        self._steps = steps
        self._step_t.fill_(float(steps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        # gradients -> flat staging buffer.  A parameter without a gradient is SKIPPED, as torch.optim.Adam does (no
        # moment decay, no update): the kernel then runs over the contiguous runs of parameters that have one.  (One
        # difference remains: the bias-correction step count is shared by all parameters, torch keeps one per parameter.)
        have = [(v, p.grad) for v, p in zip(self._grad_views, self._params) if p.grad is not None]
        partial = len(have) != len(self._params)
        if partial:
            self.flat_grad.zero_()                              # (all-reduced with the other ranks' buffers below)
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1:
            dist.all_reduce(self.flat_grad, group=self._group)
            scale = 1.0 / dist.get_world_size(self._group)
        stream = L.stream_ptr(self.flat_param.device)
        if self._capturable:
            if partial:
                raise RuntimeError("capturable FlatAdam needs a gradient for every parameter")
            L.check(lib.sahs_adam_step_dev(self.flat_param.data_ptr(), self.flat_grad.data_ptr(),
                                           self.flat_exp_avg.data_ptr(), self.flat_exp_avg_sq.data_ptr(), self._n,
                                           self._hyper.data_ptr(), scale, stream), "adam_step_dev")
            L.check(lib.sahs_adam_advance(self._hyper.data_ptr(), stream), "adam_advance")
            return loss
        self._steps += 1
        for gi, b, e in (self._live_runs() if partial else self._segments):
            g = self.param_groups[gi]
            if e <= b:
                continue
            L.check(lib.sahs_adam_step(self.flat_param.data_ptr() + 4 * b, self.flat_grad.data_ptr() + 4 * b,
                                       self.flat_exp_avg.data_ptr() + 4 * b, self.flat_exp_avg_sq.data_ptr() + 4 * b,
                                       e - b, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                       float(g["eps"]), self._steps, scale, stream), "adam_step")
        self._step_t += 1
        return loss

    def _live_runs(self):
        """(group, begin, end) for every maximal run of consecutive parameters that have a gradient."""
        runs, i = [], 0
        for gi, g in enumerate(self.param_groups):
            cur = None
            for _ in g["params"]:
                p, off = self._params[i], self._offs[i]
                end = self._offs[i + 1] if i + 1 < len(self._offs) else self._n
                if p.grad is not None:
                    cur = [gi, off, end] if cur is None else [gi, cur[1], end]
                elif cur is not None:
                    runs.append(tuple(cur))
                    cur = None
                i += 1
            if cur is not None:
                runs.append(tuple(cur))
        return runs

    def current_lr(self) -> float:
        """Learning rate of the next step (capturable mode: computed from the device-side schedule; one sync)."""
        if not self._capturable:
            return float(self.param_groups[0]["lr"])
        h = self._hyper.cpu().tolist()
        return h[0] * (h[1] ** (h[7] / h[2]))

    def steps_done(self) -> int:
        return int(self._hyper[6].item()) - 1 if self._capturable else self._steps

    def averaged_gradients(self) -> Optional[torch.Tensor]:
        """The flat gradient buffer of the last step (summed over ranks; multiply by 1 / world size for the mean)."""
        return self.flat_grad
