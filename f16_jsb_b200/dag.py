"""DAG optimizer (Dynamic-Alpha Gradient) for device-resident AM-PPO training.

The optimizer train.py selects by default for AM-PPO (`--am_ppo_optimizer DAG`, train.py:164;
stable_baselines3/ppo/ppo.py:259-261), defined in stable_baselines3/ppo/optim/sgd.py:87-344: per parameter
tensor the gradient is normalised by its L2 norm, stretched by an adaptive alpha and squashed,
    update = k * s_t * tanh(alpha / s_t * g / (||g|| + eps)),
alpha follows alpha_hat = kappa (||g||+eps)/(sigma+eps) (d_l/d)^beta (p*/(sat+eps))^eta s_t by an EMA (rho),
sigma = ||g||/sqrt(d_l), `sat` is the fraction of |alpha g/||g||| above tau (refreshed every `sat_every` steps),
and a global RMS-shrink factor s_t in [s_min, 1] follows the update RMS against its warm-up level.

This version keeps every statistic on the device: the reference reads the update RMS back to the host each
step (`.item()`, sgd.py:322) and keeps s_t as a Python float; here s_t and the RMS averages are 0-dim float64
device tensors, so a step never synchronises. The exact-sigma branch (Apex multi_tensor_std, sgd.py:75-85,
off by default and unavailable without Apex) is not provided.
"""
import math
from typing import Callable, Iterable, Optional

import torch
from torch.optim.optimizer import Optimizer


class DAG(Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0, weight_decay: float = 0.0,
                 k_val: float = 2.0, k_sched: Optional[Callable[[int], float]] = None, nesterov: bool = False, maximize: bool = False,
                 hyper: Optional[dict] = None, shrink: Optional[dict] = None, sat_every: int = 10):
        h = dict(tau=1.25, p_star=0.10, kappa=None, beta=1 / 3, eta=0.3, rho=0.1, eps=1e-5, alpha_min=1e-12, alpha_max=1e12)
        if hyper:
            h.update(hyper)
        if h["kappa"] is None:          # tau / Phi^-1(1 - p*/2): a unit Gaussian saturates with probability p*
            h["kappa"] = h["tau"] / torch.distributions.Normal(0.0, 1.0).icdf(torch.tensor(1 - h["p_star"] / 2)).item()
        self.h = h
        s = dict(lambda_rms=0.3, s_min=0.1, gamma=1.0, ema_beta=0.98, warmup_steps=500)
        if shrink:
            s.update(shrink)
        self.s_cfg = s
        self.k_val, self.k_sched = float(k_val), k_sched
        self.sat_every = max(1, int(sat_every))
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                                      maximize=maximize))
        self.d_total = sum(p.numel() for g in self.param_groups for p in g["params"] if p.requires_grad)
        self.global_step = 0
        self._s_t = None          # 0-dim float64 tensors on the parameters' device, created at the first step
        self._rms_t = None
        self._rms0 = None

    @property
    def s_t(self) -> float:
        return 1.0 if self._s_t is None else float(self._s_t)

    # ---- checkpointing: the RMS-shrink state and the step counter are optimizer state too
    def state_dict(self):
        sd = super().state_dict()
        sd["dag"] = {"global_step": int(self.global_step), "k_val": float(self.k_val),
                     "s_t": None if self._s_t is None else self._s_t.detach().cpu().clone(),
                     "rms_t": None if self._rms_t is None else self._rms_t.detach().cpu().clone(),
                     "rms0": None if self._rms0 is None else self._rms0.detach().cpu().clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("dag", None)
        super().load_state_dict(state_dict)
        dev = next((p.device for g in self.param_groups for p in g["params"]), torch.device("cpu"))
        for key, st in self.state.items():                      # the stacked group statistics live on the parameters' device
            if isinstance(key, str) and key.startswith("_group_"):
                for k, v in st.items():
                    if torch.is_tensor(v):
                        st[k] = v.to(dev)
        if extra is not None:
            self.global_step = int(extra["global_step"])
            self.k_val = float(extra["k_val"])
            self._s_t = None if extra["s_t"] is None else extra["s_t"].to(dev)
            self._rms_t = None if extra["rms_t"] is None else extra["rms_t"].to(dev)
            self._rms0 = None if extra["rms0"] is None else extra["rms0"].to(dev)

    def set_k_val(self, new_k: float) -> None:
        self.k_val, self.k_sched = float(new_k), None

    @staticmethod
    def cosine_decay(k0: float, total_steps: int) -> Callable[[int], float]:
        return lambda step: k0 * 0.5 * (1.0 + math.cos(math.pi * min(step, total_steps) / float(total_steps)))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            with torch.enable_grad():
                closure()
        if self.k_sched is not None:
            self.k_val = float(self.k_sched(self.global_step))
        h, sc = self.h, self.s_cfg
        total_sq, total_n = None, 0
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            if self._s_t is None:
                self._s_t = torch.ones((), dtype=torch.float64, device=dev)
            # stacked per-tensor statistics of the group, keyed by the group's INDEX so that they survive
            # state_dict() / load_state_dict() into a new process (the reference keeps them per parameter, sgd.py:196-214)
            st = self.state.setdefault("_group_%d" % gi, {})
            if "alpha" not in st or st["alpha"].numel() != len(params):
                st["alpha"] = torch.ones(len(params), dtype=torch.float32, device=dev)
                st["sat"] = torch.zeros(len(params), dtype=torch.float32, device=dev)
                st["d"] = torch.tensor([float(p.numel()) for p in params], dtype=torch.float32, device=dev)
            grads = [p.grad for p in params]
            norms = torch.stack(torch._foreach_norm(grads, 2))
            sigma = norms / torch.sqrt(st["d"])
            s_t = self._s_t.to(torch.float32)
            alpha_hat = (h["kappa"] * (norms + h["eps"]) / (sigma + h["eps"]) * (st["d"] / self.d_total) ** h["beta"]
                         * (h["p_star"] / (st["sat"] + h["eps"])) ** h["eta"] * s_t)
            alpha = ((1 - h["rho"]) * st["alpha"] + h["rho"] * alpha_hat).clamp_(h["alpha_min"], h["alpha_max"])
            st["alpha"] = alpha
            scale = (alpha / s_t / (norms + h["eps"])).unbind()              # alpha / s_t / (||g|| + eps), one scalar per tensor
            scaled = torch._foreach_mul(grads, list(scale))
            if self.global_step % self.sat_every == 0:
                # fraction of |scaled| above tau per tensor (sgd.py: (x.abs() > tau).float().mean()) from one flat pass and a
                # prefix sum instead of four launches per parameter tensor; counts are integers < 2^24: exact in float32
                if "ends" not in st or st["ends"].numel() != len(params):
                    st["ends"] = torch.cumsum(st["d"].to(torch.int64), 0) - 1
                hits = torch.cumsum((torch.cat([t.reshape(-1) for t in scaled]).abs_() > h["tau"]).to(torch.int32), 0)[st["ends"]]
                st["sat"] = torch.diff(hits, prepend=hits.new_zeros(1)).to(torch.float32) / st["d"]
            torch._foreach_tanh_(scaled)
            updates = torch._foreach_mul(scaled, (self.k_val * self._s_t).to(torch.float32))
            if group["weight_decay"]:
                torch._foreach_add_(updates, params, alpha=group["weight_decay"])
            m = group["momentum"]
            if m != 0.0:
                bufs = []
                for p in params:
                    ps = self.state[p]
                    if "momentum_buffer" not in ps:
                        ps["momentum_buffer"] = torch.zeros_like(p)
                    bufs.append(ps["momentum_buffer"])
                torch._foreach_mul_(bufs, m)
                torch._foreach_add_(bufs, updates, alpha=1 - group["dampening"])
                updates = torch._foreach_add(updates, bufs, alpha=m) if group["nesterov"] else bufs
            if group["maximize"]:
                updates = torch._foreach_neg(updates)
            torch._foreach_add_(params, updates, alpha=-group["lr"])
            # sum of squares of the whole update (sgd.py:318-322) from one multi-tensor norm instead of three launches
            # per parameter tensor; the squares are added in float64
            sq = (torch.stack(torch._foreach_norm(updates, 2)).double() ** 2).sum()
            total_sq = sq if total_sq is None else total_sq + sq
            total_n += sum(p.numel() for p in params)
        if total_n:
            rms_now = torch.sqrt(total_sq / total_n)
            b = sc["ema_beta"]
            self._rms_t = rms_now if self._rms_t is None else b * self._rms_t + (1 - b) * rms_now
            if self.global_step < sc["warmup_steps"]:
                self._rms0 = self._rms_t if self._rms0 is None else b * self._rms0 + (1 - b) * self._rms_t
            else:
                if self._rms0 is None:
                    self._rms0 = self._rms_t
                ratio = (self._rms_t / (sc["lambda_rms"] * self._rms0)).clamp(0.0, 1.0)
                self._s_t = sc["s_min"] + (1 - sc["s_min"]) * ratio ** sc["gamma"]
        self.global_step += 1
        return None
