"""AM-PPO on the GPU-resident F-16 env: rollout and update without leaving the device (SURVEY.md 8(f) row 3,
BASELINE.json configs[4]).

Restates what `python train.py --algorithm ppo --use_am_ppo` runs in the reference - SB3's on-policy loop
(stable_baselines3/common/on_policy_algorithm.py:162-341) with the modified PPO.train of
stable_baselines3/ppo/ppo.py:271-455 and its advantage modulation `dynago_transform_advantages` (:29-99) -
against device tensors: observations come from F16BatchedEnv.step as views of device memory, transitions go to
GpuRolloutBuffer (frame-only storage, CUDA GAE and stack-rebuilding gather, rollout.py), the policy is
LMAActorCritic (lma.py), the optimizer DAG (dag.py) or Adam. Nothing is copied to the host inside the loop;
the only synchronisations are the optional logging reads.

Advantage modulation (AM-PPO / "DynAGo" controller), for a batch of raw GAE advantages A:
    N = ||A||_2, sigma = std(A) + eps
    alpha_hat = kappa (N + eps) / (sigma + eps) (p* / (sat + eps))^eta
    alpha <- clamp((1 - rho) alpha + rho alpha_hat)            (only when the EMAs are updated: once per rollout)
    Z = alpha A / (N + eps);  sat <- (1 - rho_sat) sat + rho_sat mean(|Z| > tau)
    A_mod = |A| (kappa_f tanh(Z) + v_shift)
The policy loss uses (optionally re-normalised) A_mod, the value target is A_mod + V_old (ppo.py:331-372).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .constants import ACTION_HIGH, ACTION_LOW
from .dag import DAG
from .lma import LMAActorCritic, LMAConfig, PolicyForwardKernel
from .rollout import GpuRolloutBuffer


@dataclass
class DynagoConfig:
    """train.py:165-178 defaults."""
    tau: float = 1.25
    p_star: float = 0.10
    kappa: float = 2.0              # both the controller's kappa and the tanh formula's (ppo.py:212-218)
    eta: float = 0.3
    rho: float = 0.1
    eps: float = 1e-5
    alpha_min: float = 1e-12
    alpha_max: float = 1e12
    rho_sat: float = 0.98
    alpha_init: float = 1.0
    sat_init: float = 0.10
    v_shift: float = 0.0
    kappa_controller: Optional[float] = None

    def params(self) -> Dict[str, float]:
        return {"kappa": self.kappa if self.kappa_controller is None else self.kappa_controller, "tau": self.tau, "p_star": self.p_star,
                "eta": self.eta, "rho": self.rho, "eps": self.eps, "alpha_min": self.alpha_min, "alpha_max": self.alpha_max,
                "rho_sat": self.rho_sat}


def modulate_advantages(adv: torch.Tensor, params: Dict[str, float], alpha_state: torch.Tensor, sat_state: torch.Tensor,
                        kappa_formula: float, v_shift: float, update_ema: bool = True, group=None) -> torch.Tensor:
    """dynago_transform_advantages (stable_baselines3/ppo/ppo.py:29-99): same arguments, same in-place update
    of the two 1-element EMA state tensors, no host synchronisation. With `group` (a torch.distributed process
    group) the batch statistics are formed from all-reduced sums so that every rank modulates with the moments
    of the whole rollout."""
    n = adv.numel()
    if n <= 1:
        return adv.clone() if n > 0 else torch.zeros((), device=adv.device)
    eps = params["eps"]
    if group is None:
        norm = torch.linalg.norm(adv)
        sigma = (torch.std(adv) + eps) if update_ema else None       # only the controller's target needs it
        count = None
    else:
        import torch.distributed as dist
        a64 = adv.double()
        sums = torch.stack([(a64 * a64).sum(), a64.sum(), torch.tensor(float(n), dtype=torch.float64, device=adv.device)])
        dist.all_reduce(sums, group=group)
        count = sums[2]
        norm = torch.sqrt(sums[0]).float()
        sigma = torch.sqrt(((sums[0] - sums[1] * sums[1] / count) / (count - 1)).clamp_min(0)).float() + eps
    alpha_prev, sat_prev = alpha_state[0].clone(), sat_state[0].clone()
    if update_ema:
        alpha_hat = params["kappa"] * (norm + eps) / (sigma + eps) * (params["p_star"] / (sat_prev + eps)) ** params["eta"]
        alpha = torch.clamp((1 - params["rho"]) * alpha_prev + params["rho"] * alpha_hat, params["alpha_min"], params["alpha_max"])
        alpha_state[0] = alpha.detach()
    else:                      # minibatch calls (ppo.py:331-345): the frozen alpha, no statistics of the controller
        alpha = alpha_prev
    z = alpha * (adv / (norm + eps))
    if update_ema:
        if group is None:
            observed = (z.abs() > params["tau"]).float().mean()
        else:
            import torch.distributed as dist
            hits = (z.abs() > params["tau"]).double().sum()
            dist.all_reduce(hits, group=group)
            observed = (hits / count).float()
        sat_state[0] = ((1 - params["rho_sat"]) * sat_prev + params["rho_sat"] * observed).detach()
    return adv.abs() * (kappa_formula * torch.tanh(z) + v_shift)


@dataclass
class AMPPOConfig:
    """train.py:143-185 defaults."""
    n_steps: int = 2048
    batch_size: int = 256
    n_epochs: int = 10
    learning_rate: float = 9e-5
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.2
    ent_coef: float = 0.0
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    use_am_ppo: bool = True
    optimizer: str = "DAG"            # "DAG" | "Adam"
    norm_adv: bool = True
    dynago: DynagoConfig = field(default_factory=DynagoConfig)
    lma: LMAConfig = field(default_factory=LMAConfig)
    seed: int = 1
    cuda_graph: bool = True           # replay the rollout's policy forward (~50 small launches) as one CUDA graph
    tf32: bool = False                # TF32 tensor-core matmuls for the policy (the reference computes in FP32)
    fused_policy_forward: bool = True  # rollout: the whole policy forward as one kernel (reference-shaped policies only)


class AMPPO:
    """PPO / AM-PPO over an F16BatchedEnv (stacked or ring observation layout)."""

    def __init__(self, env, cfg: Optional[AMPPOConfig] = None, group=None):
        self.env, self.cfg, self.group = env, cfg or AMPPOConfig(), group
        c = self.cfg
        self.device = env.device
        torch.manual_seed(c.seed)
        self.policy = LMAActorCritic(c.lma).to(self.device)
        if c.optimizer.lower() == "dag" and c.use_am_ppo:
            self.optimizer = DAG(self.policy.parameters(), lr=c.learning_rate)                  # ppo.py:259-261
        else:
            self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=c.learning_rate, eps=1e-5)   # ppo.py:253, policies.py
        self.buffer = GpuRolloutBuffer(c.n_steps, env.num_envs, device=self.device, gae_lambda=c.gae_lambda, gamma=c.gamma)
        self.alpha_state = torch.tensor([c.dynago.alpha_init], dtype=torch.float32, device=self.device)
        self.sat_state = torch.tensor([c.dynago.sat_init], dtype=torch.float32, device=self.device)
        self.act_low = torch.as_tensor(ACTION_LOW, device=self.device)
        self.act_high = torch.as_tensor(ACTION_HIGH, device=self.device)
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(c.seed)
        self._obs = None
        self._episode_starts = None
        self._graph = None
        self._fused_act = None
        if c.fused_policy_forward and self.device.type == "cuda" and PolicyForwardKernel.supported(self.policy):
            self._fused_act = PolicyForwardKernel(self.policy, self.act_low, self.act_high)
        if c.tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
        self.num_timesteps = 0
        self.n_updates = 0
        self.last_stats: Dict[str, float] = {}

    # ------------------------------------------------------------------ rollout (on_policy_algorithm.py:162-275)
    def _act(self, obs: torch.Tensor):
        if self._fused_act is not None:          # one launch for the whole forward (csrc/f16_lma_policy.cu)
            noise = torch.randn((obs.shape[0], 4), dtype=torch.float32, device=self.device)
            return self._fused_act(obs, noise)
        actions, values, log_probs = self.policy(obs)
        clipped = torch.maximum(torch.minimum(actions, self.act_high), self.act_low)              # :216
        return actions, values, log_probs, clipped

    def _capture_act(self) -> None:
        """The per-step policy forward is ~50 launches of tiny kernels; at a few thousand envs the rollout is
        launch-bound, so it is captured once (static observation buffer in, static outputs out) and replayed.
        The optimizers update the weights in place, so the graph stays valid across updates."""
        self._g_obs = torch.zeros((self.env.num_envs,) + tuple(self._obs.shape[1:]), dtype=torch.float32, device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._act(self._g_obs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_out = self._act(self._g_obs)

    @torch.no_grad()
    def collect_rollouts(self) -> None:
        c, env = self.cfg, self.env
        if self._obs is None:
            self._obs = env.reset()
            self._episode_starts = torch.ones(env.num_envs, dtype=torch.uint8, device=self.device)
        self.policy.eval()
        if self._fused_act is not None:
            self._fused_act.refresh()            # the update moved the weights: re-pack them (in place, the graph keeps its pointers)
        if c.cuda_graph and self._graph is None:
            self._capture_act()
        self.buffer.reset()
        for step in range(c.n_steps):
            if self._graph is not None:
                self._g_obs.copy_(self._obs)            # the env shifts its observation tensor in place: keep this step's
                self._graph.replay()
                obs = self._g_obs
                actions, values, log_probs, clipped = self._g_out
            else:
                obs = self._obs.contiguous().clone()
                actions, values, log_probs, clipped = self._act(obs)
            new_obs, rewards, dones, truncated = env.step(clipped, auto_reset=True)
            self._note_truncations(step, truncated)
            self.buffer.add(obs, actions, rewards, self._episode_starts, values, log_probs)
            self._obs, self._episode_starts = new_obs, dones.clone()
            self.num_timesteps += env.num_envs
        self._bootstrap_truncations()
        last_values = self.policy.predict_values(self._obs.contiguous())
        self.buffer.compute_returns_and_advantage(last_values, self._episode_starts)

    # Time-limit bootstrap (on_policy_algorithm.py:236-245): reward += gamma * V(terminal observation) for episodes cut
    # at step 1200. The reference does it inside the step loop; asking "was anybody truncated?" there costs one host
    # synchronisation per env-step (it was 60 % of the rollout's host time). The policy does not change during a
    # rollout, so the terminal observations of truncated envs are parked in a side buffer by index arithmetic on the
    # device and all of them are valued once, after the last step - the same numbers, one synchronisation per rollout.
    def _truncation_store(self):
        if getattr(self, "_tr_obs", None) is None:
            n, c = self.env.num_envs, self.cfg
            cap = n * (c.n_steps // 1200 + 1)                       # an env is truncated at most once per 1200 steps
            self._tr_cap = cap
            self._tr_obs = torch.zeros((cap + 1,) + tuple(self.env.terminal_obs.shape[1:]), dtype=torch.float32, device=self.device)
            self._tr_flat = torch.zeros(cap + 1, dtype=torch.int64, device=self.device)
            self._tr_count = torch.zeros((), dtype=torch.int64, device=self.device)
            self._tr_env = torch.arange(n, dtype=torch.int64, device=self.device)
        return self._tr_obs

    def _note_truncations(self, step: int, truncated: torch.Tensor) -> None:
        self._truncation_store()
        if truncated.is_cuda and truncated.dtype == torch.uint8 and self.env.terminal_obs.is_contiguous():
            import ctypes as C

            from . import _lib
            p = lambda t: C.c_void_p(t.data_ptr())          # noqa: E731
            with torch.cuda.device(self.device):            # one launch: slots are handed out by an atomic counter on the device
                _lib.check(_lib.load().f16_rollout_park_truncated(
                    self.env.num_envs, step, self._tr_cap, p(truncated), p(self.env.terminal_obs), p(self._tr_obs), p(self._tr_flat), p(self._tr_count),
                    C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "f16_rollout_park_truncated")
            return
        tr = truncated.to(torch.int64)
        slot = torch.where(tr > 0, self._tr_count + torch.cumsum(tr, 0) - 1, self._tr_cap)      # non-truncated rows go to the spare slot
        self._tr_obs.index_copy_(0, slot, self.env.terminal_obs)
        self._tr_flat.index_copy_(0, slot, self._tr_env + step * self.env.num_envs)
        self._tr_count += tr.sum()

    def _bootstrap_truncations(self) -> None:
        if getattr(self, "_tr_obs", None) is None:
            return
        n = int(self._tr_count.item())
        if n > self._tr_cap:
            raise RuntimeError("more truncated episodes (%d) than an env can have in one rollout (%d)" % (n, self._tr_cap))
        if n:
            v = self.policy.predict_values(self._tr_obs[:n])
            self.buffer.rewards.view(-1).index_add_(0, self._tr_flat[:n], self.cfg.gamma * v)      # rewards are (step, env)
        self._tr_count.zero_()

    # ------------------------------------------------------------------ update (ppo.py:271-455)
    def train(self) -> None:
        c, d = self.cfg, self.cfg.dynago
        params = d.params()
        self.policy.train()
        if c.use_am_ppo:      # EMAs move once per rollout, on all raw advantages (ppo.py:289-306)
            with torch.no_grad():
                modulate_advantages(self.buffer.advantages.flatten(), params, self.alpha_state, self.sat_state, d.kappa, d.v_shift,
                                    update_ema=True, group=self.group)
        pg, vl, kl, cf = [], [], [], []
        for _ in range(c.n_epochs):
            for batch in self.buffer.get(c.batch_size, generator=self.generator):
                values, log_prob, entropy = self.policy.evaluate_actions(batch.observations, batch.actions)
                adv_raw = batch.advantages
                if c.use_am_ppo:
                    adv_mod = modulate_advantages(adv_raw, params, self.alpha_state, self.sat_state, d.kappa, d.v_shift, update_ema=False)
                    adv = (adv_mod - adv_mod.mean()) / (adv_mod.std() + 1e-8) if (c.norm_adv and adv_mod.numel() > 1) else adv_mod
                    target_values = adv_mod + batch.old_values                                  # ppo.py:368-369
                else:
                    adv = (adv_raw - adv_raw.mean()) / (adv_raw.std() + 1e-8) if adv_raw.numel() > 1 else adv_raw
                    target_values = batch.returns
                ratio = torch.exp(log_prob - batch.old_log_prob)
                policy_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - c.clip_range, 1 + c.clip_range)).mean()
                value_loss = F.mse_loss(target_values, values)
                entropy_loss = -entropy.mean()
                loss = policy_loss + c.ent_coef * entropy_loss + c.vf_coef * value_loss
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                if self.group is not None:
                    self._allreduce_grads()
                torch.nn.utils.clip_grad_norm_(self.policy.parameters(), c.max_grad_norm)
                self.optimizer.step()
                with torch.no_grad():
                    log_ratio = log_prob - batch.old_log_prob
                    pg.append(policy_loss.detach()); vl.append(value_loss.detach())
                    kl.append(((torch.exp(log_ratio) - 1) - log_ratio).mean()); cf.append(((ratio - 1).abs() > c.clip_range).float().mean())
            self.n_updates += 1
        with torch.no_grad():
            self.last_stats = {"policy_gradient_loss": float(torch.stack(pg).mean()), "value_loss": float(torch.stack(vl).mean()),
                               "approx_kl": float(torch.stack(kl).mean()), "clip_fraction": float(torch.stack(cf).mean()),
                               "std": float(torch.exp(self.policy.log_std).mean()), "alpha_A_ema": float(self.alpha_state[0]),
                               "prev_saturation_A_ema": float(self.sat_state[0]), "n_updates": self.n_updates}

    def _allreduce_grads(self) -> None:
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        grads = [p.grad for p in self.policy.parameters() if p.grad is not None]
        flat = torch.cat([g.flatten() for g in grads])
        dist.all_reduce(flat, group=self.group)
        flat /= world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    # ------------------------------------------------------------------ checkpoint / resume (train.py always saves its model)
    def save(self, path: str) -> None:
        """Everything the learner needs to continue as if it had not stopped: policy, optimizer (for DAG including the
        per-tensor alpha / saturation statistics, the RMS-shrink state and the step counter), the two advantage-modulation
        EMAs, the sampling generator and the counters. The policy's state_dict keys are the reference policy's."""
        torch.save({"policy": self.policy.state_dict(), "optimizer": self.optimizer.state_dict(),
                    "optimizer_class": type(self.optimizer).__name__,
                    "alpha_state": self.alpha_state.cpu(), "sat_state": self.sat_state.cpu(),
                    "generator": self.generator.get_state(), "num_timesteps": self.num_timesteps, "n_updates": self.n_updates,
                    "cfg": self.cfg}, path)

    def load(self, path: str) -> "AMPPO":
        ck = torch.load(path, map_location=self.device, weights_only=False)
        if ck["optimizer_class"] != type(self.optimizer).__name__:
            raise ValueError("checkpoint was written with %s, this learner uses %s" % (ck["optimizer_class"], type(self.optimizer).__name__))
        self.policy.load_state_dict(ck["policy"])
        self.optimizer.load_state_dict(ck["optimizer"])
        self.alpha_state.copy_(ck["alpha_state"])
        self.sat_state.copy_(ck["sat_state"])
        self.generator.set_state(ck["generator"].cpu())
        self.num_timesteps, self.n_updates = int(ck["num_timesteps"]), int(ck["n_updates"])
        self._graph = None                      # the captured policy forward holds the old weights' addresses only; recapture lazily
        return self

    def learn(self, total_timesteps: int) -> "AMPPO":
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            self.train()
        return self
