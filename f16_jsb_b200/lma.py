"""LMA feature extractor and actor-critic heads for device-resident training on the F-16 env.

What the reference builds for `PPO('MlpPolicy', env, policy_kwargs=...)` (train.py:21-32,81-84):
StackedLMAFeaturesExtractor (jsbsim_gym/LMA_features.py:636-777) = per-frame JSBSimFeatureExtractor
(jsbsim_gym/features.py:37-67, here the CUDA kernel behind include/f16_features.h) -> LMAFeaturesExtractor
(:410-530): Linear 17->64 + ReLU + sinusoidal positions, head stacking 4 x 16 along the sequence, re-chunk of the
640 values into L'=5 tokens of 128, Linear 128->32 + ReLU, two pre-LayerNorm blocks (4-head attention over the 5
tokens, GELU MLP 32->128->32), flatten to 160 features; then SB3's ActorCriticPolicy heads
(stable_baselines3/common/policies.py:416-760, torch_layers.py MlpExtractor): tanh MLPs pi [64,64] / vf [128,64],
Linear action mean, state-independent log-std, Linear value.

Parameter names follow the reference's state_dict keys, so checkpoints interchange
(`lma_extractor.initial_transform.input_embedding.weight`, `lma_extractor.lma_blocks.0.attn.c_attn.weight`, ...;
heads: `mlp_extractor.policy_net.0.weight`, `action_net.weight`, `value_net.weight`, `log_std`). The code is
this repo's own: the stacking / re-chunk is one permute, attention is one fused SDPA call over (B, 4, 5, 8).
"""
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .constants import NUM_FEATURES, NUM_STACKED_FRAMES
from .features import FEATURES_PER_FRAME, jsbsim_features


def closest_divisor(total: int, target: int, max_delta: int = 100) -> int:
    """Divisor of `total` nearest to `target`, smaller one first on ties within +-max_delta
    (find_closest_divisor, jsbsim_gym/LMA_features.py:23-96)."""
    if total <= 0:
        raise ValueError("total must be positive")
    target = max(1, int(target))
    if total % target == 0:
        return target
    for delta in range(1, max_delta + 1):
        lo, hi = target - delta, target + delta
        if lo > 0 and total % lo == 0:
            return lo
        if total % hi == 0:
            return hi
    divisors = [d for i in range(1, int(math.isqrt(total)) + 1) if total % i == 0 for d in (i, total // i)]
    return min(divisors, key=lambda d: (abs(d - target), d))


@dataclass
class LMAConfig:
    """Shapes of the extractor (LMAConfigRL, jsbsim_gym/LMA_features.py:99-169; defaults = train.py:21-32)."""
    seq_len: int = NUM_STACKED_FRAMES
    in_features: int = FEATURES_PER_FRAME
    embed_dim: int = 64
    num_heads_stacking: int = 4
    num_heads_latent: int = 4
    ff_hidden: int = 128
    num_layers: int = 2
    dropout: float = 0.1
    bias: bool = True

    def __post_init__(self):
        if self.embed_dim % self.num_heads_stacking:
            raise ValueError("embed_dim must be divisible by num_heads_stacking")
        self.d_new = self.embed_dim // 2                                 # StackedLMAFeaturesExtractor :706
        if self.d_new % self.num_heads_latent:
            raise ValueError("d_new must be divisible by num_heads_latent")
        total = self.seq_len * self.embed_dim
        self.l_new = closest_divisor(total, self.seq_len // 2)           # :705, LMAConfigRL.__post_init__
        self.c_new = total // self.l_new
        self.features_dim = self.l_new * self.d_new


def sinusoidal_positions(seq_len: int, dim: int) -> torch.Tensor:
    pos = torch.arange(seq_len, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros(seq_len, dim)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe


def features17_torch(frames: torch.Tensor) -> torch.Tensor:
    """PyTorch form of the per-frame transform for tensors that are not on a GPU (tests; the CUDA kernel
    f16_features17 is what runs on device tensors)."""
    pos, mach, ab, rates = frames[..., 0:3], frames[..., 3:4], frames[..., 4:6], frames[..., 6:9]
    pt, psi, goal = frames[..., 9:11], frames[..., 11:12], frames[..., 12:15]
    disp = goal - pos
    dist = torch.sqrt((disp[..., :2] ** 2).sum(-1, keepdim=True))
    rel = torch.atan2(disp[..., 1:2], disp[..., 0:1]) - psi
    return torch.cat([1 / (1 + dist * 1e-3), disp[..., 2:3] / 15000, pos[..., 2:3] / 15000, mach, rates, torch.cos(ab), torch.sin(ab),
                      torch.cos(pt), torch.sin(pt), torch.cos(rel), torch.sin(rel)], dim=-1)


class _LatentAttentionFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) v over the five latent tokens by the hand-written kernels of
    csrc/f16_lma_attention.cu (include/f16_lma.h): (B, T, 3*H*dh) fused projection in, (B, T, H*dh) out."""

    @staticmethod
    def forward(ctx, qkv: torch.Tensor, heads: int, dropout_p: float):
        import ctypes as C

        from . import _lib
        qkv = qkv.contiguous()
        b, t, d3 = qkv.shape
        d = d3 // 3
        y = torch.empty((b, t, d), dtype=torch.float32, device=qkv.device)
        # the mask is a function of (seed, sample, head): the backward kernel regenerates it
        seed = int(torch.empty((), dtype=torch.int64).random_()) if dropout_p > 0 else 0
        stream = C.c_void_p(torch.cuda.current_stream(qkv.device).cuda_stream)
        with _lib.device_guard(qkv.device):
            _lib.check(_lib.load().f16_lma_attention_forward(b, t, heads, d // heads, C.c_void_p(qkv.data_ptr()), C.c_void_p(y.data_ptr()),
                                                             float(dropout_p), seed, stream), "f16_lma_attention_forward")
        ctx.save_for_backward(qkv)
        ctx.meta = (heads, float(dropout_p), seed)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib
        (qkv,) = ctx.saved_tensors
        heads, p, seed = ctx.meta
        b, t, d3 = qkv.shape
        dy = dy.contiguous()
        dqkv = torch.empty_like(qkv)
        stream = C.c_void_p(torch.cuda.current_stream(qkv.device).cuda_stream)
        with _lib.device_guard(qkv.device):
            _lib.check(_lib.load().f16_lma_attention_backward(b, t, heads, d3 // 3 // heads, C.c_void_p(qkv.data_ptr()), C.c_void_p(dy.data_ptr()),
                                                              C.c_void_p(dqkv.data_ptr()), p, seed, stream), "f16_lma_attention_backward")
        return dqkv, None, None


def latent_attention(qkv: torch.Tensor, heads: int, dropout_p: float = 0.0) -> torch.Tensor:
    """(B, T, 3*D) -> (B, T, D). CUDA float32 tensors with T = 5 and D / heads = 8 (the reference's LMA shape)
    go through the hand-written kernels; anything else through torch's scaled_dot_product_attention."""
    b, t, d3 = qkv.shape
    d = d3 // 3
    if qkv.is_cuda and qkv.dtype == torch.float32 and t == 5 and d // heads == 8:
        return _LatentAttentionFn.apply(qkv, heads, dropout_p)
    q, k, v = qkv.view(b, t, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    return F.scaled_dot_product_attention(q, k, v, dropout_p=dropout_p).transpose(1, 2).reshape(b, t, d)


def _new_seed() -> int:
    return int(torch.empty((), dtype=torch.int64).random_())


def _stream(t: torch.Tensor):
    import ctypes as C
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _DropoutAddFn(torch.autograd.Function):
    """z + dropout(x) in one pass (csrc/f16_lma_elementwise.cu); the mask is regenerated from the seed in the backward."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, z: torch.Tensor, p: float):
        import ctypes as C

        from . import _lib
        x, z = x.contiguous(), z.contiguous()
        y = torch.empty_like(x)
        seed = _new_seed()
        with _lib.device_guard(x.device):
            _lib.check(_lib.load().f16_lma_dropout_add_forward(x.numel(), C.c_void_p(x.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()),
                                                               float(p), seed, _stream(x)), "f16_lma_dropout_add_forward")
        ctx.meta = (float(p), seed)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib
        p, seed = ctx.meta
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(dy)
            with _lib.device_guard(dy.device):
                _lib.check(_lib.load().f16_lma_dropout_backward(dy.numel(), C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()), p, seed, _stream(dy)),
                           "f16_lma_dropout_backward")
        return dx, (dy if ctx.needs_input_grad[1] else None), None


def dropout_add(x: torch.Tensor, z: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """z + dropout(x, p). Training on CUDA float32 tensors goes through the fused kernel; anything else through torch."""
    if (training and p > 0 and x.is_cuda and x.dtype == torch.float32 and z.dtype == torch.float32 and x.shape == z.shape
            and x.numel() % 8 == 0 and _LinearFn.use_fused_elementwise):
        return _DropoutAddFn.apply(x, z, p)
    return z + F.dropout(x, p, training)


class _EmbedActFn(torch.autograd.Function):
    """dropout(relu(a) + positions) in one pass; backward da = keep * dy * (a > 0) from the saved pre-activation. With
    stack_heads = H > 1 the result (B, T, C) comes back head-stacked - y.view(B, T, H, C/H).permute(0, 2, 1, 3) flattened to
    (B, T * C) - and the backward reads dy in that order: the permutation copies are never made."""

    @staticmethod
    def forward(ctx, a: torch.Tensor, pos: torch.Tensor, p: float, stack_heads: int = 1):
        import ctypes as C

        from . import _lib
        a, pos = a.contiguous(), pos.contiguous()
        seq, ch = pos.shape
        y = torch.empty_like(a) if stack_heads == 1 else torch.empty((a.numel() // (seq * ch), seq * ch), dtype=a.dtype, device=a.device)
        seed = _new_seed()
        with _lib.device_guard(a.device):
            _lib.check(_lib.load().f16_lma_embed_act_forward(a.numel() // ch, ch, seq, stack_heads, C.c_void_p(a.data_ptr()), C.c_void_p(pos.data_ptr()),
                                                             C.c_void_p(y.data_ptr()), float(p), seed, _stream(a)), "f16_lma_embed_act_forward")
        ctx.save_for_backward(a)
        ctx.meta = (float(p), seed, ch, seq, stack_heads)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib
        (a,) = ctx.saved_tensors
        p, seed, ch, seq, stack_heads = ctx.meta
        dy = dy.contiguous()
        da = torch.empty_like(a)
        with _lib.device_guard(a.device):
            _lib.check(_lib.load().f16_lma_embed_act_backward(a.numel() // ch, ch, seq, stack_heads, C.c_void_p(a.data_ptr()), C.c_void_p(dy.data_ptr()),
                                                              C.c_void_p(da.data_ptr()), p, seed, _stream(a)), "f16_lma_embed_act_backward")
        return da, None, None, None


class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int, dropout: float, bias: bool):
        super().__init__()
        self.heads, self.p = heads, dropout
        self.c_attn = Linear(dim, 3 * dim, bias=bias)
        self.c_proj = Linear(dim, dim, bias=bias)

    def forward(self, z: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """drop(proj(attention(z))), or residual + that in one pass when a residual is given."""
        y = self.c_proj(latent_attention(self.c_attn(z), self.heads, self.p if self.training else 0.0))
        if residual is None:
            return F.dropout(y, self.p, self.training)
        return dropout_add(y, residual, self.p, self.training)


class _MLP(nn.Module):
    def __init__(self, dim: int, hidden: int, dropout: float, bias: bool):
        super().__init__()
        self.p = dropout
        self.c_fc = Linear(dim, hidden, bias=bias)
        self.c_proj = Linear(hidden, dim, bias=bias)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = self.c_proj(F.gelu(self.c_fc(x)))
        if residual is None:
            return F.dropout(y, self.p, self.training)
        return dropout_add(y, residual, self.p, self.training)


def _linear_tc(x2: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x2 (rows, in) @ weight (out, in)^T + bias on the tensor cores (csrc/f16_lma_linear.cu): float32, contiguous."""
    import ctypes as C

    from . import _lib
    y = torch.empty((x2.shape[0], weight.shape[0]), dtype=torch.float32, device=x2.device)
    stream = C.c_void_p(torch.cuda.current_stream(x2.device).cuda_stream)
    with _lib.device_guard(x2.device):
        _lib.check(_lib.load().f16_lma_linear_forward(x2.shape[0], x2.shape[1], weight.shape[0], C.c_void_p(x2.data_ptr()),
                                                      C.c_void_p(weight.data_ptr()), C.c_void_p(bias.data_ptr() if bias is not None else 0),
                                                      C.c_void_p(y.data_ptr()), stream), "f16_lma_linear_forward")
    return y


_TC_SHAPES = {}
_WG_SHAPES = {}


def _tc_supported(in_features: int, out_features: int) -> bool:
    key = (in_features, out_features)
    if key not in _TC_SHAPES:
        from . import _lib
        _TC_SHAPES[key] = bool(_lib.load().f16_lma_linear_supported(in_features, out_features))
    return _TC_SHAPES[key]


def _wgrad_tc_supported(in_features: int, out_features: int) -> bool:
    key = (in_features, out_features)
    if key not in _WG_SHAPES:
        from . import _lib
        _WG_SHAPES[key] = bool(_lib.load().f16_lma_linear_wgrad_tc_supported(in_features, out_features))
    return _WG_SHAPES[key]


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b for the policy's tall-skinny layers (<= 160 features, 10^5..10^6 rows). Forward and input gradient
    (dx = dy W, the same kernel with W^T as the weight) run on the tensor cores with split TF32 operands
    (csrc/f16_lma_linear.cu: FP32-accurate, 4.4-5.1 TB/s where the library's FP32 GEMMs reach 1.0-1.9); the weight /
    bias gradients, whose reduction axis is the batch, come from csrc/f16_lma_wgrad.cu (include/f16_lma.h). Shapes the
    tensor-core kernels do not build (the 4- and 1-wide output heads; for the weight gradient also 17 input features and
    160 -> 128) stay torch matmuls / the FP32 slab kernel."""

    use_tc = os.environ.get("F16_LMA_TC", "1") != "0"        # class-wide switches (A/B measurements, tests)
    use_wgrad_tc = os.environ.get("F16_LMA_WGRAD_TC", "1") != "0"
    use_fused_elementwise = os.environ.get("F16_LMA_FUSED_ELEMENTWISE", "1") != "0"
    pad_narrow_wgrad = os.environ.get("F16_LMA_PAD_NARROW_WGRAD", "1") != "0"

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        k, n = weight.shape[1], weight.shape[0]
        if _LinearFn.use_tc and _tc_supported(k, n):
            x2 = x.reshape(-1, k)
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            if x2.data_ptr() % 16 == 0:
                return _linear_tc(x2, weight.contiguous(), bias).view(*x.shape[:-1], n)
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib
        x, weight = ctx.saved_tensors
        dx = dw = db = None
        k, n = weight.shape[1], weight.shape[0]
        dy2 = dy.reshape(-1, n)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if ctx.needs_input_grad[0]:
            if _LinearFn.use_tc and _tc_supported(n, k) and dy2.data_ptr() % 16 == 0:
                dx = _linear_tc(dy2, weight.t().contiguous(), None).view(*dy.shape[:-1], k)
            else:
                dx = dy.matmul(weight)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            x2 = x.reshape(-1, x.shape[-1]).contiguous()
            dw = torch.empty_like(weight)
            db = torch.empty(weight.shape[0], dtype=weight.dtype, device=weight.device) if ctx.has_bias else None
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            tc = (_LinearFn.use_wgrad_tc and _wgrad_tc_supported(k, n) and x2.data_ptr() % 16 == 0 and dy2.data_ptr() % 16 == 0)
            # the 17-feature embedding: rows of 68 bytes cannot be fetched by TMA, and the FP32 slab kernel needs 300 us for this
            # layer's 1.3 M rows; a zero-padded 32-wide copy (one pass, ~45 us) lets the tensor-core kernel take it (~100 us)
            pad = (not tc and _LinearFn.use_wgrad_tc and _LinearFn.pad_narrow_wgrad and k < 32 and x2.shape[0] >= 65536
                   and _wgrad_tc_supported(32, n) and dy2.data_ptr() % 16 == 0)
            if pad:
                x2 = F.pad(x2, (0, 32 - k))
                dw_out = torch.empty((n, 32), dtype=weight.dtype, device=weight.device)
            else:
                dw_out = dw
            fn = "f16_lma_linear_wgrad_tc" if (tc or pad) else "f16_lma_linear_wgrad"
            with _lib.device_guard(x.device):
                _lib.check(getattr(_lib.load(), fn)(x2.shape[0], x2.shape[1], dy2.shape[1], C.c_void_p(x2.data_ptr()),
                                                    C.c_void_p(dy2.data_ptr()), C.c_void_p(dw_out.data_ptr()),
                                                    C.c_void_p(db.data_ptr() if db is not None else 0), stream), fn)
            if pad:
                dw.copy_(dw_out[:, :k])
        return dx, dw, db


class Linear(nn.Linear):
    """nn.Linear (same parameters, same state_dict keys) that, for CUDA float32 tensors with many rows and gradients
    enabled, runs forward, input gradient and weight gradient through the hand-written kernels (_LinearFn)."""

    min_rows = 4096      # below this the library GEMM is as good

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if (x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32 and torch.is_grad_enabled()
                and self.weight.requires_grad and x.numel() // x.shape[-1] >= self.min_rows):
            return _LinearFn.apply(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class _LayerNorm32Fn(torch.autograd.Function):
    """LayerNorm over 32-channel rows by the kernels of csrc/f16_lma_norm.cu (include/f16_lma.h): statistics are
    recomputed in the backward, which also reduces the weight / bias gradients."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float):
        import ctypes as C

        from . import _lib
        x = x.contiguous()
        y = torch.empty_like(x)
        rows = x.numel() // 32
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with _lib.device_guard(x.device):
            _lib.check(_lib.load().f16_lma_layernorm_forward(rows, 32, C.c_void_p(x.data_ptr()), C.c_void_p(weight.data_ptr()),
                                                             C.c_void_p(bias.data_ptr() if bias is not None else 0), float(eps),
                                                             C.c_void_p(y.data_ptr()), stream), "f16_lma_layernorm_forward")
        ctx.save_for_backward(x, weight)
        ctx.meta = (float(eps), bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib
        x, weight = ctx.saved_tensors
        eps, has_bias = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty_like(weight)
        db = torch.empty_like(weight) if has_bias else None
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with _lib.device_guard(x.device):
            _lib.check(_lib.load().f16_lma_layernorm_backward(x.numel() // 32, 32, C.c_void_p(x.data_ptr()), C.c_void_p(weight.data_ptr()),
                                                              C.c_void_p(dy.data_ptr()), eps, C.c_void_p(dx.data_ptr()),
                                                              C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr() if db is not None else 0),
                                                              stream), "f16_lma_layernorm_backward")
        return dx, dw, db, None


def layer_norm32(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last axis. CUDA float32 tensors with 32 channels (the reference's LMA shape) go through the
    hand-written kernels; anything else through torch."""
    if x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 32 and weight.dtype == torch.float32:
        return _LayerNorm32Fn.apply(x, weight, bias, eps)
    return F.layer_norm(x, weight.shape, weight, bias, eps)


class _Norm(nn.Module):
    def __init__(self, dim: int, bias: bool):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim)) if bias else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return layer_norm32(x, self.weight, self.bias, 1e-5)


class _Block(nn.Module):
    def __init__(self, cfg: LMAConfig):
        super().__init__()
        self.ln_1 = _Norm(cfg.d_new, cfg.bias)
        self.attn = _Attention(cfg.d_new, cfg.num_heads_latent, cfg.dropout, cfg.bias)
        self.ln_2 = _Norm(cfg.d_new, cfg.bias)
        self.mlp = _MLP(cfg.d_new, cfg.ff_hidden, cfg.dropout, cfg.bias)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        z = self.attn(self.ln_1(z), residual=z)
        return self.mlp(self.ln_2(z), residual=z)


class _InitialTransform(nn.Module):
    def __init__(self, cfg: LMAConfig):
        super().__init__()
        self.cfg = cfg
        self.input_embedding = Linear(cfg.in_features, cfg.embed_dim, bias=cfg.bias)
        self.embed_layer_2 = Linear(cfg.c_new, cfg.d_new, bias=cfg.bias)
        self.register_buffer("positions", sinusoidal_positions(cfg.seq_len, cfg.embed_dim), persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        b = x.shape[0]
        a = self.input_embedding(x)
        # head stacking (split the 64 channels in 4 heads, lay the heads one after the other along the
        # sequence) and re-chunking into L' tokens of C' values is a single permutation of the 640 values
        h = c.num_heads_stacking
        if (self.training and c.dropout > 0 and a.is_cuda and a.dtype == torch.float32 and (c.embed_dim // h) % 8 == 0 and torch.is_grad_enabled()
                and _LinearFn.use_fused_elementwise):
            y = _EmbedActFn.apply(a, self.positions, c.dropout, h).view(b, c.l_new, c.c_new)      # written in stacked order
        else:
            y = F.dropout(F.relu(a) + self.positions, c.dropout, self.training)
            y = y.view(b, c.seq_len, h, c.embed_dim // h).permute(0, 2, 1, 3).reshape(b, c.l_new, c.c_new)
        return F.relu(self.embed_layer_2(y))


class _LMACore(nn.Module):
    def __init__(self, cfg: LMAConfig):
        super().__init__()
        self.initial_transform = _InitialTransform(cfg)
        self.lma_blocks = nn.ModuleList([_Block(cfg) for _ in range(cfg.num_layers)])

    def forward(self, feats: torch.Tensor) -> torch.Tensor:
        z = self.initial_transform(feats)
        for blk in self.lma_blocks:
            z = blk(z)
        return z.flatten(1)


class LMAExtractor(nn.Module):
    """(B, 10, 15) stacked observations -> (B, 160) features."""

    def __init__(self, cfg: Optional[LMAConfig] = None):
        super().__init__()
        self.cfg = cfg or LMAConfig()
        self.lma_extractor = _LMACore(self.cfg)
        self.features_dim = self.cfg.features_dim

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        assert obs.shape[1:] == (self.cfg.seq_len, NUM_FEATURES), obs.shape
        with torch.no_grad():     # the transform has no parameters and observations carry no gradient
            feats = jsbsim_features(obs) if obs.is_cuda else features17_torch(obs.float())
        return self.lma_extractor(feats)


def _mlp(sizes, act=nn.Tanh) -> nn.Sequential:
    layers = []
    for i in range(len(sizes) - 1):
        layers += [Linear(sizes[i], sizes[i + 1]), act()]
    return nn.Sequential(*layers)


class _MlpExtractor(nn.Module):
    def __init__(self, in_dim: int, pi, vf):
        super().__init__()
        self.policy_net = _mlp([in_dim] + list(pi))
        self.value_net = _mlp([in_dim] + list(vf))


class LMAActorCritic(nn.Module):
    """ActorCriticPolicy of the reference run: shared LMA extractor, pi [64,64] / vf [128,64] tanh MLPs
    (train.py:84), diagonal Gaussian with a state-independent log-std initialised to 0
    (stable_baselines3/common/distributions.py:125-190), orthogonal init with gains sqrt(2) / 0.01 / 1
    (policies.py:580-600)."""

    def __init__(self, cfg: Optional[LMAConfig] = None, pi=(64, 64), vf=(128, 64), action_dim: int = 4, log_std_init: float = 0.0,
                 ortho_init: bool = True):
        super().__init__()
        self.features_extractor = LMAExtractor(cfg)
        d = self.features_extractor.features_dim
        self.mlp_extractor = _MlpExtractor(d, pi, vf)
        self.action_net = Linear(pi[-1], action_dim)
        self.value_net = Linear(vf[-1], 1)
        self.log_std = nn.Parameter(torch.full((action_dim,), float(log_std_init)))
        if ortho_init:
            for mod, gain in ((self.features_extractor, math.sqrt(2)), (self.mlp_extractor, math.sqrt(2)), (self.action_net, 0.01),
                              (self.value_net, 1.0)):
                for m in mod.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.orthogonal_(m.weight, gain=gain)
                        if m.bias is not None:
                            m.bias.data.zero_()

    def _heads(self, obs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        f = self.features_extractor(obs)
        mean = self.action_net(self.mlp_extractor.policy_net(f))
        value = self.value_net(self.mlp_extractor.value_net(f)).squeeze(-1)
        return mean, value

    def _log_prob(self, mean: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        var = torch.exp(2 * self.log_std)
        return (-((actions - mean) ** 2) / (2 * var) - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def forward(self, obs: torch.Tensor, deterministic: bool = False):
        """policy(obs) -> actions, values, log_probs (policies.py:636-658)."""
        mean, value = self._heads(obs)
        actions = mean if deterministic else mean + torch.exp(self.log_std) * torch.randn_like(mean)
        return actions, value, self._log_prob(mean, actions)

    def evaluate_actions(self, obs: torch.Tensor, actions: torch.Tensor):
        """-> values, log_prob, entropy (policies.py:719-745)."""
        mean, value = self._heads(obs)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum(-1).expand(mean.shape[0])
        return value, self._log_prob(mean, actions), entropy

    def predict_values(self, obs: torch.Tensor) -> torch.Tensor:
        f = self.features_extractor(obs)
        return self.value_net(self.mlp_extractor.value_net(f)).squeeze(-1)


class PolicyPacker:
    """The packed parameter buffer f16_lma_policy_forward reads (include/f16_lma.h: layout from f16_lma_policy_entry; Linear
    weights transposed, pair-interleaved where the kernel uses packed FMAs). Works on any device - the layout is plain index
    arithmetic - so that it can be checked without a GPU (the layout tests under tests/). `refresh()` re-packs IN PLACE."""

    @staticmethod
    def shape_supported(policy: "LMAActorCritic") -> bool:
        c = policy.features_extractor.cfg
        pi = [m.out_features for m in policy.mlp_extractor.policy_net if isinstance(m, nn.Linear)]
        vf = [m.out_features for m in policy.mlp_extractor.value_net if isinstance(m, nn.Linear)]
        acts = {type(m) for net in (policy.mlp_extractor.policy_net, policy.mlp_extractor.value_net) for m in net if not isinstance(m, nn.Linear)}
        return ((c.seq_len, c.in_features, c.embed_dim, c.num_heads_stacking, c.num_heads_latent, c.ff_hidden, c.num_layers, c.bias)
                == (10, 17, 64, 4, 4, 128, 2, True) and (c.l_new, c.c_new, c.d_new) == (5, 128, 32) and pi == [64, 64] and vf == [128, 64]
                and acts == {nn.Tanh} and policy.action_net.out_features == 4 and next(policy.parameters()).dtype == torch.float32)

    @staticmethod
    def entries():
        """[(in_features, out_features, weight_offset, bias_offset)] as the library reports them."""
        import ctypes as C

        from . import _lib
        L = _lib.load()
        out = []
        for i in range(L.f16_lma_policy_entries()):
            fin, fout, wo, bo = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
            _lib.check(L.f16_lma_policy_entry(i, C.byref(fin), C.byref(fout), C.byref(wo), C.byref(bo)), "f16_lma_policy_entry")
            out.append((fin.value, fout.value, wo.value, bo.value))
        return out

    def __init__(self, policy: "LMAActorCritic"):
        from . import _lib
        if not PolicyPacker.shape_supported(policy):
            raise ValueError("f16_lma_policy_forward is built for the reference run's policy shape (train.py:21-32,84)")
        self.policy = policy
        self.device = next(policy.parameters()).device
        self.packed = torch.zeros(int(_lib.load().f16_lma_policy_packed_size()), dtype=torch.float32, device=self.device)
        core = policy.features_extractor.lma_extractor
        it = core.initial_transform
        mods = [it.positions, it.input_embedding, it.embed_layer_2]
        for blk in core.lma_blocks:
            mods += [blk.ln_1, blk.attn.c_attn, blk.attn.c_proj, blk.ln_2, blk.mlp.c_fc, blk.mlp.c_proj]
        mods += [policy.mlp_extractor.policy_net[0], policy.mlp_extractor.policy_net[2], policy.action_net,
                 policy.mlp_extractor.value_net[0], policy.mlp_extractor.value_net[2], policy.value_net]
        ents = PolicyPacker.entries()
        assert len(mods) == len(ents)
        self._plan = []                              # (source tensor, destination view, how)
        for i, (m, (fin, fout, wo, bo)) in enumerate(zip(mods, ents)):
            if isinstance(m, torch.Tensor):          # the position table
                assert tuple(m.shape) == (fin, fout) and bo < 0
                self._plan.append((m, self.packed[wo:wo + fin * fout].view(fin, fout), False))
            elif isinstance(m, nn.Linear):
                assert (m.in_features, m.out_features) == (fin, fout) and m.bias is not None, (i, m)
                if fin % 2 == 0 and fout % 4 == 0:      # pair-interleaved for the packed FMAs (include/f16_lma.h)
                    self._plan.append((m.weight, self.packed[wo:wo + fin * fout].view(fin // 2, 2, fout // 4, 2, 2), "pairs"))
                else:
                    self._plan.append((m.weight, self.packed[wo:wo + fin * fout].view(fin, fout), True))
                self._plan.append((m.bias, self.packed[bo:bo + fout], False))
            else:                                    # _Norm
                assert fout == 0 and m.weight.numel() == fin and m.bias is not None, (i, m)
                self._plan.append((m.weight, self.packed[wo:wo + fin], False))
                self._plan.append((m.bias, self.packed[bo:bo + fin], False))
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        for src, dst, how in self._plan:
            if how == "pairs":      # W[n][k] -> [k / 2][(n % 4) / 2][n / 4][n % 2][k % 2]
                k2, _, nc, _, _ = dst.shape
                dst.copy_(src.view(nc, 2, 2, k2, 2).permute(3, 1, 0, 2, 4))
            else:
                dst.copy_(src.t() if how else src)


class PolicyForwardKernel:
    """ActorCriticPolicy.forward for the rollout (policies.py:636-658) as one launch of f16_lma_policy_forward
    (csrc/f16_lma_policy.cu): (N, 10, 15) observations -> actions, values, log_probs, clipped actions.

    The kernel reads the parameters from one packed buffer (PolicyPacker); `refresh()` re-packs it in place from the module -
    call it whenever the optimizer has moved the weights (once per rollout). Buffers are allocated once, so a CUDA graph
    that captured `__call__` stays valid. Only the reference run's shape is built (`supported`)."""

    @staticmethod
    def supported(policy: "LMAActorCritic") -> bool:
        return PolicyPacker.shape_supported(policy) and next(policy.parameters()).is_cuda

    def __init__(self, policy: "LMAActorCritic", act_low: torch.Tensor, act_high: torch.Tensor):
        if not PolicyForwardKernel.supported(policy):
            raise ValueError("f16_lma_policy_forward is built for the reference run's policy shape on a CUDA device (train.py:21-32,84)")
        self.policy = policy
        self._packer = PolicyPacker(policy)
        self.device, self.packed = self._packer.device, self._packer.packed
        self.act_low = act_low.to(self.device, torch.float32).contiguous()
        self.act_high = act_high.to(self.device, torch.float32).contiguous()
        self._out = {}

    def refresh(self) -> None:
        self._packer.refresh()

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, noise: Optional[torch.Tensor] = None, features: bool = False):
        """-> actions (N, 4), values (N,), log_probs (N,), clipped (N, 4) [, features (N, 160)]; noise (N, 4) standard normal
        draws, or None for the deterministic action (the mean). Outputs are reused from call to call for a given N."""
        import ctypes as C

        from . import _lib
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and tuple(obs.shape[1:]) == (10, NUM_FEATURES), obs.shape
        n = obs.shape[0]
        if n not in self._out:
            e = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)      # noqa: E731
            self._out[n] = (e(n, 4), e(n), e(n), e(n, 4), e(n, self.policy.features_extractor.features_dim))
        actions, values, log_probs, clipped, feats = self._out[n]
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (n, 4)
        p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)                  # noqa: E731
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().f16_lma_policy_forward(
                n, p(obs), p(self.packed), self.packed.numel(), p(noise), p(self.policy.log_std), p(self.act_low), p(self.act_high),
                p(actions), p(clipped), p(values), p(log_probs), p(feats if features else None),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "f16_lma_policy_forward")
        return (actions, values, log_probs, clipped) + ((feats,) if features else ())
