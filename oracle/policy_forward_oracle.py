"""NumPy restatement (float64) of what f16_lma_policy_forward computes (TEST INFRASTRUCTURE ONLY).

The checker of the rollout's fused policy forward (include/f16_lma.h, csrc/f16_lma_policy.cu): it reads the SAME packed
parameter buffer the kernel reads - by the layout the header documents and nothing else (plain index arithmetic below, no code
shared with f16_jsb_b200/lma.py's packer) - and restates the reference's forward pass:

  JSBSimFeatureExtractor.forward                jsbsim_gym/features.py:37-67
  _InitialTransformRL (embedding, ReLU, sinusoidal positions, head stacking, re-chunking, second embedding + ReLU)
                                                jsbsim_gym/LMA_features.py:221-279
  LMA block x 2 (LayerNorm eps 1e-5 biased variance :172-185, latent attention over 5 tokens with 4 heads of 8 :315-354,
  MLP with exact GELU :357-385, residuals :386-407; dropout is the identity at inference)
  MlpExtractor pi [64,64] / vf [128,64] with tanh, action_net, value_net   train.py:84; stable_baselines3/common/policies.py:560-658
  DiagGaussianDistribution: actions = mean + exp(log_std) * noise, log_prob summed over the action
                                                stable_baselines3/common/distributions.py:125-190
  clip to the action box                        stable_baselines3/common/on_policy_algorithm.py:216

Pinned two ways in tests/test_policy_oracle.py (CPU): against f16_jsb_b200.lma.LMAActorCritic - itself pinned to the reference's
own modules' recorded outputs (tests/golden/learner_golden.pt, tests/test_learner.py) - on random parameters packed by the
product's packer, and directly against those recorded reference features with the reference's recorded weights.
Only tests/ may import this file.
"""
import math

import numpy as np

T, FI, FO, EMB, HS, LT, CN, D, HEADS, DH, FF = 10, 15, 17, 64, 4, 5, 128, 32, 4, 8, 128
_erf = np.vectorize(math.erf, otypes=[np.float64])


def unpack(packed, entries):
    """-> list of (W [out][in] or table, bias or None) per entry of f16_lma_policy_entry, by the header's layout rules."""
    p = np.asarray(packed, dtype=np.float64)
    out = []
    for i, (fin, fout, wo, bo) in enumerate(entries):
        if bo < 0:                                   # entry 0: the position table [10][64]
            out.append((p[wo:wo + fin * fout].reshape(fin, fout).copy(), None))
        elif fout == 0:                              # LayerNorm: weight [in], bias [in]
            out.append((p[wo:wo + fin].copy(), p[bo:bo + fin].copy()))
        else:
            w = np.empty((fout, fin))
            n, k = np.meshgrid(np.arange(fout), np.arange(fin), indexing="ij")
            if fin % 2 == 0 and fout % 4 == 0:       # pair-interleaved
                off = (k // 2) * 2 * fout + ((n % 4) // 2) * fout + (n // 4) * 4 + (n % 2) * 2 + k % 2
            else:                                    # plain transposed [in][out]
                off = k * fout + n
            assert sorted(off.ravel().tolist()) == list(range(fin * fout)), "the layout must be a permutation"
            w[n, k] = p[wo + off]
            out.append((w, p[bo:bo + fout].copy()))
    return out


def features17(frames):
    """(..., 15) -> (..., 17): jsbsim_gym/features.py:37-67."""
    f = np.asarray(frames, dtype=np.float64)
    pos, mach, ab, rates, pt, psi, goal = f[..., 0:3], f[..., 3:4], f[..., 4:6], f[..., 6:9], f[..., 9:11], f[..., 11:12], f[..., 12:15]
    disp = goal - pos
    dist = np.sqrt((disp[..., :2] ** 2).sum(-1, keepdims=True))
    rel = np.arctan2(disp[..., 1:2], disp[..., 0:1]) - psi
    return np.concatenate([1 / (1 + dist * 1e-3), disp[..., 2:3] / 15000, pos[..., 2:3] / 15000, mach, rates, np.cos(ab), np.sin(ab),
                           np.cos(pt), np.sin(pt), np.cos(rel), np.sin(rel)], axis=-1)


def _linear(x, wb):
    return x @ wb[0].T + wb[1]


def _layernorm(x, wb):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5) * wb[0] + wb[1]


def _attention(qkv):
    b = qkv.shape[0]
    q, k, v = [qkv[..., i * D:(i + 1) * D].reshape(b, LT, HEADS, DH).transpose(0, 2, 1, 3) for i in range(3)]
    s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(DH)
    s = np.exp(s - s.max(-1, keepdims=True))
    return ((s / s.sum(-1, keepdims=True)) @ v).transpose(0, 2, 1, 3).reshape(b, LT, D)


def forward(obs, packed, entries, log_std, noise=None, low=None, high=None):
    """obs (N, 10, 15) -> dict(actions, clipped, values, log_probs, features, mean), float64."""
    P = unpack(packed, entries)
    x = features17(obs)                                                        # (N, 10, 17)
    b = x.shape[0]
    y = np.maximum(_linear(x, P[1]), 0.0) + P[0][0]                            # ReLU, then the positions
    y = y.reshape(b, T, HS, EMB // HS).transpose(0, 2, 1, 3).reshape(b, LT, CN)   # head stacking, re-chunking
    z = np.maximum(_linear(y, P[2]), 0.0)
    for blk in range(2):
        ln1, c_attn, c_proj, ln2, c_fc, m_proj = P[3 + 6 * blk: 9 + 6 * blk]
        z = z + _linear(_attention(_linear(_layernorm(z, ln1), c_attn)), c_proj)
        h = _linear(_layernorm(z, ln2), c_fc)
        z = z + _linear(0.5 * h * (1.0 + _erf(h / math.sqrt(2.0))), m_proj)
    feats = z.reshape(b, LT * D)
    mean = _linear(np.tanh(_linear(np.tanh(_linear(feats, P[15])), P[16])), P[17])
    values = _linear(np.tanh(_linear(np.tanh(_linear(feats, P[18])), P[19])), P[20])[:, 0]
    ls = np.asarray(log_std, dtype=np.float64)
    actions = mean if noise is None else mean + np.exp(ls) * np.asarray(noise, dtype=np.float64)
    log_probs = (-((actions - mean) ** 2) / (2 * np.exp(2 * ls)) - ls - 0.5 * math.log(2 * math.pi)).sum(-1)
    clipped = actions if low is None else np.maximum(np.minimum(actions, np.asarray(high, dtype=np.float64)), np.asarray(low, dtype=np.float64))
    return {"actions": actions, "clipped": clipped, "values": values, "log_probs": log_probs, "features": feats, "mean": mean}
