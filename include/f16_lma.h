/* f16_lma.h - C ABI of the latent-attention kernels (libf16b200.so), SURVEY.md 8(f) rows 2-3.
 *
 * The attention inside the reference's LMA feature extractor (LatentAttention_RL.forward,
 * jsbsim_gym/LMA_features.py:315-354) runs over L' = 5 latent tokens with 4 heads of 8 channels
 * (train.py:21-32). The reference calls F.scaled_dot_product_attention (:337-342), whose library kernels tile
 * the sequence in blocks of 64: at a sequence length of 5 they spend two thirds of an AM-PPO update and 40 % of
 * a rollout step (measured, tools/profile_amppo_update.py). Here one thread owns one (sample, head): 120 floats
 * of q, k, v in registers, 25 scores, softmax, optional dropout on the probabilities, 40 outputs - read
 * straight from the fused projection's (B, T, 3*H*DH) output and written in (B, T, H*DH) order, so no
 * transposes are materialised. The backward kernel recomputes the probabilities and regenerates the same
 * Philox dropout mask from (seed, sample, head).
 *
 * Device pointers, caller's stream. Supported shape: T = 5, DH = 8 (any H); other shapes are refused. */
#ifndef F16_LMA_H
#define F16_LMA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* qkv: [B][T][3*H*DH] float32 (q | k | v along the last axis) -> y: [B][T][H*DH].
 * dropout_p in [0,1): probability of zeroing an attention weight (0 at inference); kept weights are scaled by
 * 1/(1-p). */
int f16_lma_attention_forward(int64_t batch, int seq_len, int heads, int head_dim, const float* qkv, float* y,
                              float dropout_p, uint64_t seed, void* stream);
/* dy: [B][T][H*DH] -> dqkv: [B][T][3*H*DH] (every element written). Same dropout_p and seed as the forward. */
int f16_lma_attention_backward(int64_t batch, int seq_len, int heads, int head_dim, const float* qkv, const float* dy,
                               float* dqkv, float dropout_p, uint64_t seed, void* stream);
/* The keep mask the two kernels use, as floats (0 or 1/(1-p')): [B][H][T][T]. For tests. */
int f16_lma_attention_mask(int64_t batch, int seq_len, int heads, float* mask, float dropout_p, uint64_t seed, void* stream);

/* LayerNorm over the last axis of the latent tokens (class LayerNorm, jsbsim_gym/LMA_features.py:172-185: weight,
 * optional bias, eps 1e-5; used as ln_1 / ln_2 of every block, :398-400), rows x 32 float32, contiguous. torch's LayerNorm kernels spend a quarter of an AM-PPO
 * update on these 128-byte rows (measured); here a warp normalises 32 rows per pass with coalesced
 * 128-bit accesses and the statistics are recomputed in the backward instead of being stored.
 * backward: dx (every element written), dweight[32] and dbias[32] (or NULL) are overwritten with the sums over rows.
 * Only dim = 32 is built; pointers 16-byte aligned. */
int f16_lma_layernorm_forward(int64_t rows, int dim, const float* x, const float* weight, const float* bias, float eps, float* y,
                              void* stream);
int f16_lma_layernorm_backward(int64_t rows, int dim, const float* x, const float* weight, const float* dy, float eps, float* dx,
                               float* dweight, float* dbias, void* stream);

/* Weight and bias gradients of one of the policy's Linear layers (torch.nn.Linear of the LMA extractor,
 * jsbsim_gym/LMA_features.py:221-279,315-385, and of SB3's MlpExtractor / action_net / value_net,
 * stable_baselines3/common/torch_layers.py:MlpExtractor, common/policies.py:560-600):
 * dweight[out][in] = sum over rows of dy[row][out] * x[row][in], dbias[out] = sum over rows of dy[row][out] (or NULL).
 * x: rows x in_features, dy: rows x out_features, row-major float32; both outputs are overwritten. Any shape; built
 * for in / out <= 160 and 10^5..10^6 rows, where the reduction runs over the batch. */
int f16_lma_linear_wgrad(int64_t rows, int in_features, int out_features, const float* x, const float* dy, float* dweight,
                         float* dbias, void* stream);

/* Forward of the same layers (and, with weight = W^T, their input gradient dx = dy W) on the tensor cores:
 * y[row][out] = sum_in x[row][in] * weight[out][in] + bias[out] (bias may be NULL), replacing torch.nn.functional.linear
 * as called by the reference's Linear modules (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 MlpExtractor,
 * stable_baselines3/common/torch_layers.py). tcgen05.mma kind::tf32 with both operands split into a TF32 head and an
 * FP32 remainder (three MMAs per k-step), FP32 accumulation in tensor memory: FP32-accurate like a reordered FP32 sum.
 * Row-major float32, contiguous; x, weight, y and bias 16-byte aligned. Shapes: out_features a multiple of 32;
 * in_features <= 32, or a multiple of 32 - f16_lma_linear_supported says whether a shape is built (the caller keeps the
 * library GEMM for the others: the 4- and 1-wide output heads). Layers whose W, in both parts, does not fit shared memory
 * next to the rings (160 -> 128, 128 -> 160) are computed in up to four groups of output columns, one launch each. */
int f16_lma_linear_supported(int in_features, int out_features);
int f16_lma_linear_forward(int64_t rows, int in_features, int out_features, const float* x, const float* weight,
                           const float* bias, float* y, void* stream);

/* The same weight / bias gradients on the tensor cores (csrc/f16_lma_wgrad_tc.cu): D[out][in] accumulates in tensor memory
 * over 16-row chunks of dy and x (tcgen05.mma kind::tf32 on MN-major operands, TF32 head / remainder split: three MMAs
 * per 8 rows), is read out and restarted every 64 rows so that the tensor core's truncating accumulation never runs
 * long, and the CTAs' partial results meet in float atomics on the zero-initialised outputs. Same arguments and
 * semantics as f16_lma_linear_wgrad; x and dy 16-byte aligned. Built for out_features in {32, 64, 96, 128} and
 * in_features a multiple of 32 up to 160 whose partial sums fit shared memory (f16_lma_linear_wgrad_tc_supported); the
 * caller keeps f16_lma_linear_wgrad for the rest (17 input features, the 4- and 1-wide heads, 160 -> 128). */
int f16_lma_linear_wgrad_tc_supported(int in_features, int out_features);
int f16_lma_linear_wgrad_tc(int64_t rows, int in_features, int out_features, const float* x, const float* dy, float* dweight,
                            float* dbias, void* stream);

/* The two fused elementwise steps of the extractor's training pass (csrc/f16_lma_elementwise.cu). The keep mask is a function
 * of (seed, element index): the backward calls regenerate it from the same dropout_p and seed instead of reading a stored
 * mask; kept elements are scaled by 1 / (1 - p'), p' = round(p * 65536) / 65536. float32, contiguous, 16-byte aligned.
 *   embed activation (_InitialTransform: ReLU, sinusoidal positions, embedding dropout - jsbsim_gym/LMA_features.py:221-279):
 *     y[r][c] = keep * (max(a[r][c], 0) + pos[r mod seq_len][c]),   da[r][c] = keep * dy[r][c] * (a[r][c] > 0);  channels % 8 == 0
 *     With stack_heads = H > 1, y and dy are in the head-stacked order the extractor re-chunks into latent tokens (:255-270):
 *     element (b, t, h * channels/H + c) at b * seq_len * channels + h * seq_len * channels/H + t * channels/H + c - the
 *     view / permute / reshape copy is not materialised, forward or backward. a, da and the keep mask keep the natural order.
 *   residual dropout (LMA block: z + drop(attn(ln(z))), z + drop(mlp(ln(z))) - :386-407):
 *     y = z + keep * x,   dx = keep * dy (dz = dy is the caller's);  n % 8 == 0 */
int f16_lma_embed_act_forward(int64_t rows, int channels, int seq_len, int stack_heads, const float* a, const float* pos, float* y,
                              float dropout_p, uint64_t seed, void* stream);
int f16_lma_embed_act_backward(int64_t rows, int channels, int seq_len, int stack_heads, const float* a, const float* dy, float* da,
                               float dropout_p, uint64_t seed, void* stream);
int f16_lma_dropout_add_forward(int64_t n, const float* x, const float* z, float* y, float dropout_p, uint64_t seed, void* stream);
int f16_lma_dropout_backward(int64_t n, const float* dy, float* dx, float dropout_p, uint64_t seed, void* stream);

/* The rollout's policy forward in one kernel (csrc/f16_lma_policy.cu): what ActorCriticPolicy.forward computes once per env-step
 * while a rollout is collected (stable_baselines3/common/on_policy_algorithm.py:203-216 -> common/policies.py:636-658) for the
 * reference run's policy (train.py:21-32,84): feature transform (jsbsim_gym/features.py:37-67), LMA extractor
 * (jsbsim_gym/LMA_features.py:221-279,315-407; inference: dropout is the identity), pi [64,64] / vf [128,64] tanh MLPs, mean and
 * value heads, actions = mean + exp(log_std) * noise (noise NULL: actions = mean), log-probability of the diagonal Gaussian
 * (common/distributions.py:125-190), and the clip to [act_low, act_high] (on_policy_algorithm.py:216; clipped may be NULL).
 *   obs [n][10][15] -> actions [n][4], clipped [n][4], values [n], log_probs [n], features [n][160] (NULL unless wanted)
 * The parameters come as one packed float32 buffer of f16_lma_policy_packed_size() floats; f16_lma_policy_entry(i) gives, for entry
 * i of f16_lma_policy_entries(), the shape and the offsets at which the caller stores it: Linear layers as the TRANSPOSED weight
 * followed by the bias [out]; LayerNorms (out_features 0) as weight [in] and bias [in]; entry 0 is the sinusoidal position
 * table [10][64] (bias_offset -1). Transposed weight of a layer with an even number of inputs and a multiple of four outputs:
 * PAIR-INTERLEAVED for the packed FP32 FMAs - W[out = n][in = k] at weight_offset +
 *     (k / 2) * 2 * out_features + ((n % 4) / 2) * out_features + (n / 4) * 4 + (n % 2) * 2 + k % 2;
 * the other two (17 -> 64 embedding, 64 -> 1 value head): plain [in][out], W[n][k] at weight_offset + k * out_features + n. Order: positions, input_embedding, embed_layer_2, per block {ln_1, attn.c_attn, attn.c_proj, ln_2,
 * mlp.c_fc, mlp.c_proj} x 2, policy_net.0, policy_net.2, action_net, value_net.0, value_net.2, value head.
 * Only this shape is built (LMAConfigRL defaults of train.py); FP32 FMA arithmetic, sums re-ordered against torch's. */
int64_t f16_lma_policy_packed_size(void);
int f16_lma_policy_entries(void);
int f16_lma_policy_entry(int index, int* in_features, int* out_features, int64_t* weight_offset, int64_t* bias_offset);
int f16_lma_policy_forward(int64_t n_envs, const float* obs, const float* packed, int64_t packed_len, const float* noise,
                           const float* log_std, const float* act_low, const float* act_high, float* actions, float* clipped,
                           float* values, float* log_probs, float* features, void* stream);
#ifdef __cplusplus
}
#endif
#endif
