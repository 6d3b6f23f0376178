// f16_lma_policy.cu - the rollout's policy forward as ONE kernel (include/f16_lma.h).
//
// What the reference runs once per env-step while it collects a rollout (stable_baselines3/common/on_policy_algorithm.py:
// 203-216 -> ActorCriticPolicy.forward, common/policies.py:636-658): the per-frame feature transform
// (jsbsim_gym/features.py:37-67), the LMA extractor (jsbsim_gym/LMA_features.py:221-279 initial transform, :315-407 two
// blocks of latent attention + MLP), the two tanh MLPs (train.py:84: pi [64,64], vf [128,64]), the 4-wide mean and 1-wide
// value heads, the diagonal-Gaussian sample, its log-probability and the clip to the action box. In torch that is ~50
// launches of 2-5 us kernels on a few thousand rows (library FP32 GEMMs at 32x32 tiles, measured 0.36 ms per step at 4 096
// envs even when replayed as a CUDA graph); here a CTA takes 16 envs through the whole network with every activation in
// shared memory (100 KB: two CTAs per SM) and the weights (58 K floats, transposed once per rollout by the caller so that
// consecutive threads read consecutive output columns) served by L1 / L2.
//
// Dense layers: a thread owns one output column and a strip of rows; per four input features it reads four weights (coalesced)
// and, per row, one 128-bit shared-memory broadcast - RC rows x 4 FMAs per RC + 4 loads. FP32 FMA throughout (the reference
// computes in FP32; sums are re-ordered, nothing is rounded to TF32). LayerNorm: one warp per 32-channel row. Attention: one
// thread per (env, head, query) over the five latent tokens.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/f16_lma.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int E = 16;                    // envs per CTA
constexpr int NT = 128;                  // threads per CTA
constexpr int T = 10, FI = 15, FO = 17;  // frames per observation, floats per frame, features per frame
constexpr int EMB = 64, HS = 4;          // embedding width, heads of the head stacking
constexpr int LT = 5, CN = 128, D = 32;  // latent tokens, their input width, latent width
constexpr int HEADS = 4, DH = 8, FF = 128, BLOCKS = 2;
constexpr int FEAT = LT * D;             // 160
constexpr int PI0 = 64, PI1 = 64, VF0 = 128, VF1 = 64, ACT = 4;

// packed parameter layout (floats). Linear weights are stored transposed: Wt[in][out].
struct Lin { int w, b; };
constexpr int lin_size(int in, int out) { return in * out + out; }
constexpr int O_POS = 0;                                  // positions [T][EMB]
constexpr int O_EMBED = O_POS + T * EMB;                  // 17 -> 64
constexpr int O_EMBED2 = O_EMBED + lin_size(FO, EMB);     // 128 -> 32
constexpr int O_BLOCK0 = O_EMBED2 + lin_size(CN, D);
// per block: ln1 w, ln1 b, c_attn 32 -> 96, c_proj 32 -> 32, ln2 w, ln2 b, c_fc 32 -> 128, mlp c_proj 128 -> 32
constexpr int B_LN1 = 0, B_ATTN = B_LN1 + 2 * D, B_PROJ = B_ATTN + lin_size(D, 3 * D), B_LN2 = B_PROJ + lin_size(D, D),
              B_FC = B_LN2 + 2 * D, B_MPROJ = B_FC + lin_size(D, FF), BLOCK_SIZE = B_MPROJ + lin_size(FF, D);
constexpr int O_PI0 = O_BLOCK0 + BLOCKS * BLOCK_SIZE;
constexpr int O_PI1 = O_PI0 + lin_size(FEAT, PI0);
constexpr int O_ACT = O_PI1 + lin_size(PI0, PI1);
constexpr int O_VF0 = O_ACT + lin_size(PI1, ACT);
constexpr int O_VF1 = O_VF0 + lin_size(FEAT, VF0);
constexpr int O_VAL = O_VF1 + lin_size(VF0, VF1);
constexpr int PACKED = O_VAL + lin_size(VF1, 1);
static_assert(O_EMBED % 4 == 0 && O_EMBED2 % 4 == 0 && O_BLOCK0 % 4 == 0 && BLOCK_SIZE % 4 == 0 && B_ATTN % 4 == 0 && B_PROJ % 4 == 0 && B_FC % 4 == 0 &&
              B_MPROJ % 4 == 0 && O_PI0 % 4 == 0 && O_PI1 % 4 == 0 && O_ACT % 4 == 0 && O_VF0 % 4 == 0 && O_VF1 % 4 == 0,
              "transposed weights are read as float4: 16-byte aligned offsets");

// shared memory (floats). Rows are padded by four floats (strides 132 / 36 / 100 / 68): the four rows a warp's row groups read
// at a time then sit in different banks. The stacked embedding and the MLP's hidden layer share one region, the frames /
// features and the heads' hidden layers live where q | k | v do.
constexpr int PAD = 4;
constexpr int TS = CN + PAD, ZS = D + PAD, QS = 3 * D + PAD;      // 132, 36, 100
constexpr int S_TOK = 0;                       // [E*LT][TS] stacked embedding; later h [E*LT][FF + PAD]
constexpr int S_Z = S_TOK + E * LT * TS;       // [E*LT][ZS] residual stream; env e's 160 features = its 5 rows
constexpr int S_LN = S_Z + E * LT * ZS;        // [E*LT][ZS]
constexpr int S_ATT = S_LN + E * LT * ZS;      // [E*LT][ZS]
constexpr int S_QKV = S_ATT + E * LT * ZS;     // [E*LT][QS]; earlier the frames / features [E*T][FO] + [E*T][FI]; later the heads
constexpr int S_TOTAL = S_QKV + E * LT * QS;
static_assert(FF == CN, "the MLP's hidden layer reuses the stacked embedding's rows");
static_assert(E * T * (FO + FI) <= E * LT * QS, "frames and features must fit the q|k|v region");
static_assert(E * ((PI0 + PAD) + (PI1 + PAD) + (VF0 + PAD) + (VF1 + PAD) + ACT) <= E * LT * QS, "head activations must fit the q|k|v region");

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3 };

template <int A>
__device__ __forceinline__ float activate(float v) {
  if (A == ACT_RELU) return fmaxf(v, 0.0f);
  if (A == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (A == ACT_TANH) return tanhf(v);
  return v;
}

// out(r, n .. n + CC - 1) = act(sum_k in(r, k) * wt[k * N + n ..] + bias[n ..]) for r < ROWS, handed to `store(r, n, values)`.
// A thread owns CC adjacent output columns (N / CC column threads x G = NT / (N / CC) row groups) and accumulates RC rows at a
// time (rows g, g + G, ...): per four input features CC x 4 weights come as four 128-bit loads (a warp's column threads read one
// contiguous span, row groups share it), each row's four inputs as one 128-bit shared-memory load shared by the row group:
// RC + 4 loads per 4 * RC * CC FMAs. in(r, k) = in[r * IS + k], or with SEG > 0 in[r * IS + (k / SEG) * (SEG + PAD) + k % SEG]
// (a row made of padded SEG-wide pieces: the 160 features of an env are its five 32-wide latent rows).
template <int K, int N, int ROWS, int IS, int RC, int CC, int A, int SEG = 0, class Store>
__device__ __forceinline__ void dense(const float* __restrict__ in, const float* __restrict__ wt, const float* __restrict__ bias, Store store) {
  static_assert(N % CC == 0 && (CC == 4 || CC == 1), "column tile");
  constexpr int NC = N / CC;
  constexpr int G = NT / NC;
  static_assert(G >= 1, "at most NT column threads");
  const int nc = threadIdx.x % NC, g = threadIdx.x / NC;
  if (g >= G) return;
  const int n = nc * CC;
  float bn[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) bn[c] = __ldg(bias + n + c);
  for (int r0 = g; r0 < ROWS; r0 += G * RC) {
    float acc[RC][CC];
#pragma unroll
    for (int i = 0; i < RC; ++i)
#pragma unroll
      for (int c = 0; c < CC; ++c) acc[i][c] = 0.0f;
    if constexpr (K % 4 == 0 && IS % 4 == 0) {
#pragma unroll 2
      for (int k = 0; k < K; k += 4) {
        float w[4][CC];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (CC == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(wt + (k + j) * N + n));
            w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
          } else {
            w[j][0] = __ldg(wt + (k + j) * N + n);
          }
        }
        const int ko = SEG > 0 ? (k / SEG) * (SEG + PAD) + k % SEG : k;
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          const int r = r0 + i * G;
          if (r < ROWS) {
            const float4 x = *reinterpret_cast<const float4*>(in + r * IS + ko);
#pragma unroll
            for (int c = 0; c < CC; ++c) {
              acc[i][c] = fmaf(x.x, w[0][c], acc[i][c]);
              acc[i][c] = fmaf(x.y, w[1][c], acc[i][c]);
              acc[i][c] = fmaf(x.z, w[2][c], acc[i][c]);
              acc[i][c] = fmaf(x.w, w[3][c], acc[i][c]);
            }
          }
        }
      }
    } else {
      static_assert(SEG == 0, "segmented rows need K % 4 == 0");
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float w[CC];
        if constexpr (CC == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(wt + k * N + n));
          w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
          w[0] = __ldg(wt + k * N + n);
        }
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          const int r = r0 + i * G;
          if (r < ROWS) {
            const float x = in[r * IS + k];
#pragma unroll
            for (int c = 0; c < CC; ++c) acc[i][c] = fmaf(x, w[c], acc[i][c]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RC; ++i) {
      const int r = r0 + i * G;
      if (r < ROWS) {
        float v[CC];
#pragma unroll
        for (int c = 0; c < CC; ++c) v[c] = activate<A>(acc[i][c] + bn[c]);
        store(r, n, v);
      }
    }
  }
}

__device__ __forceinline__ void put4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void add4(float* p, const float* v) {
  float4 o = *reinterpret_cast<float4*>(p);
  o.x += v[0]; o.y += v[1]; o.z += v[2]; o.w += v[3];
  *reinterpret_cast<float4*>(p) = o;
}

// LayerNorm over rows of D = 32 channels (class LayerNorm, LMA_features.py:172-185: biased variance, eps 1e-5): a warp per row
__device__ __forceinline__ void layernorm_rows(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                               float* __restrict__ y, int rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float wl = __ldg(w + lane), bl = __ldg(b + lane);
  for (int r = warp; r < rows; r += NT / 32) {
    const float v = x[r * ZS + lane];
    float s = v;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    const float d = v - mean;
    float q = d * d;
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    y[r * ZS + lane] = d * rsqrtf(q * (1.0f / D) + 1e-5f) * wl + bl;
  }
}

// softmax(q k^T / sqrt(DH)) v over the LT latent tokens: a thread per (env, head, query)
__device__ __forceinline__ void attention_rows(const float* __restrict__ qkv, float* __restrict__ y, int envs) {
  for (int w = threadIdx.x; w < envs * HEADS * LT; w += NT) {
    const int e = w / (HEADS * LT), h = (w / LT) % HEADS, i = w % LT;
    const float* base = qkv + e * LT * QS + h * DH;
    float q[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] = base[i * QS + d];
    float s[LT], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < LT; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int d = 0; d < DH; ++d) a = fmaf(q[d], base[j * QS + D + d], a);
      s[j] = a * 0.35355339059327376220f;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < LT; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.0f / sum;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.0f;
#pragma unroll
    for (int j = 0; j < LT; ++j) {
      const float p = s[j] * inv;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, base[j * QS + 2 * D + d], o[d]);
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) y[(e * LT + i) * ZS + h * DH + d] = o[d];
  }
}

struct PolicyArgs {
  int64_t n;
  const float* obs;       // [n][T][FI]
  const float* p;         // packed parameters
  const float* noise;     // [n][ACT] standard normal draws, or NULL: actions = mean
  const float* log_std;   // [ACT]
  const float* low;       // [ACT]
  const float* high;      // [ACT]
  float* actions;         // [n][ACT]
  float* clipped;         // [n][ACT] or NULL
  float* values;          // [n]
  float* log_probs;       // [n]
  float* features;        // [n][FEAT] or NULL (tests)
};

__global__ void __launch_bounds__(NT, 2) lma_policy_forward_kernel(PolicyArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* tok = sm + S_TOK;
  float* z = sm + S_Z;
  float* ln = sm + S_LN;
  float* att = sm + S_ATT;
  float* qkv = sm + S_QKV;
  float* feat = qkv;                     // [E*T][FO]
  float* frames = qkv + E * T * FO;      // [E*T][FI]
  const float* P = a.p;

  for (int64_t e0 = (int64_t)blockIdx.x * E; e0 < a.n; e0 += (int64_t)gridDim.x * E) {
    const int envs = (int)((a.n - e0) < E ? (a.n - e0) : E);
    // frames in (coalesced), one thread per frame: the 17 features (jsbsim_gym/features.py:37-67)
    for (int i = threadIdx.x; i < E * T * FI; i += NT) frames[i] = (i < envs * T * FI) ? a.obs[e0 * T * FI + i] : 0.0f;
    __syncthreads();
    for (int f = threadIdx.x; f < E * T; f += NT) {
      const float* o = frames + f * FI;
      const float dx = o[12] - o[0], dy = o[13] - o[1], dz = o[14] - o[2];
      const float distance = sqrtf(dx * dx + dy * dy);
      const float rel = atan2f(dy, dx) - o[11];
      float ca, sa, cb, sb, cp, sp, ct, st, cr, sr;
      sincosf(o[4], &sa, &ca);
      sincosf(o[5], &sb, &cb);
      sincosf(o[9], &sp, &cp);
      sincosf(o[10], &st, &ct);
      sincosf(rel, &sr, &cr);
      float* y = feat + f * FO;
      y[0] = 1.0f / (1.0f + distance * 1e-3f);
      y[1] = dz / 15000.0f;
      y[2] = o[2] / 15000.0f;
      y[3] = o[3];
      y[4] = o[6]; y[5] = o[7]; y[6] = o[8];
      y[7] = ca; y[8] = cb; y[9] = sa; y[10] = sb;
      y[11] = cp; y[12] = ct; y[13] = sp; y[14] = st;
      y[15] = cr; y[16] = sr;
    }
    __syncthreads();
    // embedding + ReLU + positions, written head-stacked: (t, h, c) -> flat h * 160 + t * 16 + c of the env's 640 values,
    // which read as LT tokens of CN values (LMA_features.py:221-279)
    dense<FO, EMB, E * T, FO, 10, 4, ACT_RELU>(feat, P + O_EMBED, P + O_EMBED + FO * EMB, [&](int r, int n, const float* v) {
      const int e = r / T, t = r % T;
      const int j = (n / (EMB / HS)) * (T * (EMB / HS)) + t * (EMB / HS) + (n % (EMB / HS));
      const float4 ps = __ldg(reinterpret_cast<const float4*>(P + O_POS + t * EMB + n));
      const float o[4] = {v[0] + ps.x, v[1] + ps.y, v[2] + ps.z, v[3] + ps.w};
      put4(tok + (e * LT + j / CN) * TS + j % CN, o);
    });
    __syncthreads();
    dense<CN, D, E * LT, TS, 5, 4, ACT_RELU>(tok, P + O_EMBED2, P + O_EMBED2 + CN * D, [&](int r, int n, const float* v) { put4(z + r * ZS + n, v); });
    __syncthreads();
#pragma unroll 1
    for (int blk = 0; blk < BLOCKS; ++blk) {
      const float* B = P + O_BLOCK0 + blk * BLOCK_SIZE;
      layernorm_rows(z, B + B_LN1, B + B_LN1 + D, ln, E * LT);
      __syncthreads();
      dense<D, 3 * D, E * LT, ZS, 8, 4, ACT_NONE>(ln, B + B_ATTN, B + B_ATTN + D * 3 * D, [&](int r, int n, const float* v) { put4(qkv + r * QS + n, v); });
      __syncthreads();
      attention_rows(qkv, att, E);
      __syncthreads();
      dense<D, D, E * LT, ZS, 5, 4, ACT_NONE>(att, B + B_PROJ, B + B_PROJ + D * D, [&](int r, int n, const float* v) { add4(z + r * ZS + n, v); });
      __syncthreads();
      layernorm_rows(z, B + B_LN2, B + B_LN2 + D, ln, E * LT);
      __syncthreads();
      dense<D, FF, E * LT, ZS, 10, 4, ACT_GELU>(ln, B + B_FC, B + B_FC + D * FF, [&](int r, int n, const float* v) { put4(tok + r * TS + n, v); });
      __syncthreads();
      dense<FF, D, E * LT, TS, 5, 4, ACT_NONE>(tok, B + B_MPROJ, B + B_MPROJ + FF * D, [&](int r, int n, const float* v) { add4(z + r * ZS + n, v); });
      __syncthreads();
    }
    if (a.features)
      for (int i = threadIdx.x; i < envs * FEAT; i += NT) a.features[e0 * FEAT + i] = z[(i / D) * ZS + i % D];
    // heads: env e's 160 features are its LT rows of z
    float* p0 = qkv;                           // [E][PI0 + PAD]
    float* p1 = p0 + E * (PI0 + PAD);          // [E][PI1 + PAD]
    float* v0 = p1 + E * (PI1 + PAD);          // [E][VF0 + PAD]
    float* v1 = v0 + E * (VF0 + PAD);          // [E][VF1 + PAD]
    float* mean = v1 + E * (VF1 + PAD);        // [E][ACT]
    dense<FEAT, PI0, E, LT * ZS, 2, 4, ACT_TANH, D>(z, P + O_PI0, P + O_PI0 + FEAT * PI0, [&](int r, int n, const float* v) { put4(p0 + r * (PI0 + PAD) + n, v); });
    dense<FEAT, VF0, E, LT * ZS, 4, 4, ACT_TANH, D>(z, P + O_VF0, P + O_VF0 + FEAT * VF0, [&](int r, int n, const float* v) { put4(v0 + r * (VF0 + PAD) + n, v); });
    __syncthreads();
    dense<PI0, PI1, E, PI0 + PAD, 2, 4, ACT_TANH>(p0, P + O_PI1, P + O_PI1 + PI0 * PI1, [&](int r, int n, const float* v) { put4(p1 + r * (PI1 + PAD) + n, v); });
    dense<VF0, VF1, E, VF0 + PAD, 2, 4, ACT_TANH>(v0, P + O_VF1, P + O_VF1 + VF0 * VF1, [&](int r, int n, const float* v) { put4(v1 + r * (VF1 + PAD) + n, v); });
    __syncthreads();
    dense<PI1, ACT, E, PI1 + PAD, 1, 4, ACT_NONE>(p1, P + O_ACT, P + O_ACT + PI1 * ACT, [&](int r, int n, const float* v) { put4(mean + r * ACT + n, v); });
    dense<VF1, 1, E, VF1 + PAD, 1, 1, ACT_NONE>(v1, P + O_VAL, P + O_VAL + VF1, [&](int r, int, const float* v) {
      if (r < envs) a.values[e0 + r] = v[0];
    });
    __syncthreads();
    // sample, log-probability (common/distributions.py:125-190: Normal(mean, exp(log_std)), summed over the action), clip
    if (threadIdx.x < envs) {
      const int r = threadIdx.x;
      float lp = 0.0f;
#pragma unroll
      for (int j = 0; j < ACT; ++j) {
        const float ls = __ldg(a.log_std + j), m = mean[r * ACT + j];
        const float act = a.noise ? m + expf(ls) * a.noise[(e0 + r) * ACT + j] : m;
        const float d = act - m;
        lp += -(d * d) / (2.0f * expf(2.0f * ls)) - ls - 0.91893853320467274178f;
        a.actions[(e0 + r) * ACT + j] = act;
        if (a.clipped) a.clipped[(e0 + r) * ACT + j] = fmaxf(fminf(act, __ldg(a.high + j)), __ldg(a.low + j));
      }
      a.log_probs[e0 + r] = lp;
    }
    __syncthreads();
  }
}

struct Entry { int in, out, w, b; };
const Entry kEntries[] = {
    {T, EMB, O_POS, -1},                                        // 0: positions (a [T][EMB] table, not transposed)
    {FO, EMB, O_EMBED, O_EMBED + FO * EMB},                     // 1: input_embedding
    {CN, D, O_EMBED2, O_EMBED2 + CN * D},                       // 2: embed_layer_2
#define F16_BLOCK_ENTRIES(o)                                                                                  \
    {D, 0, (o) + B_LN1, (o) + B_LN1 + D},                       /* ln_1: weight, bias */                      \
    {D, 3 * D, (o) + B_ATTN, (o) + B_ATTN + D * 3 * D},         /* attn.c_attn */                             \
    {D, D, (o) + B_PROJ, (o) + B_PROJ + D * D},                 /* attn.c_proj */                             \
    {D, 0, (o) + B_LN2, (o) + B_LN2 + D},                       /* ln_2 */                                    \
    {D, FF, (o) + B_FC, (o) + B_FC + D * FF},                   /* mlp.c_fc */                                \
    {FF, D, (o) + B_MPROJ, (o) + B_MPROJ + FF * D}              /* mlp.c_proj */
    F16_BLOCK_ENTRIES(O_BLOCK0),                                // 3-8
    F16_BLOCK_ENTRIES(O_BLOCK0 + BLOCK_SIZE),                   // 9-14
#undef F16_BLOCK_ENTRIES
    {FEAT, PI0, O_PI0, O_PI0 + FEAT * PI0},                     // 15: policy_net.0
    {PI0, PI1, O_PI1, O_PI1 + PI0 * PI1},                       // 16: policy_net.2
    {PI1, ACT, O_ACT, O_ACT + PI1 * ACT},                       // 17: action_net
    {FEAT, VF0, O_VF0, O_VF0 + FEAT * VF0},                     // 18: value_net.0
    {VF0, VF1, O_VF1, O_VF1 + VF0 * VF1},                       // 19: value_net.2
    {VF1, 1, O_VAL, O_VAL + VF1},                               // 20: value head
};
constexpr int kNumEntries = (int)(sizeof(kEntries) / sizeof(kEntries[0]));
}  // namespace

extern "C" int64_t f16_lma_policy_packed_size(void) { return PACKED; }
extern "C" int f16_lma_policy_entries(void) { return kNumEntries; }

extern "C" int f16_lma_policy_entry(int index, int* in_features, int* out_features, int64_t* weight_offset, int64_t* bias_offset) {
  if (index < 0 || index >= kNumEntries) return f16_internal_fail("f16_lma_policy_entry: index out of range");
  if (in_features) *in_features = kEntries[index].in;
  if (out_features) *out_features = kEntries[index].out;
  if (weight_offset) *weight_offset = kEntries[index].w;
  if (bias_offset) *bias_offset = kEntries[index].b;
  return 0;
}

extern "C" int f16_lma_policy_forward(int64_t n_envs, const float* obs, const float* packed, int64_t packed_len, const float* noise,
                                      const float* log_std, const float* act_low, const float* act_high, float* actions, float* clipped,
                                      float* values, float* log_probs, float* features, void* stream) {
  if (n_envs <= 0) return f16_internal_fail("f16_lma_policy_forward: n_envs must be positive");
  if (packed_len != PACKED) return f16_internal_fail("f16_lma_policy_forward: packed_len does not match f16_lma_policy_packed_size()");
  if (!obs || !packed || !log_std || !actions || !values || !log_probs) return f16_internal_fail("f16_lma_policy_forward: NULL pointer");
  if (clipped && (!act_low || !act_high)) return f16_internal_fail("f16_lma_policy_forward: clipped actions need act_low and act_high");
  const size_t smem = (size_t)S_TOTAL * sizeof(float);
  {   // per device, so set on every call (a host-side attribute, legal during stream capture)
    cudaError_t e = cudaFuncSetAttribute(lma_policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  }
  PolicyArgs a = {n_envs, obs, packed, noise, log_std, act_low, act_high, actions, clipped, values, log_probs, features};
  int64_t tiles = (n_envs + E - 1) / E;
  unsigned grid = (unsigned)(tiles < 148 * 2 ? tiles : 148 * 2);
  lma_policy_forward_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
