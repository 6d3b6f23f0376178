// f16_lma_policy.cu - the rollout's policy forward as ONE kernel (include/f16_lma.h).
//
// What the reference runs once per env-step while it collects a rollout (stable_baselines3/common/on_policy_algorithm.py:
// 203-216 -> ActorCriticPolicy.forward, common/policies.py:636-658): the per-frame feature transform
// (jsbsim_gym/features.py:37-67), the LMA extractor (jsbsim_gym/LMA_features.py:221-279 initial transform, :315-407 two
// blocks of latent attention + MLP), the two tanh MLPs (train.py:84: pi [64,64], vf [128,64]), the 4-wide mean and 1-wide
// value heads, the diagonal-Gaussian sample, its log-probability and the clip to the action box. In torch that is ~50
// launches of 2-5 us kernels on a few thousand rows (library FP32 GEMMs at 32x32 tiles, measured 0.29 ms per step at 4 096
// envs even when replayed as a CUDA graph); here a CTA of 128 threads takes 16 envs through the whole network with every
// activation in shared memory (87 KB + a 24 KB weight stage: two CTAs per SM). The weights (75 K floats, transposed once per
// rollout by the caller so that consecutive threads read consecutive output columns) are staged layer by layer: the next
// layer's travel from L2 in registers while the current layer's FMAs run, and are stored to shared memory between two
// barriers (reading them straight from global memory left every k-step waiting on L2: 2.6 long-scoreboard stalls per issued
// instruction at 7 warps per SM, profiles/).
//
// Dense layers: a thread owns four adjacent output columns and up to ten rows; per four input features it reads 4 x 4 weights
// and, per row, four inputs as 128-bit shared-memory loads and issues 8 packed FMAs per row (fma.rn.f32x2 on the partial sums
// over even and odd input features; the weights are staged pair-interleaved so that no operand needs packing). FP32 throughout
// (the reference computes in FP32; sums are re-ordered, nothing is rounded to TF32). LayerNorm: one warp per 32-channel row.
// Attention: one thread per (env, head, query) over the five latent tokens.
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
// at a time then sit in different banks. Regions are reused as their contents die: the stacked embedding, q | k | v and the
// MLP's hidden layer share R0; LayerNorm output and attention output share LA; frames / features and the heads' hidden layers
// share MISC. WS holds the weights of the layer (or K-chunk of a layer) being computed.
constexpr int PAD = 4;
constexpr int TS = CN + PAD, ZS = D + PAD, QS = 3 * D + PAD;      // 132, 36, 100
constexpr int HEAD_FLOATS = E * ((PI0 + PAD) + (PI1 + PAD) + (VF0 + PAD) + (VF1 + PAD) + ACT);
constexpr int MISC_FLOATS = HEAD_FLOATS > E * T * (FO + FI) ? HEAD_FLOATS : E * T * (FO + FI);
constexpr int WS_FLOATS = 6144;                // 24 KB: the largest staged piece is 5 120 floats
constexpr int WS_PER_THREAD = WS_FLOATS / 4 / NT;      // float4 registers a thread carries for the next piece
constexpr int S_R0 = 0;                        // [E*LT][TS]
constexpr int S_Z = S_R0 + E * LT * TS;        // [E*LT][ZS] residual stream; env e's 160 features = its 5 rows
constexpr int S_LA = S_Z + E * LT * ZS;        // [E*LT][ZS]
constexpr int S_MISC = S_LA + E * LT * ZS;
constexpr int S_WS = S_MISC + MISC_FLOATS;
constexpr int S_TOTAL = S_WS + WS_FLOATS;
static_assert(FF == CN, "the MLP's hidden layer reuses the stacked embedding's rows");
static_assert(QS <= TS, "q | k | v rows fit the stacked embedding's rows");
static_assert(S_Z % 4 == 0 && S_LA % 4 == 0 && S_MISC % 4 == 0 && S_WS % 4 == 0 && WS_FLOATS % (4 * NT) == 0, "128-bit accesses");
static_assert(2 * (S_TOTAL * 4 + 1024) <= 227 * 1024, "two CTAs per SM");

#ifndef F16_POL_RC_FC
#define F16_POL_RC_FC 5         // rows per pass of the 32 -> 128 layer (10: fewer weight reads, more registers in flight)
#endif

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3 };

template <int A>
__device__ __forceinline__ float activate(float v) {
  if (A == ACT_RELU) return fmaxf(v, 0.0f);
  if (A == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (A == ACT_TANH) return tanhf(v);
  return v;
}

// The weights of the next piece travel through registers: loaded from global memory (L2) before the current piece is computed,
// stored to WS after it - their latency hides under the FMAs instead of stalling every k-step of a thinly occupied SM.
struct Prefetch {
  float4 r[WS_PER_THREAD];
  __device__ __forceinline__ void load(const float* __restrict__ src, int floats) {
#pragma unroll
    for (int i = 0; i < WS_PER_THREAD; ++i) {
      const int j = threadIdx.x + i * NT;
      if (4 * j < floats) r[i] = __ldg(reinterpret_cast<const float4*>(src) + j);
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ ws, int floats) const {
#pragma unroll
    for (int i = 0; i < WS_PER_THREAD; ++i) {
      const int j = threadIdx.x + i * NT;
      if (4 * j < floats) reinterpret_cast<float4*>(ws)[j] = r[i];
    }
  }
};

// Packed FP32: fma.rn.f32x2 does two FMAs per issued instruction (d.lo = a.lo * b.lo + c.lo, d.hi likewise).
using u64 = unsigned long long;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum2(u64 v) { return __uint_as_float((unsigned)v) + __uint_as_float((unsigned)(v >> 32)); }

// One piece of a dense layer on packed FMAs: for the RC rows r0, r0 + G, ... of this thread and its four adjacent output columns
// n .. n + 3, acc(r, c) += (sum over even k, sum over odd k) of in(r, k) * W(k, n + c), k = k0 .. k0 + KC - 1 - the two halves of a
// 64-bit accumulator are the partial sums over the even and the odd input features (added at the end, sum2). The inputs' pairs
// (k, k + 1) are the halves of a 128-bit shared-memory load as they lie; the weights are staged PAIR-INTERLEAVED so that theirs
// are too: element (k, n) of a layer with N outputs sits at
//     (k / 2) * 2N + ((n % 4) / 2) * N + (n / 4) * 4 + (n % 2) * 2 + k % 2
// i.e. per pair of input features first every column thread's (n, n + 1) x (k, k + 1) quad, then every thread's (n + 2, n + 3)
// quad: a thread's four 128-bit loads per four input features are contiguous across the column threads (no bank conflicts,
// row groups share them). Per four input features: 4 + RC loads, 8 RC packed FMAs = 16 RC FMAs, no packing instructions.
// The k-steps are software-pipelined two deep (a scheduler holds two warps here: other warps do not hide shared-memory latency).
// in(r, k) = in[r * IS + k], or with SEG > 0 in[r * IS + (k / SEG) * (SEG + PAD) + k % SEG] (a row made of padded SEG-wide
// pieces: the 160 features of an env are its five 32-wide latent rows).
template <int KC, int N, int ROWS, int IS, int RC, int SEG>
__device__ __forceinline__ void dense_piece2(const float* __restrict__ in, const float* __restrict__ ws, int k0, int r0, int n, u64 (&acc)[RC][4]) {
  static_assert(KC % 4 == 0 && IS % 4 == 0 && N % 4 == 0, "128-bit loads");
  constexpr int G = NT / (N / 4);
  constexpr bool FULL = ROWS % G == 0 && (ROWS / G) % RC == 0;      // every (thread, pass, i) is a row: no predicates
  const float* wcol = ws + n;                                        // (n / 4) * 4
  ulonglong2 xa[RC], xb[RC], wa[4], wb[4];
  auto fetch = [&](int k, ulonglong2 (&x)[RC], ulonglong2 (&w)[4]) {
    const float* wp = wcol + (k >> 1) * (2 * N);
    w[0] = *reinterpret_cast<const ulonglong2*>(wp);                 // (k, k+1) x (n, n+1)
    w[1] = *reinterpret_cast<const ulonglong2*>(wp + N);             // (k, k+1) x (n+2, n+3)
    w[2] = *reinterpret_cast<const ulonglong2*>(wp + 2 * N);         // (k+2, k+3) x (n, n+1)
    w[3] = *reinterpret_cast<const ulonglong2*>(wp + 3 * N);         // (k+2, k+3) x (n+2, n+3)
    const int kg = k0 + k;
    const int ko = SEG > 0 ? (kg / SEG) * (SEG + PAD) + kg % SEG : kg;
#pragma unroll
    for (int i = 0; i < RC; ++i) {
      const int r = r0 + i * G;
      if (FULL || r < ROWS) x[i] = *reinterpret_cast<const ulonglong2*>(in + r * IS + ko);
      else x[i] = make_ulonglong2(0ull, 0ull);
    }
  };
  auto fmas = [&](const ulonglong2 (&x)[RC], const ulonglong2 (&w)[4]) {
#pragma unroll
    for (int i = 0; i < RC; ++i) {
      acc[i][0] = fma2(x[i].x, w[0].x, acc[i][0]);
      acc[i][1] = fma2(x[i].x, w[0].y, acc[i][1]);
      acc[i][2] = fma2(x[i].x, w[1].x, acc[i][2]);
      acc[i][3] = fma2(x[i].x, w[1].y, acc[i][3]);
      acc[i][0] = fma2(x[i].y, w[2].x, acc[i][0]);
      acc[i][1] = fma2(x[i].y, w[2].y, acc[i][1]);
      acc[i][2] = fma2(x[i].y, w[3].x, acc[i][2]);
      acc[i][3] = fma2(x[i].y, w[3].y, acc[i][3]);
    }
  };
  fetch(0, xa, wa);
#pragma unroll 1
  for (int k = 0; k < KC; k += 8) {
    if (k + 4 < KC) fetch(k + 4, xb, wb);
    fmas(xa, wa);
    if (k + 8 < KC) fetch(k + 8, xa, wa);
    if (k + 4 < KC) fmas(xb, wb);
  }
}

// A whole layer whose weights are in WS: out(r, n ..) = act(acc + bias) handed to `store(r, n, values)`. Layers with an even
// number of inputs and a multiple of four outputs run on packed FMAs (pair-interleaved weights, dense_piece2); the 17-feature
// embedding and the 1-wide value head keep plain [K][N] weights and scalar FMAs.
template <int K, int N, int ROWS, int IS, int RC, int CC, int A, class Store>
__device__ __forceinline__ void dense(const float* __restrict__ in, const float* __restrict__ ws, const float* __restrict__ bias, Store store) {
  static_assert(N % CC == 0 && (CC == 4 || CC == 1), "column tile");
  constexpr int NC = N / CC, G = NT / NC;
  static_assert(G >= 1, "at most NT column threads");
  const int nc = threadIdx.x % NC, g = threadIdx.x / NC, n = nc * CC;
  if (g >= G) return;
  float bn[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) bn[c] = __ldg(bias + n + c);
  for (int r0 = g; r0 < ROWS; r0 += G * RC) {
    float out[RC][CC];
    if constexpr (CC == 4 && K % 4 == 0 && IS % 4 == 0) {
      u64 acc[RC][4];
#pragma unroll
      for (int i = 0; i < RC; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0ull;
      dense_piece2<K, N, ROWS, IS, RC, 0>(in, ws, 0, r0, n, acc);
#pragma unroll
      for (int i = 0; i < RC; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[i][c] = sum2(acc[i][c]);
    } else {
#pragma unroll
      for (int i = 0; i < RC; ++i)
#pragma unroll
        for (int c = 0; c < CC; ++c) out[i][c] = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float w[CC];
        if constexpr (CC == 4) {
          const float4 v = *reinterpret_cast<const float4*>(ws + k * N + n);
          w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
          w[0] = ws[k * N + n];
        }
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          const int r = r0 + i * G;
          if (r < ROWS) {
            const float x = in[r * IS + k];
#pragma unroll
            for (int c = 0; c < CC; ++c) out[i][c] = fmaf(x, w[c], out[i][c]);
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
        for (int c = 0; c < CC; ++c) v[c] = activate<A>(out[i][c] + bn[c]);
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

// LayerNorm over rows of D = 32 channels (class LayerNorm, LMA_features.py:172-185: biased variance, eps 1e-5): eight lanes per
// row (four channels each), four rows per warp at a time - three shuffles per statistic instead of five, and four rows'
// dependent chains in flight (a warp per row left the scheduler waiting on ten serial shuffles per row)
__device__ __forceinline__ void layernorm_rows(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                               float* __restrict__ y, int rows) {
  static_assert(D == 32 && (E * LT) % (NT / 32 * 4) == 0, "eight lanes x four channels per row; whole warps of rows");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 3, c4 = (lane & 7) * 4;
  const float4 wl = __ldg(reinterpret_cast<const float4*>(w + c4)), bl = __ldg(reinterpret_cast<const float4*>(b + c4));
  for (int r = warp * 4 + sub; r < rows; r += NT / 32 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(x + r * ZS + c4);
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 4; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    float q = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
    for (int o = 4; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * (1.0f / D) + 1e-5f);
    *reinterpret_cast<float4*>(y + r * ZS + c4) = make_float4(d0 * rs * wl.x + bl.x, d1 * rs * wl.y + bl.y, d2 * rs * wl.z + bl.z, d3 * rs * wl.w + bl.w);
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

// the 17 features of one frame (jsbsim_gym/features.py:37-67)
__device__ __forceinline__ void frame_features(const float* __restrict__ o, float* __restrict__ y) {
  const float dx = o[12] - o[0], dy = o[13] - o[1], dz = o[14] - o[2];
  const float distance = sqrtf(dx * dx + dy * dy);
  const float rel = atan2f(dy, dx) - o[11];
  float ca, sa, cb, sb, cp, sp, ct, st, cr, sr;
  sincosf(o[4], &sa, &ca);
  sincosf(o[5], &sb, &cb);
  sincosf(o[9], &sp, &cp);
  sincosf(o[10], &st, &ct);
  sincosf(rel, &sr, &cr);
  y[0] = 1.0f / (1.0f + distance * 1e-3f);
  y[1] = dz / 15000.0f;
  y[2] = o[2] / 15000.0f;
  y[3] = o[3];
  y[4] = o[6]; y[5] = o[7]; y[6] = o[8];
  y[7] = ca; y[8] = cb; y[9] = sa; y[10] = sb;
  y[11] = cp; y[12] = ct; y[13] = sp; y[14] = st;
  y[15] = cr; y[16] = sr;
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

// K-chunks of the two wide head layers (their weights do not fit WS in one piece)
constexpr int PI0_KC = 80, VF0_KC = 40, VF1_KC = 64;
static_assert(PI0_KC % 4 == 0 && VF0_KC % 4 == 0 && VF1_KC % 4 == 0, "chunks are whole k-steps");
static_assert(PI0_KC * PI0 <= WS_FLOATS && VF0_KC * VF0 <= WS_FLOATS && VF1_KC * VF1 <= WS_FLOATS && CN * D <= WS_FLOATS && D * FF <= WS_FLOATS &&
              FF * D <= WS_FLOATS && D * 3 * D <= WS_FLOATS && FO * EMB <= WS_FLOATS && PI0 * PI1 <= WS_FLOATS, "every staged piece fits WS");

__global__ void __launch_bounds__(NT, 2) lma_policy_forward_kernel(PolicyArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* r0buf = sm + S_R0;              // stacked embedding [E*LT][TS] -> q | k | v [E*LT][QS] -> hidden [E*LT][TS]
  float* z = sm + S_Z;
  float* la = sm + S_LA;
  float* misc = sm + S_MISC;
  float* ws = sm + S_WS;
  float* feat = misc;                    // [E*T][FO]
  float* frames = misc + E * T * FO;     // [E*T][FI]
  float* p0 = misc;                      // [E][PI0 + PAD]
  float* p1 = p0 + E * (PI0 + PAD);      // [E][PI1 + PAD]
  float* v0 = p1 + E * (PI1 + PAD);      // [E][VF0 + PAD]
  float* v1 = v0 + E * (VF0 + PAD);      // [E][VF1 + PAD]
  float* mean = v1 + E * (VF1 + PAD);    // [E][ACT]
  const float* P = a.p;
  Prefetch pf;

  // A stage: the previous stage's FMAs are done with WS -> `pre` work that only touches activations, the prefetched weights go
  // to WS, the weights after them start their trip -> barrier -> the layer's FMAs. Two barriers per stage.
#define F16_STAGE(pre, cur_floats, next_src, next_floats, compute) \
  __syncthreads();                                                 \
  pre;                                                             \
  pf.store(ws, (cur_floats));                                      \
  pf.load((next_src), (next_floats));                              \
  __syncthreads();                                                 \
  compute;

  pf.load(P + O_EMBED, FO * EMB);
  for (int64_t e0 = (int64_t)blockIdx.x * E; e0 < a.n; e0 += (int64_t)gridDim.x * E) {
    const int envs = (int)((a.n - e0) < E ? (a.n - e0) : E);
    for (int i = threadIdx.x; i < E * T * FI; i += NT) frames[i] = (i < envs * T * FI) ? a.obs[e0 * T * FI + i] : 0.0f;
    // embedding + ReLU + positions, written head-stacked: (t, h, c) -> flat h * 160 + t * 16 + c of the env's 640 values,
    // which read as LT tokens of CN values (LMA_features.py:221-279)
    F16_STAGE(for (int f = threadIdx.x; f < E * T; f += NT) frame_features(frames + f * FI, feat + f * FO),
              FO * EMB, P + O_EMBED2, CN * D,
              (dense<FO, EMB, E * T, FO, 10, 4, ACT_RELU>(feat, ws, P + O_EMBED + FO * EMB, [&](int r, int n, const float* v) {
                const int e = r / T, t = r % T;
                const int j = (n / (EMB / HS)) * (T * (EMB / HS)) + t * (EMB / HS) + (n % (EMB / HS));
                const float4 ps = __ldg(reinterpret_cast<const float4*>(P + O_POS + t * EMB + n));
                const float o[4] = {v[0] + ps.x, v[1] + ps.y, v[2] + ps.z, v[3] + ps.w};
                put4(r0buf + (e * LT + j / CN) * TS + j % CN, o);
              })))
    F16_STAGE(, CN * D, P + O_BLOCK0 + B_ATTN, D * 3 * D,
              (dense<CN, D, E * LT, TS, 5, 4, ACT_RELU>(r0buf, ws, P + O_EMBED2 + CN * D, [&](int r, int n, const float* v) { put4(z + r * ZS + n, v); })))
#pragma unroll 1
    for (int blk = 0; blk < BLOCKS; ++blk) {
      const float* B = P + O_BLOCK0 + blk * BLOCK_SIZE;
      const float* after = blk + 1 < BLOCKS ? B + BLOCK_SIZE + B_ATTN : P + O_PI0;         // what follows this block's last layer
      const int after_floats = blk + 1 < BLOCKS ? D * 3 * D : PI0_KC * PI0;
      F16_STAGE(layernorm_rows(z, B + B_LN1, B + B_LN1 + D, la, E * LT), D * 3 * D, B + B_PROJ, D * D,
                (dense<D, 3 * D, E * LT, ZS, 8, 4, ACT_NONE>(la, ws, B + B_ATTN + D * 3 * D, [&](int r, int n, const float* v) { put4(r0buf + r * QS + n, v); })))
      F16_STAGE(attention_rows(r0buf, la, E), D * D, B + B_FC, D * FF,
                (dense<D, D, E * LT, ZS, 5, 4, ACT_NONE>(la, ws, B + B_PROJ + D * D, [&](int r, int n, const float* v) { add4(z + r * ZS + n, v); })))
      F16_STAGE(layernorm_rows(z, B + B_LN2, B + B_LN2 + D, la, E * LT), D * FF, B + B_MPROJ, FF * D,
                (dense<D, FF, E * LT, ZS, F16_POL_RC_FC, 4, ACT_GELU>(la, ws, B + B_FC + D * FF, [&](int r, int n, const float* v) { put4(r0buf + r * TS + n, v); })))
      F16_STAGE(, FF * D, after, after_floats,
                (dense<FF, D, E * LT, TS, 5, 4, ACT_NONE>(r0buf, ws, B + B_MPROJ + FF * D, [&](int r, int n, const float* v) { add4(z + r * ZS + n, v); })))
    }
    // heads: env e's 160 features are its LT rows of z. The two 160-wide layers run in K-chunks with the accumulators kept in
    // registers across the chunks (one row pass each).
    {
      constexpr int NC = PI0 / 4, G = NT / NC, RC = E / G;
      static_assert(RC * G == E, "one row pass");
      const int n = (threadIdx.x % NC) * 4, g = threadIdx.x / NC;
      u64 acc[RC][4] = {};
#pragma unroll
      for (int ch = 0; ch < FEAT / PI0_KC; ++ch) {
        const bool last = ch + 1 == FEAT / PI0_KC;
        F16_STAGE(if (ch == 0 && a.features) for (int i = threadIdx.x; i < envs * FEAT; i += NT) a.features[e0 * FEAT + i] = z[(i / D) * ZS + i % D],
                  PI0_KC * PI0, last ? P + O_VF0 : P + O_PI0 + (ch + 1) * PI0_KC * PI0, last ? VF0_KC * VF0 : PI0_KC * PI0,
                  (dense_piece2<PI0_KC, PI0, E, LT * ZS, RC, D>(z, ws, ch * PI0_KC, g, n, acc)))
      }
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = tanhf(sum2(acc[i][c]) + __ldg(P + O_PI0 + FEAT * PI0 + n + c));
        put4(p0 + (g + i * G) * (PI0 + PAD) + n, v);
      }
    }
    {
      constexpr int NC = VF0 / 4, G = NT / NC, RC = E / G;
      static_assert(RC * G == E, "one row pass");
      const int n = (threadIdx.x % NC) * 4, g = threadIdx.x / NC;
      u64 acc[RC][4] = {};
#pragma unroll
      for (int ch = 0; ch < FEAT / VF0_KC; ++ch) {
        const bool last = ch + 1 == FEAT / VF0_KC;
        F16_STAGE(, VF0_KC * VF0, last ? P + O_PI1 : P + O_VF0 + (ch + 1) * VF0_KC * VF0, last ? PI0 * PI1 : VF0_KC * VF0,
                  (dense_piece2<VF0_KC, VF0, E, LT * ZS, RC, D>(z, ws, ch * VF0_KC, g, n, acc)))
      }
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = tanhf(sum2(acc[i][c]) + __ldg(P + O_VF0 + FEAT * VF0 + n + c));
        put4(v0 + (g + i * G) * (VF0 + PAD) + n, v);
      }
    }
    F16_STAGE(, PI0 * PI1, P + O_VF1, VF1_KC * VF1,
              (dense<PI0, PI1, E, PI0 + PAD, 2, 4, ACT_TANH>(p0, ws, P + O_PI1 + PI0 * PI1, [&](int r, int n, const float* v) { put4(p1 + r * (PI1 + PAD) + n, v); })))
    {
      constexpr int NC = VF1 / 4, G = NT / NC, RC = E / G;
      static_assert(RC * G == E, "one row pass");
      const int n = (threadIdx.x % NC) * 4, g = threadIdx.x / NC;
      u64 acc[RC][4] = {};
#pragma unroll
      for (int ch = 0; ch < VF0 / VF1_KC; ++ch) {
        const bool last = ch + 1 == VF0 / VF1_KC;
        F16_STAGE(, VF1_KC * VF1, last ? P + O_ACT : P + O_VF1 + (ch + 1) * VF1_KC * VF1, last ? PI1 * ACT : VF1_KC * VF1,
                  (dense_piece2<VF1_KC, VF1, E, VF0 + PAD, RC, 0>(v0, ws, ch * VF1_KC, g, n, acc)))
      }
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = tanhf(sum2(acc[i][c]) + __ldg(P + O_VF1 + VF0 * VF1 + n + c));
        put4(v1 + (g + i * G) * (VF1 + PAD) + n, v);
      }
    }
    F16_STAGE(, PI1 * ACT, P + O_VAL, VF1,
              (dense<PI1, ACT, E, PI1 + PAD, 1, 4, ACT_NONE>(p1, ws, P + O_ACT + PI1 * ACT, [&](int r, int n, const float* v) { put4(mean + r * ACT + n, v); })))
    F16_STAGE(, VF1, P + O_EMBED, FO * EMB,
              (dense<VF1, 1, E, VF1 + PAD, 1, 1, ACT_NONE>(v1, ws, P + O_VAL + VF1, [&](int r, int, const float* v) {
                if (r < envs) a.values[e0 + r] = v[0];
              })))
    // sample, log-probability (common/distributions.py:125-190: Normal(mean, exp(log_std)), summed over the action), clip
    if (threadIdx.x < envs) {            // mean was written by this stage's predecessor, two barriers ago
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
    __syncthreads();                     // MISC (mean) and WS are free for the next tile
  }
#undef F16_STAGE
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
