// Device-resident batched greedy RNN-T decode (fp32 CUDA-core variant).
//
// Replaces the Python loops of Transducer.decode_greedy (reference
// libreasr/lib/models.py:403-443) and Transducer.transcribe_stream (models.py:528-571),
// which evaluate Joint -> log_softmax -> max -> (Predictor step) one symbol at a time with
// a host sync per symbol (models.py:421), for B independent utterances at once in ONE
// persistent cooperative kernel: hypothesis, frame pointer and predictor state never leave
// HBM/SMEM and the host is not involved until all utterances are done.
//
// Every step of the loop is four grid-wide phases (cooperative_groups grid sync between):
//   A  pp = W1p * g (only utterances whose predictor output changed) and
//      z = tanh(pp + ep[b, t_b])          -- Joint first Linear + Tanh (models.py:126,136-139);
//                                            the encoder half ep = W1e*enc + b1 is hoisted
//   B  logits tile = W2 * z + b2, per-tile (max, argmax, sum exp) partials
//                                         -- second Linear + log_softmax + max (models.py:418-420)
//   R  every CTA redundantly folds the partials, applies the blank / max_iters rule
//      (models.py:408-437) to its private copy of the control state (no extra sync)
//   C.. one phase per predictor layer: GRU cell (haste/nbrc.py:46-56) + BatchNorm eval
//      (custom_rnn.py:210-213); layer 0's input projection is a table lookup
//      (Embedding -> Linear -> kernel_0 folded at finalize, models.py:182-183)
// All contractions are the 8-unit x 32-batch split-K tiles of tile_gemm.cuh.
//
// LM shallow fusion (LMFuser, reference libreasr/lib/lm.py:43-83; call sites models.py:401,431,440,
// 478,558,569) runs inside the same loop when the handle has a language model:
//   F  (only when a non-blank decision meets a valid LM row) fused score tiles
//      alpha * standardised LM row + theta * standardised joint row, blank pinned to -10, and their
//      arg max -- the standardisation constants come from per-tile (sum, sum of squares) partials
//   C.. LM layer i (torch.nn.LSTM cell; layer 0's input projection is a table like the predictor's)
//      shares a grid phase with predictor layer i; the LM's output projection shares the next
//      step's phase A.  The fuser state (h, c, LM row, validity) is per utterance / stream.
#include <cooperative_groups.h>

#include "kernels.h"
#include "tile_gemm.cuh"

namespace cg = cooperative_groups;

namespace rnnt {
namespace {

struct Ctrl {
  double nlp[kDecodeMaxBatch];
  int t[kDecodeMaxBatch];
  int it[kDecodeMaxBatch];
  int ntok[kDecodeMaxBatch];
  int tok[kDecodeMaxBatch];
  int n_eval[kDecodeMaxBatch];
  int len[kDecodeMaxBatch];
  unsigned char active[kDecodeMaxBatch];
  unsigned char emit[kDecodeMaxBatch];
  // LM fusion
  int am[kDecodeMaxBatch];          // arg max of the current evaluation (after fusion when it applies)
  float j_mean[kDecodeMaxBatch], j_rstd[kDecodeMaxBatch];    // standardisation of the joint row (pending fusions)
  float lm_mean[kDecodeMaxBatch], lm_rstd[kDecodeMaxBatch];  // standardisation of the held LM row
  unsigned char pend[kDecodeMaxBatch];      // non-blank decision waiting for the fused arg max
  unsigned char lm_valid[kDecodeMaxBatch];  // the fuser holds an LM row (lm.py:58 `torch.is_tensor(lm_logits)`)
};

constexpr float kLmMinVal = -10.0f;  // lm.py:15
constexpr float kLmStdEps = 1e-5f;   // utils.py:162

// mean and 1 / (std + eps) (unbiased std, torch.Tensor.std) of n values from their sum and sum of squares
__device__ __forceinline__ void standardize_consts(double s, double q, int n, float* mean, float* rstd) {
  const double m = s / n;
  double var = (q - s * m) / (n - 1);
  if (var < 0.0) var = 0.0;
  *mean = (float)m;
  *rstd = 1.0f / ((float)sqrt(var) + kLmStdEps);
}

// ---- phase: GRU layer l for tile job (ut, bt) ------------------------------------------
// hT_in/hT_out feature-major [H][Bp]; x_in (layers >= 1) the BatchNorm'ed output below;
// x_out receives BatchNorm(h_new).  tok/emit index by batch column.
__device__ __forceinline__ void phase_gru(const DecodeWeights& w, int l, int job, int B, int Bp,
                                          const float* hT_in, float* hT_out, const float* x_in, float* x_out,
                                          const int* tok, const unsigned char* emit, float* smem) {
  const int H = w.H, nut = H / TG_UNITS;
  const int ut = job % nut, b0 = (job / nut) * kBatchTile;
  float vh[3], vx[3];
  tile_gemm<3>(w.Rt[l], 3 * H, ut * TG_UNITS * 3, hT_in, Bp, b0, H, smem, vh);
  if (l > 0) tile_gemm<3>(w.Kt[l], 3 * H, ut * TG_UNITS * 3, x_in, Bp, b0, H, smem, vx);
  const int unit = ut * TG_UNITS + (threadIdx.x >> 5);
  const int b = b0 + (threadIdx.x & 31);
  if (b >= B) return;
  const size_t si = (size_t)unit * Bp + b;
  const float h_old = hT_in[si];
  float h_new = h_old;
  if (emit[b]) {
    if (l == 0) {
      const float* row = w.table0 + (size_t)tok[b] * (3 * H);
      vx[0] = row[unit]; vx[1] = row[H + unit]; vx[2] = row[2 * H + unit];
    } else {
#pragma unroll
      for (int g = 0; g < 3; ++g) vx[g] += w.kbias[l][unit * 3 + g];
    }
    const float* rb = w.rbias[l] + unit * 3;
    const float z = sigmoidf_acc(vx[0] + (vh[0] + rb[0]));
    const float r = sigmoidf_acc(vx[1] + (vh[1] + rb[1]));
    const float g = tanhf(vx[2] + r * (vh[2] + rb[2]));
    h_new = z * h_old + (1.0f - z) * g;
  }
  hT_out[si] = h_new;
  x_out[si] = h_new * w.bn_scale[l][unit] + w.bn_shift[l][unit];
}

// ---- phase A: pp (for updated columns) and z = tanh(pp + ep[b, t_b]) --------------------
__device__ __forceinline__ void phase_pp(const DecodeWeights& w, int job, int B, int Bp, int T, bool any_upd,
                                         const float* gT, float* ppT, float* zT, const float* ep,
                                         const int* t_cur, const unsigned char* upd, const unsigned char* active,
                                         float* smem) {
  const int J = w.J, njt = J / 32;
  const int jt = job % njt, b0 = (job / njt) * kBatchTile;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (any_upd) tile_gemm<4>(w.W1p_t, J, jt * 32, gT, Bp, b0, w.H, smem, v);
  const int j0 = jt * 32 + (threadIdx.x >> 5) * 4;
  const int b = b0 + (threadIdx.x & 31);
  if (b >= B) return;
  float pp[4];
  if (any_upd && upd[b]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      pp[g] = v[g];
      ppT[(size_t)(j0 + g) * Bp + b] = v[g];
    }
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) pp[g] = ppT[(size_t)(j0 + g) * Bp + b];
  }
  if (active[b]) {
    const float4 e = *reinterpret_cast<const float4*>(ep + ((size_t)b * T + t_cur[b]) * J + j0);
    zT[(size_t)(j0 + 0) * Bp + b] = tanhf(pp[0] + e.x);
    zT[(size_t)(j0 + 1) * Bp + b] = tanhf(pp[1] + e.y);
    zT[(size_t)(j0 + 2) * Bp + b] = tanhf(pp[2] + e.z);
    zT[(size_t)(j0 + 3) * Bp + b] = tanhf(pp[3] + e.w);
  }
}

// ---- phase B: 32 vocabulary rows x 32 batch columns of logits + softmax partials --------
__device__ __forceinline__ void phase_logits(const DecodeWeights& w, int job, int B, int Bp, const float* zT,
                                             float* part, float* logits_out /*[B][V] or null*/,
                                             float* trace, int trace_cap, const int* n_eval,
                                             const unsigned char* active, float* smem,
                                             float* logitT = nullptr /*[V][Bp]*/, double* jpart = nullptr) {
  const int V = w.V, nvt = V / 32;
  const int vt = job % nvt, b0 = (job / nvt) * kBatchTile;
  float v[4];
  tile_gemm<4>(w.W2_t, V, vt * 32, zT, Bp, b0, w.J, smem, v);
  const int uu = threadIdx.x >> 5, bb = threadIdx.x & 31;
  const int v0 = vt * 32 + uu * 4;
  const int b = b0 + bb;
  float* lt = smem;  // [32 rows][33]
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    v[g] += w.b2[v0 + g];
    lt[(uu * 4 + g) * 33 + bb] = v[g];
  }
  if (b < B) {
    if (logits_out) {
#pragma unroll
      for (int g = 0; g < 4; ++g) logits_out[(size_t)b * V + v0 + g] = v[g];
    }
    if (trace && active[b] && n_eval[b] < trace_cap) {
      float* tr = trace + ((size_t)b * trace_cap + n_eval[b]) * V + v0;
#pragma unroll
      for (int g = 0; g < 4; ++g) tr[g] = v[g];
    }
    if (logitT) {
#pragma unroll
      for (int g = 0; g < 4; ++g) logitT[(size_t)(v0 + g) * Bp + b] = v[g];
    }
  }
  __syncthreads();
  if (jpart && threadIdx.x >= 32 && threadIdx.x < 64 && b < B) {   // second warp: (sum, sum of squares) for the fuser
    double sd = 0.0, qd = 0.0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const double x = (double)lt[r * 33 + bb];
      sd += x; qd += x * x;
    }
    jpart[((size_t)vt * Bp + b) * 2] = sd;
    jpart[((size_t)vt * Bp + b) * 2 + 1] = qd;
  }
  if (part && threadIdx.x < 32 && b < B) {
    float m = lt[bb];
    int am = 0;
#pragma unroll 4
    for (int r = 1; r < 32; ++r) {
      const float x = lt[r * 33 + bb];
      if (x > m) { m = x; am = r; }
    }
    float s = 0.f;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) s += expf(lt[r * 33 + bb] - m);
    float4 o;
    o.x = m; o.y = __int_as_float(vt * 32 + am); o.z = s; o.w = 0.f;
    *reinterpret_cast<float4*>(part + ((size_t)vt * Bp + b) * 4) = o;
  }
  __syncthreads();
}

// ---- LM layer l for tile job (ut, bt): torch.nn.LSTM cell, gate order i,f,g,o (lm.py:23,33-36) ----
__device__ __forceinline__ void phase_lm_layer(const LmWeights& lm, int l, int job, int B, int Bp,
                                               const float* hT_in, float* hT_out, float* cT, const float* x_in,
                                               const int* tok, const unsigned char* emit, float* smem) {
  const int Hl = lm.Hl, nut = Hl / TG_UNITS;
  const int ut = job % nut, b0 = (job / nut) * kBatchTile;
  float vh[4], vx[4];
  tile_gemm<4>(lm.Rt[l], 4 * Hl, ut * TG_UNITS * 4, hT_in, Bp, b0, Hl, smem, vh);
  if (l > 0) tile_gemm<4>(lm.Wt[l], 4 * Hl, ut * TG_UNITS * 4, x_in, Bp, b0, Hl, smem, vx);
  const int unit = ut * TG_UNITS + (threadIdx.x >> 5);
  const int b = b0 + (threadIdx.x & 31);
  if (b >= B) return;
  const size_t si = (size_t)unit * Bp + b;
  float h_new = hT_in[si];
  if (emit[b]) {   // only streams that emitted a token advance their LM (lm.py:50-54)
    if (l == 0) {  // Embedding -> W_ih0 (+ both biases) folded into a [V][4Hl] table
      const float* row = lm.table0 + (size_t)tok[b] * (4 * Hl);
#pragma unroll
      for (int g = 0; g < 4; ++g) vx[g] = row[g * Hl + unit];
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) vx[g] += lm.bias[l][unit * 4 + g];
    }
    const float ig = sigmoidf_acc(vx[0] + vh[0]);
    const float fg = sigmoidf_acc(vx[1] + vh[1]);
    const float gg = tanhf(vx[2] + vh[2]);
    const float og = sigmoidf_acc(vx[3] + vh[3]);
    const float c_new = fg * cT[si] + ig * gg;
    cT[si] = c_new;
    h_new = og * tanhf(c_new);
  }
  hT_out[si] = h_new;
}

// ---- LM output projection tile (lm.py:37-40): raw logits + (sum, sum of squares) partials.  log_softmax is
// a per-row shift, which standardisation removes, so the fuser works on the raw logits. ----
__device__ __forceinline__ void phase_lm_logits(const LmWeights& lm, int V, int job, int B, int Bp, const float* hT,
                                                float* logitsT, double* lmpart, float* smem) {
  const int nvt = V / 32;
  const int vt = job % nvt, b0 = (job / nvt) * kBatchTile;
  float v[4];
  tile_gemm<4>(lm.Wo_t, V, vt * 32, hT, Bp, b0, lm.Hl, smem, v);
  const int uu = threadIdx.x >> 5, bb = threadIdx.x & 31;
  const int v0 = vt * 32 + uu * 4;
  const int b = b0 + bb;
  float* lt = smem;  // [32 rows][33]
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    v[g] += lm.bo[v0 + g];
    lt[(uu * 4 + g) * 33 + bb] = v[g];
    if (b < B) logitsT[(size_t)(v0 + g) * Bp + b] = v[g];
  }
  __syncthreads();
  if (threadIdx.x < 32 && b < B) {
    double sd = 0.0, qd = 0.0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const double x = (double)lt[r * 33 + bb];
      sd += x; qd += x * x;
    }
    lmpart[((size_t)vt * Bp + b) * 2] = sd;
    lmpart[((size_t)vt * Bp + b) * 2 + 1] = qd;
  }
  __syncthreads();
}

// ---- phase F: fused score tile alpha * LM row + theta * joint row (lm.py:59-77) and its (max, argmax) ----
__device__ __forceinline__ void phase_fuse(const LmWeights& lm, int V, int job, int B, int Bp, const float* lm_logitsT,
                                           const float* logitT, const Ctrl& c, float* fpart, float* smem) {
  const int nvt = V / 32;
  const int vt = job % nvt, b0 = (job / nvt) * kBatchTile;
  const int uu = threadIdx.x >> 5, bb = threadIdx.x & 31;
  const int v0 = vt * 32 + uu * 4;
  const int b = b0 + bb;
  float* lt = smem;  // [32 rows][33]
  const bool on = b < B && c.pend[b];
  if (on) {
    const float lmean = c.lm_mean[b], lrstd = c.lm_rstd[b], jmean = c.j_mean[b], jrstd = c.j_rstd[b];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int v = v0 + g;
      float a = (lm_logitsT[(size_t)v * Bp + b] - lmean) * lrstd;
      float j = (logitT[(size_t)v * Bp + b] - jmean) * jrstd;
      if (v == 0) { a = kLmMinVal; j = kLmMinVal; }   // lm.py:54,61: entry 0 (blank) of both rows
      lt[(uu * 4 + g) * 33 + bb] = __fadd_rn(__fmul_rn(lm.alpha, a), __fmul_rn(lm.theta, j));
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && on) {
    float m = lt[bb];
    int am = 0;
#pragma unroll 4
    for (int r = 1; r < 32; ++r) {
      const float x = lt[r * 33 + bb];
      if (x > m) { m = x; am = r; }
    }
    fpart[((size_t)vt * Bp + b) * 2] = m;
    fpart[((size_t)vt * Bp + b) * 2 + 1] = __int_as_float(vt * 32 + am);
  }
  __syncthreads();
}

// ---- the persistent loop ------------------------------------------------------------------
__global__ void __launch_bounds__(TG_THREADS, 1) decode_greedy_kernel(DecodeArgs p) {
  extern __shared__ __align__(16) float smem[];
  Ctrl& c = *reinterpret_cast<Ctrl*>(smem + TG_SMEM_FLOATS);
  __shared__ int s_flags[3];
  cg::grid_group grid = cg::this_grid();
  const DecodeWeights& w = p.w;
  const int tid = threadIdx.x, B = p.B, Bp = p.Bp, T = p.T;
  const int nbt = Bp / kBatchTile;
  const int jobs_gru = (w.H / TG_UNITS) * nbt, jobs_pp = (w.J / 32) * nbt, jobs_lg = (w.V / 32) * nbt;
  const int nvt = w.V / 32;

  for (int b = tid; b < kDecodeMaxBatch; b += blockDim.x) {
    const int len = (b < B) ? (p.lens_T ? min(p.lens_T[b], T) : T) : 0;
    c.len[b] = len;
    c.t[b] = 0; c.it[b] = 0; c.ntok[b] = 0; c.n_eval[b] = 0; c.nlp[b] = 0.0;
    c.tok[b] = w.bos;
    c.active[b] = (len > 0);
    c.emit[b] = (b < B);  // doubles as "pp must be (re)computed" on the first pass
  }
  __syncthreads();

  // ---- LM fuser state (lm.py:43-48,81-83) ----
  const LmWeights& lm = p.lm;
  const bool lm_on = lm.L > 0;
  const int Hl = lm.Hl;
  const size_t lhb = (size_t)Hl * Bp;
  auto lm_h = [&](int l, int buf) { return p.lms.h + ((size_t)l * 2 + buf) * lhb; };
  if (lm_on) {
    for (int b = tid; b < kDecodeMaxBatch; b += blockDim.x) {
      const bool v = b < B && p.lms.valid[b] != 0.f;
      c.lm_valid[b] = v;
      c.lm_mean[b] = v ? p.lms.stats[b] : 0.f;
      c.lm_rstd[b] = v ? p.lms.stats[Bp + b] : 0.f;
      c.pend[b] = 0;
    }
    __syncthreads();
  }

  int par = 0;      // ping-pong parity of the predictor state buffers
  int lm_par = 0;   // ... and of the LM hidden state buffers
  bool lm_lin_pending = false;   // the LM advanced; its output projection has not run yet
  // predictor layers (and, after an emitted token, the LM layers: layer i of both shares a grid phase)
  auto run_predictor = [&](bool with_lm) {
    const int nlm = with_lm ? lm.L : 0;
    const int nph = w.Lp > nlm ? w.Lp : nlm;
    const int jobs_lm = lm_on ? (Hl / TG_UNITS) * nbt : 0;
    for (int l = 0; l < nph; ++l) {
      if (l < w.Lp) {
        float* x_out = (l == w.Lp - 1) ? p.gT : (l & 1 ? p.xT + (size_t)w.H * Bp : p.xT);
        const float* x_in = (l == 0) ? nullptr : ((l - 1) & 1 ? p.xT + (size_t)w.H * Bp : p.xT);
        for (int job = blockIdx.x; job < jobs_gru; job += gridDim.x)
          phase_gru(w, l, job, B, Bp, p.hT[l][par], p.hT[l][par ^ 1], x_in, x_out, c.tok, c.emit, smem);
      }
      if (l < nlm) {
        for (int job = blockIdx.x; job < jobs_lm; job += gridDim.x)
          phase_lm_layer(lm, l, job, B, Bp, lm_h(l, lm_par), lm_h(l, lm_par ^ 1), p.lms.c + (size_t)l * lhb,
                         l > 0 ? lm_h(l - 1, lm_par ^ 1) : nullptr, c.tok, c.emit, smem);
      }
      grid.sync();
    }
    par ^= 1;
    if (nlm) { lm_par ^= 1; lm_lin_pending = true; }
  };
  // standardisation constants of the LM row just projected (LMFuser.advance, lm.py:50-54); identical in every CTA
  auto fold_lm_stats = [&]() {
    if (tid < B && c.emit[tid]) {
      double sd = 0.0, qd = 0.0;
      for (int vt = 0; vt < nvt; ++vt) {
        sd += p.lmpart[((size_t)vt * Bp + tid) * 2];
        qd += p.lmpart[((size_t)vt * Bp + tid) * 2 + 1];
      }
      standardize_consts(sd, qd, w.V, &c.lm_mean[tid], &c.lm_rstd[tid]);
      c.lm_valid[tid] = 1;
    }
    __syncthreads();
  };

  if (!p.use_state_in) run_predictor(false);  // feed BOS from the learnable initial state (models.py:397-398)

  bool any_upd = true;
  while (true) {
    for (int job = blockIdx.x; job < jobs_pp; job += gridDim.x)
      phase_pp(w, job, B, Bp, T, any_upd, p.gT, p.ppT, p.zT, p.ep, c.t, c.emit, c.active, smem);
    if (lm_lin_pending)   // independent of phase A: the LM's output projection for the tokens emitted last step
      for (int job = blockIdx.x; job < jobs_lg; job += gridDim.x)
        phase_lm_logits(lm, w.V, job, B, Bp, lm_h(lm.L - 1, lm_par), p.lms.logits, p.lmpart, smem);
    grid.sync();
    if (lm_lin_pending) { fold_lm_stats(); lm_lin_pending = false; }
    for (int job = blockIdx.x; job < jobs_lg; job += gridDim.x)
      phase_logits(w, job, B, Bp, p.zT, p.part, nullptr, p.trace, p.trace_cap, c.n_eval, c.active, smem,
                   lm_on ? p.logitT : nullptr, lm_on ? p.jpart : nullptr);
    grid.sync();

    // ---- R: fold partials, apply the greedy rule; identical in every CTA ----
    if (tid < B) {
      const int b = tid;
      c.pend[b] = 0;
      if (c.active[b]) {
        float M = -INFINITY, S = 0.f;
        int am = 0;
        for (int vt = 0; vt < nvt; ++vt) {
          const float4 q = *reinterpret_cast<const float4*>(p.part + ((size_t)vt * Bp + b) * 4);
          if (q.x > M) {
            S = S * expf(M - q.x) + q.z;
            M = q.x;
            am = __float_as_int(q.y);
          } else {
            S += q.z * expf(q.x - M);
          }
        }
        const float lse = M + logf(S);
        const float prob = M - lse;  // log_softmax value of the arg max (models.py:418-420); pre-fusion (models.py:422)
        const int ne = c.n_eval[b];
        if (blockIdx.x == 0 && p.trace && ne < p.trace_cap) p.trace_lse[(size_t)b * p.trace_cap + ne] = lse;
        c.n_eval[b] = ne + 1;
        c.nlp[b] += (double)prob;
        c.am[b] = am;
        if (lm_on && am != w.blank && c.lm_valid[b]) {   // LMFuser.fuse applies (models.py:427-431, lm.py:58)
          double sd = 0.0, qd = 0.0;
          for (int vt = 0; vt < nvt; ++vt) {
            sd += p.jpart[((size_t)vt * Bp + b) * 2];
            qd += p.jpart[((size_t)vt * Bp + b) * 2 + 1];
          }
          standardize_consts(sd, qd, w.V, &c.j_mean[b], &c.j_rstd[b]);
          c.pend[b] = 1;
        }
      }
    }
    __syncthreads();
    if (lm_on) {
      if (tid == 0) {
        int ap = 0;
        for (int b = 0; b < B; ++b) ap |= c.pend[b];
        s_flags[2] = ap;
      }
      __syncthreads();
      if (s_flags[2]) {   // ---- F: fused arg max for the pending decisions ----
        for (int job = blockIdx.x; job < jobs_lg; job += gridDim.x)
          phase_fuse(lm, w.V, job, B, Bp, p.lms.logits, p.logitT, c, p.fpart, smem);
        grid.sync();
        if (tid < B && c.pend[tid]) {
          float m = -INFINITY;
          int am = 0;
          for (int vt = 0; vt < nvt; ++vt) {   // ascending vocabulary order, strict >: first maximum like torch.max
            const float x = p.fpart[((size_t)vt * Bp + tid) * 2];
            if (x > m) { m = x; am = __float_as_int(p.fpart[((size_t)vt * Bp + tid) * 2 + 1]); }
          }
          c.am[tid] = am;
        }
        __syncthreads();
      }
    }
    if (tid < B) {
      const int b = tid;
      unsigned char emit = 0;
      if (c.active[b]) {
        const int am = c.am[b], t = c.t[b];
        const int it = c.it[b] + 1;
        bool advance;
        if (am == w.blank && !c.pend[b]) {
          advance = true;
        } else {
          const int n = c.ntok[b];
          if (blockIdx.x == 0 && n < p.U_cap) p.tokens[(size_t)b * p.U_cap + n] = am;
          c.ntok[b] = n + 1;
          c.tok[b] = am;
          emit = 1;
          advance = (it >= p.max_iters);
        }
        if (advance) {
          if (blockIdx.x == 0 && p.iters) p.iters[(size_t)b * T + t] = (uint8_t)it;
          c.t[b] = t + 1;
          c.it[b] = 0;
          if (t + 1 >= c.len[b]) c.active[b] = 0;
        } else {
          c.it[b] = it;
        }
      }
      c.emit[b] = emit;
    }
    __syncthreads();
    if (tid == 0) {
      int ae = 0, aa = 0;
      for (int b = 0; b < B; ++b) { ae |= c.emit[b]; aa |= c.active[b]; }
      s_flags[0] = ae; s_flags[1] = aa;
    }
    __syncthreads();
    const bool any_emit = s_flags[0] != 0, any_active = s_flags[1] != 0;
    if (any_emit) run_predictor(lm_on);
    any_upd = any_emit;
    if (!any_active) break;
  }

  // ---- epilogue: results, state back in buffer 0, trace -> log_softmax ----
  if (lm_lin_pending) {   // the stream's fuser must hold the row of the last emitted token (lm.py:50-54)
    for (int job = blockIdx.x; job < jobs_lg; job += gridDim.x)
      phase_lm_logits(lm, w.V, job, B, Bp, lm_h(lm.L - 1, lm_par), p.lms.logits, p.lmpart, smem);
    grid.sync();
    fold_lm_stats();
  }
  if (blockIdx.x == 0 && tid < B) {
    p.ntok[tid] = c.ntok[tid];
    if (p.neg_logp) p.neg_logp[tid] = -c.nlp[tid];
    if (lm_on) {
      p.lms.valid[tid] = c.lm_valid[tid] ? 1.f : 0.f;
      p.lms.stats[tid] = c.lm_mean[tid];
      p.lms.stats[Bp + tid] = c.lm_rstd[tid];
    }
  }
  const size_t gsz = (size_t)gridDim.x * blockDim.x, gid = (size_t)blockIdx.x * blockDim.x + tid;
  if (par) {
    const size_t n = (size_t)w.H * Bp;
    for (int l = 0; l < w.Lp; ++l)
      for (size_t i = gid; i < n; i += gsz) p.hT[l][0][i] = p.hT[l][1][i];
  }
  if (lm_on && lm_par) {
    for (int l = 0; l < lm.L; ++l)
      for (size_t i = gid; i < lhb; i += gsz) lm_h(l, 0)[i] = lm_h(l, 1)[i];
  }
  if (p.trace) {
    grid.sync();
    const float* lse = p.trace_lse;
    const size_t n = (size_t)B * p.trace_cap * w.V;
    for (size_t i = gid; i < n; i += gsz) {
      const size_t be = i / w.V;
      const int b = (int)(be / p.trace_cap), e = (int)(be % p.trace_cap);
      if (e < c.n_eval[b]) p.trace[i] -= lse[be];
    }
  }
}

// ---- standalone wrappers (Predictor.forward / Joint.forward entry points) ------------------
__global__ void __launch_bounds__(TG_THREADS) gru_layer_kernel(PredictArgs p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_tok[kDecodeMaxBatch];
  __shared__ unsigned char s_emit[kDecodeMaxBatch];
  for (int b = threadIdx.x; b < kDecodeMaxBatch; b += blockDim.x) {
    s_tok[b] = (b < p.B && p.tokens) ? p.tokens[b] : 0;
    s_emit[b] = 1;
  }
  __syncthreads();
  phase_gru(p.w, p.layer, blockIdx.x, p.B, p.Bp, p.hT_in, p.hT_out, p.xT_in, p.xT_out, s_tok, s_emit, smem);
}

__global__ void __launch_bounds__(TG_THREADS) joint_hidden_kernel(JointArgs p) {
  // zT = tanh(W1p * g + W1e * e + b1)   (Joint: cat + Linear + Tanh, models.py:136-139,126)
  extern __shared__ __align__(16) float smem[];
  const int J = p.w.J, njt = J / 32;
  const int jt = blockIdx.x % njt, b0 = (blockIdx.x / njt) * kBatchTile;
  float vp[4], ve[4];
  tile_gemm<4>(p.w.W1p_t, J, jt * 32, p.gT, p.Bp, b0, p.w.H, smem, vp);
  tile_gemm<4>(p.w.W1e_t, J, jt * 32, p.eT, p.Bp, b0, p.w.H, smem, ve);
  const int j0 = jt * 32 + (threadIdx.x >> 5) * 4;
  const int b = b0 + (threadIdx.x & 31);
  if (b >= p.B) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) p.zT[(size_t)(j0 + g) * p.Bp + b] = tanhf((vp[g] + ve[g]) + p.w.b1[j0 + g]);
}

__global__ void __launch_bounds__(TG_THREADS) joint_logits_kernel(JointArgs p) {
  extern __shared__ __align__(16) float smem[];
  phase_logits(p.w, blockIdx.x, p.B, p.Bp, p.zT, nullptr, p.logits, nullptr, 0, nullptr, nullptr, smem);
}

}  // namespace

size_t decode_smem_bytes() { return TG_SMEM_BYTES + sizeof(Ctrl) + 16; }

cudaError_t configure_decode(int device, int* max_coop_blocks) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(decode_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem_bytes())) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gru_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(joint_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(joint_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES)) != cudaSuccess) return e;
  int per_sm = 0, sms = 0, coop = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_greedy_kernel, TG_THREADS, decode_smem_bytes())) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device)) != cudaSuccess) return e;
  if (!coop || per_sm < 1) return cudaErrorCooperativeLaunchTooLarge;
  *max_coop_blocks = sms;  // one persistent CTA per SM
  return cudaSuccess;
}

cudaError_t launch_decode(const DecodeArgs& a, int grid_blocks, cudaStream_t st) {
  DecodeArgs args = a;
  void* kargs[] = {&args};
  return launch_persistent((const void*)decode_greedy_kernel, dim3(grid_blocks), dim3(TG_THREADS), kargs, decode_smem_bytes(), st);
}

cudaError_t launch_gru_layer(const PredictArgs& a, cudaStream_t st) {
  const int jobs = (a.w.H / TG_UNITS) * (a.Bp / kBatchTile);
  gru_layer_kernel<<<jobs, TG_THREADS, TG_SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_joint(const JointArgs& a, cudaStream_t st) {
  const int nbt = a.Bp / kBatchTile;
  joint_hidden_kernel<<<(a.w.J / 32) * nbt, TG_THREADS, TG_SMEM_BYTES, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  joint_logits_kernel<<<(a.w.V / 32) * nbt, TG_THREADS, TG_SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

}  // namespace rnnt
