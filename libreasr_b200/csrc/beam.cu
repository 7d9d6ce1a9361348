// Batched RNN-T beam search with hypotheses and predictor state resident in HBM (BASELINE.json configs[3], [4]).
//
// PARITY UNPINNED IN THE REFERENCE: iceychris/LibreASR has no beam search (libreasr/lib/models.py:8 is an unused
// PriorityQueue import).  The algorithm is defined by oracle/beam.py (a breadth-first, iteration-capped RNN-T beam search
// that keeps the reference's max_iters rule, models.py:369) on the reference's own Predictor / Joint modules
// (models.py:116-187); this file is its device implementation:
//
//   * B utterances x W hypothesis slots = R rows advance in lock-step, one frame at a time, max_iters expansions per
//     frame.  Every dense contraction of an expansion -- pp = g W1p^T, logits = tanh(pp + ep[t]) W2^T + b2,
//     R0 h0, K1 BN(h0'), R1 h1 -- is ONE tcgen05 3xFP16 GEMM over all R rows (gemm_tc.cu), with the operand images
//     written directly by the producing kernels (parent-state gathers fused into the image conversion).
//   * All control is on the device: per-row log-softmax + top-W (beam_topk_kernel), per-utterance candidate
//     selection / blank bookkeeping (beam_select_kernel), end-of-frame merge of equal token sequences + top-W
//     (beam_merge_kernel).  Token sequences live in a per-utterance node pool (parent pointers), hypotheses are
//     (score fp64, sequence hash, node, state slot); the host only enqueues the fixed launch sequence of a frame.
#include "kernels.h"
#include "tc_common.cuh"

namespace rnnt {
namespace {

constexpr unsigned long long kHashMul = 0x9E3779B97F4A7C15ull;

// fp32 rows (optionally gathered through row_map) -> hi/lo operand image with 128-row tiles
__global__ void __launch_bounds__(256) beam_rows_to_image_kernel(const float* __restrict__ src, int ld, const int* __restrict__ row_map, int R, int K,
                                                                 uint8_t* __restrict__ img) {
  const int KB = (K + kImgK - 1) / kImgK;
  const int rt = blockIdx.y, kb = blockIdx.x;
  uint8_t* hi_t = img + img_tile_offset(rt, kb, 0, KB, 128);
  uint8_t* lo_t = img + img_tile_offset(rt, kb, 1, KB, 128);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int row = rt * 128 + r;
    const int k0 = kb * kImgK + c * 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (row < R) {
      const float* s = src + (size_t)(row_map ? row_map[row] : row) * ld + k0;
      const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16x3(x[j], h[j], l[j]);
    const uint32_t off = img_elem_offset(r, c * 8);
    *reinterpret_cast<uint4*>(hi_t + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo_t + off) = *reinterpret_cast<const uint4*>(l);
  }
}

// z = tanh(pp + ep[b, t_b]) (Joint.forward, models.py:132-140, first Linear split into pred / enc halves) -> operand image
__global__ void __launch_bounds__(256) beam_z_image_kernel(const float* __restrict__ pp, const float* __restrict__ ep, const int32_t* __restrict__ lens_T, int t,
                                                           int T, int W, int R, int J, uint8_t* __restrict__ img) {
  const int KB = J / kImgK;
  const int rt = blockIdx.y, kb = blockIdx.x;
  uint8_t* hi_t = img + img_tile_offset(rt, kb, 0, KB, 128);
  uint8_t* lo_t = img + img_tile_offset(rt, kb, 1, KB, 128);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int row = rt * 128 + r;
    const int k0 = kb * kImgK + c * 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (row < R) {
      const int b = row / W;
      const int len = lens_T ? min(lens_T[b], T) : T;
      const int tt = min(t, max(len - 1, 0));
      const float* p0 = pp + (size_t)row * J + k0;
      const float* e0 = ep + ((size_t)b * T + tt) * J + k0;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = tanhf(p0[j] + e0[j]);
    }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16x3(x[j], h[j], l[j]);
    const uint32_t off = img_elem_offset(r, c * 8);
    *reinterpret_cast<uint4*>(hi_t + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo_t + off) = *reinterpret_cast<const uint4*>(l);
  }
}

// One block per row: log-sum-exp of the V logits, log p(blank), and the W best non-blank log-probabilities
// (value descending, ties: lower token id) -- the per-parent candidates of oracle/beam.py.
constexpr int BK_THREADS = 256, BK_MAX_PER = 16, BK_MAXW = 8;
__global__ void __launch_bounds__(BK_THREADS) beam_topk_kernel(const float* __restrict__ logits, int V, int W, int blank, float* __restrict__ cand_val,
                                                               int* __restrict__ cand_idx, float* __restrict__ lp_blank) {
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)r * V;
  float v[BK_MAX_PER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < BK_MAX_PER; ++i) {
    const int k = tid + i * BK_THREADS;
    v[i] = k < V ? x[k] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  __shared__ float sred[BK_THREADS / 32];
  __shared__ float sval[BK_THREADS / 32];
  __shared__ int sidx[BK_THREADS / 32];
  __shared__ float bcast[2];
  __shared__ int bidx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) sred[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = sred[0];
    for (int w = 1; w < BK_THREADS / 32; ++w) mm = fmaxf(mm, sred[w]);
    bcast[0] = mm;
  }
  __syncthreads();
  m = bcast[0];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < BK_MAX_PER; ++i)
    if (tid + i * BK_THREADS < V) s += expf(v[i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float ss = 0.f;
    for (int w = 0; w < BK_THREADS / 32; ++w) ss += sred[w];
    bcast[1] = m + logf(ss);
  }
  __syncthreads();
  const float lse = bcast[1];
  if (tid == 0) lp_blank[r] = x[blank] - lse;
  {   // blank never competes
    const int i = blank / BK_THREADS;
    if (blank % BK_THREADS == tid) {
#pragma unroll
      for (int q = 0; q < BK_MAX_PER; ++q)
        if (q == i) v[q] = -INFINITY;
    }
  }
  for (int j = 0; j < W; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < BK_MAX_PER; ++i) {
      const int k = tid + i * BK_THREADS;
      if (k < V && (v[i] > bv || (v[i] == bv && k < bi))) { bv = v[i]; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sval[warp] = bv; sidx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float fv = sval[0];
      int fi = sidx[0];
      for (int w = 1; w < BK_THREADS / 32; ++w)
        if (sval[w] > fv || (sval[w] == fv && sidx[w] < fi)) { fv = sval[w]; fi = sidx[w]; }
      cand_val[(size_t)r * W + j] = fv - lse;
      cand_idx[(size_t)r * W + j] = fi;
      bidx = fi;
    }
    __syncthreads();
    const int win = bidx;
    if (win % BK_THREADS == tid) {
      const int i = win / BK_THREADS;
#pragma unroll
      for (int q = 0; q < BK_MAX_PER; ++q)
        if (q == i) v[q] = -INFINITY;
    }
    __syncthreads();
  }
}

// Expansion `it` of the current frame for utterance b (one thread; at most W*W candidates):
//   every valid hypothesis of the expanding generation leaves the frame with its blank (-> leave list),
//   the W best (parent, token) candidates form the next generation (stable: parent slot, then token rank).
__global__ void beam_select_kernel(BeamArgs a, int it, int t, int gi, int gn) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int W = a.W, R = a.B * W;
  const int len = a.lens_T ? min(a.lens_T[b], a.T) : a.T;
  BeamHyp* cur = a.hyp + (size_t)gi * R + (size_t)b * W;
  BeamHyp* nxt = a.hyp + (size_t)gn * R + (size_t)b * W;
  BeamLeave* lv = a.leave + (size_t)b * a.leave_cap;
  int nl = it == 0 ? 0 : a.n_leave[b];
  if (t >= len) {   // past the end of this utterance: its hypotheses pass through unchanged
    if (it == 0)
      for (int w = 0; w < W; ++w)
        if (cur[w].valid) lv[nl++] = BeamLeave{cur[w].score, cur[w].hash, cur[w].node, gi, w};
    for (int w = 0; w < W; ++w) { nxt[w].valid = 0; a.sel_row[b * W + w] = b * W; a.sel_tok[b * W + w] = 0; }
    a.n_leave[b] = nl;
    return;
  }
  double bs[BK_MAXW];
  int bp[BK_MAXW], bk[BK_MAXW], nb = 0;
  for (int w = 0; w < W; ++w) {
    if (!cur[w].valid) continue;
    const int r = b * W + w;
    lv[nl++] = BeamLeave{cur[w].score + (double)a.lp_blank[r], cur[w].hash, cur[w].node, gi, w};
    for (int j = 0; j < W; ++j) {
      const double sc = cur[w].score + (double)a.cand_val[(size_t)r * W + j];
      // stable insertion: an earlier candidate wins ties
      int pos = nb;
      while (pos > 0 && sc > bs[pos - 1]) --pos;
      if (pos >= W) continue;
      const int last = nb < W ? nb : W - 1;
      for (int q = last; q > pos; --q) { bs[q] = bs[q - 1]; bp[q] = bp[q - 1]; bk[q] = bk[q - 1]; }
      bs[pos] = sc; bp[pos] = w; bk[pos] = a.cand_idx[(size_t)r * W + j];
      if (nb < W) ++nb;
    }
  }
  int nn = a.n_nodes[b];
  for (int j = 0; j < W; ++j) {
    if (j < nb) {
      const BeamHyp& par = cur[bp[j]];
      const int node = nn++;
      a.node_parent[(size_t)b * a.node_cap + node] = par.node;
      a.node_token[(size_t)b * a.node_cap + node] = bk[j];
      nxt[j].score = bs[j];
      nxt[j].hash = par.hash * kHashMul + (unsigned long long)(bk[j] + 1);
      nxt[j].node = node;
      nxt[j].valid = 1;
      a.sel_row[b * W + j] = b * W + bp[j];
      a.sel_tok[b * W + j] = bk[j];
    } else {
      nxt[j].valid = 0;
      a.sel_row[b * W + j] = b * W;
      a.sel_tok[b * W + j] = 0;
    }
  }
  a.n_nodes[b] = nn;
  a.n_leave[b] = nl;
}

// GRU cell of predictor layer 0 + BatchNorm eval (haste/nbrc.py:46-56; input = table row of the token)
__global__ void __launch_bounds__(256) beam_gru0_kernel(BeamArgs a, const float* __restrict__ rec, const float* __restrict__ h_par, float* __restrict__ h_new,
                                                        float* __restrict__ x1, int use_map) {
  const int H = a.w.H, R = a.B * a.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)R * H; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / H), u = (int)(i % H);
    const int pr = use_map ? a.sel_row[r] : r;
    const float* row = a.w.table0 + (size_t)a.sel_tok[r] * (3 * H);
    const float* rc = rec + (size_t)r * 3 * H + (size_t)u * 3;          // incl. recurrent bias (GEMM bias)
    const float hp = h_par[(size_t)pr * H + u];
    const float z = sigmoidf_acc(row[u] + rc[0]);
    const float rr = sigmoidf_acc(row[H + u] + rc[1]);
    const float g = tanhf(row[2 * H + u] + rr * rc[2]);
    const float hn = z * hp + (1.0f - z) * g;
    h_new[i] = hn;
    x1[i] = hn * a.w.bn_scale[0][u] + a.w.bn_shift[0][u];
  }
}
// ... layer 1: input product K1 BN(h0') (+ bias) and recurrent product R1 h1 (+ bias) from GEMMs
__global__ void __launch_bounds__(256) beam_gru1_kernel(BeamArgs a, const float* __restrict__ kin, const float* __restrict__ rec, const float* __restrict__ h_par,
                                                        float* __restrict__ h_new, float* __restrict__ g_new, int use_map) {
  const int H = a.w.H, R = a.B * a.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)R * H; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / H), u = (int)(i % H);
    const int pr = use_map ? a.sel_row[r] : r;
    const float* kx = kin + (size_t)r * 3 * H + (size_t)u * 3;
    const float* rc = rec + (size_t)r * 3 * H + (size_t)u * 3;
    const float hp = h_par[(size_t)pr * H + u];
    const float z = sigmoidf_acc(kx[0] + rc[0]);
    const float rr = sigmoidf_acc(kx[1] + rc[1]);
    const float g = tanhf(kx[2] + rr * rc[2]);
    const float hn = z * hp + (1.0f - z) * g;
    h_new[i] = hn;
    g_new[i] = hn * a.w.bn_scale[1][u] + a.w.bn_shift[1][u];
  }
}

// End of a frame for utterance b: leave list (+ the last generation, which hit the iteration cap) -> merge equal token
// sequences (logaddexp, state of the better one, position of the first) -> W best -> generation 0 of the next frame.
__global__ void beam_merge_kernel(BeamArgs a, int g_last, int g_stage) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int W = a.W, R = a.B * W;
  BeamLeave* lv = a.leave + (size_t)b * a.leave_cap;
  int nl = a.n_leave[b];
  const BeamHyp* last = a.hyp + (size_t)g_last * R + (size_t)b * W;
  for (int w = 0; w < W; ++w)
    if (last[w].valid) lv[nl++] = BeamLeave{last[w].score, last[w].hash, last[w].node, g_last, w};
  // merge (in place, keeping first-occurrence order)
  int nm = 0;
  for (int i = 0; i < nl; ++i) {
    int j = 0;
    for (; j < nm; ++j)
      if (lv[j].hash == lv[i].hash) break;
    if (j == nm) { lv[nm++] = lv[i]; continue; }
    const bool keep_old = lv[j].score >= lv[i].score;
    const double hi = keep_old ? lv[j].score : lv[i].score, lo = keep_old ? lv[i].score : lv[j].score;
    const double tot = hi + log1p(exp(lo - hi));
    if (!keep_old) { lv[j].node = lv[i].node; lv[j].gen = lv[i].gen; lv[j].slot = lv[i].slot; }
    lv[j].score = tot;
  }
  // W best, stable
  int order[BK_MAXW], no = 0;
  for (int i = 0; i < nm; ++i) {
    int pos = no;
    while (pos > 0 && lv[i].score > lv[order[pos - 1]].score) --pos;
    if (pos >= W) continue;
    const int lastp = no < W ? no : W - 1;
    for (int q = lastp; q > pos; --q) order[q] = order[q - 1];
    order[pos] = i;
    if (no < W) ++no;
  }
  BeamHyp* out = a.hyp + (size_t)g_stage * R + (size_t)b * W;
  for (int w = 0; w < W; ++w) {
    if (w < no) {
      const BeamLeave& e = lv[order[w]];
      out[w].score = e.score; out[w].hash = e.hash; out[w].node = e.node; out[w].valid = 1;
      a.copy_src[b * W + w] = e.gen * R + b * W + e.slot;
    } else {
      out[w].valid = 0;
      a.copy_src[b * W + w] = g_stage * R + b * W + w;   // self: nothing to copy
    }
  }
}

// predictor state (h0, h1, g) of the surviving hypotheses -> generation slots of the next frame
__global__ void __launch_bounds__(256) beam_copy_states_kernel(BeamArgs a, int g_stage) {
  const int H = a.w.H, R = a.B * a.W;
  const size_t plane = (size_t)a.n_gen * R * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)R * H; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / H), u = (int)(i % H);
    const size_t src = (size_t)a.copy_src[r] * H + u, dst = ((size_t)g_stage * R + r) * H + u;
    if (src == dst) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k) a.state[k * plane + dst] = a.state[k * plane + src];
  }
}

__global__ void beam_init_kernel(BeamArgs a, int g0, int g_init) {
  const int b = blockIdx.x, W = a.W, R = a.B * W, H = a.w.H;
  const size_t plane = (size_t)a.n_gen * R * H;
  // learnable initial state in the "parent" generation, BOS as the token of every row
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    const int w = i / H, u = i % H;
    const size_t d = ((size_t)g_init * R + b * W + w) * H + u;
    a.state[0 * plane + d] = a.w.h0[0][u];
    a.state[1 * plane + d] = a.w.h0[1][u];
  }
  if (threadIdx.x == 0) {
    for (int w = 0; w < W; ++w) {
      a.sel_row[b * W + w] = b * W + w;
      a.sel_tok[b * W + w] = a.w.bos;
      BeamHyp& h = a.hyp[(size_t)g0 * R + b * W + w];
      h.score = 0.0; h.hash = 0ull; h.node = 0; h.valid = w == 0;
    }
    a.node_parent[(size_t)b * a.node_cap] = -1;
    a.node_token[(size_t)b * a.node_cap] = -1;
    a.n_nodes[b] = 1;
    a.n_leave[b] = 0;
  }
}

__global__ void beam_finish_kernel(BeamArgs a, int g_final, int32_t* tokens, int U_cap, int32_t* ntok, double* score) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int W = a.W, R = a.B * W;
  const BeamHyp* hy = a.hyp + (size_t)g_final * R + (size_t)b * W;
  int best = -1;
  for (int w = 0; w < W; ++w)
    if (hy[w].valid && (best < 0 || hy[w].score > hy[best].score)) best = w;
  int n = 0;
  if (best >= 0) {
    for (int nd = hy[best].node; nd > 0; nd = a.node_parent[(size_t)b * a.node_cap + nd]) ++n;
    int k = n;
    for (int nd = hy[best].node; nd > 0; nd = a.node_parent[(size_t)b * a.node_cap + nd]) {
      --k;
      if (k < U_cap) tokens[(size_t)b * U_cap + k] = a.node_token[(size_t)b * a.node_cap + nd];
    }
    if (score) score[b] = hy[best].score;
  } else if (score) {
    score[b] = 0.0;
  }
  ntok[b] = n < U_cap ? n : U_cap;
}

}  // namespace

size_t beam_state_floats(int B, int W, int H, int n_gen) { return (size_t)3 * n_gen * B * W * H; }

#define BEAM_CK(x)                  \
  do {                              \
    cudaError_t e_ = (x);           \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

// One whole beam decode: enqueues the fixed launch sequence (no host synchronisation).
cudaError_t launch_beam_search(const BeamArgs& a, const BeamBuffers& bf, int max_iters, int32_t* tokens, int U_cap, int32_t* ntok, double* score,
                               int* launches, cudaStream_t st) {
  const int R = a.B * a.W, H = a.w.H, J = a.w.J, V = a.w.V, NG = a.n_gen;
  const size_t plane = (size_t)NG * R * H;
  auto S = [&](int k, int g) { return a.state + k * plane + (size_t)g * R * H; };
  const dim3 gimgH((unsigned)(H / kImgK), (unsigned)ceil_div(R, 128)), gimgJ((unsigned)(J / kImgK), (unsigned)ceil_div(R, 128));
  const int ew_blocks = (int)std::min<int64_t>(ceil_div((int64_t)R * H, 256), 148 * 8);
  int n = 0;
  // predictor step: parents in generation gp (rows through sel_row when use_map), results into generation gn
  auto predictor = [&](int gp, int gn, int use_map) -> cudaError_t {
    beam_rows_to_image_kernel<<<gimgH, 256, 0, st>>>(S(0, gp), H, use_map ? a.sel_row : nullptr, R, H, bf.a_img);
    BEAM_CK(launch_gemm_tc(bf.a_img, bf.r_img[0], a.w.rbias[0], bf.rec, 3 * H, R, 3 * H, H, st));
    beam_gru0_kernel<<<ew_blocks, 256, 0, st>>>(a, bf.rec, S(0, gp), S(0, gn), bf.x1, use_map);
    beam_rows_to_image_kernel<<<gimgH, 256, 0, st>>>(bf.x1, H, nullptr, R, H, bf.a_img);
    BEAM_CK(launch_gemm_tc(bf.a_img, bf.k1_img, a.w.kbias[1], bf.kin, 3 * H, R, 3 * H, H, st));
    beam_rows_to_image_kernel<<<gimgH, 256, 0, st>>>(S(1, gp), H, use_map ? a.sel_row : nullptr, R, H, bf.a_img);
    BEAM_CK(launch_gemm_tc(bf.a_img, bf.r_img[1], a.w.rbias[1], bf.rec, 3 * H, R, 3 * H, H, st));
    beam_gru1_kernel<<<ew_blocks, 256, 0, st>>>(a, bf.kin, bf.rec, S(1, gp), S(1, gn), S(2, gn), use_map);
    n += 8;
    return cudaGetLastError();
  };
  int base = 0;
  beam_init_kernel<<<a.B, 256, 0, st>>>(a, base, NG - 1);
  ++n;
  BEAM_CK(predictor(NG - 1, base, 0));   // BOS from the learnable initial state (models.py:397-398)
  for (int t = 0; t < a.T; ++t) {
    for (int it = 0; it < max_iters; ++it) {
      const int gi = (base + it) % NG, gn = (base + it + 1) % NG;
      beam_rows_to_image_kernel<<<gimgH, 256, 0, st>>>(S(2, gi), H, nullptr, R, H, bf.a_img);
      BEAM_CK(launch_gemm_tc(bf.a_img, bf.w1p_img, nullptr, bf.pp, J, R, J, H, st));
      beam_z_image_kernel<<<gimgJ, 256, 0, st>>>(bf.pp, a.ep, a.lens_T, t, a.T, a.W, R, J, bf.a_img);
      BEAM_CK(launch_gemm_tc(bf.a_img, bf.w2_img, a.w.b2, bf.logits, V, R, V, J, st));
      beam_topk_kernel<<<R, BK_THREADS, 0, st>>>(bf.logits, V, a.W, a.w.blank, a.cand_val, a.cand_idx, a.lp_blank);
      beam_select_kernel<<<a.B, 32, 0, st>>>(a, it, t, gi, gn);
      n += 6;
      BEAM_CK(predictor(gi, gn, 1));
    }
    const int g_last = (base + max_iters) % NG, g_stage = (base + max_iters + 1) % NG;
    beam_merge_kernel<<<a.B, 32, 0, st>>>(a, g_last, g_stage);
    beam_copy_states_kernel<<<ew_blocks, 256, 0, st>>>(a, g_stage);
    n += 2;
    base = g_stage;
  }
  beam_finish_kernel<<<a.B, 32, 0, st>>>(a, base, tokens, U_cap, ntok, score);
  ++n;
  if (launches) *launches = n;
  return cudaGetLastError();
}

}  // namespace rnnt
