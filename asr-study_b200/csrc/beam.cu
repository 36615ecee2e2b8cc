// K8 — CTC prefix beam search (top path) for sm_100a.
// Replaces tf.nn.ctc_beam_search_decoder(y_pred, seq_len, beam_width, top_paths=1, merge_repeated)[0][0]
// (core/ctc_utils.py:44-50; width 100 by default, 400 via utils/core_utils.py:70-71).
//
// One WARP per utterance, faithful to the insertion ORDER of TF 1.3's CTCBeamSearchDecoder::Step()
// (tensorflow/core/util/ctc/ctc_beam_search.h), because that order is observable: a leaf that was
// evicted earlier in the same frame and is then re-proposed by its (earlier-processed) parent either
// re-enters with label-only mass or has its OLD mass reset and cannot spawn children in this frame.
// A "rank everything, keep the best W" formulation cannot reproduce that (it was tried first and
// disagreed with the oracle for narrow beams), so per frame the warp does:
//   A  scores = logits - max;  old <- new for every leaf ("branches", kept sorted by total, descending)
//   B  parent slot of every branch (global node -> slot table) and the child index table
//   C  mass update of every branch from itself and its still-active parent        (lanes = branches)
//   D  grow: branches in order; lanes = labels; a lane is handled sequentially only if its candidate
//      beats the CURRENT worst leaf (the threshold only rises) or if it is an existing branch object;
//      inserting into a full beam evicts the current worst leaf
//   E  survivors are rank-sorted into the next branch list; new leaves get their trie node through a
//      per-utterance hash keyed by (parent node, label) so a prefix that returns is the same node.
// Scores are fp32 like TF.  Ties between equal totals are broken by slot order (TF: heap order).
#include "common.cuh"
#include <math_constants.h>

namespace {

__device__ __forceinline__ float lse2f(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return m;
  // TF: max + log1pf(expf(-|a-b|)) with glibc's (practically correctly rounded) float functions.  Evaluated in
  // fp64 and rounded to fp32 at the same two points so the device and the CPU oracle agree bit for bit; with the
  // device's own expf/log1pf (1-2 ulp) a 999-frame, width-100 search drifts onto a different beam at near ties.
  const float e = (float)exp((double)(fminf(a, b) - m));
  return m + (float)log1p((double)e);
}

struct Br {           // branch arrays (one buffer)
  int* node; int* label; int* par; int* pslot;
  float* ob; float* ol; float* ot; float* nb; float* nl; float* nt;
};

__global__ void __launch_bounds__(32)
beam_kernel(const float* __restrict__ logits, int T, int N, int C, const int* __restrict__ in_len, int blank, int W,
            int merge_repeated, int* __restrict__ out_labels, int* __restrict__ out_len, int* __restrict__ ws,
            long long ws_ints_per_utt, int max_nodes, int hash_size) {
  extern __shared__ unsigned long long sm64[];
  const int n = blockIdx.x, lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  const float NEG = -CUDART_INF_F;
  const int len = min(max(in_len[n], 0), T);
  int* base = ws + (size_t)n * ws_ints_per_utt;
  int* node_parent = base;
  int* node_label = base + max_nodes;
  int* slot_of_node = base + 2 * (size_t)max_nodes;
  unsigned long long* hash = reinterpret_cast<unsigned long long*>(base + 3 * (size_t)max_nodes + (max_nodes & 1));

  // ---- shared memory carve-up ----
  float* fbase = reinterpret_cast<float*>(sm64);
  float* inp = fbase;                          fbase += (C + 3) & ~3;
  Br br[2];
  for (int b = 0; b < 2; ++b) {
    br[b].ob = fbase; br[b].ol = fbase + W; br[b].ot = fbase + 2 * W;
    br[b].nb = fbase + 3 * W; br[b].nl = fbase + 4 * W; br[b].nt = fbase + 5 * W;
    fbase += 6 * W;
  }
  float* nw_nt = fbase;                        fbase += W;
  float* skey = fbase;                         fbase += 2 * W;        // gather keys for the sort
  int* ibase = reinterpret_cast<int*>(fbase);
  for (int b = 0; b < 2; ++b) {
    br[b].node = ibase; br[b].label = ibase + W; br[b].par = ibase + 2 * W; br[b].pslot = ibase + 3 * W;
    ibase += 4 * W;
  }
  int* nw_label = ibase;                       ibase += W;
  int* nw_pidx = ibase;                        ibase += W;            // index of the parent branch (this frame)
  int* nw_free = ibase;                        ibase += W;            // free-slot stack
  int* ssrc = ibase;                           ibase += 2 * W;        // gather: source id (branch j or W + new slot)
  short* child_idx = reinterpret_cast<short*>(ibase);                 // [W][C]

  int* outp = out_labels + (size_t)n * T;
  for (int i = lane; i < hash_size; i += 32) hash[i] = 0ull;
  int cur = 0, m = 1, nodes = 1;
  if (lane == 0) {
    node_parent[0] = -1; node_label[0] = -1; slot_of_node[0] = 0;
    br[0].node[0] = 0; br[0].label[0] = -1; br[0].par[0] = -1;
    br[0].nb[0] = 0.0f; br[0].nl[0] = NEG; br[0].nt[0] = 0.0f;
  }
  __syncwarp();

  for (int t = 0; t < len; ++t) {
    Br& B = br[cur];
    Br& X = br[cur ^ 1];
    // ---- A. frame scores, old <- new --------------------------------------------------------------
    {
      const float* row = logits + ((size_t)t * N + n) * C;
      float mx = NEG;
      for (int k = lane; k < C; k += 32) mx = fmaxf(mx, row[k]);
      mx = asr::warp_max(mx);
      for (int k = lane; k < C; k += 32) inp[k] = row[k] - mx;
    }
    for (int i = lane; i < m; i += 32) { B.ob[i] = B.nb[i]; B.ol[i] = B.nl[i]; B.ot[i] = B.nt[i]; }
    for (int i = lane; i < m * C; i += 32) child_idx[i] = -1;
    __syncwarp();
    // ---- B. parent slots, child table ----------------------------------------------------------------
    for (int i = lane; i < m; i += 32) {
      const int par = B.par[i];
      const int ps = (par >= 0) ? slot_of_node[par] : -1;
      B.pslot[i] = ps;
      if (ps >= 0) child_idx[ps * C + B.label[i]] = (short)i;
    }
    __syncwarp();
    // ---- C. mass update ----------------------------------------------------------------------------------
    for (int i = lane; i < m; i += 32) {
      const int lab = B.label[i], ps = B.pslot[i];
      float nl = B.ol[i];
      if (lab >= 0) {
        if (ps >= 0) nl = lse2f(nl, (lab == B.label[ps]) ? B.ob[ps] : B.ot[ps]);
        nl += inp[lab];
      }
      const float nbk = B.ot[i] + inp[blank];
      B.nb[i] = nbk; B.nl[i] = nl; B.nt[i] = lse2f(nbk, nl);
    }
    __syncwarp();
    // ---- D. grow ---------------------------------------------------------------------------------------------
    int size = m, nw_cnt = 0, nfree = 0;
    float bot_val = NEG;
    int bot_id = -1;                       // < W: branch index ; >= W: W + new slot
    auto find_bottom = [&]() {
      float bv = CUDART_INF_F;
      int bi = 0x7fffffff;
      for (int i = lane; i < m; i += 32) {
        const float v = B.nt[i];
        if (v != NEG && (v < bv || (v == bv && i < bi))) { bv = v; bi = i; }
      }
      for (int i = lane; i < nw_cnt; i += 32) {
        const float v = nw_nt[i];
        if (v != NEG && (v < bv || (v == bv && W + i < bi))) { bv = v; bi = W + i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      bot_val = bv; bot_id = bi;
    };
    if (size == W) find_bottom();
    for (int i = 0; i < m; ++i) {
      const float b_ot = B.ot[i];
      if (!(b_ot > NEG && (size < W || b_ot > bot_val))) continue;
      const int b_label = B.label[i];
      const float b_ob = B.ob[i];
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const bool lab_ok = (c < C) && (c != blank);
        const int j = lab_ok ? (int)child_idx[i * C + c] : -1;
        float tot = NEG;
        if (lab_ok) {
          const float prev = (c == b_label) ? b_ob : b_ot;
          if (prev > NEG) tot = inp[c] + prev;
        }
        const bool pass0 = lab_ok && tot > NEG && (size < W || tot > bot_val);
        unsigned mask = __ballot_sync(FULL, lab_ok && (pass0 || j >= 0));
        while (mask) {
          const int l0 = __ffs(mask) - 1;
          mask &= mask - 1;
          const float tot0 = __shfl_sync(FULL, tot, l0);
          const int j0 = __shfl_sync(FULL, j, l0);
          if (j0 >= 0 && B.nt[j0] != NEG) continue;               // child currently active: TF skips it
          const bool pass = tot0 > NEG && (size < W || tot0 > bot_val);
          if (pass) {
            if (size == W) {                                       // evict the current worst leaf
              if (lane == 0) {
                if (bot_id < W) { B.nb[bot_id] = NEG; B.nl[bot_id] = NEG; B.nt[bot_id] = NEG; }
                else { nw_nt[bot_id - W] = NEG; nw_free[nfree] = bot_id - W; }
              }
              if (bot_id >= W) ++nfree;
              --size;
              __syncwarp();
            }
            if (j0 >= 0) {                                         // an existing object re-enters (label mass only)
              if (lane == 0) { B.nb[j0] = NEG; B.nl[j0] = tot0; B.nt[j0] = tot0; }
            } else {
              int slot;
              if (nfree > 0) { slot = nw_free[nfree - 1]; --nfree; }
              else slot = nw_cnt++;
              if (lane == 0) { nw_nt[slot] = tot0; nw_label[slot] = c0 + l0; nw_pidx[slot] = i; }
            }
            ++size;
            __syncwarp();
            if (size == W) find_bottom();
          } else if (j0 >= 0) {                                    // TF: deactivate child -> oldp AND newp reset
            if (lane == 0) { B.ob[j0] = NEG; B.ol[j0] = NEG; B.ot[j0] = NEG; }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
    // ---- E. next beam: gather survivors, rank-sort by total (descending), assign nodes / slots --------------
    int cnt = 0;
    {
      // warp-level compaction (order: branches then new slots)
      for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        const bool keep = (i < m) && (B.nt[i] != NEG);
        const unsigned bal = __ballot_sync(FULL, keep);
        if (keep) { const int pos = cnt + __popc(bal & ((1u << lane) - 1)); skey[pos] = B.nt[i]; ssrc[pos] = i; }
        cnt += __popc(bal);
      }
      for (int i0 = 0; i0 < nw_cnt; i0 += 32) {
        const int i = i0 + lane;
        const bool keep = (i < nw_cnt) && (nw_nt[i] != NEG);
        const unsigned bal = __ballot_sync(FULL, keep);
        if (keep) { const int pos = cnt + __popc(bal & ((1u << lane) - 1)); skey[pos] = nw_nt[i]; ssrc[pos] = W + i; }
        cnt += __popc(bal);
      }
    }
    __syncwarp();
    // evicted branches leave the node -> slot table
    for (int i = lane; i < m; i += 32)
      if (B.nt[i] == NEG) slot_of_node[B.node[i]] = -1;
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
      const float k = skey[i];
      int rank = 0;
      for (int q = 0; q < cnt; ++q) {
        const float kq = skey[q];
        rank += (kq > k) || (kq == k && q < i);
      }
      const int src = ssrc[i];
      if (src < W) {
        X.node[rank] = B.node[src]; X.label[rank] = B.label[src]; X.par[rank] = B.par[src];
        X.nb[rank] = B.nb[src]; X.nl[rank] = B.nl[src]; X.nt[rank] = B.nt[src];
      } else {
        const int s = src - W;
        X.node[rank] = -1 - s;                                    // resolved below (needs a fresh id)
        X.label[rank] = nw_label[s]; X.par[rank] = B.node[nw_pidx[s]];
        X.nb[rank] = NEG; X.nl[rank] = nw_nt[s]; X.nt[rank] = nw_nt[s];
      }
    }
    __syncwarp();
    // trie identity for the new leaves: find-or-insert (parent node, label); ids handed out in slot order
    for (int i0 = 0; i0 < cnt; i0 += 32) {
      const int i = i0 + lane;
      const bool isnew = (i < cnt) && (X.node[i] < 0);
      const unsigned bal = __ballot_sync(FULL, isnew);
      if (isnew) {
        const int fresh = nodes + __popc(bal & ((1u << lane) - 1));
        const int pnode = X.par[i], c = X.label[i];
        const unsigned kk = (unsigned)pnode * 64u + (unsigned)c + 1u;
        unsigned h = (kk * 2654435761u) & (unsigned)(hash_size - 1);
        int node = fresh;
        for (int probe = 0; probe < hash_size; ++probe) {
          const unsigned long long want = ((unsigned long long)kk << 32) | (unsigned)fresh;
          const unsigned long long old = atomicCAS(&hash[h], 0ull, want);
          if (old == 0ull) break;
          if ((unsigned)(old >> 32) == kk) { node = (int)(old & 0xffffffffu); break; }
          h = (h + 1) & (unsigned)(hash_size - 1);
        }
        if (node == fresh && node < max_nodes) { node_parent[node] = pnode; node_label[node] = c; }
        X.node[i] = node;
      }
      nodes += __popc(bal);
    }
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) slot_of_node[X.node[i]] = i;
    __threadfence_block();
    __syncwarp();
    m = cnt;
    cur ^= 1;
  }

  // best leaf = branch 0 (sorted); walk up the trie
  if (lane == 0) {
    const int node = br[cur].node[0];
    int cnt = 0, prev = -1;
    for (int v = node; v > 0; v = node_parent[v]) {
      const int lab = node_label[v];
      if (!merge_repeated || lab != prev) ++cnt;
      prev = lab;
    }
    int pos = cnt;
    prev = -1;
    for (int v = node; v > 0; v = node_parent[v]) {
      const int lab = node_label[v];
      if (!merge_repeated || lab != prev) outp[--pos] = lab;
      prev = lab;
    }
    out_len[n] = cnt;
    nodes = cnt;
  }
  nodes = __shfl_sync(FULL, nodes, 0);
  for (int i = nodes + lane; i < T; i += 32) outp[i] = -1;
}

inline int next_pow2(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
struct BeamDims { int max_nodes, hash_size; long long ints_per_utt; };
inline BeamDims beam_dims(int T, int W) {
  BeamDims d;
  d.max_nodes = T * W + 2;
  d.hash_size = next_pow2(2LL * d.max_nodes);
  d.ints_per_utt = 3LL * d.max_nodes + (d.max_nodes & 1) + 2LL * d.hash_size;
  d.ints_per_utt = (d.ints_per_utt + 3) & ~3LL;
  return d;
}

}  // namespace

extern "C" size_t asr_ctc_beam_workspace_bytes(int32_t T, int32_t N, int32_t C, int32_t beam_width) {
  if (T <= 0 || N <= 0 || C <= 0 || beam_width <= 0) return 0;
  return (size_t)N * beam_dims(T, beam_width).ints_per_utt * sizeof(int);
}

extern "C" int32_t asr_ctc_beam(const float* logits, int32_t T, int32_t N, int32_t C, const int32_t* in_len,
                                int32_t blank, int32_t beam_width, int32_t merge_repeated, int32_t* out_labels,
                                int32_t* out_len, void* ws, void* stream) {
  ASR_CHECK_ARG(logits && in_len && out_labels && out_len && ws, "asr_ctc_beam: null argument");
  ASR_CHECK_ARG(T >= 1 && N >= 1 && C >= 2 && C <= 64 && blank >= 0 && blank < C, "asr_ctc_beam: bad shape (C <= 64)");
  ASR_CHECK_ARG(beam_width >= 1 && beam_width <= 1024, "asr_ctc_beam: beam_width must be in [1, 1024]");
  const int W = beam_width;
  const BeamDims d = beam_dims(T, W);
  const size_t fl = (size_t)((C + 3) & ~3) + 12 * (size_t)W + W + 2 * W;
  const size_t in = 8 * (size_t)W + 3 * W + 2 * W;
  const size_t smem = (fl + in) * 4 + (((size_t)W * C * 2 + 15) & ~(size_t)15) + 64;
  ASR_CHECK_ARG(smem <= 220 * 1024, "asr_ctc_beam: beam_width %d x %d classes needs %zu B of shared memory", W, C, smem);
  ASR_CUDA(cudaFuncSetAttribute(beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  beam_kernel<<<N, 32, smem, (cudaStream_t)stream>>>(logits, T, N, C, in_len, blank, W, merge_repeated, out_labels,
                                                      out_len, (int*)ws, d.ints_per_utt, d.max_nodes, d.hash_size);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
