// K8 — CTC prefix beam search (top path) for sm_100a.
// Replaces tf.nn.ctc_beam_search_decoder(y_pred, seq_len, beam_width, top_paths=1, merge_repeated)[0][0]
// (core/ctc_utils.py:44-50; width 100 by default, 400 via utils/core_utils.py:70-71).
//
// One CTA per utterance; per frame, in parallel over the beam:
//   1. scores = logits - max(logits) (TF 1.3 ctc_beam_search.h Step()); every leaf updates its
//      (blank, label, total) mass from itself and from its parent if the parent is still in the beam;
//   2. every (leaf, label) pair whose child is not already in the beam proposes a candidate
//      total = score[label] + (label == leaf.label ? leaf.blank_old : leaf.total_old);
//   3. leaves + candidates are ranked with one shared-memory bitonic sort on 64-bit keys
//      (order-preserving float bits << 32 | tie-break), the best `beam_width` survive.
// TF inserts candidates one by one against a moving threshold; because a child never outscores its
// parent's old total and the threshold only rises, that procedure selects exactly the global top-W
// (ties aside: here ties prefer existing leaves, then lower beam slot / label).
// Trie identity follows TF (a prefix that leaves the beam and comes back is the SAME node, so its
// children see it as an active parent again): nodes are found through a per-utterance open-addressing
// hash keyed by (parent node, label) in the L2-resident workspace.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int BEAM_THREADS = 512;

__device__ __forceinline__ float lse2f(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return m;
  return m + log1pf(expf(fminf(a, b) - m));
}
__device__ __forceinline__ uint32_t ord_bits(float f) {      // monotone float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// per-utterance global workspace (ints): node_parent[max_nodes], node_label[max_nodes],
// slot_of_node[max_nodes] (current beam slot or -1), then the hash table (u64: key in the high 32
// bits, 0 = empty; node id in the low 32).

template <int SORT_N>
__global__ void __launch_bounds__(BEAM_THREADS)
beam_kernel(const float* __restrict__ logits, int T, int N, int C, const int* __restrict__ in_len, int blank, int W,
            int merge_repeated, int* __restrict__ out_labels, int* __restrict__ out_len, int* __restrict__ ws,
            long long ws_ints_per_utt, int max_nodes, int hash_size) {
  extern __shared__ unsigned long long sm64[];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int len = min(max(in_len[n], 0), T);
  int* base = ws + (size_t)n * ws_ints_per_utt;
  int* node_parent = base;
  int* node_label = base + max_nodes;
  int* slot_of_node = base + 2 * (size_t)max_nodes;
  unsigned long long* hash = reinterpret_cast<unsigned long long*>(base + 3 * (size_t)max_nodes + (max_nodes & 1));

  unsigned long long* keys = sm64;                               // [SORT_N]
  float* inp = reinterpret_cast<float*>(keys + SORT_N);          // [C]
  // beam state, double buffered: node, label, parent_node, blank, label-mass, total
  int* st_i = reinterpret_cast<int*>(inp + ((C + 1) & ~1));      // [2][3][W]
  float* st_f = reinterpret_cast<float*>(st_i + 2 * 3 * W);      // [2][3][W]
  float* nw_f = st_f + 2 * 3 * W;                                // new probs of existing leaves [3][W]
  unsigned* childmask = reinterpret_cast<unsigned*>(nw_f + 3 * W);   // [2][W]  (64 labels)
  int* sel_old = reinterpret_cast<int*>(childmask + 2 * W);      // [W] new slot of each old slot, or -1
  int* scan = sel_old + W;                                       // [W + 1]
  __shared__ int s_nb, s_nodes;

  int* outp = out_labels + (size_t)n * T;
  for (int i = tid; i < hash_size; i += BEAM_THREADS) hash[i] = 0ull;
  if (tid == 0) {
    s_nb = 1;
    s_nodes = 1;
    node_parent[0] = -1; node_label[0] = -1; slot_of_node[0] = 0;
    st_i[0] = 0; st_i[W] = -1; st_i[2 * W] = -1;                 // root: node 0, label -1, parent -1
    st_f[0] = 0.0f; st_f[W] = -CUDART_INF_F; st_f[2 * W] = 0.0f; // blank = 0, label = -inf, total = 0
  }
  __syncthreads();
  int cur = 0;
  const float NEG = -CUDART_INF_F;

  for (int t = 0; t < len; ++t) {
    const int nb = s_nb;
    int* c_node = st_i + cur * 3 * W; int* c_label = c_node + W; int* c_par = c_label + W;
    float* c_pb = st_f + cur * 3 * W; float* c_pl = c_pb + W; float* c_pt = c_pl + W;
    int* x_node = st_i + (cur ^ 1) * 3 * W; int* x_label = x_node + W; int* x_par = x_label + W;
    float* x_pb = st_f + (cur ^ 1) * 3 * W; float* x_pl = x_pb + W; float* x_pt = x_pl + W;
    float* n_pb = nw_f; float* n_pl = nw_f + W; float* n_pt = nw_f + 2 * W;

    // 0. frame scores (minus max), reset per-step tables
    if (tid < 32) {
      const float* row = logits + ((size_t)t * N + n) * C;
      float m = NEG;
      for (int k = tid; k < C; k += 32) m = fmaxf(m, row[k]);
      m = asr::warp_max(m);
      for (int k = tid; k < C; k += 32) inp[k] = row[k] - m;
    }
    for (int i = tid; i < 2 * W; i += BEAM_THREADS) childmask[i] = 0u;
    for (int i = tid; i < W; i += BEAM_THREADS) sel_old[i] = -1;
    __syncthreads();

    // 1. existing leaves: parent lookup, child masks, new masses, sort keys
    for (int s = tid; s < W; s += BEAM_THREADS) {
      unsigned long long key = 0ull;
      if (s < nb) {
        const int lab = c_label[s], par = c_par[s];
        int ps = -1;
        if (par >= 0) ps = slot_of_node[par];
        float nl = c_pl[s];
        if (lab >= 0) {
          if (ps >= 0) {
            atomicOr(&childmask[2 * ps + (lab >> 5)], 1u << (lab & 31));
            const float prev = (lab == c_label[ps]) ? c_pb[ps] : c_pt[ps];
            nl = lse2f(nl, prev);
          }
          nl += inp[lab];
        }
        const float nbk = c_pt[s] + inp[blank];
        const float nt = lse2f(nbk, nl);
        n_pb[s] = nbk; n_pl[s] = nl; n_pt[s] = nt;
        key = ((unsigned long long)ord_bits(nt) << 32) | (unsigned)(0x7fffffff - s);
      }
      keys[s] = key;
    }
    __syncthreads();
    // 2. candidate children
    for (int idx = tid; idx < SORT_N - W; idx += BEAM_THREADS) {
      unsigned long long key = 0ull;
      const int p = idx / C, c = idx - p * C;
      if (p < nb && c != blank && !((childmask[2 * p + (c >> 5)] >> (c & 31)) & 1u)) {
        const float prev = (c == c_label[p]) ? c_pb[p] : c_pt[p];
        if (prev > NEG) {
          const float tot = inp[c] + prev;
          key = ((unsigned long long)ord_bits(tot) << 32) | (unsigned)(0x3fffffff - idx);
        }
      }
      keys[W + idx] = key;
    }
    __syncthreads();
    // 3. bitonic sort, descending
    for (int k = 2; k <= SORT_N; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < SORT_N; i += BEAM_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = keys[i], b = keys[ixj];
            const bool desc = ((i & k) == 0);
            if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    // 4. survivors -> next beam
    int is_new = 0, valid = 0, payload = 0;
    if (tid < W) {
      const unsigned long long key = keys[tid];
      valid = key != 0ull;
      const unsigned lo = (unsigned)(key & 0xffffffffu);
      if (valid) {
        if (lo > 0x3fffffffu) { payload = 0x7fffffff - (int)lo; is_new = 0; sel_old[payload] = tid; }
        else { payload = 0x3fffffff - (int)lo; is_new = 1; }
      }
    }
    // exclusive scan of is_new over the first W threads (W <= BEAM_THREADS)
    if (tid <= W) scan[tid] = 0;
    __syncthreads();
    if (tid < W) scan[tid + 1] = is_new;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i <= W; ++i) scan[i] += scan[i - 1];
    }
    __syncthreads();
    const int nodes0 = s_nodes;
    if (tid < W && valid) {
      if (!is_new) {
        const int s = payload;
        x_node[tid] = c_node[s]; x_label[tid] = c_label[s]; x_par[tid] = c_par[s];
        x_pb[tid] = n_pb[s]; x_pl[tid] = n_pl[s]; x_pt[tid] = n_pt[s];
      } else {
        const int p = payload / C, c = payload - p * C;
        const int pnode = c_node[p];
        const float prev = (c == c_label[p]) ? c_pb[p] : c_pt[p];
        const float tot = inp[c] + prev;
        // trie identity: find-or-insert (pnode, c)
        const unsigned kk = (unsigned)pnode * 64u + (unsigned)c + 1u;
        unsigned h = (kk * 2654435761u) & (unsigned)(hash_size - 1);
        int node = -1;
        const int fresh = nodes0 + scan[tid];
        for (int probe = 0; probe < hash_size; ++probe) {
          const unsigned long long want = ((unsigned long long)kk << 32) | (unsigned)fresh;
          const unsigned long long old = atomicCAS(&hash[h], 0ull, want);
          if (old == 0ull) { node = fresh; break; }
          if ((unsigned)(old >> 32) == kk) { node = (int)(old & 0xffffffffu); break; }
          h = (h + 1) & (unsigned)(hash_size - 1);
        }
        if (node == fresh && node < max_nodes) { node_parent[node] = pnode; node_label[node] = c; }
        x_node[tid] = node; x_label[tid] = c; x_par[tid] = pnode;
        x_pb[tid] = NEG; x_pl[tid] = tot; x_pt[tid] = tot;
      }
    }
    __syncthreads();
    // evictions, then new slot table
    for (int s = tid; s < nb; s += BEAM_THREADS)
      if (sel_old[s] < 0) slot_of_node[c_node[s]] = -1;
    __syncthreads();
    if (tid < W && valid) slot_of_node[x_node[tid]] = tid;
    if (tid == 0) {
      int cnt = 0;
      for (int i = 0; i < W; ++i) cnt += (keys[i] != 0ull);
      s_nb = cnt;
      s_nodes = nodes0 + scan[W];
    }
    __threadfence_block();
    __syncthreads();
    cur ^= 1;
  }

  // best leaf = slot 0 after the sort (for len == 0 the root): walk up the trie
  if (tid == 0) {
    int* c_node = st_i + cur * 3 * W;
    float* c_pt = st_f + cur * 3 * W + 2 * W;
    int best = 0;
    for (int s = 1; s < s_nb; ++s)
      if (c_pt[s] > c_pt[best]) best = s;
    int node = c_node[best];
    int cnt = 0, prev = -1;
    // first pass: count (after merge), second pass: write reversed
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
    s_nb = cnt;
  }
  __syncthreads();
  for (int i = s_nb + tid; i < T; i += BEAM_THREADS) outp[i] = -1;
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
  ASR_CHECK_ARG(beam_width >= 1 && beam_width <= BEAM_THREADS, "asr_ctc_beam: beam_width must be in [1, %d]", BEAM_THREADS);
  const int W = beam_width;
  const long long need = (long long)W + (long long)W * C;
  const int sort_n = next_pow2(need) < 512 ? 512 : next_pow2(need);
  ASR_CHECK_ARG(sort_n <= 16384, "asr_ctc_beam: beam_width * num_classes too large (%lld)", need);
  const BeamDims d = beam_dims(T, W);
  const size_t smem = (size_t)sort_n * 8 + (size_t)((C + 1) & ~1) * 4 + (size_t)(2 * 3 * W) * 4 * 2 + (size_t)3 * W * 4 +
                      (size_t)2 * W * 4 + (size_t)W * 4 + (size_t)(W + 2) * 4 + 64;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kern) -> int32_t {
    ASR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<N, BEAM_THREADS, smem, st>>>(logits, T, N, C, in_len, blank, W, merge_repeated, out_labels, out_len,
                                        (int*)ws, d.ints_per_utt, d.max_nodes, d.hash_size);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
  };
  switch (sort_n) {
    case 16384: return launch(beam_kernel<16384>);
    case 8192: return launch(beam_kernel<8192>);
    case 4096: return launch(beam_kernel<4096>);
    case 2048: return launch(beam_kernel<2048>);
    case 1024: return launch(beam_kernel<1024>);
    default: return launch(beam_kernel<512>);
  }
}
