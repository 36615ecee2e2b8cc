/* Oracle, C restatement of oracle/ctc.py:beam_decode_single — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * The Python restatement of TF 1.3's CTCBeamSearchDecoder (tensorflow/core/util/ctc/ctc_beam_search.h, top path only;
 * call site core/ctc_utils.py:43-50 of the reference) is pinned on TensorFlow's own known-answer test
 * (tests/golden/ctc_tf_beam_known_answer.json) but takes ~20 s per 10 s clip at width 100 in the interpreter.  This file
 * is the same algorithm, statement for statement, in C, so that the label-error-rate parity of BASELINE config 5 can be
 * checked on hundreds of full-length clips; tests/test_oracle_ctc.py holds it to the Python oracle (identical label
 * sequences on random and peaky posteriors, several widths).  float32 scores; exp / log1p evaluated in double and rounded
 * to float, exactly like the Python version ("what a correctly rounded libm returns").
 *
 * Order semantics kept from the Python version (they decide ties): `leaves` is an ordered list; the per-frame visiting
 * order is a STABLE sort by descending score; bottom() is the FIRST minimum, the result the FIRST maximum; a leaf that is
 * pushed out is removed in place.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC oracle/ctc_beam.c -o oracle/_build/libctc_beam_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

typedef struct Entry {
  struct Entry* parent;
  struct Entry** children; /* C - 1 children in label order (blank skipped), or NULL */
  int label;
  float ob, ol, ot, nb, nl, nt;
} Entry;

typedef struct Pool {
  Entry* block;
  size_t used, cap;
  struct Pool* next;
} Pool;

static Entry* pool_new(Pool** head, Entry* parent, int label) {
  Pool* p = *head;
  if (!p || p->used == p->cap) {
    Pool* q = (Pool*)malloc(sizeof(Pool));
    q->cap = 1 << 14;
    q->used = 0;
    q->block = (Entry*)malloc(q->cap * sizeof(Entry));
    q->next = p;
    *head = p = q;
  }
  Entry* e = &p->block[p->used++];
  e->parent = parent;
  e->children = NULL;
  e->label = label;
  e->ob = e->ol = e->ot = e->nb = e->nl = e->nt = NEG_INF;
  return e;
}

static float lse2(float a, float b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  const float m = a >= b ? a : b, mn = a >= b ? b : a;
  const float d = mn - m;
  const float e = (float)exp((double)d);
  return m + (float)log1p((double)e);
}

static int active(const Entry* e) { return e->nt != NEG_INF; }

/* stable merge sort of entry pointers by descending nt */
static void sort_desc(Entry** a, Entry** tmp, int n) {
  if (n < 2) return;
  const int h = n / 2;
  sort_desc(a, tmp, h);
  sort_desc(a + h, tmp, n - h);
  int i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = (a[j]->nt > a[i]->nt) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(Entry*));
}

static int bottom_index(Entry** leaves, int n) {
  int b = 0;
  for (int i = 1; i < n; ++i)
    if (leaves[i]->nt < leaves[b]->nt) b = i;
  return b;
}

/* logits: [T, C] row-major float32.  out: label ids, capacity T.  returns the number of labels. */
int ctc_beam_oracle_single(const float* logits, int T, int C, int seq_len, int blank, int beam_width, int merge_repeated,
                           int32_t* out) {
  Pool* pool = NULL;
  Entry* root = pool_new(&pool, NULL, -1);
  root->nt = 0.0f;
  root->nb = 0.0f;
  const int W = beam_width;
  Entry** leaves = (Entry**)malloc((size_t)(W + 1) * sizeof(Entry*));
  Entry** branches = (Entry**)malloc((size_t)(W + 1) * sizeof(Entry*));
  Entry** tmp = (Entry**)malloc((size_t)(W + 1) * sizeof(Entry*));
  float* inp = (float*)malloc((size_t)C * sizeof(float));
  int nleaves = 1;
  leaves[0] = root;
  if (seq_len > T) seq_len = T;
  for (int t = 0; t < seq_len; ++t) {
    const float* row = logits + (size_t)t * C;
    float mx = row[0];
    for (int k = 1; k < C; ++k)
      if (row[k] > mx) mx = row[k];
    for (int k = 0; k < C; ++k) inp[k] = row[k] - mx;
    const int nb = nleaves;
    memcpy(branches, leaves, (size_t)nb * sizeof(Entry*));
    sort_desc(branches, tmp, nb);
    for (int i = 0; i < nb; ++i) {
      Entry* b = branches[i];
      b->ob = b->nb; b->ol = b->nl; b->ot = b->nt;
    }
    nleaves = 0;
    for (int i = 0; i < nb; ++i) {
      Entry* b = branches[i];
      if (b->parent) {
        if (active(b->parent)) {
          const float prev = (b->label == b->parent->label) ? b->parent->ob : b->parent->ot;
          b->nl = lse2(b->nl, prev);
        }
        b->nl = b->nl + inp[b->label];
      }
      b->nb = b->ot + inp[blank];
      b->nt = lse2(b->nb, b->nl);
      leaves[nleaves++] = b;
    }
    for (int i = 0; i < nb; ++i) {
      Entry* b = branches[i];
      if (!(b->ot > NEG_INF && (nleaves < W || b->ot > leaves[bottom_index(leaves, nleaves)]->nt))) continue;
      if (!b->children) {
        b->children = (Entry**)malloc((size_t)(C - 1) * sizeof(Entry*));
        int j = 0;
        for (int k = 0; k < C; ++k)
          if (k != blank) b->children[j++] = pool_new(&pool, b, k);
      }
      for (int j = 0; j < C - 1; ++j) {
        Entry* c = b->children[j];
        if (active(c)) continue;
        const float prev = (c->label == b->label) ? b->ob : b->ot;
        c->nb = NEG_INF;
        c->nl = (prev != NEG_INF) ? inp[c->label] + prev : NEG_INF;
        c->nt = c->nl;
        if (c->nt > NEG_INF && (nleaves < W || c->nt > leaves[bottom_index(leaves, nleaves)]->nt)) {
          if (nleaves == W) {
            const int wi = bottom_index(leaves, nleaves);
            Entry* worst = leaves[wi];
            memmove(leaves + wi, leaves + wi + 1, (size_t)(nleaves - wi - 1) * sizeof(Entry*));
            --nleaves;
            worst->nb = worst->nl = worst->nt = NEG_INF;
          }
          leaves[nleaves++] = c;
        } else {
          c->ob = c->ol = c->ot = NEG_INF;
          c->nb = c->nl = c->nt = NEG_INF;
        }
      }
    }
  }
  int bi = 0;
  for (int i = 1; i < nleaves; ++i)
    if (leaves[i]->nt > leaves[bi]->nt) bi = i;
  int n = 0, prev = -1;
  for (Entry* e = leaves[bi]; e->parent; e = e->parent) {
    if (!merge_repeated || e->label != prev) out[n++] = e->label;
    prev = e->label;
  }
  for (int i = 0; i < n / 2; ++i) {
    const int32_t v = out[i];
    out[i] = out[n - 1 - i];
    out[n - 1 - i] = v;
  }
  /* children arrays live outside the pool */
  for (Pool* p = pool; p;) {
    for (size_t i = 0; i < p->used; ++i) free(p->block[i].children);
    Pool* nx = p->next;
    free(p->block);
    free(p);
    p = nx;
  }
  free(leaves); free(branches); free(tmp); free(inp);
  return n;
}

/* logits: [N, T, C] batch-major; out: [N, T] (-1 padded); out_len: [N].  Utterances in parallel (OpenMP). */
void ctc_beam_oracle_batch(const float* logits, int N, int T, int C, const int32_t* seq_len, int blank, int beam_width,
                           int merge_repeated, int32_t* out, int32_t* out_len) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < N; ++n) {
    int32_t* o = out + (size_t)n * T;
    const int k = ctc_beam_oracle_single(logits + (size_t)n * T * C, T, C, seq_len[n], blank, beam_width, merge_repeated, o);
    for (int i = k; i < T; ++i) o[i] = -1;
    out_len[n] = k;
  }
}
