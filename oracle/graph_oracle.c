/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C restatement of the reference's relation builders, used as the exact (integer) checker
 * for the CUDA builder at sizes where the dense torch restatement is too slow:
 *
 *   semantics 0: construct_edges_from_states_batch   src/dynamics/dataset/graph.py:91-156
 *   semantics 1: construct_edges_from_states         src/dynamics/dataset/graph.py:38-89
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/graph_cases.npz (outputs of the
 * reference itself).  Compile: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py).
 *
 * Output: relation rows in the reference's order (graph, receiver, sender ascending — the order
 * of adj_matrix.nonzero(), graph.py:151) as flat arrays recv/send (graph-local ids) with
 * per-graph counts n_edges[B].  Returns the total number of relations, or -1 if it exceeds cap.
 * Distance ties inside a row's top-k are broken towards the lower sender id.
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float d; int32_t j; } cand_t;

static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->j > y->j) - (x->j < y->j);
}

int64_t agx_oracle_edges(const float* pos, const uint8_t* mask, const uint8_t* tool, const float* thr2,
                         int32_t B, int32_t N, int32_t topk, int32_t cta, int32_t sem,
                         int32_t* recv, int32_t* send, int64_t cap, int32_t* n_edges) {
  int64_t total = 0;
  if (topk > N) topk = N;                                   /* graph.py:128 */
  uint8_t* adj = (uint8_t*)malloc((size_t)N * N);
  cand_t* row = (cand_t*)malloc(sizeof(cand_t) * (size_t)N);
  if (!adj || !row) { free(adj); free(row); return -2; }
  for (int32_t b = 0; b < B; ++b) {
    const float* p = pos + (size_t)b * N * 3;
    const uint8_t* m = mask + (size_t)b * N;
    const uint8_t* t = tool + (size_t)b * N;
    int probe = 0;
    for (int32_t i = 0; i < N; ++i) {
      for (int32_t j = 0; j < N; ++j) {
        const float dx = p[3 * i] - p[3 * j], dy = p[3 * i + 1] - p[3 * j + 1], dz = p[3 * i + 2] - p[3 * j + 2];
        float d = (dx * dx + dy * dy) + dz * dz;            /* :109-110 */
        if (!(m[i] && m[j])) d = 1e10f;                     /* :111-114 */
        if (t[i] && t[j]) d = 1e10f;                        /* :115-118 */
        row[j].d = d; row[j].j = j;
        adj[(size_t)i * N + j] = (d - thr2[b]) < 0.0f;      /* :125 */
      }
      qsort(row, (size_t)N, sizeof(cand_t), cand_cmp);      /* :129 top-k smallest */
      for (int32_t k = topk; k < N; ++k) adj[(size_t)i * N + row[k].j] = 0;   /* :130-132 */
      if (t[i]) for (int32_t j = 0; j < N; ++j) if (!t[j] && adj[(size_t)i * N + j]) probe = 1;  /* :123,:135 */
    }
    if (cta) {
      for (int32_t i = 0; i < N; ++i) for (int32_t j = 0; j < N; ++j) {
        const int m1 = t[i] && m[j];                        /* obj_tool_mask_1: tool receiver, valid sender */
        const int m2 = t[j] && m[i];                        /* obj_tool_mask_2: tool sender, valid receiver */
        uint8_t* a = &adj[(size_t)i * N + j];
        if (sem == 0) {                                     /* :136-144 */
          if (m1) *a = 0;
          if (m2) *a = probe ? 1 : 0;
        } else {                                            /* :78-80 */
          if (m1) *a = 0;
          if (m2) *a = 1;
          if (t[i] && t[j]) *a = 0;
        }
      }
    }
    int32_t cnt = 0;
    for (int32_t i = 0; i < N; ++i) for (int32_t j = 0; j < N; ++j) if (adj[(size_t)i * N + j]) {
      if (total >= cap) { free(adj); free(row); return -1; }
      recv[total] = i; send[total] = j; ++total; ++cnt;
    }
    n_edges[b] = cnt;
  }
  free(adj); free(row);
  return total;
}
