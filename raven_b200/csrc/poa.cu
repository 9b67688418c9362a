// raven_b200 — windowed partial-order-alignment consensus on sm_100a.
//
// Replaces the per-window consensus of racon::Polisher::Polish (un-vendored;
// the reference calls it at RavenLib/src/polish.cc:43-51 with w = 500,
// trim = true, m/n/g from PolishCfg polish.hpp:13-17): racon::Window::
// GenerateConsensus over spoa's Graph / AlignmentEngine (global alignment,
// linear gaps). Semantics restated in SURVEY.md App. A.4-A.5 and, as the
// parity oracle, in oracle/spoa_graph.cpp + oracle/racon_window.cpp.
//
// One warp per window, many windows in flight (the graph surgery between two
// layers is sequential pointer work; occupancy hides its latency):
//   * lane 0 keeps the partial order graph (nodes, in/out edge lists in
//     insertion order, aligned-node sets, per-node sequence counts) in a
//     per-window scratch slab, extracts the sub-graph a partial layer aligns
//     to, and sorts it topologically exactly like spoa's DFS;
//   * all 32 lanes fill the DP rows: coalesced int16 rows, the diagonal
//     neighbour through a shuffle, and the horizontal gap recurrence
//     H[i][j] = max(M[j], H[i][j-1] + g) solved as a warp max-scan of
//     M[j] - j*g (exact in integers);
//   * lane 0 walks the traceback with spoa's preference (diagonal over
//     predecessors in in-edge order, then vertical, then horizontal), merges
//     the layer into the graph, and finally runs the heaviest-bundle consensus
//     with branch completion and the coverage-based TGS trimming.
// A node's Coverage() (distinct sequence labels on its incident edges) equals
// the number of sequences (of >= 2 bases) whose path visits it, so a per-node
// counter replaces spoa's per-edge label lists.
// No tensor-core work exists here (no dense contraction).
#include <algorithm>

#include "engine.cuh"

namespace rvn {

namespace {

constexpr uint16_t kNone = 0xFFFF;
constexpr int kPoaStatusOk = 1;        // polished
constexpr int kPoaStatusChimeric = 2;  // trimming skipped (racon's warning)
constexpr int kPoaStatusCapacity = 64; // scratch too small: host retries bigger
constexpr int kPoaStatusInvalid = 128; // bad layer coordinates / letters

struct PoaShape {
  uint32_t ncap;   // node capacity
  uint32_t ecap;   // edge capacity
  uint32_t lmax;   // longest layer
  uint32_t rows;   // DP rows = ncap + 1
  uint32_t width;  // DP row stride (lmax + 1 rounded up to 32)
};

// byte offsets of the arrays inside one window's scratch slab
struct PoaLayout {
  size_t code, n_aligned, marks, ignored, in_sub, sink;
  size_t in_head, in_tail, out_head, aligned, cov, r2n, n2r, stack, order;
  size_t e_tail, e_head, e_next_in, e_next_out, e_weight;
  size_t aln_node, aln_pos, pred, score, H, bytes;
};

__host__ __device__ inline PoaLayout MakePoaLayout(const PoaShape& s) {
  PoaLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 15) & ~size_t(15);
    return at;
  };
  L.code = take(s.ncap);
  L.n_aligned = take(s.ncap);
  L.marks = take(s.ncap);
  L.ignored = take(s.ncap);
  L.in_sub = take(s.ncap);
  L.sink = take(s.ncap);
  L.in_head = take(2ULL * s.ncap);
  L.in_tail = take(2ULL * s.ncap);
  L.out_head = take(2ULL * s.ncap);
  L.aligned = take(2ULL * 3 * s.ncap);
  L.cov = take(2ULL * s.ncap);
  L.r2n = take(2ULL * s.ncap);
  L.n2r = take(2ULL * s.ncap);
  L.stack = take(2ULL * 4 * s.ncap);
  L.order = take(2ULL * 4096);
  L.e_tail = take(2ULL * s.ecap);
  L.e_head = take(2ULL * s.ecap);
  L.e_next_in = take(2ULL * s.ecap);
  L.e_next_out = take(2ULL * s.ecap);
  L.e_weight = take(4ULL * s.ecap);
  L.aln_node = take(2ULL * (s.ncap + s.lmax + 2));
  L.aln_pos = take(2ULL * (s.ncap + s.lmax + 2));
  L.pred = take(4ULL * s.ncap);
  L.score = take(8ULL * s.ncap);
  L.H = take(2ULL * s.rows * s.width);
  L.bytes = o;
  return L;
}

struct PoaGraph {
  uint8_t *code, *n_aligned, *marks, *ignored, *in_sub, *sink;
  uint16_t *in_head, *in_tail, *out_head, *aligned, *cov, *r2n, *n2r, *stack, *order;
  uint16_t *e_tail, *e_head, *e_next_in, *e_next_out;
  int32_t* e_weight;
  int16_t *aln_node, *aln_pos;
  int32_t* pred;
  long long* score;
  int16_t* H;
  uint32_t n_nodes, n_edges, ncap, ecap, stack_cap;
  bool overflow;
};

__device__ __forceinline__ int CodeOf(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return -1;
  }
}

__device__ uint16_t AddNode(PoaGraph& g, uint32_t code) {
  if (g.n_nodes >= g.ncap) {
    g.overflow = true;
    return 0;
  }
  const uint32_t v = g.n_nodes++;
  g.code[v] = static_cast<uint8_t>(code);
  g.n_aligned[v] = 0;
  g.in_head[v] = g.in_tail[v] = g.out_head[v] = kNone;
  g.cov[v] = 0;
  return static_cast<uint16_t>(v);
}

// spoa Graph::AddEdge: bump the weight of an existing edge or append a new
// one at the END of both adjacency lists (insertion order is observable)
__device__ void AddEdge(PoaGraph& g, uint16_t tail, uint16_t head, int32_t w) {
  uint16_t last_out = kNone;
  for (uint16_t e = g.out_head[tail]; e != kNone; e = g.e_next_out[e]) {
    if (g.e_head[e] == head) {
      g.e_weight[e] += w;
      return;
    }
    last_out = e;
  }
  if (g.n_edges >= g.ecap) {
    g.overflow = true;
    return;
  }
  const uint16_t e = static_cast<uint16_t>(g.n_edges++);
  g.e_tail[e] = tail;
  g.e_head[e] = head;
  g.e_weight[e] = w;
  g.e_next_in[e] = g.e_next_out[e] = kNone;
  if (last_out == kNone) {
    g.out_head[tail] = e;
  } else {
    g.e_next_out[last_out] = e;
  }
  if (g.in_head[head] == kNone) {
    g.in_head[head] = e;
  } else {
    g.e_next_in[g.in_tail[head]] = e;
  }
  g.in_tail[head] = e;
}

struct SeqView {
  const uint8_t* bases;
  const uint8_t* quals;  // nullptr = constant weight
  uint32_t len;
  int32_t flat;          // weight without qualities: layers 1, backbone 0
  __device__ __forceinline__ int32_t weight(uint32_t i) const {
    return quals ? static_cast<int32_t>(quals[i]) - 33 : flat;
  }
};

// a chain of fresh nodes for s[begin, end); returns its first node or kNone
__device__ uint16_t AddChain(PoaGraph& g, const SeqView& s, uint32_t begin,
                             uint32_t end) {
  if (begin == end) return kNone;
  uint16_t first = kNone, prev = kNone;
  for (uint32_t i = begin; i < end && !g.overflow; ++i) {
    const uint16_t cur = AddNode(g, CodeOf(s.bases[i]));
    if (prev != kNone) AddEdge(g, prev, cur, s.weight(i - 1) + s.weight(i));
    if (first == kNone) first = cur;
    prev = cur;
  }
  return first;
}

// spoa Graph::TopologicalSort restricted to the nodes with in_sub != 0
// (Graph::Subgraph keeps relative ids and adjacency order, so the restricted
// DFS over the full graph yields exactly the sub-graph's order)
__device__ uint32_t TopoSort(PoaGraph& g) {
  const uint32_t n = g.n_nodes;
  for (uint32_t v = 0; v < n; ++v) {
    g.marks[v] = 0;
    g.ignored[v] = 0;
  }
  uint32_t rank = 0, sp = 0;
  for (uint32_t root = 0; root < n; ++root) {
    if (!g.in_sub[root] || g.marks[root] != 0) continue;
    g.stack[sp++] = static_cast<uint16_t>(root);
    while (sp > 0) {
      const uint16_t cur = g.stack[sp - 1];
      bool valid = true;
      if (g.marks[cur] != 2) {
        for (uint16_t e = g.in_head[cur]; e != kNone; e = g.e_next_in[e]) {
          const uint16_t t = g.e_tail[e];
          if (g.in_sub[t] && g.marks[t] != 2) {
            if (sp >= g.stack_cap) {
              g.overflow = true;
              return 0;
            }
            g.stack[sp++] = t;
            valid = false;
          }
        }
        if (!g.ignored[cur]) {
          for (uint32_t a = 0; a < g.n_aligned[cur]; ++a) {
            const uint16_t t = g.aligned[3 * cur + a];
            if (g.in_sub[t] && g.marks[t] != 2) {
              if (sp >= g.stack_cap) {
                g.overflow = true;
                return 0;
              }
              g.stack[sp++] = t;
              g.ignored[t] = 1;
              valid = false;
            }
          }
        }
        if (valid) {
          g.marks[cur] = 2;
          if (!g.ignored[cur]) {
            g.n2r[cur] = static_cast<uint16_t>(rank);
            g.r2n[rank++] = cur;
            for (uint32_t a = 0; a < g.n_aligned[cur]; ++a) {
              const uint16_t t = g.aligned[3 * cur + a];
              if (g.in_sub[t]) {
                g.n2r[t] = static_cast<uint16_t>(rank);
                g.r2n[rank++] = t;
              }
            }
          }
        } else {
          g.marks[cur] = 1;
        }
      }
      if (valid) --sp;
    }
  }
  return rank;
}

// spoa Graph::ExtractSubgraph(nodes[end], nodes[begin]): backward reachability
// from `from` through in-edges and aligned nodes, keeping ids >= min_id
__device__ void MarkSubgraph(PoaGraph& g, uint32_t from, uint32_t min_id) {
  for (uint32_t v = 0; v < g.n_nodes; ++v) g.in_sub[v] = 0;
  uint32_t sp = 0;
  g.stack[sp++] = static_cast<uint16_t>(from);
  while (sp > 0) {
    const uint16_t cur = g.stack[--sp];
    if (g.in_sub[cur] || cur < min_id) continue;
    for (uint16_t e = g.in_head[cur]; e != kNone; e = g.e_next_in[e]) {
      if (sp >= g.stack_cap) {
        g.overflow = true;
        return;
      }
      g.stack[sp++] = g.e_tail[e];
    }
    for (uint32_t a = 0; a < g.n_aligned[cur]; ++a) {
      if (sp >= g.stack_cap) {
        g.overflow = true;
        return;
      }
      g.stack[sp++] = g.aligned[3 * cur + a];
    }
    g.in_sub[cur] = 1;
  }
}

// spoa Graph::AddAlignment (alignment non-empty)
__device__ void AddAlignment(PoaGraph& g, const SeqView& s, uint32_t aln_len) {
  // valid = sequence positions that appear in the alignment
  int32_t first_pos = -1, last_pos = -1;
  for (uint32_t i = 0; i < aln_len; ++i) {
    if (g.aln_pos[i] != -1) {
      if (first_pos < 0) first_pos = g.aln_pos[i];
      last_pos = g.aln_pos[i];
    }
  }
  const uint32_t before = g.n_nodes;
  uint16_t begin = AddChain(g, s, 0, first_pos);
  uint16_t prev = before == g.n_nodes ? kNone : static_cast<uint16_t>(g.n_nodes - 1);
  const uint16_t last = AddChain(g, s, last_pos + 1, s.len);
  // count this sequence on the chains (coverage = sequences through a node)
  const bool counts = s.len >= 2;
  if (counts) {
    for (uint32_t v = before; v < g.n_nodes; ++v) g.cov[v] += 1;
  }
  for (uint32_t i = 0; i < aln_len && !g.overflow; ++i) {
    const int32_t pos = g.aln_pos[i];
    if (pos == -1) continue;
    const uint32_t code = CodeOf(s.bases[pos]);
    uint16_t cur = kNone;
    if (g.aln_node[i] == -1) {
      cur = AddNode(g, code);
    } else {
      const uint16_t jt = static_cast<uint16_t>(g.aln_node[i]);
      if (g.code[jt] == code) {
        cur = jt;
      } else {
        for (uint32_t a = 0; a < g.n_aligned[jt]; ++a) {
          const uint16_t kt = g.aligned[3 * jt + a];
          if (g.code[kt] == code) {
            cur = kt;
            break;
          }
        }
        if (cur == kNone) {  // a new letter for this column
          cur = AddNode(g, code);
          if (g.overflow) break;
          const uint32_t na = g.n_aligned[jt];
          for (uint32_t a = 0; a < na; ++a) {
            const uint16_t kt = g.aligned[3 * jt + a];
            g.aligned[3 * kt + g.n_aligned[kt]++] = cur;
            g.aligned[3 * cur + g.n_aligned[cur]++] = kt;
          }
          g.aligned[3 * jt + g.n_aligned[jt]++] = cur;
          g.aligned[3 * cur + g.n_aligned[cur]++] = jt;
        }
      }
    }
    if (g.overflow) break;
    if (counts) g.cov[cur] += 1;
    if (begin == kNone) begin = cur;
    if (prev != kNone) AddEdge(g, prev, cur, s.weight(pos - 1) + s.weight(pos));
    prev = cur;
  }
  if (last != kNone && !g.overflow) {
    AddEdge(g, prev, last, s.weight(last_pos) + s.weight(last_pos + 1));
  }
}

__global__ void __launch_bounds__(32)
PoaKernel(uint32_t n_windows, const uint32_t* __restrict__ win_list,
          const uint32_t* __restrict__ win_first, const uint64_t* __restrict__ seq_off,
          const uint8_t* __restrict__ bases, const uint8_t* __restrict__ quals,
          const uint32_t* __restrict__ seq_begin, const uint32_t* __restrict__ seq_end,
          int m, int n, int gap, int trim, int tgs, PoaShape shape,
          uint8_t* __restrict__ scratch, size_t scratch_stride,
          uint8_t* __restrict__ cons, const uint64_t* __restrict__ cons_off,
          uint32_t* __restrict__ cons_len, uint32_t* __restrict__ cov_out,
          uint8_t* __restrict__ status, unsigned long long* __restrict__ cells) {
  if (blockIdx.x >= n_windows) return;
  const uint32_t lane = threadIdx.x;
  const uint32_t w = win_list ? win_list[blockIdx.x] : blockIdx.x;
  const uint32_t s0 = win_first[w], s1 = win_first[w + 1];
  const uint32_t nseq = s1 - s0;
  const uint32_t L0 = static_cast<uint32_t>(seq_off[s0 + 1] - seq_off[s0]);
  uint8_t* out = cons + cons_off[w];
  uint32_t* cov_dst = cov_out ? cov_out + cons_off[w] : nullptr;

  if (nseq < 3) {  // fewer than 3 sequences: the backbone, "unpolished"
    for (uint32_t i = lane; i < L0; i += 32) {
      out[i] = bases[seq_off[s0] + i];
      if (cov_dst) cov_dst[i] = 0;
    }
    if (lane == 0) {
      cons_len[w] = L0;
      status[w] = 0;
    }
    return;
  }

  // only ACGT(U) letters are valid (racon feeds 2-bit reads)
  {
    bool bad_letter = false;
    for (uint64_t i = seq_off[s0] + lane; i < seq_off[s1]; i += 32) {
      if (CodeOf(bases[i]) < 0) bad_letter = true;
    }
    if (__any_sync(0xffffffffu, bad_letter)) {
      if (lane == 0) {
        cons_len[w] = 0;
        status[w] = kPoaStatusInvalid;
      }
      return;
    }
  }

  uint8_t* base = scratch + scratch_stride * blockIdx.x;
  const PoaLayout L = MakePoaLayout(shape);
  __shared__ PoaGraph g;
  __shared__ uint32_t sh_rows, sh_fail;
  if (lane == 0) {
    g.code = base + L.code;
    g.n_aligned = base + L.n_aligned;
    g.marks = base + L.marks;
    g.ignored = base + L.ignored;
    g.in_sub = base + L.in_sub;
    g.sink = base + L.sink;
    g.in_head = reinterpret_cast<uint16_t*>(base + L.in_head);
    g.in_tail = reinterpret_cast<uint16_t*>(base + L.in_tail);
    g.out_head = reinterpret_cast<uint16_t*>(base + L.out_head);
    g.aligned = reinterpret_cast<uint16_t*>(base + L.aligned);
    g.cov = reinterpret_cast<uint16_t*>(base + L.cov);
    g.r2n = reinterpret_cast<uint16_t*>(base + L.r2n);
    g.n2r = reinterpret_cast<uint16_t*>(base + L.n2r);
    g.stack = reinterpret_cast<uint16_t*>(base + L.stack);
    g.order = reinterpret_cast<uint16_t*>(base + L.order);
    g.e_tail = reinterpret_cast<uint16_t*>(base + L.e_tail);
    g.e_head = reinterpret_cast<uint16_t*>(base + L.e_head);
    g.e_next_in = reinterpret_cast<uint16_t*>(base + L.e_next_in);
    g.e_next_out = reinterpret_cast<uint16_t*>(base + L.e_next_out);
    g.e_weight = reinterpret_cast<int32_t*>(base + L.e_weight);
    g.aln_node = reinterpret_cast<int16_t*>(base + L.aln_node);
    g.aln_pos = reinterpret_cast<int16_t*>(base + L.aln_pos);
    g.pred = reinterpret_cast<int32_t*>(base + L.pred);
    g.score = reinterpret_cast<long long*>(base + L.score);
    g.H = reinterpret_cast<int16_t*>(base + L.H);
    g.n_nodes = g.n_edges = 0;
    g.ncap = shape.ncap;
    g.ecap = shape.ecap;
    g.stack_cap = 4 * shape.ncap;
    g.overflow = false;
    sh_fail = 0;

    // letters + layer coordinates are validated up front
    bool bad = nseq - 1 > 4095;
    for (uint32_t s = s0; s < s1 && !bad; ++s) {
      const uint32_t len = static_cast<uint32_t>(seq_off[s + 1] - seq_off[s]);
      if (s > s0 && (seq_begin[s] >= seq_end[s] || seq_end[s] >= L0)) bad = true;
      if (s > s0 && len > shape.lmax) bad = true;
    }
    if (bad) sh_fail = kPoaStatusInvalid;

    if (!bad) {
      // backbone chain (spoa AddAlignment with an empty alignment)
      // racon gives the backbone a dummy quality of '!' (weight 0) in any case
      SeqView bb{bases + seq_off[s0], quals ? quals + seq_off[s0] : nullptr, L0, 0};
      AddChain(g, bb, 0, L0);
      if (L0 >= 2) {
        for (uint32_t v = 0; v < g.n_nodes; ++v) g.cov[v] += 1;
      }
      // layers by begin position, stable
      for (uint32_t i = 0; i + 1 < nseq; ++i) g.order[i] = static_cast<uint16_t>(i + 1);
      for (uint32_t a = 1; a + 1 < nseq; ++a) {
        const uint16_t v = g.order[a];
        uint32_t b = a;
        while (b > 0 && seq_begin[s0 + g.order[b - 1]] > seq_begin[s0 + v]) {
          g.order[b] = g.order[b - 1];
          --b;
        }
        g.order[b] = v;
      }
    }
  }
  __syncwarp();

  const uint32_t offset = static_cast<uint32_t>(0.01 * L0);
  const uint32_t W = shape.width;
  unsigned long long my_cells = 0;

  for (uint32_t li = 0; li + 1 < nseq; ++li) {
    if (sh_fail || g.overflow) break;
    const uint32_t s = s0 + g.order[li];
    const SeqView sv{bases + seq_off[s], quals ? quals + seq_off[s] : nullptr,
                     static_cast<uint32_t>(seq_off[s + 1] - seq_off[s]), 1};
    const uint32_t len = sv.len;
    if (len == 0) continue;

    // ---- (lane 0) the graph this layer aligns to, in topological order ----
    if (lane == 0) {
      const uint32_t lb = seq_begin[s], le = seq_end[s];
      if (lb < offset && le > L0 - offset) {
        for (uint32_t v = 0; v < g.n_nodes; ++v) g.in_sub[v] = 1;
      } else {
        MarkSubgraph(g, le, lb);
      }
      uint32_t rows = g.overflow ? 0 : TopoSort(g);
      // a node is a sink if it has no out-edge inside the (sub)graph
      for (uint32_t r = 0; r < rows; ++r) {
        const uint16_t v = g.r2n[r];
        uint8_t sink = 1;
        for (uint16_t e = g.out_head[v]; e != kNone; e = g.e_next_out[e]) {
          if (g.in_sub[g.e_head[e]]) {
            sink = 0;
            break;
          }
        }
        g.sink[v] = sink;
      }
      // int16 cells: |score| <= max|m,n,g| * (rows + columns)
      const int amax = max(max(abs(m), abs(n)), abs(gap));
      if (static_cast<long long>(amax) * (rows + len + 2) > 32000) g.overflow = true;
      sh_rows = rows;
    }
    __syncwarp();
    const uint32_t rows = sh_rows;
    if (g.overflow) break;
    my_cells += static_cast<unsigned long long>(rows + 1) * (len + 1);

    // ---- DP (all lanes) ----
    int16_t* H = g.H;
    for (uint32_t j = lane; j <= len; j += 32) H[j] = static_cast<int16_t>(j * gap);
    __syncwarp();
    int best_score = -2147483647, best_row = -1;
    for (uint32_t r = 0; r < rows; ++r) {
      const uint16_t v = g.r2n[r];
      const uint32_t vcode = g.code[v];
      int16_t* row = H + static_cast<size_t>(r + 1) * W;
      // column 0
      int col0 = 0;
      {
        bool any = false;
        int pen = -2147483647;
        for (uint16_t e = g.in_head[v]; e != kNone; e = g.e_next_in[e]) {
          const uint16_t t = g.e_tail[e];
          if (!g.in_sub[t]) continue;
          any = true;
          pen = max(pen, static_cast<int>(H[static_cast<size_t>(g.n2r[t] + 1) * W]));
        }
        col0 = (any ? pen : 0) + gap;
      }
      int carry = col0;  // running max of M[k] - k*g over k < chunk start (k = 0 term)
      for (uint32_t c0 = 1; c0 <= len; c0 += 32) {
        const uint32_t j = c0 + lane;
        int M = -1000000;
        if (j <= len) {
          const int match = (static_cast<uint32_t>(CodeOf(sv.bases[j - 1])) == vcode) ? m : n;
          bool any = false;
          for (uint16_t e = g.in_head[v]; e != kNone; e = g.e_next_in[e]) {
            const uint16_t t = g.e_tail[e];
            if (!g.in_sub[t]) continue;
            any = true;
            const int16_t* pr = H + static_cast<size_t>(g.n2r[t] + 1) * W;
            M = max(M, max(static_cast<int>(pr[j - 1]) + match,
                           static_cast<int>(pr[j]) + gap));
          }
          if (!any) {
            M = max(static_cast<int>(H[j - 1]) + match, static_cast<int>(H[j]) + gap);
          }
        }
        // horizontal recurrence as a max-scan of T = M - j*g
        int T = j <= len ? M - static_cast<int>(j) * gap : -1000000000;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, T, d);
          if (static_cast<int>(lane) >= d) T = max(T, o);
        }
        T = max(T, carry);
        if (j <= len) row[j] = static_cast<int16_t>(T + static_cast<int>(j) * gap);
        carry = __shfl_sync(0xffffffffu, T, 31);
      }
      if (lane == 0) row[0] = static_cast<int16_t>(col0);
      __syncwarp();
      if (g.sink[v]) {
        const int sc = row[len];
        if (best_score < sc) {
          best_score = sc;
          best_row = static_cast<int>(r + 1);
        }
      }
    }

    // ---- (lane 0) traceback + merge ----
    if (lane == 0) {
      uint32_t alen = 0;
      uint32_t i = best_row < 0 ? 0 : static_cast<uint32_t>(best_row), j = len;
      if (best_row < 0) j = 0;
      const uint32_t acap = shape.ncap + shape.lmax + 2;
      while (!(i == 0 && j == 0)) {
        const int h = H[static_cast<size_t>(i) * W + j];
        uint32_t pi = 0, pj = 0;
        bool found = false;
        if (i != 0 && j != 0) {
          const uint16_t v = g.r2n[i - 1];
          const int match =
              (static_cast<uint32_t>(CodeOf(sv.bases[j - 1])) == g.code[v]) ? m : n;
          bool any = false;
          for (uint16_t e = g.in_head[v]; e != kNone && !found; e = g.e_next_in[e]) {
            const uint16_t t = g.e_tail[e];
            if (!g.in_sub[t]) continue;
            any = true;
            const uint32_t p = g.n2r[t] + 1;
            if (h == H[static_cast<size_t>(p) * W + (j - 1)] + match) {
              pi = p;
              pj = j - 1;
              found = true;
            }
          }
          if (!any && h == H[j - 1] + match) {
            pi = 0;
            pj = j - 1;
            found = true;
          }
        }
        if (!found && i != 0) {
          const uint16_t v = g.r2n[i - 1];
          bool any = false;
          for (uint16_t e = g.in_head[v]; e != kNone && !found; e = g.e_next_in[e]) {
            const uint16_t t = g.e_tail[e];
            if (!g.in_sub[t]) continue;
            any = true;
            const uint32_t p = g.n2r[t] + 1;
            if (h == H[static_cast<size_t>(p) * W + j] + gap) {
              pi = p;
              pj = j;
              found = true;
            }
          }
          if (!any && h == H[j] + gap) {
            pi = 0;
            pj = j;
            found = true;
          }
        }
        if (!found && j != 0 && h == H[static_cast<size_t>(i) * W + j - 1] + gap) {
          pi = i;
          pj = j - 1;
          found = true;
        }
        if (!found || alen >= acap) {  // cannot happen for a consistent matrix
          g.overflow = true;
          break;
        }
        g.aln_node[alen] = i == pi ? -1 : static_cast<int16_t>(g.r2n[i - 1]);
        g.aln_pos[alen] = j == pj ? -1 : static_cast<int16_t>(j - 1);
        ++alen;
        i = pi;
        j = pj;
      }
      // reverse into emission order
      for (uint32_t a = 0; a < alen / 2; ++a) {
        const int16_t tn = g.aln_node[a], tp = g.aln_pos[a];
        g.aln_node[a] = g.aln_node[alen - 1 - a];
        g.aln_pos[a] = g.aln_pos[alen - 1 - a];
        g.aln_node[alen - 1 - a] = tn;
        g.aln_pos[alen - 1 - a] = tp;
      }
      if (!g.overflow) {
        if (alen == 0) {
          // spoa: an empty alignment appends the sequence as a separate chain
          const uint32_t before = g.n_nodes;
          AddChain(g, sv, 0, sv.len);
          if (sv.len >= 2) {
            for (uint32_t v = before; v < g.n_nodes; ++v) g.cov[v] += 1;
          }
        } else {
          AddAlignment(g, sv, alen);
        }
      }
    }
    __syncwarp();
  }

  // ---- (lane 0) consensus: heaviest bundle + branch completion + trimming ----
  if (lane == 0) {
    uint32_t st = 0;
    uint32_t clen = 0;
    if (sh_fail) {
      st = sh_fail;
    } else if (g.overflow) {
      st = kPoaStatusCapacity;
    } else {
      for (uint32_t v = 0; v < g.n_nodes; ++v) g.in_sub[v] = 1;
      const uint32_t rows = TopoSort(g);
      if (g.overflow) {
        st = kPoaStatusCapacity;
      } else {
        for (uint32_t v = 0; v < g.n_nodes; ++v) {
          g.pred[v] = -1;
          g.score[v] = -1;
        }
        auto relax = [&](uint16_t v) {
          for (uint16_t e = g.in_head[v]; e != kNone; e = g.e_next_in[e]) {
            const uint16_t t = g.e_tail[e];
            const long long wgt = g.e_weight[e];
            if (g.score[v] < wgt ||
                (g.score[v] == wgt && g.score[g.pred[v]] <= g.score[t])) {
              g.score[v] = wgt;
              g.pred[v] = t;
            }
          }
          if (g.pred[v] != -1) g.score[v] += g.score[g.pred[v]];
        };
        int32_t mx = -1;
        for (uint32_t r = 0; r < rows; ++r) {
          const uint16_t v = g.r2n[r];
          relax(v);
          if (mx < 0 || g.score[mx] < g.score[v]) mx = v;
        }
        while (g.out_head[mx] != kNone) {  // branch completion
          const uint32_t rank = g.n2r[mx];
          for (uint16_t e = g.out_head[mx]; e != kNone; e = g.e_next_out[e]) {
            const uint16_t hd = g.e_head[e];
            for (uint16_t f = g.in_head[hd]; f != kNone; f = g.e_next_in[f]) {
              if (g.e_tail[f] != mx) g.score[g.e_tail[f]] = -1;
            }
          }
          int32_t nmx = -1;
          for (uint32_t r = rank + 1; r < rows; ++r) {
            const uint16_t v = g.r2n[r];
            g.score[v] = -1;
            g.pred[v] = -1;
            for (uint16_t e = g.in_head[v]; e != kNone; e = g.e_next_in[e]) {
              const uint16_t t = g.e_tail[e];
              if (g.score[t] == -1) continue;
              const long long wgt = g.e_weight[e];
              if (g.score[v] < wgt ||
                  (g.score[v] == wgt && g.score[g.pred[v]] <= g.score[t])) {
                g.score[v] = wgt;
                g.pred[v] = t;
              }
            }
            if (g.pred[v] != -1) g.score[v] += g.score[g.pred[v]];
            if (nmx < 0 || g.score[nmx] < g.score[v]) nmx = v;
          }
          mx = nmx;
        }
        // walk back; r2n is free now and holds the reversed path
        uint32_t plen = 0;
        int32_t v = mx;
        while (g.pred[v] != -1) {
          g.r2n[plen++] = static_cast<uint16_t>(v);
          v = g.pred[v];
        }
        g.r2n[plen++] = static_cast<uint16_t>(v);
        // coverage of a consensus base = its node + the aligned nodes
        auto coverage = [&](uint32_t idx) -> uint32_t {
          const uint16_t nd = g.r2n[plen - 1 - idx];
          uint32_t cvg = g.cov[nd];
          for (uint32_t a = 0; a < g.n_aligned[nd]; ++a) cvg += g.cov[g.aligned[3 * nd + a]];
          return cvg;
        };
        uint32_t cb = 0, ce = plen;  // [cb, ce)
        st = kPoaStatusOk;
        if (tgs && trim) {
          const uint32_t avg = (nseq - 1) / 2;
          int32_t b = 0, e2 = static_cast<int32_t>(plen) - 1;
          for (; b < static_cast<int32_t>(plen); ++b) {
            if (coverage(b) >= avg) break;
          }
          for (; e2 >= 0; --e2) {
            if (coverage(e2) >= avg) break;
          }
          if (b >= e2) {
            st |= kPoaStatusChimeric;
          } else {
            cb = b;
            ce = e2 + 1;
          }
        }
        clen = ce - cb;
        for (uint32_t i = 0; i < clen; ++i) {
          const uint16_t nd = g.r2n[plen - 1 - (cb + i)];
          out[i] = "ACGT"[g.code[nd]];
          if (cov_dst) cov_dst[i] = coverage(cb + i);
        }
      }
    }
    cons_len[w] = clen;
    status[w] = static_cast<uint8_t>(st);
  }
  // DP cells of this window (lanes hold the same count)
  if (lane == 0 && cells) atomicAdd(cells, my_cells);
}

}  // namespace
}  // namespace rvn

#include "poa_fast.cuh"

namespace rvn {

// Host driver: windows in batches sized by the scratch budget; windows that
// outgrow the first-tier capacity are retried with the exact upper bound.
void PoaBatch(Ctx& c, uint32_t n_windows, const uint32_t* h_win_first,
              const uint64_t* h_seq_off, const uint8_t* h_bases, const uint8_t* h_quals,
              const uint32_t* h_seq_begin, const uint32_t* h_seq_end, int m, int n,
              int gap, bool trim, bool tgs, bool want_coverage) {
  c.poa_valid = false;
  if (gap >= 0) {
    throw InvalidArgument("[racon::Polisher::Create] error: gap penalty must be negative");
  }
  const uint32_t n_seqs = n_windows ? h_win_first[n_windows] : 0;
  const uint64_t n_bases = n_seqs ? h_seq_off[n_seqs] : 0;
  TimerBegin(c, "poa_h2d");
  uint32_t* d_wf = c.po_win_first.reserve(n_windows + 2ULL);
  uint64_t* d_so = c.po_seq_off.reserve(n_seqs + 2ULL);
  uint8_t* d_bases = c.po_bases.reserve(n_bases + 16);
  uint8_t* d_quals = h_quals ? c.po_quals.reserve(n_bases + 16) : nullptr;
  uint32_t* d_sb = c.po_seq_begin.reserve(n_seqs + 2ULL);
  uint32_t* d_se = c.po_seq_end.reserve(n_seqs + 2ULL);
  RVN_CUDA(cudaMemcpyAsync(d_wf, h_win_first, (n_windows + 1ULL) * 4, cudaMemcpyHostToDevice, c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_so, h_seq_off, (n_seqs + 1ULL) * 8, cudaMemcpyHostToDevice, c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_bases, h_bases, n_bases, cudaMemcpyHostToDevice, c.stream));
  if (h_quals) RVN_CUDA(cudaMemcpyAsync(d_quals, h_quals, n_bases, cudaMemcpyHostToDevice, c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_sb, h_seq_begin, n_seqs * 4ULL, cudaMemcpyHostToDevice, c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_se, h_seq_end, n_seqs * 4ULL, cudaMemcpyHostToDevice, c.stream));
  TimerEnd(c);

  // output slots: a consensus never has more bases than the window holds
  c.po_cons_off.assign(n_windows + 1ULL, 0);
  // per window: longest layer, backbone length, all bases - the kernel is chosen
  // PER WINDOW (one long layer must not send a whole batch to the generic kernel)
  std::vector<uint32_t> w_lmax(n_windows, 1), w_bb(n_windows, 1);
  std::vector<uint64_t> w_tot(n_windows, 1);
  for (uint32_t w = 0; w < n_windows; ++w) {
    const uint32_t s0 = h_win_first[w], s1 = h_win_first[w + 1];
    uint64_t tot = 0;
    for (uint32_t s = s0; s < s1; ++s) {
      const uint64_t len = h_seq_off[s + 1] - h_seq_off[s];
      tot += len;
      if (s > s0) w_lmax[w] = std::max<uint32_t>(w_lmax[w], static_cast<uint32_t>(len));
      else w_bb[w] = std::max<uint32_t>(w_bb[w], static_cast<uint32_t>(len));
    }
    w_tot[w] = std::max<uint64_t>(tot, 1);
    if (tot > 65000 || w_lmax[w] > 32000) {
      throw LimitError("a POA window holds more than 65000 bases");
    }
    c.po_cons_off[w + 1] = c.po_cons_off[w] + tot;
  }
  const uint64_t out_total = c.po_cons_off[n_windows];
  uint64_t* d_coff = c.po_d_cons_off.reserve(n_windows + 2ULL);
  RVN_CUDA(cudaMemcpyAsync(d_coff, c.po_cons_off.data(), (n_windows + 1ULL) * 8, cudaMemcpyHostToDevice, c.stream));
  uint8_t* d_cons = c.po_cons.reserve(out_total + 16);
  uint32_t* d_cov = want_coverage ? c.po_cov.reserve(out_total + 16) : nullptr;
  uint32_t* d_clen = c.po_cons_len.reserve(n_windows + 2ULL);
  uint8_t* d_status = c.po_status.reserve(n_windows + 2ULL);
  uint64_t* d_cells = c.m_counter.reserve((1u << 16) + 8);
  RVN_CUDA(cudaMemsetAsync(d_cells, 0, 8, c.stream));
  RVN_CUDA(cudaMemsetAsync(d_status, 0, n_windows + 1ULL, c.stream));

  c.po_h_status.assign(n_windows, 0);
  TimerBegin(c, "poa");
  // group 0: every layer fits the register-row kernel; group 1: the generic kernel
  for (int group = 0; group < 2; ++group) {
    std::vector<uint32_t> todo;
    for (uint32_t w = 0; w < n_windows; ++w) {
      if ((w_lmax[w] + 1 <= kFastCols) == (group == 0)) todo.push_back(w);
    }
    for (int tier = 0; tier < 2 && !todo.empty(); ++tier) {
      uint32_t lmax = 1, bb_max = 1;
      uint64_t total_max = 1;
      for (uint32_t w : todo) {
        lmax = std::max(lmax, w_lmax[w]);
        bb_max = std::max(bb_max, w_bb[w]);
        total_max = std::max(total_max, w_tot[w]);
      }
      PoaShape shape;
      shape.lmax = lmax;
      shape.ncap = tier == 0 ? std::min<uint64_t>(total_max, 3ULL * bb_max + lmax + 64)
                             : static_cast<uint32_t>(total_max);
      shape.ncap = std::min<uint32_t>(shape.ncap, 65000);
      shape.ecap = std::min<uint32_t>(3 * shape.ncap, 65000);
      shape.rows = shape.ncap + 1;
      shape.width = (lmax + 1 + 31) & ~31u;
      const size_t fast_smem = ((shape.ncap + 15) & ~15u) + 4ULL * shape.ncap + 16;
      const bool fast = group == 0 && fast_smem <= 220 * 1024;  // the register-row kernel
      const size_t stride = ((fast ? MakeFastLayout(shape).bytes : MakePoaLayout(shape).bytes) +
                             255) & ~size_t(255);
      if (fast) {
        RVN_CUDA(cudaFuncSetAttribute(PoaKernelFast,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(fast_smem)));
      }
      const size_t budget = 48ULL << 30;  // scratch budget per batch
      const uint32_t batch = static_cast<uint32_t>(
          std::max<size_t>(1, std::min<size_t>(todo.size(), budget / stride)));
      uint8_t* scratch = c.po_scratch.reserve(stride * batch);
      uint32_t* d_list = c.po_list.reserve(n_windows + 1ULL);
      RVN_CUDA(cudaMemcpyAsync(d_list, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, c.stream));
      for (size_t b0 = 0; b0 < todo.size(); b0 += batch) {
        const uint32_t nb = static_cast<uint32_t>(std::min<size_t>(batch, todo.size() - b0));
        if (fast) {
          PoaKernelFast<<<nb, 32, fast_smem, c.stream>>>(
              nb, d_list + b0, d_wf, d_so, d_bases, d_quals, d_sb, d_se, m, n, gap, trim,
              tgs, shape, scratch, stride, d_cons, d_coff, d_clen, d_cov, d_status,
              reinterpret_cast<unsigned long long*>(d_cells));
        } else {
          PoaKernel<<<nb, 32, 0, c.stream>>>(
              nb, d_list + b0, d_wf, d_so, d_bases, d_quals, d_sb, d_se, m, n, gap, trim,
              tgs, shape, scratch, stride, d_cons, d_coff, d_clen, d_cov, d_status,
              reinterpret_cast<unsigned long long*>(d_cells));
        }
        RVN_LAUNCH_CHECK();
        ++c.launches;
      }
      RVN_CUDA(cudaMemcpyAsync(c.po_h_status.data(), d_status, n_windows, cudaMemcpyDeviceToHost, c.stream));
      RVN_CUDA(cudaStreamSynchronize(c.stream));  // (also: todo / d_list are reusable)
      std::vector<uint32_t> again;
      for (uint32_t w : todo) {
        if (c.po_h_status[w] == kPoaStatusCapacity) again.push_back(w);
      }
      todo.swap(again);
    }
  }
  TimerEnd(c);
  for (uint32_t w = 0; w < n_windows; ++w) {
    if (c.po_h_status[w] == kPoaStatusCapacity) {
      throw LimitError("a POA graph outgrew 65000 nodes/edges or the int16 score range");
    }
    if (c.po_h_status[w] == kPoaStatusInvalid) {
      throw InvalidArgument(
          "[racon::Window::AddLayer] error: layer begin and end positions are invalid");
    }
  }

  // results to the host, compacted
  TimerBegin(c, "poa_d2h");
  c.po_h_clen.assign(n_windows, 0);
  c.po_h_cons.resize(out_total);
  RVN_CUDA(cudaMemcpyAsync(c.po_h_clen.data(), d_clen, n_windows * 4ULL, cudaMemcpyDeviceToHost, c.stream));
  RVN_CUDA(cudaMemcpyAsync(c.po_h_cons.data(), d_cons, out_total, cudaMemcpyDeviceToHost, c.stream));
  if (want_coverage) {
    c.po_h_cov.resize(out_total);
    RVN_CUDA(cudaMemcpyAsync(c.po_h_cov.data(), d_cov, out_total * 4, cudaMemcpyDeviceToHost, c.stream));
  }
  c.po_cells = ReadU64(c, d_cells);
  TimerEnd(c);
  // compact the per-window slots
  c.po_out_off.assign(n_windows + 1ULL, 0);
  for (uint32_t w = 0; w < n_windows; ++w) c.po_out_off[w + 1] = c.po_out_off[w] + c.po_h_clen[w];
  c.po_out_cons.resize(c.po_out_off[n_windows]);
  if (want_coverage) c.po_out_cov.resize(c.po_out_off[n_windows]);
  for (uint32_t w = 0; w < n_windows; ++w) {
    std::copy(c.po_h_cons.begin() + c.po_cons_off[w],
              c.po_h_cons.begin() + c.po_cons_off[w] + c.po_h_clen[w],
              c.po_out_cons.begin() + c.po_out_off[w]);
    if (want_coverage) {
      std::copy(c.po_h_cov.begin() + c.po_cons_off[w],
                c.po_h_cov.begin() + c.po_cons_off[w] + c.po_h_clen[w],
                c.po_out_cov.begin() + c.po_out_off[w]);
    }
  }
  c.po_n_windows = n_windows;
  c.po_has_cov = want_coverage;
  c.poa_valid = true;
}

}  // namespace rvn
