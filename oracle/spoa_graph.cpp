// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/spoa/spoa.hpp).
// Partial order graph + global sequence-to-graph alignment + heaviest-bundle
// consensus, restated from spoa's published behaviour.
#include "spoa/spoa.hpp"

#include <algorithm>
#include <limits>
#include <stack>
#include <unordered_set>

namespace spoa {

std::uint32_t Graph::Node::Coverage() const {
  std::unordered_set<std::uint32_t> labels;
  for (const auto& e : inedges) {
    for (auto l : e->labels) labels.emplace(l);
  }
  for (const auto& e : outedges) {
    for (auto l : e->labels) labels.emplace(l);
  }
  return labels.size();
}

Graph::Graph()
    : num_codes_(0), coder_(256, -1), decoder_(256, -1), sequences_(), nodes_(),
      edges_(), rank_to_node_(), consensus_() {}

Graph::Node* Graph::AddNode(std::uint32_t code) {
  nodes_.emplace_back(new Node(nodes_.size(), code));
  return nodes_.back().get();
}

void Graph::AddEdge(Node* tail, Node* head, std::uint32_t weight) {
  for (const auto& it : tail->outedges) {
    if (it->head == head) {
      it->AddSequence(sequences_.size(), weight);
      return;
    }
  }
  edges_.emplace_back(new Edge(tail, head, sequences_.size(), weight));
  tail->outedges.emplace_back(edges_.back().get());
  head->inedges.emplace_back(edges_.back().get());
}

// a chain of fresh nodes for sequence[begin, end); returns its first node
Graph::Node* Graph::AddSequence(const char* sequence,
                                const std::vector<std::uint32_t>& weights,
                                std::uint32_t begin, std::uint32_t end) {
  if (begin == end) {
    return nullptr;
  }
  Node* prev = nullptr;
  for (std::uint32_t i = begin; i < end; ++i) {
    Node* curr = AddNode(coder_[static_cast<std::uint8_t>(sequence[i])]);
    if (prev) {
      AddEdge(prev, curr, weights[i - 1] + weights[i]);  // both ends vote
    }
    prev = curr;
  }
  return nodes_[nodes_.size() - (end - begin)].get();
}

void Graph::AddAlignment(const Alignment& alignment, const char* sequence,
                         std::uint32_t sequence_len, std::uint32_t weight) {
  AddAlignment(alignment, sequence, sequence_len,
               std::vector<std::uint32_t>(sequence_len, weight));
}

void Graph::AddAlignment(const Alignment& alignment, const char* sequence,
                         std::uint32_t sequence_len, const char* quality) {
  std::vector<std::uint32_t> weights;
  for (std::uint32_t i = 0; i < sequence_len; ++i) {
    weights.emplace_back(static_cast<std::uint32_t>(quality[i]) - 33);  // Phred
  }
  AddAlignment(alignment, sequence, sequence_len, weights);
}

void Graph::AddAlignment(const Alignment& alignment, const char* sequence,
                         std::uint32_t sequence_len,
                         const std::vector<std::uint32_t>& weights) {
  if (sequence_len == 0) {
    return;
  }
  // letters get their code on first appearance
  for (std::uint32_t i = 0; i < sequence_len; ++i) {
    const auto c = static_cast<std::uint8_t>(sequence[i]);
    if (coder_[c] == -1) {
      coder_[c] = num_codes_;
      decoder_[num_codes_++] = c;
    }
  }

  if (alignment.empty()) {
    sequences_.emplace_back(AddSequence(sequence, weights, 0, sequence_len));
    TopologicalSort();
    return;
  }

  std::vector<std::uint32_t> valid;
  for (const auto& it : alignment) {
    if (it.second != -1) {
      valid.emplace_back(it.second);
    }
  }

  // unaligned head and tail become fresh chains
  const std::uint32_t before = nodes_.size();
  Node* begin = AddSequence(sequence, weights, 0, valid.front());
  Node* prev = before == nodes_.size() ? nullptr : nodes_.back().get();
  Node* last = AddSequence(sequence, weights, valid.back() + 1, sequence_len);

  for (const auto& it : alignment) {
    if (it.second == -1) {
      continue;
    }
    const std::uint32_t code = coder_[static_cast<std::uint8_t>(sequence[it.second])];
    Node* curr = nullptr;
    if (it.first == -1) {
      curr = AddNode(code);
    } else {
      Node* jt = nodes_[it.first].get();
      if (jt->code == code) {
        curr = jt;
      } else {
        for (const auto& kt : jt->aligned_nodes) {
          if (kt->code == code) {
            curr = kt;
            break;
          }
        }
        if (!curr) {  // a new letter for this column: link it to the whole column
          curr = AddNode(code);
          for (const auto& kt : jt->aligned_nodes) {
            kt->aligned_nodes.emplace_back(curr);
            curr->aligned_nodes.emplace_back(kt);
          }
          jt->aligned_nodes.emplace_back(curr);
          curr->aligned_nodes.emplace_back(jt);
        }
      }
    }
    if (!begin) {
      begin = curr;
    }
    if (prev) {
      AddEdge(prev, curr, weights[it.second - 1] + weights[it.second]);
    }
    prev = curr;
  }
  if (last) {
    AddEdge(prev, last, weights[valid.back()] + weights[valid.back() + 1]);
  }
  sequences_.emplace_back(begin);
  TopologicalSort();
}

// iterative DFS over in-edges; aligned nodes are pushed together and emitted
// contiguously
void Graph::TopologicalSort() {
  rank_to_node_.clear();
  std::vector<std::uint8_t> marks(nodes_.size(), 0);
  std::vector<bool> ignored(nodes_.size(), false);
  std::stack<Node*> stack;
  for (const auto& it : nodes_) {
    if (marks[it->id] != 0) {
      continue;
    }
    stack.push(it.get());
    while (!stack.empty()) {
      Node* curr = stack.top();
      bool is_valid = true;
      if (marks[curr->id] != 2) {
        for (const auto& jt : curr->inedges) {
          if (marks[jt->tail->id] != 2) {
            stack.push(jt->tail);
            is_valid = false;
          }
        }
        if (!ignored[curr->id]) {
          for (const auto& jt : curr->aligned_nodes) {
            if (marks[jt->id] != 2) {
              stack.push(jt);
              ignored[jt->id] = true;
              is_valid = false;
            }
          }
        }
        if (is_valid) {
          marks[curr->id] = 2;
          if (!ignored[curr->id]) {
            rank_to_node_.emplace_back(curr);
            for (const auto& jt : curr->aligned_nodes) {
              rank_to_node_.emplace_back(jt);
            }
          }
        } else {
          marks[curr->id] = 1;
        }
      }
      if (is_valid) {
        stack.pop();
      }
    }
  }
}

void Graph::ExtractSubgraph(const Node* begin, const Node* end,
                            std::vector<bool>* dst) const {
  std::stack<const Node*> stack;
  stack.push(begin);
  while (!stack.empty()) {
    const Node* curr = stack.top();
    stack.pop();
    if (!(*dst)[curr->id] && curr->id >= end->id) {
      for (const auto& it : curr->inedges) {
        stack.push(it->tail);
      }
      for (const auto& it : curr->aligned_nodes) {
        stack.push(it);
      }
      (*dst)[curr->id] = true;
    }
  }
}

Graph Graph::Subgraph(std::uint32_t begin, std::uint32_t end,
                      std::vector<const Node*>* subgraph_to_graph) const {
  std::vector<bool> is_in(nodes_.size(), false);
  ExtractSubgraph(nodes_[end].get(), nodes_[begin].get(), &is_in);

  Graph sub{};
  sub.num_codes_ = num_codes_;
  sub.coder_ = coder_;
  sub.decoder_ = decoder_;

  subgraph_to_graph->assign(nodes_.size(), nullptr);
  std::vector<Node*> graph_to_subgraph(nodes_.size(), nullptr);
  for (const auto& it : nodes_) {
    if (!is_in[it->id]) {
      continue;
    }
    sub.AddNode(it->code);
    graph_to_subgraph[it->id] = sub.nodes_.back().get();
    (*subgraph_to_graph)[sub.nodes_.back()->id] = it.get();
  }
  for (const auto& it : nodes_) {
    if (!is_in[it->id]) {
      continue;
    }
    Node* jt = graph_to_subgraph[it->id];
    for (const auto& kt : it->inedges) {
      if (graph_to_subgraph[kt->tail->id]) {
        sub.AddEdge(graph_to_subgraph[kt->tail->id], jt, kt->weight);
      }
    }
    for (const auto& kt : it->aligned_nodes) {
      if (graph_to_subgraph[kt->id]) {
        jt->aligned_nodes.emplace_back(graph_to_subgraph[kt->id]);
      }
    }
  }
  sub.TopologicalSort();
  return sub;
}

void Graph::UpdateAlignment(const std::vector<const Node*>& subgraph_to_graph,
                            Alignment* alignment) const {
  for (auto& it : *alignment) {
    if (it.first != -1) {
      it.first = subgraph_to_graph[it.first]->id;
    }
  }
}

Graph::Node* Graph::BranchCompletion(std::uint32_t rank,
                                     std::vector<std::int64_t>* scores,
                                     std::vector<std::int32_t>* predecessors) {
  Node* start = rank_to_node_[rank];
  for (const auto& it : start->outedges) {
    for (const auto& jt : it->head->inedges) {
      if (jt->tail != start) {
        (*scores)[jt->tail->id] = -1;
      }
    }
  }
  Node* max = nullptr;
  for (std::uint32_t i = rank + 1; i < rank_to_node_.size(); ++i) {
    Node* it = rank_to_node_[i];
    (*scores)[it->id] = -1;
    (*predecessors)[it->id] = -1;
    for (const auto& jt : it->inedges) {
      if ((*scores)[jt->tail->id] == -1) {
        continue;
      }
      if ((*scores)[it->id] < jt->weight ||
          ((*scores)[it->id] == jt->weight &&
           (*scores)[(*predecessors)[it->id]] <= (*scores)[jt->tail->id])) {
        (*scores)[it->id] = jt->weight;
        (*predecessors)[it->id] = jt->tail->id;
      }
    }
    if ((*predecessors)[it->id] != -1) {
      (*scores)[it->id] += (*scores)[(*predecessors)[it->id]];
    }
    if (max == nullptr || (*scores)[max->id] < (*scores)[it->id]) {
      max = it;
    }
  }
  return max;
}

void Graph::TraverseHeaviestBundle() {
  consensus_.clear();
  if (rank_to_node_.empty()) {
    return;
  }
  std::vector<std::int32_t> predecessors(nodes_.size(), -1);
  std::vector<std::int64_t> scores(nodes_.size(), -1);
  Node* max = nullptr;
  for (const auto& it : rank_to_node_) {
    for (const auto& jt : it->inedges) {
      if (scores[it->id] < jt->weight ||
          (scores[it->id] == jt->weight &&
           scores[predecessors[it->id]] <= scores[jt->tail->id])) {
        scores[it->id] = jt->weight;
        predecessors[it->id] = jt->tail->id;
      }
    }
    if (predecessors[it->id] != -1) {
      scores[it->id] += scores[predecessors[it->id]];
    }
    if (max == nullptr || scores[max->id] < scores[it->id]) {
      max = it;
    }
  }
  if (!max->outedges.empty()) {
    std::vector<std::uint32_t> node_id_to_rank(nodes_.size(), 0);
    for (std::uint32_t i = 0; i < rank_to_node_.size(); ++i) {
      node_id_to_rank[rank_to_node_[i]->id] = i;
    }
    while (!max->outedges.empty()) {
      max = BranchCompletion(node_id_to_rank[max->id], &scores, &predecessors);
    }
  }
  while (predecessors[max->id] != -1) {
    consensus_.emplace_back(max);
    max = nodes_[predecessors[max->id]].get();
  }
  consensus_.emplace_back(max);
  std::reverse(consensus_.begin(), consensus_.end());
}

std::string Graph::GenerateConsensus(std::vector<std::uint32_t>* coverages) {
  TraverseHeaviestBundle();
  std::string dst;
  for (const auto& it : consensus_) {
    dst += static_cast<char>(decoder_[it->code]);
  }
  if (coverages) {
    coverages->clear();
    for (const auto& it : consensus_) {
      coverages->emplace_back(it->Coverage());
      for (const auto& jt : it->aligned_nodes) {
        coverages->back() += jt->Coverage();
      }
    }
  }
  return dst;
}

// ---------------------------------------------------------------------------
// global alignment, linear gaps
// ---------------------------------------------------------------------------
namespace {
bool g_use_simd = false;
}
void AlignmentEngine::UseSimd(bool on) { g_use_simd = on; }

#if defined(__AVX2__)
}  // namespace spoa
#include <immintrin.h>
namespace spoa {

namespace {

// v shifted up by one 16-bit lane, lane 0 = `fill` (across the two 128-bit halves)
inline __m256i ShiftUp1(__m256i v, __m256i fill) {
  const __m256i t = _mm256_permute2x128_si256(fill, v, 0x21);  // [fill.hi | v.lo]
  return _mm256_alignr_epi8(v, t, 14);
}
inline __m256i ShiftUpN(__m256i v, __m256i fill, int n) {  // n in {1, 2, 4, 8}
  const __m256i t = _mm256_permute2x128_si256(fill, v, 0x21);
  switch (n) {
    case 1: return _mm256_alignr_epi8(v, t, 14);
    case 2: return _mm256_alignr_epi8(v, t, 12);
    case 4: return _mm256_alignr_epi8(v, t, 8);
    default: return t;  // 8 lanes = one 128-bit half
  }
}

}  // namespace

// Same cell values as the scalar loops of Align, 16 int16 cells at a time:
//   x[j]   = max over predecessors p of max(H[p][j-1] + profile[j], H[p][j] + g)
//   H[i][j] = max(x[j], H[i][j-1] + g) = j*g + prefix_max_k<=j (x[k] - k*g)
// H16_ rows are `stride` wide (a multiple of 16, 16 spare lanes on the left so
// that column -1 reads are harmless).
bool AlignmentEngine::FillSimd16(const char* sequence, std::uint32_t sequence_len,
                                 const Graph& graph) {
  const auto& rank_to_node = graph.rank_to_node();
  const std::uint64_t width = sequence_len + 1ULL;
  const std::uint64_t height = graph.nodes().size() + 1ULL;
  const int amax = std::max(std::max(std::abs(m_), std::abs(n_)), std::abs(g_));
  if (static_cast<std::uint64_t>(amax) * (2 * width + height + 32) > 32000) return false;
  const std::uint64_t stride = ((width + 15) / 16) * 16 + 16;  // 16 guard lanes in front
  constexpr std::int16_t kNeg = -32000;
  H16_.assign(stride * height, kNeg);
  profile16_.assign(graph.num_codes() * stride, 0);
  for (std::uint32_t c = 0; c < graph.num_codes(); ++c) {
    for (std::uint64_t j = 0; j < sequence_len; ++j) {
      profile16_[c * stride + 16 + (j + 1)] =
          (static_cast<std::int32_t>(c) == graph.coder(sequence[j])) ? m_ : n_;
    }
  }
  auto row_of = [&](std::uint64_t i) { return H16_.data() + i * stride + 16; };
  for (std::uint64_t j = 0; j < width; ++j) row_of(0)[j] = static_cast<std::int16_t>(j * g_);
  for (std::uint64_t i = 1; i < height; ++i) {
    const auto& edges = rank_to_node[i - 1]->inedges;
    std::int32_t penalty = edges.empty() ? 0 : -1000000;
    for (const auto& it : edges) {
      penalty = std::max<std::int32_t>(penalty, row_of(node_id_to_rank_[it->tail->id] + 1)[0]);
    }
    row_of(i)[0] = static_cast<std::int16_t>(penalty + g_);
  }
  // j * g per lane, and the per-vector step 16 * g
  alignas(32) std::int16_t ramp[16];
  for (int l = 0; l < 16; ++l) ramp[l] = static_cast<std::int16_t>(l * g_);
  const __m256i vramp = _mm256_load_si256(reinterpret_cast<const __m256i*>(ramp));
  const __m256i vg = _mm256_set1_epi16(g_);
  const __m256i vneg = _mm256_set1_epi16(kNeg);
  const std::uint64_t nvec = (width + 15) / 16;
  for (const auto& it : rank_to_node) {
    const std::uint64_t i = node_id_to_rank_[it->id] + 1;
    std::int16_t* row = row_of(i);
    const std::int16_t* prof = profile16_.data() + it->code * stride + 16;
    const std::int16_t col0 = row[0];
    const std::size_t np = std::max<std::size_t>(1, it->inedges.size());
    std::int16_t carry = kNeg;  // prefix max of x[k] - k*g over the vectors before
    for (std::uint64_t v = 0; v < nvec; ++v) {
      const std::uint64_t j0 = v * 16;
      __m256i x = vneg;
      for (std::size_t p = 0; p < np; ++p) {
        const std::uint64_t pred_i =
            it->inedges.empty() ? 0 : node_id_to_rank_[it->inedges[p]->tail->id] + 1;
        const std::int16_t* pred = row_of(pred_i);
        const __m256i up = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(pred + j0));
        const __m256i dg = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(pred + j0 - 1));
        const __m256i pf = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(prof + j0));
        x = _mm256_max_epi16(x, _mm256_max_epi16(_mm256_add_epi16(dg, pf),
                                                 _mm256_add_epi16(up, vg)));
      }
      // y[l] = x[l] - (j0 + l) * g ; column 0 is fixed by the column initialisation
      const __m256i base = _mm256_set1_epi16(static_cast<std::int16_t>(j0 * g_));
      const __m256i off = _mm256_add_epi16(base, vramp);
      __m256i y = _mm256_sub_epi16(x, off);
      if (v == 0) {
        alignas(32) std::int16_t tmp[16];
        _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), y);
        tmp[0] = col0;  // (0 * g = 0)
        y = _mm256_load_si256(reinterpret_cast<const __m256i*>(tmp));
      }
      // inclusive prefix maximum over the 16 lanes, then over the vectors before
      y = _mm256_max_epi16(y, ShiftUpN(y, vneg, 1));
      y = _mm256_max_epi16(y, ShiftUpN(y, vneg, 2));
      y = _mm256_max_epi16(y, ShiftUpN(y, vneg, 4));
      y = _mm256_max_epi16(y, ShiftUpN(y, vneg, 8));
      y = _mm256_max_epi16(y, _mm256_set1_epi16(carry));
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(row + j0), _mm256_add_epi16(y, off));
      alignas(32) std::int16_t last[16];
      _mm256_store_si256(reinterpret_cast<__m256i*>(last), y);
      carry = last[15];
    }
    row[0] = col0;
  }
  return true;
}
#else
bool AlignmentEngine::FillSimd16(const char*, std::uint32_t, const Graph&) { return false; }
#endif

Alignment AlignmentEngine::Align(const char* sequence, std::uint32_t sequence_len,
                                 const Graph& graph, std::int32_t* score) {
  if (graph.nodes().empty() || sequence_len == 0) {
    return Alignment();
  }
  const auto& rank_to_node = graph.rank_to_node();
  const std::uint64_t width = sequence_len + 1ULL;
  const std::uint64_t height = graph.nodes().size() + 1ULL;
  constexpr std::int32_t kNegInf = std::numeric_limits<std::int32_t>::min() + 1024;
  cells_ += width * height;

  // profile[code][j] = score of aligning sequence[j-1] with a node of that code
  profile_.assign(graph.num_codes() * width, 0);
  for (std::uint32_t c = 0; c < graph.num_codes(); ++c) {
    for (std::uint64_t j = 0; j < sequence_len; ++j) {
      profile_[c * width + (j + 1)] =
          (static_cast<std::int32_t>(c) == graph.coder(sequence[j])) ? m_ : n_;
    }
  }
  node_id_to_rank_.assign(graph.nodes().size(), 0);
  for (std::uint32_t i = 0; i < rank_to_node.size(); ++i) {
    node_id_to_rank_[rank_to_node[i]->id] = i;
  }

  const bool simd = g_use_simd && FillSimd16(sequence, sequence_len, graph);
  const std::uint64_t stride16 = ((width + 15) / 16) * 16 + 16;
  auto HH = [&](std::uint64_t i, std::uint64_t j) -> std::int32_t {
    return simd ? static_cast<std::int32_t>(H16_[i * stride16 + 16 + j]) : H_[i * width + j];
  };
  std::int32_t max_score = kNegInf;
  std::int64_t max_i = -1, max_j = -1;
  if (!simd) {
    H_.assign(width * height, 0);
    for (std::uint64_t j = 1; j < width; ++j) {
      H_[j] = static_cast<std::int32_t>(j) * g_;
    }
    for (std::uint64_t i = 1; i < height; ++i) {
      const auto& edges = rank_to_node[i - 1]->inedges;
      std::int32_t penalty = edges.empty() ? 0 : kNegInf;
      for (const auto& it : edges) {
        const std::uint64_t pred_i = node_id_to_rank_[it->tail->id] + 1;
        penalty = std::max(penalty, H_[pred_i * width]);
      }
      H_[i * width] = penalty + g_;
    }

    for (const auto& it : rank_to_node) {
      const std::int32_t* prof = &profile_[it->code * width];
      const std::uint64_t i = node_id_to_rank_[it->id] + 1;
      std::uint64_t pred_i =
          it->inedges.empty() ? 0 : node_id_to_rank_[it->inedges[0]->tail->id] + 1;
      std::int32_t* row = &H_[i * width];
      const std::int32_t* pred = &H_[pred_i * width];
      for (std::uint64_t j = 1; j < width; ++j) {
        row[j] = std::max(pred[j - 1] + prof[j], pred[j] + g_);
      }
      for (std::uint32_t p = 1; p < it->inedges.size(); ++p) {
        pred_i = node_id_to_rank_[it->inedges[p]->tail->id] + 1;
        pred = &H_[pred_i * width];
        for (std::uint64_t j = 1; j < width; ++j) {
          row[j] = std::max(pred[j - 1] + prof[j], std::max(row[j], pred[j] + g_));
        }
      }
      for (std::uint64_t j = 1; j < width; ++j) {
        row[j] = std::max(row[j - 1] + g_, row[j]);
      }
      if (it->outedges.empty() && max_score < row[width - 1]) {  // sinks, first maximum
        max_score = row[width - 1];
        max_i = i;
        max_j = width - 1;
      }
    }
  } else {
    for (const auto& it : rank_to_node) {  // sinks, first maximum (rank order)
      const std::uint64_t i = node_id_to_rank_[it->id] + 1;
      if (it->outedges.empty() && max_score < HH(i, width - 1)) {
        max_score = HH(i, width - 1);
        max_i = i;
        max_j = width - 1;
      }
    }
  }
  if (max_i == -1 && max_j == -1) {
    return Alignment();
  }
  if (score) {
    *score = max_score;
  }

  // traceback: diagonal (predecessors in in-edge order), vertical, horizontal
  Alignment alignment;
  std::uint64_t i = max_i, j = max_j;
  std::uint64_t prev_i = 0, prev_j = 0;
  while (!(i == 0 && j == 0)) {
    const std::int32_t h = HH(i, j);
    bool found = false;
    if (i != 0 && j != 0) {
      const auto& it = rank_to_node[i - 1];
      const std::int32_t match = profile_[it->code * width + j];
      const std::size_t np = std::max<std::size_t>(1, it->inedges.size());
      for (std::size_t p = 0; p < np; ++p) {
        const std::uint64_t pred_i =
            it->inedges.empty() ? 0 : node_id_to_rank_[it->inedges[p]->tail->id] + 1;
        if (h == HH(pred_i, j - 1) + match) {
          prev_i = pred_i;
          prev_j = j - 1;
          found = true;
          break;
        }
      }
    }
    if (!found && i != 0) {
      const auto& it = rank_to_node[i - 1];
      const std::size_t np = std::max<std::size_t>(1, it->inedges.size());
      for (std::size_t p = 0; p < np; ++p) {
        const std::uint64_t pred_i =
            it->inedges.empty() ? 0 : node_id_to_rank_[it->inedges[p]->tail->id] + 1;
        if (h == HH(pred_i, j) + g_) {
          prev_i = pred_i;
          prev_j = j;
          found = true;
          break;
        }
      }
    }
    if (!found && j != 0 && h == HH(i, j - 1) + g_) {
      prev_i = i;
      prev_j = j - 1;
      found = true;
    }
    alignment.emplace_back(
        i == prev_i ? -1 : static_cast<std::int32_t>(rank_to_node[i - 1]->id),
        j == prev_j ? -1 : static_cast<std::int32_t>(j - 1));
    i = prev_i;
    j = prev_j;
  }
  std::reverse(alignment.begin(), alignment.end());
  return alignment;
}

}  // namespace spoa
