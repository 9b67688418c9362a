// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the part of `spoa` (partial order alignment) that racon's
// window consensus uses; spoa is an un-vendored dependency of the reference
// (it arrives through racon's floating `library` branch, Raven.deps.cmake:
// 39-44; version not stated in the tree, believed 4.0.x). Restated from its
// published algorithm:
//   Graph::AddAlignment / TopologicalSort / Subgraph / UpdateAlignment /
//   GenerateConsensus (heaviest bundle + branch completion, coverages)
//   AlignmentEngine (global = kNW, linear gaps): sequence-to-DAG DP with the
//   traceback preference diagonal > vertical > horizontal, predecessors in
//   in-edge order.
// The reference reaches it only through racon::Polisher::Polish
// (RavenLib/src/polish.cc:43-51). PARITY UNPINNED: upstream holds no golden
// vectors for this path.
#ifndef ORACLE_SPOA_HPP_
#define ORACLE_SPOA_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace spoa {

using Alignment = std::vector<std::pair<std::int32_t, std::int32_t>>;

class Graph {
 public:
  struct Edge;
  struct Node {
    Node(std::uint32_t id, std::uint32_t code) : id(id), code(code) {}
    std::uint32_t Coverage() const;  // distinct sequence labels on incident edges
    std::uint32_t id;
    std::uint32_t code;
    std::vector<Edge*> inedges;
    std::vector<Edge*> outedges;
    std::vector<Node*> aligned_nodes;
  };
  struct Edge {
    Edge(Node* tail, Node* head, std::uint32_t label, std::uint32_t weight)
        : tail(tail), head(head), labels(1, label), weight(weight) {}
    void AddSequence(std::uint32_t label, std::uint32_t w) {
      labels.emplace_back(label);
      weight += w;
    }
    Node* tail;
    Node* head;
    std::vector<std::uint32_t> labels;
    std::int64_t weight;
  };

  Graph();
  Graph(const Graph&) = delete;
  Graph& operator=(const Graph&) = delete;
  Graph(Graph&&) = default;
  Graph& operator=(Graph&&) = default;

  // weights: one per base (quality - 33), or all 1 when absent
  void AddAlignment(const Alignment& alignment, const char* sequence,
                    std::uint32_t sequence_len,
                    const std::vector<std::uint32_t>& weights);
  void AddAlignment(const Alignment& alignment, const char* sequence,
                    std::uint32_t sequence_len, const char* quality);
  void AddAlignment(const Alignment& alignment, const char* sequence,
                    std::uint32_t sequence_len, std::uint32_t weight = 1);

  // backward-reachable part between backbone positions [begin, end]
  Graph Subgraph(std::uint32_t begin, std::uint32_t end,
                 std::vector<const Node*>* subgraph_to_graph) const;
  void UpdateAlignment(const std::vector<const Node*>& subgraph_to_graph,
                       Alignment* alignment) const;

  std::string GenerateConsensus(std::vector<std::uint32_t>* coverages);

  const std::vector<std::unique_ptr<Node>>& nodes() const { return nodes_; }
  const std::vector<Node*>& rank_to_node() const { return rank_to_node_; }
  std::uint32_t num_codes() const { return num_codes_; }
  std::int32_t coder(char c) const { return coder_[static_cast<std::uint8_t>(c)]; }
  char decoder(std::uint32_t code) const { return decoder_[code]; }
  std::uint64_t num_edges() const { return edges_.size(); }

 private:
  Node* AddNode(std::uint32_t code);
  void AddEdge(Node* tail, Node* head, std::uint32_t weight);
  Node* AddSequence(const char* sequence, const std::vector<std::uint32_t>& weights,
                    std::uint32_t begin, std::uint32_t end);
  void TopologicalSort();
  void ExtractSubgraph(const Node* begin, const Node* end,
                       std::vector<bool>* dst) const;
  void TraverseHeaviestBundle();
  Node* BranchCompletion(std::uint32_t rank, std::vector<std::int64_t>* scores,
                         std::vector<std::int32_t>* predecessors);

  std::uint32_t num_codes_;
  std::vector<std::int32_t> coder_;
  std::vector<std::int32_t> decoder_;
  std::vector<Node*> sequences_;
  std::vector<std::unique_ptr<Node>> nodes_;
  std::vector<std::unique_ptr<Edge>> edges_;
  std::vector<Node*> rank_to_node_;
  std::vector<Node*> consensus_;
};

// global alignment (kNW) of a sequence to the graph, linear gap penalty g
class AlignmentEngine {
 public:
  AlignmentEngine(std::int8_t m, std::int8_t n, std::int8_t g) : m_(m), n_(n), g_(g) {}
  // fill the matrix with AVX2 int16 rows (16 cells per instruction, the horizontal
  // gap recurrence as a prefix maximum) like upstream spoa's SIMD engine, whenever
  // the worst-case score fits int16; same cell values, same traceback
  static void UseSimd(bool on);
  Alignment Align(const char* sequence, std::uint32_t sequence_len,
                  const Graph& graph, std::int32_t* score = nullptr);
  // DP cells evaluated so far (for GCUPS figures)
  std::uint64_t cells() const { return cells_; }

 private:
  std::int8_t m_, n_, g_;
  std::uint64_t cells_ = 0;
  bool FillSimd16(const char* sequence, std::uint32_t sequence_len, const Graph& graph);
  std::vector<std::int32_t> H_;
  std::vector<std::int16_t> H16_, profile16_;
  std::vector<std::int32_t> profile_;
  std::vector<std::uint32_t> node_id_to_rank_;
};

}  // namespace spoa

#endif  // ORACLE_SPOA_HPP_
