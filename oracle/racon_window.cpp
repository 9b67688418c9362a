// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/racon/window.hpp).
#include "racon/window.hpp"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>

namespace racon {

Window::Window(std::uint64_t id, std::uint32_t rank, WindowType type,
               const char* backbone, std::uint32_t backbone_len, const char* quality,
               std::uint32_t quality_len)
    : id_(id), rank_(rank), type_(type) {
  sequences_.emplace_back(backbone, backbone_len);
  qualities_.emplace_back(quality, quality_len);
  positions_.emplace_back(0, 0);
}

void Window::AddLayer(const char* sequence, std::uint32_t sequence_len,
                      const char* quality, std::uint32_t quality_len,
                      std::uint32_t begin, std::uint32_t end) {
  if (sequence_len == 0 || begin == end) {
    return;
  }
  if (quality != nullptr && sequence_len != quality_len) {
    throw std::invalid_argument(
        "[racon::Window::AddLayer] error: unequal quality size");
  }
  if (begin >= end || begin > sequences_.front().second ||
      end > sequences_.front().second) {
    throw std::invalid_argument(
        "[racon::Window::AddLayer] error: layer begin and end positions are invalid");
  }
  sequences_.emplace_back(sequence, sequence_len);
  qualities_.emplace_back(quality, quality_len);
  positions_.emplace_back(begin, end);
}

bool Window::GenerateConsensus(spoa::AlignmentEngine* engine, bool trim) {
  if (sequences_.size() < 3) {
    consensus_ = std::string(sequences_.front().first, sequences_.front().second);
    return false;
  }

  spoa::Graph graph{};
  if (qualities_.front().first == nullptr) {
    // racon gives the backbone a dummy quality of '!' (weight 0) in any case
    graph.AddAlignment(spoa::Alignment(), sequences_.front().first,
                       sequences_.front().second, 0U);
  } else {
    graph.AddAlignment(spoa::Alignment(), sequences_.front().first,
                       sequences_.front().second, qualities_.front().first);
  }

  // layers by begin position (stable: equal begins keep their arrival order)
  std::vector<std::uint32_t> rank;
  for (std::uint32_t i = 0; i < sequences_.size(); ++i) {
    rank.emplace_back(i);
  }
  auto by_begin = [&](std::uint32_t lhs, std::uint32_t rhs) {
    return positions_[lhs].first < positions_[rhs].first;
  };
  if (std::getenv("ORC_WINDOW_UNSTABLE_SORT")) {  // experiment knob
    std::sort(rank.begin() + 1, rank.end(), by_begin);
  } else {
    std::stable_sort(rank.begin() + 1, rank.end(), by_begin);
  }

  const std::uint32_t offset = 0.01 * sequences_.front().second;
  for (std::uint32_t j = 1; j < sequences_.size(); ++j) {
    const std::uint32_t i = rank[j];
    spoa::Alignment alignment;
    if (positions_[i].first < offset &&
        positions_[i].second > sequences_.front().second - offset) {
      alignment = engine->Align(sequences_[i].first, sequences_[i].second, graph);
    } else {
      std::vector<const spoa::Graph::Node*> mapping;
      auto subgraph = graph.Subgraph(positions_[i].first, positions_[i].second, &mapping);
      alignment = engine->Align(sequences_[i].first, sequences_[i].second, subgraph);
      subgraph.UpdateAlignment(mapping, &alignment);
    }
    if (qualities_[i].first == nullptr) {
      graph.AddAlignment(alignment, sequences_[i].first, sequences_[i].second);
    } else {
      graph.AddAlignment(alignment, sequences_[i].first, sequences_[i].second,
                         qualities_[i].first);
    }
  }

  coverages_.clear();
  consensus_ = graph.GenerateConsensus(&coverages_);

  if (type_ == WindowType::kTGS && trim) {
    const std::uint32_t average_coverage = (sequences_.size() - 1) / 2;
    std::int32_t begin = 0, end = consensus_.size() - 1;
    for (; begin < static_cast<std::int32_t>(consensus_.size()); ++begin) {
      if (coverages_[begin] >= average_coverage) {
        break;
      }
    }
    for (; end >= 0; --end) {
      if (coverages_[end] >= average_coverage) {
        break;
      }
    }
    if (begin >= end) {
      chimeric_warning_ = true;  // upstream warns "might be chimeric" and keeps it all
    } else {
      consensus_ = consensus_.substr(begin, end - begin + 1);
      // keep the coverages aligned with the returned letters
      coverages_ = std::vector<std::uint32_t>(coverages_.begin() + begin,
                                              coverages_.begin() + end + 1);
    }
  } else if (type_ == WindowType::kNGS) {
    std::uint32_t i = 0;
    for (; i < consensus_.size() && consensus_[i] == 'N'; ++i) {}
    std::uint32_t j = consensus_.size();
    for (; j > i && consensus_[j - 1] == 'N'; --j) {}
    consensus_ = consensus_.substr(i, j - i);
    coverages_ = std::vector<std::uint32_t>(coverages_.begin() + i, coverages_.begin() + j);
  }
  return true;
}

}  // namespace racon
