// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of racon's per-window consensus (racon::Window), the unit the
// "POA windows/s" metric counts. racon is an un-vendored dependency of the
// reference (`GIT_TAG library`, Raven.deps.cmake:39-44); the reference reaches
// it through racon::Polisher::Polish (RavenLib/src/polish.cc:43-51) with
// w = 500, trim = true, e = 0.3 (polish.cc:44). Restated from its published
// behaviour. PARITY UNPINNED per stage.
#ifndef ORACLE_RACON_WINDOW_HPP_
#define ORACLE_RACON_WINDOW_HPP_

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "spoa/spoa.hpp"

namespace racon {

enum class WindowType { kNGS, kTGS };

class Window {
 public:
  Window(std::uint64_t id, std::uint32_t rank, WindowType type, const char* backbone,
         std::uint32_t backbone_len, const char* quality, std::uint32_t quality_len);

  // layer = read segment aligned to backbone positions [begin, end]
  void AddLayer(const char* sequence, std::uint32_t sequence_len, const char* quality,
                std::uint32_t quality_len, std::uint32_t begin, std::uint32_t end);

  // false = fewer than 3 sequences: the consensus is the backbone ("unpolished")
  bool GenerateConsensus(spoa::AlignmentEngine* engine, bool trim);

  const std::string& consensus() const { return consensus_; }
  const std::vector<std::uint32_t>& coverages() const { return coverages_; }
  std::uint64_t id() const { return id_; }
  std::uint32_t rank() const { return rank_; }
  std::size_t num_sequences() const { return sequences_.size(); }
  bool chimeric_warning() const { return chimeric_warning_; }

 private:
  std::uint64_t id_;
  std::uint32_t rank_;
  WindowType type_;
  std::string consensus_;
  std::vector<std::uint32_t> coverages_;
  std::vector<std::pair<const char*, std::uint32_t>> sequences_;
  std::vector<std::pair<const char*, std::uint32_t>> qualities_;
  std::vector<std::pair<std::uint32_t, std::uint32_t>> positions_;
  bool chimeric_warning_ = false;
};

}  // namespace racon

#endif  // ORACLE_RACON_WINDOW_HPP_
