// ORACLE — TEST INFRASTRUCTURE ONLY. See stage1_port.cpp.
#ifndef ORACLE_STAGE1_PORT_HPP_
#define ORACLE_STAGE1_PORT_HPP_

#include <cstdint>
#include <memory>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "biosoup/overlap.hpp"
#include "ram/minimizer_engine.hpp"
#include "thread_pool/thread_pool.hpp"

namespace oracle {

biosoup::Overlap ReverseOverlap(const biosoup::Overlap& o);
std::uint32_t OverlapLength(const biosoup::Overlap& o);

void AddLayers(std::uint32_t id, std::vector<std::uint16_t>& data,
               const biosoup::Overlap* first, const biosoup::Overlap* last);

struct Stage1Result {
  std::vector<std::vector<biosoup::Overlap>> overlaps;  // per read, <= kMax
  std::vector<std::vector<std::uint16_t>> piles;        // Pile::data_
  std::vector<std::uint32_t> occurrences;               // per index batch
  std::uint64_t num_mapped = 0;  // overlaps returned by all Map calls
};

Stage1Result FindOverlapsAndCreatePiles(
    const std::shared_ptr<thread_pool::ThreadPool>& pool,
    ram::MinimizerEngine& engine,
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& reads,
    double freq, std::size_t max_overlaps, bool minhash,
    std::uint64_t index_batch_bases = 1ULL << 32,
    std::uint64_t query_batch_bases = 1ULL << 30);

}  // namespace oracle

#endif  // ORACLE_STAGE1_PORT_HPP_
