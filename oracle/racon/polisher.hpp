// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of `racon::Polisher` (un-vendored dependency, `GIT_TAG
// library`, Raven.deps.cmake:39-44) with the interface the reference uses:
//   racon::Polisher::Create(pool, q, e, w, trim, m, n, g, cuda_poa_batches,
//       cuda_banded_alignment, cuda_alignment_batches)   RavenLib/src/polish.cc:43-48
//   polisher->Polish(targets, sequences, drop_unpolished)   polish.cc:51
// Pipeline: map every read to the targets (ram, k=15 w=5,
// f=0.001), keep its longest overlap, drop it if the length ratio error > e,
// globally align read and target segments, cut the alignment at multiples of w
// into (target, read) breaking points, build one window per w target bases
// with the read segments as layers, consensus per window (racon::Window over
// spoa), stitch, tag names " LN:i: RC:i: XC:f:" (parsed by polish.cc:55-59).
// PARITY UNPINNED per stage; the alignment PATH among equally optimal ones is
// implementation defined upstream (edlib) and documented in nw_path.cpp.
#ifndef ORACLE_RACON_POLISHER_HPP_
#define ORACLE_RACON_POLISHER_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "thread_pool/thread_pool.hpp"

namespace racon {

class Polisher {
 public:
  static std::unique_ptr<Polisher> Create(
      std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q, double e,
      std::uint32_t w, bool trim, std::int8_t m, std::int8_t n, std::int8_t g,
      std::uint32_t cuda_poa_batches = 0, bool cuda_banded_alignment = false,
      std::uint32_t cuda_alignment_batches = 0);

  std::vector<std::unique_ptr<biosoup::NucleicAcid>> Polish(
      const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& targets,
      const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences,
      bool drop_unpolished);

  // counters of the last Polish() (benchmarks)
  std::uint64_t num_windows() const { return num_windows_; }
  std::uint64_t num_polished_windows() const { return num_polished_windows_; }

 private:
  Polisher(std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q, double e,
           std::uint32_t w, bool trim, std::int8_t m, std::int8_t n, std::int8_t g);

  std::shared_ptr<thread_pool::ThreadPool> thread_pool_;
  double q_, e_;
  std::uint32_t w_;
  bool trim_;
  std::int8_t m_, n_, g_;
  std::uint64_t num_windows_ = 0, num_polished_windows_ = 0;
};

// global alignment path of query vs target: 'M' (column with both), 'I' (query
// only), 'D' (target only) per alignment column
std::string GlobalAlignmentPath(const std::string& query, const std::string& target);

std::int64_t GlobalDistance(const std::string& query, const std::string& target);

// (target, query) breaking points of an alignment cut at multiples of `w`
std::vector<std::pair<std::uint32_t, std::uint32_t>> BreakingPoints(
    const std::string& path, std::uint32_t q_begin, std::uint32_t t_begin,
    std::uint32_t t_end, std::uint32_t w);

}  // namespace racon

#endif  // ORACLE_RACON_POLISHER_HPP_
