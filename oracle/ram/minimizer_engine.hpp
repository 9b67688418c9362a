// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the
// product path (raven_b200/); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of `ram::MinimizerEngine`, the un-vendored dependency that
// holds the reference's overlap arithmetic. ram is NOT in /root/reference
// (pulled transitively by racon's floating `library` branch,
// Raven.deps.cmake:39-44; version not stated anywhere in the tree, believed
// 2.1.x). What follows restates its published algorithm
// and is anchored on the reference's own call sites:
//   ctor      RavenLib/src/construct.cc:661-662, assemble.cc:753
//   Minimize  RavenLib/src/construct.cc:42-43,363
//   Filter    RavenLib/src/construct.cc:44,372
//   Map       RavenLib/src/construct.cc:62,377-381
// PARITY UNPINNED per stage: upstream holds no golden sketch / overlap vectors
// (SURVEY.md §8c); the only upstream pin is end-to-end (raven_test.cpp:66).
#ifndef RAM_MINIMIZER_ENGINE_HPP_
#define RAM_MINIMIZER_ENGINE_HPP_

#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "biosoup/overlap.hpp"
#include "thread_pool/thread_pool.hpp"

namespace ram {

class MinimizerEngine {
 public:
  MinimizerEngine(std::shared_ptr<thread_pool::ThreadPool> thread_pool = nullptr,
                  std::uint32_t k = 15,  // element of [1, 31]
                  std::uint32_t w = 5,
                  std::uint32_t bandwidth = 500,
                  std::uint32_t chain = 4,
                  std::uint32_t matches = 100,
                  std::uint32_t gap = 10000);

  MinimizerEngine(const MinimizerEngine&) = delete;
  MinimizerEngine& operator=(const MinimizerEngine&) = delete;
  MinimizerEngine(MinimizerEngine&&) = default;
  MinimizerEngine& operator=(MinimizerEngine&&) = default;
  ~MinimizerEngine() = default;

  // transform set of sequences to minimizer index
  // minhash = pick only the smallest sequence->data.size() / k minimizers
  void Minimize(
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last,
      bool minhash = false);

  // set occurrence frequency threshold
  void Filter(double frequency);

  // find overlaps in preconstructed minimizer index
  // micromizers = smallest sequence->data.size() / k minimizers
  std::vector<biosoup::Overlap> Map(
      const std::unique_ptr<biosoup::NucleicAcid>& sequence,
      bool avoid_equal,      // ignore overlaps in which lhs_id == rhs_id
      bool avoid_symmetric,  // ignore overlaps in which lhs_id > rhs_id
      bool minhash = false,  // only lhs
      std::vector<std::uint32_t>* filtered = nullptr) const;

  // ---- record types (exposed: the oracle's tests compare them 1:1 with the
  //      GPU records) ----
  struct Kmer {
    Kmer() = default;
    Kmer(std::uint64_t value, std::uint64_t origin)
        : value(value), origin(origin) {}
    std::uint32_t id() const { return static_cast<std::uint32_t>(origin >> 32); }
    std::uint32_t position() const {
      return static_cast<std::uint32_t>(origin) >> 1;
    }
    bool strand() const { return origin & 1; }
    std::uint64_t value;
    std::uint64_t origin;
  };

  struct Match {
    Match() = default;
    Match(std::uint64_t group, std::uint64_t positions)
        : group(group), positions(positions) {}
    std::uint32_t rhs_id() const {
      return static_cast<std::uint32_t>(group >> 33);
    }
    bool strand() const { return (group >> 32) & 1; }
    std::uint32_t diagonal() const { return static_cast<std::uint32_t>(group); }
    std::uint32_t lhs_position() const {
      return static_cast<std::uint32_t>(positions >> 32);
    }
    std::uint32_t rhs_position() const {
      return static_cast<std::uint32_t>(positions);
    }
    std::uint64_t group;
    std::uint64_t positions;
  };

  // ---- oracle-only introspection ----
  std::vector<Kmer> Sketch(const std::unique_ptr<biosoup::NucleicAcid>& sequence,
                           bool minhash) const {
    return Minimize(sequence, minhash);
  }
  std::vector<Match> Matches(
      const std::unique_ptr<biosoup::NucleicAcid>& sequence, bool avoid_equal,
      bool avoid_symmetric, bool minhash,
      std::vector<std::uint32_t>* filtered = nullptr) const;
  std::vector<biosoup::Overlap> ChainMatches(std::uint64_t lhs_id,
                                             std::vector<Match>&& matches) const {
    return Chain(lhs_id, std::move(matches));
  }
  std::uint32_t occurrence() const { return occurrence_; }
  std::uint64_t num_keys() const;        // distinct minimizer values indexed
  std::uint64_t num_minimizers() const;  // records indexed
  // (value, count) of every distinct key, ascending by value
  void Keys(std::vector<std::uint64_t>* values,
            std::vector<std::uint32_t>* counts) const;
  std::uint32_t k() const { return k_; }
  std::uint32_t w() const { return w_; }

 private:
  // one bucket of the index: keys sorted ascending, postings grouped per key
  // in insertion order (= read order, then position order)
  struct Index {
    std::uint32_t Find(std::uint64_t key, const std::uint64_t** dst) const;
    std::vector<std::uint64_t> keys;
    std::vector<std::uint64_t> begins;  // keys.size() + 1
    std::vector<std::uint64_t> origins;
  };

  std::vector<Kmer> Minimize(
      const std::unique_ptr<biosoup::NucleicAcid>& sequence,
      bool minhash = false) const;

  std::vector<biosoup::Overlap> Chain(std::uint64_t lhs_id,
                                      std::vector<Match>&& matches) const;

  std::uint32_t k_;
  std::uint32_t w_;
  std::uint32_t bandwidth_;
  std::uint32_t chain_;
  std::uint32_t matches_;
  std::uint64_t gap_;
  std::uint32_t occurrence_;
  std::vector<Index> index_;
  std::shared_ptr<thread_pool::ThreadPool> thread_pool_;
};

}  // namespace ram

#endif  // RAM_MINIMIZER_ENGINE_HPP_
