// raven-b200: drop-in `ram/minimizer_engine.hpp`.
//
// The reference includes this header from the un-vendored `ram` library
// (RavenLib/include/raven/graph/construct.h:1) and uses
//   ram::MinimizerEngine{thread_pool, k, w}        construct.cc:661-662, assemble.cc:753
//   Minimize(first, last, minhash)                 construct.cc:42-43,363; assemble.cc:754,777
//   Filter(frequency)                              construct.cc:44,372; assemble.cc:755,778
//   Map(sequence, avoid_equal, avoid_symmetric, minhash, &filtered)
//                                                  construct.cc:62,377-381; assemble.cc:757,780
// This class keeps those signatures and forwards to the B200 engine through
// the C ABI (include/raven_b200.h). `Map` stays callable concurrently from pool
// workers like the reference's (a mutex serialises the device calls); the
// batched entry points below are the fast path our FindOverlapsAndCreatePiles
// replacement uses (include/raven_b200/construct_b200.hpp).
#ifndef RAM_MINIMIZER_ENGINE_HPP_
#define RAM_MINIMIZER_ENGINE_HPP_

#include <cstdint>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "biosoup/overlap.hpp"
#include "thread_pool/thread_pool.hpp"

struct rvn_ctx;

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
  MinimizerEngine(MinimizerEngine&&) noexcept;
  MinimizerEngine& operator=(MinimizerEngine&&) noexcept;
  ~MinimizerEngine();

  // transform set of sequences to minimizer index
  // minhash = pick only the smallest sequence->data.size() / k minimizers
  void Minimize(
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last,
      bool minhash = false);

  // set occurrence frequency threshold (throws std::invalid_argument)
  void Filter(double frequency);

  // find overlaps in preconstructed minimizer index
  std::vector<biosoup::Overlap> Map(
      const std::unique_ptr<biosoup::NucleicAcid>& sequence,
      bool avoid_equal,      // ignore overlaps in which lhs_id == rhs_id
      bool avoid_symmetric,  // ignore overlaps in which lhs_id > rhs_id
      bool minhash = false,  // only lhs
      std::vector<std::uint32_t>* filtered = nullptr) const;

  // ---- B200 extensions ----
  // the whole read set on the device (ids must equal positions); lets stage 1
  // run without per-batch uploads
  void Upload(const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences);
  // any range, with its own ids (stage 2: the valid reads, sorted by id)
  void Upload(std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
              std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last);
  rvn_ctx* context() const { return ctx_; }
  std::mutex& mutex() const { return *mutex_; }
  std::uint32_t occurrence() const { return occurrence_; }

 private:
  void UploadRange(
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
      std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last);

  rvn_ctx* ctx_;
  std::unique_ptr<std::mutex> mutex_;
  std::uint32_t occurrence_;
  // id -> (position in the uploaded set, address of the uploaded object)
  std::unordered_map<std::uint32_t,
                     std::pair<std::uint32_t, const biosoup::NucleicAcid*>>
      uploaded_;
  std::shared_ptr<thread_pool::ThreadPool> thread_pool_;
};

}  // namespace ram

#endif  // RAM_MINIMIZER_ENGINE_HPP_
