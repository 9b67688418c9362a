// raven-b200: B200 replacement of raven::FindOverlapsAndCreatePiles
// (RavenLib/src/construct.cc:14-121) with the reference's exact signature
// (RavenLib/include/raven/graph/construct.h:22-29). One call = the whole stage
// on the device (sketch, index, filter, map/chain, piles, gather, truncation);
// the results are written back into RavenLib's own containers.
//
// Needs RavenLib's headers ("raven/pile.h") on the include path: it is meant to
// be compiled inside RavenLib, replacing the body of
// construct.cc:14-121 with a call to raven_b200::FindOverlapsAndCreatePiles.
#ifndef RAVEN_B200_CONSTRUCT_B200_HPP_
#define RAVEN_B200_CONSTRUCT_B200_HPP_

#include <cstdint>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "biosoup/overlap.hpp"
#include "biosoup/timer.hpp"
#include "cereal/access.hpp"
#include "ram/minimizer_engine.hpp"
#include "raven/pile.h"
#include "raven_b200.h"
#include "thread_pool/thread_pool.hpp"

namespace raven_b200 {

namespace detail {

// raven::Pile keeps its histogram private and only lets AddLayers touch it;
// the one door it leaves open is `friend cereal::access` + serialize(), which
// hands out references to every field. This "archive" uses it to move the
// device-computed histogram in.
struct PileDoor {
  const std::uint16_t* src;
  std::size_t n;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t&, std::uint32_t&, std::uint16_t&,
                  bool&, bool&, bool&, bool&, std::vector<std::uint16_t>& data,
                  Ts&...) {
    if (data.size() != n) {
      throw std::logic_error("[raven_b200] pile size mismatch");
    }
    data.assign(src, src + n);
  }
};

}  // namespace detail

inline void FindOverlapsAndCreatePiles(
    const std::shared_ptr<thread_pool::ThreadPool>& /*thread_pool*/,
    ram::MinimizerEngine& minimizer_engine,
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences,
    double freq, std::vector<std::unique_ptr<raven::Pile>>& piles,
    std::vector<std::vector<biosoup::Overlap>>& overlaps,
    std::size_t kMaxNumOverlaps = 32, bool useMinhash = false) {
  piles.reserve(sequences.size());
  for (const auto& it : sequences) {
    piles.emplace_back(new raven::Pile(it->id, it->inflated_len));
  }
  if (overlaps.size() < sequences.size()) overlaps.resize(sequences.size());

  biosoup::Timer timer;
  timer.Start();
  minimizer_engine.Upload(sequences);  // ids must equal positions (they do in raven)
  std::lock_guard<std::mutex> lock(minimizer_engine.mutex());
  rvn_ctx* ctx = minimizer_engine.context();
  int rc = rvn_find_overlaps_and_create_piles(ctx, freq, kMaxNumOverlaps,
                                              useMinhash, 0, 0);
  if (rc == RVN_ERR_INVALID) throw std::invalid_argument(rvn_last_error(ctx));
  if (rc != RVN_OK) throw std::runtime_error(rvn_last_error(ctx));

  const rvn_overlap* o = nullptr;
  const std::uint64_t *ooff = nullptr, *poff = nullptr;
  const std::uint16_t* pile = nullptr;
  std::uint64_t n_mapped = 0;
  rc = rvn_stage1_results(ctx, &o, &ooff, &pile, &poff, &n_mapped);
  if (rc != RVN_OK) throw std::runtime_error(rvn_last_error(ctx));

  for (std::size_t i = 0; i < sequences.size(); ++i) {
    auto& dst = overlaps[i];
    for (std::uint64_t e = ooff[i]; e < ooff[i + 1]; ++e) {
      dst.emplace_back(o[e].lhs_id, o[e].lhs_begin, o[e].lhs_end, o[e].rhs_id,
                       o[e].rhs_begin, o[e].rhs_end, o[e].score, o[e].strand != 0);
    }
    detail::PileDoor door{pile + poff[i], static_cast<std::size_t>(poff[i + 1] - poff[i])};
    cereal::access::member_serialize(door, *piles[i]);
  }
  std::cerr << "[raven::Graph::Construct] minimized + mapped sequences (B200) "
            << std::fixed << timer.Stop() << "s" << std::endl;
}

}  // namespace raven_b200

#endif  // RAVEN_B200_CONSTRUCT_B200_HPP_
