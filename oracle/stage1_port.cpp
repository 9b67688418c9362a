// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("port") of the stage-1 all-vs-all loop of the reference:
//   raven::FindOverlapsAndCreatePiles   RavenLib/src/construct.cc:14-121
//   raven::Pile::Pile / AddLayers       RavenLib/src/pile.cc:19-62
//   raven::OverlapReverse / GetOverlapLength   RavenLib/src/overlap_utils.cc:5-12
// Pinned against the reference's own compiled sources (oracle/_ref, built by
// oracle/Makefile from /root/reference in place) by tests/test_oracle_ref.py.
// The batch thresholds (2^32 index bases, 2^30 query bases) are parameters so
// that the multi-batch / multi-flush schedule can be exercised at test sizes.
#include "stage1_port.hpp"

#include <algorithm>
#include <future>

namespace oracle {

biosoup::Overlap ReverseOverlap(const biosoup::Overlap& o) {  // overlap_utils.cc:5-8
  return biosoup::Overlap(o.rhs_id, o.rhs_begin, o.rhs_end, o.lhs_id,
                          o.lhs_begin, o.lhs_end, o.score, o.strand);
}

std::uint32_t OverlapLength(const biosoup::Overlap& o) {  // overlap_utils.cc:10-12
  return std::max(o.rhs_end - o.rhs_begin, o.lhs_end - o.lhs_begin);
}

// pile.cc:33-62 with kPSS = 4 (pile.h:21): +1 on bins [(b>>4)+1, (e>>4)-1) per
// overlap side that belongs to read `id`, saturating at 65535, done as the
// reference does it: a sweep over sorted begin/end marks (begin marks sort
// before end marks of the same bin; the depth counter is unsigned and wraps
// exactly like the reference's on degenerate intervals).
void AddLayers(std::uint32_t id, std::vector<std::uint16_t>& data,
               const biosoup::Overlap* first, const biosoup::Overlap* last) {
  if (first >= last) {
    return;
  }
  std::vector<std::uint32_t> marks;
  for (auto it = first; it != last; ++it) {
    if (it->lhs_id == id) {
      marks.emplace_back(((it->lhs_begin >> 4) + 1) << 1);
      marks.emplace_back(((it->lhs_end >> 4) - 1) << 1 | 1);
    } else if (it->rhs_id == id) {
      marks.emplace_back(((it->rhs_begin >> 4) + 1) << 1);
      marks.emplace_back(((it->rhs_end >> 4) - 1) << 1 | 1);
    }
  }
  std::sort(marks.begin(), marks.end());
  std::uint32_t depth = 0, prev = 0;
  for (auto m : marks) {
    if (depth > 0) {
      for (std::uint32_t b = prev; b < (m >> 1); ++b) {
        std::uint32_t v = data[b] + depth;
        data[b] = v < 65535U ? v : 65535U;
      }
    }
    prev = m >> 1;
    depth += (m & 1) ? -1 : 1;
  }
}

Stage1Result FindOverlapsAndCreatePiles(
    const std::shared_ptr<thread_pool::ThreadPool>& pool,
    ram::MinimizerEngine& engine,
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& reads,
    double freq, std::size_t max_overlaps, bool minhash,
    std::uint64_t index_batch_bases, std::uint64_t query_batch_bases) {
  Stage1Result res;
  const std::uint32_t n = reads.size();
  res.piles.resize(n);
  for (std::uint32_t i = 0; i < n; ++i) {  // pile.cc:19-31: len >> 4 bins
    res.piles[i].assign(reads[i]->inflated_len >> 4, 0);
  }
  res.overlaps.resize(n);

  std::uint64_t bases = 0;
  for (std::uint32_t last = 0, first = 0; last < n; ++last) {
    bases += reads[last]->inflated_len;
    if (last != n - 1 && bases < index_batch_bases) {
      continue;
    }
    bases = 0;

    engine.Minimize(reads.begin() + first, reads.begin() + last + 1, minhash);
    engine.Filter(freq);
    res.occurrences.emplace_back(engine.occurrence());

    std::vector<std::uint32_t> seen(n);
    for (std::uint32_t r = 0; r < n; ++r) {
      seen[r] = res.overlaps[r].size();
    }

    // every read up to the end of this index batch is a query
    std::vector<std::future<std::vector<biosoup::Overlap>>> pending;
    for (std::uint32_t q = 0; q <= last; ++q) {
      pending.emplace_back(pool->Submit(
          [&](std::uint32_t q) { return engine.Map(reads[q], true, true, true); },
          q));
      bases += reads[q]->inflated_len;
      if (q != last && bases < query_batch_bases) {
        continue;
      }
      bases = 0;

      // serial gather in query order, forward then mirrored record
      for (auto& f : pending) {
        for (const auto& o : f.get()) {
          ++res.num_mapped;
          res.overlaps[o.lhs_id].emplace_back(o);
          res.overlaps[o.rhs_id].emplace_back(ReverseOverlap(o));
        }
      }
      pending.clear();

      std::vector<std::future<void>> tasks;
      for (std::uint32_t r = 0; r < n; ++r) {
        if (res.overlaps[r].empty() || res.overlaps[r].size() == seen[r]) {
          continue;
        }
        tasks.emplace_back(pool->Submit(
            [&](std::uint32_t r) {
              auto& list = res.overlaps[r];
              AddLayers(r, res.piles[r], list.data() + seen[r],
                        list.data() + list.size());
              seen[r] = std::min(list.size(), max_overlaps);
              if (list.size() < max_overlaps) {
                return;
              }
              // same call as construct.cc:98-102: unstable, key not unique
              std::sort(list.begin(), list.end(),
                        [](const biosoup::Overlap& a, const biosoup::Overlap& b) {
                          return OverlapLength(a) > OverlapLength(b);
                        });
              std::vector<biosoup::Overlap> kept(list.begin(),
                                                 list.begin() + max_overlaps);
              kept.swap(list);
            },
            r));
      }
      for (auto& t : tasks) {
        t.wait();
      }
    }
    first = last + 1;
  }
  return res;
}

}  // namespace oracle
