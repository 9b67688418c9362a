// raven-b200: drop-in header for the un-vendored dependency `biosoup`.
// Field order is pinned by the reference's aggregate use in
// RavenLib/src/overlap_utils.cc:5-8 and PythonLib/src/ravenpy.cc:96-122.
#ifndef BIOSOUP_OVERLAP_HPP_
#define BIOSOUP_OVERLAP_HPP_

#include <cstdint>
#include <string>

namespace biosoup {

struct Overlap {
  Overlap() = default;

  Overlap(std::uint32_t lhs_id, std::uint32_t lhs_begin, std::uint32_t lhs_end,
          std::uint32_t rhs_id, std::uint32_t rhs_begin, std::uint32_t rhs_end,
          std::uint32_t score, bool strand = true)
      : lhs_id(lhs_id),
        lhs_begin(lhs_begin),
        lhs_end(lhs_end),
        rhs_id(rhs_id),
        rhs_begin(rhs_begin),
        rhs_end(rhs_end),
        score(score),
        strand(strand),
        alignment() {}

  Overlap(std::uint32_t lhs_id, std::uint32_t lhs_begin, std::uint32_t lhs_end,
          std::uint32_t rhs_id, std::uint32_t rhs_begin, std::uint32_t rhs_end,
          std::uint32_t score, const std::string& alignment,
          bool strand = true)
      : lhs_id(lhs_id),
        lhs_begin(lhs_begin),
        lhs_end(lhs_end),
        rhs_id(rhs_id),
        rhs_begin(rhs_begin),
        rhs_end(rhs_end),
        score(score),
        strand(strand),
        alignment(alignment) {}

  std::uint32_t lhs_id;
  std::uint32_t lhs_begin;
  std::uint32_t lhs_end;
  std::uint32_t rhs_id;
  std::uint32_t rhs_begin;
  std::uint32_t rhs_end;
  std::uint32_t score;  // based on k-mer matches or alignment score
  bool strand;          // (optional) Watson-Crick strand
  std::string alignment;
};

}  // namespace biosoup

#endif  // BIOSOUP_OVERLAP_HPP_
