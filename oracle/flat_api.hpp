// ORACLE — TEST INFRASTRUCTURE ONLY.
// Flat (ctypes-friendly) plumbing shared by the oracle library and the
// compiled-reference library: read sets built from packed 2-bit words, and a
// "bag" of named byte arrays for variable-sized results.
#ifndef ORACLE_FLAT_API_HPP_
#define ORACLE_FLAT_API_HPP_

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "biosoup/nucleic_acid.hpp"
#include "biosoup/overlap.hpp"

struct orc_bag {
  std::map<std::string, std::vector<std::uint8_t>> arrays;

  template <typename T>
  void Put(const std::string& name, const std::vector<T>& v) {
    auto& dst = arrays[name];
    dst.resize(v.size() * sizeof(T));
    if (!v.empty()) {
      std::memcpy(dst.data(), v.data(), dst.size());
    }
  }
};

struct orc_reads {
  std::vector<std::unique_ptr<biosoup::NucleicAcid>> seqs;
};

// reads from the biosoup wire format: words[woff[i] .. woff[i+1]) hold read i
// (32 bases per u64, LSB first), len[i] bases, id = i; optional block quality
inline orc_reads* MakeReads(const std::uint64_t* words,
                            const std::uint64_t* woff, const std::uint32_t* len,
                            std::uint32_t n, const std::uint8_t* bq,
                            const std::uint64_t* bq_off) {
  auto* r = new orc_reads();
  r->seqs.reserve(n);
  for (std::uint32_t i = 0; i < n; ++i) {
    auto s = std::make_unique<biosoup::NucleicAcid>();
    s->id = i;
    s->name = std::to_string(i);
    s->deflated_data.assign(words + woff[i], words + woff[i + 1]);
    s->inflated_len = len[i];
    s->is_reverse_complement = false;
    if (bq) {
      s->block_quality.assign(bq + bq_off[i], bq + bq_off[i + 1]);
    }
    r->seqs.emplace_back(std::move(s));
  }
  return r;
}

// 8 x u32 per overlap: lhs_id lhs_begin lhs_end rhs_id rhs_begin rhs_end score strand
inline void PushOverlap(std::vector<std::uint32_t>& dst,
                        const biosoup::Overlap& o) {
  dst.insert(dst.end(), {o.lhs_id, o.lhs_begin, o.lhs_end, o.rhs_id,
                         o.rhs_begin, o.rhs_end, o.score,
                         static_cast<std::uint32_t>(o.strand)});
}

#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

#define ORC_BAG_ACCESSORS(prefix)                                             \
  ORC_EXPORT std::int64_t prefix##_bag_len(orc_bag* b, const char* name) {    \
    auto it = b->arrays.find(name);                                           \
    return it == b->arrays.end() ? -1                                         \
                                 : static_cast<std::int64_t>(it->second.size()); \
  }                                                                           \
  ORC_EXPORT const void* prefix##_bag_ptr(orc_bag* b, const char* name) {     \
    auto it = b->arrays.find(name);                                           \
    return it == b->arrays.end() ? nullptr : it->second.data();               \
  }                                                                           \
  ORC_EXPORT void prefix##_bag_free(orc_bag* b) { delete b; }                 \
  ORC_EXPORT orc_reads* prefix##_reads_create(                                \
      const std::uint64_t* words, const std::uint64_t* woff,                  \
      const std::uint32_t* len, std::uint32_t n, const std::uint8_t* bq,      \
      const std::uint64_t* bq_off) {                                          \
    return MakeReads(words, woff, len, n, bq, bq_off);                        \
  }                                                                           \
  ORC_EXPORT void prefix##_reads_free(orc_reads* r) { delete r; }

#endif  // ORACLE_FLAT_API_HPP_
