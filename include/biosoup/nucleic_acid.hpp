// raven-b200: drop-in header for the un-vendored dependency `biosoup`.
//
// The reference (lbcb-sci/raven) includes "biosoup/nucleic_acid.hpp" from a
// FetchContent dependency that is not in its tree (Raven.deps.cmake:39-44).
// This is our own implementation of the public surface RavenLib uses:
//   fields            RavenLib/include/raven/graph/graph.h:13-18
//   ctors / methods   RavenLib/src/construct.cc:185, common.cc:248,
//                     PythonLib/src/ravenpy.cc:75-91
// The 2-bit layout (A0 C1 G2 T3, 32 bases per u64, base i at bits
// [(i<<1)&63, +1] of word i>>5) is the GPU input format: `deflated_data` is
// uploaded as-is, never repacked (include/raven_b200.h: rvn_reads_upload).
#ifndef BIOSOUP_NUCLEIC_ACID_HPP_
#define BIOSOUP_NUCLEIC_ACID_HPP_

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace biosoup {

// IUPAC letter -> 2-bit code. ACGT(U) are the only letters pinned by the
// reference fixture; ambiguity codes collapse onto one of their members.
inline std::uint8_t NucleotideCode(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 'R': case 'r': return 0;  // A|G
    case 'Y': case 'y': return 3;  // C|T
    case 'K': case 'k': return 2;  // G|T
    case 'M': case 'm': return 1;  // A|C
    case 'S': case 's': return 1;  // C|G
    case 'W': case 'w': return 0;  // A|T
    case 'B': case 'b': return 1;
    case 'D': case 'd': return 0;
    case 'H': case 'h': return 3;
    case 'V': case 'v': return 2;
    case 'N': case 'n': return 0;
    case '-': return 0;
    default: return 255;
  }
}

class NucleicAcid {
 public:
  NucleicAcid() = default;

  NucleicAcid(const std::string& name, const std::string& data)
      : NucleicAcid(name.c_str(), name.size(), data.c_str(), data.size()) {}

  NucleicAcid(const char* name, std::uint32_t name_len, const char* data,
              std::uint32_t data_len)
      : id(num_objects++),
        name(name, name_len),
        deflated_data(),
        block_quality(),
        inflated_len(data_len),
        is_reverse_complement(0) {
    deflated_data.assign((static_cast<std::uint64_t>(data_len) + 31) >> 5, 0);
    for (std::uint32_t i = 0; i < data_len; ++i) {
      std::uint64_t c = NucleotideCode(data[i]);
      if (c == 255ULL) {
        throw std::invalid_argument(
            "[biosoup::NucleicAcid::NucleicAcid] error: not a nucleotide");
      }
      deflated_data[i >> 5] |= c << ((i << 1) & 63);
    }
  }

  NucleicAcid(const std::string& name, const std::string& data,
              const std::string& quality)
      : NucleicAcid(name.c_str(), name.size(), data.c_str(), data.size(),
                    quality.c_str(), quality.size()) {}

  NucleicAcid(const char* name, std::uint32_t name_len, const char* data,
              std::uint32_t data_len, const char* quality,
              std::uint32_t quality_len)
      : NucleicAcid(name, name_len, data, data_len) {
    // one byte per 64 bases: integer mean of Phred values
    block_quality.reserve((quality_len + 63) / 64);
    for (std::uint32_t i = 0; i < quality_len; i += 64) {
      std::uint32_t j = std::min(i + 64, quality_len);
      std::uint32_t sum = 0;
      for (std::uint32_t l = i; l < j; ++l) {
        sum += static_cast<std::uint32_t>(quality[l] - '!');
      }
      block_quality.emplace_back(sum / (j - i));
    }
  }

  NucleicAcid(const NucleicAcid&) = default;
  NucleicAcid& operator=(const NucleicAcid&) = default;
  NucleicAcid(NucleicAcid&&) = default;
  NucleicAcid& operator=(NucleicAcid&&) = default;
  ~NucleicAcid() = default;

  std::uint64_t Code(std::uint32_t i) const {
    std::uint64_t x = 0;
    if (is_reverse_complement) {
      i = inflated_len - i - 1;
      x = 3;
    }
    return ((deflated_data[i >> 5] >> ((i << 1) & 63)) & 3) ^ x;
  }

  std::uint8_t Score(std::uint32_t i) const {
    if (is_reverse_complement) {
      i = inflated_len - i - 1;
    }
    return block_quality[i >> 6];
  }

  std::string InflateData(std::uint32_t i = 0, std::uint32_t len = -1) const {
    if (i >= inflated_len) {
      return std::string{};
    }
    len = std::min(len, inflated_len - i);
    std::string dst;
    dst.reserve(len);
    for (; len; ++i, --len) {
      dst += "ACGT"[Code(i)];
    }
    return dst;
  }

  std::string InflateQuality(std::uint32_t i = 0,
                             std::uint32_t len = -1) const {
    if (block_quality.empty() || i >= inflated_len) {
      return std::string{};
    }
    len = std::min(len, inflated_len - i);
    std::string dst;
    dst.reserve(len);
    for (; len; ++i, --len) {
      dst += static_cast<char>(Score(i) + '!');
    }
    return dst;
  }

  void ReverseAndComplement() { is_reverse_complement ^= 1; }

  // defined by the application translation unit, as in the reference
  // (RavenExe/src/main.cc:12, RavenTest/src/raven_test.cpp:17)
  static std::atomic<std::uint32_t> num_objects;

  std::uint32_t id;
  std::string name;
  std::vector<std::uint64_t> deflated_data;
  std::vector<std::uint8_t> block_quality;
  std::uint32_t inflated_len;
  bool is_reverse_complement;
};

}  // namespace biosoup

#endif  // BIOSOUP_NUCLEIC_ACID_HPP_
