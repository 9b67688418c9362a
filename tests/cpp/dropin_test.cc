// Drop-in check: the reference's OWN stage-1 function (RavenLib construct.cc,
// compiled in place) running on the B200 engine through our
// ram::MinimizerEngine facade, next to our batched replacement with the same
// signature. Both results are dumped for tests/test_gpu_dropin.py, which
// compares them with the CPU oracle.
//   usage: dropin_test <reads.bin> <out.bin> <k> <w> <freq> <kmax> <minhash>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "raven/graph/construct.h"
#include "raven/graph/serialization/binary.h"
#include "raven/pile.h"
#include "raven_b200/construct_b200.hpp"

std::atomic<std::uint32_t> biosoup::NucleicAcid::num_objects{0};

namespace raven {
void StoreGraphToFile(const Graph&) { throw std::logic_error("no checkpoints here"); }
}  // namespace raven

namespace {

struct PileProbe {
  std::vector<std::uint16_t> data;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t&, std::uint32_t&, std::uint16_t&,
                  bool&, bool&, bool&, bool&, std::vector<std::uint16_t>& d, Ts&...) {
    data = d;
  }
};

template <typename T>
std::vector<T> ReadVec(std::ifstream& f) {
  std::uint64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 8);
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}

template <typename T>
void WriteVec(std::ofstream& f, const std::vector<T>& v) {
  std::uint64_t n = v.size();
  f.write(reinterpret_cast<const char*>(&n), 8);
  f.write(reinterpret_cast<const char*>(v.data()), n * sizeof(T));
}

void Dump(std::ofstream& out,
          const std::vector<std::vector<biosoup::Overlap>>& overlaps,
          const std::vector<std::unique_ptr<raven::Pile>>& piles) {
  std::vector<std::uint32_t> ovl;
  std::vector<std::uint64_t> off{0}, poff{0};
  std::vector<std::uint16_t> pile;
  for (std::size_t i = 0; i < overlaps.size(); ++i) {
    for (const auto& o : overlaps[i]) {
      ovl.insert(ovl.end(), {o.lhs_id, o.lhs_begin, o.lhs_end, o.rhs_id, o.rhs_begin,
                             o.rhs_end, o.score, static_cast<std::uint32_t>(o.strand)});
    }
    off.push_back(ovl.size() / 8);
    PileProbe probe;
    auto visit = cereal::fields(probe);
    cereal::access::member_serialize(visit, *piles[i]);
    pile.insert(pile.end(), probe.data.begin(), probe.data.end());
    poff.push_back(pile.size());
  }
  WriteVec(out, ovl);
  WriteVec(out, off);
  WriteVec(out, pile);
  WriteVec(out, poff);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 8) {
    std::cerr << "usage: dropin_test reads.bin out.bin k w freq kmax minhash\n";
    return 2;
  }
  std::ifstream in(argv[1], std::ios::binary);
  auto words = ReadVec<std::uint64_t>(in);
  auto woff = ReadVec<std::uint64_t>(in);
  auto lens = ReadVec<std::uint32_t>(in);
  const std::uint32_t k = std::stoul(argv[3]), w = std::stoul(argv[4]);
  const double freq = std::stod(argv[5]);
  const std::size_t kmax = std::stoull(argv[6]);
  const bool minhash = std::stoi(argv[7]) != 0;

  std::vector<std::unique_ptr<biosoup::NucleicAcid>> seqs;
  for (std::size_t i = 0; i < lens.size(); ++i) {
    auto s = std::make_unique<biosoup::NucleicAcid>();
    s->id = i;
    s->name = std::to_string(i);
    s->deflated_data.assign(words.begin() + woff[i], words.begin() + woff[i + 1]);
    s->inflated_len = lens[i];
    s->is_reverse_complement = false;
    seqs.emplace_back(std::move(s));
  }
  auto pool = std::make_shared<thread_pool::ThreadPool>(4);
  std::ofstream out(argv[2], std::ios::binary);
  try {
    {  // A: the reference's construct.cc over the facade (per-read Map calls)
      ram::MinimizerEngine engine{pool, k, w};
      std::vector<std::unique_ptr<raven::Pile>> piles;
      std::vector<std::vector<biosoup::Overlap>> overlaps(seqs.size());
      raven::FindOverlapsAndCreatePiles(pool, engine, seqs, freq, piles, overlaps,
                                        kmax, minhash);
      Dump(out, overlaps, piles);
      // error behaviour of the facade mirrors the reference engine
      bool threw = false;
      try {
        engine.Filter(1.5);
      } catch (const std::invalid_argument&) {
        threw = true;
      }
      if (!threw) throw std::logic_error("Filter(1.5) did not throw invalid_argument");
    }
    {  // B: our batched replacement, same signature
      ram::MinimizerEngine engine{pool, k, w};
      std::vector<std::unique_ptr<raven::Pile>> piles;
      std::vector<std::vector<biosoup::Overlap>> overlaps(seqs.size());
      raven_b200::FindOverlapsAndCreatePiles(pool, engine, seqs, freq, piles,
                                             overlaps, kmax, minhash);
      Dump(out, overlaps, piles);
    }
  } catch (const std::exception& e) {
    std::cerr << "dropin_test: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
