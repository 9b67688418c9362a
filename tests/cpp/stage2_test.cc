// Drop-in check of the second overlap stage with the identity filter on: the
// reference's OWN functions (RavenLib construct.cc compiled in place: per-read
// Map calls through the ram::MinimizerEngine facade, one host edlibAlign per
// overlap) next to the batched B200 replacements with the same signatures
// (raven_b200::ResolveContainedReads, raven_b200::FindOverlapsAndRepetetiveRegions:
// one device map per batch, one batched edit-distance call). Both end states
// (every overlap list and every field of every pile) are dumped; they must be
// identical (tests/test_gpu_dropin.py).
//   usage: stage2_test <reads.bin> <out.bin> <k> <w> <freq> <identity>
#include <atomic>
#include <cstdint>
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

using Region = std::pair<std::uint32_t, std::uint32_t>;

struct PileDump {
  std::vector<std::uint32_t>* out;
  void operator()(std::uint32_t& id, std::uint32_t& b, std::uint32_t& e, std::uint16_t& med,
                  bool& inv, bool& cont, bool& chim, bool& rep, std::vector<std::uint16_t>& data,
                  std::vector<bool>& kmers, std::vector<Region>& cr, std::vector<Region>& rr) {
    out->insert(out->end(), {id, b, e, med, inv, cont, chim, rep,
                             static_cast<std::uint32_t>(data.size())});
    for (auto v : data) out->push_back(v);
    out->push_back(static_cast<std::uint32_t>(kmers.size()));
    for (std::size_t i = 0; i < kmers.size(); ++i) {
      if (kmers[i]) out->push_back(static_cast<std::uint32_t>(i));
    }
    out->push_back(0xFFFFFFFFu);
    for (const auto& v : {cr, rr}) {
      out->push_back(static_cast<std::uint32_t>(v.size()));
      for (const auto& r : v) {
        out->push_back(r.first);
        out->push_back(r.second);
      }
    }
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

}  // namespace

int main(int argc, char** argv) {
  if (argc < 7) {
    std::cerr << "usage: stage2_test reads.bin out.bin k w freq identity\n";
    return 2;
  }
  std::ifstream in(argv[1], std::ios::binary);
  const auto words = ReadVec<std::uint64_t>(in);
  const auto woff = ReadVec<std::uint64_t>(in);
  const auto lens = ReadVec<std::uint32_t>(in);
  const std::uint32_t k = std::stoul(argv[3]), w = std::stoul(argv[4]);
  const double freq = std::stod(argv[5]), identity = std::stod(argv[6]);
  auto pool = std::make_shared<thread_pool::ThreadPool>(4);
  std::ofstream out(argv[2], std::ios::binary);
  try {
    for (int mode = 0; mode < 2; ++mode) {
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
      ram::MinimizerEngine engine{pool, k, w};
      std::vector<std::unique_ptr<raven::Pile>> piles;
      std::vector<std::vector<biosoup::Overlap>> overlaps(seqs.size());
      raven_b200::FindOverlapsAndCreatePiles(pool, engine, seqs, freq, piles, overlaps, 32, false);
      if (mode == 0) {
        raven::TrimAndAnnotatePiles(pool, piles, overlaps);
      } else {
        raven_b200::TrimAndAnnotatePiles(pool, piles, overlaps, engine);
      }
      if (mode == 0) {
        raven::ResolveContainedReads(piles, overlaps, seqs, pool, identity);
      } else {
        raven_b200::ResolveContainedReads(piles, overlaps, seqs, pool, identity, engine);
      }
      raven::ResolveChimericSequences(pool, piles, overlaps, seqs);
      if (mode == 0) {
        raven::FindOverlapsAndRepetetiveRegions(pool, engine, freq, k, identity, piles, overlaps,
                                                seqs);
      } else {
        raven_b200::FindOverlapsAndRepetetiveRegions(pool, engine, freq, k, identity, piles,
                                                     overlaps, seqs);
      }
      std::vector<std::uint32_t> ovl, pd;
      std::vector<std::uint64_t> off{0};
      for (const auto& list : overlaps) {
        for (const auto& o : list) {
          ovl.insert(ovl.end(), {o.lhs_id, o.lhs_begin, o.lhs_end, o.rhs_id, o.rhs_begin,
                                 o.rhs_end, o.score, static_cast<std::uint32_t>(o.strand)});
        }
        off.push_back(ovl.size() / 8);
      }
      for (const auto& p : piles) {
        PileDump d{&pd};
        auto visit = cereal::fields(d);
        cereal::access::member_serialize(visit, *p);
      }
      std::vector<std::uint32_t> order;
      for (const auto& s : seqs) order.push_back(s->id);
      WriteVec(out, ovl);
      WriteVec(out, off);
      WriteVec(out, pd);
      WriteVec(out, order);
    }
  } catch (const std::exception& e) {
    std::cerr << "stage2_test: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
