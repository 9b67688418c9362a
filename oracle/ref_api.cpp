// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" surface of oracle/_ref/libraven_ref.so: the reference's OWN
// sources (RavenLib/src/{construct,pile,overlap_utils,graph}.cc), compiled in
// place from /root/reference by oracle/Makefile, over this repo's drop-in
// dependency headers (include/) and the oracle restatement of ram
// (oracle/ram/). It pins the in-tree half of the hot path (batch schedule,
// gather order, Pile::AddLayers, truncation sort) against the real code; the
// ram arithmetic underneath is still the restatement (ram is not in the tree).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "flat_api.hpp"
#include "ram/minimizer_engine.hpp"
#include "raven/graph/assemble.h"
#include "raven/graph/common.h"
#include "raven/graph/construct.h"
#include "raven/graph/polish.hpp"
#include "raven/graph/overlap_utils.h"
#include "raven/graph/serialization/binary.h"
#include "raven/pile.h"

std::atomic<std::uint32_t> biosoup::NucleicAcid::num_objects{0};

namespace raven {
// construct.cc only reaches these with checkpoints = true, which the pinning
// entry points never pass
void StoreGraphToFile(const Graph&) {
  throw std::logic_error("checkpoints are not part of the pinned path");
}
}  // namespace raven

namespace {

// Pile keeps its histogram private; the only door the reference leaves open is
// `friend cereal::access` + serialize(). This archive just captures fields.
struct KmerProbe {
  bool any = false;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t&, std::uint32_t&, std::uint16_t&,
                  bool&, bool&, bool&, bool&, std::vector<std::uint16_t>&,
                  std::vector<bool>& kmers, Ts&...) {
    for (bool b : kmers) any = any || b;
  }
};

struct PileProbe {
  std::vector<std::uint16_t> data;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t&, std::uint32_t&,
                  std::uint16_t&, bool&, bool&, bool&, bool&,
                  std::vector<std::uint16_t>& d, Ts&...) {
    data = d;
  }
};

struct PileDataSetter {  // a given histogram into a fresh pile
  const std::uint16_t* src;
  std::size_t n;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t&, std::uint32_t&,
                  std::uint16_t&, bool&, bool&, bool&, bool&,
                  std::vector<std::uint16_t>& d, Ts&...) {
    d.assign(src, src + n);
  }
};

struct PileFieldsProbe {
  std::uint32_t begin = 0, end = 0;
  std::uint16_t median = 0;
  bool invalid = false;
  std::vector<std::uint16_t> data;
  template <typename... Ts>
  void operator()(std::uint32_t&, std::uint32_t& b, std::uint32_t& e,
                  std::uint16_t& m, bool& inv, bool&, bool&, bool&,
                  std::vector<std::uint16_t>& d, Ts&...) {
    begin = b;
    end = e;
    median = m;
    invalid = inv;
    data = d;
  }
};

}  // namespace

ORC_BAG_ACCESSORS(ref)

ORC_EXPORT orc_bag* ref_stage1(orc_reads* r, std::uint32_t k, std::uint32_t w,
                               double freq, std::uint64_t max_overlaps,
                               int minhash, std::uint32_t threads) {
  auto pool = std::make_shared<thread_pool::ThreadPool>(threads ? threads : 1);
  ram::MinimizerEngine engine{pool, k, w};
  std::vector<std::unique_ptr<raven::Pile>> piles;
  std::vector<std::vector<biosoup::Overlap>> overlaps(r->seqs.size());

  auto t0 = std::chrono::steady_clock::now();
  raven::FindOverlapsAndCreatePiles(pool, engine, r->seqs, freq, piles,
                                    overlaps, max_overlaps, minhash);
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
                 .count();

  auto* bag = new orc_bag();
  std::vector<std::uint32_t> ovl;
  std::vector<std::uint64_t> ovl_off{0}, pile_off{0};
  std::vector<std::uint16_t> pile;
  for (std::size_t i = 0; i < overlaps.size(); ++i) {
    for (const auto& o : overlaps[i]) {
      PushOverlap(ovl, o);
    }
    ovl_off.emplace_back(ovl.size() / 8);
    PileProbe probe;
    auto visit = cereal::fields(probe);
    cereal::access::member_serialize(visit, *piles[i]);
    pile.insert(pile.end(), probe.data.begin(), probe.data.end());
    pile_off.emplace_back(pile.size());
  }
  bag->Put("overlaps", ovl);
  bag->Put("ovl_off", ovl_off);
  bag->Put("pile", pile);
  bag->Put("pile_off", pile_off);
  std::vector<std::uint32_t> occ{engine.occurrence()};
  bag->Put("occurrences", occ);
  std::vector<double> t{s};
  bag->Put("seconds", t);
  return bag;
}

// raven::Pile::AddLayers on a fresh pile of `len` bases
ORC_EXPORT orc_bag* ref_pile_add_layers(std::uint32_t id, std::uint32_t len,
                                        const std::uint32_t* ovl,
                                        std::uint64_t n, std::uint32_t rounds) {
  std::vector<biosoup::Overlap> v;
  for (std::uint64_t i = 0; i < n; ++i) {
    const std::uint32_t* o = ovl + 8 * i;
    v.emplace_back(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7] != 0);
  }
  raven::Pile p(id, len);
  for (std::uint32_t i = 0; i < rounds; ++i) {
    p.AddLayers(v.begin(), v.end());
  }
  PileProbe probe;
  auto visit = cereal::fields(probe);
  cereal::access::member_serialize(visit, p);
  auto* bag = new orc_bag();
  bag->Put("pile", probe.data);
  return bag;
}

// raven::Pile::FindValidRegion(coverage) + FindMedian (TrimAndAnnotatePiles,
// construct.cc:134-139) on piles with the given histograms: fields afterwards,
// and the histograms as UpdateValidRegion left them
ORC_EXPORT void ref_pile_trim(const std::uint16_t* data, const std::uint64_t* off,
                              std::uint32_t n, std::uint32_t coverage, std::uint32_t* begin,
                              std::uint32_t* end, std::uint16_t* median,
                              std::uint8_t* invalid, std::uint16_t* data_out) {
  for (std::uint32_t i = 0; i < n; ++i) {
    const std::size_t bins = off[i + 1] - off[i];
    raven::Pile p(i, static_cast<std::uint32_t>(bins << 4));
    PileDataSetter set{data + off[i], bins};
    auto put = cereal::fields(set);
    cereal::access::member_serialize(put, p);
    p.FindValidRegion(static_cast<std::uint16_t>(coverage));
    if (!p.is_invalid()) p.FindMedian();
    PileFieldsProbe probe;
    auto get = cereal::fields(probe);
    cereal::access::member_serialize(get, p);
    begin[i] = probe.begin;
    end[i] = probe.end;
    median[i] = probe.median;
    invalid[i] = probe.invalid ? 1 : 0;
    std::copy(probe.data.begin(), probe.data.end(), data_out + off[i]);
  }
}

// raven::Pile::AddKmers on a fresh pile of read `index`: which positions mark
// their bin (one position at a time, so keep[] is per position)
ORC_EXPORT void ref_kmer_complexity(orc_reads* r, const std::uint32_t* read_index,
                                    const std::uint32_t* pos, std::uint64_t n,
                                    std::uint32_t k, std::uint8_t* keep) {
  for (std::uint64_t t = 0; t < n; ++t) {
    const auto& seq = r->seqs[read_index[t]];
    raven::Pile p(seq->id, seq->inflated_len);
    p.AddKmers(std::vector<std::uint32_t>{pos[t]}, k, seq);
    KmerProbe probe;
    auto visit = cereal::fields(probe);
    cereal::access::member_serialize(visit, p);
    keep[t] = probe.any;
  }
}

// The reference's only golden value: RavenTest.Assemble
// (RavenTest/src/raven_test.cpp:50-67) = ConstructGraph(useMinhash) ->
// Assemble -> Polish(2 rounds, m3 n-5 g-4) on the lambda reads, then the edit
// distance of the reverse-complemented first unitig to NC_001416 (1137).
// The reference's own construct/assemble/polish/common sources run here over
// the oracle restatements of ram / racon / spoa / edlib.
ORC_EXPORT orc_bag* ref_assemble(orc_reads* r, int minhash, std::uint32_t rounds,
                                 std::uint32_t threads) {
  biosoup::NucleicAcid::num_objects = r->seqs.size();
  raven::Graph graph;
  auto pool = std::make_shared<thread_pool::ThreadPool>(threads ? threads : 1);
  raven::OverlapPhaseCfg cfg{};
  cfg.useMinhash = minhash != 0;
  raven::ConstructGraph(graph, r->seqs, pool, false, cfg);
  raven::Assemble(pool, graph, false);
  raven::PolishCfg pcfg{};
  pcfg.num_rounds = rounds;
  raven::Polish(pool, graph, false, r->seqs, pcfg);
  auto unitigs = raven::GetUnitigs(graph);
  auto* bag = new orc_bag();
  std::vector<char> flat;
  std::vector<std::uint64_t> off{0};
  std::vector<char> names;
  for (const auto& u : unitigs) {
    auto s = u->InflateData();
    flat.insert(flat.end(), s.begin(), s.end());
    off.emplace_back(flat.size());
    names.insert(names.end(), u->name.begin(), u->name.end());
    names.push_back('\n');
  }
  bag->Put("unitigs", flat);
  bag->Put("unitig_off", off);
  bag->Put("names", names);
  return bag;
}

ORC_EXPORT std::uint32_t ref_overlap_length(const std::uint32_t* o) {
  return raven::GetOverlapLength(
      biosoup::Overlap(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7] != 0));
}
