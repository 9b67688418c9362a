// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" surface of liboracle.so for tests/, smoke() and bench.py's CPU
// legs (ctypes). Nothing under raven_b200/ may link or load this library.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "flat_api.hpp"
#include "ram/minimizer_engine.hpp"
#include "stage1_port.hpp"

std::atomic<std::uint32_t> biosoup::NucleicAcid::num_objects{0};

struct orc_engine {
  std::shared_ptr<thread_pool::ThreadPool> pool;
  std::unique_ptr<ram::MinimizerEngine> eng;
};

ORC_BAG_ACCESSORS(orc)

ORC_EXPORT orc_engine* orc_engine_create(std::uint32_t k, std::uint32_t w,
                                         std::uint32_t bandwidth,
                                         std::uint32_t chain,
                                         std::uint32_t matches,
                                         std::uint32_t gap,
                                         std::uint32_t threads) {
  auto* e = new orc_engine();
  e->pool = std::make_shared<thread_pool::ThreadPool>(std::max(1U, threads));
  e->eng = std::make_unique<ram::MinimizerEngine>(e->pool, k, w, bandwidth,
                                                  chain, matches, gap);
  return e;
}

ORC_EXPORT void orc_engine_free(orc_engine* e) { delete e; }

// per-read sketches of reads [first, last): value/origin + offsets
ORC_EXPORT orc_bag* orc_sketch(orc_engine* e, orc_reads* r, std::uint32_t first,
                               std::uint32_t last, int minhash) {
  auto* bag = new orc_bag();
  std::vector<std::uint64_t> value, origin, off{0};
  for (std::uint32_t i = first; i < last; ++i) {
    for (const auto& m : e->eng->Sketch(r->seqs[i], minhash)) {
      value.emplace_back(m.value);
      origin.emplace_back(m.origin);
    }
    off.emplace_back(value.size());
  }
  bag->Put("value", value);
  bag->Put("origin", origin);
  bag->Put("offsets", off);
  return bag;
}

ORC_EXPORT void orc_engine_minimize(orc_engine* e, orc_reads* r,
                                    std::uint32_t first, std::uint32_t last,
                                    int minhash) {
  e->eng->Minimize(r->seqs.begin() + first, r->seqs.begin() + last, minhash);
}

// returns 0 and writes the occurrence threshold, or -1 on invalid frequency
ORC_EXPORT int orc_engine_filter(orc_engine* e, double f,
                                 std::uint32_t* occurrence) {
  try {
    e->eng->Filter(f);
  } catch (const std::invalid_argument&) {
    return -1;
  }
  *occurrence = e->eng->occurrence();
  return 0;
}

ORC_EXPORT orc_bag* orc_engine_keys(orc_engine* e) {
  auto* bag = new orc_bag();
  std::vector<std::uint64_t> values;
  std::vector<std::uint32_t> counts;
  e->eng->Keys(&values, &counts);
  bag->Put("values", values);
  bag->Put("counts", counts);
  std::vector<std::uint64_t> totals{e->eng->num_keys(),
                                    e->eng->num_minimizers()};
  bag->Put("totals", totals);
  return bag;
}

// Map every read in [first, last) against the current index
ORC_EXPORT orc_bag* orc_engine_map(orc_engine* e, orc_reads* r,
                                   std::uint32_t first, std::uint32_t last,
                                   int avoid_equal, int avoid_symmetric,
                                   int minhash, int want_matches) {
  auto* bag = new orc_bag();
  std::vector<std::uint32_t> ovl, filtered;
  std::vector<std::uint64_t> ovl_off{0}, filt_off{0};
  std::vector<std::uint64_t> mgroup, mpos, moff{0};
  for (std::uint32_t i = first; i < last; ++i) {
    std::vector<std::uint32_t> f;
    if (want_matches) {
      auto m = e->eng->Matches(r->seqs[i], avoid_equal, avoid_symmetric,
                               minhash, nullptr);
      for (const auto& it : m) {
        mgroup.emplace_back(it.group);
        mpos.emplace_back(it.positions);
      }
      moff.emplace_back(mgroup.size());
    }
    for (const auto& o : e->eng->Map(r->seqs[i], avoid_equal, avoid_symmetric,
                                     minhash, &f)) {
      PushOverlap(ovl, o);
    }
    ovl_off.emplace_back(ovl.size() / 8);
    filtered.insert(filtered.end(), f.begin(), f.end());
    filt_off.emplace_back(filtered.size());
  }
  bag->Put("overlaps", ovl);
  bag->Put("ovl_off", ovl_off);
  bag->Put("filtered", filtered);
  bag->Put("filt_off", filt_off);
  if (want_matches) {
    bag->Put("match_group", mgroup);
    bag->Put("match_pos", mpos);
    bag->Put("match_off", moff);
  }
  return bag;
}

// chain a caller-supplied hit list (for chain-kernel unit tests)
ORC_EXPORT orc_bag* orc_engine_chain(orc_engine* e, std::uint32_t lhs_id,
                                     const std::uint64_t* group,
                                     const std::uint64_t* positions,
                                     std::uint64_t n) {
  std::vector<ram::MinimizerEngine::Match> m;
  m.reserve(n + 1);
  for (std::uint64_t i = 0; i < n; ++i) {
    m.emplace_back(group[i], positions[i]);
  }
  auto* bag = new orc_bag();
  std::vector<std::uint32_t> ovl;
  for (const auto& o : e->eng->ChainMatches(lhs_id, std::move(m))) {
    PushOverlap(ovl, o);
  }
  bag->Put("overlaps", ovl);
  return bag;
}

static void PutStage1(orc_bag* bag, const oracle::Stage1Result& res,
                      double seconds) {
  std::vector<std::uint32_t> ovl;
  std::vector<std::uint64_t> ovl_off{0}, pile_off{0};
  std::vector<std::uint16_t> pile;
  for (std::size_t i = 0; i < res.overlaps.size(); ++i) {
    for (const auto& o : res.overlaps[i]) {
      PushOverlap(ovl, o);
    }
    ovl_off.emplace_back(ovl.size() / 8);
    pile.insert(pile.end(), res.piles[i].begin(), res.piles[i].end());
    pile_off.emplace_back(pile.size());
  }
  bag->Put("overlaps", ovl);
  bag->Put("ovl_off", ovl_off);
  bag->Put("pile", pile);
  bag->Put("pile_off", pile_off);
  bag->Put("occurrences", res.occurrences);
  std::vector<std::uint64_t> stats{res.num_mapped};
  bag->Put("num_mapped", stats);
  std::vector<double> t{seconds};
  bag->Put("seconds", t);
}

// stage 1 of the overlap phase (port of construct.cc:14-121)
ORC_EXPORT orc_bag* orc_stage1(orc_engine* e, orc_reads* r, double freq,
                               std::uint64_t max_overlaps, int minhash,
                               std::uint64_t index_batch_bases,
                               std::uint64_t query_batch_bases) {
  auto* bag = new orc_bag();
  auto t0 = std::chrono::steady_clock::now();
  auto res = oracle::FindOverlapsAndCreatePiles(
      e->pool, *e->eng, r->seqs, freq, max_overlaps, minhash,
      index_batch_bases, query_batch_bases);
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
                 .count();
  PutStage1(bag, res, s);
  return bag;
}

// Pile::AddLayers on one pile (pile.cc:33-62); overlaps as 8 x u32 records
ORC_EXPORT void orc_pile_add_layers(std::uint32_t id, std::uint16_t* data,
                                    std::uint32_t bins,
                                    const std::uint32_t* ovl, std::uint64_t n) {
  std::vector<biosoup::Overlap> v;
  for (std::uint64_t i = 0; i < n; ++i) {
    const std::uint32_t* o = ovl + 8 * i;
    v.emplace_back(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7] != 0);
  }
  std::vector<std::uint16_t> d(data, data + bins);
  oracle::AddLayers(id, d, v.data(), v.data() + v.size());
  std::copy(d.begin(), d.end(), data);
}

// the truncation step alone (construct.cc:98-107) on 8 x u32 records, in place
ORC_EXPORT std::uint64_t orc_truncate(std::uint32_t* ovl, std::uint64_t n,
                                      std::uint64_t max_overlaps) {
  if (n < max_overlaps) {
    return n;
  }
  struct Rec {
    std::uint32_t f[8];
  };
  Rec* first = reinterpret_cast<Rec*>(ovl);
  std::sort(first, first + n, [](const Rec& a, const Rec& b) {
    return std::max(a.f[5] - a.f[4], a.f[2] - a.f[1]) >
           std::max(b.f[5] - b.f[4], b.f[2] - b.f[1]);
  });
  return max_overlaps;
}

// the real std::sort on u64 elements compared by their high 32 bits,
// descending - the checker for raven_b200/csrc/introsort.cuh
ORC_EXPORT void orc_std_sort_hi32_desc(std::uint64_t* data, std::uint64_t n) {
  std::sort(data, data + n, [](std::uint64_t a, std::uint64_t b) {
    return (a >> 32) > (b >> 32);
  });
}

// Pile::AddKmers' low-complexity test (pile.cc:64-120), restated with plain
// strings: keep[i] = 1 iff the reference would set kmers_[pos >> 4]
ORC_EXPORT void orc_kmer_complexity(orc_reads* r, const std::uint32_t* read_index,
                                    const std::uint32_t* pos, std::uint64_t n,
                                    std::uint32_t k, std::uint8_t* keep) {
  auto unique_concat = [](const std::vector<std::string>& tok) {
    std::string out;
    for (std::size_t i = 0; i < tok.size(); ++i) {
      if (i == 0 || tok[i] != tok[i - 1]) out += tok[i];
    }
    return out;
  };
  for (std::uint64_t t = 0; t < n; ++t) {
    std::string kmer = r->seqs[read_index[t]]->InflateData(pos[t], k);
    keep[t] = 0;
    std::vector<std::string> tok;
    for (char c : kmer) tok.emplace_back(1, c);
    kmer = unique_concat(tok);
    if (kmer.size() < k / 2 + 1) continue;
    tok.clear();
    for (std::size_t i = 0; i < kmer.size(); ++i) {
      if (i % 2 == 1) tok.back() += kmer[i]; else tok.emplace_back(1, kmer[i]);
    }
    kmer = unique_concat(tok);
    if (kmer.size() < k / 2 + 1) continue;
    tok.clear();
    for (std::size_t i = 0; i < kmer.size(); ++i) {
      if (!tok.empty() && i % 2 == 0) tok.back() += kmer[i]; else tok.emplace_back(1, kmer[i]);
    }
    kmer = unique_concat(tok);
    if (kmer.size() < k / 2 + 1) continue;
    keep[t] = 1;
  }
}

ORC_EXPORT void orc_reads_set_ids(orc_reads* r, const std::uint32_t* ids) {
  for (std::size_t i = 0; i < r->seqs.size(); ++i) r->seqs[i]->id = ids[i];
}

// ---------------------------------------------------------------------------
// racon window consensus (POA) over a flat batch of windows; the same layout
// as rvn_poa_batch (include/raven_b200.h): window w owns sequences
// [win_first[w], win_first[w+1]), the first one is the backbone; sequence s is
// bases[seq_off[s] .. seq_off[s+1]) (+ quals at the same offsets, or NULL),
// layered at backbone positions [seq_begin[s], seq_end[s]].
// ---------------------------------------------------------------------------
#include "racon/window.hpp"

ORC_EXPORT orc_bag* orc_poa_batch(std::uint32_t n_windows,
                                  const std::uint32_t* win_first,
                                  const std::uint64_t* seq_off, const char* bases,
                                  const char* quals, const std::uint32_t* seq_begin,
                                  const std::uint32_t* seq_end, int m, int n, int g,
                                  int trim, int tgs, std::uint32_t threads) {
  std::vector<std::string> cons(n_windows);
  std::vector<std::vector<std::uint32_t>> covs(n_windows);
  std::vector<std::uint8_t> status(n_windows, 0);
  std::vector<std::uint64_t> cells(n_windows, 0);
  auto pool = std::make_shared<thread_pool::ThreadPool>(std::max(1U, threads));
  std::vector<std::future<void>> fut;
  auto t0 = std::chrono::steady_clock::now();
  for (std::uint32_t w = 0; w < n_windows; ++w) {
    fut.emplace_back(pool->Submit(
        [&](std::uint32_t w) {
          const std::uint32_t s0 = win_first[w], s1 = win_first[w + 1];
          racon::Window win(w, 0, tgs ? racon::WindowType::kTGS : racon::WindowType::kNGS,
                            bases + seq_off[s0], seq_off[s0 + 1] - seq_off[s0],
                            quals ? quals + seq_off[s0] : nullptr,
                            quals ? seq_off[s0 + 1] - seq_off[s0] : 0);
          for (std::uint32_t s = s0 + 1; s < s1; ++s) {
            const std::uint32_t len = seq_off[s + 1] - seq_off[s];
            win.AddLayer(bases + seq_off[s], len, quals ? quals + seq_off[s] : nullptr,
                         quals ? len : 0, seq_begin[s], seq_end[s]);
          }
          spoa::AlignmentEngine engine(m, n, g);
          status[w] = win.GenerateConsensus(&engine, trim != 0) ? 1 : 0;
          if (win.chimeric_warning()) status[w] |= 2;
          cons[w] = win.consensus();
          covs[w] = win.coverages();
          // unpolished windows (backbone returned) carry no coverage: zeros
          if (!(status[w] & 1)) covs[w].assign(cons[w].size(), 0);
          cells[w] = engine.cells();
        },
        w));
  }
  for (auto& f : fut) f.get();
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  auto* bag = new orc_bag();
  std::vector<char> flat;
  std::vector<std::uint64_t> off{0}, coff{0};
  std::vector<std::uint32_t> cov;
  for (std::uint32_t w = 0; w < n_windows; ++w) {
    flat.insert(flat.end(), cons[w].begin(), cons[w].end());
    off.emplace_back(flat.size());
    cov.insert(cov.end(), covs[w].begin(), covs[w].end());
    coff.emplace_back(cov.size());
  }
  bag->Put("consensus", flat);
  bag->Put("cons_off", off);
  bag->Put("coverage", cov);
  bag->Put("cov_off", coff);
  bag->Put("status", status);
  bag->Put("cells", cells);
  std::vector<double> t{s};
  bag->Put("seconds", t);
  return bag;
}

#include "racon/polisher.hpp"
ORC_EXPORT std::int64_t orc_nw_path(const char* q, int nq, const char* t, int nt, char* out) {
  const std::string p = racon::GlobalAlignmentPath(std::string(q, nq), std::string(t, nt));
  std::memcpy(out, p.data(), p.size());
  return static_cast<std::int64_t>(p.size());
}

// racon::Polisher::Create(...)->Polish(targets, sequences, false) of the ORACLE
// (oracle/racon_polisher.cpp): the checker of the product's polisher facade and
// the CPU leg of the polishing bench
ORC_EXPORT orc_bag* orc_polish(orc_reads* targets, orc_reads* sequences, double q, double e,
                               std::uint32_t w, int trim, int m, int n, int g,
                               std::uint32_t threads) {
  auto pool = std::make_shared<thread_pool::ThreadPool>(std::max(1U, threads));
  auto t0 = std::chrono::steady_clock::now();
  auto polisher = racon::Polisher::Create(pool, q, e, w, trim != 0, m, n, g);
  auto out = polisher->Polish(targets->seqs, sequences->seqs, false);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  auto* bag = new orc_bag();
  std::vector<char> seq, names;
  std::vector<std::uint64_t> off{0};
  for (const auto& s : out) {
    const std::string d = s->InflateData();
    seq.insert(seq.end(), d.begin(), d.end());
    off.emplace_back(seq.size());
    names.insert(names.end(), s->name.begin(), s->name.end());
    names.push_back('\n');
  }
  bag->Put("sequences", seq);
  bag->Put("seq_off", off);
  bag->Put("names", names);
  std::vector<double> st{static_cast<double>(polisher->num_windows()),
                         static_cast<double>(polisher->num_polished_windows()), secs};
  bag->Put("stats", st);
  return bag;
}

ORC_EXPORT void orc_reads_set_names(orc_reads* r, const char* prefix) {
  for (std::size_t i = 0; i < r->seqs.size(); ++i) {
    r->seqs[i]->name = std::string(prefix) + std::to_string(i);
  }
}

// AVX2 int16 matrix fill in the oracle's spoa (the CPU legs of the benches time
// the SIMD engine like upstream spoa's; the tests check both ways agree)
ORC_EXPORT void orc_spoa_use_simd(int on) { spoa::AlignmentEngine::UseSimd(on != 0); }
