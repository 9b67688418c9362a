// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/racon/polisher.hpp).
#include "racon/polisher.hpp"

#include <algorithm>
#include <future>
#include <cstdlib>
#include <stdexcept>

#include "biosoup/overlap.hpp"
#include "racon/window.hpp"
#include "ram/minimizer_engine.hpp"
#include "spoa/spoa.hpp"

namespace racon {

Polisher::Polisher(std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q,
                   double e, std::uint32_t w, bool trim, std::int8_t m, std::int8_t n,
                   std::int8_t g)
    : thread_pool_(thread_pool ? thread_pool
                               : std::make_shared<thread_pool::ThreadPool>(1)),
      q_(q), e_(e), w_(w), trim_(trim), m_(m), n_(n), g_(g) {}

std::unique_ptr<Polisher> Polisher::Create(
    std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q, double e,
    std::uint32_t w, bool trim, std::int8_t m, std::int8_t n, std::int8_t g,
    std::uint32_t cuda_poa_batches, bool cuda_banded_alignment,
    std::uint32_t cuda_alignment_batches) {
  if (w == 0) {
    throw std::invalid_argument("[racon::Polisher::Create] error: invalid window length");
  }
  if (g >= 0) {
    throw std::invalid_argument(
        "[racon::Polisher::Create] error: gap penalty must be negative");
  }
  if (cuda_poa_batches > 0 || cuda_alignment_batches > 0) {
    // the CPU build of the reference's dependency has no CUDA path
    throw std::logic_error("[racon::Polisher::Create] error: CUDA support is not available");
  }
  (void)cuda_banded_alignment;
  return std::unique_ptr<Polisher>(new Polisher(thread_pool, q, e, w, trim, m, n, g));
}

namespace {

std::uint32_t Span(const biosoup::Overlap& o) {
  return std::max(o.lhs_end - o.lhs_begin, o.rhs_end - o.rhs_begin);
}

}  // namespace

std::vector<std::unique_ptr<biosoup::NucleicAcid>> Polisher::Polish(
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& targets,
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences,
    bool drop_unpolished) {
  num_windows_ = num_polished_windows_ = 0;
  if (targets.empty() || sequences.empty()) {
    return {};
  }

  // ---- 1. map every read to the targets, keep its longest overlap ----
  // ids inside this routine are positions in `targets` / `sequences`
  std::vector<std::unique_ptr<biosoup::NucleicAcid>> tcopy, scopy;
  std::uint64_t total_len = 0;
  const std::uint32_t saved_counter = biosoup::NucleicAcid::num_objects;
  for (std::uint32_t i = 0; i < targets.size(); ++i) {
    tcopy.emplace_back(new biosoup::NucleicAcid(*targets[i]));
    tcopy.back()->id = i;
  }
  for (std::uint32_t i = 0; i < sequences.size(); ++i) {
    scopy.emplace_back(new biosoup::NucleicAcid(*sequences[i]));
    scopy.back()->id = i;
    total_len += sequences[i]->inflated_len;
  }
  const WindowType window_type =
      static_cast<double>(total_len) / sequences.size() <= 1000 ? WindowType::kNGS
                                                                 : WindowType::kTGS;

  std::vector<biosoup::Overlap> best(scopy.size());
  std::vector<bool> has(scopy.size(), false);
  ram::MinimizerEngine engine{thread_pool_, 15, 5};
  std::uint64_t bytes = 0;
  for (std::uint32_t i = 0, j = 0; i < tcopy.size(); ++i) {
    bytes += tcopy[i]->inflated_len;
    if (i != tcopy.size() - 1 && bytes < (1ULL << 32)) {
      continue;
    }
    bytes = 0;
    engine.Minimize(tcopy.begin() + j, tcopy.begin() + i + 1);
    engine.Filter(0.001);
    std::vector<std::future<std::vector<biosoup::Overlap>>> futures;
    for (std::uint32_t k = 0; k < scopy.size(); ++k) {
      futures.emplace_back(thread_pool_->Submit(
          [&](std::uint32_t k) { return engine.Map(scopy[k], false, false); }, k));
    }
    for (std::uint32_t k = 0; k < scopy.size(); ++k) {
      static const bool by_score = std::getenv("ORC_BEST_SCORE") != nullptr;
      for (const auto& o : futures[k].get()) {
        if (!has[k] || (by_score ? best[k].score < o.score : Span(best[k]) < Span(o))) {
          best[k] = o;
          has[k] = true;
        }
      }
    }
    j = i + 1;
  }

  // ---- 2. align, cut into breaking points ----
  struct Aligned {
    std::uint32_t q, t;
    bool strand;
    std::vector<std::pair<std::uint32_t, std::uint32_t>> bp;
  };
  std::vector<std::future<Aligned>> afut;
  for (std::uint32_t k = 0; k < scopy.size(); ++k) {
    if (!has[k]) continue;
    const auto& o = best[k];
    const double ql = o.lhs_end - o.lhs_begin, tl = o.rhs_end - o.rhs_begin;
    if (1 - std::min(ql, tl) / std::max(ql, tl) > e_) continue;
    afut.emplace_back(thread_pool_->Submit(
        [&](biosoup::Overlap o) {
          Aligned a{o.lhs_id, o.rhs_id, o.strand, {}};
          const auto& seq = scopy[o.lhs_id];
          std::uint32_t qb = o.lhs_begin, qe = o.lhs_end;
          biosoup::NucleicAcid view(*seq);
          if (!o.strand) {  // work on the reverse complement of the read
            view.ReverseAndComplement();
            qb = seq->inflated_len - o.lhs_end;
            qe = seq->inflated_len - o.lhs_begin;
          }
          const std::string q = view.InflateData(qb, qe - qb);
          const std::string t =
              tcopy[o.rhs_id]->InflateData(o.rhs_begin, o.rhs_end - o.rhs_begin);
          const std::string path = GlobalAlignmentPath(q, t);
          a.bp = BreakingPoints(path, qb, o.rhs_begin, o.rhs_end, w_);
          return a;
        },
        o));
  }
  std::vector<Aligned> aligned;
  for (auto& f : afut) aligned.emplace_back(f.get());

  // ---- 3. windows ----
  std::vector<std::string> backbones, dummy_q;
  std::vector<std::unique_ptr<Window>> windows;
  std::vector<std::uint64_t> first_window(tcopy.size() + 1, 0);
  std::vector<std::string> layer_data, layer_qual;  // payloads outlive the windows
  {
    std::uint64_t nw = 0;
    for (const auto& t : tcopy) nw += (t->inflated_len + w_ - 1) / w_;
    backbones.reserve(nw);
    dummy_q.reserve(nw);
  }
  for (std::uint32_t i = 0; i < tcopy.size(); ++i) {
    std::uint32_t k = 0;
    for (std::uint32_t j = 0; j < tcopy[i]->inflated_len; j += w_, ++k) {
      const std::uint32_t len = std::min(j + w_, tcopy[i]->inflated_len) - j;
      backbones.emplace_back(tcopy[i]->InflateData(j, len));
      dummy_q.emplace_back(std::string(len, '!'));
      windows.emplace_back(new Window(i, k, window_type, backbones.back().c_str(), len,
                                      dummy_q.back().c_str(), len));
    }
    first_window[i + 1] = first_window[i] + k;
  }
  std::vector<std::uint32_t> coverage(tcopy.size(), 0);
  std::uint64_t n_layers = 0;
  for (const auto& a : aligned) n_layers += a.bp.size() / 2;
  layer_data.reserve(n_layers);
  layer_qual.reserve(n_layers);
  for (const auto& a : aligned) {
    ++coverage[a.t];
    biosoup::NucleicAcid view(*scopy[a.q]);
    if (!a.strand) view.ReverseAndComplement();
    for (std::size_t j = 0; j + 1 < a.bp.size(); j += 2) {
      const std::uint32_t qb = a.bp[j].second, qe = a.bp[j + 1].second;
      if (qe - qb < 0.02 * w_) continue;
      if (!view.block_quality.empty()) {
        double avg = 0;
        for (std::uint32_t k = qb; k < qe; ++k) avg += view.Score(k);
        avg /= qe - qb;
        if (avg < q_) continue;
      }
      const std::uint64_t wid = first_window[a.t] + a.bp[j].first / w_;
      const std::uint32_t wstart = (a.bp[j].first / w_) * w_;
      layer_data.emplace_back(view.InflateData(qb, qe - qb));
      const char* qual = nullptr;
      std::uint32_t qual_len = 0;
      if (!view.block_quality.empty()) {
        layer_qual.emplace_back(view.InflateQuality(qb, qe - qb));
        qual = layer_qual.back().c_str();
        qual_len = layer_qual.back().size();
      }
      windows[wid]->AddLayer(layer_data.back().c_str(), layer_data.back().size(), qual,
                             qual_len, a.bp[j].first - wstart,
                             a.bp[j + 1].first - wstart - 1);
    }
  }

  // ---- 4. consensus per window ----
  std::vector<std::future<bool>> cfut;
  for (std::uint64_t i = 0; i < windows.size(); ++i) {
    cfut.emplace_back(thread_pool_->Submit(
        [&](std::uint64_t i) {
          spoa::AlignmentEngine engine(m_, n_, g_);
          return windows[i]->GenerateConsensus(&engine, trim_);
        },
        i));
  }

  // ---- 5. stitch ----
  biosoup::NucleicAcid::num_objects = saved_counter;  // the copies were scratch
  std::vector<std::unique_ptr<biosoup::NucleicAcid>> dst;
  std::string polished;
  std::uint32_t n_polished = 0;
  for (std::uint64_t i = 0; i < windows.size(); ++i) {
    n_polished += cfut[i].get() ? 1 : 0;
    polished += windows[i]->consensus();
    if (i == windows.size() - 1 || windows[i + 1]->rank() == 0) {
      const double ratio = n_polished / static_cast<double>(windows[i]->rank() + 1);
      num_windows_ += windows[i]->rank() + 1;
      num_polished_windows_ += n_polished;
      if (!drop_unpolished || ratio > 0) {
        std::string tags = " LN:i:" + std::to_string(polished.size());
        tags += " RC:i:" + std::to_string(coverage[windows[i]->id()]);
        tags += " XC:f:" + std::to_string(ratio);
        dst.emplace_back(new biosoup::NucleicAcid(
            targets[windows[i]->id()]->name + tags, polished));
      }
      n_polished = 0;
      polished.clear();
    }
  }
  return dst;
}

}  // namespace racon
