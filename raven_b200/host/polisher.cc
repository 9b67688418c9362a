// raven_b200 — racon::Polisher facade (include/racon/polisher.hpp).
// Reference contract: RavenLib/src/polish.cc:43-51; behaviour of the stages:
// SURVEY.md App. A.4. GPU: mapping (rvn_minimize/rvn_filter/rvn_map) and the
// per-window consensus (rvn_poa_batch). Host pool: best overlap per read,
// global alignment path (edlibAlign, EDLIB_TASK_PATH), breaking points, window
// assembly, stitching.
#include "racon/polisher.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <future>
#include <stdexcept>
#include <string>

#include "edlib.h"
#include "raven_b200.h"

namespace racon {

namespace {

void Check(rvn_ctx* ctx, int rc) {
  if (rc == RVN_OK) return;
  std::string msg = std::string("[racon::Polisher::Polish] error: ") + rvn_last_error(ctx);
  if (rc == RVN_ERR_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

struct Hit {  // the overlap kept for one read
  bool valid = false;
  rvn_overlap o{};
};

struct Piece {  // one layer: read segment [qb, qe) laid over a window
  std::uint64_t window;
  std::uint32_t read, qb, qe, begin, end;
  bool strand;
};

}  // namespace

Polisher::Polisher(std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q,
                   double e, std::uint32_t w, bool trim, std::int8_t m, std::int8_t n,
                   std::int8_t g)
    : thread_pool_(thread_pool ? thread_pool
                               : std::make_shared<thread_pool::ThreadPool>(1)),
      q_(q), e_(e), w_(w), trim_(trim), m_(m), n_(n), g_(g), ctx_(nullptr) {
  if (rvn_ctx_create(0, nullptr, &ctx_) != RVN_OK) {
    throw std::runtime_error(
        "[racon::Polisher::Create] error: no usable CUDA device (no CPU fallback)");
  }
}

Polisher::~Polisher() {
  if (ctx_) rvn_ctx_destroy(ctx_);
}

std::unique_ptr<Polisher> Polisher::Create(
    std::shared_ptr<thread_pool::ThreadPool> thread_pool, double q, double e,
    std::uint32_t w, bool trim, std::int8_t m, std::int8_t n, std::int8_t g,
    std::uint32_t /*cuda_poa_batches*/, bool /*cuda_banded_alignment*/,
    std::uint32_t /*cuda_alignment_batches*/) {
  if (w == 0) {
    throw std::invalid_argument("[racon::Polisher::Create] error: invalid window length");
  }
  if (g >= 0) {
    throw std::invalid_argument(
        "[racon::Polisher::Create] error: gap penalty must be negative");
  }
  return std::unique_ptr<Polisher>(new Polisher(thread_pool, q, e, w, trim, m, n, g));
}

std::vector<std::unique_ptr<biosoup::NucleicAcid>> Polisher::Polish(
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& targets,
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences,
    bool drop_unpolished) {
  num_windows_ = num_polished_windows_ = 0;
  poa_seconds_ = 0;
  if (targets.empty() || sequences.empty()) {
    return {};
  }
  const std::uint32_t T = targets.size(), S = sequences.size();

  // ---- 1. GPU: index the targets, map every read (ram k=15 w=5 f=0.001) ----
  // device set = targets (ids 0..T-1) followed by reads (ids T..T+S-1)
  std::vector<std::uint64_t> words, off{0};
  std::vector<std::uint32_t> lens;
  std::uint64_t total_len = 0;
  auto append = [&](const biosoup::NucleicAcid& s) {
    if (s.is_reverse_complement) {
      std::vector<std::uint64_t> w((static_cast<std::uint64_t>(s.inflated_len) + 31) >> 5, 0);
      for (std::uint32_t i = 0; i < s.inflated_len; ++i) w[i >> 5] |= s.Code(i) << ((i << 1) & 63);
      words.insert(words.end(), w.begin(), w.end());
    } else {
      words.insert(words.end(), s.deflated_data.begin(), s.deflated_data.end());
    }
    off.push_back(words.size());
    lens.push_back(s.inflated_len);
  };
  for (const auto& t : targets) append(*t);
  for (const auto& s : sequences) {
    append(*s);
    total_len += s->inflated_len;
  }
  const bool tgs = static_cast<double>(total_len) / S > 1000;
  Check(ctx_, rvn_engine_configure(ctx_, 15, 5, 500, 4, 100, 10000));
  Check(ctx_, rvn_reads_upload(ctx_, words.data(), off.data(), lens.data(), T + S));

  std::vector<Hit> best(S);
  std::uint64_t bytes = 0;
  for (std::uint32_t i = 0, j = 0; i < T; ++i) {
    bytes += targets[i]->inflated_len;
    if (i != T - 1 && bytes < (1ULL << 32)) continue;
    bytes = 0;
    Check(ctx_, rvn_minimize(ctx_, j, i + 1, 0));
    std::uint32_t occ = 0;
    Check(ctx_, rvn_filter(ctx_, 0.001, &occ));
    Check(ctx_, rvn_map(ctx_, T, T + S, 0, 0, 0, 0));
    const rvn_overlap* o = nullptr;
    const std::uint64_t* ooff = nullptr;
    std::uint64_t n = 0;
    Check(ctx_, rvn_map_results(ctx_, &o, &ooff, &n, nullptr, nullptr));
    for (std::uint32_t k = 0; k < S; ++k) {
      for (std::uint64_t x = ooff[k]; x < ooff[k + 1]; ++x) {
        const auto span = [](const rvn_overlap& v) {
          return std::max(v.lhs_end - v.lhs_begin, v.rhs_end - v.rhs_begin);
        };
        if (!best[k].valid || span(best[k].o) < span(o[x])) {
          best[k].o = o[x];
          best[k].valid = true;
        }
      }
    }
    j = i + 1;
  }

  // ---- 2. host pool: alignment path -> breaking points -> layer pieces ----
  std::vector<std::uint64_t> first_window(T + 1ULL, 0);
  for (std::uint32_t i = 0; i < T; ++i) {
    first_window[i + 1] = first_window[i] + (targets[i]->inflated_len + w_ - 1) / w_;
  }
  std::vector<std::future<std::vector<Piece>>> futures;
  std::vector<std::uint32_t> coverage(T, 0);
  for (std::uint32_t k = 0; k < S; ++k) {
    if (!best[k].valid) continue;
    const rvn_overlap o = best[k].o;
    const double ql = o.lhs_end - o.lhs_begin, tl = o.rhs_end - o.rhs_begin;
    if (1 - std::min(ql, tl) / std::max(ql, tl) > e_) continue;
    ++coverage[o.rhs_id];
    futures.emplace_back(thread_pool_->Submit(
        [&](std::uint32_t k, rvn_overlap o) {
          std::vector<Piece> pieces;
          const auto& seq = *sequences[k];
          biosoup::NucleicAcid view(seq);
          std::uint32_t qb = o.lhs_begin, qe = o.lhs_end;
          if (!o.strand) {  // the reverse complement of the read is aligned
            view.ReverseAndComplement();
            qb = seq.inflated_len - o.lhs_end;
            qe = seq.inflated_len - o.lhs_begin;
          }
          const std::string q = view.InflateData(qb, qe - qb);
          const std::string t =
              targets[o.rhs_id]->InflateData(o.rhs_begin, o.rhs_end - o.rhs_begin);
          EdlibAlignResult r = edlibAlign(
              q.c_str(), q.size(), t.c_str(), t.size(),
              edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, nullptr, 0));
          if (r.status != EDLIB_STATUS_OK) {
            edlibFreeAlignResult(r);
            return pieces;
          }
          // breaking points: first / one-past-last aligned (target, read) pair
          // of every window the overlap touches
          std::int64_t next_end;  // last target position of the current window
          {
            const std::uint32_t tb = o.rhs_begin;
            next_end = std::min<std::int64_t>(
                static_cast<std::int64_t>(tb / w_ + 1) * w_ - 1,
                static_cast<std::int64_t>(o.rhs_end) - 1);
          }
          bool found = false;
          std::uint32_t ft = 0, fq = 0, lt = 0, lq = 0;
          std::int64_t qp = static_cast<std::int64_t>(qb) - 1;
          std::int64_t tp = static_cast<std::int64_t>(o.rhs_begin) - 1;
          auto close = [&]() {
            if (tp != next_end) return;
            if (found && lq - fq >= 0.02 * w_) {
              bool keep = true;
              if (!view.block_quality.empty()) {
                double avg = 0;
                for (std::uint32_t x = fq; x < lq; ++x) avg += view.Score(x);
                avg /= lq - fq;
                keep = !(avg < q_);
              }
              const std::uint32_t ws0 = (ft / w_) * w_;
              if (ft - ws0 >= lt - ws0 - 1) keep = false;  // racon skips begin == end
              if (keep) {
                const std::uint32_t ws = (ft / w_) * w_;
                pieces.push_back(Piece{first_window[o.rhs_id] + ft / w_, k, fq, lq,
                                       ft - ws, lt - ws - 1, o.strand != 0});
              }
            }
            found = false;
            next_end = std::min<std::int64_t>(next_end + w_,
                                              static_cast<std::int64_t>(o.rhs_end) - 1);
          };
          for (int a = 0; a < r.alignmentLength; ++a) {
            const unsigned char op = r.alignment[a];
            if (op == EDLIB_EDOP_MATCH || op == EDLIB_EDOP_MISMATCH) {
              ++qp;
              ++tp;
              if (!found) {
                found = true;
                ft = tp;
                fq = qp;
              }
              lt = tp + 1;
              lq = qp + 1;
              close();
            } else if (op == EDLIB_EDOP_INSERT) {
              ++qp;
            } else {
              ++tp;
              close();
            }
          }
          edlibFreeAlignResult(r);
          return pieces;
        },
        k, o));
  }
  std::vector<Piece> pieces;
  for (auto& f : futures) {
    auto p = f.get();
    pieces.insert(pieces.end(), p.begin(), p.end());
  }

  // ---- 3. windows in the flat layout of rvn_poa_batch ----
  const std::uint64_t n_windows = first_window[T];
  std::vector<std::uint32_t> layers_of(n_windows + 1, 0);
  for (const auto& p : pieces) ++layers_of[p.window + 1];
  std::vector<std::uint32_t> win_first(n_windows + 1, 0);
  for (std::uint64_t w = 0; w < n_windows; ++w) {
    win_first[w + 1] = win_first[w] + 1 + layers_of[w + 1];
  }
  const std::uint32_t n_seqs = win_first[n_windows];
  std::vector<std::uint64_t> seq_off(n_seqs + 1ULL, 0);
  std::vector<std::uint32_t> seq_begin(n_seqs, 0), seq_end(n_seqs, 0), slot(n_windows, 0);
  // sizes first (arrival order of the layers = order of `pieces`)
  std::vector<std::uint32_t> seq_len(n_seqs, 0);
  {
    std::uint64_t w = 0;
    for (std::uint32_t i = 0; i < T; ++i) {
      for (std::uint32_t j = 0; j < targets[i]->inflated_len; j += w_, ++w) {
        seq_len[win_first[w]] = std::min(j + w_, targets[i]->inflated_len) - j;
      }
    }
  }
  std::vector<std::uint32_t> piece_seq(pieces.size());
  for (std::size_t x = 0; x < pieces.size(); ++x) {
    const auto& p = pieces[x];
    const std::uint32_t s = win_first[p.window] + 1 + slot[p.window]++;
    piece_seq[x] = s;
    seq_len[s] = p.qe - p.qb;
    seq_begin[s] = p.begin;
    seq_end[s] = p.end;
  }
  for (std::uint32_t s = 0; s < n_seqs; ++s) seq_off[s + 1] = seq_off[s] + seq_len[s];
  std::string bases(seq_off[n_seqs], 'A'), quals;
  bool any_quality = false;
  for (const auto& s : sequences) any_quality = any_quality || !s->block_quality.empty();
  if (any_quality) quals.assign(seq_off[n_seqs], '!');
  {
    std::uint64_t w = 0;
    for (std::uint32_t i = 0; i < T; ++i) {
      for (std::uint32_t j = 0; j < targets[i]->inflated_len; j += w_, ++w) {
        const std::uint32_t s = win_first[w];
        const std::string bb = targets[i]->InflateData(j, seq_len[s]);
        std::memcpy(&bases[seq_off[s]], bb.data(), bb.size());
      }
    }
  }
  for (std::size_t x = 0; x < pieces.size(); ++x) {
    const auto& p = pieces[x];
    biosoup::NucleicAcid view(*sequences[p.read]);
    if (!p.strand) view.ReverseAndComplement();
    const std::string d = view.InflateData(p.qb, p.qe - p.qb);
    std::memcpy(&bases[seq_off[piece_seq[x]]], d.data(), d.size());
    if (any_quality) {
      if (!view.block_quality.empty()) {
        const std::string qv = view.InflateQuality(p.qb, p.qe - p.qb);
        std::memcpy(&quals[seq_off[piece_seq[x]]], qv.data(), qv.size());
      } else {
        std::memset(&quals[seq_off[piece_seq[x]]], '!' + 1, p.qe - p.qb);  // weight 1
      }
    }
  }

  // ---- 4. GPU: consensus of every window ----
  const auto t0 = std::chrono::steady_clock::now();
  Check(ctx_, rvn_poa_batch(ctx_, static_cast<std::uint32_t>(n_windows), win_first.data(),
                            seq_off.data(), bases.data(),
                            any_quality ? quals.data() : nullptr, seq_begin.data(),
                            seq_end.data(), m_, n_, g_, trim_, tgs, 0));
  const char* cons = nullptr;
  const std::uint64_t* cons_off = nullptr;
  const std::uint8_t* status = nullptr;
  Check(ctx_, rvn_poa_results(ctx_, &cons, &cons_off, &status, nullptr, nullptr));
  poa_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // ---- 5. stitch ----
  std::vector<std::unique_ptr<biosoup::NucleicAcid>> dst;
  for (std::uint32_t i = 0; i < T; ++i) {
    const std::uint64_t w0 = first_window[i], w1 = first_window[i + 1];
    std::string polished(cons + cons_off[w0], cons + cons_off[w1]);
    std::uint32_t n_polished = 0;
    for (std::uint64_t w = w0; w < w1; ++w) n_polished += status[w] & 1;
    num_windows_ += w1 - w0;
    num_polished_windows_ += n_polished;
    const double ratio = w1 > w0 ? n_polished / static_cast<double>(w1 - w0) : 0;
    if (!drop_unpolished || ratio > 0) {
      std::string tags = " LN:i:" + std::to_string(polished.size());
      tags += " RC:i:" + std::to_string(coverage[i]);
      tags += " XC:f:" + std::to_string(ratio);
      dst.emplace_back(new biosoup::NucleicAcid(targets[i]->name + tags, polished));
    }
  }
  return dst;
}

}  // namespace racon
