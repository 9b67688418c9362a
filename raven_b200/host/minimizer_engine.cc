// raven_b200 — ram::MinimizerEngine facade over the C ABI (include/raven_b200.h).
// Same signatures, argument meaning and error behaviour as the reference's
// dependency (call sites: RavenLib/src/construct.cc:42-44,62,363,372,377-381,
// 661-662; assemble.cc:753-780).
#include "ram/minimizer_engine.hpp"

#include <stdexcept>
#include <string>

#include "raven_b200.h"

namespace ram {

namespace {

void Check(rvn_ctx* ctx, int rc, const char* what) {
  if (rc == RVN_OK) return;
  std::string msg = std::string("[ram::MinimizerEngine::") + what + "] error: " +
                    (ctx ? rvn_last_error(ctx) : "no context");
  if (rc == RVN_ERR_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

}  // namespace

MinimizerEngine::MinimizerEngine(
    std::shared_ptr<thread_pool::ThreadPool> thread_pool, std::uint32_t k,
    std::uint32_t w, std::uint32_t bandwidth, std::uint32_t chain,
    std::uint32_t matches, std::uint32_t gap)
    : ctx_(nullptr),
      mutex_(new std::mutex()),
      occurrence_(-1),
      thread_pool_(thread_pool) {
  int rc = rvn_ctx_create(0, nullptr, &ctx_);
  if (rc != RVN_OK) {
    // no CPU fallback: the B200 engine is the only implementation
    throw std::runtime_error(
        "[ram::MinimizerEngine::MinimizerEngine] error: no usable CUDA device");
  }
  Check(ctx_, rvn_engine_configure(ctx_, k, w, bandwidth, chain, matches, gap),
        "MinimizerEngine");
}

MinimizerEngine::MinimizerEngine(MinimizerEngine&& o) noexcept
    : ctx_(o.ctx_),
      mutex_(std::move(o.mutex_)),
      occurrence_(o.occurrence_),
      uploaded_(std::move(o.uploaded_)),
      thread_pool_(std::move(o.thread_pool_)) {
  o.ctx_ = nullptr;
}

MinimizerEngine& MinimizerEngine::operator=(MinimizerEngine&& o) noexcept {
  if (this != &o) {
    if (ctx_) rvn_ctx_destroy(ctx_);
    ctx_ = o.ctx_;
    o.ctx_ = nullptr;
    mutex_ = std::move(o.mutex_);
    occurrence_ = o.occurrence_;
    uploaded_ = std::move(o.uploaded_);
    thread_pool_ = std::move(o.thread_pool_);
  }
  return *this;
}

MinimizerEngine::~MinimizerEngine() {
  if (ctx_) rvn_ctx_destroy(ctx_);
}

void MinimizerEngine::UploadRange(
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last) {
  // the same objects as last time (stage 1 -> identity filter): nothing to move
  {
    const std::size_t n = static_cast<std::size_t>(last - first);
    bool same = n > 0 && n == uploaded_.size();
    std::uint32_t pos = 0;
    for (auto it = first; same && it != last; ++it, ++pos) {
      const auto f = uploaded_.find((*it)->id);
      same = f != uploaded_.end() && f->second.first == pos && f->second.second == it->get();
    }
    if (same) return;
  }
  // deflated_data is already the device format: concatenate, never repack
  std::vector<std::uint64_t> words, off{0};
  std::vector<std::uint32_t> lens, ids;
  uploaded_.clear();
  std::uint32_t pos = 0;
  for (auto it = first; it != last; ++it, ++pos) {
    const auto& s = **it;
    if (s.is_reverse_complement) {
      throw std::invalid_argument(
          "[ram::MinimizerEngine::Minimize] error: reverse-complemented view");
    }
    words.insert(words.end(), s.deflated_data.begin(), s.deflated_data.end());
    off.push_back(words.size());
    lens.push_back(s.inflated_len);
    ids.push_back(s.id);
    uploaded_[s.id] = {pos, &s};
  }
  Check(ctx_,
        rvn_reads_upload_ids(ctx_, words.data(), off.data(), lens.data(),
                             ids.data(), pos),
        "Minimize");
}

void MinimizerEngine::Upload(
    const std::vector<std::unique_ptr<biosoup::NucleicAcid>>& sequences) {
  std::lock_guard<std::mutex> lock(*mutex_);
  UploadRange(sequences.begin(), sequences.end());
}

void MinimizerEngine::Upload(
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last) {
  std::lock_guard<std::mutex> lock(*mutex_);
  UploadRange(first, last);
}

void MinimizerEngine::Minimize(
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator first,
    std::vector<std::unique_ptr<biosoup::NucleicAcid>>::const_iterator last,
    bool minhash) {
  std::lock_guard<std::mutex> lock(*mutex_);
  occurrence_ = -1;
  if (first >= last) {
    UploadRange(first, first);
    Check(ctx_, rvn_minimize(ctx_, 0, 0, minhash), "Minimize");
    return;
  }
  UploadRange(first, last);
  Check(ctx_, rvn_minimize(ctx_, 0, static_cast<std::uint32_t>(last - first), minhash),
        "Minimize");
}

void MinimizerEngine::Filter(double frequency) {
  std::lock_guard<std::mutex> lock(*mutex_);
  Check(ctx_, rvn_filter(ctx_, frequency, &occurrence_), "Filter");
}

std::vector<biosoup::Overlap> MinimizerEngine::Map(
    const std::unique_ptr<biosoup::NucleicAcid>& sequence, bool avoid_equal,
    bool avoid_symmetric, bool minhash,
    std::vector<std::uint32_t>* filtered) const {
  std::lock_guard<std::mutex> lock(*mutex_);
  const auto& s = *sequence;
  auto it = uploaded_.find(s.id);
  if (it != uploaded_.end() && it->second.second == &s && !s.is_reverse_complement) {
    const std::uint32_t pos = it->second.first;
    Check(ctx_,
          rvn_map(ctx_, pos, pos + 1, avoid_equal, avoid_symmetric, minhash,
                  filtered != nullptr),
          "Map");
  } else {
    // a read outside the indexed batch (construct.cc:59 maps reads 0..i against
    // the index of reads j..i): it rides in the spare device slot
    std::vector<std::uint64_t> rc_words;
    const std::uint64_t* words = s.deflated_data.data();
    if (s.is_reverse_complement) {
      rc_words.assign((static_cast<std::uint64_t>(s.inflated_len) + 31) >> 5, 0);
      for (std::uint32_t i = 0; i < s.inflated_len; ++i) {
        rc_words[i >> 5] |= s.Code(i) << ((i << 1) & 63);
      }
      words = rc_words.data();
    }
    Check(ctx_,
          rvn_map_external(ctx_, words, s.inflated_len, s.id, avoid_equal,
                           avoid_symmetric, minhash, filtered != nullptr),
          "Map");
  }
  const rvn_overlap* o = nullptr;
  const std::uint64_t* off = nullptr;
  const std::uint32_t* f = nullptr;
  const std::uint64_t* foff = nullptr;
  std::uint64_t n = 0;
  Check(ctx_, rvn_map_results(ctx_, &o, &off, &n, &f, &foff), "Map");
  std::vector<biosoup::Overlap> dst;
  dst.reserve(n);
  for (std::uint64_t i = 0; i < n; ++i) {
    dst.emplace_back(o[i].lhs_id, o[i].lhs_begin, o[i].lhs_end, o[i].rhs_id,
                     o[i].rhs_begin, o[i].rhs_end, o[i].score, o[i].strand != 0);
  }
  if (filtered) {
    filtered->insert(filtered->end(), f + foff[0], f + foff[1]);
  }
  return dst;
}

}  // namespace ram
