// Bench utility (not part of the reference interface): seeded synthetic long
// reads straight into the biosoup wire format (2-bit words), multi-threaded and
// independent of the thread count (every read has its own counter-based RNG
// stream). Model: SURVEY.md §8d — i.i.d. uniform genome; read start uniform,
// strand Bernoulli(0.5), genomic span ~ clip(Normal(mu, 0.2 mu), min_len,
// 5 mu); per genome base: deletion / substitution / (base kept +) insertion.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {
  std::uint64_t s;
  explicit Rng(std::uint64_t seed) : s(seed) {}
  std::uint64_t next() {  // splitmix64
    std::uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  double normal() {  // Box-Muller
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

struct Result {
  std::vector<std::uint64_t> words;
  std::vector<std::uint64_t> word_off;
  std::vector<std::uint32_t> lens;
};

template <typename F>
void ParallelFor(std::uint64_t n, unsigned threads, F&& f) {
  std::atomic<std::uint64_t> next{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < threads; ++t) {
    th.emplace_back([&] {
      while (true) {
        std::uint64_t b = next.fetch_add(64);
        if (b >= n) break;
        for (std::uint64_t i = b; i < std::min(n, b + 64); ++i) f(i);
      }
    });
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) void* synth_reads(
    std::uint64_t seed, std::uint64_t genome_len, std::uint32_t n_reads,
    std::uint32_t mean_len, double sub, double ins, double del,
    std::uint32_t min_len, std::uint32_t threads) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  // genome: 2 bits per base, 32 bases per word, chunk-seeded
  const std::uint64_t gwords = (genome_len + 31) / 32;
  std::vector<std::uint64_t> genome(gwords);
  ParallelFor((gwords + 4095) / 4096, threads, [&](std::uint64_t c) {
    Rng r(seed * 0x2545F4914F6CDD1DULL + c + 1);
    for (std::uint64_t i = c * 4096; i < std::min(gwords, (c + 1) * 4096); ++i) {
      genome[i] = r.next();
    }
  });
  auto base = [&](std::uint64_t p) -> std::uint32_t {
    return (genome[p >> 5] >> ((p & 31) << 1)) & 3;
  };

  std::vector<std::vector<std::uint64_t>> packed(n_reads);
  auto* res = new Result();
  res->lens.assign(n_reads, 0);
  const double max_len = 5.0 * mean_len;
  ParallelFor(n_reads, threads, [&](std::uint64_t i) {
    Rng r((seed ^ 0xA5A5A5A5DEADBEEFULL) + 0x632BE59BD9B4E019ULL * (i + 1));
    double l = mean_len + 0.2 * mean_len * r.normal();
    l = std::min(std::max(l, static_cast<double>(min_len)), max_len);
    std::uint64_t span = std::min<std::uint64_t>(static_cast<std::uint64_t>(l), genome_len);
    std::uint64_t start = r.next() % (genome_len - span + 1);
    const bool rc = r.next() & 1;
    auto& out = packed[i];
    out.assign((span + span / 8 + 64) / 32 + 2, 0);
    std::uint64_t n = 0;
    auto push = [&](std::uint64_t c) {
      if ((n >> 5) >= out.size()) out.resize(out.size() * 2, 0);
      out[n >> 5] |= c << ((n & 31) << 1);
      ++n;
    };
    for (std::uint64_t j = 0; j < span; ++j) {
      std::uint32_t c = rc ? 3 - base(start + span - 1 - j) : base(start + j);
      const double u = r.uniform();
      if (u < del) continue;
      if (u < del + sub) c = (c + 1 + r.next() % 3) & 3;
      push(c);
      if (r.uniform() < ins) push(r.next() & 3);
    }
    out.resize((n + 31) / 32);
    res->lens[i] = static_cast<std::uint32_t>(n);
  });

  res->word_off.assign(n_reads + 1ULL, 0);
  for (std::uint32_t i = 0; i < n_reads; ++i) {
    res->word_off[i + 1] = res->word_off[i] + packed[i].size();
  }
  res->words.resize(res->word_off[n_reads]);
  ParallelFor(n_reads, threads, [&](std::uint64_t i) {
    std::copy(packed[i].begin(), packed[i].end(),
              res->words.begin() + res->word_off[i]);
    std::vector<std::uint64_t>().swap(packed[i]);
  });
  return res;
}

// Draft contigs of the SAME genome (polishing targets of the C3-shaped bench):
// the genome cut into pieces of contig_len bases, each with draft-like errors.
__attribute__((visibility("default"))) void* synth_contigs(
    std::uint64_t seed, std::uint64_t genome_len, std::uint64_t contig_len, double sub,
    double ins, double del, std::uint32_t threads) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  const std::uint64_t gwords = (genome_len + 31) / 32;
  std::vector<std::uint64_t> genome(gwords);
  ParallelFor((gwords + 4095) / 4096, threads, [&](std::uint64_t c) {
    Rng r(seed * 0x2545F4914F6CDD1DULL + c + 1);
    for (std::uint64_t i = c * 4096; i < std::min(gwords, (c + 1) * 4096); ++i) {
      genome[i] = r.next();
    }
  });
  auto base = [&](std::uint64_t p) -> std::uint32_t {
    return (genome[p >> 5] >> ((p & 31) << 1)) & 3;
  };
  const std::uint32_t n = static_cast<std::uint32_t>((genome_len + contig_len - 1) / contig_len);
  std::vector<std::vector<std::uint64_t>> packed(n);
  auto* res = new Result();
  res->lens.assign(n, 0);
  ParallelFor(n, threads, [&](std::uint64_t i) {
    Rng r((seed ^ 0x0123456789ABCDEFULL) + 0x9E3779B97F4A7C15ULL * (i + 1));
    const std::uint64_t start = i * contig_len, span = std::min(contig_len, genome_len - start);
    auto& out = packed[i];
    out.assign((span + span / 8 + 64) / 32 + 2, 0);
    std::uint64_t m = 0;
    auto push = [&](std::uint64_t c) {
      if ((m >> 5) >= out.size()) out.resize(out.size() * 2, 0);
      out[m >> 5] |= c << ((m & 31) << 1);
      ++m;
    };
    for (std::uint64_t j = 0; j < span; ++j) {
      std::uint32_t c = base(start + j);
      const double u = r.uniform();
      if (u < del) continue;
      if (u < del + sub) c = (c + 1 + r.next() % 3) & 3;
      push(c);
      if (r.uniform() < ins) push(r.next() & 3);
    }
    out.resize((m + 31) / 32);
    res->lens[i] = static_cast<std::uint32_t>(m);
  });
  res->word_off.assign(n + 1ULL, 0);
  for (std::uint32_t i = 0; i < n; ++i) res->word_off[i + 1] = res->word_off[i] + packed[i].size();
  res->words.resize(res->word_off[n]);
  for (std::uint32_t i = 0; i < n; ++i) {
    std::copy(packed[i].begin(), packed[i].end(), res->words.begin() + res->word_off[i]);
  }
  return res;
}
__attribute__((visibility("default"))) std::uint32_t synth_n_reads(void* h) {
  return static_cast<std::uint32_t>(static_cast<Result*>(h)->lens.size());
}

__attribute__((visibility("default"))) std::uint64_t synth_n_words(void* h) {
  return static_cast<Result*>(h)->words.size();
}
__attribute__((visibility("default"))) const std::uint64_t* synth_words(void* h) {
  return static_cast<Result*>(h)->words.data();
}
__attribute__((visibility("default"))) const std::uint64_t* synth_word_off(void* h) {
  return static_cast<Result*>(h)->word_off.data();
}
__attribute__((visibility("default"))) const std::uint32_t* synth_lens(void* h) {
  return static_cast<Result*>(h)->lens.data();
}
__attribute__((visibility("default"))) void synth_free(void* h) {
  delete static_cast<Result*>(h);
}

}  // extern "C"
