// raven_b200 — batched global (NW) edit distance on sm_100a.
//
// Replaces the edlibAlign(lhs, rhs, edlibDefaultAlignConfig()) calls of the
// identity filter (RavenLib/src/construct.cc:162-217 ResolveContainedReads and
// :385-424 FindOverlapsAndRepetetiveRegions; score = 1 - ed / max(len)): up to
// 3e7 alignments of 10-15 kb substrings on the HiFi configuration. Only the
// integer distance is needed, so any exact method is bit-identical with edlib.
//
// Myers/Hyyro bit-vector blocks (64 query rows per u64) inside an Ukkonen band
// |row - column| <= k. One THREAD per pair: the band of a column spans at most
// NB blocks whose vertical delta vectors (pv, mv) live in registers as a window
// that slides down one block at a time; the match masks of the window's blocks
// sit in shared memory ([symbol][slot][thread], conflict free), built on the
// fly from the 2-bit packed reads (no inflated strings anywhere); the reverse
// complement of the rhs substring is read in place. A pair whose distance
// exceeds the band of one tier (k = 64, 192, 448, 960) moves to the next; what
// is left (or has no bound at all) runs the unbanded kernel with its vectors in
// global scratch. Integer ALU bound (about 40 instructions per block and
// column), no HBM pressure: 0.25 B per base once.
#include <algorithm>
#include <numeric>

#include "engine.cuh"
#include "myers.cuh"

namespace rvn {

namespace {

struct PairDesc {
  uint64_t q_word;   // first word of the lhs read in `words`
  uint64_t t_word;   // first word of the rhs read
  uint32_t q_begin, q_len;
  uint32_t t_begin, t_len;
  uint32_t strand;   // 1 = same strand; 0 = rhs substring reverse complemented
  uint32_t slot;     // index of the pair in the caller's arrays
};

// match mask of query block b (rows 64b .. 64b+63 of the lhs substring)
__device__ __forceinline__ uint64_t BlockEq(const uint64_t* __restrict__ qw, uint64_t q_words,
                                            uint32_t q_begin, uint32_t m, uint32_t b,
                                            uint32_t sym) {
  uint64_t w0, w1;
  Bases64(qw, static_cast<uint64_t>(q_begin) + 64ULL * b, q_words, &w0, &w1);
  uint64_t eq = EqMask32(w0, sym) | (EqMask32(w1, sym) << 32);
  const uint32_t rows = min(64u, m - 64u * b);
  if (rows < 64) eq &= (1ULL << rows) - 1ULL;
  return eq;
}

// target symbol of column j (1-based) of the rhs substring, reverse
// complemented in place when !strand
__device__ __forceinline__ uint32_t TargetSym(const uint64_t* __restrict__ tw, const PairDesc& d,
                                              uint32_t j) {
  if (d.strand) return BaseAt(tw, static_cast<uint64_t>(d.t_begin) + j - 1);
  return 3u - BaseAt(tw, static_cast<uint64_t>(d.t_begin) + d.t_len - j);
}

// Banded distance, window of NB blocks in registers. k <= 32 * (NB - 2).
// out[slot] = distance if it is <= k, else -1 ("beyond this tier").
template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
BandedMyersKernel(const uint64_t* __restrict__ words, const uint64_t* __restrict__ woff_unused,
                  const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ list,
                  uint32_t n_list, int k, const int32_t* __restrict__ limit,
                  int32_t* __restrict__ out, uint64_t total_words) {
  extern __shared__ uint64_t sh_eq[];  // [4][NB][THREADS]
  (void)woff_unused;
  const uint32_t gi = blockIdx.x * THREADS + threadIdx.x;
  if (gi >= n_list) return;
  const PairDesc d = pairs[list[gi]];
  const int m = static_cast<int>(d.q_len), n = static_cast<int>(d.t_len);
  // the caller's bound for this pair (-1: none): never look beyond it
  const int lim = limit[d.slot];
  int kk = k;
  if (lim >= 0 && lim < kk) kk = lim;
  int result = -1;
  if (m == 0 || n == 0) {
    result = (m + n <= kk) ? m + n : -1;
    out[d.slot] = result;
    return;
  }
  if (abs(m - n) > kk) {
    out[d.slot] = -1;
    return;
  }
  const uint64_t* qw = words + d.q_word;
  const uint64_t* tw = words + d.t_word;
  const uint64_t q_words = total_words - d.q_word;  // reads beyond: zero bases, masked rows
  const int blocks = (m + 63) >> 6;
  const uint64_t last_high = 1ULL << ((m - 1) & 63);

  uint64_t pv[NB], mv[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    pv[i] = ~0ULL;
    mv[i] = 0;
  }
  int wlo = 0;                                   // first block of the window
  int hi = (min(m, max(1, kk)) - 1) >> 6;        // last active block (column 0: rows 1..k)
  int score = min(m, (hi + 1) << 6);             // D[bottom row of block hi][column]
  auto eq_at = [&](uint32_t sym, int slot) -> uint64_t& {
    return sh_eq[(sym * NB + slot) * THREADS + threadIdx.x];
  };
  for (int b = 0; b <= hi; ++b) {
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) eq_at(s, b % NB) = BlockEq(qw, q_words, d.q_begin, m, b, s);
  }

  for (int j = 1; j <= n; ++j) {
    const int want_hi = (min(m, j + kk) - 1) >> 6;
    if (hi < want_hi) {  // a block enters the band: vertical deltas all +1
      ++hi;
      // (its register slot hi - wlo holds pv = ~0, mv = 0: set at start / when sliding)
      score += min(64, m - (hi << 6));
#pragma unroll
      for (uint32_t s = 0; s < 4; ++s) eq_at(s, hi % NB) = BlockEq(qw, q_words, d.q_begin, m, hi, s);
    }
    const int want_lo = (max(1, j - kk) - 1) >> 6;
    if (want_lo > wlo) {  // the window slides down one block
#pragma unroll
      for (int i = 0; i + 1 < NB; ++i) {
        pv[i] = pv[i + 1];
        mv[i] = mv[i + 1];
      }
      pv[NB - 1] = ~0ULL;
      mv[NB - 1] = 0;
      ++wlo;
    }
    const uint32_t sym = TargetSym(tw, d, j);
    int h = 1;  // row 0 for wlo == 0; an upper bound once the band has left row 0
    const int top = hi - wlo;  // last active register slot
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      if (i <= top) {
        const int b = wlo + i;
        h = Step(eq_at(sym, b % NB), h, b == blocks - 1 ? last_high : (1ULL << 63), pv[i], mv[i]);
      }
    }
    score += h;
  }
  if (hi == blocks - 1 && score <= kk) result = score;
  out[d.slot] = result;
}

// Unbanded distance for what the tiers left: vectors of all query blocks in
// global scratch, [block][thread of the launch] (coalesced across the warp).
__global__ void __launch_bounds__(64)
FullMyersKernel(const uint64_t* __restrict__ words, const PairDesc* __restrict__ pairs,
                const uint32_t* __restrict__ list, uint32_t n_list, uint32_t max_blocks,
                uint64_t* __restrict__ scratch, int32_t* __restrict__ out, uint64_t total_words) {
  const uint32_t gi = blockIdx.x * 64 + threadIdx.x;
  if (gi >= n_list) return;
  const PairDesc d = pairs[list[gi]];
  const int m = static_cast<int>(d.q_len), n = static_cast<int>(d.t_len);
  if (m == 0 || n == 0) {
    out[d.slot] = m + n;
    return;
  }
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * 64;
  uint64_t* pv = scratch + gi;                                // pv[b * stride]
  uint64_t* mv = scratch + static_cast<uint64_t>(max_blocks) * stride + gi;
  const uint64_t* qw = words + d.q_word;
  const uint64_t* tw = words + d.t_word;
  const uint64_t q_words = total_words - d.q_word;
  const int blocks = (m + 63) >> 6;
  const uint64_t last_high = 1ULL << ((m - 1) & 63);
  for (int b = 0; b < blocks; ++b) {
    pv[b * stride] = ~0ULL;
    mv[b * stride] = 0;
  }
  int score = m;
  for (int j = 1; j <= n; ++j) {
    const uint32_t sym = TargetSym(tw, d, j);
    int h = 1;
    for (int b = 0; b < blocks; ++b) {
      uint64_t p = pv[b * stride], q = mv[b * stride];
      h = Step(BlockEq(qw, q_words, d.q_begin, m, b, sym), h,
               b == blocks - 1 ? last_high : (1ULL << 63), p, q);
      pv[b * stride] = p;
      mv[b * stride] = q;
    }
    score += h;
  }
  out[d.slot] = score;
}

template <int NB, int THREADS>
void LaunchTier(Ctx& c, const PairDesc* d_pairs, const uint32_t* d_list, uint32_t n_list, int k,
                const int32_t* d_limit, int32_t* d_out) {
  auto kern = BandedMyersKernel<NB, THREADS>;
  const size_t smem = sizeof(uint64_t) * 4 * NB * THREADS;
  RVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  kern<<<CeilDiv(n_list, THREADS), THREADS, smem, c.stream>>>(
      c.d_words.get(), c.d_woff.get(), d_pairs, d_list, n_list, k, d_limit, d_out, c.n_words + 2);
  RVN_LAUNCH_CHECK();
  ++c.launches;
}

}  // namespace

// Global edit distances of n substring pairs of the uploaded reads. limit[i] >= 0:
// distances above it are reported as -1 (the caller only needs "beyond").
void EditDistanceBatch(Ctx& c, uint64_t n, const uint32_t* lhs_read, const uint32_t* lhs_begin,
                       const uint32_t* lhs_len, const uint32_t* rhs_read,
                       const uint32_t* rhs_begin, const uint32_t* rhs_len,
                       const uint8_t* strand, const int32_t* limit, int32_t* out) {
  if (n == 0) return;
  if (n >= 0xFFFFFFFFULL) throw LimitError("2^32 or more pairs");
  std::vector<PairDesc> h(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t a = lhs_read[i], b = rhs_read[i];
    if (a >= c.n_reads || b >= c.n_reads) throw InvalidArgument("read index out of bounds");
    if (a < c.res_first || a >= c.res_last || b < c.res_first || b >= c.res_last) {
      throw StateError("the bases of these reads were not uploaded (rvn_reads_upload_range)");
    }
    if (static_cast<uint64_t>(lhs_begin[i]) + lhs_len[i] > c.h_len[a] ||
        static_cast<uint64_t>(rhs_begin[i]) + rhs_len[i] > c.h_len[b]) {
      throw InvalidArgument("substring beyond the end of its read");
    }
    h[i] = PairDesc{c.h_woff[a], c.h_woff[b], lhs_begin[i], lhs_len[i], rhs_begin[i],
                    rhs_len[i], strand[i] ? 1u : 0u, static_cast<uint32_t>(i)};
  }
  TimerBegin(c, "edit_distance");
  DevBuf<PairDesc> d_pairs;
  DevBuf<uint32_t> d_list;
  DevBuf<int32_t> d_out, d_limit;
  d_pairs.reserve(n);
  d_list.reserve(n);
  d_out.reserve(n);
  d_limit.reserve(n);
  std::vector<int32_t> h_limit(n), h_out(n, -1);
  for (uint64_t i = 0; i < n; ++i) h_limit[i] = limit ? limit[i] : -1;
  RVN_CUDA(cudaMemcpyAsync(d_pairs.get(), h.data(), n * sizeof(PairDesc), cudaMemcpyHostToDevice,
                           c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_limit.get(), h_limit.data(), n * 4, cudaMemcpyHostToDevice, c.stream));

  // work list, longest pairs first (lanes of a warp then carry similar work)
  std::vector<uint32_t> todo(n);
  std::iota(todo.begin(), todo.end(), 0u);
  auto cost = [&](uint32_t i) { return static_cast<uint64_t>(h[i].q_len) + h[i].t_len; };
  std::sort(todo.begin(), todo.end(), [&](uint32_t a, uint32_t b) { return cost(a) > cost(b); });

  static const int kTierK[4] = {64, 192, 448, 960};
  for (int tier = 0; tier <= 4 && !todo.empty(); ++tier) {
    const uint32_t m = static_cast<uint32_t>(todo.size());
    RVN_CUDA(cudaMemcpyAsync(d_list.get(), todo.data(), m * 4ULL, cudaMemcpyHostToDevice, c.stream));
    if (tier == 0) LaunchTier<4, 128>(c, d_pairs.get(), d_list.get(), m, kTierK[0], d_limit.get(), d_out.get());
    if (tier == 1) LaunchTier<8, 128>(c, d_pairs.get(), d_list.get(), m, kTierK[1], d_limit.get(), d_out.get());
    if (tier == 2) LaunchTier<16, 64>(c, d_pairs.get(), d_list.get(), m, kTierK[2], d_limit.get(), d_out.get());
    if (tier == 3) LaunchTier<32, 64>(c, d_pairs.get(), d_list.get(), m, kTierK[3], d_limit.get(), d_out.get());
    if (tier == 4) {
      // unbanded: scratch for the vectors of every query block of every pair in flight
      uint32_t max_blocks = 1;
      for (uint32_t i : todo) max_blocks = std::max(max_blocks, (h[i].q_len + 63) / 64);
      const uint64_t threads = static_cast<uint64_t>(CeilDiv(m, 64)) * 64;
      uint64_t* scratch = c.m_scratch64.reserve(2ULL * max_blocks * threads + 16);
      FullMyersKernel<<<CeilDiv(m, 64), 64, 0, c.stream>>>(c.d_words.get(), d_pairs.get(),
                                                          d_list.get(), m, max_blocks, scratch,
                                                          d_out.get(), c.n_words + 2);
      RVN_LAUNCH_CHECK();
      ++c.launches;
    }
    RVN_CUDA(cudaMemcpyAsync(h_out.data(), d_out.get(), n * 4, cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaStreamSynchronize(c.stream));
    // pairs still open: not found within this tier's band and allowed to look further
    std::vector<uint32_t> next;
    for (uint32_t i : todo) {
      if (h_out[i] >= 0 || tier == 4) continue;
      const int lim = h_limit[i];
      if (lim >= 0 && lim <= kTierK[tier]) continue;  // beyond the caller's bound: stays -1
      next.push_back(i);
    }
    todo.swap(next);
  }
  TimerEnd(c);
  for (uint64_t i = 0; i < n; ++i) {
    int32_t v = h_out[i];
    if (limit && limit[i] >= 0 && v > limit[i]) v = -1;
    out[i] = v;
  }
}

}  // namespace rvn
