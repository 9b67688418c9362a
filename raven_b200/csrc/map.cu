// raven_b200 — seed lookup, hit expansion and chaining on sm_100a.
//
// Replaces ram::MinimizerEngine::Map + Chain (un-vendored; call sites
// RavenLib/src/construct.cc:59-64 -> Map(seq,1,1,1) and :377-381 ->
// Map(seq,1,1,0,&filtered); algorithm SURVEY.md App. A.2).
//
//   probe   one thread per query minimizer: bucket table -> sorted run ->
//           (first posting, count) or "filtered" if count > occurrence
//   expand  hits (ram "Match": group = (rhs_id<<1|same_strand)<<32|diagonal,
//           positions = lhs_pos<<32|rhs_pos) written per query read
//   chain   ONE CTA PER QUERY READ, hits resident in shared memory:
//           bitonic sort by (group, positions) -> diagonal-band intervals by
//           binary searches + block scans -> second sort by (band, positions)
//           -> one thread per band runs ram's patience/LIS with its exact
//           binary-search predicate, gap split and covered-bases test ->
//           overlaps written through a reserved slab, then re-ordered by
//           query id so that the output order equals the reference's.
//   Reads whose hits do not fit in shared memory take the same code path
//   over a global-memory scratch slab.
//
// The chain result is a pure function of the MULTISET of hits of a query
// (both reference sorts are total orders here: equal (group, positions)
// pairs cannot occur), so hit generation order is free.
#include <algorithm>

#include "engine.cuh"

namespace rvn {

namespace {

constexpr int kThreads = 256;

struct IndexView {
  const uint64_t* val;
  const uint64_t* org;
  const uint32_t* bucket;
  uint64_t n;
  int shift;
  uint32_t occurrence;
};

// first record with value v and the run length capped at occurrence+1
__device__ __forceinline__ void Lookup(const IndexView& ix, uint64_t v,
                                       uint32_t* first, uint32_t* count) {
  const uint64_t b = v >> ix.shift;
  uint32_t lo = ix.bucket[b], hi = ix.bucket[b + 1];
  while (hi - lo > 8) {  // long buckets: bisect down to a short scan
    const uint32_t mid = lo + (hi - lo) / 2;
    if (ix.val[mid] < v) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  // here every record before lo is < v; the run (if any) starts in [lo, hi]
  const uint32_t end = ix.bucket[b + 1];
  while (lo < end && ix.val[lo] < v) ++lo;
  if (lo >= end || ix.val[lo] != v) {
    *first = 0;
    *count = 0;
    return;
  }
  *first = lo;
  if (ix.occurrence != 0xFFFFFFFFu &&
      static_cast<uint64_t>(lo) + ix.occurrence < ix.n &&
      ix.val[static_cast<uint64_t>(lo) + ix.occurrence] == v) {
    *count = ix.occurrence + 1;  // over the threshold, exact length not needed
    return;
  }
  uint32_t n = 1;
  while (static_cast<uint64_t>(lo) + n < ix.n && ix.val[lo + n] == v) ++n;
  *count = n;
}

__device__ __forceinline__ bool KeepPosting(uint32_t lhs_id, uint64_t origin,
                                            bool avoid_equal,
                                            bool avoid_symmetric) {
  const uint32_t rhs_id = static_cast<uint32_t>(origin >> 32);
  if (avoid_equal && lhs_id == rhs_id) return false;
  if (avoid_symmetric && lhs_id > rhs_id) return false;
  return true;
}

__global__ void __launch_bounds__(kThreads)
ProbeKernel(IndexView ix, const uint64_t* __restrict__ q_val,
            const uint64_t* __restrict__ q_org, uint64_t q_begin, uint64_t n_q,
            bool avoid_equal, bool avoid_symmetric,
            uint32_t* __restrict__ cnt, uint32_t* __restrict__ first,
            uint8_t* __restrict__ filt) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n_q) return;
  const uint64_t v = q_val[q_begin + i];
  const uint32_t lhs_id = static_cast<uint32_t>(q_org[q_begin + i] >> 32);
  uint32_t f, n;
  Lookup(ix, v, &f, &n);
  uint32_t kept = 0;
  uint8_t over = 0;
  if (n > ix.occurrence) {
    over = 1;
    n = 0;
  } else {
    for (uint32_t j = 0; j < n; ++j) {
      kept += KeepPosting(lhs_id, ix.org[f + j], avoid_equal, avoid_symmetric);
    }
  }
  cnt[i] = kept;
  first[i] = f;
  filt[i] = over;
  // the posting count is re-derived in ExpandKernel from the run itself
}

__global__ void __launch_bounds__(kThreads)
ExpandKernel(IndexView ix, const uint64_t* __restrict__ q_val,
             const uint64_t* __restrict__ q_org, uint64_t q_begin, uint64_t n_q,
             bool avoid_equal, bool avoid_symmetric,
             const uint32_t* __restrict__ cnt,
             const uint32_t* __restrict__ first,
             const uint64_t* __restrict__ hit_off, uint64_t* __restrict__ h_grp,
             uint64_t* __restrict__ h_pos) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n_q) return;
  uint32_t left = cnt[i];
  if (left == 0) return;
  const uint64_t v = q_val[q_begin + i];
  const uint64_t lo = q_org[q_begin + i];
  const uint32_t lhs_id = static_cast<uint32_t>(lo >> 32);
  const uint64_t lhs_pos = static_cast<uint32_t>(lo) >> 1;
  uint64_t dst = hit_off[i];
  for (uint64_t j = first[i]; left > 0 && j < ix.n && ix.val[j] == v; ++j) {
    const uint64_t o = ix.org[j];
    if (!KeepPosting(lhs_id, o, avoid_equal, avoid_symmetric)) continue;
    const uint64_t rhs_id = o >> 32;
    const uint64_t strand = (lo & 1) == (o & 1);
    const uint64_t rhs_pos = static_cast<uint32_t>(o) >> 1;
    const uint64_t diagonal =
        !strand ? rhs_pos + lhs_pos : rhs_pos - lhs_pos + (3ULL << 30);
    h_grp[dst] = (((rhs_id << 1) | strand) << 32) | diagonal;
    h_pos[dst] = (lhs_pos << 32) | rhs_pos;
    ++dst;
    --left;
  }
}

// per-read offsets out of per-record offsets
__global__ void GatherU64(const uint64_t* __restrict__ src,
                          const uint64_t* __restrict__ idx, uint64_t idx_base,
                          uint64_t n, uint64_t* __restrict__ dst) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = src[idx[i] - idx_base];
}

// positions of over-frequent query minimizers, compacted in sketch order
__global__ void FilteredFlagsToU32(const uint8_t* __restrict__ filt, uint64_t n,
                                   uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = filt[i];
}
__global__ void ScatterFiltered(const uint8_t* __restrict__ filt,
                                const uint64_t* __restrict__ pos,
                                const uint64_t* __restrict__ q_org,
                                uint64_t q_begin, uint64_t n,
                                uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !filt[i]) return;
  out[pos[i]] = static_cast<uint32_t>(q_org[q_begin + i]) >> 1;
}

// ---------------------------------------------------------------------------
// chaining
// ---------------------------------------------------------------------------

struct ChainParams {
  uint32_t k, bandwidth, chain, matches, gap;
};

// sorts (A[i], B[i]) pairs ascending by (A, B); npad is a power of two
template <int THREADS>
__device__ void BitonicSortPairs(uint64_t* A, uint64_t* B, uint32_t npad) {
  for (uint32_t size = 2; size <= npad; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < (npad >> 1); t += THREADS) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const bool up = (i & size) == 0;
        const uint64_t ai = A[i], aj = A[j], bi = B[i], bj = B[j];
        const bool gt = ai > aj || (ai == aj && bi > bj);
        if (gt == up) {
          A[i] = aj;
          A[j] = ai;
          B[i] = bj;
          B[j] = bi;
        }
      }
      __syncthreads();
    }
  }
}

// Everything a CTA needs to chain the hits of one query read. IdxT = u16 for
// the shared-memory path (n <= 65534), u32 for the global scratch path.
template <typename IdxT>
struct ChainWork {
  uint64_t* G;   // npad   group, later band tag
  uint64_t* P;   // npad   positions
  IdxT* LB;      // n + 1  lower bounds, later LIS "minimal"+indices (n + nb + 1)
  IdxT* PD;      // n      predecessor
  IdxT* IB;      // n/4+1  band begin
  IdxT* IE;      // n/4+1  band end
  uint32_t* CNT; // n/4+1  overlaps per band
};

template <typename IdxT, int THREADS>
__device__ uint32_t ChainRead(const ChainWork<IdxT>& wk, uint32_t n,
                              uint32_t npad, uint32_t lhs_id,
                              const ChainParams& cp, uint32_t* sm32,
                              rvn_overlap* __restrict__ ovl_raw,
                              unsigned long long* __restrict__ ovl_counter,
                              uint64_t ovl_cap, uint64_t* out_base) {
  uint64_t* G = wk.G;
  uint64_t* P = wk.P;
  __shared__ unsigned long long sh_base;

  // 1. order by (group, positions); padding is all-ones and sorts last, and
  //    G[n] doubles as the reference's stop dummy
  BitonicSortPairs<THREADS>(G, P, npad);

  // 2. lower bounds: LB[i] = first j with G[i] - G[j] <= bandwidth
  for (uint32_t i = threadIdx.x; i <= n; i += THREADS) {
    const uint64_t gi = i < n ? G[i] : ~0ULL;
    uint32_t lo = 0, hi = i;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (gi - G[mid] <= cp.bandwidth) {
        hi = mid;
      } else {
        lo = mid + 1;
      }
    }
    wk.LB[i] = static_cast<IdxT>(lo);
  }
  __syncthreads();

  // 3. events: at i (1..n) the window [LB[i-1], i) closes; it qualifies with
  //    >= 4 hits; a qualifying window overlapping the previous qualifying
  //    one extends it. Done chunk-wise with carried state.
  uint32_t carry_prevq = 0;  // last qualifying event index so far (0 = none)
  uint32_t carry_nb = 0;     // bands opened so far
  for (uint32_t base = 1; base <= n; base += THREADS) {
    const uint32_t i = base + threadIdx.x;
    uint32_t e = 0, jp = 0;
    if (i <= n) {
      jp = wk.LB[i - 1];
      const uint64_t gi = i < n ? G[i] : ~0ULL;
      e = (gi - G[jp] > cp.bandwidth) && (i - jp >= 4);
    }
    // previous qualifying event (exclusive max-scan of e ? i : 0)
    const uint32_t incl = BlockInclusiveMax<uint32_t, THREADS>(e ? i : 0u, sm32);
    uint32_t prevq_incl_before;  // max over lanes < this one
    {
      // shift by one lane: recompute exclusive from inclusive of neighbours
      __shared__ uint32_t sh_incl[THREADS];
      sh_incl[threadIdx.x] = incl;
      __syncthreads();
      prevq_incl_before = threadIdx.x ? sh_incl[threadIdx.x - 1] : 0u;
      __syncthreads();
    }
    const uint32_t prevq = max(carry_prevq, prevq_incl_before);
    const uint32_t start = e && (prevq == 0 || prevq <= jp);
    uint32_t tot;
    const uint32_t ex = BlockExclusiveSum<uint32_t, THREADS>(start, sm32, &tot);
    if (e) {
      const uint32_t band = carry_nb + ex + start - 1;  // band this event feeds
      if (start) {
        wk.IB[band] = static_cast<IdxT>(jp);
        // the previous band (if any) ended at the previous qualifying event
        if (prevq != 0) wk.IE[band - 1] = static_cast<IdxT>(prevq);
      }
    }
    carry_nb += tot;
    // last qualifying event of the chunk
    {
      __shared__ uint32_t sh_last;
      if (threadIdx.x == THREADS - 1) sh_last = max(carry_prevq, incl);
      __syncthreads();
      carry_prevq = sh_last;
      __syncthreads();
    }
  }
  const uint32_t nb = carry_nb;
  if (threadIdx.x == 0 && nb) wk.IE[nb - 1] = static_cast<IdxT>(carry_prevq);
  __syncthreads();
  if (nb == 0) return 0;

  // 4. tag hits with their band (2b+1) or the gap before band b (2b), then
  //    order each band by positions with one more (tag, positions) sort
  for (uint32_t i = threadIdx.x; i < npad; i += THREADS) {
    if (i >= n) {
      G[i] = ~0ULL;
      continue;
    }
    // band with the largest begin <= i
    uint32_t lo = 0, hi = nb;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (wk.IB[mid] <= i) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    // lo = number of bands beginning at or before i
    uint64_t tag;
    if (lo == 0) {
      tag = 0;
    } else if (i < wk.IE[lo - 1]) {
      tag = 2ULL * (lo - 1) + 1;
    } else {
      tag = 2ULL * lo;
    }
    // keep rhs_id/strand of the band reachable after the sort: stash the
    // top 32 group bits in the low half of the tag word
    G[i] = (tag << 32) | (G[i] >> 32);
  }
  __syncthreads();
  BitonicSortPairs<THREADS>(G, P, npad);

  // 5. one thread per band: LIS + gap split + covered bases
  IdxT* MINI = wk.LB;  // re-used: per band (len + 1) entries at IB[b] + b
  for (uint32_t b = threadIdx.x; b < nb; b += THREADS) {
    const uint32_t jb = wk.IB[b], ie = wk.IE[b];
    const uint32_t len = ie - jb;
    uint32_t emitted = 0;
    uint32_t longest = 0;
    if (len >= cp.chain) {
      const uint64_t* Pb = P + jb;
      IdxT* minimal = MINI + jb + b;
      IdxT* pred = wk.PD + jb;
      const bool strand = G[jb] & 1;
      minimal[0] = 0;
      for (uint32_t t = 0; t < len; ++t) {
        const uint32_t cl = static_cast<uint32_t>(Pb[t] >> 32);
        const uint32_t cr = static_cast<uint32_t>(Pb[t]);
        uint32_t lo = 1, hi = longest;
        while (lo <= hi) {
          const uint32_t mid = lo + (hi - lo) / 2;
          const uint64_t tail = Pb[minimal[mid]];
          const uint32_t tl = static_cast<uint32_t>(tail >> 32);
          const uint32_t tr = static_cast<uint32_t>(tail);
          if (tl < cl && (strand ? tr < cr : tr > cr)) {
            lo = mid + 1;
          } else {
            hi = mid - 1;
          }
        }
        pred[t] = minimal[lo - 1];
        minimal[lo] = static_cast<IdxT>(t);
        longest = max(longest, lo);
      }
      if (longest >= cp.chain) {
        // unroll the chain into minimal[0 .. longest)
        uint32_t j = minimal[longest];
        for (uint32_t i = 0; i < longest; ++i) {
          const uint32_t pj = pred[j];
          minimal[longest - 1 - i] = static_cast<IdxT>(j);
          j = pj;
        }
      } else {
        longest = 0;
      }
    }
    // count the overlaps this band emits (walk repeated when writing)
    if (longest) {
      const uint64_t* Pb = P + jb;
      const IdxT* idx = MINI + jb + b;
      const bool strand = G[jb] & 1;
      for (uint32_t kk = 1, l = 0; kk <= longest; ++kk) {
        const uint32_t prev = static_cast<uint32_t>(Pb[idx[kk - 1]] >> 32);
        const uint32_t cur =
            kk < longest ? static_cast<uint32_t>(Pb[idx[kk]] >> 32) : 0xFFFFFFFFu;
        if (cur - prev > cp.gap) {
          if (kk - l >= cp.chain) {
            uint32_t lm = 0, lb_ = 0, le = 0, rm = 0, rb_ = 0, re = 0;
            for (uint32_t m = l; m < kk; ++m) {
              const uint64_t pp = Pb[idx[m]];
              const uint32_t lp = static_cast<uint32_t>(pp >> 32);
              if (lp > le) {
                lm += le - lb_;
                lb_ = lp;
              }
              le = lp + cp.k;
              uint32_t rp = static_cast<uint32_t>(pp);
              rp = strand ? rp : (1U << 31) - (rp + cp.k - 1);
              if (rp > re) {
                rm += re - rb_;
                rb_ = rp;
              }
              re = rp + cp.k;
            }
            lm += le - lb_;
            rm += re - rb_;
            if (min(lm, rm) >= cp.matches) ++emitted;
          }
          l = kk;
        }
      }
    }
    wk.CNT[b] = emitted;
    // stash the chain length where the writer finds it
    wk.IE[b] = static_cast<IdxT>(longest);
  }
  __syncthreads();

  // 6. place the overlaps of all bands in band order
  uint32_t carry = 0;
  __shared__ uint32_t sh_total;
  // first pass: total
  {
    uint32_t local = 0;
    for (uint32_t b = threadIdx.x; b < nb; b += THREADS) local += wk.CNT[b];
    uint32_t tot;
    BlockExclusiveSum<uint32_t, THREADS>(local, sm32, &tot);
    if (threadIdx.x == 0) {
      sh_total = tot;
      sh_base = tot ? atomicAdd(ovl_counter, static_cast<unsigned long long>(tot))
                    : 0ULL;
    }
    __syncthreads();
  }
  const uint32_t total = sh_total;
  const uint64_t base = sh_base;
  *out_base = base;
  if (total == 0 || base + total > ovl_cap) return total;

  for (uint32_t b0 = 0; b0 < nb; b0 += THREADS) {
    const uint32_t b = b0 + threadIdx.x;
    const uint32_t mine = b < nb ? wk.CNT[b] : 0;
    uint32_t tot;
    const uint32_t ex = BlockExclusiveSum<uint32_t, THREADS>(mine, sm32, &tot);
    if (mine) {
      rvn_overlap* dst = ovl_raw + base + carry + ex;
      const uint32_t jb = wk.IB[b];
      const uint32_t longest = wk.IE[b];
      const uint64_t* Pb = P + jb;
      const IdxT* idx = MINI + jb + b;
      const bool strand = G[jb] & 1;
      const uint32_t rhs_id = static_cast<uint32_t>(G[jb] & 0xFFFFFFFFu) >> 1;
      for (uint32_t kk = 1, l = 0; kk <= longest; ++kk) {
        const uint32_t prev = static_cast<uint32_t>(Pb[idx[kk - 1]] >> 32);
        const uint32_t cur =
            kk < longest ? static_cast<uint32_t>(Pb[idx[kk]] >> 32) : 0xFFFFFFFFu;
        if (cur - prev > cp.gap) {
          if (kk - l >= cp.chain) {
            uint32_t lm = 0, lb_ = 0, le = 0, rm = 0, rb_ = 0, re = 0;
            for (uint32_t m = l; m < kk; ++m) {
              const uint64_t pp = Pb[idx[m]];
              const uint32_t lp = static_cast<uint32_t>(pp >> 32);
              if (lp > le) {
                lm += le - lb_;
                lb_ = lp;
              }
              le = lp + cp.k;
              uint32_t rp = static_cast<uint32_t>(pp);
              rp = strand ? rp : (1U << 31) - (rp + cp.k - 1);
              if (rp > re) {
                rm += re - rb_;
                rb_ = rp;
              }
              re = rp + cp.k;
            }
            lm += le - lb_;
            rm += re - rb_;
            if (min(lm, rm) >= cp.matches) {
              const uint64_t pf = Pb[idx[l]], pl = Pb[idx[kk - 1]];
              rvn_overlap o;
              o.lhs_id = lhs_id;
              o.lhs_begin = static_cast<uint32_t>(pf >> 32);
              o.lhs_end = cp.k + static_cast<uint32_t>(pl >> 32);
              o.rhs_id = rhs_id;
              o.rhs_begin = strand ? static_cast<uint32_t>(pf)
                                   : static_cast<uint32_t>(pl);
              o.rhs_end = cp.k + (strand ? static_cast<uint32_t>(pl)
                                         : static_cast<uint32_t>(pf));
              o.score = min(lm, rm);
              o.strand = strand;
              *dst++ = o;
            }
          }
          l = kk;
        }
      }
    }
    carry += tot;
  }
  return total;
}

constexpr uint32_t kChainSmemCap = 8191;  // hits per read on the smem path
constexpr uint32_t kChainMaxGroup = 1024;  // larger (rhs, strand) groups -> generic path

// ---------------------------------------------------------------------------
// Fast path: one CTA per query read, everything in shared memory.
//
// Bands never span two (rhs_id, strand) pairs (group keys of different pairs
// differ by >= 2^30 > bandwidth), so the reference's global sort by group is
// not needed: hits are first split by pair with a shared-memory hash table
// (pairs with < 4 hits can never form a band and are dropped on the spot),
// then ONE THREAD PER PAIR orders its few dozen hits by (diagonal, positions)
// with an insertion sort, walks the reference's window loop, and for every
// closed band sorts by positions, runs the patience/LIS recurrence, splits at
// gaps and tests covered bases — all without a barrier. Overlaps are staged
// in the (by then dead) hash table and written out ordered by (pair, band,
// walk) = the reference's emission order.
// ---------------------------------------------------------------------------
struct FastLayout {
  uint32_t n, hs, ngmax;
  size_t p2, d2, hk, hc, gl, goff, gcnt, bytes;
};

__host__ __device__ inline FastLayout MakeFastLayout(uint32_t n) {
  FastLayout L;
  L.n = n;
  uint32_t hs = 8;
  while (hs < n + 1) hs <<= 1;
  L.hs = hs;
  L.ngmax = n / 4 + 1;
  size_t o = 0;
  L.p2 = o;   o += 8ULL * n;                 // positions, grouped
  L.d2 = o;   o += 4ULL * n;                 // diagonals, grouped (later LIS scratch)
  o = (o + 15) & ~size_t(15);                // staging is read as uint4
  L.hk = o;   o += 4ULL * hs;                // hash keys      } later: overlap staging
  L.hc = o;   o += 4ULL * hs;                // hash counters  }   (8*hs bytes)
  uint32_t gpad = 2;
  while (gpad < L.ngmax) gpad <<= 1;
  L.gl = o;   o += 8ULL * gpad;              // (gid << 32 | slot), later staging keys
  L.goff = o; o += 2ULL * L.ngmax;
  L.gcnt = o; o += 2ULL * L.ngmax;
  L.bytes = (o + 15) & ~size_t(15);
  return L;
}

__device__ __forceinline__ uint32_t HashGid(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// one band [jb, ie) of a pair: sort by positions, LIS, gap split, emit
__device__ __forceinline__ void FastBand(uint64_t* P, uint32_t* D, uint32_t jb,
                                         uint32_t ie, bool strand, uint32_t lhs_id,
                                         uint32_t rhs_id, const ChainParams& cp,
                                         uint32_t q, uint32_t* seq,
                                         rvn_overlap* stage, uint32_t* stage_key,
                                         uint32_t* stage_cnt) {
  const uint32_t len = ie - jb;
  if (len < cp.chain) return;
  uint64_t* Pb = P + jb;
  for (uint32_t a = 1; a < len; ++a) {  // insertion sort by positions
    const uint64_t p = Pb[a];
    uint32_t b = a;
    while (b > 0 && Pb[b - 1] > p) {
      Pb[b] = Pb[b - 1];
      --b;
    }
    Pb[b] = p;
  }
  // the band's diagonals are dead: their storage holds minimal[1..len] and
  // predecessor[0..len) as u16 (minimal[0] is always 0)
  uint16_t* mini = reinterpret_cast<uint16_t*>(D + jb);  // mini[x-1] = minimal[x]
  uint16_t* pred = mini + len;
  uint32_t longest = 0;
  for (uint32_t t = 0; t < len; ++t) {
    const uint32_t cl = static_cast<uint32_t>(Pb[t] >> 32);
    const uint32_t cr = static_cast<uint32_t>(Pb[t]);
    uint32_t lo = 1, hi = longest;
    while (lo <= hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      const uint64_t tail = Pb[mini[mid - 1]];
      const uint32_t tl = static_cast<uint32_t>(tail >> 32);
      const uint32_t tr = static_cast<uint32_t>(tail);
      if (tl < cl && (strand ? tr < cr : tr > cr)) {
        lo = mid + 1;
      } else {
        hi = mid - 1;
      }
    }
    pred[t] = lo > 1 ? mini[lo - 2] : 0;
    mini[lo - 1] = static_cast<uint16_t>(t);
    longest = max(longest, lo);
  }
  if (longest < cp.chain) return;
  {
    uint32_t j = mini[longest - 1];
    for (uint32_t i = 0; i < longest; ++i) {
      const uint32_t pj = pred[j];
      mini[longest - 1 - i] = static_cast<uint16_t>(j);
      j = pj;
    }
  }
  const uint16_t* idx = mini;
  for (uint32_t kk = 1, l = 0; kk <= longest; ++kk) {
    const uint32_t prev = static_cast<uint32_t>(Pb[idx[kk - 1]] >> 32);
    const uint32_t cur =
        kk < longest ? static_cast<uint32_t>(Pb[idx[kk]] >> 32) : 0xFFFFFFFFu;
    if (cur - prev > cp.gap) {
      if (kk - l >= cp.chain) {
        uint32_t lm = 0, lb_ = 0, le = 0, rm = 0, rb_ = 0, re = 0;
        for (uint32_t m = l; m < kk; ++m) {
          const uint64_t pp = Pb[idx[m]];
          const uint32_t lp = static_cast<uint32_t>(pp >> 32);
          if (lp > le) {
            lm += le - lb_;
            lb_ = lp;
          }
          le = lp + cp.k;
          uint32_t rp = static_cast<uint32_t>(pp);
          rp = strand ? rp : (1U << 31) - (rp + cp.k - 1);
          if (rp > re) {
            rm += re - rb_;
            rb_ = rp;
          }
          re = rp + cp.k;
        }
        lm += le - lb_;
        rm += re - rb_;
        if (min(lm, rm) >= cp.matches) {
          const uint64_t pf = Pb[idx[l]], pl = Pb[idx[kk - 1]];
          rvn_overlap o;
          o.lhs_id = lhs_id;
          o.lhs_begin = static_cast<uint32_t>(pf >> 32);
          o.lhs_end = cp.k + static_cast<uint32_t>(pl >> 32);
          o.rhs_id = rhs_id;
          o.rhs_begin = strand ? static_cast<uint32_t>(pf) : static_cast<uint32_t>(pl);
          o.rhs_end = cp.k + (strand ? static_cast<uint32_t>(pl)
                                     : static_cast<uint32_t>(pf));
          o.score = min(lm, rm);
          o.strand = strand;
          const uint32_t slot = atomicAdd(stage_cnt, 1u);
          stage[slot] = o;
          stage_key[slot] = (q << 16) | (*seq)++;
        }
      }
      l = kk;
    }
  }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ChainKernelFast(const uint64_t* __restrict__ h_grp,
                const uint64_t* __restrict__ h_pos,
                const uint64_t* __restrict__ read_hit_off,
                const uint32_t* __restrict__ lhs_ids,
                const uint32_t* __restrict__ read_list, ChainParams cp,
                rvn_overlap* __restrict__ ovl_raw,
                unsigned long long* __restrict__ ovl_counter, uint64_t ovl_cap,
                uint64_t* __restrict__ ovl_loc,
                uint32_t* __restrict__ fallback_list,
                unsigned int* __restrict__ fallback_cnt) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t sm32[34];
  __shared__ uint32_t sh_stage_cnt, sh_bail;
  __shared__ unsigned long long sh_base;
  const uint32_t r = read_list[blockIdx.x];
  const uint64_t hb = read_hit_off[r];
  const uint32_t n = static_cast<uint32_t>(read_hit_off[r + 1] - hb);
  const FastLayout L = MakeFastLayout(n);
  uint64_t* P2 = reinterpret_cast<uint64_t*>(smem + L.p2);
  uint32_t* D2 = reinterpret_cast<uint32_t*>(smem + L.d2);
  uint32_t* HK = reinterpret_cast<uint32_t*>(smem + L.hk);
  uint32_t* HC = reinterpret_cast<uint32_t*>(smem + L.hc);
  uint64_t* GL = reinterpret_cast<uint64_t*>(smem + L.gl);
  uint16_t* GOFF = reinterpret_cast<uint16_t*>(smem + L.goff);
  uint16_t* GCNT = reinterpret_cast<uint16_t*>(smem + L.gcnt);
  const uint32_t hmask = L.hs - 1;
  const uint64_t* hg = h_grp + hb;
  const uint64_t* hp = h_pos + hb;

  // ---- 1a. hash table of (rhs_id, strand) pairs with their hit counts ----
  for (uint32_t i = threadIdx.x; i < L.hs; i += THREADS) {
    HK[i] = 0xFFFFFFFFu;
    HC[i] = 0;
  }
  if (threadIdx.x == 0) {
    sh_stage_cnt = 0;
    sh_bail = 0;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
    const uint32_t gid = static_cast<uint32_t>(hg[i] >> 32);
    uint32_t s = HashGid(gid) & hmask;
    while (true) {
      const uint32_t prev = atomicCAS(&HK[s], 0xFFFFFFFFu, gid);
      if (prev == 0xFFFFFFFFu || prev == gid) break;
      s = (s + 1) & hmask;
    }
    atomicAdd(&HC[s], 1u);
  }
  __syncthreads();

  // ---- 1b. pairs with >= 4 hits: list + grouped offsets ----
  uint32_t carry = 0;  // low 16: pairs so far, high 16: hits so far
  for (uint32_t b = 0; b < L.hs; b += THREADS) {
    const uint32_t s = b + threadIdx.x;
    const uint32_t cnt = s < L.hs ? HC[s] : 0;
    const uint32_t keep = cnt >= 4;
    if (cnt > kChainMaxGroup) sh_bail = 1;
    uint32_t tot;
    const uint32_t ex = BlockExclusiveSum<uint32_t, THREADS>(
        keep ? ((cnt << 16) | 1u) : 0u, sm32, &tot);
    if (s < L.hs) {
      if (keep) {
        const uint32_t at = carry + ex;
        GL[at & 0xFFFF] = (static_cast<uint64_t>(HK[s]) << 32) | s;
        HC[s] = (cnt << 16) | (at >> 16);  // (count, offset)
      } else {
        HC[s] = 0xFFFFFFFFu;  // dropped
      }
    }
    carry += tot;
  }
  const uint32_t ng = carry & 0xFFFF;
  __syncthreads();
  if (sh_bail) {  // a very large pair: the generic kernel takes this read
    if (threadIdx.x == 0) fallback_list[atomicAdd(fallback_cnt, 1u)] = r;
    return;
  }
  if (ng == 0) {
    if (threadIdx.x == 0) ovl_loc[r] = 0;
    return;
  }

  // ---- 1c. pairs in ascending key order (= the reference's emission order) ----
  uint32_t gpad = 2;
  while (gpad < ng) gpad <<= 1;
  for (uint32_t i = ng + threadIdx.x; i < gpad; i += THREADS) GL[i] = ~0ULL;
  __syncthreads();
  for (uint32_t size = 2; size <= gpad; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < (gpad >> 1); t += THREADS) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const uint64_t a = GL[i], c2 = GL[j];
        if ((a > c2) == ((i & size) == 0)) {
          GL[i] = c2;
          GL[j] = a;
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t q = threadIdx.x; q < ng; q += THREADS) {
    const uint32_t s = static_cast<uint32_t>(GL[q]);
    const uint32_t v = HC[s];
    GOFF[q] = static_cast<uint16_t>(v & 0xFFFF);
    GCNT[q] = static_cast<uint16_t>(v >> 16);
    HC[s] = v & 0xFFFF;  // fill cursor
  }
  __syncthreads();

  // ---- 1d. scatter the hits of kept pairs (order inside a pair is free) ----
  for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
    const uint64_t g = hg[i];
    const uint32_t gid = static_cast<uint32_t>(g >> 32);
    uint32_t s = HashGid(gid) & hmask;
    while (HK[s] != gid) s = (s + 1) & hmask;
    if (HC[s] == 0xFFFFFFFFu) continue;
    const uint32_t at = atomicAdd(&HC[s], 1u);
    D2[at] = static_cast<uint32_t>(g);
    P2[at] = hp[i];
  }
  __syncthreads();

  // ---- 2. one thread per pair ----
  rvn_overlap* stage = reinterpret_cast<rvn_overlap*>(smem + L.hk);
  uint32_t* stage_key = reinterpret_cast<uint32_t*>(smem + L.gl);
  // GL is dead once every thread has its (gid, offset, count) in registers
  uint32_t my_gid[(kChainSmemCap / 4 + THREADS) / THREADS];
  {
    uint32_t t = 0;
    for (uint32_t q = threadIdx.x; q < ng; q += THREADS) my_gid[t++] = GL[q] >> 32;
  }
  __syncthreads();
  {
    uint32_t t = 0;
    for (uint32_t q = threadIdx.x; q < ng; q += THREADS, ++t) {
      const uint32_t gid = my_gid[t];
      const uint32_t off = GOFF[q], m = GCNT[q];
      uint32_t* D = D2 + off;
      uint64_t* P = P2 + off;
      for (uint32_t a = 1; a < m; ++a) {  // insertion sort by (diagonal, positions)
        const uint32_t d = D[a];
        const uint64_t p = P[a];
        uint32_t b = a;
        while (b > 0 && (D[b - 1] > d || (D[b - 1] == d && P[b - 1] > p))) {
          D[b] = D[b - 1];
          P[b] = P[b - 1];
          --b;
        }
        D[b] = d;
        P[b] = p;
      }
      // the reference's window loop; index m plays the stop dummy
      const bool strand = gid & 1;
      const uint32_t rhs_id = gid >> 1;
      uint32_t seq = 0;
      bool open = false;
      uint32_t ob = 0, oe = 0;
      for (uint32_t i = 1, j = 0; i <= m; ++i) {
        if (i == m || D[i] - D[j] > cp.bandwidth) {
          if (i - j >= 4) {
            if (open && oe > j) {
              oe = i;
            } else {
              if (open) {
                FastBand(P, D, ob, oe, strand, lhs_ids[r], rhs_id, cp, q, &seq,
                         stage, stage_key, &sh_stage_cnt);
              }
              ob = j;
              oe = i;
              open = true;
            }
          }
          ++j;
          while (j < i && (i == m || D[i] - D[j] > cp.bandwidth)) ++j;
        }
      }
      if (open) {
        FastBand(P, D, ob, oe, strand, lhs_ids[r], rhs_id, cp, q, &seq, stage,
                 stage_key, &sh_stage_cnt);
      }
    }
  }
  __syncthreads();

  // ---- 3. emission order, reserved slab, write ----
  const uint32_t total = sh_stage_cnt;
  if (threadIdx.x == 0) {
    sh_base = total ? atomicAdd(ovl_counter, static_cast<unsigned long long>(total))
                    : 0ULL;
    ovl_loc[r] = total ? (static_cast<uint64_t>(sh_base) << 24) | total : 0;
  }
  if (total == 0) return;
  uint64_t* order = P2;  // dead: (key << 32 | staging slot)
  uint32_t opad = 2;
  while (opad < total) opad <<= 1;
  for (uint32_t i = threadIdx.x; i < opad; i += THREADS) {
    order[i] = i < total ? (static_cast<uint64_t>(stage_key[i]) << 32) | i : ~0ULL;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= opad; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < (opad >> 1); t += THREADS) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const uint64_t a = order[i], c2 = order[j];
        if ((a > c2) == ((i & size) == 0)) {
          order[i] = c2;
          order[j] = a;
        }
      }
      __syncthreads();
    }
  }
  const uint64_t base = sh_base;
  if (base + total > ovl_cap) return;  // host reports the overflow
  const uint4* src = reinterpret_cast<const uint4*>(stage);
  uint4* dst = reinterpret_cast<uint4*>(ovl_raw + base);
  for (uint32_t i = threadIdx.x; i < total * 2; i += THREADS) {
    const uint32_t from = static_cast<uint32_t>(order[i >> 1]);
    dst[i] = src[from * 2 + (i & 1)];
  }
}

// global-memory path for reads with more hits than shared memory holds: one
// CTA per listed read, arrays in a scratch slab (u32 indices)
__global__ void __launch_bounds__(kThreads)
ChainKernelGlobal(const uint64_t* __restrict__ h_grp,
                  const uint64_t* __restrict__ h_pos,
                  const uint64_t* __restrict__ read_hit_off,
                  const uint32_t* __restrict__ lhs_ids,
                  const uint32_t* __restrict__ big_reads,
                  const uint64_t* __restrict__ slab64_off,
                  const uint64_t* __restrict__ slab32_off,
                  uint64_t* __restrict__ slab64, uint32_t* __restrict__ slab32,
                  ChainParams cp, rvn_overlap* __restrict__ ovl_raw,
                  unsigned long long* __restrict__ ovl_counter,
                  uint64_t ovl_cap, uint64_t* __restrict__ ovl_loc) {
  __shared__ uint32_t sm32[34];
  const uint32_t r = big_reads[blockIdx.x];
  const uint64_t hb = read_hit_off[r];
  const uint32_t n = static_cast<uint32_t>(read_hit_off[r + 1] - hb);
  uint32_t npad = 8;
  while (npad < n + 1) npad <<= 1;

  ChainWork<uint32_t> wk;
  wk.G = slab64 + slab64_off[blockIdx.x];
  wk.P = wk.G + npad;
  const uint32_t nbmax = n / 4 + 1;
  wk.LB = slab32 + slab32_off[blockIdx.x];
  wk.PD = wk.LB + (n + nbmax + 2);
  wk.IB = wk.PD + (n + 1);
  wk.IE = wk.IB + nbmax;
  wk.CNT = wk.IE + nbmax;

  for (uint32_t i = threadIdx.x; i < npad; i += kThreads) {
    wk.G[i] = i < n ? h_grp[hb + i] : ~0ULL;
    wk.P[i] = i < n ? h_pos[hb + i] : ~0ULL;
  }
  __syncthreads();
  uint64_t base = 0;
  const uint32_t total = ChainRead<uint32_t, kThreads>(
      wk, n, npad, lhs_ids[r], cp, sm32, ovl_raw, ovl_counter, ovl_cap,
      &base);
  if (threadIdx.x == 0) ovl_loc[r] = total ? (base << 24) | total : 0;
}

__global__ void OverlapCounts(const uint64_t* __restrict__ ovl_loc, uint64_t n,
                              uint32_t* __restrict__ cnt) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = static_cast<uint32_t>(ovl_loc[i] & 0xFFFFFF);
}

// move every read's overlaps from its reserved slab to query order
__global__ void ReorderOverlaps(const rvn_overlap* __restrict__ raw,
                                const uint64_t* __restrict__ ovl_loc,
                                const uint64_t* __restrict__ ovl_off,
                                uint32_t n_reads, rvn_overlap* __restrict__ out) {
  const uint32_t r = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (r >= n_reads) return;
  const uint64_t loc = ovl_loc[r];
  const uint32_t cnt = static_cast<uint32_t>(loc & 0xFFFFFF);
  const uint64_t src = loc >> 24, dst = ovl_off[r];
  const uint4* s = reinterpret_cast<const uint4*>(raw + src);
  uint4* d = reinterpret_cast<uint4*>(out + dst);
  for (uint32_t i = threadIdx.x & 31; i < cnt * 2; i += 32) d[i] = s[i];
}

}  // namespace

void MapRange(Ctx& c, uint32_t first, uint32_t last, bool avoid_equal,
              bool avoid_symmetric, bool minhash, bool want_filtered,
              bool fetch) {
  if (!c.i_valid) throw StateError("Map before Minimize");
  c.r_valid = false;
  const uint32_t nr = last - first;

  // ---- query records ----
  const uint64_t *qv, *qo, *d_read_off;
  const std::vector<uint64_t>* h_read_off;
  uint64_t off_base_read;  // index of `first` inside the offsets arrays
  if (minhash) {
    if (!(c.q_valid && c.q_first <= first && last <= c.q_last)) {
      EnsureMicromizers(c, first, last);
    }
    qv = c.q_val.get();
    qo = c.q_org.get();
    d_read_off = c.q_off.get();
    h_read_off = &c.h_q_off;
    off_base_read = first - c.q_first;
  } else {
    if (!(c.s_valid && c.s_first <= first && last <= c.s_last)) {
      EnsureSketch(c, first, last);
    }
    qv = c.s_val.get();
    qo = c.s_org.get();
    d_read_off = c.s_off.get();
    h_read_off = &c.h_s_off;
    off_base_read = first - c.s_first;
  }
  const uint64_t q_begin = (*h_read_off)[off_base_read];
  const uint64_t n_q = (*h_read_off)[off_base_read + nr] - q_begin;

  IndexView ix{c.i_val.get(), c.i_org.get(), c.i_bucket.get(), c.i_n,
               static_cast<int>(2 * c.prm.k) - c.i_bucket_bits, c.occurrence};

  // ---- probe + expand ----
  TimerBegin(c, "probe");
  uint32_t* cnt = c.m_cnt.reserve(n_q + 1);
  uint32_t* frst = c.m_first.reserve(n_q + 1);
  uint8_t* filt = c.m_filt.reserve(n_q + 1);
  uint64_t* hit_off = c.m_hit_off.reserve(n_q + 2);
  uint64_t n_hits = 0;
  if (n_q > 0) {
    ProbeKernel<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(
        ix, qv, qo, q_begin, n_q, avoid_equal, avoid_symmetric, cnt, frst, filt);
    RVN_LAUNCH_CHECK();
    ++c.launches;
    ExclusiveScanU32(c, cnt, hit_off, n_q);
    n_hits = ReadU64(c, hit_off + n_q);
  } else {
    RVN_CUDA(cudaMemsetAsync(hit_off, 0, sizeof(uint64_t), c.stream));
  }
  TimerEnd(c);
  TimerBegin(c, "expand");
  uint64_t* hg = c.h_grp.reserve(n_hits + 1);
  uint64_t* hp = c.h_pos.reserve(n_hits + 1);
  if (n_hits > 0) {
    ExpandKernel<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(
        ix, qv, qo, q_begin, n_q, avoid_equal, avoid_symmetric, cnt, frst,
        hit_off, hg, hp);
    RVN_LAUNCH_CHECK();
    ++c.launches;
  }
  // per-read hit ranges
  uint64_t* read_hit_off = c.m_read_hit_off.reserve(nr + 2ULL);
  GatherU64<<<CeilDiv(nr + 1ULL, kThreads), kThreads, 0, c.stream>>>(
      hit_off, d_read_off + off_base_read, q_begin, nr + 1ULL, read_hit_off);
  RVN_LAUNCH_CHECK();
  ++c.launches;
  std::vector<uint64_t> h_rho(nr + 1ULL);
  RVN_CUDA(cudaMemcpyAsync(h_rho.data(), read_hit_off,
                           (nr + 1ULL) * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, c.stream));
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  TimerEnd(c);

  // ---- filtered positions (stage 2 only) ----
  c.r_filt_off.reserve(nr + 2ULL);
  for (uint32_t i = 0; i <= nr; ++i) c.r_filt_off.get()[i] = 0;
  uint64_t n_filtered = 0;
  if (want_filtered && n_q > 0) {
    uint32_t* f32 = c.m_first.get();  // `first` is dead after ExpandKernel
    FilteredFlagsToU32<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(
        filt, n_q, f32);
    uint64_t* fpos = c.m_filt_off.reserve(n_q + 2);
    ExclusiveScanU32(c, f32, fpos, n_q);
    n_filtered = ReadU64(c, fpos + n_q);
    uint32_t* fout = c.m_filtered.reserve(n_filtered + 1);
    ScatterFiltered<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(
        filt, fpos, qo, q_begin, n_q, fout);
    RVN_LAUNCH_CHECK();
    c.launches += 2;
    // per-read offsets of the filtered list
    uint64_t* froff = c.m_ovl_off.reserve(nr + 2ULL);
    GatherU64<<<CeilDiv(nr + 1ULL, kThreads), kThreads, 0, c.stream>>>(
        fpos, d_read_off + off_base_read, q_begin, nr + 1ULL, froff);
    RVN_LAUNCH_CHECK();
    ++c.launches;
    uint32_t* hf = c.r_filtered.reserve(n_filtered + 1);
    RVN_CUDA(cudaMemcpyAsync(hf, fout, n_filtered * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaMemcpyAsync(c.r_filt_off.get(), froff,
                             (nr + 1ULL) * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaStreamSynchronize(c.stream));
  }

  if (c.keep_hits) {
    uint64_t* g = c.r_hit_grp.reserve(n_hits + 1);
    uint64_t* p = c.r_hit_pos.reserve(n_hits + 1);
    uint64_t* o = c.r_hit_off.reserve(nr + 2ULL);
    RVN_CUDA(cudaMemcpyAsync(g, hg, n_hits * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaMemcpyAsync(p, hp, n_hits * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, c.stream));
    for (uint32_t i = 0; i <= nr; ++i) o[i] = h_rho[i];
    RVN_CUDA(cudaStreamSynchronize(c.stream));
    c.r_n_hits = n_hits;
  }

  // ---- chain ----
  TimerBegin(c, "chain");
  ChainParams cp{c.prm.k, c.prm.bandwidth, c.prm.chain, c.prm.matches, c.prm.gap};
  const uint32_t* lhs_ids = c.d_ids.get() + first;
  const uint64_t ovl_cap = n_hits / std::max(1u, std::min(c.prm.chain, 4u)) + 16;
  rvn_overlap* raw = c.m_ovl_raw.reserve(ovl_cap);
  uint64_t* loc = c.m_ovl_loc.reserve(nr + 1ULL);
  uint64_t* counter = c.m_counter.reserve((1u << 16) + 8);
  RVN_CUDA(cudaMemsetAsync(counter, 0, sizeof(uint64_t), c.stream));
  RVN_CUDA(cudaMemsetAsync(loc, 0, (nr + 1ULL) * sizeof(uint64_t), c.stream));

  // size classes by hit count (n + 1 <= 256 << k), then the generic path
  constexpr int kClasses = 6;
  std::vector<uint32_t> cls[kClasses];
  std::vector<uint32_t> big;
  const bool fast_ok = c.prm.chain >= 4;  // staging capacity argument (n / 4)
  for (uint32_t i = 0; i < nr; ++i) {
    const uint64_t n = h_rho[i + 1] - h_rho[i];
    if (n < 4) continue;  // cannot form a band; ovl_loc stays 0
    if (n > kChainSmemCap || !fast_ok) {
      big.push_back(i);
      continue;
    }
    int k = 0;
    while ((256u << k) < n + 1) ++k;
    cls[k].push_back(i);
  }
  {
    size_t total = 0;
    for (auto& v : cls) total += v.size();
    uint32_t* d_list = c.m_first.reserve(std::max<size_t>(total, n_q) + 1);
    uint32_t* d_fb = c.m_fallback.reserve(nr + 4ULL);
    RVN_CUDA(cudaMemsetAsync(d_fb, 0, 2 * sizeof(uint32_t), c.stream));
    std::vector<uint32_t> flat;
    flat.reserve(total);
    for (int k = kClasses - 1; k >= 0; --k) {
      flat.insert(flat.end(), cls[k].begin(), cls[k].end());
    }
    if (total) {
      RVN_CUDA(cudaMemcpyAsync(d_list, flat.data(), total * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c.stream));
      RVN_CUDA(cudaStreamSynchronize(c.stream));  // flat goes out of scope
    }
    size_t off = 0;
    auto* ctr = reinterpret_cast<unsigned long long*>(counter);
    for (int k = kClasses - 1; k >= 0; --k) {  // largest class first
      const unsigned cnt = static_cast<unsigned>(cls[k].size());
      if (cnt == 0) continue;
      const size_t smem = MakeFastLayout((256u << k) - 1).bytes;
      const uint32_t* lst = d_list + off;
      off += cnt;
      if (k >= 3) {
        auto kern = ChainKernelFast<256, 2>;
        RVN_CUDA(cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(MakeFastLayout(kChainSmemCap).bytes)));
        kern<<<cnt, 256, smem, c.stream>>>(hg, hp, read_hit_off, lhs_ids, lst, cp,
                                           raw, ctr, ovl_cap, loc, d_fb + 2, d_fb);
      } else {
        auto kern = ChainKernelFast<128, 8>;
        RVN_CUDA(cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(MakeFastLayout(2047).bytes)));
        kern<<<cnt, 128, smem, c.stream>>>(hg, hp, read_hit_off, lhs_ids, lst, cp,
                                           raw, ctr, ovl_cap, loc, d_fb + 2, d_fb);
      }
      RVN_LAUNCH_CHECK();
      ++c.launches;
    }
    if (total) {
      // reads the fast kernel handed back (a pair with > kChainMaxGroup hits)
      std::vector<uint32_t> fb(1);
      RVN_CUDA(cudaMemcpyAsync(fb.data(), d_fb, sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c.stream));
      RVN_CUDA(cudaStreamSynchronize(c.stream));
      const uint32_t nfb = fb[0];
      if (nfb) {
        fb.resize(nfb);
        RVN_CUDA(cudaMemcpyAsync(fb.data(), d_fb + 2, nfb * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, c.stream));
        RVN_CUDA(cudaStreamSynchronize(c.stream));
        std::sort(fb.begin(), fb.end());
        big.insert(big.end(), fb.begin(), fb.end());
      }
    }
  }
  if (!big.empty()) {
    std::vector<uint64_t> off64(big.size() + 1, 0), off32(big.size() + 1, 0);
    for (size_t i = 0; i < big.size(); ++i) {
      const uint64_t n = h_rho[big[i] + 1] - h_rho[big[i]];
      if (n >= 0x7FFFFFFFULL) throw LimitError("a query has 2^31 or more hits");
      uint64_t npad = 8;
      while (npad < n + 1) npad <<= 1;
      const uint64_t nbmax = n / 4 + 1;
      off64[i + 1] = off64[i] + 2 * npad;
      off32[i + 1] = off32[i] + (n + nbmax + 2) + (n + 1) + 3 * nbmax + 4;
    }
    uint64_t* slab64 = c.m_scratch64.reserve(off64.back() + 2 * (big.size() + 1) + 8);
    uint32_t* slab32 = c.m_scratch32.reserve(off32.back() + big.size() + 8);
    // offsets and the read list ride at the tail of the slabs
    uint64_t* d_off64 = slab64 + off64.back();
    uint64_t* d_off32 = d_off64 + big.size() + 1;
    uint32_t* d_big = slab32 + off32.back();
    RVN_CUDA(cudaMemcpyAsync(d_off64, off64.data(), (big.size() + 1) * 8,
                             cudaMemcpyHostToDevice, c.stream));
    RVN_CUDA(cudaMemcpyAsync(d_off32, off32.data(), (big.size() + 1) * 8,
                             cudaMemcpyHostToDevice, c.stream));
    RVN_CUDA(cudaMemcpyAsync(d_big, big.data(), big.size() * 4,
                             cudaMemcpyHostToDevice, c.stream));
    ChainKernelGlobal<<<static_cast<unsigned>(big.size()), kThreads, 0,
                        c.stream>>>(
        hg, hp, read_hit_off, lhs_ids, d_big, d_off64, d_off32, slab64, slab32, cp,
        raw, reinterpret_cast<unsigned long long*>(counter), ovl_cap, loc);
    RVN_LAUNCH_CHECK();
    ++c.launches;
    RVN_CUDA(cudaStreamSynchronize(c.stream));  // host vectors go out of scope
  }

  // ---- query order ----
  uint32_t* ocnt = c.m_cnt.reserve(nr + 1ULL);
  uint64_t* ooff = c.m_ovl_off.reserve(nr + 2ULL);
  uint64_t n_ovl = 0;
  if (nr > 0) {
    OverlapCounts<<<CeilDiv(nr, kThreads), kThreads, 0, c.stream>>>(loc, nr, ocnt);
    RVN_LAUNCH_CHECK();
    ++c.launches;
    ExclusiveScanU32(c, ocnt, ooff, nr);
    n_ovl = ReadU64(c, ooff + nr);
  } else {
    RVN_CUDA(cudaMemsetAsync(ooff, 0, sizeof(uint64_t), c.stream));
  }
  if (n_ovl > ovl_cap) throw LimitError("overlap slab overflow");
  rvn_overlap* ordered = c.m_ovl.reserve(n_ovl + 1);
  if (n_ovl > 0) {
    ReorderOverlaps<<<CeilDiv(nr, kThreads / 32), kThreads, 0, c.stream>>>(
        raw, loc, ooff, nr, ordered);
    RVN_LAUNCH_CHECK();
    ++c.launches;
  }
  TimerEnd(c);

  // ---- results to the host ----
  if (fetch) {
    rvn_overlap* ho = c.r_ovl.reserve(n_ovl + 1);
    uint64_t* hoff = c.r_ovl_off.reserve(nr + 2ULL);
    RVN_CUDA(cudaMemcpyAsync(ho, ordered, n_ovl * sizeof(rvn_overlap),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaMemcpyAsync(hoff, ooff, (nr + 1ULL) * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaStreamSynchronize(c.stream));
  }
  c.r_n_ovl = n_ovl;
  c.m_hits = n_hits;
  c.m_first_read = first;
  c.m_last_read = last;
  c.r_valid = fetch;

  uint64_t qbases = 0;
  for (uint32_t r = first; r < last; ++r) qbases += c.h_len[r];
  c.stats.query_bases += qbases;
  c.stats.query_records += n_q;
  c.stats.hits += n_hits;
  c.stats.overlaps += n_ovl;
  (void)n_filtered;
}

}  // namespace rvn
