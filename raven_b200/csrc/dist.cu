// raven_b200 — multi-GPU building blocks of the stage-1 overlap path.
//
// One context per rank; the collectives themselves are the caller's
// (raven_b200/distributed.py: torch.distributed over NCCL). Reads are sharded
// by id (contiguous ranges); the minimizer index is partitioned by KEY:
//   owner(value) = value mod n_parts
// (minimizers are minima of hashes, so their HIGH bits are skewed towards zero;
// the low bits stay uniform). Every key's postings live on exactly one rank,
// in the reference's order (records arrive in source-rank = read order and
// the build sort is stable).
//   1. SketchSplit   sketch own reads, split index / query records by owner
//        -> all-to-all of 16-byte minimizer records
//   2. BuildIndexFrom on the received records; IndexHistogram -> all-reduce ->
//        ONE global occurrence threshold per batch
//   3. HitsSplit     probe + expand the received queries, split the hits by
//        the owner of their lhs read
//        -> all-to-all of seed hits ("minimizer-bucket hits", north star)
//   4. ChainOwned    group received hits by read, chain (same kernels as one GPU)
//        -> all-gather of overlaps (32 B each, query order = rank order)
//   5. Stage1Finish  piles + gather/truncate with the reference's flush schedule
//        on the gathered list (cheap, replicated: every rank ends with the
//        complete, identical result).
#include <algorithm>

#include "engine.cuh"

namespace rvn {

namespace {

constexpr int kThreads = 256;

__global__ void OwnerFlags(const uint64_t* __restrict__ val, uint64_t n, uint32_t parts,
                           uint32_t p, uint32_t* __restrict__ flag) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = static_cast<uint32_t>(val[i] % parts) == p ? 1u : 0u;
}

__global__ void ScatterFlagged(const uint64_t* __restrict__ val,
                               const uint64_t* __restrict__ org,
                               const uint32_t* __restrict__ flag,
                               const uint64_t* __restrict__ pos, uint64_t n,
                               uint64_t base, uint64_t* __restrict__ out_val,
                               uint64_t* __restrict__ out_org) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  out_val[base + pos[i]] = val[i];
  out_org[base + pos[i]] = org[i];
}

struct IndexView2 {
  const uint64_t* val;
  const uint64_t* org;
  const uint32_t* bucket;
  uint64_t n;
  int shift;
  uint32_t occurrence;
};

__device__ __forceinline__ void Lookup2(const IndexView2& ix, uint64_t v, uint32_t* first,
                                        uint32_t* count) {
  const uint64_t b = v >> ix.shift;
  uint32_t lo = ix.bucket[b], hi = ix.bucket[b + 1];
  while (hi - lo > 8) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (ix.val[mid] < v) lo = mid + 1; else hi = mid;
  }
  const uint32_t end = ix.bucket[b + 1];
  while (lo < end && ix.val[lo] < v) ++lo;
  if (lo >= end || ix.val[lo] != v) {
    *first = 0;
    *count = 0;
    return;
  }
  *first = lo;
  if (ix.occurrence != 0xFFFFFFFFu && static_cast<uint64_t>(lo) + ix.occurrence < ix.n &&
      ix.val[static_cast<uint64_t>(lo) + ix.occurrence] == v) {
    *count = ix.occurrence + 1;
    return;
  }
  uint32_t n = 1;
  while (static_cast<uint64_t>(lo) + n < ix.n && ix.val[lo + n] == v) ++n;
  *count = n;
}

__device__ __forceinline__ bool Keep2(uint32_t lhs_id, uint64_t origin, bool ae, bool as) {
  const uint32_t rhs_id = static_cast<uint32_t>(origin >> 32);
  if (ae && lhs_id == rhs_id) return false;
  if (as && lhs_id > rhs_id) return false;
  return true;
}

// per query record: hits kept, first posting, owner of the lhs read
__global__ void ProbeOwned(IndexView2 ix, const uint64_t* __restrict__ q_val,
                           const uint64_t* __restrict__ q_org, uint64_t n_q, bool ae,
                           bool as, const uint32_t* __restrict__ bounds, uint32_t parts,
                           uint32_t* __restrict__ cnt, uint32_t* __restrict__ first,
                           uint8_t* __restrict__ dest) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n_q) return;
  const uint64_t v = q_val[i];
  const uint32_t lhs_id = static_cast<uint32_t>(q_org[i] >> 32);
  uint32_t f, n;
  Lookup2(ix, v, &f, &n);
  uint32_t kept = 0;
  if (n <= ix.occurrence) {
    for (uint32_t j = 0; j < n; ++j) kept += Keep2(lhs_id, ix.org[f + j], ae, as);
  }
  cnt[i] = kept;
  first[i] = f;
  uint32_t d = 0;
  while (d + 1 < parts && lhs_id >= bounds[d + 1]) ++d;
  dest[i] = static_cast<uint8_t>(d);
}

__global__ void MaskCounts(const uint32_t* __restrict__ cnt,
                           const uint8_t* __restrict__ dest, uint64_t n, uint32_t p,
                           uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dest[i] == p ? cnt[i] : 0u;
}

__global__ void ExpandOwned(IndexView2 ix, const uint64_t* __restrict__ q_val,
                            const uint64_t* __restrict__ q_org, uint64_t n_q, bool ae,
                            bool as, const uint32_t* __restrict__ cnt,
                            const uint32_t* __restrict__ first,
                            const uint8_t* __restrict__ dest, uint32_t p,
                            const uint64_t* __restrict__ off, uint64_t base,
                            uint64_t* __restrict__ h_grp, uint64_t* __restrict__ h_pos,
                            uint32_t* __restrict__ h_lhs) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n_q || dest[i] != p) return;
  uint32_t left = cnt[i];
  if (left == 0) return;
  const uint64_t v = q_val[i];
  const uint64_t lo = q_org[i];
  const uint32_t lhs_id = static_cast<uint32_t>(lo >> 32);
  const uint64_t lhs_pos = static_cast<uint32_t>(lo) >> 1;
  uint64_t dst = base + off[i];
  for (uint64_t j = first[i]; left > 0 && j < ix.n && ix.val[j] == v; ++j) {
    const uint64_t o = ix.org[j];
    if (!Keep2(lhs_id, o, ae, as)) continue;
    const uint64_t rhs_id = o >> 32;
    const uint64_t strand = (lo & 1) == (o & 1);
    const uint64_t rhs_pos = static_cast<uint32_t>(o) >> 1;
    const uint64_t diagonal =
        !strand ? rhs_pos + lhs_pos : rhs_pos - lhs_pos + (3ULL << 30);
    h_grp[dst] = (((rhs_id << 1) | strand) << 32) | diagonal;
    h_pos[dst] = (lhs_pos << 32) | rhs_pos;
    h_lhs[dst] = lhs_id;
    ++dst;
    --left;
  }
}

__global__ void CountByRead(const uint32_t* __restrict__ lhs, uint64_t n, uint32_t first,
                            uint32_t* __restrict__ cnt) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[lhs[i] - first], 1u);
}

__global__ void ScatterByRead(const uint64_t* __restrict__ grp,
                              const uint64_t* __restrict__ pos,
                              const uint32_t* __restrict__ lhs, uint64_t n, uint32_t first,
                              const uint64_t* __restrict__ off, uint32_t* __restrict__ cursor,
                              uint64_t* __restrict__ out_grp, uint64_t* __restrict__ out_pos) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = lhs[i] - first;
  const uint64_t at = off[r] + atomicAdd(&cursor[r], 1u);
  out_grp[at] = grp[i];
  out_pos[at] = pos[i];
}

__global__ void OverlapCountsPerRead(const uint64_t* __restrict__ off, uint32_t n,
                                     uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = static_cast<uint32_t>(off[i + 1] - off[i]);
}

}  // namespace

void DistSketchSplit(Ctx& c, uint32_t first, uint32_t last, int which, uint32_t parts,
                     const uint64_t** d_val, const uint64_t** d_org, uint64_t* counts) {
  if (parts == 0 || parts > 16) throw InvalidArgument("1..16 partitions");
  EnsureSketch(c, first, last);
  const uint64_t* sv = c.s_val.get();
  const uint64_t* so = c.s_org.get();
  uint64_t n = c.s_n;
  if (which == 1) {
    EnsureMicromizers(c, first, last);
    sv = c.q_val.get();
    so = c.q_org.get();
    n = c.q_n;
  }
  DevBuf<uint64_t>& ov = which == 1 ? c.ds_qsplit_val : c.ds_split_val;
  DevBuf<uint64_t>& oo = which == 1 ? c.ds_qsplit_org : c.ds_split_org;
  uint64_t* out_val = ov.reserve(n + 1);
  uint64_t* out_org = oo.reserve(n + 1);
  uint32_t* flag = c.m_cnt.reserve(n + 1);
  uint64_t* pos = c.m_hit_off.reserve(n + 2);
  TimerBegin(c, "dist_split");
  uint64_t base = 0;
  for (uint32_t p = 0; p < parts; ++p) {
    uint64_t cnt = 0;
    if (n) {
      OwnerFlags<<<CeilDiv(n, kThreads), kThreads, 0, c.stream>>>(sv, n, parts, p, flag);
      ExclusiveScanU32(c, flag, pos, n);
      cnt = ReadU64(c, pos + n);
      ScatterFlagged<<<CeilDiv(n, kThreads), kThreads, 0, c.stream>>>(sv, so, flag, pos, n, base,
                                                                     out_val, out_org);
      RVN_LAUNCH_CHECK();
      c.launches += 2;
    }
    counts[p] = cnt;
    base += cnt;
  }
  TimerEnd(c);
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  *d_val = out_val;
  *d_org = out_org;
}

void DistHitsSplit(Ctx& c, const uint64_t* d_qval, const uint64_t* d_qorg, uint64_t n_q,
                   bool ae, bool as, uint32_t parts, const uint32_t* h_bounds,
                   const uint64_t** d_grp, const uint64_t** d_pos, const uint32_t** d_lhs,
                   uint64_t* counts) {
  if (!c.i_valid) throw StateError("no index");
  if (parts == 0 || parts > 16) throw InvalidArgument("1..16 partitions");
  IndexView2 ix{c.i_val.get(), c.i_org.get(), c.i_bucket.get(), c.i_n,
                static_cast<int>(2 * c.prm.k) - c.i_bucket_bits, c.occurrence};
  uint32_t* d_bounds = c.m_bounds.reserve(parts + 2);
  RVN_CUDA(cudaMemcpyAsync(d_bounds, h_bounds, (parts + 1) * 4, cudaMemcpyHostToDevice, c.stream));
  uint32_t* cnt = c.m_cnt.reserve(n_q + 1);
  uint32_t* frst = c.m_first.reserve(n_q + 1);
  uint8_t* dest = c.m_filt.reserve(n_q + 1);
  uint32_t* masked = c.ds_masked.reserve(n_q + 1);
  uint64_t* off = c.m_hit_off.reserve(n_q + 2);
  TimerBegin(c, "dist_probe");
  if (n_q) {
    ProbeOwned<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(ix, d_qval, d_qorg, n_q, ae, as,
                                                                 d_bounds, parts, cnt, frst, dest);
    RVN_LAUNCH_CHECK();
    ++c.launches;
  }
  // totals per destination first (buffers are sized once)
  std::vector<uint64_t> tot(parts, 0);
  uint64_t n_hits = 0;
  for (uint32_t p = 0; p < parts && n_q; ++p) {
    MaskCounts<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(cnt, dest, n_q, p, masked);
    ExclusiveScanU32(c, masked, off, n_q);
    tot[p] = ReadU64(c, off + n_q);
    n_hits += tot[p];
  }
  TimerEnd(c);
  TimerBegin(c, "dist_expand");
  uint64_t* hg = c.h_grp.reserve(n_hits + 1);
  uint64_t* hp = c.h_pos.reserve(n_hits + 1);
  uint32_t* hl = c.ds_hit_lhs.reserve(n_hits + 1);
  uint64_t base = 0;
  for (uint32_t p = 0; p < parts && n_q; ++p) {
    if (tot[p]) {
      MaskCounts<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(cnt, dest, n_q, p, masked);
      ExclusiveScanU32(c, masked, off, n_q);
      ExpandOwned<<<CeilDiv(n_q, kThreads), kThreads, 0, c.stream>>>(
          ix, d_qval, d_qorg, n_q, ae, as, cnt, frst, dest, p, off, base, hg, hp, hl);
      RVN_LAUNCH_CHECK();
      c.launches += 2;
    }
    counts[p] = tot[p];
    base += tot[p];
  }
  TimerEnd(c);
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  c.stats.query_records += n_q;
  c.stats.hits += n_hits;
  *d_grp = hg;
  *d_pos = hp;
  *d_lhs = hl;
}

// hits of the owned reads [first, last) (any order) -> overlaps in query order
void DistChainOwned(Ctx& c, const uint64_t* d_grp, const uint64_t* d_pos,
                    const uint32_t* d_lhs, uint64_t n_hits, uint32_t first, uint32_t last,
                    const rvn_overlap** d_ovl, const uint32_t** d_ovl_cnt, uint64_t* n_ovl) {
  const uint32_t nr = last - first;
  TimerBegin(c, "dist_group");
  uint32_t* rcnt = c.ds_read_cnt.reserve(nr + 2ULL);
  uint32_t* cursor = c.ds_read_cursor.reserve(nr + 2ULL);
  uint64_t* roff = c.m_read_hit_off.reserve(nr + 2ULL);
  RVN_CUDA(cudaMemsetAsync(rcnt, 0, (nr + 1ULL) * 4, c.stream));
  RVN_CUDA(cudaMemsetAsync(cursor, 0, (nr + 1ULL) * 4, c.stream));
  uint64_t* gg = c.ds_grouped_grp.reserve(n_hits + 1);
  uint64_t* gp = c.ds_grouped_pos.reserve(n_hits + 1);
  if (n_hits) {
    CountByRead<<<CeilDiv(n_hits, kThreads), kThreads, 0, c.stream>>>(d_lhs, n_hits, first, rcnt);
  }
  ExclusiveScanU32(c, rcnt, roff, nr);
  if (n_hits) {
    ScatterByRead<<<CeilDiv(n_hits, kThreads), kThreads, 0, c.stream>>>(
        d_grp, d_pos, d_lhs, n_hits, first, roff, cursor, gg, gp);
    RVN_LAUNCH_CHECK();
    c.launches += 2;
  }
  std::vector<uint64_t> h_rho(nr + 1ULL);
  RVN_CUDA(cudaMemcpyAsync(h_rho.data(), roff, (nr + 1ULL) * 8, cudaMemcpyDeviceToHost, c.stream));
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  TimerEnd(c);
  const uint64_t n = ChainGroupedHits(c, gg, gp, roff, h_rho, first, nr, n_hits, n_hits);
  uint32_t* ocnt = c.ds_ovl_cnt.reserve(nr + 2ULL);
  if (nr) {
    OverlapCountsPerRead<<<CeilDiv(nr, kThreads), kThreads, 0, c.stream>>>(c.m_ovl_off.get(), nr, ocnt);
    RVN_LAUNCH_CHECK();
    ++c.launches;
  }
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  uint64_t qbases = 0;
  for (uint32_t r = first; r < last; ++r) qbases += c.h_len[r];
  c.stats.query_bases += qbases;
  c.stats.overlaps += n;
  *d_ovl = c.m_ovl.get();
  *d_ovl_cnt = ocnt;
  *n_ovl = n;
}

// Stage-1 tail, replicated on every rank: piles + gather/truncate with the
// reference's flush schedule (construct.cc:51-112). Begin once, Add once per
// index batch with the complete ordered overlap list of queries [0, n_query)
// (device) and its absolute per-read offsets (host, n_query + 1), then End.
void DistStage1Begin(Ctx& c) {
  if (!c.ids_identity) throw StateError("stage 1 needs read ids equal to their index");
  const uint32_t n = c.n_reads;
  c.st_valid = false;
  c.st_pile_off.assign(n + 1ULL, 0);
  for (uint32_t i = 0; i < n; ++i) c.st_pile_off[i + 1] = c.st_pile_off[i] + (c.h_len[i] >> 4);
  const uint64_t total_bins = c.st_pile_off[n];
  uint16_t* d_pile = c.p_data.reserve(total_bins + 1);
  uint64_t* d_poff = c.p_off.reserve(n + 1ULL);
  RVN_CUDA(cudaMemsetAsync(d_pile, 0, (total_bins + 1) * 2, c.stream));
  RVN_CUDA(cudaMemcpyAsync(d_poff, c.st_pile_off.data(), (n + 1ULL) * 8, cudaMemcpyHostToDevice,
                           c.stream));
  GatherReset(c);
  c.st_mapped = 0;
}

void DistStage1Add(Ctx& c, const rvn_overlap* d_ovl, const uint64_t* h_ovl_off,
                   uint32_t n_query, uint64_t kmax, uint64_t qb) {
  if (qb == 0) qb = 1ULL << 30;
  if (n_query > c.n_reads) throw InvalidArgument("query range out of bounds");
  const uint32_t n = c.n_reads;
  uint64_t* d_rel = c.ds_rel_off.reserve(n + 2ULL);
  std::vector<uint64_t> rel;
  uint64_t bases = 0;
  for (uint32_t k = 0, k0 = 0; k < n_query; ++k) {
    bases += c.h_len[k];
    if (k != n_query - 1 && bases < qb) continue;
    bases = 0;
    const uint64_t b = h_ovl_off[k0], e = h_ovl_off[k + 1];
    rel.assign(k + 2 - k0, 0);
    for (uint32_t r = k0; r <= k + 1; ++r) rel[r - k0] = h_ovl_off[r] - b;
    RVN_CUDA(cudaMemcpyAsync(d_rel, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, c.stream));
    RVN_CUDA(cudaStreamSynchronize(c.stream));
    PileAddLayersDevice(c, c.p_data.get(), c.p_off.get(), c.st_pile_off.data(), n, d_ovl + b,
                        e - b);
    GatherFlush(c, d_ovl + b, d_rel, e - b, k0, k + 1, kmax);
    k0 = k + 1;
  }
  c.st_mapped += h_ovl_off[n_query];
}

void DistStage1End(Ctx& c) {
  GatherFetch(c);
  const uint64_t total_bins = c.st_pile_off[c.n_reads];
  c.st_pile.resize(total_bins);
  RVN_CUDA(cudaMemcpyAsync(c.st_pile.data(), c.p_data.get(), total_bins * 2,
                           cudaMemcpyDeviceToHost, c.stream));
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  TimerCollect(c);
  c.stats.occurrence = c.occurrence;
  c.st_valid = true;
}

}  // namespace rvn
