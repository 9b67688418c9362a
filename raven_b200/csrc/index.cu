// raven_b200 — minimizer index build and frequency filter on sm_100a.
//
// Replaces the index half of ram::MinimizerEngine::Minimize(first,last,
// minhash) and ram::MinimizerEngine::Filter (un-vendored; call sites
// RavenLib/src/construct.cc:42-44,363,372).
//
// The reference keeps 2^14 hash buckets, radix-sorts each bucket by value and
// fills an unordered_map per bucket. On B200 the whole batch is ONE stable
// radix sort of the (value, origin) records by value — the input is already
// in (read, position) order, so equal values keep exactly the reference's
// posting order — plus a direct-address bucket table over the top bits of
// the (uniformly mixed) value: a probe is one table read and one short scan
// of a sorted run, no hashing, no pointer chasing.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>

#include "engine.cuh"

namespace rvn {

namespace {

constexpr int kThreads = 256;

// bucket[b] = index of the first record whose (value >> shift) >= b
__global__ void BuildBucketTable(const uint64_t* __restrict__ val, uint64_t n,
                                 int shift, uint32_t n_buckets,
                                 uint32_t* __restrict__ bucket) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint64_t lo = i == 0 ? 0 : (val[i - 1] >> shift) + 1;
  const uint64_t hi = i == n ? n_buckets : (val[i] >> shift);
  // record i is the first one of buckets (prev_bucket, this_bucket]
  for (uint64_t b = lo; b <= hi; ++b) {
    bucket[b] = static_cast<uint32_t>(i);
  }
}

// flag the first record of every run of equal values
__global__ void FlagRunStarts(const uint64_t* __restrict__ val, uint64_t n,
                              uint32_t* __restrict__ flag) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || val[i] != val[i - 1]) ? 1u : 0u;
}

__global__ void ScatterRunStarts(const uint32_t* __restrict__ flag,
                                 const uint64_t* __restrict__ pos, uint64_t n,
                                 uint32_t* __restrict__ run_start) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) run_start[pos[i]] = static_cast<uint32_t>(i);
}

constexpr uint32_t kHistBins = 1u << 16;

// histogram of run lengths; lengths >= kHistBins-1 land in the last bin.
// Nearly every run has length 1..3, so short lengths are counted in a
// shared-memory histogram per CTA (one global atomic per bin per CTA).
constexpr uint32_t kSmemBins = 1024;
__global__ void __launch_bounds__(kThreads)
RunLengthHistogram(const uint32_t* __restrict__ run_start, uint64_t n_keys,
                   uint64_t n, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[kSmemBins];
  for (uint32_t i = threadIdx.x; i < kSmemBins; i += kThreads) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kThreads;
  for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
       j < n_keys; j += stride) {
    const uint64_t end = j + 1 < n_keys ? run_start[j + 1] : n;
    uint64_t len = end - run_start[j];
    if (len > kHistBins - 1) len = kHistBins - 1;
    if (len < kSmemBins) {
      atomicAdd(&sh[len], 1u);
    } else {
      atomicAdd(&hist[len], 1ULL);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kSmemBins; i += kThreads) {
    if (sh[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(sh[i]));
  }
}

__global__ void CollectLongRuns(const uint32_t* __restrict__ run_start,
                                uint64_t n_keys, uint64_t n,
                                unsigned long long* __restrict__ counter,
                                uint32_t* __restrict__ out) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_keys) return;
  const uint64_t end = j + 1 < n_keys ? run_start[j + 1] : n;
  const uint64_t len = end - run_start[j];
  if (len >= kHistBins - 1) {
    out[atomicAdd(counter, 1ULL)] = static_cast<uint32_t>(len);
  }
}

}  // namespace

void BuildIndex(Ctx& c, uint32_t first, uint32_t last, bool minhash) {
  c.i_valid = false;
  c.occurrence = 0xFFFFFFFFu;
  EnsureSketch(c, first, last);
  const uint64_t* src_val = c.s_val.get();
  const uint64_t* src_org = c.s_org.get();
  uint64_t n = c.s_n;
  if (minhash) {
    EnsureMicromizers(c, first, last);
    src_val = c.q_val.get();
    src_org = c.q_org.get();
    n = c.q_n;
  }
  uint64_t bases = 0;
  for (uint32_t r = first; r < last; ++r) bases += c.h_len[r];
  c.i_first = first;
  c.i_last = last;
  BuildIndexFrom(c, src_val, src_org, n, bases);
}

void BuildIndexFrom(Ctx& c, const uint64_t* src_val, const uint64_t* src_org, uint64_t n,
                    uint64_t index_bases) {
  c.i_valid = false;
  c.occurrence = 0xFFFFFFFFu;
  if (n >= 0xFFFFFFFFULL) {
    throw LimitError("index batch holds 2^32 or more minimizers");
  }
  c.i_n = n;
  c.i_keys = 0;

  TimerBegin(c, "index_sort");
  uint64_t* kv = c.i_val.reserve(n + 1);
  uint64_t* ko = c.i_org.reserve(n + 1);
  uint64_t* kv_alt = c.i_val_alt.reserve(n + 1);
  uint64_t* ko_alt = c.i_org_alt.reserve(n + 1);
  if (n > 0) {
    // stable LSD radix sort on the 2k value bits; the sketch arrays stay
    // untouched (they still serve the queries of this batch)
    RVN_CUDA(cudaMemcpyAsync(kv_alt, src_val, n * sizeof(uint64_t),
                             cudaMemcpyDeviceToDevice, c.stream));
    RVN_CUDA(cudaMemcpyAsync(ko_alt, src_org, n * sizeof(uint64_t),
                             cudaMemcpyDeviceToDevice, c.stream));
    cub::DoubleBuffer<uint64_t> keys(kv_alt, kv);
    cub::DoubleBuffer<uint64_t> vals(ko_alt, ko);
    size_t tmp_bytes = 0;
    RVN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, vals, n, 0,
                                             static_cast<int>(2 * c.prm.k),
                                             c.stream));
    void* tmp = c.sort_tmp.reserve(tmp_bytes + 16);
    RVN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, vals, n, 0,
                                             static_cast<int>(2 * c.prm.k),
                                             c.stream));
    c.launches += 2 * ((2 * c.prm.k + 7) / 8) + 1;
    if (keys.Current() != kv) {
      std::swap(c.i_val.p, c.i_val_alt.p);
      std::swap(c.i_val.cap, c.i_val_alt.cap);
      std::swap(c.i_org.p, c.i_org_alt.p);
      std::swap(c.i_org.cap, c.i_org_alt.cap);
      kv = c.i_val.get();
      ko = c.i_org.get();
    }
  }
  TimerEnd(c);

  TimerBegin(c, "index_table");
  // bucket table over the top bits of the value
  int bits = 8;
  while (bits < 28 && (1ULL << (bits + 1)) <= n) ++bits;
  bits = std::min<int>(bits, 2 * c.prm.k);
  c.i_bucket_bits = bits;
  const int shift = static_cast<int>(2 * c.prm.k) - bits;
  const uint32_t n_buckets = 1u << bits;
  uint32_t* bucket = c.i_bucket.reserve(n_buckets + 2ULL);
  BuildBucketTable<<<CeilDiv(n + 1, kThreads), kThreads, 0, c.stream>>>(
      kv, n, shift, n_buckets, bucket);
  RVN_LAUNCH_CHECK();
  ++c.launches;

  // distinct keys: run starts, compacted
  if (n > 0) {
    uint32_t* flag = c.m_cnt.reserve(n);
    uint64_t* pos = c.m_hit_off.reserve(n + 1);
    FlagRunStarts<<<CeilDiv(n, kThreads), kThreads, 0, c.stream>>>(kv, n, flag);
    RVN_LAUNCH_CHECK();
    ExclusiveScanU32(c, flag, pos, n);
    c.i_keys = ReadU64(c, pos + n);
    uint32_t* rs = c.i_run_start.reserve(c.i_keys + 1);
    ScatterRunStarts<<<CeilDiv(n, kThreads), kThreads, 0, c.stream>>>(flag, pos,
                                                                      n, rs);
    RVN_LAUNCH_CHECK();
    c.launches += 2;
  }
  TimerEnd(c);

  c.stats.index_bases = index_bases;
  c.stats.index_records = n;
  c.stats.index_keys = c.i_keys;
  c.i_valid = true;
}

uint64_t* IndexHistogram(Ctx& c) {
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
  uint64_t* hist = c.m_counter.reserve(kHistBins + 8);
  RVN_CUDA(cudaMemsetAsync(hist, 0, (kHistBins + 8) * sizeof(uint64_t), c.stream));
  if (c.i_keys) {
    RunLengthHistogram<<<std::min<unsigned>(CeilDiv(c.i_keys, kThreads), 148 * 16),
                         kThreads, 0, c.stream>>>(
        c.i_run_start.get(), c.i_keys, c.i_n,
        reinterpret_cast<unsigned long long*>(hist));
    RVN_LAUNCH_CHECK();
    ++c.launches;
  }
  return hist;
}

// same arithmetic as the reference engine: index = (1 - f) * #keys, truncated
// towards zero; the value at that ascending rank, plus one
uint32_t ThresholdFromHistogram(Ctx& c, const uint64_t* h, uint64_t n_keys,
                                double frequency, bool* needs_long_runs) {
  (void)c;
  *needs_long_runs = false;
  std::size_t rank = static_cast<std::size_t>((1 - frequency) * static_cast<double>(n_keys));
  if (rank >= n_keys) rank = n_keys - 1;
  uint64_t cum = 0;
  for (uint32_t len = 0; len + 1 < kHistBins; ++len) {
    cum += h[len];
    if (cum > rank) return len + 1;
  }
  *needs_long_runs = true;  // the rank falls among runs of >= 65535 postings
  return 0;
}

// occurrence_ = (run length at ascending rank (1-f)*#keys) + 1
uint32_t FilterIndex(Ctx& c, double frequency) {
  if (!(0 <= frequency && frequency <= 1)) {
    throw InvalidArgument(
        "[ram::MinimizerEngine::Filter] error: invalid frequency");
  }
  if (!c.i_valid) throw StateError("Filter before Minimize");
  if (frequency == 0 || c.i_keys == 0) {
    c.occurrence = 0xFFFFFFFFu;
    return c.occurrence;
  }
  TimerBegin(c, "filter");
  uint64_t* hist = IndexHistogram(c);
  TimerEnd(c);
  std::vector<uint64_t> h(kHistBins);
  RVN_CUDA(cudaMemcpyAsync(h.data(), hist, kHistBins * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, c.stream));
  RVN_CUDA(cudaStreamSynchronize(c.stream));
  bool long_runs = false;
  uint32_t occ = ThresholdFromHistogram(c, h.data(), c.i_keys, frequency, &long_runs);
  if (long_runs) {
    uint64_t cum = 0;
    for (uint32_t len = 0; len + 1 < kHistBins; ++len) cum += h[len];
    std::size_t rank = static_cast<std::size_t>((1 - frequency) * static_cast<double>(c.i_keys));
    if (rank >= c.i_keys) rank = c.i_keys - 1;
    const uint64_t n_long = h[kHistBins - 1];
    uint32_t* out = c.m_cnt.reserve(n_long + 1);
    RVN_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t), c.stream));
    CollectLongRuns<<<CeilDiv(c.i_keys, kThreads), kThreads, 0, c.stream>>>(
        c.i_run_start.get(), c.i_keys, c.i_n,
        reinterpret_cast<unsigned long long*>(hist), out);
    RVN_LAUNCH_CHECK();
    ++c.launches;
    std::vector<uint32_t> lens(n_long);
    RVN_CUDA(cudaMemcpyAsync(lens.data(), out, n_long * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, c.stream));
    RVN_CUDA(cudaStreamSynchronize(c.stream));
    std::sort(lens.begin(), lens.end());
    occ = lens[rank - cum] + 1;
  }
  c.occurrence = occ;
  c.stats.occurrence = c.occurrence;
  return c.occurrence;
}

}  // namespace rvn
