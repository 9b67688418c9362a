// raven-b200 host-side implementation of the `edlib` C surface the reference
// calls (include/edlib.h). Exact global (NW) edit distance by Myers/Hyyro
// bit-vector blocks (only the NW mode is on the reference's path and only NW
// is served). The edit distance is unique, so any exact algorithm is
// bit-identical with upstream edlib on `editDistance` (reference use:
// construct.cc:190-199,407-416 - identity = 1 - ed/max(len)). EDLIB_TASK_PATH
// (used by the polisher for the read-to-unitig alignment) walks back over the
// stored vertical deltas.
#include "edlib.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int kWord = 64;

inline int CeilDiv(int a, int b) { return (a + b - 1) / b; }

// One column step of one 64-row block. hin/hout in {-1, 0, +1}.
inline int Step(std::uint64_t eq, int hin, std::uint64_t high,
                std::uint64_t& pv, std::uint64_t& mv) {
  std::uint64_t xv = eq | mv;
  if (hin < 0) {
    eq |= 1ULL;
  }
  std::uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
  std::uint64_t ph = mv | ~(xh | pv);
  std::uint64_t mh = pv & xh;
  int hout = 0;
  if (ph & high) {
    hout = 1;
  } else if (mh & high) {
    hout = -1;
  }
  ph <<= 1;
  mh <<= 1;
  if (hin < 0) {
    mh |= 1ULL;
  } else if (hin > 0) {
    ph |= 1ULL;
  }
  pv = mh | ~(xv | ph);
  mv = ph & xv;
  return hout;
}

// Column-wise state kept for the traceback: vertical deltas of every block and
// the cell value at the bottom of every block, after each target column.
struct Trace {
  int blocks = 0;
  std::vector<std::uint64_t> pv, mv;  // [column][block]
  std::vector<std::int32_t> bottom;   // [column][block] = D[last row of block][column]
};

// Global distance between query (rows) and target (columns).
int GlobalDistance(const unsigned char* q, int m, const unsigned char* t,
                   int n, Trace* trace = nullptr) {
  if (m == 0) {
    return n;
  }
  if (n == 0) {
    return m;
  }
  const int blocks = CeilDiv(m, kWord);
  std::vector<std::uint64_t> peq(static_cast<std::size_t>(256) * blocks, 0);
  for (int i = 0; i < m; ++i) {
    peq[static_cast<std::size_t>(q[i]) * blocks + i / kWord] |=
        1ULL << (i % kWord);
  }
  const int last_bits = m - (blocks - 1) * kWord;
  const std::uint64_t last_high = 1ULL << (last_bits - 1);

  std::vector<std::uint64_t> pv(blocks, ~0ULL), mv(blocks, 0);
  std::vector<std::int32_t> bottom(blocks);
  for (int b = 0; b < blocks; ++b) bottom[b] = std::min(m, (b + 1) * kWord);
  if (trace) {
    trace->blocks = blocks;
    trace->pv.resize(static_cast<std::size_t>(n) * blocks);
    trace->mv.resize(static_cast<std::size_t>(n) * blocks);
    trace->bottom.resize(static_cast<std::size_t>(n) * blocks);
  }
  for (int j = 0; j < n; ++j) {
    const std::uint64_t* eq = peq.data() + static_cast<std::size_t>(t[j]) * blocks;
    int h = 1;  // D[0][j] - D[0][j-1] = +1 in global mode
    for (int b = 0; b < blocks; ++b) {
      h = Step(eq[b], h, b == blocks - 1 ? last_high : (1ULL << 63), pv[b],
               mv[b]);
      bottom[b] += h;
    }
    if (trace) {
      const std::size_t at = static_cast<std::size_t>(j) * blocks;
      std::copy(pv.begin(), pv.end(), trace->pv.begin() + at);
      std::copy(mv.begin(), mv.end(), trace->mv.begin() + at);
      std::copy(bottom.begin(), bottom.end(), trace->bottom.begin() + at);
    }
  }
  return bottom[blocks - 1];
}

}  // namespace

extern "C" {

EdlibAlignConfig edlibNewAlignConfig(
    int k, EdlibAlignMode mode, EdlibAlignTask task,
    const EdlibEqualityPair* additionalEqualities,
    int additionalEqualitiesLength) {
  EdlibAlignConfig c;
  c.k = k;
  c.mode = mode;
  c.task = task;
  c.additionalEqualities = additionalEqualities;
  c.additionalEqualitiesLength = additionalEqualitiesLength;
  return c;
}

EdlibAlignConfig edlibDefaultAlignConfig(void) {
  return edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE, NULL, 0);
}

EdlibAlignResult edlibAlign(const char* query, int queryLength,
                            const char* target, int targetLength,
                            const EdlibAlignConfig config) {
  EdlibAlignResult r;
  std::memset(&r, 0, sizeof(r));
  r.status = EDLIB_STATUS_OK;
  r.editDistance = -1;
  r.alphabetLength = 0;
  if (config.mode != EDLIB_MODE_NW || queryLength < 0 || targetLength < 0) {
    // only the global mode is on the reference's path
    r.status = EDLIB_STATUS_ERROR;
    return r;
  }
  const unsigned char* q = reinterpret_cast<const unsigned char*>(query);
  const unsigned char* t = reinterpret_cast<const unsigned char*>(target);
  {
    bool seen[256] = {};
    for (int i = 0; i < queryLength; ++i) seen[q[i]] = true;
    for (int i = 0; i < targetLength; ++i) seen[t[i]] = true;
    for (bool s : seen) r.alphabetLength += s;
  }

  const bool want_path = config.task == EDLIB_TASK_PATH;
  Trace trace;
  int d = GlobalDistance(q, queryLength, t, targetLength, want_path ? &trace : nullptr);
  if (config.k >= 0 && d > config.k) {
    return r;  // editDistance stays -1
  }
  r.editDistance = d;
  r.numLocations = 1;
  r.endLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.startLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.endLocations[0] = targetLength - 1;
  r.startLocations[0] = 0;

  if (want_path) {
    // Walk back from (m, n). Among equally good moves: up (a query symbol
    // alone, EDLIB_EDOP_INSERT), then left (a target symbol alone,
    // EDLIB_EDOP_DELETE), then the diagonal - the order we recall of edlib's
    // own traceback (upstream's choice is not pinned by the reference tree).
    const int m = queryLength, n = targetLength;
    auto cell = [&](int i, int j) -> int {  // D[i][j], i rows of query, j columns
      if (j == 0) return i;
      if (i == 0) return j;
      const int b = (i - 1) / kWord;
      const std::size_t at = static_cast<std::size_t>(j - 1) * trace.blocks + b;
      const int last_row = std::min(m, (b + 1) * kWord);  // 1-based row of the block bottom
      // rows i+1 .. last_row lie below cell i inside the block
      const int below = last_row - i;
      int v = trace.bottom[at];
      if (below > 0) {
        const int lo_bit = (i - 1) % kWord + 1;  // first bit below row i
        const std::uint64_t mask = (below >= 64 ? ~0ULL : ((1ULL << below) - 1)) << lo_bit;
        v -= __builtin_popcountll(trace.pv[at] & mask);
        v += __builtin_popcountll(trace.mv[at] & mask);
      }
      return v;
    };
    std::vector<unsigned char> ops;
    ops.reserve(static_cast<std::size_t>(m) + n);
    int i = m, j = n, cur = d;
    while (i > 0 || j > 0) {
      if (i > 0 && cell(i - 1, j) + 1 == cur) {
        ops.push_back(EDLIB_EDOP_INSERT);
        --i;
        --cur;
      } else if (j > 0 && cell(i, j - 1) + 1 == cur) {
        ops.push_back(EDLIB_EDOP_DELETE);
        --j;
        --cur;
      } else {
        const bool eq = q[i - 1] == t[j - 1];
        ops.push_back(eq ? EDLIB_EDOP_MATCH : EDLIB_EDOP_MISMATCH);
        --i;
        --j;
        if (!eq) --cur;
      }
    }
    std::reverse(ops.begin(), ops.end());
    r.alignmentLength = static_cast<int>(ops.size());
    r.alignment = static_cast<unsigned char*>(std::malloc(ops.size() + 1));
    std::memcpy(r.alignment, ops.data(), ops.size());
  }
  return r;
}

void edlibFreeAlignResult(EdlibAlignResult result) {
  std::free(result.endLocations);
  std::free(result.startLocations);
  std::free(result.alignment);
}

char* edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength,
                            EdlibCigarFormat cigarFormat) {
  std::string out;
  static const char ext[] = {'=', 'I', 'D', 'X'};
  static const char stdc[] = {'M', 'I', 'D', 'M'};
  const char* tbl = cigarFormat == EDLIB_CIGAR_EXTENDED ? ext : stdc;
  int run = 0;
  char last = 0;
  for (int i = 0; i <= alignmentLength; ++i) {
    char c = i < alignmentLength ? tbl[alignment[i] & 3] : 0;
    if (i == alignmentLength || (run > 0 && c != last)) {
      out += std::to_string(run);
      out += last;
      run = 0;
    }
    last = c;
    ++run;
  }
  char* s = static_cast<char*>(std::malloc(out.size() + 1));
  std::memcpy(s, out.c_str(), out.size() + 1);
  return s;
}

}  // extern "C"
