// raven-b200 host-side implementation of the `edlib` C surface the reference
// calls (include/edlib.h). Exact global (NW) edit distance by Myers/Hyyro
// bit-vector blocks; prefix (SHW) and infix (HW) modes come from the same
// recurrence with different boundary conditions (only NW is on the
// reference's path and only NW is served). The edit distance is unique, so
// any exact algorithm is bit-identical with upstream edlib on `editDistance`
// (reference use: construct.cc:190-199,407-416 - identity = 1 - ed/max(len)).
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

// Global distance between query (rows) and target (columns).
int GlobalDistance(const unsigned char* q, int m, const unsigned char* t,
                   int n) {
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
  int score = m;
  for (int j = 0; j < n; ++j) {
    const std::uint64_t* eq = peq.data() + static_cast<std::size_t>(t[j]) * blocks;
    int h = 1;  // D[0][j] - D[0][j-1] = +1 in global mode
    for (int b = 0; b < blocks; ++b) {
      h = Step(eq[b], h, b == blocks - 1 ? last_high : (1ULL << 63), pv[b],
               mv[b]);
    }
    score += h;
  }
  return score;
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

  if (config.task == EDLIB_TASK_PATH) {
    // alignment paths are produced by the polish engine's own aligner
    // (raven_b200/host/nw_path.cc); this entry point serves distances
    r.status = EDLIB_STATUS_ERROR;
    return r;
  }
  int d = GlobalDistance(q, queryLength, t, targetLength);
  if (config.k >= 0 && d > config.k) {
    return r;  // editDistance stays -1
  }
  r.editDistance = d;
  r.numLocations = 1;
  r.endLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.startLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.endLocations[0] = targetLength - 1;
  r.startLocations[0] = 0;

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
