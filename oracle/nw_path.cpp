// ORACLE — TEST INFRASTRUCTURE ONLY.
// Global (NW, unit cost) alignment PATH, the role of
// edlibAlign(..., {k=-1, EDLIB_MODE_NW, EDLIB_TASK_PATH}) inside racon
// (SURVEY.md App. A.4 step 3). Exact: Ukkonen band doubling until the
// distance fits the band. Among equally optimal paths the choice follows what
// we recall of edlib's traceback: walking back from the end, a cell first tries
// the cell above ('I', query base only), then the cell to the left ('D', target
// base only), then the diagonal. Of the six fixed orders this one (and its
// mirror) lands closest to the reference's end-to-end golden value (1141 vs
// 1137 on RavenTest.Assemble); the residue is unpinned.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "racon/polisher.hpp"

namespace racon {

std::string GlobalAlignmentPath(const std::string& q, const std::string& t) {
  const std::int64_t n = q.size(), m = t.size();
  if (n == 0) return std::string(m, 'D');
  if (m == 0) return std::string(n, 'I');
  const std::int64_t diff = std::llabs(n - m);
  // tie preference (experiment knob: ORC_NW_PREF=231 = I, D, diagonal)
  int pref[3] = {2, 3, 1};
  if (const char* e = std::getenv("ORC_NW_PREF")) {
    for (int z = 0; z < 3 && e[z]; ++z) pref[z] = e[z] - '0';
  }
  for (std::int64_t k = std::max<std::int64_t>(64, diff + 1);; k *= 2) {
    // band: diagonals j - i in [lo, hi]
    const std::int64_t lo = -k + std::min<std::int64_t>(0, m - n);
    const std::int64_t hi = k + std::max<std::int64_t>(0, m - n);
    const std::int64_t width = hi - lo + 1;
    const std::int32_t kInf = 1 << 29;
    std::vector<std::int32_t> prev(width, kInf), cur(width, kInf);
    std::vector<std::uint8_t> dir(static_cast<std::size_t>(n + 1) * width, 0);
    for (std::int64_t j = 0; j <= std::min<std::int64_t>(m, hi); ++j) {
      prev[j - lo] = static_cast<std::int32_t>(j);
      dir[j - lo] = 3;
    }
    for (std::int64_t i = 1; i <= n; ++i) {
      std::fill(cur.begin(), cur.end(), kInf);
      const std::int64_t jb = std::max<std::int64_t>(0, i + lo);
      const std::int64_t je = std::min<std::int64_t>(m, i + hi);
      for (std::int64_t j = jb; j <= je; ++j) {
        const std::int64_t c = j - i - lo;  // column inside the band of row i
        std::int32_t best = kInf;
        std::uint8_t d = 0;
        if (j == 0) {
          best = static_cast<std::int32_t>(i);
          d = 2;
        } else {
          // candidates: 1 diagonal, 2 up ('I'), 3 left ('D'); ties by `pref`
          std::int32_t v[4] = {kInf, prev[c] + (q[i - 1] != t[j - 1]),
                               c + 1 < width ? prev[c + 1] + 1 : kInf,
                               c - 1 >= 0 ? cur[c - 1] + 1 : kInf};
          for (int z = 0; z < 3; ++z) {
            const int which = pref[z];
            if (v[which] < best) { best = v[which]; d = which; }
          }
        }
        cur[c] = best;
        dir[static_cast<std::size_t>(i) * width + c] = d;
      }
      prev.swap(cur);
    }
    const std::int64_t cend = m - n - lo;
    const std::int32_t dist = (cend >= 0 && cend < width) ? prev[cend] : kInf;
    if (dist <= k || k > n + m) {
      std::string path;
      std::int64_t i = n, j = m;
      while (i > 0 || j > 0) {
        const std::uint8_t d = dir[static_cast<std::size_t>(i) * width + (j - i - lo)];
        if (d == 1) { path += 'M'; --i; --j; }
        else if (d == 2) { path += 'I'; --i; }
        else { path += 'D'; --j; }
      }
      std::reverse(path.begin(), path.end());
      return path;
    }
  }
}

// racon Overlap::find_breaking_points (cigar walk): for every window boundary
// of the target (multiples of w, and the overlap end) the first and the
// one-past-last aligned (target, query) pair inside that window
std::vector<std::pair<std::uint32_t, std::uint32_t>> BreakingPoints(
    const std::string& path, std::uint32_t q_begin, std::uint32_t t_begin,
    std::uint32_t t_end, std::uint32_t w) {
  std::vector<std::int64_t> window_ends;
  for (std::uint32_t i = 0; i < t_end; i += w) {
    if (i > t_begin) window_ends.emplace_back(static_cast<std::int64_t>(i) - 1);
  }
  window_ends.emplace_back(static_cast<std::int64_t>(t_end) - 1);

  std::vector<std::pair<std::uint32_t, std::uint32_t>> dst;
  std::size_t wi = 0;
  bool found_first = false;
  std::pair<std::uint32_t, std::uint32_t> first{0, 0}, last{0, 0};
  std::int64_t q_ptr = static_cast<std::int64_t>(q_begin) - 1;
  std::int64_t t_ptr = static_cast<std::int64_t>(t_begin) - 1;
  auto close_window = [&]() {
    if (wi < window_ends.size() && t_ptr == window_ends[wi]) {
      if (found_first) {
        dst.emplace_back(first);
        dst.emplace_back(last);
      }
      found_first = false;
      ++wi;
    }
  };
  for (char op : path) {
    if (op == 'M') {
      ++q_ptr;
      ++t_ptr;
      if (!found_first) {
        found_first = true;
        first = {static_cast<std::uint32_t>(t_ptr), static_cast<std::uint32_t>(q_ptr)};
      }
      last = {static_cast<std::uint32_t>(t_ptr + 1), static_cast<std::uint32_t>(q_ptr + 1)};
      close_window();
    } else if (op == 'I') {
      ++q_ptr;
    } else {  // 'D'
      ++t_ptr;
      close_window();
    }
  }
  return dst;
}

}  // namespace racon
