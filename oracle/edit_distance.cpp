// ORACLE — TEST INFRASTRUCTURE ONLY.
// Global edit distance by the textbook two-row dynamic programme. This is the
// checker for edlibAlign(..., edlibDefaultAlignConfig()).editDistance as the
// reference uses it (RavenLib/src/construct.cc:190-197,407-414): unit costs,
// global (NW) mode. The optimum is unique, so any exact method is the oracle.
#include <algorithm>
#include <cstdint>
#include <vector>

extern "C" __attribute__((visibility("default"))) int orc_edit_distance(
    const char* a, int na, const char* b, int nb) {
  std::vector<int> prev(nb + 1), cur(nb + 1);
  for (int j = 0; j <= nb; ++j) prev[j] = j;
  for (int i = 1; i <= na; ++i) {
    cur[0] = i;
    for (int j = 1; j <= nb; ++j) {
      int sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
      cur[j] = std::min(sub, std::min(prev[j], cur[j - 1]) + 1);
    }
    prev.swap(cur);
  }
  return prev[nb];
}
