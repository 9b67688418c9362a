// ORACLE — TEST INFRASTRUCTURE ONLY.
// The edlib C surface (include/edlib.h) over the oracle's own plain dynamic
// programmes, for oracle/_ref: the reference's construct.cc / assemble.cc call
// edlibAlign(..., edlibDefaultAlignConfig()) for the identity filter
// (construct.cc:190-199,407-416; assemble.cc:271-277). Deliberately NOT the
// product's raven_b200/host/edlib.cc (bit-vector): the two implementations
// check each other (tests/test_oracle.py::test_product_edlib_equals_oracle).
#include <cstdlib>
#include <cstring>
#include <string>

#include "edlib.h"
#include "racon/polisher.hpp"

#define ORC_EXPORT __attribute__((visibility("default")))

extern "C" {

ORC_EXPORT EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                     const EdlibEqualityPair* eq, int n_eq) {
  EdlibAlignConfig c;
  c.k = k;
  c.mode = mode;
  c.task = task;
  c.additionalEqualities = eq;
  c.additionalEqualitiesLength = n_eq;
  return c;
}

ORC_EXPORT EdlibAlignConfig edlibDefaultAlignConfig(void) {
  return edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE, NULL, 0);
}

ORC_EXPORT EdlibAlignResult edlibAlign(const char* query, int queryLength, const char* target,
                            int targetLength, const EdlibAlignConfig config) {
  EdlibAlignResult r;
  std::memset(&r, 0, sizeof(r));
  r.status = EDLIB_STATUS_OK;
  r.editDistance = -1;
  if (config.mode != EDLIB_MODE_NW || queryLength < 0 || targetLength < 0) {
    r.status = EDLIB_STATUS_ERROR;
    return r;
  }
  const std::string q(query, queryLength), t(target, targetLength);
  std::string path;
  int d;
  if (config.task == EDLIB_TASK_PATH) {
    path = racon::GlobalAlignmentPath(q, t);
    d = 0;
    int i = 0, j = 0;
    for (char op : path) {
      if (op == 'M') {
        d += q[i] != t[j];
        ++i;
        ++j;
      } else if (op == 'I') {
        ++d;
        ++i;
      } else {
        ++d;
        ++j;
      }
    }
  } else {
    d = static_cast<int>(racon::GlobalDistance(q, t));
  }
  if (config.k >= 0 && d > config.k) return r;
  r.editDistance = d;
  r.numLocations = 1;
  r.endLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.startLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.endLocations[0] = targetLength - 1;
  r.startLocations[0] = 0;
  if (config.task == EDLIB_TASK_PATH) {
    r.alignmentLength = static_cast<int>(path.size());
    r.alignment = static_cast<unsigned char*>(std::malloc(path.size() + 1));
    int i = 0, j = 0;
    for (std::size_t x = 0; x < path.size(); ++x) {
      if (path[x] == 'M') {
        r.alignment[x] = q[i] == t[j] ? EDLIB_EDOP_MATCH : EDLIB_EDOP_MISMATCH;
        ++i;
        ++j;
      } else if (path[x] == 'I') {
        r.alignment[x] = EDLIB_EDOP_INSERT;
        ++i;
      } else {
        r.alignment[x] = EDLIB_EDOP_DELETE;
        ++j;
      }
    }
  }
  return r;
}

ORC_EXPORT void edlibFreeAlignResult(EdlibAlignResult result) {
  std::free(result.endLocations);
  std::free(result.startLocations);
  std::free(result.alignment);
}

}  // extern "C"
