/* raven-b200: drop-in C header for the un-vendored dependency `edlib`.
 * Surface used by the reference: edlibAlign / edlibDefaultAlignConfig /
 * edlibFreeAlignResult / EdlibAlignResult.{status,editDistance} /
 * EDLIB_STATUS_OK (RavenLib/src/construct.cc:190-199,407-416;
 * assemble.cc:271-277; graph_repr.cc:250-254,361-365; raven_test.cpp:39-42).
 * Host implementation: raven_b200/host/edlib.cc (Myers bit-vector, exact).
 * The batched GPU path is rvn_edit_distance_batch (include/raven_b200.h). */
#ifndef EDLIB_H
#define EDLIB_H

#ifdef __cplusplus
extern "C" {
#endif

#define EDLIB_STATUS_OK 0
#define EDLIB_STATUS_ERROR 1

typedef enum { EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW } EdlibAlignMode;
typedef enum { EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH } EdlibAlignTask;
typedef enum { EDLIB_CIGAR_STANDARD, EDLIB_CIGAR_EXTENDED } EdlibCigarFormat;

#define EDLIB_EDOP_MATCH 0
#define EDLIB_EDOP_INSERT 1
#define EDLIB_EDOP_DELETE 2
#define EDLIB_EDOP_MISMATCH 3

typedef struct {
  char first;
  char second;
} EdlibEqualityPair;

typedef struct {
  int k; /* -1: no bound */
  EdlibAlignMode mode;
  EdlibAlignTask task;
  const EdlibEqualityPair* additionalEqualities;
  int additionalEqualitiesLength;
} EdlibAlignConfig;

typedef struct {
  int status;
  int editDistance; /* -1 if larger than k */
  int* endLocations;
  int* startLocations;
  int numLocations;
  unsigned char* alignment; /* EDLIB_EDOP_* per column, TASK_PATH only */
  int alignmentLength;
  int alphabetLength;
} EdlibAlignResult;

EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode,
                                     EdlibAlignTask task,
                                     const EdlibEqualityPair* additionalEqualities,
                                     int additionalEqualitiesLength);
EdlibAlignConfig edlibDefaultAlignConfig(void);
EdlibAlignResult edlibAlign(const char* query, int queryLength,
                            const char* target, int targetLength,
                            const EdlibAlignConfig config);
void edlibFreeAlignResult(EdlibAlignResult result);
char* edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength,
                            EdlibCigarFormat cigarFormat);

#ifdef __cplusplus
}
#endif

#endif /* EDLIB_H */
