/* raven_b200 — C ABI of the B200-native overlap engine.
 *
 * This is the drop-in boundary for the reference's overlap hot path. The
 * reference has no FFI layer of its own: RavenLib calls C++ classes of
 * un-vendored dependencies directly. Every entry point below replaces one of
 * those call sites (file:line are into lbcb-sci/raven @ v1.8.3):
 *
 *   rvn_engine_configure   ram::MinimizerEngine ctor   RavenLib/src/construct.cc:661-662
 *   rvn_reads_upload       biosoup::NucleicAcid fields RavenLib/include/raven/graph/graph.h:13-18
 *   rvn_minimize           MinimizerEngine::Minimize   RavenLib/src/construct.cc:42-43,363
 *   rvn_filter             MinimizerEngine::Filter     RavenLib/src/construct.cc:44,372
 *   rvn_map / rvn_map_external
 *                          MinimizerEngine::Map        RavenLib/src/construct.cc:59-64,377-381
 *   rvn_pile_add_layers    raven::Pile::AddLayers      RavenLib/src/pile.cc:33-62
 *   rvn_kmer_complexity    raven::Pile::AddKmers       RavenLib/src/pile.cc:64-120
 *   rvn_poa_batch          racon::Polisher::Polish (window consensus: racon::Window +
 *                          spoa)                       RavenLib/src/polish.cc:43-51
 *   rvn_edit_distance_batch
 *                          edlibAlign(lhs, rhs, edlibDefaultAlignConfig()) of the
 *                          identity filter      RavenLib/src/construct.cc:176-199,
 *                                                                        :393-416
 *   rvn_stage1_pile_regions
 *                          raven::Pile::FindValidRegion + FindMedian of
 *                          raven::TrimAndAnnotatePiles RavenLib/src/construct.cc:123-152
 *   rvn_align_breaking_points
 *                          edlibAlign(read, unitig, NW, EDLIB_TASK_PATH) + the window
 *                          breaking points of racon::Polisher::Polish
 *                                                      RavenLib/src/polish.cc:43-51
 *   rvn_find_overlaps_and_create_piles
 *                          raven::FindOverlapsAndCreatePiles
 *                                                      RavenLib/src/construct.cc:14-121
 *
 * The C++ facade over this ABI (include/ram/minimizer_engine.hpp and
 * raven_b200/host/) keeps the reference's class and function signatures, so
 * RavenLib's own sources compile against it unchanged.
 *
 * Conventions: plain pointers and sizes, POD structs, `int` status (0 = ok,
 * negative = error; text via rvn_last_error). No exceptions or C++ types
 * cross this line. Output arrays are owned by the context and stay valid
 * until the next call on the same context that produces the same kind of
 * output, or rvn_ctx_destroy. A context is not thread-safe; use one per
 * host thread / GPU. There is NO CPU fallback: every compute entry point
 * fails with RVN_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RAVEN_B200_H_
#define RAVEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVN_OK 0
#define RVN_ERR_INVALID (-1) /* bad argument (mirrors std::invalid_argument) */
#define RVN_ERR_CUDA (-2)    /* CUDA runtime / no device */
#define RVN_ERR_STATE (-3)   /* call order (e.g. map before minimize) */
#define RVN_ERR_LIMIT (-4)   /* input beyond a documented limit */

typedef struct rvn_ctx rvn_ctx;

/* biosoup::Overlap without the alignment string
 * (field order: RavenLib/src/overlap_utils.cc:5-8) */
typedef struct rvn_overlap {
  uint32_t lhs_id, lhs_begin, lhs_end;
  uint32_t rhs_id, rhs_begin, rhs_end;
  uint32_t score;
  uint32_t strand; /* 1 = same strand */
} rvn_overlap;

/* Counters of the last rvn_map / stage-1 call; the roofline's algorithmic
 * bytes are computed from these (DESIGN.md "algorithmic bytes"). */
typedef struct rvn_stats {
  uint64_t index_bases;     /* bases sketched for the index */
  uint64_t index_records;   /* minimizer records indexed */
  uint64_t index_keys;      /* distinct minimizer values */
  uint64_t query_bases;     /* bases sketched as queries */
  uint64_t query_records;   /* query minimizers probed */
  uint64_t hits;            /* seed hits expanded */
  uint64_t overlaps;        /* overlaps emitted by chaining */
  uint64_t pile_bins;       /* pile bins updated */
  uint64_t kernel_launches; /* kernels of this library launched */
  uint32_t occurrence;      /* frequency threshold in force */
  uint32_t reserved;
} rvn_stats;

int rvn_version(void);

/* device: CUDA ordinal. stream: a cudaStream_t to run on (e.g. torch's
 * current stream) or NULL for a stream owned by the context. */
int rvn_ctx_create(int device, void* stream, rvn_ctx** out);
void rvn_ctx_destroy(rvn_ctx* ctx);
const char* rvn_last_error(const rvn_ctx* ctx);

/* ram::MinimizerEngine(pool, k, w, bandwidth, chain, matches, gap);
 * k is clamped to [1, 31] like the reference engine. Drops any index. */
int rvn_engine_configure(rvn_ctx* ctx, uint32_t k, uint32_t w,
                         uint32_t bandwidth, uint32_t chain, uint32_t matches,
                         uint32_t gap);

/* The read set, in the biosoup wire format: read i is words[word_off[i] ..
 * word_off[i+1]) (32 bases per word, base j at bits [(j<<1)&63,+1] of word
 * j>>5), lens[i] bases, id = i. Host buffers; copied to the device. */
int rvn_reads_upload(rvn_ctx* ctx, const uint64_t* words,
                     const uint64_t* word_off, const uint32_t* lens,
                     uint32_t n_reads);

/* As rvn_reads_upload, but only the bases of reads [first,last) are copied to
 * the device (lengths of all reads are): a rank of a partitioned run only
 * sketches its own range. Sketching any other read is RVN_ERR_STATE. */
int rvn_reads_upload_range(rvn_ctx* ctx, const uint64_t* words,
                           const uint64_t* word_off, const uint32_t* lens,
                           uint32_t n_reads, uint32_t first, uint32_t last);


/* Same, with explicit sequence ids (biosoup::NucleicAcid::id): origins and
 * overlap ids carry ids[i] instead of i. raven re-sorts its sequence vector
 * before stage 2 (construct.cc:324-349), so id != position there. */
int rvn_reads_upload_ids(rvn_ctx* ctx, const uint64_t* words,
                         const uint64_t* word_off, const uint32_t* lens,
                         const uint32_t* ids, uint32_t n_reads);

/* Minimize(reads[first..last), minhash): sketch and index. */
int rvn_minimize(rvn_ctx* ctx, uint32_t first, uint32_t last, int minhash);

/* Filter(frequency): occurrence threshold of the current index.
 * frequency outside [0,1] -> RVN_ERR_INVALID (reference throws). */
int rvn_filter(rvn_ctx* ctx, double frequency, uint32_t* occurrence);

/* Map(read, avoid_equal, avoid_symmetric, minhash[, &filtered]) for every
 * read in [first, last) against the current index. Overlaps come back
 * grouped by query in ascending id, each group in the reference engine's
 * emission order; ovl_off has (last-first)+1 entries. `filtered` are the
 * positions of over-frequent query minimizers (only if want_filtered). */
int rvn_map(rvn_ctx* ctx, uint32_t first, uint32_t last, int avoid_equal,
            int avoid_symmetric, int minhash, int want_filtered);
/* Map one read that is NOT in the uploaded set (e.g. a query of an earlier
 * index batch, construct.cc:59; plasmids vs unitigs, assemble.cc:757,780).
 * Results through rvn_map_results with a single query. */
int rvn_map_external(rvn_ctx* ctx, const uint64_t* words, uint32_t len,
                     uint32_t id, int avoid_equal, int avoid_symmetric,
                     int minhash, int want_filtered);
int rvn_map_results(rvn_ctx* ctx, const rvn_overlap** overlaps,
                    const uint64_t** ovl_off, uint64_t* n_overlaps,
                    const uint32_t** filtered, const uint64_t** filt_off);

/* Pile::AddLayers on a batch of piles: pile i has bins[i] uint16 counters at
 * data + bin_off[i]; every overlap adds +1 to bins [(begin>>4)+1, (end>>4)-1)
 * of the pile of its lhs and of its rhs read, saturating at 65535.
 * Precondition (holds for every chained overlap): *_end >= 16. */
int rvn_pile_add_layers(rvn_ctx* ctx, uint16_t* data, const uint64_t* bin_off,
                        uint32_t n_piles, const rvn_overlap* overlaps,
                        uint64_t n_overlaps);

/* The low-complexity test of Pile::AddKmers (RavenLib/src/pile.cc:64-120) for
 * n (read index, position) pairs - the positions rvn_map reports as filtered
 * (construct.cc:377-383): keep[i] = 1 iff the k-mer at positions[i] survives
 * the three compressions, i.e. iff the reference sets kmers_[position >> 4]. */
int rvn_kmer_complexity(rvn_ctx* ctx, const uint32_t* read_index,
                        const uint32_t* positions, uint64_t n, uint32_t kmer_len,
                        uint8_t* keep);

/* The identity filter's edlibAlign(lhs, rhs, edlibDefaultAlignConfig()).editDistance
 * (RavenLib/src/construct.cc:176-199 and :393-416) for n_pairs overlaps at once:
 * pair i aligns bases [lhs_begin, +lhs_len) of uploaded read lhs_read[i] with
 * [rhs_begin, +rhs_len) of read rhs_read[i], the latter reverse complemented when
 * strand[i] == 0 (construct.cc:184-188). Reads are addressed by their index in
 * the uploaded set. limit (may be NULL): limit[i] >= 0 bounds the search -
 * distance[i] = -1 if the distance exceeds it (the filter only needs to know that
 * 1 - ed/max(len) < identity); limit[i] < 0: exact distance, whatever it takes. */
int rvn_edit_distance_batch(rvn_ctx* ctx, uint64_t n_pairs, const uint32_t* lhs_read,
                            const uint32_t* lhs_begin, const uint32_t* lhs_len,
                            const uint32_t* rhs_read, const uint32_t* rhs_begin,
                            const uint32_t* rhs_len, const uint8_t* strand,
                            const int32_t* limit, int32_t* distance);

/* raven::TrimAndAnnotatePiles' first two steps for every pile (RavenLib/src/construct.cc:
 * 123-152: Pile::FindValidRegion(coverage) and Pile::FindMedian, pile.cc:122-172), on
 * the piles the last rvn_find_overlaps_and_create_piles call left on the device - no
 * round trip of the histograms. Per read i: invalid[i] = 1 if the reference marks the
 * pile invalid (no run of bins >= coverage followed by a lower bin, or one shorter than
 * 1260 >> 4 bins; begin/end then stay 0 / bins like the reference's fields); else
 * [begin[i], end[i]) is the valid region in bins and median[i] the element of rank
 * size/2 of it. The caller zeroes the bins outside the region like UpdateValidRegion.
 * RVN_ERR_STATE if another call has replaced the device piles since. */
int rvn_stage1_pile_regions(rvn_ctx* ctx, uint32_t coverage, uint32_t* begin, uint32_t* end,
                            uint16_t* median, uint8_t* invalid);

/* The read-to-target alignments of racon::Polisher::Polish (RavenLib/src/polish.cc:43-51:
 * edlibAlign(query, target, EDLIB_MODE_NW, EDLIB_TASK_PATH) per read) and the walk along
 * each path that cuts it at the target's windows, for n_pairs alignments at once.
 * Pair i: query = bases [q_begin, +q_len) of uploaded read q_read[i], reverse
 * complemented as a whole when strand[i] == 0; target = bases [t_begin, +t_len) of read
 * t_read[i]. The path is the ONE upstream edlib returns (its traceback order and its
 * Hirschberg split of large problems are kept), so the cuts are the reference's.
 * Output: distance[i] = the edit distance, and for every window x of `window` target
 * bases that the target substring touches (x = t_begin/window .. (t_begin+t_len-1)/window)
 * the slot s = bp_off[i] + (x - t_begin/window) of breaking_points (4 values per slot):
 *   [4s+0] target position of the first match/mismatch column inside the window,
 *   [4s+1] its query position (0-based inside the oriented query substring),
 *   [4s+2], [4s+3] one past the last such pair;
 * all four 0xFFFFFFFF when no match/mismatch falls into the window. bp_off (n_pairs + 1
 * entries, computed by the caller) must count exactly those windows. */
int rvn_align_breaking_points(rvn_ctx* ctx, uint64_t n_pairs, const uint32_t* q_read,
                              const uint32_t* q_begin, const uint32_t* q_len,
                              const uint8_t* strand, const uint32_t* t_read,
                              const uint32_t* t_begin, const uint32_t* t_len, uint32_t window,
                              const uint64_t* bp_off, int32_t* distance,
                              uint32_t* breaking_points);

/* raven::FindOverlapsAndCreatePiles over the uploaded read set: index batches
 * of >= index_batch_bases (reference: 1<<32), query flushes of >=
 * query_batch_bases (reference: 1<<30); pass 0 for the reference values.
 * Results: per read the kept overlaps (<= max_overlaps after the reference's
 * truncation rule) and the pile histogram (len>>4 uint16 bins). */
int rvn_find_overlaps_and_create_piles(rvn_ctx* ctx, double frequency,
                                       uint64_t max_overlaps, int minhash,
                                       uint64_t index_batch_bases,
                                       uint64_t query_batch_bases);
int rvn_stage1_results(rvn_ctx* ctx, const rvn_overlap** overlaps,
                       const uint64_t** ovl_off, const uint16_t** pile,
                       const uint64_t** pile_off, uint64_t* n_mapped);

/* racon::Window::GenerateConsensus over a batch of windows - the unit of the
 * "POA windows/s" metric; stands in for the consensus phase of
 * racon::Polisher::Polish (RavenLib/src/polish.cc:43-51; raven passes w = 500,
 * trim = true, m/n/g = PolishCfg polish.hpp:13-17). Flat layout: window w owns
 * sequences [win_first[w], win_first[w+1]); the first is the backbone (draft),
 * the others are layers aligned to backbone positions [seq_begin, seq_end]
 * (inclusive, < backbone length); sequence s = bases[seq_off[s]..seq_off[s+1])
 * as ACGT letters, quals (Phred+33) at the same offsets or NULL (layers weigh 1,
 * the backbone 0 like racon's dummy '!' quality).
 * g must be negative (racon throws otherwise -> RVN_ERR_INVALID).
 * Results: consensus letters per window (cons_off has n_windows+1 entries),
 * status bit 0 = polished (>= 3 sequences), bit 1 = trimming skipped
 * ("might be chimeric"), coverage per consensus base if requested. */
int rvn_poa_batch(rvn_ctx* ctx, uint32_t n_windows, const uint32_t* win_first,
                  const uint64_t* seq_off, const char* bases, const char* quals,
                  const uint32_t* seq_begin, const uint32_t* seq_end, int8_t m,
                  int8_t n, int8_t g, int trim, int tgs, int want_coverage);
int rvn_poa_results(rvn_ctx* ctx, const char** consensus, const uint64_t** cons_off,
                    const uint8_t** status, const uint32_t** coverage,
                    uint64_t* cells);

/* ---- introspection (parity tests, profiling) ---- */

/* Per-read sketches of reads [first,last) exactly as the reference engine's
 * Minimize(sequence, minhash) returns them: value[], origin[] and
 * (last-first)+1 offsets. */
int rvn_sketch(rvn_ctx* ctx, uint32_t first, uint32_t last, int minhash,
               const uint64_t** value, const uint64_t** origin,
               const uint64_t** offsets, uint64_t* n_records);

/* The index as sorted (value, origin) records + distinct-key statistics. */
int rvn_index_records(rvn_ctx* ctx, const uint64_t** value,
                      const uint64_t** origin, uint64_t* n_records,
                      uint64_t* n_keys);

/* Seed hits of the last rvn_map call with keep_hits set via rvn_set_option:
 * group[], positions[] and per-query offsets, unsorted within a query. */
int rvn_map_hits(rvn_ctx* ctx, const uint64_t** group,
                 const uint64_t** positions, const uint64_t** hit_off,
                 uint64_t* n_hits);

/* The engine's stable LSD radix sort (raven_b200/csrc/radix.cu - it orders the
 * minimizer index, construct.cc:42-43) on host arrays, in place: keys of 4 or 8
 * bytes, values of 0 (u32 keys only), 4 or 8 bytes, key bits [begin_bit,
 * end_bit); equal keys keep their input order. */
int rvn_debug_sort_pairs(rvn_ctx* ctx, int key_bytes, int val_bytes, void* keys,
                         void* vals, uint64_t n, int begin_bit, int end_bit,
                         int descending);

int rvn_get_stats(rvn_ctx* ctx, rvn_stats* out);

/* options: "keep_hits" (0/1); "tier_min_records" (stage 1 splits an index batch of
 * at least that many minimizers into a probe-able tier and bare keys, default
 * 2^18; 0 = always); "self_join" (0/1, default 1: the seed hits of a stage-1 flush
 * whose reads are inside the index batch come from a self-join over the sorted
 * index instead of a probe per micromizer); "async_upload" (0/1, default 0: with 1,
 * rvn_reads_upload* returns while the packed bases still travel - in chunks, on a
 * copy stream - and the sketch kernel of the next call starts on the reads that
 * have arrived; the caller must keep `words` unchanged until the next call on the
 * context has returned; meant for pinned host memory); "reset_stats".
 * Unknown name -> RVN_ERR_INVALID. */
int rvn_set_option(rvn_ctx* ctx, const char* name, int64_t value);

/* Device time (ms, CUDA events on the context's stream) of the phases of the
 * last stage-1 / minimize / map call: names[] are static strings. */
int rvn_get_timings(rvn_ctx* ctx, const char* const** names,
                    const float** ms, uint32_t* n);

/* ---- multi-GPU building blocks -------------------------------------------
 * One context per rank (one process per GPU); the collectives between the
 * steps are the caller's (raven_b200/distributed.py: torch.distributed over
 * NCCL). With n_parts ranks: index keys are owned by value mod n_parts; a read
 * (as a query, and its pile and overlap list) by id mod n_parts; sketching is
 * split by any contiguous read ranges, ascending with the rank.
 * Every d_* pointer is DEVICE memory, owned by the context (outputs: valid
 * until the next call on it) or by the caller (inputs). Per index batch of
 * raven::FindOverlapsAndCreatePiles (RavenLib/src/construct.cc:36-112):
 *   sketch_split -> all-to-all -> index -> histogram -> all-reduce ->
 *   set_occurrence -> hits_split -> all-to-all -> chain -> overlaps_split ->
 *   all-to-all -> stage1_add; then stage1_end and rvn_dist_stage1_results
 *   (the piles and overlap lists of the reads this rank owns).
 * "runs": an all-to-all delivers one run per source rank; run_off (host,
 * n_runs + 1 entries) are their offsets in the received array. */
int rvn_dist_sketch_split(rvn_ctx* ctx, uint32_t first, uint32_t last, int minhash,
                          uint32_t n_parts, const uint64_t** d_value,
                          const uint64_t** d_origin, uint64_t* counts);
int rvn_dist_index(rvn_ctx* ctx, const uint64_t* d_value, const uint64_t* d_origin,
                   uint64_t n_records, uint64_t index_bases);
/* The same with the tiers of stage 1 (index.cu): records whose value exceeds
 * value_limit - the largest micromizer value of ANY read that can query this batch,
 * i.e. the maximum over the ranks of rvn_dist_max_threshold - are only counted for
 * the occurrence threshold. ~0: no tiers. */
int rvn_dist_index_limited(rvn_ctx* ctx, const uint64_t* d_value, const uint64_t* d_origin,
                           uint64_t n_records, uint64_t index_bases, uint64_t value_limit);
/* Largest micromizer value of the reads [first, last) this rank sketches (0 if empty). */
int rvn_dist_max_threshold(rvn_ctx* ctx, uint32_t first, uint32_t last, uint64_t* value);
/* run-length histogram of this rank's keys (u64 bins; bin i = keys with i
 * postings, last bin = longer runs) */
int rvn_dist_histogram(rvn_ctx* ctx, const uint64_t** d_hist, uint32_t* n_bins,
                       uint64_t* n_keys);
/* MinimizerEngine::Filter on the summed (host) histogram: one global threshold */
int rvn_dist_set_occurrence(rvn_ctx* ctx, const uint64_t* hist, uint64_t n_keys,
                            double frequency, uint32_t* occurrence);
/* seed hits of the received query records (sorted by read id; reads
 * [0, n_query_reads)) against this rank's keys, written as n_parts runs by the
 * owner of the query read; d_lhs = query read of every hit */
int rvn_dist_hits_split(rvn_ctx* ctx, const uint64_t* d_qvalue,
                        const uint64_t* d_qorigin, uint64_t n_queries,
                        int avoid_equal, int avoid_symmetric, uint32_t n_parts,
                        uint32_t n_query_reads, const uint64_t** d_group,
                        const uint64_t** d_positions, const uint32_t** d_lhs,
                        uint64_t* counts);
/* chains the hits of the owned reads below n_query_reads (runs sorted by query
 * read): overlaps in query order */
int rvn_dist_chain(rvn_ctx* ctx, const uint64_t* d_group, const uint64_t* d_positions,
                   const uint32_t* d_lhs, uint64_t n_hits, uint32_t n_runs,
                   const uint64_t* run_off, uint32_t n_parts, uint32_t rank,
                   uint32_t n_query_reads, const rvn_overlap** d_overlaps,
                   uint64_t* n_overlaps);
/* the overlaps of the last rvn_dist_chain as n_parts runs: run d != rank holds
 * those whose rhs read rank d owns, run `rank` holds all of them */
int rvn_dist_overlaps_split(rvn_ctx* ctx, uint32_t n_parts, uint32_t rank,
                            const rvn_overlap** d_overlaps, uint64_t* counts);
/* piles + per-read overlap lists of the owned reads: begin once, add once per
 * index batch (runs sorted by query read; reference flush schedule), end */
int rvn_dist_stage1_begin(rvn_ctx* ctx, uint32_t n_parts, uint32_t rank);
int rvn_dist_stage1_add(rvn_ctx* ctx, const rvn_overlap* d_overlaps,
                        uint64_t n_overlaps, uint32_t n_runs, const uint64_t* run_off,
                        uint32_t n_query_reads, uint64_t max_overlaps,
                        uint64_t query_batch_bases);
int rvn_dist_stage1_end(rvn_ctx* ctx);
/* owned read j is read rank + j * n_parts; arrays as rvn_stage1_results, over
 * the n_owned owned reads; n_mapped = overlaps found by this rank's queries */
int rvn_dist_stage1_results(rvn_ctx* ctx, const rvn_overlap** overlaps,
                            const uint64_t** overlap_off, const uint16_t** pile,
                            const uint64_t** pile_off, uint32_t* n_owned,
                            uint64_t* n_mapped);

/* ---- peer-memory exchange: the all-to-alls of the schedule as direct DMA
 * writes into the destination rank's receive arena over NVLink (CUDA IPC; one
 * process per GPU on one node). export: (re)allocate this rank's arena, 64-byte
 * handle out; import: map the peers' arenas from all n_parts handles (own slot
 * ignored); put: asynchronous copy of device memory into rank dest's arena at
 * dst_offset; put_flush: wait for this rank's puts. All arenas must have the
 * same size (the caller checks offsets against it); the caller brackets a
 * round of puts with barriers and closes the peers before any re-export. */
int rvn_dist_arena_export(rvn_ctx* ctx, uint64_t bytes, void* handle64);
int rvn_dist_arena_import(rvn_ctx* ctx, uint32_t n_parts, uint32_t rank,
                          const void* handles);
int rvn_dist_arena_close_peers(rvn_ctx* ctx);
int rvn_dist_arena(rvn_ctx* ctx, void** d_arena, uint64_t* bytes);
int rvn_dist_put(rvn_ctx* ctx, uint32_t dest, uint64_t dst_offset, const void* d_src,
                 uint64_t bytes);
int rvn_dist_put_flush(rvn_ctx* ctx);

#ifdef __cplusplus
}
#endif

#endif /* RAVEN_B200_H_ */
