// Internal helpers shared by the host and device halves of libw2b (not part of the ABI).
#pragma once
#include <stdint.h>

#include <exception>
#include <new>

void w2b_set_error(const char *fmt, ...);

// Nothing is thrown across the C ABI: entry points that allocate host memory or start threads run their body
// through this guard (std::bad_alloc -> W2B_ENOMEM, anything else -> W2B_EINVAL, message in w2b_last_error()).
template <class F>
static inline int w2b_guarded(const char *name, F &&body) {
  try {
    return body();
  } catch (const std::bad_alloc &) {
    w2b_set_error("%s: out of host memory", name);
    return 6;  // W2B_ENOMEM
  } catch (const std::exception &ex) {
    w2b_set_error("%s: %s", name, ex.what());
    return 1;  // W2B_EINVAL
  } catch (...) {
    w2b_set_error("%s: unexpected exception", name);
    return 1;
  }
}
// InitUnigramTable (src/word2bits.cpp:112-128) in boundary form: start[i] = first table
// slot owned by word i, start[V] = 1e8.  Same libm pow() and the same double arithmetic
// as the reference loop, so expanding it reproduces the 1e8-entry table bit for bit.
void w2b_unigram_bounds(const int64_t *cn, int64_t V, int32_t *start);
// expTable (:614-618), host expf.
void w2b_exptable(float *out /*1000*/);
// Sub-sampling thresholds `ran` (:403-404), float32.
void w2b_keep_thresholds(const int64_t *cn, int64_t V, int64_t train_words, float sample, float *out /*V*/);
// Streaming mode: copy tokens [max(cursor,0), +L) of every unfinished shard into its slice stage[i*L ..) and
// report where the slice sits (xlate = global index - staging index, limit = global end, eof flag).  A plain
// memcpy of tens of MB per step, spread over a few host threads (nthreads <= 0: pick from the size).
void w2b_gather_slices(const int32_t *ids, long long n_tokens, long long L, int nshards, const long long *cursor,
                       const int *done, int32_t *stage, long long *xlate, long long *limit, int *limit_is_eof,
                       int nthreads);
