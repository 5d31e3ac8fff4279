// Internal helpers shared by the host and device halves of libw2b (not part of the ABI).
#pragma once
#include <stdint.h>

void w2b_set_error(const char *fmt, ...);
// InitUnigramTable (src/word2bits.cpp:112-128) in boundary form: start[i] = first table
// slot owned by word i, start[V] = 1e8.  Same libm pow() and the same double arithmetic
// as the reference loop, so expanding it reproduces the 1e8-entry table bit for bit.
void w2b_unigram_bounds(const int64_t *cn, int64_t V, int32_t *start);
// expTable (:614-618), host expf.
void w2b_exptable(float *out /*1000*/);
// Sub-sampling thresholds `ran` (:403-404), float32.
void w2b_keep_thresholds(const int64_t *cn, int64_t V, int64_t train_words, float sample, float *out /*V*/);
