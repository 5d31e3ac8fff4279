/* libw2b — C ABI of the B200-native Word2Bits training path.
 *
 * The reference (agnusmaximus/Word2Bits, src/word2bits.cpp) has no FFI: it is one
 * executable whose hot path is `void *TrainModelThread(void *id)` (:363-516) reading and
 * writing file-scope globals (:45-61).  This header is the seam a maintainer would cut
 * there: every entry point below names the reference code it replaces.  Plain C types
 * only; the caller owns host buffers, the library owns device memory; one context is
 * driven by one host thread; all calls are synchronous; every function returns 0 on
 * success or a non-zero W2B_E* code, and w2b_last_error() gives the message (the
 * reference's convention is printf + exit(1), which the CLI wrapper reproduces).
 */
#ifndef W2B_H
#define W2B_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2B_OK 0
#define W2B_EINVAL 1   /* bad argument / unsupported configuration */
#define W2B_ECUDA 2    /* CUDA runtime error (no device, launch failure, OOM) */
#define W2B_EIO 3      /* file not found / unreadable / unwritable */
#define W2B_ESTATE 4   /* call sequence error (e.g. train before set_corpus) */
#define W2B_ENCCL 5    /* NCCL error / NCCL not loadable */
#define W2B_ENOMEM 6   /* host allocation failed (nothing is thrown across the ABI) */

#define W2B_TABLE_SIZE 100000000 /* table_size, :60 */
#define W2B_MAX_SENTENCE 1000    /* MAX_SENTENCE_LENGTH, :32 */
#define W2B_MAX_WINDOW 512   /* a sentence has at most 1000 words (:32): beyond ~500 every word is context anyway */
#define W2B_MAX_NEGATIVE 63   /* 1 + negative targets per position fit one 64-entry trace record / two lanes-worth of loss terms */

#define W2B_MODE_FAST 0   /* production: one warp per shard, all shards concurrent (Hogwild) */
#define W2B_MODE_STRICT 1 /* parity: shards one after another, sequential IEEE op order */

/* ------------------------------------------------------------------ host glue (no GPU)
 * Corpus reader + vocabulary: replaces ReadWord/SearchVocab/AddWordToVocab/SortVocab/
 * LearnVocabFromTrainFile (:131-301).  Results (word order, counts, train_words,
 * file_size) are identical; the text is tokenised ONCE into an int32 id stream. */
typedef struct w2b_corpus w2b_corpus;
int w2b_corpus_load(const char *train_file, int min_count, w2b_corpus **out);
void w2b_corpus_free(w2b_corpus *c);
int64_t w2b_corpus_vocab_size(const w2b_corpus *c);  /* vocab_size, :50 */
int64_t w2b_corpus_train_words(const w2b_corpus *c); /* train_words, :51 */
int64_t w2b_corpus_file_size(const w2b_corpus *c);   /* file_size, :299 */
const char *w2b_corpus_word(const w2b_corpus *c, int64_t i);
const int64_t *w2b_corpus_counts(const w2b_corpus *c); /* vocab[i].cn */
int64_t w2b_corpus_num_tokens(const w2b_corpus *c);    /* in-vocab tokens incl. </s> */
const int32_t *w2b_corpus_tokens(const w2b_corpus *c);
/* Shard i of n starts where the reference's fseek(file_size/n*i) (:377) puts thread i:
 * first[i] = id of the first token read there (a suffix fragment when the seek lands
 * mid-word; -1 if it is out of vocabulary), start[i] = index of the next regular token. */
int w2b_corpus_shards(const w2b_corpus *c, int n, int64_t *start, int32_t *first);
/* Vector file writer (:560-576): header "%lld %lld\n", then "<word> " + D values
 * ("%lf " text or raw float32) + "\n" per word. */
int w2b_write_vectors(const char *path, const w2b_corpus *c, const float *vectors, int64_t V,
                      int64_t D, int binary);

/* Packed vector file (the README's storage claim is realised only by gzip in the
 * reference).  Header "<V> <D> <bitlevel>\n", then per word "<word> " + ceil(D*bitlevel/8) bytes
 * (value j occupies bits [j*bitlevel, (j+1)*bitlevel) little-endian: bit 0 = sign (1 = negative),
 * bit 1 (bitlevel 2) = magnitude (1 = .75)) + "\n".  bitlevel 1 and 2 only.  unpack restores the exact
 * float32 levels, so `unpack -> w2b_write_vectors(binary=1)` feeds compute_accuracy unchanged. */
int w2b_write_packed(const char *path, const w2b_corpus *c, const float *vectors, int64_t V, int64_t D,
                     int bitlevel);
int w2b_read_packed_header(const char *path, int64_t *V, int64_t *D, int *bitlevel);
int w2b_read_packed(const char *path, float *vectors /* V*D */, char *words /* V*max_word */, int max_word);

/* Host-side arithmetic of the path, callable (and tested) without a GPU.  The device path uses exactly
 * these functions for what it uploads: unigram boundaries (InitUnigramTable :112-128 in boundary form:
 * start[i] = first of the 1e8 table slots owned by word i, start[V] = 1e8), expTable (:614-618, same libm
 * expf), the sub-sampling thresholds `ran` (:403-404, float32), and the k-step / 2^j-step jump constants
 * of the LCG r*25214903917+11 (:352,:405,:428,:455) that let 32 lanes take 32 draws at once
 * (ja/jc: 65 entries, r_k = r*ja[k]+jc[k]; pa/pc: 64 entries for 2^j steps). */
int w2b_host_unigram_bounds(const int64_t *cn, int64_t V, int32_t *start /* V+1 */);
int w2b_host_exptable(float *out /* 1000 */);
int w2b_host_keep_thresholds(const int64_t *cn, int64_t V, int64_t train_words, float sample, float *out /* V */);
int w2b_host_lcg_tables(uint64_t *ja /*65*/, uint64_t *jc /*65*/, uint64_t *pa /*64*/, uint64_t *pc /*64*/);

/* The host half of a streaming step (w2b_set_corpus resident = 0): the next L tokens of every unfinished shard are
 * gathered into one pinned staging buffer (slice i at stage[i*L]) by a few host threads before the single H2D copy.
 * xlate[i] = global token index - staging index, limit[i] = global end of the slice, limit_is_eof[i] = slice reaches
 * the end of the stream.  Outputs of finished shards are left untouched.  nthreads <= 0: chosen from the size. */
int w2b_host_gather_slices(const int32_t *ids, int64_t n_tokens, int64_t L, int nshards, const int64_t *cursor,
                           const int32_t *done, int32_t *stage /* nshards*L */, int64_t *xlate, int64_t *limit,
                           int32_t *limit_is_eof, int nthreads);

/* ------------------------------------------------------------------------ device path */
typedef struct w2b_ctx w2b_ctx;

typedef struct {
  int64_t vocab_size;  /* V, incl. </s> at 0 */
  int64_t layer1_size; /* -size */
  int32_t window;      /* -window */
  int32_t negative;    /* -negative */
  int32_t bitlevel;    /* -bitlevel */
  float alpha;         /* -alpha (starting_alpha, :524) */
  float sample;        /* -sample */
  float reg;           /* -reg */
  int64_t iter;        /* -iter: enters the learning-rate schedule (:391) */
  int32_t num_shards;  /* -threads: TOTAL number of corpus shards S */
  int32_t shard_begin; /* this context trains shards [shard_begin, shard_end) of S */
  int32_t shard_end;   /*   (0,0 = all; used to split S across GPUs) */
  int32_t device;      /* CUDA device ordinal */
  int32_t mode;        /* W2B_MODE_FAST | W2B_MODE_STRICT */
  int32_t group;       /* register kernel: target rows in flight per CTA step (0 = default) */
  int32_t plain_store; /* reserved, must be 0 (round 1's racy load/add/store variant of the register kernel is gone) */
  int32_t kernel;      /* fast mode: 0 = warp-per-shard kernel when applicable (default; csrc/w2b_warp.cuh),
                          1 = register kernel (one CTA per shard; also serves strict mode and D > 1024) */
  int32_t slots;       /* warp kernel: shared-memory row slots per warp (0 = as many as fit, at most 16) */
  int32_t prefetch;    /* warp kernel: 0 = the positions of a shard strictly one after another, like a reference
                          thread (default; measured free on B200); 1 = rows of position p+1 are fetched before p's
                          updates have landed (a context row shared by neighbours is read one update stale) */
  int32_t sync_mode;   /* multi-GPU exchange (w2b_sync): 0 = replicas are averaged (default); 1 = every rank's
                          updates since the last exchange are summed onto the common base (needs two more tables;
                          call w2b_nccl_init after w2b_init_tables / w2b_checkpoint_load) */
} w2b_config;

typedef struct {
  double loss;            /* sum of shard losses accumulated by this call */
  int64_t words;          /* word_count advanced (reference semantics, :399) */
  int64_t positions;      /* trained positions (cw > 0) */
  int64_t context_rows;   /* sum of cw */
  int64_t target_rows;    /* processed targets (skips excluded) */
  int64_t shards_done;    /* shards that have reached their end */
  float alpha;            /* alpha after the call */
  int64_t word_count_actual;
  float kernel_ms;        /* CUDA-event time of the training kernel(s) in this call */
  int32_t launches;       /* kernels launched by this call */
  int64_t h2d_bytes;      /* host->device bytes copied by this call (token slices, shard states) */
  int64_t d2h_bytes;      /* device->host bytes copied by this call (shard states, alpha, counter) */
} w2b_step_stats;

/* One record per loop iteration that reaches the window draw (:428). */
typedef struct {
  int32_t center, b, cw, ntargets;
  int32_t targets[64];
  float alpha;
} w2b_trace_rec;

const char *w2b_last_error(void);
int w2b_device_count(int *n);

/* Geometry the production (warp-per-shard) kernel would run with for a configuration: pure host arithmetic (no
 * CUDA call).  warp = 0: the configuration runs the register kernel instead (D > 1024, strict mode, kernel = 1). */
typedef struct {
  int32_t warp;           /* 1 = the warp kernel applies */
  int32_t slots;          /* K: shared-memory row slots of a warp's ring (K-2 loads in flight) */
  int32_t queue_entries;  /* job queue capacity (two positions) */
  int32_t warps_per_sm;   /* resident 1-warp CTAs per SM the registers are sized for */
  int32_t sentence_in_smem; /* 1: the shard's sentence buffer (4000 B) is part of smem_bytes; 0: global scratch */
  int32_t reserved;
  int64_t smem_bytes;     /* dynamic shared memory per warp */
} w2b_warp_plan;
int w2b_warp_plan_query(const w2b_config *cfg, w2b_warp_plan *out);

/* Number of shards that keeps every SM busy for this configuration (SMs x resident warps of the production kernel);
 * the CLI's default for -threads (the reference's default of 12 is a CPU core count). */
int w2b_suggest_shards(const w2b_config *cfg, int *out);
int w2b_create(const w2b_config *cfg, w2b_ctx **out); /* globals :45-61 -> context */
int w2b_destroy(w2b_ctx *ctx);

/* vocab[].cn + train_words -> sub-sampling thresholds (:403-404) and the 1e8-entry
 * unigram table (InitUnigramTable, :112-128; boundaries on the host with the same libm
 * pow(), expanded on the device). */
int w2b_set_vocab_counts(w2b_ctx *ctx, const int64_t *cn, int64_t V, int64_t train_words);
/* The id stream + shard starts (replaces each thread's fopen/fseek/ReadWordIndex,
 * :376-377,:396).  resident=1 uploads the whole stream once; resident=0 keeps the host
 * pointer (must stay valid) and w2b_train_step copies each shard's next slice. */
int w2b_set_corpus(w2b_ctx *ctx, const int32_t *ids, int64_t n, const int64_t *shard_start,
                   const int32_t *shard_first, int resident);
/* InitNet (:343-361) by LCG jump-ahead on the device + expTable (:614-618, host expf). */
int w2b_init_tables(w2b_ctx *ctx);

/* Re-arms every shard (seed = shard id :368, cursor = shard start :377) — what the
 * per-epoch pthread_create does (:532-535). */
int w2b_epoch_begin(w2b_ctx *ctx);
/* Advances every unfinished shard by >= words_per_shard words, whole sentences only
 * (<=0: to the end of the shard).  The per-step equivalent of TrainModelThread. */
int w2b_train_step(w2b_ctx *ctx, int64_t words_per_shard, w2b_step_stats *stats);
/* epoch_begin + steps until all shards are done; *loss = "Epoch Loss" (:537-539). */
int w2b_train_epoch(w2b_ctx *ctx, double *loss, w2b_step_stats *stats);

/* Parity hooks */
int w2b_trace(w2b_ctx *ctx, int shard, int64_t max_iterations, w2b_trace_rec *out, int64_t cap,
              int64_t *n_out); /* draws only; does not touch u/v or shard state */
int w2b_strict_prefix(w2b_ctx *ctx, int shard, int64_t max_iterations, double *loss); /* strict mode: first k iterations of a shard */
int w2b_apply_position(w2b_ctx *ctx, const int32_t *context_ids, int cw, const int32_t *targets,
                       int ntargets, float *f_out); /* Appendix-A steps 5-7 for explicit ids */
int w2b_get_state(w2b_ctx *ctx, float *alpha, int64_t *word_count_actual);
int w2b_set_state(w2b_ctx *ctx, float alpha, int64_t word_count_actual);
int w2b_download_raw(w2b_ctx *ctx, float *u, float *v);   /* fp32 master tables */
int w2b_upload_raw(w2b_ctx *ctx, const float *u, const float *v);
int w2b_download_table(w2b_ctx *ctx, int32_t *table);    /* 1e8 entries */
int w2b_download_exptable(w2b_ctx *ctx, float *t);       /* 1000 entries */

/* Resumable checkpoint (-save-every-epoch only keeps the non-resumable quantized
 * sum, :540-557): fp32 master tables + learning rate + global word counter + epochs done. */
int w2b_checkpoint_save(w2b_ctx *ctx, const char *path, int64_t epochs_done);
int w2b_checkpoint_load(w2b_ctx *ctx, const char *path, int64_t *epochs_done);

/* quantize(u+v) (:568-569), V*D floats into host memory. */
int w2b_export(w2b_ctx *ctx, float *out);
/* quantize() itself on the device, for known-answer tests (:73-108). */
int w2b_quantize(w2b_ctx *ctx, const float *in, float *out, int64_t n, int bitlevel);

/* Analogy evaluator, replaces src/compute-accuracy.c:63-189: same inputs
 * (word2vec-binary vector file, optional re-quantisation, vocabulary threshold, question stream),
 * same report text; all questions are scored on the GPU: one Q x V x D TF32 tensor-core contraction
 * (tcgen05 + TMA) as a filter with a proven error bound, then an fp32 re-score of the surviving candidates in the
 * reference's operation order, so the arg-max (ties included) is the reference's.  questions_file NULL = stdin.  report may be NULL. */
typedef struct {
  int64_t questions_total, questions_seen, correct;
  int64_t semantic_correct, semantic_seen, syntactic_correct, syntactic_seen;
  int64_t vocab, size;
  float gpu_ms; /* normalise + query build + scoring kernels, CUDA events */
  int64_t candidates; /* (question, word) pairs the tensor-core filter let through (within 2 eps of the running best) */
  int64_t rescored;   /* ... of which still within 2 eps of the final best: scored again in fp32, reference order */
} w2b_accuracy;
int w2b_compute_accuracy(const char *vectors_file, int bitlevel, int64_t threshold, const char *questions_file,
                         int device, w2b_accuracy *acc, char *report, int64_t report_cap);

/* Multi-GPU replica averaging; G=1 contexts never touch NCCL. */
int w2b_device_ptrs(w2b_ctx *ctx, void **u, void **v, int64_t *elems);
int w2b_nccl_unique_id(void *id128);                                    /* ncclGetUniqueId */
int w2b_nccl_init(w2b_ctx *ctx, const void *id128, int rank, int nranks); /* ncclCommInitRank */
int w2b_sync(w2b_ctx *ctx); /* all-reduce-average u, v; exact global word_count_actual — one NCCL group, no host round trip */
int w2b_sync_timed(w2b_ctx *ctx, float *ms); /* same; *ms = device time of the exchange (CUDA events) */
/* Fingerprints of u and v (sum of the 32-bit patterns mod 2^64): equal on every rank right after w2b_sync. */
int w2b_table_checksum(w2b_ctx *ctx, uint64_t *u_sum, uint64_t *v_sum);
int w2b_scale_tables(w2b_ctx *ctx, float s); /* u*=s, v*=s (for host-driven all-reduce) */

#ifdef __cplusplus
}
#endif
#endif
