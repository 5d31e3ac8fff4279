// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Library wrapper around the UNMODIFIED reference translation unit
// (/root/reference/src/word2bits.cpp), compiled where it lies.  The reference's
// `main` is renamed by the preprocessor so that its file-scope functions and
// globals (quantize :73, InitUnigramTable :112, LearnVocabFromTrainFile :265,
// InitNet :343, TrainModelThread :363, u/v/table/expTable/alpha :45-61) become
// callable from tests through the small extern "C" surface below.
//
// The setup that the reference performs inside main() (:612-618: vocab /
// vocab_hash allocation and the expTable fill) and inside TrainModel()
// (:521,:524: thread_losses, starting_alpha) has to be re-done here because
// those functions are not called.
//
// Build: see oracle/Makefile (outputs go to oracle/_ref/, git-ignored).
#ifndef W2B_REF_SOURCE
#error "compile with -DW2B_REF_SOURCE=\"/root/reference/src/word2bits.cpp\""
#endif
#define main w2b_reference_main
#include W2B_REF_SOURCE
#undef main

#include <stdint.h>

extern "C" {

// Configure the reference's globals the way main() would from argv.
void ref_configure(const char *train, long long size, int win, int neg, int bits,
                   int threads, long long iters, int mincount, float a, float smp,
                   float rg) {
  strncpy(train_file, train, MAX_STRING - 1);
  strcpy(output_file, "unused");
  layer1_size = size; window = win; negative = neg; bitlevel = bits;
  num_threads = threads; iter = iters; min_count = mincount;
  alpha = a; sample = smp; reg = rg;
  debug_mode = 0; binary = 1; classes = 0; save_every_epoch = 0;
  word_count_actual = 0; train_words = 0; file_size = 0;
  vocab_max_size = 1000; vocab_size = 0;
  if (vocab) free(vocab);
  vocab = (struct vocab_word *)calloc(vocab_max_size, sizeof(struct vocab_word));
  if (!vocab_hash) vocab_hash = (int *)calloc(vocab_hash_size, sizeof(int));
  if (!expTable) {
    expTable = (real *)malloc((EXP_TABLE_SIZE + 1) * sizeof(real));
    for (int k = 0; k < EXP_TABLE_SIZE; k++) {
      real e = exp((k / (real)EXP_TABLE_SIZE * 2 - 1) * MAX_EXP);
      expTable[k] = e;
      expTable[k] = expTable[k] / (expTable[k] + 1);
    }
  }
  if (u) { free(u); u = NULL; }
  if (v) { free(v); v = NULL; }
  if (table) { free(table); table = NULL; }
  if (thread_losses) free(thread_losses);
  thread_losses = (double *)calloc(threads > 0 ? threads : 1, sizeof(double));
  starting_alpha = alpha;
}

void ref_learn_vocab(void) { LearnVocabFromTrainFile(); }
void ref_init_net(void) { InitNet(); }
void ref_init_unigram(void) { InitUnigramTable(); }

long long ref_vocab_size(void) { return vocab_size; }
long long ref_train_words(void) { return train_words; }
long long ref_file_size(void) { return file_size; }
long long ref_word_count_actual(void) { return word_count_actual; }
void ref_set_word_count_actual(long long x) { word_count_actual = x; }
const char *ref_vocab_word(long long i) { return vocab[i].word; }
long long ref_vocab_cn(long long i) { return vocab[i].cn; }
float *ref_u(void) { return u; }
float *ref_v(void) { return v; }
int *ref_table(void) { return table; }
int ref_table_size(void) { return table_size; }
float *ref_exptable(void) { return expTable; }
float ref_get_alpha(void) { return alpha; }
void ref_set_alpha(float a) { alpha = a; }
float ref_quantize(float x, int b) { return quantize(x, b); }
float ref_sigmoid(float x) { return sigmoid(x); }
double ref_thread_loss(int id) { return thread_losses[id]; }

// TrainModelThread ends in pthread_exit (:515), so it must run on its own thread.
void ref_train_thread(long long id) {
  pthread_t t;
  thread_losses[id] = 0;
  pthread_create(&t, NULL, TrainModelThread, (void *)id);
  pthread_join(t, NULL);
}

// One epoch exactly as TrainModel's loop body does it (:534-538); returns the epoch loss.
double ref_train_epoch(void) {
  pthread_t *pt = (pthread_t *)malloc(num_threads * sizeof(pthread_t));
  memset(thread_losses, 0, sizeof(double) * num_threads);
  for (long a = 0; a < num_threads; a++) pthread_create(&pt[a], NULL, TrainModelThread, (void *)a);
  for (long a = 0; a < num_threads; a++) pthread_join(pt[a], NULL);
  double s = 0;
  for (long a = 0; a < num_threads; a++) s += thread_losses[a];
  free(pt);
  return s;
}

}  // extern "C"
