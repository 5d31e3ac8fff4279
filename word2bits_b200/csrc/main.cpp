// word2bits — drop-in command line for the B200 training path.
//
// Same flags, defaults, progress lines, exit codes and output files as the reference's main()
// / TrainModel() (src/word2bits.cpp:518-621).  Differences, all additive:
//   -threads N   number of corpus shards (one CUDA CTA each).  When the flag is absent the
//                shard count that fills the GPU is used instead of the reference's 12.
//   -gpu N       CUDA device ordinal (default 0).
//   -strict 1    parity mode: shards one after another, sequential IEEE arithmetic.
//   -binary 2    packed output: bitlevel bits per value (bitlevel 1 and 2), see w2b_write_packed.
//   -checkpoint F  write a resumable checkpoint (fp32 u, v, alpha, word counter) to F after every epoch.
//   -resume F      continue from checkpoint F at the epoch it was written after.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "w2b.h"

static int arg_pos(const char *str, int argc, char **argv) {  // ArgPos, :579-589
  for (int a = 1; a < argc; a++)
    if (!strcmp(str, argv[a])) {
      if (a == argc - 1) {
        printf("Argument missing for %s\n", str);
        exit(1);
      }
      return a;
    }
  return -1;
}

static void die(const char *what) {
  printf("%s: %s\n", what, w2b_last_error());
  exit(1);
}

static int g_bitlevel = 1;
static int write_vectors(const std::string &path, w2b_ctx *ctx, w2b_corpus *corpus, long long V, long long D,
                         int binary, std::vector<float> &buf) {
  if (w2b_export(ctx, buf.data())) return 1;
  if (binary == 2) return w2b_write_packed(path.c_str(), corpus, buf.data(), V, D, g_bitlevel);
  return w2b_write_vectors(path.c_str(), corpus, buf.data(), V, D, binary);
}

int main(int argc, char **argv) {
  int i;
  std::string train_file, output_file;
  int binary = 0, debug_mode = 2, window = 5, min_count = 5, num_threads = 0, bitlevel = 1, negative = 5;
  long long layer1_size = 100, iter = 5, classes = 0;
  bool save_every_epoch = false;
  float alpha = 0.05f, sample = 1e-3f, reg = 0;
  int device = 0, strict = 0;
  if ((i = arg_pos("-save-every-epoch", argc, argv)) > 0) save_every_epoch = atoi(argv[i + 1]);
  if ((i = arg_pos("-bitlevel", argc, argv)) > 0) bitlevel = atoi(argv[i + 1]);
  if ((i = arg_pos("-size", argc, argv)) > 0) layer1_size = atoi(argv[i + 1]);
  if ((i = arg_pos("-reg", argc, argv)) > 0) reg = atof(argv[i + 1]);
  if ((i = arg_pos("-train", argc, argv)) > 0) train_file = argv[i + 1];
  if ((i = arg_pos("-debug", argc, argv)) > 0) debug_mode = atoi(argv[i + 1]);
  if ((i = arg_pos("-binary", argc, argv)) > 0) binary = atoi(argv[i + 1]);
  if ((i = arg_pos("-alpha", argc, argv)) > 0) alpha = atof(argv[i + 1]);
  if ((i = arg_pos("-output", argc, argv)) > 0) output_file = argv[i + 1];
  if ((i = arg_pos("-window", argc, argv)) > 0) window = atoi(argv[i + 1]);
  if ((i = arg_pos("-sample", argc, argv)) > 0) sample = atof(argv[i + 1]);
  if ((i = arg_pos("-negative", argc, argv)) > 0) negative = atoi(argv[i + 1]);
  if ((i = arg_pos("-threads", argc, argv)) > 0) num_threads = atoi(argv[i + 1]);
  if ((i = arg_pos("-iter", argc, argv)) > 0) iter = atoi(argv[i + 1]);
  if ((i = arg_pos("-min-count", argc, argv)) > 0) min_count = atoi(argv[i + 1]);
  if ((i = arg_pos("-classes", argc, argv)) > 0) classes = atoi(argv[i + 1]);
  if ((i = arg_pos("-gpu", argc, argv)) > 0) device = atoi(argv[i + 1]);
  if ((i = arg_pos("-strict", argc, argv)) > 0) strict = atoi(argv[i + 1]);
  std::string ckpt_file, resume_file;
  if ((i = arg_pos("-checkpoint", argc, argv)) > 0) ckpt_file = argv[i + 1];
  if ((i = arg_pos("-resume", argc, argv)) > 0) resume_file = argv[i + 1];
  g_bitlevel = bitlevel;

  printf("Starting training using file %s\n", train_file.c_str());  // :523
  w2b_corpus *corpus = nullptr;
  if (w2b_corpus_load(train_file.c_str(), min_count, &corpus)) {
    printf("%s\n", w2b_last_error());  // "ERROR: training data file not found!" (:272)
    exit(1);
  }
  const long long V = w2b_corpus_vocab_size(corpus);
  const long long train_words = w2b_corpus_train_words(corpus);
  if (debug_mode > 0) {  // :295-298
    printf("Vocab size: %lld\n", V);
    printf("Words in train file: %lld\n", train_words);
  }
  if (output_file.empty()) return 0;  // :527

  w2b_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.vocab_size = V;
  cfg.layer1_size = layer1_size;
  cfg.window = window;
  cfg.negative = negative;
  cfg.bitlevel = bitlevel;
  cfg.alpha = alpha;
  cfg.sample = sample;
  cfg.reg = reg;
  cfg.iter = iter;
  cfg.device = device;
  cfg.mode = strict ? W2B_MODE_STRICT : W2B_MODE_FAST;
  cfg.num_shards = num_threads > 0 ? num_threads : 1;
  if (num_threads <= 0) {
    int s = 0;
    if (w2b_suggest_shards(&cfg, &s)) die("w2b_suggest_shards");
    // never cut the corpus into shards shorter than a few sentences
    while (s > 1 && train_words / s < 4000) s /= 2;
    cfg.num_shards = s;
  }
  w2b_ctx *ctx = nullptr;
  if (w2b_create(&cfg, &ctx)) {
    if (strstr(w2b_last_error(), "cudaMalloc")) printf("Memory allocation failed\n");  // :347
    die("w2b_create");
  }
  if (w2b_set_vocab_counts(ctx, w2b_corpus_counts(corpus), V, train_words)) die("w2b_set_vocab_counts");
  std::vector<int64_t> start(cfg.num_shards);
  std::vector<int32_t> first(cfg.num_shards);
  if (w2b_corpus_shards(corpus, cfg.num_shards, start.data(), first.data())) die("w2b_corpus_shards");
  if (w2b_set_corpus(ctx, w2b_corpus_tokens(corpus), w2b_corpus_num_tokens(corpus), start.data(), first.data(), 1))
    die("w2b_set_corpus");
  if (w2b_init_tables(ctx)) die("w2b_init_tables");

  std::vector<float> buf((size_t)V * layer1_size);
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  long long words_done = 0;
  long long first_epoch = 0;
  if (!resume_file.empty()) {
    int64_t done = 0;
    if (w2b_checkpoint_load(ctx, resume_file.c_str(), &done)) die("w2b_checkpoint_load");
    first_epoch = done;
  }
  for (int iteration = (int)first_epoch; iteration < iter; iteration++) {
    printf("Starting epoch: %d\n", iteration);  // :533
    if (w2b_epoch_begin(ctx)) die("w2b_epoch_begin");
    double epoch_loss = 0;
    for (;;) {
      w2b_step_stats st;
      // 50k words per shard per step: progress lines 5x less often than the reference (:379), steps long
      // enough that the whole-sentence overshoot at a step boundary (<= ~1.2k words) stays ~1 %
      if (w2b_train_step(ctx, debug_mode > 1 ? 50000 : 0, &st)) die("w2b_train_step");
      epoch_loss += st.loss;
      words_done += st.words;
      if (debug_mode > 1) {  // :384-387 (Words/sec here is wall-clock and whole-job, not per CPU thread)
        struct timespec now;
        clock_gettime(CLOCK_MONOTONIC, &now);
        double secs = (now.tv_sec - t0.tv_sec) + (now.tv_nsec - t0.tv_nsec) * 1e-9;
        printf("%cAlpha: %f  Progress: %.2f%%  Cost: %f Words/sec: %.2fk  ", 13, st.alpha,
               st.word_count_actual / (float)(iter * train_words + 1) * 100, st.loss,
               words_done / (secs + 1e-9) / 1000);
        fflush(stdout);
      }
      if (st.shards_done >= cfg.num_shards) break;
    }
    printf("Epoch Loss: %lf\n", epoch_loss);  // :539
    if (!ckpt_file.empty() && w2b_checkpoint_save(ctx, ckpt_file.c_str(), iteration + 1)) die("w2b_checkpoint_save");
    if (classes == 0 && save_every_epoch) {     // :540-557
      char name[4200];
      snprintf(name, sizeof name, "%s_epoch%d", output_file.c_str(), iteration);
      if (write_vectors(name, ctx, corpus, V, layer1_size, binary, buf)) die("write");
    }
  }
  if (classes == 0) {  // :560-576
    if (write_vectors(output_file, ctx, corpus, V, layer1_size, binary, buf)) die("write");
  } else {
    FILE *fo = fopen(output_file.c_str(), "wb");  // the reference creates an empty file (:561-562)
    if (fo) fclose(fo);
  }
  w2b_destroy(ctx);
  w2b_corpus_free(corpus);
  return 0;
}
