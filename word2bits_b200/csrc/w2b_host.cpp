// libw2b host side (no CUDA): corpus reader + vocabulary, shard-start resolution,
// unigram boundaries, expTable, vector-file writer.  These replace the reference's
// host glue (src/word2bits.cpp:112-301, :560-576, :614-618) with identical results;
// the text is read once through mmap and tokenised into an int32 id stream instead of
// being re-parsed with fgetc by every thread in every epoch (:396).
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <numeric>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "w2b.h"
#include "w2b_internal.h"

void w2b_unigram_bounds(const int64_t *cn, int64_t V, int32_t *start) {
  const double N = (double)W2B_TABLE_SIZE;
  double total = 0;
  for (int64_t a = 0; a < V; ++a) total += pow((double)cn[a], 0.75);
  double d1 = pow((double)cn[0], 0.75) / total;  // cumulative share of words 0..i
  int64_t a = 0;                                  // next table slot to examine
  start[0] = 0;
  int64_t i = 0;
  for (; i + 1 < V; ++i) {
    // The reference assigns slot a to word i and only then tests a/1e8 > d1 (:121-125):
    // word i keeps every slot up to and including the first one that passes the test.
    // Jump close with the closed form, then settle with the reference's own comparison.
    const int64_t guess = (int64_t)(d1 * N) - 2;  // the test is monotone in a
    if (guess > a) a = guess;
    while (a < W2B_TABLE_SIZE && !((double)a / N > d1)) ++a;
    if (a >= W2B_TABLE_SIZE) break;
    ++a;
    start[i + 1] = (int32_t)a;
    d1 += pow((double)cn[i + 1], 0.75) / total;
  }
  for (int64_t j = i + 1; j <= V; ++j) start[j] = W2B_TABLE_SIZE;
}

void w2b_exptable(float *out) {
  for (int i = 0; i < 1000; ++i) {
    float e = expf((i / (float)1000 * 2 - 1) * 6);
    out[i] = e / (e + 1);
  }
}

// Sub-sampling threshold `ran` of every word (:403-404), float32 throughout as in the reference:
// (sqrt(cn / (sample*train_words)) + 1) * (sample*train_words) / cn.
void w2b_keep_thresholds(const int64_t *cn, int64_t V, int64_t train_words, float sample, float *out) {
  const float S = sample * (float)train_words;
  for (int64_t w = 0; w < V; ++w) out[w] = (sqrtf((float)cn[w] / S) + 1.f) * S / (float)cn[w];
}

extern "C" int w2b_host_unigram_bounds(const int64_t *cn, int64_t V, int32_t *start) {
  if (!cn || !start || V < 1) { w2b_set_error("w2b_host_unigram_bounds: bad argument"); return W2B_EINVAL; }
  w2b_unigram_bounds(cn, V, start);
  return W2B_OK;
}
extern "C" int w2b_host_exptable(float *out) {
  if (!out) { w2b_set_error("w2b_host_exptable: null argument"); return W2B_EINVAL; }
  w2b_exptable(out);
  return W2B_OK;
}
extern "C" int w2b_host_keep_thresholds(const int64_t *cn, int64_t V, int64_t train_words, float sample, float *out) {
  if (!cn || !out || V < 1) { w2b_set_error("w2b_host_keep_thresholds: bad argument"); return W2B_EINVAL; }
  w2b_keep_thresholds(cn, V, train_words, sample, out);
  return W2B_OK;
}

void w2b_gather_slices(const int32_t *ids, long long n_tokens, long long L, int nshards, const long long *cursor,
                       const int *done, int32_t *stage, long long *xlate, long long *limit, int *limit_is_eof,
                       int nthreads) {
  auto gather = [=](int lo, int hi) {  // shard i only touches its own outputs and its own slice of the staging buffer
    for (int i = lo; i < hi; ++i) {
      if (done[i]) continue;
      const long long b = std::max<long long>(cursor[i], 0);  // cursor -1 = pending override token
      const long long e = std::min<long long>(b + L, n_tokens);
      if (e > b) memcpy(stage + (long long)i * L, ids + b, (size_t)(e - b) * sizeof(int32_t));
      xlate[i] = b - (long long)i * L;
      limit[i] = e;
      limit_is_eof[i] = (e == n_tokens);
    }
  };
  if (nthreads <= 0) {  // >= 1 M tokens per thread, at most 8 threads / half the host / one per shard
    const long long by_size = L * nshards / (1 << 20);
    const long long by_host = std::max(1u, std::thread::hardware_concurrency() / 2);
    nthreads = (int)std::min<long long>(std::min<long long>(8, by_host), std::min<long long>(nshards, by_size));
  }
  nthreads = std::min(nthreads, nshards);
  if (nthreads <= 1) {
    gather(0, nshards);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t)
    th.emplace_back(gather, (int)((long long)nshards * t / nthreads), (int)((long long)nshards * (t + 1) / nthreads));
  gather(0, (int)((long long)nshards / nthreads));
  for (auto &x : th) x.join();
}

extern "C" int w2b_host_gather_slices(const int32_t *ids, int64_t n_tokens, int64_t L, int nshards, const int64_t *cursor,
                                      const int32_t *done, int32_t *stage, int64_t *xlate, int64_t *limit,
                                      int32_t *limit_is_eof, int nthreads) {
  if (!ids || !cursor || !done || !stage || !xlate || !limit || !limit_is_eof || L < 1 || nshards < 1 || n_tokens < 0) {
    w2b_set_error("w2b_host_gather_slices: bad argument");
    return W2B_EINVAL;
  }
  std::vector<long long> cur(cursor, cursor + nshards), xl(nshards, 0), lim(nshards, 0);
  std::vector<int> dn(done, done + nshards), eof(nshards, 0);
  w2b_gather_slices(ids, n_tokens, L, nshards, cur.data(), dn.data(), stage, xl.data(), lim.data(), eof.data(), nthreads);
  for (int i = 0; i < nshards; ++i) {
    if (dn[i]) continue;
    xlate[i] = xl[i];
    limit[i] = lim[i];
    limit_is_eof[i] = eof[i];
  }
  return W2B_OK;
}

// ----------------------------------------------------------------------------- tokeniser
namespace {

constexpr int kMaxWord = 4096;  // MAX_STRING :29
constexpr int64_t kCkptEvery = 4096;

// One ReadWord call (:131-155) over the mapped file.  false at EOF; a word cut short by
// EOF is dropped, as both callers of the reference do (:279, :180).
inline bool next_token(const uint8_t *buf, int64_t n, int64_t &pos, char *word, int &len, int64_t &begin) {
  int a = 0;
  while (pos < n) {
    const int ch = buf[pos++];
    if (ch == 13) continue;
    if (ch == ' ' || ch == '\t' || ch == '\n') {
      if (a > 0) {
        if (ch == '\n') --pos;  // the newline is read again as </s>
        word[a] = 0;
        len = (int)strlen(word);  // an embedded NUL ends the C string, as in the reference
        return true;
      }
      if (ch == '\n') {
        memcpy(word, "</s>", 5);
        len = 4;
        begin = pos - 1;
        return true;
      }
      continue;
    }
    if (a == 0) begin = pos - 1;
    word[a++] = (char)ch;
    if (a >= kMaxWord - 1) --a;
  }
  return false;
}

inline uint64_t hash_bytes(const char *s, int len) {
  uint64_t h = 1469598103934665603ULL;
  for (int i = 0; i < len; ++i) h = (h ^ (uint8_t)s[i]) * 1099511628211ULL;
  return h;
}

struct WordMap {
  std::vector<uint32_t> slot;   // entry index + 1, 0 = empty
  std::vector<uint64_t> off;    // arena offset per entry
  std::vector<uint32_t> wlen;
  std::vector<int64_t> count;
  std::string arena;
  uint64_t mask = 0;

  WordMap() { slot.assign(1u << 16, 0); mask = slot.size() - 1; }
  const char *str(uint32_t e) const { return arena.data() + off[e]; }
  void grow() {
    std::vector<uint32_t> ns(slot.size() * 2, 0);
    const uint64_t m = ns.size() - 1;
    for (uint32_t e = 0; e < off.size(); ++e) {
      uint64_t h = hash_bytes(str(e), (int)wlen[e]) & m;
      while (ns[h]) h = (h + 1) & m;
      ns[h] = e + 1;
    }
    slot.swap(ns);
    mask = m;
  }
  int64_t find(const char *w, int len) const {
    uint64_t h = hash_bytes(w, len) & mask;
    while (slot[h]) {
      const uint32_t e = slot[h] - 1;
      if ((int)wlen[e] == len && !memcmp(str(e), w, len)) return e;
      h = (h + 1) & mask;
    }
    return -1;
  }
  uint32_t insert(const char *w, int len) {
    if ((off.size() + 1) * 2 > slot.size()) grow();
    const uint32_t e = (uint32_t)off.size();
    off.push_back(arena.size());
    wlen.push_back((uint32_t)len);
    count.push_back(0);
    arena.append(w, len);
    arena.push_back('\0');
    uint64_t h = hash_bytes(w, len) & mask;
    while (slot[h]) h = (h + 1) & mask;
    slot[h] = e + 1;
    return e;
  }
};

}  // namespace

struct w2b_corpus {
  const uint8_t *buf = nullptr;
  int64_t file_size = 0;
  int fd = -1;
  WordMap map;
  std::vector<int32_t> final_id;  // map entry -> vocab id or -1
  std::vector<const char *> words;
  std::vector<int64_t> cn;
  int64_t train_words = 0;
  std::vector<int32_t> ids;
  // every kCkptEvery raw tokens: byte offset of the token and #in-vocab tokens before it
  std::vector<int64_t> ck_begin, ck_comp;
};

// Pass 1 over one chunk [begin, end) of the mapped file: chunk-local vocabulary in first-appearance
// order + the chunk's tokens as local ids.  Chunks start right after a whitespace byte, so the
// sequential reader would be in the same (empty-word) state there.
struct ChunkResult {
  WordMap map;
  std::vector<uint32_t> raw;        // local entry id per token
  std::vector<int64_t> ck_begin;    // byte offset of every kCkptEvery-th token of the chunk
};

static void tokenize_chunk(const uint8_t *buf, int64_t begin, int64_t end, ChunkResult *out) {
  out->raw.reserve((size_t)((end - begin) / 5 + 16));
  char word[kMaxWord];
  int len = 0;
  int64_t pos = begin, tb = 0;
  while (next_token(buf, end, pos, word, len, tb)) {
    int64_t e = out->map.find(word, len);
    if (e < 0) e = out->map.insert(word, len);
    out->map.count[e]++;
    if ((int64_t)out->raw.size() % kCkptEvery == 0) out->ck_begin.push_back(tb);
    out->raw.push_back((uint32_t)e);
  }
}

extern "C" int w2b_corpus_load(const char *path, int min_count, w2b_corpus **out) {
  if (!out) {
    w2b_set_error("w2b_corpus_load: null out");
    return W2B_EINVAL;
  }
  *out = nullptr;
  int fd = path ? open(path, O_RDONLY) : -1;
  if (fd < 0) {
    w2b_set_error("ERROR: training data file not found!");  // :272
    return W2B_EIO;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {  // a directory opens fine but cannot be mapped
    close(fd);
    w2b_set_error("ERROR: training data file not found!");
    return W2B_EIO;
  }
  w2b_corpus *c = new w2b_corpus();
  c->fd = fd;
  c->file_size = st.st_size;  // == ftell at EOF, :299
  if (st.st_size > 0) {
    void *m = mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      close(fd);
      delete c;
      w2b_set_error("mmap failed for %s", path);
      return W2B_EIO;
    }
    c->buf = (const uint8_t *)m;
  }
  const bool dbg = getenv("W2B_TOKENIZER_DEBUG") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!dbg) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[tokenizer] %-10s %.3f s\n", what, std::chrono::duration<double>(now - t_start).count());
    t_start = now;
  };
  // ---- pass 1, parallel over chunks of the file (the reference's single fgetc loop, :277-293,
  // runs at ~7 M words/s; the GPU consumes 30 M words/s)
  const int64_t n = c->file_size;
  int nthreads = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("W2B_TOKENIZER_THREADS")) nthreads = atoi(e);
  int64_t min_chunk = 4 << 20;
  if (const char *e = getenv("W2B_TOKENIZER_MIN_CHUNK")) min_chunk = atoll(e);
  nthreads = std::max(1, std::min(nthreads, 64));
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, n / std::max<int64_t>(min_chunk, 1)));
  std::vector<int64_t> cut(nthreads + 1, n);
  cut[0] = 0;
  for (int t = 1; t < nthreads; ++t) {
    int64_t p = n / nthreads * t;
    if (p < cut[t - 1]) p = cut[t - 1];
    // advance to just after the next whitespace byte: a clean token boundary
    while (p < n && !(c->buf[p] == ' ' || c->buf[p] == '\t' || c->buf[p] == '\n')) ++p;
    cut[t] = p < n ? p + 1 : n;
  }
  std::vector<ChunkResult> chunks(nthreads);
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t)
      th.emplace_back(tokenize_chunk, c->buf, cut[t], cut[t + 1], &chunks[t]);
    tokenize_chunk(c->buf, cut[0], cut[1], &chunks[0]);
    for (auto &x : th) x.join();
  }
  lap("pass1");
  // ---- merge in file order: global first-appearance order = chunk order, then local order; </s> first (:276)
  WordMap &map = c->map;
  map.insert("</s>", 4);
  std::vector<std::vector<uint32_t>> l2g(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    const WordMap &lm = chunks[t].map;
    l2g[t].resize(lm.off.size());
    for (uint32_t e = 0; e < lm.off.size(); ++e) {
      int64_t g = map.find(lm.str(e), (int)lm.wlen[e]);
      if (g < 0) g = map.insert(lm.str(e), (int)lm.wlen[e]);
      map.count[g] += lm.count[e];
      l2g[t][e] = (uint32_t)g;
    }
  }
  lap("merge");
  // SortVocab (:215-242): </s> pinned at 0, the rest by count descending, ties in
  // first-appearance order (what glibc's qsort yields here; asserted against the
  // reference in tests), then the min_count cut.
  const size_t m = map.off.size();
  std::vector<uint32_t> order(m);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin() + 1, order.end(),
                   [&](uint32_t a, uint32_t b) { return map.count[a] > map.count[b]; });
  c->final_id.assign(m, -1);
  for (size_t k = 0; k < m; ++k) {
    const uint32_t e = order[k];
    if (map.count[e] < min_count && k != 0) continue;
    c->final_id[e] = (int32_t)c->words.size();
    c->words.push_back(nullptr);
    c->cn.push_back(map.count[e]);
    c->train_words += map.count[e];
  }
  for (size_t e = 0; e < m; ++e)
    if (c->final_id[e] >= 0) c->words[c->final_id[e]] = map.str((uint32_t)e);
  lap("sort");
  // ---- compact to the in-vocab stream (parallel per chunk), remembering where every checkpoint lands
  std::vector<int64_t> kept(nthreads, 0), base(nthreads + 1, 0);
  {
    auto count_kept = [&](int t) {
      int64_t k = 0;
      for (uint32_t le : chunks[t].raw) k += c->final_id[l2g[t][le]] >= 0;
      kept[t] = k;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(count_kept, t);
    count_kept(0);
    for (auto &x : th) x.join();
  }
  for (int t = 0; t < nthreads; ++t) base[t + 1] = base[t] + kept[t];
  c->ids.resize((size_t)base[nthreads]);
  std::vector<std::vector<int64_t>> ck_comp(nthreads);
  {
    auto fill = [&](int t) {
      int64_t w = base[t];
      const auto &raw = chunks[t].raw;
      ck_comp[t].reserve(chunks[t].ck_begin.size());
      for (size_t k = 0; k < raw.size(); ++k) {
        if ((int64_t)k % kCkptEvery == 0) ck_comp[t].push_back(w);
        const int32_t id = c->final_id[l2g[t][raw[k]]];
        if (id >= 0) c->ids[(size_t)w++] = id;
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(fill, t);
    fill(0);
    for (auto &x : th) x.join();
  }
  for (int t = 0; t < nthreads; ++t) {
    c->ck_begin.insert(c->ck_begin.end(), chunks[t].ck_begin.begin(), chunks[t].ck_begin.end());
    c->ck_comp.insert(c->ck_comp.end(), ck_comp[t].begin(), ck_comp[t].end());
  }
  lap("compact");
  *out = c;
  return W2B_OK;
}

extern "C" void w2b_corpus_free(w2b_corpus *c) {
  if (!c) return;
  if (c->buf) munmap((void *)c->buf, c->file_size);
  if (c->fd >= 0) close(c->fd);
  delete c;
}

extern "C" int64_t w2b_corpus_vocab_size(const w2b_corpus *c) { return (int64_t)c->words.size(); }
extern "C" int64_t w2b_corpus_train_words(const w2b_corpus *c) { return c->train_words; }
extern "C" int64_t w2b_corpus_file_size(const w2b_corpus *c) { return c->file_size; }
extern "C" const char *w2b_corpus_word(const w2b_corpus *c, int64_t i) { return c->words[i]; }
extern "C" const int64_t *w2b_corpus_counts(const w2b_corpus *c) { return c->cn.data(); }
extern "C" int64_t w2b_corpus_num_tokens(const w2b_corpus *c) { return (int64_t)c->ids.size(); }
extern "C" const int32_t *w2b_corpus_tokens(const w2b_corpus *c) { return c->ids.data(); }

extern "C" int w2b_corpus_shards(const w2b_corpus *c, int n, int64_t *start, int32_t *first) {
  if (!c || !start || !first) {
    w2b_set_error("w2b_corpus_shards: null argument");
    return W2B_EINVAL;
  }
  if (n < 1) {
    w2b_set_error("shard count must be >= 1");
    return W2B_EINVAL;
  }
  char word[kMaxWord];
  for (int i = 0; i < n; ++i) {
    const int64_t off = c->file_size / (int64_t)n * (int64_t)i;  // :377
    int64_t pos = off, begin = 0;
    int len = 0;
    first[i] = -1;
    if (!next_token(c->buf, c->file_size, pos, word, len, begin)) {
      start[i] = (int64_t)c->ids.size();
      continue;
    }
    const int64_t e = c->map.find(word, len);
    if (e >= 0) first[i] = c->final_id[e];
    // index of the first regular in-vocab token that begins at or after `pos`:
    // restart from the last checkpoint at or before it and count forward
    size_t k = std::upper_bound(c->ck_begin.begin(), c->ck_begin.end(), pos) - c->ck_begin.begin();
    if (k == 0) {
      start[i] = 0;
      continue;
    }
    --k;
    int64_t p2 = c->ck_begin[k], comp = c->ck_comp[k], b2 = 0;
    int l2 = 0;
    for (;;) {
      if (!next_token(c->buf, c->file_size, p2, word, l2, b2)) break;
      if (b2 >= pos) break;
      const int64_t e2 = c->map.find(word, l2);
      if (e2 >= 0 && c->final_id[e2] >= 0) ++comp;
    }
    start[i] = comp;
  }
  return W2B_OK;
}

extern "C" int w2b_write_vectors(const char *path, const w2b_corpus *c, const float *vec, int64_t V, int64_t D,
                                 int binary) {
  if (!path || !c || !vec || V < 0 || V > (int64_t)c->words.size() || D < 1) {
    w2b_set_error("w2b_write_vectors: bad argument");
    return W2B_EINVAL;
  }
  FILE *fo = fopen(path, "wb");
  if (!fo) {
    w2b_set_error("cannot open %s for writing", path);
    return W2B_EIO;
  }
  std::vector<char> iobuf(1 << 20);  // per call: two writers may run on two host threads
  setvbuf(fo, iobuf.data(), _IOFBF, iobuf.size());
  fprintf(fo, "%lld %lld\n", (long long)V, (long long)D);
  for (int64_t a = 0; a < V; ++a) {
    fprintf(fo, "%s ", c->words[a]);
    const float *row = vec + a * D;
    if (binary) fwrite(row, sizeof(float), D, fo);
    else
      for (int64_t b = 0; b < D; ++b) fprintf(fo, "%lf ", row[b]);
    fprintf(fo, "\n");
  }
  const bool bad = ferror(fo) != 0;
  if (fclose(fo) != 0 || bad) {  // the reference ignores write errors; a truncated vector file is worse
    w2b_set_error("short write to %s (disk full?)", path);
    return W2B_EIO;
  }
  return W2B_OK;
}

// ------------------------------------------------------------------------- packed vector files
static inline int level_code(float x, int bits) {  // inverse of quantize() for bitlevel 1 / 2
  const int neg = x < 0.f;
  if (bits == 1) return neg;
  return neg | ((fabsf(x) > 0.5f) ? 2 : 0);
}
static inline float level_value(int code, int bits) {
  const float m = (bits == 1) ? (1.0f / 3) : ((code & 2) ? 0.75f : 0.25f);
  return (code & 1) ? -m : m;
}

extern "C" int w2b_write_packed(const char *path, const w2b_corpus *c, const float *vec, int64_t V, int64_t D,
                                int bitlevel) {
  if (bitlevel != 1 && bitlevel != 2) {
    w2b_set_error("packed format supports bitlevel 1 and 2");
    return W2B_EINVAL;
  }
  if (!path || !c || !vec || V < 0 || V > (int64_t)c->words.size() || D < 1) {
    w2b_set_error("w2b_write_packed: bad argument");
    return W2B_EINVAL;
  }
  FILE *fo = fopen(path, "wb");
  if (!fo) {
    w2b_set_error("cannot open %s for writing", path);
    return W2B_EIO;
  }
  fprintf(fo, "%lld %lld %d\n", (long long)V, (long long)D, bitlevel);
  const int64_t nbytes = (D * bitlevel + 7) / 8;
  std::vector<uint8_t> row(nbytes);
  for (int64_t a = 0; a < V; ++a) {
    fprintf(fo, "%s ", c->words[a]);
    std::fill(row.begin(), row.end(), 0);
    for (int64_t j = 0; j < D; ++j) {
      const int code = level_code(vec[a * D + j], bitlevel);
      const int64_t bit = j * bitlevel;
      row[bit >> 3] |= (uint8_t)(code << (bit & 7));  // bitlevel divides 8: a value never straddles bytes
    }
    fwrite(row.data(), 1, nbytes, fo);
    fputc('\n', fo);
  }
  const bool bad = ferror(fo) != 0;
  if (fclose(fo) != 0 || bad) {
    w2b_set_error("short write to %s (disk full?)", path);
    return W2B_EIO;
  }
  return W2B_OK;
}

extern "C" int w2b_read_packed_header(const char *path, int64_t *V, int64_t *D, int *bitlevel) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    w2b_set_error("cannot open %s", path);
    return W2B_EIO;
  }
  long long v = 0, d = 0;
  int b = 0;
  const int n = fscanf(f, "%lld %lld %d", &v, &d, &b);
  fclose(f);
  if (n != 3 || (b != 1 && b != 2) || v < 0 || d < 1) {
    w2b_set_error("%s is not a packed vector file", path);
    return W2B_EIO;
  }
  *V = v; *D = d; *bitlevel = b;
  return W2B_OK;
}

extern "C" int w2b_read_packed(const char *path, float *vec, char *words, int max_word) {
  int64_t V, D;
  int bits;
  int rc = w2b_read_packed_header(path, &V, &D, &bits);
  if (rc) return rc;
  FILE *f = fopen(path, "rb");
  if (!f) {
    w2b_set_error("cannot open %s", path);
    return W2B_EIO;
  }
  int ch;
  while ((ch = fgetc(f)) != EOF && ch != '\n') {}
  const int64_t nbytes = (D * bits + 7) / 8;
  std::vector<uint8_t> row(nbytes);
  for (int64_t a = 0; a < V; ++a) {
    int k = 0;
    while ((ch = fgetc(f)) != EOF && ch != ' ')
      if (words && k < max_word - 1) words[a * max_word + k++] = (char)ch;
    if (words) words[a * max_word + k] = 0;
    if (fread(row.data(), 1, nbytes, f) != (size_t)nbytes) {
      fclose(f);
      w2b_set_error("%s is truncated", path);
      return W2B_EIO;
    }
    for (int64_t j = 0; j < D; ++j) {
      const int64_t bit = j * bits;
      vec[a * D + j] = level_value((row[bit >> 3] >> (bit & 7)) & ((1 << bits) - 1), bits);
    }
    fgetc(f);  // '\n'
  }
  fclose(f);
  return W2B_OK;
}
