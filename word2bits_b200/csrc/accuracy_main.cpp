// compute_accuracy — drop-in for the reference's evaluator (src/compute-accuracy.c), scored on the GPU.
//   ./compute_accuracy <FILE> <bitlevel> <threshold> < questions-words.txt
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "w2b.h"

int main(int argc, char **argv) {
  if (argc < 2) {  // :74-77
    printf("Usage: ./compute-accuracy <FILE> <bitlevel> <threshold>\nwhere FILE contains word projections, and threshold is used to reduce vocabulary of the model for fast approximate evaluation (0 = off, otherwise typical value is 30000)\n");
    return 0;
  }
  const int bitlevel = argc > 2 ? atoi(argv[2]) : 0;
  const long long threshold = argc > 3 ? atoll(argv[3]) : 0;
  std::vector<char> report(1 << 20);
  w2b_accuracy acc;
  const int rc = w2b_compute_accuracy(argv[1], bitlevel, threshold, nullptr, 0, &acc, report.data(), (long long)report.size());
  if (rc) {
    printf("%s\n", w2b_last_error());  // "Input file not found" (:83)
    return -1;
  }
  fputs(report.data(), stdout);
  return 0;
}
