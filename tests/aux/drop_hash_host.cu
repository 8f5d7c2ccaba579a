// Host-side print-out of the dropout mask hash of csrc/common.cuh (the functions are __host__ __device__), used by
// tests/test_dropout_hash.py to pin the numpy restatement in tests/test_dropout_gpu.py to the C++ implementation.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

int main(int argc, char** argv) {
  const unsigned seed = (unsigned)strtoul(argv[1], nullptr, 0), site = (unsigned)strtoul(argv[2], nullptr, 0);
  const unsigned rows = (unsigned)strtoul(argv[3], nullptr, 0), cols = (unsigned)strtoul(argv[4], nullptr, 0);
  const unsigned thresh = (unsigned)strtoul(argv[5], nullptr, 0);
  const unsigned ss = s3d::drop_site_seed(seed, site);
  for (unsigned r = 0; r < rows; ++r) {
    for (unsigned c = 0; c < cols; ++c) putchar(s3d::drop_keep(ss, r * 7919u + 3u, c, thresh) ? '1' : '0');
    putchar('\n');
  }
  return 0;
}
