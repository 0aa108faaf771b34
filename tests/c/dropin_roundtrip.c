/* A C caller of the reference's C ABI (pco_c/include/cpcodec_generated.h:33-64; this repo's include/cpcodec.h keeps the three
 * functions verbatim): guarantee -> compress into a caller buffer -> decompress into a caller buffer -> compare, for f64, i32
 * and u64, plus the two error cases the reference's own C test checks (bad dtype, destination too small).
 * Exit code 0: all round trips bit-exact.  3: the library reported an error on the first compress (no CUDA device: the
 * library has no CPU fallback).  1: anything else. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpcodec.h"

static int roundtrip(const void *nums, size_t n, unsigned char dtype, size_t elem, int *first_error) {
  size_t bound = pco_standalone_guarantee_file_size(n, dtype);
  if (bound == 0) { printf("FAIL: guarantee returned 0\n"); return 1; }
  unsigned char *cbuf = (unsigned char *)malloc(bound);
  void *dbuf = malloc(n * elem + 8);
  struct PcoChunkConfig config;
  config.compression_level = 8;
  config.max_page_n = 0; /* library default */
  size_t clen = 0, dn = 0;
  enum PcoError rc = pco_standalone_simple_compress_into(nums, n, dtype, &config, cbuf, bound, &clen);
  if (rc != PcoSuccess) {
    printf("compress_into error %d\n", (int)rc);
    if (first_error) *first_error = (int)rc;
    free(cbuf); free(dbuf);
    return 3;
  }
  rc = pco_standalone_simple_decompress_into(cbuf, clen, dtype, dbuf, n, &dn);
  int bad = rc != PcoSuccess || dn != n || memcmp(dbuf, nums, n * elem) != 0;
  printf("dtype %u: %zu numbers -> %zu bytes -> %zu numbers, rc %d: %s\n", (unsigned)dtype, n, clen, dn, (int)rc, bad ? "FAIL" : "ok");
  if (!bad) {
    /* destination too small -> PcoDecompressionError (pco_c/src/lib.rs:110-112) */
    rc = pco_standalone_simple_decompress_into(cbuf, clen, dtype, dbuf, n - 1, &dn);
    if (rc != PcoDecompressionError) { printf("FAIL: short destination gave %d\n", (int)rc); bad = 1; }
    /* compressed buffer too small -> PcoCompressionError */
    rc = pco_standalone_simple_compress_into(nums, n, dtype, &config, cbuf, 8, &clen);
    if (rc != PcoCompressionError) { printf("FAIL: short compressed buffer gave %d\n", (int)rc); bad = 1; }
  }
  free(cbuf); free(dbuf);
  return bad ? 1 : 0;
}

int main(void) {
  double f[] = {10.0, 20.0, 30.0, 40.0, 50.0, 60.0};
  int i32[1000];
  unsigned long long u64[5000];
  for (int i = 0; i < 1000; i++) i32[i] = (i * 37) % 101 - 50;
  unsigned long long acc = 1000;
  for (int i = 0; i < 5000; i++) { acc += (unsigned long long)((i * 2654435761u) % 97); u64[i] = acc; }
  if (pco_standalone_guarantee_file_size(10, 99) != 0) { printf("FAIL: unknown dtype must give 0\n"); return 1; }
  size_t nw = 0;
  unsigned char tmp[64];
  if (pco_standalone_simple_compress_into(f, 6, 99, NULL, tmp, sizeof(tmp), &nw) != PcoInvalidType) { printf("FAIL: bad dtype\n"); return 1; }
  int first_error = 0;
  int rc = roundtrip(f, 6, PCO_TYPE_F64, sizeof(double), &first_error);
  if (rc == 3) return 3;
  if (rc) return 1;
  if (roundtrip(i32, 1000, PCO_TYPE_I32, sizeof(int), NULL)) return 1;
  if (roundtrip(u64, 5000, PCO_TYPE_U64, sizeof(unsigned long long), NULL)) return 1;
  printf("all ok\n");
  return 0;
}
