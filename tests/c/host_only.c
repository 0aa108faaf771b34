/* The C-ABI from plain C99 (-Wall -Wextra -pedantic), restricted to the entry points that need no device: header validity,
 * linkage, pco_standalone_guarantee_file_size (pco_c/include/cpcodec_generated.h:27-31) and pco_b200_choose_mode.
 * Built and run by tests/test_c_host_only.py. */
#include <stdio.h>
#include <string.h>

#include "cpcodec.h"
#include "pco_b200.h"

int main(void) {
  PcoB200ModeChoice m;
  unsigned int mult[3000];
  double dec[3000];
  size_t i;
  unsigned long long state = 88172645463325252ULL;
  for (i = 0; i < 3000; i++) { /* xorshift: any spread-out multipliers do */
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    mult[i] = (unsigned int)((state >> 20) % 1000000u) * 77u;
    dec[i] = (double)((state >> 24) % 100000u) / 100.0;
  }
  if (pco_standalone_guarantee_file_size(0, PCO_TYPE_U32) == 0) return 1;
  if (pco_standalone_guarantee_file_size(1000, PCO_TYPE_U64) < 8000) return 2;
  if (pco_b200_choose_mode(mult, 3000, PCO_TYPE_U32, &m) != PCO_B200_OK) return 3;
  if (m.mode_spec != PCO_B200_MODE_TRY_INT_MULT || m.int_mult_base != 77) return 4;
  if (pco_b200_choose_mode(dec, 3000, PCO_TYPE_F64, &m) != PCO_B200_OK) return 5;
  if (m.mode_spec != PCO_B200_MODE_TRY_FLOAT_MULT || m.float_mult_inv_base != 100.0) return 6;
  if (pco_b200_choose_mode(dec, 5, PCO_TYPE_F64, &m) != PCO_B200_OK || m.mode_spec != PCO_B200_MODE_CLASSIC) return 7;
  if (pco_b200_choose_mode(dec, 5, 99, &m) == PCO_B200_OK) return 8;
  printf("HOST_ONLY_OK\n");
  return 0;
}
