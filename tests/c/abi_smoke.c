/* The C ABI from plain C99 (no Python, no C++): header validity, error codes, planning, the .srk round trip.
   Built and run by tests/test_abi.py::test_abi_from_plain_c; no device is touched. */
#include "srack_b200.h"
#include <stdio.h>
#include <string.h>
int main(void) {
  srk_audio_config cfg = {48000, 1024, 2};
  srk_patch* p = NULL;
  if (srk_patch_create(&cfg, &p) != SRK_OK) return 1;
  srk_module *osc = NULL, *out = NULL, *smp = NULL;
  if (srk_module_create_by_name(p, "Oscillator", &osc) || srk_module_create(p, SRK_KIND_OUTPUT, &out) ||
      srk_module_create_by_name(p, "Sample", &smp)) return 2;
  if (srk_connect(out, 0, osc, 0) != SRK_OK) return 3;
  if (srk_connect(out, 2, osc, 0) != SRK_ERR_PORT) return 4;        /* the reference's Err(()) */
  if (srk_connect(osc, 0, osc, 0) != SRK_ERR_SELF_LOOP) return 5;
  if (srk_plan(p) != SRK_OK) return 6;
  size_t n = 0;
  srk_module* plan[8];
  if (srk_plan_get(p, plan, 8, &n) != SRK_OK || n != 3 || plan[0] != osc || plan[1] != out || plan[2] != smp) return 7; /* emission order, synth.rs:193-211 */
  float wave[4] = {0.f, .5f, -.5f, 1.f};
  if (srk_set_sample(smp, wave, 4, 8000.f) != SRK_OK) return 8;
  const void* bytes = NULL; size_t nb = 0;
  if (srk_patch_save_srk(p, &bytes, &nb) != SRK_OK || nb < 100) return 9;
  srk_patch* q = NULL;
  size_t skipped = 99;
  if (srk_patch_create(&cfg, &q) || srk_patch_load_srk(q, bytes, nb, &skipped) != SRK_OK || skipped != 0) return 10;
  if (srk_module_count(q) != 3 || strcmp(srk_get_name(srk_module_at(q, 0)), "Sample") != 0) return 11;  /* reversed list */
  printf("%s: ok (%zu bytes of .srk)\n", srk_version(), nb);
  srk_patch_destroy(q);
  srk_patch_destroy(p);
  return 0;
}
