// mask_test.cu — host-only exhaustive check of mask.cuh: for every mode (the four pre-training masks of
// data/dataset_origin.py:138-176 and the two fine-tune variants of .../sc/data_loader.py:394-408), every (A, t_len) of a
// small layout and every tile size, the per-row interval and the tile predicates the kernels use for skipping must agree
// with a brute-force evaluation of mask_allowed().  Runs on the CPU (no GPU needed): ./tests/mask_test
#include <cstdio>

#include "../mask.cuh"

using namespace mv;

int main() {
  long checked = 0, bad = 0;
  const int modes[6] = {MODE_BIDIR, MODE_S2S, MODE_BAR, MODE_NONCROSS, MODE_S2S_FT, MODE_BAR_FT};
  for (int L = 9; L <= 41; L += 8)
    for (int A = 2; A < L - 1; A += 3)
      for (int tl = 1; tl <= L - A; ++tl)
        for (int mi = 0; mi < 6; ++mi) {
          const int mode = modes[mi];
          // per-row interval == the set of allowed keys
          for (int q = 0; q < L; ++q) {
            int lo, hi;
            mask_row_interval(mode, q, A, tl, L, lo, hi);
            for (int k = 0; k < L; ++k) {
              const bool in = k >= lo && k < hi;
              ++checked;
              if (in != mask_allowed(mode, q, k, A, tl)) {
                if (bad++ < 10) printf("interval mismatch mode %d L %d A %d t %d q %d k %d [%d,%d)\n", mode, L, A, tl, q, k, lo, hi);
              }
            }
          }
          // tile predicates (inclusive bounds) == any / all over the tile
          for (int T = 3; T <= 8; T += 5)
            for (int q0 = 0; q0 < L; q0 += T)
              for (int k0 = 0; k0 < L; k0 += T) {
                const int q1 = q0 + T - 1 < L - 1 ? q0 + T - 1 : L - 1, k1 = k0 + T - 1 < L - 1 ? k0 + T - 1 : L - 1;
                bool any = false, all = true;
                for (int q = q0; q <= q1; ++q)
                  for (int k = k0; k <= k1; ++k) {
                    const bool ok = mask_allowed(mode, q, k, A, tl);
                    any = any || ok;
                    all = all && ok;
                  }
                ++checked;
                const bool any_p = tile_any_allowed(mode, q0, q1, k0, k1, A, tl), all_p = tile_all_allowed(mode, q0, q1, k0, k1, A, tl);
                // exactness contract: a tile is skipped only if NOTHING is allowed (any_p may be conservatively true), and the
                // unmasked fast path is taken only if EVERYTHING is allowed (all_p may be conservatively false)
                if ((any && !any_p) || (all_p && !all)) {
                  if (bad++ < 10)
                    printf("tile mismatch mode %d L %d A %d t %d q[%d,%d] k[%d,%d]: any %d/%d all %d/%d\n", mode, L, A, tl, q0, q1, k0, k1,
                           any, any_p, all, all_p);
                }
                if (any_p && !any) ++checked;   // conservative (allowed by the contract): counted, not an error
              }
        }
  printf("%s: %ld checks, %ld mismatches\n", bad ? "MASK TEST FAILED" : "MASK TEST PASSED", checked, bad);
  return bad ? 1 : 0;
}
