/*
 * abi_driver.c -- a plain C caller of libbandup_gpu.so that makes exactly the sequence of calls
 * the patched BandUP main loop makes through fortran/bandup_gpu_mod.f90:
 *   init -> set_cells -> set_grid -> [per SC-kpt: folding_g0 -> submit_kpt -> fetch_kpt] -> finalize
 * with two SC k-points in flight (double buffering).  No Python, no torch: this is the C ABI as a
 * Fortran/C host sees it.  Input comes from a small binary file written by the test; results go
 * to stdout as text for the test to compare with the oracle.
 *
 * file layout (little endian):
 *   int32  n_kpts, nener_expected
 *   double B_sc[9] (Fortran order), b_pc[9] (Fortran order), E_start, E_end, delta_e
 *   per k-point: int32 n_spinor, n_pw, n_bands, n_fold
 *                double folding_G[n_fold][3]; int32 g_frac[n_pw][3]; double energies[n_bands];
 *                complex64 coeffs[n_bands][n_pw][n_spinor]
 *
 * build: gcc -O2 -std=c11 -I include tests/abi_driver.c -L bandup_b200/lib -lbandup_gpu -o abi_driver
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bandup_gpu.h"

#define CHECK(call)                                                                       \
  do {                                                                                    \
    int rc_ = (call);                                                                     \
    if (rc_ != BANDUP_GPU_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, bandup_gpu_last_error(h));            \
      return 10 + rc_;                                                                    \
    }                                                                                     \
  } while (0)

typedef struct {
  int32_t n_spinor, n_pw, n_bands, n_fold;
  double *folding_G, *energies;
  int32_t *g_frac, *g0;
  void *coeffs; /* pinned */
} Kpt;

static int rd(void *dst, size_t n, FILE *f) { return fread(dst, 1, n, f) == n ? 0 : 1; }

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t n_kpts = 0, nener_expected = 0;
  double B_sc[9], b_pc[9], win[3];
  if (rd(&n_kpts, 4, f) || rd(&nener_expected, 4, f) || rd(B_sc, 72, f) || rd(b_pc, 72, f) || rd(win, 24, f))
    return 4;

  int h = -1;
  if (bandup_gpu_abi_version() != BANDUP_GPU_ABI_VERSION) return 5;
  CHECK(bandup_gpu_init(0, &h));
  int32_t M[9], nener = 0;
  CHECK(bandup_gpu_set_cells(h, B_sc, b_pc, 0.0, M));
  CHECK(bandup_gpu_set_grid(h, win[0], win[1], win[2], -1.0, &nener));
  if (nener != nener_expected) return 6;
  printf("M %d %d %d %d %d %d %d %d %d\n", M[0], M[1], M[2], M[3], M[4], M[5], M[6], M[7], M[8]);

  Kpt *k = (Kpt *)calloc((size_t)n_kpts, sizeof(Kpt));
  for (int i = 0; i < n_kpts; ++i) {
    if (rd(&k[i].n_spinor, 16, f)) return 7;
    const size_t nf = (size_t)k[i].n_fold, npw = (size_t)k[i].n_pw, nb = (size_t)k[i].n_bands;
    const size_t cbytes = nb * npw * (size_t)k[i].n_spinor * 8;
    k[i].folding_G = (double *)malloc(nf * 24);
    k[i].g_frac = (int32_t *)malloc(npw * 12);
    k[i].energies = (double *)malloc(nb * 8);
    k[i].g0 = (int32_t *)malloc(nf * 12);
    CHECK(bandup_gpu_pinned_alloc(cbytes, &k[i].coeffs));
    if (rd(k[i].folding_G, nf * 24, f) || rd(k[i].g_frac, npw * 12, f) || rd(k[i].energies, nb * 8, f) ||
        rd(k[i].coeffs, cbytes, f))
      return 8;
  }
  fclose(f);

  const int depth = bandup_gpu_pipeline_depth(h);
  for (int i = 0; i < n_kpts + depth; ++i) {
    if (i >= depth) { /* fetch the k-point submitted `depth` iterations ago */
      Kpt *q = &k[i - depth];
      double *w = (double *)malloc(sizeof(double) * (size_t)q->n_fold * (size_t)q->n_bands);
      double *dN = (double *)malloc(sizeof(double) * (size_t)q->n_fold * (size_t)nener);
      int32_t *ns = (int32_t *)malloc(sizeof(int32_t) * (size_t)q->n_fold);
      CHECK(bandup_gpu_fetch_kpt(h, i - depth + 1, w, dN, ns, NULL, NULL, 0));
      for (int fo = 0; fo < q->n_fold; ++fo) {
        printf("K %d fold %d nsel %d W", i - depth + 1, fo, ns[fo]);
        for (int b = 0; b < q->n_bands; ++b) printf(" %.17g", w[(size_t)fo * q->n_bands + b]);
        printf("\nK %d fold %d dN", i - depth + 1, fo);
        for (int e = 0; e < nener; ++e) printf(" %.17g", dN[(size_t)fo * nener + e]);
        printf("\n");
      }
      free(w); free(dN); free(ns);
    }
    if (i < n_kpts) {
      Kpt *q = &k[i];
      double resid = 0.0;
      CHECK(bandup_gpu_folding_g0(h, q->folding_G, q->n_fold, q->g0, &resid));
      CHECK(bandup_gpu_submit_kpt(h, i + 1, q->n_spinor, q->n_pw, q->n_bands, BANDUP_GPU_LAYOUT_FORTRAN_INMEM,
                                  (int64_t)q->n_pw * q->n_spinor * 8, q->coeffs, q->g_frac, q->energies,
                                  q->n_fold, q->g0));
    }
  }
  printf("launches %lld\n", (long long)bandup_gpu_launch_count(h));
  for (int i = 0; i < n_kpts; ++i) bandup_gpu_pinned_free(k[i].coeffs);
  CHECK(bandup_gpu_finalize(h));
  return 0;
}
