/*
 * bandup_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain C (+OpenMP) restatement of the arithmetic of BandUP's `unfold`
 * hot path, written from a reading of the reference's Fortran.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (bandup_b200/, libbandup_gpu.so)
 * never does.
 *
 * PARITY STATUS: pinned against the reference's own Fortran, EXECUTED BY AN
 * INTERPRETER, not compiled.  The reference ships no tests, golden vectors or
 * fixtures for this path and this image has no
 * Fortran front end, so oracle/_ref cannot be built.  Instead
 * oracle/f90interp.py interprets the reference's source files where they lie
 * (/root/reference/src: get_rec_latt, check_if_pc_and_SC_are_commensurate,
 * populate_wf_with_G_vec_coords, the renormalisation block of
 * read_wavefunction, select_coeffs_to_calc_spectral_weights, perform_unfolding
 * with everything below it, get_delta_Ns_for_EBS with a smearing,
 * calc_spectral_function, get_delta_Ns_for_output, read_wavecar on synthetic
 * WAVECAR files) on seeded inputs; tests/golden/make_fortran_golden.py is the
 * generating script and tests/golden/fortran_golden.npz the committed result
 * (13 cases, two of them the reference as a kind_cplx_coeffs = dp build, two
 * the fold search get_geom_unfolding_relations; 2.4 M Fortran statements).  tests/test_fortran_golden.py holds every function of this file
 * against those vectors: plane-wave list, renormalised coefficients, selected
 * indices, rho and the symmetry averages bit for bit; weights, delta_N and the
 * spectral function to 1e-15.  What that does NOT pin: code generation of a
 * particular Fortran compiler (the interpreter evaluates every operation in the
 * kind Fortran prescribes, without reassociation or fused multiply-add -- what
 * gfortran -O2 does on baseline x86-64, src/Makefile:23-30), the order OpenMP
 * threads append selected indices in (ascending here), the rho-file writer
 * (write_unf_dens_op: held by the reference's own Python reader instead).  The interpreter itself
 * is held against a COMPILED BandUP where the reference tree allows it: real_seq
 * + write_band_struc, interpreted, print all 36 603 data lines of
 * utils/sample_input_files_task_plot/unfolded_EBS_symmetry-averaged.dat text for
 * text (tests/golden/make_sample_plot_writer_check.py).
 * Also asserted in tests/: agreement with an independent NumPy restatement
 * (oracle/oracle_np.py) and analytic invariants (sum over cosets of P = 1,
 * perfect-supercell weights in {0,1}, float lattice test == integer coset test).
 *
 * Conventions: 3x3 matrices are row-major with ROW i = vector i, i.e.
 * m[3*i+j] == fortran_matrix(i+1, j+1).  All reference citations are relative
 * to /root/reference/src.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BO_EPS_DP 2.220446049250313e-16 /* epsilon(1.0_dp) */

/* constants_and_types_mod.f90:63-66 */
static const double BO_DEFAULT_TOL_VEC_EQ = 1e-5;
static const double BO_MAX_TOL_VEC_EQ = 1e-3;

int bo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void bo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- math_mod.f90:196-231  inverse_of_3x3_matrix ------------------------ */
int bo_inverse_3x3(const double *m, double *inv) {
  double c[9];
  c[0] = (m[4] * m[8] - m[5] * m[7]);
  c[1] = -(m[3] * m[8] - m[5] * m[6]);
  c[2] = (m[3] * m[7] - m[4] * m[6]);
  double det = m[0] * c[0] + m[1] * c[1] + m[2] * c[2];
  if (fabs(det) <= 1.0e-10) {
    for (int i = 0; i < 9; ++i) inv[i] = 0.0;
    return 0;
  }
  c[3] = -(m[1] * m[8] - m[2] * m[7]);
  c[4] = (m[0] * m[8] - m[2] * m[6]);
  c[5] = -(m[0] * m[7] - m[1] * m[6]);
  c[6] = (m[1] * m[5] - m[2] * m[4]);
  c[7] = -(m[0] * m[5] - m[2] * m[3]);
  c[8] = (m[0] * m[4] - m[1] * m[3]);
  /* inverse = transpose(cofactor)/det */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[3 * i + j] = c[3 * j + i] / det;
  return 1;
}

static void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---- crystals_mod.f90:186-202  get_rec_latt ------------------------------ */
void bo_rec_latt(const double *latt, double *rec) {
  const double twopi = 2.0 * (4.0 * atan(1.0));
  double c[3];
  cross3(latt + 3, latt + 6, c);
  double vpc = fabs(latt[0] * c[0] + latt[1] * c[1] + latt[2] * c[2]);
  for (int j = 0; j < 3; ++j) rec[j] = twopi * c[j] / vpc;
  cross3(latt + 6, latt + 0, c);
  for (int j = 0; j < 3; ++j) rec[3 + j] = twopi * c[j] / vpc;
  cross3(latt + 0, latt + 3, c);
  for (int j = 0; j < 3; ++j) rec[6 + j] = twopi * c[j] / vpc;
}

/* ---- crystals_mod.f90:41-73  check_if_pc_and_SC_are_commensurate ---------
 * M = transpose(matmul(b_pc, inverse(B_SC))); int_M = nint(M).
 * Returns 1 when every |M - int_M| <= tol (default 1e-5). */
int bo_commensurate(const double *b_pc, const double *B_sc, double *M,
                    int32_t *int_M, double tol) {
  double Binv[9], P[9];
  if (tol <= 0.0) tol = 1e-5; /* default_tol_for_int_commens_test */
  bo_inverse_3x3(B_sc, Binv);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += b_pc[3 * i + k] * Binv[3 * k + j];
      P[3 * i + j] = s;
    }
  int ok = 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      M[3 * i + j] = P[3 * j + i];
      int_M[3 * i + j] = (int32_t)lround(M[3 * i + j]); /* nint */
      if (fabs(M[3 * i + j] - (double)int_M[3 * i + j]) > tol) ok = 0;
    }
  return ok;
}

/* ---- math_mod.f90:234-248  coords_cart_vec_in_new_basis ------------------
 * aux = inverse(transpose(basis)); coords(i) = dot(aux(i,:), v).
 * The inverse is recomputed per call in the reference; here the caller may
 * hoist it (same arithmetic, same result). */
static void inv_transposed_basis(const double *basis, double *aux) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = basis[3 * j + i];
  bo_inverse_3x3(t, aux);
}

void bo_coords_in_basis(const double *v, const double *basis, double *out) {
  double aux[9];
  inv_transposed_basis(basis, aux);
  for (int i = 0; i < 3; ++i)
    out[i] = aux[3 * i] * v[0] + aux[3 * i + 1] * v[1] + aux[3 * i + 2] * v[2];
}

/* ---- math_mod.f90:250-282  same_vector ----------------------------------- */
int bo_same_vector(const double *v1, const double *v2, double tol) {
  tol = fabs(tol);
  double d[3];
  for (int i = 0; i < 3; ++i) {
    d[i] = v2[i] - v1[i];
    if (fabs(d[i]) > tol) return 0;
  }
  double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  return dist > tol ? 0 : 1;
}

/* ---- crystals_mod.f90:204-255  vec_in_latt -------------------------------
 * frac = coords of vec in latt; r = mod(frac, 1) (sign of dividend == fmod);
 * reduced = r . latt; true iff reduced equals one of the 27 combinations
 * {0,1,-1}^3 . latt (loop order ig3, ig2, ig1 with 0,1,-1 each). */
static int vec_in_latt_pre(const double *vec, const double *latt,
                           const double *aux, double tol) {
  double f[3], red[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < 3; ++i)
    f[i] = aux[3 * i] * vec[0] + aux[3 * i + 1] * vec[1] + aux[3 * i + 2] * vec[2];
  for (int i = 0; i < 3; ++i) {
    double r = fmod(f[i], 1.0);
    for (int j = 0; j < 3; ++j) red[j] += r * latt[3 * i + j];
  }
  static const int seq[3] = {0, 1, -1};
  for (int a3 = 0; a3 < 3; ++a3)
    for (int a2 = 0; a2 < 3; ++a2)
      for (int a1 = 0; a1 < 3; ++a1) {
        double g[3];
        for (int j = 0; j < 3; ++j)
          g[j] = seq[a1] * latt[j] + seq[a2] * latt[3 + j] + seq[a3] * latt[6 + j];
        if (bo_same_vector(red, g, tol)) return 1;
      }
  return 0;
}

int bo_vec_in_latt(const double *vec, const double *latt, double tol) {
  double aux[9];
  inv_transposed_basis(latt, aux);
  return vec_in_latt_pre(vec, latt, aux, tol);
}

/* ---- io_routines_mod.f90:1317-1330  band renormalisation -----------------
 * coeffs layout == Fortran pw_coeffs(n_spinor, n_pw, n_bands) in memory:
 * [band][pw][spinor] complex64.
 * norm_mode 0 ("faithful_fp32_sequential"): per spinor a complex(sp)
 *   dot_product accumulated sequentially in single precision, spinor results
 *   summed in single precision, real part widened to double.
 * norm_mode 1 ("exact_fp64"): the same sum accumulated in double.
 * In both modes: if s > epsilon: C <- complex64( (1/sqrt(|s|)) * C ), the
 * product evaluated in double as the reference's mixed-kind expression does.
 * norms_out (optional) receives s per band. */
void bo_renormalize(float complex *coeffs, int n_spinor, int64_t n_pw,
                    int n_bands, int norm_mode, double *norms_out) {
#pragma omp parallel for schedule(guided)
  for (int ib = 0; ib < n_bands; ++ib) {
    float complex *c = coeffs + (int64_t)ib * n_pw * n_spinor;
    double s;
    if (norm_mode == 0) {
      float tot_re = 0.0f;
      for (int is = 0; is < n_spinor; ++is) {
        volatile float acc = 0.0f; /* volatile: forbid reassociation/vectorised partial sums */
        for (int64_t ig = 0; ig < n_pw; ++ig) {
          float re = crealf(c[ig * n_spinor + is]);
          float im = cimagf(c[ig * n_spinor + is]);
          float p1 = re * re;
          float p2 = im * im;
          float t = p1 + p2;
          acc = acc + t;
        }
        tot_re = tot_re + acc;
      }
      s = (double)tot_re;
    } else {
      double acc = 0.0;
      for (int is = 0; is < n_spinor; ++is)
        for (int64_t ig = 0; ig < n_pw; ++ig) {
          double re = (double)crealf(c[ig * n_spinor + is]);
          double im = (double)cimagf(c[ig * n_spinor + is]);
          acc += re * re + im * im;
        }
      s = acc;
    }
    if (norms_out) norms_out[ib] = s;
    if (s > BO_EPS_DP) {
      double f = 1.0 / sqrt(fabs(s));
      for (int64_t e = 0; e < n_pw * n_spinor; ++e) {
        double re = f * (double)crealf(c[e]);
        double im = f * (double)cimagf(c[e]);
        c[e] = (float)re + (float)im * I;
      }
    }
  }
}

/* ---- band_unfolding_routines_mod.f90:550-612
 *      select_coeffs_to_calc_spectral_weights ------------------------------
 * folding_G = Scoords(k) - K (Cartesian).  tol starts at 0.9e-5 and is raised
 * by 1e-6 per sweep until at least one plane wave passes vec_in_latt or tol
 * exceeds 1e-3.  sel_out receives 1-based indices in ASCENDING order (the
 * reference appends in thread-arrival order; with one thread that is
 * ascending).  Returns n_sel (0 == "pckpt cannot be parsed"). */
int64_t bo_select_coeffs(const double *G_cart, int64_t n_pw,
                         const double *folding_G, const double *b_pc,
                         int32_t *sel_out, double *tol_used) {
  double aux[9];
  inv_transposed_basis(b_pc, aux);
  unsigned char *flag = (unsigned char *)malloc((size_t)(n_pw > 0 ? n_pw : 1));
  double tol = 0.9 * BO_DEFAULT_TOL_VEC_EQ;
  int64_t n_sel = 0;
  while (n_sel == 0 && tol <= BO_MAX_TOL_VEC_EQ) {
    tol = tol + 0.1 * BO_DEFAULT_TOL_VEC_EQ;
#pragma omp parallel for schedule(guided)
    for (int64_t ig = 0; ig < n_pw; ++ig) {
      double trial[3];
      for (int j = 0; j < 3; ++j) trial[j] = G_cart[3 * ig + j] - folding_G[j];
      flag[ig] = (unsigned char)vec_in_latt_pre(trial, b_pc, aux, tol);
    }
    n_sel = 0;
    for (int64_t ig = 0; ig < n_pw; ++ig)
      if (flag[ig]) sel_out[n_sel++] = (int32_t)(ig + 1);
  }
  if (tol_used) *tol_used = tol;
  free(flag);
  return n_sel;
}

/* ---- band_unfolding_routines_mod.f90:615-651 (+ :1335-1345 spinor sum)
 *      spectral_weight_for_coeff -------------------------------------------
 * P(band) = sum_spinor sum_{i in sel} dble( abs(coeff) )**2 with abs() in
 * single precision, square and sum in double.  Outer loop over spinor
 * components as perform_unfolding does. */
void bo_spectral_weights(const float complex *coeffs, int n_spinor,
                         int64_t n_pw, int n_bands, const int32_t *sel,
                         int64_t n_sel, double *weights) {
  for (int ib = 0; ib < n_bands; ++ib) weights[ib] = 0.0;
  for (int is = 0; is < n_spinor; ++is) {
#pragma omp parallel for schedule(guided)
    for (int ib = 0; ib < n_bands; ++ib) {
      const float complex *c = coeffs + (int64_t)ib * n_pw * n_spinor;
      double w = 0.0;
      for (int64_t i = 0; i < n_sel; ++i) {
        int64_t ig = (int64_t)sel[i] - 1;
        float a = cabsf(c[ig * n_spinor + is]);
        double ad = (double)a;
        w = w + ad * ad;
      }
      weights[ib] = weights[ib] + w;
    }
  }
}

/* ---- math_mod.f90:181-193  integral_delta_x_minus_x0 --------------------- */
double bo_integral_delta(double x0, double lower_lim, double upper_lim,
                         double std_dev) {
  double x1 = (lower_lim - x0) / (sqrt(2.0) * std_dev);
  double x2 = (upper_lim - x0) / (sqrt(2.0) * std_dev);
  return 0.5 * (erf(x2) - erf(x1));
}

/* ---- band_unfolding_routines_mod.f90:701-716  lambda --------------------- */
double bo_lambda(double pc_ener, double sc_ener, double delta_e, double sigma) {
  return bo_integral_delta(sc_ener, pc_ener - 0.5 * delta_e,
                           pc_ener + 0.5 * delta_e, sigma);
}

/* ---- lists_and_seqs_mod.f90:188-200  real_seq ----------------------------
 * Returns the number of terms; fills out[] when it is non-NULL. */
int64_t bo_real_seq(double first, double last, double incr, double *out) {
  int64_t n = (int64_t)lround((last - first) / incr + 1.0);
  if (out)
    for (int64_t i = 1; i <= n; ++i) out[i - 1] = first + (double)(i - 1) * incr;
  return n;
}

/* ---- band_unfolding_routines_mod.f90:719-786
 *      get_delta_Ns_for_EBS -> calc_delta_N_pckpt ---------------------------
 * smearing <= 0 means "not passed": sigma = epsilon(1.0_dp) (box).  Otherwise
 * sigma = max(epsilon, |smearing|).  delta_e = grid[1]-grid[0]. */
void bo_delta_N(const double *energies, int64_t nener, const double *sc_ener,
                const double *weights, int n_bands, double smearing,
                double *dN) {
  double sigma = BO_EPS_DP;
  if (smearing > 0.0) sigma = fmax(sigma, fabs(smearing));
  double delta_e = energies[1] - energies[0];
#pragma omp parallel for schedule(guided)
  for (int64_t ie = 0; ie < nener; ++ie) {
    double E = energies[ie];
    double acc = 0.0;
    for (int ib = 0; ib < n_bands; ++ib)
      acc = acc + weights[ib] * bo_lambda(E, sc_ener[ib], delta_e, sigma);
    dN[ie] = acc;
  }
}

/* ---- math_mod.f90:159-178  delta (a Gaussian; std_dev given) --------------- */
double bo_delta(double x, double std_dev) {
  const double twopi = 2.0 * (4.0 * atan(1.0));
  double stddev = fabs(std_dev);
  double r = x / stddev;
  return exp(-(r * r) / 2.0) / (stddev * sqrt(twopi));
}

/* ---- band_unfolding_routines_mod.f90:654-698  calc_spectral_function -------
 * std_dev <= 0 means "not passed": sigma = 0.025. */
void bo_spectral_function(const double *energies, int64_t nener, const double *sc_ener,
                          const double *weights, int n_bands, double std_dev, double *sf) {
  double sigma = std_dev > 0.0 ? std_dev : 0.025;
#pragma omp parallel for schedule(guided)
  for (int64_t ie = 0; ie < nener; ++ie) {
    double E = energies[ie], acc = 0.0;
    for (int ib = 0; ib < n_bands; ++ib) acc = acc + weights[ib] * bo_delta(sc_ener[ib] - E, sigma);
    sf[ie] = acc;
  }
}

/* ---- vasp_related_modules/read_wavecar_mod.f90:151-210
 *      get_max_nb1_nb2_and_nb3 ---------------------------------------------- */
static double norm3(const double *a) {
  return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
}
static double angle3(const double *a, const double *b) {
  /* math_mod.f90:150-157 */
  return acos((a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (norm3(a) * norm3(b)));
}

void bo_max_nb(const double *B, double ecut, double c, int *nbmax) {
  const double *b1 = B, *b2 = B + 3, *b3 = B + 6;
  double x[3];
  double phi12 = angle3(b1, b2), phi13 = angle3(b1, b3), phi23 = angle3(b2, b3);
  double m1 = norm3(b1), m2 = norm3(b2), m3 = norm3(b3);
  double r = sqrt(ecut * c);
  cross3(b1, b2, x);
  double s = cos(angle3(b3, x));
  int a1 = (int)(r / (m1 * fabs(sin(phi12))) + 1.0);
  int a2 = (int)(r / (m2 * fabs(sin(phi12))) + 1.0);
  int a3 = (int)(r / (m3 * fabs(s)) + 1.0);
  cross3(b1, b3, x);
  s = cos(angle3(b2, x));
  int c1 = (int)(r / (m1 * fabs(sin(phi13))) + 1.0);
  int c2 = (int)(r / (m2 * fabs(s)) + 1.0);
  int c3 = (int)(r / (m3 * fabs(sin(phi13))) + 1.0);
  cross3(b2, b3, x);
  s = cos(angle3(b1, x));
  int d1 = (int)(r / (m1 * fabs(s)) + 1.0);
  int d2 = (int)(r / (m2 * fabs(sin(phi23))) + 1.0);
  int d3 = (int)(r / (m3 * fabs(sin(phi23))) + 1.0);
  nbmax[0] = a1 > c1 ? (a1 > d1 ? a1 : d1) : (c1 > d1 ? c1 : d1);
  nbmax[1] = a2 > c2 ? (a2 > d2 ? a2 : d2) : (c2 > d2 ? c2 : d2);
  nbmax[2] = a3 > c3 ? (a3 > d3 ? a3 : d3) : (c3 > d3 ? c3 : d3);
}

/* ---- read_wavecar_mod.f90:213-304  n_Gvecs_2b_generated +
 *      populate_wf_with_G_vec_coords ----------------------------------------
 * Enumerates ig3 (slowest) -> ig2 -> ig1 (fastest), each 0..nbmax then
 * -nbmax..-1, keeping G when |(K+G).B|^2 / c < ecut.  Pass g_frac/g_cart NULL
 * to count only.  c defaults to the float32-rounded literal 0.262465831
 * (read_wavecar_mod.f90:107) when c <= 0.  capacity bounds the fill. */
int64_t bo_gen_gvecs(const double *kfrac, double ecut, const double *B,
                     double c, int32_t *g_frac, double *g_cart,
                     int64_t capacity) {
  if (c <= 0.0) c = (double)0.262465831f;
  int nb[3];
  bo_max_nb(B, ecut, c, nb);
  const double *b1 = B, *b2 = B + 3, *b3 = B + 6;
  int64_t n = 0;
  for (int i3 = 0; i3 <= 2 * nb[2]; ++i3) {
    int p3 = i3 > nb[2] ? i3 - 2 * nb[2] - 1 : i3;
    for (int i2 = 0; i2 <= 2 * nb[1]; ++i2) {
      int p2 = i2 > nb[1] ? i2 - 2 * nb[1] - 1 : i2;
      for (int i1 = 0; i1 <= 2 * nb[0]; ++i1) {
        int p1 = i1 > nb[0] ? i1 - 2 * nb[0] - 1 : i1;
        double sk[3];
        for (int j = 0; j < 3; ++j)
          sk[j] = (kfrac[0] + p1) * b1[j] + (kfrac[1] + p2) * b2[j] +
                  (kfrac[2] + p3) * b3[j];
        double gtot = sqrt(sk[0] * sk[0] + sk[1] * sk[1] + sk[2] * sk[2]);
        double etot = gtot * gtot / c;
        if (etot < ecut) {
          if (g_frac && n < capacity) {
            g_frac[3 * n] = p1;
            g_frac[3 * n + 1] = p2;
            g_frac[3 * n + 2] = p3;
          }
          if (g_cart && n < capacity)
            for (int j = 0; j < 3; ++j)
              g_cart[3 * n + j] = p1 * b1[j] + p2 * b2[j] + p3 * b3[j];
          ++n;
        }
      }
    }
  }
  return n;
}

/* ---- band_unfolding_routines_mod.f90:1046-1161  calc_rho (unfold branch) --
 * rho is n_bands x n_bands complex64, row-major here: rho[m1*nb + m2] ==
 * fortran rho(m1+1, m2+1).  m1 runs over 1..nb-1 only (reference quirk);
 * dot_product(a,b) = sum conj(a)*b accumulated in complex single precision.
 * Returns 1 when delta_N < epsilon (the reference's fatal branch). */
int bo_calc_rho(const float complex *coeffs, int n_spinor, int64_t n_pw,
                int n_bands, const int32_t *sel, int64_t n_sel,
                const double *sc_ener, double delta_N, double pc_ener,
                double delta_e, double sigma, float complex *rho) {
  if (delta_N < BO_EPS_DP) return 1;
  if (sigma <= 0.0) sigma = BO_EPS_DP;
  const double min_pref = 1e-5;
  memset(rho, 0, sizeof(float complex) * (size_t)n_bands * (size_t)n_bands);
#pragma omp parallel for schedule(guided)
  for (int m1 = 0; m1 < n_bands - 1; ++m1) {
    double l1 = bo_lambda(pc_ener, sc_ener[m1], delta_e, sigma);
    if (l1 < min_pref) continue;
    const float complex *c1 = coeffs + (int64_t)m1 * n_pw * n_spinor;
    for (int m2 = m1; m2 < n_bands; ++m2) {
      double l2 = bo_lambda(pc_ener, sc_ener[m2], delta_e, sigma);
      double pref = (l1 * l2) / delta_N;
      if (pref < min_pref) continue;
      const float complex *c2 = coeffs + (int64_t)m2 * n_pw * n_spinor;
      float complex r = 0.0f;
      for (int is = 0; is < n_spinor; ++is) {
        float are = 0.0f, aim = 0.0f;
        for (int64_t i = 0; i < n_sel; ++i) {
          int64_t ig = (int64_t)sel[i] - 1;
          float ar = crealf(c1[ig * n_spinor + is]), ai = -cimagf(c1[ig * n_spinor + is]);
          float br = crealf(c2[ig * n_spinor + is]), bi = cimagf(c2[ig * n_spinor + is]);
          float pr = ar * br - ai * bi;
          float pi = ar * bi + ai * br;
          are = are + pr;
          aim = aim + pi;
        }
        r = (crealf(r) + are) + (cimagf(r) + aim) * I;
      }
      /* rho(m1,m2) = prefactor * rho(m1,m2): real(dp) * complex(sp) -> dp, rounded to sp */
      float rr = (float)(pref * (double)crealf(r));
      float ri = (float)(pref * (double)cimagf(r));
      rho[(int64_t)m1 * n_bands + m2] = rr + ri * I;
      rho[(int64_t)m2 * n_bands + m1] = rr - ri * I;
    }
  }
  return 0;
}

/* ---- band_unfolding_routines_mod.f90:1136-1154  calc_rho, -dont_unfold branch --
 * rho(m1,m2) = 1 where lambda(m1) >= 1e-5 and lambda(m1) lambda(m2) / delta_N >= 1e-5 (m1 and m2
 * over ALL bands), then rho = rho / real(n_picked_bands, dp): complex(sp) / real(dp) formed in dp
 * and stored to sp. */
int bo_calc_rho_dont_unfold(int n_bands, const double *sc_ener, double delta_N, double pc_ener,
                            double delta_e, double sigma, float complex *rho) {
  if (delta_N < BO_EPS_DP) return 1;
  if (sigma <= 0.0) sigma = BO_EPS_DP;
  const double min_pref = 1e-5;
  memset(rho, 0, sizeof(float complex) * (size_t)n_bands * (size_t)n_bands);
  int n_picked = 0;
  for (int m1 = 0; m1 < n_bands; ++m1) {
    double l1 = bo_lambda(pc_ener, sc_ener[m1], delta_e, sigma);
    if (l1 < min_pref) continue;
    ++n_picked;
    for (int m2 = 0; m2 < n_bands; ++m2) {
      double l2 = bo_lambda(pc_ener, sc_ener[m2], delta_e, sigma);
      if ((l1 * l2) / delta_N < min_pref) continue;
      rho[(int64_t)m1 * n_bands + m2] = 1.0f;
    }
  }
  if (n_picked > 0)
    for (int64_t i = 0; i < (int64_t)n_bands * n_bands; ++i) {
      double re = (double)crealf(rho[i]) / (double)n_picked, im = (double)cimagf(rho[i]) / (double)n_picked;
      rho[i] = (float)re + (float)im * I;
    }
  return 0;
}

/* ---- band_unfolding_routines_mod.f90:789-855, :975-1043  symmetry average --
 * avrgd = 0; do i_needed_dirs: avrgd = avrgd + weight * x(i_needed_dirs) -- dp, direction order.
 * in [n_dirs][n_values], out [n_values]. */
void bo_symm_average(const double *weights, int n_dirs, int64_t n_values,
                     const double *in, double *out) {
  for (int64_t r = 0; r < n_values; ++r) {
    double acc = 0.0;
    for (int d = 0; d < n_dirs; ++d) {
      volatile double t = weights[d] * in[(int64_t)d * n_values + r]; /* no fma contraction */
      acc = acc + t;
    }
    out[r] = acc;
  }
}

/* ---- band_unfolding_routines_mod.f90:857-973  symmetry-averaged rho --------
 * Entries of all directions concatenated, direction d = [dir_off[d], dir_off[d+1]), each
 * direction in the reference's append order (pckpt, energy ascending, m1 outer, m2 inner;
 * pckpt 0-based, iener/m1/m2 1-based).  Operators at energies whose AVERAGED delta_N is below
 * min_dN are skipped (:906; the literal 1E-3 is default real).  The first direction holding
 * (m1,m2) creates the element as cmplx(weight*rho, kind=sp) (:934-941), later ones add
 * weight*rho in dp and store to sp (:952-957).  Output: (pckpt, iener) ascending and, inside an
 * operator, the reference's append order (direction, then m1/m2).  Returns the number of
 * averaged entries (<= dir_off[n_dirs], which is the capacity the caller provides). */
typedef struct {
  int64_t group, pair, seq;
  float re, im;
} bo_avg_entry;

static int bo_avg_cmp(const void *a, const void *b) {
  const bo_avg_entry *x = (const bo_avg_entry *)a, *y = (const bo_avg_entry *)b;
  if (x->group != y->group) return x->group < y->group ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);
}

int64_t bo_symm_average_rho(const double *weights, int n_dirs, int nener, int n_bands,
                            const double *avg_dN, double min_dN, const int64_t *dir_off,
                            const int32_t *pckpt, const int32_t *iener, const int32_t *m1,
                            const int32_t *m2, const float *rho_re, const float *rho_im,
                            int32_t *pckpt_out, int32_t *iener_out, int32_t *m1_out,
                            int32_t *m2_out, float *re_out, float *im_out) {
  const int64_t n_in = dir_off[n_dirs];
  if (n_in == 0) return 0;
  bo_avg_entry *out = (bo_avg_entry *)malloc(sizeof(bo_avg_entry) * (size_t)n_in);
  int64_t cap = 16;
  while (cap < 2 * n_in) cap <<= 1;
  int64_t *table = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap); /* open addressing -> out index */
  for (int64_t i = 0; i < cap; ++i) table[i] = -1;
  int64_t n_out = 0;
  for (int d = 0; d < n_dirs; ++d) {
    for (int64_t i = dir_off[d]; i < dir_off[d + 1]; ++i) {
      const int64_t group = (int64_t)pckpt[i] * nener + (iener[i] - 1);
      if (avg_dN[group] < min_dN) continue;
      const int64_t pair = (int64_t)(m1[i] - 1) * n_bands + (m2[i] - 1);
      uint64_t h = ((uint64_t)group * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)pair * 0xC2B2AE3D27D4EB4Full);
      int64_t slot = (int64_t)(h & (uint64_t)(cap - 1));
      while (table[slot] >= 0 && !(out[table[slot]].group == group && out[table[slot]].pair == pair))
        slot = (slot + 1) & (cap - 1);
      const double tr = weights[d] * (double)rho_re[i], ti = weights[d] * (double)rho_im[i];
      if (table[slot] < 0) {
        table[slot] = n_out;
        out[n_out].group = group;
        out[n_out].pair = pair;
        out[n_out].seq = n_out;
        out[n_out].re = (float)tr;
        out[n_out].im = (float)ti;
        ++n_out;
      } else {
        bo_avg_entry *e = &out[table[slot]];
        volatile double sr = (double)e->re + tr, si = (double)e->im + ti;
        e->re = (float)sr;
        e->im = (float)si;
      }
    }
  }
  qsort(out, (size_t)n_out, sizeof(bo_avg_entry), bo_avg_cmp);
  for (int64_t i = 0; i < n_out; ++i) {
    pckpt_out[i] = (int32_t)(out[i].group / nener);
    iener_out[i] = (int32_t)(out[i].group % nener) + 1;
    m1_out[i] = (int32_t)(out[i].pair / n_bands) + 1;
    m2_out[i] = (int32_t)(out[i].pair % n_bands) + 1;
    re_out[i] = out[i].re;
    im_out[i] = out[i].im;
  }
  free(table);
  free(out);
  return n_out;
}

/* ---- Integer coset test (the arithmetic the GPU kernel uses; NOT in the
 * reference).  Used by tests to assert float selection == integer selection.
 * adjT = adjugate(transpose(M)) (integer), D = |det M|.
 * key(g) = (g . adjT) mod D, component-wise in [0, D).
 * G - g0 lies on the pc reciprocal lattice  <=>  key(G) == key(g0). */
void bo_coset_setup(const int32_t *M, int64_t *adjT, int64_t *D) {
  int64_t T[9], c[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = M[3 * j + i];
  c[0] = T[4] * T[8] - T[5] * T[7];
  c[1] = -(T[3] * T[8] - T[5] * T[6]);
  c[2] = T[3] * T[7] - T[4] * T[6];
  c[3] = -(T[1] * T[8] - T[2] * T[7]);
  c[4] = T[0] * T[8] - T[2] * T[6];
  c[5] = -(T[0] * T[7] - T[1] * T[6]);
  c[6] = T[1] * T[5] - T[2] * T[4];
  c[7] = -(T[0] * T[5] - T[2] * T[3]);
  c[8] = T[0] * T[4] - T[1] * T[3];
  int64_t det = T[0] * c[0] + T[1] * c[1] + T[2] * c[2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) adjT[3 * i + j] = c[3 * j + i];
  *D = det < 0 ? -det : det;
}

void bo_coset_key(const int32_t *g, const int64_t *adjT, int64_t D, int64_t *key) {
  for (int j = 0; j < 3; ++j) {
    int64_t v = (int64_t)g[0] * adjT[j] + (int64_t)g[1] * adjT[3 + j] +
                (int64_t)g[2] * adjT[6 + j];
    v %= D;
    if (v < 0) v += D;
    key[j] = v;
  }
}

int64_t bo_select_coeffs_int(const int32_t *g_frac, int64_t n_pw,
                             const int32_t *g0, const int32_t *M,
                             int32_t *sel_out) {
  int64_t adjT[9], D, k0[3], k[3], n = 0;
  bo_coset_setup(M, adjT, &D);
  bo_coset_key(g0, adjT, D, k0);
  for (int64_t ig = 0; ig < n_pw; ++ig) {
    bo_coset_key(g_frac + 3 * ig, adjT, D, k);
    if (k[0] == k0[0] && k[1] == k0[1] && k[2] == k0[2])
      sel_out[n++] = (int32_t)(ig + 1);
  }
  return n;
}

/* ---- The reference's per-SC-k-point sequence, for timing and end-to-end
 * parity: renormalise (destructive, like the reference), then per folding
 * pc-kpt: select -> weights -> delta_N.  Mirrors main_BandUP.f90:145-172 with
 * perform_unfolding's a6/a8 part.  times[3] accumulates seconds in the
 * reference's own report categories (read_wf-renormalise / spectral weights
 * (selection + |C|^2) / N(k,E)).  Returns the number of folds with an empty
 * selection. */
static double now_s(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

int bo_unfold_kpt(float complex *coeffs, int n_spinor, int64_t n_pw,
                  int n_bands, const double *G_cart, const double *sc_ener,
                  const double *b_pc, const double *folding_G, int n_fold,
                  const double *energies, int64_t nener, double smearing,
                  int norm_mode, double *weights /*[n_fold][n_bands]*/,
                  double *dN /*[n_fold][nener]*/, int32_t *n_selected,
                  double *times) {
  int empty = 0;
  double t0 = now_s();
  bo_renormalize(coeffs, n_spinor, n_pw, n_bands, norm_mode, NULL);
  double t1 = now_s();
  if (times) times[0] += t1 - t0;
  int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_pw > 0 ? n_pw : 1));
  for (int f = 0; f < n_fold; ++f) {
    double ta = now_s();
    int64_t ns = bo_select_coeffs(G_cart, n_pw, folding_G + 3 * f, b_pc, sel, NULL);
    n_selected[f] = (int32_t)ns;
    if (ns == 0) {
      ++empty;
      for (int ib = 0; ib < n_bands; ++ib) weights[(int64_t)f * n_bands + ib] = 0.0;
      for (int64_t ie = 0; ie < nener; ++ie) dN[(int64_t)f * nener + ie] = 0.0;
      continue;
    }
    bo_spectral_weights(coeffs, n_spinor, n_pw, n_bands, sel, ns,
                        weights + (int64_t)f * n_bands);
    double tb = now_s();
    bo_delta_N(energies, nener, sc_ener, weights + (int64_t)f * n_bands, n_bands,
               smearing, dN + (int64_t)f * nener);
    double tc = now_s();
    if (times) {
      times[1] += tb - ta;
      times[2] += tc - tb;
    }
  }
  free(sel);
  return empty;
}

/* ---- a `kind_cplx_coeffs = dp` build of the reference (constants_and_types_mod.f90:60; the
 * build a double-precision WAVECAR, nprec = 45210, needs: read_wavecar_mod.f90:363-377) --------
 * The same three loops with complex(dp) coefficients: the dot_product of the renormalisation
 * (io_routines_mod.f90:1321-1324) accumulates sequentially in double, the rescale
 * (:1325-1327) stays in double, abs() in spectral_weight_for_coeff
 * (band_unfolding_routines_mod.f90:638-644) is the double-precision cabs. */
void bo_renormalize_dp(double complex *coeffs, int n_spinor, int64_t n_pw,
                       int n_bands, double *norms_out) {
#pragma omp parallel for schedule(guided)
  for (int ib = 0; ib < n_bands; ++ib) {
    double complex *c = coeffs + (int64_t)ib * n_pw * n_spinor;
    double s = 0.0;
    for (int is = 0; is < n_spinor; ++is) {
      double acc = 0.0;
      for (int64_t ig = 0; ig < n_pw; ++ig) {
        double re = creal(c[ig * n_spinor + is]);
        double im = cimag(c[ig * n_spinor + is]);
        acc += re * re + im * im;
      }
      s += acc;
    }
    if (norms_out) norms_out[ib] = s;
    if (s > BO_EPS_DP) {
      double f = 1.0 / sqrt(fabs(s));
      for (int64_t e = 0; e < n_pw * n_spinor; ++e) c[e] = f * creal(c[e]) + f * cimag(c[e]) * I;
    }
  }
}

void bo_spectral_weights_dp(const double complex *coeffs, int n_spinor,
                            int64_t n_pw, int n_bands, const int32_t *sel,
                            int64_t n_sel, double *weights) {
  for (int ib = 0; ib < n_bands; ++ib) weights[ib] = 0.0;
  for (int is = 0; is < n_spinor; ++is) {
#pragma omp parallel for schedule(guided)
    for (int ib = 0; ib < n_bands; ++ib) {
      const double complex *c = coeffs + (int64_t)ib * n_pw * n_spinor;
      double w = 0.0;
      for (int64_t i = 0; i < n_sel; ++i) {
        int64_t ig = (int64_t)sel[i] - 1;
        double a = cabs(c[ig * n_spinor + is]);
        w = w + a * a;
      }
      weights[ib] = weights[ib] + w;
    }
  }
}

int bo_unfold_kpt_dp(double complex *coeffs, int n_spinor, int64_t n_pw,
                     int n_bands, const double *G_cart, const double *sc_ener,
                     const double *b_pc, const double *folding_G, int n_fold,
                     const double *energies, int64_t nener, double smearing,
                     double *weights /*[n_fold][n_bands]*/,
                     double *dN /*[n_fold][nener]*/, int32_t *n_selected,
                     double *norms_out) {
  int empty = 0;
  bo_renormalize_dp(coeffs, n_spinor, n_pw, n_bands, norms_out);
  int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_pw > 0 ? n_pw : 1));
  for (int f = 0; f < n_fold; ++f) {
    int64_t ns = bo_select_coeffs(G_cart, n_pw, folding_G + 3 * f, b_pc, sel, NULL);
    n_selected[f] = (int32_t)ns;
    if (ns == 0) {
      ++empty;
      for (int ib = 0; ib < n_bands; ++ib) weights[(int64_t)f * n_bands + ib] = 0.0;
      for (int64_t ie = 0; ie < nener; ++ie) dN[(int64_t)f * nener + ie] = 0.0;
      continue;
    }
    bo_spectral_weights_dp(coeffs, n_spinor, n_pw, n_bands, sel, ns, weights + (int64_t)f * n_bands);
    bo_delta_N(energies, nener, sc_ener, weights + (int64_t)f * n_bands, n_bands, smearing,
               dN + (int64_t)f * nener);
  }
  free(sel);
  return empty;
}

/* ---- Synthetic coefficient generator shared (bit-for-bit) with the CUDA
 * generator in bandup_b200/csrc/synth.cuh: a counter-based hash so any
 * (seed, ikpt, band, element) value can be produced without a file.
 * Values are uniform in (-1,1) times a per-element envelope 1/(1+e mod 7). */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

void bo_synth_fill_rows(float complex *coeffs, int64_t row_elems, int band0,
                        int n_rows, uint64_t seed, uint64_t ikpt) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n_rows; ++r) {
    const int ib = band0 + r;
    float complex *c = coeffs + (int64_t)r * row_elems;
    uint64_t key = splitmix64(seed ^ (ikpt * 0xD1B54A32D192ED03ull) ^
                              ((uint64_t)ib * 0x8CB92BA72F3D8DD7ull));
    for (int64_t e = 0; e < row_elems; ++e) {
      uint64_t h = splitmix64(key + (uint64_t)e);
      float re = (float)((int32_t)(h & 0xFFFFFFFFu)) * (1.0f / 2147483648.0f);
      float im = (float)((int32_t)(h >> 32)) * (1.0f / 2147483648.0f);
      float env = 1.0f / (float)(1 + (int)(e % 7));
      c[e] = re * env + im * env * I;
    }
  }
}

void bo_synth_fill(float complex *coeffs, int64_t row_elems, int n_bands,
                   uint64_t seed, uint64_t ikpt) {
  bo_synth_fill_rows(coeffs, row_elems, 0, n_bands, seed, ikpt);
}
