/*
 * bandup_gpu.h -- C ABI of libbandup_gpu.so, the B200 (sm_100a) drop-in for the
 * weight loop of BandUP's module `band_unfolding`.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference's src/).  The reference has no FFI seam on this path today:
 * the procedures below are ordinary Fortran module procedures called from the
 * main loop (main_BandUP.f90:145-172).  The binding a maintainer adds is the
 * ISO_C_BINDING module fortran/bandup_gpu_mod.f90; the
 * reference's only C-interop precedent, spglib_f08, sets the style: plain
 * pointers, C ints by value, integer status returned.
 *
 * Conventions
 *  - Every function returns BANDUP_GPU_OK (0) or a positive error code; nothing
 *    aborts or throws across the boundary (the Fortran caller prints and
 *    `stop`s exactly where the reference does).
 *  - 3x3 matrices are passed in FORTRAN MEMORY ORDER (column-major) exactly as
 *    `crystal%rec_latt_vecs` lies in memory, where Fortran row i is lattice
 *    vector i: element (i,j) is at [ (j-1)*3 + (i-1) ].
 *  - Handles are small ints (Fortran integer(c_int)).  One caller thread per
 *    handle; handles on different devices are independent.
 *  - The caller owns every host array; the library owns device memory and its
 *    pinned result staging.
 */
#ifndef BANDUP_GPU_H_
#define BANDUP_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BANDUP_GPU_ABI_VERSION 5

/* status codes */
#define BANDUP_GPU_OK 0
#define BANDUP_GPU_ERR_INVALID_ARG 1
#define BANDUP_GPU_ERR_CUDA 2
/* band_unfolding_routines_mod.f90:132 (verify_commens): cells not commensurate */
#define BANDUP_GPU_ERR_NOT_COMMENSURATE 3
/* select_coeffs...:582-601 gave up: folding vector is not an SC reciprocal-lattice
 * vector within max_tol_for_vec_equality (1e-3) */
#define BANDUP_GPU_ERR_FOLDING_VECTOR 4
#define BANDUP_GPU_ERR_TOO_MANY_FOLDS 5
#define BANDUP_GPU_ERR_BAD_HANDLE 6
#define BANDUP_GPU_ERR_NOT_CONFIGURED 7 /* set_cells / set_grid not called yet */
#define BANDUP_GPU_ERR_UNKNOWN_KPT 8    /* fetch of an ikpt that is not in flight */
#define BANDUP_GPU_ERR_NO_DEVICE 9      /* no usable sm_100 device: there is NO CPU fallback */
#define BANDUP_GPU_ERR_CAPACITY 10      /* caller-provided output buffer too small */
/* calc_rho :1070-1074 'ERROR (calc_rho): delta_N = 0.' */
#define BANDUP_GPU_ERR_DELTA_N_ZERO 11
#define BANDUP_GPU_ERR_INT_OVERFLOW 12  /* Miller indices x adjugate would overflow */
#define BANDUP_GPU_ERR_NCCL 13          /* libnccl not found, or a NCCL call failed */

/* DISTINCT cosets (pc-kpts whose folding vectors differ by more than a pc reciprocal vector)
 * unfolded by ONE pass over the coefficients; a k-point with more of them (up to |det M|) is
 * worked off in ceil(n/64) passes inside the same launches, transparently.  pc-kpts that share
 * a coset are computed once and their rows replicated.  MAX_PCKPTS: most pc-kpts per submit. */
#define BANDUP_GPU_MAX_FOLDS 64
#define BANDUP_GPU_MAX_PCKPTS 4096

/* coefficient layouts accepted by submit_kpt / unfold_device */
/* pw_coeffs(n_spinor, n_pw, n_bands) as it lies in Fortran memory:
 * [band][pw][spinor] complex64 (constants_and_types_mod.f90:151) */
#define BANDUP_GPU_LAYOUT_FORTRAN_INMEM 0
/* raw WAVECAR band records: [band][spinor][pw] complex64, consecutive bands
 * band_stride_bytes (= nrecl) apart (read_wavecar_mod.f90:570-577) */
#define BANDUP_GPU_LAYOUT_VASP_RECORD 1

/* ---- lifetime --------------------------------------------------------- */

/* Creates a context on CUDA device `device` (no counterpart in the reference: called once where
 * program BandUP_main starts, main_BandUP.f90:52, and released before its end, :199).  Fails with
 * ERR_NO_DEVICE when the device is absent or is not compute capability 10.x. */
int bandup_gpu_init(int device, int *handle);
int bandup_gpu_finalize(int handle);
/* Human-readable text of the last error on this handle (never NULL). */
const char *bandup_gpu_last_error(int handle);
int bandup_gpu_abi_version(void);

/* Pinned host memory for the wavefunction reader to read into (where read_wavecar fills
 * wf%pw_coeffs, read_wavecar_mod.f90:537-583), so that submit_kpt's upload is a true asynchronous
 * DMA (optional). */
int bandup_gpu_pinned_alloc(size_t bytes, void **ptr);
int bandup_gpu_pinned_free(void *ptr);
/* The same for memory the caller already owns -- a Fortran ALLOCATEd wf%pw_coeffs, the buffer a
 * reader fills: page-locks [ptr, ptr + bytes) in place (cudaHostRegister) so uploads from it are
 * asynchronous DMA at full PCIe rate instead of staged copies.  Register once, reuse the buffer
 * for every k-point (locking costs about as much as allocating); unregister before freeing it. */
int bandup_gpu_host_register(void *ptr, size_t bytes);
int bandup_gpu_host_unregister(void *ptr);

/* ---- run-wide set-up -------------------------------------------------- */

/* Replaces check_if_pc_and_SC_are_commensurate (crystals_mod.f90:41-73) as
 * called from verify_commens: M = transpose(b_pc . inverse(B_sc)),
 * int_M = nint(M), commensurate iff max|M - int_M| <= tol (tol <= 0: 1e-5).
 * Stores int_M, det M and the integer adjugate used by the coset kernel.
 * M_out (nullable) receives int_M in Fortran order.  */
int bandup_gpu_set_cells(int handle, const double *B_sc, const double *b_pc,
                         double tol, int32_t *M_out);

/* Replaces real_seq(E_start, E_end, delta_e) (lists_and_seqs_mod.f90:188-200)
 * + the sigma choice of get_delta_Ns_for_EBS (band_unfolding_routines_mod.f90:
 * 774-777): smearing <= 0 means "argument absent" -> sigma = epsilon(1.0_dp)
 * (box integration); otherwise sigma = max(epsilon, |smearing|).
 * nener_out (nullable) receives the number of grid points. */
int bandup_gpu_set_grid(int handle, double E_start, double E_end, double delta_e,
                        double smearing, int32_t *nener_out);
/* Copies the energy grid (nener doubles) the kernels use: real_seq(E_start, E_end, delta_e)
 * (lists_and_seqs_mod.f90:188-200 as called at main_BandUP.f90:119-122). */
int bandup_gpu_get_grid(int handle, double *energies);

/* Replaces the float part of select_coeffs_to_calc_spectral_weights
 * (band_unfolding_routines_mod.f90:571-576): folding_G = Scoords(k) - K in
 * Cartesian 1/Angstrom, [n_fold][3].  Writes g0 = nint(folding_G . inverse(B_sc))
 * ([n_fold][3], SC Miller indices) and, when non-NULL, the largest Cartesian
 * residual |g0.B_sc - folding_G|.  Returns ERR_FOLDING_VECTOR when a residual
 * exceeds 1e-3 (the reference's max_tol_for_vec_equality). */
int bandup_gpu_folding_g0(int handle, const double *folding_G, int n_fold,
                          int32_t *g0, double *max_residual);

/* Replaces the search of get_geom_unfolding_relations / get_GUR_not_public
 * (band_unfolding_routines_mod.f90:306-483): for every pc-kpt (Cartesian, [n_pckpts][3]) the first
 * SC-kpt of the list -- the first n_sckpts_to_skip ignored -- having a symmetry-equivalent point
 * eqv with  vec_in_latt(pc_kpt - eqv, B_matrix_SC, tol)  (crystals_mod.f90:204-255, the
 * reference's floating-point recipe).  The equivalent points of SC-kpt s are
 * eqv_cart[eqv_offsets[s] .. eqv_offsets[s+1]) (Cartesian; they come from the host's symmetry
 * analysis, symmetry_mod.f90:311-484; without symmetry: the SC-kpt itself).  tol starts at 1e-5 and
 * grows by 5e-7 per attempt until every pc-kpt folds or 1e-3 is passed (:455-479).
 *   fold_sckpt [n_pckpts]  0-based SC-kpt index, -1 = folds onto none
 *   fold_eqv   [n_pckpts]  (nullable) index into eqv_cart of the point it folded onto
 *   tol_used   (nullable)  tolerance of the last attempt
 * ERR_FOLDING_VECTOR when some pc-kpt folds onto none (main_BandUP.f90:114); the maps are still
 * filled in. */
int bandup_gpu_fold_map(int handle, int n_pckpts, const double *pckpts_cart, int n_sckpts,
                        const int32_t *eqv_offsets, const double *eqv_cart, int n_sckpts_to_skip,
                        int32_t *fold_sckpt, int32_t *fold_eqv, double *tol_used);

/* ---- per SC-k-point, host buffers (the drop-in path) -------------------- */

/* Replaces, for ALL pc-kpts folding onto SC-kpt `ikpt` at once:
 *   the renormalisation in read_wavefunction   (io_routines_mod.f90:1317-1330)
 *   select_coeffs_to_calc_spectral_weights      (band_unfolding_routines_mod.f90:550-612)
 *   spectral_weight_for_coeff + spinor sum      (:615-651, :1335-1345)
 *   get_delta_Ns_for_EBS                        (:760-786, :719-757)
 * Asynchronous: enqueues upload + kernels and returns; up to
 * bandup_gpu_pipeline_depth() k-points may be in flight (double buffering); a
 * further submit is refused with ERR_INVALID_ARG until one of them has been fetched.
 *   coeffs         host pointer (pinned or pageable), layout as `layout`;
 *                  NOT renormalised by the caller (renormalised input is also
 *                  fine: the division by the band norm is then a no-op)
 *   g_frac         int32 [n_pw][3] == wf%G_frac (read_wavecar_mod.f90:296-298)
 *   band_energies  double [n_bands] == wf%band_energies
 *   g0             int32 [n_fold][3] from bandup_gpu_folding_g0
 */
int bandup_gpu_submit_kpt(int handle, int ikpt, int n_spinor, int n_pw,
                          int n_bands, int layout, int64_t band_stride_bytes,
                          const void *coeffs, const int32_t *g_frac,
                          const double *band_energies, int n_fold,
                          const int32_t *g0);

/* The VASP reader's path without its host work: `band_records` are the n_bands raw WAVECAR
 * coefficient records of the k-point exactly as they lie in the file (nrecl bytes apart,
 * spinor-up block then spinor-down block; read_wavecar_mod.f90:570-577), `nplane` is the count
 * in the k-point's header record, and the plane-wave list is regenerated ON THE DEVICE from
 * the k-point, ENCUT and the SC lattice given to bandup_gpu_set_cells, in the reader's own
 * enumeration order (n_Gvecs_2b_generated + populate_wf_with_G_vec_coords, :213-304),
 * including its scalar/spinor decision (nint(nplane / n_Gvecs) == 2, :446-452) and its
 * adaptive adjustment of 2m/hbar^2 (steps of 1e-10, at most 1000) when the count is off
 * (:459-472).  two_m_over_hbar_sqrd <= 0 selects the reference's constant (:107).
 * n_spinor_out / n_pw_out (nullable) receive what was decided.  Otherwise as
 * bandup_gpu_submit_kpt; ERR_INVALID_ARG when no adjustment yields the file's count. */
int bandup_gpu_submit_kpt_vasp(int handle, int ikpt, int nplane, int n_bands, int64_t nrecl,
                               const void *band_records, const double *kpt_frac_coords,
                               double ecut, double two_m_over_hbar_sqrd,
                               const double *band_energies, int n_fold, const int32_t *g0,
                               int *n_spinor_out, int *n_pw_out);
/* The plane-wave list (int32 [n_pw][3]) of a submitted or still resident k-point: wf%G_frac as
 * read_wavecar leaves it (read_wavecar_mod.f90:275-302). */
int bandup_gpu_fetch_gvecs(int handle, int ikpt, int32_t *g_frac, int64_t capacity);

/* Blocks until k-point `ikpt` is done and copies out:
 *   weights    [n_fold][n_bands]  P_mK(k)           (spectral_weight, :1333-1345)
 *   dN         [n_fold][nener]    delta_N(k, E_i)   (%dN, :1355-1360)
 *   n_selected [n_fold]  size(selected_coeff_indices); 0 <=> the reference's
 *              "pckpt cannot be parsed" branch (main_BandUP.f90:167-171)
 *   norms      [n_bands] (nullable) sum_G |C|^2 before renormalisation
 *   selected   (nullable) 1-based ascending plane-wave indices, the folds'
 *              lists concatenated (CSR by n_selected); selected_capacity is
 *              the number of int32 available.
 * Any of weights / dN / n_selected may be NULL.  ERR_CAPACITY (selected_capacity too
 * small; the message names the size needed) consumes nothing: the k-point stays in
 * flight and may be fetched again with a larger buffer. */
int bandup_gpu_fetch_kpt(int handle, int ikpt, double *weights, double *dN,
                         int32_t *n_selected, double *norms, int32_t *selected,
                         int64_t selected_capacity);

/* SC k-points that may be in flight at once (submitted, not fetched): 2, so that the reader can
 * fill k-point j+1 (read_wavefunction, main_BandUP.f90:150-153) while j uploads and computes. */
int bandup_gpu_pipeline_depth(int handle);

/* args%perform_unfold (cla_wrappers_mod.f90, flag -dont_unfold): with 0 every spectral weight
 * is 1 (band_unfolding_routines_mod.f90:1346-1351) and delta_N counts the SC bands.  Default 1. */
int bandup_gpu_set_perform_unfold(int handle, int perform_unfold);

/* Replaces calc_spectral_function (band_unfolding_routines_mod.f90:654-698) for the resident
 * SC-kpt `ikpt` (see the residency rule below): sf[fold][iener] = sum_m P_m * Gaussian(E_m - E_i)
 * with standard deviation std_dev (<= 0: the reference's default 0.025 eV).  The reference
 * compiles this step out (calc_spec_func_explicitly = .FALSE.); it is offered on request.
 * sf [n_fold][nener] HOST. */
int bandup_gpu_spectral_function(int handle, int ikpt, double std_dev, double *sf);

/* ---- projected variant: unfolding-density operator ---------------------- */
/* (`bandup unfold --orbitals` = -write_unf_dens_op, consumed by `projected-unfold`)
 *
 * The device buffers of SC-kpt `ikpt` stay valid after bandup_gpu_fetch_kpt until
 * bandup_gpu_pipeline_depth() further submits have reused them; the calls below work
 * on that resident k-point.
 *
 * bandup_gpu_rho_count replaces, for every pc-kpt of the k-point, the energy selection
 * and the calc_rho loop of perform_unfolding (band_unfolding_routines_mod.f90:1372-1428,
 * calc_rho :1046-1161): for every energy with dN >= min_dN (reference: 1e-3) the operator
 *   rho(m1,m2) = lambda(m1) lambda(m2) / dN * sum_spinor sum_{G in sel} conj(C_m1) C_m2
 * between the bands with lambda >= 1e-5 (m1 = 1..n_bands-1 only, prefactor >= 1e-5); lambda is
 * the box (sigma = epsilon) whatever smearing the grid was given, because perform_unfolding
 * calls calc_rho without std_dev (:1404-1407).  After bandup_gpu_set_perform_unfold(0) it is the
 * other branch of calc_rho (:1136-1154): rho(m1,m2) = 1 / (number of bands with lambda >= 1e-5)
 * for every coupled pair, m1 over ALL bands.
 * It blocks until done and returns
 *   n_energies  number of (pc-kpt, energy) operators with at least one coupled band pair
 *   n_entries   number of stored elements, |rho| >= 1e-5, both triangles (:1416-1425)
 * min_dN < epsilon(1.0_dp) is refused with ERR_DELTA_N_ZERO (calc_rho's fatal branch). */
int bandup_gpu_rho_count(int handle, int ikpt, double min_dN, int64_t *n_energies,
                         int64_t *n_entries);
/* Copies the elements counted by the last bandup_gpu_rho_count(ikpt), in the order
 * perform_unfolding appends them to rhos(iener2)%rho / %band_indices
 * (band_unfolding_routines_mod.f90:1416-1427; pc-kpt, energy ascending, m1 outer, m2 inner):
 *   fold   0-based index into the submit's pc-kpt list
 *   iener  1-based index into the energy grid   (iener_in_full_pc_egrid)
 *   m1,m2  1-based band indices                 (band_indices)
 *   rho_re, rho_im  complex(sp) value           (rho) */
int bandup_gpu_fetch_rho(int handle, int ikpt, int64_t capacity, int32_t *fold, int32_t *iener,
                         int32_t *m1, int32_t *m2, float *rho_re, float *rho_im);

/* Spinor wavefunctions: replaces get_SC_pauli_matrix_elmnts + the trace loop of
 * perform_unfolding (band_unfolding_routines_mod.f90:1164-1263, :1459-1481) for every operator
 * counted by the last bandup_gpu_rho_count(ikpt):
 *   sigma[fold][iener][i] = Re Tr(rho . sigma_i),  i = x, y, z      (0 where there is no rho)
 * the three spin columns of the reference's output file (io_routines_mod.f90:950-960); the
 * in-plane projections (calc_spin_projections, :1266-1286) are 3-vector algebra the caller
 * does from these.  sigma [n_fold][nener][3] HOST.  ERR_INVALID_ARG for n_spinor = 1. */
int bandup_gpu_pauli_vectors(int handle, int ikpt, double *sigma);

/* Replaces KptInfo.get_orbital_contribution_matrix with use_dual=True
 * (python_interface/bandupy/orbital_contributions.py:226-362):
 *   O(m1,m2) = sum_{a: ao_mask[a]} dual[m1,a] * proj[a,m2]
 * kept when |E_m2 - E_m1| <= max_band_dE (0.1 eV) and |O| >= min_abs (1e-2), else 0.
 * dual [n_bands][n_ao], proj [n_ao][n_bands], O_out [n_bands][n_bands]: complex128
 * interleaved (re,im), row-major, HOST.  ao_mask[a] != 0 selects the combined index
 * a = iat*n_orbs_per_atom + iorb of the picked atoms and orbitals.  The matrix stays on
 * the device as the handle's current operator; O_out may be NULL. */
int bandup_gpu_orb_contrib_matrix(int handle, int n_bands, int n_ao, const double *dual,
                                  const double *proj, const uint8_t *ao_mask,
                                  const double *band_energies, double max_band_dE,
                                  double min_abs, double *O_out);
/* Uploads a general operator O [n_bands][n_bands] complex128 (HOST) as the current one: the
 * `general_operator` argument of UnfDensOp.unfold (unfolding_density_op.py:121-126). */
int bandup_gpu_set_operator(int handle, int n_bands, const double *O);
/* Replaces UnfDensOp.unfold(O, multiply_by_N=True, discard_imag=True, clip_interval=(0,N))
 * (unfolding_density_op.py:121-141 as called from orbital_contributions.py:444-458) for every
 * operator counted by the last bandup_gpu_rho_count(ikpt):
 *   out[fold][iener] = clip(dN * Re Tr(rho . O), 0, dN)    (0 where there is no rho)
 * rho is Hermitian-completed from its upper triangle with a real diagonal, as the Python
 * class does; elements with |rho| < min_abs_rho are dropped (the reference's rho file keeps
 * |rho| >= 1e-4, io_routines_mod.f90:1861).  clip = 0 returns dN * Re Tr(rho.O) unclipped.
 * out [n_fold][nener] HOST. */
int bandup_gpu_unfold_operator(int handle, int ikpt, double min_abs_rho, int clip, double *out);

/* ---- symmetry averaging over the needed directions ---------------------- */
/* After the SC-kpt loop the reference averages what every pc-kpt got along the symmetry-
 * equivalent ("needed") directions of its pc-BZ direction, with the directions' weights
 * (get_delta_Ns_for_output, band_unfolding_routines_mod.f90:789-1043).  The directions and their
 * weights come from the caller's symmetry analysis (spglib in BandUP); with -no_symm_avg there
 * is one direction of weight 1 and these calls are the identity.
 *
 * bandup_gpu_symm_average replaces the dense loops (:826-853 delta_N; :984-1036 spin_proj_perp,
 * spin_proj_para, sigma):  out[r] = sum_d weights[d] * in[d][r], accumulated in dp in direction
 * order exactly as `avrgd = avrgd + weight*x` does.  in [n_dirs][n_values], out [n_values], HOST. */
int bandup_gpu_symm_average(int handle, int n_dirs, int64_t n_values, const double *weights,
                            const double *in, double *out);
/* Replaces the symmetry average of the unfolding-density operators (:857-973).  Input: the stored
 * elements of all n_dirs directions concatenated, direction d = [dir_offsets[d], dir_offsets[d+1]),
 * each direction in the order bandup_gpu_fetch_rho returns them (pckpt, iener ascending, m1 outer,
 * m2 inner; no element twice) with `pckpt` the 0-based index of the pc-kpt along the direction
 * (< n_pckpts), iener/m1/m2 1-based.  avg_dN [n_pckpts][nener] is the symmetry-averaged delta_N
 * (bandup_gpu_symm_average); operators at energies with avg_dN < min_dN are dropped (:906, the
 * reference's literal 1E-3 is default real: pass (double)1e-3f for bit parity).  Every remaining
 * (m1,m2) is created by the first direction that stores it as cmplx(weight*rho, kind=sp) and
 * updated by the later ones as sp(dp(avg) + weight*rho) (:934-957).
 * Output (HOST, `capacity` entries each; capacity = dir_offsets[n_dirs] always suffices): the
 * averaged elements by (pckpt, iener) and, inside one operator, in the reference's append order
 * (direction, then m1 outer, m2 inner).  *n_out = their number; when it exceeds `capacity` nothing
 * is copied and BANDUP_GPU_ERR_INVALID_ARG is returned.  Unsorted input is ERR_INVALID_ARG. */
int bandup_gpu_symm_average_rho(int handle, int n_dirs, int n_pckpts, int nener, int n_bands,
                                const double *weights, const double *avg_dN, double min_dN,
                                const int64_t *dir_offsets, const int32_t *pckpt,
                                const int32_t *iener, const int32_t *m1, const int32_t *m2,
                                const float *rho_re, const float *rho_im, int64_t capacity,
                                int64_t *n_out, int32_t *pckpt_out, int32_t *iener_out,
                                int32_t *m1_out, int32_t *m2_out, float *rho_re_out,
                                float *rho_im_out);

/* ---- per SC-k-point, device-resident buffers ---------------------------- */

/* Scratch bytes unfold_device needs for these sizes. */
size_t bandup_gpu_workspace_bytes(int handle, int n_spinor, int n_pw, int n_bands,
                                  int n_fold);

/* Same computation as submit+fetch -- the body of the loop over the pc-kpts folding onto one
 * SC-kpt, main_BandUP.f90:145-172 -- but on buffers already in HBM; everything
 * is enqueued on `stream` (a cudaStream_t; NULL = default stream), nothing is
 * synchronised.  All d_* pointers are device pointers; g0 is a HOST pointer.
 * d_norms, d_slot_out (uint8 [n_pw]: fold id or 255) may be NULL; d_slot_out needs
 * <= BANDUP_GPU_MAX_FOLDS distinct cosets (ERR_TOO_MANY_FOLDS otherwise). */
int bandup_gpu_unfold_device(int handle, void *stream, int n_spinor, int n_pw,
                             int n_bands, int layout, int64_t band_stride_bytes,
                             const void *d_coeffs, const int32_t *d_g_frac,
                             const double *d_band_energies, int n_fold,
                             const int32_t *g0, double *d_weights, double *d_dN,
                             int32_t *d_n_selected, double *d_norms,
                             uint8_t *d_slot_out, void *d_workspace,
                             size_t workspace_bytes);

/* One SC k-point of a batch for bandup_gpu_unfold_device_batch: the arguments of
 * bandup_gpu_unfold_device as a struct (d_* device pointers, g0 HOST [n_fold][3]). */
typedef struct bandup_gpu_kpt_desc {
  int32_t n_spinor, n_pw, n_bands, layout, n_fold, reserved;
  int64_t band_stride_bytes;
  const void *d_coeffs;
  const int32_t *d_g_frac;
  const double *d_band_energies;
  const int32_t *g0;
  double *d_weights;      /* [n_fold][n_bands] */
  double *d_dN;           /* [n_fold][nener]   */
  int32_t *d_n_selected;  /* [n_fold]          */
  double *d_norms;        /* [n_bands] or NULL */
} bandup_gpu_kpt_desc;

size_t bandup_gpu_batch_workspace_bytes(int handle, int n_kpts, const bandup_gpu_kpt_desc *descs);
/* The same computation as n_kpts calls of bandup_gpu_unfold_device, but every kernel is
 * launched ONCE for the whole batch (three launches in total): the k-points' (band, chunk)
 * tiles form one persistent grid, so small k-points (graphene: 12 MB) are no longer
 * launch-bound and large ones lose the per-k-point ramp and tail.  All k-points of a batch
 * share n_spinor and layout. */
int bandup_gpu_unfold_device_batch(int handle, void *stream, int n_kpts,
                                   const bandup_gpu_kpt_desc *descs, void *d_workspace,
                                   size_t workspace_bytes);

/* ---- multi-GPU: one process per GPU, SC k-points sharded, delta_N gathered ---- */
/* SC k-points are independent and every pc-kpt folds onto exactly one SC-kpt
 * (band_unfolding_routines_mod.f90:317-324, 396-403), so ranks exchange nothing but their
 * disjoint delta_N rows at the end: one ncclAllGather over NVLink.  libnccl is resolved at run
 * time (dlopen; $BANDUP_GPU_NCCL_LIB overrides the search).
 *   rank 0 calls bandup_gpu_comm_unique_id and hands the 128 bytes to the other processes by
 *   whatever means the host has (MPI_Bcast in a Fortran host, a file, torch.distributed ...);
 *   every rank calls bandup_gpu_comm_init with them. */
#define BANDUP_GPU_COMM_ID_BYTES 128
int bandup_gpu_comm_unique_id(void *id);
int bandup_gpu_comm_init(int handle, int n_ranks, int rank, const void *id);
int bandup_gpu_comm_finalize(int handle);
/* Collective.  The SC-kpt loop of the reference (main_BandUP.f90:131-177) is what the ranks share;
 * the table it fills, delta_N%...%pckpt(:)%dN, is consumed after the loop by get_delta_Ns_for_output
 * and the writers (:187-191), which is where every rank needs all rows.
 * dN_local [n_local][nener] + the global pc-kpt row id of every row (HOST); every
 * rank receives all rows in ascending row id: dN_all [n_total][nener], row_ids_all [n_total]
 * (HOST, capacity_rows rows).  With dN_all == NULL only *n_total is returned (the counts are
 * exchanged, no rows move).  A row id
 * contributed twice is an error (a pc-kpt unfolded by two ranks). */
int bandup_gpu_gather_dN(int handle, const double *dN_local, const int64_t *row_ids_local,
                         int64_t n_local, double *dN_all, int64_t *row_ids_all,
                         int64_t capacity_rows, int64_t *n_total);

/* The same collective for rows of any length (the spin columns perform_unfolding fills next to dN,
 * band_unfolding_routines_mod.f90:1431-1508: sigma rows of 3*nener doubles). */
int bandup_gpu_gather_rows(int handle, int row_len, const double *rows_local,
                           const int64_t *row_ids_local, int64_t n_local, double *rows_all,
                           int64_t *row_ids_all, int64_t capacity_rows, int64_t *n_total);

/* The device-resident form of the collective, for callers whose rows never leave HBM
 * (bandup_gpu_unfold_device[_batch]): every rank contributes `count` doubles at d_send (pad to
 * the largest shard), d_recv [n_ranks][count] receives rank r's block at r*count.  One
 * ncclAllGather enqueued on the caller's `stream` (cudaStream_t); nothing is synchronised and
 * nothing touches the host. */
int bandup_gpu_allgather_device(int handle, void *stream, const double *d_send, double *d_recv,
                                int64_t count);

/* ---- instrumentation ---------------------------------------------------- */

/* Kernel ids for bandup_gpu_kernel_time */
#define BANDUP_GPU_K_COSET 0
#define BANDUP_GPU_K_WEIGHTS 1
#define BANDUP_GPU_K_FINALIZE 2  /* fused into K_DELTA_N: reports 0 launches */
#define BANDUP_GPU_K_DELTA_N 3
#define BANDUP_GPU_K_RHO 4
#define BANDUP_GPU_K_SYMM_AVG 5
#define BANDUP_GPU_K_COUNT 6

/* When enabled, every launch is bracketed by CUDA events on its own stream (the device-side
 * counterpart of the reference's `times` bookkeeping, print_final_times,
 * io_routines_mod.f90:1923-1962). */
int bandup_gpu_set_profiling(int handle, int enabled);
/* Same, for the kernels whose id bit is set only (bit k = kernel id k): an event pair between two
 * back-to-back launches costs a few microseconds of device time, so a timed region that wants one
 * kernel's duration brackets that kernel alone. */
int bandup_gpu_set_profiling_mask(int handle, unsigned kernel_mask);
/* Synchronises the device, then returns summed milliseconds and launch count
 * of kernel `which` since the last reset. */
int bandup_gpu_kernel_time(int handle, int which, double *total_ms,
                           int64_t *launches);
int bandup_gpu_reset_kernel_times(int handle);
/* Total number of kernels this handle has launched (never reset). */
int64_t bandup_gpu_launch_count(int handle);

/* Tuning knobs of the weights kernel (0 = keep default): tile size in 16-byte
 * vectors per (band, chunk) tile (rounded up to whole 1024-vector stages),
 * resident CTAs per SM and depth of each CTA's shared-memory stage ring. */
int bandup_gpu_set_tuning(int handle, int chunk_vecs, int ctas_per_sm, int n_stages);

/* Further switches (not while k-points are in flight; the first two exist for A/B measurements):
 *   OPT_K2_VARIANT        accumulators of the weights kernel: 0 = automatic, 1 = per-thread
 *                         shared-memory bins (any number of cosets), 2 = registers whenever a
 *                         batch has <= 4 distinct cosets per SC-kpt
 *   OPT_DEPENDENT_LAUNCH  programmatic dependent launch inside a batch: bit 0 = K2 behind K1,
 *                         bit 1 = K3 behind K2; 3 (default) = both, 0 = plain stream order
 *   OPT_COMPLEX128        the reference's `kind_cplx_coeffs = dp` build
 *                         (constants_and_types_mod.f90:60; needed for a double-precision WAVECAR,
 *                         nprec = 45210, read_wavecar_mod.f90:363-377): 1 = EVERY coefficient
 *                         pointer handed to this handle refers to complex128 elements (16 bytes;
 *                         band_stride_bytes a multiple of 16, device blocks 16-byte aligned).
 *                         Norms, |C|^2 and all sums are then fp64 throughout; rho and the Pauli
 *                         vectors are formed from the renormalised coefficients rounded to
 *                         complex64 (their outputs are single precision in this ABI).  Default 0. */
#define BANDUP_GPU_OPT_K2_VARIANT 0
#define BANDUP_GPU_OPT_DEPENDENT_LAUNCH 1
#define BANDUP_GPU_OPT_COMPLEX128 2
int bandup_gpu_set_option(int handle, int option, int value);

/* Fills a device buffer with the synthetic coefficients bench.py and the tests
 * use (same hash as oracle/bandup_oracle.c:bo_synth_fill): n_bands rows of
 * row_elems complex64, rows band_stride_bytes apart. */
int bandup_gpu_synth_fill_device(int handle, void *stream, void *d_coeffs,
                                 int64_t row_elems, int n_bands,
                                 int64_t band_stride_bytes, uint64_t seed,
                                 uint64_t ikpt);

#ifdef __cplusplus
}
#endif
#endif /* BANDUP_GPU_H_ */
