!! bandup_gpu_mod.f90 -- ISO_C_BINDING interface of libbandup_gpu.so (include/bandup_gpu.h).
!!
!! This is the binding a BandUP maintainer adds to src/ to call the B200 path from
!! the main loop of main_BandUP.f90 (see INTEGRATION.md for the patch and the link
!! line).  Style follows the reference's only C interop, module spglib_f08
!! (src/external/spglib-1.5.2/example/spglib_f08.f90): `interface` blocks with
!! bind(c), scalars by value, arrays as assumed-size, C ints as status.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran front end
!! (gfortran/flang/ifx absent).  tests/test_abi.py checks every bind(c) name,
!! argument count and kind against include/bandup_gpu.h so the two cannot drift, and
!! bandup_b200/unfold.py (ctypes) makes the identical sequence of C-ABI calls.
module bandup_gpu
use, intrinsic :: iso_c_binding
implicit none
private
public :: bandup_gpu_init, bandup_gpu_finalize, bandup_gpu_last_error, &
          bandup_gpu_abi_version, bandup_gpu_pinned_alloc, bandup_gpu_pinned_free, &
          bandup_gpu_host_register, bandup_gpu_host_unregister, &
          bandup_gpu_set_cells, bandup_gpu_set_grid, bandup_gpu_get_grid, &
          bandup_gpu_folding_g0, bandup_gpu_fold_map, bandup_gpu_submit_kpt, bandup_gpu_submit_kpt_vasp, &
          bandup_gpu_fetch_kpt, bandup_gpu_fetch_gvecs, &
          bandup_gpu_pipeline_depth, bandup_gpu_set_profiling, &
          bandup_gpu_kernel_time, bandup_gpu_reset_kernel_times, &
          bandup_gpu_launch_count, bandup_gpu_set_tuning, bandup_gpu_set_option, &
          bandup_gpu_rho_count, bandup_gpu_fetch_rho, bandup_gpu_pauli_vectors, &
          bandup_gpu_spectral_function, bandup_gpu_set_perform_unfold, &
          bandup_gpu_orb_contrib_matrix, bandup_gpu_set_operator, bandup_gpu_unfold_operator, &
          bandup_gpu_symm_average, bandup_gpu_symm_average_rho, &
          bandup_gpu_comm_unique_id, bandup_gpu_comm_init, bandup_gpu_comm_finalize, bandup_gpu_gather_dN, &
          bandup_gpu_gather_rows, bandup_gpu_allgather_device, &
          bandup_gpu_unfold_kpt_f, bandup_gpu_error_message
public :: BANDUP_GPU_OK, BANDUP_GPU_LAYOUT_FORTRAN_INMEM, BANDUP_GPU_LAYOUT_VASP_RECORD, &
          BANDUP_GPU_MAX_FOLDS, BANDUP_GPU_MAX_PCKPTS, BANDUP_GPU_COMM_ID_BYTES, &
          BANDUP_GPU_ERR_INVALID_ARG, BANDUP_GPU_ERR_CUDA, BANDUP_GPU_ERR_NOT_COMMENSURATE, &
          BANDUP_GPU_ERR_FOLDING_VECTOR, BANDUP_GPU_ERR_TOO_MANY_FOLDS, BANDUP_GPU_ERR_BAD_HANDLE, &
          BANDUP_GPU_ERR_NOT_CONFIGURED, BANDUP_GPU_ERR_UNKNOWN_KPT, BANDUP_GPU_ERR_NO_DEVICE, &
          BANDUP_GPU_ERR_CAPACITY, BANDUP_GPU_ERR_DELTA_N_ZERO, BANDUP_GPU_ERR_INT_OVERFLOW, &
          BANDUP_GPU_ERR_NCCL, &
          BANDUP_GPU_OPT_K2_VARIANT, BANDUP_GPU_OPT_DEPENDENT_LAUNCH, BANDUP_GPU_OPT_COMPLEX128

integer(c_int), parameter :: BANDUP_GPU_OK = 0
integer(c_int), parameter :: BANDUP_GPU_LAYOUT_FORTRAN_INMEM = 0
integer(c_int), parameter :: BANDUP_GPU_LAYOUT_VASP_RECORD = 1
integer(c_int), parameter :: BANDUP_GPU_MAX_FOLDS = 64
integer(c_int), parameter :: BANDUP_GPU_MAX_PCKPTS = 4096
integer(c_int), parameter :: BANDUP_GPU_COMM_ID_BYTES = 128
! status codes (include/bandup_gpu.h); the message is bandup_gpu_error_message(handle)
integer(c_int), parameter :: BANDUP_GPU_ERR_INVALID_ARG = 1
integer(c_int), parameter :: BANDUP_GPU_ERR_CUDA = 2
integer(c_int), parameter :: BANDUP_GPU_ERR_NOT_COMMENSURATE = 3
integer(c_int), parameter :: BANDUP_GPU_ERR_FOLDING_VECTOR = 4
integer(c_int), parameter :: BANDUP_GPU_ERR_TOO_MANY_FOLDS = 5
integer(c_int), parameter :: BANDUP_GPU_ERR_BAD_HANDLE = 6
integer(c_int), parameter :: BANDUP_GPU_ERR_NOT_CONFIGURED = 7
integer(c_int), parameter :: BANDUP_GPU_ERR_UNKNOWN_KPT = 8
integer(c_int), parameter :: BANDUP_GPU_ERR_NO_DEVICE = 9
integer(c_int), parameter :: BANDUP_GPU_ERR_CAPACITY = 10
integer(c_int), parameter :: BANDUP_GPU_ERR_DELTA_N_ZERO = 11
integer(c_int), parameter :: BANDUP_GPU_ERR_INT_OVERFLOW = 12
integer(c_int), parameter :: BANDUP_GPU_ERR_NCCL = 13
! bandup_gpu_set_option
integer(c_int), parameter :: BANDUP_GPU_OPT_K2_VARIANT = 0
integer(c_int), parameter :: BANDUP_GPU_OPT_DEPENDENT_LAUNCH = 1
integer(c_int), parameter :: BANDUP_GPU_OPT_COMPLEX128 = 2   ! kind_cplx_coeffs = dp builds

interface

    function bandup_gpu_abi_version() bind(c, name='bandup_gpu_abi_version')
        import c_int
        integer(c_int) :: bandup_gpu_abi_version
    end function bandup_gpu_abi_version

    function bandup_gpu_init(device, handle) bind(c, name='bandup_gpu_init')
        import c_int
        integer(c_int), intent(in), value :: device
        integer(c_int), intent(out) :: handle
        integer(c_int) :: bandup_gpu_init
    end function bandup_gpu_init

    function bandup_gpu_finalize(handle) bind(c, name='bandup_gpu_finalize')
        import c_int
        integer(c_int), intent(in), value :: handle
        integer(c_int) :: bandup_gpu_finalize
    end function bandup_gpu_finalize

    function bandup_gpu_last_error(handle) bind(c, name='bandup_gpu_last_error')
        import c_int, c_ptr
        integer(c_int), intent(in), value :: handle
        type(c_ptr) :: bandup_gpu_last_error
    end function bandup_gpu_last_error

    function bandup_gpu_pinned_alloc(bytes, ptr) bind(c, name='bandup_gpu_pinned_alloc')
        import c_int, c_size_t, c_ptr
        integer(c_size_t), intent(in), value :: bytes
        type(c_ptr), intent(out) :: ptr
        integer(c_int) :: bandup_gpu_pinned_alloc
    end function bandup_gpu_pinned_alloc

    function bandup_gpu_pinned_free(ptr) bind(c, name='bandup_gpu_pinned_free')
        import c_int, c_ptr
        type(c_ptr), intent(in), value :: ptr
        integer(c_int) :: bandup_gpu_pinned_free
    end function bandup_gpu_pinned_free

    ! page-lock an array Fortran allocated (c_loc(wf%pw_coeffs), storage_size/8 * size bytes)
    function bandup_gpu_host_register(ptr, bytes) bind(c, name='bandup_gpu_host_register')
        import c_int, c_ptr, c_size_t
        type(c_ptr), intent(in), value :: ptr
        integer(c_size_t), intent(in), value :: bytes
        integer(c_int) :: bandup_gpu_host_register
    end function bandup_gpu_host_register

    function bandup_gpu_host_unregister(ptr) bind(c, name='bandup_gpu_host_unregister')
        import c_int, c_ptr
        type(c_ptr), intent(in), value :: ptr
        integer(c_int) :: bandup_gpu_host_unregister
    end function bandup_gpu_host_unregister

    ! crystal%rec_latt_vecs(1:3,1:3) is passed as it lies in memory (column-major,
    ! row i = vector i); M_out receives int_M the same way.
    function bandup_gpu_set_cells(handle, B_sc, b_pc, tol, M_out) &
        bind(c, name='bandup_gpu_set_cells')
        import c_int, c_double, c_int32_t
        integer(c_int), intent(in), value :: handle
        real(c_double), intent(in) :: B_sc(3,3), b_pc(3,3)
        real(c_double), intent(in), value :: tol
        integer(c_int32_t), intent(out) :: M_out(3,3)
        integer(c_int) :: bandup_gpu_set_cells
    end function bandup_gpu_set_cells

    function bandup_gpu_set_grid(handle, E_start, E_end, delta_e, smearing, nener_out) &
        bind(c, name='bandup_gpu_set_grid')
        import c_int, c_double, c_int32_t
        integer(c_int), intent(in), value :: handle
        real(c_double), intent(in), value :: E_start, E_end, delta_e, smearing
        integer(c_int32_t), intent(out) :: nener_out
        integer(c_int) :: bandup_gpu_set_grid
    end function bandup_gpu_set_grid

    function bandup_gpu_get_grid(handle, energies) bind(c, name='bandup_gpu_get_grid')
        import c_int, c_double
        integer(c_int), intent(in), value :: handle
        real(c_double), intent(out) :: energies(*)
        integer(c_int) :: bandup_gpu_get_grid
    end function bandup_gpu_get_grid

    function bandup_gpu_folding_g0(handle, folding_G, n_fold, g0, max_residual) &
        bind(c, name='bandup_gpu_folding_g0')
        import c_int, c_double, c_int32_t
        integer(c_int), intent(in), value :: handle
        real(c_double), intent(in) :: folding_G(3,*)
        integer(c_int), intent(in), value :: n_fold
        integer(c_int32_t), intent(out) :: g0(3,*)
        real(c_double), intent(out) :: max_residual
        integer(c_int) :: bandup_gpu_folding_g0
    end function bandup_gpu_folding_g0

    ! get_geom_unfolding_relations (band_unfolding_routines_mod.f90:430-483) on the device
    function bandup_gpu_fold_map(handle, n_pckpts, pckpts_cart, n_sckpts, eqv_offsets, eqv_cart, &
                                 n_sckpts_to_skip, fold_sckpt, fold_eqv, tol_used) &
        bind(c, name='bandup_gpu_fold_map')
        import c_int, c_int32_t, c_double
        integer(c_int), intent(in), value :: handle, n_pckpts, n_sckpts, n_sckpts_to_skip
        real(c_double), intent(in) :: pckpts_cart(3,*), eqv_cart(3,*)
        integer(c_int32_t), intent(in) :: eqv_offsets(*)
        integer(c_int32_t), intent(out) :: fold_sckpt(*), fold_eqv(*)
        real(c_double), intent(out) :: tol_used
        integer(c_int) :: bandup_gpu_fold_map
    end function bandup_gpu_fold_map

    ! coeffs = c_loc(wf%pw_coeffs), g_frac = c_loc(wf%G_frac): type(vec3d_int) is three
    ! default integers, i.e. int32 [n_pw][3] in memory.
    function bandup_gpu_submit_kpt(handle, ikpt, n_spinor, n_pw, n_bands, layout, &
                                   band_stride_bytes, coeffs, g_frac, band_energies, &
                                   n_fold, g0) bind(c, name='bandup_gpu_submit_kpt')
        import c_int, c_int64_t, c_int32_t, c_double, c_ptr
        integer(c_int), intent(in), value :: handle, ikpt, n_spinor, n_pw, n_bands, layout
        integer(c_int64_t), intent(in), value :: band_stride_bytes
        type(c_ptr), intent(in), value :: coeffs, g_frac
        real(c_double), intent(in) :: band_energies(*)
        integer(c_int), intent(in), value :: n_fold
        integer(c_int32_t), intent(in) :: g0(3,*)
        integer(c_int) :: bandup_gpu_submit_kpt
    end function bandup_gpu_submit_kpt

    ! raw WAVECAR band records + plane-wave list regenerated on the device (read_wavecar_mod.f90:213-304)
    function bandup_gpu_submit_kpt_vasp(handle, ikpt, nplane, n_bands, nrecl, band_records, &
                                        kpt_frac_coords, ecut, two_m_over_hbar_sqrd, band_energies, &
                                        n_fold, g0, n_spinor_out, n_pw_out) &
        bind(c, name='bandup_gpu_submit_kpt_vasp')
        import c_int, c_int64_t, c_int32_t, c_double, c_ptr
        integer(c_int), intent(in), value :: handle, ikpt, nplane, n_bands
        integer(c_int64_t), intent(in), value :: nrecl
        type(c_ptr), intent(in), value :: band_records
        real(c_double), intent(in) :: kpt_frac_coords(3)
        real(c_double), intent(in), value :: ecut, two_m_over_hbar_sqrd
        real(c_double), intent(in) :: band_energies(*)
        integer(c_int), intent(in), value :: n_fold
        integer(c_int32_t), intent(in) :: g0(3,*)
        integer(c_int), intent(out) :: n_spinor_out, n_pw_out
        integer(c_int) :: bandup_gpu_submit_kpt_vasp
    end function bandup_gpu_submit_kpt_vasp

    ! wf%G_frac of a k-point submitted through bandup_gpu_submit_kpt_vasp (the list was regenerated
    ! on the device, read_wavecar_mod.f90:275-302): g_frac(3, capacity), capacity >= n_pw
    function bandup_gpu_fetch_gvecs(handle, ikpt, g_frac, capacity) bind(c, name='bandup_gpu_fetch_gvecs')
        import c_int, c_int32_t, c_int64_t
        integer(c_int), intent(in), value :: handle, ikpt
        integer(c_int32_t), intent(out) :: g_frac(3,*)
        integer(c_int64_t), intent(in), value :: capacity
        integer(c_int) :: bandup_gpu_fetch_gvecs
    end function bandup_gpu_fetch_gvecs

    ! weights(n_bands, n_fold), dN(nener, n_fold) in Fortran order.  Optional C outputs
    ! are passed as type(c_ptr): c_null_ptr to skip, c_loc(array) otherwise.
    function bandup_gpu_fetch_kpt(handle, ikpt, weights, dN, n_selected, norms, selected, &
                                  selected_capacity) bind(c, name='bandup_gpu_fetch_kpt')
        import c_int, c_int64_t, c_ptr
        integer(c_int), intent(in), value :: handle, ikpt
        type(c_ptr), intent(in), value :: weights, dN, n_selected, norms, selected
        integer(c_int64_t), intent(in), value :: selected_capacity
        integer(c_int) :: bandup_gpu_fetch_kpt
    end function bandup_gpu_fetch_kpt

    function bandup_gpu_pipeline_depth(handle) bind(c, name='bandup_gpu_pipeline_depth')
        import c_int
        integer(c_int), intent(in), value :: handle
        integer(c_int) :: bandup_gpu_pipeline_depth
    end function bandup_gpu_pipeline_depth

    ! projected variant (-write_unf_dens_op): energy selection + calc_rho of perform_unfolding
    ! (band_unfolding_routines_mod.f90:1372-1428) for the resident SC-kpt ikpt
    function bandup_gpu_rho_count(handle, ikpt, min_dN, n_energies, n_entries) &
        bind(c, name='bandup_gpu_rho_count')
        import c_int, c_double, c_int64_t
        integer(c_int), intent(in), value :: handle, ikpt
        real(c_double), intent(in), value :: min_dN
        integer(c_int64_t), intent(out) :: n_energies, n_entries
        integer(c_int) :: bandup_gpu_rho_count
    end function bandup_gpu_rho_count

    function bandup_gpu_fetch_rho(handle, ikpt, capacity, fold, iener, m1, m2, rho_re, rho_im) &
        bind(c, name='bandup_gpu_fetch_rho')
        import c_int, c_int64_t, c_int32_t, c_float
        integer(c_int), intent(in), value :: handle, ikpt
        integer(c_int64_t), intent(in), value :: capacity
        integer(c_int32_t), intent(out) :: fold(*), iener(*), m1(*), m2(*)
        real(c_float), intent(out) :: rho_re(*), rho_im(*)
        integer(c_int) :: bandup_gpu_fetch_rho
    end function bandup_gpu_fetch_rho

    ! symmetry average over the needed directions (get_delta_Ns_for_output,
    ! band_unfolding_routines_mod.f90:826-853, :984-1036): out(r) = sum_d weights(d)*x(r,d)
    function bandup_gpu_symm_average(handle, n_dirs, n_values, weights, x, out) &
        bind(c, name='bandup_gpu_symm_average')
        import c_int, c_int64_t, c_double
        integer(c_int), intent(in), value :: handle, n_dirs
        integer(c_int64_t), intent(in), value :: n_values
        real(c_double), intent(in) :: weights(*), x(*)
        real(c_double), intent(out) :: out(*)
        integer(c_int) :: bandup_gpu_symm_average
    end function bandup_gpu_symm_average

    ! symmetry-averaged unfolding-density operators (:857-973); avg_dN(nener, n_pckpts)
    function bandup_gpu_symm_average_rho(handle, n_dirs, n_pckpts, nener, n_bands, weights, avg_dN, min_dN, &
                                         dir_offsets, pckpt, iener, m1, m2, rho_re, rho_im, capacity, n_out, &
                                         pckpt_out, iener_out, m1_out, m2_out, rho_re_out, rho_im_out) &
        bind(c, name='bandup_gpu_symm_average_rho')
        import c_int, c_int64_t, c_int32_t, c_double, c_float
        integer(c_int), intent(in), value :: handle, n_dirs, n_pckpts, nener, n_bands
        real(c_double), intent(in) :: weights(*), avg_dN(*)
        real(c_double), intent(in), value :: min_dN
        integer(c_int64_t), intent(in) :: dir_offsets(*)
        integer(c_int32_t), intent(in) :: pckpt(*), iener(*), m1(*), m2(*)
        real(c_float), intent(in) :: rho_re(*), rho_im(*)
        integer(c_int64_t), intent(in), value :: capacity
        integer(c_int64_t), intent(out) :: n_out
        integer(c_int32_t), intent(out) :: pckpt_out(*), iener_out(*), m1_out(*), m2_out(*)
        real(c_float), intent(out) :: rho_re_out(*), rho_im_out(*)
        integer(c_int) :: bandup_gpu_symm_average_rho
    end function bandup_gpu_symm_average_rho

    ! spinor wavefunctions: sigma(3, nener, n_fold) = Re Tr(rho . sigma_i)
    ! (get_SC_pauli_matrix_elmnts + band_unfolding_routines_mod.f90:1459-1481)
    function bandup_gpu_pauli_vectors(handle, ikpt, sigma) bind(c, name='bandup_gpu_pauli_vectors')
        import c_int, c_double
        integer(c_int), intent(in), value :: handle, ikpt
        real(c_double), intent(out) :: sigma(3,*)
        integer(c_int) :: bandup_gpu_pauli_vectors
    end function bandup_gpu_pauli_vectors

    ! calc_spectral_function (:654-698) for the resident SC-kpt; sf(nener, n_fold)
    function bandup_gpu_spectral_function(handle, ikpt, std_dev, sf) &
        bind(c, name='bandup_gpu_spectral_function')
        import c_int, c_double
        integer(c_int), intent(in), value :: handle, ikpt
        real(c_double), intent(in), value :: std_dev
        real(c_double), intent(out) :: sf(*)
        integer(c_int) :: bandup_gpu_spectral_function
    end function bandup_gpu_spectral_function

    ! args%perform_unfold: 0 == -dont_unfold
    function bandup_gpu_set_perform_unfold(handle, perform_unfold) &
        bind(c, name='bandup_gpu_set_perform_unfold')
        import c_int
        integer(c_int), intent(in), value :: handle, perform_unfold
        integer(c_int) :: bandup_gpu_set_perform_unfold
    end function bandup_gpu_set_perform_unfold

    ! projected-unfold without the text round trip (bandupy/orbital_contributions.py:304-362,
    ! unfolding_density_op.py:121-141); complex(c_double_complex) arrays in C (row-major) order
    function bandup_gpu_orb_contrib_matrix(handle, n_bands, n_ao, dual, proj, ao_mask, band_energies, &
                                           max_band_dE, min_abs, O_out) &
        bind(c, name='bandup_gpu_orb_contrib_matrix')
        import c_int, c_double, c_int8_t, c_ptr
        integer(c_int), intent(in), value :: handle, n_bands, n_ao
        real(c_double), intent(in) :: dual(*), proj(*), band_energies(*)
        integer(c_int8_t), intent(in) :: ao_mask(*)
        real(c_double), intent(in), value :: max_band_dE, min_abs
        type(c_ptr), intent(in), value :: O_out
        integer(c_int) :: bandup_gpu_orb_contrib_matrix
    end function bandup_gpu_orb_contrib_matrix

    function bandup_gpu_set_operator(handle, n_bands, O) bind(c, name='bandup_gpu_set_operator')
        import c_int, c_double
        integer(c_int), intent(in), value :: handle, n_bands
        real(c_double), intent(in) :: O(*)
        integer(c_int) :: bandup_gpu_set_operator
    end function bandup_gpu_set_operator

    function bandup_gpu_unfold_operator(handle, ikpt, min_abs_rho, clip, out) &
        bind(c, name='bandup_gpu_unfold_operator')
        import c_int, c_double
        integer(c_int), intent(in), value :: handle, ikpt
        real(c_double), intent(in), value :: min_abs_rho
        integer(c_int), intent(in), value :: clip
        real(c_double), intent(out) :: out(*)
        integer(c_int) :: bandup_gpu_unfold_operator
    end function bandup_gpu_unfold_operator

    ! multi-GPU (one BandUP process per GPU): rank 0 makes the id, MPI_Bcast's its 128 bytes,
    ! every rank calls comm_init; gather_dN is the path's only collective (ncclAllGather)
    function bandup_gpu_comm_unique_id(id) bind(c, name='bandup_gpu_comm_unique_id')
        import c_int, c_char
        character(kind=c_char), intent(out) :: id(128)
        integer(c_int) :: bandup_gpu_comm_unique_id
    end function bandup_gpu_comm_unique_id

    function bandup_gpu_comm_init(handle, n_ranks, rank, id) bind(c, name='bandup_gpu_comm_init')
        import c_int, c_char
        integer(c_int), intent(in), value :: handle, n_ranks, rank
        character(kind=c_char), intent(in) :: id(128)
        integer(c_int) :: bandup_gpu_comm_init
    end function bandup_gpu_comm_init

    function bandup_gpu_comm_finalize(handle) bind(c, name='bandup_gpu_comm_finalize')
        import c_int
        integer(c_int), intent(in), value :: handle
        integer(c_int) :: bandup_gpu_comm_finalize
    end function bandup_gpu_comm_finalize

    function bandup_gpu_gather_dN(handle, dN_local, row_ids_local, n_local, dN_all, row_ids_all, &
                                  capacity_rows, n_total) bind(c, name='bandup_gpu_gather_dN')
        import c_int, c_int64_t, c_double, c_ptr
        integer(c_int), intent(in), value :: handle
        real(c_double), intent(in) :: dN_local(*)
        integer(c_int64_t), intent(in) :: row_ids_local(*)
        integer(c_int64_t), intent(in), value :: n_local
        type(c_ptr), intent(in), value :: dN_all, row_ids_all
        integer(c_int64_t), intent(in), value :: capacity_rows
        integer(c_int64_t), intent(out) :: n_total
        integer(c_int) :: bandup_gpu_gather_dN
    end function bandup_gpu_gather_dN

    ! the same for rows of row_len doubles (sigma: 3*nener)
    function bandup_gpu_gather_rows(handle, row_len, rows_local, row_ids_local, n_local, rows_all, row_ids_all, &
                                    capacity_rows, n_total) bind(c, name='bandup_gpu_gather_rows')
        import c_int, c_int64_t, c_double, c_ptr
        integer(c_int), intent(in), value :: handle, row_len
        real(c_double), intent(in) :: rows_local(*)
        integer(c_int64_t), intent(in) :: row_ids_local(*)
        integer(c_int64_t), intent(in), value :: n_local
        type(c_ptr), intent(in), value :: rows_all, row_ids_all
        integer(c_int64_t), intent(in), value :: capacity_rows
        integer(c_int64_t), intent(out) :: n_total
        integer(c_int) :: bandup_gpu_gather_rows
    end function bandup_gpu_gather_rows

    !! device-resident rows (c_ptr = device addresses), enqueued on the caller's CUDA stream
    function bandup_gpu_allgather_device(handle, stream, d_send, d_recv, count) &
        bind(c, name='bandup_gpu_allgather_device')
        import c_int, c_int64_t, c_ptr
        integer(c_int), intent(in), value :: handle
        type(c_ptr), intent(in), value :: stream, d_send, d_recv
        integer(c_int64_t), intent(in), value :: count
        integer(c_int) :: bandup_gpu_allgather_device
    end function bandup_gpu_allgather_device

    function bandup_gpu_set_profiling(handle, enabled) bind(c, name='bandup_gpu_set_profiling')
        import c_int
        integer(c_int), intent(in), value :: handle, enabled
        integer(c_int) :: bandup_gpu_set_profiling
    end function bandup_gpu_set_profiling

    function bandup_gpu_kernel_time(handle, which, total_ms, launches) &
        bind(c, name='bandup_gpu_kernel_time')
        import c_int, c_double, c_int64_t
        integer(c_int), intent(in), value :: handle, which
        real(c_double), intent(out) :: total_ms
        integer(c_int64_t), intent(out) :: launches
        integer(c_int) :: bandup_gpu_kernel_time
    end function bandup_gpu_kernel_time

    function bandup_gpu_reset_kernel_times(handle) bind(c, name='bandup_gpu_reset_kernel_times')
        import c_int
        integer(c_int), intent(in), value :: handle
        integer(c_int) :: bandup_gpu_reset_kernel_times
    end function bandup_gpu_reset_kernel_times

    function bandup_gpu_launch_count(handle) bind(c, name='bandup_gpu_launch_count')
        import c_int, c_int64_t
        integer(c_int), intent(in), value :: handle
        integer(c_int64_t) :: bandup_gpu_launch_count
    end function bandup_gpu_launch_count

    function bandup_gpu_set_tuning(handle, chunk_vecs, ctas_per_sm, n_stages) &
        bind(c, name='bandup_gpu_set_tuning')
        import c_int
        integer(c_int), intent(in), value :: handle, chunk_vecs, ctas_per_sm, n_stages
        integer(c_int) :: bandup_gpu_set_tuning
    end function bandup_gpu_set_tuning

    function bandup_gpu_set_option(handle, option, value) &
        bind(c, name='bandup_gpu_set_option')
        import c_int
        integer(c_int), intent(in), value :: handle, option, value
        integer(c_int) :: bandup_gpu_set_option
    end function bandup_gpu_set_option

end interface

contains

    !! Last error text of a handle as a Fortran string (for `write(*,*) ...; stop`).
    function bandup_gpu_error_message(handle) result(msg)
        integer(c_int), intent(in) :: handle
        character(len=:), allocatable :: msg
        type(c_ptr) :: p
        character(kind=c_char), dimension(:), pointer :: chars
        integer :: n, i
        p = bandup_gpu_last_error(handle)
        if(.not. c_associated(p))then
            msg = ''
            return
        endif
        call c_f_pointer(p, chars, [1024])
        n = 0
        do while(n < 1024)
            if(chars(n+1) == c_null_char) exit
            n = n + 1
        enddo
        allocate(character(len=n) :: msg)
        do i=1,n
            msg(i:i) = chars(i)
        enddo
    end function bandup_gpu_error_message

    !! Convenience wrapper with the reference's own argument kinds: everything the body of
    !! `if(pckpt_folds)` computes (main_BandUP.f90:145-172) for ALL pc-kpts folding onto the
    !! current SC-kpt, synchronously.  pw_coeffs is wf%pw_coeffs (n_spinor, n_pw, n_bands),
    !! G_frac_flat is wf%G_frac viewed as integer(3, n_pw), folding_G(3, n_fold) holds
    !! Scoords(k) - K in Cartesian coordinates (band_unfolding_routines_mod.f90:571-576).
    !! On return n_selected(f) == 0 marks a pc-kpt that "cannot be parsed" (:161-172).
    subroutine bandup_gpu_unfold_kpt_f(handle, ikpt, pw_coeffs, G_frac_flat, band_energies, &
                                       folding_G, weights, dN, n_selected, ierr)
        integer(c_int), intent(in) :: handle, ikpt
        complex(c_float_complex), dimension(:,:,:), contiguous, target, intent(in) :: pw_coeffs
        integer(c_int32_t), dimension(:,:), contiguous, target, intent(in) :: G_frac_flat
        real(c_double), dimension(:), intent(in) :: band_energies
        real(c_double), dimension(:,:), intent(in) :: folding_G
        real(c_double), dimension(:,:), contiguous, target, intent(out) :: weights, dN
        integer(c_int32_t), dimension(:), contiguous, target, intent(out) :: n_selected
        integer(c_int), intent(out) :: ierr
        integer(c_int) :: n_spinor, n_pw, n_bands, n_fold
        integer(c_int32_t), dimension(:,:), allocatable :: g0
        real(c_double) :: residual

        n_spinor = size(pw_coeffs, dim=1)
        n_pw = size(pw_coeffs, dim=2)
        n_bands = size(pw_coeffs, dim=3)
        n_fold = size(folding_G, dim=2)
        allocate(g0(3, n_fold))
        ierr = bandup_gpu_folding_g0(handle, folding_G, n_fold, g0, residual)
        if(ierr /= BANDUP_GPU_OK) return
        ierr = bandup_gpu_submit_kpt(handle, ikpt, n_spinor, n_pw, n_bands, &
                                     BANDUP_GPU_LAYOUT_FORTRAN_INMEM, &
                                     int(n_pw, c_int64_t)*int(n_spinor, c_int64_t)*8_c_int64_t, &
                                     c_loc(pw_coeffs), c_loc(G_frac_flat), band_energies, &
                                     n_fold, g0)
        if(ierr /= BANDUP_GPU_OK) return
        ierr = bandup_gpu_fetch_kpt(handle, ikpt, c_loc(weights), c_loc(dN), c_loc(n_selected), &
                                    c_null_ptr, c_null_ptr, 0_c_int64_t)
    end subroutine bandup_gpu_unfold_kpt_f

end module bandup_gpu
