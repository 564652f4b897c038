"""Known-answer tests of oracle/f90interp.py, the Fortran-subset interpreter that produces
tests/golden/fortran_golden.npz from the reference's sources.  Each test feeds it a few lines of
Fortran whose result under the language's rules is known in closed form -- kinds and mixed-kind
promotion, assignment conversion, literals, integer arithmetic, loops, allocatables, optional and
keyword arguments, generic resolution, derived types, source form."""
import math

import numpy as np
import pytest

from oracle import f90interp as F

PRELUDE = """
module kinds
implicit none
integer, parameter :: sp = selected_real_kind(6, 37), dp = selected_real_kind(15, 307)
real(kind=dp), parameter :: third_sp_literal = 0.333333333, third_dp_literal = 0.333333333_dp
end module kinds
"""


def interp(tmp_path, body, name="m.f90"):
    p = tmp_path / name
    p.write_text(PRELUDE + body)
    it = F.Interp()
    it.load(p)
    return it


def test_default_real_literals_are_single_precision(tmp_path):
    it = interp(tmp_path, "")
    assert it.module_value("third_sp_literal").value == np.float64(np.float32(0.333333333))
    assert it.module_value("third_dp_literal").value == 0.333333333
    assert it.module_value("dp").value == 8 and it.module_value("sp").value == 4


def test_mixed_kind_promotion_follows_fortran_not_numpy(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
function f(i, x) result(r)
  integer, intent(in) :: i
  real(kind=sp), intent(in) :: x
  real(kind=dp) :: r
  r = i*x + 0.1          ! integer*real(sp) and a default-real literal: all single precision
end function f
function g(i, x) result(r)
  integer, intent(in) :: i
  real(kind=sp), intent(in) :: x
  real(kind=dp) :: r
  r = i*x + 0.1_dp       ! the dp literal promotes the sum (the product is still formed in sp)
end function g
function h(z, s) result(r)
  complex(kind=sp), intent(in) :: z
  real(kind=dp), intent(in) :: s
  complex(kind=sp) :: r
  r = s*z                ! dp * complex(sp) is complex(dp), rounded back on assignment
end function h
end module t
""")
    x = np.float32(1.1)
    want_f = np.float64(np.float32(np.float32(3) * x) + np.float32(0.1))
    assert it.call("f", np.int32(3), x) == want_f
    want_g = np.float64(np.float32(np.float32(3) * x)) + 0.1
    assert it.call("g", np.int32(3), x) == want_g and want_f != want_g
    z = np.complex64(0.3 - 0.7j)
    s = np.float64(1.0 / 3.0)
    r = it.call("h", z, s)
    assert r.dtype == np.complex64
    assert r == np.complex64(complex(s * np.float64(z.real), s * np.float64(z.imag)))


def test_integer_division_power_mod_nint(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
subroutine s(out)
  real(kind=dp), dimension(1:8), intent(out) :: out
  integer :: i
  i = -7
  out(1) = i/2                 ! truncates toward zero
  out(2) = mod(i, 3)           ! sign of the dividend
  out(3) = mod(-0.25_dp, 1.0_dp)
  out(4) = nint(2.5_dp)        ! half away from zero
  out(5) = nint(-2.5_dp)
  out(6) = 2**10
  out(7) = 1.5_dp**2           ! x*x
  out(8) = 7/2*2.0_dp          ! (7/2) is integer division: 3*2.0
end subroutine s
end module t
""")
    out = F.Var(np.zeros(8), None)
    it.call("s", out)
    assert out.value.tolist() == [-3.0, -1.0, -0.25, 3.0, -3.0, 1024.0, 2.25, 6.0]


def test_complex_abs_is_single_precision_and_squared_in_double(tmp_path):
    """The expression of spectral_weight_for_coeff: abs() in the kind of its argument, **2.0_dp in dp."""
    it = interp(tmp_path, """
module t
use kinds
contains
function w(c) result(r)
  complex(kind=sp), intent(in) :: c
  real(kind=dp) :: r
  r = abs(c)**2.0_dp
end function w
end module t
""")
    c = np.complex64(0.123456789 + 0.987654321j)
    a = np.float32(math.hypot(float(c.real), float(c.imag)))      # correctly rounded fp32 |c|
    assert it.call("w", c) == np.float64(a) * np.float64(a)
    assert it.call("w", c) != abs(complex(c)) ** 2                # not the double-precision modulus


def test_dot_product_and_sum_accumulate_sequentially_in_their_own_kind(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
function d(a, b) result(r)
  complex(kind=sp), dimension(:), intent(in) :: a, b
  complex(kind=sp) :: r
  r = dot_product(a, b)
end function d
function s(a) result(r)
  real(kind=sp), dimension(:), intent(in) :: a
  real(kind=sp) :: r
  r = sum(a)
end function s
end module t
""")
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(200) + 1j * rng.standard_normal(200)).astype(np.complex64)
    b = (rng.standard_normal(200) + 1j * rng.standard_normal(200)).astype(np.complex64)
    re = im = np.float32(0)
    for x, y in zip(a, b):                                         # sum conjg(a)*b, every step rounded to fp32
        re = np.float32(re + np.float32(np.float32(x.real * y.real) + np.float32(x.imag * y.imag)))
        im = np.float32(im + np.float32(np.float32(x.real * y.imag) - np.float32(x.imag * y.real)))
    got = it.call("d", a, b)
    assert got.dtype == np.complex64 and got.real == re and got.imag == im
    v = rng.standard_normal(1000).astype(np.float32)
    acc = np.float32(0)
    for x in v:
        acc = np.float32(acc + x)
    assert it.call("s", v) == acc


def test_do_loops_exit_cycle_and_loop_variable_after_the_loop(tmp_path):
    it = interp(tmp_path, """
module t
contains
subroutine s(n, total, last, trips)
  integer, intent(in) :: n
  integer, intent(out) :: total, last, trips
  integer :: i
  total = 0
  trips = 0
  do i = 1, n
    if (mod(i, 2) == 0) cycle
    if (i > 7) exit
    total = total + i
  enddo
  last = i
  do i = 10, 1, -3
    trips = trips + 1
  enddo
  do while (trips < 100)
    trips = trips*2
  enddo
end subroutine s
end module t
""")
    total, last, trips = F.Var(np.int32(0), None), F.Var(np.int32(0), None), F.Var(np.int32(0), None)
    it.call("s", np.int32(20), total, last, trips)
    assert (int(total.value), int(last.value), int(trips.value)) == (1 + 3 + 5 + 7, 9, 128)
    it.call("s", np.int32(4), total, last, trips)
    assert (int(total.value), int(last.value)) == (1 + 3, 5)      # a completed loop leaves n+1


def test_allocatables_intent_out_and_generic_resolution(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
interface push
  module procedure push_int, push_real
end interface push
contains
subroutine push_int(item, list)
  integer, intent(in) :: item
  integer, dimension(:), allocatable, intent(inout) :: list
  integer, dimension(:), allocatable :: tmp
  if (.not. allocated(list)) then
    allocate(list(1:1))
    list(1) = item
  else
    allocate(tmp(1:size(list)+1))
    tmp(1:size(list)) = list
    tmp(size(list)+1) = item
    deallocate(list)
    allocate(list(1:size(tmp)))
    list = tmp
  endif
end subroutine push_int
subroutine push_real(item, list)
  real(kind=dp), intent(in) :: item
  real(kind=dp), dimension(:), allocatable, intent(inout) :: list
  if (.not. allocated(list)) allocate(list(1:1))
  list(1) = -item
end subroutine push_real
subroutine fresh(list, stat)
  integer, dimension(:), allocatable, intent(out) :: list   ! deallocated on entry
  integer, intent(out) :: stat
  deallocate(list, stat=stat)                                 ! so this fails quietly
end subroutine fresh
subroutine driver(ints, reals, stat)
  integer, dimension(:), allocatable, intent(inout) :: ints
  real(kind=dp), dimension(:), allocatable, intent(inout) :: reals
  integer, intent(out) :: stat
  integer :: i
  do i = 1, 4
    call push(item=i*i, list=ints)
  enddo
  call push(2.5_dp, reals)
  call fresh(ints, stat)
  call push(list=ints, item=42)
end subroutine driver
end module t
""")
    d_i = F.parse_declaration("integer, dimension(:), allocatable :: ints", it._eval_module_const)[0]
    d_r = F.parse_declaration("real(kind=8), dimension(:), allocatable :: reals", it._eval_module_const)[0]
    ints, reals, stat = F.Var(None, d_i), F.Var(None, d_r), F.Var(np.int32(0), None)
    it.call("driver", ints, reals, stat)
    assert ints.value.tolist() == [42] and reals.value.tolist() == [-2.5] and int(stat.value) != 0


def test_optional_and_keyword_arguments(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
function scale(x, factor, offset) result(r)
  real(kind=dp), intent(in) :: x
  real(kind=dp), intent(in), optional :: factor, offset
  real(kind=dp) :: r, f
  f = 2.0_dp
  if (present(factor)) f = factor
  r = f*x
  if (present(offset)) r = r + offset
end function scale
function outer(x, offset) result(r)
  real(kind=dp), intent(in) :: x
  real(kind=dp), intent(in), optional :: offset
  real(kind=dp) :: r
  r = scale(x, offset=offset)          ! an absent optional stays absent when passed on
end function outer
end module t
""")
    x = np.float64(3.0)
    assert it.call("scale", x) == 6.0
    assert it.call("scale", x, offset=np.float64(1.0)) == 7.0
    assert it.call("scale", offset=np.float64(1.0), factor=np.float64(-1.0), x=x) == -2.0
    assert it.call("outer", x) == 6.0 and it.call("outer", x, np.float64(0.5)) == 6.5


def test_derived_types_sections_and_array_intrinsics(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
type :: vec
  real(kind=dp), dimension(1:3) :: coord
end type vec
type :: box
  integer :: n
  type(vec), dimension(:), allocatable :: pts
  real(kind=dp), dimension(1:3,1:3) :: latt
end type box
contains
subroutine fill(b, n)
  type(box), intent(inout) :: b
  integer, intent(in) :: n
  integer :: i
  b%n = n
  allocate(b%pts(1:n))
  do i = 1, n
    b%pts(i)%coord(:) = i*b%latt(2,:) + (/ (0.5_dp*i, i=1,3) /)
  enddo
end subroutine fill
function cell(b) result(m)
  type(box), intent(in) :: b
  real(kind=dp), dimension(1:3,1:3) :: m
  m = transpose(matmul(b%latt, b%latt))
  if (any(m > 1000.0_dp) .or. all(m == 0.0_dp)) m = -1.0_dp
end function cell
end module t
""")
    b = it.new_struct("box")
    b["latt"][:] = np.arange(9.0).reshape(3, 3)
    it.call("fill", b, np.int32(2))
    assert int(b["n"]) == 2 and len(b["pts"]) == 2
    assert b["pts"][1]["coord"].tolist() == (2 * b["latt"][1] + np.array([0.5, 1.0, 1.5])).tolist()
    assert np.array_equal(it.call("cell", b), (b["latt"] @ b["latt"]).T)


def test_source_form(tmp_path):
    """Continuation lines with comments between them, `!$omp` lines, strings holding `!`, keywords in
    any case, `;`, preprocessor branches."""
    it = interp(tmp_path, """
MODULE T
USE kinds
CONTAINS
FUNCTION F(X) RESULT(R)
  REAL(KIND=dp), INTENT(IN) :: X
  REAL(KIND=dp) :: R
  CHARACTER(LEN=20) :: msg
  msg = 'not a comment ! here'
  R = X + &     ! trailing comment
      ! a comment line inside a continued statement
      & 1.0_DP
  !$omp parallel do &
  !$omp shared(r)
  R = R*2.0_dp ; R = R + 1.0_dp
#if defined (__NEVER__)
  R = -1.0_dp
#else
  IF (R .GT. 0.0_dp .AND. msg(1:3) == 'not') R = R + 100.0_dp
#endif
END FUNCTION F
END MODULE T
""")
    assert it.call("f", np.float64(1.0)) == (1.0 + 1.0) * 2 + 1 + 100


def test_statements_of_a_line_range_run_in_a_given_environment(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
subroutine s(a)
  real(kind=dp), dimension(:), intent(inout) :: a
  integer :: i
  do i = 1, size(a)
    a(i) = a(i)*a(i)
  enddo
  a = a + 1.0_dp
end subroutine s
end module t
""")
    src = (tmp_path / "m.f90").read_text().splitlines()
    first = next(i for i, l in enumerate(src, 1) if "do i = 1" in l)
    last = next(i for i, l in enumerate(src, 1) if "enddo" in l)
    a = np.array([1.0, 2.0, 3.0])
    it.exec_statements(tmp_path / "m.f90", first, last, {"a": (a, None), "i": (None, "integer :: i")})
    assert a.tolist() == [1.0, 4.0, 9.0]


def test_unsupported_constructs_are_refused_not_guessed(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
contains
subroutine s(a)
  real(kind=dp), dimension(0:2) :: a
  a(0) = 1.0_dp
end subroutine s
subroutine u(x)
  real(kind=dp), intent(out) :: x
  x = undefined_thing + 1.0_dp
end subroutine u
end module t
""")
    with pytest.raises(NotImplementedError):
        it.call("s", np.zeros(3))
    with pytest.raises(F.FortranError):
        it.call("u", F.Var(np.float64(0), None))


def test_unformatted_direct_and_stream_reads(tmp_path):
    """The two access modes read_wavecar uses: direct records (rec=, recl in bytes) and a stream
    position (pos=), with implied-do item lists, array sections and 8-byte integers."""
    rec = 64
    data = np.zeros(3 * rec, dtype=np.uint8)
    data[0:24] = np.array([float(rec), 2.0, 45200.0]).view(np.uint8)
    data[rec:rec + 8 + 32] = np.concatenate([np.array([7.0]), np.array([1 + 2j, 3 - 4j])
                                             .view(np.float64)]).view(np.uint8)
    data[2 * rec:2 * rec + 32] = np.array([1 + 1j, 2 + 2j, 3 + 3j, 4 + 4j], dtype=np.complex64).view(np.uint8)
    path = tmp_path / "records.bin"
    path.write_bytes(data.tobytes())
    it = interp(tmp_path, """
module t
use kinds
contains
subroutine rd(file, head, z, c, ios_past_end)
  character(len=*), intent(in) :: file
  real(kind=dp), dimension(1:4), intent(out) :: head
  complex(kind=dp), dimension(1:2), intent(out) :: z
  complex(kind=sp), dimension(1:2,1:2), intent(out) :: c
  integer, intent(out) :: ios_past_end
  integer(kind=selected_int_kind(18)) :: nrecl, pos
  real(kind=dp) :: x
  integer :: u, ios, j
  u = 11
  nrecl = 24
  open(unit=u, file=file, access='direct', recl=nrecl, iostat=ios, status='old')
  read(unit=u, rec=1, iostat=ios) x
  close(unit=u)
  nrecl = nint(x)
  open(unit=u, file=file, access='direct', recl=nrecl, iostat=ios, status='old')
  read(unit=u, rec=1) (head(j), j=1,3)
  read(unit=u, rec=2) head(4), (z(j), j=1,2)
  read(unit=u, rec=9, iostat=ios_past_end) x
  close(u)
  open(unit=u, file=file, access='stream', iostat=ios, status='old')
  pos = 1 + 2*nrecl
  read(unit=u, pos=pos) (c(j,:), j=1,2)
  close(u)
end subroutine rd
end module t
""")
    head, z = F.Var(np.zeros(4), None), F.Var(np.zeros(2, dtype=np.complex128), None)
    c, ios = F.Var(np.zeros((2, 2), dtype=np.complex64), None), F.Var(np.int32(0), None)
    it.call("rd", str(path), head, z, c, ios)
    assert head.value.tolist() == [64.0, 2.0, 45200.0, 7.0]
    assert z.value.tolist() == [1 + 2j, 3 - 4j]
    assert c.value.tolist() == [[1 + 1j, 2 + 2j], [3 + 3j, 4 + 4j]]      # c(1,:) then c(2,:)
    assert int(ios.value) != 0 and not it.units


def test_local_types_components_of_array_sections_and_intent_out_structures(tmp_path):
    it = interp(tmp_path, """
module t
use kinds
real(kind=dp), dimension(1:2,1:2), parameter :: eye = reshape(real((/ 1, 0, 0, 1 /), kind=dp), shape(eye))
type :: bag
  integer :: n
  integer, dimension(:), allocatable :: items
end type bag
contains
subroutine fill(b, n)
  type(bag), intent(out) :: b          ! allocatable components are deallocated on entry
  integer, intent(in) :: n
  allocate(b%items(1:n))               ! ... so this may be called twice on the same object
  b%items = n
  b%n = n
end subroutine fill
function flags(n) result(c)
  integer, intent(in) :: n
  integer :: c
  type :: entry
    logical :: seen
    integer :: weight
  end type entry
  type(entry), dimension(:), allocatable :: list
  integer :: i
  allocate(list(1:n))
  do i = 1, n
    list(i)%seen = mod(i, 3) == 0
    list(i)%weight = i
  enddo
  c = 100*count(.not. list(:)%seen) + sum(list(:)%weight)
end function flags
end module t
""")
    assert np.array_equal(it.module_value("eye").value, np.eye(2))
    b = it.new_struct("bag")
    it.call("fill", b, np.int32(3))
    it.call("fill", b, np.int32(2))
    assert b["items"].tolist() == [2, 2] and int(b["n"]) == 2
    assert int(it.call("flags", np.int32(7))) == 100 * 5 + 28
