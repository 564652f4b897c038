"""GPU parity of the symmetry averaging over the needed directions (K8/K9), the post-loop step
of the unfolding path (get_delta_Ns_for_output, band_unfolding_routines_mod.f90:789-1043),
through the C ABI against the C oracle: bit-exact values and the reference's element order."""
import numpy as np
import pytest

from bandup_b200 import synth
from bandup_b200.unfold import MIN_DN_SYMM_AVG, BandupGpuError, PwWavefunction, UnfoldDensityOps
from tests._symm_cases import make_case

pytestmark = pytest.mark.gpu


def _ops_per_dir(case):
    k, e, m1, m2, rho = case["cols"]
    off = case["off"]
    return [UnfoldDensityOps(0, k[a:b], e[a:b], m1[a:b], m2[a:b], rho[a:b]) for a, b in zip(off[:-1], off[1:])]


def _same(res, want):
    assert len(res.rho) == len(want[0])
    for got, w in zip((res.fold, res.iener, res.m1, res.m2), want[:4]):
        assert np.array_equal(got, w)
    assert np.array_equal(res.rho.view(np.uint32), want[4].view(np.uint32))      # bit-equal complex64


@pytest.mark.parametrize("shape", [(3, 12), (5, 501, 3), (200, 501)])
@pytest.mark.parametrize("n_dirs", [1, 4, 12])
def test_dense_average_is_bit_exact(unfolder, shape, n_dirs):
    from oracle import oracle as O
    rng = np.random.default_rng(n_dirs * 100 + len(shape))
    w = rng.dirichlet(np.ones(n_dirs))
    x = rng.random((n_dirs,) + shape) * np.exp(rng.normal(size=(n_dirs,) + shape) * 4)
    x[rng.random(x.shape) < 0.3] = 0.0
    got = unfolder.symm_average(w, x)
    assert got.shape == shape
    assert np.array_equal(got, O.symm_average(w, x))
    if n_dirs == 1:
        assert np.array_equal(got, w[0] * x[0])


@pytest.mark.parametrize("seed,n_dirs,equal", [(7, 4, False), (8, 1, False), (9, 6, True), (10, 12, False)])
def test_rho_average_matches_oracle_bitwise(unfolder, seed, n_dirs, equal):
    from oracle import oracle as O
    case = make_case(seed=seed, n_dirs=n_dirs, equal_weights=equal)
    avg = unfolder.symm_average(case["weights"], case["dN"])
    res = unfolder.symm_average_rho(case["weights"], _ops_per_dir(case), avg, case["nb"])
    want = O.symm_average_rho(case["weights"], case["off"], *case["cols"], avg, case["nb"])
    assert len(want[0]) > 0
    _same(res, want)
    assert res.n_energies == len({(int(k), int(e)) for k, e in zip(want[0], want[1])})
    assert all(avg[k, e - 1] >= MIN_DN_SYMM_AVG for k, e in zip(res.fold, res.iener))


def test_rho_average_large_sparse(unfolder):
    """C4-like sizes: 100 pc-kpts x 501 energies x 500 bands, 8 directions, ~2 M stored elements."""
    from oracle import oracle as O
    rng = np.random.default_rng(44)
    n_dirs, n_pck, nener, nb = 8, 100, 501, 500
    w = rng.dirichlet(np.ones(n_dirs))
    space = n_pck * nener * nb * nb
    # same-bin band pairs: a pool of keys the directions mostly share, so the merge really merges
    pool = np.unique(rng.integers(0, space, size=400_000))
    cols, off = [[], [], [], [], []], [0]
    for d in range(n_dirs):
        key = np.sort(pool[rng.random(pool.size) < 0.6])
        g, pair = key // (nb * nb), key % (nb * nb)
        for c, v in zip(cols, (g // nener, g % nener + 1, pair // nb + 1, pair % nb + 1,
                               (rng.normal(size=key.size) + 1j * rng.normal(size=key.size)))):
            c.append(v)
        off.append(off[-1] + key.size)
    arrays = [np.concatenate(c).astype(np.int32) for c in cols[:4]] + [np.concatenate(cols[4]).astype(np.complex64)]
    avg = rng.random((n_pck, nener))
    avg[rng.random(avg.shape) < 0.2] = 5e-4
    ops = [UnfoldDensityOps(0, *[a[lo:hi] for a in arrays]) for lo, hi in zip(off[:-1], off[1:])]
    res = unfolder.symm_average_rho(w, ops, avg, nb)
    want = O.symm_average_rho(w, np.array(off), *arrays, avg, nb)
    assert len(want[0]) > 200_000
    _same(res, want)


def test_rho_average_refuses_unsorted_and_out_of_range(unfolder):
    case = make_case(seed=3, n_dirs=2)
    avg = unfolder.symm_average(case["weights"], case["dN"])
    ops = _ops_per_dir(case)
    bad = UnfoldDensityOps(0, ops[1].fold[::-1].copy(), ops[1].iener[::-1].copy(), ops[1].m1[::-1].copy(),
                           ops[1].m2[::-1].copy(), ops[1].rho[::-1].copy())
    with pytest.raises(BandupGpuError) as ei:
        unfolder.symm_average_rho(case["weights"], [ops[0], bad], avg, case["nb"])
    assert ei.value.code == 1 and "order" in str(ei.value)
    with pytest.raises(BandupGpuError) as ei:
        unfolder.symm_average_rho(case["weights"], ops, avg, case["nb"] - 1)      # m > n_bands
    assert ei.value.code == 1 and "out of range" in str(ei.value)
    empty = UnfoldDensityOps(0, *[np.zeros(0, dtype=np.int32)] * 4, np.zeros(0, dtype=np.complex64))
    res = unfolder.symm_average_rho([0.5, 0.5], [empty, empty], avg, case["nb"])
    assert len(res.rho) == 0


def test_average_of_unfolded_directions(unfolder):
    """End to end on real operators: the pc-kpts one SC-kpt unfolds to stand in for the needed
    directions of ONE pc-kpt (what symmetry-equivalent directions are to BandUP); delta_N and rho
    come from the device path, the average is compared with the oracle's on the same inputs, and
    with equal operators the average reproduces them."""
    from oracle import oracle as O
    sc = synth.make_scenario("si333_vacancy", scale=0.25)
    unfolder.verify_commens(sc.b_pc, sc.B_sc)
    unfolder.set_energy_grid(*sc.energy_window())
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    ener[2] = ener[1]
    ener = np.sort(ener)
    fG = sc.folding_G(0)
    fG = np.concatenate([fG, fG[:1]])                          # the last "direction" repeats the first
    r = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG, ikpt=2)
    rho = unfolder.calc_rho(2)
    n_dirs = fG.shape[0]
    w = np.full(n_dirs, 1.0 / n_dirs)
    dN = r.dN[:, None, :]                                      # (n_dirs, 1 pc-kpt, nener)
    avg = unfolder.symm_average(w, dN)
    assert np.array_equal(avg, O.symm_average(w, dN))
    ops, off = [], [0]
    for d in range(n_dirs):
        s = rho.fold == d
        ops.append(UnfoldDensityOps(0, np.zeros(int(s.sum()), dtype=np.int32), rho.iener[s], rho.m1[s],
                                    rho.m2[s], rho.rho[s]))
        off.append(off[-1] + int(s.sum()))
    res = unfolder.symm_average_rho(w, ops, avg, coeffs.shape[0])
    cat = [np.concatenate([getattr(o, a) for o in ops]) for a in ("fold", "iener", "m1", "m2", "rho")]
    _same(res, O.symm_average_rho(w, np.array(off), *cat, avg, coeffs.shape[0]))
    assert len(res.rho) > 0
    # two identical directions with weight 1/2 each give the operator back (x/2 + x/2 == x in sp/dp)
    two = [ops[0], ops[-1]]
    avg2 = unfolder.symm_average([0.5, 0.5], dN[[0, -1]])
    assert np.array_equal(avg2[0], r.dN[0])
    res2 = unfolder.symm_average_rho([0.5, 0.5], two, avg2, coeffs.shape[0])
    keep = avg2[0][ops[0].iener - 1] >= MIN_DN_SYMM_AVG
    assert np.array_equal(res2.rho, ops[0].rho[keep]) and np.array_equal(res2.m2, ops[0].m2[keep])
