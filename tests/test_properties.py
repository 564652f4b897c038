"""Property-based checks (hypothesis) of the host logic either side of the GPU path: file
round trips, the SC-kpt partition, the ticket counter, the ES10.3 formatter and the symmetry
averaging oracle.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from bandup_b200 import bandup_files as BF
from bandup_b200.sharding import KptTickets, partition_kpts

finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(min_value=0.0, max_value=1e9, allow_nan=False), min_size=0, max_size=60),
       st.integers(min_value=1, max_value=9))
def test_partition_tiles_the_list_in_order(costs, world):
    blocks = partition_kpts(costs, world)
    assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == len(costs)
    for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
        assert a <= b == c <= d
    if costs and sum(costs) > 0 and world > 1:
        # no block is more than one k-point heavier than the ideal share plus the heaviest k-point
        share = sum(costs) / world
        assert max(sum(costs[a:b]) for a, b in blocks) <= share + 2 * max(costs) + 1e-6


@settings(max_examples=40, deadline=None)
@given(st.integers(min_value=0, max_value=50), st.integers(min_value=1, max_value=6))
def test_round_robin_tickets_cover_every_kpoint_once(n, world):
    taken = []
    for r in range(world):
        t = KptTickets(n, offset=r, stride=world)
        while (i := t.take()) is not None:
            taken.append(i)
    assert sorted(taken) == list(range(n))


@settings(max_examples=200, deadline=None)
@given(st.one_of(finite, st.floats(min_value=-1e-30, max_value=1e-30), st.sampled_from([0.0, 1e-3, 9.9995e-4, 1.0, -1.0])))
def test_es10_3_is_fortran_es10_3(x):
    s = BF._es(x)
    assert len(s) == 10
    if "*" in s:                       # only a negative number with a three-digit exponent overflows
        assert x < 0 and (abs(x) < 1e-99 or abs(x) >= 1e100)
        return
    body = s.strip()
    if "E" not in body:                # d.ddd+eee / d.ddd-eee
        body = body[:5 + (body[0] == "-")] + "E" + body[5 + (body[0] == "-"):]
    assert abs(float(body) - x) <= 5.001e-4 * abs(x) + 1e-300
    mant = body.lstrip("-").split("E")[0]
    assert mant[1] == "." and len(mant) == 5 and (mant[0] != "0" or float(body) == 0.0)


@settings(max_examples=25, deadline=None)
@given(st.integers(min_value=1, max_value=4), st.integers(min_value=1, max_value=7), st.integers(min_value=0, max_value=2**31 - 1))
def test_band_structure_file_round_trip(n_dirs, nener, seed):
    import tempfile
    from pathlib import Path
    rng = np.random.default_rng(seed)
    dirs = []
    for _ in range(n_dirs):
        a, b = rng.normal(size=3), rng.normal(size=3)
        dirs.append(BF.kpts_line(a, b, int(rng.integers(1, 5))))
    kp = BF.PcKptsPath(dirs, float(rng.uniform(0, 2)))
    n = sum(len(d) for d in dirs)
    grid = np.sort(rng.uniform(-20, 5, nener))
    dN = rng.uniform(0, 3, (n, nener)) * (rng.random((n, nener)) < 0.6)
    with tempfile.TemporaryDirectory() as tmp:
        p = Path(tmp) / "ebs.dat"
        BF.write_band_struc(p, kp, grid, dN, e_fermi=1.25)
        tab = np.atleast_2d(BF.read_band_struc(p))
    assert tab.shape == (n * nener, 3)
    assert np.allclose(tab[:, 2].reshape(n, nener), dN, rtol=6e-4, atol=1e-30)
    assert np.allclose(tab[:, 1].reshape(n, nener), np.tile(grid - 1.25, (n, 1)), atol=5.1e-5)
    coords = tab[:, 0].reshape(n, nener)[:, 0]
    assert np.all(np.diff(coords) >= -1e-4) and abs(coords[0] - kp.zero_of_kpts_scale) < 5.1e-5


@settings(max_examples=30, deadline=None)
@given(st.integers(min_value=1, max_value=6), st.integers(min_value=0, max_value=2**31 - 1))
def test_symmetry_average_oracle_properties(n_dirs, seed):
    """One direction of weight one is the identity; equal operators in every direction average to
    themselves (weights summing to one, to single-precision rounding); the averaged list is ordered
    by (pc-kpt, energy) and holds no element twice."""
    from oracle import oracle_np as ON
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.ones(n_dirs))
    ents = []
    for k in range(2):
        for ie in range(1, 5):
            for m1 in range(1, 4):
                for m2 in range(1, 4):
                    if rng.random() < 0.5:
                        ents.append((k, ie, m1, m2, np.complex64(rng.normal() + 1j * rng.normal())))
    avg_dN = np.ones((2, 4))
    out = ON.symm_average_rho(w, [ents] * n_dirs, avg_dN)
    assert [e[:4] for e in out] == [e[:4] for e in ents]
    assert np.allclose([e[4] for e in out], [e[4] for e in ents], rtol=2e-6 * n_dirs, atol=1e-30)
    one = ON.symm_average_rho([1.0], [ents], avg_dN)
    assert all(a[4] == b[4] for a, b in zip(one, ents))
    avg_dN[1, 2] = 1e-4                                       # below the cut: that operator disappears
    cut = ON.symm_average_rho(w, [ents] * n_dirs, avg_dN)
    assert all((k, ie) != (1, 3) for k, ie, *_ in cut) and len(cut) == len([e for e in ents if (e[0], e[1]) != (1, 3)])
