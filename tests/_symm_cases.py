"""Seeded inputs for the symmetry-averaging tests (shared by the CPU oracle tests and the GPU
parity tests): n_dirs needed directions with weights, dense per-direction columns, and sparse
per-direction unfolding-density operators in the reference's append order."""
import numpy as np


def make_case(seed=7, n_dirs=4, n_pck=3, nener=12, nb=9, p_op=0.6, p_el=0.2, equal_weights=False):
    rng = np.random.default_rng(seed)
    w = np.full(n_dirs, 1.0 / n_dirs) if equal_weights else rng.dirichlet(np.ones(n_dirs))
    dN = rng.random((n_dirs, n_pck, nener)) * 2.0
    dN[rng.random(dN.shape) < 0.3] = 0.0                       # most of a delta_N column is empty
    per_dir, off = [], [0]
    cols = [[], [], [], [], []]
    for d in range(n_dirs):
        ents = []
        for k in range(n_pck):
            for ie in range(1, nener + 1):
                if rng.random() > p_op:
                    continue
                for m1 in range(1, nb + 1):
                    for m2 in range(1, nb + 1):
                        if rng.random() > p_el:
                            continue
                        ents.append((k, ie, m1, m2, np.complex64(rng.normal() + 1j * rng.normal())))
        per_dir.append(ents)
        off.append(off[-1] + len(ents))
        for e in ents:
            for c, v in zip(cols, e):
                c.append(v)
    arrays = [np.array(c, dtype=np.int32) for c in cols[:4]] + [np.array(cols[4], dtype=np.complex64)]
    return dict(weights=w, dN=dN, per_dir=per_dir, off=np.array(off, dtype=np.int64), cols=arrays, nb=nb,
                n_pck=n_pck, nener=nener)
