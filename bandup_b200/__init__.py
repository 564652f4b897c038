"""bandup_b200 -- B200-native (sm_100a) unfolding hot path of BandUP behind a C ABI.

    csrc/            CUDA kernels (K0 plane-wave lists, K1 coset classification, K2 fused norm +
                     masked |C|^2 reduction, K3 finalisation + delta_N, K4-K6 projected/spinor
                     variant, K7 fold map, K8/K9 symmetry averaging) and the C ABI of
                     libbandup_gpu.so (include/bandup_gpu.h)
    host/            BandUP_gpu.x: the native C++ host mirroring the reference's unfold driver
    unfold.py        Python mirror of the reference's band_unfolding module (ctypes over the C ABI)
    wavecar.py       VASP WAVECAR container: raw band records handed to the GPU untouched
    bandup_files.py  BandUP's text inputs/outputs and a file-level `unfold` task
    sharding.py      SC-k-point sharding across GPUs (static blocks or a shared counter), gather of
                     the disjoint delta_N rows
    synth.py         synthetic inputs shaped like BASELINE.json's configurations (tests, bench)
    build.py         nvcc/g++ build of lib/libbandup_gpu.so and lib/BandUP_gpu.x

Nothing in this package computes on the CPU and nothing imports the oracle; importing
`bandup_b200.unfold` fails loudly when the CUDA library has not been built.
"""
__version__ = "0.1.0"
