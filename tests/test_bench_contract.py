"""bench.py's one-JSON-line contract (what the driver parses)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout                  # exactly ONE line on stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run(["--impl", "reference", "--workload", "graphene4x4", "--steps", "2", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "coeff_GB_per_s" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    f = d["e2e_file"]
    assert f["value"] > 0 and f["unit"] == "GB/s" and f["sc_kpts"] >= 1 and f["coeff_bytes"] < f["file_bytes"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_b200_arm_json_line():
    d = _run(["--workload", "si333_vacancy", "--steps", "3", "--warmup", "3", "--kpts-per-step", "2",
              "--cpu-seconds", "1"])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d)
    # the timed region runs for at least 0.5 s: more steps than asked for, and `steps` says how many
    assert d["steps"] >= 3 and d["steps_requested"] == 3 and d["timed_region_ms"] >= 450.0
    assert d["n_gpus"] == 1 and d["gpu_launches"] == 3 * d["steps"] and d["value"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert set(rf["kernel_ms_per_step"]) == {"k1_coset_classify", "k2_weights_reduce", "k3_finalize_delta_n"}
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 2 * 500 * 19000 * 8 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                    # host<->device copies are inside the e2e region
    assert e["e2e_checked"] is True and len(e["h2d_GBps_per_rank"]) == 1 and e["gpu_of_rank"] == [0]
    pb = e["pinned_block_probe"]
    assert len(pb["candidates_GBps"]) in (2, 3, 4) and len(pb["kept"]) == 2     # spares are best effort
    f = d["e2e_file"]
    assert f["value"] > 0 and f["checked_vs_oracle_max_rel_err"] <= 1e-6 and f["page_cache"] == "warm"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    # the reference arm names the same configuration
    r = _run(["--impl", "reference", "--workload", "si333_vacancy", "--steps", "1", "--warmup", "0", "--kpts-per-step", "2",
              "--no-e2e-file"])
    assert r["config"] == d["config"] and r["metric"] == d["metric"] and r["unit"] == d["unit"]


def test_both_arms_print_the_same_config_object():
    """--impl reference names the B200 arm's configuration (it times a bounded sample of it): the
    `config` objects come from one function, so the driver's same-config check sees equal dicts."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from bandup_b200 import synth
    from oracle import oracle as O
    args = types.SimpleNamespace(workload="graphene4x4", complex128=False, kpts_per_step=0, batch_gb=40.0,
                                 per_kpt_launch=False, k2_variant=0, dependent_launch=3, kpt_offset=0)
    sc = synth.make_scenario("graphene4x4")
    kps = bench.batch_kpts(args, sc)
    assert kps == sc.n_sc_kpts                            # the whole job of a small cell is one batch
    my_k = list(range(kps))
    cfg = bench.workload_config(args, sc, 1, kps, [sc.gvecs(i).shape[0] for i in my_k],
                                [len(sc.folding_G(i)) for i in my_k], len(O.real_seq(*sc.energy_window())))
    d = _run(["--impl", "reference", "--workload", "graphene4x4", "--steps", "1", "--warmup", "0", "--no-e2e-file"])
    assert d["config"] == cfg and "model" not in cfg and cfg["kpts_per_step_per_gpu"] == kps
    args.workload = "sige555"
    assert bench.batch_kpts(args, synth.make_scenario("sige555")) == 16
