"""CPU checks of host-side logic that has no GPU dependency: numpy.percentile's interpolation rule as restated for the
slide-level fit, the histogram rank location, and the JSON contract of bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(-1e3, 1e3, allow_nan=False), min_size=1, max_size=200), st.floats(0.0, 100.0))
def test_percentile_index_and_lerp_match_numpy(vals, pct):
    from stainlib_b200.normalization.slide_fit import lerp_np, percentile_index
    s = np.sort(np.asarray(vals, dtype=np.float64))
    lo, hi, fr = percentile_index(len(s), pct)
    got = lerp_np(s[lo], s[hi], fr)
    want = np.percentile(s, pct)
    assert got == pytest.approx(want, rel=1e-12, abs=1e-12)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 50), min_size=1, max_size=64).filter(lambda h: sum(h) > 0), st.data())
def test_locate_rank_in_histogram(hist, data):
    from stainlib_b200.normalization.slide_fit import _locate
    hist = np.asarray(hist, dtype=np.int64)
    rank = data.draw(st.integers(0, int(hist.sum()) - 1))
    b, rem = _locate(hist, rank)
    expanded = np.repeat(np.arange(len(hist)), hist)           # the sorted multiset the histogram stands for
    assert expanded[rank] == b
    assert rem == rank - int(hist[:b].sum()) and 0 <= rem < hist[b]


def test_key_inverses_are_monotone():
    from stainlib_b200.normalization.slide_fit import angle_from_key, conc_from_key
    keys = np.linspace(1, (1 << 23) - 1, 4001).astype(np.int64)          # key 0 is the branch cut itself (angle pi)
    ang = np.array([angle_from_key(int(k)) for k in keys])
    con = np.array([conc_from_key(int(k)) for k in keys])
    assert np.all(np.diff(ang) > 0) and ang[0] >= -np.pi - 1e-12 and ang[-1] <= np.pi + 1e-12
    assert np.all(np.diff(con) > 0) and conc_from_key(0) == 0.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (CPU only: runs here)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "macenko256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "Mpx/s" and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
    # the reference's own files when a copy is reachable (/root/reference here, baseline/_ref on the GPU box), else the port
    have_ref = os.path.isdir("/root/reference/stainlib") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "stainlib"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    # one BLAS/OpenMP thread per worker process (oversubscription used to cost the baseline 5-8x)
    assert d["cpu_baseline"]["threads_per_worker"] == 1 and d["cpu_baseline"]["cores"] == os.cpu_count()


def test_traffic_entries_are_tied_to_the_sources(tmp_path, monkeypatch):
    """bench.py reports an ncu DRAM-traffic figure only while the library's sources are the ones the capture was taken from
    (profiles/traffic.json carries a hash of csrc/ + the header + the nvcc flags; rebuilding the same sources keeps it)."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from stainlib_b200.build import source_sha16
    sha = source_sha16()
    assert isinstance(sha, str) and len(sha) == 16 and sha == source_sha16() == bench.src_sha16()
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "traffic.json").write_text(json.dumps({"w": {"k_now": {"dram_bytes": 123, "src_sha16": sha},
                                                          "k_old": {"dram_bytes": 456, "src_sha16": "0" * 16},
                                                          "k_legacy": 789}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.ncu_traffic("w", "k_now") == 123
    assert bench.ncu_traffic("w", "k_old") is None
    assert bench.ncu_traffic("w", "k_legacy") is None
    assert bench.ncu_traffic("w", "absent") is None and bench.ncu_traffic("other", "k_now") is None
