"""bench.py's reference arm runs on host cores, so its side of the JSON contract can be checked without a GPU:
one line on stdout, the keys the driver reads, `impl: reference`, a cpu_baseline describing the run and an e2e
object that repeats the line's own value.  (The GPU arm needs a B200; its line is recorded under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    # the arm picks the workload's own grid (384^3) when the host can hold it; the contract is checked on the small sample
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT, env=dict(os.environ, LUMA_BENCH_CPU_CASE="c2_128"))
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "MLUPS" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.decode().strip() == ""


def test_cpu_baseline_case_table_and_host_probes():
    """the reference arm runs the workload's own per-GPU grid first and only falls back to stated smaller samples; every case it
    names exists in the oracle's table; the host probes return sane values"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    # bench.py redirects fd 1 at import; keep this process's stdout
    saved = os.dup(1)
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    from oracle.cases import BENCH_CASES
    for workload, cases in bench.CPU_CASES.items():
        assert cases, workload
        for name, need_gb, what in cases:
            assert name in BENCH_CASES and need_gb > 0 and what
    assert bench.CPU_CASES["c5"][0][0] == "c5_384" and bench.CPU_CASES["c5"][0][2].startswith("the workload")
    assert bench.CPU_CASES["c2"][0][0] == "c2_256" and bench.CPU_CASES["c2"][0][2].startswith("the workload")
    c = BENCH_CASES["c5_384"]
    assert (c.N, c.M, c.K, c.Q) == (384, 384, 384, 19)
    threads, desc = bench.physical_cores()
    assert 1 <= threads <= (os.cpu_count() or 1) and "physical cores" in desc
    assert bench.host_mem_available_gb() > 0.5
    # both arms name the workload with the same string (the driver compares them)
    assert bench.workload_name("c5", 8) == "BASELINE configs[4]: weak-scaling sweep, 3D lid-driven cavity D3Q19 BGK Re=1000 (the case of configs[1]), 3072x384x384 cells (384^3 per GPU, x-slabs)"
    assert bench.DEFAULT_WORKLOAD == "c5"
