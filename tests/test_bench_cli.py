"""bench.py's reference arm (`--impl reference`: the CPU port of the reference's step, no GPU involved) runs here: the line it prints
must carry the contract keys the driver reads.  A bounded 256-sample batch keeps it to a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_a_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "256"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "mlp_train_step_samples_per_s" and line["unit"] == "samples/s" and line["dtype"] == "f32"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["sample_batch"] == 256 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--cpu-sample", "256"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


def test_numa_binding_is_a_noop_without_topology():
    """bench.bind_to_gpu_numa_node must never fail a run: no GPU / no sysfs topology -> None and the affinity is left alone"""
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    assert bench.bind_to_gpu_numa_node(0) is None or isinstance(bench.bind_to_gpu_numa_node(0), int)
    assert os.sched_getaffinity(0) <= before


def test_sweep_tables_regenerate_from_the_committed_bench_line(tmp_path):
    """profiles/r2_opbench_*.txt are generated from the `sweeps` key of a bench line (tools/sweeps.py --from-bench-json): the committed
    N=1 line must carry all four configs and the writer must reproduce the committed op table."""
    sys.path.insert(0, ROOT)
    from tools import sweeps
    line = [ln for ln in open(os.path.join(ROOT, "profiles", "r2p_bench_n1.json")).read().splitlines() if ln.startswith("{")][-1]
    sw = json.loads(line)["sweeps"]
    for key in ("gemm", "ops", "chained", "sine_net"):
        assert key in sw, key
    assert {r["n"] for r in sw["gemm"]["rows"]} == {512, 1024, 2048, 4096, 8192, 16384}
    assert {r["mode"] for r in sw["gemm"]["rows"]} == {"3xf16", "3xtf32", "tf32"} and {r["layout"] for r in sw["gemm"]["rows"]} == {"NN", "NT", "TN"}
    assert len(sw["ops"]["ops"]) >= 34 and "one_launch" in sw["sine_net"]
    sweeps.write_tables(sw, str(tmp_path / "x"))
    assert open(tmp_path / "x_opbench_ops_16384.txt").read() == open(os.path.join(ROOT, "profiles", "r2_opbench_ops_16384.txt")).read()
