"""CPU tests of bench.py's reference arm (the driver runs it on the GPU box beside the product's arm): one JSON line
on stdout with the contract's keys, never the product library, rank 0 only under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-reads", "40000", *args], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_contract_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads/sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert abs(d["value"] - 40000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["config"]["workload"].startswith("C3") and d["config"]["reads_per_step"] == 40000
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert "NOT parasail" in cb["kind_note"] and cb["sample"] and cb["gcups"] > 0
    # the scalar leg of the same port beside it: same work, one alignment at a time
    assert cb["scalar"]["value"] > 0 and cb["scalar"]["gcups"] > 0


def test_reference_arm_other_configs_and_ranks():
    d = json.loads(_run(None, "--config", "C5"))
    assert d["config"]["workload"].startswith("C5") and d["config"]["adapter_len"] == 40
    # under torchrun only rank 0 runs and prints; the other ranks exit 0 without work
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2").strip() == ""
    d = json.loads(_run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2"))
    assert d["n_gpus"] == 2 and d["impl"] == "reference"


def test_oracle_file_leg(tmp_path):
    """The ingest legs' CPU baseline: the port on a gzipped FASTQ file, phases timed, same table as the plain file call."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import oracle
    cfg = oracle.synth_cfg(seed=5, read_len=150, adapter_len=20, region_len=99, n_variants=500)
    ad = oracle.synth_adapters(cfg)
    leg = bench.oracle_file_leg(oracle, cfg, ad, 0.75, str(tmp_path), 30000, 2)
    assert leg["value"] > 0 and leg["value_if_overlapped"] >= leg["value"] and leg["cores"] == 2 and leg["kind"] == "port"
    assert abs(leg["seconds_inflate_one_zlib_stream"] + leg["seconds_fastq_framing"] + leg["seconds_closures_all_threads_avx2"]
               - leg["seconds"]) < 0.25 * leg["seconds"] + 0.05
    assert os.listdir(tmp_path) == []
    path = str(tmp_path / "same.fq.gz")
    oracle.write_fastq(cfg, 0, 30000, path, bgzf=True, level=1)
    assert leg["table_rows"] == len(oracle.find_variants_file(path, ad, n_threads=2))
