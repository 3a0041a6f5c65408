"""bench.py's reference arm (CPU): the JSON line the driver parses, and the torchrun convention that only rank 0 works."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=REPO, timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-seconds/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("audio-seconds/sec encoded") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "utterances" in cb["sample"]
    assert d["config"]["model"] == "microsoft/wavlm-large" and "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_both_arms_print_the_same_config_and_every_baseline_workload_is_listed():
    """The reference arm has to run on the B200 arm's `config` (same object for both), and the default line has to carry
    every BASELINE config as a workload."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in bench.WORKLOADS:
        a, b = bench.workload_config(name, 1), bench.workload_config(name, 1)
        assert a == b and a["model"] == bench.WORKLOADS[name]["model"] and "workload" in a
    assert set(bench.DEFAULT_EXTRAS) == {"whisper-large-v3", "hubert-xlarge", "xls-r-2b", "wavlm-large-c1", "wavlm-large-sweep"}
    assert bench.WORKLOADS["wavlm-large-corpus"].get("strong") and bench.workload_config("wavlm-large-corpus", 8)["global_batch"] == 4096
    for name in bench.DEFAULT_EXTRAS:
        assert bench.WORKLOADS[name]["model"] in bench.GOLDEN      # every benched model has a committed HF golden for the in-bench check
    for f in bench.GOLDEN.values():
        assert os.path.isfile(os.path.join(REPO, "tests", "golden", f)), f


def test_gpus_n_without_torchrun_relaunches_one_rank_per_gpu(monkeypatch):
    """`python bench.py --gpus 4` typed by hand execs the driver's torchrun launch (127.0.0.1 rendezvous) with the same flags;
    under torchrun (WORLD_SIZE set) it does not."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    seen = {}

    class Exec(Exception):
        pass

    def fake_exec(prog, argv):
        seen["argv"] = list(argv)
        raise Exec()

    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(bench.os, "execvp", fake_exec)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "4", "--steps", "7", "--warmup", "3"])
    try:
        bench.main()
        assert False, "expected the relaunch"
    except Exec:
        pass
    a = seen["argv"]
    assert a[1:3] == ["-m", "torch.distributed.run"] and "--nproc-per-node=4" in a and a[a.index("--master-addr") + 1] == "127.0.0.1"
    assert a[-6:] == ["--gpus", "4", "--steps", "7", "--warmup", "3"] and a[-7].endswith("bench.py")
    # under torchrun the same flags run in place
    called = {}
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setattr(bench, "run_ours", lambda args: called.setdefault("gpus", args.gpus))
    bench.main()
    assert called["gpus"] == 4
