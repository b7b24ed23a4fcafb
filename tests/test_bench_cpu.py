"""CPU checks of bench.py's host logic: the reference arm's JSON contract (it runs the oracle port on the host cores, the one
place besides tests/ and smoke() that may execute oracle/), and the per-kernel algorithmic work model behind `roofline`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                          "1", "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    # "reference": the unmodified reference modules (checkout / baseline/_ref install); "port": the oracle restatement
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("supervised", 2, 1)          # the same object the GPU arm prints


def test_workload_splits():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.batch_split("supervised", 24) == [12, 12, 0] and bench.batch_split("mean_teacher", 48) == [12, 12, 24]
    assert bench.batch_split("dcase2024", 60) == [12, 6, 6, 12, 24]          # pretrained.yaml:8
    s = bench.batch_split("dcase2024", 24)
    assert sum(s) == 24 and len(s) == 5 and min(s) >= 1
    cm = bench.class_masks_2024(s)
    assert cm.shape == (24, 27) and cm[0, 10:].all() and not cm[0, :10].any() and cm[-1, :10].all()


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_work_model_covers_the_kernel_families_of_a_step():
    sys.path.insert(0, ROOT)
    import bench
    line = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_n1.json")))
    heavy = [k["kernel"] for k in line["kernel_breakdown_ms"] if k["ms_per_step"] >= 0.05]
    assert "gru_seq_bwd" in heavy and len(heavy) >= 10
    for name in heavy:
        flops, nbytes = bench.kernel_work(name, 24)
        assert flops > 0 and nbytes > 0, name
    # algorithmic front-end bytes per clip: 160 000 samples in + 128 x 626 log-mel out, fp32
    assert bench.kernel_work("logmel", 1)[1] == 160000 * 4 + 128 * 626 * 4 == 960512
    # names the parser must not choke on
    for name in ("conv3x3_tc5_128to128_F8", "conv_wgrad_tc5_64to128_F16", "conv3x3_16to32_F64", "gemm_tc5_NN_3744x256x384_k2",
                 "gemm_TN_384x256x3744_x2", "bnglu_tc5_bwd_c64", "glu_wgrad_tc5_c128", "some_future_kernel"):
        f, b = bench.kernel_work(name, 24)
        assert f >= 0 and b >= 0
