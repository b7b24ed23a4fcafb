"""Assemble profiles/r2_summary.md (+ copies of the evidence files) from a tools/gpu_r2.sh session in gpurun_out/.

    python tools/r2_summary.py r2z
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def jl(path):
    try:
        for line in reversed(open(path).read().strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"
    md = ["# Round-2 evidence summary (B200, session `%s`)" % tag, "",
          "Numbers printed under ncu are never bench values; bench lines come from un-profiled runs of `bench.py`.", ""]
    # ---- bench lines
    md += ["## Bench lines (`bench.py`, N = 1 unless noted)", "",
           "| workload | value clips/s | ms/step | e2e clips/s | repeats min / median ms | GPU-library bar (best) | × library | CPU reference |",
           "|---|---|---|---|---|---|---|---|"]
    files = [("supervised", "%s_bench_n1.json" % tag), ("mean_teacher", "%s_bench_mean_teacher.json" % tag),
             ("inference", "%s_bench_inference.json" % tag), ("dcase2024", "%s_bench_dcase2024.json" % tag)]
    sup = None
    for wl, f in files:
        d = jl(os.path.join(OUT, f))
        if d is None:
            continue
        shutil.copy(os.path.join(OUT, f), os.path.join(PROF, "r2_bench_%s.json" % wl))
        if wl == "supervised":
            sup = d
        lib = d.get("gpu_library_baseline") or {}
        cpu = d.get("cpu_baseline") or {}
        md.append("| %s | %s | %s | %s | %s / %s | %s | %s | %s |" % (
            wl, d["value"], d["ms_per_step"], d["e2e"]["value"], d["repeats"]["ms_per_step_min"],
            d["repeats"]["ms_per_step_median"], lib.get("best", "-"), d.get("vs_gpu_library", "-"),
            ("%s clips/s (%s, %s threads)" % (cpu.get("value"), cpu.get("kind"), cpu.get("cores"))) if cpu else "-"))
    multi = (("r2N_bench_n2_supervised.json", "r2_bench_n2.json", "supervised N=2 (fused NVLink all-reduce + Adam)"),
             ("r2N_bench_n2_mean_teacher.json", "r2_bench_mean_teacher_n2.json", "mean_teacher N=2 (48 clips/GPU, nvls)"),
             ("r2N_bench_n2_dcase2024.json", "r2_bench_dcase2024_n2.json", "dcase2024 N=2 (nvls all-reduce, clip, Adam)"),
             ("r2M_bench_n8_nvls.json", "r2_bench_supervised_n8.json", "supervised N=8 (fused NVLink all-reduce + Adam)"),
             ("r2M_bench_n8_eager.json", "r2_bench_supervised_n8_nccl.json", "supervised N=8 (SEDK_AR_MODE=eager: NCCL)"),
             ("r2m_bench_mean_teacher_n8.json", "r2_bench_mean_teacher_n8.json",
              "mean_teacher N=8 (48 clips/GPU; NCCL schedule, before the store-free first block)"),
             ("r2m_bench_inference_n8.json", "r2_bench_inference_n8.json",
              "inference N=8 (64 clips/GPU; before the store-free first block)"))
    for f, dst, label in multi:
        d = jl(os.path.join(OUT, f))
        if d is not None:
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, dst))
            md.append("| %s | %s | %s | %s | %s / %s | | | |" % (label, d["value"], d["ms_per_step"], d["e2e"]["value"],
                                                             d["repeats"]["ms_per_step_min"], d["repeats"]["ms_per_step_median"]))
    ref = jl(os.path.join(OUT, "%s_bench_ref.json" % tag))
    if ref is not None:
        shutil.copy(os.path.join(OUT, "%s_bench_ref.json" % tag), os.path.join(PROF, "r2_bench_reference.json"))
        md.append("| `--impl reference` (%s) | %s | %s | | | | | |" % (ref["cpu_baseline"]["kind"], ref["value"], ref["ms_per_step"]))
    md.append("")
    if sup is not None:
        md += ["## Kernel families of the supervised step (eager pass with CUDA events around every launcher)", "",
               "| family | ms/step | launches | algorithmic MB | GFLOP | GB/s | TFLOP/s | bound | frac of peak | DRAM / algorithmic |",
               "|---|---|---|---|---|---|---|---|---|---|"]
        for f in sup["kernel_families"]:
            md.append("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % tuple(
                f.get(k, "-") for k in ("family", "ms_per_step", "launches_per_step", "algorithmic_MB", "algorithmic_GFLOP",
                                        "GBps", "TFLOPs", "bound", "frac", "dram_over_algorithmic")))
        md += ["", "GRU: %s us per dependent time step.  Front end: %s." % (sup["gru_us_per_time_step"], sup["frontend"]),
               "", "`roofline`: `%s`" % json.dumps(sup["roofline"]), ""]
    # ---- ncu
    for name, title, mode in (("%s_launches_supervised.csv" % tag, "ncu launch list of one eager supervised step", "launches"),
                              ("%s_hot_raw.csv" % tag, "ncu --set full metrics of the hot kernels", "raw")):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), mode, p], capture_output=True, text=True)
            md += ["## " + title, "", r.stdout.strip(), ""]
            if mode == "launches":
                shutil.copy(p, os.path.join(PROF, "r2_launches_supervised.csv"))
            else:
                subprocess.run("gzip -c %s > %s" % (p, os.path.join(PROF, "r2_hot_raw.csv.gz")), shell=True)
                t = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "traffic", p], capture_output=True, text=True)
                try:
                    new = json.loads(t.stdout)
                    old = json.load(open(os.path.join(PROF, "traffic.json")))
                    old.update(new)
                    json.dump(old, open(os.path.join(PROF, "traffic.json"), "w"), indent=1, sort_keys=True)
                except Exception:
                    pass
    for src, dst in (("%s_hot_sass.txt" % tag, "r2_hot_sass.txt"), ("%s_pytest_gpu.log" % tag, "r2_pytest_gpu.log"),
                     ("%s_smoke.log" % tag, "r2_smoke.log"), ("%s_frontend.txt" % tag, "r2_frontend_ab.txt"),
                     ("%s_timeline.txt" % tag, "r2_timeline.txt"), ("r2_parity_margins.txt", "r2_parity_margins.txt"),
                     ("r2b_ubench.txt", "r2_ubench.txt"), ("r2a_library_bar.json", "r2_library_bar_first_pass.json")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    for title, f in (("Store-free first block: A/B against the kernels it replaces (tools/diag_l0.py, tools/bench_ab.py)", "r2_layer0_ab.txt"),
                     ("Fused NVLink all-reduce + Adam: check and timing at 2 and 8 ranks (tools/nvls_check.py)", "r2_nvls_check.txt"),
                     ("Host-side input path on the build container's CPU (tools/bench_io.py)", "r2_io_cpu.txt")):
        fp = os.path.join(PROF, f)
        if os.path.exists(fp):
            md += ["", "## " + title, "", "```", open(fp).read().rstrip(), "```"]
    open(os.path.join(PROF, "r2_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md)[:3000])


if __name__ == "__main__":
    main()
