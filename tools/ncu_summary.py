"""Turn `ncu --page raw --csv` exports (and a launch list) into a compact markdown summary for profiles/."""
import collections
import csv
import re
import sys

KEYS = [("gpu__time_duration.sum", "t_us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("sass__inst_executed_register_spilling", "spill_inst"), ("smsp__inst_executed.sum", "inst")]


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[hi], rows[hi + 1], rows[hi + 2:]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void sedk::<unnamed>::", "").replace("sedk::<unnamed>::", "").replace("void ", "")


def raw_table(path):
    hdr, units, data = load(path)
    out = ["| kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|" + "---|" * len(KEYS)]
    for r in data:
        vals = []
        for full, _ in KEYS:
            if full in hdr:
                v = r[hdr.index(full)]
                u = units[hdr.index(full)]
                try:
                    f = float(v.replace(",", ""))
                    if full.endswith("time_duration.sum") and u == "ns":
                        f /= 1000.0
                    if "bytes" in full and u == "byte":
                        f /= 1e6
                    if "bytes" in full and u == "Kbyte":
                        f /= 1e3
                    vals.append("%.4g" % f)
                except ValueError:
                    vals.append(v)
            else:
                vals.append("-")
        out.append("| `%s` | " % short(r[hdr.index("Kernel Name")])[:60] + " | ".join(vals) + " |")
    return "\n".join(out)


def launch_table(path):
    hdr, units, data = load(path)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in [units] + data:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1000.0
        elif r[ui] == "ms":
            v *= 1000.0
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out = ["total %.1f us over %d launches (cold-cache, serialised under ncu: compare shares)" % (tot, sum(a[0] for a in agg.values())),
           "", "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k[:70], c, t, 100 * t / tot))
    return "\n".join(out)


def family(name):
    """ncu kernel name -> the launcher family name bench.py's per-kernel breakdown uses."""
    n = short(name)
    m = re.search(r"<(?:\(int\))?(\d+)", name)
    c = m.group(1) if m else ""
    if "gru_fwd" in n:
        return "gru_seq_fwd"
    if "gru_bwd" in n:
        return "gru_seq_bwd"
    if "logmel" in n:
        return "logmel"
    if "bnglu_tc5_fwd" in n:
        return "bnglu_tc5_fwd_c" + c
    if "bnglu_tc5_bwd" in n:
        return "bnglu_tc5_bwd_c" + c
    if "bnglu_small_fwd" in n or "bnglu_fwd" in n:
        return "bnglu_pool_fwd_c" + c
    if "bnglu_small_bwd" in n or "bnglu_bwd" in n:
        return "bnglu_pool_bwd_c" + c
    for k in ("l0_x0", "l0_stats", "l0_fwd", "l0_bwd", "l0_finish"):
        if k + "_kernel" in n:
            return k
    if "conv0_fwd" in n:
        return "conv0_fwd"
    if "conv0_wgrad" in n:
        return "conv0_wgrad"
    return n.split("<")[0].split("::")[-1]


def traffic_json(path):
    """{family: mean dram__bytes_read.sum + dram__bytes_write.sum per launch} from a `--page raw --csv` export."""
    import json
    hdr, units, data = load(path)
    ki = hdr.index("Kernel Name")
    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = {}
    for r in data:
        try:
            b = float(r[ri].replace(",", "")) * scale.get(units[ri], 1.0) + float(r[wi].replace(",", "")) * scale.get(units[wi], 1.0)
        except ValueError:
            continue
        a = agg.setdefault(family(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += b
    return json.dumps({k: round(v[1] / v[0]) for k, v in sorted(agg.items())}, indent=1)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    print(raw_table(path) if mode == "raw" else traffic_json(path) if mode == "traffic" else launch_table(path))
