"""Top SASS instructions by warp-stall samples / executed count from `ncu --page source --csv` (per kernel)."""
import csv
import sys


def main(path, want=None, top=24):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    hdr = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and hdr is not None and len(r) == len(hdr):
            cur["rows"].append(r)
    si, ni, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    for kid, k in enumerate(kernels):
        if want is not None and kid not in want:
            continue
        data = [(float(r[si] or 0), float(r[ii] or 0), idx, r[ni].strip()) for idx, r in enumerate(k["rows"])]
        ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
        print("=== kernel %d: %s\n    %d SASS instructions, %.0f stall samples, %.0f warp-instructions executed" %
              (kid, k["name"][:90], len(data), ts, ti))
        print("  -- top by stall samples (share of samples | share of executed instr | index | SASS)")
        for d in sorted(data, key=lambda d: -d[0])[:top]:
            print("  %5.1f%% %5.1f%% #%-5d %s" % (100 * d[0] / ts, 100 * d[1] / ti, d[2], d[3][:100]))
        ops = {}
        for d in data:
            op = d[3].split()[0] if not d[3].startswith("@") else d[3].split()[1]
            op = op.split(".")[0]
            o = ops.setdefault(op, [0.0, 0.0])
            o[0] += d[0]
            o[1] += d[1]
        print("  -- by opcode (stall share | instr share)")
        for op, (s, i) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
            print("  %5.1f%% %5.1f%%  %s" % (100 * s / ts, 100 * i / ti, op))


if __name__ == "__main__":
    want = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
    main(sys.argv[1], want)
