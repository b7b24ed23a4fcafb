"""Per-kernel SASS opcode census of libsedk.so (cuobjdump -sass): which kernels use tcgen05 (UTC*MMA), TMEM (LDTM / STTM),
TMA (UTMALDG / UBLKCP), mbarriers (SYNCS), legacy tensor cores (HMMA), packed fp32 (FFMA2 / FADD2 / FMUL2), clusters.

    python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "desed_task_b200", "lib", "libsedk.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCBAR", "LDGMC", "HMMA", "FFMA2", "FADD2", "FMUL2",
       "FFMA", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "UCGABAR", "MAPA", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {"n": 0}
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["n"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    kernels[cur][o] = kernels[cur].get(o, 0) + 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, dem)) if len(dem) == len(kernels) else {k: k for k in kernels}
    print("# SASS opcode census of desed_task_b200/lib/libsedk.so (sm_100a); columns = static instruction counts per kernel")
    print("# kernel | total | " + " | ".join(OPS))
    tot = {o: 0 for o in OPS}
    for k in sorted(kernels, key=lambda k: names[k]):
        c = kernels[k]
        short = re.sub(r"\(.*", "", names[k]).replace("sedk::(anonymous namespace)::", "").replace("void ", "")
        print("%-60s | %5d | %s" % (short[:60], c["n"], " | ".join(str(c.get(o, 0)) for o in OPS)))
        for o in OPS:
            tot[o] += c.get(o, 0)
    print("%-60s | %5s | %s" % ("TOTAL", "", " | ".join(str(tot[o]) for o in OPS)))
    uses = lambda o: sorted({re.sub(r"<.*", "", re.sub(r"\(.*", "", names[k]).split("::")[-1]) for k in kernels if kernels[k].get(o)})
    print("\n# kernels issuing tcgen05.mma (UTCHMMA):", ", ".join(uses("UTCHMMA")))
    print("# kernels using TMEM loads/stores (LDTM/STTM):", ", ".join(sorted(set(uses("LDTM")) | set(uses("STTM")))))
    print("# kernels using TMA (UTMALDG tensor / UBLKCP bulk):", ", ".join(sorted(set(uses("UTMALDG")) | set(uses("UBLKCP")))))
    print("# kernels on legacy mma.sync (HMMA):", ", ".join(uses("HMMA")))
    print("# kernels using packed fp32 (FFMA2/FADD2/FMUL2):", ", ".join(sorted(set(uses("FFMA2")) | set(uses("FADD2")) | set(uses("FMUL2")))))
    print("# kernels with cluster barriers / DSMEM (UCGABAR / MAPA):", ", ".join(sorted(set(uses("UCGABAR")) | set(uses("MAPA")))))
    print("# kernels reducing through the NVSwitch (LDGMC = multimem.ld_reduce; multimem.st is an STG to the multicast address):",
          ", ".join(uses("LDGMC")))


if __name__ == "__main__":
    main()
