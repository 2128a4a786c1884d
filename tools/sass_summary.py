"""Per-kernel SASS mnemonic counts of gptorch_b200/lib/libgpb200.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

FP64 has no tcgen05 kind, so the tensor path of this library is DMMA (mma.sync m8n8k4 f64) fed by TMA (UTMALDG) with
mbarrier synchronisation (SYNCS); LDGSTS is cp.async.  The table is the evidence for which kernels use which."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gptorch_b200", "lib", "libgpb200.so")
WATCH = ["DMMA", "DFMA", "UTMALDG", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "MUFU", "BAR", "SHFL", "UTCHMMA", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[name][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel, %s, arch %s, %d kernels" % (os.path.relpath(LIB, ROOT), ",".join(sorted(arch)), len(kernels)))
    print("# %-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % w for w in WATCH))
    agg = collections.OrderedDict()
    for (mangled, cnt), nice in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", nice).replace("void ", "").replace("gpb::", "")
        base = re.sub(r"<.*", "", short)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(cnt)
        if "--all" in sys.argv:
            print("%-80s %7d " % (short[:80], cnt["total"]) + " ".join("%7d" % cnt[w] for w in WATCH))
    if "--all" not in sys.argv:
        for base, (n, cnt) in agg.items():
            label = "%s  [%d instantiation%s]" % (base, n, "" if n == 1 else "s")
            print("%-80s %7d " % (label[:80], cnt["total"]) + " ".join("%7d" % cnt[w] for w in WATCH))
    tot = collections.Counter()
    for _, cnt in agg.values():
        tot.update(cnt)
    print("%-80s %7d " % ("TOTAL", tot["total"]) + " ".join("%7d" % tot[w] for w in WATCH))


if __name__ == "__main__":
    main()
