#!/usr/bin/env python3
"""profiles/<round>_sass.md: `cuobjdump -sass` instruction histogram of every kernel in libtfm_b200.so.

    python tools/sass_histogram.py r2

Runs in the authoring container (no GPU).  The mnemonics that identify the Blackwell-specific paths are
called out per kernel: TLD4 (texture gather), UTMALDG (TMA bulk tensor load), SYNCS (mbarrier), FFMA2 /
FMUL2 / FADD2 (packed fp32x2 arithmetic, sm_100+), LDG.E.64 (paired tap loads)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "transform_b200", "lib", "libtfm_b200.so")
MARK = ["TLD4", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDG.E.64", "LDG.E.128", "STG.E.128", "DFMA",
        "MUFU", "LDS", "STS", "BAR", "HMMA", "UTC"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            kernels[name][m.group(1).rstrip(";")] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (k, cnt), dn in zip(kernels.items(), demangled):
        total = sum(cnt.values())
        marks = {m: sum(v for op, v in cnt.items() if op.startswith(m)) for m in MARK}
        rows.append((dn, total, marks, cnt))
    keep = [r for r in rows if "k_" in r[0]]
    with open(os.path.join(ROOT, "profiles", rnd + "_sass.md"), "w") as f:
        f.write("# SASS instruction histogram of libtfm_b200.so (cuobjdump -sass, sm_100a)\n\n")
        f.write("Static instruction counts per kernel.  TLD4 = texture gather, UTMALDG = `cp.async.bulk.tensor` (TMA), "
                "SYNCS = mbarrier arrive / try_wait, FFMA2 / FMUL2 / FADD2 = packed fp32x2 arithmetic (sm_100+), "
                "LDG.E.64 = paired tap loads.  No HMMA / UTC*MMA anywhere: nothing on this path is a contraction.\n\n")
        f.write("| kernel | SASS instr | " + " | ".join(MARK) + " |\n|---|---:|" + "---:|" * len(MARK) + "\n")
        for dn, total, marks, cnt in sorted(keep, key=lambda r: r[0]):
            short = re.sub(r"\(tfm::\w+Params.*$", "", dn).replace("tfm::", "").replace("void ", "")
            f.write("| `%s` | %d | %s |\n" % (short[:100], total, " | ".join(str(marks[m]) if marks[m] else "" for m in MARK)))
        f.write("\n## Headline kernels, top mnemonics\n\n")
        for pat in ("k_resample_pipe", "k_resample_tex4<tfm::Fast, true", "k_resample_tex4_batch<tfm::Fast, false, 1, true",
                    "k_jac_gauss_polar_rows_pair", "k_jac_gauss_polar_rows<2>", "k_jac_gauss_polar<2>",
                    "k_jac_gauss_const<2, 2>", "k_apply_f64x2", "k_apply_linear_nd<double>"):
            for dn, total, marks, cnt in keep:
                if pat in dn.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false"):
                    f.write("* `%s` (%d): %s\n" % (dn[:90], total, ", ".join("%s %d" % kv for kv in cnt.most_common(14))))
                    break
    print("wrote profiles/%s_sass.md: %d kernels" % (rnd, len(keep)))


if __name__ == "__main__":
    main()
