#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: instruction mix, stall samples per SASS region.

    python tools/src_profile.py gpurun_out/x_src.csv [pixels] [chunk]
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    return rows[0], col, data


def opname(src):
    parts = src.split()
    if not parts:
        return ""
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    return op.rstrip(";")


def main():
    path = sys.argv[1]
    npx = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0 * 8192.0
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    name, col, data = load(path)
    ie, isamp, isrc = col["Instructions Executed"], col["# Samples"], col["Source"]
    stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ie]) for r in data)
    tots = sum(int(r[isamp]) for r in data)
    print(name[1][:120] if len(name) > 1 else name)
    print("warp instructions %d = %.1f per thread-pixel; samples %d" % (tot, tot * 32 / npx, tots))
    byop, sop = collections.Counter(), collections.Counter()
    for r in data:
        byop[opname(r[isrc])] += int(r[ie])
        sop[opname(r[isrc])] += int(r[isamp])
    print("-- instruction mix (share of executed, share of stall samples)")
    for op, n in byop.most_common(28):
        print("  %-26s %6.1f/px %5.1f%% %5.1f%%" % (op, n * 32 / npx, 100.0 * n / tot, 100.0 * sop[op] / max(tots, 1)))
    st = collections.Counter()
    for r in data:
        for h, i in stall_cols:
            try:
                st[h] += int(r[i])
            except ValueError:
                pass
    print("-- stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * n / max(tots, 1)) for h, n in st.most_common(8)))
    print("-- regions of %d SASS instructions (executed only)" % chunk)
    for i in range(0, len(data), chunk):
        c = data[i:i + chunk]
        n = sum(int(r[ie]) for r in c)
        s = sum(int(r[isamp]) for r in c)
        if n == 0:
            continue
        ops = collections.Counter()
        for r in c:
            ops[opname(r[isrc])] += int(r[ie])
        print("  %5d  %6.1f/px  samples %5.1f%%  %s" % (i, n * 32 / npx, 100.0 * s / max(tots, 1),
                                                        " ".join("%s:%.1f" % (o, k * 32 / npx) for o, k in ops.most_common(5))))
    print("-- hottest instructions by samples")
    hot = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:25]
    for k in sorted(hot):
        r = data[k]
        print("  %5d %5.1f%% x%-9d %s" % (k, 100.0 * int(r[isamp]) / max(tots, 1), int(r[ie]), r[isrc][:90]))


if __name__ == "__main__":
    main()
