"""Development aid: per-source-line stall samples / executed instructions of one kernel captured with
`ncu --set full --import-source on`, by joining the report's SASS page with `nvdisasm -g` line info of the SAME build.

  python tools/sass_lines.py <file.ncu-rep> <kernel substring> [cubin dir or .so] [top N]

The .so (default atlas_engine_b200/libatlas_rt.so) must be the binary that was profiled.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_page(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    ia, iss, ii, isamp = H.index("Address"), H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
    res = []
    for r in rows[hdr + 1:]:
        if len(r) <= isamp or not r[ia].startswith("0x"):
            break
        res.append((int(r[ia], 16), r[iss].strip(), int(r[ii] or 0), int(r[isamp] or 0)))
    base = res[0][0]
    return [(a - base, s, i, n) for a, s, i, n in res]


def line_table(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    table = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        inside, cur = False, None
        for ln in txt.splitlines():
            if ln.startswith("//-----") and ".text." in ln:
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
            if m and cur:
                table[int(m.group(1), 16)] = cur
        if table:
            break
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "atlas_engine_b200", "libatlas_rt.so")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    sass = sass_page(rep, kernel)
    lines = line_table(so, kernel)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    missing = 0
    for off, src, inst, samp in sass:
        key = lines.get(off)
        if key is None:
            missing += 1
            key = ("?", 0)
        a = agg[key]
        a[0] += samp
        a[1] += inst
        a[2] += 1
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"{len(sass)} SASS instructions, {missing} without line info; {ts} samples, {ti} warp instructions executed")
    srcs = {}
    print(f"{'samples':>8s} {'%':>5s} {'inst exec':>11s} {'%':>5s} {'sass':>5s}  line")
    for (f, l), (s, i, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = os.path.join(ROOT, "atlas_engine_b200", "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:110] if 0 < l <= len(srcs[f]) else ""
        print(f"{s:8d} {100 * s / max(ts, 1):5.1f} {i:11d} {100 * i / max(ti, 1):5.1f} {c:5d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
