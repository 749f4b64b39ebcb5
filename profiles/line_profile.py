"""Per-CUDA-source-line attribution of an ncu capture (taken with --import-source on / -lineinfo): joins the SASS page of the
report with `nvdisasm -g` of the object file (same instruction order) and sums executed instructions, samples and active lanes
per source line.   usage: line_profile.py report.ncu-rep object.o mangled_function_substring source.cu [top_n]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, fn, src_path = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and fn in l)
    end = next((i for i, l in enumerate(dis) if i > start and l.startswith(".text.")), len(dis))
    cur, inst = None, []
    base = os.path.basename(src_path)
    for l in dis[start:end]:
        if "//## File" in l:
            m = re.search(r'//## File "([^"]*)", line (\d+)', l)
            cur = int(m.group(2)) if m and os.path.basename(m.group(1)) == base else -1
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            inst.append(cur)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[h], rows[h + 1:]
    data = data[:len(inst)] if len(data) >= len(inst) else data
    assert len(data) == len(inst), (len(data), len(inst))
    iex, ith, ism = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for ln, r in zip(inst, data):
        a = agg[ln]
        a[0] += int(r[iex] or 0); a[1] += int(r[ith] or 0); a[2] += int(r[ism] or 0)
    tot, tots, totth = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values()), sum(a[1] for a in agg.values())
    src = open(src_path).read().split("\n")
    print(f"# {os.path.basename(rep)} / {fn}: {tot} warp instructions, {totth / max(tot, 1):.1f} active lanes on average, {tots} samples")
    print("# %inst  %samples  lanes  line: source")
    for ln, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        t = src[ln - 1].strip()[:100] if ln and 0 < ln <= len(src) else "(inlined from another file)"
        print(f"{a[0] / tot * 100:5.1f}  {a[2] / max(tots, 1) * 100:5.1f}  {a[1] / max(a[0], 1):4.1f}  L{ln}: {t}")


if __name__ == "__main__":
    main()
