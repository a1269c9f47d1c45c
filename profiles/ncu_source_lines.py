"""Aggregates `ncu --page source --csv --print-source sass,cuda` per CUDA source line: stall samples and instructions.
usage: python profiles/ncu_source_lines.py dump.csv [top_n] [function-substring]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
rows = csv.reader(open(path))
agg = {}; hdr = None; fn = ""; keep = True; line = None; src = ""
for r in rows:
    if len(r) == 2 and r[0] == "Function Name": fn = r[1]; keep = want in fn; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not keep: continue
    d = dict(zip(hdr, r))
    # rows with a line number are source lines (aggregates); rows with '-' line are SASS
    ln = r[0]
    if ln != "-" and ln != "":
        try:
            samples = int(r[hdr.index("# Samples")] or 0); inst = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError: continue
        stalls = {k: int(d[k] or 0) for k in d if k.startswith("stall_") and "Not Issued" not in k}
        key = (fn[:40], int(ln))
        a = agg.setdefault(key, [r[1][:110], 0, 0, collections.Counter()])
        a[1] += samples; a[2] += inst; a[3].update(stalls)
tot = sum(a[1] for a in agg.values()) or 1
toti = sum(a[2] for a in agg.values()) or 1
print(f"total samples {tot}, warp-instructions {toti}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    top3 = ", ".join(f"{k[6:]}={v}" for k, v in a[3].most_common(3))
    print(f"L{ln:4d} {100*a[1]/tot:5.1f}% samp {100*a[2]/toti:5.1f}% inst | {top3:45s} | {a[0]}")
