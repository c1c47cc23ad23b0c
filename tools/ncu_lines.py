"""Per-source-line instruction / stall summary of one kernel of an ncu report (captured with --import-source on from a
-lineinfo build): `python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]`. Read on the CPU side; the report
itself stays in gpurun_out/ (scratch)."""
import csv, io, os, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
fpath, hdr, lines, seen_fn = "?", None, [], None
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if len(r) >= 2 and r[0] == "Function Name":
        if seen_fn is None:
            seen_fn = r[1]
        cur_fn = r[1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ii or not r[0].isdigit() or cur_fn != seen_fn:
        continue
    try:
        lines.append((int(r[ii]), int(r[si] or 0), fpath, int(r[0]), r[1].strip()))
    except ValueError:
        pass
ti = sum(l[0] for l in lines); ts = sum(l[1] for l in lines)
print("%s: %d warp instructions, %d stall samples over %d source lines" % (seen_fn, ti, ts, len(lines)))
for i, s_, f, ln, text in sorted(lines, reverse=True)[:top]:
    print("%6.2f%% inst %6.2f%% stall  %s:%d: %s" % (100.0 * i / max(ti, 1), 100.0 * s_ / max(ts, 1), f, ln, text[:110]))
