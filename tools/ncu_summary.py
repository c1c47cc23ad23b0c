#!/usr/bin/env python3
"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md + .json.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_itx8x8 [note]"""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    # a .csv argument = the `ncu -i rep --page raw --csv` dump made on the GPU box (reports of a dozen kernels exceed what gpurun copies back)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT.get(units[i], 1.0) if "bytes" in k else v
                d[k + "|unit"] = "byte" if "bytes" in k else units[i]
        if "dram__bytes_read.sum" in d:
            d["dram_bytes_per_launch"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        res.append(d)
    json.dump({"report": rep, "note": note, "launches": res}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu summary: %s\n\n%s\n\nSource: `%s` (ncu --set full --clock-control none; per-launch, cold-cache, serialised)\n\n" % (out, note, rep))
        for d in res:
            f.write("## %s\n\n| metric | value |\n|---|---|\n" % d["kernel"][:120])
            for k in KEYS + ["dram_bytes_per_launch"]:
                if k in d:
                    f.write("| %s | %.6g %s |\n" % (k, d[k], d.get(k + "|unit", "byte")))
            f.write("\n")
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
