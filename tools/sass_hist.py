#!/usr/bin/env python3
"""SASS opcode histogram per kernel of dav1d_b200/libb200av1.so (cuobjdump -sass; no GPU needed).
usage: tools/sass_hist.py [out.md]   -> profiles/<round>_sass_opcodes.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dav1d_b200", "libb200av1.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.md")
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    with open(out_path, "w") as fh:
        fh.write("# SASS opcode histograms, libb200av1.so (sm_100a), `tools/sass_hist.py`\n\n"
                 "Static instruction counts per kernel (`cuobjdump -sass`), ten most frequent opcodes, and the markers that matter here:\n"
                 "`IDP` = dp4a / dp2a, `VIADDMNMX` / `VIMNMX` = fused add + clamp, `LDL` / `STL` = local memory (spills or dynamically\n"
                 "indexed arrays), `PREEXIT` / `ACQBULK` = programmatic dependent launch, `I2F` / `MUFU` / `F2I` = float-reciprocal integer\n"
                 "division, `UTMALDG` / `UBLKCP` = TMA (none: tiles are small, ragged and clamped; see DESIGN.md §4).\n\n")
        fh.write("| kernel | instructions | top opcodes | IDP | VIADDMNMX+VIMNMX | LDL+STL | MUFU | PDL |\n|---|---|---|---|---|---|---|---|\n")
        for k, c in kernels.items():
            n = sum(c.values())
            if n < 40:
                continue
            top = ", ".join("%s %d" % (o, v) for o, v in c.most_common(10))
            short = re.sub(r"\(.*", "", names.get(k, k)).replace("void ", "")
            tmpl = re.search(r"<[^>]*>", names.get(k, ""))
            fh.write("| `%s%s` | %d | %s | %d | %d | %d | %d | %s |\n" % (short, "", n, top, c["IDP"], c["VIADDMNMX"] + c["VIMNMX"],
                                                                       c["LDL"] + c["STL"], c["MUFU"], "yes" if c["PREEXIT"] else "no"))
    print("wrote", out_path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
