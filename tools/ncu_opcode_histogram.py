"""Dynamic SASS opcode histogram of one kernel from an ncu report captured with --section SourceCounters:
    ncu -i <report>.ncu-rep --page source --csv > src.csv ; python tools/ncu_opcode_histogram.py src.csv out.json
Sums "Instructions Executed" (warp level) per opcode over the kernel AND the device functions it calls (the source page lists
them all), groups the opcodes by issue pipe, and keeps the stall samples of the hottest opcodes."""
import collections
import csv
import json
import sys

PIPE = {"IMAD.WIDE": "fma-heavy (32x32->64 multiply)", "IMAD.HI": "fma-heavy (32x32->64 multiply)", "IMAD.MOV": "fma-lite / alu (register move)",
        "IMAD.IADD": "fma (add through the multiplier)", "IMAD.SHL": "fma (shift through the multiplier)", "IMAD.X": "fma (multiply-add with carry)",
        "IMAD": "fma (32-bit multiply-add)", "IADD3": "alu (integer add / carry chain)", "LOP3": "alu (logic)", "SHF": "alu (shift)", "SEL": "alu (select)",
        "PRMT": "alu (byte permute)", "ISETP": "alu (compare)", "MOV": "alu (move)", "LDL": "lsu (local load: spill / table)", "STL": "lsu (local store: spill / table)",
        "LDG": "lsu (global load)", "STG": "lsu (global store)", "LD": "lsu (generic load)", "ST": "lsu (generic store)", "LDS": "lsu (shared load)",
        "STS": "lsu (shared store)", "CALL": "branch (call)", "RET": "branch (return)", "BRA": "branch", "BAR": "barrier", "SHFL": "lsu (shuffle)"}


def pipe_of(op):
    for k in sorted(PIPE, key=len, reverse=True):
        if op.startswith(k):
            return PIPE[k]
    return "other"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    name = rows[0][1] if rows and len(rows[0]) > 1 else ""
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= iex or not r[isrc]:
            continue
        tok = r[isrc].split()
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        ops[op] += int(float(r[iex] or 0))
        smp[op] += int(float(r[ismp] or 0))
    tot, stot = sum(ops.values()), sum(smp.values())
    pipes = collections.Counter()
    for op, v in ops.items():
        pipes[pipe_of(op)] += v
    out = {"kernel": name, "warp_instructions_executed": tot, "sass_lines": len(rows) - 2,
           "by_pipe_pct": {k: round(100.0 * v / tot, 2) for k, v in pipes.most_common()},
           "top_opcodes": [{"op": op, "executed": v, "pct": round(100.0 * v / tot, 2), "stall_samples_pct": round(100.0 * smp[op] / max(1, stot), 2)}
                           for op, v in ops.most_common(24)]}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out["by_pipe_pct"]))


if __name__ == "__main__":
    main()
