"""Where a long one-thread-per-proof kernel spends its instructions and its time, from an ncu report captured with
--section SourceCounters:   ncu -i <report>.ncu-rep --page source --csv --print-source sass > sass.csv
                            python tools/ncu_segment_profile.py sass.csv out.json
The SASS listing of a kernel with out-of-line device functions is one flat list; it is cut at every RET / EXIT, so a segment is
(the tail of) one device function.  Per segment: static size, share of the executed warp instructions, share of the warp-state
samples (time), the dominant opcodes and the 32-bit immediates (which identify the routine: field / scalar-field constants)."""
import collections
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")

    def opcode(r):
        tok = r[isrc].split()
        return tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "")
    segs, cur = [], []
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        cur.append(r)
        if opcode(r).startswith(("RET", "EXIT")):
            segs.append(cur)
            cur = []
    if cur:
        segs.append(cur)
    num = lambda r, i: int(float(r[i] or 0))
    tot_ex = sum(num(r, iex) for s in segs for r in s)
    tot_sm = sum(num(r, ismp) for s in segs for r in s)
    out = []
    for i, s in enumerate(segs):
        ex, sm = sum(num(r, iex) for r in s), sum(num(r, ismp) for r in s)
        ops, imm = collections.Counter(), collections.Counter()
        for r in s:
            ops[opcode(r).split(".")[0]] += num(r, iex)
            for m in re.findall(r"0x[0-9a-f]{6,8}\b", r[isrc]):
                imm[m] += 1
        out.append({"segment": i, "static_instructions": len(s), "executed_pct": round(100.0 * ex / max(1, tot_ex), 1),
                    "samples_pct": round(100.0 * sm / max(1, tot_sm), 1),
                    "top_opcodes_pct": {o: round(100.0 * v / max(1, ex)) for o, v in ops.most_common(5)},
                    "immediates": [m for m, _ in imm.most_common(6)]})
    out.sort(key=lambda d: -d["samples_pct"])
    res = {"kernel": rows[0][1] if len(rows[0]) > 1 else "", "warp_instructions_executed": tot_ex, "samples": tot_sm, "segments": out[:12]}
    json.dump(res, open(sys.argv[2], "w"), indent=1)
    for d in out[:8]:
        print(d)


if __name__ == "__main__":
    main()
