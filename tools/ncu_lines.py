"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view, executed
instruction counts per address) with `nvdisasm -g` line info of the same cubin.

  cuobjdump -xelf all glimpsw_b200/libswrb.so && nvdisasm -g -c swrb.sm_100a.cubin > all.disasm
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_resolve > src.csv
  python tools/ncu_lines.py src.csv all.disasm k_resolveILb1 [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main(src_csv, disasm, fun_pat, top=45):
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hi]
    ci, cs, ct = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
    insts = []
    for r in rows[hi + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):     # next kernel instance in the same dump
            break
        if len(r) > ci:
            insts.append((r[cs].strip(), int(r[ci] or 0), int(r[ct] or 0)))
    # line info in program order
    lines, cur, on = [], ("?", 0), False
    for l in open(disasm):
        if l.startswith(".text."):
            on = fun_pat in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((cur, m.group(2).strip()))
    if len(lines) != len(insts):
        print(f"warning: {len(lines)} disasm instructions vs {len(insts)} profiled", file=sys.stderr)
    per_line, per_op = defaultdict(lambda: [0, 0]), defaultdict(int)
    total = sum(n for _, n, _ in insts)
    for (loc, _), (sass, n, smp) in zip(lines, insts):
        per_line[loc][0] += n
        per_line[loc][1] += smp
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        per_op[op.split(".")[0]] += n
    print(f"total warp instructions {total}")
    for loc, (n, smp) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{loc[0]}:{loc[1]:<5} {n:>10} {100.0 * n / total:5.1f}%  samples {smp}")
    print("-- by opcode")
    for op, n in sorted(per_op.items(), key=lambda kv: -kv[1])[:30]:
        print(f"{op:<10} {n:>10} {100.0 * n / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 45)
