"""Per-kernel SASS summary of libswrb.so (cuobjdump -sass): instruction count and the mnemonics that show how each kernel
touches memory (global reductions / atomics, shared-memory atomics, 128-bit loads and stores, warp votes and shuffles).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "glimpsw_b200", "libswrb.so")
INTEREST = ["REDG", "RED.", "ATOMG", "ATOMS", "ATOM.", "LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDS.128", "STS.128", "VOTE", "SHFL", "MATCH",
            "MUFU.RCP", "MUFU.RSQ", "FFMA", "IMAD", "BAR.SYNC", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "I2F", "F2I", "VIMNMX", "VIADD"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), cuobjdump -sass; counts are static instructions")
    for (name, insts), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("swrb::", "")
        cnt = collections.Counter()
        for i in insts:
            op = i.split()[1] if i.startswith("@") else i.split()[0]
            for k in INTEREST:
                if op.startswith(k):
                    cnt[k] += 1
        keys = " ".join(f"{k}={v}" for k, v in cnt.items() if v)
        print(f"{short:60s} {len(insts):6d} instr  {keys}")


if __name__ == "__main__":
    main()
