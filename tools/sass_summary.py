"""Per-kernel SASS mnemonic counts of libvecsearch_b200.so (cuobjdump -sass): the proof that the hot kernels are
sm_100a-native (tcgen05 = UTCHMMA / UTCBAR / LDTM, TMA = UTMALDG / UTMAPF / UBLKCP, mbarrier = SYNCS).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-image-similarity-search_b200", "libvecsearch_b200.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "UTCCP", "LDS.128", "HMMA", "FFMA",
        "FMUL2", "REDUX", "ACQBULK", "MEMBAR", "ERRBAR", "RED", "ATOMG", "LDC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=900).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if cur and m:
            op = m.group(1)
            per[cur][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                per[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("LDS") and ".128" in op:
                per[cur]["LDS.128"] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: architectures {arch}, {len(per)} kernels")
    print("# counts of selected SASS mnemonics per kernel (static instruction counts)")
    total = collections.Counter()
    for (name, c), dn in zip(per.items(), demangled):
        short = re.sub(r"\(.*", "", dn)[:110]
        cells = " ".join(f"{w}={c[w]}" for w in WANT if c[w])
        total.update({w: c[w] for w in WANT})
        print(f"{short:110s} instr={sum(v for k, v in c.items() if '.' not in k):6d}  {cells}")
    print("# totals: " + " ".join(f"{w}={total[w]}" for w in WANT if total[w]))


if __name__ == "__main__":
    sys.exit(main())
