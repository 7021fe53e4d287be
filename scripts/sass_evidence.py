"""SASS evidence for libpulpo_b200.so (no GPU needed): per kernel, the Blackwell-specific / design-relevant
mnemonics the DESIGN.md claims rest on, from `cuobjdump -sass` of the in-tree library.
    python scripts/sass_evidence.py > profiles/r2_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pulpo_b200", "lib", "libpulpo_b200.so")
COLS = [("UTMALDG", r"^UTMALDG"), ("SYNCS (mbarrier)", r"^SYNCS"), ("LDGSTS (cp.async)", r"^LDGSTS"),
        ("RED .128 f32x4", r"^REDG?\..*F32x4|^RED\..*F32x4"), ("RED/ATOM other", r"^(REDG?|ATOMG?)\b(?!.*F32x4)"),
        ("FFMA2", r"^FFMA2"), ("FMUL2", r"^FMUL2"), ("FADD2", r"^FADD2"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)\b"),
        ("LDG", r"^LDG"), ("LDS", r"^LDS"), ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("CCTL.IVALL", r"^CCTL"),
        ("MEMBAR", r"^MEMBAR")]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def kernel_name(d):
    """`void pulpo::k<(int)9, (bool)1>(args...)` -> `k<9, 1>`"""
    d = d.replace("void ", "").replace("pulpo::", "")
    depth, end = 0, len(d)
    for i, ch in enumerate(d):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            end = i
            break
    return re.sub(r"\((int|bool)\)", "", d[:end])


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    cnt, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?(\S+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for name, pat in COLS:
                if re.search(pat, op):
                    cnt[cur][name] += 1
    names = sorted(total)
    dm = demangle(names)
    print("# SASS evidence (round 2) -- `cuobjdump -sass pulpo_b200/lib/libpulpo_b200.so`, sm_100a\n")
    print("Static instruction counts per kernel (whole kernel, all paths).  TMA = `UTMALDG` + `SYNCS` mbarriers; vector")
    print("reductions = `RED*.F32x4`; cp.async = `LDGSTS`; packed fp32 = `FFMA2 / FMUL2 / FADD2` (Blackwell f32x2).  No")
    print("tensor-core (`UTC*MMA`, `HMMA`) instruction appears anywhere: none of these kernels is a contraction.\n")
    print("| kernel | regs | smem B | instrs | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(COLS))
    for n in names:
        short = kernel_name(dm.get(n, n))
        r = regs.get(n, ("", ""))
        print("| `%s` | %s | %s | %d | " % (short, r[0], r[1], total[n]) + " | ".join(str(cnt[n][c]) if cnt[n][c] else "" for c, _ in COLS) + " |")
    tc = len(re.findall(r"UTC\w*MMA|HMMA|HGMMA", txt))
    print("\nTensor-core mnemonics (`UTC*MMA|HMMA|HGMMA`): %d." % tc)


if __name__ == "__main__":
    main()
