"""Static per-opcode histogram of the RK45 attempt (the hot loop) of a trace kernel, from the built library.

    python scripts/sass_hist.py [--lib blackhole_geodesic_calculator_b200/lib/libbhgeo.so] [--kernel 4,1,0,0]
                                [--dump FILE]

The attempt is straight-line code (branch-free sincos / reciprocal / clamps), so it shows up as a run of large
basic blocks full of FP64-pipe instructions.  The script splits the kernel's SASS into basic blocks at branch
instructions and branch targets, marks the blocks with >= 8 FP64-pipe instructions, takes the span from the first to
the last such block that belongs to the densest cluster (gaps of small blocks, e.g. the predicated cold pow call,
are bridged), and prints the opcode histogram of that span: FP64 pipe (DFMA / DMUL / DADD / DSETP / ...) against
everything else.  `--dump` writes the span's instructions (address, opcode, operands) to a file for profiles/.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "F2F.F64", "F2F.F32.F64", "I2F.F64", "F2I.F64", "DMMA")
INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")


def is_fp64(op):
    return op.startswith(FP64) or ".F64" in op and op.startswith(("F2F", "I2F", "F2I", "FRND"))


def kernel_sass(lib, tmpl):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    a = [int(v) for v in tmpl.split(",")]
    name = "trace_kernelILi%dELi%dELb%dELb%dELb%dELb%dEEE" % tuple(a)
    parts = txt.split("Function : ")
    for p in parts[1:]:
        if name in p.splitlines()[0]:
            return p
    raise SystemExit("kernel %s not found in %s" % (name, lib))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "lib", "libbhgeo.so"))
    ap.add_argument("--kernel", default="4,1,0,0,1,0")
    ap.add_argument("--dump", default=None)
    a = ap.parse_args()
    ins = []
    for ln in kernel_sass(a.lib, a.kernel).splitlines():
        m = INS.match(ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {ad: i for i, (ad, _) in enumerate(ins)}
    # basic-block leaders
    leaders = {0}
    for i, (ad, tx) in enumerate(ins):
        body = re.sub(r"^@!?U?P\d+\s+", "", tx)
        op = body.split()[0]
        if op.startswith(("BRA", "BSSY", "CALL", "RET", "EXIT", "BSYNC", "WARPSYNC", "JMP", "BRX")):
            if not op.startswith(("BSSY", "BSYNC", "WARPSYNC")) and i + 1 < len(ins):
                leaders.add(i + 1)
            for t in re.findall(r"0x([0-9a-f]+)", body):
                ti = addr_index.get(int(t, 16))
                if ti is not None and op.startswith(("BRA", "JMP")):
                    leaders.add(ti)
    leaders = sorted(leaders)
    blocks = []
    for bi, s in enumerate(leaders):
        e = leaders[bi + 1] if bi + 1 < len(leaders) else len(ins)
        nf = sum(is_fp64(re.sub(r"^@!?U?P\d+\s+", "", tx).split()[0]) for _, tx in ins[s:e])
        blocks.append((s, e, nf))
    big = [i for i, (s, e, nf) in enumerate(blocks) if nf >= 8]
    # cluster big blocks separated by < 40 small-block instructions; keep the cluster with most FP64 instructions
    clusters, cur = [], [big[0]]
    for b in big[1:]:
        gap = blocks[b][0] - blocks[cur[-1]][1]
        if gap < 40:
            cur.append(b)
        else:
            clusters.append(cur)
            cur = [b]
    clusters.append(cur)
    best = max(clusters, key=lambda c: sum(blocks[i][2] for i in c))
    s, e = blocks[best[0]][0], blocks[best[-1]][1]
    hist = collections.Counter()
    for _, tx in ins[s:e]:
        body = re.sub(r"^@!?U?P\d+\s+", "", tx)
        op = body.split()[0]
        key = op.split(".")[0]
        if op.startswith("IMAD.MOV"):
            key = "IMAD.MOV"
        hist[key] += 1
    n = e - s
    nf = sum(v for k, v in hist.items() if is_fp64(k))
    print(f"kernel <{a.kernel}>: {len(ins)} instructions; attempt span 0x{ins[s][0]:x}..0x{ins[e - 1][0]:x}: "
          f"{n} instructions, FP64 pipe {nf}, other {n - nf}")
    print("FP64 pipe:", ", ".join(f"{k} {v}" for k, v in hist.most_common() if is_fp64(k)))
    print("other    :", ", ".join(f"{k} {v}" for k, v in hist.most_common() if not is_fp64(k)))
    print(f"issue slots if an FP64 instruction holds the port 2 cycles: {2 * nf + n - nf}; FP64 share {2 * nf / (2 * nf + n - nf):.3f}")
    if a.dump:
        with open(a.dump, "w") as f:
            f.write(f"# attempt span of trace_kernel<{a.kernel}> ({n} instructions, {nf} FP64-pipe); scripts/sass_hist.py\n")
            for ad, tx in ins[s:e]:
                f.write(f"{ad:06x}  {tx}\n")


if __name__ == "__main__":
    sys.exit(main())
