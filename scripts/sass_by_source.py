"""Static instruction budget of a trace kernel by source region (needs -lineinfo, which build() always passes).

    python scripts/sass_by_source.py [--kernel 4,1,0,0] [--by callee|line]

Extracts the sm_100a cubin from the built library, disassembles it with inline line info (`nvdisasm -gi`) and
attributes every instruction to (a) the line of the kernel body in trace_kernel.cuh that (transitively) produced it
and (b) the outermost inlined callee, then prints FP64-pipe / other instruction counts per region.  This is how the
service path (finish + init) was sized against the attempt loop before spending GPU time.
"""
import argparse
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?) ;")
LOC = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')


def disassemble(lib):
    tmp = tempfile.mkdtemp(prefix="bhg_sass_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    return subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "lib", "libbhgeo.so"))
    ap.add_argument("--kernel", default="4,1,0,0,1,0")
    ap.add_argument("--by", default="callee", choices=["callee", "line"])
    a = ap.parse_args()
    t = [int(v) for v in a.kernel.split(",")]
    name = "_ZN3bhg12trace_kernelILi%dELi%dELb%dELb%dELb%dELb%dEEEvNS_9TraceArgsE" % tuple(t)
    txt = disassemble(a.lib)
    start = txt.index(".text." + name + ":")
    end = txt.find("//--------------------- .text.", start)
    body = txt[start:end if end > 0 else None].splitlines()
    # source text for naming regions
    src = {}
    for f in ("trace_kernel.cuh", "geodesic_core.cuh"):
        with open(os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "csrc", f)) as fh:
            src[f] = fh.read().splitlines()
    chain = []          # current chain of (file, line), innermost first
    pending = []
    counts = collections.defaultdict(lambda: [0, 0])
    for ln in body:
        m = LOC.search(ln)
        if m:
            if not pending or pending[-1][1] is None:
                pending = []
            pending.append(((os.path.basename(m.group(1)), int(m.group(2))),
                            (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None))
            continue
        mi = INS.match(ln)
        if not mi:
            continue
        if pending:
            chain = [pending[0][0]] + [p[1] for p in pending if p[1]]
            pending = []
        op = re.sub(r"^@!?U?P\d+\s+", "", mi.group(2)).split()[0]
        fp = op.startswith(FP64)
        # kernel-body frame = last frame in trace_kernel.cuh with line >= 322 (the kernel)
        body_line, callee = None, None
        for i in range(len(chain) - 1, -1, -1):
            f, l = chain[i]
            if f == "trace_kernel.cuh" and l >= 322:
                body_line = l
                callee = chain[i - 1] if i > 0 else None
                break
        if body_line is None:
            key = "(library / no line)"
        elif a.by == "line":
            key = "L%d %s" % (body_line, src["trace_kernel.cuh"][body_line - 1].strip()[:70])
        else:
            region = ("loop head" if body_line < 364 else "finish" if body_line < 416 else
                      "refill+init" if body_line < 452 else "attempt")
            if callee:
                cs = src.get(callee[0], [""] * (callee[1] + 1))[callee[1] - 1].strip()[:60]
                key = "%-11s L%d -> %s:%d %s" % (region, body_line, callee[0][:4], callee[1], cs)
            else:
                key = "%-11s L%d %s" % (region, body_line, src["trace_kernel.cuh"][body_line - 1].strip()[:60])
        counts[key][0 if fp else 1] += 1
    tot = [sum(v[0] for v in counts.values()), sum(v[1] for v in counts.values())]
    print(f"kernel <{a.kernel}>: FP64-pipe {tot[0]}, other {tot[1]}")
    agg = collections.defaultdict(lambda: [0, 0])
    for k, v in sorted(counts.items()):
        print(f"{v[0]:5d} {v[1]:5d}  {k}")
        r = k.split()[0]
        agg[r][0] += v[0]
        agg[r][1] += v[1]
    print("--- per region (FP64, other, slots = 2*FP64 + other)")
    for r, v in agg.items():
        print(f"{r:14s} {v[0]:5d} {v[1]:5d} {2 * v[0] + v[1]:6d}")


if __name__ == "__main__":
    sys.exit(main())
