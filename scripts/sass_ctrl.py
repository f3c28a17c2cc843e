"""Hot-loop SASS with the scheduling control fields decoded (stall count, yield, barriers) - sm_70+ encoding:
high word bits 41-44 stall, 45 yield, 46-48 write barrier, 49-51 read barrier, 52-57 wait mask, 58-61 reuse.

    python scripts/sass_ctrl.py [--kernel 4,1,0,0,1] [--start 0x4260 --end 0x7900]
"""
import argparse, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "lib", "libbhgeo.so"))
    ap.add_argument("--kernel", default="4,1,0,0,1,0")
    ap.add_argument("--start", default="0")
    ap.add_argument("--end", default="0xffffff")
    a = ap.parse_args()
    t = [int(v) for v in a.kernel.split(",")]
    name = "trace_kernelILi%dELi%dELb%dELb%dELb%dELb%dEEE" % tuple(t)
    txt = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    part = [p for p in txt.split("Function : ")[1:] if name in p.splitlines()[0]][0]
    lines = part.splitlines()
    lo, hi = int(a.start, 16), int(a.end, 16)
    i = 0
    while i < len(lines):
        m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"^\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            ad = int(m.group(1), 16)
            if m2 and lo <= ad <= hi:
                h = int(m2.group(1), 16)
                stall = (h >> 41) & 0xf; yld = (h >> 45) & 1; wb = (h >> 46) & 7; rb = (h >> 49) & 7; wm = (h >> 52) & 0x3f
                print("%05x S%02d %s W%s R%s M%02x  %s" % (ad, stall, "Y" if yld else "-", wb if wb != 7 else "-", rb if rb != 7 else "-", wm, m.group(2)))
            i += 2
        else:
            i += 1

if __name__ == "__main__":
    main()
