"""SASS of the per-knot hot loops of the K3 kernels in the built library (no GPU needed).

    python tools/sass_hot_loops.py [kernel-name-substring ...] > profiles/k3_hot_loops_<tag>.sass

For every selected kernel: instruction count, LDL/STL sites in the whole kernel, and every INNERMOST loop (a backward
branch whose body holds no other backward branch) of more than 64 instructions with its static opcode mix and its
local-memory sites; then the listing of those loops.  The loop addresses are the ones `tools/ncu_summary.py` prints
for an `ncu --import-source on` capture of the same build, so sample shares and listings can be put side by side.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("TS_B200_LIB", os.path.join(ROOT, "tortoisesat.jl_b200", "libtortoise_b200.so"))
DEFAULT = ["k3_wide_diag_kernel", "k3_alilqr_diag_kernel", "k3_pair_diag_kernel"]


def functions(sass):
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def innermost_loops(ins, min_len=64):
    back = []
    for a, t in ins:
        if opcode(t) != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            back.append((int(m.group(1), 16), a))
    loops = []
    for lo, hi in back:
        if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in back):
            continue
        body = [(a, t) for a, t in ins if lo <= a <= hi]
        if len(body) > min_len:
            loops.append((lo, hi, body))
    return sorted(loops)


def main():
    want = sys.argv[1:] or DEFAULT
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fns = functions(sass)
    print("SASS of the hot loops of the K3 kernels in %s (cuobjdump -sass, sm_100a; tools/sass_hot_loops.py)" % os.path.relpath(LIB, ROOT))
    listings = []
    for w in want:
        for name, ins in fns.items():
            if w not in name:
                continue
            ldl = sum(1 for _, t in ins if opcode(t) == "LDL")
            stl = sum(1 for _, t in ins if opcode(t) == "STL")
            print("\n%s: %d instructions, LDL %d / STL %d sites in the whole kernel" % (name, len(ins), ldl, stl))
            for lo, hi, body in innermost_loops(ins):
                mix = {}
                for _, t in body:
                    mix[opcode(t)] = mix.get(opcode(t), 0) + 1
                top = ", ".join("%s %d" % kv for kv in sorted(mix.items(), key=lambda kv: -kv[1])[:8])
                loc = mix.get("LDL", 0) + mix.get("STL", 0)
                fp64 = sum(mix.get(o, 0) for o in ("DFMA", "DMUL", "DADD"))
                print("  innermost loop 0x%05x-0x%05x  %4d instructions  FP64 %3d  LDL+STL %d  | %s" % (lo, hi, len(body), fp64, loc, top))
                if fp64 >= 100:
                    listings.append((name, lo, hi, body))
    for name, lo, hi, body in listings:
        print("\n" + "=" * 100)
        print("%s  loop 0x%05x-0x%05x  %d instructions" % (name, lo, hi, len(body)))
        print("=" * 100)
        for a, t in body:
            print("/*%05x*/  %s ;" % (a, t))


if __name__ == "__main__":
    main()
