"""Static view of a kernel's SASS by source line (no GPU needed):
    python tools/sass_lines.py <cubin or libw2b.so> <kernel name substring> [first_line last_line]
Extracts the cubin when given the .so, disassembles the kernel with nvdisasm -g and prints how many
SASS instructions each source line of word2bits_b200/csrc/*.cuh produced, plus an opcode histogram of the
selected line range (e.g. the consumer row loop).  Instruction counts are static, not execution counts."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def cubin_of(path):
    if path.endswith(".cubin"):
        return path
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=tmp, stdout=subprocess.DEVNULL)
    cands = sorted((os.path.getsize(os.path.join(tmp, f)), os.path.join(tmp, f)) for f in os.listdir(tmp))
    return cands[-1][1]


def symbol_index(cubin, needle):
    out = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
    for line in out.splitlines():
        m = re.match(r"\s*(0x[0-9a-f]+)\s+\S+\s+\S+\s+0x12\s+\S+\s+\S+\s+(\S+)", line)
        if m and needle in m.group(2):
            return m.group(1), m.group(2)
    raise SystemExit("no kernel matching %r" % needle)


def main():
    cubin = cubin_of(sys.argv[1])
    idx, name = symbol_index(cubin, sys.argv[2])
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
    txt = subprocess.run(["nvdisasm", "-c", "-g", "-fun", idx, cubin], capture_output=True, text=True).stdout
    cur = ("?", 0)
    per_line = collections.Counter()
    ops = collections.Counter()
    total = 0
    for line in txt.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        total += 1
        per_line[cur] += 1
        if cur[0] == "w2b_ring.cuh" and lo <= cur[1] <= hi:
            ops[m.group(1).split(".")[0]] += 1
    print(name, "-", total, "SASS instructions")
    for (f, l), n in sorted(per_line.items()):
        if f.startswith("w2b_") and (f != "w2b_ring.cuh" or lo <= l <= hi):
            print("%-18s %5d %5d" % (f, l, n))
    sel = sum(n for (f, l), n in per_line.items() if f == "w2b_ring.cuh" and lo <= l <= hi)
    print("w2b_ring.cuh lines %d-%d: %d instructions; opcodes: %s" % (lo, hi, sel, ops.most_common(25)))


if __name__ == "__main__":
    main()
