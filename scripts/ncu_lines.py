"""Attribute ncu warp-stall samples of a kernel to source lines (the source page of a CSV export only
shows the file of the __global__ function; the per-agent solver is inlined from other headers).

usage: python scripts/ncu_lines.py report.ncu-rep lib.so 'qp_kernelILi4ELi15' [launch_index] [top]
Joins `ncu --page source --print-source sass` (samples per SASS instruction) with
`nvdisasm --print-line-info` of the cubin inside lib.so by instruction order.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    # one cubin per translation unit: take the one that defines the kernel
    dis = ""
    for f in sorted(os.listdir(tmp)):
        d = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if any(ln.startswith(".text.") and kern in ln for ln in d.splitlines()):
            dis = d
            break
    lines, cur, on = [], ("?", 0), False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks = raw.split('"Kernel Name"')[1:]
    # launches of OTHER kernels are in the report too: keep the blocks whose instruction count matches
    parsed = [list(csv.reader(io.StringIO('"Kernel Name"' + b))) for b in blocks]
    same = [r for r in parsed if len(r) - 2 == len(lines)]
    rows = same[launch] if same else parsed[launch]
    hdr = rows[1]
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = rows[2:]
    if len(body) != len(lines):
        print(f"warning: {len(body)} sampled instructions vs {len(lines)} disassembled", file=sys.stderr)
    by_line = collections.Counter()
    ex_line = collections.Counter()
    tot = 0
    for (off, loc, txt), r in zip(lines, body):
        s = int(r[i_s] or 0)
        by_line[loc] += s
        ex_line[loc] += int(r[i_e] or 0)
        tot += s
    print(f"kernel {kern}, launch {launch}: {tot} samples, {len(lines)} instructions")
    by_file = collections.Counter()
    for (f, l), s in by_line.items():
        by_file[f] += s
    for f, s in by_file.most_common():
        print(f"  {f:24s} {100 * s / tot:5.1f} %")
    print("top lines:")
    for (f, l), s in by_line.most_common(top):
        print(f"  {f}:{l:<5d} {100 * s / tot:5.2f} %   inst {ex_line[(f, l)]}")


if __name__ == "__main__":
    main()
