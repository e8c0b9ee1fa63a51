"""Excerpt of the reference's own trajectories2file dump (dmpc/cpp_results/trajectories (200-agents).txt, written
by DMPC::trajectories2file, dmpc/cpp/dmpc.cpp:2088-2126) as a golden vector for the text-format writer:
header line, po, pf (3 lines each) and the position / velocity / acceleration blocks of the first agent.
Run here (the reference is not available on the GPU box):  python tests/golden/make_golden_traj.py"""
import os

SRC = "/root/reference/dmpc/cpp_results/trajectories (200-agents).txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traj200_excerpt.txt")
lines = open(SRC).read().split("\n")
n_cmd = int(lines[0].split()[1])
keep = lines[:7]
for b in range(3):
    s = 7 + 3 * n_cmd * b
    keep += lines[s:s + 3]
open(DST, "w").write("\n".join(keep) + "\n")
print(DST, len(keep), "lines")
