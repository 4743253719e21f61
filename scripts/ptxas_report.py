"""Registers / spills / shared memory of every kernel instantiation, from the ptxas -v logs the Makefile
keeps (pymuscle_b200/csrc/build/*.ptxas.log).  `python scripts/ptxas_report.py [--spills]`."""
import glob, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
only_spills = "--spills" in sys.argv
rows = []
for path in sorted(glob.glob(os.path.join(root, "pymuscle_b200", "csrc", "build", "*.ptxas.log"))):
    name = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack, st, ld = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            rows.append((os.path.basename(path)[:-10], name, int(m.group(1)), stack, st, ld))
            name = None
names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.split("\n")
for (tu, _, regs, stack, st, ld), nm in zip(rows, names):
    if only_spills and not (st or ld):
        continue
    nm = re.sub(r"^void mb::|\(mb::\w+(<\w+>)?( const)?\)$|\(int, double\*\)$", "", nm)
    print(f"{tu:22s} {regs:4d} regs  stack {stack:4d}  spill st/ld {st:4d}/{ld:4d}  {nm}")
print(f"{len(rows)} kernels, {sum(1 for r in rows if r[4] or r[5])} with spills")
