"""Development aid: registers / spills per kernel from the `QC_PTXAS_V=1 csrc/build.sh` logs.   python tests/tools/ptxas_table.py [filter]"""
import glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(ROOT, "deepreinforcementlearningcontrolofquantumcartpoles_b200", "csrc", "build", "*.log"))):
    txt = open(log).read()
    for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("qc::", "").replace("(qc::StepParams)", "").replace("void ", "")
        if flt in name: print("%-60s regs %3s  stack %4s  spill st/ld %4s/%4s" % (name, m.group(5), m.group(2), m.group(3), m.group(4)))
