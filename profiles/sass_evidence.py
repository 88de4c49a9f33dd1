"""Grep-able SASS / ptxas evidence of the built library (run here, after `make -C nemo-fmi-devel_b200/csrc`):
    python profiles/sass_evidence.py > profiles/r2_sass_evidence.txt
Per kernel: static instruction count, opcode histogram (top 14), the counts of the Blackwell-path mnemonics (UTMALDG = TMA bulk
tensor load, SYNCS = mbarrier, MUFU.RCP64H = the division seed, BAR = __syncthreads) and the ptxas register / spill line."""
import collections
import glob
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nemo-fmi-devel_b200", "csrc")


def main():
    lib = os.path.join(ROOT, "nemo-fmi-devel_b200", "libnemo_fct.so")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, check=True, stdout=subprocess.DEVNULL)
        print("# libnemo_fct.so: sm_100a cubins:", ", ".join(sorted(os.path.basename(f) for f in glob.glob(td + "/*.cubin"))))
        for cubin in sorted(glob.glob(td + "/*.cubin")):
            txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
            cur, ops = None, collections.OrderedDict()
            for line in txt.splitlines():
                m = re.search(r"Function : (\S+)", line)
                if m:
                    cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
                    ops[cur] = collections.Counter()
                    continue
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if m and cur:
                    ops[cur][m.group(1)] += 1
            for name, c in ops.items():
                short = re.sub(r"nemo::\(anonymous namespace\)::", "", name)
                short = re.sub(r"\((nemo::)?FctArgs.*", "(...)", short)[:110]
                base = collections.Counter()
                for k, v in c.items():
                    base[k.split(".")[0]] += v
                tot = sum(c.values())
                key = {k: sum(v for kk, v in c.items() if kk.startswith(k)) for k in ("UTMALDG", "SYNCS", "MUFU.RCP64H", "BAR", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "DSETP", "FSEL")}
                print("%-112s insts %5d | %s | %s" % (short, tot, " ".join("%s=%d" % kv for kv in key.items() if kv[1]), " ".join("%s:%d" % kv for kv in base.most_common(8))))
    print("\n# ptxas -v (registers / spills / stack), from csrc/*.ptxas.log")
    for log in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        name = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"nemo::\(anonymous namespace\)::", "", name)
                name = re.sub(r"\((nemo::)?FctArgs.*", "(...)", name)[:100]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                spill = m.groups()
            m = re.search(r"Used (\d+) registers", line)
            if m and name:
                print("%-102s regs %3s  stack %s B, spill stores %s B, spill loads %s B" % (name, m.group(1), *spill))
                name = None


if __name__ == "__main__":
    main()
