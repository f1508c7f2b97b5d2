"""Per-kernel counts of the SASS mnemonics that prove the Blackwell data paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM (tcgen05.ld), UTMALDG (TMA tensor loads, .GATHER4 included), UBLKCP (cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier).
   python tools/sass_summary.py efficient-gnns_b200/libb200gnn.so > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG(?:\.[0-9A-Z.]+)?|UTMASTG|UBLKCP(?:\.[A-Z.]+)?|LDGSTS(?:\.[A-Z0-9.]+)?|SYNCS\.[A-Z0-9.]+|UTCBAR|HMMA)")


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    fn, idx = None, -1
    counts = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            idx += 1
            fn = re.sub(r"\(.*", "", names[idx]) if idx < len(names) else m.group(1)
            counts.setdefault(fn, collections.Counter())
            continue
        if fn is None:
            continue
        for op in PAT.findall(line.split("/*")[1] if line.count("/*") >= 2 else line):
            counts[fn][op] += 1
    print(f"# {path}: SASS mnemonics per kernel (cuobjdump -sass, CUDA 12.9, sm_100a)")
    for fn, c in counts.items():
        if c:
            print(f"{fn}\n    " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main(sys.argv[1])
