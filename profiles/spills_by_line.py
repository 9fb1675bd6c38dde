"""Where do a kernel's register spills sit?  Cross-compiles one .cu for sm_100a (no GPU needed), disassembles it with
line info and counts local-memory loads/stores (LDL/STL) per source line and kernel.

    python profiles/spills_by_line.py fmov_pose_b200/csrc/mlp_fine.cu [-DFMOV_...] [--kernel fine_bwd]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    args = [a for a in sys.argv[1:]]
    kernel = None
    if "--kernel" in args:
        i = args.index("--kernel")
        kernel = args[i + 1]
        del args[i:i + 2]
    src, flags = args[0], args[1:]
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "k.cubin")
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"] + flags +
                       ["-cubin", "-o", cubin, src], check=True, capture_output=True)
        sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], check=True, capture_output=True, text=True).stdout
    fn, cur = None, None
    cnt = collections.Counter()
    for ln in sass.split("\n"):
        m = re.search(r"\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
        m = re.search(r"\b(LDL|STL)(\.\S+)?\s", ln)
        if m and (kernel is None or (fn and kernel in fn)):
            cnt[(fn, cur, m.group(1))] += 1
    tot = collections.Counter()
    for (fn, cur, kind), n in sorted(cnt.items(), key=lambda kv: (kv[0][0] or "", kv[0][1] or ("", 0))):
        print(f"{fn[:48]:48s} {cur[0]}:{cur[1]:<5d} {kind} x{n}")
        tot[(fn, kind)] += n
    for (fn, kind), n in sorted(tot.items()):
        print(f"TOTAL {fn[:60]:60s} {kind} {n}")


if __name__ == "__main__":
    main()
