#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of slim_b200/lib/libslim.so, the count of the instructions that show
which hardware paths the kernel uses (TMA bulk copies UBLKCP, cp.async LDGSTS, mbarrier SYNCS, DSMEM MAPA + cluster
stores and generic LD.E loads of peer shared memory, cluster barriers UCGABAR (split: UCGABAR_ARV / UCGABAR_WAIT), fp64 tensor-core DMMA, fp64 FMA, reductions RED / atomics ATOM) plus one sample
line of each.  Runs anywhere (cuobjdump only reads the ELF):   python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "slim_b200" / "lib" / "libslim.so"
KEYS = ["UBLKCP", "LDGSTS", "LDGDEPBAR", "SYNCS", "MAPA", "UCGABAR", "UCGABAR_ARV", "UCGABAR_WAIT", "LD.E", "DMMA", "DFMA", "HMMA", "UTCMMA", "RED", "REDG", "ATOM", "ATOMG", "PRMT",
        "I2F.F64", "SHFL", "BAR.SYNC", "LDG", "LDS", "STS", "ST.E"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    fn, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            body[fn] = []
        elif fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            body[fn].append(re.sub(r"/\*[0-9a-f]{4,}\*/", "", line.split(";")[0]).strip())
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (static), sm_100a")
    want = sys.argv[1:] or ["cd_gram_batch_kernel<slimb200::GaPacked, 16, 8, 2, 512, true>", "cd_gram_batch_kernel<slimb200::GaPacked, 16, 8, 2, 512, false>",
                            "cd_gram_kernel<slimb200::GaPacked, 4>", "cd_gram_kernel<slimb200::GaPacked, 1>",
                            "cd_gram_kernel<slimb200::GaStair, 1", "cd_hybrid_kernel<slimb200::GaStair, false>",
                            "gram_build_kernel<slimb200::GbPacked, false>", "gram_build_kernel<slimb200::GbStair, false>",
                            "cd_cluster_kernel<false, true>",
                            "cd_solve_kernel<128, true, false>", "fslim_neighbors_kernel<false>", "predict_topn_kernel",
                            "place_columns_kernel", "fill_rows_kernel"]
    for fn, ins in body.items():
        if not any(w in fn for w in want):
            continue
        print(f"\n== {fn}   ({len(ins)} instructions)")
        for k in KEYS:
            hits = [i for i in ins if re.search(r"(^|\s)" + re.escape(k) + r"(\.|\s|$)", i)]
            if hits:
                print(f"   {k:10s} {len(hits):5d}    e.g.  {hits[0][:90]}")


if __name__ == "__main__":
    main()
