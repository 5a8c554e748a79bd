#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel family.
usage: python tools/summarize_launches.py launches.csv[.gz] [top]"""
import collections
import csv
import gzip
import re
import sys


def family(name):
    n = name
    own = re.search(r"((?:msda|tfb200)::(?:\w+::)*\w+)", n) or re.search(
        r"<unnamed>::((?:add_dropout_ln|colsum|relu_dropout|sampling_prep|frozen_bn_act|flat_adamw|lsa|match_cost|set_loss|"
        r"refine_boxes|detect_postprocess|finish)\w*)", n)
    if own and "native::<unnamed>" not in n:
        return "OWN " + own.group(1)
    anon = re.match(r"(?:void )?<unnamed>::(\w+)", n)           # kernels of libmsda_b200.so in anonymous namespaces
    if anon:
        return "OWN " + anon.group(1)
    for key, fam in (("cutlass_80", "LIB cublas sm80 mma.sync gemm (cutlass_80 s1688)"), ("cutlass3x", "LIB cutlass3x sm100 gemm/conv"),
                     ("cutlass", "LIB cutlass other"), ("cudnn", "LIB cudnn"), ("xmma", "LIB cudnn xmma"),
                     ("fmha", "LIB fused attention (sdpa, sm80 kernel)"), ("cublas", "LIB cublas misc"),
                     ("nchwToNhwc", "LIB cudnn layout"), ("convolve", "LIB cudnn conv"), ("nccl", "LIB nccl")):
        if key in n:
            return fam
    m = re.search(r"(?:native|at)::(?:native::)?(?:<unnamed>::)?(\w+)(?:<[^>]*?(\w+Functor|sum_functor|FillFunctor|direct_copy|Copy|neg_kernel|clamp|threshold|random)\w*)?", n)
    if m:
        return "ATEN " + m.group(1) + (":" + m.group(2) if m.group(2) else "")
    return "OTHER " + re.sub(r"[<(].*", "", n)[:50]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        fam = family(r[ki])
        tot[fam] += v
        cnt[fam] += 1
    total = sum(tot.values())
    print(f"{sum(cnt.values())} launches, {total / 1e3:.3f} ms of kernel time (serialised, cold-cache: compare SHARES)")
    groups = collections.Counter()
    gcnt = collections.Counter()
    for fam, v in tot.items():
        groups[fam.split()[0]] += v
        gcnt[fam.split()[0]] += cnt[fam]
    for gname, v in groups.most_common():
        print(f"  {gname:6s} {v / 1e3:8.3f} ms {100 * v / total:5.1f} %  {gcnt[gname]:5d} launches")
    for fam, v in tot.most_common(top):
        print(f"{v:10.1f} us {100 * v / total:5.1f} % n={cnt[fam]:5d}  {fam}")


if __name__ == "__main__":
    main()
