#!/usr/bin/env python
"""Intervals of a kernel timeline (tools/kernel_trace.py JSON) during which NO tensor-core kernel runs, and what runs
inside them.  CUPTI adds a few microseconds to every kernel of a dependent chain, so chains of tiny kernels look longer
here than inside the un-profiled graph replay: read the list as "which heads / tails of the chains have no tensor-core
work", not as exact durations.  `python tools/trace_gaps.py gpurun_out/ktrace.json`"""
import collections
import json
import sys


def main(path):
    ev = json.load(open(path))
    tc = sorted((s, s + d) for n, s, d, st in ev if ("tapconv_tc" in n or "tapwgrad_tc" in n or "resunit" in n))
    u = []
    for a, b in tc:
        if u and a <= u[-1][1]:
            u[-1][1] = max(u[-1][1], b)
        else:
            u.append([a, b])
    span = max(s + d for n, s, d, st in ev)
    cov = sum(b - a for a, b in u)
    print(f"{len(ev)} kernels, span {span:.0f} us; some tensor-core kernel runs during {cov:.0f} us, none during {span - cov:.0f} us")
    gaps, prev = [], 0.0
    for a, b in u:
        if a > prev:
            gaps.append((prev, a))
        prev = b
    gaps.append((prev, span))
    gaps = [g for g in gaps if g[1] - g[0] > 3]

    def what(g):
        c = collections.Counter()
        for n, s, d, st in ev:
            lo, hi = max(s, g[0]), min(s + d, g[1])
            if hi > lo:
                c[n.split("(")[0].replace("void ", "").replace("artic::", "")[:40]] += hi - lo
        return c

    tot = collections.Counter()
    print("start us   length us   kernels inside (us, summed over concurrent streams)")
    for g in sorted(gaps, key=lambda g: -(g[1] - g[0])):
        c = what(g)
        tot.update(c)
        if g[1] - g[0] > 8:
            print(f"{g[0]:8.0f} {g[1] - g[0]:10.1f}   " + ", ".join(f"{k} {v:.0f}" for k, v in c.most_common(4)))
    print("kernel time inside all gaps:")
    for k, v in tot.most_common(20):
        print(f"{v:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
