"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.
    python tools/launch_list.py profiles/r02_launches_bench_c2.csv > profiles/r02_launch_list_bench_c2.md"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    per = collections.OrderedDict()
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        per.setdefault(r[ki], []).append(v)
        seq.append((r[ki], v))
    total = sum(sum(v) for v in per.values())
    ours = sum(sum(v) for k, v in per.items() if "tip::" in k or k.startswith("void dsa_") or "dsa_pack" in k)
    print("| launches | mean us | max us | share | kernel |\n|---:|---:|---:|---:|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {len(v)} | {sum(v) / len(v):.1f} | {max(v):.1f} | {100 * sum(v) / total:.1f}% | `{k[:100]}` |")
    print(f"\nlibb200tip kernels: {100 * ours / total:.1f}% of the summed kernel time ({ours:.0f} of {total:.0f} us).")
    # timed steps = the launches between two L2-flush fills; graph replays carry no exhaustive-scan launch
    steps, cur = [], None
    for k, v in seq:
        if "FillFunctor<unsigned char>" in k:
            if cur:
                steps.append(cur)
            cur = []
        elif cur is not None and "tip::" in k:
            cur.append((k, v))
    if cur:
        steps.append(cur)
    replay = [st for st in steps if st and not any("rerank_scan" in k for k, _ in st)]
    eager = [st for st in steps if any("rerank_scan" in k for k, _ in st)]
    for title, group in (("One CUDA-graph replay step (serialised and cold under ncu)", replay), ("One eagerly launched step (the profile pass)", eager)):
        if group:
            st = group[-1]
            print(f"\n{title}:\n")
            for k, v in st:
                print(f"* {v:7.1f} us  `{k.split('(')[0][-60:]}`")
            print(f"* {sum(v for _, v in st):7.1f} us  sum of {len(st)} launches")

if __name__ == "__main__":
    main(sys.argv[1])
