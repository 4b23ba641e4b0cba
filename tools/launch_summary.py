"""Per-kernel totals of ONE score-network forward from an ncu launch list (profiles/*_launches_cfg2_forward.csv).

    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 546 -c 420 --csv --log-file launches.csv \
        python tools/profile_forward.py 64 256 2
    python tools/launch_summary.py launches.csv

A forward is delimited by two consecutive `node_features_kernel` launches (the first kernel of the embedder); times are
cold-cache and serialised, so compare SHARES with bench.py's event-timed step, not absolute values.
"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    data = rows[1:]
    marks = [i for i, r in enumerate(data) if "node_features_kernel" in r[ix["Kernel Name"]]]
    if not marks:
        raise SystemExit("no node_features_kernel launch in the list: --launch-skip too large?")
    fw = data[marks[0]:marks[1]] if len(marks) > 1 else data[marks[0]:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in fw:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("s2s::", "").replace("<unnamed>::", "").replace("void ", "")
        unit = r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        ms = v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"one forward: {len(fw)} launches, {total:.3f} ms (serialised, cold cache)")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:8.3f} ms  x{n:4d}  avg {1e3 * ms / n:8.1f} us  {100 * ms / total:5.1f} %  {k[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
