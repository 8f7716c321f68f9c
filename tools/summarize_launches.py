"""Summarises an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
tools/profile_frame.py: per kernel of the LAST frame in the capture the launch count, summed duration, share, SM-weighted
share (duration x min(1, CTAs / 148)) and DRAM bytes; prints a markdown table and, with --traffic KEY FILE, records the
frame's DRAM bytes under KEY in a JSON file (profiles/r02_traffic.json: what bench.py reports as roofline.traffic).

    python tools/summarize_launches.py gpurun_out/r02_launches.csv [--traffic c512 profiles/r02_traffic.json]
"""
import collections
import csv
import json
import re
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = recs.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row["Grid Size"]})
        d[row["Metric Name"]] = (float(row["Metric Value"].replace(",", "")), row["Metric Unit"])
    return list(recs.values())


def t_us(d):
    v, u = d["gpu__time_duration.sum"]
    return v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)


def nbytes(d, k):
    v, u = d[k]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    recs = load(sys.argv[1])
    starts = [i for i, d in enumerate(recs) if "yuv420_to_rgb" in d["name"]]
    fr = recs[starts[-1]:]
    for i, d in enumerate(fr):
        if "pack_rgb_yuv420" in d["name"]:
            fr = fr[:i + 1]
            break
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for d in fr:
        n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("vsd::", "")
        g = [int(x) for x in re.findall(r"\d+", d["grid"])]
        share = min(1.0, g[0] * g[1] * g[2] / 148.0)
        a = agg[n]
        a[0] += 1
        a[1] += t_us(d)
        a[2] += t_us(d) * share
        a[3] += nbytes(d, "dram__bytes_read.sum")
        a[4] += nbytes(d, "dram__bytes_write.sum")
    tot = sum(a[1] for a in agg.values())
    tots = sum(a[2] for a in agg.values())
    rd = sum(a[3] for a in agg.values())
    wr = sum(a[4] for a in agg.values())
    print(f"launches in the frame: {len(fr)}; serialised kernel time {tot / 1e3:.2f} ms; DRAM read {rd / 1e9:.2f} GB, written {wr / 1e9:.2f} GB\n")
    print("| kernel | launches | time (us) | share | SM-weighted share | DRAM read (MB) | DRAM written (MB) |")
    print("|---|---|---|---|---|---|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f} % | {100 * a[2] / tots:.1f} % | {a[3] / 1e6:.0f} | {a[4] / 1e6:.0f} |")
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        key, path = sys.argv[i + 1], sys.argv[i + 2]
        try:
            with open(path) as f:
                js = json.load(f)
        except FileNotFoundError:
            js = {}
        js[key] = rd + wr
        js.setdefault("_source", {})[key] = f"{sys.argv[1]}: dram__bytes_read.sum + dram__bytes_write.sum over the {len(fr)} launches of one frame"
        with open(path, "w") as f:
            json.dump(js, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
