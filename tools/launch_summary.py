"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys

def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"^void ", "", name); name = re.sub(r"\(.*$", "", name)[:70]
        v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        out.append((name, v, row["Grid Size"]))
    return out

def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g in rows:
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"total {tot/1e3:.1f} ms over {len(rows)} launches")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print(f"{v/1e3:9.2f} ms {100*v/tot:5.1f}%  n={n:4d}  avg {v/n:8.1f} us  {k}")
    if len(sys.argv) > 3:
        pat = sys.argv[3]
        i0 = next(i for i, r in enumerate(rows) if pat in r[0])
        for n, v, g in rows[i0:i0 + int(sys.argv[4]) if len(sys.argv) > 4 else i0 + 30]:
            print(f"{v:9.1f} us {g:>20}  {n}")

if __name__ == "__main__":
    main()
