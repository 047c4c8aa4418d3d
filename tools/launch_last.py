"""Per-kernel times of the LAST MSM in an `ncu --csv --metrics gpu__time_duration.sum` launch list (starts at k_msm_digits)."""
import csv, re, sys
def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        seq.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("zkmsm::", "").replace("void ", ""), v))
    return seq
for path in sys.argv[1:]:
    seq = load(path)
    idx = [i for i, (n, _) in enumerate(seq) if "k_msm_digits" in n]
    last = seq[idx[-1]:] if idx else seq
    keys = ("k_ba_forward", "k_ba_invert", "k_ba_backward", "k_accumulate")
    print(path, "total %.0f us" % sum(v for _, v in last))
    print("   " + "  ".join("%s=%.0f" % (n[:28], v) for n, v in last if any(k in n for k in keys)))
