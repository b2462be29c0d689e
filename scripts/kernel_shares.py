"""profiles/r02_kernel_shares.json from an ncu launch list of one denoising step (scripts/profile_step.py): share of the step per
kernel template, the dominant one first.  bench.py reads it to decide which kernel its `roofline` line describes.
Usage: python scripts/kernel_shares.py <launches.csv> <out.json>"""
import collections
import csv
import json
import re
import sys


def main(csv_path, out_path):
    rows = list(csv.reader(open(csv_path)))
    i = [k for k, r in enumerate(rows) if "Kernel Name" in r][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[i + 1:]:
        if len(r) < 5:
            continue
        name = r[4].split("(")[0].replace("sg::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    kernels = [{"kernel": k, "launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    top = kernels[0]["kernel"]
    m = re.match(r"attn_kernel<(\d+), (\d+)", top)
    dominant = f"attn_kernel<{m.group(1)},{m.group(2)}>" if m else top
    json.dump({"source": csv_path.split("/")[-1], "total_us": round(tot, 1), "launches": sum(v[0] for v in agg.values()),
               "dominant": dominant, "kernels": kernels}, open(out_path, "w"), indent=1)
    print(dominant, kernels[0])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
