"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list of scripts/profile_step.py, joined
with the shape trace it dumps (gpurun_out/trace_step.json).  Usage:
    python scripts/summarize_launches.py profiles/r01_final_launches_step_B4R3.csv profiles/r01_final_trace_step_shapes.json"""
import collections
import csv
import json
import sys


def main(csv_path, trace_path):
    rows = list(csv.reader(open(csv_path)))
    i = [k for k, r in enumerate(rows) if "Kernel Name" in r][0]
    launches = [(r[4].split("(")[0].replace("sg::", "").replace("void ", ""), float(r[-1]) / 1e3) for r in rows[i + 1:]]
    tot = sum(t for _, t in launches)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in launches:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"total {tot / 1e3:.2f} ms over {len(launches)} launches\n")
    print("| share | ms | launches | kernel |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot >= 0.001:
            print(f"| {v[1] / tot * 100:.1f}% | {v[1] / 1e3:.2f} | {v[0]} | `{k[:60]}` |")
    trace = json.load(open(trace_path))
    gq = collections.deque(t for t in trace if t[0] == "gemm")
    aq = collections.deque(t for t in trace if t[0] == "attn")
    gs, ats = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
    for n, t in launches:
        if n.startswith(("gemm_kernel", "gemm_pair_kernel")) and gq:
            _, m, nn, k, segs, mode = gq.popleft()
            gs[(m, nn, k, mode)][0] += 1
            gs[(m, nn, k, mode)][1] += t
        elif n.startswith("attn_kernel") and aq:
            _, lanes, heads, d, nq, nk = aq.popleft()
            ats[(lanes, heads, d, nq, nk)][0] += 1
            ats[(lanes, heads, d, nq, nk)][1] += t
    print("\n| M (pixels) | N (Cout) | K | epilogue | launches | total us | TFLOP/s |\n|---|---|---|---|---|---|---|")
    for (m, nn, k, mode), v in sorted(gs.items(), key=lambda kv: -kv[1][1])[:26]:
        fl = 2.0 * m * nn * k * v[0]
        print(f"| {m} | {nn} | {k} | {('linear', 'geglu', 'nchw')[mode]} | {v[0]} | {v[1]:.0f} | {fl / v[1] / 1e6:.0f} |")
    print("\n| lanes | head dim | Nq | Nk | launches | total us | TFLOP/s (4 Nq Nk C) |\n|---|---|---|---|---|---|---|")
    for (lanes, heads, d, nq, nk), v in sorted(ats.items(), key=lambda kv: -kv[1][1]):
        fl = 4.0 * lanes * nq * nk * heads * d * v[0]
        print(f"| {lanes} | {d} | {nq} | {nk} | {v[0]} | {v[1]:.0f} | {fl / v[1] / 1e6:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
