"""Summarise an .ncu-rep here (no GPU needed): key metrics per kernel + top stall instructions.
usage: python tools/ncu_top.py gpurun_out/prof.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
            "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print(" | ".join(f"{hdr[i].split('.')[0][:28]}={r[i][:60]}" for i in idx))


def source(rep, topn):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for k in kernels:
        data = k["rows"]
        tot = sum(int(r[isamp]) for r in data)
        print(f"\n=== {k['name'][:110]}  samples={tot} instrs={len(data)}")
        agg = {}
        for r in data:
            for j in stalls:
                agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j])
        print("  stall totals:", sorted(((v, n) for n, v in agg.items() if v), reverse=True)[:8])
        top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
        for i in sorted(top):
            r = data[i]
            st = sorted(((int(r[j]), hdr[j]) for j in stalls if int(r[j]) > 0), reverse=True)[:2]
            print(f"  {i:5d} {r[isamp]:>5s} {r[iex]:>8s}  {r[isrc].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw_metrics(rep)
    source(rep, int(sys.argv[2]) if len(sys.argv) > 2 else 25)
