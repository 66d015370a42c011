"""Join the GEMM shape log of one step (bench.py --ncu --graph 0 with LD_GEMM_LOG=file) with the ncu launch list of the same run
(gpu__time_duration per launch, in launch order) and aggregate the GEMM kernel time by shape.

    python tools/gemm_shapes_join.py gpurun_out/r2_gemm_log.json gpurun_out/r2_step_launches_ncu.csv > profiles/r2_gemm_time_by_shape.txt"""
import collections
import csv
import json
import sys

log = json.load(open(sys.argv[1]))
lines = [l for l in open(sys.argv[2]) if not l.startswith("==")]
durs = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    if "gemm_bf16" not in name:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    durs.append((v / 1000.0 if row["Metric Unit"] in ("ns", "nsecond") else v, "2sm" if "2sm" in name else "1cta", row.get("Grid Size", "")))
print("# GEMM launches in the shape log: %d, in the ncu list: %d" % (len(log), len(durs)))
n = min(len(log), len(durs))
agg = collections.OrderedDict()
for e, (us, kern, grid) in zip(log[:n], durs[:n]):
    M, N, K, nb, a_mn, b_mn, sk, caller, act, dd, res, conv = e
    key = (kern, M, N, K, nb, a_mn, b_mn, sk, conv or 0)
    a = agg.setdefault(key, [0, 0.0, caller, grid])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# total GEMM kernel time %.2f ms (serialised, cold); rows sorted by time; flops = 2 M N K x batches" % (tot / 1000))
print("%-5s %7s %6s %6s %5s %4s %3s %4s %6s %9s %7s %8s  %s" % ("kern", "M", "N", "K", "batch", "mn", "sk", "conv", "count", "total us", "us/each", "TFLOP/s", "caller"))
for key, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    kern, M, N, K, nb, a_mn, b_mn, sk, conv = key
    fl = 2.0 * M * N * K * nb
    print("%-5s %7d %6d %6d %5d %2d%2d %3d %4d %6d %9.1f %7.1f %8.1f  %s" % (kern, M, N, K, nb, a_mn, b_mn, sk, conv, a[0], a[1], a[1] / a[0],
                                                                      fl * a[0] / (a[1] * 1e-6) / 1e12, a[2][:70]))
