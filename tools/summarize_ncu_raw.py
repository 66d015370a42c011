"""Table from an `ncu --set full` report: per kernel launch duration, DRAM bytes and GB/s, tensor-pipe %, SM throughput.

    ncu -i gpurun_out/r1_ops.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/summarize_ncu_raw.py /tmp/raw.csv [HBM_peak_GB/s]
"""
import csv, sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6546.9
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def get(r, name, default=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col.get(name)
    return units[i] if i is not None else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


tensor_names = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"]
print("%-34s %10s %8s %9s %9s %9s %8s %8s %8s %7s %6s" % ("kernel", "grid", "us", "rd MB", "wr MB", "GB/s", "%HBMpk", "tensor%", "SM thr%", "IPC", "regs"))
for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    dur = to_us(get(r, "gpu__time_duration.sum", 0.0), unit("gpu__time_duration.sum"))
    rd = to_bytes(get(r, "dram__bytes_read.sum", 0.0), unit("dram__bytes_read.sum"))
    wr = to_bytes(get(r, "dram__bytes_write.sum", 0.0), unit("dram__bytes_write.sum"))
    gbs = (rd + wr) / (dur * 1e-6) / 1e9 if dur else 0.0
    tens = None
    for tn in tensor_names:
        tens = get(r, tn)
        if tens is not None:
            break
    smthr = get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    ipc = get(r, "sm__inst_executed.avg.per_cycle_active")
    regs = get(r, "launch__registers_per_thread")
    print("%-34s %10s %8.1f %9.2f %9.2f %9.0f %8.1f %8s %8s %7s %6s" % (
        name[:34], r[col["Grid Size"]].replace(" ", ""), dur, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak,
        "-" if tens is None else "%.1f" % tens, "-" if smthr is None else "%.1f" % smthr,
        "-" if ipc is None else "%.2f" % ipc, "-" if regs is None else "%d" % regs))
