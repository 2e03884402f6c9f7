"""Summarise an .ncu-rep (ncu --set full) per launch: duration, DRAM bytes, tensor/SM pipe utilisation, registers.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.csv]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in KEYS if k in idx]
    extra = [h for h in hdr if "pipe_tensor" in h and h not in cols][:6]
    cols += extra
    out = [["id", "kernel"] + ["%s [%s]" % (c, units[idx[c]]) for c in cols]]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "")[:60]
        out.append([r[idx["ID"]], name] + [r[idx[c]] for c in cols])
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)

if __name__ == "__main__":
    main()
