"""One training step's tensor/recurrent launches from an `ncu --set full` report -> the small JSON bench.py reads for
`roofline.traffic` (per-launch duration, DRAM bytes, tensor-pipe activity).
usage: python tools/ncu_step_json.py gpurun_out/x.ncu-rep profiles/out.json "<how it was captured>" """
import csv, io, json, subprocess, sys

ORDER = ["conv1.fwd", "conv2.fwd", "conv3.fwd", "rnn.fwd", "rnn.bwd", "conv3.wgrad", "conv3.dgrad", "conv2.wgrad",
         "conv2.dgrad", "conv1.wgrad"]


def main():
    rep, out, how = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key, scale_unit=None):
        v = float(r[ix[key]].replace(",", ""))
        u = units[ix[key]]
        if scale_unit == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        if scale_unit == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1.0))
        return v

    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    assert len(data) == len(ORDER), "expected %d launches, got %d" % (len(ORDER), len(data))
    res = {"source": how}
    for name, r in zip(ORDER, data):
        kn = r[ix["Kernel Name"]]
        assert ("wgrad" in kn) == ("wgrad" in name) and ("rnn" in kn) == ("rnn" in name), (name, kn)
        res[name] = {"ms": round(val(r, "gpu__time_duration.sum", "ms"), 4),
                     "dram_read_MB": round(val(r, "dram__bytes_read.sum", "MB"), 1),
                     "dram_write_MB": round(val(r, "dram__bytes_write.sum", "MB"), 1),
                     "tensor_pipe_active_pct": round(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                     "registers": int(val(r, "launch__registers_per_thread"))}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
