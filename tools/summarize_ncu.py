"""Summarise an `ncu --set full` report (captured by tools/profile_kernels.py) into one line per profiled launch:
   python tools/summarize_ncu.py gpurun_out/r2l_strict.ncu-rep name1,name2,... > profiles/ncu_full_summary_strict_r02.txt
The names are the profile_kernels.py selection in launch order (the report itself only knows kernel template names)."""
import csv
import io
import json
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu_wavefront_pct"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp_inst")]


def main():
    rep, names = sys.argv[1], sys.argv[2].split(",")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    facts = {}
    for i, r in enumerate(rows[2:]):
        name = names[i] if i < len(names) else "launch%d" % i
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for metric, short in WANT:
            if metric in hdr:
                j = hdr.index(metric)
                v = r[j].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    continue
                u = units[j]
                if short in ("dram_read", "dram_write"):
                    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                if short == "us":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
                d[short] = v
        facts[name] = d
    for n, d in facts.items():
        print("%-10s %7.1f us  tensor %5.1f%%  issue %5.1f%%  lsu %5.1f%%  dram %7.1f MB r / %6.1f MB w (%4.1f%% of peak)  L2 hit %5.1f%%  warps %5.1f%%  regs %3d  grid %5d  %s"
              % (n, d.get("us", 0), d.get("tensor_pipe_pct", 0), d.get("issue_active_pct", 0), d.get("lsu_wavefront_pct", 0),
                 d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, d.get("dram_pct", 0), d.get("l2_hit_pct", 0),
                 d.get("warps_active_pct", 0), int(d.get("regs", 0)), int(d.get("grid", 0)), d["kernel"]))
    if len(sys.argv) > 3:
        json.dump(facts, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
