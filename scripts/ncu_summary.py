"""Condense one `ncu --set full` capture of the advance kernel into the summary CSV committed under profiles/ and the
DRAM-traffic record bench.py reads (profiles/roofline_traffic.json).
usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME_ncu_summary.csv [rows_of_the_launch]"""
import csv, io, json, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__icc_request_hit_rate.pct',
        'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def main(rep, out, rows=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = r[0], r[1], r[2]
    get = dict(zip(hdr, zip(vals, units)))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name", get["Kernel Name"][0], ""])
        for h in hdr:
            if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h):
                w.writerow([h, get[h][0], get[h][1]])
    def gb(name):
        v, u = get[name]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    print(f"time {get['gpu__time_duration.sum']}, dram read {rd/1e9:.4f} GB write {wr/1e9:.4f} GB")
    if rows:
        print(json.dumps({"dram_bytes_per_row": round((rd + wr) / rows, 1)}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
