"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
W = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
     ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak (elapsed)"),
     ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active % (while SM active)"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
     ("sm__ops_path_tensor_src_fp64.sum.per_second", "fp64 tensor ops/ns (x2 = GFLOP/s)"),
     ("sm__ops_path_tensor_src_int8.sum.per_second", "int8 tensor ops/ns"),
     ("sm__inst_executed_pipe_tc.sum", "tcgen05 (UTC*) instructions"),
     ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
     ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe %"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
     ("smsp__sass_inst_executed_op_shared_ld.sum", "LDS instructions"),
     ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %")]
for d in data:
    print("=" * 100)
    print(d[idx["Kernel Name"]][:160])
    for key, label in W:
        if key in idx:
            print(f"  {label:48s} {d[idx[key]]:>20s} {units[idx[key]]}")
