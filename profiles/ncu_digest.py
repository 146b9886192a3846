"""Text digest of an .ncu-rep (ncu -i ... --page raw --csv): one block per profiled launch with the counters the
DESIGN/README tables quote.   python profiles/ncu_digest.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed.sum", "thread instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
]
STALLS = "smsp__average_warps_issue_stalled_"
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}
print(f"== {sys.argv[1].split('/')[-1]}: ncu --set full --clock-control none (per-launch values; cold caches, serialised launches)")
for r in rows[2:]:
    if len(r) != len(h):
        continue
    print(f"\n-- {r[col['Kernel Name']][:110]}")
    for k, name in WANT:
        if k in col and r[col[k]] != "":
            print(f"   {name:32s} {r[col[k]]:>16s} {u[col[k]]}")
    st = sorted(((float(r[i].replace(',', '')), k[len(STALLS):-len('_per_issue_active.ratio')]) for k, i in col.items()
                 if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")), reverse=True)
    print("   stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
    for k in ("smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed.sum"):
        if k in col and r[col[k]] not in ("", "n/a"):
            v = float(r[col[k]].replace(",", ""))
            lanes = v if k.endswith(".ratio") else v / max(1.0, float(r[col["smsp__inst_executed.sum"]].replace(",", "")))
            print(f"   active lanes per warp instruction {lanes:.1f} / 32")
            break
