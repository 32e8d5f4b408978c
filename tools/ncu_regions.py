"""Summarise an ncu report of nerf_mlp_tc_kernel: key metrics + stall samples per kernel region."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            print(f"{k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = rows[2:]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
# region boundaries: mbarrier try-wait instructions with many samples and LDTM / UTCHMMA markers
print("total samples", tot, "instructions", len(data))
marks = [i for i, r in enumerate(data) if any(t in r[isrc] for t in ("LDTM", "UTCHMMA", "UBLKCP", "SYNCS.PHASECHK", "MUFU.SIN", "USETMAXREG", "STG"))]
last = None
for i in marks:
    op = data[i][isrc].split()[0] if not data[i][isrc].strip().startswith("@") else data[i][isrc].split()[1]
    print(i, op, data[i][iS], data[i][iI])
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:25]
print("--- hottest")
for i in sorted(top):
    print(i, data[i][isrc][:90], data[i][iS], data[i][iI])
if len(sys.argv) > 2:
    bounds = [int(x) for x in sys.argv[2].split(",")]
    for a, b in zip(bounds[:-1], bounds[1:]):
        s = sum(int(r[iS]) for r in data[a:b]); ins = sum(int(r[iI]) for r in data[a:b])
        print(f"region [{a},{b}) samples {s} ({100*s/tot:.1f}%) warp-instr {ins/1e6:.1f}M")
