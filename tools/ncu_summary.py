#!/usr/bin/env python3
"""Markdown summary of an `ncu --set full` report of tools/profile_conv.py (one column per layer class; the second of the
two launches of every class, i.e. the warm one):   python tools/ncu_summary.py gpurun_out/n18/prof_v18.ncu-rep > profiles/...md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
names = sys.argv[2].split(",") if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
sel = data[1::2]
want = [
    ("gpu__time_duration.sum", "kernel time"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor pipe (UTCHMMA bf16) % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe: LSU wavefronts % of peak"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "shared-memory pipe: tensor-core operand wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory pipe: LSU wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_op_ldgsts.sum", "LDGSTS warp instructions"),
    ("smsp__sass_inst_executed_op_utcmma.sum", "UTCHMMA instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]
kn = ix["Kernel Name"]
print("| metric | unit | " + " | ".join(names[i] if names else "launch %d" % i for i in range(len(sel))) + " |")
print("|---|---|" + "---|" * len(sel))
print("| kernel | | " + " | ".join(r[kn].replace("void tsg::", "").replace("(tsg::TcParams)", "") for r in sel) + " |")
for key, label in want:
    if key not in ix:
        continue
    vals = []
    for r in sel:
        v = r[ix[key]]
        try:
            f = float(v.replace(",", ""))
            v = "%.1f" % f if abs(f) < 1e4 else "%.3g" % f
        except ValueError:
            pass
        vals.append(v)
    print("| %s (`%s`) | %s | %s |" % (label, key, units[ix[key]], " | ".join(vals)))
# derived: shared-memory pipe utilisation = (LSU + tensor-core wavefronts) / (SMs * cycles)
try:
    cyc = [float(r[ix["sm__cycles_elapsed.max"]].replace(",", "")) for r in sel]
    lsu = [float(r[ix["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]].replace(",", "")) for r in sel]
    tc = [float(r[ix["l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]].replace(",", "")) for r in sel]
    print("| shared-memory data pipe busy = (LSU + tensor-core wavefronts) / (148 SMs x cycles) | % | " +
          " | ".join("%.1f" % (100 * (a + b) / (148 * c)) for a, b, c in zip(lsu, tc, cyc)) + " |")
    xb = [float(r[ix["l1tex__m_xbar2l1tex_read_bytes.sum"]].replace(",", "")) for r in sel]
    un = units[ix["l1tex__m_xbar2l1tex_read_bytes.sum"]]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un]
    t = [float(r[ix["gpu__time_duration.sum"]].replace(",", "")) for r in sel]
    tun = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[units[ix["gpu__time_duration.sum"]]]
    print("| L2 -> SM throughput | TB/s | " + " | ".join("%.2f" % (b * mul / (x * tun) / 1e12) for b, x in zip(xb, t)) + " |")
except Exception as e:  # noqa
    print("<!-- derived rows unavailable: %r -->" % (e,))
