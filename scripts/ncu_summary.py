"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/: key metrics, stall mix,
and the SASS lines with the most stall samples.   python scripts/ncu_summary.py rep.ncu-rep out.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
with open(out, "w") as f:
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write(f"== kernel: {d.get('Kernel Name', '?')}\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k:72s} {d[k]} {units[hdr.index(k)]}\n")
        st = sorted(((float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")
                     and "not_issued" not in h and d[h] not in ("", "n/a")), reverse=True)
        f.write("-- warp stall reasons (warps stalled per issue-active cycle)\n")
        for v, h in st[:8]:
            f.write(f"   {v:7.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = src.splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')] + [len(lines)]
    prev, n = None, -1
    for a, b in zip(starts[:-1], starts[1:]):       # source tables of the captured launches (ncu may print each twice)
        if lines[a:b] == prev:
            continue
        prev, n = lines[a:b], n + 1
        rd = [r for r in csv.DictReader(io.StringIO("\n".join(lines[a:b]))) if (r.get("# Samples") or "0").isdigit()]
        tot = sum(int(r["# Samples"] or 0) for r in rd)
        f.write(f"-- launch {n}: top SASS instructions by stall samples (total samples {tot})\n")
        top = sorted(rd, key=lambda r: -int(r["# Samples"] or 0))[:25 if n == 0 else 8]
        for r in top:
            reasons = sorted(((int(r[k] or 0), k) for k in r if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
            f.write(f"   {int(r['# Samples']):6d}  {r['Source'][:70]:70s} {reasons}\n")
        # per-opcode sample share
        agg = {}
        for r in rd:
            op = r["Source"].split()[0] if r["Source"] else "?"
            if op.startswith("@"):
                op = r["Source"].split()[1]
            op = op.split(".")[0]
            agg[op] = agg.get(op, 0) + int(r["# Samples"] or 0)
        f.write(f"-- launch {n}: stall samples by opcode\n")
        for op, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
            f.write(f"   {100.0 * v / max(tot, 1):5.1f}%  {op}\n")
print(open(out).read())
