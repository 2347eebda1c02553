#!/usr/bin/env python
"""Summaries of ncu outputs (run here, no GPU needed).
  python profiles/summarize.py launches gpurun_out/r01b_launches.csv
  python profiles/summarize.py kernel gpurun_out/r01b_k_chain.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki]
        name = name[:name.index('(')] if '(' in name else name
        a = agg.setdefault(name[:110], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print("per-kernel device time (%s), cold-cache serialised launches under ncu: compare SHARES" % data[0][ui])
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-112s n=%4d total=%12.0f share=%6.3f avg=%10.0f" % (k, c, t, t / tot, t / c))


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(vals, units)))
        print("kernel:", d.get('Kernel Name', ('?',))[0][:150])
        for k in KEEP:
            if k in d:
                print("  %-88s %18s %s" % (k, d[k][0], d[k][1]))


def sass(path, rays=None):
    """Dynamic instruction mix from the source page (ncu -i X --page source --csv [| gzip]): warp instructions
    executed per opcode, their share, and -- given the rays per launch -- thread instructions per ray."""
    import gzip
    op = gzip.open if path.endswith('.gz') else open
    rows = list(csv.reader(op(path, 'rt')))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hi]
    si, ei, ti, wi = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    agg = collections.OrderedDict()
    tot = tt = ts = 0
    for r in rows[hi + 1:]:
        if len(r) <= ti:
            continue
        toks = r[si].split()
        if not toks:
            continue
        o = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        o = o.rstrip(';')
        base = o.split('.')[0]
        key = o if base in ('MUFU', 'IMAD', 'LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL') else base
        key = 'IMAD.MOV' if key.startswith('IMAD.MOV') else ('IMAD' if key.startswith('IMAD') else key)
        e, t, smp = int(r[ei] or 0), int(r[ti] or 0), int(r[wi] or 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += e; a[1] += t; a[2] += smp
        tot += e; tt += t; ts += smp
    print(rows[0][1][:150] if rows and len(rows[0]) > 1 else '')
    print("warp instructions executed: %d; thread instructions: %d%s; stall samples: %d" %
          (tot, tt, ("; per ray: %.1f" % (tt / rays)) if rays else "", ts))
    print("%-14s %14s %7s %10s %8s" % ("opcode", "warp instr", "share", "per ray" if rays else "", "samples"))
    for k, (e, t, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if e == 0 and smp == 0:
            continue
        print("%-14s %14d %7.3f %10s %8.3f" % (k, e, e / max(tot, 1), ("%.2f" % (t / rays)) if rays else "", smp / max(ts, 1)))


if __name__ == '__main__':
    if sys.argv[1] == 'sass':
        sass(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
        sys.exit(0)
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
