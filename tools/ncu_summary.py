#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + top source lines by stall samples.
Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [n_lines]"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_sample_buffer_full', 'sm__cycles_elapsed.max']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                print(f'  {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}')
        stall = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
        if not stall:
            stall = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if 'warp_issue_stalled' in h and h.endswith('.pct')]
        for v, h in sorted(stall, reverse=True)[:8]:
            print(f'  stall {h:66s} {v:.3f}')
    src = run([rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'])
    cur_file, hdr, out, tot = '', None, [], 0.0
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]; continue
        if r[0] == 'Line No':
            hdr = r
            ci = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); te = hdr.index('Avg. Threads Executed')
            sb = {h: k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
            continue
        if hdr is None or not r[0] or not r[0].isdigit():
            continue
        try:
            s = float(r[ci] or 0)
        except ValueError:
            continue
        tot += s
        top = sorted(((float(r[k] or 0), h) for h, k in sb.items()), reverse=True)[:2]
        out.append((s, cur_file, int(r[0]), r[1], r[ie], r[te], ' '.join(f'{h[6:]}={v:.0f}' for v, h in top if v)))
    print(f'-- source lines by stall samples (total {tot:.0f}); columns: samples% file:line inst avg-threads top-stalls | source')
    for s, f, ln, text, ie, te, st in sorted(out, reverse=True)[:nlines]:
        print(f'{100 * s / max(tot, 1):6.2f}% {f}:{ln:<5d} {ie:>10s} {te:>3s} {st:32s}| {text.strip()[:110]}')


if __name__ == '__main__':
    main()
