"""Turns an `ncu --page raw --csv` export (gpurun_out/*_raw.csv) into the committed summary under
profiles/ and the per-launch DRAM traffic table bench.py reads (profiles/traffic.json).
Usage: python tools/ncu_summary.py <raw.csv> <kernel substring> <batch> <out.txt> [note]"""
import csv
import json
import os
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % (occupancy achieved)'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 data-pipe LSU wavefronts % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts.avg', 'L1 LSU wavefronts per SM'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'of which shared-memory wavefronts (all SMs)'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 sector hit rate %'),
    ('lts__t_sectors.sum.pct_of_peak_sustained_elapsed', 'L2 sectors % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard (warps per issue)'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
    ('smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'stall sleeping'),
    ('launch__registers_per_thread', 'registers per thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic shared memory per block'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
]


def main():
    raw, kernel, batch, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    note = sys.argv[5] if len(sys.argv) > 5 else ''
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    found = [r for r in rows[2:] if kernel in r[hdr.index('Kernel Name')]]
    if not found:
        raise SystemExit('kernel %r not in %s' % (kernel, raw))
    lines = ['# ncu --set full --clock-control none, %s' % os.path.basename(raw)]
    if note:
        lines.append('# ' + note)
    dram = None
    for r in found:
        col = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        lines.append('kernel: ' + col['Kernel Name'][0][:160])
        for key, label in KEYS:
            if key in col:
                lines.append('  %-46s %s %s' % (label, col[key][0], col[key][1]))
        to_bytes = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        rd = float(col['dram__bytes_read.sum'][0]) * to_bytes[col['dram__bytes_read.sum'][1]]
        wr = float(col['dram__bytes_write.sum'][0]) * to_bytes[col['dram__bytes_write.sum'][1]]
        dram = rd + wr
        lines.append('  %-46s %.1f MB' % ('DRAM traffic per launch (read + write)', dram / 1e6))
    open(out, 'w').write('\n'.join(lines) + '\n')
    tj = os.path.join(os.path.dirname(out), 'traffic.json')
    table = json.load(open(tj)) if os.path.isfile(tj) else []
    table = [t for t in table if not (t['kernel'] == kernel and t['batch'] == batch)]
    table.append({'kernel': kernel, 'batch': batch, 'dram_bytes': dram,
                  'source': os.path.basename(out)})
    json.dump(table, open(tj, 'w'), indent=1)
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
