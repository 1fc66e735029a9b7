"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small JSON/markdown for profiles/."""
import csv, io, json, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'dram__cycles_active.avg',
        'launch__shared_mem_per_block_dynamic']


def to_bytes(v, unit):
    v = float(v)
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = {'value': r[i], 'unit': units[i]}
        rd = d.get('dram__bytes_read.sum'); wr = d.get('dram__bytes_write.sum')
        if rd and wr:
            d['dram_bytes_total'] = to_bytes(rd['value'], rd['unit']) + to_bytes(wr['value'], wr['unit'])
        # top warp stall reasons
        st = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_warp_active.pct')]
        st.sort(key=lambda x: -x[1])
        d['top_stalls_pct_of_warp_active'] = {h.split('issue_stalled_')[1].replace('_per_warp_active.pct', ''): v for h, v in st[:6]}
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
