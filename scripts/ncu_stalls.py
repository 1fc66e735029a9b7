"""Per-instruction stall summary of an .ncu-rep captured with --import-source on (read on the CPU box).
usage: python scripts/ncu_stalls.py REP [min_exec_for_hot_group]"""
import collections, csv, io, subprocess, sys


def main(rep, hot=500000):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, r = rows[0], rows[2]
    for key in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
                'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
                'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
                'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size'):
        if key in h:
            print(key, r[h.index(key)], rows[1][h.index(key)])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}; data = rows[2:]
    keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    grp = collections.defaultdict(collections.Counter); cnt = collections.Counter(); ins = collections.Counter(); op = collections.Counter()
    for d in data:
        ex = int(d[ix['Instructions Executed']] or 0)
        g = 'hot' if ex > hot else 'rest'
        for k in keys:
            grp[g][k] += int(d[ix[k]] or 0)
        cnt[g] += int(d[ix['# Samples']] or 0); ins[g] += ex
        if g == 'hot':
            s = d[ix['Source']].split()
            o = s[1] if s[0].startswith('@') else s[0]
            op[o.split('.')[0]] += ex
    for g in grp:
        print(g, 'samples', cnt[g], 'inst', ins[g], {k.replace('stall_', ''): v for k, v in grp[g].most_common(10)})
    tot = sum(op.values())
    print('hot opcode shares (%):', [(k, round(100.0 * v / tot, 1)) for k, v in op.most_common(22)])
    top = sorted(data, key=lambda d: -int(d[ix['# Samples']] or 0))[:25]
    for d in top:
        n = int(d[ix['# Samples']] or 0)
        st = {k.replace('stall_', ''): int(d[ix[k]] or 0) for k in keys if int(d[ix[k]] or 0) > 0.2 * n}
        print(d[ix['Address']][-5:], str(n).rjust(6), d[ix['Instructions Executed']].rjust(9), d[ix['Source']][:64].ljust(64), st)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 500000)
