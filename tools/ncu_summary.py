"""Summarise an .ncu-rep: key raw metrics + executed-instruction histogram + hottest source lines."""
import sys, csv, subprocess, collections, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print('== kernel:', d.get('Kernel Name', '?')[:90])
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_registers', 'sm__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_elapsed',
            'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']
    for k in keys:
        if k in d:
            print(f'  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}')
    for k, v in d.items():
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
            try:
                if float(v) > 0.2:
                    print(f'  stall {k[34:-28]:30s} {float(v):6.2f}')
            except ValueError:
                pass
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
for nth, hi in enumerate(heads):                     # one block of SASS lines per profiled launch, in the order of the kernels above
    h = rows[hi]
    ci, si, ss = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    end = heads[nth + 1] if nth + 1 < len(heads) else len(rows)
    ops, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[hi + 1:end]:
        if len(r) <= ci or not r[si].strip():
            continue
        toks = r[si].split()
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        try:
            n = int(r[ci] or 0)
        except ValueError:
            continue
        ops[op] += n; tot += n; samp[op] += int(r[ss] or 0); tots += int(r[ss] or 0)
    if not tot:
        continue
    print(f'== launch {nth}: executed warp instructions: {tot}')
    for op, n in ops.most_common(top):
        print(f'  {op:26s} {n:12d} {100 * n / tot:5.1f}%   stall samples {100 * samp[op] / max(tots, 1):5.1f}%')
