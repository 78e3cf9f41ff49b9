"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall reasons, loop instruction mix."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('=' * 100)
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print('%-72s %s %s' % (w, r[i], units[i]))
    print('-- stall reasons (warps per issue-active cycle)')
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            st.append((float(r[i] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    for v, h in sorted(st, reverse=True)[:9]:
        print('   %-28s %.3f' % (h, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = -1
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        blocks.append({'name': r[1], 'rows': []}); continue
    if r and r[0] == 'Address':
        blocks[-1]['hdr'] = r; continue
    if blocks and 'hdr' in blocks[-1] and len(r) >= len(blocks[-1]['hdr']):
        blocks[-1]['rows'].append(r)
for b in blocks:
    ci = {h: i for i, h in enumerate(b['hdr'])}
    ex = [int(r[ci['Instructions Executed']] or 0) for r in b['rows']]
    if not ex:
        continue
    mx = sorted(ex)[-max(1, len(ex) // 20)]
    ops = collections.Counter()
    for r, e in zip(b['rows'], ex):
        if e >= 0.4 * mx:
            toks = r[ci['Source']].split()
            op = toks[1] if toks[0].startswith('@') else toks[0]
            ops[op.split('.')[0]] += e / mx
    print('=' * 100)
    print(b['name'][:90])
    print('-- per-iteration instruction mix of the hot loop (executed / ref count %d)' % mx)
    print('   ' + ', '.join('%s %.1f' % (k, v) for k, v in ops.most_common(16)), ' | sum %.1f' % sum(ops.values()))
    print('-- top stall-sample instructions')
    srt = sorted(b['rows'], key=lambda r: -int(r[ci['# Samples']] or 0))
    tot = sum(int(r[ci['# Samples']] or 0) for r in b['rows'])
    for r in srt[:10]:
        print('   %5.1f%%  %s' % (100.0 * int(r[ci['# Samples']] or 0) / max(tot, 1), r[ci['Source']][:90]))
