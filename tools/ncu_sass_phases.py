"""Per-phase view of one kernel of an ncu report: the SASS listing is cut at every barrier / warp sync / branch /
call and the executed instructions, stall samples and opcode classes of each stretch are summed.
usage: python tools/ncu_sass_phases.py report.ncu-rep kernel_regex [min_samples]"""
import csv, io, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
min_smp = int(sys.argv[3]) if len(sys.argv) > 3 else 20
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
col = {k: hdr.index(k) for k in ('Source', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_math', 'stall_barrier',
                                 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_not_selected', 'stall_lg', 'stall_dispatch',
                                 'stall_branch_resolving', 'stall_no_inst')}
keys = ['n', 'ex', 'smp', 'lsb', 'math', 'bar', 'ssb', 'wait', 'mio', 'nsel', 'lg', 'noinst', 'LDG', 'LDS', 'STS', 'STG', 'FP2', 'FP', 'INT', 'SHFL', 'MUFU', 'MMA']
seg = dict.fromkeys(keys, 0)
tot = dict.fromkeys(keys, 0)
print(' '.join(f'{k:>6s}' for k in keys), ' end of stretch')
def flush(tag):
    global seg
    if seg['smp'] >= min_smp: print(' '.join(f'{seg[k]:6d}' if k != 'ex' else f'{seg[k] // 1000:5d}K' for k in keys), ' ', tag)
    for k in keys: tot[k] += seg[k]
    seg = dict.fromkeys(keys, 0)
for n, r in enumerate(rows[hi + 1:]):
    if len(r) <= col['stall_no_inst']: continue
    ins = r[col['Source']].strip()
    if not ins or not (r[col['Instructions Executed']] or '0').isdigit(): continue
    f = lambda k: int(r[col[k]] or 0)
    seg['n'] += 1; seg['ex'] += f('Instructions Executed'); seg['smp'] += f('# Samples')
    for k, c in (('lsb', 'stall_long_sb'), ('math', 'stall_math'), ('bar', 'stall_barrier'), ('ssb', 'stall_short_sb'), ('wait', 'stall_wait'),
                 ('mio', 'stall_mio'), ('nsel', 'stall_not_selected'), ('lg', 'stall_lg'), ('noinst', 'stall_no_inst')): seg[k] += f(c)
    p = ins.split()
    op = (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
    ex = f('Instructions Executed')
    cls = ('LDG' if op in ('LDG', 'LD', 'LDGSTS') else 'LDS' if op in ('LDS', 'LDSM') else 'STS' if op == 'STS' else 'STG' if op in ('STG', 'ST', 'RED', 'ATOMG') else
           'FP2' if op in ('FADD2', 'FMUL2', 'FFMA2') else 'FP' if op in ('FADD', 'FMUL', 'FFMA', 'FMNMX', 'FSEL', 'FSETP', 'DADD', 'DMUL', 'DFMA') else
           'SHFL' if op == 'SHFL' else 'MUFU' if op == 'MUFU' else 'MMA' if op.endswith('MMA') else 'INT')
    seg[cls] += ex // 1000
    if op in ('BAR', 'WARPSYNC', 'CALL', 'BRA', 'EXIT', 'RET', 'BSYNC'): flush(f'{n} {ins[:48]}')
flush('end')
print(' '.join(f'{tot[k]:6d}' if k != 'ex' else f'{tot[k] // 1000:5d}K' for k in keys), '  TOTAL (opcode classes in K warp-instr)')
