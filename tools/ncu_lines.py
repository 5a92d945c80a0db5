"""Rank CUDA source lines of an ncu report by executed instructions / stall samples.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
lines, ops = [], collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= iex: continue
    if r[0].strip().isdigit():
        try: lines.append((int(r[iex]), int(r[ismp]), int(r[0]), r[1].strip()))
        except ValueError: pass
    elif r[2].startswith('0x') and r[iex].isdigit():
        s = r[3].split(); o = s[0] if not s[0].startswith('@') else s[1]
        ops[o.split('.')[0]] += int(r[iex])
tot = sum(l[0] for l in lines) or 1; tots = sum(l[1] for l in lines) or 1
print('total warp-inst (source rows)', tot, 'samples', tots)
for n, s, ln, src in sorted(lines, reverse=True)[:top]:
    print(f'{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% smp  L{ln:<4d} {src[:120]}')
t2 = sum(ops.values()) or 1
print('--- opcodes'); print(', '.join(f'{o} {100*n/t2:.1f}%' for o, n in ops.most_common(22)))
