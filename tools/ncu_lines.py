"""Join ncu per-SASS-instruction execution counts with source lines (nvdisasm line info).
usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [top] [nth kernel of the report, default 0]"""
import sys, csv, subprocess, io, re, os, tempfile, glob, collections
rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
nth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
hi = his[nth]
rows = rows[:his[nth + 1]] if nth + 1 < len(his) else rows
h = rows[hi]
ci, si, ss = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
prof = [(r[si].strip(), int(r[ci] or 0), int(r[ss] or 0)) for r in rows[hi + 1:] if len(r) > ci and r[0] != 'Address']
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for cubin in glob.glob(os.path.join(tmp, '*.cubin')):
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    if kname not in dis:
        continue
    cur, infn, seq = None, False, []
    for ln in dis.split('\n'):
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            infn = kname in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', ln)
        if m:
            seq.append((cur, m.group(1).strip()))
    if seq:
        lines = seq
        break
assert lines, 'kernel not found'
print(len(prof), 'profiled instrs;', len(lines), 'disassembled')
agg, agg_s = collections.Counter(), collections.Counter()
tot = sum(p[1] for p in prof)
tots = sum(p[2] for p in prof)
for (loc, txt), (ptxt, n, s) in zip(lines, prof):
    agg[loc] += n
    agg_s[loc] += s
srcs = {}
for loc, n in agg.most_common(top):
    if loc is None:
        print(f'{n:11d} {100*n/tot:5.1f}%  stall {100*agg_s[loc]/max(tots,1):5.1f}%  <no line>')
        continue
    f = loc[0]
    if f not in srcs:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(so)), 'csrc', f)) + glob.glob('/usr/local/cuda/include/**/' + f, recursive=True)
        srcs[f] = open(cands[0]).read().split('\n') if cands else []
    text = srcs[f][loc[1] - 1].strip()[:100] if len(srcs[f]) >= loc[1] else ''
    print(f'{n:11d} {100*n/tot:5.1f}%  stall {100*agg_s[loc]/max(tots,1):5.1f}%  {f}:{loc[1]}: {text}')
