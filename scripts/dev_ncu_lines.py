"""Join an ncu SASS source page (CSV) with nvdisasm line info of the same build: stall samples per source line.
usage: python scripts/dev_ncu_lines.py gpurun_out/prof.ncu-rep <kernel mangled-name substring> [top N]   (CPU; needs ncu, cuobjdump, nvdisasm)"""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'da4ml_b200', '_binary', 'libda4ml_b200_cmvm.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel with their (file, line, inlined-at chain)
lines = []
inside = False
cur = ('?', 0)
for ln in dis:
    if ln.startswith('.text.'):
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        lines.append((cur, m.group(2).strip()))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
cols = rows[hdr]
ci = {c: i for i, c in enumerate(cols)}
body = rows[hdr + 1:]
print(f'ncu instructions: {len(body)}, nvdisasm instructions: {len(lines)}')
agg = collections.Counter()
agg_inst = collections.Counter()
tot = 0
for k, r in enumerate(body):
    if k >= len(lines):
        break
    s = int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
    agg[lines[k][0]] += s
    agg_inst[lines[k][0]] += int(r[ci['Instructions Executed']] or 0)
    tot += s
print('total samples', tot)
for (f, l), s in agg.most_common(top):
    print(f'{100*s/tot:5.1f}%  {s:8d}  inst {agg_inst[(f,l)]:12d}  {f}:{l}')
