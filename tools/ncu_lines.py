"""Dev tool: attribute ncu warp-stall samples to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object-or-cubin with -lineinfo> <mangled function name>"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, kre, obj, fun = sys.argv[1:5]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kre}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, iss, isrc = hdr.index('Address'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
samples, seen = [], set()
for r in rows[2:]:
    if len(r) <= iss or r[ia] in seen:
        continue
    seen.add(r[ia])
    try:
        samples.append((int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia]), int(r[iss]), r[isrc]))
    except ValueError:
        pass
base = min(a for a, _, _ in samples)
tmp = tempfile.mkdtemp()
if obj.endswith('.cubin'):
    cubin = obj
else:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith('.cubin')][0])
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, None, False
for ln in dis:
    if ln.startswith('.text.'):
        inside = (ln.strip() == f'.text.{fun}:')
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.Counter()
tot = 0
for a, s, src in samples:
    agg[line_of.get(a - base, ('?', 0))] += s
    tot += s
print('total samples', tot, 'instructions', len(samples), 'mapped', sum(1 for a, _, _ in samples if (a - base) in line_of))
srcs = {}
for (f, l), s in agg.most_common(45):
    text = ''
    path = os.path.join('/root/repo/spinor_gpe_b200/csrc', f)
    if os.path.exists(path):
        if path not in srcs:
            srcs[path] = open(path).read().splitlines()
        if 0 < l <= len(srcs[path]):
            text = srcs[path][l - 1].strip()[:90]
    print(f'{100 * s / tot:5.1f}%  {f}:{l}  {text}')
