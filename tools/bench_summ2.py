"""Dev tool: one-line digest of a bench.py JSON line read from stdin."""
import json
import sys

for raw in sys.stdin:
    raw = raw.strip()
    if not raw.startswith('{'):
        continue
    d = json.loads(raw)
    r = d.get('roofline', {})
    print(json.dumps({'steps_per_s': round(d['value'], 1), 'ms_per_step': round(d['ms_per_step'], 4),
                      'col_ms': round(r.get('col_pass_ms', 0), 4), 'row_ms': round(r.get('row_pass_ms', 0), 4),
                      'e2e': round(d['e2e']['value'], 1), 'energy_steps_per_s': round(d['energy_tracking']['value'], 1),
                      'energy_unwrapped_ms': round(d.get('energy_unwrapped', {}).get('ms', 0), 1),
                      'atoms': d.get('atom_number_check'), 'sm_mhz': d['clocks']['sm_mhz']}))
