import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line); r=d["roofline"]
    print(round(d["value"],1),"steps/s  col",round(r["col_pass_ms"],4),"row",round(r["row_pass_ms"],4),"whole_frac",round(r["whole_step_frac"],3),"e2e",round(d["e2e"]["value"],1),"energy",round(d["energy_tracking"]["value"],1))
