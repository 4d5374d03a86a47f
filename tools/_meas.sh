python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['full_frame'])); print(d['value'], d['e2e']['value'], d['tracked_frames_per_sec_per_camera'])"
bash tools/_run_tmp.sh > /dev/null 2>&1
python - <<PY
import csv,collections
lines=open("gpurun_out/mtr_list.csv").read().splitlines()
i0=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[i0:]))
d=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    d[r["Kernel Name"].split("(")[0]][r["Metric Name"]].append(float(r["Metric Value"].replace(",","")))
for k,m in d.items():
    print(k, {mn:(round(sum(v)/len(v),1), len(v)) for mn,v in m.items()})
PY
