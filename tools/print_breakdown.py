import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'breakdown_ms' in d and d.get('transcripts','aggregate')=='aggregate':
        print(d.get('probe'), d.get('m'), d.get('proofs'), round(d.get('wall_ms',0),3), {k:round(v,3) for k,v in d['breakdown_ms'].items() if v})
