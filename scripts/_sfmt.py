import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d["shape"], round(d["own_fwd_ms"],4), round(d["lib_fwd_ms"],4), round(d["own_fwdbwd_ms"],4), round(d["lib_fwdbwd_ms"],4), d["max_abs_diff"])
