"""SCB solve timings on the default grid (bench.py's scb_metrics), with and without clusters."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for env in ("1", ""):
    if env: os.environ["RSG_SCB_NO_CLUSTER"] = env
    else: os.environ.pop("RSG_SCB_NO_CLUSTER", None)
    m = bench.scb_metrics(0)
    print("cluster" if not env else "one CTA", json.dumps({k: (v if not isinstance(v, dict) else {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}) for k, v in m.items() if k != "grid"}))
