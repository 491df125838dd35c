import sys; sys.path.insert(0, "/root/repo")
import subprocess, json
for fuse in ("1", "0"):
    code = f"import viscy_b200.functional as F; F.FUSE_GRN_SUMSQ = bool({fuse}); import runpy, sys; sys.argv=['bench.py','--no-secondary','--no-cpu-baseline']; runpy.run_path('/root/repo/bench.py', run_name='__main__')"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out); print("FUSE_GRN_SUMSQ", fuse, d["value"], d["ms_per_step"], d["roofline"]["per_launch_ms"]["fc1+gelu"])
