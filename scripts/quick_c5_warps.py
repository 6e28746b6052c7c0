import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for w in sys.argv[1:] or ["16", "15", "14", "13", "12"]:
    out = subprocess.run([sys.executable, os.path.join(here, "quick_c5.py"), "16384", "200", "3"], env=dict(os.environ, THRL_PWL_WARPS=w), capture_output=True, text=True)
    print("warps", w, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
