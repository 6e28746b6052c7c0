"""Device-side timing of the interval-table kernel under a few settings: `python scripts/quick_pwc.py [runs] [epochs]`."""
import json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "16384"
E = sys.argv[2] if len(sys.argv) > 2 else "40"
here = os.path.dirname(os.path.abspath(__file__))
for env in ({"THRL_PWC_WARPS": "16"}, {"THRL_PWC_WARPS": "20"}, {}):
    for case in ("noisy_aa", "noisy_qr", "cac"):
        out = subprocess.run([sys.executable, os.path.join(here, "quick_mixed.py"), R, E, case], env=dict(os.environ, **env), capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print("%-28s %-9s %-6s %.3e agent-steps/s (%.1f ms)" % (env, case, d["kernel"], d["agent_steps_per_s"], d["ms"]), flush=True)
        except Exception:
            print(env, case, "FAILED", out.stdout[-300:], out.stderr[-300:], flush=True)
