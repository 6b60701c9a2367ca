set -e
for v in ""; do
  export PSB_NVCC_EXTRA="$v"
  if [ -n "$v" ]; then export PSB_ABI_LIMITS=4,4,16; else unset PSB_ABI_LIMITS; fi
  python -c "from pysages_b200 import build; build.build(force=True)" > /dev/null 2>&1
  echo "variant [$v]"
  python - <<'PY'
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "2000", "--no-cpu", "--no-e2e", "--no-multi"], capture_output=True, text=True)
d = json.loads(out.stdout.strip().splitlines()[-1])
print("headline us", d["launch"]["pipelined_us_per_step"], "plain", d["launch"]["plain_launch_us_per_step"])
for r in d["series"]:
    print("  %-80s %8.3f" % (r["workload"][:80], r["us_per_step"]))
PY
done
