#!/bin/bash
# Developer aid: A/B several builds of the library on one box.  tools/ab_run.sh TAG lib1.so lib2.so ...
# Per build: smoke (oracle check), the GPU parity file, one configs[1] bench line (10 steps).  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
T=$1; shift
for lib in "$@"; do
  name=$(basename "$lib" .so)
  export DGAN_LIB=$PWD/$lib
  timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_${name}_smoke.log 2>&1
  rc=$?
  echo "smoke rc=$rc" >> gpurun_out/${T}_${name}_smoke.log
  if [ $rc -ne 0 ]; then echo "$name: smoke failed (rc=$rc)"; tail -5 gpurun_out/${T}_${name}_smoke.log; continue; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no_extra --cpu_sample 0 > gpurun_out/${T}_${name}_bench.json 2> gpurun_out/${T}_${name}_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_${name}_bench.json"))
    print("$name: %.0f images/s  e2e %.0f  " % (d["value"], d["e2e"]["value"]) + "  ".join("%s %.1f" % (k["kernel"], k["avg_us"]) for k in d.get("kernels", [])))
except Exception as e:
    print("$name: bench failed", e)
PY
  if [ "${AB_TESTS:-1}" = "1" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_${name}_tests.log
    echo "$name tests: $(tail -1 gpurun_out/${T}_${name}_tests.log)"
  fi
done
