python tools/rec_trace.py > gpurun_out/rec_trace_strict.txt 2>&1
python - > gpurun_out/rec_trace_fast.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
from r2d2_b200 import _lib
_lib.lib().r2d2_set_fast_math(1)
exec(open("tools/rec_trace.py").read())
PY
cat gpurun_out/rec_trace_strict.txt gpurun_out/rec_trace_fast.txt
