#!/bin/bash
# cProfile of the README-config TDVP run (host-side cost of the launch-latency-bound path)
python -m cProfile -o /tmp/prof tools/readme_bench.py --steps 30 --cpu 0 > /dev/null
python - <<'PY'
import pstats, io
for key, n in (("cumulative", 50), ("tottime", 30)):
    s = io.StringIO(); p = pstats.Stats("/tmp/prof", stream=s); p.sort_stats(key).print_stats(n)
    print(s.getvalue()[:10000])
PY
