#!/usr/bin/env python
"""Condenses bench.py JSON lines on stdin to one short line each (GPU-box log helper)."""
import json
import sys

for l in sys.stdin:
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    ph = d["roofline"]["phase_ms_per_step"]
    print(f"{sys.argv[1] if len(sys.argv) > 1 else '':>10s} {d['value'] / 1e9:7.2f} G/s  {d['ms_per_step']:7.3f} ms/step  "
          f"pd {ph['PushDeposit']:7.3f}  filt {ph['CurrentFiltering']:6.3f}  fs {ph['FieldSolver']:6.3f}  "
          f"comm {ph['Communications']:6.3f}  sort {ph['ParticleSort']:6.3f}  frac {d['roofline']['frac']:.3f}")
