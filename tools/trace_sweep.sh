export LAYERS=6 REPS=8
for st in 0 1; do STASH=$st timeout 60 python tools/trace_stack.py | grep -B1 SUMMARY; done
for i in 1 2; do for w in stack dgrad_nocs dgrad; do timeout 100 python tools/ring_probe.py $w | grep -v "^$"; done; done
