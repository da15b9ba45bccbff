export LAYERS=6 REPS=8
for st in 0 1; do for qd in 85 340; do echo "stash=$st quads per cluster=$qd"; STASH=$st QUADS=$qd timeout 100 python tools/trace_stack.py | grep "SM clock\|CTA pairs\|^kernel"; done; done
