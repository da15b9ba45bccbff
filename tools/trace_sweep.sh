export LAYERS=6 REPS=8
for st in 0 1; do
  for n in 0 1; do STASH=$st PAPR_STACK_NSPLIT=$n timeout 60 python tools/trace_stack.py | grep -B1 SUMMARY; done
done
for n in 0 1; do for w in stack dgrad_nocs; do PAPR_STACK_NSPLIT=$n timeout 100 python tools/ring_probe.py $w; done; done 2>&1 | grep -v "^$"
