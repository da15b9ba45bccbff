for i in 1 2 3; do for n in 0 1; do echo -n "bits_ilp=$n "; PAPR_STACK_BITS_ILP=$n timeout 100 python tools/ring_probe.py stack | grep -v "^$"; done; done
