export PYTHONUNBUFFERED=1
P="timeout 120 python tools/ring_probe.py"
$P hbm
for what in stack dgrad; do
  $P $what
  PAPR_DBG_STACK_RING=64 $P $what
  PAPR_DBG_STACK_GRID=74 $P $what
  PAPR_DBG_STACK_RING=64 PAPR_DBG_STACK_GRID=74 $P $what
  PAPR_DBG_STACK_RING=64 PAPR_DBG_STACK_GRID=64 $P $what
done
$P wgrad
PAPR_DBG_WGRAD_GRID=74 $P wgrad
PAPR_DBG_WGRAD_GRID=84 $P wgrad
