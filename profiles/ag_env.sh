#!/bin/bash
# is the +0.1 ms per step seen at N >= 2 tied to how the NCCL process group is initialised?  (world = 1, NCCL initialised)
export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29620
run() { echo "== $1 $(env $1 python profiles/ag_probe.py 2>&1 | grep -E "^world" | cut -c1-80)"; }
for i in 1 2 3 4; do run "AG_NONE=1"; run "AG_LAZY=1"; run "AG_LAZY=1 AG_WARM=1"; done
