for cfg in "RXB_SPLIT=1 RXB_ROWLIST=1" "RXB_SPLIT=0 RXB_ROWLIST=1" "RXB_SPLIT=0 RXB_ROWLIST=0"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --quick --no-parity 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$cfg', round(d['ms_per_step'],3), 'cg',k['qeq_cg'],'spmv',k['spmv'],'bnd',k['spmv_boundary'])"
done
