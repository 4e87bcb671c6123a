timeout 300 python -m pytest tests/test_sharded_gpu.py -x -q -k "four" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
for mode in by-index slabbed; do
  echo "=== peer $mode"
  timeout 300 $TR scripts/exp_shard_phases.py 10000000 $mode 5 2>&1 | grep "nl_shard_exchange\|shard phases\|^mode\|rror" | tail -3
done
echo "=== nccl by-index"
NL_SHARD_PEER=0 timeout 300 $TR scripts/exp_shard_phases.py 10000000 by-index 5 2>&1 | grep "nl_shard_exchange\|shard phases\|^mode\|rror" | tail -3
