# A/B of the shard exchange at 2+ GPUs: NVLink peer copies (default) against ncclSend / ncclRecv (NL_SHARD_PEER=0), both input modes
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
for envs in "NL_SHARD_PEER=1" "NL_SHARD_PEER=0"; do
  for mode in by-index slabbed; do
    echo "=== $envs $mode"
    env $envs timeout 300 $TR scripts/exp_shard_phases.py 10000000 $mode 5 2>&1 | grep "nl_shard_exchange\|shard phases\|^mode\|rror" | tail -4
  done
done
