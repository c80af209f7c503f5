run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 8 --steps 20 --warmup 5 --no-retrieval 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', d['ms_per_step'], d['config']['launch'], round(d['config']['eager_ms_per_step'],4), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; }
run TRICOLO_B200_SHARD_SYNC=barrier TRICOLO_B200_SHARDED_BWD=pc
run TRICOLO_B200_SHARD_SYNC=barrier TRICOLO_B200_SHARDED_BWD=sharedg
run TRICOLO_B200_SHARD_SYNC=flags TRICOLO_B200_PUSH=k1
run TRICOLO_B200_SHARD_SYNC=flags TRICOLO_B200_PUSH=fwd
