# 1/2/4/8-GPU weak scaling on ONE 8-GPU box, back to back (gpurun --gpus 8)
cd $GRAFT_REPO_ROOT
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --also-modes "" > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline --also-modes "" > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  tail -c 300 gpurun_out/scale_$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --math bf16 --no-cpu-baseline --also-modes "" > gpurun_out/scale_8_bf16.json 2> gpurun_out/scale_8_bf16.err
python - <<'P'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/scale_{n}.json').read().strip().splitlines()[-1]); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
    except Exception as e: print(n, 'ERR', e)
P
