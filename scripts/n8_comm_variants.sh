for c in nccl nccl-overlap p2p p2p-overlap; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e --comm $c > gpurun_out/n8_$c.json 2> gpurun_out/n8_$c.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/n8_$c.json") if l.startswith("{")][-1])
    print("$c", round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],4), "ms; sustained", round(d["sustained"]["value"]/1e6,2), round(d["sustained"]["ms_per_step"],4), d["allreduce_check"]["status"], d["allreduce_check"]["backend"])
except Exception as e:
    print("$c failed", e)
PY
done
