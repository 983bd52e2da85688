for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset AIDET_B200_LIB; else export AIDET_B200_LIB=$PWD/aidet_b200/libexp_riou_$v.so; fi
  AIDET_BENCH_NO_MCAST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --workload iou --no-cpu --no-e2e > gpurun_out/r2u_n2_$v.json 2> gpurun_out/r2u_n2_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2u_n2_$v.json')); m=d['multi_gpu']; print('$v', round(d['value'],1), {k:(round(v['value'],1), v.get('max_abs_err')) for k,v in m.items() if 'value' in v}, round(d['compute_only']['value'],1))"
done
