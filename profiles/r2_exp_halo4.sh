run() { # name, env...
  name=$1; shift
  env "$@" VB_TRACE=2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  echo "== $name: $(python -c "import json; d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])")"
  grep "vb mark\|vb halo" gpurun_out/exp_$name.err | tail -100 | grep "sweep of\|push of\|last mark\|keys of block 0" | sort | uniq -c | sort -k4,4 -k6,6n | head -30
}
run nooverlap VB_HALO_OVERLAP=0
run ce_1phase VB_HALO_PHASES=1
run push_nooverlap VB_HALO_OVERLAP=0 VB_HALO_CE=0
